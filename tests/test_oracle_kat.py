"""Pins the CPU oracle against the closed-form known answers of SURVEY.md Appendix E and the structural
invariants of the reference architecture (the reference itself ships no golden vectors: "parity unpinned")."""
import json
import math
import os

import pytest
import torch

import oracle
from oracle.unet2d import get_timestep_embedding

GOLD = os.path.join(os.path.dirname(__file__), "golden")


def test_alphas_cumprod_kat():
    s = oracle.DDPMScheduler(num_train_timesteps=1000)
    ac = s.alphas_cumprod
    assert ac.dtype == torch.float32
    for i, want in ((0, 0.9998999834), (1, 0.9997800589), (499, 7.8587234020e-02), (999, 4.0358303522e-05)):
        assert ac[i].item() == pytest.approx(want, rel=2e-7)
    assert ac.double().sum().item() == pytest.approx(275.513195, abs=2e-6)
    s2 = oracle.DDPMScheduler(num_train_timesteps=2000)
    assert s2.alphas_cumprod[999].item() == pytest.approx(6.1605386436e-03, rel=2e-7)
    assert s2.alphas_cumprod[1999].item() == pytest.approx(1.6288478344e-09, rel=2e-6)


@pytest.mark.parametrize("t,want", [
    (999, (6.352818571e-03, 9.999797940e-01, 1.283513702e-04, 9.899486899e-01, 1.414212286e-01)),
    (500, (2.789205015e-01, 9.603142142e-01, 3.058036556e-03, 9.941043854e-01, 1.002560183e-01)),
    (1, (9.998900294e-01, 1.483041234e-02, 5.452302098e-01, 4.547152817e-01, 7.384767756e-03)),
    (0, (9.999499917e-01, 1.000082958e-02, 1.0, 0.0, None)),
])
def test_step_coefficients_kat(t, want):
    from polyp_image_generator_b200.scheduler import step_coefficients
    s = oracle.DDPMScheduler()
    c = step_coefficients(s.alphas_cumprod, t, t - 1)
    got = (c["sa"], c["sb"], c["c0"], c["ct"], c["sigma"])
    for g, w in zip(got, want):
        if w is None:
            assert g == 0.0
        else:
            assert g == pytest.approx(w, rel=3e-7, abs=1e-12)


def test_oracle_step_matches_closed_form():
    s = oracle.DDPMScheduler()
    s.set_timesteps(1000)
    torch.manual_seed(0)
    x, e, z = torch.randn(2, 3, 8, 8), torch.randn(2, 3, 8, 8), torch.randn(2, 3, 8, 8)
    out = s.step(e, torch.tensor(500), x, variance_noise=z)
    sa, sb, c0, ct, sg = 2.789205015e-01, 9.603142142e-01, 3.058036556e-03, 9.941043854e-01, 1.002560183e-01
    x0 = ((x - sb * e) / sa).clamp(-1, 1)
    want = c0 * x0 + ct * x + sg * z
    assert torch.allclose(out.prev_sample, want, rtol=1e-5, atol=1e-6)
    out0 = s.step(e, torch.tensor(0), x)
    assert torch.allclose(out0.prev_sample, out0.pred_original_sample)


def test_set_timesteps_and_errors():
    s = oracle.DDPMScheduler()
    s.set_timesteps(1000)
    assert s.timesteps[0].item() == 999 and s.timesteps[-1].item() == 0 and len(s.timesteps) == 1000
    s.set_timesteps(50)
    assert s.timesteps.tolist() == list(range(980, -1, -20))
    with pytest.raises(ValueError):
        s.set_timesteps(1001)


def test_timestep_embedding_kat():
    e = get_timestep_embedding(torch.tensor([500]), 128, True, 0)[0]
    for i, w in ((0, -0.88384926), (1, 0.84850687), (63, 0.99833357), (64, -0.46777180), (65, -0.52918434),
                 (127, 0.05770700)):
        assert e[i].item() == pytest.approx(w, abs=2e-6)
    assert e.double().sum().item() == pytest.approx(26.608001, abs=1e-4)
    e = get_timestep_embedding(torch.tensor([500]), 128, False, 1)[0]
    for i, w in ((0, -0.46777180), (1, -0.99968141), (63, 0.04997916), (64, -0.88384926), (127, 0.99875027)):
        assert e[i].item() == pytest.approx(w, abs=2e-6)


def test_cosine_schedule_kat():
    from polyp_image_generator_b200.training import get_cosine_schedule_with_warmup
    p = torch.nn.Parameter(torch.zeros(1))
    opt = torch.optim.SGD([p], lr=1.0)
    sch = get_cosine_schedule_with_warmup(opt, 500, 10000)
    f = sch.lr_lambdas[0]
    for step, want in ((0, 0.0), (250, 0.5), (500, 1.0), (5250, 0.5), (10000, 0.0)):
        assert f(step) == pytest.approx(want, abs=1e-9)


def test_unet_structure():
    m = oracle.UNet2DModel(**oracle.polyp_unet_config(64))
    assert sum(p.numel() for p in m.parameters()) == 113_673_219
    sd = m.state_dict()
    assert len(sd) == 450
    for k in ("conv_in.weight", "time_embedding.linear_1.weight", "down_blocks.4.attentions.1.to_out.0.bias",
              "down_blocks.2.resnets.0.conv_shortcut.weight", "mid_block.attentions.0.group_norm.weight",
              "up_blocks.1.attentions.2.to_q.weight", "up_blocks.4.upsamplers.0.conv.weight", "conv_norm_out.bias",
              "down_blocks.4.downsamplers.0.conv.weight", "conv_out.bias"):
        assert k in sd, k
    assert "down_blocks.5.downsamplers.0.conv.weight" not in sd
    shortcuts = [k for k in sd if k.endswith("conv_shortcut.weight")]
    assert len(shortcuts) == 20
    assert sd["up_blocks.0.resnets.0.conv1.weight"].shape == (512, 1024, 3, 3)
    assert sd["up_blocks.3.resnets.2.conv1.weight"].shape == (256, 384, 3, 3)
    c = oracle.UNet2DModel(**oracle.celebahq_unet_config(256))
    assert sum(p.numel() for p in c.parameters()) == 113_673_219
    assert c.mid_block.attentions[0].heads == 1 and c.mid_block.attentions[0].dim_head == 512


def test_lora_structure_and_key_grammar():
    m = oracle.UNet2DModel(**oracle.polyp_unet_config(64))
    cfg = oracle.LoraConfig(r=8, lora_alpha=8, target_modules=["to_q", "to_k", "to_v", "to_out.0"], lora_dropout=0.3,
                            init_lora_weights="gaussian")
    wrapped = oracle.add_adapter(m, cfg)
    assert len(wrapped) == 24
    assert sum(p.numel() for p in m.parameters() if p.requires_grad) == 196_608
    sd = oracle.lora_state_dict(m)
    assert len(sd) == 48
    assert "mid_block.attentions.0.to_q.lora_A.default.weight" in sd
    assert sd["mid_block.attentions.0.to_q.lora_A.default.weight"].shape == (8, 512)
    assert sd["mid_block.attentions.0.to_out.0.lora_B.default.weight"].shape == (512, 8)
    mods = oracle.recover_lora_modules(sd)
    assert len(mods) == 24 and "up_blocks.1.attentions.2.to_out.0" in mods
    assert "mid_block.attentions.0.to_q.base_layer.weight" in m.state_dict()


def test_golden_vectors_reproduce():
    """tests/golden/*.pt were written by tests/golden/make_golden.py from this oracle; re-derive and compare."""
    path = os.path.join(GOLD, "unet32_fwd_bwd.pt")
    if not os.path.exists(path):
        pytest.skip("golden fixture not generated")
    from golden.make_golden import small_case
    g = torch.load(path)
    case = small_case()
    assert torch.allclose(case["pred"], g["pred"], rtol=1e-4, atol=1e-5)
    assert case["loss"] == pytest.approx(g["loss"], rel=1e-5)
    for k, v in g["grad_norms"].items():
        assert case["grad_norms"][k] == pytest.approx(v, rel=1e-3, abs=1e-9)


def test_beta_schedule_tables_kat():
    """make_betas against an independent float64 derivation and the well-known Stable-Diffusion end points."""
    import numpy as np
    sl = oracle.make_betas("scaled_linear", 0.00085, 0.012, 1000)
    ref = np.linspace(0.00085 ** 0.5, 0.012 ** 0.5, 1000) ** 2
    assert np.allclose(sl.double().numpy(), ref, rtol=3e-6)
    ac = torch.cumprod(1 - sl, 0)
    assert ac[0].item() == pytest.approx(0.99915, abs=1e-6) and ac[999].item() == pytest.approx(0.0046601, rel=2e-3)
    cos = oracle.make_betas("squaredcos_cap_v2", 0, 0, 1000)
    bar = lambda t: math.cos((t + 0.008) / 1.008 * math.pi / 2) ** 2
    assert cos[0].item() == pytest.approx(1 - bar(1e-3) / bar(0.0), rel=1e-5)
    assert cos[999].item() == pytest.approx(0.999) and (cos[1:] >= cos[:-1] - 1e-7).all()
    with pytest.raises(NotImplementedError):
        oracle.make_betas("sigmoid", 1e-4, 0.02, 10)


def test_oracle_ddim_closed_form_properties():
    """eta = 0 DDIM with the TRUE noise walks x_t = sa x0 + sb eps back along the same (x0, eps) line; eta = 1 has the
    DDPM posterior standard deviation (step_coefficients KAT above: 1.002560183e-01 at t = 500 -> 499)."""
    s = oracle.DDIMScheduler(clip_sample=False)
    s.set_timesteps(1000)
    torch.manual_seed(0)
    x0, eps = torch.randn(2, 3, 8, 8).clamp(-1, 1), torch.randn(2, 3, 8, 8)
    ac = s.alphas_cumprod
    xt = ac[500] ** 0.5 * x0 + (1 - ac[500]) ** 0.5 * eps
    out = s.step(eps, torch.tensor(500), xt, eta=0.0)
    assert torch.allclose(out.pred_original_sample, x0, atol=2e-5)
    assert torch.allclose(out.prev_sample, ac[499] ** 0.5 * x0 + (1 - ac[499]) ** 0.5 * eps, atol=2e-5)
    z = torch.randn_like(x0)
    d = s.step(eps, torch.tensor(500), xt, eta=1.0, variance_noise=z).prev_sample - \
        s.step(eps, torch.tensor(500), xt, eta=1.0, variance_noise=torch.zeros_like(z)).prev_sample
    assert torch.allclose(d, 1.002560183e-01 * z, rtol=1e-5, atol=1e-7)
    s.set_timesteps(50)                      # strided: t = 980 -> 960, last step t = 0 -> final_alpha_cumprod = 1
    assert s.timesteps[0].item() == 980 and s.timesteps[-1].item() == 0
    x1 = ac[0] ** 0.5 * x0 + (1 - ac[0]) ** 0.5 * eps
    assert torch.allclose(s.step(eps, torch.tensor(0), x1).prev_sample, x0, atol=2e-5)


def test_ddim_golden_trajectory():
    """tests/golden/ddim.json (make_golden_session3.py): the oracle reproduces its committed 25-step trajectory."""
    gold = json.load(open(os.path.join(GOLD, "ddim.json")))
    s = oracle.DDIMScheduler()
    s.set_timesteps(25)
    assert s.timesteps.tolist() == gold["timesteps"]
    g = torch.Generator().manual_seed(11)
    x = torch.randn(1, 3, 8, 8, generator=g)
    for t, want in zip(s.timesteps, gold["trajectory_sums"]):
        eps = torch.cos(x * 2.0 - float(t) * 0.02)
        x = s.step(eps, t, x, eta=gold["eta"], use_clipped_model_output=True, generator=g).prev_sample
        assert x.double().sum().item() == pytest.approx(want, rel=1e-5, abs=1e-5)
    assert torch.allclose(x.flatten(), torch.tensor(gold["final"]), rtol=1e-5, atol=1e-6)


def test_oracle_unipc_closed_form_and_golden():
    """UniPC (Zhao et al. 2023) integrates the data-prediction ODE exactly when the x0 prediction is constant: fed the
    TRUE noise of a fixed (x0, e) line, every iterate is alpha_t x0 + sigma_t e and the final sample (sigma = 0) is x0 --
    for both solver orders, any step count; plus the committed 12-step trajectory (make_golden_unipc.py)."""
    torch.manual_seed(0)
    x0, e = torch.randn(2, 3, 8, 8).double(), torch.randn(2, 3, 8, 8).double()
    for order in (1, 2):
        for n in (25, 7, 2):
            s = oracle.UniPCMultistepScheduler(solver_order=order)
            s.set_timesteps(n)
            a, sg = s._sigma_to_alpha_sigma_t(s.sigmas[0].double())
            x = (a * x0 + sg * e).float()
            for i, t in enumerate(s.timesteps):
                a, sg = s._sigma_to_alpha_sigma_t(s.sigmas[i].double())
                eps = ((x.double() - a * x0) / sg).float()         # the noise that makes the x0 prediction exact
                x = s.step(eps, t, x).prev_sample
                a, sg = s._sigma_to_alpha_sigma_t(s.sigmas[i + 1].double())
                assert torch.allclose(x.double(), a * x0 + sg * e, atol=5e-4), (order, n, i)
            assert torch.allclose(x.double(), x0, atol=5e-4)
    # timestep / sigma grid of the default configuration: linspace over [0, T-1], rounded, descending, last dropped
    s = oracle.UniPCMultistepScheduler()
    s.set_timesteps(25)
    assert s.timesteps.tolist()[:3] == [999, 959, 919] and s.timesteps[-1].item() == 40 and s.sigmas[-1].item() == 0.0
    ac = s.alphas_cumprod
    assert s.sigmas[0].item() == pytest.approx(float(((1 - ac[999]) / ac[999]) ** 0.5), rel=1e-6)
    gold = json.load(open(os.path.join(GOLD, "unipc.json")))
    s = oracle.UniPCMultistepScheduler(**gold["kwargs"])
    s.set_timesteps(gold["steps"])
    assert s.timesteps.tolist() == gold["timesteps"]
    assert torch.allclose(s.sigmas, torch.tensor(gold["sigmas"]), rtol=1e-6)
    x = torch.randn(1, 3, 8, 8, generator=torch.Generator().manual_seed(13))
    for t, want in zip(s.timesteps, gold["trajectory_sums"]):
        x = s.step(torch.cos(x * 2.0 - float(t) * 0.02), t, x).prev_sample
        assert x.double().sum().item() == pytest.approx(want, rel=1e-5, abs=1e-5)
    assert torch.allclose(x.flatten(), torch.tensor(gold["final"]), rtol=1e-5, atol=1e-6)


def test_resize_golden_from_pillow():
    """tests/golden/resize.json was written by Pillow / torchvision themselves; the oracle restatement must match it."""
    import numpy as np
    from oracle import preprocess as opre
    sys_path_hack = os.path.join(GOLD)
    gold = json.load(open(os.path.join(GOLD, "resize.json")))
    import importlib.util
    spec = importlib.util.spec_from_file_location("mk3", os.path.join(sys_path_hack, "make_golden_session3.py"))
    mk3 = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mk3)
    img = mk3.pattern(gold["h"], gold["w"])
    assert np.array_equal(opre.resize_bilinear_u8(img, gold["size"], gold["size"]), np.array(gold["resized"], dtype=np.uint8))
    t = opre.transform(img, gold["size"], True)
    assert t.double().sum().item() == pytest.approx(gold["transform_flipped_sum"], abs=1e-9)
    assert torch.equal(t[0, 0], torch.tensor(gold["transform_flipped_first_row"]))


def test_oracle_attention_block_equals_torch_multihead_attention():
    """An INDEPENDENT implementation that is present in this image: torch.nn.MultiheadAttention splits the embedding
    into heads the way diffusers' Attention does (head h = channels [h*d, (h+1)*d)), scales by d^-1/2 and applies the
    output projection -- so the oracle's attention block (GroupNorm -> q/k/v -> SDPA -> to_out.0 -> + residual) must
    equal x + MHA(GroupNorm(x)) with the same weights.  Pins the head split / scale / token order of oracle/unet2d.py."""
    from oracle.unet2d import Attention
    torch.manual_seed(0)
    for C, d, hw in ((64, 8, (4, 4)), (128, 8, (7, 7)), (64, 64, (8, 8))):
        heads = C // d
        att = Attention(C, heads, d, norm_num_groups=32).double()
        mha = torch.nn.MultiheadAttention(C, heads, bias=True, batch_first=True).double()
        with torch.no_grad():
            mha.in_proj_weight.copy_(torch.cat([att.to_q.weight, att.to_k.weight, att.to_v.weight], 0))
            mha.in_proj_bias.copy_(torch.cat([att.to_q.bias, att.to_k.bias, att.to_v.bias], 0))
            mha.out_proj.weight.copy_(att.to_out[0].weight)
            mha.out_proj.bias.copy_(att.to_out[0].bias)
        x = torch.randn(2, C, *hw, dtype=torch.float64)
        tok = att.group_norm(x).flatten(2).transpose(1, 2)                  # [B, T, C], token = h * W + w
        want = x + mha(tok, tok, tok, need_weights=False)[0].transpose(1, 2).reshape(x.shape)
        got = att(x)
        assert torch.allclose(got, want, rtol=1e-10, atol=1e-10), (C, d)
