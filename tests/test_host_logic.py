"""Host-side logic of the product on CPU: C-ABI surface, scheduler tables, tap tables, arena layout, and the UNet
forward/backward *programs* verified against the oracle through the CPU emulation of the op contracts
(tests/emu_ops.py).  No CUDA compute happens here."""
import ctypes
import os
import re

import pytest
import torch

import oracle

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_capi_exports_every_declared_symbol(built_lib):
    hdr = open(os.path.join(ROOT, "include", "ddpm_b200.h")).read()
    declared = set(re.findall(r"\b(ddpm_[a-z0-9_]+)\s*\(", hdr))
    declared -= {"ddpm_conv_args", "ddpm_wgrad_args"}
    assert len(declared) >= 25
    lib = ctypes.CDLL(str(built_lib))
    for name in sorted(declared):
        assert hasattr(lib, name), f"{name} declared in include/ddpm_b200.h but not exported"
    from polyp_image_generator_b200 import _capi
    assert set(_capi.SIGNATURES) | {"ddpm_last_error"} == declared
    assert _capi.load().ddpm_abi_version() == _capi.ABI_VERSION == 3


def test_capi_argument_validation_without_gpu(built_lib):
    """Entry points reject bad arguments before touching the device (error string, negative rc)."""
    from polyp_image_generator_b200 import _capi
    lib = _capi.load()
    assert lib.ddpm_add_noise(None, None, None, None, None, None, 1, 1, 1000, None) < 0
    assert "null pointer" in _capi.last_error()
    a = _capi.ConvArgs()
    assert lib.ddpm_conv_gemm(ctypes.byref(a), None) < 0
    assert lib.ddpm_attn_fwd(1, 8, 1, 8, 1, 1, 4, 1, 24, 1.0, None) < 0
    assert "head_dim=24" in _capi.last_error()


def test_product_refuses_cpu_tensors(built_lib):
    from polyp_image_generator_b200 import DDPMScheduler, UNet2DModel, ops as ops_mod
    assert ops_mod._backend is None or ops_mod._backend.name == "cuda"
    s = DDPMScheduler()
    x = torch.zeros(1, 3, 4, 4)
    with pytest.raises(RuntimeError, match="CUDA"):
        s.add_noise(x, x, torch.tensor([1]))
    cfg = oracle.polyp_unet_config(32)
    cfg["block_out_channels"] = (64, 64, 64, 64, 64, 64)
    m = UNet2DModel(**cfg)
    with pytest.raises(RuntimeError, match="CUDA"):
        m(torch.zeros(1, 3, 32, 32), 1)


def test_scheduler_tables_and_timesteps_match_oracle():
    from polyp_image_generator_b200 import DDPMScheduler
    for T in (1000, 2000):
        a, b = DDPMScheduler(num_train_timesteps=T), oracle.DDPMScheduler(num_train_timesteps=T)
        assert torch.equal(a.alphas_cumprod, b.alphas_cumprod) and torch.equal(a.betas, b.betas)
        assert torch.equal(a.timesteps, b.timesteps)
        assert a.config.num_train_timesteps == T and a.init_noise_sigma == 1.0
    a, b = DDPMScheduler(), oracle.DDPMScheduler()
    for n in (1000, 250, 50, 7):
        a.set_timesteps(n)
        b.set_timesteps(n)
        assert torch.equal(a.timesteps, b.timesteps)
        for t in a.timesteps.tolist()[:5] + a.timesteps.tolist()[-3:]:
            assert a.previous_timestep(t) == int(b.previous_timestep(torch.tensor(t)))
    with pytest.raises(ValueError, match="cannot be larger"):
        a.set_timesteps(1001)
    with pytest.raises(NotImplementedError):
        DDPMScheduler(beta_schedule="sigmoid")


def test_scheduler_step_and_add_noise_through_emulation(emu_backend):
    from polyp_image_generator_b200 import DDPMScheduler
    a, b = DDPMScheduler(), oracle.DDPMScheduler()
    a.set_timesteps(1000)
    b.set_timesteps(1000)
    torch.manual_seed(0)
    x, e, z = torch.randn(2, 3, 8, 8), torch.randn(2, 3, 8, 8), torch.randn(2, 3, 8, 8)
    t = torch.tensor([3, 998])
    assert torch.allclose(a.add_noise(x, e, t), b.add_noise(x, e, t), rtol=1e-6, atol=1e-7)
    for tt in (999, 500, 1, 0):
        pa = a.step(e, tt, x, variance_noise=z).prev_sample
        pb = b.step(e, torch.tensor(tt), x, variance_noise=z).prev_sample
        assert torch.allclose(pa, pb, rtol=1e-6, atol=1e-6)
    ga, gb = torch.Generator().manual_seed(5), torch.Generator().manual_seed(5)
    pa = a.step(e, 10, x, generator=ga).prev_sample
    pb = b.step(e, torch.tensor(10), x, generator=gb).prev_sample
    assert torch.allclose(pa, pb, rtol=1e-6, atol=1e-6)   # consumes the CPU generator like randn_tensor


def test_tap_tables():
    from polyp_image_generator_b200.ops import taps_3x3, taps_s2d
    t = taps_3x3(128)
    assert t[0] == (0, -1, -1, 0) and t[4] == (0, 0, 0, 4 * 128) and t[8] == (0, 1, 1, 8 * 128)
    n = 5
    for pad in (0, 1):
        for (dn, dh, dw, wk), (r, s) in zip(taps_s2d(64, n, pad), [(r, s) for r in range(3) for s in range(3)]):
            phase = dn // n
            ph, pw = phase // 2, phase % 2
            assert 2 * dh + ph == r - pad and 2 * dw + pw == s - pad and wk == (r * 3 + s) * 64


def _small_cfg(S=32, **kw):
    cfg = oracle.polyp_unet_config(S)
    cfg["block_out_channels"] = (64, 64, 128, 128, 128, 128)
    cfg.update(kw)
    return cfg


def test_unet_state_dict_and_arena(emu_backend):
    from polyp_image_generator_b200 import UNet2DModel
    m = UNet2DModel(**oracle.polyp_unet_config(64))
    om = oracle.UNet2DModel(**oracle.polyp_unet_config(64))
    sd, osd = m.state_dict(), om.state_dict()
    assert list(sd.keys()) == list(osd.keys())
    assert all(sd[k].shape == osd[k].shape for k in sd)
    assert sum(p.numel() for p in m.parameters()) == 113_673_219
    m.load_state_dict(osd)
    m._ensure_arena()
    P = m._plan
    assert P.temb_total == 9984
    base = m._arena.data_ptr()
    for p, off, phys in P.layout:
        assert p.data_ptr() == base + 4 * off
    # parameters are views into the arena with unchanged logical values; conv weights are channels_last
    assert torch.equal(m.conv_in.weight.detach(), osd["conv_in.weight"])
    w = m.down_blocks[2].resnets[0].conv1.weight
    assert w.shape == (256, 128, 3, 3) and w.stride() == (9 * 128, 1, 3 * 128, 128)
    assert torch.equal(w.detach(), osd["down_blocks.2.resnets.0.conv1.weight"])
    # q, k, v weights (and biases) are adjacent so the fused qkv GEMM / wgrad see one matrix
    a = m.mid_block.attentions[0]
    assert a.to_k.weight.data_ptr() == a.to_q.weight.data_ptr() + 4 * 512 * 512
    assert a.to_v.bias.data_ptr() == a.to_q.bias.data_ptr() + 4 * 1024
    # .to()/state_dict round trip keeps values and re-flattens lazily
    m2 = UNet2DModel(**oracle.polyp_unet_config(64))
    m2.load_state_dict(m.state_dict())
    assert torch.equal(m2.conv_out.weight, m.conv_out.weight)


@pytest.mark.parametrize("variant", ["polyp", "celebahq", "celebahq_1head", "polyp_96"])
def test_unet_forward_backward_programs_match_oracle(emu_backend, variant):
    from polyp_image_generator_b200 import UNet2DModel
    if variant == "polyp":
        cfg, S = _small_cfg(32), 32
    elif variant == "polyp_96":      # like the reference's 224 = 7 * 32: odd maps (3x3) at the bottom, 36 / 9 tokens
        cfg, S = _small_cfg(96), 96
    else:
        cfg = oracle.celebahq_unet_config(64)
        cfg["block_out_channels"] = (64, 64, 128, 128, 128, 128)
        if variant == "celebahq":
            cfg["attention_head_dim"] = 16
        S = 64
    torch.manual_seed(0)
    om = oracle.UNet2DModel(**cfg)
    m = UNet2DModel(**cfg)
    m.load_state_dict(om.state_dict())
    x, t = torch.randn(2, 3, S, S), torch.tensor([10, 700])
    y, yo = m(x, t).sample, om(x, t).sample
    assert ((y - yo).norm() / yo.norm()).item() < 1e-5
    tgt = torch.randn_like(y)
    torch.nn.functional.mse_loss(y, tgt).backward()
    torch.nn.functional.mse_loss(yo, tgt).backward()
    og = dict(om.named_parameters())
    tot = sum(p.grad.norm() ** 2 for p in om.parameters()) ** 0.5
    for n, p in m.named_parameters():
        g = og[n].grad
        assert p.grad is not None, n
        assert ((p.grad - g).norm() / (g.norm() + 1e-6 * tot)).item() < 1e-3, n
    # python-int and 0-dim timesteps, return_dict=False, no_grad path
    with torch.no_grad():
        y1 = m(x, 37, return_dict=False)[0]
        y2 = m(x, torch.tensor(37)).sample
        yo1 = om(x, 37).sample
    assert torch.equal(y1, y2) and ((y1 - yo1).norm() / yo1.norm()).item() < 1e-5


def test_unet_errors_mirror_reference_behaviour():
    from polyp_image_generator_b200 import UNet2DModel
    with pytest.raises(ValueError, match="same number"):
        UNet2DModel(down_block_types=("DownBlock2D",), up_block_types=("UpBlock2D", "UpBlock2D"),
                    block_out_channels=(64,))
    with pytest.raises(NotImplementedError):
        UNet2DModel(**_small_cfg(), time_embedding_type="fourier")
    m = UNet2DModel(**_small_cfg())
    with pytest.raises(TypeError):
        m(torch.zeros(1, 3, 32, 32), 1, encoder_hidden_states=None)   # train_from_scratch.py:98 dead branch


def test_train_step_recipe_matches_oracle_loop(emu_backend):
    """One full iteration of train_from_scratch.py:83-116 (add_noise, UNet, MSE, clip, AdamW) vs the oracle."""
    from polyp_image_generator_b200 import DDPMScheduler, UNet2DModel
    from polyp_image_generator_b200.training import train_step
    cfg = _small_cfg(32)
    torch.manual_seed(0)
    om = oracle.UNet2DModel(**cfg)
    m = UNet2DModel(**cfg)
    m.load_state_dict(om.state_dict())
    opt, oopt = torch.optim.AdamW(m.parameters(), lr=1e-4), torch.optim.AdamW(om.parameters(), lr=1e-4)
    x0, noise, t = torch.randn(2, 3, 32, 32).clamp(-1, 1), torch.randn(2, 3, 32, 32), torch.tensor([5, 900])
    loss = train_step(m, DDPMScheduler(), opt, x0, noise, t)
    osch = oracle.DDPMScheduler()
    pred = om(osch.add_noise(x0, noise, t), t, return_dict=False)[0]
    oloss = torch.nn.functional.mse_loss(pred, noise)
    oloss.backward()
    torch.nn.utils.clip_grad_norm_(om.parameters(), 1.0)
    oopt.step()
    assert loss.item() == pytest.approx(oloss.item(), rel=1e-5)
    osd = om.state_dict()
    # Adam's first update is lr*sign(g) wherever |g| >> eps, so elements whose gradient is numerically zero may
    # land on either side: every weight moves by at most lr, and all but a sliver of them identically.
    diffs = torch.cat([(v - osd[k]).abs().reshape(-1) for k, v in m.state_dict().items()])
    assert diffs.max().item() <= 2.0 * 1e-4 * 1.01 + 1e-7
    assert (diffs > 2e-6).float().mean().item() < 2e-3
    assert all(p.grad is None for p in m.parameters())   # zero_grad(set_to_none)


def test_sampling_shards_cover_single_gpu_image_set():
    from polyp_image_generator_b200.ddp import shard_sampling_batches
    for total, bs, world in ((256, 32, 8), (1024, 20, 8), (100, 20, 3), (7, 20, 4)):
        seen = []
        for r in range(world):
            seen += shard_sampling_batches(total, bs, r, world)
        seen.sort()
        assert [b for b, _, _ in seen] == list(range(len(seen)))
        assert sum(c for _, _, c in seen) == total
        assert all(s == b * bs for b, s, _ in seen)


def test_pipeline_matches_oracle_pipeline(emu_backend):
    """DDPMPipeline RNG contract: x_T and every step's z come from the CPU generator in diffusers' order."""
    from polyp_image_generator_b200 import DDPMPipeline, DDPMScheduler, UNet2DModel
    cfg = _small_cfg(32)
    cfg["block_out_channels"] = (64, 64, 64, 64, 64, 64)
    torch.manual_seed(0)
    om = oracle.UNet2DModel(**cfg)
    m = UNet2DModel(**cfg)
    m.load_state_dict(om.state_dict())
    pa = DDPMPipeline(unet=m, scheduler=DDPMScheduler())
    pb = oracle.DDPMPipeline(unet=om, scheduler=oracle.DDPMScheduler())
    ia = pa(batch_size=2, generator=torch.Generator("cpu").manual_seed(3), num_inference_steps=4, output_type="np").images
    ib = pb(batch_size=2, generator=torch.Generator("cpu").manual_seed(3), num_inference_steps=4, output_type="np").images
    assert ia.shape == (2, 32, 32, 3)
    assert abs(ia - ib).max() <= 1.0 / 255 + 1e-6
    pil = pa(batch_size=1, generator=torch.Generator("cpu").manual_seed(3), num_inference_steps=2).images
    assert len(pil) == 1 and pil[0].size == (32, 32)
    with pytest.raises(ValueError):
        pa(batch_size=1, num_inference_steps=1001)


def test_fused_adamw_matches_clip_plus_torch_adamw(emu_backend):
    """optim.FusedAdamW (flat-arena clip + AdamW, train_from_scratch.py:106-108,273) vs clip_grad_norm_ + torch AdamW."""
    from polyp_image_generator_b200 import FusedAdamW, UNet2DModel
    cfg = _small_cfg(32)
    torch.manual_seed(0)
    a, b = UNet2DModel(**cfg), UNet2DModel(**cfg)
    b.load_state_dict(a.state_dict())
    oa = FusedAdamW(a.parameters(), lr=3e-4, weight_decay=0.01, max_grad_norm=1.0)
    ob = torch.optim.AdamW(b.parameters(), lr=3e-4, weight_decay=0.01)
    x, t, tgt = torch.randn(2, 3, 32, 32), torch.tensor([3, 600]), torch.randn(2, 3, 32, 32)
    for it in range(3):
        for m in (a, b):
            (torch.nn.functional.mse_loss(m(x, t).sample, tgt) * (40.0 if it == 0 else 1.0)).backward()
        oa.step()
        torch.nn.utils.clip_grad_norm_(b.parameters(), 1.0)
        ob.step()
        oa.zero_grad()
        ob.zero_grad()
        sa, sb = a.state_dict(), b.state_dict()
        for k in sa:
            assert torch.allclose(sa[k], sb[k], rtol=1e-5, atol=1e-6), (it, k)   # lr = 3e-4: <= 0.4 % of one update
    assert float(oa._scal[0]) == 3.0
    a.eval()
    with torch.no_grad():
        a(x, 5)
    key = a._wcache_key
    assert key is not None
    (torch.nn.functional.mse_loss(a.train()(x, t).sample, tgt)).backward()
    a.eval()
    with torch.no_grad():
        a(x, 5)
    oa.step()                                   # writes the arena behind the version counters ...
    assert a._wcache_key is None                # ... and tells the model to refresh its bf16 operands
    # foreign / frozen parameters are refused instead of silently skipped
    a.conv_in.weight.requires_grad_(False)
    with pytest.raises(NotImplementedError):
        FusedAdamW([p for p in a.parameters() if p.requires_grad], lr=1e-4).step()


def test_sampling_bookkeeping_file_names_seeds_and_shards(emu_backend, tmp_path):
    """train_from_scratch.py:39-66 / train_with_lora_per_class.py:59-88,262-290: ragged last batch, seed + batch_id,
    1-based <n>.png names, rank shards reproduce the single-process image set, top-up counts existing files."""
    from types import SimpleNamespace
    import numpy as np
    from PIL import Image
    from polyp_image_generator_b200 import DDPMPipeline, DDPMScheduler, UNet2DModel
    from polyp_image_generator_b200.sampling import count_samples, evaluate, top_up
    cfg = _small_cfg(32)
    cfg["block_out_channels"] = (64, 64, 64, 64, 64, 64)
    torch.manual_seed(0)
    pipe = DDPMPipeline(unet=UNet2DModel(**cfg), scheduler=DDPMScheduler())
    single = SimpleNamespace(output_dir=str(tmp_path / "one"), eval_batch_size=3, seed=11)
    paths = evaluate(single, 0, pipe, "AD", 7, num_inference_steps=2, verbose=False)
    assert [os.path.basename(p) for p in paths] == [f"{i}.png" for i in range(1, 8)]
    # batch 2 (the ragged one: image 7) was drawn from seed 11 + 2
    want = pipe(batch_size=1, generator=torch.Generator("cpu").manual_seed(13), num_inference_steps=2,
                output_type="uint8").images[0].numpy()
    assert np.array_equal(np.asarray(Image.open(paths[6])), want)
    # two ranks, no communication: same files, same pixels
    shard = SimpleNamespace(output_dir=str(tmp_path / "two"), eval_batch_size=3, seed=11)
    p0 = evaluate(shard, 0, pipe, "AD", 7, rank=0, world=2, num_inference_steps=2, verbose=False)
    p1 = evaluate(shard, 0, pipe, "AD", 7, rank=1, world=2, num_inference_steps=2, verbose=False)
    assert sorted(os.path.basename(p) for p in p0 + p1) == sorted(os.path.basename(p) for p in paths)
    assert [os.path.basename(p) for p in p1] == ["4.png", "5.png", "6.png"]
    for p in paths:
        q = os.path.join(shard.output_dir, "samples", "AD", os.path.basename(p))
        assert np.array_equal(np.asarray(Image.open(p)), np.asarray(Image.open(q)))
    # top-up: 7 files present, 9 wanted -> the reference regenerates 1.png, 2.png; continue_numbering appends 8, 9
    d = os.path.join(single.output_dir, "samples", "AD")
    assert count_samples(d) == 7 and top_up(single, pipe, "AD", 7, num_inference_steps=2, verbose=False) == []
    again = top_up(single, pipe, "AD", 9, num_inference_steps=2, verbose=False)
    assert [os.path.basename(p) for p in again] == ["1.png", "2.png"] and count_samples(d) == 7
    more = top_up(single, pipe, "AD", 9, continue_numbering=True, num_inference_steps=2, verbose=False)
    assert [os.path.basename(p) for p in more] == ["8.png", "9.png"] and count_samples(d) == 9
    fresh = top_up(single, pipe, "HP", 2, num_inference_steps=2, verbose=False)
    assert [os.path.basename(p) for p in fresh] == ["1.png", "2.png"]


def test_beta_schedules_and_ddim_scheduler_match_oracle(emu_backend):
    """scaled_linear / cosine tables and the DDIM strided sampler, against the oracle restatement."""
    from polyp_image_generator_b200 import DDIMScheduler, DDPMScheduler
    for sched, (b0, b1) in (("linear", (1e-4, 0.02)), ("scaled_linear", (0.00085, 0.012)), ("squaredcos_cap_v2", (1e-4, 0.02))):
        a = DDPMScheduler(beta_schedule=sched, beta_start=b0, beta_end=b1)
        b = oracle.DDPMScheduler(beta_schedule=sched, beta_start=b0, beta_end=b1)
        assert torch.equal(a.betas, b.betas) and torch.equal(a.alphas_cumprod, b.alphas_cumprod)
    # SD-style forward noising (the noise_scheduler of train_with_lora_all_classes.py:314 uses this table)
    a, b = DDPMScheduler(beta_schedule="scaled_linear", beta_start=0.00085, beta_end=0.012), \
        oracle.DDPMScheduler(beta_schedule="scaled_linear", beta_start=0.00085, beta_end=0.012)
    torch.manual_seed(1)
    x, e, z = torch.randn(2, 3, 8, 8), torch.randn(2, 3, 8, 8), torch.randn(2, 3, 8, 8)
    t = torch.tensor([0, 999])
    assert torch.allclose(a.add_noise(x, e, t), b.add_noise(x, e, t), rtol=1e-6, atol=1e-7)
    a, b = DDIMScheduler(), oracle.DDIMScheduler()
    with pytest.raises(ValueError, match="set_timesteps"):
        a.step(e, 10, x)
    for n in (50, 7, 1000):
        a.set_timesteps(n)
        b.set_timesteps(n)
        assert torch.equal(a.timesteps, b.timesteps)
        for tt in (a.timesteps.tolist()[0], a.timesteps.tolist()[len(a.timesteps) // 2], a.timesteps.tolist()[-1]):
            for eta, clipped in ((0.0, False), (0.0, True), (0.5, False), (1.0, True)):
                ra = a.step(e, tt, x, eta=eta, use_clipped_model_output=clipped, variance_noise=z if eta > 0 else None)
                rb = b.step(e, torch.tensor(tt), x, eta=eta, use_clipped_model_output=clipped,
                            variance_noise=z if eta > 0 else None)
                assert torch.allclose(ra.prev_sample, rb.prev_sample, rtol=1e-6, atol=1e-6), (n, tt, eta, clipped)
                assert torch.allclose(ra.pred_original_sample, rb.pred_original_sample, rtol=1e-6, atol=1e-6)
    ga, gb = torch.Generator().manual_seed(9), torch.Generator().manual_seed(9)
    assert torch.allclose(a.step(e, 500, x, eta=1.0, generator=ga).prev_sample,
                          b.step(e, torch.tensor(500), x, eta=1.0, generator=gb).prev_sample, rtol=1e-6, atol=1e-6)
    with pytest.raises(ValueError, match="Cannot pass both"):
        a.step(e, 500, x, eta=1.0, generator=ga, variance_noise=z)
    with pytest.raises(ValueError, match="cannot be larger"):
        a.set_timesteps(1001)


def test_ddim_pipeline_matches_oracle_loop_and_round_trips(emu_backend, tmp_path):
    from polyp_image_generator_b200 import DDIMPipeline, DDIMScheduler, UNet2DModel
    cfg = _small_cfg(32)
    cfg["block_out_channels"] = (64, 64, 64, 64, 64, 64)
    torch.manual_seed(0)
    om = oracle.UNet2DModel(**cfg)
    m = UNet2DModel(**cfg)
    m.load_state_dict(om.state_dict())
    pipe = DDIMPipeline(unet=m, scheduler=DDIMScheduler())
    got = pipe(batch_size=2, generator=torch.Generator("cpu").manual_seed(4), num_inference_steps=5, eta=0.3,
               output_type="np").images
    # the oracle loop (diffusers DDIMPipeline.__call__): x_T from the generator, then step(..., eta, generator)
    osch = oracle.DDIMScheduler()
    g = torch.Generator("cpu").manual_seed(4)
    img = oracle.randn_tensor((2, 3, 32, 32), generator=g)
    osch.set_timesteps(5)
    with torch.no_grad():
        for t in osch.timesteps:
            img = osch.step(om(img, t).sample, t, img, eta=0.3, use_clipped_model_output=False, generator=g).prev_sample
    want = (img / 2 + 0.5).clamp(0, 1).permute(0, 2, 3, 1).numpy()
    assert abs(got - want).max() <= 1.0 / 255 + 1e-6
    pipe.save_pretrained(str(tmp_path / "p"))
    back = DDIMPipeline.from_pretrained(str(tmp_path / "p"))
    assert type(back.scheduler).__name__ == "DDIMScheduler" and back.scheduler.config.set_alpha_to_one is True


def test_unipc_scheduler_and_pipeline_match_oracle(emu_backend, tmp_path):
    """UniPCMultistepScheduler (train_with_lora_all_classes.py:314) -- the tabulated-coefficient
    host logic against the oracle restatement, bit for bit on the emulated kernels; add_noise; the sampling loop."""
    from polyp_image_generator_b200 import UNet2DModel, UniPCMultistepScheduler, UniPCPipeline
    sd = dict(beta_schedule="scaled_linear", beta_start=0.00085, beta_end=0.012, steps_offset=1)
    for kw in (dict(), sd, dict(solver_order=1), dict(disable_corrector=[0, 3])):
        for n in (25, 4, 2, 1):
            a, b = UniPCMultistepScheduler(**kw), oracle.UniPCMultistepScheduler(**kw)
            a.set_timesteps(n)
            b.set_timesteps(n)
            assert torch.equal(a.timesteps, b.timesteps) and torch.equal(a.sigmas, b.sigmas)
            g = torch.Generator().manual_seed(n)
            xa = xb = torch.randn(2, 3, 8, 8, generator=g)
            for t in a.timesteps:
                eps = torch.randn(2, 3, 8, 8, generator=g)
                xa, xb = a.step(eps, t, xa).prev_sample, b.step(eps, t, xb).prev_sample
                assert torch.equal(xa, xb), (kw, n, int(t))
            assert a.step_index == b.step_index == n
    with pytest.raises(ValueError, match="set_timesteps"):
        UniPCMultistepScheduler().step(xa, 10, xa)
    for bad in (dict(solver_type="bh1"), dict(prediction_type="v_prediction"), dict(solver_order=3),
                dict(use_karras_sigmas=True)):
        with pytest.raises(NotImplementedError):
            UniPCMultistepScheduler(**bad)
    # forward noising on the UniPC scheduler (train_with_lora_all_classes.py:137): alpha_t x0 + sigma_t noise
    a = UniPCMultistepScheduler(**sd)
    x, e, t = torch.randn(3, 4, 8, 8), torch.randn(3, 4, 8, 8), torch.tensor([0, 500, 999])
    ac = a.alphas_cumprod[t].view(-1, 1, 1, 1)
    assert torch.allclose(a.add_noise(x, e, t), ac ** 0.5 * x + (1 - ac) ** 0.5 * e, rtol=1e-5, atol=1e-6)
    # sampling loop + on-disk round trip
    cfg = _small_cfg(32)
    cfg["block_out_channels"] = (64, 64, 64, 64, 64, 64)
    torch.manual_seed(0)
    om = oracle.UNet2DModel(**cfg)
    m = UNet2DModel(**cfg)
    m.load_state_dict(om.state_dict())
    pipe = UniPCPipeline(unet=m, scheduler=UniPCMultistepScheduler())
    got = pipe(batch_size=2, generator=torch.Generator("cpu").manual_seed(4), num_inference_steps=5,
               output_type="np").images
    osch = oracle.UniPCMultistepScheduler()
    img = oracle.randn_tensor((2, 3, 32, 32), generator=torch.Generator("cpu").manual_seed(4))
    osch.set_timesteps(5)
    with torch.no_grad():
        for t in osch.timesteps:
            img = osch.step(om(img, t).sample, t, img).prev_sample
    want = (img / 2 + 0.5).clamp(0, 1).permute(0, 2, 3, 1).numpy()
    assert abs(got - want).max() <= 1.0 / 255 + 1e-6
    pipe.save_pretrained(str(tmp_path / "p"))
    back = UniPCPipeline.from_pretrained(str(tmp_path / "p"))
    assert type(back.scheduler).__name__ == "UniPCMultistepScheduler" and back.scheduler.config.solver_order == 2


def test_bench_has_no_rank_conditional_collectives():
    """Every rank must walk bench.py's GPU arm through the same sequence of collectives: a training step (DDP
    all-reduce), a barrier or a dist.* call under `if rank == 0` deadlocks the multi-GPU run (it did, once)."""
    import ast
    tree = ast.parse(open(os.path.join(ROOT, "bench.py")).read())
    collective = {"step", "eager_step", "e2e_iter", "run_sampling", "run_lora_finetune", "barrier", "gstep", "net"}

    def mentions_rank(node):
        return any(isinstance(n, ast.Name) and n.id == "rank" for n in ast.walk(node))

    def calls_in(nodes):
        out = []
        for st in nodes:
            for n in ast.walk(st):
                if isinstance(n, ast.Call):
                    f = n.func
                    if isinstance(f, ast.Name) and f.id in collective:
                        out.append(f.id)
                    if isinstance(f, ast.Attribute) and isinstance(f.value, ast.Name) and f.value.id == "dist":
                        out.append("dist." + f.attr)
        return out

    fn = next(n for n in tree.body if isinstance(n, ast.FunctionDef) and n.name == "run_b200")
    bad = []
    for node in ast.walk(fn):
        if isinstance(node, ast.If) and mentions_rank(node.test):
            # `if rank != 0: return` (no collective after it in that branch) is fine; anything else is checked
            bad += [(node.lineno, c) for c in calls_in(node.body) + calls_in(node.orelse)]
    assert not bad, f"collective-bearing calls under a rank condition: {bad}"


def test_fused_adamw_device_lr_follows_a_schedule(emu_backend):
    """lr_tensor (what GraphedTrainStep(lr_scheduler=...) refreshes between replays) overrides the group's lr, so a
    captured optimizer kernel follows get_cosine_schedule_with_warmup (train_from_scratch.py:274-278, :113)."""
    from polyp_image_generator_b200 import FusedAdamW, UNet2DModel
    from polyp_image_generator_b200.training import get_cosine_schedule_with_warmup
    cfg = _small_cfg(32)
    torch.manual_seed(0)
    a, b = UNet2DModel(**cfg), UNet2DModel(**cfg)
    b.load_state_dict(a.state_dict())
    oa = FusedAdamW(a.parameters(), lr=1e-3, max_grad_norm=1.0)
    ob = FusedAdamW(b.parameters(), lr=1e-3, max_grad_norm=1.0)
    sa = get_cosine_schedule_with_warmup(oa, 2, 6)
    lr_t = torch.full((1,), float(oa.param_groups[0]["lr"]))     # LambdaLR already applied warm-up step 0 (lr = 0)
    ob.lr_tensor = lr_t
    ob.param_groups[0]["lr"] = 123.0            # must be ignored once the device scalar is set
    x, t, tgt = torch.randn(2, 3, 32, 32), torch.tensor([3, 600]), torch.randn(2, 3, 32, 32)
    for _ in range(4):
        for m in (a, b):
            torch.nn.functional.mse_loss(m(x, t).sample, tgt).backward()
        oa.step(); ob.step()
        oa.zero_grad(); ob.zero_grad()
        sa.step()
        ob.param_groups[0]["lr"] = oa.param_groups[0]["lr"]   # what lr_scheduler.step() leaves in the group ...
        lr_t.fill_(float(oa.param_groups[0]["lr"]))           # ... copied into the device scalar
        ob.param_groups[0]["lr"] = 123.0
    assert sa.get_last_lr()[0] == pytest.approx(1e-3 * 0.5 * (1 + __import__("math").cos(__import__("math").pi * 0.5)))
    for p, q in zip(a.parameters(), b.parameters()):
        assert torch.allclose(p, q, rtol=1e-6, atol=1e-7)


def test_train_loop_mirrors_train_from_scratch(emu_backend, tmp_path):
    """training.train_loop = train_from_scratch.py:68-133: epochs x batches of the step recipe, a pipeline per epoch,
    evaluate / save_pretrained at the listed epochs, files where the reference puts them."""
    from types import SimpleNamespace
    from polyp_image_generator_b200 import DDPMPipeline, DDPMScheduler, UNet2DModel
    from polyp_image_generator_b200.training import get_cosine_schedule_with_warmup, train_loop
    cfg = _small_cfg(32)
    cfg["block_out_channels"] = (64, 64, 64, 64, 64, 64)
    torch.manual_seed(0)
    m = UNet2DModel(**cfg)
    w0 = m.conv_out.weight.detach().clone()
    g = torch.Generator().manual_seed(1)
    loader = [(torch.randn(2, 3, 32, 32, generator=g).clamp(-1, 1), torch.zeros(2)) for _ in range(2)]
    conf = SimpleNamespace(output_dir=str(tmp_path / "run"), num_epochs=2, eval_batch_size=2, seed=0)
    opt = torch.optim.AdamW(m.parameters(), lr=1e-4)
    sched = get_cosine_schedule_with_warmup(opt, 1, 4)
    seen = []
    hist = train_loop(conf, m, DDPMScheduler(), opt, loader, sched, "AD", 3, evaluate_epochs=(1,), save_epochs=(1,),
                      num_inference_steps=2, on_epoch_end=lambda e, l, p: seen.append((e, type(p).__name__)))
    assert len(hist) == 2 and all(h > 0 for h in hist) and seen == [(0, "DDPMPipeline"), (1, "DDPMPipeline")]
    assert sched.last_epoch == 4 and not torch.equal(m.conv_out.weight.detach(), w0)
    samples = sorted(os.listdir(os.path.join(conf.output_dir, "samples", "AD")))
    assert samples == ["1.png", "2.png", "3.png"]
    saved = os.path.join(conf.output_dir, "models", "model_AD")
    assert sorted(os.listdir(saved)) == ["model_index.json", "scheduler", "unet"]
    back = DDPMPipeline.from_pretrained(saved)
    assert torch.equal(back.unet.conv_out.weight.detach(), m.conv_out.weight.detach())


def test_per_class_resume_plan(tmp_path):
    """train_with_lora_per_class.py:252-293: which classes still need training, full sampling, a top-up, or nothing."""
    from polyp_image_generator_b200.sampling import per_class_resume_plan
    root = tmp_path / "run"
    for d in ("lora_AD", "model_AD", "lora_HP", "model_HP", "lora_ASS", "model_ASS", "lora_X"):
        (root / d).mkdir(parents=True)
    (root / "samples" / "AD").mkdir(parents=True)
    for i in range(1, 6):
        (root / "samples" / "AD" / f"{i}.png").write_bytes(b"x")
    (root / "samples" / "HP").mkdir()
    for i in range(1, 4):
        (root / "samples" / "HP" / f"{i}.png").write_bytes(b"x")
    plan = per_class_resume_plan(str(root), ["AD", "HP", "ASS", "X", "NEW"], [5, 10, 7, 4, 2])
    assert plan == [("AD", "done", 0), ("HP", "top_up", 7), ("ASS", "generate", 7), ("X", "train", 4), ("NEW", "train", 2)]
    assert per_class_resume_plan(str(tmp_path / "missing"), ["AD"], [3]) == [("AD", "train", 3)]


def test_product_configs_equal_the_oracle_configs_and_bench_gpu_arm_does_not_import_the_oracle():
    import ast
    from polyp_image_generator_b200.model import celebahq_unet_config, polyp_unet_config
    for s in (64, 128, 256):
        assert polyp_unet_config(s) == oracle.polyp_unet_config(s)
        assert celebahq_unet_config(s) == oracle.celebahq_unet_config(s)
    # only the CPU-baseline leg of bench.py (cpu_oracle_train) may touch oracle/
    tree = ast.parse(open(os.path.join(ROOT, "bench.py")).read())
    for fn in (n for n in tree.body if isinstance(n, ast.FunctionDef)):
        uses = any(isinstance(n, (ast.Import, ast.ImportFrom)) and
                   any((a.name or "").split(".")[0] == "oracle" for a in n.names) or
                   (isinstance(n, ast.ImportFrom) and (n.module or "").split(".")[0] == "oracle")
                   for n in ast.walk(fn))
        assert uses == (fn.name == "cpu_oracle_train"), fn.name
    # and the product package never imports it
    pkg = os.path.join(ROOT, "polyp_image_generator_b200")
    for f in os.listdir(pkg):
        if f.endswith(".py"):
            src = open(os.path.join(pkg, f)).read()
            assert "import oracle" not in src and "from oracle" not in src, f


def test_capi_argument_validation_of_session3_entry_points(built_lib):
    """The entry points added for the optimizer, wide-head attention, samplers, input transform and the GroupNorm
    statistics fusion reject bad arguments before touching the device (negative rc + message, no GPU needed)."""
    from polyp_image_generator_b200 import _capi
    lib = _capi.load()
    fake = ctypes.c_void_p(0x1000)          # 16-byte aligned, never dereferenced: validation fails first
    odd = ctypes.c_void_p(0x1004)

    def bad(rc, needle):
        assert rc < 0 and needle in _capi.last_error(), (rc, _capi.last_error())

    bad(lib.ddpm_bgemm(None, 8, 0, 0, 0, fake, 8, 0, 0, 0, fake, 8, 0, 0, 0, 16, 16, 16, 1, 1, 1.0, None), "null pointer")
    bad(lib.ddpm_bgemm(fake, 8, 0, 0, 0, fake, 8, 0, 0, 0, fake, 8, 0, 0, 0, 0, 16, 16, 1, 1, 1.0, None), "bad shape")
    bad(lib.ddpm_bgemm(fake, 8, 0, 0, 0, fake, 8, 0, 0, 0, fake, 12, 0, 0, 0, 16, 16, 16, 1, 1, 1.0, None), "16-byte aligned")
    bad(lib.ddpm_bgemm(fake, 8, 0, 0, 0, fake, 8, 0, 0, 0, fake, 8, 0, 0, 0, 16, 16, 16, 70000, 1, 1.0, None), "gridDim.z")
    bad(lib.ddpm_softmax_rows(None, 8, fake, 8, 4, 8, None), "bad argument")
    bad(lib.ddpm_softmax_rows(fake, 4, fake, 8, 4, 8, None), "bad argument")              # lds < t
    bad(lib.ddpm_softmax_rows_bwd(fake, 8, None, 8, fake, 4, 8, 1.0, None), "bad argument")
    bad(lib.ddpm_resize_h_u8(None, fake, 4, 8, 3, 4, fake, fake, 3, None), "null pointer")
    bad(lib.ddpm_resize_h_u8(fake, fake, 4, 8, 2, 4, fake, fake, 3, None), "bad shape")   # c must be 1 or 3
    bad(lib.ddpm_resize_v_normalize(fake, fake, 1, 8, 8, 3, 0, fake, fake, 3, None, None), "bad shape")
    bad(lib.ddpm_ddim_step(None, fake, None, fake, None, 16, 1.0, 0.1, 1.0, 0.1, 0.0, 1.0, 0, None), "bad argument")
    bad(lib.ddpm_ddim_step(fake, odd, None, fake, None, 16, 1.0, 0.1, 1.0, 0.1, 0.0, 1.0, 0, None), "16-byte aligned")
    bad(lib.ddpm_adamw_flat(None, fake, fake, fake, 16, fake, None, 1.0, None, 1e-3, 0.9, 0.999, 1e-8, 0.01, None),
        "bad argument")
    bad(lib.ddpm_adamw_flat(fake, odd, fake, fake, 16, fake, None, 1.0, None, 1e-3, 0.9, 0.999, 1e-8, 0.01, None),
        "16-byte aligned")
    bad(lib.ddpm_sumsq_f32(odd, 16, fake, None), "16-byte aligned")
    bad(lib.ddpm_gn_stats_from_csum(None, 128, None, 0, 2, 32, fake, None), "bad argument")
    bad(lib.ddpm_gn_stats_from_csum(fake, 100, None, 0, 2, 32, fake, None), "bad argument")   # channels % groups
    bad(lib.ddpm_gn_bwd_dparams(None, None, 2, 128, 32, 64, 1e-5, fake, fake, None), "bad argument")
    a = _capi.ConvArgs()
    assert lib.ddpm_conv_gemm_workspace_elems(ctypes.byref(a)) == 0


def test_from_pretrained_reads_a_hub_style_celebahq_config(emu_backend, tmp_path):
    """The on-disk layout of google/ddpm-celebahq-256 (config.json as written by an old diffusers,
    lists instead of tuples, `attention_head_dim: null`, a `.bin` weight file) loads into the drop-in UNet."""
    import json
    from polyp_image_generator_b200 import UNet2DModel
    d = tmp_path / "unet"
    d.mkdir()
    hub = {
        "_class_name": "UNet2DModel", "_diffusers_version": "0.0.4", "act_fn": "silu", "attention_head_dim": None,
        "block_out_channels": [64, 64, 128, 128, 128, 128], "center_input_sample": False,
        "down_block_types": ["DownBlock2D", "DownBlock2D", "DownBlock2D", "DownBlock2D", "AttnDownBlock2D", "DownBlock2D"],
        "downsample_padding": 0, "flip_sin_to_cos": False, "freq_shift": 1, "in_channels": 3, "layers_per_block": 2,
        "mid_block_scale_factor": 1, "norm_eps": 1e-06, "norm_num_groups": 32, "out_channels": 3, "sample_size": 64,
        "time_embedding_type": "positional",
        "up_block_types": ["UpBlock2D", "AttnUpBlock2D", "UpBlock2D", "UpBlock2D", "UpBlock2D", "UpBlock2D"],
    }
    (d / "config.json").write_text(json.dumps(hub, indent=2))
    cfg = oracle.celebahq_unet_config(64)
    cfg["block_out_channels"] = (64, 64, 128, 128, 128, 128)
    torch.manual_seed(4)
    om = oracle.UNet2DModel(**cfg)
    torch.save(om.state_dict(), str(d / "diffusion_pytorch_model.bin"))
    m = UNet2DModel.from_pretrained(str(d))
    assert m.config.attention_head_dim is None and m.config.downsample_padding == 0 and m.config.freq_shift == 1
    assert list(m.state_dict().keys()) == list(om.state_dict().keys())
    x, t = torch.randn(1, 3, 64, 64), torch.tensor([77])
    with torch.no_grad():
        y, yo = m.eval()(x, t).sample, om.eval()(x, t).sample
    assert ((y - yo).norm() / yo.norm()).item() < 1e-5
    # and our own save -> load round trip keeps the architecture knobs
    m.save_pretrained(str(tmp_path / "again"))
    m2 = UNet2DModel.from_pretrained(str(tmp_path / "again"))
    assert vars(m2.config) == vars(m.config)


def test_from_pretrained_accepts_the_deprecated_attention_key_names(emu_backend, tmp_path):
    """Real google/ddpm-* hub checkpoints (the celebahq architecture of BASELINE configs[3]/[4]) predate diffusers'
    AttentionBlock -> Attention refactor: their attention weights are stored as `attentions.N.{query,key,value,
    proj_attn}`; diffusers renames them at load time (`_convert_deprecated_attention_blocks`) and so does the drop-in."""
    import json
    from polyp_image_generator_b200 import UNet2DModel
    from polyp_image_generator_b200.unet import convert_deprecated_attention_keys
    cfg = oracle.celebahq_unet_config(64)
    cfg["block_out_channels"] = (64, 64, 128, 128, 128, 128)
    torch.manual_seed(4)
    om = oracle.UNet2DModel(**cfg)
    ren = {".to_q.": ".query.", ".to_k.": ".key.", ".to_v.": ".value.", ".to_out.0.": ".proj_attn."}
    old_sd = {}
    for k, v in om.state_dict().items():
        for new, old in ren.items():
            if ".attentions." in k and new in k:
                k = k.replace(new, old)
        old_sd[k] = v
    assert sum(".query." in k or ".proj_attn." in k for k in old_sd) == 6 * 2 * 2
    assert set(convert_deprecated_attention_keys(old_sd)) == set(om.state_dict())
    d = tmp_path / "unet"
    d.mkdir()
    hub = {k: (list(v) if isinstance(v, tuple) else v) for k, v in cfg.items()}
    hub.update({"_class_name": "UNet2DModel", "_diffusers_version": "0.0.4"})
    (d / "config.json").write_text(json.dumps(hub))
    torch.save(old_sd, str(d / "diffusion_pytorch_model.bin"))
    m = UNet2DModel.from_pretrained(str(d))
    for (n, p), (no, po) in zip(m.named_parameters(), om.named_parameters()):
        assert n == no and torch.equal(p.detach(), po.detach()), n
    x, t = torch.randn(1, 3, 64, 64), torch.tensor([10])
    with torch.no_grad():
        assert torch.allclose(m(x, t).sample, om(x, t).sample, rtol=1e-3, atol=1e-4)
    # a freshly built model takes the old names through load_state_dict as well; new names pass through untouched
    m2 = UNet2DModel(**cfg)
    assert not m2.load_state_dict(old_sd).missing_keys
    assert not m2.load_state_dict(om.state_dict()).unexpected_keys


def test_fused_adamw_state_dict_round_trip_resumes_moments_and_step(emu_backend):
    """A restored FusedAdamW continues from the saved moments and step count (not from zero)."""
    from polyp_image_generator_b200 import FusedAdamW, UNet2DModel
    cfg = _small_cfg(32)
    torch.manual_seed(0)
    a, b = UNet2DModel(**cfg), UNet2DModel(**cfg)
    b.load_state_dict(a.state_dict())
    x, t, tgt = torch.randn(2, 3, 32, 32), torch.tensor([3, 600]), torch.randn(2, 3, 32, 32)
    oa = FusedAdamW(a.parameters(), lr=3e-4, max_grad_norm=1.0)

    def one(m, o):
        torch.nn.functional.mse_loss(m(x, t).sample, tgt).backward()
        o.step()
        o.zero_grad()

    for _ in range(2):
        one(a, oa)
    b.load_state_dict(a.state_dict())
    ob = FusedAdamW(b.parameters(), lr=3e-4, max_grad_norm=1.0)
    ob.load_state_dict(oa.state_dict())
    one(a, oa)
    one(b, ob)
    assert float(ob._scal[0]) == 3.0
    for (n, p), (_, q) in zip(a.named_parameters(), b.named_parameters()):
        assert torch.allclose(p, q, rtol=0, atol=1e-7), n


def test_ddp_skips_arena_buckets_for_frozen_models_and_resets_progress():
    """ddp.DistributedDataParallel: LoRA runs (frozen arena) exchange no arena buckets, and bucket progress left by an
    aborted backward is forgotten at the start of the next one."""
    from types import SimpleNamespace
    from polyp_image_generator_b200.ddp import DistributedDataParallel as D
    fake = D.__new__(D)
    torch.nn.Module.__init__(fake)
    p = torch.nn.Parameter(torch.zeros(4), requires_grad=False)
    fake.module = SimpleNamespace(_plan=SimpleNamespace(layout=[(p, 0, None)], temb_w_off=1 << 30))
    fake.bucket_elems, fake._done_upto, fake._arena_trainable = 1, 12345, None
    calls = []
    fake._allreduce_mean = lambda t: calls.append(t)
    fake._stream_ctx = lambda G: None
    D._begin(fake)
    assert fake._done_upto is None
    D._progress(fake, torch.zeros(8), 0)
    assert calls == [] and fake._done_upto is None


def test_fp32_faithful_inference_program_matches_fp32_oracle(emu_backend):
    """inference_precision = "fp32" (north_star: eps within 1e-4 of the fp32 reference): the split-bf16 forward program
    -- K-concat [hi | lo | hi] x [W_hi | W_hi | W_lo] convs, split GroupNorm / attention, two-call shortcut over a
    channel concat, stride-2 taps, LoRA folded into the effective weight -- against the fp32 oracle through the CPU
    emulation of the op contracts."""
    from polyp_image_generator_b200 import LoraConfig, UNet2DModel
    for variant in ("polyp", "polyp_lora"):
        cfg = _small_cfg(32)
        torch.manual_seed(3)
        om = oracle.UNet2DModel(**cfg)
        m = UNet2DModel(**cfg)
        m.load_state_dict(om.state_dict())
        if variant == "polyp_lora":
            oracle.add_adapter(om, oracle.LoraConfig(r=8, lora_alpha=8, init_lora_weights="gaussian"))
            m.add_adapter(LoraConfig(r=8, lora_alpha=8, init_lora_weights="gaussian"))
            sd = {k: torch.randn_like(v) * 0.05 for k, v in oracle.lora_state_dict(om).items()}
            om.load_state_dict(sd, strict=False)
            m.load_state_dict(sd, strict=False)
        m.eval()
        m.inference_precision = "fp32"
        x, t = torch.randn(2, 3, 32, 32), torch.tensor([7, 903])
        with torch.no_grad():
            got = m(x, t).sample
            want = om(x, t).sample
        r = ((got - want).norm() / want.norm()).item()
        assert r < 1e-4, (variant, r)
        m.inference_precision = "bf16"
        with torch.no_grad():
            assert ((m(x, t).sample - want).norm() / want.norm()).item() < 1e-3     # emulation runs the bf16 path in fp32


def test_zero_pool_hands_out_disjoint_zero_slices_and_grows():
    """unet._ZeroPool: one fill per pass; slices are zero, 128-byte aligned and disjoint; a pass that needs more than the
    previous one gets fresh zero tensors and the next pass a larger buffer; buffers of earlier passes are never reused
    (gradients of an accumulating loop may still alias them)."""
    from polyp_image_generator_b200.unet import _ZeroPool
    dev = torch.device("cpu")
    zp = _ZeroPool().begin(dev)
    a, b = zp.take((3, 5), dev), zp.take((7,), dev)          # first pass: nothing learned yet -> plain zero tensors
    assert a.shape == (3, 5) and b.shape == (7,) and not a.any() and not b.any()
    a.fill_(1.0)
    zp.begin(dev)
    assert zp.need == 2 * _ZeroPool.ALIGN
    c, d = zp.take((3, 5), dev), zp.take((7,), dev)
    assert c.untyped_storage().data_ptr() == d.untyped_storage().data_ptr() == zp.buf.untyped_storage().data_ptr()
    assert (d.data_ptr() - c.data_ptr()) == 4 * _ZeroPool.ALIGN and not c.any() and not d.any()
    c.fill_(2.0)
    assert not d.any()
    e = zp.take((100,), dev)                                  # more than the pass had last time: own tensor
    assert e.untyped_storage().data_ptr() != zp.buf.untyped_storage().data_ptr() and not e.any()
    old = zp.buf
    zp.begin(dev)
    assert zp.buf is not old and zp.buf.numel() >= 2 * _ZeroPool.ALIGN + 100 and not zp.buf.any()
    assert (c == 2.0).all()                                   # the previous pass's slices are left alone


def test_bias_gradient_fusion_equals_separate_reductions(emu_backend, monkeypatch):
    """unet._colsum_target: the bias gradient of the conv that consumes a block's result rides on the block's last
    GroupNorm-backward pass (out_c).  Same gradients as the separate column reductions, and fewer reduce_hw launches."""
    from polyp_image_generator_b200 import UNet2DModel

    def grads(fusion):
        monkeypatch.setenv("DDPM_BIAS_FUSION", fusion)
        torch.manual_seed(11)
        m = UNet2DModel(**_small_cfg(32))
        x, t = torch.randn(2, 3, 32, 32), torch.tensor([5, 700])
        calls = {"n": 0}
        real = emu_backend.reduce_hw

        def counted(*a, **k):
            calls["n"] += 1
            return real(*a, **k)
        monkeypatch.setattr(emu_backend, "reduce_hw", counted)
        m(x, t).sample.square().mean().backward()
        monkeypatch.setattr(emu_backend, "reduce_hw", real)
        return {n: p.grad.clone() for n, p in m.named_parameters()}, calls["n"]

    g1, n1 = grads("1")
    g0, n0 = grads("0")
    assert n1 < n0, (n1, n0)
    for k in g0:
        assert torch.allclose(g1[k], g0[k], rtol=1e-4, atol=1e-6), k
