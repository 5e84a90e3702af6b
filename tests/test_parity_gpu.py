"""GPU parity tests (-m gpu): the sm_100a kernels, called through the C-ABI, against the CPU oracle on identical
inputs, seeds and noise tensors.  Tolerances are BASELINE.json's north_star: scheduler / elementwise ops 1e-6
relative in fp32 (we assert bit-exactness where the op order is reproducible), UNet eps-prediction and gradients
2e-2 relative with bf16 tensor-core compute.  Nothing here reads /root/reference."""
import os

import pytest
import torch
import torch.nn.functional as F

import oracle

pytestmark = pytest.mark.gpu

GOLD = os.path.join(os.path.dirname(__file__), "golden")


def rel(a, b):
    a, b = a.detach().float().cpu(), b.detach().float().cpu()
    return ((a - b).norm() / (b.norm() + 1e-30)).item()


@pytest.fixture(scope="module")
def dev():
    assert torch.cuda.is_available(), "GPU tests need a CUDA device"
    from polyp_image_generator_b200 import ops
    assert ops.get().name == "cuda"
    return torch.device("cuda:0")


# ---- elementwise / scheduler -----------------------------------------------------------------------------------
@pytest.mark.parametrize("shape", [(4, 3, 64, 64), (64, 3, 128, 128), (3, 3, 7, 5), (1, 3, 1, 1)])
def test_add_noise_bit_exact(dev, shape):
    from polyp_image_generator_b200 import DDPMScheduler
    torch.manual_seed(0)
    x0, nz = torch.randn(shape), torch.randn(shape)
    t = torch.randint(0, 1000, (shape[0],))
    want = oracle.DDPMScheduler().add_noise(x0, nz, t)
    got = DDPMScheduler().add_noise(x0.to(dev), nz.to(dev), t.to(dev))
    assert torch.equal(got.cpu(), want)


def test_add_noise_empty_and_errors(dev):
    from polyp_image_generator_b200 import DDPMScheduler
    s = DDPMScheduler()
    e = torch.zeros(0, 3, 8, 8, device=dev)
    assert s.add_noise(e, e, torch.zeros(0, dtype=torch.int64, device=dev)).shape == (0, 3, 8, 8)
    with pytest.raises(ValueError):
        s.add_noise(torch.zeros(2, 3, 8, 8, device=dev), torch.zeros(2, 3, 8, 8, device=dev),
                    torch.zeros(3, dtype=torch.int64, device=dev))


@pytest.mark.parametrize("n_steps", [1000, 50])
def test_scheduler_step_bit_exact(dev, n_steps):
    from polyp_image_generator_b200 import DDPMScheduler
    a, b = DDPMScheduler(), oracle.DDPMScheduler()
    a.set_timesteps(n_steps)
    b.set_timesteps(n_steps)
    torch.manual_seed(1)
    x, e, z = torch.randn(4, 3, 32, 32), torch.randn(4, 3, 32, 32), torch.randn(4, 3, 32, 32)
    ts = a.timesteps.tolist()
    for t in [ts[0], ts[len(ts) // 2], ts[-2], ts[-1]]:
        want = b.step(e, torch.tensor(t), x, variance_noise=z)
        got = a.step(e.to(dev), t, x.to(dev), variance_noise=z.to(dev))
        assert torch.equal(got.prev_sample.cpu(), want.prev_sample), t
        assert torch.equal(got.pred_original_sample.cpu(), want.pred_original_sample), t


@pytest.mark.parametrize("n_steps", [50, 1000])
def test_ddim_step_bit_exact(dev, n_steps):
    """DDIMScheduler.step in one kernel, bit-identical to the oracle's torch expressions."""
    from polyp_image_generator_b200 import DDIMScheduler
    a, b = DDIMScheduler(), oracle.DDIMScheduler()
    a.set_timesteps(n_steps)
    b.set_timesteps(n_steps)
    torch.manual_seed(2)
    x, e, z = torch.randn(4, 3, 32, 32), torch.randn(4, 3, 32, 32), torch.randn(4, 3, 32, 32)
    ts = a.timesteps.tolist()
    for t in [ts[0], ts[len(ts) // 2], ts[-1]]:
        for eta, clipped in ((0.0, False), (0.0, True), (0.7, False), (1.0, True)):
            zz = z if eta > 0 else None
            want = b.step(e, torch.tensor(t), x, eta=eta, use_clipped_model_output=clipped, variance_noise=zz)
            got = a.step(e.to(dev), t, x.to(dev), eta=eta, use_clipped_model_output=clipped,
                         variance_noise=zz.to(dev) if zz is not None else None)
            assert torch.equal(got.prev_sample.cpu(), want.prev_sample), (t, eta, clipped)
            assert torch.equal(got.pred_original_sample.cpu(), want.pred_original_sample), (t, eta, clipped)
    # ragged size (scalar tail) and the scaled_linear table
    a2 = DDIMScheduler(beta_schedule="scaled_linear", beta_start=0.00085, beta_end=0.012, clip_sample=False)
    b2 = oracle.DDIMScheduler(beta_schedule="scaled_linear", beta_start=0.00085, beta_end=0.012, clip_sample=False)
    a2.set_timesteps(25)
    b2.set_timesteps(25)
    x, e = torch.randn(1, 3, 7, 5), torch.randn(1, 3, 7, 5)
    want = b2.step(e, torch.tensor(480), x).prev_sample
    assert torch.equal(a2.step(e.to(dev), 480, x.to(dev)).prev_sample.cpu(), want)


@pytest.mark.parametrize("n_steps", [25, 3])
def test_unipc_step_bit_exact(dev, n_steps):
    """UniPCMultistepScheduler.step (train_with_lora_all_classes.py:314): x0 conversion, corrector
    and predictor kernels, whole trajectories bit-identical to the oracle's torch expressions; ragged size; golden."""
    import json
    from polyp_image_generator_b200 import UniPCMultistepScheduler
    sd = dict(beta_schedule="scaled_linear", beta_start=0.00085, beta_end=0.012, steps_offset=1)
    for kw, shape in ((dict(), (4, 3, 32, 32)), (sd, (1, 3, 7, 5)), (dict(solver_order=1), (2, 3, 16, 16))):
        a, b = UniPCMultistepScheduler(**kw), oracle.UniPCMultistepScheduler(**kw)
        a.set_timesteps(n_steps)
        b.set_timesteps(n_steps)
        g = torch.Generator().manual_seed(5)
        xb = torch.randn(shape, generator=g)
        xa = xb.to(dev)
        for t in a.timesteps.tolist():
            eps = torch.randn(shape, generator=g)
            xa = a.step(eps.to(dev), t, xa).prev_sample
            xb = b.step(eps, torch.tensor(t), xb).prev_sample
            assert torch.equal(xa.cpu(), xb), (kw, t)
    gold = json.load(open(os.path.join(GOLD, "unipc.json")))
    s = UniPCMultistepScheduler(**gold["kwargs"])
    s.set_timesteps(gold["steps"])
    assert s.timesteps.tolist() == gold["timesteps"]
    x = torch.randn(1, 3, 8, 8, generator=torch.Generator().manual_seed(13)).to(dev)
    for t, want in zip(s.timesteps.tolist(), gold["trajectory_sums"]):
        x = s.step(torch.cos(x.cpu() * 2.0 - float(t) * 0.02).to(dev), t, x).prev_sample
        assert x.double().sum().item() == pytest.approx(want, rel=1e-5, abs=1e-5)
    assert torch.allclose(x.cpu().flatten(), torch.tensor(gold["final"]), rtol=1e-5, atol=1e-6)


def test_ddim_golden_trajectory_on_device(dev):
    """The committed oracle trajectory (tests/golden/ddim.json) through the DDIM step kernel: bit-identical op order, so
    the trajectory sums agree to fp32 rounding of the stand-in network."""
    import json
    from polyp_image_generator_b200 import DDIMScheduler
    gold = json.load(open(os.path.join(GOLD, "ddim.json")))
    s = DDIMScheduler()
    s.set_timesteps(25)
    assert s.timesteps.tolist() == gold["timesteps"]
    g = torch.Generator().manual_seed(11)
    x = torch.randn(1, 3, 8, 8, generator=g).to(dev)
    for t, want in zip(s.timesteps.tolist(), gold["trajectory_sums"]):
        eps = torch.cos(x.cpu() * 2.0 - float(t) * 0.02).to(dev)      # the stand-in runs on the CPU like the oracle's
        x = s.step(eps, t, x, eta=gold["eta"], use_clipped_model_output=True, generator=g).prev_sample
        assert x.double().sum().item() == pytest.approx(want, rel=1e-5, abs=1e-5)
    assert torch.allclose(x.cpu().flatten(), torch.tensor(gold["final"]), rtol=1e-5, atol=1e-6)


def test_scheduler_golden_trajectory(dev):
    """50-step trajectory with a deterministic stand-in for the UNet; CPU generator consumed in diffusers' order."""
    import json
    from polyp_image_generator_b200 import DDPMScheduler
    gold = json.load(open(os.path.join(GOLD, "scheduler.json")))
    s = DDPMScheduler()
    s.set_timesteps(50)
    assert s.timesteps.tolist() == gold["timesteps"]
    g = torch.Generator().manual_seed(7)
    x = torch.randn(1, 3, 8, 8, generator=g).to(dev)
    for i, t in enumerate(s.timesteps.tolist()):
        eps = torch.sin(x.cpu() * 3.0 + float(t) * 0.01).to(dev)
        x = s.step(eps, t, x, generator=g).prev_sample
        assert x.double().sum().item() == pytest.approx(gold["trajectory_sums"][i], rel=1e-6, abs=1e-6)
    assert torch.allclose(x.cpu().flatten(), torch.tensor(gold["final"]), rtol=1e-5, atol=1e-6)


def test_mse_loss_fwd_bwd(dev):
    from polyp_image_generator_b200.training import mse_loss
    torch.manual_seed(2)
    for shape in [(64, 3, 128, 128), (2, 3, 5, 7)]:
        p, q = torch.randn(shape), torch.randn(shape)
        pr = p.clone().requires_grad_(True)
        lo = F.mse_loss(pr, q)
        (lo * 3.0).backward()
        pg = p.to(dev).requires_grad_(True)
        l = mse_loss(pg, q.to(dev))
        (l * 3.0).backward()
        assert l.item() == pytest.approx(lo.item(), rel=2e-6)
        assert rel(pg.grad, pr.grad) < 1e-6


def test_philox_step_statistics_and_linearity(dev):
    """In-kernel noise: z ~ N(0,1) (moments), deterministic per (seed, offset), and prev is affine in sigma."""
    from polyp_image_generator_b200 import ops
    o = ops.get()
    n = 1 << 22
    zero = torch.zeros(n, device=dev)
    z1 = o.scheduler_step_philox(zero, zero, 1.0, 0.0, 0.0, 0.0, 1.0, 0.0, 99, 1)
    z1b = o.scheduler_step_philox(zero, zero, 1.0, 0.0, 0.0, 0.0, 1.0, 0.0, 99, 1)
    z2 = o.scheduler_step_philox(zero, zero, 1.0, 0.0, 0.0, 0.0, 1.0, 0.0, 99, 2)
    assert torch.equal(z1, z1b) and not torch.equal(z1, z2)
    assert abs(z1.mean().item()) < 3e-3 and abs(z1.std().item() - 1) < 3e-3
    assert abs((z1 ** 4).mean().item() - 3.0) < 0.05
    assert abs((z1 * z2).mean().item()) < 3e-3
    half = o.scheduler_step_philox(zero, zero, 1.0, 0.0, 0.0, 0.0, 0.5, 0.0, 99, 1)
    assert torch.allclose(half, 0.5 * z1, rtol=1e-6, atol=1e-7)


# ---- UNet -----------------------------------------------------------------------------------------------------
def _small_cfg(S=32):
    cfg = oracle.polyp_unet_config(S)
    cfg["block_out_channels"] = (64, 64, 128, 128, 128, 128)
    return cfg


def _grad_report(m, om, x, t, noise):
    """-> (whole-gradient rel error vs the fp32 oracle, worst per-tensor noise-floor ratio, its name).  The ratio is
    (product vs fp32 oracle) / max(rounding-matched oracle vs fp32 oracle, 5e-3) per tensor, over every tensor that
    carries >= 1e-6 of the gradient norm: see tests/parity_util.py and tests/test_parity_full_gpu.py."""
    from parity_util import _grad_table, _noise_floor_ratio, _oracle_grads
    _, g_o = _oracle_grads(om, x, t, noise, rounded=False)
    _, g_r = _oracle_grads(om, x, t, noise, rounded=True)
    whole, rows = _grad_table(m, g_o)
    ratio, name, _ = _noise_floor_ratio(rows, g_r, g_o)
    return whole, ratio, name


def test_unet_golden_small_case(dev):
    """Committed golden vector (tests/golden/make_golden.py): eps prediction, loss and gradients."""
    from golden.make_golden import small_cfg
    from polyp_image_generator_b200 import UNet2DModel
    from polyp_image_generator_b200.training import mse_loss
    gold = torch.load(os.path.join(GOLD, "unet32_fwd_bwd.pt"))
    torch.manual_seed(1234)
    om = oracle.UNet2DModel(**small_cfg())   # same seed -> same init as the golden generator
    m = UNet2DModel(**small_cfg())
    m.load_state_dict(om.state_dict())
    m.to(dev).train()
    pred = m(gold["noisy"].to(dev), gold["t"].to(dev)).sample
    assert rel(pred, gold["pred"]) < 2e-2
    loss = mse_loss(pred, gold["noise"].to(dev))
    assert loss.item() == pytest.approx(gold["loss"], rel=2e-2)
    loss.backward()
    params = dict(m.named_parameters())
    tot = sum(v ** 2 for v in gold["grad_norms"].values()) ** 0.5
    num = 0.0
    for n, p in params.items():      # whole-gradient norm check via the stored per-tensor norms
        num += (p.grad.norm().item() - gold["grad_norms"][n]) ** 2
    assert (num ** 0.5) / tot < 2e-2
    for k, g in gold["grads"].items():       # the stored tensors that carry >= 1 % of the gradient norm each
        if g.norm().item() >= 1e-2 * tot:
            print(f"golden grad {k}: rel {rel(params[k].grad, g):.4f}")
            assert rel(params[k].grad, g) < 2e-2, k


@pytest.mark.parametrize("variant,S,B", [("polyp_small", 32, 3), ("celebahq_small", 64, 2), ("celebahq_1head", 64, 3),
                                         ("polyp_full", 64, 4), ("celebahq_full", 64, 2)])
def test_unet_forward_backward_vs_oracle(dev, variant, S, B):
    from polyp_image_generator_b200 import UNet2DModel
    from polyp_image_generator_b200.training import mse_loss
    if variant == "polyp_small":
        cfg = _small_cfg(S)
    elif variant == "celebahq_small":
        cfg = oracle.celebahq_unet_config(S)
        cfg["block_out_channels"] = (64, 64, 128, 128, 128, 128)
        cfg["attention_head_dim"] = 16
    elif variant == "celebahq_1head":          # attention_head_dim=None: ONE head as wide as the block (bgemm.cu path)
        cfg = oracle.celebahq_unet_config(S)
        cfg["block_out_channels"] = (64, 64, 128, 128, 128, 128)
    elif variant == "celebahq_full":           # BASELINE configs[3]/[4] architecture (1 head x 512), reduced resolution
        cfg = oracle.celebahq_unet_config(S)
    else:
        cfg = oracle.polyp_unet_config(S)      # BASELINE configs[0]: the real 113.7 M-parameter model at 64x64, batch 4
    torch.manual_seed(0)
    om = oracle.UNet2DModel(**cfg)
    m = UNet2DModel(**cfg)
    m.load_state_dict(om.state_dict())
    m.to(dev).train()
    x = torch.randn(B, 3, S, S)
    t = torch.randint(0, 1000, (B,))
    noise = torch.randn(B, 3, S, S)
    pred = m(x.to(dev), t.to(dev), return_dict=False)[0]
    pred_o = om(x, t).sample
    assert pred.dtype == torch.float32 and pred.shape == pred_o.shape
    assert rel(pred, pred_o) < 2e-2
    loss = mse_loss(pred, noise.to(dev))
    loss.backward()
    total, worst, name = _grad_report(m, om, x, t, noise)
    assert total < 2e-2, f"whole-gradient rel error {total}"
    assert worst < 3.0, f"{name}: {worst} x the bf16-storage noise floor of that tensor"
    # inference path (no tape) gives the same prediction; python-int timestep broadcast
    with torch.no_grad():
        p2 = m(x.to(dev), t.to(dev)).sample
        p3 = m(x[:1].to(dev), int(t[0])).sample
    # fp32-atomic GroupNorm statistics make runs differ in the last bit; at random init the ~70-layer network
    # amplifies those bf16 rounding flips, so run-to-run agreement is at the bf16 noise level, not bit-exact
    assert rel(p2, pred) < 1.5e-2
    assert rel(p3, pred_o[:1]) < 2e-2


def test_unet_linearity_of_backward_and_gradscaler_compat(dev):
    """Size-independent property: gradients scale linearly with the loss scale (GradScaler's 65536)."""
    from polyp_image_generator_b200 import UNet2DModel
    from polyp_image_generator_b200.training import mse_loss
    torch.manual_seed(3)
    m = UNet2DModel(**_small_cfg(32)).to(dev).train()
    x, t, nz = torch.randn(2, 3, 32, 32, device=dev), torch.tensor([1, 999], device=dev), torch.randn(2, 3, 32, 32, device=dev)
    mse_loss(m(x, t).sample, nz).backward()
    g1 = torch.cat([p.grad.reshape(-1) for p in m.parameters()]).clone()
    m.zero_grad()
    scaler = torch.amp.GradScaler("cuda")
    scaler.scale(mse_loss(m(x, t).sample, nz)).backward()
    g2 = torch.cat([p.grad.reshape(-1) for p in m.parameters()])
    assert torch.isfinite(g2).all()
    assert rel(g2 / 65536.0, g1) < 2e-2
    opt = torch.optim.AdamW(m.parameters(), lr=1e-4)
    torch.nn.utils.clip_grad_norm_(m.parameters(), 1.0)
    scaler.step(opt)
    scaler.update()


def test_training_reduces_loss(dev):
    """End-to-end sanity on the drop-in objects: a few AdamW steps on a fixed batch reduce the loss."""
    from polyp_image_generator_b200 import DDPMScheduler, UNet2DModel
    from polyp_image_generator_b200.training import train_step
    torch.manual_seed(4)
    m = UNet2DModel(**_small_cfg(32)).to(dev).train()
    opt = torch.optim.AdamW(m.parameters(), lr=2e-4)
    s = DDPMScheduler()
    x0 = torch.randn(8, 3, 32, 32, device=dev).clamp(-1, 1)
    nz = torch.randn(8, 3, 32, 32, device=dev)
    t = torch.randint(0, 1000, (8,), device=dev)
    losses = [train_step(m, s, opt, x0, nz, t).item() for _ in range(12)]
    assert losses[-1] < 0.7 * losses[0], losses


@pytest.mark.parametrize("m,n,k,heads,batch", [(256, 256, 512, 1, 3), (64, 64, 512, 1, 2), (16, 16, 128, 2, 2),
                                               (49, 49, 72, 3, 2), (196, 200, 136, 1, 2), (300, 520, 264, 2, 1)])
def test_bgemm_all_layouts(dev, m, n, k, heads, batch):
    """ddpm_bgemm (batched tcgen05 GEMM of the wide-head attention): the four operand layouts, ragged edges
    (TMA zero fill), head / batch strides, bf16 and fp32 outputs, against a float64 matmul of the same bf16 inputs."""
    from polyp_image_generator_b200 import ops as ops_mod
    o = ops_mod.get()
    g = torch.Generator(device=dev).manual_seed(m * 131 + n * 7 + k)
    r8 = lambda x: (x + 7) // 8 * 8
    for a_mn in (0, 1):
        for b_mn in (0, 1):
            # storage: K-major A is [batch][heads][m][k8], MN-major A is [batch][heads][k][m8] (likewise B)
            ash = (batch, heads, k, r8(m)) if a_mn else (batch, heads, m, r8(k))
            bsh = (batch, heads, k, r8(n)) if b_mn else (batch, heads, n, r8(k))
            A = torch.randn(ash, device=dev, generator=g).to(torch.bfloat16)
            Bm = torch.randn(bsh, device=dev, generator=g).to(torch.bfloat16)
            Al = (A[..., :m].transpose(-1, -2) if a_mn else A[..., :k]).double()      # logical m x k
            Bl = (Bm[..., :n] if b_mn else Bm[..., :k].transpose(-1, -2)).double()    # logical k x n
            want = 0.5 * (Al @ Bl)
            for dt in (torch.float32, torch.bfloat16):
                C = torch.full((batch, heads, m, r8(n)), float("nan"), device=dev, dtype=dt)
                o.bgemm(A, ash[3], ash[2] * ash[3], heads * ash[2] * ash[3], a_mn,
                        Bm, bsh[3], bsh[2] * bsh[3], heads * bsh[2] * bsh[3], b_mn,
                        C, r8(n), m * r8(n), heads * m * r8(n), m, n, k, heads, batch, alpha=0.5)
                got = C[..., :n].double()
                tol = 1e-5 if dt == torch.float32 else 6e-3
                assert torch.isfinite(got).all(), (a_mn, b_mn, dt)
                assert ((got - want).norm() / want.norm()).item() < tol, (a_mn, b_mn, dt)
                if r8(n) > n:      # columns past n are never written
                    assert torch.isnan(C[..., n:].float()).all()


@pytest.mark.parametrize("b,t,heads,d", [(3, 256, 1, 512), (2, 64, 1, 512), (2, 16, 1, 512), (2, 49, 2, 128),
                                         (1, 196, 1, 256)])
def test_wide_head_attention_vs_sdpa(dev, b, t, heads, d):
    """AttnProcessor2_0's scaled_dot_product_attention for wide heads (celebahq: 1 x 512), forward and backward, against
    torch SDPA in fp32 on the CPU (same bf16 qkv)."""
    from polyp_image_generator_b200 import ops as ops_mod
    o = ops_mod.get()
    torch.manual_seed(t + d)
    C = heads * d
    qkv = (torch.randn(b * t, 3 * C) * 0.7).to(torch.bfloat16)
    d_o = torch.randn(b * t, C).to(torch.bfloat16)
    scale = d ** -0.5
    out, aux = o.attn_fwd(qkv.to(dev), b, t, heads, d, scale)
    dqkv = o.attn_bwd(qkv.to(dev), out, d_o.to(dev), aux, b, t, heads, d, scale)
    ref = qkv.float().clone().requires_grad_(True)
    q, k, v = [x.reshape(b, t, heads, d).transpose(1, 2) for x in ref.split(C, dim=1)]
    want = F.scaled_dot_product_attention(q, k, v, scale=scale)
    want = want.transpose(1, 2).reshape(b * t, C)
    want.backward(d_o.float())
    assert rel(out, want) < 1e-2
    for i, name in enumerate("qkv"):
        assert rel(dqkv[:, i * C:(i + 1) * C], ref.grad[:, i * C:(i + 1) * C]) < 1.5e-2, name


@pytest.mark.parametrize("b,t,heads,d", [(8, 256, 1, 512), (8, 64, 1, 512), (2, 256, 2, 256), (1, 130, 1, 128),
                                         (1, 100, 3, 384), (2, 7, 1, 128), (1, 256, 1, 1024), (3, 129, 2, 128)])
def test_fused_wide_head_attention_matches_unfused_and_sdpa(dev, b, t, heads, d, monkeypatch):
    """attn_wide.cu (ONE tcgen05 kernel: S in TMEM, softmax into shared memory, O = P V) against the three-launch
    composition it replaces (bgemm + softmax_rows + bgemm) and against fp32 SDPA: output, the probabilities kept for
    the backward pass, and the inference call that keeps none."""
    from polyp_image_generator_b200 import ops as ops_mod
    o = ops_mod.get()
    assert o.lib.ddpm_attn_wide_supported(t, heads, d) == 1
    torch.manual_seed(3 * t + d)
    C = heads * d
    qkv = (torch.randn(b * t, 3 * C) * 0.7).to(torch.bfloat16)
    scale = d ** -0.5
    monkeypatch.setenv("DDPM_ATTN_FUSED", "0")
    l0 = o.launches
    out_u, p_u = o.attn_fwd(qkv.to(dev), b, t, heads, d, scale)
    monkeypatch.setenv("DDPM_ATTN_FUSED", "1")
    l1 = o.launches
    out_f, p_f = o.attn_fwd(qkv.to(dev), b, t, heads, d, scale)
    assert o.launches - l1 == 1 and l1 - l0 > 1          # the fused path IS one launch
    out_i, p_i = o.attn_fwd(qkv.to(dev), b, t, heads, d, scale, need_aux=False)
    assert p_i is None and torch.equal(out_i, out_f)
    q, k, v = [x.reshape(b, t, heads, d).transpose(1, 2) for x in qkv.float().split(C, dim=1)]
    want = F.scaled_dot_product_attention(q, k, v, scale=scale).transpose(1, 2).reshape(b * t, C)
    want_p = torch.softmax(q @ k.transpose(-1, -2) * scale, dim=-1)
    assert rel(out_f, want) < 1e-2
    assert rel(out_f, out_u) < 6e-3                      # both round P and O to bf16 once
    assert rel(p_f[..., :t], want_p) < 6e-3
    assert rel(p_f[..., :t], p_u[..., :t]) < 6e-3
    assert torch.isfinite(out_f.float()).all()


@pytest.mark.parametrize("hw,cin,cout,k3,c1", [(64, 128, 128, True, 0), (64, 128, 256, False, 0), (32, 256, 256, True, 0),
                                               (128, 128, 128, True, 64), (16, 128, 128, True, 0)])
def test_conv_epilogue_statistics_and_single_pass_groupnorm(dev, hw, cin, cout, k3, c1):
    """conv_gemm(csum=...): the producing conv reduces the per-(sample, channel) moments of its stored output (halo and
    generic kernels); gn_fwd_from_csum then equals the two-phase gn_fwd (stats, y and the backward coef table)."""
    from polyp_image_generator_b200 import ops as ops_mod
    from polyp_image_generator_b200.ops import taps_1x1, taps_3x3
    o = ops_mod.get()
    torch.manual_seed(hw + cout)
    n = 3
    x0 = (torch.randn(n, hw, hw, cin - c1, device=dev) * 0.5).to(torch.bfloat16)
    x1 = (torch.randn(n, hw, hw, c1, device=dev) * 0.5).to(torch.bfloat16) if c1 else None
    taps = taps_3x3(cin) if k3 else taps_1x1()
    w = (torch.randn(cout, len(taps) * cin, device=dev) * 0.05).to(torch.bfloat16)
    bias = torch.randn(cout, device=dev)
    res = (torch.randn(n, hw, hw, cout, device=dev)).to(torch.bfloat16)
    cs = torch.zeros(n, cout // 4, 2, device=dev)          # per (sample, 4-channel granule)
    y = o.conv_gemm(x0, x1, taps, w, cout, (n, hw, hw), bias=bias, res=res, csum=cs)
    y_plain = o.conv_gemm(x0, x1, taps, w, cout, (n, hw, hw), bias=bias, res=res)
    # the statistics do not disturb the output (not bit-equal: the plain launch may take the split-K path, which
    # accumulates in another order)
    assert rel(y, y_plain) < 4e-3
    yf = y.float()
    want = torch.stack([yf.sum((1, 2)), (yf * yf).sum((1, 2))], -1).reshape(n, cout // 4, 4, 2).sum(2)
    assert rel(cs, want) < 1e-5
    gam, bet = torch.randn(cout, device=dev), torch.randn(cout, device=dev)
    st_a, ya, co_a = o.gn_fwd(y, None, 32, 1e-5, gam, bet, True, want_coef=True)
    st_b, yb, co_b = o.gn_fwd_from_csum(y, None, cs, None, 32, 1e-5, gam, bet, True, want_coef=True)
    assert rel(st_b, st_a) < 1e-5 and rel(co_b, co_a) < 1e-4
    assert rel(yb, ya) < 4e-3                                          # bf16 outputs, fp32 statistics from two orders
    # concat of two tensors whose statistics came from two different producers
    cs2 = torch.zeros(n, cout // 4, 2, device=dev)
    y2 = o.conv_gemm(x0, x1, taps, w, cout, (n, hw, hw), csum=cs2)
    g2, b2 = torch.randn(2 * cout, device=dev), torch.randn(2 * cout, device=dev)
    st_c, yc = o.gn_fwd(y, y2, 32, 1e-6, g2, b2, False)
    st_d, yd = o.gn_fwd_from_csum(y, y2, cs, cs2, 32, 1e-6, g2, b2, False)
    assert rel(st_d, st_c) < 1e-5 and rel(yd, yc) < 4e-3


def test_fused_adamw_vs_torch_adamw(dev):
    """clip_grad_norm_(1.0) + AdamW.step() (train_from_scratch.py:106-108) as two streaming kernels over the flat
    arena, against torch's own clip + AdamW on identical gradients: fp32, <= 1e-6 relative to the update size."""
    from polyp_image_generator_b200 import FusedAdamW, UNet2DModel
    from polyp_image_generator_b200.training import mse_loss
    cfg = _small_cfg(32)
    torch.manual_seed(3)
    m = UNet2DModel(**cfg).to(dev).train()
    x, t = torch.randn(4, 3, 32, 32, device=dev), torch.randint(0, 1000, (4,), device=dev)
    tgt = torch.randn(4, 3, 32, 32, device=dev)
    opt = FusedAdamW(m.parameters(), lr=2e-4, weight_decay=0.01, max_grad_norm=1.0)
    shadow = [p.detach().clone().requires_grad_(True) for p in m.parameters()]
    ref = torch.optim.AdamW(shadow, lr=2e-4, weight_decay=0.01)
    for it in range(4):
        (mse_loss(m(x, t, return_dict=False)[0], tgt) * (30.0 if it == 0 else 1.0)).backward()
        for q, p in zip(shadow, m.parameters()):
            q.grad = p.grad.detach().clone()
        gn = torch.nn.utils.clip_grad_norm_(shadow, 1.0)
        if it == 0:
            assert gn.item() > 1.0                      # step 0 exercises the clip branch
        ref.step()
        opt.step()
        opt.zero_grad()
        worst = max((p.detach() - q.detach()).abs().max().item() for q, p in zip(shadow, m.parameters()))
        assert worst <= 2e-4 * 5e-3, (it, worst)        # 0.5 % of one lr-sized update (fp32 op-order noise)
        with torch.no_grad():                           # keep both on the same trajectory
            for q, p in zip(shadow, m.parameters()):
                q.copy_(p)
    # the sum-of-squares kernel against torch
    from polyp_image_generator_b200 import ops as ops_mod
    v = torch.randn(1_000_003, device=dev)
    out = torch.zeros(1, device=dev)
    ops_mod.get().sumsq(v, out)
    assert abs(out.item() - (v.double() ** 2).sum().item()) / out.item() < 1e-5


def test_pipeline_sampling_vs_oracle(dev):
    """DDPMPipeline: same CPU generator -> same images as the oracle pipeline (8 strided steps, small UNet)."""
    from polyp_image_generator_b200 import DDPMPipeline, DDPMScheduler, UNet2DModel
    cfg = _small_cfg(32)
    torch.manual_seed(5)
    om = oracle.UNet2DModel(**cfg)
    m = UNet2DModel(**cfg)
    m.load_state_dict(om.state_dict())
    m.to(dev).eval()
    pa = DDPMPipeline(unet=m, scheduler=DDPMScheduler())
    pb = oracle.DDPMPipeline(unet=om, scheduler=oracle.DDPMScheduler())
    ia = pa(batch_size=2, generator=torch.Generator("cpu").manual_seed(11), num_inference_steps=8, output_type="np").images
    ib = pb(batch_size=2, generator=torch.Generator("cpu").manual_seed(11), num_inference_steps=8, output_type="np").images
    assert ia.shape == ib.shape == (2, 32, 32, 3)
    assert abs(ia - ib).mean() < 2e-2      # bf16 UNet inside an 8-step chain, images in [0, 1]
    pil = pa(batch_size=1, generator=torch.Generator("cpu").manual_seed(11), num_inference_steps=2).images
    assert pil[0].size == (32, 32)
    dev_noise = pa(batch_size=2, num_inference_steps=3, output_type="uint8").images   # in-kernel Philox noise path
    assert dev_noise.dtype == torch.uint8 and dev_noise.shape == (2, 32, 32, 3)


def test_lora_forward_backward_and_merge_vs_oracle(dev):
    """configs[3]-style LoRA (r=8, alpha=8, to_q/to_k/to_v/to_out.0), dropout off for parity."""
    from polyp_image_generator_b200 import LoraConfig, UNet2DModel
    from polyp_image_generator_b200.lora import lora_state_dict, merge_adapter
    from polyp_image_generator_b200.training import mse_loss
    cfg = _small_cfg(64)
    torch.manual_seed(7)
    om = oracle.UNet2DModel(**cfg)
    m = UNet2DModel(**cfg)
    m.load_state_dict(om.state_dict())
    tg = ["to_q", "to_k", "to_v", "to_out.0"]
    oracle.add_adapter(om, oracle.LoraConfig(r=8, lora_alpha=8, target_modules=tg, init_lora_weights="gaussian"))
    m.add_adapter(LoraConfig(r=8, lora_alpha=8, target_modules=tg, init_lora_weights="gaussian"))
    sd = {k: torch.randn_like(v) * 0.05 for k, v in oracle.lora_state_dict(om).items()}
    om.load_state_dict(sd, strict=False)
    m.load_state_dict(sd, strict=False)
    m.to(dev).train()
    x, t, nz = torch.randn(3, 3, 64, 64), torch.tensor([4, 400, 900]), torch.randn(3, 3, 64, 64)
    pred = m(x.to(dev), t.to(dev)).sample
    pred_o = om(x, t).sample
    assert rel(pred, pred_o) < 2e-2
    mse_loss(pred, nz.to(dev)).backward()
    for n, p in m.named_parameters():
        assert p.requires_grad or p.grad is None, n
    # Adapter gradients.  The adapters sit in the 1..16-token attentions at the bottom of the UNet, behind ~60 bf16
    # storage points in each direction: ANY bf16 realisation of the graph (the rounding-matched oracle) differs from
    # the fp32 oracle by several per cent per adapter tensor at random init.  End to end the product must stay within
    # 2x of that floor, tensor by tensor; the LoRA path itself is checked tightly below (self-consistency) and
    # layer-locally at 256x256 in tests/test_parity_full_gpu.py::test_celebahq_256_lora_step_vs_oracle.
    from parity_util import _grad_table, _noise_floor_ratio, _oracle_grads
    _, g_o = _oracle_grads(om, x, t, nz, rounded=False)
    _, g_r = _oracle_grads(om, x, t, nz, rounded=True)
    whole, rows = _grad_table(m, g_o)
    ratio, worst_name, cnt = _noise_floor_ratio(rows, g_r, g_o, floor=1e-6)
    floor_whole = (sum((g_r[n] - g_o[n]).norm().item() ** 2 for n in g_o) /
                   sum(g_o[n].norm().item() ** 2 for n in g_o)) ** 0.5
    print(f"LoRA gradients: whole vs fp32 {whole:.4f}, bf16-storage floor {floor_whole:.4f}, worst tensor ratio "
          f"{ratio:.2f} ({worst_name}), {cnt} tensors")
    assert len(rows) == 48
    assert whole < 2.0 * max(floor_whole, 1e-2), (whole, floor_whole)
    assert ratio < 3.0, (worst_name, ratio)

    # Self-consistency (common-mode noise cancels): with the base projections ALSO trainable, the same backward
    # yields dW = dY^T x, and the adapter gradients must equal dA = s B^T dW, dB = s dW A^T  (s = alpha / r = 1).
    for n, p in m.named_parameters():
        p.grad = None
        if n.endswith("base_layer.weight"):
            p.requires_grad_(True)
    pred2 = m(x.to(dev), t.to(dev)).sample
    mse_loss(pred2, nz.to(dev)).backward()
    named = dict(m.named_parameters())
    checked = 0
    for n, p in named.items():
        if ".lora_A." not in n:
            continue
        base = n.split(".lora_A.")[0]
        dW = named[base + ".base_layer.weight"].grad.float()
        A, Bm = named[base + ".lora_A.default.weight"], named[base + ".lora_B.default.weight"]
        if dW.norm() < 1e-9:          # q/k of a 1-token attention: softmax over one key has zero gradient
            continue
        assert rel(A.grad, Bm.detach().t().float() @ dW) < 3e-2, n
        assert rel(Bm.grad, dW @ A.detach().t().float()) < 3e-2, n
        checked += 1
    assert checked >= 12
    for n, p in m.named_parameters():
        if n.endswith("base_layer.weight"):
            p.requires_grad_(False)
            p.grad = None
    assert len(lora_state_dict(m)) == 48
    merge_adapter(m)
    oracle.merge_adapter(om)
    for (n, p), (_, po) in zip(m.named_parameters(), om.named_parameters()):
        assert torch.allclose(p.detach().cpu(), po, rtol=0, atol=1e-5), n      # merged weights within 1e-5
    m.eval()
    with torch.no_grad():
        assert rel(m(x.to(dev), t.to(dev)).sample, pred) < 1e-2


def test_lora_dropout_kernel(dev):
    """Keep-rate, scaling, determinism in (seed, offset), and forward/backward mask agreement."""
    from polyp_image_generator_b200 import ops
    o = ops.get()
    x = torch.ones(1 << 20, device=dev, dtype=torch.bfloat16)
    y = o.dropout(x, 0.3, 11, 5)
    y2 = o.dropout(x, 0.3, 11, 5)
    y3 = o.dropout(x, 0.3, 11, 6)
    assert torch.equal(y, y2) and not torch.equal(y, y3)
    kept = (y != 0).float().mean().item()
    assert abs(kept - 0.7) < 5e-3
    assert torch.allclose(y[y != 0].float(), torch.tensor(1 / 0.7), rtol=1e-2)
    g = torch.randn(1 << 20, device=dev).to(torch.bfloat16)
    add = torch.randn(1 << 20, device=dev).to(torch.bfloat16)
    back = o.dropout(g, 0.3, 11, 5, add=add)
    want = add.float() + g.float() * (y != 0).float() / 0.7
    assert rel(back, want) < 5e-3
    assert torch.equal(o.dropout(x, 0.0, 1, 1), x)
    # device-resident step counter (CUDA-graph replays): tick 0 == no tick, tick 1 draws another mask
    tick = torch.zeros(1, device=dev, dtype=torch.int64)
    assert torch.equal(o.dropout(x, 0.3, 11, 5, tick=tick), y)
    tick.add_(1)
    y4 = o.dropout(x, 0.3, 11, 5, tick=tick)
    assert not torch.equal(y4, y) and abs((y4 == 0).float().mean().item() - 0.3) < 5e-3


def test_full_size_properties_128(dev):
    """BASELINE configs[1] size (128x128, full model): properties that need no CPU oracle run -- finite output,
    batch independence (sample i's eps does not depend on its batch mates), run-to-run stability."""
    from polyp_image_generator_b200 import UNet2DModel
    torch.manual_seed(6)
    m = UNet2DModel(**oracle.polyp_unet_config(128)).to(dev).eval()
    x = torch.randn(8, 3, 128, 128, device=dev)
    t = torch.randint(0, 1000, (8,), device=dev)
    with torch.no_grad():
        y = m(x, t).sample
        y2 = m(x, t).sample
        y_sub = m(x[2:5], t[2:5]).sample
    assert torch.isfinite(y).all() and y.shape == x.shape
    assert rel(y, y2) < 1.5e-2        # see test_unet_forward_backward_vs_oracle: bf16-noise level, not bit-exact
    assert rel(y_sub, y[2:5]) < 1.5e-2


def test_bias_gradient_fusion_on_off_agree_on_the_full_model(dev, monkeypatch):
    """The bias gradient of a conv = pixel sums of the gradient entering it.  With DDPM_BIAS_FUSION=1 (default) the
    GroupNorm-backward pass that PRODUCES that gradient accumulates the sums (gn_bwd_apply out_c); with 0 a separate
    column reduction reads the stored tensor again.  Full polyp model at 64x64: every bias gradient and the whole
    gradient agree to the run-to-run noise of the bf16 path."""
    from polyp_image_generator_b200 import UNet2DModel
    from polyp_image_generator_b200.training import mse_loss
    torch.manual_seed(21)
    m = UNet2DModel(**oracle.polyp_unet_config(64)).to(dev).train()
    x = torch.randn(4, 3, 64, 64, device=dev)
    noise = torch.randn_like(x)
    t = torch.randint(0, 1000, (4,), device=dev)

    def grads(flag):
        monkeypatch.setenv("DDPM_BIAS_FUSION", flag)
        m.zero_grad(set_to_none=True)
        mse_loss(m(x, t, return_dict=False)[0], noise).backward()
        return {n: p.grad.detach().float().clone() for n, p in m.named_parameters()}

    g1, g0, g1b = grads("1"), grads("0"), grads("1")
    flat = lambda g: torch.cat([v.reshape(-1) for v in g.values()])
    noise_floor = rel(flat(g1b), flat(g1))
    assert rel(flat(g0), flat(g1)) < max(3 * noise_floor, 2e-3), (rel(flat(g0), flat(g1)), noise_floor)
    worst = 0.0
    for n in g1:
        if n.endswith(".bias") and ("conv" in n) and g1[n].norm() > 1e-3 * flat(g1).norm():
            worst = max(worst, rel(g0[n], g1[n]))
    assert worst < 2e-2, worst


def test_lora_on_time_emb_proj_vs_oracle_and_self_consistency(dev):
    """LoRA on the per-block time-embedding projections (config_diffusion.py:37 candidate target; fp32 side computation
    on the tiled SIMT linears): forward against the oracle, and -- with the base projections also trainable -- the
    identities dB = s dW A^T, dA = s B^T dW between the adapter gradients and the base weight gradient of the SAME
    backward pass (common-mode bf16 noise of the UNet cancels; the LoRA arithmetic itself is fp32)."""
    from polyp_image_generator_b200 import LoraConfig, UNet2DModel
    from polyp_image_generator_b200.training import mse_loss
    cfg = _small_cfg(64)
    torch.manual_seed(9)
    om = oracle.UNet2DModel(**cfg)
    m = UNet2DModel(**cfg)
    m.load_state_dict(om.state_dict())
    tg = ["to_q", "to_v", "time_emb_proj"]
    oracle.add_adapter(om, oracle.LoraConfig(r=8, lora_alpha=16, target_modules=tg, init_lora_weights="gaussian"))
    m.add_adapter(LoraConfig(r=8, lora_alpha=16, target_modules=tg, init_lora_weights="gaussian"))
    sd = {k: torch.randn_like(v) * 0.05 for k, v in oracle.lora_state_dict(om).items()}
    om.load_state_dict(sd, strict=False)
    m.load_state_dict(sd, strict=False)
    assert list(m.state_dict().keys()) == list(om.state_dict().keys())
    m.to(dev).train()
    x, t, nz = torch.randn(3, 3, 64, 64), torch.tensor([4, 400, 900]), torch.randn(3, 3, 64, 64)
    pred = m(x.to(dev), t.to(dev)).sample
    assert rel(pred, om(x, t).sample) < 2e-2
    for n, p in m.named_parameters():
        if n.endswith("time_emb_proj.base_layer.weight"):
            p.requires_grad_(True)
    pred = m(x.to(dev), t.to(dev)).sample
    mse_loss(pred, nz.to(dev)).backward()
    named = dict(m.named_parameters())
    checked = 0
    for n, p in named.items():
        if ".time_emb_proj.lora_A." not in n:
            continue
        base = n.split(".lora_A.")[0]
        dW = named[base + ".base_layer.weight"].grad.float()
        A, Bm = named[base + ".lora_A.default.weight"], named[base + ".lora_B.default.weight"]
        s = 16 / 8
        assert rel(Bm.grad, s * dW @ A.detach().t()) < 1e-4, base
        assert rel(A.grad, s * Bm.detach().t() @ dW) < 1e-4, base
        checked += 1
    assert checked >= 20
    m.eval()
    m.inference_precision = "fp32"
    om.eval()
    with torch.no_grad():
        assert rel(m(x.to(dev), t.to(dev)).sample, om(x, t).sample) < 1e-4
