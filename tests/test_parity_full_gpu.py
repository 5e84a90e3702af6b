"""Full-model GPU parity at the geometries the benchmark and BASELINE.json configs[1..4] actually run (-m gpu).

  * polyp UNet2DModel (113.7 M parameters) at 128x128 (bench.py / configs[1], [2]) and at 224x224
    (/root/reference/generator_model/config_diffusion.py:6 image_size = 224: 7x7 / 14x14 maps, 196 / 49 tokens),
  * celebahq-architecture UNet at 256x256 (configs[3], [4]) incl. one LoRA step,
  * a >= 200-step reverse-diffusion chain on the full model against the oracle pipeline.

Two comparisons per case:
  (1) against the fp32 oracle (the reference arithmetic): eps-prediction and the whole gradient within north_star's
      2e-2 -- this measures the bf16 precision choice end to end;
  (2) against the ROUNDING-MATCHED oracle (oracle/rounding.py: same fp32 graph, bf16 at the product's storage points):
      every parameter tensor's gradient within 2e-2 of its own norm, for every tensor that carries at least
      `GRAD_FLOOR` of the whole-gradient norm (below that a tensor is numerically zero next to its neighbours: the
      q/k projections of a 1-token attention, for example, have an exactly-zero true gradient).
The measured numbers of every case are appended to gpurun_out/parity_report.jsonl (DESIGN.md §2 quotes them).
"""
import json
import os

import pytest
import torch
import torch.nn.functional as F

import oracle
from oracle.rounding import bf16_storage_points

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GRAD_FLOOR = 1e-3        # per-tensor bound applies to tensors with ||g_n|| >= GRAD_FLOOR * ||g||


def rel(a, b):
    a, b = a.detach().float().cpu(), b.detach().float().cpu()
    return ((a - b).norm() / (b.norm() + 1e-30)).item()


def _report(rec):
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    with open(os.path.join(ROOT, "gpurun_out", "parity_report.jsonl"), "a") as f:
        f.write(json.dumps(rec) + "\n")
    print("PARITY", json.dumps(rec), flush=True)


@pytest.fixture(scope="module")
def dev():
    assert torch.cuda.is_available(), "GPU tests need a CUDA device"
    from polyp_image_generator_b200 import ops
    assert ops.get().name == "cuda"
    return torch.device("cuda:0")


def _polyp_like_batch(B, S, seed):
    """Inputs with the dynamic range of the training data: images in [-1, 1] noised at random timesteps."""
    g = torch.Generator().manual_seed(seed)
    x0 = (torch.rand(B, 3, S, S, generator=g) * 2 - 1) * torch.linspace(0.3, 1.0, B).view(B, 1, 1, 1)
    noise = torch.randn(B, 3, S, S, generator=g)
    t = torch.randint(0, 1000, (B,), generator=g)
    noisy = oracle.DDPMScheduler().add_noise(x0, noise, t)
    return noisy, t, noise


def _oracle_grads(om, x, t, noise, rounded):
    for p in om.parameters():
        p.grad = None
    if rounded:
        with bf16_storage_points():
            pred = om(x, t).sample
            F.mse_loss(pred, noise).backward()
    else:
        pred = om(x, t).sample
        F.mse_loss(pred, noise).backward()
    return pred.detach(), {n: p.grad.detach().clone() for n, p in om.named_parameters() if p.grad is not None}


def _grad_table(m, ref):
    """-> (whole-gradient rel error, [(name, rel to own norm, share of whole norm)] sorted by rel error)."""
    tot = sum(g.norm().item() ** 2 for g in ref.values()) ** 0.5
    num = 0.0
    rows = []
    for n, p in m.named_parameters():
        if not p.requires_grad:
            continue
        g = ref[n]
        d = (p.grad.detach().float().cpu() - g).norm().item()
        num += d * d
        rows.append((n, d / max(g.norm().item(), 1e-30), g.norm().item() / tot))
    rows.sort(key=lambda r: -r[1])
    return num ** 0.5 / tot, rows


CASES = [("polyp", 128, 4), ("polyp", 224, 2), ("celebahq", 256, 2), ("polyp", 64, 4)]


@pytest.mark.parametrize("arch,S,B", CASES)
def test_full_model_forward_backward_at_benchmark_geometry(dev, arch, S, B):
    from polyp_image_generator_b200 import UNet2DModel
    from polyp_image_generator_b200.training import mse_loss
    cfg = oracle.polyp_unet_config(S) if arch == "polyp" else oracle.celebahq_unet_config(S)
    torch.manual_seed(0)
    om = oracle.UNet2DModel(**cfg)
    m = UNet2DModel(**cfg)
    m.load_state_dict(om.state_dict())
    m.to(dev).train()
    x, t, noise = _polyp_like_batch(B, S, 100 + S)
    pred = m(x.to(dev), t.to(dev), return_dict=False)[0]
    loss = mse_loss(pred, noise.to(dev))
    loss.backward()
    pred_o, g_o = _oracle_grads(om, x, t, noise, rounded=False)
    pred_r, g_r = _oracle_grads(om, x, t, noise, rounded=True)
    whole_o, rows_o = _grad_table(m, g_o)
    whole_r, rows_r = _grad_table(m, g_r)
    big_r = [r for r in rows_r if r[2] >= GRAD_FLOOR]
    big_o = [r for r in rows_o if r[2] >= GRAD_FLOOR]
    rec = {"case": f"{arch} {S}x{S} B={B}", "eps_vs_fp32": rel(pred, pred_o), "eps_vs_rounded": rel(pred, pred_r),
           "eps_rounded_vs_fp32": rel(pred_r, pred_o), "loss": loss.item(),
           "loss_fp32": F.mse_loss(pred_o, noise).item(),
           "grad_whole_vs_fp32": whole_o, "grad_whole_vs_rounded": whole_r,
           "grad_worst_tensor_vs_rounded": big_r[0][:2] if big_r else None,
           "grad_worst_tensor_vs_fp32": big_o[0][:2] if big_o else None,
           "grad_tensors_checked": len(big_r), "grad_tensors_total": len(rows_r),
           "grad_p99_vs_rounded": sorted(r[1] for r in big_r)[int(0.99 * (len(big_r) - 1))] if big_r else None,
           "by_floor_vs_rounded": {f: [len([r for r in rows_r if r[2] >= f]),
                                       max([r[1] for r in rows_r if r[2] >= f], default=0.0)]
                                   for f in (1e-2, 1e-3, 1e-4, 1e-5, 1e-6, 0.0)},
           "by_floor_vs_fp32": {f: [len([r for r in rows_o if r[2] >= f]),
                                    max([r[1] for r in rows_o if r[2] >= f], default=0.0)]
                                for f in (1e-2, 1e-3, 1e-4, 1e-5, 1e-6, 0.0)},
           "worst10_vs_rounded": [(n, round(a, 4), float("%.2e" % b)) for n, a, b in rows_r[:10]]}
    _report(rec)
    assert pred.dtype == torch.float32 and pred.shape == pred_o.shape
    assert rec["eps_vs_fp32"] < 2e-2, rec
    assert rec["grad_whole_vs_fp32"] < 2e-2, rec
    assert rec["eps_vs_rounded"] < 2e-2 and rec["grad_whole_vs_rounded"] < 2e-2, rec
    assert not big_r or big_r[0][1] < 2e-2, big_r[:8]


def test_celebahq_256_lora_step_vs_oracle(dev):
    """configs[3]: celebahq-architecture UNet at 256x256, r=8 / alpha=8 adapters on to_q/to_k/to_v/to_out.0 (dropout off
    for parity): forward, adapter gradients and one AdamW step against the oracle."""
    from polyp_image_generator_b200 import LoraConfig, UNet2DModel
    from polyp_image_generator_b200.training import mse_loss
    S, B = 256, 2
    cfg = oracle.celebahq_unet_config(S)
    torch.manual_seed(0)
    om = oracle.UNet2DModel(**cfg)
    m = UNet2DModel(**cfg)
    m.load_state_dict(om.state_dict())
    tg = ["to_q", "to_k", "to_v", "to_out.0"]
    oracle.add_adapter(om, oracle.LoraConfig(r=8, lora_alpha=8, target_modules=tg, init_lora_weights="gaussian"))
    m.add_adapter(LoraConfig(r=8, lora_alpha=8, target_modules=tg, init_lora_weights="gaussian"))
    g = torch.Generator().manual_seed(5)
    sd = {k: torch.randn(v.shape, generator=g) * 0.05 for k, v in oracle.lora_state_dict(om).items()}
    om.load_state_dict(sd, strict=False)
    m.load_state_dict(sd, strict=False)
    m.to(dev).train()
    x, t, noise = _polyp_like_batch(B, S, 77)
    pred = m(x.to(dev), t.to(dev)).sample
    mse_loss(pred, noise.to(dev)).backward()
    pred_o, g_o = _oracle_grads(om, x, t, noise, rounded=False)
    pred_r, g_r = _oracle_grads(om, x, t, noise, rounded=True)
    assert all(p.grad is None for n, p in m.named_parameters() if not p.requires_grad)
    whole_o, rows_o = _grad_table(m, g_o)
    whole_r, rows_r = _grad_table(m, g_r)
    big_r = [r for r in rows_r if r[2] >= GRAD_FLOOR]
    rec = {"case": "celebahq 256x256 B=2 LoRA r=8", "eps_vs_fp32": rel(pred, pred_o), "eps_vs_rounded": rel(pred, pred_r),
           "lora_grad_whole_vs_fp32": whole_o, "lora_grad_whole_vs_rounded": whole_r,
           "lora_grad_worst_tensor_vs_rounded": big_r[0][:2] if big_r else None,
           "lora_grad_worst_tensor_vs_fp32": rows_o[0][:2], "tensors_checked": len(big_r), "tensors": len(rows_r)}
    _report(rec)
    assert len(rows_r) == 48
    assert rec["eps_vs_fp32"] < 2e-2 and rec["eps_vs_rounded"] < 2e-2, rec
    assert whole_r < 2e-2, rec
    assert not big_r or big_r[0][1] < 2e-2, big_r[:8]
    assert whole_o < 2e-2, rec


def test_sampling_chain_250_steps_full_model(dev):
    """DDPMPipeline on the full 113.7 M-parameter UNet, 250 strided reverse steps from the same CPU generator as the
    oracle pipeline (the reference's RNG contract): image-level agreement after the whole chain."""
    from polyp_image_generator_b200 import DDPMPipeline, DDPMScheduler, UNet2DModel
    S, B, steps = 64, 2, int(os.environ.get("DDPM_TEST_CHAIN_STEPS", "250"))
    cfg = oracle.polyp_unet_config(S)
    torch.manual_seed(5)
    om = oracle.UNet2DModel(**cfg)
    m = UNet2DModel(**cfg)
    m.load_state_dict(om.state_dict())
    m.to(dev).eval()
    om.eval()
    pa = DDPMPipeline(unet=m, scheduler=DDPMScheduler())
    pb = oracle.DDPMPipeline(unet=om, scheduler=oracle.DDPMScheduler())
    raw = pa(batch_size=B, generator=torch.Generator("cpu").manual_seed(11), num_inference_steps=steps,
             output_type="pt_raw").images           # x_0 in [-1, 1] on the device, before the uint8 epilogue
    ia = (raw / 2 + 0.5).clamp(0, 1).cpu().permute(0, 2, 3, 1).numpy()
    ib = pb(batch_size=B, generator=torch.Generator("cpu").manual_seed(11), num_inference_steps=steps,
            output_type="np").images
    with bf16_storage_points():
        ic = pb(batch_size=B, generator=torch.Generator("cpu").manual_seed(11), num_inference_steps=steps,
                output_type="np").images
    u8 = lambda im: (im * 255).round().astype("int32")
    rec = {"case": f"sampling chain polyp {S}x{S} B={B} {steps} steps",
           "mean_abs_vs_fp32": float(abs(ia - ib).mean()), "max_abs_vs_fp32": float(abs(ia - ib).max()),
           "mean_abs_vs_rounded": float(abs(ia - ic).mean()),
           "mean_abs_rounded_vs_fp32": float(abs(ic - ib).mean()),
           "uint8_levels_mean_vs_fp32": float(abs(u8(ia) - u8(ib)).mean()),
           "uint8_frac_within_2_levels": float((abs(u8(ia) - u8(ib)) <= 2).mean())}
    _report(rec)
    assert ia.shape == ib.shape == (B, S, S, 3)
    assert rec["mean_abs_vs_fp32"] < 1e-2, rec        # images in [0, 1]
    assert rec["uint8_frac_within_2_levels"] > 0.9, rec


def test_two_rank_nccl_gradients_equal_single_rank(dev):
    """N-rank NCCL data-parallel gradients == 1-rank gradients on the concatenated batch, through
    the real NCCL + side-stream + CUDA-graph path (tests/nccl_worker.py, 2 ranks).  Needs 2 GPUs."""
    import subprocess
    import sys
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs (gpurun --gpus 2)")
    port = 29500 + os.getpid() % 1000
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr",
           "127.0.0.1", "--master-port", str(port), os.path.join(ROOT, "tests", "nccl_worker.py")]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=600)
    print(r.stdout[-4000:], r.stderr[-4000:])
    assert r.returncode == 0, r.stderr[-2000:]
    assert "NCCL_PARITY_OK" in r.stdout
