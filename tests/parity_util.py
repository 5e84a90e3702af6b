"""Shared helpers of the GPU parity tests (TEST INFRASTRUCTURE): oracle gradients in fp32 and in the rounding-matched
mode (oracle/rounding.py), per-tensor gradient tables, and the bf16-storage noise-floor ratio."""
import torch
import torch.nn.functional as F

import oracle
from oracle.rounding import bf16_storage_points


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


def _noise_floor_ratio(rows_prod_vs_fp32, g_r, g_o, floor=1e-6):
    """Per tensor: (product vs fp32 oracle) / max(rounding-matched oracle vs fp32 oracle, 5e-3).  The denominator is what
    ANY implementation that stores bf16 at the product's storage points differs from fp32 by, so a ratio of ~1 means
    the tensor carries no error beyond the precision choice.  -> (worst ratio, its name, #tensors)."""
    worst, name, cnt = 0.0, "", 0
    for n, e, share in rows_prod_vs_fp32:
        if share < floor:
            continue
        base = ((g_r[n] - g_o[n]).norm() / (g_o[n].norm() + 1e-30)).item()
        r = e / max(base, 5e-3)
        cnt += 1
        if r > worst:
            worst, name = r, n
    return worst, name, cnt


