"""Generates the committed golden vectors from the CPU oracle (the reference ships none).

    python tests/golden/make_golden.py

Writes tests/golden/unet32_fwd_bwd.pt (a small-channel UNet2DModel case: inputs, eps prediction, loss, per-tensor
gradient norms, a few full gradients) and tests/golden/scheduler.json (closed-form scheduler scalars and a
50-step sampling trajectory checksum).  Seeds are fixed; torch CPU fp32.
"""
import json
import os
import sys

import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
import oracle  # noqa: E402


def small_cfg():
    cfg = oracle.polyp_unet_config(32)
    cfg["block_out_channels"] = (64, 64, 128, 128, 128, 128)
    return cfg


def small_case():
    torch.manual_seed(1234)
    m = oracle.UNet2DModel(**small_cfg())
    x0 = torch.randn(2, 3, 32, 32).clamp(-1, 1)
    noise = torch.randn(2, 3, 32, 32)
    t = torch.tensor([25, 640])
    s = oracle.DDPMScheduler()
    noisy = s.add_noise(x0, noise, t)
    pred = m(noisy, t).sample
    loss = torch.nn.functional.mse_loss(pred, noise)
    loss.backward()
    grads = {n: p.grad.clone() for n, p in m.named_parameters()}
    keep = ["conv_in.weight", "conv_out.weight", "mid_block.attentions.0.to_v.weight",
            "down_blocks.2.resnets.0.conv_shortcut.weight", "up_blocks.5.resnets.2.norm2.weight",
            "time_embedding.linear_1.weight", "up_blocks.0.resnets.0.time_emb_proj.bias"]
    return {"state_dict": {k: v.clone() for k, v in m.state_dict().items()}, "x0": x0, "noise": noise, "t": t,
            "noisy": noisy, "pred": pred.detach(), "loss": loss.item(),
            "grad_norms": {n: g.norm().item() for n, g in grads.items()}, "grads": {k: grads[k] for k in keep}}


def scheduler_case():
    s = oracle.DDPMScheduler()
    s.set_timesteps(50)
    g = torch.Generator().manual_seed(7)
    x = torch.randn(1, 3, 8, 8, generator=g)
    sums = []
    for t in s.timesteps:
        eps = torch.sin(x * 3.0 + float(t) * 0.01)     # deterministic stand-in for the UNet
        x = s.step(eps, t, x, generator=g).prev_sample
        sums.append(x.double().sum().item())
    return {"timesteps": s.timesteps.tolist(), "trajectory_sums": sums, "final": x.flatten().tolist()}


if __name__ == "__main__":
    case = small_case()
    torch.save({k: case[k] for k in ("x0", "noise", "t", "noisy", "pred", "loss", "grad_norms", "grads")},
               os.path.join(HERE, "unet32_fwd_bwd.pt"))
    with open(os.path.join(HERE, "scheduler.json"), "w") as f:
        json.dump(scheduler_case(), f)
    print("wrote golden fixtures")
