"""Golden vectors for the rows added in session 3 (DDIM sampler, Pillow-exact input transform).

    python tests/golden/make_golden_session3.py

Writes tests/golden/ddim.json (a 25-step DDIM trajectory of the CPU oracle with a deterministic stand-in for the UNet,
eta = 0.3, CPU generator consumed in diffusers' order) and tests/golden/resize.json (Pillow ITSELF -- a third-party
dependency of the reference that is installed here -- resizing a fixed 37x29 RGB pattern to 16x16, plus the
torchvision Resize/flip/ToTensor/Normalize output checksum).  The kept fixtures let the GPU box check against what was
generated here even if its Pillow build differed.
"""
import json
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
import oracle  # noqa: E402


def ddim_case():
    s = oracle.DDIMScheduler()
    s.set_timesteps(25)
    g = torch.Generator().manual_seed(11)
    x = torch.randn(1, 3, 8, 8, generator=g)
    sums = []
    for t in s.timesteps:
        eps = torch.cos(x * 2.0 - float(t) * 0.02)     # deterministic stand-in for the UNet
        x = s.step(eps, t, x, eta=0.3, use_clipped_model_output=True, generator=g).prev_sample
        sums.append(x.double().sum().item())
    return {"timesteps": s.timesteps.tolist(), "eta": 0.3, "trajectory_sums": sums, "final": x.flatten().tolist()}


def pattern(h=29, w=37):
    yy, xx = np.mgrid[0:h, 0:w]
    return np.stack([(xx * 7 + yy * 3) % 256, (xx * xx + yy * 5) % 256, (255 - (xx * 2 + yy * yy) % 256)], -1).astype(np.uint8)


def resize_case():
    from PIL import Image
    import torchvision.transforms as T
    img = pattern()
    out = np.asarray(Image.fromarray(img).resize((16, 16), Image.BILINEAR))
    tv = T.Compose([T.Resize((16, 16)), T.RandomHorizontalFlip(p=1.0), T.ToTensor(), T.Normalize([0.5], [0.5])])
    t = tv(Image.fromarray(img))
    return {"h": 29, "w": 37, "size": 16, "resized": out.tolist(), "transform_flipped_sum": t.double().sum().item(),
            "transform_flipped_first_row": t[0, 0].tolist()}


if __name__ == "__main__":
    with open(os.path.join(HERE, "ddim.json"), "w") as f:
        json.dump(ddim_case(), f)
    with open(os.path.join(HERE, "resize.json"), "w") as f:
        json.dump(resize_case(), f)
    print("wrote session-3 golden fixtures")
