"""Generates the UniPC golden fixture (run once, in the build container):

    python tests/golden/make_golden_unipc.py

Writes tests/golden/unipc.json: a 12-step UniPC (bh2, order 2, SD beta table) trajectory of the CPU oracle with a
deterministic stand-in for the UNet.  The oracle itself is anchored by the closed-form test in tests/test_oracle_kat.py
(UniPC reproduces x0 exactly when the data prediction is constant along the trajectory).
"""
import json
import os
import sys

import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
import oracle  # noqa: E402

KW = dict(beta_schedule="scaled_linear", beta_start=0.00085, beta_end=0.012, steps_offset=1)


def unipc_case(n=12):
    s = oracle.UniPCMultistepScheduler(**KW)
    s.set_timesteps(n)
    x = torch.randn(1, 3, 8, 8, generator=torch.Generator().manual_seed(13))
    sums = []
    for t in s.timesteps:
        eps = torch.cos(x * 2.0 - float(t) * 0.02)
        x = s.step(eps, t, x).prev_sample
        sums.append(x.double().sum().item())
    return {"kwargs": KW, "steps": n, "timesteps": s.timesteps.tolist(), "sigmas": s.sigmas.tolist(),
            "trajectory_sums": sums, "final": x.flatten().tolist()}


if __name__ == "__main__":
    with open(os.path.join(HERE, "unipc.json"), "w") as f:
        json.dump(unipc_case(), f)
    print("wrote tests/golden/unipc.json")
