"""Hardware probe: shifted (non-1024-aligned) SWIZZLE_128B UMMA operand descriptors (kernel: profiles/scratch/shift_probe.cu -- it was linked into the library during round-1 bring-up as ddpm_debug_shift_probe; it is no longer part of the product ABI, so this script is a record of the experiment, not a runnable tool)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from polyp_image_generator_b200 import _capi

lib = _capi.load()
torch.manual_seed(0)
a = torch.randn(256, 64, device="cuda").to(torch.bfloat16)
b = torch.randn(64, 64, device="cuda").to(torch.bfloat16)
for mode in (0, 1):
    for shift in (0, 1, 2, 3, 5, 7, 8, 9, 16, 17, 65, 128):
        out = torch.full((128, 64), float("nan"), device="cuda")
        rc = lib.ddpm_debug_shift_probe(a.data_ptr(), b.data_ptr(), out.data_ptr(), shift, mode,
                                        torch.cuda.current_stream().cuda_stream)
        torch.cuda.synchronize()
        want = a[shift:shift + 128].float() @ b.float().t()
        err = ((out - want).norm() / want.norm()).item()
        # which row offset does the result actually correspond to?
        best = min(range(0, 129), key=lambda s: ((out - a[s:s + 128].float() @ b.float().t()).norm()).item())
        print(f"base_offset_mode={mode} shift={shift:3d} rc={rc} rel_err={err:.3e} best_matching_shift={best}", flush=True)
