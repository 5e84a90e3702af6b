"""Per-op CUDA-event breakdown of ONE reverse-diffusion step's UNet forward (eval, no grad) at a given batch."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from polyp_image_generator_b200.model import polyp_unet_config
from polyp_image_generator_b200 import UNet2DModel
from polyp_image_generator_b200 import ops as ops_mod

B = int(sys.argv[1]) if len(sys.argv) > 1 else 32
S = 128
torch.manual_seed(0)
m = UNet2DModel(**polyp_unet_config(S)).to("cuda").eval()
x = torch.randn(B, 3, S, S, device="cuda")
t = torch.full((B,), 500, device="cuda", dtype=torch.int64)
with torch.no_grad():
    for _ in range(3):
        m(x, t)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(5):
        m(x, t)
    e1.record()
    torch.cuda.synchronize()
    print(f"eager forward batch {B}: {e0.elapsed_time(e1) / 5:.3f} ms")
    pr = ops_mod.OpProfiler(ops_mod.get())
    pr.by_shape = len(sys.argv) > 2
    pr.start()
    for _ in range(3):
        m(x, t)
    tab = pr.stop()
tot = 0.0
for k, v in sorted(tab.items(), key=lambda kv: -kv[1]["ms"]):
    ms = v["ms"] / 3
    tot += ms
    tf = v["flops"] / 3 / (ms * 1e-3) / 1e12 if v["flops"] else 0
    gb = v["bytes"] / 3 / (ms * 1e-3) / 1e9 if v["bytes"] else 0
    print(f"  {k:40s} calls {v['calls'] // 3:4d} {ms:8.3f} ms  {tf:7.1f} TFLOP/s {gb:7.0f} GB/s")
print(f"  sum of ops {tot:.3f} ms")
