"""One eager (no CUDA graph) pass of a hot-path workload inside a cudaProfilerStart/Stop range, for ncu launch lists:

    ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file X.csv \
        python profiles/step_eager.py {sampling|lora|train} [batch] [size]

sampling: one reverse-diffusion step (UNet forward in eval mode + scheduler step with in-kernel noise), batch 32, 128^2
lora    : one LoRA fine-tune step of the celebahq-architecture UNet, batch 8, 256^2 (train_with_lora_all_classes.py:120-180)
train   : one from-scratch training step, batch 64, 128^2 (train_from_scratch.py:83-116)
Without ncu it prints the eager wall time of the pass (host-issue bound for the small kernels)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from polyp_image_generator_b200 import DDPMScheduler, LoraConfig, UNet2DModel  # noqa: E402
from polyp_image_generator_b200.model import celebahq_unet_config, polyp_unet_config  # noqa: E402
from polyp_image_generator_b200.training import mse_loss  # noqa: E402

mode = sys.argv[1] if len(sys.argv) > 1 else "sampling"
dev = torch.device("cuda:0")
torch.manual_seed(0)


def ranged(fn, warm=3):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.profiler.start()
    e0.record()
    fn()
    e1.record()
    torch.cuda.synchronize()
    torch.cuda.profiler.stop()
    print(f"{mode}: eager pass {e0.elapsed_time(e1):.3f} ms")


if mode == "sampling":
    B = int(sys.argv[2]) if len(sys.argv) > 2 else 32
    S = int(sys.argv[3]) if len(sys.argv) > 3 else 128
    m = UNet2DModel(**polyp_unet_config(S)).to(dev).eval()
    sch = DDPMScheduler(num_train_timesteps=1000)
    sch.set_timesteps(1000)
    x = torch.randn(B, 3, S, S, device=dev)
    tt = torch.full((B,), 500, device=dev, dtype=torch.int64)

    def fn():
        with torch.no_grad():
            eps = m(x, tt, return_dict=False)[0]
            sch.step(eps, 500, x)
    ranged(fn)
else:
    lora = mode == "lora"
    B = int(sys.argv[2]) if len(sys.argv) > 2 else (8 if lora else 64)
    S = int(sys.argv[3]) if len(sys.argv) > 3 else (256 if lora else 128)
    m = UNet2DModel(**(celebahq_unet_config(S) if lora else polyp_unet_config(S))).to(dev)
    if lora:
        m.add_adapter(LoraConfig(r=8, lora_alpha=8, target_modules=["to_q", "to_k", "to_v", "to_out.0"],
                                 lora_dropout=0.3, init_lora_weights="gaussian"))
        m.to(dev)
    m.train()
    params = [p for p in m.parameters() if p.requires_grad]
    if lora:
        opt = torch.optim.AdamW(params, lr=1e-4, fused=True)
    else:
        from polyp_image_generator_b200.optim import FusedAdamW
        opt = FusedAdamW(m.parameters(), lr=1e-4, max_grad_norm=1.0)
    sch = DDPMScheduler(num_train_timesteps=1000)
    clean = torch.randn(B, 3, S, S, device=dev).clamp(-1, 1)
    noise = torch.randn(B, 3, S, S, device=dev)
    t = torch.randint(0, 1000, (B,), device=dev, dtype=torch.int64)

    def fn():
        noisy = sch.add_noise(clean, noise, t)
        pred = m(noisy, t, return_dict=False)[0]
        loss = mse_loss(pred, noise)
        opt.zero_grad(set_to_none=True)
        loss.backward()
        if lora:
            torch.nn.utils.clip_grad_norm_(params, 1.0)
        opt.step()
    ranged(fn)
