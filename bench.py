#!/usr/bin/env python
"""bench.py -- DDPM UNet training throughput (img/s @128^2) on N B200s, with roofline and CPU baseline.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference] [--batch B] [--size S]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N --steps K --warmup W

Workload (BASELINE.json configs[1]): UNet2DModel 128x128 train_from_scratch DDPM step, bf16 tensor-core compute,
batch 64 per GPU, data-parallel over N GPUs.  One "step" = add_noise -> UNet forward -> MSE -> backward ->
(gradient all-reduce) -> clip_grad_norm_(1.0) -> AdamW -> zero_grad, i.e. the loop body of
/root/reference/generator_model/train_from_scratch.py:83-116 over the drop-in objects.

Prints ONE JSON line (rank 0).  `value`: inputs already resident in HBM.  `e2e`: same step through the public API
with the batch coming from pinned host memory every step and the loss read back (loss.item()) every step.
`roofline`: the dominant kernel (tcgen05 implicit-GEMM conv fprop/dgrad) timed live with CUDA events inside the
timed region.  `cpu_baseline` / --impl reference: the oracle (pure-PyTorch restatement of the reference's
diffusers path) on the host cores, bounded sample.
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import torch

METRIC = "ddpm_unet_train_images_per_sec_128px"
UNIT = "img/s"
FWD_GFLOP_PER_IMG = {64: 31.035, 128: 124.138, 256: 497.028}   # SURVEY.md §8(d), algorithmic


def synthetic_polyp_batch(n: int, size: int, seed: int) -> torch.Tensor:
    """Polyp-shaped synthetic RGB in [-1, 1]: mucosa base colour + low-frequency noise, one bright
    shaded ellipse with specular dots, dark endoscope vignette, random horizontal flip; fp32 NCHW."""
    g = torch.Generator().manual_seed(seed)
    yy, xx = torch.meshgrid(torch.linspace(0, 1, size), torch.linspace(0, 1, size), indexing="ij")
    base = torch.tensor([0.75, 0.35, 0.30]).view(1, 3, 1, 1)
    low = torch.nn.functional.interpolate(torch.randn(n, 3, 8, 8, generator=g), size=(size, size), mode="bilinear",
                                          align_corners=False) * 0.06
    img = base + low
    cx, cy = torch.rand(n, generator=g) * 0.4 + 0.3, torch.rand(n, generator=g) * 0.4 + 0.3
    ax, ay = torch.rand(n, generator=g) * 0.2 + 0.1, torch.rand(n, generator=g) * 0.2 + 0.1
    th = torch.rand(n, generator=g) * 3.14159
    dx, dy = xx[None] - cx.view(-1, 1, 1), yy[None] - cy.view(-1, 1, 1)
    u = dx * torch.cos(th).view(-1, 1, 1) + dy * torch.sin(th).view(-1, 1, 1)
    v = -dx * torch.sin(th).view(-1, 1, 1) + dy * torch.cos(th).view(-1, 1, 1)
    r2 = (u / ax.view(-1, 1, 1)) ** 2 + (v / ay.view(-1, 1, 1)) ** 2
    bump = torch.clamp(1 - r2, min=0).sqrt()
    img = img + 0.25 * bump[:, None] * torch.tensor([1.0, 0.8, 0.7]).view(1, 3, 1, 1)
    for _ in range(3):
        sx, sy = cx + (torch.rand(n, generator=g) - 0.5) * ax, cy + (torch.rand(n, generator=g) - 0.5) * ay
        d2 = (xx[None] - sx.view(-1, 1, 1)) ** 2 + (yy[None] - sy.view(-1, 1, 1)) ** 2
        img = img + 0.5 * torch.exp(-d2 / 2e-4)[:, None]
    vign = ((xx - 0.5) ** 2 + (yy - 0.5) ** 2).sqrt()
    img = img * torch.clamp(1.15 - 1.6 * vign, 0, 1)[None, None]
    flip = torch.rand(n, generator=g) < 0.5
    img[flip] = img[flip].flip(-1)
    return ((img.clamp(0, 1) - 0.5) / 0.5).contiguous()


class ClockSampler(threading.Thread):
    """Samples SM clock / throttle reasons through NVML while the timed region runs."""

    def __init__(self, index: int, period_s: float = 0.1):
        super().__init__(daemon=True)
        self.index, self.period = index, period_s
        self.samples, self.reasons, self.max_mhz = [], set(), None
        self._stop_evt = threading.Event()
        self.ok = False
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
            self.ok = True
        except Exception as e:  # noqa: BLE001
            self.err = str(e)

    def run(self):
        if not self.ok:
            return
        nv = self.nv
        names = {
            getattr(nv, "nvmlClocksEventReasonHwSlowdown", 0x8): "hw_slowdown",
            getattr(nv, "nvmlClocksEventReasonSwPowerCap", 0x4): "sw_power_cap",
            getattr(nv, "nvmlClocksEventReasonHwThermalSlowdown", 0x40): "hw_thermal_slowdown",
            getattr(nv, "nvmlClocksEventReasonSwThermalSlowdown", 0x20): "sw_thermal_slowdown",
            getattr(nv, "nvmlClocksEventReasonHwPowerBrakeSlowdown", 0x80): "hw_power_brake_slowdown",
        }
        while not self._stop_evt.is_set():
            try:
                self.samples.append(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM))
                try:
                    mask = nv.nvmlDeviceGetCurrentClocksEventReasons(self.h)
                except Exception:  # noqa: BLE001
                    mask = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                for bit, name in names.items():
                    if mask & bit:
                        self.reasons.add(name)
            except Exception:  # noqa: BLE001
                pass
            self._stop_evt.wait(self.period)

    def stop(self):
        self._stop_evt.set()
        if self.is_alive():
            self.join(timeout=2)
        if not self.samples:
            return {"sm_mhz": None, "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons)}
        return {"sm_mhz": statistics.median(self.samples), "sm_max_mhz": self.max_mhz,
                "reasons": sorted(self.reasons)}


# ---------------------------------------------------------------------------------------------------------------
# CPU arm: the oracle on the host cores
# ---------------------------------------------------------------------------------------------------------------
def cpu_oracle_train(size: int, batch: int, steps: int, warmup: int, budget_s: float = 30.0, precision: str = "bf16"):
    """Reference CPU path = oracle restatement of the diffusers graph driven by train_from_scratch.py:83-116.
    precision "bf16": torch.amp.autocast("cpu") exactly as the reference does at :95 on a CPU device;
    precision "fp32": the same loop without autocast (what the reference's arithmetic is on fp32 weights)."""
    import contextlib
    import oracle
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    torch.manual_seed(0)
    model = oracle.UNet2DModel(**oracle.polyp_unet_config(size))
    sched = oracle.DDPMScheduler(num_train_timesteps=1000)
    opt = torch.optim.AdamW(model.parameters(), lr=1e-4)
    clean = synthetic_polyp_batch(batch, size, 1234)
    g = torch.Generator().manual_seed(4321)
    times = []
    t_begin = time.perf_counter()
    for i in range(warmup + steps):
        noise = torch.randn(clean.shape, generator=g)
        t = torch.randint(0, 1000, (batch,), generator=g, dtype=torch.int64)
        t0 = time.perf_counter()
        noisy = sched.add_noise(clean, noise, t)
        with (torch.amp.autocast("cpu") if precision == "bf16" else contextlib.nullcontext()):
            pred = model(noisy, t, return_dict=False)[0]
            loss = torch.nn.functional.mse_loss(pred, noise)
        loss.backward()
        torch.nn.utils.clip_grad_norm_(model.parameters(), 1.0)
        opt.step()
        opt.zero_grad()
        _ = loss.item()
        dt = time.perf_counter() - t0
        if i >= warmup:
            times.append(dt)
        if time.perf_counter() - t_begin > budget_s and len(times) >= 1:
            break
    ms = 1e3 * sum(times) / len(times)
    return {"value": batch / (ms / 1e3), "ms_per_step": ms, "steps": len(times), "cores": cores, "precision": precision,
            "sample": f"{len(times)} step(s) of batch {batch} at {size}x{size} "
                      f"({'bf16 CPU autocast' if precision == 'bf16' else 'fp32, no autocast'}, AdamW, clip 1.0)"}


def cpu_oracle_both(size: int, batch: int, steps: int, budget_s: float):
    """Both precisions of the CPU arm; the FASTER one is the quoted baseline."""
    runs = [cpu_oracle_train(size, batch, steps, 1, budget_s=budget_s / 2, precision=p) for p in ("bf16", "fp32")]
    best = max(runs, key=lambda r: r["value"])
    detail = {r["precision"]: {"img_per_s": round(r["value"], 4), "ms_per_step": round(r["ms_per_step"], 1),
                               "steps": r["steps"]} for r in runs}
    return best, detail


def run_reference(args, rank):
    if rank != 0:
        return
    r, both = cpu_oracle_both(args.size, args.cpu_batch, max(1, min(args.steps, 3)), budget_s=150.0)
    line = {
        "impl": "reference", "metric": METRIC, "value": round(r["value"], 4), "unit": UNIT, "n_gpus": args.gpus,
        "steps": r["steps"], "warmup": 1, "ms_per_step": round(r["ms_per_step"], 2), "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": r["precision"], "data": "synthetic",
        "config": {"workload": f"UNet2DModel {args.size}x{args.size} DDPM train step (train_from_scratch.py:83-116)",
                   "per_gpu_batch": args.batch, "note": "CPU arm runs a bounded sample: batch "
                   f"{args.cpu_batch} per step on the host cores (oracle port of the reference's diffusers path; "
                   "diffusers itself is not installable offline)"},
        "cpu_baseline": {"value": round(r["value"], 4), "unit": UNIT, "cores": r["cores"], "kind": "port",
                         "sample": r["sample"], "precisions": both},
        "e2e": {"value": round(r["value"], 4), "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


# ---------------------------------------------------------------------------------------------------------------
# B200 arm
# ---------------------------------------------------------------------------------------------------------------

LORA_BWD_FRACTION = 0.741      # dgrad only downstream of down_blocks.4.attentions.0, wgrad only LoRA


def run_lora_finetune(args, dev, world, rank, barrier, batch=None):
    """train_with_lora_all_classes.py:120-180 over the drop-in objects: celebahq-architecture UNet (1 head x 512,
    downsample_padding 0), r=8 / alpha=8 / dropout 0.3 adapters on to_q, to_k, to_v, to_out.0, everything else frozen;
    add_noise -> forward -> MSE -> backward (LoRA gradients only) -> all-reduce -> clip 1.0 -> AdamW, graph-replayed."""
    import torch.distributed as dist
    from polyp_image_generator_b200 import DDPMScheduler, LoraConfig, UNet2DModel
    from polyp_image_generator_b200.graphs import GraphedTrainStep
    from polyp_image_generator_b200.model import celebahq_unet_config
    S, B = args.lora_size, (batch or args.lora_batch)
    torch.manual_seed(1)
    model = UNet2DModel(**celebahq_unet_config(S)).to(dev)
    model.add_adapter(LoraConfig(r=8, lora_alpha=8, target_modules=["to_q", "to_k", "to_v", "to_out.0"],
                                 lora_dropout=0.3, init_lora_weights="gaussian"))
    model.to(dev).train()
    params = [p for p in model.parameters() if p.requires_grad]
    n_train = sum(p.numel() for p in params)
    opt = torch.optim.AdamW(params, lr=1e-4, fused=True, capturable=True)
    net = model
    if world > 1:
        from polyp_image_generator_b200.ddp import DistributedDataParallel
        net = DistributedDataParallel(model)
    sched = DDPMScheduler(num_train_timesteps=1000)
    clean = synthetic_polyp_batch(B, S, 777 + rank).to(dev)
    gen = torch.Generator(device=dev).manual_seed(888 + rank)
    noise = torch.randn(clean.shape, device=dev, generator=gen)
    t = torch.randint(0, 1000, (B,), device=dev, dtype=torch.int64, generator=gen)
    step = GraphedTrainStep(net, sched, opt, clean.shape, max_grad_norm=1.0)
    for _ in range(3):
        step(clean, noise, t)
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    n_steps = max(args.steps, 10)
    e0.record()
    for _ in range(n_steps):
        loss = step(clean, noise, t)
    e1.record()
    barrier()
    ms = torch.tensor([e0.elapsed_time(e1) / n_steps], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    ms = float(ms[0])
    gflop = FWD_GFLOP_PER_IMG.get(S, 0.0) * (1.0 + LORA_BWD_FRACTION) * B
    out = {"workload": f"celebahq-architecture UNet2DModel LoRA fine-tune step, {S}x{S}, batch {B}/GPU, bf16, r=8 "
                       "alpha=8 dropout=0.3 on to_q/to_k/to_v/to_out.0",
           "trainable_params": n_train, "ms_per_step": round(ms, 3), "images_per_sec": round(world * B / (ms * 1e-3), 2),
           "algorithmic_gflop_per_step": round(gflop, 1), "tflops": round(gflop / ms, 1), "steps_timed": n_steps,
           "final_loss": round(float(loss.item()), 5), "cuda_graph": True}
    del step, model, net, opt
    torch.cuda.empty_cache()
    return out


def run_b200(args, rank, world, local_rank):
    import torch.distributed as dist
    from polyp_image_generator_b200 import DDPMScheduler, UNet2DModel
    from polyp_image_generator_b200 import ops as ops_mod
    from polyp_image_generator_b200.training import mse_loss
    from polyp_image_generator_b200.model import polyp_unet_config      # the reference's constructor kwargs

    if not torch.cuda.is_available():
        raise RuntimeError("bench.py --impl b200 needs a CUDA device (no CPU fallback)")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    S, B = args.size, args.batch
    torch.manual_seed(0)
    model = UNet2DModel(**polyp_unet_config(S)).to(dev)
    model.train()
    sched = DDPMScheduler(num_train_timesteps=1000)
    if args.torch_optimizer:
        opt = torch.optim.AdamW(model.parameters(), lr=1e-4, fused=True, capturable=not args.no_graph)
    else:   # clip_grad_norm_(1.0) + AdamW as two streaming kernels over the flat arena (optim.py)
        from polyp_image_generator_b200 import FusedAdamW
        opt = FusedAdamW(model.parameters(), lr=1e-4, max_grad_norm=1.0)
    net = model
    if world > 1:
        from polyp_image_generator_b200.ddp import DistributedDataParallel
        net = DistributedDataParallel(model)
    ops = ops_mod.get()

    clean_host = synthetic_polyp_batch(B, S, 1234 + rank).pin_memory()
    clean_dev = clean_host.to(dev)
    gen = torch.Generator(device=dev).manual_seed(4321 + rank)
    noise_dev = torch.randn(clean_dev.shape, device=dev, generator=gen)
    t_dev = torch.randint(0, 1000, (B,), device=dev, dtype=torch.int64, generator=gen)
    params = [p for p in model.parameters() if p.requires_grad]

    def step(clean, noise, t):
        noisy = sched.add_noise(clean, noise, t)
        pred = net(noisy, t, return_dict=False)[0]
        loss = mse_loss(pred, noise)
        loss.backward()
        if args.torch_optimizer:
            torch.nn.utils.clip_grad_norm_(params, 1.0)
        opt.step()
        opt.zero_grad()
        return loss

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- instrument the dominant kernel (conv_gemm) with CUDA events on the launch stream ----
    # classes: "halo" = conv_halo_kernel (3x3 stride-1, W >= 64: the dominant kernel), "gemm" = every other
    # ddpm_conv_gemm launch (generic implicit GEMM: low-res 3x3, 1x1, linears, boundary convs), "wgrad" = ddpm_conv_wgrad
    prof = {"on": False, "ev": {"halo": [], "gemm": [], "wgrad": []}, "flops": {"halo": 0.0, "gemm": 0.0, "wgrad": 0.0},
            "k1": {"ev": [], "bytes": 0.0, "flops": 0.0}}
    orig_conv_gemm, orig_conv_wgrad = ops.conv_gemm, ops.conv_wgrad

    def conv_gemm_timed(x0, x1, taps, wgt, cout, grid, **kw):
        if not prof["on"]:
            return orig_conv_gemm(x0, x1, taps, wgt, cout, grid, **kw)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        out = orig_conv_gemm(x0, x1, taps, wgt, cout, grid, **kw)
        e1.record()
        # class by the kernel the launcher picks: 3x3 stride-1 taps at a width the halo-resident kernel serves
        is_halo = (len(taps) == 9 and ops.halo_strips(grid[2]) > 0 and kw.get("src_n", 0) in (0, grid[0]) and
                   all(tp[:3] == (0, i // 3 - 1, i % 3 - 1) for i, tp in enumerate(taps)))
        cls = "halo" if is_halo else "gemm"
        prof["ev"][cls].append((e0, e1))
        fl = ops_mod.algorithmic_conv_flops(x0, x1, taps, getattr(wgt, "_ddpm_alg_cout", cout), grid)
        prof["flops"][cls] += fl
        if len(taps) == 1 and grid[0] * grid[1] * grid[2] >= 64 * 64 * 64:
            # 1x1 convs / boundary GEMMs over >= 64x64 maps: HBM-bound by bytes (2-6 k-blocks of reduction), so their
            # roofline is the copy bandwidth: algorithmic bytes = bf16 input + output (+ residual), weights negligible
            cin = x0.shape[-1] + (x1.shape[-1] if x1 is not None else 0)
            npx = grid[0] * grid[1] * grid[2]
            by = npx * (2.0 * cin + (4.0 if kw.get("out_f32") else 2.0) * cout + (2.0 * cout if kw.get("res") is not None else 0.0))
            prof["k1"]["ev"].append((e0, e1))
            prof["k1"]["bytes"] += by
            prof["k1"]["flops"] += fl
        return out

    def conv_wgrad_timed(dy, x0, x1, taps, dw, grid, **kw):
        if not prof["on"]:
            return orig_conv_wgrad(dy, x0, x1, taps, dw, grid, **kw)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        out = orig_conv_wgrad(dy, x0, x1, taps, dw, grid, **kw)
        e1.record()
        prof["ev"]["wgrad"].append((e0, e1))
        prof["flops"]["wgrad"] += ops_mod.algorithmic_conv_flops(x0, x1, taps, dy.shape[-1], grid)
        return out

    ops.conv_gemm = conv_gemm_timed
    ops.conv_wgrad = conv_wgrad_timed

    for _ in range(args.warmup):
        step(clean_dev, noise_dev, t_dev)
    barrier()
    l_step0 = ops.launches
    step(clean_dev, noise_dev, t_dev)
    launches_per_step = ops.launches - l_step0
    eager_step = step
    if not args.no_graph:
        # the timed steps replay the SAME launches from one CUDA graph (graphs.py): host issue time is ~44 ms per
        # step in eager mode, as much as the GPU needs to execute it
        from polyp_image_generator_b200.graphs import GraphedTrainStep
        gstep = GraphedTrainStep(net, sched, opt, clean_dev.shape, max_grad_norm=1.0)
        step = gstep
    barrier()

    # ---- timed region 1: device-resident inputs ----
    sampler = ClockSampler(local_rank)
    sampler.start()
    e_begin, e_end = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    e_begin.record()
    t_host0 = time.perf_counter()
    for _ in range(args.steps):
        loss = step(clean_dev, noise_dev, t_dev)
    host_issue_ms = 1e3 * (time.perf_counter() - t_host0) / args.steps   # CPU time to ENQUEUE one step (no sync)
    e_end.record()
    barrier()
    clocks = sampler.stop()
    launches = launches_per_step * args.steps
    ms_total = e_begin.elapsed_time(e_end)
    final_loss = float(loss.item())

    # ---- timed region 2: end to end (pinned host batch in, loss out, every step) ----
    # The loop body of train_from_scratch.py:84-115 through the public API, software-pipelined one step deep the way a
    # production input pipeline is: step i+1's batch crosses PCIe on a copy stream while step i computes, and step i's
    # loss is read from a pinned buffer while step i+1 is already queued -- every step still pays its own 12.6 MB H2D
    # copy and its own D2H loss read INSIDE the timed region; none is skipped, they are just not serialised with the
    # GPU work (the reference's `loss.item()` right after `backward()` drains the queue every step).
    copy_stream = torch.cuda.Stream(device=dev)
    in_bufs = [torch.empty_like(clean_dev) for _ in range(2)]
    in_evs = [torch.cuda.Event() for _ in range(2)]
    loss_pinned = [torch.zeros(1, dtype=torch.float32).pin_memory() for _ in range(2)]
    loss_evs = [torch.cuda.Event() for _ in range(2)]

    def issue_copy(k):
        copy_stream.wait_stream(torch.cuda.current_stream())     # the buffer's previous consumer has been enqueued
        with torch.cuda.stream(copy_stream):
            in_bufs[k].copy_(clean_host, non_blocking=True)      # train_from_scratch.py:84 (pinned host -> device)
            in_evs[k].record(copy_stream)

    def e2e_loop(n_iters):
        losses = []
        issue_copy(0)
        for it in range(n_iters):
            k = it & 1
            torch.cuda.current_stream().wait_event(in_evs[k])
            if it + 1 < n_iters:
                issue_copy(k ^ 1)       # starts when step it-1 (the last reader of that buffer) is done: under step it
            clean = in_bufs[k]
            noise = torch.randn(clean.shape, device=dev)                      # :85
            t = torch.randint(0, 1000, (B,), device=dev, dtype=torch.int64)   # :88-91
            loss = step(clean, noise, t)
            loss_pinned[k].copy_(loss.detach().reshape(1), non_blocking=True)  # :115 loss.item(), read one step later
            loss_evs[k].record()
            if it > 0:
                loss_evs[k ^ 1].synchronize()
                losses.append(float(loss_pinned[k ^ 1][0]))
        loss_evs[(n_iters - 1) & 1].synchronize()
        losses.append(float(loss_pinned[(n_iters - 1) & 1][0]))
        return losses

    e2e_loop(2)             # untimed: the graph capture emptied the eager allocator pool (first draws re-allocate)
    barrier()
    e2_begin, e2_end = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e2_begin.record()
    e2e_losses = e2e_loop(args.steps)
    e2_end.record()
    barrier()
    assert len(e2e_losses) == args.steps and all(x == x for x in e2e_losses)
    ms_e2e = e2_begin.elapsed_time(e2_end)

    # ---- roofline of the dominant kernel class: the same step, eager, with CUDA events around every conv GEMM launch
    # (a graph replay cannot host per-kernel events; same kernels, same inputs, same stream)
    barrier()
    # per-launch events need every kernel serialised on ONE stream: the timed steps run weight gradients on a second
    # stream (parallel graph branches), which would smear concurrent kernels into each other's event brackets
    os.environ["DDPM_WGRAD_STREAM"] = "0"
    for _ in range(2):      # the graph capture emptied the eager allocator pool: refill it before timing launches
        eager_step(clean_dev, noise_dev, t_dev)
    barrier()
    prof["on"] = True
    e3_begin, e3_end = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e3_begin.record()
    n_prof_steps = max(1, min(args.steps, 5))
    for _ in range(n_prof_steps):
        eager_step(clean_dev, noise_dev, t_dev)
    e3_end.record()
    barrier()
    prof["on"] = False
    ms_prof_total = e3_begin.elapsed_time(e3_end)
    cls_ms = {k: sum(a.elapsed_time(b) for a, b in v) for k, v in prof["ev"].items()}
    cls_tf = {k: (prof["flops"][k] / (cls_ms[k] * 1e-3) / 1e12 if cls_ms[k] > 0 else 0.0) for k in cls_ms}
    k1_ms = sum(a.elapsed_time(b) for a, b in prof["k1"]["ev"])
    gemm_ms = cls_ms["halo"]
    gemm_launches = len(prof["ev"]["halo"])
    gemm_flops = prof["flops"]["halo"]
    all_ms = sum(cls_ms.values())
    all_tf = sum(prof["flops"].values()) / (all_ms * 1e-3) / 1e12 if all_ms > 0 else 0.0
    ops.conv_wgrad = orig_conv_wgrad
    ops.conv_gemm = orig_conv_gemm
    step = eager_step
    # HBM-bound kernel classes of the same single-stream eager step: algorithmic bytes over
    # CUDA-event time, for the launches that are large enough to be bandwidth- rather than latency-sized (>= 64x64)
    hbm_kernels = {}
    # EVERY rank runs these steps: under DDP each step all-reduces the gradients, so a rank-0-only pass would leave
    # rank 0 waiting in a collective the other ranks never enter
    pr = ops_mod.OpProfiler(ops)
    pr.by_shape = True
    pr.start()
    try:
        for _ in range(2):
            eager_step(clean_dev, noise_dev, t_dev)
    finally:
        table = pr.stop()
    barrier()
    if rank == 0:
        agg = {}
        for key, v in table.items():
            name, _, shape = key.partition("|")
            if not v["bytes"] or v["ms"] <= 0:
                continue
            if name.startswith("gn_") and not (shape.startswith("128x128") or shape.startswith("64x64")):
                continue
            if name not in ("gn_fwd_from_csum", "gn_fwd", "gn_bwd_apply", "gn_bwd", "adamw_flat", "sumsq", "add_noise",
                            "mse_fwd_bwd"):
                continue
            a = agg.setdefault(name, {"bytes": 0.0, "ms": 0.0, "calls": 0})
            a["bytes"] += v["bytes"]; a["ms"] += v["ms"]; a["calls"] += v["calls"]
        hbm_peak = None
        try:
            with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
                hbm_peak = json.load(f).get("hbm_gbs")
        except Exception:  # noqa: BLE001
            pass
        hbm_peak = hbm_peak or 6500.0
        for name, a in agg.items():
            gbps = a["bytes"] / (a["ms"] * 1e-3) / 1e9
            hbm_kernels[name] = {"gbps": round(gbps, 1), "frac_of_copy_peak": round(gbps / hbm_peak, 3),
                                 "ms_per_step": round(a["ms"] / 2, 3), "calls_per_step": a["calls"] // 2}
        hbm_kernels["_peak_gbs"] = hbm_peak
        hbm_kernels["_note"] = ("GroupNorm rows: launches at >= 64x64 only (smaller maps are latency-sized); "
                                "add_noise / mse at 12.6 MB are latency-sized too")
    os.environ.pop("DDPM_WGRAD_STREAM", None)

    # ---- secondary metric: reverse-diffusion sampling (BASELINE configs[2]: 256 images over 8 GPUs = 32 / GPU) ----
    sampling = None
    if not args.no_sampling:
        from polyp_image_generator_b200 import DDPMPipeline
        model.eval()
        pipe = DDPMPipeline(unet=model, scheduler=DDPMScheduler(num_train_timesteps=1000))
        sb, n_steps = args.sampling_batch, args.sampling_steps

        def run_sampling(gen):
            # num_inference_steps = n_steps: the same per-step work as the 1000-step loop (one UNet forward + one
            # scheduler step + one noise draw per step), timed over n_steps steps and reported per step
            pipe(batch_size=sb, generator=gen, num_inference_steps=3, output_type="uint8")        # warm-up
            barrier()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            pipe(batch_size=sb, generator=gen, num_inference_steps=n_steps, output_type="uint8")
            e1.record()
            barrier()
            return e0.elapsed_time(e1) / n_steps

        ms_dev = run_sampling(None)                                          # in-kernel Philox noise
        ms_ref = run_sampling(torch.Generator("cpu").manual_seed(0 + rank))  # reference RNG contract (CPU draws)
        ts = torch.tensor([ms_dev, ms_ref], device=dev, dtype=torch.float64)
        if world > 1:
            dist.all_reduce(ts, op=dist.ReduceOp.MAX)
        ms_dev, ms_ref = float(ts[0]), float(ts[1])
        sampling = {
            "per_gpu_batch": sb, "steps_timed": n_steps,
            "ms_per_step_device_noise": round(ms_dev, 3), "ms_per_step_cpu_generator": round(ms_ref, 3),
            "images_per_sec_1000_steps_device_noise": round(world * sb / (ms_dev * 1000 * 1e-3), 3),
            "images_per_sec_1000_steps_cpu_generator": round(world * sb / (ms_ref * 1000 * 1e-3), 3),
            "unit": "img/s (whole job, 1000 reverse steps per image)",
            "fwd_tflops_device_noise": round(FWD_GFLOP_PER_IMG.get(S, 0.0) * sb / ms_dev, 1),
        }
        model.train()

    # ---- secondary metric: LoRA fine-tune step at 256x256 on the celebahq-architecture UNet (BASELINE configs[3]) ----
    lora = None
    if not args.no_lora:
        try:
            lora = run_lora_finetune(args, dev, world, rank, barrier)
            if args.lora_batch2 > 0:     # "B = 8/GPU (reference train_batch_size) plus a B = 32 point"
                lora["batch_%d_point" % args.lora_batch2] = run_lora_finetune(args, dev, world, rank, barrier,
                                                                              batch=args.lora_batch2)
        except Exception as e:  # noqa: BLE001 -- a secondary section must not take the headline line down with it
            lora = {"error": repr(e)[:300]} if lora is None else dict(lora, batch2_error=repr(e)[:300])

    # ---- optional per-op breakdown (after the timed regions; CUDA events around every C-ABI op) ----
    if args.breakdown:       # all ranks step together (DDP collectives); rank 0 reports
        os.environ["DDPM_WGRAD_STREAM"] = "0"       # per-op events: one stream
        ops.conv_gemm = orig_conv_gemm
        model.train()
        for _ in range(2):
            step(clean_dev, noise_dev, t_dev)
        pr = ops_mod.OpProfiler(ops)
        pr.by_shape = args.breakdown_by_shape
        pr.start()
        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        ev0.record()
        for _ in range(2):
            step(clean_dev, noise_dev, t_dev)
        ev1.record()
        table = pr.stop()
        barrier()
        if rank == 0:
            wall = ev0.elapsed_time(ev1) / 2
            rows = sorted(table.items(), key=lambda kv: -kv[1]["ms"])
            acc = sum(v["ms"] for _, v in rows) / 2
            out = {"ms_per_step_profiled": wall, "ms_in_ops": acc, "ops": {}}
            print(f"[breakdown] step {wall:.2f} ms, in C-ABI ops {acc:.2f} ms (rest = torch optimizer/clip/alloc)",
                  file=sys.stderr)
            for name, v in rows:
                ms = v["ms"] / 2
                tf = v["flops"] / 2 / (ms * 1e-3) / 1e12 if v["flops"] and ms > 0 else None
                gb = v["bytes"] / 2 / (ms * 1e-3) / 1e9 if v["bytes"] and ms > 0 else None
                out["ops"][name] = {"calls": v["calls"] // 2, "ms": round(ms, 3), "tflops": tf and round(tf, 1),
                                    "gbps": gb and round(gb, 1)}
                print(f"[breakdown] {name:36s} calls {v['calls'] // 2:5d}  {ms:8.3f} ms  "
                      f"{'%.1f TFLOP/s' % tf if tf else ''}{'%.0f GB/s' % gb if gb else ''}", file=sys.stderr)
            os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
            with open(os.path.join(ROOT, "gpurun_out", "op_breakdown.json"), "w") as f:
                json.dump(out, f, indent=1)
    os.environ.pop("DDPM_WGRAD_STREAM", None)

    tt = torch.tensor([ms_total, ms_e2e], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
    ms_total, ms_e2e = float(tt[0]), float(tt[1])

    if rank != 0:
        return
    peaks = {}
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            peaks = json.load(f)
    except Exception:  # noqa: BLE001
        pass
    peak_tf = peaks.get("bf16_tflops_sustained")
    peak_src = "measured (MEASURED_PEAKS.json bf16_tflops_sustained)"
    if peak_tf is None:
        peak_tf, peak_src = 1400.0, "fallback (B200_PROFILING.md sustained ~1.4 PFLOP/s)"
    achieved_tf = gemm_flops / (gemm_ms * 1e-3) / 1e12 if gemm_ms > 0 else 0.0
    ms_step = ms_total / args.steps
    value = world * B / (ms_step * 1e-3)
    e2e_value = world * B / (ms_e2e / args.steps * 1e-3)
    step_gflop = 3.0 * FWD_GFLOP_PER_IMG.get(S, 0.0) * B

    cpu = None
    if world == 1 and not args.no_cpu_baseline:
        r, both = cpu_oracle_both(S, args.cpu_batch, 3, budget_s=60.0)
        cpu = {"value": round(r["value"], 4), "unit": UNIT, "cores": r["cores"], "kind": "port", "sample": r["sample"],
               "precisions": both}

    # DRAM traffic of ONE launch of the dominant kernel, from the `ncu --set full` capture of the shipped build that
    # profiles/r2_ncu_halo_pair.json summarises (dram__bytes_read.sum + dram__bytes_write.sum; 128->128 3x3 @128^2, B 64)
    traffic, traffic_note = None, "no ncu summary found (profiles/r2_ncu_halo_pair.json)"
    try:
        with open(os.path.join(ROOT, "profiles", "r2_ncu_halo_pair.json")) as f:
            cap = json.load(f)
        if S == 128 and B == 64:
            traffic = cap["dram_bytes_per_launch"]
        traffic_note = cap.get("note", "")
    except Exception:  # noqa: BLE001
        pass
    total_alg = step_gflop * 1e9
    halo_share_flops = gemm_flops / n_prof_steps / total_alg if total_alg > 0 else 0.0
    gn_rows = {k: v for k, v in hbm_kernels.items() if k.startswith("gn_")}
    gn_bytes = sum(v["gbps"] * v["ms_per_step"] for v in gn_rows.values())
    gn_ms = sum(v["ms_per_step"] for v in gn_rows.values())
    roofline = {
        "kernel": "conv_halo_pair_kernel (tcgen05 cta_group::2 halo-resident 3x3 stride-1 conv, fprop + dgrad, "
                  f"W >= 32): {100 * halo_share_flops:.0f} % of the step's algorithmic FLOPs",
        "bound": "tensor",
        "achieved": round(achieved_tf, 2), "peak": peak_tf, "unit": "TFLOP/s",
        "frac": round(achieved_tf / peak_tf, 4),
        "traffic": traffic, "traffic_note": traffic_note,
        "peak_source": peak_src,
        "flops_counted": "algorithmic (SURVEY.md §8d): stride-2 dgrads at the FLOPs of the stride-2 conv (not of the "
                         "zero-inserted correlation), conv_in / conv_out at 3 image channels (not the padded operands)",
        "launches_timed": gemm_launches, "kernel_ms_per_step": round(gemm_ms / n_prof_steps, 3),
        "share_of_step_time": round(gemm_ms / ms_prof_total, 4),
        "share_of_step_flops": round(halo_share_flops, 4),
        "timed_in": "eager single-stream re-run of the same step with per-launch CUDA events (graph replays cannot "
                    "host them; the timed steps additionally overlap weight gradients on a second stream)",
        "whole_step_tflops": round(step_gflop / ms_step, 2),
        "whole_step_frac_of_peak": round(step_gflop / ms_step / peak_tf, 4),
        "all_conv_gemms_tflops": round(all_tf, 1), "all_conv_gemms_frac_of_peak": round(all_tf / peak_tf, 4),
        "groupnorm_ge64_gbps": round(gn_bytes / gn_ms, 1) if gn_ms > 0 else None,
        "groupnorm_ge64_frac_of_copy_peak": round(gn_bytes / gn_ms / hbm_kernels.get("_peak_gbs", 6536.0), 4)
        if gn_ms > 0 else None,
        "other_conv_kernels": {
            "generic_gemm_tflops (1x1, stride-2, <= 16x16 3x3, linears, boundary convs)": round(cls_tf["gemm"], 1),
            "generic_gemm_frac_of_peak": round(cls_tf["gemm"] / peak_tf, 4),
            "generic_gemm_ms_per_step": round(cls_ms["gemm"] / n_prof_steps, 3),
            "generic_1x1_ge64 (memory-bound class: 1x1 shortcut convs, conv_in, conv_out dgrad at >= 64x64)": {
                "ms_per_step": round(k1_ms / n_prof_steps, 3), "launches_per_step": len(prof["k1"]["ev"]) // n_prof_steps,
                "gbps": round(prof["k1"]["bytes"] / (k1_ms * 1e-3) / 1e9, 1) if k1_ms > 0 else None,
                "frac_of_copy_peak": round(prof["k1"]["bytes"] / (k1_ms * 1e-3) / 1e9 /
                                           hbm_kernels.get("_peak_gbs", 6536.0), 4) if k1_ms > 0 else None,
                "tflops": round(prof["k1"]["flops"] / (k1_ms * 1e-3) / 1e12, 1) if k1_ms > 0 else None},
            "generic_rest_tflops (stride-2, <= 16x16 3x3, linears)": round(
                (prof["flops"]["gemm"] - prof["k1"]["flops"]) / ((cls_ms["gemm"] - k1_ms) * 1e-3) / 1e12, 1)
            if cls_ms["gemm"] > k1_ms else None,
            "wgrad_tflops (row-resident + generic)": round(cls_tf["wgrad"], 1),
            "wgrad_frac_of_peak": round(cls_tf["wgrad"] / peak_tf, 4),
            "wgrad_ms_per_step": round(cls_ms["wgrad"] / n_prof_steps, 3),
        },
        "hbm_bound_kernels": hbm_kernels,
    }

    line = {
        "metric": METRIC, "value": round(value, 2), "unit": UNIT, "n_gpus": world, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": round(ms_step, 3), "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "bf16", "data": "synthetic",
        "config": {
            "workload": f"UNet2DModel {S}x{S} train_from_scratch DDPM step, bf16 tensor-core compute, "
                        f"batch {B}/GPU, {'DDP dp%d' % world if world > 1 else 'single GPU'}",
            "per_gpu_batch": B, "global_batch": B * world, "image_size": S, "num_train_timesteps": 1000,
            "params": 113673219, "optimizer": ("torch AdamW(fused) + clip_grad_norm_(1.0)" if args.torch_optimizer else
                                            "FusedAdamW: global-norm clip(1.0) + AdamW, 2 streaming kernels over the flat arena"),
            "l2": "per-step working set (activations + 455 MB weights/grads) is far larger than the 126 MB L2",
            "algorithmic_gflop_per_step": round(step_gflop, 1), "final_loss": round(final_loss, 5),
            "host_issue_ms_per_step": round(host_issue_ms, 2),
            "cuda_graph": not args.no_graph,
        },
        "clocks": clocks,
        "e2e": {"value": round(e2e_value, 2), "unit": UNIT, "h2d_bytes_per_step": clean_host.numel() * 4,
                "d2h_bytes_per_step": 4, "ms_per_step": round(ms_e2e / args.steps, 3),
                "pipelining": "one step deep: batch i+1 is copied on a second stream while step i runs, loss i is read "
                              "from pinned memory while step i+1 is queued; every step's copy and read are inside the "
                              "timed region"},
        "gpu_launches": launches,
        "roofline": roofline,
        "cpu_baseline": cpu,
        "sampling": sampling,
        "lora_finetune": lora,
    }
    print(json.dumps(line), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--batch", type=int, default=64, help="per-GPU batch (BASELINE configs[1]: 64)")
    ap.add_argument("--size", type=int, default=128)
    ap.add_argument("--cpu-batch", type=int, default=4, help="batch of the bounded CPU sample")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-graph", action="store_true", help="issue every kernel from Python instead of a CUDA graph")
    ap.add_argument("--torch-optimizer", action="store_true", help="torch clip_grad_norm_ + AdamW(fused) instead of "
                    "optim.FusedAdamW")
    ap.add_argument("--no-lora", action="store_true", help="skip the secondary LoRA fine-tune measurement")
    ap.add_argument("--lora-batch", type=int, default=8, help="per-GPU batch of the LoRA measurement "
                    "(config_diffusion.py:7 train_batch_size = 8)")
    ap.add_argument("--lora-batch2", type=int, default=32, help="second per-GPU batch of the LoRA measurement (0: skip)")
    ap.add_argument("--lora-size", type=int, default=256)
    ap.add_argument("--no-sampling", action="store_true", help="skip the secondary sampling measurement")
    ap.add_argument("--sampling-batch", type=int, default=32, help="images per GPU in the sampling measurement")
    ap.add_argument("--sampling-steps", type=int, default=50, help="reverse steps timed (reported per step; "
                    "SURVEY.md §8d: >= 50)")
    ap.add_argument("--breakdown", action="store_true", help="after timing, print a per-op CUDA-event breakdown")
    ap.add_argument("--breakdown-by-shape", action="store_true", help="split conv / GroupNorm rows by layer shape")
    args = ap.parse_args()
    if args.warmup < 3 and args.impl == "b200":
        args.warmup = 3

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))

    if args.impl == "reference":
        run_reference(args, rank)
        return
    if world > 1:
        import torch.distributed as dist
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        torch.cuda.set_device(local_rank)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    try:
        run_b200(args, rank, world, local_rank)
    finally:
        if world > 1:
            import torch.distributed as dist
            dist.destroy_process_group()


if __name__ == "__main__":
    main()
