"""Per-kernel microbenchmarks on the UNet's real shapes (run under gpurun; CUDA-event timing, L2 flushed between
iterations by rotating through more input buffers than fit in the 126 MB L2).

    python profiles/bench_kernels.py [conv|wgrad|gn|small|all] [--batch 64] [--iters 10]

Prints one line per shape: achieved TFLOP/s (GEMM kernels) or GB/s (streaming kernels) and the fraction of the
measured peak from MEASURED_PEAKS.json.
"""
from __future__ import annotations

import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import torch

from polyp_image_generator_b200 import ops as ops_mod
from polyp_image_generator_b200.ops import taps_1x1, taps_3x3


def peaks():
    try:
        return json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:  # noqa: BLE001
        return {"bf16_tflops": 1590.0, "hbm_gbs": 6650.0}


WARMUP = 3


def timeit(fn, iters, nbuf):
    for i in range(WARMUP):
        fn(i % nbuf)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(iters):
        fn(i % nbuf)
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters


def bf(*shape):
    return (torch.randn(*shape, device="cuda") * 0.5).to(torch.bfloat16)


# (h, cin, cout, taps, count-per-forward) at S=128 -- SURVEY.md Appendix D
CONV_SHAPES = [
    (128, 128, 128, 9), (128, 256, 128, 9), (128, 256, 128, 1),
    (64, 128, 128, 9), (64, 256, 256, 9), (64, 384, 128, 9), (64, 256, 128, 9),
    (32, 256, 256, 9), (32, 512, 256, 9), (32, 128, 256, 9),
    (16, 256, 256, 9), (16, 512, 512, 9), (16, 768, 256, 9), (16, 512, 256, 9),
    (8, 512, 512, 9), (8, 1024, 512, 9), (8, 1024, 512, 1),
    (4, 512, 512, 9), (4, 1024, 512, 9),
]


def bench_conv(ops, B, iters, pk, first=None):
    print(f"# conv_gemm (fprop / dgrad), batch {B}; peak = {pk['bf16_tflops']} TFLOP/s (burst, kernel timed alone)")
    for (h, cin, cout, taps) in CONV_SHAPES[:first]:
        per = B * h * h * cin * 2
        nbuf = max(2, min(8, int(300e6 // per) + 1))
        xs = [bf(B, h, h, cin) for _ in range(nbuf)]
        w = bf(cout, taps * cin) * 0.1
        bias = torch.randn(cout, device="cuda")
        out = torch.empty(B, h, h, cout, device="cuda", dtype=torch.bfloat16)
        tp = taps_3x3(cin) if taps == 9 else taps_1x1()
        ms = timeit(lambda i: ops.conv_gemm(xs[i], None, tp, w, cout, (B, h, h), bias=bias, out=out), iters, nbuf)
        fl = 2.0 * B * h * h * cout * cin * taps
        tf = fl / ms / 1e9
        print(f"conv {h:3d}x{h:<3d} {cin:4d}->{cout:<4d} k{taps}  {ms:8.3f} ms  {tf:7.1f} TFLOP/s  "
              f"{tf / pk['bf16_tflops']:.3f} of peak", flush=True)


def bench_convgn(ops, B, iters, pk, first=None):
    """dgrad conv with the GroupNorm-backward first half fused into the epilogue vs the plain conv."""
    print(f"# conv_gemm dgrad: plain vs GroupNorm-backward epilogue fusion, batch {B}")
    for (h, cin, cout) in [(128, 128, 128), (128, 128, 256), (64, 128, 128), (32, 256, 256), (16, 256, 256)][:first]:
        g = bf(B, h, h, cin)
        w = bf(cout, 9 * cin) * 0.1
        x = bf(B, h, h, cout)
        gam, bet = torch.ones(cout, device="cuda"), torch.zeros(cout, device="cuda")
        _, _, coef = ops.gn_fwd(x, None, 32, 1e-5, gam, bet, True, want_coef=True)
        out = torch.empty(B, h, h, cout, device="cuda", dtype=torch.bfloat16)
        sums = torch.zeros(B, cout, 2, device="cuda")
        tp = taps_3x3(cin)
        ms_p = timeit(lambda i: ops.conv_gemm(g, None, tp, w, cout, (B, h, h), out=out), iters, 1)
        ms_f = timeit(lambda i: ops.conv_gemm(g, None, tp, w, cout, (B, h, h), out=out,
                                              gn=(x, None, coef, True, sums)), iters, 1)
        fl = 2.0 * B * h * h * cout * cin * 9
        print(f"convgn {h:3d}x{h:<3d} {cin:4d}->{cout:<4d} plain {ms_p:7.3f} ms {fl / ms_p / 1e9:7.1f} TFLOP/s | "
              f"fused {ms_f:7.3f} ms {fl / ms_f / 1e9:7.1f} TFLOP/s  (+{ms_f - ms_p:.3f} ms)", flush=True)


def bench_epi(ops, B, iters, pk, first=None):
    """Halo conv 3x3 with each epilogue fusion switched on separately: which one costs what (the epilogue warps have one
    tile's mainloop -- 9216 tensor clocks at 128->128 -- to drain 256 x 128 accumulators)."""
    print(f"# conv_gemm 3x3 epilogue variants, batch {B}")
    for (h, cin, cout) in [(128, 128, 128), (128, 256, 128), (64, 128, 128), (64, 256, 256)][:first]:
        nb = B if h <= 128 else max(2, B // 4)
        x = bf(nb, h, h, cin)
        w = bf(cout, 9 * cin) * 0.1
        bias = torch.randn(cout, device="cuda")
        temb = torch.randn(nb, cout, device="cuda")
        res = bf(nb, h, h, cout)
        gx = bf(nb, h, h, cout)
        gam, bet = torch.ones(cout, device="cuda"), torch.zeros(cout, device="cuda")
        _, _, coef = ops.gn_fwd(gx, None, 32, 1e-5, gam, bet, True, want_coef=True)
        out = torch.empty(nb, h, h, cout, device="cuda", dtype=torch.bfloat16)
        cs = torch.zeros(nb, cout // 4, 2, device="cuda")      # forward statistics: per 4-channel granule
        gs = torch.zeros(nb, cout, 2, device="cuda")           # backward fusion sums: per channel
        tp = taps_3x3(cin)
        grid = (nb, h, h)
        variants = [
            ("plain", {}), ("bias", dict(bias=bias)), ("bias+temb", dict(bias=bias, temb=temb)),
            ("bias+res", dict(bias=bias, res=res)), ("bias+temb+csum", dict(bias=bias, temb=temb, csum=cs)),
            ("bias+res+csum", dict(bias=bias, res=res, csum=cs)), ("gn-bwd fusion", dict(gn=(gx, None, coef, True, gs))),
        ]
        fl = 2.0 * nb * h * h * cout * cin * 9
        for name, kw in variants:
            ms = timeit(lambda i: ops.conv_gemm(x, None, tp, w, cout, grid, out=out, **kw), iters, 1)
            print(f"epi {h:3d}x{h:<3d} {cin:4d}->{cout:<4d} {name:16s} {ms:7.3f} ms {fl / ms / 1e9:7.1f} TFLOP/s "
                  f"{fl / ms / 1e9 / pk['bf16_tflops']:.3f} of peak", flush=True)


def bench_hires(ops, B, iters, pk, first=None):
    """224^2 / 256^2 shapes of BASELINE configs[3]/[4] and the reference's default image_size (batch 8)."""
    nb = 8
    print(f"# conv_gemm / conv_wgrad at 224^2 and 256^2, batch {nb}")
    for (h, cin, cout, taps) in [(256, 128, 128, 9), (256, 256, 128, 9), (224, 128, 128, 9), (112, 128, 128, 9),
                                 (256, 256, 128, 1), (128, 128, 128, 9)][:first]:
        x = bf(nb, h, h, cin)
        w = bf(cout, taps * cin) * 0.1
        bias = torch.randn(cout, device="cuda")
        out = torch.empty(nb, h, h, cout, device="cuda", dtype=torch.bfloat16)
        tp = taps_3x3(cin) if taps == 9 else taps_1x1()
        fl = 2.0 * nb * h * h * cout * cin * taps
        ms = timeit(lambda i: ops.conv_gemm(x, None, tp, w, cout, (nb, h, h), bias=bias, out=out), iters, 1)
        dy = bf(nb, h, h, cout)
        dw = torch.zeros(cout, taps * cin, device="cuda")
        ms_w = timeit(lambda i: ops.conv_wgrad(dy, x, None, tp, dw, (nb, h, h)), iters, 1)
        print(f"hires {h:3d}x{h:<3d} {cin:4d}->{cout:<4d} k{taps} conv {ms:7.3f} ms {fl / ms / 1e9:7.1f} TFLOP/s | "
              f"wgrad {ms_w:7.3f} ms {fl / ms_w / 1e9:7.1f} TFLOP/s", flush=True)


def bench_wgrad(ops, B, iters, pk, first=None):
    print(f"# conv_wgrad, batch {B}")
    for (h, cin, cout, taps) in CONV_SHAPES[:first]:
        if cout % 64:
            continue
        x = bf(B, h, h, cin)
        dy = bf(B, h, h, cout)
        dw = torch.zeros(cout, taps * cin, device="cuda")
        tp = taps_3x3(cin) if taps == 9 else taps_1x1()
        ms = timeit(lambda i: ops.conv_wgrad(dy, x, None, tp, dw, (B, h, h)), iters, 1)
        fl = 2.0 * B * h * h * cout * cin * taps
        tf = fl / ms / 1e9
        print(f"wgrad {h:3d}x{h:<3d} {cin:4d}->{cout:<4d} k{taps}  {ms:8.3f} ms  {tf:7.1f} TFLOP/s  "
              f"{tf / pk['bf16_tflops']:.3f} of peak", flush=True)


def bench_gn(ops, B, iters, pk, first=None):
    print(f"# GroupNorm+SiLU, batch {B}; peak = {pk['hbm_gbs']} GB/s (measured copy)")
    for (h, c0, c1) in [(128, 128, 0), (128, 128, 128), (64, 128, 0), (64, 256, 128), (32, 256, 0), (32, 256, 256),
                        (16, 512, 256), (8, 512, 512), (4, 512, 0)][:first]:
        C = c0 + c1
        xa = bf(B, h, h, c0)
        xb = bf(B, h, h, c1) if c1 else None
        gam, bet = torch.ones(C, device="cuda"), torch.zeros(C, device="cuda")
        elems = B * h * h * C
        ms_s = timeit(lambda i: ops.gn_stats(xa, xb, 32), iters, 1)
        stats = ops.gn_stats(xa, xb, 32)
        y = torch.empty(B, h, h, C, device="cuda", dtype=torch.bfloat16)
        ms_a = timeit(lambda i: ops.gn_apply(xa, xb, 32, stats, 1e-5, gam, bet, True, out=y), iters, 1)
        ms_f = timeit(lambda i: ops.gn_fwd(xa, xb, 32, 1e-5, gam, bet, True, out=y), iters, 1)
        dy = bf(B, h, h, C)
        ms_b = timeit(lambda i: ops.gn_bwd(xa, xb, 32, stats, 1e-5, gam, bet, True, dy), iters, 1)
        sums = torch.zeros(B, C, 2, device="cuda")
        ms_ba = timeit(lambda i: ops.gn_bwd_apply(xa, xb, 32, stats, 1e-5, gam, dy, sums), iters, 1)
        ms_ba2 = timeit(lambda i: ops.gn_bwd_apply(xa, xb, 32, stats, 1e-5, gam, dy, sums, add0=dy), iters, 1)
        print(f"   bwd_apply {ms_ba:7.3f} ms {6 * elems / ms_ba / 1e6:6.0f} GB/s | with add0 {ms_ba2:7.3f} ms "
              f"{8 * elems / ms_ba2 / 1e6:6.0f} GB/s ({8 * elems / ms_ba2 / 1e6 / pk['hbm_gbs']:.2f})")
        print(f"gn {h:3d}x{h:<3d} c{c0}+{c1:<4d} stats {ms_s:7.3f} ms {2 * elems / ms_s / 1e6:6.0f} GB/s | "
              f"apply {ms_a:7.3f} ms {4 * elems / ms_a / 1e6:6.0f} GB/s ({4 * elems / ms_a / 1e6 / pk['hbm_gbs']:.2f}) | "
              f"fused fwd {ms_f:7.3f} ms {4 * elems / ms_f / 1e6:6.0f} GB/s ({4 * elems / ms_f / 1e6 / pk['hbm_gbs']:.2f}) | "
              f"bwd {ms_b:7.3f} ms {6 * elems / ms_b / 1e6:6.0f} GB/s ({6 * elems / ms_b / 1e6 / pk['hbm_gbs']:.2f}; 6 B/elem: x, dy, dx)",
              flush=True)


def bench_small(ops, B, iters, pk):
    print(f"# 3-channel convs + elementwise, batch {B}, 128x128")
    S, C = 128, 128
    x = torch.randn(B, 3, S, S, device="cuda")
    ms = timeit(lambda i: ops.im2col3(x), iters, 1)
    print(f"im2col3          {ms:7.3f} ms  ({(x.numel() * 4 + B * S * S * 64 * 2) / ms / 1e6:6.0f} GB/s algorithmic)")
    pat = ops.im2col3(x)
    wf = bf(C, 64) * 0.1
    b_in = torch.randn(C, device="cuda")
    ms = timeit(lambda i: ops.conv_gemm(pat, None, taps_1x1(), wf, C, (B, S, S), bias=b_in), iters, 1)
    print(f"conv_in GEMM     {ms:7.3f} ms  ({(B * S * S * (64 + C) * 2) / ms / 1e6:6.0f} GB/s algorithmic)")
    a = bf(B, S, S, C)
    wo = bf(32, 9 * C) * 0.1
    ms = timeit(lambda i: ops.conv_gemm(a, None, taps_3x3(C), wo, 32, (B, S, S), out_f32=True), iters, 1)
    print(f"conv_out GEMM    {ms:7.3f} ms")
    nz = torch.randn_like(x)
    t = torch.randint(0, 1000, (B,), device="cuda")
    tab = torch.rand(1000, device="cuda")
    ms = timeit(lambda i: ops.add_noise(x, nz, t, tab, tab), iters, 1)
    print(f"add_noise        {ms:7.4f} ms  {12 * x.numel() / ms / 1e6:6.0f} GB/s ({12 * x.numel() / ms / 1e6 / pk['hbm_gbs']:.2f})")
    ms = timeit(lambda i: ops.mse_fwd_bwd(x, nz), iters, 1)
    print(f"mse fwd+bwd      {ms:7.4f} ms  {12 * x.numel() / ms / 1e6:6.0f} GB/s ({12 * x.numel() / ms / 1e6 / pk['hbm_gbs']:.2f})")
    ms = timeit(lambda i: ops.scheduler_step(x, nz, nz, 0.5, 0.5, 0.1, 0.9, 0.1, 1.0), iters, 1)
    print(f"scheduler step   {ms:7.4f} ms  {16 * x.numel() / ms / 1e6:6.0f} GB/s ({16 * x.numel() / ms / 1e6 / pk['hbm_gbs']:.2f})")
    ms = timeit(lambda i: ops.scheduler_step_philox(x, nz, 0.5, 0.5, 0.1, 0.9, 0.1, 1.0, 1, i), iters, 1)
    print(f"step (philox)    {ms:7.4f} ms  {12 * x.numel() / ms / 1e6:6.0f} GB/s ({12 * x.numel() / ms / 1e6 / pk['hbm_gbs']:.2f})")


def bench_linear(ops, B, iters, pk):
    """The fp32 time-embedding linears at the sizes of the step; weights rotate through > L2 of buffers (cold reads)."""
    print(f"# time-embedding linears (fp32 SIMT), batch {B}")
    for k, n in ((128, 512), (512, 512), (512, 9984)):
        nbuf = max(2, int(200e6 // (n * k * 4)) + 1) if n * k * 4 > 4e6 else 2
        ws = [torch.randn(n, k, device="cuda") * 0.05 for _ in range(nbuf)]
        dws = [torch.zeros(n, k, device="cuda") for _ in range(min(nbuf, 4))]
        x = torch.randn(B, k, device="cuda")
        dy = torch.randn(B, n, device="cuda")
        b = torch.randn(n, device="cuda")
        db = torch.zeros(n, device="cuda")
        wbytes = n * k * 4
        ms = timeit(lambda i: ops.linear_f32(x, ws[i % nbuf], b, True), iters, nbuf)
        print(f"linear  {B}x{k}->{n}  fwd   {ms * 1e3:8.1f} us  ({wbytes / ms / 1e6:6.0f} GB/s of weights)")
        ms = timeit(lambda i: ops.linear_f32_wgrad(x, dy, dws[i % len(dws)], db, True), iters, len(dws))
        print(f"linear  {B}x{k}->{n}  wgrad {ms * 1e3:8.1f} us  ({2 * wbytes / ms / 1e6:6.0f} GB/s of dW read+write)")
        ms = timeit(lambda i: ops.linear_f32_dgrad(dy, ws[i % nbuf], x, True), iters, nbuf)
        print(f"linear  {B}x{k}->{n}  dgrad {ms * 1e3:8.1f} us  ({wbytes / ms / 1e6:6.0f} GB/s of weights)")


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("what", nargs="?", default="all")
    ap.add_argument("--batch", type=int, default=64)
    ap.add_argument("--iters", type=int, default=10)
    ap.add_argument("--first", type=int, default=None, help="only the first N conv shapes (short ncu runs)")
    ap.add_argument("--warmup", type=int, default=3, help="untimed launches per shape (0 for ncu captures)")
    a = ap.parse_args()
    WARMUP = a.warmup
    ops = ops_mod.get()
    pk = peaks()
    torch.manual_seed(0)
    if a.what in ("conv", "all"):
        bench_conv(ops, a.batch, a.iters, pk, a.first)
    if a.what in ("convgn", "all"):
        bench_convgn(ops, a.batch, a.iters, pk, a.first)
    if a.what in ("wgrad", "all"):
        bench_wgrad(ops, a.batch, a.iters, pk, a.first)
    if a.what in ("epi", "all"):
        bench_epi(ops, a.batch, a.iters, pk, a.first)
    if a.what in ("hires", "all"):
        bench_hires(ops, a.batch, a.iters, pk, a.first)
    if a.what in ("gn", "all"):
        bench_gn(ops, a.batch, a.iters, pk, a.first)
    if a.what in ("small", "all"):
        bench_small(ops, a.batch, a.iters, pk)
    if a.what in ("linear", "all"):
        bench_linear(ops, a.batch, a.iters, pk)
