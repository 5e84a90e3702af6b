"""GPU bring-up harness: runs every kernel group against torch fp32 references and prints diagnostics.

    python tests/gpu_bringup.py            # all groups, each in its own subprocess (a trap poisons the context)
    python tests/gpu_bringup.py gemm       # one group in-process

Not a pytest file (tests/test_*_gpu.py hold the parity tests proper); this is the tool used to debug kernels on a
gpurun box where nothing streams back until the call ends.
"""
from __future__ import annotations

import os
import subprocess
import sys
import time
import traceback

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

import torch
import torch.nn.functional as F

GROUPS = ["elementwise", "gn", "misc", "smallconv", "attention", "gemm", "wgrad", "gnfuse"]


def rel(a, b):
    a, b = a.float(), b.float()
    return ((a - b).norm() / (b.norm() + 1e-30)).item()


def report(name, got, want, tol):
    r = rel(got, want)
    mx = (got.float() - want.float()).abs().max().item()
    ok = r <= tol and bool(torch.isfinite(got.float()).all())
    print(f"  [{'OK ' if ok else 'BAD'}] {name:58s} rel={r:.3e} maxabs={mx:.3e} (tol {tol:g})", flush=True)
    return ok


def bf(x):
    return x.to(torch.bfloat16)


def g_elementwise(ops):
    ok = True
    dev = "cuda"
    torch.manual_seed(0)
    from oracle.ddpm import DDPMScheduler as OS
    osch = OS()
    for shape in [(4, 3, 64, 64), (3, 3, 7, 5), (64, 3, 128, 128)]:
        x0 = torch.randn(shape, device=dev)
        nz = torch.randn(shape, device=dev)
        t = torch.randint(0, 1000, (shape[0],), device=dev)
        ac = osch.alphas_cumprod.to(dev)
        sa, sb = ac ** 0.5, (1 - ac) ** 0.5
        got = ops.add_noise(x0, nz, t, sa, sb)
        want = osch.add_noise(x0, nz, t)
        ok &= report(f"add_noise {shape}", got, want, 0.0)
        pred = torch.randn(shape, device=dev)
        ls, dp = ops.mse_fwd_bwd(pred, nz)
        p2 = pred.clone().requires_grad_(True)
        l2 = F.mse_loss(p2, nz)
        l2.backward()
        ok &= report(f"mse loss {shape}", ls / pred.numel(), l2.detach().reshape(1), 1e-6)
        ok &= report(f"mse grad {shape}", dp, p2.grad, 1e-6)
        osch.set_timesteps(1000)
        for tt in (999, 500, 1, 0):
            z = torch.randn(shape, device=dev)
            want = osch.step(pred, torch.tensor(tt), x0, variance_noise=z)
            from polyp_image_generator_b200.scheduler import step_coefficients
            c = step_coefficients(osch.alphas_cumprod, tt, tt - 1)
            got, gx0 = ops.scheduler_step(pred, x0, z if tt > 0 else None, c["sa"], c["sb"], c["c0"], c["ct"],
                                          c["sigma"], 1.0, want_x0=True)
            ok &= report(f"step t={tt} {shape}", got, want.prev_sample, 0.0)
            ok &= report(f"step x0 t={tt} {shape}", gx0, want.pred_original_sample, 0.0)
    x = torch.randn(2, 3, 16, 16, device=dev)
    got = ops.to_uint8_nhwc(x)
    want = ((x / 2 + 0.5).clamp(0, 1).permute(0, 2, 3, 1) * 255).round().to(torch.uint8)
    ok &= report("to_uint8", got, want, 0.0)
    e1 = torch.randn(1 << 20, device=dev)
    pz = ops.scheduler_step_philox(e1 * 0, e1 * 0, 1.0, 0.0, 0.0, 0.0, 1.0, 0.0, 1234, 7)
    print(f"  philox z: mean={pz.mean().item():+.4f} std={pz.std().item():.4f} (want 0, 1)")
    ok &= abs(pz.mean().item()) < 0.01 and abs(pz.std().item() - 1) < 0.01
    return ok


def g_gn(ops):
    ok = True
    dev = "cuda"
    torch.manual_seed(1)
    cases = [(2, 16, 16, 128, 0, True), (2, 8, 8, 512, 256, True), (3, 8, 8, 256, 128, True), (2, 4, 4, 512, 0, False),
             (4, 32, 32, 256, 0, True), (2, 8, 8, 512, 512, True), (9, 64, 64, 128, 0, True), (5, 128, 128, 128, 128, True),
             (70, 16, 16, 256, 0, False), (3, 5, 7, 64, 0, True)]
    for (n, h, w, c0, c1, silu) in cases:
        C = c0 + c1
        xa = bf(torch.randn(n, h, w, c0, device=dev) * 1.5 + 0.3)
        xb = bf(torch.randn(n, h, w, c1, device=dev) - 0.2) if c1 else None
        gamma = torch.randn(C, device=dev) * 0.5 + 1
        beta = torch.randn(C, device=dev) * 0.2
        eps = 1e-5
        stats = ops.gn_stats(xa, xb, 32)
        y = ops.gn_apply(xa, xb, 32, stats, eps, gamma, beta, silu)
        xcat = torch.cat([xa, xb], -1) if c1 else xa
        xr = xcat.float().permute(0, 3, 1, 2).contiguous().requires_grad_(True)
        gr = gamma.clone().requires_grad_(True)
        br = beta.clone().requires_grad_(True)
        yr = F.group_norm(xr, 32, gr, br, eps)
        if silu:
            yr = F.silu(yr)
        ok &= report(f"gn fwd n{n} {h}x{w} c{c0}+{c1} silu={silu}", y, yr.permute(0, 2, 3, 1), 6e-3)
        stats_f, y_f = ops.gn_fwd(xa, xb, 32, eps, gamma, beta, silu)
        ok &= report("   gn fused fwd y", y_f, yr.permute(0, 2, 3, 1), 6e-3)
        ok &= report("   gn fused fwd stats", stats_f, stats, 1e-5)
        dy = bf(torch.randn(n, h, w, C, device=dev))
        add0 = bf(torch.randn(n, h, w, C, device=dev))
        dgam = torch.zeros(C, device=dev)
        dbet = torch.zeros(C, device=dev)
        dx0, dx1 = ops.gn_bwd(xa, xb, 32, stats, eps, gamma, beta, silu, dy, add0=add0, dgamma=dgam, dbeta=dbet)
        yr.backward(dy.float().permute(0, 3, 1, 2))
        dxr = xr.grad.permute(0, 2, 3, 1) + add0.float()
        got = torch.cat([dx0, dx1], -1) if c1 else dx0
        ok &= report("   gn bwd dx", got, dxr, 8e-3)
        ok &= report("   gn bwd dgamma", dgam, gr.grad, 5e-3)
        ok &= report("   gn bwd dbeta", dbet, br.grad, 5e-3)
    return ok


def g_misc(ops):
    ok = True
    dev = "cuda"
    torch.manual_seed(2)
    from oracle.unet2d import get_timestep_embedding
    t = torch.tensor([0, 1, 500, 999, 37], device=dev)
    for flip, shift in ((True, 0.0), (False, 1.0)):
        got = ops.timestep_embedding(t, 128, flip, shift)
        want = get_timestep_embedding(t.cpu(), 128, flip, shift).to(dev)
        ok &= report(f"timestep_embedding flip={flip} shift={shift}", got, want, 1e-6)
    m, k, n = 5, 128, 512
    x = torch.randn(m, k, device=dev)
    w = torch.randn(n, k, device=dev) * 0.1
    b = torch.randn(n, device=dev)
    for silu in (False, True):
        xr = x.clone().requires_grad_(True)
        wr = w.clone().requires_grad_(True)
        yr = F.linear(F.silu(xr) if silu else xr, wr, b)
        y = ops.linear_f32(x, w, b, silu)
        ok &= report(f"linear_f32 silu={silu}", y, yr, 1e-5)
        dy = torch.randn(m, n, device=dev)
        yr.backward(dy)
        dw = torch.zeros_like(w)
        db = torch.zeros_like(b)
        ops.linear_f32_wgrad(x, dy, dw, db, silu)
        ok &= report("   wgrad", dw, wr.grad, 1e-5)
        ok &= report("   bgrad", db, dy.sum(0), 1e-5)
        dx = ops.linear_f32_dgrad(dy, w, x, silu)
        ok &= report("   dgrad", dx, xr.grad, 1e-5)
    xh = bf(torch.randn(3, 8, 8, 256, device=dev))
    onc = torch.empty(3, 256, device=dev)
    oc = torch.zeros(256, device=dev)
    ops.reduce_hw(xh, onc, oc)
    ok &= report("reduce_hw nc", onc, xh.float().sum((1, 2)), 1e-5)
    ok &= report("reduce_hw c", oc, xh.float().sum((0, 1, 2)), 1e-5)
    xs = bf(torch.randn(2, 8, 8, 64, device=dev))
    s2d = ops.space_to_depth(xs)
    want = torch.stack([xs[:, ph::2, pw::2] for ph in (0, 1) for pw in (0, 1)], 0).reshape(8, 4, 4, 64)
    ok &= report("space_to_depth", s2d, want, 0.0)
    zi = ops.zero_insert2x(xs, 16, 16)
    wz = torch.zeros(2, 16, 16, 64, device=dev, dtype=torch.bfloat16)
    wz[:, ::2, ::2] = xs
    ok &= report("zero_insert2x", zi, wz, 0.0)
    up = ops.upsample2x(xs)
    ok &= report("upsample2x", up, xs.repeat_interleave(2, 1).repeat_interleave(2, 2), 0.0)
    addt = bf(torch.randn(2, 4, 4, 64, device=dev))
    sp = ops.sumpool2x(xs, addt)
    wsp = xs.float().reshape(2, 4, 2, 4, 2, 64).sum((2, 4)) + addt.float()
    ok &= report("sumpool2x", sp, wsp, 4e-3)
    wm = torch.randn(96, 9, 160, device=dev)
    wf = torch.empty(96, 9 * 160, device=dev, dtype=torch.bfloat16)
    wd = torch.empty(160, 9 * 96, device=dev, dtype=torch.bfloat16)
    ops.prep_weight(wm, wf, wd, 96, 9, 160)
    ok &= report("prep_weight wf", wf, bf(wm).reshape(96, -1), 0.0)
    ok &= report("prep_weight wd", wd, bf(wm).flip(1).permute(2, 1, 0).reshape(160, -1), 0.0)
    return ok


def g_smallconv(ops):
    """The 3-channel boundary convs as one-k-block tcgen05 GEMMs (im2col3 + conv_gemm / conv_wgrad + nhwc_to_nchw)."""
    from polyp_image_generator_b200.ops import taps_1x1, taps_3x3
    ok = True
    dev = "cuda"
    torch.manual_seed(3)
    torch.backends.cudnn.allow_tf32 = False
    for (n, h, w, C) in [(2, 16, 24, 128), (3, 64, 64, 128), (1, 7, 5, 64)]:
        grid = (n, h, w)
        x = torch.randn(n, 3, h, w, device=dev)
        wt = torch.randn(C, 3, 3, 3, device=dev) * 0.2
        b = torch.randn(C, device=dev)
        wf = torch.zeros(C, 64, device=dev, dtype=torch.bfloat16)
        wf[:, :27] = wt.permute(0, 2, 3, 1).reshape(C, 27)                      # [co][tap][ci]
        csum = torch.zeros(3, device=dev)
        pat = ops.im2col3(x, chan_sum=csum)
        xb = bf(x).float()
        want_pat = F.unfold(xb, 3, padding=1).view(n, 3, 9, h, w).permute(0, 3, 4, 2, 1).reshape(n, h, w, 27)
        ok &= report(f"im2col3 n{n} {h}x{w}", pat[..., :27], want_pat, 0.0)
        ok &= report("   im2col3 zero padding", pat[..., 27:].float().abs().sum().reshape(1), torch.zeros(1, device=dev), 0.0)
        ok &= report("   channel sums", csum, x.sum((0, 2, 3)), 1e-5)
        got = ops.conv_gemm(pat, None, taps_1x1(), wf, C, grid, bias=b)
        ok &= report("   conv_in fwd", got, F.conv2d(x, wt, b, padding=1).permute(0, 2, 3, 1), 4e-3)
        dy = bf(torch.randn(n, h, w, C, device=dev))
        R = torch.zeros(C, 64, device=dev)
        ops.conv_wgrad(dy, pat, None, taps_1x1(), R, grid)
        wr = wt.clone().requires_grad_(True)
        F.conv2d(xb, wr, None, padding=1).backward(dy.float().permute(0, 3, 1, 2))
        ok &= report("   conv_in wgrad", R[:, :27], wr.grad.permute(0, 2, 3, 1).reshape(C, 27), 2e-3)
        # conv_out: N = 3 padded to 32 output columns, fp32 NHWC result -> NCHW
        a = bf(torch.randn(n, h, w, C, device=dev))
        wo = torch.randn(3, C, 3, 3, device=dev) * 0.05
        bo = torch.randn(3, device=dev)
        wof = torch.zeros(32, 9 * C, device=dev, dtype=torch.bfloat16)
        wof[:3] = wo.permute(0, 2, 3, 1).reshape(3, 9 * C)
        b32 = torch.zeros(32, device=dev)
        b32[:3] = bo
        o32 = ops.conv_gemm(a, None, taps_3x3(C), wof, 32, grid, bias=b32, out_f32=True)
        got = ops.nhwc_to_nchw_f32(o32, 3)
        ar = a.float().permute(0, 3, 1, 2).contiguous().requires_grad_(True)
        wor = bf(wo).float().requires_grad_(True)
        yr = F.conv2d(ar, wor, bo, padding=1)
        ok &= report("   conv_out fwd", got, yr, 1e-4)
        dyo = torch.randn(n, 3, h, w, device=dev)
        yr.backward(bf(dyo).float())
        pd = ops.im2col3(dyo)
        wd = torch.zeros(C, 64, device=dev, dtype=torch.bfloat16)
        wd[:, :27] = wo.permute(0, 2, 3, 1).reshape(3, 9, C).flip(1).permute(2, 1, 0).reshape(C, 27)
        da = ops.conv_gemm(pd, None, taps_1x1(), wd, C, grid)
        ok &= report("   conv_out dgrad", da, ar.grad.permute(0, 2, 3, 1), 4e-3)
        R = torch.zeros(C, 64, device=dev)
        ops.conv_wgrad(a, pd, None, taps_1x1(), R, grid)
        dwo = R[:, :27].view(C, 9, 3).flip(1).permute(2, 1, 0)                  # [co][tap][ci]
        ok &= report("   conv_out wgrad", dwo, wor.grad.permute(0, 2, 3, 1).reshape(3, 9, C), 2e-3)
    return ok


def g_attention(ops):
    ok = True
    dev = "cuda"
    torch.manual_seed(4)
    for (b, t, heads, d) in [(2, 64, 64, 8), (3, 16, 64, 8), (1, 64, 8, 64), (2, 49, 4, 16), (2, 256, 64, 8)]:
        Cc = heads * d
        qkv = bf(torch.randn(b * t, 3 * Cc, device=dev))
        scale = d ** -0.5
        o, lse = ops.attn_fwd(qkv, b, t, heads, d, scale)
        qr = qkv.float().clone().requires_grad_(True)
        q, k, v = [z.reshape(b, t, heads, d).transpose(1, 2) for z in qr.split(Cc, 1)]
        orf = F.scaled_dot_product_attention(q, k, v, scale=scale).transpose(1, 2).reshape(b * t, Cc)
        ok &= report(f"attn fwd b{b} t{t} h{heads} d{d}", o, orf, 5e-3)
        do = bf(torch.randn(b * t, Cc, device=dev))
        orf.backward(do.float())
        dqkv = ops.attn_bwd(qkv, o, do, lse, b, t, heads, d, scale)
        ok &= report("   attn bwd", dqkv, qr.grad, 1e-2)
    return ok


def conv_ref(x, w4, bias, stride=1, pad=1):
    return F.conv2d(x.float().permute(0, 3, 1, 2), w4.float(), bias, stride=stride, padding=pad).permute(0, 2, 3, 1)


def g_gemm(ops):
    from polyp_image_generator_b200.ops import taps_1x1, taps_3x3, taps_s2d
    ok = True
    dev = "cuda"
    torch.manual_seed(5)
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    # --- probe: one-hot A, identity-ish B (diagnoses layout / descriptor errors) ---
    M, K, N = 128, 64, 128
    X = torch.zeros(1, 1, M, K, device=dev, dtype=torch.bfloat16)
    X[0, 0, 5, 3] = 1.0
    X[0, 0, 77, 40] = 2.0
    Wt = torch.zeros(N, K, device=dev, dtype=torch.bfloat16)
    for i in range(K):
        Wt[i, i] = 1.0
    out = ops.conv_gemm(X, None, taps_1x1(), Wt, N, (1, 1, M))
    torch.cuda.synchronize()
    nzi = out.float().reshape(M, N).nonzero()
    print("  probe nonzeros (want [5,3]=1, [77,40]=2):",
          [(int(r), int(c), float(out.reshape(M, N)[r, c])) for r, c in nzi[:8]], flush=True)
    # --- linear mode ---
    for (M, K, N) in [(128, 64, 128), (256, 128, 128), (4096, 512, 512), (4096, 512, 1536), (300, 192, 128),
                      (64, 512, 256)]:
        x = bf(torch.randn(1, 1, M, K, device=dev))
        w = bf(torch.randn(N, K, device=dev) * 0.05)
        b = torch.randn(N, device=dev)
        out = ops.conv_gemm(x, None, taps_1x1(), w, N, (1, 1, M), bias=b)
        want = x.float().reshape(M, K) @ w.float().t() + b
        ok &= report(f"linear M{M} K{K} N{N}", out.reshape(M, N), want, 4e-3)
    os.environ["DDPM_BLOCK_N"] = "256"
    x = bf(torch.randn(1, 1, 1024, 256, device=dev))
    w = bf(torch.randn(512, 256, device=dev) * 0.05)
    out = ops.conv_gemm(x, None, taps_1x1(), w, 512, (1, 1, 1024))
    ok &= report("linear BLOCK_N=256", out.reshape(1024, 512), x.float().reshape(1024, 256) @ w.float().t(), 4e-3)
    os.environ.pop("DDPM_BLOCK_N")
    # fp32 output
    out = ops.conv_gemm(x, None, taps_1x1(), w, 512, (1, 1, 1024), out_f32=True)
    ok &= report("linear fp32 out", out.reshape(1024, 512), x.float().reshape(1024, 256) @ w.float().t(), 1e-5)
    # --- 3x3 convs over the UNet's resolutions ---
    for (n, h, w_, cin, cout) in [(2, 16, 16, 128, 128), (1, 128, 128, 128, 128), (3, 32, 32, 256, 256),
                                  (4, 8, 8, 512, 512), (9, 4, 4, 512, 512), (2, 64, 64, 128, 256),
                                  (2, 14, 14, 128, 128), (3, 7, 7, 64, 128), (1, 200, 136, 64, 128),
                                  (3, 70, 96, 128, 256), (5, 64, 64, 256, 128), (2, 128, 128, 256, 128),
                                  (1, 65, 64, 64, 64), (2, 3, 200, 64, 128), (32, 4, 4, 1024, 512), (16, 8, 8, 512, 1024)]:
        x = bf(torch.randn(n, h, w_, cin, device=dev))
        w4 = bf(torch.randn(cout, cin, 3, 3, device=dev) * 0.03)
        wk = w4.permute(0, 2, 3, 1).reshape(cout, 9 * cin).contiguous()
        b = torch.randn(cout, device=dev)
        temb = torch.randn(n, cout, device=dev)
        res = bf(torch.randn(n, h, w_, cout, device=dev))
        out = ops.conv_gemm(x, None, taps_3x3(cin), wk, cout, (n, h, w_), bias=b, temb=temb, res=res)
        want = conv_ref(x, w4, b) + temb[:, None, None, :] + res.float()
        ok &= report(f"conv3x3 n{n} {h}x{w_} {cin}->{cout} +bias+temb+res", out, want, 4e-3)
    # --- halo kernel with two concatenated sources at high resolution ---
    n, h, w_, c0, c1, cout = 2, 64, 64, 128, 64, 128
    xa, xb = bf(torch.randn(n, h, w_, c0, device=dev)), bf(torch.randn(n, h, w_, c1, device=dev))
    w4 = bf(torch.randn(cout, c0 + c1, 3, 3, device=dev) * 0.03)
    wk = w4.permute(0, 2, 3, 1).reshape(cout, -1).contiguous()
    out = ops.conv_gemm(xa, xb, taps_3x3(c0 + c1), wk, cout, (n, h, w_))
    ok &= report("conv3x3 concat 128+64 -> 128 @64x64 (halo)", out, conv_ref(torch.cat([xa, xb], -1), w4, None), 4e-3)
    # --- concat of two sources + 1x1 ---
    n, h, w_, c0, c1, cout = 2, 16, 16, 256, 128, 256
    xa, xb = bf(torch.randn(n, h, w_, c0, device=dev)), bf(torch.randn(n, h, w_, c1, device=dev))
    w4 = bf(torch.randn(cout, c0 + c1, 3, 3, device=dev) * 0.03)
    wk = w4.permute(0, 2, 3, 1).reshape(cout, -1).contiguous()
    out = ops.conv_gemm(xa, xb, taps_3x3(c0 + c1), wk, cout, (n, h, w_))
    ok &= report("conv3x3 concat 256+128 -> 256", out, conv_ref(torch.cat([xa, xb], -1), w4, None), 4e-3)
    w1 = bf(torch.randn(cout, c0 + c1, device=dev) * 0.05)
    out = ops.conv_gemm(xa, xb, taps_1x1(), w1, cout, (n, h, w_))
    ok &= report("conv1x1 concat", out, torch.cat([xa, xb], -1).float() @ w1.float().t(), 4e-3)
    # channel-slice input / output views
    big = bf(torch.randn(n, h, w_, 512, device=dev))
    outbig = torch.zeros(n, h, w_, 512, device=dev, dtype=torch.bfloat16)
    ops.conv_gemm(big[..., 128:384], None, taps_1x1(), w1[:, :256].contiguous(), cout, (n, h, w_), out=outbig[..., 256:])
    ok &= report("conv1x1 sliced views", outbig[..., 256:], big[..., 128:384].float() @ w1[:, :256].float().t(), 4e-3)
    # --- stride 2 via space-to-depth, both paddings ---
    for pad in (1, 0):
        n, h, w_, c = 2, 16, 16, 128
        x = bf(torch.randn(n, h, w_, c, device=dev))
        w4 = bf(torch.randn(c, c, 3, 3, device=dev) * 0.03)
        wk = w4.permute(0, 2, 3, 1).reshape(c, -1).contiguous()
        s2d = ops.space_to_depth(x)
        out = ops.conv_gemm(s2d, None, taps_s2d(c, n, pad), wk, c, (n, h // 2, w_ // 2), src_n=4 * n)
        xin = x.float().permute(0, 3, 1, 2)
        if pad == 0:
            xin = F.pad(xin, (0, 1, 0, 1))
        want = F.conv2d(xin, w4.float(), None, stride=2, padding=pad).permute(0, 2, 3, 1)
        ok &= report(f"conv3x3 stride2 pad={pad}", out, want, 4e-3)
    # --- dgrad through prep_weight ---
    n, h, w_, cin, cout = 2, 16, 16, 128, 256
    wm = torch.randn(cout, 9, cin, device=dev) * 0.03
    wd = torch.empty(cin, 9 * cout, device=dev, dtype=torch.bfloat16)
    ops.prep_weight(wm, None, wd, cout, 9, cin)
    dy = bf(torch.randn(n, h, w_, cout, device=dev))
    dx = ops.conv_gemm(dy, None, taps_3x3(cout), wd, cin, (n, h, w_))
    w4 = bf(wm).float().reshape(cout, 3, 3, cin).permute(0, 3, 1, 2)
    xr = torch.zeros(n, cin, h, w_, device=dev, requires_grad=True)
    F.conv2d(xr, w4, None, padding=1).backward(dy.float().permute(0, 3, 1, 2))
    ok &= report("conv3x3 dgrad", dx, xr.grad.permute(0, 2, 3, 1), 4e-3)
    return ok


def g_wgrad(ops):
    from polyp_image_generator_b200.ops import taps_1x1, taps_3x3, taps_s2d
    ok = True
    dev = "cuda"
    torch.manual_seed(6)
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    # probe
    P, Co, Ci = 128, 128, 128
    dy = torch.zeros(1, 1, P, Co, device=dev, dtype=torch.bfloat16)
    x = torch.zeros(1, 1, P, Ci, device=dev, dtype=torch.bfloat16)
    dy[0, 0, 9, 70] = 1.0
    x[0, 0, 9, 33] = 3.0
    dy[0, 0, 100, 2] = 1.0
    x[0, 0, 100, 127] = 5.0
    dw = torch.zeros(Co, Ci, device=dev)
    ops.conv_wgrad(dy, x, None, taps_1x1(), dw, (1, 1, P), accumulate=False)
    torch.cuda.synchronize()
    nzi = dw.nonzero()
    print("  probe nonzeros (want [70,33]=3, [2,127]=5):", [(int(r), int(c), float(dw[r, c])) for r, c in nzi[:8]],
          flush=True)
    for (M, Co, Ci, splits) in [(128, 128, 128, 1), (1024, 128, 128, 1), (4096, 512, 512, 0), (4096, 256, 192, 4),
                                (300, 128, 64, 1), (100, 128, 128, 1)]:
        dy = bf(torch.randn(1, 1, M, Co, device=dev))
        x = bf(torch.randn(1, 1, M, Ci, device=dev))
        dw = torch.zeros(Co, Ci, device=dev)
        ops.conv_wgrad(dy, x, None, taps_1x1(), dw, (1, 1, M), accumulate=False, splits=splits)
        want = dy.float().reshape(M, Co).t() @ x.float().reshape(M, Ci)
        ok &= report(f"linear wgrad M{M} Co{Co} Ci{Ci} splits={splits}", dw, want, 2e-3)
    for (n, h, w_, cin, cout) in [(2, 16, 16, 128, 128), (1, 64, 64, 128, 256), (4, 8, 8, 512, 512),
                                  (9, 4, 4, 256, 512), (2, 14, 14, 128, 128), (3, 7, 7, 64, 128),
                                  # row-resident kernel (W >= 64, W % 16 == 0, channels % 128 == 0)
                                  (2, 128, 128, 128, 128), (1, 64, 64, 256, 128), (1, 256, 256, 128, 128),
                                  (1, 80, 80, 128, 128), (3, 5, 64, 128, 256), (1, 64, 96, 384, 128), (3, 32, 32, 256, 256),
                                  (2, 16, 16, 512, 256), (2, 48, 48, 128, 128), (2, 6, 32, 128, 128)]:
        x = bf(torch.randn(n, h, w_, cin, device=dev))
        dy = bf(torch.randn(n, h, w_, cout, device=dev))
        dw = torch.zeros(cout, 9 * cin, device=dev)
        db = torch.zeros(cout, device=dev)
        ops.conv_wgrad(dy, x, None, taps_3x3(cin), dw, (n, h, w_), accumulate=True, dbias=db)
        wr = torch.zeros(cout, cin, 3, 3, device=dev, requires_grad=True)
        F.conv2d(x.float().permute(0, 3, 1, 2), wr, None, padding=1).backward(dy.float().permute(0, 3, 1, 2))
        ok &= report(f"conv3x3 wgrad n{n} {h}x{w_} {cin}->{cout}", dw, wr.grad.permute(0, 2, 3, 1).reshape(cout, -1),
                     2e-3)
        ok &= report("   bias grad", db, dy.float().sum((0, 1, 2)), 2e-3)
    # concat
    n, h, w_, c0, c1, cout = 2, 16, 16, 256, 128, 128
    xa, xb = bf(torch.randn(n, h, w_, c0, device=dev)), bf(torch.randn(n, h, w_, c1, device=dev))
    dy = bf(torch.randn(n, h, w_, cout, device=dev))
    dw = torch.zeros(cout, 9 * (c0 + c1), device=dev)
    ops.conv_wgrad(dy, xa, xb, taps_3x3(c0 + c1), dw, (n, h, w_))
    wr = torch.zeros(cout, c0 + c1, 3, 3, device=dev, requires_grad=True)
    F.conv2d(torch.cat([xa, xb], -1).float().permute(0, 3, 1, 2), wr, None, padding=1).backward(
        dy.float().permute(0, 3, 1, 2))
    ok &= report("conv3x3 wgrad concat", dw, wr.grad.permute(0, 2, 3, 1).reshape(cout, -1), 2e-3)
    n, h, w_, c0, c1, cout = 2, 64, 64, 128, 128, 128     # row-resident kernel, two sources
    xa, xb = bf(torch.randn(n, h, w_, c0, device=dev)), bf(torch.randn(n, h, w_, c1, device=dev))
    dy = bf(torch.randn(n, h, w_, cout, device=dev))
    dw = torch.zeros(cout, 9 * (c0 + c1), device=dev)
    ops.conv_wgrad(dy, xa, xb, taps_3x3(c0 + c1), dw, (n, h, w_))
    wr = torch.zeros(cout, c0 + c1, 3, 3, device=dev, requires_grad=True)
    F.conv2d(torch.cat([xa, xb], -1).float().permute(0, 3, 1, 2), wr, None, padding=1).backward(
        dy.float().permute(0, 3, 1, 2))
    ok &= report("conv3x3 wgrad concat 64x64 (row kernel)", dw, wr.grad.permute(0, 2, 3, 1).reshape(cout, -1), 2e-3)
    # stride 2
    for pad in (1, 0):
        n, h, w_, c = 2, 16, 16, 128
        x = bf(torch.randn(n, h, w_, c, device=dev))
        dy = bf(torch.randn(n, h // 2, w_ // 2, c, device=dev))
        s2d = ops.space_to_depth(x)
        dw = torch.zeros(c, 9 * c, device=dev)
        ops.conv_wgrad(dy, s2d, None, taps_s2d(c, n, pad), dw, (n, h // 2, w_ // 2), src_n=4 * n)
        wr = torch.zeros(c, c, 3, 3, device=dev, requires_grad=True)
        xin = x.float().permute(0, 3, 1, 2)
        if pad == 0:
            xin = F.pad(xin, (0, 1, 0, 1))
        F.conv2d(xin, wr, None, stride=2, padding=pad).backward(dy.float().permute(0, 3, 1, 2))
        ok &= report(f"conv3x3 stride2 wgrad pad={pad}", dw, wr.grad.permute(0, 2, 3, 1).reshape(c, -1), 2e-3)
    return ok


def g_gnfuse(ops):
    """GroupNorm backward split: first half in the dgrad conv epilogue (dz + per-(n,c) sums), second half streaming."""
    from polyp_image_generator_b200.ops import taps_3x3
    ok = True
    dev = "cuda"
    torch.manual_seed(11)
    #        n, h,   w,   c0,  c1,  cin_of_dgrad
    cases = [(2, 128, 128, 128, 0, 128), (3, 32, 32, 256, 0, 256), (4, 8, 8, 512, 0, 512), (2, 64, 64, 128, 128, 128),
             (5, 16, 16, 256, 128, 256), (1, 64, 64, 128, 0, 256)]
    for (n, h, w, c0, c1, cg) in cases:
        C = c0 + c1
        grid = (n, h, w)
        xa = bf(torch.randn(n, h, w, c0, device=dev) * 1.3 + 0.2)
        xb = bf(torch.randn(n, h, w, c1, device=dev) - 0.1) if c1 else None
        gamma = torch.randn(C, device=dev) * 0.5 + 1
        beta = torch.randn(C, device=dev) * 0.2
        eps = 1e-5
        stats, _, coef = ops.gn_fwd(xa, xb, 32, eps, gamma, beta, True, want_coef=True)
        g = bf(torch.randn(n, h, w, cg, device=dev))
        wd = bf(torch.randn(C, 9 * cg, device=dev) * 0.05)       # dgrad operand: [C (= conv input chans), 9*cg]
        add0 = bf(torch.randn(n, h, w, C, device=dev))
        # unfused reference path through the same kernels
        d_y = ops.conv_gemm(g, None, taps_3x3(cg), wd, C, grid)
        dg_r, db_r = torch.zeros(C, device=dev), torch.zeros(C, device=dev)
        r0, r1 = ops.gn_bwd(xa, xb, 32, stats, eps, gamma, beta, True, d_y, add0=add0, dgamma=dg_r, dbeta=db_r)
        # fused path
        sums = torch.zeros(n, C, 2, device=dev)
        dz = ops.conv_gemm(g, None, taps_3x3(cg), wd, C, grid, gn=(xa, xb, coef, True, sums))
        dg_f, db_f = torch.zeros(C, device=dev), torch.zeros(C, device=dev)
        o_nc, o_c = torch.zeros(n, C, device=dev), torch.zeros(C, device=dev)
        f0, f1 = ops.gn_bwd_apply(xa, xb, 32, stats, eps, gamma, dz, sums, add0=add0, dgamma=dg_f, dbeta=db_f,
                                  out_nc=o_nc, out_c=o_c)
        # fp32 autograd reference of the GroupNorm+SiLU backward on the bf16 dy the conv produced
        xcat = torch.cat([xa, xb], -1) if c1 else xa
        xr = xcat.float().permute(0, 3, 1, 2).contiguous().requires_grad_(True)
        gr, br = gamma.clone().requires_grad_(True), beta.clone().requires_grad_(True)
        F.silu(F.group_norm(xr, 32, gr, br, eps)).backward(d_y.float().permute(0, 3, 1, 2))
        want = xr.grad.permute(0, 2, 3, 1) + add0.float()
        got = torch.cat([f0, f1], -1) if c1 else f0
        ref_k = torch.cat([r0, r1], -1) if c1 else r0
        tag = f"gnfuse n{n} {h}x{w} c{c0}+{c1} <- {cg}"
        ok &= report(tag + " dx vs autograd", got, want, 1e-2)
        ok &= report("   dx vs unfused kernels", got, ref_k, 1e-2)
        ok &= report("   dgamma", dg_f, gr.grad, 1e-2)
        ok &= report("   dbeta", db_f, br.grad, 1e-2)
        ok &= report("   fused pixel sums [n, c]", o_nc, want.sum((1, 2)), 1e-2)
        ok &= report("   fused pixel sums [c]", o_c, want.sum((0, 1, 2)), 1e-2)
    return ok


def run_group(name):
    from polyp_image_generator_b200 import ops as ops_mod
    ops = ops_mod.get()
    t0 = time.time()
    try:
        ok = globals()["g_" + name](ops)
        torch.cuda.synchronize()
    except Exception:
        traceback.print_exc()
        ok = False
    print(f"GROUP {name}: {'PASS' if ok else 'FAIL'} ({time.time() - t0:.1f}s, {ops.launches} launches)", flush=True)
    return ok


if __name__ == "__main__":
    if len(sys.argv) > 1:
        sys.exit(0 if run_group(sys.argv[1]) else 1)
    results = {}
    for g in GROUPS:
        print(f"=== {g} ===", flush=True)
        try:
            r = subprocess.run([sys.executable, os.path.abspath(__file__), g], timeout=300)
            results[g] = r.returncode == 0
        except subprocess.TimeoutExpired:
            print(f"GROUP {g}: TIMEOUT", flush=True)
            results[g] = False
    print("SUMMARY", results, flush=True)
    sys.exit(0 if all(results.values()) else 1)
