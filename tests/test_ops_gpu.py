"""Op-level GPU parity tests (-m gpu): every C-ABI kernel group against a plain PyTorch fp32 statement of the same op on
the same bf16 inputs (torch on the same device is the checker here, never the product path).

These are the checks that used to live in the bring-up script tests/gpu_bringup.py, now collected by pytest, plus the
geometries the benchmark and BASELINE configs[1..4] actually run: the halo-resident conv at W = 128 / 224 / 256 (CTA
pair and single CTA), the row-resident wgrad in both segment shapes (hb = 1 for W >= 128, hb > 1 below), every conv
shape class of SURVEY.md Appendix D, GroupNorm forward / backward, head_dim-8 attention and the uint8 epilogue.
"""
import os

import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu

DEV = "cuda"


def rel(a, b):
    a, b = a.float(), b.float()
    return ((a - b).norm() / (b.norm() + 1e-30)).item()


class Check:
    """Collects every comparison of a test (all are printed with -s) and fails once at the end."""

    def __init__(self):
        self.bad = []

    def __call__(self, name, got, want, tol):
        r = rel(got, want)
        mx = (got.float() - want.float()).abs().max().item() if got.numel() else 0.0
        ok = r <= tol and bool(torch.isfinite(got.float()).all())
        print(f"  [{'OK ' if ok else 'BAD'}] {name:62s} rel={r:.3e} maxabs={mx:.3e} (tol {tol:g})", flush=True)
        if not ok:
            self.bad.append((name, r, tol))
        return ok

    def done(self):
        assert not self.bad, self.bad


def bf(x):
    return x.to(torch.bfloat16)


@pytest.fixture(scope="module")
def ops():
    assert torch.cuda.is_available(), "GPU tests need a CUDA device"
    from polyp_image_generator_b200 import ops as ops_mod
    o = ops_mod.get()
    assert o.name == "cuda"
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    return o


class env:
    """Temporarily set launcher knobs (DDPM_HALO_PAIR=0 ...) read by the C library through getenv."""

    def __init__(self, **kv):
        self.kv, self.old = kv, {}

    def __enter__(self):
        for k, v in self.kv.items():
            self.old[k] = os.environ.get(k)
            os.environ[k] = str(v)

    def __exit__(self, *a):
        for k, v in self.old.items():
            if v is None:
                os.environ.pop(k, None)
            else:
                os.environ[k] = v


def conv_ref(x, w4, bias, stride=1, pad=1):
    return F.conv2d(x.float().permute(0, 3, 1, 2), w4.float(), bias, stride=stride, padding=pad).permute(0, 2, 3, 1)


# ---- elementwise ------------------------------------------------------------------------------------------------
def test_to_uint8_nhwc_values(ops):
    """DDPMPipeline's (x/2 + 0.5).clamp(0, 1) -> (x * 255).round() -> uint8 NHWC: VALUES, incl. the
    clamp edges, exact .5 ties (numpy rounds half to even) and ragged sizes."""
    torch.manual_seed(0)
    for shape in [(2, 3, 16, 16), (1, 3, 7, 5), (3, 3, 128, 128), (1, 4, 9, 3)]:
        x = torch.randn(shape, device=DEV) * 1.2
        x.view(-1)[:6] = torch.tensor([-1.0, 1.0, -3.0, 3.0, 0.0, 1.0 / 255.0], device=DEV)
        # values that land exactly on k + 0.5 after *255 (x/2 + 0.5 = (k + 0.5)/255)
        k = torch.arange(0, 40, device=DEV, dtype=torch.float32)
        n = min(k.numel(), x.numel() - 6)
        x.view(-1)[6:6 + n] = (((k[:n] + 0.5) / 255.0) - 0.5) * 2.0
        got = ops.to_uint8_nhwc(x)
        img = (x / 2 + 0.5).clamp(0, 1).cpu().permute(0, 2, 3, 1).numpy()
        want = torch.from_numpy((img * 255).round().astype("uint8"))
        assert got.dtype == torch.uint8 and got.shape == want.shape
        diff = (got.cpu().int() - want.int()).abs()
        assert int(diff.max()) == 0, (shape, int(diff.max()), int((diff > 0).sum()))


def test_elementwise_kernels_vs_oracle(ops):
    import oracle
    from polyp_image_generator_b200.scheduler import step_coefficients
    ck = Check()
    torch.manual_seed(0)
    osch = oracle.DDPMScheduler()
    for shape in [(4, 3, 64, 64), (3, 3, 7, 5), (64, 3, 128, 128), (2, 3, 224, 224)]:
        x0, nz = torch.randn(shape), torch.randn(shape)
        t = torch.randint(0, 1000, (shape[0],))
        ac = osch.alphas_cumprod
        got = ops.add_noise(x0.to(DEV), nz.to(DEV), t.to(DEV), (ac ** 0.5).to(DEV), ((1 - ac) ** 0.5).to(DEV))
        ck(f"add_noise {shape}", got.cpu(), osch.add_noise(x0, nz, t), 0.0)
        pred = torch.randn(shape)
        ls, dp = ops.mse_fwd_bwd(pred.to(DEV), nz.to(DEV))
        p2 = pred.clone().requires_grad_(True)
        l2 = F.mse_loss(p2, nz)
        l2.backward()
        ck(f"mse loss {shape}", (ls / pred.numel()).cpu(), l2.detach().reshape(1), 1e-6)
        ck(f"mse grad {shape}", dp.cpu(), p2.grad, 1e-6)
        osch.set_timesteps(1000)
        for tt in (999, 500, 1, 0):
            z = torch.randn(shape)
            want = osch.step(pred, torch.tensor(tt), x0, variance_noise=z)
            c = step_coefficients(osch.alphas_cumprod, tt, tt - 1)
            got, gx0 = ops.scheduler_step(pred.to(DEV), x0.to(DEV), z.to(DEV) if tt > 0 else None, c["sa"], c["sb"],
                                          c["c0"], c["ct"], c["sigma"], 1.0, want_x0=True)
            ck(f"step t={tt} {shape}", got.cpu(), want.prev_sample, 0.0)
            ck(f"step x0 t={tt} {shape}", gx0.cpu(), want.pred_original_sample, 0.0)
    ck.done()


# ---- GroupNorm --------------------------------------------------------------------------------------------------
GN_CASES = [(2, 16, 16, 128, 0, True), (2, 8, 8, 512, 256, True), (3, 8, 8, 256, 128, True), (2, 4, 4, 512, 0, False),
            (4, 32, 32, 256, 0, True), (2, 8, 8, 512, 512, True), (9, 64, 64, 128, 0, True), (5, 128, 128, 128, 128, True),
            (70, 16, 16, 256, 0, False), (3, 5, 7, 64, 0, True), (2, 224, 224, 128, 0, True), (2, 7, 7, 512, 512, True),
            (1, 256, 256, 128, 128, True), (2, 14, 14, 512, 0, False), (2, 112, 112, 256, 128, True)]


@pytest.mark.parametrize("n,h,w,c0,c1,silu", GN_CASES)
def test_groupnorm_fwd_bwd_vs_torch(ops, n, h, w, c0, c1, silu):
    """gn_stats + gn_apply, the fused gn_fwd (team / solo kernels) and gn_bwd against F.group_norm (+ SiLU) autograd in
    fp32 on the same bf16 input, two-source channel concat included."""
    ck = Check()
    torch.manual_seed(1)
    C = c0 + c1
    xa = bf(torch.randn(n, h, w, c0, device=DEV) * 1.5 + 0.3)
    xb = bf(torch.randn(n, h, w, c1, device=DEV) - 0.2) if c1 else None
    gamma = torch.randn(C, device=DEV) * 0.5 + 1
    beta = torch.randn(C, device=DEV) * 0.2
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
    ck("gn_stats + gn_apply", y, yr.permute(0, 2, 3, 1), 6e-3)
    stats_f, y_f = ops.gn_fwd(xa, xb, 32, eps, gamma, beta, silu)
    ck("gn_fwd y", y_f, yr.permute(0, 2, 3, 1), 6e-3)
    ck("gn_fwd stats", stats_f, stats, 1e-5)
    # statistics against fp64 moments of the same bf16 values
    xg = xcat.double().reshape(n, h * w, 32, C // 32)
    want_stats = torch.stack([xg.sum((1, 3)), (xg * xg).sum((1, 3))], -1)
    ck("gn_fwd stats vs fp64", stats_f.double(), want_stats, 1e-5)
    dy = bf(torch.randn(n, h, w, C, device=DEV))
    add0 = bf(torch.randn(n, h, w, C, device=DEV))
    dgam = torch.zeros(C, device=DEV)
    dbet = torch.zeros(C, device=DEV)
    dx0, dx1 = ops.gn_bwd(xa, xb, 32, stats, eps, gamma, beta, silu, dy, add0=add0, dgamma=dgam, dbeta=dbet)
    yr.backward(dy.float().permute(0, 3, 1, 2))
    dxr = xr.grad.permute(0, 2, 3, 1) + add0.float()
    got = torch.cat([dx0, dx1], -1) if c1 else dx0
    ck("gn_bwd dx", got, dxr, 8e-3)
    ck("gn_bwd dgamma", dgam, gr.grad, 5e-3)
    ck("gn_bwd dbeta", dbet, br.grad, 5e-3)
    ck.done()


GNFUSE_CASES = [(2, 128, 128, 128, 0, 128), (3, 32, 32, 256, 0, 256), (4, 8, 8, 512, 0, 512), (2, 64, 64, 128, 128, 128),
                (5, 16, 16, 256, 128, 256), (1, 64, 64, 128, 0, 256), (1, 256, 256, 128, 0, 128),
                (1, 224, 224, 128, 128, 128), (2, 112, 112, 256, 0, 128)]


@pytest.mark.parametrize("n,h,w,c0,c1,cg", GNFUSE_CASES)
def test_groupnorm_backward_fused_in_dgrad_epilogue(ops, n, h, w, c0, c1, cg):
    """GroupNorm backward split: first half in the dgrad conv epilogue (dz + per-(n, c) sums), second half streaming
    (gn_bwd_apply), against fp32 autograd of SiLU(GroupNorm(x)) and against the unfused kernels."""
    from polyp_image_generator_b200.ops import taps_3x3
    ck = Check()
    torch.manual_seed(11)
    C = c0 + c1
    grid = (n, h, w)
    xa = bf(torch.randn(n, h, w, c0, device=DEV) * 1.3 + 0.2)
    xb = bf(torch.randn(n, h, w, c1, device=DEV) - 0.1) if c1 else None
    gamma = torch.randn(C, device=DEV) * 0.5 + 1
    beta = torch.randn(C, device=DEV) * 0.2
    eps = 1e-5
    stats, _, coef = ops.gn_fwd(xa, xb, 32, eps, gamma, beta, True, want_coef=True)
    g = bf(torch.randn(n, h, w, cg, device=DEV))
    wd = bf(torch.randn(C, 9 * cg, device=DEV) * 0.05)       # dgrad operand: [C (= conv input chans), 9*cg]
    add0 = bf(torch.randn(n, h, w, C, device=DEV))
    d_y = ops.conv_gemm(g, None, taps_3x3(cg), wd, C, grid)
    dg_r, db_r = torch.zeros(C, device=DEV), torch.zeros(C, device=DEV)
    r0, r1 = ops.gn_bwd(xa, xb, 32, stats, eps, gamma, beta, True, d_y, add0=add0, dgamma=dg_r, dbeta=db_r)
    sums = torch.zeros(n, C, 2, device=DEV)
    dz = ops.conv_gemm(g, None, taps_3x3(cg), wd, C, grid, gn=(xa, xb, coef, True, sums))
    dg_f, db_f = torch.zeros(C, device=DEV), torch.zeros(C, device=DEV)
    o_nc, o_c = torch.zeros(n, C, device=DEV), torch.zeros(C, device=DEV)
    f0, f1 = ops.gn_bwd_apply(xa, xb, 32, stats, eps, gamma, dz, sums, add0=add0, dgamma=dg_f, dbeta=db_f,
                              out_nc=o_nc, out_c=o_c)
    xcat = torch.cat([xa, xb], -1) if c1 else xa
    xr = xcat.float().permute(0, 3, 1, 2).contiguous().requires_grad_(True)
    gr, br = gamma.clone().requires_grad_(True), beta.clone().requires_grad_(True)
    F.silu(F.group_norm(xr, 32, gr, br, eps)).backward(d_y.float().permute(0, 3, 1, 2))
    want = xr.grad.permute(0, 2, 3, 1) + add0.float()
    got = torch.cat([f0, f1], -1) if c1 else f0
    ref_k = torch.cat([r0, r1], -1) if c1 else r0
    ck("dx vs autograd", got, want, 1e-2)
    ck("dx vs unfused kernels", got, ref_k, 1e-2)
    ck("dgamma", dg_f, gr.grad, 1e-2)
    ck("dbeta", db_f, br.grad, 1e-2)
    ck("fused pixel sums [n, c]", o_nc, want.sum((1, 2)), 1e-2)
    ck("fused pixel sums [c]", o_c, want.sum((0, 1, 2)), 1e-2)
    ck.done()


# ---- time embedding / layout helpers ------------------------------------------------------------------------------
def test_time_embedding_linears_and_layout_helpers(ops):
    from oracle.unet2d import get_timestep_embedding
    ck = Check()
    torch.manual_seed(2)
    t = torch.tensor([0, 1, 500, 999, 37], device=DEV)
    for flip, shift in ((True, 0.0), (False, 1.0)):
        got = ops.timestep_embedding(t, 128, flip, shift)
        want = get_timestep_embedding(t.cpu(), 128, flip, shift).to(DEV)
        ck(f"timestep_embedding flip={flip} shift={shift}", got, want, 1e-6)
    m, k, n = 5, 128, 512
    x = torch.randn(m, k, device=DEV)
    w = torch.randn(n, k, device=DEV) * 0.1
    b = torch.randn(n, device=DEV)
    for silu in (False, True):
        xr = x.clone().requires_grad_(True)
        wr = w.clone().requires_grad_(True)
        yr = F.linear(F.silu(xr) if silu else xr, wr, b)
        ck(f"linear_f32 silu={silu}", ops.linear_f32(x, w, b, silu), yr, 1e-5)
        dy = torch.randn(m, n, device=DEV)
        yr.backward(dy)
        dw = torch.zeros_like(w)
        db = torch.zeros_like(b)
        ops.linear_f32_wgrad(x, dy, dw, db, silu)
        ck("   wgrad", dw, wr.grad, 1e-5)
        ck("   bgrad", db, dy.sum(0), 1e-5)
        ck("   dgrad", ops.linear_f32_dgrad(dy, w, x, silu), xr.grad, 1e-5)
    xh = bf(torch.randn(3, 8, 8, 256, device=DEV))
    onc = torch.empty(3, 256, device=DEV)
    oc = torch.zeros(256, device=DEV)
    ops.reduce_hw(xh, onc, oc)
    ck("reduce_hw nc", onc, xh.float().sum((1, 2)), 1e-5)
    ck("reduce_hw c", oc, xh.float().sum((0, 1, 2)), 1e-5)
    xs = bf(torch.randn(2, 8, 8, 64, device=DEV))
    want = torch.stack([xs[:, ph::2, pw::2] for ph in (0, 1) for pw in (0, 1)], 0).reshape(8, 4, 4, 64)
    ck("space_to_depth", ops.space_to_depth(xs), want, 0.0)
    wz = torch.zeros(2, 16, 16, 64, device=DEV, dtype=torch.bfloat16)
    wz[:, ::2, ::2] = xs
    ck("zero_insert2x", ops.zero_insert2x(xs, 16, 16), wz, 0.0)
    ck("upsample2x", ops.upsample2x(xs), xs.repeat_interleave(2, 1).repeat_interleave(2, 2), 0.0)
    addt = bf(torch.randn(2, 4, 4, 64, device=DEV))
    wsp = xs.float().reshape(2, 4, 2, 4, 2, 64).sum((2, 4)) + addt.float()
    ck("sumpool2x", ops.sumpool2x(xs, addt), wsp, 4e-3)
    wm = torch.randn(96, 9, 160, device=DEV)
    wf = torch.empty(96, 9 * 160, device=DEV, dtype=torch.bfloat16)
    wd = torch.empty(160, 9 * 96, device=DEV, dtype=torch.bfloat16)
    ops.prep_weight(wm, wf, wd, 96, 9, 160)
    ck("prep_weight wf", wf, bf(wm).reshape(96, -1), 0.0)
    ck("prep_weight wd", wd, bf(wm).flip(1).permute(2, 1, 0).reshape(160, -1), 0.0)
    ck.done()


# ---- time-embedding linears at the sizes of the step (tiled fp32 SIMT GEMM, split reduction in the input gradient) ----
@pytest.mark.parametrize("m,k,n", [(64, 512, 9984), (32, 512, 9984), (64, 128, 512), (64, 512, 512), (70, 100, 333),
                                   (1, 8, 16), (130, 40, 65), (512, 8, 512)])
def test_time_embedding_linears_at_step_sizes(ops, m, k, n):
    ck = Check()
    torch.manual_seed(m * 7 + n)
    x = torch.randn(m, k, device=DEV)
    w = torch.randn(n, k, device=DEV) / k ** 0.5
    b = torch.randn(n, device=DEV)
    for silu in (False, True):
        xr = x.double().requires_grad_(True)
        wr = w.double().requires_grad_(True)
        yr = F.linear(F.silu(xr) if silu else xr, wr, b.double())
        ck(f"linear_f32 {m}x{k}->{n} silu={silu}", ops.linear_f32(x, w, b, silu), yr, 1e-5)
        ck("   no bias", ops.linear_f32(x, w, None, silu), yr - b.double(), 1e-5)
        dy = torch.randn(m, n, device=DEV)
        yr.backward(dy.double())
        dw = torch.zeros_like(w)
        db = torch.zeros_like(b)
        ops.linear_f32_wgrad(x, dy, dw, db, silu)
        ck("   wgrad", dw, wr.grad, 1e-5)
        ck("   bgrad", db, dy.double().sum(0), 1e-5)
        ops.linear_f32_wgrad(x, dy, dw, db, silu)          # accumulates
        ck("   wgrad accumulates", dw, 2 * wr.grad, 1e-5)
        ck("   dgrad", ops.linear_f32_dgrad(dy, w, x, silu), xr.grad, 1e-5)
    ck.done()


def test_batched_weight_preparation_pairs_and_scalar_paths(ops):
    """ddpm_prep_weights_batched over a table mixing even (bf16x2 path) and odd (scalar path) layer shapes, ragged
    against the 64x64 tile, against the per-layer definition: wf[co][tap][ci], wd[ci][T-1-tap][co], both rounded once."""
    ck = Check()
    torch.manual_seed(5)
    shapes = [(128, 9, 128), (96, 9, 160), (512, 1, 1024), (70, 9, 66), (33, 3, 65), (64, 9, 64), (256, 9, 384)]
    entries, keep = [], []
    for cout, taps, cin in shapes:
        wm = torch.randn(cout, taps, cin, device=DEV)
        wf = torch.full((cout, taps * cin + 64), 7.0, device=DEV, dtype=torch.bfloat16)     # wider row: ldwf > K
        wd = torch.full((cin, taps * cout), 7.0, device=DEV, dtype=torch.bfloat16)
        entries.append((wm.reshape(-1), wf[:, :taps * cin], wd, cout, taps, cin))
        keep.append((wm, wf, wd))
    table = ops.build_prep_table(entries, torch.device(DEV))
    for with_d in (True, False):
        for _, wf, wd in keep:
            wf.fill_(7.0)
            wd.fill_(7.0)
        ops.prep_weights_batched(*table, with_d)
        for (cout, taps, cin), (wm, wf, wd) in zip(shapes, keep):
            ck(f"wf {cout}x{taps}x{cin} with_d={with_d}", wf[:, :taps * cin], bf(wm).reshape(cout, -1), 0.0)
            ck("   wf padding untouched", wf[:, taps * cin:], torch.full_like(wf[:, taps * cin:], 7.0), 0.0)
            want_d = bf(wm).flip(1).permute(2, 1, 0).reshape(cin, -1) if with_d else torch.full_like(wd, 7.0)
            ck("   wd", wd, want_d, 0.0)
    ck.done()


# ---- 3-channel boundary convs -------------------------------------------------------------------------------------
@pytest.mark.parametrize("n,h,w,C", [(2, 16, 24, 128), (3, 64, 64, 128), (1, 7, 5, 64), (2, 128, 128, 128),
                                     (1, 224, 224, 128), (1, 256, 256, 128)])
def test_boundary_convs_as_gemms(ops, n, h, w, C):
    """conv_in / conv_out (3 image channels) as one-k-block tcgen05 GEMMs: im2col3 + conv_gemm / conv_wgrad +
    nhwc_to_nchw, forward, dgrad and wgrad against F.conv2d autograd."""
    from polyp_image_generator_b200.ops import taps_1x1, taps_3x3
    ck = Check()
    torch.manual_seed(3)
    grid = (n, h, w)
    x = torch.randn(n, 3, h, w, device=DEV)
    wt = torch.randn(C, 3, 3, 3, device=DEV) * 0.2
    b = torch.randn(C, device=DEV)
    wf = torch.zeros(C, 64, device=DEV, dtype=torch.bfloat16)
    wf[:, :27] = wt.permute(0, 2, 3, 1).reshape(C, 27)                      # [co][tap][ci]
    csum = torch.zeros(3, device=DEV)
    pat = ops.im2col3(x, chan_sum=csum)
    xb = bf(x).float()
    want_pat = F.unfold(xb, 3, padding=1).view(n, 3, 9, h, w).permute(0, 3, 4, 2, 1).reshape(n, h, w, 27)
    ck("im2col3", pat[..., :27], want_pat, 0.0)
    ck("im2col3 zero padding", pat[..., 27:].float().abs().sum().reshape(1), torch.zeros(1, device=DEV), 0.0)
    ck("channel sums", csum, x.sum((0, 2, 3)), 1e-5)
    got = ops.conv_gemm(pat, None, taps_1x1(), wf, C, grid, bias=b)
    ck("conv_in fwd", got, F.conv2d(xb, bf(wt).float(), b, padding=1).permute(0, 2, 3, 1), 4e-3)
    dy = bf(torch.randn(n, h, w, C, device=DEV))
    R = torch.zeros(C, 64, device=DEV)
    ops.conv_wgrad(dy, pat, None, taps_1x1(), R, grid)
    wr = wt.clone().requires_grad_(True)
    F.conv2d(xb, wr, None, padding=1).backward(dy.float().permute(0, 3, 1, 2))
    ck("conv_in wgrad", R[:, :27], wr.grad.permute(0, 2, 3, 1).reshape(C, 27), 2e-3)
    a = bf(torch.randn(n, h, w, C, device=DEV))
    wo = torch.randn(3, C, 3, 3, device=DEV) * 0.05
    bo = torch.randn(3, device=DEV)
    wof = torch.zeros(32, 9 * C, device=DEV, dtype=torch.bfloat16)
    wof[:3] = wo.permute(0, 2, 3, 1).reshape(3, 9 * C)
    b32 = torch.zeros(32, device=DEV)
    b32[:3] = bo
    o32 = ops.conv_gemm(a, None, taps_3x3(C), wof, 32, grid, bias=b32, out_f32=True)
    got = ops.nhwc_to_nchw_f32(o32, 3)
    ar = a.float().permute(0, 3, 1, 2).contiguous().requires_grad_(True)
    wor = bf(wo).float().requires_grad_(True)
    yr = F.conv2d(ar, wor, bo, padding=1)
    ck("conv_out fwd (fp32 out)", got, yr, 1e-4)
    dyo = torch.randn(n, 3, h, w, device=DEV)
    yr.backward(bf(dyo).float())
    pd = ops.im2col3(dyo)
    wd = torch.zeros(C, 64, device=DEV, dtype=torch.bfloat16)
    wd[:, :27] = wo.permute(0, 2, 3, 1).reshape(3, 9, C).flip(1).permute(2, 1, 0).reshape(C, 27)
    da = ops.conv_gemm(pd, None, taps_1x1(), wd, C, grid)
    ck("conv_out dgrad", da, ar.grad.permute(0, 2, 3, 1), 4e-3)
    R = torch.zeros(C, 64, device=DEV)
    ops.conv_wgrad(a, pd, None, taps_1x1(), R, grid)
    dwo = R[:, :27].view(C, 9, 3).flip(1).permute(2, 1, 0)                  # [co][tap][ci]
    ck("conv_out wgrad", dwo, wor.grad.permute(0, 2, 3, 1).reshape(3, 9, C), 2e-3)
    ck.done()


# ---- attention (narrow heads: the reference default attention_head_dim = 8) -----------------------------------------
@pytest.mark.parametrize("b,t,heads,d", [(2, 64, 64, 8), (3, 16, 64, 8), (1, 64, 8, 64), (2, 49, 4, 16), (2, 256, 64, 8),
                                         (2, 196, 64, 8), (4, 4, 64, 8), (2, 1, 64, 8)])
def test_narrow_head_attention_vs_sdpa(ops, b, t, heads, d):
    torch.manual_seed(4)
    Cc = heads * d
    qkv = bf(torch.randn(b * t, 3 * Cc, device=DEV))
    scale = d ** -0.5
    o, lse = ops.attn_fwd(qkv, b, t, heads, d, scale)
    qr = qkv.float().clone().requires_grad_(True)
    q, k, v = [z.reshape(b, t, heads, d).transpose(1, 2) for z in qr.split(Cc, 1)]
    orf = F.scaled_dot_product_attention(q, k, v, scale=scale).transpose(1, 2).reshape(b * t, Cc)
    ck = Check()
    ck("attn fwd", o, orf, 5e-3)
    do = bf(torch.randn(b * t, Cc, device=DEV))
    orf.backward(do.float())
    dqkv = ops.attn_bwd(qkv, o, do, lse, b, t, heads, d, scale)
    ck("attn bwd", dqkv, qr.grad, 1e-2)
    ck.done()


# ---- tcgen05 conv fprop / dgrad -------------------------------------------------------------------------------------
def test_gemm_layout_probe_and_linears(ops):
    from polyp_image_generator_b200.ops import taps_1x1
    ck = Check()
    torch.manual_seed(5)
    M, K, N = 128, 64, 128
    X = torch.zeros(1, 1, M, K, device=DEV, dtype=torch.bfloat16)
    X[0, 0, 5, 3] = 1.0
    X[0, 0, 77, 40] = 2.0
    Wt = torch.zeros(N, K, device=DEV, dtype=torch.bfloat16)
    for i in range(K):
        Wt[i, i] = 1.0
    out = ops.conv_gemm(X, None, taps_1x1(), Wt, N, (1, 1, M)).float().reshape(M, N)
    want = torch.zeros(M, N, device=DEV)
    want[5, 3], want[77, 40] = 1.0, 2.0
    ck("one-hot layout probe", out, want, 0.0)
    for (M, K, N) in [(128, 64, 128), (256, 128, 128), (4096, 512, 512), (4096, 512, 1536), (300, 192, 128),
                      (64, 512, 256), (2 * 196, 512, 1536), (2 * 49, 512, 512)]:
        x = bf(torch.randn(1, 1, M, K, device=DEV))
        w = bf(torch.randn(N, K, device=DEV) * 0.05)
        b = torch.randn(N, device=DEV)
        out = ops.conv_gemm(x, None, taps_1x1(), w, N, (1, 1, M), bias=b)
        ck(f"linear M{M} K{K} N{N}", out.reshape(M, N), x.float().reshape(M, K) @ w.float().t() + b, 4e-3)
    x = bf(torch.randn(1, 1, 1024, 256, device=DEV))
    w = bf(torch.randn(512, 256, device=DEV) * 0.05)
    with env(DDPM_BLOCK_N=256):
        out = ops.conv_gemm(x, None, taps_1x1(), w, 512, (1, 1, 1024))
    ck("linear BLOCK_N=256", out.reshape(1024, 512), x.float().reshape(1024, 256) @ w.float().t(), 4e-3)
    out = ops.conv_gemm(x, None, taps_1x1(), w, 512, (1, 1, 1024), out_f32=True)
    ck("linear fp32 out", out.reshape(1024, 512), x.float().reshape(1024, 256) @ w.float().t(), 1e-5)
    ck.done()


# (n, h, w, cin, cout): every 3x3 shape class of SURVEY.md Appendix D at 128^2 (x the batch the kernels see in tests),
# the 224^2 pyramid (224, 112, 56, 28, 14, 7), 256^2, and ragged / odd sizes
CONV3_CASES = [(2, 16, 16, 128, 128), (1, 128, 128, 128, 128), (3, 32, 32, 256, 256), (4, 8, 8, 512, 512),
               (9, 4, 4, 512, 512), (2, 64, 64, 128, 256), (2, 14, 14, 128, 128), (3, 7, 7, 64, 128),
               (1, 200, 136, 64, 128), (3, 70, 96, 128, 256), (5, 64, 64, 256, 128), (2, 128, 128, 256, 128),
               (1, 65, 64, 64, 64), (2, 3, 200, 64, 128), (32, 4, 4, 1024, 512), (16, 8, 8, 512, 1024),
               (4, 128, 128, 128, 128), (3, 128, 128, 128, 128), (2, 64, 64, 256, 256), (2, 64, 64, 384, 128),
               (2, 32, 32, 512, 256), (2, 32, 32, 384, 256), (2, 16, 16, 768, 256), (2, 8, 8, 1024, 512),
               (2, 8, 8, 768, 512), (1, 256, 256, 128, 128), (2, 256, 256, 256, 128), (1, 224, 224, 128, 128),
               (2, 224, 224, 256, 128), (2, 112, 112, 128, 128), (2, 112, 112, 256, 256), (2, 56, 56, 256, 256),
               (2, 28, 28, 256, 256), (3, 255, 255, 64, 128), (1, 192, 192, 128, 128), (1, 160, 320, 64, 64)]


@pytest.mark.parametrize("n,h,w,cin,cout", CONV3_CASES)
def test_conv3x3_fprop_with_fused_epilogue(ops, n, h, w, cin, cout):
    """conv_gemm (halo-resident CTA-pair / single-CTA kernels at W >= 64, generic kernel below) + bias + temb + residual
    against F.conv2d in fp32 on the same bf16 operands.  For the halo sizes the single-CTA variant is checked too."""
    from polyp_image_generator_b200.ops import taps_3x3
    ck = Check()
    torch.manual_seed(5)
    x = bf(torch.randn(n, h, w, cin, device=DEV))
    w4 = bf(torch.randn(cout, cin, 3, 3, device=DEV) * 0.03)
    wk = w4.permute(0, 2, 3, 1).reshape(cout, 9 * cin).contiguous()
    b = torch.randn(cout, device=DEV)
    temb = torch.randn(n, cout, device=DEV)
    res = bf(torch.randn(n, h, w, cout, device=DEV))
    want = conv_ref(x, w4, b) + temb[:, None, None, :] + res.float()
    out = ops.conv_gemm(x, None, taps_3x3(cin), wk, cout, (n, h, w), bias=b, temb=temb, res=res)
    ck("conv3x3 +bias+temb+res", out, want, 4e-3)
    if w >= 64:
        with env(DDPM_HALO_PAIR=0):
            out1 = ops.conv_gemm(x, None, taps_3x3(cin), wk, cout, (n, h, w), bias=b, temb=temb, res=res)
        ck("   single-CTA halo kernel", out1, want, 4e-3)
        with env(DDPM_HALO=0):
            out2 = ops.conv_gemm(x, None, taps_3x3(cin), wk, cout, (n, h, w), bias=b, temb=temb, res=res)
        ck("   generic kernel", out2, want, 4e-3)
    ck.done()


@pytest.mark.parametrize("n,h,w,cin,cout", [(8, 128, 128, 128, 128), (5, 128, 128, 256, 128), (16, 64, 64, 256, 256),
                                            (3, 256, 256, 128, 128), (24, 32, 32, 256, 512)])
def test_halo_pair_dynamic_schedule_matches_static(ops, n, h, w, cin, cout):
    """The CTA-pair halo conv draws its tiles from a self-resetting global counter (conv_halo.cu, TileFeed): same
    outputs, bit for bit, as the static round-robin -- plain, +residual (next-tile L2 prefetch through peek()) and the
    GroupNorm-backward epilogue -- over repeated launches (the counter must be back at zero after each) and against
    F.conv2d."""
    from polyp_image_generator_b200.ops import taps_3x3
    ck = Check()
    torch.manual_seed(3)
    x = bf(torch.randn(n, h, w, cin, device=DEV))
    w4 = bf(torch.randn(cout, cin, 3, 3, device=DEV) * 0.03)
    wk = w4.permute(0, 2, 3, 1).reshape(cout, 9 * cin).contiguous()
    b = torch.randn(cout, device=DEV)
    res = bf(torch.randn(n, h, w, cout, device=DEV))
    want = conv_ref(x, w4, b)
    with env(DDPM_HALO_DYNAMIC=0):
        s_plain = ops.conv_gemm(x, None, taps_3x3(cin), wk, cout, (n, h, w), bias=b)
        s_res = ops.conv_gemm(x, None, taps_3x3(cin), wk, cout, (n, h, w), bias=b, res=res)
    ck("static vs F.conv2d", s_plain, want, 4e-3)
    for rep in range(3):
        d_plain = ops.conv_gemm(x, None, taps_3x3(cin), wk, cout, (n, h, w), bias=b)
        d_res = ops.conv_gemm(x, None, taps_3x3(cin), wk, cout, (n, h, w), bias=b, res=res)
        assert torch.equal(d_plain, s_plain), f"dynamic != static (plain, launch {rep})"
        assert torch.equal(d_res, s_res), f"dynamic != static (+residual, launch {rep})"
    ck("dynamic +res vs F.conv2d", d_res, want + res.float(), 4e-3)
    # GroupNorm-backward epilogue: dz and the per-(n, c) sums (atomics: order differs, values agree to fp32 rounding)
    gamma, beta = torch.randn(cout, device=DEV) * 0.5 + 1, torch.randn(cout, device=DEV) * 0.2
    xg = bf(torch.randn(n, h, w, cout, device=DEV))
    _, _, coef = ops.gn_fwd(xg, None, 32, 1e-5, gamma, beta, True, want_coef=True)
    outs = []
    for dyn in (0, 1, 1):
        with env(DDPM_HALO_DYNAMIC=dyn):
            sums = torch.zeros(n, cout, 2, device=DEV)
            dz = ops.conv_gemm(x, None, taps_3x3(cin), wk, cout, (n, h, w), gn=(xg, None, coef, True, sums))
            outs.append((dz, sums))
    for dz, sums in outs[1:]:
        assert torch.equal(dz, outs[0][0]), "dynamic != static (GroupNorm-backward epilogue)"
        ck("   fused sums", sums, outs[0][1], 1e-5)
    torch.cuda.synchronize()
    ck.done()


def test_conv_concat_slices_stride2_and_dgrad(ops):
    from polyp_image_generator_b200.ops import taps_1x1, taps_3x3, taps_s2d
    ck = Check()
    torch.manual_seed(5)
    for (n, h, w_, c0, c1, cout) in [(2, 64, 64, 128, 64, 128), (2, 16, 16, 256, 128, 256), (2, 128, 128, 128, 128, 128),
                                     (1, 256, 256, 128, 128, 128), (1, 224, 224, 128, 128, 128)]:
        xa, xb = bf(torch.randn(n, h, w_, c0, device=DEV)), bf(torch.randn(n, h, w_, c1, device=DEV))
        w4 = bf(torch.randn(cout, c0 + c1, 3, 3, device=DEV) * 0.03)
        wk = w4.permute(0, 2, 3, 1).reshape(cout, -1).contiguous()
        out = ops.conv_gemm(xa, xb, taps_3x3(c0 + c1), wk, cout, (n, h, w_))
        ck(f"conv3x3 concat {c0}+{c1} -> {cout} @{h}x{w_}", out, conv_ref(torch.cat([xa, xb], -1), w4, None), 4e-3)
        w1 = bf(torch.randn(cout, c0 + c1, device=DEV) * 0.05)
        out = ops.conv_gemm(xa, xb, taps_1x1(), w1, cout, (n, h, w_))
        ck(f"conv1x1 concat @{h}x{w_}", out, torch.cat([xa, xb], -1).float() @ w1.float().t(), 4e-3)
    n, h, w_, cout = 2, 16, 16, 256
    w1 = bf(torch.randn(cout, 384, device=DEV) * 0.05)
    big = bf(torch.randn(n, h, w_, 512, device=DEV))
    outbig = torch.zeros(n, h, w_, 512, device=DEV, dtype=torch.bfloat16)
    ops.conv_gemm(big[..., 128:384], None, taps_1x1(), w1[:, :256].contiguous(), cout, (n, h, w_), out=outbig[..., 256:])
    ck("conv1x1 sliced views", outbig[..., 256:], big[..., 128:384].float() @ w1[:, :256].float().t(), 4e-3)
    for pad in (1, 0):
        for (n, h, w_, c) in [(2, 16, 16, 128), (2, 128, 128, 128), (1, 224, 224, 128), (2, 14, 14, 256), (1, 256, 256, 128)]:
            x = bf(torch.randn(n, h, w_, c, device=DEV))
            w4 = bf(torch.randn(c, c, 3, 3, device=DEV) * 0.03)
            wk = w4.permute(0, 2, 3, 1).reshape(c, -1).contiguous()
            s2d = ops.space_to_depth(x)
            out = ops.conv_gemm(s2d, None, taps_s2d(c, n, pad), wk, c, (n, h // 2, w_ // 2), src_n=4 * n)
            xin = x.float().permute(0, 3, 1, 2)
            if pad == 0:
                xin = F.pad(xin, (0, 1, 0, 1))
            want = F.conv2d(xin, w4.float(), None, stride=2, padding=pad).permute(0, 2, 3, 1)
            ck(f"conv3x3 stride2 pad={pad} {h}x{w_} c{c}", out, want, 4e-3)
    for (n, h, w_, cin, cout) in [(2, 16, 16, 128, 256), (2, 128, 128, 128, 128), (1, 256, 256, 128, 256),
                                  (1, 224, 224, 256, 128)]:
        wm = torch.randn(cout, 9, cin, device=DEV) * 0.03
        wd = torch.empty(cin, 9 * cout, device=DEV, dtype=torch.bfloat16)
        ops.prep_weight(wm, None, wd, cout, 9, cin)
        dy = bf(torch.randn(n, h, w_, cout, device=DEV))
        dx = ops.conv_gemm(dy, None, taps_3x3(cout), wd, cin, (n, h, w_))
        w4 = bf(wm).float().reshape(cout, 3, 3, cin).permute(0, 3, 1, 2)
        xr = torch.zeros(n, cin, h, w_, device=DEV, requires_grad=True)
        F.conv2d(xr, w4, None, padding=1).backward(dy.float().permute(0, 3, 1, 2))
        ck(f"conv3x3 dgrad {h}x{w_} {cout}->{cin}", dx, xr.grad.permute(0, 2, 3, 1), 4e-3)
    ck.done()


# ---- wgrad --------------------------------------------------------------------------------------------------------
def test_wgrad_probe_and_linear_shapes(ops):
    from polyp_image_generator_b200.ops import taps_1x1
    ck = Check()
    torch.manual_seed(6)
    P, Co, Ci = 128, 128, 128
    dy = torch.zeros(1, 1, P, Co, device=DEV, dtype=torch.bfloat16)
    x = torch.zeros(1, 1, P, Ci, device=DEV, dtype=torch.bfloat16)
    dy[0, 0, 9, 70] = 1.0
    x[0, 0, 9, 33] = 3.0
    dy[0, 0, 100, 2] = 1.0
    x[0, 0, 100, 127] = 5.0
    dw = torch.zeros(Co, Ci, device=DEV)
    ops.conv_wgrad(dy, x, None, taps_1x1(), dw, (1, 1, P), accumulate=False)
    want = torch.zeros(Co, Ci, device=DEV)
    want[70, 33], want[2, 127] = 3.0, 5.0
    ck("one-hot wgrad probe", dw, want, 0.0)
    for (M, Co, Ci, splits) in [(128, 128, 128, 1), (1024, 128, 128, 1), (4096, 512, 512, 0), (4096, 256, 192, 4),
                                (300, 128, 64, 1), (100, 128, 128, 1)]:
        dy = bf(torch.randn(1, 1, M, Co, device=DEV))
        x = bf(torch.randn(1, 1, M, Ci, device=DEV))
        dw = torch.zeros(Co, Ci, device=DEV)
        ops.conv_wgrad(dy, x, None, taps_1x1(), dw, (1, 1, M), accumulate=False, splits=splits)
        ck(f"linear wgrad M{M} Co{Co} Ci{Ci} splits={splits}", dw, dy.float().reshape(M, Co).t() @ x.float().reshape(M, Ci),
           2e-3)
    ck.done()


# row-resident kernel: W % 16 == 0, channels % 128 == 0; segment shape hb = 1 (W >= 128) or hb = 128 / W rows (W < 128)
WGRAD_CASES = [(2, 16, 16, 128, 128), (1, 64, 64, 128, 256), (4, 8, 8, 512, 512), (9, 4, 4, 256, 512),
               (2, 14, 14, 128, 128), (3, 7, 7, 64, 128), (2, 128, 128, 128, 128), (1, 64, 64, 256, 128),
               (1, 256, 256, 128, 128), (1, 80, 80, 128, 128), (3, 5, 64, 128, 256), (1, 64, 96, 384, 128),
               (3, 32, 32, 256, 256), (2, 16, 16, 512, 256), (2, 48, 48, 128, 128), (2, 6, 32, 128, 128),
               (4, 128, 128, 256, 128), (1, 224, 224, 128, 128), (2, 112, 112, 256, 256), (2, 56, 56, 256, 256),
               (2, 28, 28, 512, 256), (2, 256, 256, 256, 128), (2, 8, 8, 1024, 512), (2, 64, 64, 256, 256)]


@pytest.mark.parametrize("n,h,w,cin,cout", WGRAD_CASES)
def test_conv3x3_wgrad_and_bias_grad(ops, n, h, w, cin, cout):
    from polyp_image_generator_b200.ops import taps_3x3
    ck = Check()
    torch.manual_seed(6)
    x = bf(torch.randn(n, h, w, cin, device=DEV))
    dy = bf(torch.randn(n, h, w, cout, device=DEV))
    dw = torch.zeros(cout, 9 * cin, device=DEV)
    db = torch.zeros(cout, device=DEV)
    ops.conv_wgrad(dy, x, None, taps_3x3(cin), dw, (n, h, w), accumulate=True, dbias=db)
    wr = torch.zeros(cout, cin, 3, 3, device=DEV, requires_grad=True)
    F.conv2d(x.float().permute(0, 3, 1, 2), wr, None, padding=1).backward(dy.float().permute(0, 3, 1, 2))
    want = wr.grad.permute(0, 2, 3, 1).reshape(cout, -1)
    ck("conv3x3 wgrad", dw, want, 2e-3)
    ck("   bias grad", db, dy.float().sum((0, 1, 2)), 2e-3)
    with env(DDPM_WGRAD_ROW=0):      # generic one-tap-per-CTA kernel on the same problem
        dw2 = torch.zeros(cout, 9 * cin, device=DEV)
        ops.conv_wgrad(dy, x, None, taps_3x3(cin), dw2, (n, h, w), accumulate=True)
    ck("   generic wgrad kernel", dw2, want, 2e-3)
    ck.done()


def test_wgrad_concat_and_stride2(ops):
    from polyp_image_generator_b200.ops import taps_3x3, taps_s2d
    ck = Check()
    torch.manual_seed(6)
    for (n, h, w_, c0, c1, cout) in [(2, 16, 16, 256, 128, 128), (2, 64, 64, 128, 128, 128), (1, 128, 128, 128, 128, 128),
                                     (1, 256, 256, 128, 128, 128)]:
        xa, xb = bf(torch.randn(n, h, w_, c0, device=DEV)), bf(torch.randn(n, h, w_, c1, device=DEV))
        dy = bf(torch.randn(n, h, w_, cout, device=DEV))
        dw = torch.zeros(cout, 9 * (c0 + c1), device=DEV)
        ops.conv_wgrad(dy, xa, xb, taps_3x3(c0 + c1), dw, (n, h, w_))
        wr = torch.zeros(cout, c0 + c1, 3, 3, device=DEV, requires_grad=True)
        F.conv2d(torch.cat([xa, xb], -1).float().permute(0, 3, 1, 2), wr, None, padding=1).backward(
            dy.float().permute(0, 3, 1, 2))
        ck(f"conv3x3 wgrad concat {c0}+{c1} @{h}x{w_}", dw, wr.grad.permute(0, 2, 3, 1).reshape(cout, -1), 2e-3)
    for pad in (1, 0):
        for (n, h, w_, c) in [(2, 16, 16, 128), (2, 128, 128, 128), (1, 224, 224, 128)]:
            x = bf(torch.randn(n, h, w_, c, device=DEV))
            dy = bf(torch.randn(n, h // 2, w_ // 2, c, device=DEV))
            s2d = ops.space_to_depth(x)
            dw = torch.zeros(c, 9 * c, device=DEV)
            ops.conv_wgrad(dy, s2d, None, taps_s2d(c, n, pad), dw, (n, h // 2, w_ // 2), src_n=4 * n)
            wr = torch.zeros(c, c, 3, 3, device=DEV, requires_grad=True)
            xin = x.float().permute(0, 3, 1, 2)
            if pad == 0:
                xin = F.pad(xin, (0, 1, 0, 1))
            F.conv2d(xin, wr, None, stride=2, padding=pad).backward(dy.float().permute(0, 3, 1, 2))
            ck(f"conv3x3 stride2 wgrad pad={pad} {h}x{w_}", dw, wr.grad.permute(0, 2, 3, 1).reshape(c, -1), 2e-3)
    ck.done()


# ---- fp32-faithful mode: split-bf16 kernels (include/ddpm_b200.h "fp32-faithful mode") ----------------------------------
def _to_split(v):
    hi = v.to(torch.bfloat16)
    lo = (v - hi.float()).to(torch.bfloat16)
    return torch.cat([hi, lo], -1).contiguous()


def _from_split(t):
    c = t.shape[-1] // 2
    return t[..., :c].float() + t[..., c:].float()


@pytest.mark.parametrize("n,h,w,cin,cout", [(2, 128, 128, 128, 128), (2, 64, 64, 256, 128), (3, 32, 32, 256, 256),
                                            (4, 8, 8, 512, 512), (2, 16, 16, 128, 256), (1, 7, 7, 64, 128)])
def test_split_conv_is_fp32_faithful(ops, n, h, w, cin, cout):
    """conv over a split tensor as ONE GEMM [hi | lo | hi] x [W_hi | W_hi | W_lo] (+ fp32 bias, temb, split residual, split
    output): against F.conv2d in fp32 (TF32 off) on the SAME fp32 values -- 1e-5, i.e. fp32-level, not bf16-level."""
    from polyp_image_generator_b200.ops import taps_3x3
    from polyp_image_generator_b200.unet import UNet2DModel
    ck = Check()
    torch.manual_seed(8)
    xv = torch.randn(n, h, w, cin, device=DEV)
    w4 = torch.randn(cout, cin, 3, 3, device=DEV) * 0.03
    b = torch.randn(cout, device=DEV)
    temb = torch.randn(n, cout, device=DEV)
    resv = torch.randn(n, h, w, cout, device=DEV)
    xs, rs = _to_split(xv), _to_split(resv)
    xq, rq = _from_split(xs), _from_split(rs)                      # the values the split tensors actually hold
    w3 = UNet2DModel._split3(w4.permute(0, 2, 3, 1).reshape(cout, 9 * cin), 9, cin)
    out = ops.conv_gemm(xs, xs[..., :cin], taps_3x3(3 * cin), w3, cout, (n, h, w), bias=b, temb=temb, res=rs,
                        split_io=True)
    want = F.conv2d(xq.permute(0, 3, 1, 2), w4, b, padding=1).permute(0, 2, 3, 1) + temb[:, None, None, :] + rq
    assert out.shape == (n, h, w, 2 * cout)
    ck("split conv3x3 +bias+temb+res", _from_split(out), want, 1e-5)
    ck("   (its bf16 hi half alone is only bf16-accurate)", out[..., :cout].float(), want, 4e-3)
    ck.done()
    assert rel(out[..., :cout].float(), want) > 5e-4


@pytest.mark.parametrize("n,h,w,c0,c1,silu", [(2, 64, 64, 128, 0, True), (2, 16, 16, 256, 128, True),
                                              (3, 8, 8, 512, 512, True), (2, 4, 4, 512, 0, False),
                                              (2, 128, 128, 128, 128, True)])
def test_split_groupnorm_is_fp32_faithful(ops, n, h, w, c0, c1, silu):
    torch.manual_seed(9)
    C = c0 + c1
    xa = torch.randn(n, h, w, c0, device=DEV) * 1.5 + 0.3
    xb = torch.randn(n, h, w, c1, device=DEV) - 0.2 if c1 else None
    gamma = torch.randn(C, device=DEV) * 0.5 + 1
    beta = torch.randn(C, device=DEV) * 0.2
    sa, sb = _to_split(xa), (_to_split(xb) if c1 else None)
    y = ops.gn_fwd_split(sa, sb, 32, 1e-5, gamma, beta, silu)
    xq = _from_split(sa) if not c1 else torch.cat([_from_split(sa), _from_split(sb)], -1)
    want = F.group_norm(xq.permute(0, 3, 1, 2), 32, gamma, beta, 1e-5)
    if silu:
        want = F.silu(want)
    assert y.shape == (n, h, w, 2 * C)
    assert rel(_from_split(y), want.permute(0, 2, 3, 1)) < 2e-5


@pytest.mark.parametrize("b,t,heads,d", [(2, 64, 64, 8), (3, 16, 64, 8), (2, 196, 64, 8), (2, 49, 4, 16)])
def test_split_attention_is_fp32_faithful(ops, b, t, heads, d):
    torch.manual_seed(10)
    C = heads * d
    qkv = torch.randn(b * t, 3 * C, device=DEV)
    qs = _to_split(qkv)
    o = ops.attn_fwd_split(qs, b, t, heads, d, d ** -0.5)
    q, k, v = [z.reshape(b, t, heads, d).transpose(1, 2) for z in _from_split(qs).split(C, 1)]
    want = F.scaled_dot_product_attention(q, k, v, scale=d ** -0.5).transpose(1, 2).reshape(b * t, C)
    assert o.shape == (b * t, 2 * C)
    assert rel(_from_split(o), want) < 2e-5


def test_split_im2col3_layout(ops):
    torch.manual_seed(11)
    x = torch.randn(2, 3, 9, 13, device=DEV)
    pat = ops.im2col3_split(x)
    want = F.unfold(x, 3, padding=1).view(2, 3, 9, 9, 13).permute(0, 3, 4, 2, 1).reshape(2, 9, 13, 27)
    assert pat.shape == (2, 9, 13, 128)
    assert torch.equal(pat[..., :27], pat[..., 64:91])
    assert rel(pat[..., :27].float() + pat[..., 32:59].float(), want) < 1e-5
    assert float(pat[..., 27:32].float().abs().sum() + pat[..., 59:64].float().abs().sum() +
                 pat[..., 91:].float().abs().sum()) == 0.0
