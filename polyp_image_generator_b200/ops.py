"""Thin torch-tensor front end of the C-ABI kernels (include/ddpm_b200.h).

Every function launches hand-written sm_100a kernels on torch's current CUDA stream; PyTorch is used only for
device memory and streams.  There is NO fallback: a non-CUDA tensor or a missing library raises.

Activation tensors are "NHWC views": shape [N, H, W, C] bf16 with stride(3) == 1 and a pixel stride
ld == stride(2) (so channel slices of wider tensors are valid inputs and outputs).
"""
from __future__ import annotations

import ctypes as C
from typing import List, Optional, Sequence, Tuple

import torch

from . import _capi

Tap = Tuple[int, int, int, int]  # (dn, dh, dw, wk)
PREP_TILE = 64                   # DDPM_PREP_TILE of include/ddpm_b200.h (tile edge of ddpm_prep_weights_batched)

# smallest map (pixels per sample) at which the GroupNorm statistics / backward fusions ride on the conv epilogues
import os as _os


def _gn_fuse_min_hw() -> int:
    # read per call (not at import) so that profiles/ab_step.py can capture one CUDA graph per setting in one process
    return int(_os.environ.get("DDPM_GN_FUSE_MIN_HW", "1024"))


def _stream() -> int:
    return torch.cuda.current_stream().cuda_stream


def _ptr(t: Optional[torch.Tensor]) -> Optional[int]:
    if t is None:
        return None
    if not t.is_cuda:
        raise RuntimeError("polyp_image_generator_b200 kernels need CUDA tensors (no CPU fallback exists)")
    return t.data_ptr()


def _nhwc(t: torch.Tensor, what: str) -> Tuple[int, int, int, int, int]:
    """-> (n, h, w, c, ld) of an NHWC bf16 view."""
    if t.dtype != torch.bfloat16 or t.dim() != 4:
        raise ValueError(f"{what}: expected a 4-d bf16 NHWC view, got {tuple(t.shape)} {t.dtype}")
    n, h, w, c = t.shape
    ld = t.stride(2) if w > 1 else (t.stride(1) if h > 1 else (t.stride(0) if n > 1 else c))
    if c > 1 and t.stride(3) != 1:
        raise ValueError(f"{what}: channel stride must be 1")
    if (w > 1 and t.stride(2) != ld) or (h > 1 and t.stride(1) != w * ld) or (n > 1 and t.stride(0) != h * w * ld):
        raise ValueError(f"{what}: not a dense-pixel NHWC view, shape {tuple(t.shape)} strides {t.stride()}")
    return n, h, w, c, ld


def taps_3x3(cin_total: int, pad: int = 1) -> List[Tap]:
    return [(0, r - pad, s - pad, (r * 3 + s) * cin_total) for r in range(3) for s in range(3)]


def taps_1x1() -> List[Tap]:
    return [(0, 0, 0, 0)]


def taps_s2d(cin_total: int, n: int, pad: int) -> List[Tap]:
    """3x3 stride-2 conv over a space-to-depth tensor whose phase p = (row parity)*2 + (col parity) lives at
    batch offset p*n.  pad = 1 (diffusers default) or 0 (downsample_padding=0, i.e. F.pad(x, (0,1,0,1)))."""
    out = []
    for r in range(3):
        for s in range(3):
            ir, ic = r - pad, s - pad          # input row = 2*ho + ir
            ph, dh = ir % 2, (ir - ir % 2) // 2
            pw, dw = ic % 2, (ic - ic % 2) // 2
            out.append(((ph * 2 + pw) * n, dh, dw, (r * 3 + s) * cin_total))
    return out


def timestep_freqs(dim: int, freq_shift: float, max_period: int = 10000) -> torch.Tensor:
    """Host-side frequency table of diffusers' get_timestep_embedding (fp32, same op order)."""
    import math
    half = dim // 2
    exponent = -math.log(max_period) * torch.arange(0, half, dtype=torch.float32)
    exponent = exponent / (half - freq_shift)
    return torch.exp(exponent)


class CudaOps:
    """The product backend: every method is one (or a few) C-ABI calls."""

    name = "cuda"

    def __init__(self):
        self.lib = _capi.load()
        self.launches = 0  # kernels launched through this object (bench.py reports it)
        self._freqs = {}
        self._halo_strips = {}

    # ---- scheduler / loss ----------------------------------------------------------------------------
    def add_noise(self, x0, noise, t, sqrt_ac, sqrt_1mac):
        out = torch.empty_like(x0)
        b = x0.shape[0]
        per = x0.numel() // max(b, 1)
        _capi.check(self.lib.ddpm_add_noise(_ptr(x0), _ptr(noise), _ptr(t), _ptr(sqrt_ac), _ptr(sqrt_1mac), _ptr(out),
                                            b, per, sqrt_ac.numel(), _stream()), "ddpm_add_noise")
        self.launches += 1
        return out

    def mse_fwd_bwd(self, pred, target, want_grad=True):
        loss_sum = torch.zeros(1, device=pred.device, dtype=torch.float32)
        dpred = torch.empty_like(pred) if want_grad else None
        _capi.check(self.lib.ddpm_mse_fwd_bwd(_ptr(pred), _ptr(target), _ptr(loss_sum), _ptr(dpred), pred.numel(),
                                              _stream()), "ddpm_mse_fwd_bwd")
        self.launches += 1
        return loss_sum, dpred

    def scale_by_device_scalar(self, x, scale):
        _capi.check(self.lib.ddpm_scale_by_device_scalar(_ptr(x), _ptr(scale), x.numel(), _stream()),
                    "ddpm_scale_by_device_scalar")
        self.launches += 1
        return x

    def sumsq(self, x, out):
        """out += sum(x^2) over a flat fp32 tensor."""
        _capi.check(self.lib.ddpm_sumsq_f32(_ptr(x), x.numel(), _ptr(out), _stream()), "ddpm_sumsq_f32")
        self.launches += 1

    def adamw_flat(self, p, g, m, v, scal, gnorm_sq, max_norm, lr_dev, lr, beta1, beta2, eps, weight_decay):
        _capi.check(self.lib.ddpm_adamw_flat(_ptr(p), _ptr(g), _ptr(m), _ptr(v), p.numel(), _ptr(scal),
                                             _ptr(gnorm_sq), float(max_norm or 0.0), _ptr(lr_dev), float(lr),
                                             float(beta1), float(beta2), float(eps), float(weight_decay), _stream()),
                    "ddpm_adamw_flat")
        self.launches += 2

    def ddim_step(self, eps, x, z, sa, sb, sap, dirc, sigma, clip, use_clipped, want_x0=False):
        prev = torch.empty_like(x)
        x0 = torch.empty_like(x) if want_x0 else None
        _capi.check(self.lib.ddpm_ddim_step(_ptr(eps), _ptr(x), _ptr(z), _ptr(prev), _ptr(x0), x.numel(), sa, sb, sap,
                                            dirc, sigma, clip, int(bool(use_clipped)), _stream()), "ddpm_ddim_step")
        self.launches += 1
        return prev, x0

    def scheduler_step(self, eps, x, z, sa, sb, c0, ct, sigma, clip, want_x0=False):
        prev = torch.empty_like(x)
        x0 = torch.empty_like(x) if want_x0 else None
        _capi.check(self.lib.ddpm_scheduler_step(_ptr(eps), _ptr(x), _ptr(z), _ptr(prev), _ptr(x0), x.numel(), sa, sb,
                                                 c0, ct, sigma, clip, _stream()), "ddpm_scheduler_step")
        self.launches += 1
        return prev, x0

    def scheduler_step_philox(self, eps, x, sa, sb, c0, ct, sigma, clip, seed, offset):
        prev = torch.empty_like(x)
        _capi.check(self.lib.ddpm_scheduler_step_philox(_ptr(eps), _ptr(x), _ptr(prev), x.numel(), sa, sb, c0, ct,
                                                        sigma, clip, seed, offset, _stream()),
                    "ddpm_scheduler_step_philox")
        self.launches += 1
        return prev

    def unipc_x0(self, eps, x, sigma_t: float, alpha_t: float):
        out = torch.empty_like(x)
        _capi.check(self.lib.ddpm_unipc_x0(_ptr(eps), _ptr(x), _ptr(out), x.numel(), sigma_t, alpha_t, _stream()),
                    "ddpm_unipc_x0")
        self.launches += 1
        return out

    def unipc_update(self, x, m0, m1, mt, cx: float, cm: float, cb: float, rk: float = 1.0, rho0: float = 0.0,
                     rho_t: float = 0.0):
        """(cx x - cm m0) - cb [rho0 (m1 - m0) / rk  (+)  rho_t (mt - m0)]; m1 / mt may be None."""
        out = torch.empty_like(x)
        _capi.check(self.lib.ddpm_unipc_update(_ptr(x), _ptr(m0), _ptr(m1), _ptr(mt), _ptr(out), x.numel(), cx, cm, cb,
                                               rk, rho0, rho_t, _stream()), "ddpm_unipc_update")
        self.launches += 1
        return out

    def to_uint8_nhwc(self, x):
        n, c, h, w = x.shape
        out = torch.empty((n, h, w, c), device=x.device, dtype=torch.uint8)
        _capi.check(self.lib.ddpm_to_uint8_nhwc(_ptr(x), _ptr(out), n, c, h, w, _stream()), "ddpm_to_uint8_nhwc")
        self.launches += 1
        return out

    # ---- tcgen05 conv / linear -------------------------------------------------------------------------
    @staticmethod
    def gn_bwd_fusable(grid: Tuple[int, int, int]) -> bool:
        """Should the GroupNorm-backward first half run in the dgrad epilogue of a conv over this pixel grid?
        Possible when a warp's 32 tile rows share a sample (H*W >= 128 or H*W % 32 == 0); worth it from 64x64 up:
        measured on B200 (profiles/bench_kernels.py convgn) the fused epilogue costs +0.058 ms at 128^2 against
        0.19 ms saved in the GroupNorm pass, but +0.040 ms at 32^2 against 0.01 ms saved."""
        hw = grid[1] * grid[2]
        return hw >= _gn_fuse_min_hw()

    def gn_stats_fusable(self, grid: Tuple[int, int, int]) -> bool:
        """Should a 3x3 stride-1 conv that PRODUCES a tensor over this grid also reduce the per-(sample, channel) moments
        its consuming GroupNorm needs (conv_gemm(csum=...))?  At >= 64x64 the GroupNorm forward is HBM-bound and drops
        from the two-phase team kernel (0.139 ms at 128^2 x 128, 59 % of the copy peak) to one streaming pass (0.102 ms,
        81 %).  Only where the conv runs on the halo-resident kernel (ddpm_conv_halo_strips(w) > 0), whose epilogue
        overlaps the next tile's mainloop: in the generic kernel the extra warp reductions are exposed (measured: +1.6 ms
        on the 256^2 LoRA step)."""
        hw = grid[1] * grid[2]
        return hw >= _gn_fuse_min_hw() and self.halo_strips(grid[2]) > 0

    def halo_strips(self, w: int) -> int:
        s = self._halo_strips.get(w)
        if s is None:
            s = self._halo_strips[w] = int(self.lib.ddpm_conv_halo_strips(int(w)))
        return s

    def conv_gemm(self, x0, x1, taps: Sequence[Tap], wgt, cout: int, grid: Tuple[int, int, int], bias=None,
                  temb=None, res=None, out=None, out_f32: bool = False, src_n: int = 0, gn=None, csum=None,
                  split_io: bool = False):
        """out[n,h,w,co] = bias + temb[n] + res + sum_taps X[pix+tap] . wgt[co, wk:wk+cin]; see ddpm_conv_gemm.
        gn = (x0, x1, coef, silu, sums) with coef from gn_fwd(want_coef=True): fuse the first half of the backward of
        y = act(GroupNorm(x0|x1)) into the epilogue (out becomes dz, sums[n, c] += (sum dz, sum dz*x)).
        csum [n, cout / 4, 2] fp32 (zero-filled): += (sum out, sum out^2) per (sample, 4-channel granule) -- the
        statistics of the GroupNorm that consumes `out`, so that its forward is one streaming pass (gn_fwd_from_csum)."""
        n, h, w = grid
        _, _, _, c0, ld0 = _nhwc(x0, "x0")
        a = _capi.ConvArgs()
        a.x0, a.c0, a.ld0 = _ptr(x0), c0, ld0
        if x1 is not None:
            _, _, _, c1, ld1 = _nhwc(x1, "x1")
            a.x1, a.c1, a.ld1 = _ptr(x1), c1, ld1
        a.n, a.h, a.w, a.src_n = n, h, w, src_n
        a.ntaps = len(taps)
        for i, (dn, dh, dw, wk) in enumerate(taps):
            a.tap_dn[i], a.tap_dh[i], a.tap_dw[i], a.tap_wk[i] = dn, dh, dw, wk
        if wgt.dtype != torch.bfloat16 or wgt.dim() != 2 or wgt.stride(1) != 1:
            raise ValueError("wgt must be a bf16 [cout, K] matrix with unit inner stride")
        a.wgt, a.cout, a.ldw, a.k_total = _ptr(wgt), cout, wgt.stride(0), wgt.shape[1]
        if split_io:          # fp32-faithful mode: out / res rows are [hi (cout) | lo (cout)] (ddpm_conv_args.split_io)
            if gn is not None or csum is not None or out_f32:
                raise ValueError("split_io excludes the GroupNorm fusions and fp32 output")
            a.split_io = 1
            if out is None:
                out = torch.empty((n, h, w, 2 * cout), device=x0.device, dtype=torch.bfloat16)
        if out is None:
            out = torch.empty((n, h, w, cout), device=x0.device, dtype=torch.float32 if out_f32 else torch.bfloat16)
        if out_f32:
            a.out_f32, a.ldo = _ptr(out), out.stride(2) if w > 1 else out.shape[3]
        else:
            a.out, a.ldo = _ptr(out), _nhwc(out, "out")[4]
        a.bias = _ptr(bias)
        if temb is not None:
            a.temb, a.ld_temb = _ptr(temb), temb.stride(0)
        if res is not None:
            a.res, a.ldr = _ptr(res), _nhwc(res, "res")[4]
        if gn is not None:
            gx0, gx1, gcoef, gsilu, gsums = gn
            _, _, _, gc0, gld0 = _nhwc(gx0, "gn x0")
            a.gn_x0, a.gn_ld0, a.gn_c0 = _ptr(gx0), gld0, gc0
            if gx1 is not None:
                a.gn_x1, a.gn_ld1 = _ptr(gx1), _nhwc(gx1, "gn x1")[4]
            if gcoef.dtype != torch.float32 or not gcoef.is_contiguous() or gcoef.numel() != n * cout * 2:
                raise ValueError("gn coef must be the contiguous fp32 [n, cout/2, 4] table written by gn_fwd")
            a.gn_coef, a.gn_silu = _ptr(gcoef), int(gsilu)
            if gsums.dtype != torch.float32 or not gsums.is_contiguous() or gsums.numel() != n * cout * 2:
                raise ValueError("gn sums must be a contiguous fp32 [n, cout, 2] tensor")
            a.gn_sums = _ptr(gsums)
        if csum is not None:
            if gn is not None or out_f32:
                raise ValueError("csum excludes the GroupNorm-backward fusion and fp32 output")
            if csum.dtype != torch.float32 or not csum.is_contiguous() or csum.numel() != n * (cout // 4) * 2:
                raise ValueError("csum must be a contiguous fp32 [n, cout / 4, 2] tensor")
            a.out_csum = _ptr(csum)
        ws_elems = self.lib.ddpm_conv_gemm_workspace_elems(C.byref(a)) if gn is None and not out_f32 else 0
        if ws_elems > 0:      # low-resolution layer: split-K over idle SMs, fp32 partial sums in a workspace
            ws = torch.empty(ws_elems, device=x0.device, dtype=torch.float32)
            a.splitk_ws, a.splitk_ws_elems = _ptr(ws), ws_elems
        _capi.check(self.lib.ddpm_conv_gemm(C.byref(a), _stream()), "ddpm_conv_gemm")
        self.launches += 3 if ws_elems > 0 else 1
        return out

    def conv_wgrad(self, dy, x0, x1, taps: Sequence[Tap], dw, grid: Tuple[int, int, int], accumulate: bool = True,
                   src_n: int = 0, splits: int = 0, dbias=None):
        """dw[co, wk+ci] (+)= sum_pix dy[pix, co] * X[pix+tap, ci]; dw: fp32 [cout, K] matrix view.
        dbias[co] += sum_pix dy[pix, co] when given (the launcher adds one column-reduction pass over dy)."""
        n, h, w = grid
        a = _capi.WgradArgs()
        _, _, _, cout, ldy = _nhwc(dy, "dy")
        a.dy, a.ldy, a.cout = _ptr(dy), ldy, cout
        _, _, _, c0, ld0 = _nhwc(x0, "x0")
        a.x0, a.c0, a.ld0 = _ptr(x0), c0, ld0
        if x1 is not None:
            _, _, _, c1, ld1 = _nhwc(x1, "x1")
            a.x1, a.c1, a.ld1 = _ptr(x1), c1, ld1
        a.n, a.h, a.w, a.src_n = n, h, w, src_n
        a.ntaps = len(taps)
        for i, (dn, dh, dw_, wk) in enumerate(taps):
            a.tap_dn[i], a.tap_dh[i], a.tap_dw[i], a.tap_wk[i] = dn, dh, dw_, wk
        if dw.dtype != torch.float32 or dw.dim() != 2 or dw.stride(1) != 1:
            raise ValueError("dw must be an fp32 [cout, K] matrix with unit inner stride")
        a.dw, a.ldw = _ptr(dw), dw.stride(0)
        a.accumulate, a.splits = int(accumulate), splits
        a.dbias = _ptr(dbias)
        _capi.check(self.lib.ddpm_conv_wgrad(C.byref(a), _stream()), "ddpm_conv_wgrad")
        self.launches += 1
        return dw

    def prep_weight(self, w, wf, wd, cout: int, taps: int, cin: int):
        """w: fp32 [cout][taps][cin] (physical); wf: bf16 [cout, taps*cin] or None; wd: bf16 [cin, taps*cout] or None."""
        _capi.check(self.lib.ddpm_prep_weight(_ptr(w), _ptr(wf), wf.stride(0) if wf is not None else 0, _ptr(wd),
                                              wd.stride(0) if wd is not None else 0, cout, taps, cin, _stream()),
                    "ddpm_prep_weight")
        self.launches += 1

    def build_prep_table(self, entries, device):
        """entries: [(w fp32 [cout][taps][cin] view, wf bf16 view, wd bf16 view, cout, taps, cin)] -> (device table,
        n, total_tiles).  The views must stay alive and in place (they are arena slices)."""
        arr = (_capi.PrepDesc * len(entries))()
        tiles = 0
        for i, (w, wf, wd, cout, taps, cin) in enumerate(entries):
            d = arr[i]
            d.w, d.wf, d.ldwf = _ptr(w), _ptr(wf), wf.stride(0)
            d.wd, d.ldwd = _ptr(wd), wd.stride(0)
            d.cout, d.taps, d.cin = cout, taps, cin
            d.tiles_x, d.tiles_y = (cin + PREP_TILE - 1) // PREP_TILE, (cout + PREP_TILE - 1) // PREP_TILE   # ddpm_b200.h
            d.tile_begin = tiles
            tiles += d.tiles_x * d.tiles_y * taps
        raw = torch.frombuffer(bytearray(bytes(arr)), dtype=torch.uint8).to(device)
        return raw, len(entries), tiles

    def prep_weights_batched(self, table, n_entries: int, total_tiles: int, with_d: bool):
        _capi.check(self.lib.ddpm_prep_weights_batched(_ptr(table), n_entries, total_tiles, int(with_d), _stream()),
                    "ddpm_prep_weights_batched")
        self.launches += 1

    # ---- fp32-faithful mode: the same ops on split-bf16 tensors (rows [hi (C) | lo (C)], value = hi + lo) -----------
    def im2col3_split(self, x):
        n, cin, h, w = x.shape
        out = torch.empty((n, h, w, 128), device=x.device, dtype=torch.bfloat16)
        _capi.check(self.lib.ddpm_im2col3_split(_ptr(x), _ptr(out), n, h, w, cin, _stream()), "ddpm_im2col3_split")
        self.launches += 1
        return out

    @staticmethod
    def _split_src(t, what):
        n, h, w, c2, ld = _nhwc(t, what)
        if c2 % 2:
            raise ValueError(f"{what}: a split tensor has an even number of stored channels")
        return n, h, w, c2 // 2, ld

    def gn_fwd_split(self, x0, x1, groups: int, eps: float, gamma, beta, silu: bool):
        """GroupNorm(+SiLU) on split tensors: statistics pass + apply pass (exact sigmoid) -> split y [n, h, w, 2C]."""
        n, h, w, c0, ld0 = self._split_src(x0, "x0")
        c1, ld1 = 0, 0
        if x1 is not None:
            _, _, _, c1, ld1 = self._split_src(x1, "x1")
        C_ = c0 + c1
        stats = torch.empty((n, groups, 2), device=x0.device, dtype=torch.float32)
        _capi.check(self.lib.ddpm_gn_stats_split(_ptr(x0), c0, ld0, _ptr(x1), c1, ld1, n, h * w, groups, _ptr(stats),
                                                 _stream()), "ddpm_gn_stats_split")
        y = torch.empty((n, h, w, 2 * C_), device=x0.device, dtype=torch.bfloat16)
        _capi.check(self.lib.ddpm_gn_apply_split(_ptr(x0), c0, ld0, _ptr(x1), c1, ld1, n, h * w, groups, _ptr(stats),
                                                 eps, _ptr(gamma), _ptr(beta), int(silu), _ptr(y), 2 * C_, _stream()),
                    "ddpm_gn_apply_split")
        self.launches += 3
        return y

    def attn_fwd_split(self, qkv, b: int, t: int, heads: int, d: int, scale: float):
        """qkv: split rows [b*t, 6*heads*d] -> split o [b*t, 2*heads*d] (narrow heads only)."""
        if d not in self.NARROW_HEAD_DIMS:
            raise NotImplementedError(f"fp32-faithful attention supports head_dim in {self.NARROW_HEAD_DIMS}, got {d}")
        o = torch.empty((b * t, 2 * heads * d), device=qkv.device, dtype=torch.bfloat16)
        _capi.check(self.lib.ddpm_attn_fwd_split(_ptr(qkv), qkv.stride(0), _ptr(o), o.stride(0), b, t, heads, d, scale,
                                                 _stream()), "ddpm_attn_fwd_split")
        self.launches += 1
        return o

    # ---- 3-channel boundary convs (as GEMMs) --------------------------------------------------------------
    def im2col3(self, x, chan_sum=None):
        """x: NCHW fp32 [n, cin<=4, h, w] -> 3x3/pad-1 patches, NHWC bf16 [n, h, w, 64] (columns tap*cin + k)."""
        n, cin, h, w = x.shape
        out = torch.empty((n, h, w, 64), device=x.device, dtype=torch.bfloat16)
        _capi.check(self.lib.ddpm_im2col3(_ptr(x), _ptr(out), n, h, w, cin, _ptr(chan_sum), _stream()), "ddpm_im2col3")
        self.launches += 1
        out._ddpm_alg_k = 9 * cin          # columns that carry data (FLOP accounting only)
        return out

    def nhwc_to_nchw_f32(self, src, cout: int):
        """src: fp32 [n, h, w, ld] -> NCHW fp32 [n, cout, h, w] of its first cout (<= 4) channels."""
        n, h, w, ld = src.shape
        out = torch.empty((n, cout, h, w), device=src.device, dtype=torch.float32)
        _capi.check(self.lib.ddpm_nhwc_to_nchw_f32(_ptr(src), ld, _ptr(out), n, h, w, cout, _stream()),
                    "ddpm_nhwc_to_nchw_f32")
        self.launches += 1
        return out

    # ---- GroupNorm ---------------------------------------------------------------------------------------
    def gn_stats(self, x0, x1, groups: int):
        n, h, w, c0, ld0 = _nhwc(x0, "x0")
        c1, ld1 = 0, 0
        if x1 is not None:
            _, _, _, c1, ld1 = _nhwc(x1, "x1")
        stats = torch.empty((n, groups, 2), device=x0.device, dtype=torch.float32)
        _capi.check(self.lib.ddpm_gn_stats(_ptr(x0), c0, ld0, _ptr(x1), c1, ld1, n, h * w, groups, _ptr(stats),
                                           _stream()), "ddpm_gn_stats")
        self.launches += 2
        return stats

    def gn_fwd_from_csum(self, x0, x1, cs0, cs1, groups: int, eps: float, gamma, beta, silu: bool,
                         want_coef: bool = False):
        """GroupNorm forward whose statistics were already reduced by the producing convs (conv_gemm(csum=...)):
        channel moments -> group stats (tiny kernel), then ONE streaming normalise pass.  Same returns as gn_fwd."""
        n, h, w, c0, _ = _nhwc(x0, "x0")
        c1 = x1.shape[-1] if x1 is not None else 0
        stats = torch.empty((n, groups, 2), device=x0.device, dtype=torch.float32)
        _capi.check(self.lib.ddpm_gn_stats_from_csum(_ptr(cs0), c0, _ptr(cs1), c1, n, groups, _ptr(stats), _stream()),
                    "ddpm_gn_stats_from_csum")
        self.launches += 1
        coef = torch.empty((n, (c0 + c1) // 2, 4), device=x0.device, dtype=torch.float32) if want_coef else None
        out = self.gn_apply(x0, x1, groups, stats, eps, gamma, beta, silu, coef=coef)
        return (stats, out, coef) if want_coef else (stats, out)

    def gn_apply(self, x0, x1, groups: int, stats, eps: float, gamma, beta, silu: bool, out=None, coef=None):
        n, h, w, c0, ld0 = _nhwc(x0, "x0")
        c1, ld1 = 0, 0
        if x1 is not None:
            _, _, _, c1, ld1 = _nhwc(x1, "x1")
        if out is None:
            out = torch.empty((n, h, w, c0 + c1), device=x0.device, dtype=torch.bfloat16)
        _capi.check(self.lib.ddpm_gn_apply(_ptr(x0), c0, ld0, _ptr(x1), c1, ld1, n, h * w, groups, _ptr(stats), eps,
                                           _ptr(gamma), _ptr(beta), int(silu), _ptr(out), _nhwc(out, "out")[4],
                                           _ptr(coef), _stream()), "ddpm_gn_apply")
        self.launches += 1
        return out

    def gn_fwd(self, x0, x1, groups: int, eps: float, gamma, beta, silu: bool, out=None, want_coef: bool = False):
        """Fused statistics + normalise/affine/(SiLU): -> (stats [n, groups, 2] fp32, y bf16 NHWC[, coef]).
        coef (want_coef): per-(sample, channel) affine table for the GroupNorm-backward fusion of conv_gemm."""
        n, h, w, c0, ld0 = _nhwc(x0, "x0")
        c1, ld1 = 0, 0
        if x1 is not None:
            _, _, _, c1, ld1 = _nhwc(x1, "x1")
        stats = torch.empty((n, groups, 2), device=x0.device, dtype=torch.float32)
        if out is None:
            out = torch.empty((n, h, w, c0 + c1), device=x0.device, dtype=torch.bfloat16)
        ws = torch.empty(n, device=x0.device, dtype=torch.int32)
        coef = torch.empty((n, (c0 + c1) // 2, 4), device=x0.device, dtype=torch.float32) if want_coef else None
        _capi.check(self.lib.ddpm_gn_fwd(_ptr(x0), c0, ld0, _ptr(x1), c1, ld1, n, h * w, groups, eps, _ptr(gamma),
                                         _ptr(beta), int(silu), _ptr(stats), _ptr(out), _nhwc(out, "out")[4],
                                         _ptr(coef), _ptr(ws), _stream()), "ddpm_gn_fwd")
        self.launches += 1
        if want_coef:
            return stats, out, coef
        return stats, out

    def gn_bwd(self, x0, x1, groups: int, stats, eps: float, gamma, beta, silu: bool, dy, add0=None, add1=None,
               dgamma=None, dbeta=None, need_dx1: bool = True, defer=None):
        """-> (dx0, dx1).  dgamma / dbeta are accumulated in place.  defer(fn): run the parameter-gradient reduction
        through fn on the caller's side stream instead of behind the main kernel."""
        n, h, w, c0, ld0 = _nhwc(x0, "x0")
        c1, ld1 = 0, 0
        if x1 is not None:
            _, _, _, c1, ld1 = _nhwc(x1, "x1")
        C_ = c0 + c1
        dx0 = torch.empty((n, h, w, c0), device=x0.device, dtype=torch.bfloat16)
        dx1 = torch.empty((n, h, w, c1), device=x0.device, dtype=torch.bfloat16) if (c1 and need_dx1) else None
        ws = torch.empty(n * C_ * 2 + n, device=x0.device, dtype=torch.float32)
        want_dp = dgamma is not None or dbeta is not None
        dg_in, db_in = (None, None) if (defer is not None and want_dp) else (dgamma, dbeta)
        _capi.check(self.lib.ddpm_gn_bwd(
            _ptr(x0), c0, ld0, _ptr(x1), c1, ld1, n, h * w, groups, _ptr(stats), eps, _ptr(gamma), _ptr(beta),
            int(silu), _ptr(dy), _nhwc(dy, "dy")[4],
            _ptr(add0), _nhwc(add0, "add0")[4] if add0 is not None else 0,
            _ptr(add1), _nhwc(add1, "add1")[4] if add1 is not None else 0,
            _ptr(dx0), c0, _ptr(dx1), c1, _ptr(dg_in), _ptr(db_in), _ptr(ws), _stream()), "ddpm_gn_bwd")
        if defer is not None and want_dp:
            defer(lambda: _capi.check(self.lib.ddpm_gn_bwd_dparams(_ptr(ws), None, n, C_, groups, h * w, eps,
                                                                   _ptr(dgamma), _ptr(dbeta), _stream()),
                                      "ddpm_gn_bwd_dparams"))
        self.launches += 2 if want_dp else 1
        return dx0, dx1

    def gn_bwd_apply(self, x0, x1, groups: int, stats, eps: float, gamma, dz, sums, add0=None, add1=None,
                     dgamma=None, dbeta=None, need_dx1: bool = True, out_nc=None, out_c=None, defer=None):
        """Second half of the GroupNorm backward (first half fused into conv_gemm(gn=...)) -> (dx0, dx1).
        out_nc[n, c] / out_c[c] are INCREMENTED by the pixel sums of dx (the caller zero-fills out_nc).
        defer(fn): as in gn_bwd."""
        n, h, w, c0, ld0 = _nhwc(x0, "x0")
        c1, ld1 = 0, 0
        if x1 is not None:
            _, _, _, c1, ld1 = _nhwc(x1, "x1")
        dx0 = torch.empty((n, h, w, c0), device=x0.device, dtype=torch.bfloat16)
        dx1 = torch.empty((n, h, w, c1), device=x0.device, dtype=torch.bfloat16) if (c1 and need_dx1) else None
        want_dp = dgamma is not None or dbeta is not None
        dg_in, db_in = (None, None) if (defer is not None and want_dp) else (dgamma, dbeta)
        _capi.check(self.lib.ddpm_gn_bwd_apply(
            _ptr(x0), c0, ld0, _ptr(x1), c1, ld1, n, h * w, groups, _ptr(stats), eps, _ptr(gamma),
            _ptr(dz), _nhwc(dz, "dz")[4], _ptr(sums),
            _ptr(add0), _nhwc(add0, "add0")[4] if add0 is not None else 0,
            _ptr(add1), _nhwc(add1, "add1")[4] if add1 is not None else 0,
            _ptr(dx0), c0, _ptr(dx1), c1, _ptr(dg_in), _ptr(db_in),
            _ptr(out_nc), out_nc.stride(0) if out_nc is not None else 0, _ptr(out_c), _stream()), "ddpm_gn_bwd_apply")
        if defer is not None and want_dp:
            defer(lambda: _capi.check(self.lib.ddpm_gn_bwd_dparams(_ptr(sums), _ptr(stats), n, c0 + c1, groups, h * w, eps,
                                                                   _ptr(dgamma), _ptr(dbeta), _stream()),
                                      "ddpm_gn_bwd_dparams"))
        self.launches += 2 if want_dp else 1
        return dx0, dx1

    # ---- attention core ----------------------------------------------------------------------------------
    NARROW_HEAD_DIMS = (8, 16, 32, 64)      # attention.cu (SIMT, one CTA per head); wider heads: bgemm.cu (tcgen05)

    def bgemm(self, a, lda, a_head, a_batch, a_mn, b, ldb, b_head, b_batch, b_mn, c, ldc, c_head, c_batch, m, n, k,
              heads, batch, alpha=1.0):
        """c[z] = alpha * a[z] (m x k) @ b[z] (k x n), z = (batch, head); a, b bf16 views, c bf16 or fp32."""
        _capi.check(self.lib.ddpm_bgemm(_ptr(a), lda, a_head, a_batch, int(a_mn), _ptr(b), ldb, b_head, b_batch,
                                        int(b_mn), _ptr(c), ldc, c_head, c_batch, int(c.dtype == torch.float32),
                                        m, n, k, heads, batch, float(alpha), _stream()), "ddpm_bgemm")
        self.launches += 1

    def _attn_wide_fwd(self, qkv, b, t, heads, d, scale, keep_probs: bool = True):
        C, ld, tp = heads * d, qkv.stride(0), (t + 7) // 8 * 8
        if _os.environ.get("DDPM_ATTN_FUSED", "1") != "0" and self.lib.ddpm_attn_wide_supported(t, heads, d):
            # ONE kernel: S in TMEM, softmax into shared memory, O = P V from there (attn_wide.cu); the probabilities
            # are written out only when a backward pass will read them
            o = torch.empty((b * t, C), device=qkv.device, dtype=torch.bfloat16)
            p = torch.empty((b, heads, t, tp), device=qkv.device, dtype=torch.bfloat16) if keep_probs else None
            _capi.check(self.lib.ddpm_attn_wide_fwd(_ptr(qkv), ld, _ptr(o), C, _ptr(p), tp, b, t, heads, d,
                                                    float(scale), _stream()), "ddpm_attn_wide_fwd")
            self.launches += 1
            return o, p
        q, k, v = qkv[:, :C], qkv[:, C:2 * C], qkv[:, 2 * C:3 * C]
        s = torch.empty((b, heads, t, tp), device=qkv.device, dtype=torch.float32)
        p = torch.empty((b, heads, t, tp), device=qkv.device, dtype=torch.bfloat16)
        o = torch.empty((b * t, C), device=qkv.device, dtype=torch.bfloat16)
        # S = scale * Q K^T
        self.bgemm(q, ld, d, t * ld, 0, k, ld, d, t * ld, 0, s, tp, t * tp, heads * t * tp, t, t, d, heads, b, scale)
        _capi.check(self.lib.ddpm_softmax_rows(_ptr(s), tp, _ptr(p), tp, b * heads * t, t, _stream()),
                    "ddpm_softmax_rows")
        # O = P V
        self.bgemm(p, tp, t * tp, heads * t * tp, 0, v, ld, d, t * ld, 1, o, C, d, t * C, t, d, t, heads, b)
        self.launches += 1
        return o, p

    def _attn_wide_bwd(self, qkv, d_o, p, b, t, heads, d, scale):
        C, ld, tp, ldg = heads * d, qkv.stride(0), p.shape[-1], d_o.stride(0)
        q, k, v = qkv[:, :C], qkv[:, C:2 * C], qkv[:, 2 * C:3 * C]
        dqkv = torch.empty_like(qkv)
        ldd = dqkv.stride(0)
        dq, dk, dv = dqkv[:, :C], dqkv[:, C:2 * C], dqkv[:, 2 * C:3 * C]
        dp = torch.empty((b, heads, t, tp), device=qkv.device, dtype=torch.float32)
        ds = torch.empty_like(p)
        ps, pb = t * tp, heads * t * tp
        # dP = dO V^T
        self.bgemm(d_o, ldg, d, t * ldg, 0, v, ld, d, t * ld, 0, dp, tp, ps, pb, t, t, d, heads, b)
        _capi.check(self.lib.ddpm_softmax_rows_bwd(_ptr(p), tp, _ptr(dp), tp, _ptr(ds), b * heads * t, t, float(scale),
                                                   _stream()), "ddpm_softmax_rows_bwd")
        self.launches += 1
        # dV = P^T dO,  dQ = dS K,  dK = dS^T Q   (dS carries the scale)
        self.bgemm(p, tp, ps, pb, 1, d_o, ldg, d, t * ldg, 1, dv, ldd, d, t * ldd, t, d, t, heads, b)
        self.bgemm(ds, tp, ps, pb, 0, k, ld, d, t * ld, 1, dq, ldd, d, t * ldd, t, d, t, heads, b)
        self.bgemm(ds, tp, ps, pb, 1, q, ld, d, t * ld, 1, dk, ldd, d, t * ldd, t, d, t, heads, b)
        return dqkv

    def attn_fwd(self, qkv, b: int, t: int, heads: int, d: int, scale: float, need_aux: bool = True):
        """qkv: bf16 [b*t, 3*heads*d] -> (o bf16 [b*t, heads*d], aux).  aux (kept for backward) is the fp32 log-sum-exp
        [b, heads, t] for narrow heads and the bf16 probabilities [b, heads, t, t8] for wide ones (None when
        need_aux is False and the fused wide-head kernel ran: inference)."""
        if d not in self.NARROW_HEAD_DIMS:
            if d % 8 or d < 64:
                raise NotImplementedError(f"attention head_dim={d}: supported are 8/16/32/64 and multiples of 8 above")
            return self._attn_wide_fwd(qkv, b, t, heads, d, scale, keep_probs=need_aux)
        o = torch.empty((b * t, heads * d), device=qkv.device, dtype=torch.bfloat16)
        lse = torch.empty((b, heads, t), device=qkv.device, dtype=torch.float32)
        _capi.check(self.lib.ddpm_attn_fwd(_ptr(qkv), qkv.stride(0), _ptr(o), o.stride(0), _ptr(lse), b, t, heads, d,
                                           scale, _stream()), "ddpm_attn_fwd")
        self.launches += 1
        return o, lse

    def attn_bwd(self, qkv, o, d_o, lse, b: int, t: int, heads: int, d: int, scale: float):
        if d not in self.NARROW_HEAD_DIMS:
            return self._attn_wide_bwd(qkv, d_o, lse, b, t, heads, d, scale)
        dqkv = torch.empty_like(qkv)
        _capi.check(self.lib.ddpm_attn_bwd(_ptr(qkv), qkv.stride(0), _ptr(o), o.stride(0), _ptr(d_o), d_o.stride(0),
                                           _ptr(lse), _ptr(dqkv), dqkv.stride(0), b, t, heads, d, scale, _stream()),
                    "ddpm_attn_bwd")
        self.launches += 1
        return dqkv

    # ---- time embedding path ---------------------------------------------------------------------------------
    def timestep_embedding(self, t, dim: int, flip_sin_to_cos: bool, freq_shift: float):
        key = (dim, float(freq_shift), str(t.device))
        freqs = self._freqs.get(key)
        if freqs is None:
            freqs = timestep_freqs(dim, freq_shift).to(t.device)
            self._freqs[key] = freqs
        out = torch.empty((t.numel(), dim), device=t.device, dtype=torch.float32)
        _capi.check(self.lib.ddpm_timestep_embedding(_ptr(t), _ptr(freqs), _ptr(out), t.numel(), dim,
                                                     int(flip_sin_to_cos), _stream()), "ddpm_timestep_embedding")
        self.launches += 1
        return out

    def linear_f32(self, x, w, bias, silu_in: bool):
        m, k = x.shape
        n = w.shape[0]
        y = torch.empty((m, n), device=x.device, dtype=torch.float32)
        _capi.check(self.lib.ddpm_linear_f32(_ptr(x), _ptr(w), _ptr(bias), _ptr(y), m, n, k, int(silu_in), _stream()),
                    "ddpm_linear_f32")
        self.launches += 1
        return y

    def linear_f32_wgrad(self, x, dy, dw, db, silu_in: bool):
        m, k = x.shape
        n = dy.shape[1]
        _capi.check(self.lib.ddpm_linear_f32_wgrad(_ptr(x), _ptr(dy), _ptr(dw), _ptr(db), m, n, k, int(silu_in),
                                                   _stream()), "ddpm_linear_f32_wgrad")
        self.launches += 1

    def linear_f32_dgrad(self, dy, w, x, silu_in: bool, dx=None):
        m, n = dy.shape
        k = w.shape[1]
        acc = dx is not None
        if dx is None:
            dx = torch.empty((m, k), device=dy.device, dtype=torch.float32)
        _capi.check(self.lib.ddpm_linear_f32_dgrad(_ptr(dy), _ptr(w), _ptr(x), _ptr(dx), m, n, k, int(silu_in),
                                                   int(acc), _stream()), "ddpm_linear_f32_dgrad")
        self.launches += 1
        return dx

    def reduce_hw(self, x, out_nc=None, out_c=None):
        """out_nc[n, c] = sum_hw x (overwritten); out_c[c] += sum_{n,hw} x."""
        n, h, w, c, ld = _nhwc(x, "x")
        _capi.check(self.lib.ddpm_reduce_hw(_ptr(x), ld, n, h * w, c, _ptr(out_nc),
                                            out_nc.stride(0) if out_nc is not None else 0, _ptr(out_c), _stream()),
                    "ddpm_reduce_hw")
        self.launches += 1

    def dropout(self, x, p: float, seed: int, offset: int, add=None, tick=None):
        """out = (add or 0) + x * keep/(1-p); keep is a pure function of (seed, offset, tick, index).  bf16, contiguous.
        tick: optional int64 device scalar (step counter) so CUDA-graph replays draw fresh masks."""
        if not x.is_contiguous() or (add is not None and not add.is_contiguous()):
            raise ValueError("dropout needs contiguous tensors")
        out = torch.empty_like(x)
        _capi.check(self.lib.ddpm_dropout(_ptr(x), _ptr(add), _ptr(out), x.numel(), float(p), seed, offset, _ptr(tick),
                                          _stream()),
                    "ddpm_dropout")
        self.launches += 1
        return out

    # ---- input transform (csrc/preprocess.cu) ------------------------------------------------------------------
    def preprocess_u8(self, frames, out_h: int, out_w: int, htab, vtab, flips):
        """uint8 [B, H, W, C] -> fp32 [B, C, out_h, out_w] in [-1, 1]: Pillow bilinear resize (two 8-bit passes), flip,
        /255, (x - 0.5)/0.5.  htab / vtab = (bounds, coeffs, ksize) device tables for the two passes."""
        b, h, w, c = frames.shape
        mid = frames
        if w != out_w:
            mid = torch.empty((b, h, out_w, c), device=frames.device, dtype=torch.uint8)
            _capi.check(self.lib.ddpm_resize_h_u8(_ptr(frames), _ptr(mid), b * h, w, c, out_w, _ptr(htab[0]),
                                                  _ptr(htab[1]), htab[2], _stream()), "ddpm_resize_h_u8")
            self.launches += 1
        out = torch.empty((b, c, out_h, out_w), device=frames.device, dtype=torch.float32)
        _capi.check(self.lib.ddpm_resize_v_normalize(_ptr(mid), _ptr(out), b, h, out_w, c, out_h, _ptr(vtab[0]),
                                                     _ptr(vtab[1]), vtab[2], _ptr(flips), _stream()),
                    "ddpm_resize_v_normalize")
        self.launches += 1
        return out

    # ---- layout helpers ----------------------------------------------------------------------------------------
    def space_to_depth(self, x):
        n, h, w, c, ld = _nhwc(x, "x")
        out = torch.empty((4 * n, (h + 1) // 2, (w + 1) // 2, c), device=x.device, dtype=torch.bfloat16)
        _capi.check(self.lib.ddpm_space_to_depth(_ptr(x), ld, _ptr(out), n, h, w, c, 0, _stream()),
                    "ddpm_space_to_depth")
        self.launches += 1
        return out

    def zero_insert2x(self, dy, h: int, w: int):
        n, ho, wo, c, ld = _nhwc(dy, "dy")
        out = torch.empty((n, h, w, c), device=dy.device, dtype=torch.bfloat16)
        _capi.check(self.lib.ddpm_zero_insert2x(_ptr(dy), ld, _ptr(out), n, ho, wo, c, h, w, _stream()),
                    "ddpm_zero_insert2x")
        self.launches += 1
        out._ddpm_alg_pixels_div = 4       # 75 % structural zeros (FLOP accounting only)
        return out

    def upsample2x(self, x):
        n, h, w, c, ld = _nhwc(x, "x")
        out = torch.empty((n, 2 * h, 2 * w, c), device=x.device, dtype=torch.bfloat16)
        _capi.check(self.lib.ddpm_upsample2x(_ptr(x), ld, _ptr(out), n, h, w, c, _stream()), "ddpm_upsample2x")
        self.launches += 1
        return out

    def sumpool2x(self, dy, add=None):
        n, h2, w2, c, ld = _nhwc(dy, "dy")
        out = torch.empty((n, h2 // 2, w2 // 2, c), device=dy.device, dtype=torch.bfloat16)
        _capi.check(self.lib.ddpm_sumpool2x(_ptr(dy), ld, _ptr(add), _nhwc(add, "add")[4] if add is not None else 0,
                                            _ptr(out), n, h2 // 2, w2 // 2, c, _stream()), "ddpm_sumpool2x")
        self.launches += 1
        return out


def algorithmic_conv_flops(x0, x1, taps, cout, grid, out_f32: bool = False) -> float:
    """ALGORITHMIC FLOPs of one conv_gemm / conv_wgrad launch (2 * B * Ho * Wo * Cout * Cin * k^2 of the
    convolution the launch implements), not the FLOPs the tensor cores execute:
      * the stride-2 dgrad runs as a stride-1 correlation over a zero-inserted gradient (4x the pixels, 75 % zeros):
        tensors produced by zero_insert2x carry `_ddpm_alg_pixels_div = 4`;
      * the 3-channel boundary convs run on padded operands: im2col3 patches (64 columns, 9 * cin used) carry
        `_ddpm_alg_k`, and conv_out's 32 padded output columns count as its real `out_channels` (tagged on the weight
        operand as `_ddpm_alg_cout`)."""
    n, h, w = grid
    cin = x0.shape[-1] + (x1.shape[-1] if x1 is not None else 0)
    k = getattr(x0, "_ddpm_alg_k", None)
    kk = float(k) if k is not None else float(cin * len(taps))
    pix = float(n) * h * w / float(getattr(x0, "_ddpm_alg_pixels_div", 1))
    return 2.0 * pix * float(cout) * kk


class OpProfiler:
    """CUDA-event timing of every C-ABI op, recorded on the launching stream (bench.py roofline, profiles/).

        prof = OpProfiler(ops.get()); prof.start(); ...; table = prof.stop()
    `table[name] = {"calls": n, "ms": total_ms, "flops": algorithmic_flops, "bytes": algorithmic_bytes}`.
    """

    def __init__(self, backend: "CudaOps"):
        self.b = backend
        self._orig = {}
        self.records = []          # (name, e0, e1, flops, bytes)

    @staticmethod
    def _algorithmic(name, args, kwargs, out):
        """Algorithmic FLOPs / bytes of one call; 0 when not modelled."""
        try:
            if name == "conv_gemm":
                x0, x1, taps, wgt, cout, grid = args[:6]
                return algorithmic_conv_flops(x0, x1, taps, getattr(wgt, "_ddpm_alg_cout", cout), grid), 0.0
            if name == "conv_wgrad":
                dy, x0, x1, taps, dw, grid = args[:6]
                return algorithmic_conv_flops(x0, x1, taps, getattr(dy, "_ddpm_alg_k", dy.shape[-1]), grid), 0.0
            if name == "gn_apply":
                return 0.0, 4.0 * out.numel()                 # bf16 read + bf16 write
            if name == "gn_fwd":
                return 0.0, 4.0 * out[1].numel()              # bf16 read (second read is an L2 hit) + bf16 write
            if name == "gn_stats":
                x0, x1 = args[0], args[1]
                return 0.0, 2.0 * (x0.numel() + (x1.numel() if x1 is not None else 0))
            if name == "gn_fwd_from_csum":
                return 0.0, 4.0 * out[1].numel()              # one streaming pass: bf16 read + bf16 write
            if name in ("gn_bwd", "gn_bwd_apply"):
                x0, x1 = args[0], args[1]
                extra = sum(kwargs[k].numel() for k in ("add0", "add1") if kwargs.get(k) is not None)
                return 0.0, 6.0 * (x0.numel() + (x1.numel() if x1 is not None else 0)) + 2.0 * extra  # x, dy, dx (+addends)
            if name == "adamw_flat":
                return 0.0, 28.0 * args[0].numel()            # p, g, m, v read; p, m, v written (fp32)
            if name == "sumsq":
                return 0.0, 4.0 * args[0].numel()
            if name in ("add_noise", "mse_fwd_bwd"):
                return 0.0, 12.0 * args[0].numel()
            if name == "scheduler_step":
                return 0.0, (16.0 if args[2] is not None else 12.0) * args[1].numel()
            if name == "scheduler_step_philox":
                return 0.0, 12.0 * args[1].numel()
        except Exception:  # noqa: BLE001
            pass
        return 0.0, 0.0

    def _detail(self, name, args):
        """Optional per-shape key ("conv_gemm|h128 c128->128 k9") when self.by_shape is set."""
        if not getattr(self, "by_shape", False):
            return ""
        try:
            if name == "conv_gemm":
                x0, x1, taps, wgt, cout, grid = args[:6]
                cin = x0.shape[-1] + (x1.shape[-1] if x1 is not None else 0)
                return f"|{grid[1]}x{grid[2]} c{cin}->{cout} k{len(taps)}"
            if name == "conv_wgrad":
                dy, x0, x1, taps, dw, grid = args[:6]
                cin = x0.shape[-1] + (x1.shape[-1] if x1 is not None else 0)
                return f"|{grid[1]}x{grid[2]} c{cin}->{dy.shape[-1]} k{len(taps)}"
            if name in ("gn_bwd", "gn_bwd_apply", "gn_apply", "gn_stats", "gn_fwd", "gn_fwd_from_csum"):
                x0, x1 = args[0], args[1]
                c = x0.shape[-1] + (x1.shape[-1] if x1 is not None else 0)
                return f"|{x0.shape[1]}x{x0.shape[2]} c{c}"
        except Exception:  # noqa: BLE001
            pass
        return ""

    def start(self):
        for name in dir(self.b):
            fn = getattr(self.b, name)
            if name.startswith("_") or not callable(fn) or name in ("lib",):
                continue
            self._orig[name] = fn

            def wrapped(*a, __fn=fn, __name=name, **k):
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                out = __fn(*a, **k)
                e1.record()
                fl, by = self._algorithmic(__name, a, k, out)
                self.records.append((__name + self._detail(__name, a), e0, e1, fl, by))
                return out

            setattr(self.b, name, wrapped)
        return self

    def stop(self):
        for name in self._orig:
            try:
                delattr(self.b, name)      # drop the instance attribute -> class method is visible again
            except AttributeError:
                pass
        self._orig = {}
        torch.cuda.synchronize()
        table = {}
        for name, e0, e1, fl, by in self.records:
            t = table.setdefault(name, {"calls": 0, "ms": 0.0, "flops": 0.0, "bytes": 0.0})
            t["calls"] += 1
            t["ms"] += e0.elapsed_time(e1)
            t["flops"] += fl
            t["bytes"] += by
        self.records = []
        return table


_backend = None


def get():
    """The active op backend.  Product code only ever gets CudaOps; tests may inject a checker via set_backend."""
    global _backend
    if _backend is None:
        _backend = CudaOps()
    return _backend


def set_backend(b) -> None:
    """Test hook (tests/emu_ops.py verifies the host-side graph wiring on CPU).  Never called by product code."""
    global _backend
    _backend = b
