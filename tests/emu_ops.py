"""CPU emulation of the op contracts in include/ddpm_b200.h -- TEST INFRASTRUCTURE ONLY.

Implements every method of polyp_image_generator_b200.ops.CudaOps with plain torch on the CPU (fp32 by default,
optionally rounding activations to bf16 like the kernels do).  It exists so the `-m "not gpu"` suite can verify the
HOST LOGIC of the product (the UNet forward/backward programs, tap tables, arena layout, LoRA wiring, DDP bucketing)
against the oracle without a GPU.  It is injected explicitly with ops.set_backend(EmuOps()) by tests; the product
never imports this file and never falls back to it.
"""
from __future__ import annotations

import torch
import torch.nn.functional as F


class EmuOps:
    name = "emu"

    def __init__(self, act_dtype=torch.float32):
        self.act = act_dtype           # dtype of "bf16" activation tensors
        self.operand_dtype = act_dtype  # dtype of the tensor-core operand arena
        self.launches = 0

    def _a(self, t):
        return t.to(self.act)

    # ---- scheduler / loss -------------------------------------------------------------------------------
    def add_noise(self, x0, noise, t, sa, sb):
        shape = (-1,) + (1,) * (x0.dim() - 1)
        return sa[t].view(shape) * x0 + sb[t].view(shape) * noise

    def mse_fwd_bwd(self, pred, target, want_grad=True):
        d = pred - target
        return (d * d).sum().reshape(1), (d * (2.0 / pred.numel())) if want_grad else None

    def scale_by_device_scalar(self, x, scale):
        x.mul_(scale.reshape(()))
        return x

    def ddim_step(self, eps, x, z, sa, sb, sap, dirc, sigma, clip, use_clipped, want_x0=False):
        f = lambda v: torch.tensor(v, dtype=torch.float32)
        x0 = (x - f(sb) * eps) / f(sa)
        if clip > 0:
            x0 = x0.clamp(-clip, clip)
        pe = (x - f(sa) * x0) / f(sb) if use_clipped else eps
        prev = f(sap) * x0 + f(dirc) * pe
        if z is not None:
            prev = prev + f(sigma) * z
        return prev, (x0 if want_x0 else None)

    def unipc_x0(self, eps, x, sigma_t, alpha_t):
        f = lambda v: torch.tensor(v, dtype=torch.float32)
        return (x - f(sigma_t) * eps) / f(alpha_t)

    def unipc_update(self, x, m0, m1, mt, cx, cm, cb, rk=1.0, rho0=0.0, rho_t=0.0):
        f = lambda v: torch.tensor(v, dtype=torch.float32)
        base = f(cx) * x - f(cm) * m0
        res = None
        if m1 is not None:
            res = f(rho0) * ((m1 - m0) / f(rk))
        if mt is not None:
            d = f(rho_t) * (mt - m0)
            res = d if res is None else res + d
        return base if res is None else base - f(cb) * res

    def scheduler_step(self, eps, x, z, sa, sb, c0, ct, sigma, clip, want_x0=False):
        f = lambda v: torch.tensor(v, dtype=torch.float32)
        x0 = (x - f(sb) * eps) / f(sa)
        if clip > 0:
            x0 = x0.clamp(-clip, clip)
        prev = f(c0) * x0 + f(ct) * x
        if z is not None:
            prev = prev + f(sigma) * z
        return prev, (x0 if want_x0 else None)

    def scheduler_step_philox(self, eps, x, sa, sb, c0, ct, sigma, clip, seed, offset):
        g = torch.Generator().manual_seed((seed + offset) & 0x7FFFFFFF)
        z = torch.randn(x.shape, generator=g) if sigma != 0 else None
        return self.scheduler_step(eps, x, z, sa, sb, c0, ct, sigma, clip)[0]

    def preprocess_u8(self, frames, out_h, out_w, htab, vtab, flips):
        def one_pass(x, tab, axis):            # x int64 [..]; gather-accumulate along `axis` with the integer tables
            bounds, coeffs, ksize = tab
            outs = []
            for o in range(bounds.shape[0]):
                lo, n = int(bounds[o, 0]), int(bounds[o, 1])
                acc = torch.full_like(x.select(axis, 0), 1 << 21)
                for i in range(n):
                    acc = acc + x.select(axis, lo + i) * int(coeffs[o, i])
                outs.append((acc >> 22).clamp(0, 255))
            return torch.stack(outs, axis)
        x = frames.to(torch.int64)
        if frames.shape[2] != out_w:
            x = one_pass(x, htab, 2)
        x = one_pass(x, vtab, 1)
        if flips is not None:
            x = torch.where(flips.bool().view(-1, 1, 1, 1), x.flip(2), x)
        t = x.permute(0, 3, 1, 2).contiguous().to(torch.float32).div(255)
        return (t - 0.5) / 0.5

    def to_uint8_nhwc(self, x):
        return ((x / 2 + 0.5).clamp(0, 1).permute(0, 2, 3, 1) * 255).round().to(torch.uint8)

    # ---- conv / linear -------------------------------------------------------------------------------------
    @staticmethod
    def _shift(X, dn, dh, dw, n, h, w):
        """out[i,j,k] = X[i+dn, j+dh, k+dw] (zero outside X)."""
        SN, H, W, C = X.shape
        out = X.new_zeros((n, h, w, C))
        i0, i1 = max(0, -dn), min(n, SN - dn)
        j0, j1 = max(0, -dh), min(h, H - dh)
        k0, k1 = max(0, -dw), min(w, W - dw)
        if i1 > i0 and j1 > j0 and k1 > k0:
            out[i0:i1, j0:j1, k0:k1] = X[i0 + dn:i1 + dn, j0 + dh:j1 + dh, k0 + dw:k1 + dw]
        return out

    @staticmethod
    def gn_bwd_fusable(grid):
        hw = grid[1] * grid[2]
        return hw >= 128 or hw % 32 == 0      # the CPU emulation exercises the fused path wherever it is POSSIBLE

    @staticmethod
    def gn_stats_fusable(grid):
        hw = grid[1] * grid[2]
        return hw >= 128 or hw % 32 == 0      # exercised wherever it is POSSIBLE

    def gn_fwd_from_csum(self, x0, x1, cs0, cs1, groups, eps, gamma, beta, silu, want_coef=False):
        cs = cs0 if cs1 is None else torch.cat([cs0, cs1], 1)          # [n, C / 4, 2]: 4-channel granules
        n, Q, _ = cs.shape
        C = 4 * Q
        stats = cs.reshape(n, groups, Q // groups, 2).sum(2)
        # the fused statistics must agree with a direct reduction over the tensor itself
        X = (x0 if x1 is None else torch.cat([x0, x1], -1)).float()
        want = torch.stack([X.reshape(n, -1, groups, C // groups).sum((1, 3)),
                            (X ** 2).reshape(n, -1, groups, C // groups).sum((1, 3))], -1)
        assert torch.allclose(stats, want, rtol=1e-4, atol=1e-3), "conv-epilogue statistics disagree with the tensor"
        r = self.gn_fwd(x0, x1, groups, eps, gamma, beta, silu, want_coef=want_coef)
        return r

    # ---- fp32-faithful mode: split tensors [.., 2C] = [hi | lo], value = hi + lo (the emulation splits at bf16) -------
    @staticmethod
    def _to_split(v):
        hi = v.to(torch.bfloat16).float()
        lo = (v - hi).to(torch.bfloat16).float()
        return torch.cat([hi, lo], -1)

    @staticmethod
    def _from_split(t):
        c = t.shape[-1] // 2
        return t[..., :c].float() + t[..., c:].float()

    def im2col3_split(self, x):
        n, cin, h, w = x.shape
        pat = F.unfold(x.float(), 3, padding=1).view(n, cin, 9, h, w).permute(0, 3, 4, 2, 1).reshape(n, h, w, 9 * cin)
        sp = self._to_split(pat)
        out = torch.zeros((n, h, w, 128), dtype=torch.float32)
        k = 9 * cin
        out[..., :k], out[..., 32:32 + k], out[..., 64:64 + k] = sp[..., :k], sp[..., k:], sp[..., :k]
        return out

    def gn_fwd_split(self, x0, x1, groups, eps, gamma, beta, silu):
        x = self._from_split(x0)
        if x1 is not None:
            x = torch.cat([x, self._from_split(x1)], -1)
        y = F.group_norm(x.permute(0, 3, 1, 2), groups, gamma, beta, eps)
        if silu:
            y = F.silu(y)
        return self._to_split(y.permute(0, 2, 3, 1).contiguous())

    def attn_fwd_split(self, qkv, b, t, heads, d, scale):
        C = heads * d
        x = self._from_split(qkv)
        q, k, v = [z.reshape(b, t, heads, d).transpose(1, 2) for z in x.split(C, dim=1)]
        o = F.scaled_dot_product_attention(q, k, v, scale=scale).transpose(1, 2).reshape(b * t, C)
        return self._to_split(o)

    def conv_gemm(self, x0, x1, taps, wgt, cout, grid, bias=None, temb=None, res=None, out=None, out_f32=False,
                  src_n=0, gn=None, csum=None, split_io=False):
        if split_io:
            assert gn is None and csum is None and not out_f32 and out is None
            acc = self.conv_gemm(x0, x1, taps, wgt, cout, grid, bias=bias, temb=temb, out_f32=True, src_n=src_n)
            if res is not None:
                acc = acc + self._from_split(res).reshape(acc.shape)
            return self._to_split(acc)
        n, h, w = grid
        X = x0 if x1 is None else torch.cat([x0, x1], -1)
        X = X.float()
        C = X.shape[-1]
        acc = torch.zeros((n, h, w, cout), dtype=torch.float32)
        Wm = wgt.float()
        for (dn, dh, dw, wk) in taps:
            acc += self._shift(X, dn, dh, dw, n, h, w) @ Wm[:cout, wk:wk + C].t()
        if bias is not None:
            acc += bias
        if temb is not None:
            acc += temb[:, None, None, :cout]
        if res is not None:
            acc += res.float()
        if gn is not None:
            gx0, gx1, gcoef, gsilu, gsums = gn
            Xg = (gx0 if gx1 is None else torch.cat([gx0, gx1], -1)).float()
            if gsilu:
                Cg = Xg.shape[-1]
                ka = torch.stack([gcoef[..., 0], gcoef[..., 1]], -1).reshape(-1, Cg)
                kb = torch.stack([gcoef[..., 2], gcoef[..., 3]], -1).reshape(-1, Cg)
                with torch.enable_grad():
                    z = (Xg * ka[:, None, None, :] + kb[:, None, None, :]).detach().requires_grad_(True)
                    F.silu(z).backward(acc)
                acc = z.grad
            dzr = self._a(acc).float()
            gsums[..., 0] += dzr.sum((1, 2))
            gsums[..., 1] += (dzr * Xg).sum((1, 2))
        r = acc if out_f32 else self._a(acc)
        if csum is not None:
            rf = r.float()
            csum[..., 0] += rf.sum((1, 2)).reshape(csum.shape[0], -1, 4).sum(-1)
            csum[..., 1] += (rf * rf).sum((1, 2)).reshape(csum.shape[0], -1, 4).sum(-1)
        if out is not None:
            out.copy_(r)
            return out
        return r

    def conv_wgrad(self, dy, x0, x1, taps, dw, grid, accumulate=True, src_n=0, splits=0, dbias=None):
        if dbias is not None:
            dbias += dy.float().reshape(-1, dy.shape[-1]).sum(0)
        n, h, w = grid
        X = (x0 if x1 is None else torch.cat([x0, x1], -1)).float()
        C = X.shape[-1]
        dyf = dy.float().reshape(-1, dy.shape[-1])
        if not accumulate:
            dw.zero_()
        for (dn, dh, dw_, wk) in taps:
            xs = self._shift(X, dn, dh, dw_, n, h, w).reshape(-1, C)
            dw[:, wk:wk + C] += dyf.t() @ xs
        return dw

    def prep_weight(self, w, wf, wd, cout, taps, cin):
        w3 = w.reshape(cout, taps, cin)
        if wf is not None:
            wf.copy_(w3.reshape(cout, taps * cin).to(wf.dtype))
        if wd is not None:
            wd.copy_(w3.flip(1).permute(2, 1, 0).reshape(cin, taps * cout).to(wd.dtype))

    # ---- 3-channel convs -------------------------------------------------------------------------------------
    @staticmethod
    def _w_from_strides(w, strides, flip, cout, cin):
        """dense [cout][9][cin] view of w[co*s0 + tap'*s1 + ci*s2], tap' = 8-tap when flip."""
        flat = w.reshape(-1)
        co = torch.arange(cout).view(-1, 1, 1)
        tap = torch.arange(9).view(1, -1, 1)
        ci = torch.arange(cin).view(1, 1, -1)
        tp = 8 - tap if flip else tap
        return flat[co * strides[0] + tp * strides[1] + ci * strides[2]]

    def sumsq(self, x, out):
        out += (x.float() ** 2).sum()

    def adamw_flat(self, p, g, m, v, scal, gnorm_sq, max_norm, lr_dev, lr, beta1, beta2, eps, weight_decay):
        scal[0] += 1
        step = float(scal[0])
        coef = 1.0
        if gnorm_sq is not None and max_norm and max_norm > 0:
            coef = min(1.0, max_norm / (float(gnorm_sq.sqrt()) + 1e-6))
        lr_ = float(lr_dev) if lr_dev is not None else lr
        gr = g * coef
        p.mul_(1 - lr_ * weight_decay)
        m.lerp_(gr, 1 - beta1)
        v.mul_(beta2).addcmul_(gr, gr, value=1 - beta2)
        bc1, bc2 = 1 - beta1 ** step, 1 - beta2 ** step
        p.addcdiv_(m, (v.sqrt() / (bc2 ** 0.5)).add_(eps), value=-lr_ / bc1)

    def im2col3(self, x, chan_sum=None):
        n, cin, h, w = x.shape
        xp = F.pad(x.float(), (1, 1, 1, 1))
        cols = [xp[:, :, r:r + h, s:s + w] for r in range(3) for s in range(3)]        # tap-major
        pat = torch.stack(cols, 1).permute(0, 3, 4, 1, 2).reshape(n, h, w, 9 * cin)  # [.., tap*cin + k]
        out = torch.zeros((n, h, w, 64), dtype=torch.float32)
        out[..., :9 * cin] = pat
        if chan_sum is not None:
            chan_sum += x.float().sum((0, 2, 3))
        return self._a(out)

    def nhwc_to_nchw_f32(self, src, cout):
        return src[..., :cout].permute(0, 3, 1, 2).contiguous().float()

    def gn_stats(self, x0, x1, groups):
        X = (x0 if x1 is None else torch.cat([x0, x1], -1)).float()
        n, h, w, C = X.shape
        xg = X.reshape(n, h * w, groups, C // groups)
        return torch.stack([xg.sum((1, 3)), (xg * xg).sum((1, 3))], -1)   # [n, groups, 2]

    @staticmethod
    def _mean_rstd(stats, m, eps):
        mean = stats[..., 0] / m
        var = (stats[..., 1] / m - mean * mean).clamp_min(0)
        return mean, torch.rsqrt(var + eps)

    def _gn_fwd_f32(self, X, groups, stats, eps, gamma, beta, silu):
        n, h, w, C = X.shape
        cpg = C // groups
        mean, rstd = self._mean_rstd(stats, cpg * h * w, eps)
        mean = mean.repeat_interleave(cpg, 1)[:, None, None, :]
        rstd = rstd.repeat_interleave(cpg, 1)[:, None, None, :]
        z = (X - mean) * rstd * gamma + beta
        return F.silu(z) if silu else z

    def gn_apply(self, x0, x1, groups, stats, eps, gamma, beta, silu, out=None):
        X = (x0 if x1 is None else torch.cat([x0, x1], -1)).float()
        r = self._a(self._gn_fwd_f32(X, groups, stats, eps, gamma, beta, silu))
        if out is not None:
            out.copy_(r)
            return out
        return r

    def gn_fwd(self, x0, x1, groups, eps, gamma, beta, silu, out=None, want_coef=False):
        stats = self.gn_stats(x0, x1, groups)
        y = self.gn_apply(x0, x1, groups, stats, eps, gamma, beta, silu, out=out)
        if not want_coef:
            return stats, y
        n, h, w, C = y.shape
        cpg = C // groups
        mean, rstd = self._mean_rstd(stats, cpg * h * w, eps)
        ka = rstd.repeat_interleave(cpg, 1) * gamma                     # [n, C]
        kb = beta - mean.repeat_interleave(cpg, 1) * ka
        coef = torch.stack([ka[:, 0::2], ka[:, 1::2], kb[:, 0::2], kb[:, 1::2]], -1).contiguous()   # [n, C/2, 4]
        return stats, y, coef

    def gn_bwd(self, x0, x1, groups, stats, eps, gamma, beta, silu, dy, add0=None, add1=None, dgamma=None,
               dbeta=None, need_dx1=True):
        X = (x0 if x1 is None else torch.cat([x0, x1], -1)).float()
        c0 = x0.shape[-1]
        with torch.enable_grad():
            Xr = X.detach().permute(0, 3, 1, 2).contiguous().requires_grad_(True)
            g = gamma.detach().clone().requires_grad_(True)
            b = beta.detach().clone().requires_grad_(True)
            y = F.group_norm(Xr, groups, g, b, eps)
            if silu:
                y = F.silu(y)
            y.backward(dy.float().permute(0, 3, 1, 2))
        dx = Xr.grad.permute(0, 2, 3, 1)
        if add0 is not None:
            dx = dx + add0.float()
        if add1 is not None:
            dx = dx + add1.float()
        if dgamma is not None:
            dgamma += g.grad
        if dbeta is not None:
            dbeta += b.grad
        dx = self._a(dx)
        dx0 = dx[..., :c0].contiguous()
        dx1 = dx[..., c0:].contiguous() if (x1 is not None and need_dx1) else None
        return dx0, dx1

    def gn_bwd_apply(self, x0, x1, groups, stats, eps, gamma, dz, sums, add0=None, add1=None, dgamma=None,
                     dbeta=None, need_dx1=True, out_nc=None, out_c=None):
        X = (x0 if x1 is None else torch.cat([x0, x1], -1)).float()
        n, h, w, C = X.shape
        c0 = x0.shape[-1]
        cpg = C // groups
        m = cpg * h * w
        mean, rstd = self._mean_rstd(stats, m, eps)
        mean_c = mean.repeat_interleave(cpg, 1)
        rstd_c = rstd.repeat_interleave(cpg, 1)
        S1 = sums[..., 0]
        S2h = rstd_c * (sums[..., 1] - mean_c * S1)                       # sum dz * xhat
        s1 = (gamma * S1).reshape(n, groups, cpg).sum(-1).repeat_interleave(cpg, 1) / m
        s2 = (gamma * S2h).reshape(n, groups, cpg).sum(-1).repeat_interleave(cpg, 1) / m
        xh = (X - mean_c[:, None, None, :]) * rstd_c[:, None, None, :]
        dx = (dz.float() * gamma - s1[:, None, None, :] - xh * s2[:, None, None, :]) * rstd_c[:, None, None, :]
        if add0 is not None:
            dx = dx + add0.float()
        if add1 is not None:
            dx = dx + add1.float()
        if dgamma is not None:
            dgamma += S2h.sum(0)
        if dbeta is not None:
            dbeta += S1.sum(0)
        if out_nc is not None:
            out_nc += dx.sum((1, 2))
        if out_c is not None:
            out_c += dx.sum((0, 1, 2))
        dx = self._a(dx)
        dx0 = dx[..., :c0].contiguous()
        dx1 = dx[..., c0:].contiguous() if (x1 is not None and need_dx1) else None
        return dx0, dx1

    # ---- attention ----------------------------------------------------------------------------------------------
    def _split(self, qkv, b, t, heads, d):
        C = heads * d
        return [z.reshape(b, t, heads, d).transpose(1, 2) for z in qkv.float().split(C, 1)]

    def attn_fwd(self, qkv, b, t, heads, d, scale, need_aux=True):
        q, k, v = self._split(qkv, b, t, heads, d)
        s = (q @ k.transpose(-1, -2)) * scale
        lse = torch.logsumexp(s, -1) * 1.4426950408889634
        o = torch.softmax(s, -1) @ v
        return self._a(o.transpose(1, 2).reshape(b * t, heads * d)), lse

    def attn_bwd(self, qkv, o, d_o, lse, b, t, heads, d, scale):
        with torch.enable_grad():
            qr = qkv.detach().float().clone().requires_grad_(True)
            q, k, v = self._split(qr, b, t, heads, d)
            out = F.scaled_dot_product_attention(q, k, v, scale=scale).transpose(1, 2).reshape(b * t, heads * d)
            out.backward(d_o.float())
        return self._a(qr.grad)

    # ---- time embedding path ---------------------------------------------------------------------------------------
    def timestep_embedding(self, t, dim, flip_sin_to_cos, freq_shift):
        from polyp_image_generator_b200.ops import timestep_freqs
        arg = t[:, None].float() * timestep_freqs(dim, freq_shift)[None, :]
        s, c = torch.sin(arg), torch.cos(arg)
        return torch.cat([c, s], -1) if flip_sin_to_cos else torch.cat([s, c], -1)

    def linear_f32(self, x, w, bias, silu_in):
        return F.linear(F.silu(x) if silu_in else x, w, bias)

    def linear_f32_wgrad(self, x, dy, dw, db, silu_in):
        xa = F.silu(x) if silu_in else x
        dw += dy.t() @ xa
        if db is not None:
            db += dy.sum(0)

    def linear_f32_dgrad(self, dy, w, x, silu_in, dx=None):
        r = dy @ w
        if silu_in:
            s = torch.sigmoid(x)
            r = r * (s * (1 + x * (1 - s)))
        if dx is not None:
            dx += r
            return dx
        return r

    def reduce_hw(self, x, out_nc=None, out_c=None):
        s = x.float().sum((1, 2))
        if out_nc is not None:
            out_nc.copy_(s)
        if out_c is not None:
            out_c += s.sum(0)

    def dropout(self, x, p, seed, offset, add=None, tick=None):
        if tick is not None:
            offset += int(tick) << 32
        g = torch.Generator().manual_seed((seed * 1000003 + offset) & 0x7FFFFFFFFFFF)
        keep = (torch.rand(x.shape, generator=g) >= p).float() / (1.0 - p)
        r = x.float() * keep
        if add is not None:
            r = r + add.float()
        return self._a(r)

    # ---- layout helpers -----------------------------------------------------------------------------------------------
    def space_to_depth(self, x):
        n, h, w, c = x.shape
        return torch.cat([x[:, ph::2, pw::2] for ph in (0, 1) for pw in (0, 1)], 0).contiguous()

    def zero_insert2x(self, dy, h, w):
        n, ho, wo, c = dy.shape
        out = dy.new_zeros((n, h, w, c))
        out[:, 0:2 * ho:2, 0:2 * wo:2] = dy
        return out

    def upsample2x(self, x):
        return x.repeat_interleave(2, 1).repeat_interleave(2, 2)

    def sumpool2x(self, dy, add=None):
        n, h2, w2, c = dy.shape
        r = dy.float().reshape(n, h2 // 2, 2, w2 // 2, 2, c).sum((2, 4))
        if add is not None:
            r = r + add.float()
        return self._a(r)
