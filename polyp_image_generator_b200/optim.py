"""Fused global-norm clip + AdamW for the B200 UNet2DModel.

Call sites replaced (paths relative to /root/reference/generator_model/):
  train_from_scratch.py:273      optimizer = torch.optim.AdamW(model.parameters(), lr=config.learning_rate)
  train_from_scratch.py:106-108  clip_grad_norm_(model.parameters(), 1.0); optimizer.step()

The UNet keeps every parameter in ONE flat fp32 arena and produces every gradient in ONE flat fp32 arena (unet.py),
so the whole update is two streaming kernels over 113.7 M elements -- a sum of squares for the global norm and one
AdamW pass that applies the clip coefficient on the fly (28 B/elem) -- instead of ~30 multi-tensor launches.
Arithmetic and op order are torch.optim.AdamW's (decoupled weight decay, lerp first moment, bias-corrected step).

    opt = FusedAdamW(model.parameters(), lr=1e-4, max_grad_norm=1.0)   # replaces AdamW + the clip_grad_norm_ line
"""
from __future__ import annotations

from typing import Optional

import torch

from . import ops as _ops


def _align4(n: int) -> int:
    return (n + 3) // 4 * 4


class FusedAdamW(torch.optim.Optimizer):
    def __init__(self, params, lr: float = 1e-3, betas=(0.9, 0.999), eps: float = 1e-8, weight_decay: float = 1e-2,
                 max_grad_norm: Optional[float] = None):
        if lr < 0 or eps < 0 or not 0 <= betas[0] < 1 or not 0 <= betas[1] < 1 or weight_decay < 0:
            raise ValueError("invalid AdamW hyper-parameter")
        super().__init__(params, dict(lr=lr, betas=tuple(betas), eps=eps, weight_decay=weight_decay))
        if len(self.param_groups) != 1:
            raise NotImplementedError("FusedAdamW updates the flat arena with one set of hyper-parameters")
        self.max_grad_norm = max_grad_norm
        self.lr_tensor: Optional[torch.Tensor] = None    # set to a device scalar to drive a schedule under CUDA graphs
        self._p = self._m = self._v = self._scal = None

    # ---- the flat views ---------------------------------------------------------------------------------------
    @staticmethod
    def _flat_of(tensors, what: str) -> torch.Tensor:
        st = tensors[0].untyped_storage()
        base = st.data_ptr()
        for t in tensors:
            if t.untyped_storage().data_ptr() != base:
                raise NotImplementedError(
                    f"FusedAdamW needs every {what} to live in the UNet's flat arena (call model(...) once after "
                    f"model.to(device); adapters / foreign parameters are not supported -- use torch.optim.AdamW)")
        n = st.nbytes() // 4
        return torch.empty(0, dtype=torch.float32, device=tensors[0].device).set_(st, 0, (n,))

    def load_state_dict(self, state_dict):
        """torch.optim.Optimizer.load_state_dict; the flat moments / step counter are adopted at the next step()."""
        super().load_state_dict(state_dict)
        for st in self.state.values():          # torch keeps same-device tensors by reference: own the restored copies
            for k in ("flat_exp_avg", "flat_exp_avg_sq", "flat_scalars"):
                if torch.is_tensor(st.get(k)):
                    st[k] = st[k].clone()
        self._p = self._m = self._v = self._scal = None

    def _bind(self):
        ps = self.param_groups[0]["params"]
        if any(not p.requires_grad for p in ps):
            raise NotImplementedError("FusedAdamW: frozen parameters in the group (LoRA runs use torch.optim.AdamW)")
        flat = self._flat_of(ps, "parameter")
        if sum(_align4(p.numel()) for p in ps) != flat.numel():
            raise NotImplementedError("FusedAdamW: the parameter list does not cover the arena (frozen parameters?)")
        if self._p is None or self._p.data_ptr() != flat.data_ptr():
            keep = self._m is not None and self._m.numel() == flat.numel() and self._m.device == flat.device
            self._p = flat
            saved = self.state.get(ps[0], {})
            if not keep and all(k in saved and torch.is_tensor(saved[k]) for k in
                                ("flat_exp_avg", "flat_exp_avg_sq", "flat_scalars")) and \
                    saved["flat_exp_avg"].numel() == flat.numel():
                # restored by load_state_dict(): continue from the saved moments and step count
                self._m = saved["flat_exp_avg"].to(device=flat.device, dtype=torch.float32).clone()
                self._v = saved["flat_exp_avg_sq"].to(device=flat.device, dtype=torch.float32).clone()
                self._scal = saved["flat_scalars"].to(device=flat.device, dtype=torch.float32).clone()
            elif not keep:
                self._m = torch.zeros_like(flat)
                self._v = torch.zeros_like(flat)
                self._scal = torch.zeros(4, dtype=torch.float32, device=flat.device)
            st = self.state[ps[0]]
            st["flat_exp_avg"], st["flat_exp_avg_sq"], st["flat_scalars"] = self._m, self._v, self._scal

    @torch.no_grad()
    def step(self, closure=None):
        loss = None
        if closure is not None:
            with torch.enable_grad():
                loss = closure()
        grp = self.param_groups[0]
        self._bind()
        grads = [p.grad for p in grp["params"]]
        if any(g is None for g in grads):
            raise RuntimeError("FusedAdamW.step(): a parameter has no gradient (run backward through the UNet first)")
        g = self._flat_of(grads, "gradient")
        if g.numel() != self._p.numel():
            raise RuntimeError("FusedAdamW: gradient arena and parameter arena differ in size")
        ops = _ops.get()
        gsq = None
        if self.max_grad_norm is not None:
            gsq = torch.zeros(1, dtype=torch.float32, device=g.device)
            ops.sumsq(g, gsq)
        b1, b2 = grp["betas"]
        ops.adamw_flat(self._p, g, self._m, self._v, self._scal, gsq, self.max_grad_norm, self.lr_tensor, grp["lr"], b1,
                       b2, grp["eps"], grp["weight_decay"])
        from .unet import arena_written
        arena_written(self._p.data_ptr())       # the kernel wrote behind the parameters' version counters
        return loss
