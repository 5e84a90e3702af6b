"""Rounding-matched mode of the oracle (TEST INFRASTRUCTURE).

The product stores activations and tensor-core operands in bf16 (fp32 accumulation, fp32 statistics); the reference
arithmetic (diffusers on fp32 weights) does not round anywhere.  At random init the ~120
chained layers of the UNet amplify every bf16 rounding, so "product vs fp32 oracle" measures the precision choice,
not the kernels.  Inside `with bf16_storage_points():` the oracle keeps its fp32 graph and op order but rounds to bf16
(round-to-nearest-even, then back to fp32) exactly where the product stores bf16:

  * conv / linear operands: the input activation and the weight matrix (bias, time embedding stay fp32),
  * every stored activation: conv outputs (after bias + temb + residual), GroupNorm(+SiLU) outputs, q|k|v,
    the attention output, and -- through autograd of the cast -- the gradients flowing through the same points.

Tests use it to assert north_star's 2e-2 per-tensor gradient bound (LoRA adapters included) against an oracle whose
only remaining difference from the product is accumulation order.
"""
from __future__ import annotations

import contextlib

import torch

_ON = False


def enabled() -> bool:
    return _ON


def act(x: torch.Tensor) -> torch.Tensor:
    """Storage point of an activation (autograd rounds the gradient at the same point)."""
    return x.to(torch.bfloat16).to(torch.float32) if _ON else x


def weight(w: torch.Tensor) -> torch.Tensor:
    """Tensor-core operand copy of a weight (the fp32 master receives the straight-through gradient)."""
    if not _ON:
        return w
    return w + (w.detach().to(torch.bfloat16).to(torch.float32) - w.detach())


@contextlib.contextmanager
def bf16_storage_points():
    global _ON
    prev, _ON = _ON, True
    try:
        yield
    finally:
        _ON = prev
