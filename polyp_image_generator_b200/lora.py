"""LoRA adapters for the B200 UNet2DModel with peft's semantics and key grammar, folded into the tcgen05 GEMMs.

Call sites replaced (paths relative to /root/reference/):
  generator_model/train_with_lora_all_classes.py:316-322  LoraConfig(r=8, lora_alpha=8, target_modules=[...],
                                                           lora_dropout=0.3, init_lora_weights="gaussian")
  generator_model/train_with_lora_all_classes.py:330      unet.add_adapter(lora_config)
  generator_model/PolypGeneratorModel.py:54-58            add_lora_config -> unet.add_adapter
  generator_model/train_with_lora_all_classes.py:29-38    save ("lora_" in key) / load_state_dict(strict=False)
  generator_model/get_lorarized_layers.py:12-19           module-path recovery from saved keys
Semantics: y = base(x) + (alpha/r) * B(A(dropout(x))); A ~ N(0, 1/r^2) ("gaussian") or
kaiming-uniform (True), B = 0; every parameter without "lora_" in its name is frozen.

How it runs here: the low-rank branch never becomes separate tiny GEMMs.  U = dropout(x) A^T (rank padded to one
64-wide k-block) is produced by one narrow GEMM, and the update (alpha/r) U B^T is *folded into the projection GEMM
itself* as 64 extra reduction columns: A-operand [x | U], B-operand [W | (alpha/r) B].  The backward reuses the same
kernels (dU = dY (sB), dA = dU^T x_d, dB = s dY^T U, dX += mask * dU A).
"""
from __future__ import annotations

import math
from dataclasses import dataclass, field
from types import SimpleNamespace
from typing import Dict, List, Optional, Sequence, Union

import torch
import torch.nn as nn

from . import ops as _ops
from .ops import taps_1x1

LORA_K = 64   # one bf16 k-block; sum of the ranks folded into one GEMM must fit


@dataclass
class LoraConfig:
    """Subset of peft.LoraConfig the reference uses."""
    r: int = 8
    lora_alpha: int = 8
    target_modules: Union[Sequence[str], str] = field(default_factory=lambda: ["to_q", "to_k", "to_v", "to_out.0"])
    lora_dropout: float = 0.0
    init_lora_weights: Union[bool, str] = True
    bias: str = "none"


class LoraLinear(nn.Module):
    """peft.tuners.lora.layer.Linear look-alike (parameter holder; state-dict keys `<path>.base_layer.{weight,bias}`,
    `<path>.lora_A.<adapter>.weight` [r, in], `<path>.lora_B.<adapter>.weight` [out, r])."""

    def __init__(self, base_layer: nn.Linear, cfg: LoraConfig, adapter_name: str = "default"):
        super().__init__()
        self.base_layer = base_layer
        self.in_features, self.out_features = base_layer.in_features, base_layer.out_features
        self.adapter_name = adapter_name
        self.r = int(cfg.r)
        self.lora_alpha = cfg.lora_alpha
        self.scaling = cfg.lora_alpha / cfg.r
        self.p = float(cfg.lora_dropout)
        self.merged = False
        dev, dt = base_layer.weight.device, base_layer.weight.dtype
        self.lora_dropout = nn.ModuleDict({adapter_name: nn.Dropout(self.p) if self.p > 0 else nn.Identity()})
        self.lora_A = nn.ModuleDict({adapter_name: nn.Linear(self.in_features, self.r, bias=False, device=dev, dtype=dt)})
        self.lora_B = nn.ModuleDict({adapter_name: nn.Linear(self.r, self.out_features, bias=False, device=dev, dtype=dt)})
        a = self.lora_A[adapter_name].weight
        if cfg.init_lora_weights is True:
            nn.init.kaiming_uniform_(a, a=math.sqrt(5))
        elif str(cfg.init_lora_weights).lower() == "gaussian":
            nn.init.normal_(a, std=1.0 / self.r)
        else:
            raise ValueError(f"Unknown initialization init_lora_weights={cfg.init_lora_weights!r}")
        nn.init.zeros_(self.lora_B[adapter_name].weight)

    @property
    def A(self):
        return self.lora_A[self.adapter_name].weight

    @property
    def B(self):
        return self.lora_B[self.adapter_name].weight

    def forward(self, *a, **k):  # pragma: no cover
        raise RuntimeError("LoraLinear inside the B200 UNet2DModel is a parameter holder; call the model itself")


def _matches(name: str, targets) -> bool:
    return any(name == t or name.endswith("." + t) for t in targets)


def add_adapter(model: nn.Module, cfg: LoraConfig, adapter_name: str = "default") -> List[str]:
    """diffusers `add_adapter` = peft.inject_adapter_in_model + freeze everything that is not a LoRA parameter."""
    if getattr(cfg, "bias", "none") != "none":
        raise NotImplementedError("LoraConfig.bias != 'none' is not on the reference path")
    if any(isinstance(m, LoraLinear) for m in model.modules()):
        raise ValueError(f"Adapter with name {adapter_name} already exists. Please use a different name.")
    targets = [cfg.target_modules] if isinstance(cfg.target_modules, str) else list(cfg.target_modules)
    # attention projections (folded into the projection GEMMs, GemmLora) and the per-block time-embedding projections
    # (fp32 side computation on the [batch, channels] vectors, unet.UNet2DModel._temb_lora_*): the targets of
    # config_diffusion.py:34-37 that exist in a UNet2DModel
    supported = ("to_q", "to_k", "to_v", "to_out.0", "time_emb_proj")
    wrapped = []
    for name, module in list(model.named_modules()):
        if not isinstance(module, nn.Linear) or not _matches(name, targets):
            continue
        if not _matches(name, supported):
            raise NotImplementedError(
                f"LoRA on '{name}' is not implemented by the B200 path (supported targets: {supported})")
        parent = model
        parts = name.split(".")
        for p in parts[:-1]:
            parent = parent[int(p)] if p.isdigit() else getattr(parent, p)
        wrapper = LoraLinear(module, cfg, adapter_name)
        if parts[-1].isdigit():
            parent[int(parts[-1])] = wrapper
        else:
            setattr(parent, parts[-1], wrapper)
        wrapped.append(name)
    if not wrapped:
        raise ValueError(f"Target modules {targets} not found in the base model. "
                         f"Please check the target modules and try again.")
    for n, p in model.named_parameters():
        p.requires_grad_("lora_" in n)
    model._plan = None      # module tree changed: rebuild the execution plan and re-flatten lazily
    model._arena = None
    return wrapped


def lora_state_dict(model: nn.Module) -> Dict[str, torch.Tensor]:
    """train_with_lora_all_classes.py:29-34: {k: v.cpu() for k, v in state_dict if "lora_" in k}."""
    return {k: v.detach().cpu() for k, v in model.state_dict().items() if "lora_" in k}


def save_lora_weights(unet: nn.Module, save_path: str) -> str:
    """train_with_lora_all_classes.py:29-34 / train_with_lora_per_class.py: `<save_path>/lora_weights.pth` holding only
    the adapter tensors (CPU), the file the reference uploads to mlflow and reloads before sampling."""
    import os
    os.makedirs(save_path, exist_ok=True)
    path = os.path.join(save_path, "lora_weights.pth")
    torch.save(lora_state_dict(unet), path)
    return path


def load_lora_weights(device, unet: nn.Module, path: str) -> None:
    """train_with_lora_all_classes.py:36-38: `load_state_dict(strict=False)` of `<path>/lora_weights.pth`."""
    import os
    weights = torch.load(os.path.join(path, "lora_weights.pth"), map_location=device)
    unet.load_state_dict(weights, strict=False)


def recover_lora_modules(state_dict: Dict[str, torch.Tensor]) -> List[str]:
    """get_lorarized_layers.py:12-24."""
    out = set()
    for key in state_dict.keys():
        if "lora_A" in key or "lora_B" in key:
            parts = key.split(".")
            for i, p in enumerate(parts):
                if p in ("lora_A", "lora_B"):
                    out.add(".".join(parts[:i]))
                    break
    return sorted(out)


def merge_adapter(model: nn.Module) -> None:
    """W <- W + (alpha/r) B A on the fp32 master weights (peft merge); the adapter branch is then skipped."""
    ops = _ops.get()
    for m in model.modules():
        if isinstance(m, LoraLinear) and not m.merged:
            with torch.no_grad():
                # delta[out, in] = sum_r B[out, r] * A^T[in, r]  (fp32 SIMT linear kernel)
                delta = ops.linear_f32(m.B.detach().contiguous(), m.A.detach().t().contiguous(), None, False)
                m.base_layer.weight.add_(delta * m.scaling)
            m.merged = True
    if hasattr(model, "invalidate_weight_cache"):
        model.invalidate_weight_cache()


def unmerge_adapter(model: nn.Module) -> None:
    ops = _ops.get()
    for m in model.modules():
        if isinstance(m, LoraLinear) and m.merged:
            with torch.no_grad():
                delta = ops.linear_f32(m.B.detach().contiguous(), m.A.detach().t().contiguous(), None, False)
                m.base_layer.weight.sub_(delta * m.scaling)
            m.merged = False
    if hasattr(model, "invalidate_weight_cache"):
        model.invalidate_weight_cache()


class GemmLora:
    """LoRA state of one (possibly fused, e.g. q|k|v) projection GEMM."""

    _seed_counter = 0

    def __init__(self, mods: Sequence[Optional[LoraLinear]], cin: int, cout_each: int):
        self.mods = list(mods)
        self.cin, self.ce = cin, cout_each
        self.cout = cout_each * len(self.mods)
        self.slots = []
        off = 0
        for i, m in enumerate(self.mods):
            if m is None:
                continue
            self.slots.append((i, m, off))
            off += m.r
        if off > LORA_K:
            raise NotImplementedError(f"sum of LoRA ranks folded into one GEMM ({off}) exceeds {LORA_K}")
        self.p = max((m.p for _, m, _ in self.slots), default=0.0)
        self.a_ext = self.at_ext = self.bt_ext = None
        self.grads: Dict[int, torch.Tensor] = {}

    @property
    def active(self):
        return any(not m.merged for _, m, _ in self.slots)

    def write_operands(self, gobj) -> None:
        """Fill the extra k-block of the projection's B operand and the narrow A^T / B^T operands."""
        dev, dt = gobj.wf.device, gobj.wf.dtype
        k = gobj.taps * gobj.cin
        if self.a_ext is None or self.a_ext.device != dev:
            self.a_ext = torch.zeros((LORA_K, self.cin), device=dev, dtype=dt)
            self.at_ext = torch.zeros((self.cin, LORA_K), device=dev, dtype=dt)
            self.bt_ext = torch.zeros((LORA_K, self.cout), device=dev, dtype=dt)
        with torch.no_grad():
            # the regions no active slot writes stay zero from one step to the next: clear only when the set of active
            # slots (or the buffers) changed -- merge_adapter / unmerge_adapter, a new device
            layout = (gobj.wf.data_ptr(), self.a_ext.data_ptr(), tuple(m.merged for _, m, _ in self.slots))
            if layout != getattr(self, "_layout", None):
                gobj.wf[:, k:].zero_()
                self.a_ext.zero_()
                self.bt_ext.zero_()
                self._layout = layout
            for i, m, off in self.slots:
                if m.merged:
                    continue
                B = m.B.detach()
                # one launch each: fp32 product, rounded once on the store into the (strided) bf16 operand view
                torch.mul(B, m.scaling, out=gobj.wf[i * self.ce:(i + 1) * self.ce, k + off:k + off + m.r])
                torch.mul(B.t(), m.scaling, out=self.bt_ext[off:off + m.r, i * self.ce:(i + 1) * self.ce])
                self.a_ext[off:off + m.r].copy_(m.A.detach())
            self.at_ext.copy_(self.a_ext.t())

    def forward_extra(self, ops, x2: torch.Tensor, training: bool, tick: Optional[torch.Tensor] = None):
        """U = dropout(x) A_ext^T as a [1, 1, M, 64] bf16 tensor (the extra k-block of the projection's A operand)."""
        M = x2.shape[2]
        p = self.p if training else 0.0
        seed = offset = 0
        xd = x2
        if p > 0.0:
            GemmLora._seed_counter += 1
            seed, offset = torch.initial_seed() & 0x7FFFFFFFFFFFFFFF, GemmLora._seed_counter
            xd = ops.dropout(x2.contiguous(), p, seed, offset, tick=tick)
        u = ops.conv_gemm(xd, None, taps_1x1(), self.a_ext, LORA_K, (1, 1, M))
        return SimpleNamespace(u=u, xd=xd, p=p, seed=seed, offset=offset, tick=tick)

    def backward(self, ops, s, x2, dy2, d_x, zeros=None):
        """Adds the adapter's contribution to d_x and stores dA / dB in self.grads (keyed by id(param)).
        zeros(shape, device): the caller's zero-filled fp32 scratch (unet._ZeroPool.take), else torch.zeros."""
        M = x2.shape[2]
        dev = dy2.device
        if zeros is None:
            def zeros(shape, device):
                return torch.zeros(shape, device=device, dtype=torch.float32)
        dU = ops.conv_gemm(dy2, None, taps_1x1(), self.bt_ext, LORA_K, (1, 1, M))
        dA_ext = zeros((LORA_K, self.cin), dev)
        ops.conv_wgrad(dU, s.xd, None, taps_1x1(), dA_ext, (1, 1, M), accumulate=True)
        dB_ext = zeros((self.cout, LORA_K), dev)
        ops.conv_wgrad(dy2, s.u, None, taps_1x1(), dB_ext, (1, 1, M), accumulate=True)
        for i, m, off in self.slots:
            if m.merged:
                continue
            self.grads[id(m.A)] = dA_ext[off:off + m.r]
            self.grads[id(m.B)] = dB_ext[i * self.ce:(i + 1) * self.ce, off:off + m.r] * m.scaling
        if s.p > 0.0:
            dxl = ops.conv_gemm(dU, None, taps_1x1(), self.at_ext, self.cin, (1, 1, M))
            return ops.dropout(dxl, s.p, s.seed, s.offset, add=d_x.contiguous(), tick=s.tick)
        return ops.conv_gemm(dU, None, taps_1x1(), self.at_ext, self.cin, (1, 1, M), res=d_x)
