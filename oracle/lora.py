"""Oracle restatement of peft LoRA injection as the reference uses it (TEST INFRASTRUCTURE).

Follows SURVEY.md Appendix C.  Reference call sites:
  /root/reference/generator_model/train_with_lora_all_classes.py:316-322 (LoraConfig), :330 (add_adapter),
  :29-38 (save / load of "lora_" keys); /root/reference/generator_model/PolypGeneratorModel.py:54-58;
  /root/reference/generator_model/get_lorarized_layers.py:12-19 (key grammar recovery).
peft itself is un-vendored and unpinned (not in requirements.txt).
"""
from __future__ import annotations

import math
from dataclasses import dataclass, field
from typing import Dict, Iterable, List, Sequence, Union

import torch
import torch.nn as nn
import torch.nn.functional as F

from . import rounding as rq


@dataclass
class LoraConfig:
    r: int = 8
    lora_alpha: int = 8
    target_modules: Union[Sequence[str], str] = field(default_factory=lambda: ["to_q", "to_k", "to_v", "to_out.0"])
    lora_dropout: float = 0.0
    init_lora_weights: Union[bool, str] = True


class LoraLinear(nn.Module):
    """peft tuners/lora/layer.py::Linear, unmerged forward."""

    def __init__(self, base_layer: nn.Linear, cfg: LoraConfig, adapter_name: str = "default"):
        super().__init__()
        self.base_layer = base_layer
        self.adapter_name = adapter_name
        self.r = cfg.r
        self.scaling = cfg.lora_alpha / cfg.r
        self.merged = False
        self.lora_dropout = nn.ModuleDict(
            {adapter_name: nn.Dropout(cfg.lora_dropout) if cfg.lora_dropout > 0.0 else nn.Identity()})
        self.lora_A = nn.ModuleDict({adapter_name: nn.Linear(base_layer.in_features, cfg.r, bias=False)})
        self.lora_B = nn.ModuleDict({adapter_name: nn.Linear(cfg.r, base_layer.out_features, bias=False)})
        a, b = self.lora_A[adapter_name].weight, self.lora_B[adapter_name].weight
        if cfg.init_lora_weights is True:
            nn.init.kaiming_uniform_(a, a=math.sqrt(5))
        elif str(cfg.init_lora_weights).lower() == "gaussian":
            nn.init.normal_(a, std=1.0 / cfg.r)
        else:
            raise ValueError(f"unsupported init_lora_weights={cfg.init_lora_weights!r}")
        nn.init.zeros_(b)

    def forward(self, x):
        n = self.adapter_name
        if rq.enabled():
            # rounding-matched mode (oracle/rounding.py): same graph, bf16 at the product's storage points -- the
            # operand copies of W, A and (alpha/r) B, and the rank-r activation U = dropout(x) A^T
            y = F.linear(x, rq.weight(self.base_layer.weight), self.base_layer.bias)
            if self.merged:
                return y
            u = rq.act(F.linear(self.lora_dropout[n](x), rq.weight(self.lora_A[n].weight)))
            return y + F.linear(u, rq.weight(self.lora_B[n].weight * self.scaling))
        y = self.base_layer(x)
        if self.merged:
            return y
        return y + self.lora_B[n](self.lora_A[n](self.lora_dropout[n](x))) * self.scaling

    def delta_weight(self):
        n = self.adapter_name
        return (self.lora_B[n].weight @ self.lora_A[n].weight) * self.scaling

    def merge(self):
        if not self.merged:
            self.base_layer.weight.data += self.delta_weight()
            self.merged = True

    def unmerge(self):
        if self.merged:
            self.base_layer.weight.data -= self.delta_weight()
            self.merged = False


def _matches(name: str, targets: Iterable[str]) -> bool:
    return any(name == t or name.endswith("." + t) for t in targets)


def add_adapter(model: nn.Module, cfg: LoraConfig, adapter_name: str = "default") -> List[str]:
    """diffusers loaders/peft.py::add_adapter = peft.inject_adapter_in_model + freeze non-LoRA parameters."""
    targets = [cfg.target_modules] if isinstance(cfg.target_modules, str) else list(cfg.target_modules)
    wrapped = []
    for name, module in list(model.named_modules()):
        if isinstance(module, nn.Linear) and _matches(name, targets) and ".base_layer" not in name:
            parent = model
            parts = name.split(".")
            for p in parts[:-1]:
                parent = getattr(parent, p) if not p.isdigit() else parent[int(p)]
            wrapper = LoraLinear(module, cfg, adapter_name)
            if parts[-1].isdigit():
                parent[int(parts[-1])] = wrapper
            else:
                setattr(parent, parts[-1], wrapper)
            wrapped.append(name)
    if not wrapped:
        raise ValueError(f"Target modules {targets} not found in the base model.")
    for n, p in model.named_parameters():
        p.requires_grad_("lora_" in n)
    return wrapped


def merge_adapter(model: nn.Module):
    for m in model.modules():
        if isinstance(m, LoraLinear):
            m.merge()


def lora_state_dict(model: nn.Module) -> Dict[str, torch.Tensor]:
    """train_with_lora_all_classes.py:29-34 filter."""
    return {k: v.cpu() for k, v in model.state_dict().items() if "lora_" in k}


def recover_lora_modules(state_dict: Dict[str, torch.Tensor]) -> List[str]:
    """get_lorarized_layers.py:12-24: module path = dotted prefix before the lora_A / lora_B token."""
    out = set()
    for key in state_dict.keys():
        if "lora_A" in key or "lora_B" in key:
            parts = key.split(".")
            for i, p in enumerate(parts):
                if p in ("lora_A", "lora_B"):
                    out.add(".".join(parts[:i]))
                    break
    return sorted(out)
