"""Drop-in `UNet2DModel` (diffusers 0.33.1 semantics) executed by hand-written sm_100a kernels.

Call sites replaced (paths relative to /root/reference/):
  generator_model/PolypGeneratorModel.py:25-48   UNet2DModel(sample_size=..., in_channels=3, ...)      (constructor)
  generator_model/train_from_scratch.py:71       model.to(device)
  generator_model/train_from_scratch.py:100      model(noisy_images, timesteps, return_dict=False)[0]  (forward)
  generator_model/train_from_scratch.py:103      scaler.scale(loss).backward()                          (backward)
  generator_model/train_from_scratch.py:106,273  model.parameters() -> clip_grad_norm_ / AdamW          (fp32 params)
  DDPMPipeline.__call__ -> unet(image, t).sample                                                        (sampling)

Design: the module tree and state-dict keys are the diffusers ones and the
parameters are ordinary fp32 `nn.Parameter`s, but they are *views into one flat fp32 arena* (conv weights stored
[Cout][kh][kw][Cin], i.e. channels_last), so that (i) one pass prepares all bf16 tensor-core operands, (ii) weight
gradients are produced directly in a flat gradient arena that DDP all-reduces in a few large buckets.
The forward and backward passes are explicit programs over the C-ABI ops (polyp_image_generator_b200.ops):
activations NHWC bf16, fp32 accumulation/statistics, NCHW fp32 only at the 3-channel model boundary.  The whole
network is a single torch.autograd.Function, so call sites keep using loss.backward() / optimizers unchanged.
"""
from __future__ import annotations

import json
import math
import os
import weakref
from dataclasses import dataclass
from types import SimpleNamespace
from typing import Dict, List, Optional, Sequence, Tuple, Union

import torch
import torch.nn as nn

from . import ops as _ops
from .ops import taps_1x1, taps_3x3, taps_s2d


@dataclass
class UNet2DOutput:
    sample: torch.Tensor


# ---------------------------------------------------------------------------------------------------------------
# module tree: parameter holders with the diffusers names (their own forward() is never used)
# ---------------------------------------------------------------------------------------------------------------
class _Holder(nn.Module):
    def forward(self, *a, **k):  # pragma: no cover
        raise RuntimeError("sub-modules of the B200 UNet2DModel are parameter holders; call the model itself")


class TimestepEmbedding(_Holder):
    def __init__(self, in_channels, time_embed_dim):
        super().__init__()
        self.linear_1 = nn.Linear(in_channels, time_embed_dim)
        self.linear_2 = nn.Linear(time_embed_dim, time_embed_dim)


class ResnetBlock2D(_Holder):
    def __init__(self, in_channels, out_channels, temb_channels, groups, eps):
        super().__init__()
        self.in_channels, self.out_channels = in_channels, out_channels
        self.norm1 = nn.GroupNorm(groups, in_channels, eps=eps)
        self.conv1 = nn.Conv2d(in_channels, out_channels, 3, padding=1)
        self.time_emb_proj = nn.Linear(temb_channels, out_channels)
        self.norm2 = nn.GroupNorm(groups, out_channels, eps=eps)
        self.conv2 = nn.Conv2d(out_channels, out_channels, 3, padding=1)
        self.conv_shortcut = nn.Conv2d(in_channels, out_channels, 1) if in_channels != out_channels else None


class Attention(_Holder):
    def __init__(self, channels, heads, dim_head, groups, eps):
        super().__init__()
        self.heads, self.dim_head = heads, dim_head
        self.group_norm = nn.GroupNorm(groups, channels, eps=eps)
        self.to_q = nn.Linear(channels, channels)
        self.to_k = nn.Linear(channels, channels)
        self.to_v = nn.Linear(channels, channels)
        self.to_out = nn.ModuleList([nn.Linear(channels, channels), nn.Dropout(0.0)])


class Downsample2D(_Holder):
    def __init__(self, channels, padding):
        super().__init__()
        self.padding = padding
        self.conv = nn.Conv2d(channels, channels, 3, stride=2, padding=padding)


class Upsample2D(_Holder):
    def __init__(self, channels):
        super().__init__()
        self.conv = nn.Conv2d(channels, channels, 3, padding=1)


class _Block(_Holder):
    pass


def _head_cfg(channels, attention_head_dim):
    d = attention_head_dim if attention_head_dim is not None else channels
    return channels // d, d


def _make_down(kind, n_layers, cin, cout, temb, add_down, eps, groups, pad, head_dim):
    b = _Block()
    b.resnets = nn.ModuleList([ResnetBlock2D(cin if i == 0 else cout, cout, temb, groups, eps) for i in range(n_layers)])
    if kind == "AttnDownBlock2D":
        h, d = _head_cfg(cout, head_dim)
        b.attentions = nn.ModuleList([Attention(cout, h, d, groups, eps) for _ in range(n_layers)])
    elif kind != "DownBlock2D":
        raise ValueError(f"{kind} does not exist.")
    b.downsamplers = nn.ModuleList([Downsample2D(cout, pad)]) if add_down else None
    return b


def _make_up(kind, n_layers, cin, prev, cout, temb, add_up, eps, groups, head_dim):
    b = _Block()
    res = []
    for i in range(n_layers):
        skip = cin if i == n_layers - 1 else cout
        rin = prev if i == 0 else cout
        res.append(ResnetBlock2D(rin + skip, cout, temb, groups, eps))
    b.resnets = nn.ModuleList(res)
    if kind == "AttnUpBlock2D":
        h, d = _head_cfg(cout, head_dim)
        b.attentions = nn.ModuleList([Attention(cout, h, d, groups, eps) for _ in range(n_layers)])
    elif kind != "UpBlock2D":
        raise ValueError(f"{kind} does not exist.")
    b.upsamplers = nn.ModuleList([Upsample2D(cout)]) if add_up else None
    return b


# ---------------------------------------------------------------------------------------------------------------
# layer records used by the forward / backward programs
# ---------------------------------------------------------------------------------------------------------------
class _Gemm:
    """One tensor-core layer: fp32 master weight (arena view, [cout][taps][cin] physical) + bf16 operand copies."""

    def __init__(self, name, weights: Sequence[nn.Parameter], biases: Sequence[Optional[nn.Parameter]], taps, cin,
                 cout_each):
        self.name = name
        self.weights, self.biases = list(weights), list(biases)
        self.taps, self.cin, self.cout_each = taps, cin, cout_each
        self.cout = cout_each * len(self.weights)
        self.extra_k = 0          # LoRA: extra (padded) reduction columns appended to the fprop operand
        self.w_off = self.b_off = -1
        self.wf = self.wd = None  # bf16 [cout, taps*cin (+extra_k)], [cin, taps*cout]
        self.lora = None

    @property
    def trainable(self):
        return any(w.requires_grad for w in self.weights)

    @property
    def bias_trainable(self):
        return any(b is not None and b.requires_grad for b in self.biases)


class _Norm:
    def __init__(self, name, mod: nn.GroupNorm):
        self.name, self.mod = name, mod
        self.groups, self.eps = mod.num_groups, mod.eps
        self.g_off = self.b_off = -1

    @property
    def trainable(self):
        return self.mod.weight.requires_grad or self.mod.bias.requires_grad


class _Tape:
    """Backward closures recorded by the forward program (executed in reverse)."""

    def __init__(self):
        self.steps = []          # list of (fn, needs) ; fn(g, state) -> g
        self.keep = []

    def add(self, fn):
        self.steps.append(fn)


# arena data_ptr -> owning model: writers that bypass autograd's version counters (optim.FusedAdamW, graph replays)
# call arena_written() so the eval-mode bf16 operand cache is refreshed
_ARENA_OWNERS: Dict[int, "weakref.ref"] = {}


def arena_written(data_ptr: int) -> None:
    ref = _ARENA_OWNERS.get(data_ptr)
    m = ref() if ref is not None else None
    if m is not None:
        m.invalidate_weight_cache()


_DEPRECATED_ATTN = ((".query.", ".to_q."), (".key.", ".to_k."), (".value.", ".to_v."), (".proj_attn.", ".to_out.0."))


def convert_deprecated_attention_keys(state_dict):
    """diffusers' ModelMixin._convert_deprecated_attention_blocks, applied at load time: checkpoints saved before the
    AttentionBlock -> Attention refactor (google/ddpm-celebahq-256 and the other google/ddpm-* repositories) store the
    attention projections as `*.attentions.N.{query,key,value,proj_attn}.{weight,bias}`; the module tree names them
    `to_q / to_k / to_v / to_out.0`.  Keys that already use the new names pass through unchanged."""
    if not any(".attentions." in k and any(old in k for old, _ in _DEPRECATED_ATTN) for k in state_dict):
        return state_dict
    out = type(state_dict)()
    for k, v in state_dict.items():
        if ".attentions." in k:
            for old, new in _DEPRECATED_ATTN:
                if old in k:
                    k = k.replace(old, new)
                    break
        out[k] = v
    return out


class _ZeroPool:
    """Zero-filled fp32 scratch for the reduction targets of ONE pass (per-conv GroupNorm moments, per-norm backward
    sums, d_temb): one fill per pass instead of one fill kernel per target (~70 launches of ~2.4 us in a training step,
    ~130 in a LoRA step).  The pass's total is learned on the first pass; a pass that asks for more gets an extra chunk."""

    ALIGN = 32      # floats: every slice starts on a 128-byte boundary

    def __init__(self):
        self.need = 0
        self.buf = None
        self.off = 0
        self.taken = 0

    def begin(self, device):
        self.need = max(self.need, self.taken)
        self.taken = 0
        self.off = 0
        self.buf = torch.zeros(self.need, device=device, dtype=torch.float32) if self.need else None
        return self

    def take(self, shape, device) -> torch.Tensor:
        n = 1
        for d in shape:
            n *= int(d)
        na = (n + self.ALIGN - 1) // self.ALIGN * self.ALIGN
        self.taken += na
        if self.buf is None or self.buf.device != device or self.off + na > self.buf.numel():
            return torch.zeros(tuple(shape), device=device, dtype=torch.float32)     # first pass / grown pass
        out = self.buf[self.off:self.off + n].view(tuple(shape))
        self.off += na
        return out


def _align(n, a=4):
    return (n + a - 1) // a * a


class UNet2DModel(nn.Module):
    """diffusers.UNet2DModel signature; unsupported options raise instead of diverging."""

    def __init__(self, sample_size: Optional[Union[int, Tuple[int, int]]] = None, in_channels: int = 3,
                 out_channels: int = 3, center_input_sample: bool = False, time_embedding_type: str = "positional",
                 time_embedding_dim: Optional[int] = None, freq_shift: int = 0, flip_sin_to_cos: bool = True,
                 down_block_types: Tuple[str, ...] = ("DownBlock2D", "AttnDownBlock2D", "AttnDownBlock2D",
                                                      "AttnDownBlock2D"),
                 mid_block_type: Optional[str] = "UNetMidBlock2D",
                 up_block_types: Tuple[str, ...] = ("AttnUpBlock2D", "AttnUpBlock2D", "AttnUpBlock2D", "UpBlock2D"),
                 block_out_channels: Tuple[int, ...] = (224, 448, 672, 896), layers_per_block: int = 2,
                 mid_block_scale_factor: float = 1, downsample_padding: int = 1, downsample_type: str = "conv",
                 upsample_type: str = "conv", dropout: float = 0.0, act_fn: str = "silu",
                 attention_head_dim: Optional[int] = 8, norm_num_groups: int = 32,
                 attn_norm_num_groups: Optional[int] = None, norm_eps: float = 1e-5,
                 resnet_time_scale_shift: str = "default", add_attention: bool = True,
                 class_embed_type: Optional[str] = None, num_class_embeds: Optional[int] = None,
                 num_train_timesteps: Optional[int] = None):
        super().__init__()
        unsupported = []
        if center_input_sample: unsupported.append("center_input_sample")
        if time_embedding_type != "positional": unsupported.append(f"time_embedding_type={time_embedding_type}")
        if mid_block_type != "UNetMidBlock2D": unsupported.append(f"mid_block_type={mid_block_type}")
        if downsample_type != "conv" or upsample_type != "conv": unsupported.append("resnet down/upsampling")
        if dropout != 0.0: unsupported.append("dropout")
        if act_fn not in ("silu", "swish"): unsupported.append(f"act_fn={act_fn}")
        if resnet_time_scale_shift != "default": unsupported.append("resnet_time_scale_shift")
        if class_embed_type is not None or num_class_embeds is not None: unsupported.append("class embedding")
        if mid_block_scale_factor != 1: unsupported.append("mid_block_scale_factor")
        if attn_norm_num_groups is not None: unsupported.append("attn_norm_num_groups")
        if downsample_padding not in (0, 1): unsupported.append(f"downsample_padding={downsample_padding}")
        if unsupported:
            raise NotImplementedError("UNet2DModel (B200 hot path) does not implement: " + ", ".join(unsupported))
        if len(down_block_types) != len(up_block_types):
            raise ValueError(f"Must provide the same number of `down_block_types` as `up_block_types`. "
                             f"`down_block_types`: {down_block_types}. `up_block_types`: {up_block_types}.")
        if len(block_out_channels) != len(down_block_types):
            raise ValueError(f"Must provide the same number of `block_out_channels` as `down_block_types`. "
                             f"`block_out_channels`: {block_out_channels}. `down_block_types`: {down_block_types}.")
        if in_channels > 4 or out_channels > 4:
            raise NotImplementedError("conv_in / conv_out kernels support at most 4 image channels")
        for c in block_out_channels:
            if c % 64 or c % norm_num_groups:
                raise NotImplementedError("block_out_channels must be multiples of 64 (tcgen05 k-block) and of groups")

        self.config = SimpleNamespace(
            sample_size=sample_size, in_channels=in_channels, out_channels=out_channels,
            center_input_sample=center_input_sample, time_embedding_type=time_embedding_type,
            time_embedding_dim=time_embedding_dim, freq_shift=freq_shift, flip_sin_to_cos=flip_sin_to_cos,
            down_block_types=tuple(down_block_types), mid_block_type=mid_block_type,
            up_block_types=tuple(up_block_types), block_out_channels=tuple(block_out_channels),
            layers_per_block=layers_per_block, mid_block_scale_factor=mid_block_scale_factor,
            downsample_padding=downsample_padding, downsample_type=downsample_type, upsample_type=upsample_type,
            dropout=dropout, act_fn=act_fn, attention_head_dim=attention_head_dim, norm_num_groups=norm_num_groups,
            attn_norm_num_groups=attn_norm_num_groups, norm_eps=norm_eps,
            resnet_time_scale_shift=resnet_time_scale_shift, add_attention=add_attention,
            class_embed_type=class_embed_type, num_class_embeds=num_class_embeds,
            num_train_timesteps=num_train_timesteps)
        self.sample_size = sample_size
        boc = tuple(block_out_channels)
        ted = time_embedding_dim or boc[0] * 4
        self._temb_dim = ted
        g, eps = norm_num_groups, norm_eps

        # construction order follows diffusers so torch.manual_seed(s) gives the same default init stream
        self.conv_in = nn.Conv2d(in_channels, boc[0], 3, padding=1)
        self.time_embedding = TimestepEmbedding(boc[0], ted)
        self.down_blocks = nn.ModuleList()
        out_ch = boc[0]
        for i, kind in enumerate(down_block_types):
            in_ch, out_ch = out_ch, boc[i]
            self.down_blocks.append(_make_down(kind, layers_per_block, in_ch, out_ch, ted, i != len(boc) - 1, eps, g,
                                               downsample_padding, attention_head_dim))
        mid = _Block()
        mid.resnets = nn.ModuleList([ResnetBlock2D(boc[-1], boc[-1], ted, g, eps)])
        if add_attention:
            h, d = _head_cfg(boc[-1], attention_head_dim)
            mid.attentions = nn.ModuleList([Attention(boc[-1], h, d, g, eps)])
        else:
            mid.attentions = nn.ModuleList([])
        mid.resnets.append(ResnetBlock2D(boc[-1], boc[-1], ted, g, eps))
        self.mid_block = mid
        self.up_blocks = nn.ModuleList()
        rev = list(reversed(boc))
        out_ch = rev[0]
        for i, kind in enumerate(up_block_types):
            prev, out_ch = out_ch, rev[i]
            in_ch = rev[min(i + 1, len(boc) - 1)]
            self.up_blocks.append(_make_up(kind, layers_per_block + 1, in_ch, prev, out_ch, ted, i != len(boc) - 1, eps,
                                           g, attention_head_dim))
        self.conv_norm_out = nn.GroupNorm(g, boc[0], eps=eps)
        self.conv_out = nn.Conv2d(boc[0], out_channels, 3, padding=1)

        self._arena: Optional[torch.Tensor] = None
        self._wgrad_stream = None
        self._prep_table = None
        self._plan = None
        self._wcache_key = None
        self._lora_layers: Dict[str, object] = {}
        self.last_launches = 0
        # Inference arithmetic (no-grad forward, i.e. sampling): "bf16" = bf16 tensor-core operands and bf16 activations
        # (fast path); "fp32" = fp32-faithful split-bf16 path (_run_forward_split): what the reference does when it
        # samples outside autocast (train_from_scratch.py:121-125, 39-66).  Training always runs the bf16 path.
        self.inference_precision = "bf16"
        self._w3_cache = None

    # ---- nn.Module conveniences expected by call sites ------------------------------------------------------
    @property
    def dtype(self):
        return torch.float32

    @property
    def device(self):
        return self.conv_in.weight.device

    # ---- LoRA (peft-style) ------------------------------------------------------------------------------------
    def add_adapter(self, adapter_config, adapter_name: str = "default"):
        from .lora import add_adapter
        return add_adapter(self, adapter_config, adapter_name)

    # ---- persistence --------------------------------------------------------------------------------------------
    def save_pretrained(self, save_directory: str, safe_serialization: bool = True):
        os.makedirs(save_directory, exist_ok=True)
        cfg = {k: (list(v) if isinstance(v, tuple) else v) for k, v in vars(self.config).items()}
        cfg.update({"_class_name": "UNet2DModel", "_diffusers_version": "0.33.1"})
        with open(os.path.join(save_directory, "config.json"), "w") as f:
            json.dump(cfg, f, indent=2)
        sd = {k: v.detach().cpu().contiguous() for k, v in self.state_dict().items()}
        if safe_serialization:
            from safetensors.torch import save_file
            save_file(sd, os.path.join(save_directory, "diffusion_pytorch_model.safetensors"))
        else:
            torch.save(sd, os.path.join(save_directory, "diffusion_pytorch_model.bin"))

    @classmethod
    def from_pretrained(cls, directory: str):
        with open(os.path.join(directory, "config.json")) as f:
            cfg = {k: v for k, v in json.load(f).items() if not k.startswith("_")}
        for k in ("down_block_types", "up_block_types", "block_out_channels"):
            cfg[k] = tuple(cfg[k])
        model = cls(**cfg)
        st = os.path.join(directory, "diffusion_pytorch_model.safetensors")
        if os.path.exists(st):
            from safetensors.torch import load_file
            sd = load_file(st)
        else:
            sd = torch.load(os.path.join(directory, "diffusion_pytorch_model.bin"), map_location="cpu")
        model.load_state_dict(convert_deprecated_attention_keys(sd))
        return model

    def load_state_dict(self, state_dict, strict: bool = True, assign: bool = False):
        """nn.Module.load_state_dict, accepting the pre-0.14 attention-block key names of the hub's google/ddpm-*
        checkpoints (see convert_deprecated_attention_keys)."""
        return super().load_state_dict(convert_deprecated_attention_keys(state_dict), strict=strict, assign=assign)

    # ---------------------------------------------------------------------------------------------------------
    # plan: layer records in execution order + arena layout
    # ---------------------------------------------------------------------------------------------------------
    def _base(self, lin):
        """nn.Linear behind an optional LoRA wrapper."""
        return getattr(lin, "base_layer", lin)

    def _build_plan(self):
        P = SimpleNamespace()
        P.gemms: List[_Gemm] = []
        P.norms: List[_Norm] = []
        P.resnets = []

        def gemm(name, mods, taps, cin, cout_each):
            mods = [self._base(m) for m in mods]
            gobj = _Gemm(name, [m.weight for m in mods], [m.bias for m in mods], taps, cin, cout_each)
            P.gemms.append(gobj)
            return gobj

        def norm(name, mod):
            nobj = _Norm(name, mod)
            P.norms.append(nobj)
            return nobj

        temb_off = [0]

        def resnet(name, r: ResnetBlock2D):
            rec = SimpleNamespace(name=name, cin=r.in_channels, cout=r.out_channels, mod=r)
            rec.norm1 = norm(name + ".norm1", r.norm1)
            rec.conv1 = gemm(name + ".conv1", [r.conv1], 9, r.in_channels, r.out_channels)
            rec.norm2 = norm(name + ".norm2", r.norm2)
            rec.conv2 = gemm(name + ".conv2", [r.conv2], 9, r.out_channels, r.out_channels)
            rec.short = gemm(name + ".conv_shortcut", [r.conv_shortcut], 1, r.in_channels, r.out_channels) \
                if r.conv_shortcut is not None else None
            rec.temb_off = temb_off[0]
            temb_off[0] += r.out_channels
            rec.temb_lora = r.time_emb_proj if hasattr(r.time_emb_proj, "base_layer") else None   # lora.LoraLinear
            P.resnets.append(rec)
            return rec

        def attention(name, a: Attention):
            c = a.group_norm.num_channels
            rec = SimpleNamespace(name=name, c=c, heads=a.heads, d=a.dim_head, mod=a)
            rec.norm = norm(name + ".group_norm", a.group_norm)
            rec.qkv = gemm(name + ".to_qkv", [a.to_q, a.to_k, a.to_v], 1, c, c)
            rec.out = gemm(name + ".to_out.0", [a.to_out[0]], 1, c, c)
            from .lora import LORA_K, GemmLora, LoraLinear
            lm = [m if isinstance(m, LoraLinear) else None for m in (a.to_q, a.to_k, a.to_v)]
            if any(m is not None for m in lm):
                rec.qkv.lora, rec.qkv.extra_k = GemmLora(lm, c, c), LORA_K
            if isinstance(a.to_out[0], LoraLinear):
                rec.out.lora, rec.out.extra_k = GemmLora([a.to_out[0]], c, c), LORA_K
            return rec

        P.down = []
        for i, blk in enumerate(self.down_blocks):
            b = SimpleNamespace(resnets=[], attns=[], down=None)
            for j, r in enumerate(blk.resnets):
                b.resnets.append(resnet(f"down_blocks.{i}.resnets.{j}", r))
                if hasattr(blk, "attentions"):
                    b.attns.append(attention(f"down_blocks.{i}.attentions.{j}", blk.attentions[j]))
            if blk.downsamplers is not None:
                d = blk.downsamplers[0]
                ch = d.conv.in_channels
                b.down = SimpleNamespace(conv=gemm(f"down_blocks.{i}.downsamplers.0.conv", [d.conv], 9, ch, ch),
                                         pad=d.padding, c=ch)
            P.down.append(b)
        m = self.mid_block
        P.mid = SimpleNamespace(r0=resnet("mid_block.resnets.0", m.resnets[0]),
                                attn=attention("mid_block.attentions.0", m.attentions[0]) if len(m.attentions) else None,
                                r1=resnet("mid_block.resnets.1", m.resnets[1]))
        P.up = []
        for i, blk in enumerate(self.up_blocks):
            b = SimpleNamespace(resnets=[], attns=[], up=None)
            for j, r in enumerate(blk.resnets):
                b.resnets.append(resnet(f"up_blocks.{i}.resnets.{j}", r))
                if hasattr(blk, "attentions"):
                    b.attns.append(attention(f"up_blocks.{i}.attentions.{j}", blk.attentions[j]))
            if blk.upsamplers is not None:
                u = blk.upsamplers[0]
                ch = u.conv.in_channels
                b.up = SimpleNamespace(conv=gemm(f"up_blocks.{i}.upsamplers.0.conv", [u.conv], 9, ch, ch), c=ch)
            P.up.append(b)
        P.norm_out = norm("conv_norm_out", self.conv_norm_out)
        P.temb_total = temb_off[0]

        # ---- arena layout (floats).  Order: [tensor-core weights][time_emb_proj weights][everything else] ----
        layout: List[Tuple[nn.Parameter, int, Optional[Tuple[int, ...]]]] = []  # (param, offset, physical shape)
        off = 0

        def place(p: nn.Parameter, phys=None):
            nonlocal off
            layout.append((p, off, phys))
            o = off
            off = _align(off + p.numel())
            return o

        for gobj in P.gemms:
            gobj.w_off = off
            for w in gobj.weights:
                phys = (w.shape[0], w.shape[2], w.shape[3], w.shape[1]) if w.dim() == 4 else None
                place(w, phys)
        P.temb_w_off = off
        for r in P.resnets:
            place(self._base(r.mod.time_emb_proj).weight)
        for gobj in P.gemms:
            gobj.b_off = off
            for b in gobj.biases:
                place(b)
        P.temb_b_off = off
        for r in P.resnets:
            place(self._base(r.mod.time_emb_proj).bias)
        for nobj in P.norms:
            nobj.g_off = place(nobj.mod.weight)
            nobj.b_off = place(nobj.mod.bias)
        te = self.time_embedding
        P.te = SimpleNamespace(w1=place(te.linear_1.weight), b1=place(te.linear_1.bias), w2=place(te.linear_2.weight),
                               b2=place(te.linear_2.bias))
        w = self.conv_in.weight
        P.cin_w = place(w, (w.shape[0], 3, 3, w.shape[1]))
        P.cin_b = place(self.conv_in.bias)
        w = self.conv_out.weight
        P.cout_w = place(w, (w.shape[0], 3, 3, w.shape[1]))
        P.cout_b = place(self.conv_out.bias)
        P.layout, P.total = layout, off
        placed = {id(p) for p, _, _ in layout}
        P.extra_params = [(n, p) for n, p in self.named_parameters() if id(p) not in placed]  # LoRA A/B
        for n, _ in P.extra_params:
            if "lora_" not in n:
                raise RuntimeError(f"parameter {n} is not covered by the execution plan")
        # bf16 operand arena
        boff = 0
        for gobj in P.gemms:
            k = gobj.taps * gobj.cin
            gobj.wf_off, gobj.wf_shape = boff, (gobj.cout, k + gobj.extra_k)
            boff = _align(boff + gobj.cout * (k + gobj.extra_k), 64)
            gobj.wd_off, gobj.wd_shape = boff, (gobj.cin, gobj.taps * gobj.cout)
            boff = _align(boff + gobj.cin * gobj.taps * gobj.cout, 64)
        # boundary convs as one-k-block GEMMs: conv_in operand [c0, 64], conv_out fprop operand [32, 9*c0] (rows >=
        # out_channels are zero) and conv_out dgrad operand [c0, 64]
        c0 = self.conv_in.out_channels
        P.cin_wf_off, boff = boff, _align(boff + c0 * 64, 64)
        P.cout_wf_off, boff = boff, _align(boff + 32 * 9 * c0, 64)
        P.cout_wd_off, boff = boff, _align(boff + c0 * 64, 64)
        P.bf16_total = boff
        return P

    def _ensure_arena(self):
        """(Re)build the flat fp32 arena on the parameters' device and point every parameter into it."""
        if self._plan is None:
            self._plan = self._build_plan()
        P = self._plan
        dev = self.conv_in.weight.device
        if dev.type != "cuda" and _ops.get().name == "cuda":
            raise RuntimeError("UNet2DModel (B200) runs on CUDA only: call model.to('cuda') first; "
                               "there is no CPU fallback for the DDPM hot path")
        ar = self._arena
        ok = ar is not None and ar.device == dev
        if ok:
            base = ar.data_ptr()
            for p, o, _ in P.layout:
                if p.data_ptr() != base + 4 * o:
                    ok = False
                    break
        if ok:
            return
        new = torch.zeros(P.total, device=dev, dtype=torch.float32)
        with torch.no_grad():
            for p, o, phys in P.layout:
                seg = new[o:o + p.numel()]
                if phys is not None:
                    view = seg.view(phys).permute(0, 3, 1, 2)     # logical [co, ci, kh, kw], channels_last storage
                else:
                    view = seg.view(p.shape)
                view.copy_(p.detach().to(device=dev, dtype=torch.float32))
                p.data = view
        self._arena = new
        _ARENA_OWNERS[new.data_ptr()] = weakref.ref(self)
        self._bf16 = torch.zeros(P.bf16_total, device=dev, dtype=getattr(_ops.get(), 'operand_dtype', torch.bfloat16))
        for gobj in P.gemms:
            gobj.wf = self._bf16[gobj.wf_off:gobj.wf_off + gobj.wf_shape[0] * gobj.wf_shape[1]].view(gobj.wf_shape)
            gobj.wd = self._bf16[gobj.wd_off:gobj.wd_off + gobj.wd_shape[0] * gobj.wd_shape[1]].view(gobj.wd_shape)
        c0 = self.conv_in.out_channels
        self._cin_wf = self._bf16[P.cin_wf_off:P.cin_wf_off + c0 * 64].view(c0, 64)
        self._cout_wf = self._bf16[P.cout_wf_off:P.cout_wf_off + 32 * 9 * c0].view(32, 9 * c0)
        self._cout_wf._ddpm_alg_cout = self.conv_out.out_channels      # rows that carry data (FLOP accounting only)
        self._cout_wd = self._bf16[P.cout_wd_off:P.cout_wd_off + c0 * 64].view(c0, 64)
        self._cout_b32 = torch.zeros(32, device=dev, dtype=torch.float32)
        self._prep_table = None
        self._wcache_key = None

    def _apply(self, fn, *a, **k):
        out = super()._apply(fn, *a, **k)
        self._arena = None   # parameters were re-created; re-flatten lazily
        return out

    def invalidate_weight_cache(self):
        """The fp32 arena was written behind autograd's version counters (fused optimizer, CUDA-graph replay)."""
        self._wcache_key = None
        self._wcache_base = None
        self._weights_epoch = getattr(self, "_weights_epoch", 0) + 1

    def _seg(self, buf, off, n):
        return buf[off:off + n]

    def _prepare_weights(self, training: bool):
        """fp32 master -> bf16 tensor-core operands (fprop layout + transposed/flipped dgrad layout)."""
        P = self._plan
        key = None
        # Frozen tensor-core weights (LoRA fine-tuning, train_with_lora_all_classes.py:330: everything but the adapters has
        # requires_grad = False) keep their bf16 operand copies across TRAINING steps too: only the adapters' extra
        # k-block is rewritten.  113 M weights re-cast per step was 0.5 ms of a 17 ms LoRA step.
        frozen = not (self.conv_in.weight.requires_grad or self.conv_out.weight.requires_grad or
                      self.conv_out.bias.requires_grad or any(g.trainable for g in P.gemms))
        if not training or frozen:
            base = (training, getattr(self, "_weights_epoch", 0)) + \
                tuple(w._version for gobj in P.gemms for w in gobj.weights) + \
                (self.conv_in.weight._version, self.conv_out.weight._version, self.conv_out.bias._version)
            key = base + tuple(p._version for _, p in P.extra_params)
            if key == self._wcache_key and not training:
                return
            if base == getattr(self, "_wcache_base", None) and self._wcache_key is not None:
                # only the adapters can have moved; a training step rewrites them unconditionally (a CUDA-graph capture
                # of the step must contain these copies whatever the version counters say at capture time)
                for gobj in P.gemms:
                    if gobj.lora is not None:
                        gobj.lora.write_operands(gobj)
                self._wcache_key = key
                return
            self._wcache_base = base
        ops = _ops.get()
        ar = self._arena
        need_d = training
        if hasattr(ops, "prep_weights_batched"):
            # every layer in ONE launch, driven by a device-resident descriptor table (built once per arena)
            if self._prep_table is None:
                entries = []
                for gobj in P.gemms:
                    k = gobj.taps * gobj.cin
                    o = gobj.w_off
                    for i, w in enumerate(gobj.weights):
                        n = w.numel()
                        rows = slice(i * gobj.cout_each, (i + 1) * gobj.cout_each)
                        cols = slice(i * gobj.taps * gobj.cout_each, (i + 1) * gobj.taps * gobj.cout_each)
                        entries.append((ar[o:o + n], gobj.wf[rows, :k], gobj.wd[:, cols], gobj.cout_each, gobj.taps,
                                        gobj.cin))
                        o = _align(o + n)
                self._prep_table = ops.build_prep_table(entries, ar.device)
            ops.prep_weights_batched(*self._prep_table, need_d)
        else:
            for gobj in P.gemms:
                k = gobj.taps * gobj.cin
                o = gobj.w_off
                for i, w in enumerate(gobj.weights):
                    n = w.numel()
                    wseg = ar[o:o + n]
                    rows = slice(i * gobj.cout_each, (i + 1) * gobj.cout_each)
                    wf = gobj.wf[rows, :k]
                    wd = gobj.wd[:, i * gobj.taps * gobj.cout_each:(i + 1) * gobj.taps * gobj.cout_each] \
                        if need_d else None
                    # wd for fused multi-weight gemms (qkv): [cin, cout_total], this weight's columns at an offset
                    ops.prep_weight(wseg, wf, wd, gobj.cout_each, gobj.taps, gobj.cin)
                    o = _align(o + n)
        for gobj in P.gemms:
            if gobj.lora is not None:
                gobj.lora.write_operands(gobj)
        # boundary convs (a few thousand elements: plain tensor copies)
        cfg = self.config
        c0, ci, co = cfg.block_out_channels[0], cfg.in_channels, cfg.out_channels
        with torch.no_grad():
            w_in = self._aview(P.cin_w, (c0, 9 * ci))                   # [c0][tap][ci]
            self._cin_wf[:, :9 * ci].copy_(w_in)
            w_out = self._aview(P.cout_w, (co, 9, c0))                  # [co][tap][ci]
            self._cout_wf[:co].copy_(w_out.reshape(co, 9 * c0))
            self._cout_b32[:co].copy_(self._aview(P.cout_b, (co,)))
            if need_d:   # dgrad operand: Wd[ci][tap'*co_n + co] = w[co][8 - tap'][ci]
                self._cout_wd[:, :9 * co].copy_(w_out.flip(1).permute(2, 1, 0).reshape(c0, 9 * co))
        self._wcache_key = key

    # ---------------------------------------------------------------------------------------------------------
    # forward
    # ---------------------------------------------------------------------------------------------------------
    def forward(self, sample: torch.Tensor, timestep: Union[torch.Tensor, float, int], class_labels=None,
                return_dict: bool = True):
        if class_labels is not None:
            raise ValueError("class_labels should be provided only when the model has a class embedding "
                             "(this UNet2DModel was built without one)")
        if sample.dim() != 4 or sample.shape[1] != self.config.in_channels:
            raise ValueError(f"expected sample of shape [B, {self.config.in_channels}, H, W], got {tuple(sample.shape)}")
        self._ensure_arena()
        dev = self.device
        n = sample.shape[0]
        ts = timestep
        if not torch.is_tensor(ts):
            ts = torch.full((n,), int(ts), dtype=torch.int64, device=dev)
        else:
            ts = ts.to(device=dev, dtype=torch.int64)
            ts = ts.reshape(1).expand(n) if ts.dim() == 0 else ts
            if ts.numel() != n:
                ts = ts * torch.ones(n, dtype=ts.dtype, device=dev)
        ts = ts.contiguous()
        x = sample.detach().to(device=dev, dtype=torch.float32).contiguous()

        params = [p for p in self.parameters() if p.requires_grad]
        if torch.is_grad_enabled() and params:
            out = _UNetFunction.apply(self, x, ts, len(params), *params)
        elif self.inference_precision == "fp32":
            out = self._run_forward_split(x, ts)
        else:
            out = self._run_forward(x, ts, None)
        if not return_dict:
            return (out,)
        return UNet2DOutput(sample=out)

    def _aview(self, off, shape):
        n = 1
        for s in shape:
            n *= s
        return self._arena[off:off + n].view(shape)

    def _bias(self, gobj: _Gemm):
        return self._arena[gobj.b_off:gobj.b_off + gobj.cout]

    def _run_forward(self, x: torch.Tensor, ts: torch.Tensor, tape: Optional[_Tape]) -> torch.Tensor:
        ops = _ops.get()
        l0 = ops.launches
        P = self._plan
        cfg = self.config
        training = tape is not None
        self._prepare_weights(training)
        ar = self._arena
        N, _, H, W = x.shape
        ted = self._temb_dim
        c0 = cfg.block_out_channels[0]

        # ---- time embedding: sinusoid -> linear_1 -> SiLU -> linear_2 ; all 32 time_emb_proj in one GEMM ----
        t_emb = ops.timestep_embedding(ts, c0, cfg.flip_sin_to_cos, float(cfg.freq_shift))
        w1, b1 = self._aview(P.te.w1, (ted, c0)), self._aview(P.te.b1, (ted,))
        w2, b2 = self._aview(P.te.w2, (ted, ted)), self._aview(P.te.b2, (ted,))
        wt, bt = self._aview(P.temb_w_off, (P.temb_total, ted)), self._aview(P.temb_b_off, (P.temb_total,))
        e1 = ops.linear_f32(t_emb, w1, b1, False)
        emb = ops.linear_f32(e1, w2, b2, True)
        temb_all = ops.linear_f32(emb, wt, bt, True)           # [N, sum C] fp32
        temb_lora = self._temb_lora_fwd(ops, emb, temb_all, training)
        st = SimpleNamespace(temb_all=temb_all, N=N, tape=tape, ops=ops,
                             d_temb_all=None, rng_tick=None, zeros=self._zero_pool("fwd", training, x.device))
        if training and any(g.lora is not None and g.lora.active and g.lora.p > 0.0 for g in P.gemms):
            # device-resident step counter for the adapter dropout: a CUDA-graph replay of the step draws new masks
            tick = getattr(self, "_rng_tick", None)
            if tick is None or tick.device != x.device:
                tick = self._rng_tick = torch.zeros(1, device=x.device, dtype=torch.int64)
            tick.add_(1)
            st.rng_tick = tick

        # ---- conv_in ----
        patches = ops.im2col3(x)                                   # [N, H, W, 64] bf16, one k-block
        cs = self._csum_for(st, (N, H, W), c0, generic=True)
        h = self._tag(ops.conv_gemm(patches, None, taps_1x1(), self._cin_wf, c0, (N, H, W),
                                    bias=self._aview(P.cin_b, (c0,)), csum=cs), cs)
        skips = [h]

        # ---- down ----
        for bi, b in enumerate(P.down):
            for j, r in enumerate(b.resnets):
                h = self._resnet_fwd(st, r, h, None, in_skip=len(skips) - 1 if h is skips[-1] else None)
                if b.attns:
                    h = self._attn_fwd(st, b.attns[j], h, in_skip=None)
                skips.append(h)
            if b.down is not None:
                h = self._down_fwd(st, b.down, h, in_skip=len(skips) - 1)
                skips.append(h)
        # ---- mid ----
        h = self._resnet_fwd(st, P.mid.r0, h, None, in_skip=len(skips) - 1)
        if P.mid.attn is not None:
            h = self._attn_fwd(st, P.mid.attn, h, in_skip=None)
        h = self._resnet_fwd(st, P.mid.r1, h, None, in_skip=None)
        # ---- up ----
        for b in P.up:
            for j, r in enumerate(b.resnets):
                sk = skips.pop()
                h = self._resnet_fwd(st, r, h, sk, in_skip=None, skip_idx=len(skips))
                if b.attns:
                    h = self._attn_fwd(st, b.attns[j], h, in_skip=None)
            if b.up is not None:
                h = self._up_fwd(st, b.up, h)
        # ---- out ----
        no = P.norm_out
        gam, bet = self._aview(no.g_off, (c0,)), self._aview(no.b_off, (c0,))
        coef_out = None
        if training and ops.gn_bwd_fusable((N, H, W)):
            stats, a, coef_out = self._gn_fwd(st, h, None, no.groups, no.eps, gam, bet, True, want_coef=True)
        else:
            stats, a = self._gn_fwd(st, h, None, no.groups, no.eps, gam, bet, True)
        o32 = ops.conv_gemm(a, None, taps_3x3(c0), self._cout_wf, 32, (N, H, W), bias=self._cout_b32, out_f32=True)
        out = ops.nhwc_to_nchw_f32(o32, cfg.out_channels)
        if training:
            tape.head = SimpleNamespace(patches=patches, t_emb=t_emb, e1=e1, emb=emb, h_last=h, stats=stats, a=a,
                                        coef=coef_out, temb_lora=temb_lora)
        self.last_launches = ops.launches - l0
        return out

    # ---------------------------------------------------------------------------------------------------------
    # fp32-faithful inference (north_star: eps within 1e-4 of the fp32 reference; SURVEY.md App. A.6: sampling and the
    # LoRA trainers run without autocast).  Activations are split-bf16 tensors [N, H, W, 2C] = [hi | lo]; every conv /
    # linear is ONE tcgen05 GEMM over the K-concat [A_hi | A_lo | A_hi] x [W_hi | W_hi | W_lo] (fp32 accumulation, the
    # dropped A_lo W_lo term is 2^-18 relative); GroupNorm / SiLU / softmax read hi + lo, compute in fp32 with exact
    # transcendentals and write split results.  3x the tensor work and 2x the activation bytes of the bf16 path.
    # ---------------------------------------------------------------------------------------------------------
    @staticmethod
    def _split3(w2d: torch.Tensor, taps: int, cin: int) -> torch.Tensor:
        """fp32 [cout, taps*cin] (physical [cout][tap][cin]) -> bf16 [cout, taps*3*cin]: per tap [W_hi | W_hi | W_lo]."""
        w = w2d.reshape(w2d.shape[0], taps, cin).float()
        hi = w.to(torch.bfloat16)
        lo = (w - hi.float()).to(torch.bfloat16)
        return torch.stack([hi, hi, lo], dim=2).reshape(w2d.shape[0], taps * 3 * cin).contiguous()

    def _effective_weight(self, gobj: _Gemm) -> torch.Tensor:
        """fp32 [cout, taps*cin] master weight of a GEMM layer, LoRA update (alpha/r) B A folded in when unmerged."""
        k = gobj.taps * gobj.cin
        rows = []
        o = gobj.w_off
        for w in gobj.weights:
            rows.append(self._arena[o:o + w.numel()].view(gobj.cout_each, k))
            o = _align(o + w.numel())
        W = torch.cat(rows, 0) if len(rows) > 1 else rows[0]
        if gobj.lora is not None and gobj.lora.active:
            W = W.clone()
            for i, m, _ in gobj.lora.slots:
                if not m.merged:
                    W[i * gobj.cout_each:(i + 1) * gobj.cout_each] += m.scaling * (m.B.detach().float() @ m.A.detach().float())
        return W

    def _split_weights(self):
        """Split operands of every layer, cached on the parameter versions."""
        P = self._plan
        key = (getattr(self, "_weights_epoch", 0),) + tuple(p._version for p in self.parameters())
        if self._w3_cache is not None and self._w3_cache[0] == key:
            return self._w3_cache[1]
        cfg = self.config
        c0, ci, co = cfg.block_out_channels[0], cfg.in_channels, cfg.out_channels
        W3 = {}
        with torch.no_grad():
            for gobj in P.gemms:
                W3[gobj.name] = self._split3(self._effective_weight(gobj), gobj.taps, gobj.cin)
            # conv_in: [c0, 128] against the im2col3_split rows: W_hi @0, W_hi @32, W_lo @64
            w_in = self._aview(P.cin_w, (c0, 9 * ci)).float()
            hi = w_in.to(torch.bfloat16)
            lo = (w_in - hi.float()).to(torch.bfloat16)
            wi = torch.zeros((c0, 128), device=w_in.device, dtype=torch.bfloat16)
            wi[:, :9 * ci], wi[:, 32:32 + 9 * ci], wi[:, 64:64 + 9 * ci] = hi, hi, lo
            W3["conv_in"] = wi
            w_out = torch.zeros((32, 9 * c0), device=w_in.device, dtype=torch.float32)
            w_out[:co] = self._aview(P.cout_w, (co, 9 * c0))
            W3["conv_out"] = self._split3(w_out, 9, c0)
            b32 = torch.zeros(32, device=w_in.device, dtype=torch.float32)
            b32[:co] = self._aview(P.cout_b, (co,))
            W3["conv_out.bias"] = b32
        self._w3_cache = (key, W3)
        return W3

    def _w3_slice(self, W3, gobj: _Gemm, lo_ch: int, n_ch: int) -> torch.Tensor:
        """Split operand of a 1x1 layer restricted to input channels [lo_ch, lo_ch + n_ch) (one source of a concat)."""
        key = (gobj.name, lo_ch, n_ch)
        w = W3.get(key)
        if w is None:
            with torch.no_grad():
                w = W3[key] = self._split3(self._effective_weight(gobj)[:, lo_ch:lo_ch + n_ch].contiguous(), 1, n_ch)
        return w

    def _run_forward_split(self, x: torch.Tensor, ts: torch.Tensor) -> torch.Tensor:
        ops = _ops.get()
        l0 = ops.launches
        P, cfg = self._plan, self.config
        W3 = self._split_weights()
        N, _, H, W = x.shape
        ted, c0 = self._temb_dim, cfg.block_out_channels[0]
        t_emb = ops.timestep_embedding(ts, c0, cfg.flip_sin_to_cos, float(cfg.freq_shift))
        e1 = ops.linear_f32(t_emb, self._aview(P.te.w1, (ted, c0)), self._aview(P.te.b1, (ted,)), False)
        emb = ops.linear_f32(e1, self._aview(P.te.w2, (ted, ted)), self._aview(P.te.b2, (ted,)), True)
        temb_all = ops.linear_f32(emb, self._aview(P.temb_w_off, (P.temb_total, ted)),
                                  self._aview(P.temb_b_off, (P.temb_total,)), True)
        self._temb_lora_fwd(ops, emb, temb_all, False)

        def conv(xs, gobj, taps_fn, cout, grid, **kw):
            """xs: split tensor [.., 2*cin]; one GEMM over [hi | lo | hi] against the layer's split operand."""
            cin = xs.shape[-1] // 2
            return ops.conv_gemm(xs, xs[..., :cin], taps_fn(3 * cin), W3[gobj.name], cout, grid,
                                 bias=self._bias(gobj), split_io=True, **kw)

        def resnet(r, x0, x1):
            n_, h_, w_, _ = x0.shape
            grid = (n_, h_, w_)
            g1, be1 = self._norm_params(r.norm1)
            g2, be2 = self._norm_params(r.norm2)
            a = ops.gn_fwd_split(x0, x1, r.norm1.groups, r.norm1.eps, g1, be1, True)
            temb = temb_all[:, r.temb_off:r.temb_off + r.cout]
            h1 = conv(a, r.conv1, taps_3x3, r.cout, grid, temb=temb)
            b = ops.gn_fwd_split(h1, None, r.norm2.groups, r.norm2.eps, g2, be2, True)
            if r.short is None:
                sc = x0
            else:
                ca = x0.shape[-1] // 2
                sc = ops.conv_gemm(x0, x0[..., :ca], taps_1x1(), self._w3_slice(W3, r.short, 0, ca), r.cout, grid,
                                   bias=self._bias(r.short), split_io=True)
                if x1 is not None:     # the concat's second source accumulates onto the first through the residual input
                    cb = x1.shape[-1] // 2
                    sc = ops.conv_gemm(x1, x1[..., :cb], taps_1x1(), self._w3_slice(W3, r.short, ca, cb), r.cout, grid,
                                       res=sc, split_io=True)
            return conv(b, r.conv2, taps_3x3, r.cout, grid, res=sc)

        def attn(at, xs):
            n_, h_, w_, c2 = xs.shape
            C, T = c2 // 2, h_ * w_
            M = n_ * T
            gam, bet = self._norm_params(at.norm)
            xn = ops.gn_fwd_split(xs, None, at.norm.groups, at.norm.eps, gam, bet, False).view(1, 1, M, 2 * C)
            qkv = conv(xn, at.qkv, lambda k: taps_1x1(), 3 * C, (1, 1, M))
            o = ops.attn_fwd_split(qkv.view(M, 6 * C), n_, T, at.heads, at.d, at.d ** -0.5).view(1, 1, M, 2 * C)
            return conv(o, at.out, lambda k: taps_1x1(), C, (1, 1, M), res=xs.view(1, 1, M, 2 * C)).view(n_, h_, w_, 2 * C)

        patches = ops.im2col3_split(x)
        h = ops.conv_gemm(patches, None, taps_1x1(), W3["conv_in"], c0, (N, H, W), bias=self._aview(P.cin_b, (c0,)),
                          split_io=True)
        skips = [h]
        for b in P.down:
            for j, r in enumerate(b.resnets):
                h = resnet(r, h, None)
                if b.attns:
                    h = attn(b.attns[j], h)
                skips.append(h)
            if b.down is not None:
                n_, h_, w_, c2 = h.shape
                s2d = ops.space_to_depth(h)
                C = c2 // 2
                h = ops.conv_gemm(s2d, s2d[..., :C], taps_s2d(3 * C, n_, b.down.pad), W3[b.down.conv.name], C,
                                  (n_, h_ // 2, w_ // 2), bias=self._bias(b.down.conv), src_n=4 * n_, split_io=True)
                skips.append(h)
        h = resnet(P.mid.r0, h, None)
        if P.mid.attn is not None:
            h = attn(P.mid.attn, h)
        h = resnet(P.mid.r1, h, None)
        for b in P.up:
            for j, r in enumerate(b.resnets):
                h = resnet(r, h, skips.pop())
                if b.attns:
                    h = attn(b.attns[j], h)
            if b.up is not None:
                up = ops.upsample2x(h)
                n_, h_, w_, _ = up.shape
                h = conv(up, b.up.conv, taps_3x3, b.up.c, (n_, h_, w_))
        no = P.norm_out
        gam, bet = self._aview(no.g_off, (c0,)), self._aview(no.b_off, (c0,))
        a = ops.gn_fwd_split(h, None, no.groups, no.eps, gam, bet, True)
        o32 = ops.conv_gemm(a, a[..., :c0], taps_3x3(3 * c0), W3["conv_out"], 32, (N, H, W),
                            bias=W3["conv_out.bias"], out_f32=True)
        out = ops.nhwc_to_nchw_f32(o32, cfg.out_channels)
        self.last_launches = ops.launches - l0
        return out

    # ---- LoRA on time_emb_proj (config_diffusion.py:37 lists it among the candidate targets) ---------------------
    # y_r += (alpha / r) * B_r (A_r dropout(SiLU(emb))) on the [batch, 512] embedding: fp32 side computation with the
    # tiled SIMT linears, per adapter (each has its own dropout mask, drawn by torch's capture-safe Philox).
    def _temb_lora_fwd(self, ops, emb, temb_all, training: bool):
        recs = [r for r in self._plan.resnets if r.temb_lora is not None and not r.temb_lora.merged]
        if not recs:
            return None
        xs = torch.nn.functional.silu(emb)
        saved = []
        for r in recs:
            m = r.temb_lora
            keep = None
            if training and m.p > 0.0:      # nn.Dropout: keep with probability 1 - p, scale the kept by 1 / (1 - p)
                keep = (torch.rand_like(xs) >= m.p).to(xs.dtype).mul_(1.0 / (1.0 - m.p))
            xd = xs * keep if keep is not None else xs
            u = ops.linear_f32(xd, m.A.detach(), None, False)                   # [N, rank]
            y = ops.linear_f32(u, m.B.detach(), None, False)                    # [N, cout]
            temb_all[:, r.temb_off:r.temb_off + r.cout].add_(y, alpha=m.scaling)
            saved.append((r, xd, u, keep))
        return saved if training else None

    def _temb_lora_bwd(self, ops, saved, d_temb_all, zeros, need_dx: bool):
        """dA / dB of every time_emb_proj adapter from the per-sample sums of d_h1 (d_temb_all) -> ({id(param): grad},
        d_xs): d_xs = the adapters' gradient w.r.t. SiLU(emb) (None unless need_dx: a trainable time-embedding MLP)."""
        grads, d_xs = {}, None
        for r, xd, u, keep in saved:
            m = r.temb_lora
            g = d_temb_all[:, r.temb_off:r.temb_off + r.cout].contiguous()
            dB = zeros((r.cout, m.r), g.device)
            ops.linear_f32_wgrad(u, g, dB, None, False)
            dU = ops.linear_f32_dgrad(g, m.B.detach(), None, False)              # [N, rank], still without alpha / r
            dU.mul_(m.scaling)
            dA = zeros((m.r, xd.shape[1]), g.device)
            ops.linear_f32_wgrad(xd, dU, dA, None, False)
            grads[id(m.A)] = dA
            grads[id(m.B)] = dB.mul_(m.scaling)
            if need_dx:
                dx = ops.linear_f32_dgrad(dU, m.A.detach(), None, False)         # [N, 512]
                if keep is not None:
                    dx.mul_(keep)
                d_xs = dx if d_xs is None else d_xs.add_(dx)
        return grads, d_xs

    def _temb_lora_active(self):
        return any(r.temb_lora is not None and not r.temb_lora.merged for r in self._plan.resnets)

    # ---- blocks: forward (each records its backward closure on the tape) -------------------------------------
    def _norm_params(self, nobj: _Norm):
        c = nobj.mod.num_channels
        return self._aview(nobj.g_off, (c,)), self._aview(nobj.b_off, (c,))

    # GroupNorm statistics ride on the epilogue of the 3x3 conv that PRODUCES the tensor (per-(sample, channel) moments,
    # conv_gemm(csum=...)); the GroupNorm forward is then one streaming pass instead of the two-phase team kernel.
    # (Not for conv_in / the stride-2 convs: they run on the generic kernel, where the reductions are exposed.)
    def _csum_for(self, st, grid, cout, generic=False):
        """generic: the producing conv runs on the generic kernel (conv_in, stride-2 convs), where the reductions are not
        hidden behind the next tile's mainloop -- taken only for memory-bound producers at >= 64x64, whose consumers then
        run the single-pass GroupNorm instead of the two-phase team kernel (in-process A/B: training step -0.8 %, sampling
        forward -0.5 %; DDPM_GN_STATS_GENERIC=0 switches it off)."""
        ops = st.ops
        if generic:
            if os.environ.get("DDPM_GN_STATS_GENERIC", "1") == "0" or grid[1] * grid[2] < 4096:
                return None
            if os.environ.get("DDPM_GN_STATS_FUSION", "1") == "0":
                return None
        elif os.environ.get("DDPM_GN_STATS_FUSION", "1") == "0" or not ops.gn_stats_fusable(grid):
            return None
        # the statistics are kept per 4-channel granule: usable when the consuming GroupNorm's groups are whole granules,
        # which holds whenever every concatenated source is a multiple of 4 * norm_num_groups channels (128 here)
        if cout % (4 * self.config.norm_num_groups):
            return None
        return st.zeros.take((grid[0], cout // 4, 2), st.temb_all.device)

    @staticmethod
    def _tag(t, csum):
        if csum is not None:
            t._ddpm_csum = csum
        return t

    @staticmethod
    def _gn_fwd(st, x0, x1, groups, eps, gam, bet, silu, want_coef=False):
        ops = st.ops
        cs0 = getattr(x0, "_ddpm_csum", None)
        cs1 = getattr(x1, "_ddpm_csum", None) if x1 is not None else None
        if cs0 is not None and (x1 is None or cs1 is not None):
            return ops.gn_fwd_from_csum(x0, x1, cs0, cs1, groups, eps, gam, bet, silu, want_coef=want_coef)
        return ops.gn_fwd(x0, x1, groups, eps, gam, bet, silu, want_coef=want_coef)

    def _resnet_fwd(self, st, r, x0, x1, in_skip=None, skip_idx=None):
        ops = st.ops
        N, H, W, _ = x0.shape
        grid = (N, H, W)
        g1, be1 = self._norm_params(r.norm1)
        g2, be2 = self._norm_params(r.norm2)
        fuse = st.tape is not None and ops.gn_bwd_fusable(grid)   # backward runs its first GN half in the dgrad epilogue
        coef1 = coef2 = None
        if fuse:
            stats1, a, coef1 = self._gn_fwd(st, x0, x1, r.norm1.groups, r.norm1.eps, g1, be1, True, want_coef=True)
        else:
            stats1, a = self._gn_fwd(st, x0, x1, r.norm1.groups, r.norm1.eps, g1, be1, True)
        temb = st.temb_all[:, r.temb_off:r.temb_off + r.cout]
        cs = self._csum_for(st, grid, r.cout)
        h1 = self._tag(ops.conv_gemm(a, None, taps_3x3(r.cin), r.conv1.wf, r.cout, grid, bias=self._bias(r.conv1),
                                     temb=temb, csum=cs), cs)
        if fuse:
            stats2, b, coef2 = self._gn_fwd(st, h1, None, r.norm2.groups, r.norm2.eps, g2, be2, True, want_coef=True)
        else:
            stats2, b = self._gn_fwd(st, h1, None, r.norm2.groups, r.norm2.eps, g2, be2, True)
        if r.short is not None:
            sc = ops.conv_gemm(x0, x1, taps_1x1(), r.short.wf, r.cout, grid, bias=self._bias(r.short))
        else:
            sc = x0
        cs = self._csum_for(st, grid, r.cout)
        out = self._tag(ops.conv_gemm(b, None, taps_3x3(r.cout), r.conv2.wf, r.cout, grid, bias=self._bias(r.conv2),
                                      res=sc, csum=cs), cs)
        if st.tape is not None:
            st.tape.add(("resnet", r, SimpleNamespace(x0=x0, x1=x1, stats1=stats1, a=a, h1=h1, stats2=stats2, b=b,
                                                       in_skip=in_skip, skip_idx=skip_idx, grid=grid, coef1=coef1,
                                                       coef2=coef2)))
        return out

    def _attn_fwd(self, st, at, x, in_skip=None):
        ops = st.ops
        N, H, W, C = x.shape
        T = H * W
        gam, bet = self._norm_params(at.norm)
        stats, xn = ops.gn_fwd(x, None, at.norm.groups, at.norm.eps, gam, bet, False)
        xn2 = xn.view(1, 1, N * T, C)
        lora_qkv = at.qkv.lora.forward_extra(ops, xn2, self.training, st.rng_tick) \
            if (at.qkv.lora and at.qkv.lora.active) else None
        qkv = ops.conv_gemm(xn2, lora_qkv.u if lora_qkv else None, self._lin_taps(at.qkv), at.qkv.wf, 3 * C,
                            (1, 1, N * T), bias=self._bias(at.qkv))
        o, lse = ops.attn_fwd(qkv.view(N * T, 3 * C), N, T, at.heads, at.d, at.d ** -0.5, need_aux=st.tape is not None)
        o2 = o.view(1, 1, N * T, C)
        lora_o = at.out.lora.forward_extra(ops, o2, self.training, st.rng_tick) \
            if (at.out.lora and at.out.lora.active) else None
        out = ops.conv_gemm(o2, lora_o.u if lora_o else None, self._lin_taps(at.out), at.out.wf, C, (1, 1, N * T),
                            bias=self._bias(at.out), res=x.view(1, 1, N * T, C)).view(N, H, W, C)
        if st.tape is not None:
            st.tape.add(("attn", at, SimpleNamespace(x=x, stats=stats, xn=xn, qkv=qkv, o=o, lse=lse, in_skip=in_skip,
                                                     lora_qkv=lora_qkv, lora_o=lora_o)))
        return out

    @staticmethod
    def _lin_taps(gobj: _Gemm):
        return taps_1x1()

    def _down_fwd(self, st, d, x, in_skip=None):
        ops = st.ops
        N, H, W, C = x.shape
        s2d = ops.space_to_depth(x)
        grid = (N, H // 2, W // 2)
        cs = self._csum_for(st, grid, C, generic=True)
        out = self._tag(ops.conv_gemm(s2d, None, taps_s2d(C, N, d.pad), d.conv.wf, C, grid, bias=self._bias(d.conv),
                                      src_n=4 * N, csum=cs), cs)
        if st.tape is not None:
            st.tape.add(("down", d, SimpleNamespace(s2d=s2d, in_skip=in_skip, shape=(N, H, W, C), grid=grid)))
        return out

    def _up_fwd(self, st, u, x):
        ops = st.ops
        N, H, W, C = x.shape
        up = ops.upsample2x(x)
        cs = self._csum_for(st, (N, 2 * H, 2 * W), C)
        out = self._tag(ops.conv_gemm(up, None, taps_3x3(C), u.conv.wf, C, (N, 2 * H, 2 * W), bias=self._bias(u.conv),
                                      csum=cs), cs)
        if st.tape is not None:
            st.tape.add(("up", u, SimpleNamespace(up=up, grid=(N, 2 * H, 2 * W))))
        return out

    # ---------------------------------------------------------------------------------------------------------
    # backward
    # ---------------------------------------------------------------------------------------------------------
    def _gview(self, G, off, shape):
        n = 1
        for s in shape:
            n *= s
        return G[off:off + n].view(shape)

    def _wgrad_views(self, G, gobj: _Gemm):
        k = gobj.taps * gobj.cin
        return self._gview(G, gobj.w_off, (gobj.cout, k)), self._gview(G, gobj.b_off, (gobj.cout,))

    def _run_backward(self, tape: _Tape, d_out: torch.Tensor):
        """d_out: fp32 NCHW grad of the prediction.  Returns the flat gradient arena (fp32)."""
        ops = _ops.get()
        l0 = ops.launches
        P = self._plan
        cfg = self.config
        hd = tape.head
        N = hd.patches.shape[0]
        c0 = cfg.block_out_channels[0]
        ted = self._temb_dim
        G = self._fresh_grad_arena()
        zp = self._zero_pool("bwd", True, G.device)
        d_temb_all = zp.take((N, P.temb_total), G.device)
        st = SimpleNamespace(ops=ops, G=G, d_temb_all=d_temb_all, skip_grads={}, N=N, wg_stream=None, keep=[],
                             defer_kw={}, zeros=zp)
        if G.is_cuda and os.environ.get("DDPM_WGRAD_STREAM", "1") != "0":
            # Weight gradients feed nothing downstream in backward, so they run on a second stream: the tensor-bound
            # wgrad GEMMs overlap the HBM-bound GroupNorm-backward kernels and the latency-bound low-resolution layers
            # (in a captured step these become parallel graph branches).  Joined before the arena is handed back.
            if self._wgrad_stream is None or self._wgrad_stream.device != G.device:
                self._wgrad_stream = torch.cuda.Stream(device=G.device)
            st.wg_stream = self._wgrad_stream
            st.defer_kw = {"defer": lambda fn: self._async_wgrad(st, fn)}     # GroupNorm dgamma / dbeta reductions too
        d_out = d_out.to(torch.float32).contiguous()

        # earliest tape step that still has trainable parameters at or before it (LoRA: stop early)
        first_needed = 0
        if not self._head_trainable():
            first_needed = len(tape.steps)
            for i, (kind, rec, _) in enumerate(tape.steps):
                if self._step_trainable(kind, rec):
                    first_needed = i
                    break

        # ---- conv_out / conv_norm_out ----
        no = P.norm_out
        co = cfg.out_channels
        grid0 = tuple(hd.a.shape[:3])
        train_out = self.conv_out.weight.requires_grad
        pd = ops.im2col3(d_out, chan_sum=self._gview(G, P.cout_b, (co,)) if train_out else None)
        if train_out:   # R[ci][tap'*co_n + co] = sum_pix a[pix, ci] * d_out[pix + off(tap'), co]
            R = zp.take((c0, 64), G.device)
            ops.conv_wgrad(hd.a, pd, None, taps_1x1(), R, grid0)
            self._gview(G, P.cout_w, (co, 9, c0)).add_(R[:, :9 * co].view(c0, 9, co).flip(1).permute(2, 1, 0))
        gam, bet = self._norm_params(no)
        tr = no.trainable
        dg_o = self._gview(G, no.g_off, (c0,)) if tr else None
        db_o = self._gview(G, no.b_off, (c0,)) if tr else None
        n_steps = len(tape.steps)
        st.bias_done = False
        if hd.coef is not None:    # SiLU / GroupNorm derivative + per-(n, c) sums in the dgrad epilogue, one streaming pass
            sums = zp.take((grid0[0], c0, 2), G.device)
            dz = ops.conv_gemm(pd, None, taps_1x1(), self._cout_wd, c0, grid0, gn=(hd.h_last, None, hd.coef, True, sums))
            st.next_bias = self._colsum_target(G, tape, n_steps - 1, first_needed)
            g, _ = ops.gn_bwd_apply(hd.h_last, None, no.groups, hd.stats, no.eps, gam, dz, sums, dgamma=dg_o, dbeta=db_o,
                                    out_c=st.next_bias, **st.defer_kw)
            st.bias_done = st.next_bias is not None
        else:
            d_a = ops.conv_gemm(pd, None, taps_1x1(), self._cout_wd, c0, grid0)
            g, _ = ops.gn_bwd(hd.h_last, None, no.groups, hd.stats, no.eps, gam, bet, True, d_a, dgamma=dg_o,
                              dbeta=db_o, **st.defer_kw)

        prog = getattr(self, "_grad_progress_hook", None)
        begin = getattr(self, "_grad_begin_hook", None)
        if begin is not None:
            begin()
        for i in range(len(tape.steps) - 1, first_needed - 1, -1):
            kind, rec, s = tape.steps[i]
            # the bias gradient of the conv that consumes this step's result (= the column sums of the gradient this step
            # returns) rides on the step's last streaming kernel where it has one: no separate pass over that tensor
            st.next_bias = self._colsum_target(G, tape, i - 1, first_needed)
            if kind == "resnet":
                g = self._resnet_bwd(st, rec, s, g)
            elif kind == "attn":
                g = self._attn_bwd(st, rec, s, g)
            elif kind == "down":
                g = self._down_bwd(st, rec, s, g)
            else:
                g = self._up_bwd(st, rec, s, g)
            if prog is not None:   # DDP: weight grads at arena offsets >= this step's first layer are final
                prog(G, rec.conv1.w_off if kind == "resnet" else (rec.qkv.w_off if kind == "attn" else rec.conv.w_off))

        if first_needed == 0 and self._head_trainable():
            # g is now the gradient of conv_in's output (skip 0 already folded in by the first resnet)
            if self.conv_in.weight.requires_grad:
                if not st.bias_done:
                    ops.reduce_hw(g, None, self._gview(G, P.cin_b, (c0,)))
                R = zp.take((c0, 64), G.device)
                ops.conv_wgrad(g, hd.patches, None, taps_1x1(), R, tuple(g.shape[:3]))
                self._gview(G, P.cin_w, (c0, 9 * cfg.in_channels)).add_(R[:, :9 * cfg.in_channels])
        if st.wg_stream is not None:      # d_temb_all and every weight gradient are complete from here on
            torch.cuda.current_stream().wait_stream(st.wg_stream)
        # ---- time_emb_proj adapters ----
        st.temb_lora_grads, d_xs_lora = {}, None
        if hd.temb_lora:
            st.temb_lora_grads, d_xs_lora = self._temb_lora_bwd(
                ops, hd.temb_lora, d_temb_all, zp.take, self.time_embedding.linear_1.weight.requires_grad)
        # ---- time-embedding MLP ----
        if self._temb_trainable() or self.time_embedding.linear_1.weight.requires_grad:
            wt = self._aview(P.temb_w_off, (P.temb_total, ted))
            w2 = self._aview(P.te.w2, (ted, ted))
            if self._temb_trainable():
                ops.linear_f32_wgrad(hd.emb, d_temb_all, self._gview(G, P.temb_w_off, (P.temb_total, ted)),
                                     self._gview(G, P.temb_b_off, (P.temb_total,)), True)
            if self.time_embedding.linear_1.weight.requires_grad:
                d_emb = ops.linear_f32_dgrad(d_temb_all, wt, hd.emb, True)
                if d_xs_lora is not None:       # the adapters read SiLU(emb) too: d_emb += d_xs * silu'(emb)
                    sg = torch.sigmoid(hd.emb)
                    d_emb.add_(d_xs_lora * (sg * (1.0 + hd.emb * (1.0 - sg))))
                ops.linear_f32_wgrad(hd.e1, d_emb, self._gview(G, P.te.w2, (ted, ted)), self._gview(G, P.te.b2, (ted,)),
                                     True)
                d_e1 = ops.linear_f32_dgrad(d_emb, w2, hd.e1, True)
                ops.linear_f32_wgrad(hd.t_emb, d_e1, self._gview(G, P.te.w1, (ted, c0)),
                                     self._gview(G, P.te.b1, (ted,)), False)
        self.last_launches_bwd = ops.launches - l0
        return G, st

    def _colsum_target(self, G, tape, i: int, first_needed: int):
        """The bias gradient that equals the pixel sums of the gradient ENTERING tape step i (i = -1: conv_in), or None."""
        if os.environ.get("DDPM_BIAS_FUSION", "1") == "0":
            return None
        if i < first_needed:
            if i == -1 and first_needed == 0 and self.conv_in.weight.requires_grad:
                return self._gview(G, self._plan.cin_b, (self.config.block_out_channels[0],))
            return None
        kind, rec, _ = tape.steps[i]
        if kind == "resnet":
            if rec.conv2.bias_trainable:
                return self._wgrad_views(G, rec.conv2)[1]
            if rec.short is not None and rec.short.bias_trainable:
                return self._wgrad_views(G, rec.short)[1]
            return None
        gobj = rec.out if kind == "attn" else rec.conv
        return self._wgrad_views(G, gobj)[1] if gobj.bias_trainable else None

    def _zero_pool(self, which: str, training: bool, device) -> _ZeroPool:
        pools = self.__dict__.setdefault("_zero_pools", {})
        key = (which, training)
        zp = pools.get(key)
        if zp is None:
            zp = pools[key] = _ZeroPool()
        return zp.begin(device)

    def _fresh_grad_arena(self) -> torch.Tensor:
        """Zero-filled flat fp32 gradient arena.  The previous step's arena is reused (one memset, no allocation) unless a
        parameter's .grad still aliases it -- gradient accumulation (train_epoch_with_accumulation) adds the new
        gradients into the old ones, so those steps get a fresh buffer."""
        P, dev = self._plan, self._arena.device
        capturing = dev.type == "cuda" and torch.cuda.is_current_stream_capturing()
        G = getattr(self, "_grad_arena", None)
        if G is not None and G.device == dev and G.numel() == P.total and not capturing:
            lo, hi = G.data_ptr(), G.data_ptr() + 4 * G.numel()
            live = any(p.grad is not None and lo <= p.grad.data_ptr() < hi for p, _, _ in P.layout)
            if not live:
                return G.zero_()
        G = torch.zeros(P.total, device=dev, dtype=torch.float32)
        if not capturing:
            self._grad_arena = G        # (a captured step owns its arena inside the graph's private pool)
        return G

    @staticmethod
    def _async_wgrad(st, fn):
        """Run fn() -- weight-gradient kernels whose results backward itself never reads -- on the wgrad stream.  The
        closure keeps its operand tensors alive until backward has joined the stream (st.keep)."""
        side = st.wg_stream
        if side is None:
            fn()
            return
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):
            fn()
        st.keep.append(fn)

    def _head_trainable(self):
        return self.conv_in.weight.requires_grad or self.time_embedding.linear_1.weight.requires_grad or \
            self._temb_trainable() or self._temb_lora_active()

    def _temb_trainable(self):
        return any(self._base(r.mod.time_emb_proj).weight.requires_grad for r in self._plan.resnets)

    def _step_trainable(self, kind, rec):
        if kind == "resnet":
            gl = [rec.conv1, rec.conv2] + ([rec.short] if rec.short else [])
            return any(x.trainable or x.bias_trainable for x in gl) or rec.norm1.trainable or rec.norm2.trainable or \
                (rec.temb_lora is not None and not rec.temb_lora.merged)
        if kind == "attn":
            return any(x.trainable or x.bias_trainable or x.lora is not None for x in (rec.qkv, rec.out)) or \
                rec.norm.trainable
        return rec.conv.trainable or rec.conv.bias_trainable

    def _norm_grads(self, G, nobj):
        if not nobj.trainable:
            return None, None
        c = nobj.mod.num_channels
        return self._gview(G, nobj.g_off, (c,)), self._gview(G, nobj.b_off, (c,))

    def _resnet_bwd(self, st, r, s, g):
        ops, G = st.ops, st.G
        grid = s.grid
        # conv2 (+ shortcut bias shares the same column sums)
        dW2, db2 = self._wgrad_views(G, r.conv2)
        done, st.bias_done = st.bias_done, False     # the producer of g already added its pixel sums to the bias gradient
        fold_bias = r.conv2.trainable and r.conv2.bias_trainable and not done    # bias gradient rides on the wgrad GEMM
        if done:
            pass
        elif r.conv2.bias_trainable and not fold_bias:
            ops.reduce_hw(g, None, db2)
        elif not r.conv2.bias_trainable and r.short is not None and r.short.bias_trainable:
            ops.reduce_hw(g, None, self._wgrad_views(G, r.short)[1])
        share_bias = r.conv2.bias_trainable and r.short is not None and r.short.bias_trainable

        def wgrad2(g=g):
            if r.conv2.trainable:
                ops.conv_wgrad(g, s.b, None, taps_3x3(r.cout), dW2, grid, dbias=db2 if fold_bias else None)
            if share_bias:
                self._wgrad_views(G, r.short)[1].copy_(db2)       # the shortcut bias sees the same column sums
        if fold_bias or done or not r.conv2.bias_trainable:
            self._async_wgrad(st, wgrad2)
        else:
            wgrad2()        # db2 came from reduce_hw on the main stream: keep the copy ordered behind it
        g2, be2 = self._norm_params(r.norm2)
        dg, dbt = self._norm_grads(G, r.norm2)
        fuse = s.coef2 is not None
        if fuse:   # SiLU/GroupNorm derivative + per-(n, c) sums in the dgrad epilogue, then one streaming pass
            sums = st.zeros.take((grid[0], r.cout, 2), g.device)
            dz = ops.conv_gemm(g, None, taps_3x3(r.cout), r.conv2.wd, r.cout, grid,
                               gn=(s.h1, None, s.coef2, True, sums))
            # the pixel sums of d_h1 (time-embedding gradient per sample, conv1 bias gradient) come out of the same pass
            d_h1, _ = ops.gn_bwd_apply(s.h1, None, r.norm2.groups, s.stats2, r.norm2.eps, g2, dz, sums,
                                       dgamma=dg, dbeta=dbt,
                                       out_nc=st.d_temb_all[:, r.temb_off:r.temb_off + r.cout],
                                       out_c=self._wgrad_views(G, r.conv1)[1] if r.conv1.bias_trainable else None,
                                       **st.defer_kw)
        else:
            d_b = ops.conv_gemm(g, None, taps_3x3(r.cout), r.conv2.wd, r.cout, grid)
            d_h1, _ = ops.gn_bwd(s.h1, None, r.norm2.groups, s.stats2, r.norm2.eps, g2, be2, True, d_b, dgamma=dg,
                                 dbeta=dbt, **st.defer_kw)
        # time embedding + conv1 bias share sum_hw(d_h1)
        dW1, db1 = self._wgrad_views(G, r.conv1)
        if not fuse:     # consumed only by the time-embedding MLP backward at the very end: off the critical path
            self._async_wgrad(st, lambda d_h1=d_h1: ops.reduce_hw(
                d_h1, st.d_temb_all[:, r.temb_off:r.temb_off + r.cout], db1 if r.conv1.bias_trainable else None))
        if r.conv1.trainable:
            self._async_wgrad(st, lambda d_h1=d_h1: ops.conv_wgrad(d_h1, s.a, None, taps_3x3(r.cin), dW1, grid))
        g1, be1 = self._norm_params(r.norm1)
        if fuse:
            sums1 = st.zeros.take((grid[0], r.cin, 2), g.device)
            d_a = ops.conv_gemm(d_h1, None, taps_3x3(r.cout), r.conv1.wd, r.cin, grid,
                                gn=(s.x0, s.x1, s.coef1, True, sums1))
        else:
            d_a = ops.conv_gemm(d_h1, None, taps_3x3(r.cout), r.conv1.wd, r.cin, grid)
        if r.short is not None:
            if r.short.trainable:
                self._async_wgrad(st, lambda g=g: ops.conv_wgrad(g, s.x0, s.x1, taps_1x1(),
                                                                 self._wgrad_views(G, r.short)[0], grid))
            d_sc = ops.conv_gemm(g, None, taps_1x1(), r.short.wd, r.cin, grid)
        else:
            d_sc = g
        extra = st.skip_grads.pop(s.in_skip, None) if s.in_skip is not None else None
        dg, dbt = self._norm_grads(G, r.norm1)
        if fuse:
            nb = st.next_bias
            oc = nb
            if nb is not None and s.x1 is not None:      # the kernel sums all c0 + c1 channels: keep the first c0
                oc = st.zeros.take((r.cin,), g.device)
            dx0, dx1 = ops.gn_bwd_apply(s.x0, s.x1, r.norm1.groups, s.stats1, r.norm1.eps, g1, d_a, sums1, add0=d_sc,
                                        add1=extra, dgamma=dg, dbeta=dbt, out_c=oc, **st.defer_kw)
            if nb is not None:
                if oc is not nb:
                    nb.add_(oc[:nb.numel()])
                st.bias_done = True
        else:
            dx0, dx1 = ops.gn_bwd(s.x0, s.x1, r.norm1.groups, s.stats1, r.norm1.eps, g1, be1, True, d_a, add0=d_sc,
                                  add1=extra, dgamma=dg, dbeta=dbt, **st.defer_kw)
        if s.x1 is not None:
            st.skip_grads[s.skip_idx] = dx1
        return dx0

    def _attn_bwd(self, st, at, s, g):
        ops, G = st.ops, st.G
        N, H, W, C = s.x.shape
        T = H * W
        M = N * T
        g2 = g.view(1, 1, M, C)
        dWo, dbo = self._wgrad_views(G, at.out)
        done, st.bias_done = st.bias_done, False
        if at.out.bias_trainable and not done:
            ops.reduce_hw(g2, None, dbo)
        o2 = s.o.view(1, 1, M, C)
        if at.out.trainable:
            self._async_wgrad(st, lambda: ops.conv_wgrad(g2, o2, None, taps_1x1(), dWo, (1, 1, M)))
        d_o = ops.conv_gemm(g2, None, taps_1x1(), at.out.wd, C, (1, 1, M))
        if s.lora_o is not None:
            d_o = at.out.lora.backward(ops, s.lora_o, o2, g2, d_o, zeros=st.zeros.take)
        dqkv = ops.attn_bwd(s.qkv.view(M, 3 * C), s.o, d_o.view(M, C), s.lse, N, T, at.heads, at.d, at.d ** -0.5)
        dq2 = dqkv.view(1, 1, M, 3 * C)
        dWq, dbq = self._wgrad_views(G, at.qkv)
        if at.qkv.bias_trainable:
            ops.reduce_hw(dq2, None, dbq)
        xn2 = s.xn.view(1, 1, M, C)
        if at.qkv.trainable:
            self._async_wgrad(st, lambda: ops.conv_wgrad(dq2, xn2, None, taps_1x1(), dWq, (1, 1, M)))
        d_xn = ops.conv_gemm(dq2, None, taps_1x1(), at.qkv.wd, C, (1, 1, M))
        if s.lora_qkv is not None:
            d_xn = at.qkv.lora.backward(ops, s.lora_qkv, xn2, dq2, d_xn, zeros=st.zeros.take)
        extra = st.skip_grads.pop(s.in_skip, None) if s.in_skip is not None else None
        gam, bet = self._norm_params(at.norm)
        dg, dbt = self._norm_grads(G, at.norm)
        dx, _ = ops.gn_bwd(s.x, None, at.norm.groups, s.stats, at.norm.eps, gam, bet, False, d_xn.view(N, H, W, C),
                           add0=g, add1=extra, dgamma=dg, dbeta=dbt, **st.defer_kw)
        return dx

    def _down_bwd(self, st, d, s, g):
        ops, G = st.ops, st.G
        N, H, W, C = s.shape
        dW, db = self._wgrad_views(G, d.conv)
        done, st.bias_done = st.bias_done, False
        fold_bias = d.conv.trainable and d.conv.bias_trainable and not done
        if d.conv.bias_trainable and not fold_bias and not done:
            ops.reduce_hw(g, None, db)
        if d.conv.trainable:
            self._async_wgrad(st, lambda g=g: ops.conv_wgrad(g, s.s2d, None, taps_s2d(C, N, d.pad), dW, s.grid,
                                                             src_n=4 * N, dbias=db if fold_bias else None))
        zi = ops.zero_insert2x(g, H, W)
        extra = st.skip_grads.pop(s.in_skip, None) if s.in_skip is not None else None
        # dgrad of the stride-2 conv = stride-1 correlation of the zero-inserted dY with the flipped taps
        shift = 1 if d.pad == 1 else 2
        taps = [(0, r - shift, q - shift, (r * 3 + q) * C) for r in range(3) for q in range(3)]
        return ops.conv_gemm(zi, None, taps, d.conv.wd, C, (N, H, W), res=extra)

    def _up_bwd(self, st, u, s, g):
        ops, G = st.ops, st.G
        C = u.c
        dW, db = self._wgrad_views(G, u.conv)
        done, st.bias_done = st.bias_done, False
        fold_bias = u.conv.trainable and u.conv.bias_trainable and not done
        if u.conv.bias_trainable and not fold_bias and not done:
            ops.reduce_hw(g, None, db)
        if u.conv.trainable:
            self._async_wgrad(st, lambda g=g: ops.conv_wgrad(g, s.up, None, taps_3x3(C), dW, s.grid,
                                                             dbias=db if fold_bias else None))
        d_up = ops.conv_gemm(g, None, taps_3x3(C), u.conv.wd, C, s.grid)
        return ops.sumpool2x(d_up)

    # ---- gradients as per-parameter views ------------------------------------------------------------------------
    def _grad_views(self, G: torch.Tensor, st) -> Dict[int, torch.Tensor]:
        out = {}
        for p, o, phys in self._plan.layout:
            seg = G[o:o + p.numel()]
            out[id(p)] = seg.view(phys).permute(0, 3, 1, 2) if phys is not None else seg.view(p.shape)
        for gobj in self._plan.gemms:
            if gobj.lora is not None:
                out.update(gobj.lora.grads)
        out.update(getattr(st, "temb_lora_grads", {}))
        return out


class _UNetFunction(torch.autograd.Function):
    """The whole UNet as one autograd node: forward records a tape, backward replays it in hand-written kernels."""

    @staticmethod
    def forward(ctx, model: UNet2DModel, x, ts, nparams, *params):
        tape = _Tape()
        out = model._run_forward(x, ts, tape)
        ctx.model, ctx.tape, ctx.params = model, tape, params
        return out

    @staticmethod
    def backward(ctx, d_out):
        model = ctx.model
        G, st = model._run_backward(ctx.tape, d_out)
        views = model._grad_views(G, st)
        hook = getattr(model, "_grad_ready_hook", None)
        if hook is not None:              # DDP: all-reduce the flat gradient arena (+ LoRA grads)
            hook(G, [g for gobj in model._plan.gemms if gobj.lora is not None for g in gobj.lora.grads.values()] +
                 list(getattr(st, "temb_lora_grads", {}).values()))
        grads = tuple(views.get(id(p)) for p in ctx.params)
        ctx.tape = None
        return (None, None, None, None) + grads
