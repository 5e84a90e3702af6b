"""Oracle restatement of diffusers==0.33.1 `UNet2DModel` (TEST INFRASTRUCTURE).

Follows SURVEY.md Appendix A.  The reference constructs the model at
/root/reference/generator_model/PolypGeneratorModel.py:25-48 and calls it at
/root/reference/generator_model/train_from_scratch.py:100; the arithmetic lives
in the un-vendored diffusers package (requirements.txt:35):
  models/unets/unet_2d.py, unet_2d_blocks.py, models/resnet.py,
  models/downsampling.py, upsampling.py, attention_processor.py, embeddings.py.
Module / parameter names reproduce the diffusers state-dict key grammar
(Appendix A.4) so that state dicts are interchangeable with the product model.
"""
from __future__ import annotations

import math
from dataclasses import dataclass
from types import SimpleNamespace
from typing import Optional, Sequence, Tuple, Union

import torch
import torch.nn as nn
import torch.nn.functional as F

from . import rounding as rq


def _conv(m: nn.Conv2d, x: torch.Tensor) -> torch.Tensor:
    """m(x); in the rounding-matched mode (oracle/rounding.py) with the bf16 operand copy of the weight."""
    if not rq.enabled():
        return m(x)
    return F.conv2d(x, rq.weight(m.weight), m.bias, m.stride, m.padding)


def _linear(m, x: torch.Tensor) -> torch.Tensor:
    if not rq.enabled() or not isinstance(m, nn.Linear):
        return m(x)
    return F.linear(x, rq.weight(m.weight), m.bias)


@dataclass
class UNet2DOutput:
    sample: torch.Tensor


def get_timestep_embedding(timesteps: torch.Tensor, embedding_dim: int, flip_sin_to_cos: bool,
                           downscale_freq_shift: float, scale: float = 1.0,
                           max_period: int = 10000) -> torch.Tensor:
    """diffusers models/embeddings.py::get_timestep_embedding (Appendix A.1 step 2)."""
    assert timesteps.dim() == 1
    half_dim = embedding_dim // 2
    exponent = -math.log(max_period) * torch.arange(0, half_dim, dtype=torch.float32, device=timesteps.device)
    exponent = exponent / (half_dim - downscale_freq_shift)
    emb = torch.exp(exponent)
    emb = timesteps[:, None].float() * emb[None, :]
    emb = scale * emb
    emb = torch.cat([torch.sin(emb), torch.cos(emb)], dim=-1)
    if flip_sin_to_cos:
        emb = torch.cat([emb[:, half_dim:], emb[:, :half_dim]], dim=-1)
    if embedding_dim % 2 == 1:
        emb = F.pad(emb, (0, 1, 0, 0))
    return emb


class Timesteps(nn.Module):
    def __init__(self, num_channels: int, flip_sin_to_cos: bool, downscale_freq_shift: float):
        super().__init__()
        self.num_channels = num_channels
        self.flip_sin_to_cos = flip_sin_to_cos
        self.downscale_freq_shift = downscale_freq_shift

    def forward(self, timesteps):
        return get_timestep_embedding(timesteps, self.num_channels, self.flip_sin_to_cos,
                                      self.downscale_freq_shift)


class TimestepEmbedding(nn.Module):
    def __init__(self, in_channels: int, time_embed_dim: int):
        super().__init__()
        self.linear_1 = nn.Linear(in_channels, time_embed_dim)
        self.act = nn.SiLU()
        self.linear_2 = nn.Linear(time_embed_dim, time_embed_dim)

    def forward(self, sample):
        return self.linear_2(self.act(self.linear_1(sample)))


class ResnetBlock2D(nn.Module):
    """Appendix A.2."""

    def __init__(self, in_channels: int, out_channels: int, temb_channels: int = 512, groups: int = 32,
                 eps: float = 1e-5, output_scale_factor: float = 1.0, dropout: float = 0.0):
        super().__init__()
        self.in_channels = in_channels
        self.out_channels = out_channels
        self.output_scale_factor = output_scale_factor
        self.norm1 = nn.GroupNorm(num_groups=groups, num_channels=in_channels, eps=eps, affine=True)
        self.conv1 = nn.Conv2d(in_channels, out_channels, kernel_size=3, stride=1, padding=1)
        self.time_emb_proj = nn.Linear(temb_channels, out_channels)
        self.norm2 = nn.GroupNorm(num_groups=groups, num_channels=out_channels, eps=eps, affine=True)
        self.dropout = nn.Dropout(dropout)
        self.conv2 = nn.Conv2d(out_channels, out_channels, kernel_size=3, stride=1, padding=1)
        self.nonlinearity = nn.SiLU()
        self.conv_shortcut = None
        if in_channels != out_channels:
            self.conv_shortcut = nn.Conv2d(in_channels, out_channels, kernel_size=1, stride=1, padding=0, bias=True)

    def forward(self, input_tensor, temb):
        hidden_states = _conv(self.conv1, rq.act(self.nonlinearity(self.norm1(input_tensor))))
        temb = self.time_emb_proj(self.nonlinearity(temb))[:, :, None, None]
        hidden_states = rq.act(hidden_states + temb)
        hidden_states = _conv(self.conv2, rq.act(self.dropout(self.nonlinearity(self.norm2(hidden_states)))))
        if self.conv_shortcut is not None:
            input_tensor = rq.act(_conv(self.conv_shortcut, input_tensor))
        return rq.act((input_tensor + hidden_states) / self.output_scale_factor)


class Attention(nn.Module):
    """Appendix A.3 (deprecated-attn-block configuration + AttnProcessor2_0)."""

    def __init__(self, query_dim: int, heads: int, dim_head: int, eps: float = 1e-5, norm_num_groups: int = 32,
                 rescale_output_factor: float = 1.0):
        super().__init__()
        assert heads * dim_head == query_dim
        self.heads = heads
        self.dim_head = dim_head
        self.scale = dim_head ** -0.5
        self.rescale_output_factor = rescale_output_factor
        self.group_norm = nn.GroupNorm(num_channels=query_dim, num_groups=norm_num_groups, eps=eps, affine=True)
        self.to_q = nn.Linear(query_dim, query_dim, bias=True)
        self.to_k = nn.Linear(query_dim, query_dim, bias=True)
        self.to_v = nn.Linear(query_dim, query_dim, bias=True)
        self.to_out = nn.ModuleList([nn.Linear(query_dim, query_dim, bias=True), nn.Dropout(0.0)])

    def forward(self, hidden_states):
        residual = hidden_states
        b, c, h, w = hidden_states.shape
        x = hidden_states.view(b, c, h * w).transpose(1, 2)
        x = rq.act(self.group_norm(x.transpose(1, 2)).transpose(1, 2))
        q, k, v = rq.act(_linear(self.to_q, x)), rq.act(_linear(self.to_k, x)), rq.act(_linear(self.to_v, x))

        def split(t):
            return t.view(b, -1, self.heads, self.dim_head).transpose(1, 2)

        o = F.scaled_dot_product_attention(split(q), split(k), split(v), attn_mask=None, dropout_p=0.0,
                                           is_causal=False, scale=self.scale)
        o = rq.act(o.transpose(1, 2).reshape(b, -1, c).to(q.dtype))
        o = self.to_out[1](_linear(self.to_out[0], o))
        o = o.transpose(-1, -2).reshape(b, c, h, w)
        o = o + residual
        return rq.act(o / self.rescale_output_factor)


class Downsample2D(nn.Module):
    def __init__(self, channels: int, padding: int = 1):
        super().__init__()
        self.padding = padding
        self.conv = nn.Conv2d(channels, channels, kernel_size=3, stride=2, padding=padding)

    def forward(self, x):
        if self.padding == 0:
            x = F.pad(x, (0, 1, 0, 1), mode="constant", value=0)
        return rq.act(_conv(self.conv, x))


class Upsample2D(nn.Module):
    def __init__(self, channels: int):
        super().__init__()
        self.conv = nn.Conv2d(channels, channels, kernel_size=3, padding=1)

    def forward(self, x):
        if x.shape[0] >= 64:
            x = x.contiguous()
        x = F.interpolate(x, scale_factor=2.0, mode="nearest")
        return rq.act(_conv(self.conv, x))


class DownBlock2D(nn.Module):
    has_attention = False

    def __init__(self, num_layers, in_channels, out_channels, temb_channels, add_downsample, eps, groups,
                 downsample_padding, attention_head_dim=None):
        super().__init__()
        self.resnets = nn.ModuleList([
            ResnetBlock2D(in_channels if i == 0 else out_channels, out_channels, temb_channels, groups, eps)
            for i in range(num_layers)])
        if self.has_attention:
            d = attention_head_dim if attention_head_dim is not None else out_channels
            self.attentions = nn.ModuleList([
                Attention(out_channels, out_channels // d, d, eps=eps, norm_num_groups=groups)
                for _ in range(num_layers)])
        self.downsamplers = None
        if add_downsample:
            self.downsamplers = nn.ModuleList([Downsample2D(out_channels, padding=downsample_padding)])

    def forward(self, hidden_states, temb):
        output_states = ()
        for i, resnet in enumerate(self.resnets):
            hidden_states = resnet(hidden_states, temb)
            if self.has_attention:
                hidden_states = self.attentions[i](hidden_states)
            output_states += (hidden_states,)
        if self.downsamplers is not None:
            for d in self.downsamplers:
                hidden_states = d(hidden_states)
            output_states += (hidden_states,)
        return hidden_states, output_states


class AttnDownBlock2D(DownBlock2D):
    has_attention = True


class UNetMidBlock2D(nn.Module):
    def __init__(self, in_channels, temb_channels, eps, groups, attention_head_dim, add_attention=True):
        super().__init__()
        self.resnets = nn.ModuleList([ResnetBlock2D(in_channels, in_channels, temb_channels, groups, eps),
                                      ResnetBlock2D(in_channels, in_channels, temb_channels, groups, eps)])
        d = attention_head_dim if attention_head_dim is not None else in_channels
        self.attentions = nn.ModuleList(
            [Attention(in_channels, in_channels // d, d, eps=eps, norm_num_groups=groups)] if add_attention else [None])

    def forward(self, hidden_states, temb):
        hidden_states = self.resnets[0](hidden_states, temb)
        for attn, resnet in zip(self.attentions, self.resnets[1:]):
            if attn is not None:
                hidden_states = attn(hidden_states)
            hidden_states = resnet(hidden_states, temb)
        return hidden_states


class UpBlock2D(nn.Module):
    has_attention = False

    def __init__(self, num_layers, in_channels, prev_output_channel, out_channels, temb_channels, add_upsample,
                 eps, groups, attention_head_dim=None):
        super().__init__()
        resnets = []
        for i in range(num_layers):
            res_skip_channels = in_channels if (i == num_layers - 1) else out_channels
            resnet_in_channels = prev_output_channel if i == 0 else out_channels
            resnets.append(ResnetBlock2D(resnet_in_channels + res_skip_channels, out_channels, temb_channels,
                                         groups, eps))
        self.resnets = nn.ModuleList(resnets)
        if self.has_attention:
            d = attention_head_dim if attention_head_dim is not None else out_channels
            self.attentions = nn.ModuleList([
                Attention(out_channels, out_channels // d, d, eps=eps, norm_num_groups=groups)
                for _ in range(num_layers)])
        self.upsamplers = None
        if add_upsample:
            self.upsamplers = nn.ModuleList([Upsample2D(out_channels)])

    def forward(self, hidden_states, res_hidden_states_tuple, temb):
        for i, resnet in enumerate(self.resnets):
            res_hidden_states = res_hidden_states_tuple[-1]
            res_hidden_states_tuple = res_hidden_states_tuple[:-1]
            hidden_states = torch.cat([hidden_states, res_hidden_states], dim=1)
            hidden_states = resnet(hidden_states, temb)
            if self.has_attention:
                hidden_states = self.attentions[i](hidden_states)
        if self.upsamplers is not None:
            for u in self.upsamplers:
                hidden_states = u(hidden_states)
        return hidden_states


class AttnUpBlock2D(UpBlock2D):
    has_attention = True


_DOWN = {"DownBlock2D": DownBlock2D, "AttnDownBlock2D": AttnDownBlock2D}
_UP = {"UpBlock2D": UpBlock2D, "AttnUpBlock2D": AttnUpBlock2D}


def polyp_unet_config(sample_size: int = 128) -> dict:
    """Keyword arguments exactly as PolypGeneratorModel.py:26-47 passes them."""
    return dict(
        sample_size=sample_size, in_channels=3, out_channels=3, layers_per_block=2,
        block_out_channels=(128, 128, 256, 256, 512, 512),
        down_block_types=("DownBlock2D", "DownBlock2D", "DownBlock2D", "DownBlock2D", "AttnDownBlock2D",
                          "DownBlock2D"),
        up_block_types=("UpBlock2D", "AttnUpBlock2D", "UpBlock2D", "UpBlock2D", "UpBlock2D", "UpBlock2D"),
    )


def celebahq_unet_config(sample_size: int = 256) -> dict:
    """google/ddpm-celebahq-256 architecture knobs."""
    cfg = polyp_unet_config(sample_size)
    cfg.update(attention_head_dim=None, downsample_padding=0, flip_sin_to_cos=False, freq_shift=1, norm_eps=1e-6)
    return cfg


class UNet2DModel(nn.Module):
    """Appendix A.1; defaults are the diffusers 0.33.1 defaults."""

    def __init__(self, sample_size: Optional[Union[int, Tuple[int, int]]] = None, in_channels: int = 3,
                 out_channels: int = 3, center_input_sample: bool = False, time_embedding_type: str = "positional",
                 freq_shift: int = 0, flip_sin_to_cos: bool = True,
                 down_block_types: Sequence[str] = ("DownBlock2D", "AttnDownBlock2D", "AttnDownBlock2D",
                                                    "AttnDownBlock2D"),
                 up_block_types: Sequence[str] = ("AttnUpBlock2D", "AttnUpBlock2D", "AttnUpBlock2D", "UpBlock2D"),
                 block_out_channels: Sequence[int] = (224, 448, 672, 896), layers_per_block: int = 2,
                 mid_block_scale_factor: float = 1, downsample_padding: int = 1, act_fn: str = "silu",
                 attention_head_dim: Optional[int] = 8, norm_num_groups: int = 32, norm_eps: float = 1e-5,
                 add_attention: bool = True):
        super().__init__()
        if time_embedding_type != "positional" or act_fn != "silu" or center_input_sample:
            raise ValueError("oracle supports the positional/silu/uncentred configuration only")
        if len(down_block_types) != len(up_block_types) or len(block_out_channels) != len(down_block_types):
            raise ValueError("down_block_types, up_block_types and block_out_channels must have equal length")
        self.config = SimpleNamespace(
            sample_size=sample_size, in_channels=in_channels, out_channels=out_channels,
            freq_shift=freq_shift, flip_sin_to_cos=flip_sin_to_cos, down_block_types=tuple(down_block_types),
            up_block_types=tuple(up_block_types), block_out_channels=tuple(block_out_channels),
            layers_per_block=layers_per_block, downsample_padding=downsample_padding,
            attention_head_dim=attention_head_dim, norm_num_groups=norm_num_groups, norm_eps=norm_eps,
            add_attention=add_attention, time_embedding_type=time_embedding_type, act_fn=act_fn)
        time_embed_dim = block_out_channels[0] * 4
        self.conv_in = nn.Conv2d(in_channels, block_out_channels[0], kernel_size=3, padding=1)
        self.time_proj = Timesteps(block_out_channels[0], flip_sin_to_cos, freq_shift)
        self.time_embedding = TimestepEmbedding(block_out_channels[0], time_embed_dim)

        self.down_blocks = nn.ModuleList()
        output_channel = block_out_channels[0]
        for i, t in enumerate(down_block_types):
            input_channel = output_channel
            output_channel = block_out_channels[i]
            is_final = i == len(block_out_channels) - 1
            self.down_blocks.append(_DOWN[t](layers_per_block, input_channel, output_channel, time_embed_dim,
                                             not is_final, norm_eps, norm_num_groups, downsample_padding,
                                             attention_head_dim))
        self.mid_block = UNetMidBlock2D(block_out_channels[-1], time_embed_dim, norm_eps, norm_num_groups,
                                        attention_head_dim, add_attention)
        self.up_blocks = nn.ModuleList()
        rev = list(reversed(block_out_channels))
        output_channel = rev[0]
        for i, t in enumerate(up_block_types):
            prev_output_channel = output_channel
            output_channel = rev[i]
            input_channel = rev[min(i + 1, len(block_out_channels) - 1)]
            is_final = i == len(block_out_channels) - 1
            self.up_blocks.append(_UP[t](layers_per_block + 1, input_channel, prev_output_channel, output_channel,
                                         time_embed_dim, not is_final, norm_eps, norm_num_groups,
                                         attention_head_dim))
        self.conv_norm_out = nn.GroupNorm(num_channels=block_out_channels[0], num_groups=norm_num_groups,
                                          eps=norm_eps)
        self.conv_act = nn.SiLU()
        self.conv_out = nn.Conv2d(block_out_channels[0], out_channels, kernel_size=3, padding=1)

    @property
    def dtype(self):
        return next(self.parameters()).dtype

    @property
    def device(self):
        return next(self.parameters()).device

    def forward(self, sample: torch.Tensor, timestep, class_labels=None, return_dict: bool = True):
        timesteps = timestep
        if not torch.is_tensor(timesteps):
            timesteps = torch.tensor([timesteps], dtype=torch.long, device=sample.device)
        elif timesteps.dim() == 0:
            timesteps = timesteps[None].to(sample.device)
        timesteps = timesteps * torch.ones(sample.shape[0], dtype=timesteps.dtype, device=timesteps.device)
        t_emb = self.time_proj(timesteps).to(dtype=self.dtype)
        emb = self.time_embedding(t_emb)

        sample = rq.act(_conv(self.conv_in, rq.act(sample)))
        res = (sample,)
        for blk in self.down_blocks:
            sample, r = blk(sample, emb)
            res += r
        sample = self.mid_block(sample, emb)
        for blk in self.up_blocks:
            n = len(blk.resnets)
            r, res = res[-n:], res[:-n]
            sample = blk(sample, r, emb)
        sample = _conv(self.conv_out, rq.act(self.conv_act(self.conv_norm_out(sample))))
        if not return_dict:
            return (sample,)
        return UNet2DOutput(sample=sample)
