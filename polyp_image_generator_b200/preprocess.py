"""Device-side input pipeline.

Replaces the per-image CPU transform of /root/reference/generator_model/PolypDiffusionDataset.py:52-59

    transforms.Compose([Resize((S, S)), RandomHorizontalFlip(), ToTensor(), Normalize([0.5], [0.5])])

and the `num_workers=0` loader of train_from_scratch.py:236.  At >1 500 img/s per GPU a PIL pipeline on one core is the
bottleneck; here the loader only has to deliver raw uint8 frames, staged through pinned memory on a side stream, and the
whole transform runs as two kernels per batch (csrc/preprocess.cu) -- bit-identical to Pillow / torchvision, including the
8-bit intermediate between the two resampling passes.

    pre = DevicePreprocessor(image_size=128)
    flips = pre.draw_flips(len(frames))            # same torch.rand(1) draws, in order, as RandomHorizontalFlip
    x = pre(frames_u8.to("cuda"), flips)           # uint8 [B, H, W, 3] -> fp32 [B, 3, 128, 128] in [-1, 1]
"""
from __future__ import annotations

import math
from typing import Dict, Iterable, Iterator, Optional, Tuple

import numpy as np
import torch

from . import ops as _ops

PRECISION_BITS = 32 - 8 - 2      # Pillow Resample.c


def pillow_bilinear_tables(in_size: int, out_size: int) -> Tuple[np.ndarray, np.ndarray, int]:
    """Integer coefficient tables of Pillow's antialiased BILINEAR resampler for in_size -> out_size (precompute_coeffs +
    normalize_coeffs_8bpc).  An identity table (one tap of weight 1) when the sizes agree: Pillow skips that pass."""
    if in_size == out_size:
        bounds = np.stack([np.arange(out_size, dtype=np.int32), np.ones(out_size, dtype=np.int32)], axis=1)
        return np.ascontiguousarray(bounds), np.full((out_size, 1), 1 << PRECISION_BITS, dtype=np.int32), 1
    scale = in_size / out_size
    filterscale = max(scale, 1.0)
    support = filterscale                     # triangle filter: support 1.0
    ksize = int(math.ceil(support)) * 2 + 1
    inv = 1.0 / filterscale
    bounds = np.zeros((out_size, 2), dtype=np.int32)
    coeffs = np.zeros((out_size, ksize), dtype=np.int32)
    for xx in range(out_size):
        center = (xx + 0.5) * scale
        lo = max(int(center - support + 0.5), 0)
        hi = min(int(center + support + 0.5), in_size)
        n = hi - lo
        w = [max(0.0, 1.0 - abs((x + lo - center + 0.5) * inv)) for x in range(n)]
        total = 0.0
        for v in w:
            total += v
        for x in range(n):
            v = w[x] / total if total != 0.0 else w[x]
            coeffs[xx, x] = int(0.5 + v * (1 << PRECISION_BITS))
        bounds[xx] = (lo, n)
    return bounds, coeffs, ksize


class DevicePreprocessor:
    """Resize -> RandomHorizontalFlip -> ToTensor -> Normalize([0.5], [0.5]) for a batch of equally sized uint8 frames."""

    def __init__(self, image_size, flip_p: float = 0.5):
        self.size = (image_size, image_size) if isinstance(image_size, int) else tuple(image_size)   # (H, W)
        self.flip_p = float(flip_p)
        self._tables: Dict[tuple, tuple] = {}

    def draw_flips(self, n: int, generator: Optional[torch.Generator] = None) -> torch.Tensor:
        """One `torch.rand(1) < p` per image, in order -- what RandomHorizontalFlip.forward draws per __getitem__ (global
        CPU RNG unless a generator is given), so a seeded run flips the same images as the reference loader."""
        return torch.tensor([bool(torch.rand(1, generator=generator) < self.flip_p) for _ in range(n)],
                            dtype=torch.bool)

    def _device_tables(self, in_size: int, out_size: int, device) -> tuple:
        key = (in_size, out_size, str(device))
        t = self._tables.get(key)
        if t is None:
            b, c, k = pillow_bilinear_tables(in_size, out_size)
            t = (torch.from_numpy(b).to(device), torch.from_numpy(c).to(device), k)
            self._tables[key] = t
        return t

    def __call__(self, frames: torch.Tensor, flips: Optional[torch.Tensor] = None) -> torch.Tensor:
        if frames.dtype != torch.uint8 or frames.dim() != 4 or frames.shape[-1] not in (1, 3):
            raise TypeError("frames must be uint8 [B, H, W, C] with C = 1 or 3 (PIL 'L' / 'RGB' memory layout)")
        ops = _ops.get()
        if ops.name == "cuda" and not frames.is_cuda:
            raise RuntimeError("DevicePreprocessor runs on CUDA tensors: stage frames with PinnedPrefetcher or .to('cuda')")
        b, h, w, c = frames.shape
        oh, ow = self.size
        if flips is not None:
            if flips.numel() != b:
                raise ValueError("flips must hold one entry per frame")
            flips = flips.to(device=frames.device, dtype=torch.uint8).contiguous()
        if b == 0:
            return torch.empty((0, c, oh, ow), device=frames.device, dtype=torch.float32)
        return ops.preprocess_u8(frames.contiguous(), oh, ow, self._device_tables(w, ow, frames.device),
                                 self._device_tables(h, oh, frames.device), flips)


class PinnedPrefetcher:
    """Iterates (frames_u8_cuda, extras) batches one ahead of the consumer: the H2D copy of batch i+1 (from pinned host
    memory, on a side stream) overlaps the training step of batch i.  `batches` yields uint8 [B, H, W, C] host tensors
    (optionally a tuple whose first element is that tensor)."""

    def __init__(self, batches: Iterable, device):
        self.batches, self.device = batches, torch.device(device)
        self.stream = torch.cuda.Stream(device=self.device)
        self._slots = [None, None]       # two pinned staging buffers: [tensor, event of the last H2D copy out of it]
        self._turn = 0

    def _stage(self, item):
        frames, rest = (item[0], tuple(item[1:])) if isinstance(item, (tuple, list)) else (item, ())
        k = self._turn
        self._turn ^= 1
        slot = self._slots[k]
        if slot is None or slot[0].shape != frames.shape:
            slot = self._slots[k] = [torch.empty(frames.shape, dtype=torch.uint8, pin_memory=True), None]
        if slot[1] is not None:
            slot[1].synchronize()        # the previous copy out of this buffer must have left before it is overwritten
        slot[0].copy_(frames)
        with torch.cuda.stream(self.stream):
            dev = slot[0].to(self.device, non_blocking=True)
        ev = torch.cuda.Event()
        ev.record(self.stream)
        slot[1] = ev
        return dev, rest, ev

    def __iter__(self) -> Iterator:
        it = iter(self.batches)
        nxt = None
        try:
            nxt = self._stage(next(it))
        except StopIteration:
            return
        while nxt is not None:
            dev, rest, ev = nxt
            try:
                nxt = self._stage(next(it))
            except StopIteration:
                nxt = None
            torch.cuda.current_stream(self.device).wait_event(ev)
            dev.record_stream(torch.cuda.current_stream(self.device))
            yield (dev, *rest) if rest else dev
