"""Oracle restatement of the reference input transform (TEST INFRASTRUCTURE).

/root/reference/generator_model/PolypDiffusionDataset.py:52-59 builds

    transforms.Compose([Resize((S, S)), RandomHorizontalFlip(), ToTensor(), Normalize([0.5], [0.5])])

on PIL RGB images.  torchvision's Resize on a PIL image is `img.resize((S, S), Image.BILINEAR)`, i.e. Pillow's
two-pass (horizontal, then vertical) antialiased triangle-filter resampler in 8-bit fixed point (Pillow
src/libImaging/Resample.c: precompute_coeffs, normalize_coeffs_8bpc, ImagingResampleHorizontal_8bpc /
ImagingResampleVertical_8bpc; PRECISION_BITS = 32 - 8 - 2).  Pillow is a third-party dependency of the reference
(requirements.txt) and IS installed here, so this restatement is pinned against Pillow itself in
tests/test_preprocess.py -- the one part of the path whose parity is not "unpinned".
"""
from __future__ import annotations

import math
from typing import Tuple

import numpy as np
import torch

PRECISION_BITS = 32 - 8 - 2


def precompute_coeffs(in_size: int, out_size: int) -> Tuple[np.ndarray, np.ndarray, int]:
    """Pillow precompute_coeffs + normalize_coeffs_8bpc for the bilinear (triangle, support 1) filter over the full
    source extent.  Returns (bounds int32 [out, 2] = (xmin, count), coeffs int32 [out, ksize], ksize)."""
    scale = filterscale = in_size / out_size
    if filterscale < 1.0:
        filterscale = 1.0
    support = 1.0 * filterscale
    ksize = int(math.ceil(support)) * 2 + 1
    bounds = np.zeros((out_size, 2), dtype=np.int32)
    kk = np.zeros((out_size, ksize), dtype=np.int32)
    ss = 1.0 / filterscale
    for xx in range(out_size):
        center = (xx + 0.5) * scale
        xmin = int(center - support + 0.5)
        if xmin < 0:
            xmin = 0
        xmax = int(center + support + 0.5)
        if xmax > in_size:
            xmax = in_size
        xmax -= xmin
        w = np.zeros(ksize, dtype=np.float64)
        ww = 0.0
        for x in range(xmax):
            a = (x + xmin - center + 0.5) * ss
            if a < 0.0:
                a = -a
            w[x] = 1.0 - a if a < 1.0 else 0.0
            ww += w[x]
        if ww != 0.0:
            for x in range(xmax):
                w[x] /= ww
        for x in range(ksize):
            v = w[x] * (1 << PRECISION_BITS)
            kk[xx, x] = int(-0.5 + v) if w[x] < 0 else int(0.5 + v)
        bounds[xx] = (xmin, xmax)
    return bounds, kk, ksize


def _clip8(acc: np.ndarray) -> np.ndarray:
    return np.clip(acc >> PRECISION_BITS, 0, 255).astype(np.uint8)


def resize_bilinear_u8(img: np.ndarray, out_h: int, out_w: int) -> np.ndarray:
    """img uint8 [H, W, C] -> uint8 [out_h, out_w, C], bit-identical to PIL Image.resize(..., BILINEAR)."""
    h, w, c = img.shape
    cur = img
    if out_w != w:                                   # horizontal pass (skipped when the width already matches)
        bounds, kk, _ = precompute_coeffs(w, out_w)
        out = np.empty((h, out_w, c), dtype=np.uint8)
        for xx in range(out_w):
            xmin, n = bounds[xx]
            acc = np.full((h, c), 1 << (PRECISION_BITS - 1), dtype=np.int64)
            for i in range(n):
                acc += cur[:, xmin + i, :].astype(np.int64) * int(kk[xx, i])
            out[:, xx, :] = _clip8(acc)
        cur = out
    if out_h != h:                                   # vertical pass
        bounds, kk, _ = precompute_coeffs(h, out_h)
        out = np.empty((out_h, cur.shape[1], c), dtype=np.uint8)
        for yy in range(out_h):
            ymin, n = bounds[yy]
            acc = np.full((cur.shape[1], c), 1 << (PRECISION_BITS - 1), dtype=np.int64)
            for i in range(n):
                acc += cur[ymin + i, :, :].astype(np.int64) * int(kk[yy, i])
            out[yy] = _clip8(acc)
        cur = out
    return cur


def transform(img: np.ndarray, size: int, flip: bool) -> torch.Tensor:
    """Resize((S, S)) -> hflip if `flip` -> ToTensor -> Normalize([0.5], [0.5]): fp32 [C, S, S] in [-1, 1]."""
    r = resize_bilinear_u8(img, size, size)
    if flip:
        r = r[:, ::-1, :]
    t = torch.from_numpy(np.ascontiguousarray(r)).permute(2, 0, 1).contiguous().to(torch.float32).div(255)
    return (t - 0.5) / 0.5
