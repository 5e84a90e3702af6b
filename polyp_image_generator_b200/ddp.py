"""Data-parallel training and sharded sampling for the DDPM hot path (one process per GPU, NCCL over NVLink).

The reference is single-GPU (no distributed code at all); BASELINE.json's north_star adds
"training shards by batch across the 8 GPUs of one box with NCCL gradient allreduce, sampling shards by batch with
no communication".  Every op of the UNet is independent across samples (GroupNorm statistics are per sample), so the
only exchange step is the mean of the weight gradients.

Because the UNet keeps all gradients in ONE flat fp32 arena (unet.py), the exchange is a handful of large
all-reduces on a side stream, issued as soon as the backward program has finished the layers a bucket covers, and
joined before the per-parameter gradient views are handed to autograd.
"""
from __future__ import annotations

from typing import Iterable, List, Optional, Tuple

import torch
import torch.distributed as dist
import torch.nn as nn


class DistributedDataParallel(nn.Module):
    """Wraps the B200 UNet2DModel; same call surface as torch's DDP for the reference's loop (`model(x, t)`,
    `.parameters()`, `.train()/.eval()`, `.module`)."""

    def __init__(self, module: nn.Module, process_group=None, bucket_cap_mb: float = 32.0,
                 broadcast_parameters: bool = True):
        super().__init__()
        if not dist.is_initialized():
            raise RuntimeError("torch.distributed must be initialised before wrapping the model")
        self.module = module
        self.pg = process_group
        self.world = dist.get_world_size(process_group)
        self.bucket_elems = int(bucket_cap_mb * 1024 * 1024 // 4)
        self._side = None
        self._pending: List[Tuple[int, int]] = []
        self._done_upto: Optional[int] = None
        self._arena_trainable: Optional[bool] = None
        module._grad_ready_hook = self._finish
        module._grad_begin_hook = self._begin
        module._grad_progress_hook = self._progress
        if broadcast_parameters:
            self.broadcast_parameters()

    # ---- parameters start identical on every rank (as torch DDP does) -----------------------------------------
    def broadcast_parameters(self, src: int = 0):
        m = self.module
        if hasattr(m, "_ensure_arena") and next(m.parameters()).device.type == "cuda":
            m._ensure_arena()
            dist.broadcast(m._arena, src, group=self.pg)
            extra = [p for _, p in m._plan.extra_params]
        else:
            extra = list(m.parameters())
        for p in extra:
            dist.broadcast(p.data, src, group=self.pg)

    def forward(self, *args, **kwargs):
        return self.module(*args, **kwargs)

    # ---- gradient exchange ---------------------------------------------------------------------------------------
    def _allreduce_mean(self, t: torch.Tensor):
        if dist.get_backend(self.pg) == "nccl":
            dist.all_reduce(t, op=dist.ReduceOp.AVG, group=self.pg)
        else:
            dist.all_reduce(t, op=dist.ReduceOp.SUM, group=self.pg)
            t.div_(self.world)

    def _stream_ctx(self, G: torch.Tensor):
        if not G.is_cuda:
            return None
        if self._side is None:
            self._side = torch.cuda.Stream(device=G.device)
        return self._side

    def _begin(self):
        """Start of a backward pass: forget bucket progress a previous (possibly aborted) pass left behind."""
        self._done_upto = None

    def _progress(self, G: torch.Tensor, final_from: int):
        """Backward finished every tensor-core weight gradient at arena offsets >= final_from: ship full buckets."""
        if self._arena_trainable is None:
            self._arena_trainable = any(p.requires_grad for p, _, _ in self.module._plan.layout)
        if not self._arena_trainable:      # LoRA runs: the arena holds no gradients, only the adapters are exchanged
            return
        hi = self._done_upto if self._done_upto is not None else self.module._plan.temb_w_off
        if hi - final_from < self.bucket_elems:
            return
        side = self._stream_ctx(G)
        seg = G[final_from:hi]
        if side is not None:
            side.wait_stream(torch.cuda.current_stream())
            wg = getattr(self.module, "_wgrad_stream", None)
            if wg is not None:            # the weight gradients themselves are produced on the model's wgrad stream
                side.wait_stream(wg)
            with torch.cuda.stream(side):
                self._allreduce_mean(seg)
            seg.record_stream(side)
        else:
            self._allreduce_mean(seg)
        self._done_upto = final_from

    def _finish(self, G: torch.Tensor, extra_grads: Iterable[torch.Tensor] = ()):
        """End of backward: reduce what is left (early layers, biases, norms, time MLP, LoRA) and join."""
        m = self.module
        arena_trainable = any(p.requires_grad for p, _, _ in m._plan.layout)
        self._arena_trainable = None       # requires_grad flags may change between steps (unfreeze_layers)
        side = self._stream_ctx(G)
        if arena_trainable:
            hi = self._done_upto if self._done_upto is not None else m._plan.temb_w_off
            if side is not None:
                side.wait_stream(torch.cuda.current_stream())
                with torch.cuda.stream(side):
                    if hi > 0:
                        self._allreduce_mean(G[:hi])
                    self._allreduce_mean(G[m._plan.temb_w_off:])
                torch.cuda.current_stream().wait_stream(side)
            else:
                if hi > 0:
                    self._allreduce_mean(G[:hi])
                self._allreduce_mean(G[m._plan.temb_w_off:])
        extra = [g for g in extra_grads if g is not None]
        if extra:
            flat = torch.cat([g.reshape(-1) for g in extra])
            self._allreduce_mean(flat)
            o = 0
            for g in extra:
                g.copy_(flat[o:o + g.numel()].view_as(g))
                o += g.numel()
        self._done_upto = None


def shard_sampling_batches(num_images: int, batch_size: int, rank: int, world: int) -> List[Tuple[int, int, int]]:
    """Sampling shards with no communication: the 1-GPU run of train_from_scratch.py::evaluate
    draws batch b with generator seed `config.seed + b` (:47-54).  Rank r takes batches r, r+world, ... so the union
    over ranks is exactly the 1-GPU image set.  Returns [(batch_id, first_image_index (0-based), count)]."""
    out = []
    start, b = 0, 0
    while start < num_images:
        cnt = min(batch_size, num_images - start)
        if b % world == rank:
            out.append((b, start, cnt))
        start += cnt
        b += 1
    return out
