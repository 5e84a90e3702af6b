"""CUDA-graph replay of the two hot loops.

One training step of the 113.7 M-parameter UNet is ~1000 kernel launches of 10-500 us; issuing them from Python takes
~44 ms of host time per step (bench.py `host_issue_ms_per_step`), i.e. as much as the GPU needs to execute them, and a
reverse-diffusion step at batch 32 (~330 launches, 7 ms of GPU work) is host-bound outright.  Both loops have static
shapes, so they are captured ONCE into a CUDA graph and replayed:

  * GraphedTrainStep -- add_noise -> UNet forward -> MSE -> backward (-> gradient all-reduce) -> clip -> AdamW,
    the body of /root/reference/generator_model/train_from_scratch.py:83-113, replayed per step with the batch, the
    noise and the timesteps copied into static device buffers;
  * GraphedUNetForward -- eps = unet(x, t) for sampling (DDPMPipeline uses it automatically on CUDA).

All kernels are the same C-ABI launches as in eager mode (they go to torch's current stream, which is the capturing
stream); memory comes from the graph's private pool, so the TMA descriptors cached by (pointer, shape) stay valid.
"""
from __future__ import annotations

from typing import Optional

import torch


class GraphedUNetForward:
    """eps = unet(x, t) with a scalar timestep, replayed from a CUDA graph (inference only)."""

    def __init__(self, unet, batch: int, height: int, width: int):
        dev = unet.device
        if dev.type != "cuda":
            raise RuntimeError("CUDA graphs need the model on a CUDA device")
        self.unet = unet
        self.x = torch.zeros((batch, unet.config.in_channels, height, width), device=dev, dtype=torch.float32)
        self.t = torch.zeros((batch,), device=dev, dtype=torch.int64)
        self._key = None
        self.graph = None
        self.out = None

    def _weights_key(self):
        return (getattr(self.unet, "_weights_epoch", 0),) + tuple(p._version for p in self.unet.parameters())

    def _capture(self):
        unet = self.unet
        side = torch.cuda.Stream(device=self.x.device)
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side), torch.no_grad():
            for _ in range(2):                      # warm-up: arena, operand cache, smem attributes, descriptors
                unet(self.x, self.t)
        torch.cuda.current_stream().wait_stream(side)
        torch.cuda.synchronize()
        self.graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(self.graph), torch.no_grad():
            self.out = unet(self.x, self.t, return_dict=False)[0]
        self._key = self._weights_key()

    def __call__(self, x: torch.Tensor, t) -> torch.Tensor:
        """Returns a STATIC output buffer that the next call overwrites."""
        if self.graph is None:
            self._capture()
        elif self._key != self._weights_key():
            # weights changed: one eager forward refreshes the bf16 operand copies IN PLACE (same addresses), so the
            # captured graph, which only reads them, stays valid
            with torch.no_grad():
                self.unet(self.x, self.t)
            self._key = self._weights_key()
        if x.data_ptr() != self.x.data_ptr():
            self.x.copy_(x)
        self.t.fill_(int(t))
        self.graph.replay()
        return self.out


class GraphedTrainStep:
    """One optimisation step as a single CUDA-graph replay.

        step = GraphedTrainStep(net, noise_scheduler, optimizer, batch_shape)   # net: UNet2DModel or its DDP wrapper
        loss = step(clean, noise, timesteps)          # device scalar (static buffer); no host sync

    Construction runs `warmup_iters` eager steps plus the captured one on whatever the static buffers hold (pass
    `warmup_batch=(clean, noise, timesteps)` to use real data instead of zeros): these ARE optimisation steps, exactly
    like the first iterations of the reference loop -- count them in the schedule or construct the object before
    loading a checkpoint.

    Do not keep an autograd graph of the same model alive across construction (e.g. the `loss` of an earlier eager step
    that was never freed): its AccumulateGrad nodes are bound to the stream of that step, and CUDA refuses a capture
    that makes the legacy default stream wait on the capturing one (cudaErrorStreamCaptureImplicit).

    The optimizer must be capturable: optim.FusedAdamW (set its lr_tensor to a device scalar to drive a schedule from
    the host: lr_tensor.fill_(value) between replays) or torch.optim.AdamW(..., fused=True, capturable=True).
    """

    def __init__(self, net, noise_scheduler, optimizer, batch_shape, max_grad_norm: Optional[float] = 1.0,
                 warmup_iters: int = 3, lr_scheduler=None, warmup_batch=None):
        from .training import mse_loss
        params = [p for p in net.parameters() if p.requires_grad]
        dev = params[0].device
        if dev.type != "cuda":
            raise RuntimeError("CUDA graphs need the model on a CUDA device")
        # The learning rate of a captured optimizer kernel is a launch constant unless it is read from device memory:
        # with lr_scheduler (train_from_scratch.py:113, get_cosine_schedule_with_warmup) the rate lives in a device
        # scalar that __call__ refreshes from the param group after every lr_scheduler.step().
        self.lr_scheduler, self._lr_t, self._opt = lr_scheduler, None, optimizer
        if lr_scheduler is not None:
            if not hasattr(optimizer, "lr_tensor"):
                raise NotImplementedError("lr_scheduler with a graphed step needs optim.FusedAdamW (device-resident lr)")
            self._lr_t = torch.full((1,), float(optimizer.param_groups[0]["lr"]), device=dev, dtype=torch.float32)
            optimizer.lr_tensor = self._lr_t
        self.clean = torch.zeros(tuple(batch_shape), device=dev, dtype=torch.float32)
        self.noise = torch.zeros(tuple(batch_shape), device=dev, dtype=torch.float32)
        self.t = torch.zeros((batch_shape[0],), device=dev, dtype=torch.int64)
        if warmup_batch is not None:
            self.clean.copy_(warmup_batch[0])
            self.noise.copy_(warmup_batch[1])
            self.t.copy_(warmup_batch[2])

        def body():
            noisy = noise_scheduler.add_noise(self.clean, self.noise, self.t)
            pred = net(noisy, self.t, return_dict=False)[0]
            loss = mse_loss(pred, self.noise)
            loss.backward()
            if max_grad_norm is not None and getattr(optimizer, "max_grad_norm", None) is None:
                torch.nn.utils.clip_grad_norm_(params, max_grad_norm)   # optim.FusedAdamW clips inside its update
            optimizer.step()
            return loss.detach()

        side = torch.cuda.Stream(device=dev)
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):
            for _ in range(warmup_iters):
                optimizer.zero_grad(set_to_none=True)
                body()
        torch.cuda.current_stream().wait_stream(side)
        torch.cuda.synchronize()
        self.graph = torch.cuda.CUDAGraph()
        optimizer.zero_grad(set_to_none=True)
        with torch.cuda.graph(self.graph):
            self.loss = body()
        self.warmup_iters = warmup_iters
        unet = getattr(net, "module", net)
        self._invalidate = getattr(unet, "invalidate_weight_cache", None)

    def __call__(self, clean: torch.Tensor, noise: torch.Tensor, timesteps: torch.Tensor) -> torch.Tensor:
        self.clean.copy_(clean, non_blocking=True)
        self.noise.copy_(noise, non_blocking=True)
        self.t.copy_(timesteps, non_blocking=True)
        self.graph.replay()
        if self._invalidate is not None:        # replays update the weights without bumping autograd versions
            self._invalidate()
        if self.lr_scheduler is not None:       # the reference steps the schedule once per iteration (:113)
            self.lr_scheduler.step()
            self._lr_t.fill_(float(self._opt.param_groups[0]["lr"]))
        return self.loss
