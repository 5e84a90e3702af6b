"""Training-step pieces of the DDPM hot path that sit next to the UNet: the fused MSE loss and the step recipe.

Call sites replaced (paths relative to /root/reference/):
  generator_model/train_from_scratch.py:101   loss = F.mse_loss(noise_pred, noise)
  generator_model/train_from_scratch.py:103   scaler.scale(loss).backward()
  generator_model/train_from_scratch.py:83-116 (train_step below restates the loop body over the drop-in objects)
"""
from __future__ import annotations

from typing import Optional

import torch

from . import ops as _ops


class _MSELoss(torch.autograd.Function):
    """mean((pred - target)^2) with the gradient produced in the same pass (ddpm_mse_fwd_bwd)."""

    @staticmethod
    def forward(ctx, pred, target):
        ops = _ops.get()
        need = pred.requires_grad
        loss_sum, dpred = ops.mse_fwd_bwd(pred.contiguous(), target.contiguous(), want_grad=need)
        ctx.dpred = dpred
        ctx.used = False
        return (loss_sum / pred.numel()).reshape(())

    @staticmethod
    def backward(ctx, g):
        if ctx.used:
            raise RuntimeError("mse_loss backward ran twice; the fused kernel keeps a single gradient buffer")
        ctx.used = True
        dpred = ctx.dpred
        ctx.dpred = None
        if dpred is None:
            return None, None
        # grad_output is a device scalar (1.0, or GradScaler's scale): fold it in without a host sync
        _ops.get().scale_by_device_scalar(dpred, g.to(torch.float32).reshape(1).contiguous())
        return dpred, None


def mse_loss(pred: torch.Tensor, target: torch.Tensor, reduction: str = "mean") -> torch.Tensor:
    """Drop-in for F.mse_loss(pred, target) on the DDPM path (fp32, reduction='mean')."""
    if reduction != "mean":
        raise NotImplementedError("only reduction='mean' is on the DDPM hot path")
    if pred.shape != target.shape:
        raise ValueError("mse_loss: pred and target must have the same shape")
    if pred.dtype != torch.float32 or target.dtype != torch.float32:
        pred, target = pred.float(), target.float()
    return _MSELoss.apply(pred, target.detach())


def train_step(model, noise_scheduler, optimizer, clean_images: torch.Tensor, noise: Optional[torch.Tensor] = None,
               timesteps: Optional[torch.Tensor] = None, lr_scheduler=None, max_grad_norm: Optional[float] = 1.0,
               scaler=None) -> torch.Tensor:
    """One iteration of train_from_scratch.py::train_loop (lines 84-113) over the drop-in objects.

    noise / timesteps may be supplied (parity tests feed both implementations the same tensors); otherwise they are
    drawn on the device as the reference does (:85, :88-91).  Returns the (detached) loss tensor; no host sync.
    """
    device = clean_images.device
    if noise is None:
        noise = torch.randn(clean_images.shape, device=device)
    if timesteps is None:
        timesteps = torch.randint(0, noise_scheduler.config.num_train_timesteps, (clean_images.shape[0],),
                                  device=device, dtype=torch.int64)
    noisy = noise_scheduler.add_noise(clean_images, noise, timesteps)
    pred = model(noisy, timesteps, return_dict=False)[0]
    loss = mse_loss(pred, noise)
    if scaler is not None:
        scaler.scale(loss).backward()
    else:
        loss.backward()
    params = [p for p in model.parameters() if p.requires_grad]
    internal_clip = getattr(optimizer, "max_grad_norm", None)       # optim.FusedAdamW clips inside its update
    if scaler is not None:
        # Reference order (train_from_scratch.py:103-111): clip_grad_norm_ runs on the STILL-SCALED gradients, then
        # scaler.step unscales and steps.  Both optimizer paths follow it: FusedAdamW's internal clip (which would see
        # the unscaled gradients) is switched off for this step and the explicit clip below takes its place.
        if max_grad_norm is not None:
            torch.nn.utils.clip_grad_norm_(params, max_grad_norm)
        if internal_clip is not None:
            optimizer.max_grad_norm = None
        try:
            scaler.step(optimizer)
        finally:
            if internal_clip is not None:
                optimizer.max_grad_norm = internal_clip
        scaler.update()
    else:
        if max_grad_norm is not None and internal_clip is None:
            torch.nn.utils.clip_grad_norm_(params, max_grad_norm)
        optimizer.step()
    optimizer.zero_grad()
    if lr_scheduler is not None:
        lr_scheduler.step()
    return loss.detach()


def train_loop(config, model, noise_scheduler, optimizer, train_dataloader, lr_scheduler, cls, imgs_to_generate,
               device=None, evaluate_epochs=(199,), save_epochs=(199,), use_grad_scaler: Optional[bool] = None,
               num_inference_steps: int = 1000, on_epoch_end=None):
    """generator_model/train_from_scratch.py:68-133 over the drop-in objects (mlflow logging left to `on_epoch_end`).

    Per batch, in the reference's order: noise / timesteps drawn on the device (:85-91), add_noise (:93), forward +
    F.mse_loss (:95-101), scaler.scale(loss).backward() (:103), clip_grad_norm_(model.parameters(), 1.0) (:106 -- on
    the still-scaled gradients, as the reference does), scaler.step / update, zero_grad, lr_scheduler.step (:110-113),
    loss.item() (:115).  After each epoch a DDPMPipeline is built (:121); at `evaluate_epochs` the per-class images
    are generated (sampling.evaluate = :39-66) and at `save_epochs` the pipeline is saved under
    <output_dir>/models/model_<cls> (:128-131).  The reference hard-codes both lists to [199].
    Returns the list of per-epoch average losses."""
    import os
    from .pipeline import DDPMPipeline
    from .sampling import evaluate
    if device is not None:
        model.to(device)
    dev = next(model.parameters()).device
    if getattr(config, "output_dir", None) is not None:
        os.makedirs(config.output_dir, exist_ok=True)
    if use_grad_scaler is None:
        use_grad_scaler = dev.type == "cuda"
    scaler = torch.amp.GradScaler(dev.type) if use_grad_scaler else None
    history = []
    for epoch in range(config.num_epochs):
        model.train()
        total_loss, n_batches = 0.0, 0
        for batch in train_dataloader:
            clean_images = (batch[0] if isinstance(batch, (tuple, list)) else batch).to(dev)
            loss = train_step(model, noise_scheduler, optimizer, clean_images, lr_scheduler=lr_scheduler,
                              max_grad_norm=1.0, scaler=scaler)
            total_loss += loss.item()
            n_batches += 1
        avg_loss = total_loss / max(n_batches, 1)
        history.append(avg_loss)
        print(f"Epoch {epoch + 1}: Loss = {avg_loss:.4f}")
        pipeline = DDPMPipeline(unet=model, scheduler=noise_scheduler)
        if epoch in evaluate_epochs:
            model.eval()
            evaluate(config, epoch, pipeline, cls, imgs_to_generate, num_inference_steps=num_inference_steps)
        if epoch in save_epochs:
            path_model = os.path.join(config.output_dir, "models", f"model_{cls}")
            pipeline.save_pretrained(path_model)
            print(f"  Model saved at {path_model}")
        if on_epoch_end is not None:
            on_epoch_end(epoch, avg_loss, pipeline)
    return history


def train_epoch_with_accumulation(model, noise_scheduler, optimizer, batches, lr_scheduler=None,
                                  accumulation_steps: int = 1, max_grad_norm: Optional[float] = 1.0,
                                  draw=None) -> float:
    """One epoch of the LoRA trainers' loop (train_with_lora_all_classes.py:123-176, train_with_lora_per_class.py
    likewise), restated over the drop-in objects for the pixel-space UNet2DModel:

        optimizer.zero_grad()
        for step, batch: loss = mse(model(add_noise(x, noise, t), t), noise) / accumulation_steps; loss.backward()
            every accumulation_steps: clip_grad_norm_(model.parameters(), 1.0); optimizer.step(); zero_grad();
                                      lr_scheduler.step()
        return mean over batches of loss * accumulation_steps

    A trailing partial group of batches leaves its gradients unapplied, exactly as the reference does.  `batches` yields
    image tensors (or tuples whose first element is one); `draw(x) -> (noise, timesteps)` lets tests inject the draws,
    default: device randn / randint as the reference (:132-135).  (The VAE / text-encoder / auxiliary cosine loss of
    the Stable-Diffusion scripts are out of scope: SURVEY.md §2.1 row 5.)"""
    if accumulation_steps < 1:
        raise ValueError("accumulation_steps must be >= 1")
    total, count = 0.0, 0
    optimizer.zero_grad()
    T = noise_scheduler.config.num_train_timesteps
    for step, batch in enumerate(batches):
        x = batch[0] if isinstance(batch, (tuple, list)) else batch
        if draw is not None:
            noise, t = draw(x)
        else:
            noise = torch.randn_like(x)
            t = torch.randint(0, T, (x.shape[0],), device=x.device)
        pred = model(noise_scheduler.add_noise(x, noise, t), t).sample
        loss = mse_loss(pred, noise) / accumulation_steps
        loss.backward()
        if (step + 1) % accumulation_steps == 0:
            if max_grad_norm is not None and getattr(optimizer, "max_grad_norm", None) is None:
                torch.nn.utils.clip_grad_norm_(list(model.parameters()), max_grad_norm)
            optimizer.step()
            optimizer.zero_grad()
            if lr_scheduler is not None:
                lr_scheduler.step()
        total += loss.item() * accumulation_steps
        count += 1
    return total / max(count, 1)


def get_cosine_schedule_with_warmup(optimizer, num_warmup_steps: int, num_training_steps: int,
                                    num_cycles: float = 0.5, last_epoch: int = -1):
    """diffusers.optimization.get_cosine_schedule_with_warmup (train_from_scratch.py:274-278); host-side scalar."""
    import math

    def lr_lambda(current_step):
        if current_step < num_warmup_steps:
            return float(current_step) / float(max(1, num_warmup_steps))
        progress = float(current_step - num_warmup_steps) / float(max(1, num_training_steps - num_warmup_steps))
        return max(0.0, 0.5 * (1.0 + math.cos(math.pi * float(num_cycles) * 2.0 * progress)))

    return torch.optim.lr_scheduler.LambdaLR(optimizer, lr_lambda, last_epoch)
