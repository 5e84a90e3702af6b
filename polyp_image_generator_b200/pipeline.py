"""Drop-in `DDPMPipeline` (diffusers 0.33.1 semantics): the 1000-step reverse-diffusion sampling loop.

Call sites replaced (paths relative to /root/reference/):
  generator_model/train_from_scratch.py:121    DDPMPipeline(unet=model, scheduler=noise_scheduler)
  generator_model/train_from_scratch.py:51-54  pipeline(batch_size=..., generator=torch.Generator('cpu').manual_seed(s)).images
  generator_model/train_from_scratch.py:130    pipeline.save_pretrained(path)
RNG contract: with a CPU generator every draw happens on the CPU generator in diffusers'
order (one draw for x_T, one per step with t > 0) and is copied to the device, so sampled images are reproducible
against the reference.  With generator=None the per-step noise is drawn inside the step kernel instead.
"""
from __future__ import annotations

import json
import os
from dataclasses import dataclass
from typing import List, Optional, Union

import torch

from . import ops as _ops


def randn_tensor(shape, generator=None, device=None, dtype=None):
    """diffusers.utils.torch_utils.randn_tensor: a CPU generator with a CUDA device draws on CPU, then moves."""
    device = torch.device(device) if device is not None else torch.device("cpu")
    rand_device = device
    if generator is not None:
        gen_device_type = generator.device.type
        if gen_device_type != device.type and gen_device_type == "cpu":
            rand_device = torch.device("cpu")
        elif gen_device_type != device.type and gen_device_type == "cuda":
            raise ValueError(f"Cannot generate a {device} tensor from a generator of type {gen_device_type}.")
    return torch.randn(tuple(shape), generator=generator, device=rand_device, dtype=dtype).to(device)


@dataclass
class ImagePipelineOutput:
    images: Union[List, "object"]


class DDPMPipeline:
    def __init__(self, unet, scheduler, use_cuda_graph: bool = True):
        self.unet = unet
        self.scheduler = scheduler
        self.use_cuda_graph = use_cuda_graph      # replay the UNet forward from a CUDA graph (graphs.py) on CUDA
        self._graphs = {}

    def _graphed_forward(self, image):
        if not image.is_cuda or not hasattr(self.unet, "_run_forward"):
            return None
        key = tuple(image.shape)
        g = self._graphs.get(key)
        if g is None:
            from .graphs import GraphedUNetForward
            g = GraphedUNetForward(self.unet, image.shape[0], image.shape[2], image.shape[3])
            self._graphs = {key: g}       # keep one shape (the private pool of a graph holds all activations)
        return g

    @property
    def device(self):
        return self.unet.device

    def to(self, device):
        self.unet.to(device)
        return self

    @torch.no_grad()
    def __call__(self, batch_size: int = 1, generator=None, num_inference_steps: int = 1000,
                 output_type: Optional[str] = "pil", return_dict: bool = True):
        s = self.unet.config.sample_size
        hw = (s, s) if isinstance(s, int) else tuple(s)
        image_shape = (batch_size, self.unet.config.in_channels, *hw)
        image = randn_tensor(image_shape, generator=generator, device=self.device, dtype=torch.float32)
        self.scheduler.set_timesteps(num_inference_steps)
        fwd = self._graphed_forward(image) if self.use_cuda_graph else None
        for t in self.scheduler._ts_list:
            model_output = fwd(image, t) if fwd is not None else self.unet(image, t).sample
            image = self.scheduler.step(model_output, t, image, generator=generator,
                                        want_pred_original_sample=False).prev_sample
        return self._finish(image, output_type, return_dict)

    def _finish(self, image, output_type, return_dict):
        if output_type == "pt_raw":           # extension: raw x_0 in [-1, 1] on the device
            return ImagePipelineOutput(images=image) if return_dict else (image,)
        u8 = _ops.get().to_uint8_nhwc(image.contiguous())  # (x/2+0.5).clamp(0,1)*255 rounded, NHWC uint8
        if output_type == "uint8":            # extension: device uint8 NHWC
            images = u8
        else:
            arr = u8.cpu().numpy()
            if output_type == "pil":
                from PIL import Image
                images = [Image.fromarray(a.squeeze() if a.shape[-1] == 1 else a) for a in arr]
            else:                              # "np"/"numpy": float32 NHWC in [0, 1] quantised to 1/255 steps
                images = arr.astype("float32") / 255.0
        if not return_dict:
            return (images,)
        return ImagePipelineOutput(images=images)

    # ---- diffusers-compatible on-disk layout --------------------------------------
    def save_pretrained(self, save_directory: str, safe_serialization: bool = True):
        os.makedirs(os.path.join(save_directory, "unet"), exist_ok=True)
        os.makedirs(os.path.join(save_directory, "scheduler"), exist_ok=True)
        with open(os.path.join(save_directory, "model_index.json"), "w") as f:
            json.dump({"_class_name": type(self).__name__, "_diffusers_version": "0.33.1",
                       "scheduler": ["diffusers", type(self.scheduler).__name__],
                       "unet": ["diffusers", "UNet2DModel"]}, f, indent=2)
        self.unet.save_pretrained(os.path.join(save_directory, "unet"), safe_serialization=safe_serialization)
        cfg = dict(vars(self.scheduler.config))
        cfg.update({"_class_name": type(self.scheduler).__name__, "_diffusers_version": "0.33.1"})
        with open(os.path.join(save_directory, "scheduler", "scheduler_config.json"), "w") as f:
            json.dump(cfg, f, indent=2)

    @classmethod
    def from_pretrained(cls, directory: str):
        from . import scheduler as _sched
        from .unet import UNet2DModel
        unet = UNet2DModel.from_pretrained(os.path.join(directory, "unet"))
        with open(os.path.join(directory, "scheduler", "scheduler_config.json")) as f:
            raw = json.load(f)
        cfg = {k: v for k, v in raw.items() if not k.startswith("_")}
        sched_cls = getattr(_sched, raw.get("_class_name", "DDPMScheduler"), _sched.DDPMScheduler)
        return cls(unet=unet, scheduler=sched_cls(**cfg))


class DDIMPipeline(DDPMPipeline):
    """diffusers.DDIMPipeline: the same reverse loop with DDIMScheduler.step (eta, use_clipped_model_output) and
    50 steps by default -- SURVEY.md §8(f) rank 4 (10-40x fewer UNet calls for config 5's 1024-image sampling)."""

    @torch.no_grad()
    def __call__(self, batch_size: int = 1, generator=None, eta: float = 0.0, num_inference_steps: int = 50,
                 use_clipped_model_output: Optional[bool] = None, output_type: Optional[str] = "pil",
                 return_dict: bool = True):
        s = self.unet.config.sample_size
        hw = (s, s) if isinstance(s, int) else tuple(s)
        image_shape = (batch_size, self.unet.config.in_channels, *hw)
        image = randn_tensor(image_shape, generator=generator, device=self.device, dtype=torch.float32)
        self.scheduler.set_timesteps(num_inference_steps)
        fwd = self._graphed_forward(image) if self.use_cuda_graph else None
        for t in self.scheduler._ts_list:
            model_output = fwd(image, t) if fwd is not None else self.unet(image, t).sample
            image = self.scheduler.step(model_output, t, image, eta=eta,
                                        use_clipped_model_output=bool(use_clipped_model_output), generator=generator,
                                        want_pred_original_sample=False).prev_sample
        return self._finish(image, output_type, return_dict)


class UniPCPipeline(DDPMPipeline):
    """The unconditional reverse loop driven by UniPCMultistepScheduler (the sampler of the LoRA scripts,
    train_with_lora_all_classes.py:314; 25 steps there, :56-61) -- SURVEY.md §8(f) rank 4.  Deterministic given x_T:
    the generator is consumed once, for the initial noise."""

    @torch.no_grad()
    def __call__(self, batch_size: int = 1, generator=None, num_inference_steps: int = 25,
                 output_type: Optional[str] = "pil", return_dict: bool = True):
        s = self.unet.config.sample_size
        hw = (s, s) if isinstance(s, int) else tuple(s)
        image_shape = (batch_size, self.unet.config.in_channels, *hw)
        image = randn_tensor(image_shape, generator=generator, device=self.device, dtype=torch.float32)
        self.scheduler.set_timesteps(num_inference_steps)
        fwd = self._graphed_forward(image) if self.use_cuda_graph else None
        for t in self.scheduler._ts_list:
            model_output = fwd(image, t) if fwd is not None else self.unet(image, t).sample
            image = self.scheduler.step(model_output, t, image).prev_sample
        return self._finish(image, output_type, return_dict)
