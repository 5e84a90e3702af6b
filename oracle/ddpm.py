"""Oracle restatement of diffusers==0.33.1 `DDPMScheduler` / `DDPMPipeline` (TEST INFRASTRUCTURE).

Follows SURVEY.md Appendix B.  Reference call sites:
  /root/reference/generator_model/train_from_scratch.py:270 (ctor), :89 (.config.num_train_timesteps),
  :93 (add_noise), :121 and :51-54 (DDPMPipeline construction and call).
The arithmetic lives in diffusers schedulers/scheduling_ddpm.py, pipelines/ddpm/pipeline_ddpm.py
and utils/torch_utils.py::randn_tensor (un-vendored, requirements.txt:35).
"""
from __future__ import annotations

from dataclasses import dataclass
from types import SimpleNamespace
from typing import Optional

import math

import numpy as np
import torch


def randn_tensor(shape, generator=None, device=None, dtype=None):
    """diffusers randn_tensor: a CPU generator with a non-CPU device draws on CPU then moves (B.4)."""
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
class DDPMSchedulerOutput:
    prev_sample: torch.Tensor
    pred_original_sample: Optional[torch.Tensor] = None


def make_betas(beta_schedule: str, beta_start: float, beta_end: float, num_train_timesteps: int) -> torch.Tensor:
    """diffusers' beta tables (scheduling_ddpm.py / scheduling_ddim.py __init__): "linear" is what
    train_from_scratch.py:270 uses; "scaled_linear" is the Stable-Diffusion table behind the noise_scheduler of
    train_with_lora_all_classes.py:314; "squaredcos_cap_v2" is betas_for_alpha_bar(cosine)."""
    if beta_schedule == "linear":
        return torch.linspace(beta_start, beta_end, num_train_timesteps, dtype=torch.float32)
    if beta_schedule == "scaled_linear":
        return torch.linspace(beta_start ** 0.5, beta_end ** 0.5, num_train_timesteps, dtype=torch.float32) ** 2
    if beta_schedule == "squaredcos_cap_v2":
        def alpha_bar(t):
            return math.cos((t + 0.008) / 1.008 * math.pi / 2) ** 2
        betas = []
        for i in range(num_train_timesteps):
            t1, t2 = i / num_train_timesteps, (i + 1) / num_train_timesteps
            betas.append(min(1 - alpha_bar(t2) / alpha_bar(t1), 0.999))
        return torch.tensor(betas, dtype=torch.float32)
    raise NotImplementedError(f"{beta_schedule} is not implemented")


class DDPMScheduler:
    """Appendix B.1-B.3 for the configuration the reference uses (all other knobs rejected)."""

    def __init__(self, num_train_timesteps: int = 1000, beta_start: float = 0.0001, beta_end: float = 0.02,
                 beta_schedule: str = "linear", variance_type: str = "fixed_small", clip_sample: bool = True,
                 prediction_type: str = "epsilon", clip_sample_range: float = 1.0,
                 timestep_spacing: str = "leading", steps_offset: int = 0):
        if variance_type != "fixed_small" or prediction_type != "epsilon" or timestep_spacing != "leading":
            raise NotImplementedError("oracle covers fixed_small / epsilon / leading only")
        self.config = SimpleNamespace(
            num_train_timesteps=num_train_timesteps, beta_start=beta_start, beta_end=beta_end,
            beta_schedule=beta_schedule, variance_type=variance_type, clip_sample=clip_sample,
            prediction_type=prediction_type, clip_sample_range=clip_sample_range,
            timestep_spacing=timestep_spacing, steps_offset=steps_offset)
        self.betas = make_betas(beta_schedule, beta_start, beta_end, num_train_timesteps)
        self.alphas = 1.0 - self.betas
        self.alphas_cumprod = torch.cumprod(self.alphas, dim=0)
        self.one = torch.tensor(1.0)
        self.init_noise_sigma = 1.0
        self.custom_timesteps = False
        self.num_inference_steps = None
        self.timesteps = torch.from_numpy(np.arange(0, num_train_timesteps)[::-1].copy())

    def scale_model_input(self, sample, timestep=None):
        return sample

    def set_timesteps(self, num_inference_steps: int, device=None):
        T = self.config.num_train_timesteps
        if num_inference_steps > T:
            raise ValueError(
                f"`num_inference_steps`: {num_inference_steps} cannot be larger than "
                f"`self.config.train_timesteps`: {T} as the unet model trained with this scheduler can only "
                f"handle maximal {T} timesteps.")
        self.num_inference_steps = num_inference_steps
        self.custom_timesteps = False
        step_ratio = T // num_inference_steps
        timesteps = (np.arange(0, num_inference_steps) * step_ratio).round()[::-1].copy().astype(np.int64)
        timesteps += self.config.steps_offset
        self.timesteps = torch.from_numpy(timesteps).to(device)

    def previous_timestep(self, timestep):
        if self.custom_timesteps or self.num_inference_steps:
            index = (self.timesteps == timestep).nonzero(as_tuple=True)[0][0]
            if index == self.timesteps.shape[0] - 1:
                prev_t = torch.tensor(-1)
            else:
                prev_t = self.timesteps[index + 1]
        else:
            prev_t = timestep - 1
        return prev_t

    def _get_variance(self, t):
        prev_t = self.previous_timestep(t)
        alpha_prod_t = self.alphas_cumprod[t]
        alpha_prod_t_prev = self.alphas_cumprod[prev_t] if prev_t >= 0 else self.one
        current_beta_t = 1 - alpha_prod_t / alpha_prod_t_prev
        variance = (1 - alpha_prod_t_prev) / (1 - alpha_prod_t) * current_beta_t
        return torch.clamp(variance, min=1e-20)

    def step(self, model_output, timestep, sample, generator=None, return_dict: bool = True,
             variance_noise: Optional[torch.Tensor] = None):
        """`variance_noise` is an oracle-only hook so tests can feed both implementations the same z."""
        t = timestep
        prev_t = self.previous_timestep(t)
        alpha_prod_t = self.alphas_cumprod[t]
        alpha_prod_t_prev = self.alphas_cumprod[prev_t] if prev_t >= 0 else self.one
        beta_prod_t = 1 - alpha_prod_t
        beta_prod_t_prev = 1 - alpha_prod_t_prev
        current_alpha_t = alpha_prod_t / alpha_prod_t_prev
        current_beta_t = 1 - current_alpha_t

        pred_original_sample = (sample - beta_prod_t ** 0.5 * model_output) / alpha_prod_t ** 0.5
        if self.config.clip_sample:
            pred_original_sample = pred_original_sample.clamp(-self.config.clip_sample_range,
                                                              self.config.clip_sample_range)
        pred_original_sample_coeff = (alpha_prod_t_prev ** 0.5 * current_beta_t) / beta_prod_t
        current_sample_coeff = current_alpha_t ** 0.5 * beta_prod_t_prev / beta_prod_t
        pred_prev_sample = pred_original_sample_coeff * pred_original_sample + current_sample_coeff * sample

        variance = 0
        if t > 0:
            if variance_noise is None:
                variance_noise = randn_tensor(model_output.shape, generator=generator,
                                              device=model_output.device, dtype=model_output.dtype)
            variance = (self._get_variance(t) ** 0.5) * variance_noise
        pred_prev_sample = pred_prev_sample + variance
        if not return_dict:
            return (pred_prev_sample, pred_original_sample)
        return DDPMSchedulerOutput(prev_sample=pred_prev_sample, pred_original_sample=pred_original_sample)

    def add_noise(self, original_samples, noise, timesteps):
        self.alphas_cumprod = self.alphas_cumprod.to(device=original_samples.device)
        alphas_cumprod = self.alphas_cumprod.to(dtype=original_samples.dtype)
        timesteps = timesteps.to(original_samples.device)
        sqrt_alpha_prod = alphas_cumprod[timesteps] ** 0.5
        sqrt_alpha_prod = sqrt_alpha_prod.flatten()
        while len(sqrt_alpha_prod.shape) < len(original_samples.shape):
            sqrt_alpha_prod = sqrt_alpha_prod.unsqueeze(-1)
        sqrt_one_minus_alpha_prod = (1 - alphas_cumprod[timesteps]) ** 0.5
        sqrt_one_minus_alpha_prod = sqrt_one_minus_alpha_prod.flatten()
        while len(sqrt_one_minus_alpha_prod.shape) < len(original_samples.shape):
            sqrt_one_minus_alpha_prod = sqrt_one_minus_alpha_prod.unsqueeze(-1)
        return sqrt_alpha_prod * original_samples + sqrt_one_minus_alpha_prod * noise

    def __len__(self):
        return self.config.num_train_timesteps


@dataclass
class DDIMSchedulerOutput:
    prev_sample: torch.Tensor
    pred_original_sample: Optional[torch.Tensor] = None


class DDIMScheduler:
    """diffusers 0.33.1 DDIMScheduler (epsilon prediction, leading spacing) -- the strided sampler of SURVEY.md §8(f)
    rank 4 on the same UNet.  eta = 0 is deterministic DDIM; eta = 1 has DDPM's posterior variance."""

    order = 1

    def __init__(self, num_train_timesteps: int = 1000, beta_start: float = 0.0001, beta_end: float = 0.02,
                 beta_schedule: str = "linear", clip_sample: bool = True, set_alpha_to_one: bool = True,
                 steps_offset: int = 0, prediction_type: str = "epsilon", clip_sample_range: float = 1.0,
                 timestep_spacing: str = "leading"):
        if prediction_type != "epsilon" or timestep_spacing != "leading":
            raise NotImplementedError("oracle covers epsilon / leading only")
        self.config = SimpleNamespace(
            num_train_timesteps=num_train_timesteps, beta_start=beta_start, beta_end=beta_end,
            beta_schedule=beta_schedule, clip_sample=clip_sample, set_alpha_to_one=set_alpha_to_one,
            steps_offset=steps_offset, prediction_type=prediction_type, clip_sample_range=clip_sample_range,
            timestep_spacing=timestep_spacing)
        self.betas = make_betas(beta_schedule, beta_start, beta_end, num_train_timesteps)
        self.alphas = 1.0 - self.betas
        self.alphas_cumprod = torch.cumprod(self.alphas, dim=0)
        self.final_alpha_cumprod = torch.tensor(1.0) if set_alpha_to_one else self.alphas_cumprod[0]
        self.init_noise_sigma = 1.0
        self.num_inference_steps = None
        self.timesteps = torch.from_numpy(np.arange(0, num_train_timesteps)[::-1].copy().astype(np.int64))

    def scale_model_input(self, sample, timestep=None):
        return sample

    def set_timesteps(self, num_inference_steps: int, device=None):
        T = self.config.num_train_timesteps
        if num_inference_steps > T:
            raise ValueError(
                f"`num_inference_steps`: {num_inference_steps} cannot be larger than `self.config.train_timesteps`:"
                f" {T} as the unet model trained with this scheduler can only handle"
                f" maximal {T} timesteps.")
        self.num_inference_steps = num_inference_steps
        step_ratio = T // num_inference_steps
        timesteps = (np.arange(0, num_inference_steps) * step_ratio).round()[::-1].copy().astype(np.int64)
        timesteps += self.config.steps_offset
        self.timesteps = torch.from_numpy(timesteps).to(device)

    def _get_variance(self, timestep, prev_timestep):
        alpha_prod_t = self.alphas_cumprod[timestep]
        alpha_prod_t_prev = self.alphas_cumprod[prev_timestep] if prev_timestep >= 0 else self.final_alpha_cumprod
        beta_prod_t = 1 - alpha_prod_t
        beta_prod_t_prev = 1 - alpha_prod_t_prev
        return (beta_prod_t_prev / beta_prod_t) * (1 - alpha_prod_t / alpha_prod_t_prev)

    def step(self, model_output, timestep, sample, eta: float = 0.0, use_clipped_model_output: bool = False,
             generator=None, variance_noise: Optional[torch.Tensor] = None, return_dict: bool = True):
        if self.num_inference_steps is None:
            raise ValueError("Number of inference steps is 'None', you need to run 'set_timesteps' after creating "
                             "the scheduler")
        prev_timestep = timestep - self.config.num_train_timesteps // self.num_inference_steps
        alpha_prod_t = self.alphas_cumprod[timestep]
        alpha_prod_t_prev = self.alphas_cumprod[prev_timestep] if prev_timestep >= 0 else self.final_alpha_cumprod
        beta_prod_t = 1 - alpha_prod_t
        pred_original_sample = (sample - beta_prod_t ** 0.5 * model_output) / alpha_prod_t ** 0.5
        pred_epsilon = model_output
        if self.config.clip_sample:
            pred_original_sample = pred_original_sample.clamp(-self.config.clip_sample_range,
                                                              self.config.clip_sample_range)
        variance = self._get_variance(timestep, prev_timestep)
        std_dev_t = eta * variance ** 0.5
        if use_clipped_model_output:
            pred_epsilon = (sample - alpha_prod_t ** 0.5 * pred_original_sample) / beta_prod_t ** 0.5
        pred_sample_direction = (1 - alpha_prod_t_prev - std_dev_t ** 2) ** 0.5 * pred_epsilon
        prev_sample = alpha_prod_t_prev ** 0.5 * pred_original_sample + pred_sample_direction
        if eta > 0:
            if variance_noise is not None and generator is not None:
                raise ValueError("Cannot pass both generator and variance_noise. Please make sure that either "
                                 "`generator` or `variance_noise` stays `None`.")
            if variance_noise is None:
                variance_noise = randn_tensor(model_output.shape, generator=generator, device=model_output.device,
                                              dtype=model_output.dtype)
            prev_sample = prev_sample + std_dev_t * variance_noise
        if not return_dict:
            return (prev_sample, pred_original_sample)
        return DDIMSchedulerOutput(prev_sample=prev_sample, pred_original_sample=pred_original_sample)

    add_noise = DDPMScheduler.add_noise

    def __len__(self):
        return self.config.num_train_timesteps


@dataclass
class ImagePipelineOutput:
    images: object


class DDPMPipeline:
    """Appendix B.4."""

    def __init__(self, unet, scheduler):
        self.unet = unet
        self.scheduler = scheduler

    @property
    def device(self):
        return self.unet.device

    @torch.no_grad()
    def __call__(self, batch_size: int = 1, generator=None, num_inference_steps: int = 1000,
                 output_type: Optional[str] = "pil", return_dict: bool = True):
        s = self.unet.config.sample_size
        if isinstance(s, int):
            image_shape = (batch_size, self.unet.config.in_channels, s, s)
        else:
            image_shape = (batch_size, self.unet.config.in_channels, *s)
        image = randn_tensor(image_shape, generator=generator, device=self.device, dtype=self.unet.dtype)
        self.scheduler.set_timesteps(num_inference_steps)
        for t in self.scheduler.timesteps:
            model_output = self.unet(image, t).sample
            image = self.scheduler.step(model_output, t, image, generator=generator).prev_sample
        image = (image / 2 + 0.5).clamp(0, 1)
        image = image.cpu().permute(0, 2, 3, 1).numpy()
        if output_type == "pil":
            image = numpy_to_pil(image)
        if not return_dict:
            return (image,)
        return ImagePipelineOutput(images=image)


def numpy_to_pil(images):
    from PIL import Image
    if images.ndim == 3:
        images = images[None, ...]
    images = (images * 255).round().astype("uint8")
    return [Image.fromarray(im) for im in images]
