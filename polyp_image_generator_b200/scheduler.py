"""Drop-in `DDPMScheduler` (diffusers 0.33.1 semantics) whose tensor arithmetic runs in single-pass sm_100a kernels.

Call sites replaced (paths relative to /root/reference/):
  generator_model/train_from_scratch.py:270  DDPMScheduler(num_train_timesteps=...)
  generator_model/train_from_scratch.py:89   noise_scheduler.config.num_train_timesteps
  generator_model/train_from_scratch.py:93   noise_scheduler.add_noise(clean_images, noise, timesteps)
  generator_model/train_from_scratch.py:51   (inside DDPMPipeline) set_timesteps / timesteps / step
Host side (this file) builds the beta tables and per-step scalar coefficients in fp32 with the same op order as
diffusers; the per-element work is ddpm_add_noise / ddpm_scheduler_step (csrc/elementwise.cu).
"""
from __future__ import annotations

from dataclasses import dataclass
from types import SimpleNamespace
from typing import Dict, Optional, Union

import numpy as np
import torch

from . import ops as _ops


@dataclass
class DDPMSchedulerOutput:
    prev_sample: torch.Tensor
    pred_original_sample: Optional[torch.Tensor] = None


def step_coefficients(alphas_cumprod: torch.Tensor, t: int, prev_t: int) -> Dict[str, float]:
    """fp32 scalars of DDPMScheduler.step (fixed_small / epsilon), same op order as diffusers (Appendix B.3)."""
    ac = alphas_cumprod.detach().to("cpu", torch.float32)
    one = torch.tensor(1.0)
    alpha_prod_t = ac[t]
    alpha_prod_t_prev = ac[prev_t] if prev_t >= 0 else one
    beta_prod_t = 1 - alpha_prod_t
    beta_prod_t_prev = 1 - alpha_prod_t_prev
    current_alpha_t = alpha_prod_t / alpha_prod_t_prev
    current_beta_t = 1 - current_alpha_t
    c0 = (alpha_prod_t_prev ** 0.5 * current_beta_t) / beta_prod_t
    ct = current_alpha_t ** 0.5 * beta_prod_t_prev / beta_prod_t
    variance = torch.clamp((1 - alpha_prod_t_prev) / (1 - alpha_prod_t) * current_beta_t, min=1e-20)
    return {
        "sa": float(alpha_prod_t ** 0.5), "sb": float(beta_prod_t ** 0.5), "c0": float(c0), "ct": float(ct),
        "sigma": float(variance ** 0.5) if t > 0 else 0.0,
    }


def make_betas(beta_schedule: str, beta_start: float, beta_end: float, num_train_timesteps: int) -> torch.Tensor:
    """diffusers' beta tables: "linear" (train_from_scratch.py:270), "scaled_linear" (the Stable-Diffusion table of the
    noise_scheduler in train_with_lora_all_classes.py:314), "squaredcos_cap_v2" (cosine)."""
    if beta_schedule == "linear":
        return torch.linspace(beta_start, beta_end, num_train_timesteps, dtype=torch.float32)
    if beta_schedule == "scaled_linear":
        return torch.linspace(beta_start ** 0.5, beta_end ** 0.5, num_train_timesteps, dtype=torch.float32) ** 2
    if beta_schedule == "squaredcos_cap_v2":
        import math
        bar = lambda t: math.cos((t + 0.008) / 1.008 * math.pi / 2) ** 2
        T = num_train_timesteps
        return torch.tensor([min(1 - bar((i + 1) / T) / bar(i / T), 0.999) for i in range(T)], dtype=torch.float32)
    raise NotImplementedError(f"{beta_schedule} is not implemented")


class DDPMScheduler:
    """Same constructor signature and attributes as diffusers.DDPMScheduler for the configuration the reference uses
    (fixed_small variance, epsilon prediction, leading spacing; linear / scaled_linear / cosine betas); other modes
    raise NotImplementedError rather than silently computing something else."""

    order = 1

    def __init__(self, num_train_timesteps: int = 1000, beta_start: float = 0.0001, beta_end: float = 0.02,
                 beta_schedule: str = "linear", trained_betas=None, variance_type: str = "fixed_small",
                 clip_sample: bool = True, prediction_type: str = "epsilon", thresholding: bool = False,
                 dynamic_thresholding_ratio: float = 0.995, clip_sample_range: float = 1.0,
                 sample_max_value: float = 1.0, timestep_spacing: str = "leading", steps_offset: int = 0,
                 rescale_betas_zero_snr: bool = False):
        if trained_betas is not None or thresholding or rescale_betas_zero_snr:
            raise NotImplementedError("trained_betas / thresholding / rescale_betas_zero_snr are not on the hot path")
        if beta_schedule not in ("linear", "scaled_linear", "squaredcos_cap_v2"):
            raise NotImplementedError(f"{beta_schedule} is not implemented for {self.__class__}")
        if variance_type != "fixed_small" or prediction_type != "epsilon" or timestep_spacing != "leading":
            raise NotImplementedError("only variance_type='fixed_small', prediction_type='epsilon', "
                                      "timestep_spacing='leading' are implemented")
        self.config = SimpleNamespace(
            num_train_timesteps=num_train_timesteps, beta_start=beta_start, beta_end=beta_end,
            beta_schedule=beta_schedule, trained_betas=None, variance_type=variance_type, clip_sample=clip_sample,
            prediction_type=prediction_type, thresholding=False, dynamic_thresholding_ratio=dynamic_thresholding_ratio,
            clip_sample_range=clip_sample_range, sample_max_value=sample_max_value,
            timestep_spacing=timestep_spacing, steps_offset=steps_offset, rescale_betas_zero_snr=False)
        self.betas = make_betas(beta_schedule, beta_start, beta_end, num_train_timesteps)
        self.alphas = 1.0 - self.betas
        self.alphas_cumprod = torch.cumprod(self.alphas, dim=0)
        self.one = torch.tensor(1.0)
        self.init_noise_sigma = 1.0
        self.custom_timesteps = False
        self.num_inference_steps = None
        self.timesteps = torch.from_numpy(np.arange(0, num_train_timesteps)[::-1].copy())
        self._tables = {}   # device -> (sqrt_ac, sqrt_1mac) fp32 device tables for add_noise
        self._coef_cache = {}
        self._philox_offset = 0

    def __len__(self):
        return self.config.num_train_timesteps

    def scale_model_input(self, sample, timestep=None):
        return sample

    # ---- timesteps -------------------------------------------------------------------------------------
    def set_timesteps(self, num_inference_steps: Optional[int] = None, device=None, timesteps=None):
        if timesteps is not None:
            raise NotImplementedError("custom timesteps are not on the hot path")
        T = self.config.num_train_timesteps
        if num_inference_steps > T:
            raise ValueError(
                f"`num_inference_steps`: {num_inference_steps} cannot be larger than `self.config.train_timesteps`:"
                f" {T} as the unet model trained with this scheduler can only handle"
                f" maximal {T} timesteps.")
        self.num_inference_steps = num_inference_steps
        self.custom_timesteps = False
        step_ratio = T // num_inference_steps
        ts = (np.arange(0, num_inference_steps) * step_ratio).round()[::-1].copy().astype(np.int64)
        ts += self.config.steps_offset
        self.timesteps = torch.from_numpy(ts).to(device)
        self._ts_list = ts.tolist()
        self._coef_cache = {}

    def previous_timestep(self, timestep):
        t = int(timestep)
        if self.custom_timesteps or self.num_inference_steps:
            ts = self._ts_list
            idx = ts.index(t)
            return -1 if idx == len(ts) - 1 else ts[idx + 1]
        return t - 1

    # ---- forward noising ---------------------------------------------------------------------------------
    def _device_tables(self, device):
        key = str(device)
        if key not in self._tables:
            ac = self.alphas_cumprod.to("cpu", torch.float32)
            self._tables[key] = ((ac ** 0.5).to(device), ((1 - ac) ** 0.5).to(device))
        return self._tables[key]

    def add_noise(self, original_samples: torch.Tensor, noise: torch.Tensor, timesteps: torch.Tensor) -> torch.Tensor:
        if original_samples.dtype != torch.float32 or noise.dtype != torch.float32:
            raise TypeError("add_noise kernel computes in fp32 (reference call site passes fp32 images and noise)")
        if original_samples.shape != noise.shape:
            raise ValueError("original_samples and noise must have the same shape")
        self.alphas_cumprod = self.alphas_cumprod.to(device=original_samples.device)  # diffusers side effect (B.2)
        sa, sb = self._device_tables(original_samples.device)
        t = timesteps.to(device=original_samples.device, dtype=torch.int64).flatten().contiguous()
        if t.numel() != original_samples.shape[0]:
            raise ValueError("timesteps must have one entry per sample")
        if original_samples.numel() == 0:
            return torch.empty_like(original_samples)
        return _ops.get().add_noise(original_samples.contiguous(), noise.contiguous(), t, sa, sb)

    # ---- reverse step ---------------------------------------------------------------------------------------
    def _coefs(self, t: int) -> Dict[str, float]:
        c = self._coef_cache.get(t)
        if c is None:
            c = step_coefficients(self.alphas_cumprod, t, self.previous_timestep(t))
            self._coef_cache[t] = c
        return c

    def _draw_philox_state(self, x: torch.Tensor):
        """(seed, offset) of one in-kernel noise draw, taken from -- and advancing -- the device's default CUDA generator,
        the way torch's own Philox consumers do: torch.manual_seed(s) makes the stream reproducible, consecutive draws
        (any scheduler / pipeline instance, any class or epoch) never repeat, and ranks seeded differently draw
        different noise.  (Ranks seeded identically draw identical noise, exactly as torch.randn would.)"""
        if not x.is_cuda:
            self._philox_offset += 1
            return 0, self._philox_offset
        idx = x.device.index if x.device.index is not None else torch.cuda.current_device()
        gen = torch.cuda.default_generators[idx]
        off = gen.get_offset()
        gen.set_offset(off + 4)           # one Philox counter block per draw: the kernel indexes elements itself
        return gen.initial_seed(), off // 4 + 1

    def step(self, model_output: torch.Tensor, timestep: Union[int, torch.Tensor], sample: torch.Tensor,
             generator=None, return_dict: bool = True, variance_noise: Optional[torch.Tensor] = None,
             want_pred_original_sample: bool = True):
        """x_{t-1} from (eps_hat, x_t).  Noise: `variance_noise` if given; else a CPU `generator` is consumed exactly
        as diffusers' randn_tensor does (draw on CPU, copy to the device); with generator=None the N(0,1) draw happens
        inside the step kernel (Philox), which removes the z read from HBM."""
        t = int(timestep)
        c = self._coefs(t)
        clip = float(self.config.clip_sample_range) if self.config.clip_sample else 0.0
        ops = _ops.get()
        if model_output.dtype != torch.float32 or sample.dtype != torch.float32:
            raise TypeError("scheduler step kernel computes in fp32 (the pipeline runs the scheduler in fp32)")
        eps, x = model_output.contiguous(), sample.contiguous()
        if t > 0 and variance_noise is None and generator is None:
            seed, offset = self._draw_philox_state(x)
            prev = ops.scheduler_step_philox(eps, x, c["sa"], c["sb"], c["c0"], c["ct"], c["sigma"], clip,
                                             seed & 0xFFFFFFFFFFFFFFFF, offset)
            x0 = None
        else:
            z = None
            if t > 0:
                if variance_noise is None:
                    from .pipeline import randn_tensor
                    variance_noise = randn_tensor(model_output.shape, generator=generator, device=model_output.device,
                                                  dtype=model_output.dtype)
                z = variance_noise.contiguous()
            prev, x0 = ops.scheduler_step(eps, x, z, c["sa"], c["sb"], c["c0"], c["ct"], c["sigma"], clip,
                                          want_x0=want_pred_original_sample)
        if not return_dict:
            return (prev, x0)
        return DDPMSchedulerOutput(prev_sample=prev, pred_original_sample=x0)


# ---------------------------------------------------------------------------------------------------------------
# DDIM: the strided sampler on the same UNet
# ---------------------------------------------------------------------------------------------------------------
@dataclass
class DDIMSchedulerOutput:
    prev_sample: torch.Tensor
    pred_original_sample: Optional[torch.Tensor] = None


def ddim_coefficients(alphas_cumprod: torch.Tensor, final_alpha_cumprod: torch.Tensor, t: int, prev_t: int,
                      eta: float) -> Dict[str, float]:
    """fp32 scalars of DDIMScheduler.step, same op order as diffusers (0-dim fp32 tensor arithmetic)."""
    ac = alphas_cumprod.detach().to("cpu", torch.float32)
    alpha_prod_t = ac[t]
    alpha_prod_t_prev = ac[prev_t] if prev_t >= 0 else final_alpha_cumprod.detach().to("cpu", torch.float32)
    beta_prod_t = 1 - alpha_prod_t
    beta_prod_t_prev = 1 - alpha_prod_t_prev
    variance = (beta_prod_t_prev / beta_prod_t) * (1 - alpha_prod_t / alpha_prod_t_prev)
    std_dev_t = eta * variance ** 0.5
    return {"sa": float(alpha_prod_t ** 0.5), "sb": float(beta_prod_t ** 0.5), "sap": float(alpha_prod_t_prev ** 0.5),
            "dir": float((1 - alpha_prod_t_prev - std_dev_t ** 2) ** 0.5), "sigma": float(std_dev_t)}


class DDIMScheduler(DDPMScheduler):
    """diffusers.DDIMScheduler for epsilon prediction / leading spacing; `step` is one single-pass kernel
    (ddpm_ddim_step).  With num_inference_steps = 50 the 1024-image-per-class sampling of BASELINE configs[4] needs 20x
    fewer UNet forwards than the 1000-step DDPM loop."""

    def __init__(self, num_train_timesteps: int = 1000, beta_start: float = 0.0001, beta_end: float = 0.02,
                 beta_schedule: str = "linear", trained_betas=None, clip_sample: bool = True,
                 set_alpha_to_one: bool = True, steps_offset: int = 0, prediction_type: str = "epsilon",
                 thresholding: bool = False, dynamic_thresholding_ratio: float = 0.995, clip_sample_range: float = 1.0,
                 sample_max_value: float = 1.0, timestep_spacing: str = "leading",
                 rescale_betas_zero_snr: bool = False):
        super().__init__(num_train_timesteps=num_train_timesteps, beta_start=beta_start, beta_end=beta_end,
                         beta_schedule=beta_schedule, trained_betas=trained_betas, clip_sample=clip_sample,
                         prediction_type=prediction_type, thresholding=thresholding,
                         dynamic_thresholding_ratio=dynamic_thresholding_ratio, clip_sample_range=clip_sample_range,
                         sample_max_value=sample_max_value, timestep_spacing=timestep_spacing,
                         steps_offset=steps_offset, rescale_betas_zero_snr=rescale_betas_zero_snr)
        del self.config.variance_type
        self.config.set_alpha_to_one = set_alpha_to_one
        self.final_alpha_cumprod = torch.tensor(1.0) if set_alpha_to_one else self.alphas_cumprod[0]
        self.num_inference_steps = None

    def previous_timestep(self, timestep):
        return int(timestep) - self.config.num_train_timesteps // self.num_inference_steps

    def step(self, model_output: torch.Tensor, timestep: Union[int, torch.Tensor], sample: torch.Tensor,
             eta: float = 0.0, use_clipped_model_output: bool = False, generator=None,
             variance_noise: Optional[torch.Tensor] = None, return_dict: bool = True,
             want_pred_original_sample: bool = True):
        if self.num_inference_steps is None:
            raise ValueError("Number of inference steps is 'None', you need to run 'set_timesteps' after creating "
                             "the scheduler")
        if model_output.dtype != torch.float32 or sample.dtype != torch.float32:
            raise TypeError("scheduler step kernel computes in fp32 (the pipeline runs the scheduler in fp32)")
        t = int(timestep)
        key = (t, float(eta))
        c = self._coef_cache.get(key)
        if c is None:
            c = ddim_coefficients(self.alphas_cumprod, self.final_alpha_cumprod, t, self.previous_timestep(t), eta)
            self._coef_cache[key] = c
        clip = float(self.config.clip_sample_range) if self.config.clip_sample else 0.0
        z = None
        if eta > 0:
            if variance_noise is not None and generator is not None:
                raise ValueError("Cannot pass both generator and variance_noise. Please make sure that either "
                                 "`generator` or `variance_noise` stays `None`.")
            if variance_noise is None:
                from .pipeline import randn_tensor
                variance_noise = randn_tensor(model_output.shape, generator=generator, device=model_output.device,
                                              dtype=model_output.dtype)
            z = variance_noise.contiguous()
        prev, x0 = _ops.get().ddim_step(model_output.contiguous(), sample.contiguous(), z, c["sa"], c["sb"], c["sap"],
                                        c["dir"], c["sigma"], clip, use_clipped_model_output,
                                        want_x0=want_pred_original_sample)
        if not return_dict:
            return (prev, x0)
        return DDIMSchedulerOutput(prev_sample=prev, pred_original_sample=x0)
