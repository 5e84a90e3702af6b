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


# ---------------------------------------------------------------------------------------------------------------
# UniPC: the multistep predictor-corrector sampler of the LoRA scripts
# ---------------------------------------------------------------------------------------------------------------
@dataclass
class SchedulerOutput:
    prev_sample: torch.Tensor


def _alpha_sigma(sigma: torch.Tensor):
    """(alpha_t, sigma_t) of a VP noise level sigma = sqrt((1 - abar) / abar), as 0-dim fp32 tensors."""
    alpha_t = 1 / ((sigma ** 2 + 1) ** 0.5)
    return alpha_t, sigma * alpha_t


def _lam(sigma: torch.Tensor):
    a, s = _alpha_sigma(sigma)
    return torch.log(a) - torch.log(s)


def unipc_update_coefficients(sigmas: torch.Tensor, i_from: int, i_to: int, i_hist: Optional[int], corrector: bool,
                              solver_type: str = "bh2") -> Dict[str, float]:
    """Scalars of one UniPC-B(h) update x_{i_to} <- x_{i_from} in data-prediction form, orders 1 and 2:

        out = (cx x - cm m0) - cb [ rho0 (m1 - m0) / rk  (+)  rho_t (mt - m0) ]

    `i_hist` is the sigma index of the older model output m1 (None: no history term), `corrector` adds the rho_t term.
    All arithmetic on 0-dim fp32 CPU tensors in diffusers' op order (UniPCMultistepScheduler.multistep_uni_{p,c}_bh_update),
    so the floats handed to the kernel are the ones diffusers broadcasts."""
    alpha_t, sigma_t = _alpha_sigma(sigmas[i_to])
    _, sigma_s0 = _alpha_sigma(sigmas[i_from])
    lambda_s0 = _lam(sigmas[i_from])
    h = _lam(sigmas[i_to]) - lambda_s0
    hh = -h
    h_phi_1 = torch.expm1(hh)
    B_h = hh if solver_type == "bh1" else torch.expm1(hh)
    out = {"cx": float(sigma_t / sigma_s0), "cm": float(alpha_t * h_phi_1), "cb": float(alpha_t * B_h),
           "rk": 1.0, "rho0": 0.0, "rho_t": 0.0}
    if i_hist is None:
        if corrector:
            out["rho_t"] = 0.5              # order-1 corrector: closed form
        return out
    rk = (_lam(sigmas[i_hist]) - lambda_s0) / h
    out["rk"] = float(rk)
    if not corrector:
        out["rho0"] = 0.5                   # order-2 predictor: closed form
        return out
    # order-2 corrector: R rho = b with R = [[1, 1], [rk, 1]], b_i = h_phi_{i+1} i! / B(h)
    rks = torch.tensor([rk, 1.0])
    h_phi_k = h_phi_1 / hh - 1
    b, fact = [], 1
    for i in (1, 2):
        b.append(h_phi_k * fact / B_h)
        fact *= i + 1
        h_phi_k = h_phi_k / hh - 1 / fact
    R = torch.stack([torch.pow(rks, 0), torch.pow(rks, 1)])
    rho = torch.linalg.solve(R, torch.tensor(b)).to(torch.float32)
    out["rho0"], out["rho_t"] = float(rho[0]), float(rho[1])
    return out


class UniPCMultistepScheduler:
    """diffusers.UniPCMultistepScheduler in its default configuration -- what `UniPCMultistepScheduler.from_pretrained(
    "CompVis/stable-diffusion-v1-4", subfolder="scheduler")` (train_with_lora_all_classes.py:314,
    train_with_lora_per_class.py:308) yields: solver_order 2, epsilon prediction, predict_x0, "bh2", lower_order_final,
    linspace spacing, final sigma 0.  Other modes raise NotImplementedError.

    `step` is two or three single-pass kernels (ddpm_unipc_x0, ddpm_unipc_update for the corrector and the predictor).
    Every scalar depends only on (step index, order), so they are tabulated once per `set_timesteps`; the loop does no
    host tensor arithmetic."""

    order = 1

    def __init__(self, num_train_timesteps: int = 1000, beta_start: float = 0.0001, beta_end: float = 0.02,
                 beta_schedule: str = "linear", trained_betas=None, solver_order: int = 2,
                 prediction_type: str = "epsilon", thresholding: bool = False, dynamic_thresholding_ratio: float = 0.995,
                 sample_max_value: float = 1.0, predict_x0: bool = True, solver_type: str = "bh2",
                 lower_order_final: bool = True, disable_corrector=(), solver_p=None, use_karras_sigmas: bool = False,
                 timestep_spacing: str = "linspace", steps_offset: int = 0, final_sigmas_type: str = "zero",
                 rescale_betas_zero_snr: bool = False):
        if trained_betas is not None or thresholding or solver_p is not None or use_karras_sigmas or \
                rescale_betas_zero_snr:
            raise NotImplementedError("trained_betas / thresholding / solver_p / karras sigmas are not on the hot path")
        if prediction_type != "epsilon" or not predict_x0 or solver_type != "bh2" or \
                timestep_spacing != "linspace" or final_sigmas_type != "zero" or solver_order not in (1, 2):
            # ("bh1" with a final sigma of 0 has B(h) = -inf on the last step and yields NaN in diffusers itself)
            raise NotImplementedError("only epsilon / predict_x0 / bh2 / linspace / final sigma 0 / order <= 2")
        self.config = SimpleNamespace(
            num_train_timesteps=num_train_timesteps, beta_start=beta_start, beta_end=beta_end,
            beta_schedule=beta_schedule, trained_betas=None, solver_order=solver_order,
            prediction_type=prediction_type, thresholding=False, dynamic_thresholding_ratio=dynamic_thresholding_ratio,
            sample_max_value=sample_max_value, predict_x0=True, solver_type=solver_type,
            lower_order_final=lower_order_final, disable_corrector=list(disable_corrector), solver_p=None,
            use_karras_sigmas=False, timestep_spacing=timestep_spacing, steps_offset=steps_offset,
            final_sigmas_type=final_sigmas_type, rescale_betas_zero_snr=False)
        self.betas = make_betas(beta_schedule, beta_start, beta_end, num_train_timesteps)
        self.alphas = 1.0 - self.betas
        self.alphas_cumprod = torch.cumprod(self.alphas, dim=0)
        self.sigmas = ((1 - self.alphas_cumprod) / self.alphas_cumprod) ** 0.5
        self.init_noise_sigma = 1.0
        self.predict_x0 = True
        self.disable_corrector = list(disable_corrector)
        self.num_inference_steps = None
        self.timesteps = torch.from_numpy(
            np.linspace(0, num_train_timesteps - 1, num_train_timesteps, dtype=np.float32)[::-1].copy())
        self._tables = {}
        self._reset()

    def _reset(self):
        self.model_outputs = [None] * self.config.solver_order
        self.timestep_list = [None] * self.config.solver_order
        self.lower_order_nums = 0
        self.last_sample = None
        self.this_order = None
        self._step_index = None
        self._coefs = {}

    def __len__(self):
        return self.config.num_train_timesteps

    @property
    def step_index(self):
        return self._step_index

    def scale_model_input(self, sample, *args, **kwargs):
        return sample

    def set_timesteps(self, num_inference_steps: int, device=None):
        T = self.config.num_train_timesteps
        ts = np.linspace(0, T - 1, num_inference_steps + 1).round()[::-1][:-1].copy().astype(np.int64)
        table = (((1 - self.alphas_cumprod) / self.alphas_cumprod) ** 0.5).cpu().numpy()
        sig = np.interp(ts, np.arange(0, len(table)), table)
        self.sigmas = torch.from_numpy(np.concatenate([sig, [0.0]]).astype(np.float32))
        self.timesteps = torch.from_numpy(ts).to(device=device, dtype=torch.int64)
        self._ts_list = ts.tolist()
        self.num_inference_steps = len(ts)
        self._reset()

    # ---- forward noising (train_with_lora_all_classes.py:137 calls add_noise on this scheduler) -------------
    def add_noise(self, original_samples: torch.Tensor, noise: torch.Tensor, timesteps: torch.Tensor) -> torch.Tensor:
        """alpha_t x0 + sigma_t noise with (alpha_t, sigma_t) from the TRAINING sigma table -- the same kernel as
        DDPMScheduler.add_noise with the two per-timestep tables swapped in."""
        if original_samples.dtype != torch.float32 or noise.dtype != torch.float32:
            raise TypeError("add_noise kernel computes in fp32")
        if original_samples.shape != noise.shape:
            raise ValueError("original_samples and noise must have the same shape")
        key = str(original_samples.device)
        if key not in self._tables:
            train_sigmas = (((1 - self.alphas_cumprod) / self.alphas_cumprod) ** 0.5).to("cpu", torch.float32)
            a, s = _alpha_sigma(train_sigmas)
            self._tables[key] = (a.to(original_samples.device), s.to(original_samples.device))
        a, s = self._tables[key]
        t = timesteps.to(device=original_samples.device, dtype=torch.int64).flatten().contiguous()
        if t.numel() != original_samples.shape[0]:
            raise ValueError("timesteps must have one entry per sample")
        if original_samples.numel() == 0:
            return torch.empty_like(original_samples)
        return _ops.get().add_noise(original_samples.contiguous(), noise.contiguous(), t, a, s)

    # ---- reverse step ---------------------------------------------------------------------------------------
    def _coef(self, kind: str, i: int, order: int) -> Dict[str, float]:
        key = (kind, i, order)
        c = self._coefs.get(key)
        if c is None:
            if kind == "x0":
                a, s = _alpha_sigma(self.sigmas[i])
                c = {"alpha_t": float(a), "sigma_t": float(s)}
            elif kind == "p":   # predictor i -> i + 1, history m1 at i - 1
                c = unipc_update_coefficients(self.sigmas, i, i + 1, i - 1 if order == 2 else None, False,
                                              self.config.solver_type)
            else:               # corrector i - 1 -> i, history m1 at i - 2
                c = unipc_update_coefficients(self.sigmas, i - 1, i, i - 2 if order == 2 else None, True,
                                              self.config.solver_type)
            self._coefs[key] = c
        return c

    def step(self, model_output: torch.Tensor, timestep: Union[int, torch.Tensor], sample: torch.Tensor,
             return_dict: bool = True):
        if self.num_inference_steps is None:
            raise ValueError("Number of inference steps is 'None', you need to run 'set_timesteps' after creating "
                             "the scheduler")
        if model_output.dtype != torch.float32 or sample.dtype != torch.float32:
            raise TypeError("scheduler step kernel computes in fp32 (the pipeline runs the scheduler in fp32)")
        if self._step_index is None:
            self._step_index = self._ts_list.index(int(timestep))
        i = self._step_index
        ops = _ops.get()
        sample = sample.contiguous()
        c = self._coef("x0", i, 0)
        x0 = ops.unipc_x0(model_output.contiguous(), sample, c["sigma_t"], c["alpha_t"])
        if i > 0 and (i - 1) not in self.disable_corrector and self.last_sample is not None:
            o = self.this_order                     # the order the predictor of the previous call used
            c = self._coef("c", i, o)
            sample = ops.unipc_update(self.last_sample, self.model_outputs[-1], self.model_outputs[-2] if o == 2 else None,
                                      x0, c["cx"], c["cm"], c["cb"], c["rk"], c["rho0"], c["rho_t"])
        self.model_outputs = self.model_outputs[1:] + [x0]
        self.timestep_list = self.timestep_list[1:] + [timestep]
        order = self.config.solver_order
        if self.config.lower_order_final:
            order = min(order, self.num_inference_steps - i)
        self.this_order = min(order, self.lower_order_nums + 1)
        self.last_sample = sample
        c = self._coef("p", i, self.this_order)
        prev = ops.unipc_update(sample, x0, self.model_outputs[-2] if self.this_order == 2 else None, None,
                                c["cx"], c["cm"], c["cb"], c["rk"], c["rho0"], 0.0)
        if self.lower_order_nums < self.config.solver_order:
            self.lower_order_nums += 1
        self._step_index += 1
        if not return_dict:
            return (prev,)
        return SchedulerOutput(prev_sample=prev)
