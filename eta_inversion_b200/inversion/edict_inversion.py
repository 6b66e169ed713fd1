"""EDICT: exact diffusion inversion via coupled transformations (modules/inversion/edict_inversion.py:17-446).

Two coupled latents (x, y) are updated alternately: each scheduler step moves one latent using the noise predicted from
the OTHER one, and after every step the pair is mixed (weight p = 0.93) so that the two copies do not drift apart.  Both
operations are invertible in closed form, which is what makes the inversion exact.  The UNet forwards (two per step) go
through the engine; the affine scheduler / mixing arithmetic on the 16 K-element latents is element-wise fp32 on the
device."""
from __future__ import annotations

import contextlib
import math
from typing import Iterator, List, Optional, Tuple

import torch

from ..editing.controller import ControllerBase, ControllerEmpty, EdictController
from ..inverse_schedulers import DDIMScheduler
from ..inverse_schedulers.schedulers import DDIMSchedulerOutput
from .diffusion_inversion import DiffusionInversion


class EdictSchedulerBase:
    """Wraps the DDIM scheduler constants; timesteps may be fractional (edict_inversion.py:82-111)."""

    def __init__(self, scheduler: Optional[DDIMScheduler] = None) -> None:
        if scheduler is None:
            scheduler = DDIMScheduler(beta_start=0.00085, beta_end=0.012, beta_schedule="scaled_linear",
                                      num_train_timesteps=1000, clip_sample=False, set_alpha_to_one=False)
        self.scheduler = scheduler

    def set_timesteps(self, num_inference_steps: int) -> None:
        self.scheduler.set_timesteps(num_inference_steps)

    @property
    def config(self):
        return self.scheduler.config

    @property
    def timesteps(self) -> torch.Tensor:
        return self.scheduler.timesteps

    @property
    def num_inference_steps(self) -> int:
        return self.scheduler.num_inference_steps

    @property
    def alphas_cumprod(self) -> torch.Tensor:
        return self.scheduler.alphas_cumprod

    def get_alpha_and_beta(self, t) -> Tuple[float, float]:
        """alpha_bar(t), 1 - alpha_bar(t).  Integer t: table lookup; t < 0: final alpha; fractional t: the reference's
        interpolation low*rem + high*(1 - rem) (sic, edict_inversion.py:102-109)."""
        if isinstance(t, int) or (torch.is_tensor(t) and t.dtype == torch.long):
            a = float(self.scheduler.alphas_cumprod[int(t)])
            return a, 1 - a
        t = float(t)
        if t < 0:
            a = float(self.scheduler.final_alpha_cumprod)
            return a, 1 - a
        low, high = math.floor(t), math.ceil(t)
        rem = t - low
        a = float(self.scheduler.alphas_cumprod[low]) * rem + float(self.scheduler.alphas_cumprod[high]) * (1 - rem)
        return a, 1 - a

    def _prev(self, timestep) -> float:
        return float(timestep) - self.config.num_train_timesteps / self.num_inference_steps

    def _get_variance(self, timestep, prev_timestep) -> float:
        a_t, a_p = self.get_alpha_and_beta(timestep)[0], self.get_alpha_and_beta(prev_timestep)[0]
        return ((1 - a_p) / (1 - a_t)) * (1 - a_t / a_p)

    def step(self, model_output, timestep, sample, eta: float = 0, variance_noise=None):
        raise NotImplementedError


class EdictScheduler(EdictSchedulerBase):
    """Backward (denoising) step written as an affine map of (sample, eps) (edict_inversion.py:144-179)."""

    def step(self, model_output, timestep, sample, eta: float = 0, variance_noise=None) -> DDIMSchedulerOutput:
        if self.num_inference_steps is None:
            raise ValueError("Number of inference steps is 'None', you need to run 'set_timesteps' after creating the scheduler")
        if int(timestep) > int(self.timesteps.max()):
            raise NotImplementedError("Need to double check what the overflow is")
        prev = self._prev(timestep)
        a_t, b_t = self.get_alpha_and_beta(timestep)
        a_p, _ = self.get_alpha_and_beta(prev)
        std = eta * self._get_variance(timestep, prev) ** 0.5
        q = (a_t / a_p) ** 0.5
        out = (1. / q) * sample - (1. / q) * (b_t ** 0.5) * model_output + ((1 - a_p - std ** 2) ** 0.5) * model_output
        if eta > 0:
            out = out + std * variance_noise
        return DDIMSchedulerOutput(out, None)


class EdictSchedulerInverse(EdictSchedulerBase):
    """Forward (inversion) step: the exact inverse of EdictScheduler.step at eta = 0 (edict_inversion.py:194-222)."""

    @property
    def timesteps(self) -> torch.Tensor:
        return self.scheduler.timesteps.flip(0)

    def step(self, model_output, timestep, sample, eta: float = 0, variance_noise=None) -> DDIMSchedulerOutput:
        if self.num_inference_steps is None:
            raise ValueError("Number of inference steps is 'None', you need to run 'set_timesteps' after creating the scheduler")
        if int(timestep) > int(self.timesteps.max()):
            raise NotImplementedError
        a_t, b_t = self.get_alpha_and_beta(timestep)
        a_p, _ = self.get_alpha_and_beta(self._prev(timestep))
        q = (a_t / a_p) ** 0.5
        return DDIMSchedulerOutput(q * sample + (b_t ** 0.5) * model_output - q * ((1 - a_p) ** 0.5) * model_output, None)


class EdictInversion(DiffusionInversion):
    dft_mix_weight = 0.93
    dft_leapfrog_steps = True
    dft_init_image_strength = 0.8

    def __init__(self, model, scheduler: Optional[str] = None, num_inference_steps: Optional[int] = None,
                 guidance_scale_bwd: Optional[float] = None, guidance_scale_fwd: Optional[float] = None,
                 verbose: bool = False, mix_weight: float = 0.93, leapfrog_steps: bool = True,
                 init_image_strength: float = 1.0, prec=torch.float32) -> None:
        guidance_scale_fwd = guidance_scale_fwd or 3.0
        guidance_scale_bwd = guidance_scale_bwd or 3.0
        super().__init__(model, scheduler, num_inference_steps, guidance_scale_bwd, guidance_scale_fwd, verbose)
        self.mix_weight, self.leapfrog_steps, self.init_image_strength = mix_weight, leapfrog_steps, init_image_strength
        self.t_limit = self.num_inference_steps - int(self.num_inference_steps * init_image_strength)
        self.bwd_t_to_i = {t.item(): i for i, t in enumerate(self.get_timesteps_backward())}
        self.fwd_t_to_i = {t.item(): i for i, t in enumerate(self.get_timesteps_forward())}
        with self.use_controller(None):
            pass

    @contextlib.contextmanager
    def use_controller(self, controller: Optional[ControllerBase]) -> Iterator[None]:
        self.controller = EdictController(ControllerEmpty() if controller is None else controller)
        self.controller.begin()
        yield
        self.controller.end()
        self.controller = EdictController(ControllerEmpty())

    def create_schedulers(self, model, scheduler, num_inference_steps: int, scheduler_inv_kwargs=None):
        sched, bwd, _ = super().create_schedulers(model, scheduler, num_inference_steps)
        return sched, EdictScheduler(bwd), EdictSchedulerInverse(bwd)  # the DDIM inverse scheduler is discarded

    # ---- the coupled pair -------------------------------------------------------------------------------------------
    def iter_latent_pair(self, i: int, latent_pair: List[torch.Tensor], is_fwd: bool = False):
        """(index, (base latent, model-input latent)) in EDICT's alternating order (edict_inversion.py:288-315)."""
        for latent_i in range(2):
            if is_fwd:
                if self.leapfrog_steps:
                    orig_i = len(self.scheduler_fwd.timesteps) - (i + 1)  # what i would be going the other way
                    latent_i = (latent_i + (orig_i + 1) % 2) % 2
                else:
                    latent_i = (latent_i + 1) % 2
            else:
                latent_i = (latent_i + i % 2) % 2
            yield latent_i, (latent_pair[latent_i], latent_pair[(latent_i + 1) % 2])

    def sync_latent_pair(self, latent_pair: List[torch.Tensor], is_fwd: bool) -> List[torch.Tensor]:
        p = self.mix_weight
        new = [l.clone() for l in latent_pair]
        if is_fwd:
            new[1] = (new[1] - (1 - p) * new[0]) / p
            new[0] = (new[0] - (1 - p) * new[1]) / p
        else:
            new[0] = p * new[0] + (1 - p) * new[1]
            new[1] = (1 - p) * new[0] + p * new[1]
        return new

    def predict_noise(self, latent, t, context, guidance_scale, is_fwd: bool = False, latent_idx: Optional[int] = None,
                      **kwargs) -> torch.Tensor:
        return super().predict_noise(latent, t, context, guidance_scale, is_fwd, **kwargs)

    def predict_step_forward_single(self, latent_idx, latent_base, latent_model_input, t, context, guidance_scale):
        noise_pred = self.predict_noise(latent_model_input, t, context, guidance_scale, is_fwd=True, latent_idx=latent_idx)
        return self.step_forward(noise_pred, t, latent_base).prev_sample.to(latent_base.dtype)

    def predict_step_backward_single(self, latent_idx, latent_base, latent_model_input, t, context, guidance_scale):
        self.controller.begin_step(latent_idx, latent_base, latent_model_input)
        noise_pred = self.predict_noise(latent_model_input, t, context, guidance_scale, is_fwd=False, latent_idx=latent_idx)
        new_latent = self.step_backward(noise_pred, t, latent_base).prev_sample.to(latent_base.dtype)
        return self.controller.end_step(latent=new_latent, noise_pred=noise_pred, t=t)

    def predict_step_forward(self, latent, t, context, guidance_scale_fwd=None):
        guidance_scale_fwd = guidance_scale_fwd or self.guidance_scale_fwd
        i = self.fwd_t_to_i[t.item()]
        pair = self.sync_latent_pair(latent, is_fwd=True)
        for idx, (base, model_input) in self.iter_latent_pair(i, pair, is_fwd=True):
            pair[idx] = self.predict_step_forward_single(idx, base, model_input, t, context, guidance_scale_fwd)
        return pair, None

    def predict_step_backward(self, latent, t, context, guidance_scale_bwd=None):
        guidance_scale_bwd = guidance_scale_bwd or self.guidance_scale_bwd
        i = self.bwd_t_to_i[t.item()]
        pair = latent
        for idx, (base, model_input) in self.iter_latent_pair(i, pair, is_fwd=False):
            pair[idx] = self.predict_step_backward_single(idx, base, model_input, t, context, guidance_scale_bwd)
        return self.sync_latent_pair(pair, is_fwd=False), None

    def get_timesteps_forward(self):
        ts = super().get_timesteps_forward()
        return ts[:-self.t_limit] if self.t_limit != 0 else ts

    def get_timesteps_backward(self):
        ts = super().get_timesteps_backward()
        return ts[self.t_limit:] if self.t_limit != 0 else ts

    def encode(self, image) -> List[torch.Tensor]:
        latent = super().encode(image)
        return [latent.clone(), latent.clone()]

    def decode(self, latent: List[torch.Tensor]) -> torch.Tensor:
        return super().decode(torch.cat(latent))

    def cat_latent(self, latents: List[List[torch.Tensor]]) -> List[torch.Tensor]:
        assert len(latents[0]) == 2
        return [torch.cat([latents[i][p] for i in range(len(latents))]) for p in range(2)]
