"""DDIM schedulers of the hot path.  Host side = constants and timestep bookkeeping; the arithmetic of every step is
the fused CUDA kernel ``etai_cfg_ddim_step`` (eta_inversion_b200/csrc/elementwise.cu).

Interfaces mirrored:
  DiffusionInverseScheduler / DDIMInverseScheduler   modules/inverse_schedulers/diffusion_inverse_scheduler.py:5-29,
                                                     modules/inverse_schedulers/scheduling_ddim_inverse.py:9-143
  DDIMScheduler (the diffusers class the reference instantiates at diffusion_inversion.py:146; used through
                 from_config / config / set_timesteps / timesteps / alphas_cumprod / final_alpha_cumprod /
                 step(eps,t,x,eta=,variance_noise=).prev_sample / _get_variance)      SURVEY.md Appendix B
"""
from __future__ import annotations

from collections import namedtuple
from typing import Optional

import numpy as np
import torch

from .. import engine as E


class FrozenConfig(dict):
    """dict with attribute access; ``{**scheduler.config}`` works like diffusers' FrozenDict."""

    def __getattr__(self, k):
        try:
            return self[k]
        except KeyError as e:
            raise AttributeError(k) from e


DDIMSchedulerOutput = namedtuple("DDIMSchedulerOutput", ("prev_sample", "pred_original_sample"))


def _f(v) -> float:
    return float(v.item() if torch.is_tensor(v) else v)


class DDIMScheduler:
    """Backward (denoising) DDIM scheduler; same constructor defaults as diffusers 0.21.1."""

    _defaults = dict(num_train_timesteps=1000, beta_start=0.0001, beta_end=0.02, beta_schedule="linear",
                     trained_betas=None, clip_sample=True, set_alpha_to_one=True, steps_offset=0,
                     prediction_type="epsilon", thresholding=False, dynamic_thresholding_ratio=0.995,
                     clip_sample_range=1.0, sample_max_value=1.0, timestep_spacing="leading",
                     rescale_betas_zero_snr=False)

    def __init__(self, **kwargs):
        cfg = {**self._defaults, **{k: v for k, v in kwargs.items() if k in self._defaults}}
        self.config = FrozenConfig(cfg)
        n = cfg["num_train_timesteps"]
        if cfg["beta_schedule"] == "linear":
            betas = torch.linspace(cfg["beta_start"], cfg["beta_end"], n, dtype=torch.float32)
        elif cfg["beta_schedule"] == "scaled_linear":
            betas = torch.linspace(cfg["beta_start"] ** 0.5, cfg["beta_end"] ** 0.5, n, dtype=torch.float32) ** 2
        else:
            raise NotImplementedError(f"beta_schedule {cfg['beta_schedule']}")
        if cfg["prediction_type"] != "epsilon" or cfg["thresholding"] or cfg["timestep_spacing"] != "leading":
            raise NotImplementedError("only epsilon prediction / leading spacing (the SD-1.x setting) is built")
        self.betas = betas
        self.alphas = 1.0 - betas
        self.alphas_cumprod = torch.cumprod(self.alphas, dim=0)  # fp32 on the host like the reference
        self.final_alpha_cumprod = torch.tensor(1.0) if cfg["set_alpha_to_one"] else self.alphas_cumprod[0]
        self.init_noise_sigma = 1.0
        self.num_inference_steps = None
        self.timesteps = torch.from_numpy(np.arange(0, n)[::-1].copy().astype(np.int64))

    @classmethod
    def from_config(cls, config, **kwargs):
        return cls(**{**dict(config), **kwargs})

    def set_timesteps(self, num_inference_steps: int, device=None) -> None:
        self.num_inference_steps = num_inference_steps
        ratio = self.config.num_train_timesteps // num_inference_steps
        ts = (np.arange(0, num_inference_steps) * ratio).round()[::-1].copy().astype(np.int64)
        self.timesteps = torch.from_numpy(ts + self.config.steps_offset)  # CPU int64, descending

    # ---- constants -------------------------------------------------------------------------------
    def alpha(self, t: int) -> float:
        t = int(t)
        return _f(self.alphas_cumprod[t]) if t >= 0 else _f(self.final_alpha_cumprod)

    def prev_timestep(self, t) -> int:
        return int(t) - self.config.num_train_timesteps // self.num_inference_steps

    def _get_variance(self, timestep, prev_timestep):
        a_t = self.alphas_cumprod[int(timestep)]
        a_p = self.alphas_cumprod[int(prev_timestep)] if prev_timestep >= 0 else self.final_alpha_cumprod
        return ((1 - a_p) / (1 - a_t)) * (1 - a_t / a_p)

    # ---- fused step -----------------------------------------------------------------------------
    def fused_step(self, eps_raw: torch.Tensor, timestep, sample: torch.Tensor, guidance: Optional[float] = None,
                   eta: float = 0.0, eta_map=None, noise_cand=None, losses=None, pin_src=None):
        """CFG + DDIM denoise step in one launch.  Returns (prev_sample, eps_cfg)."""
        if self.config.clip_sample:
            raise NotImplementedError("clip_sample=True is not used on this path (diffusion_inversion.py:132-136)")
        t, p = int(timestep), self.prev_timestep(timestep)
        return E.cfg_ddim_step(eps_raw, sample, self.alpha(t), self.alpha(p), guidance, eta,
                               _f(self._get_variance(t, p)), eta_map, noise_cand, losses, pin_src, want_eps=True)

    def step(self, model_output, timestep, sample, eta=0.0, use_clipped_model_output=False, generator=None,
             variance_noise=None, return_dict=True):
        """diffusers call shape: eps is already CFG-combined.  eta may be a float or a tensor map."""
        eta_map, eta_s = None, eta
        if torch.is_tensor(eta) or hasattr(eta, "eta"):
            m = getattr(eta, "eta", eta)
            eta_map = torch.broadcast_to(m.float(), (1,) + tuple(sample.shape[1:])).contiguous()  # shared by all rows
            eta_s = 1.0
        if float(eta_s) > 0 and variance_noise is None:  # diffusers: randn of model_output.shape, every row its own noise
            variance_noise = torch.randn(sample.shape, generator=generator, device=sample.device, dtype=torch.float32)
        eps, x = model_output.float().contiguous(), sample.float().contiguous()
        B = x.shape[0]
        if variance_noise is not None and variance_noise.shape[0] == B and B > 1:
            # the fused kernel applies ONE noise tensor to all rows (the eta-inversion loop shares it); per-row noise = per-row calls
            rows = [self.fused_step(eps[i:i + 1].contiguous(), timestep, x[i:i + 1].contiguous(), None, float(eta_s), eta_map,
                                    variance_noise[i].reshape(1, -1).float().contiguous())[0] for i in range(B)]
            return DDIMSchedulerOutput(torch.cat(rows).to(sample.dtype), None)
        cand = None if variance_noise is None else variance_noise.reshape(1, -1).float().contiguous()
        prev, _ = self.fused_step(eps, timestep, x, None, float(eta_s), eta_map, cand)
        return DDIMSchedulerOutput(prev.to(sample.dtype), None)


class DiffusionInverseScheduler:
    def set_timesteps(self, num_inference_steps: int) -> None:
        raise NotImplementedError

    def step(self, noise_pred, t, latent, *args, **kwargs):
        raise NotImplementedError


class DDIMInverseScheduler(DiffusionInverseScheduler):
    """Inverse DDIM: walks z0 -> zT.  Modes as in scheduling_ddim_inverse.py:127-138."""

    Output = namedtuple("DDIMInverseSchedulerOutput", ("prev_sample",))

    def __init__(self, scheduler: DDIMScheduler, inv_steps: str = "sameshift") -> None:
        self.scheduler = scheduler
        self.is_backward = False
        self.inv_steps = inv_steps

    @staticmethod
    def from_scheduler(scheduler: DDIMScheduler, inv_steps: str = "sameshift", **kwargs) -> "DDIMInverseScheduler":
        return DDIMInverseScheduler(DDIMScheduler.from_config({**scheduler.config, **kwargs}), inv_steps=inv_steps)

    def set_timesteps(self, num_inference_steps: int) -> None:
        self.scheduler.set_timesteps(num_inference_steps)

    @property
    def timesteps(self):
        steps = reversed(self.scheduler.timesteps)  # torch.Tensor.__reversed__ == flip(0)
        if self.scheduler.config.steps_offset != 0:
            assert steps[0] == 1
        if self.inv_steps == "shiftshift":
            steps = [self.get_timestep(s, -1) for s in steps]
        return steps

    def get_timestep(self, timestep, offset: int):
        return timestep + offset * (self.scheduler.config.num_train_timesteps // self.scheduler.num_inference_steps)

    def _endpoints(self, t):
        if not self.is_backward:
            if self.inv_steps == "sameshift":
                return self.get_timestep(t, -1), t
            if self.inv_steps in ("samesame", "shiftshift"):
                return t, self.get_timestep(t, +1)
            raise Exception(self.inv_steps)
        return t, self.get_timestep(t, -1)

    def _alphas(self, t_from, t_to):
        t_from, t_to = min(int(t_from), 999), min(int(t_to), 999)
        return self.scheduler.alpha(t_from), self.scheduler.alpha(t_to)

    def ddim_step(self, sample, model_output, timestep_from, timestep_to):
        a_f, a_t = self._alphas(timestep_from, timestep_to)
        return E.cfg_ddim_step(model_output.float().contiguous(), sample.float().contiguous(), a_f, a_t).to(sample.dtype)

    def fused_step(self, eps_raw, t, latent, guidance: Optional[float] = None):
        """CFG + inverse DDIM step in one launch. Returns (next_latent, eps_cfg)."""
        a_f, a_t = self._alphas(*self._endpoints(t))
        return E.cfg_ddim_step(eps_raw, latent, a_f, a_t, guidance, want_eps=True)

    def step(self, noise_pred, t, latent):
        t_from, t_to = self._endpoints(t)
        return DDIMInverseScheduler.Output(self.ddim_step(latent, noise_pred, t_from, t_to))


class DDPMInverseScheduler(DiffusionInverseScheduler):
    """Inverse "DDPM" scheduler of DDPM inversion / CycleDiffusion (ddpm_inverse_scheduler.py:9-203): noises z0 into a
    trajectory x_T..x_1 (independently from z0, or as a Markov chain when ``markovian_forward``) and recovers, for every
    step, the noise map z_t that makes the eta = 1 DDIM step land exactly on the sampled x_{t-1}."""

    Output = namedtuple("DDPMInverseSchedulerOutput", ("prev_sample", "variance_noise"))

    def __init__(self, scheduler: DDIMScheduler, inv_steps: str = "sameshift", eta: int = 1, markovian_forward: bool = False):
        self.scheduler, self.inv_steps, self.markovian_forward = scheduler, inv_steps, markovian_forward
        self.etas, self.t_to_idx = None, None

    @staticmethod
    def from_scheduler(scheduler: DDIMScheduler, inv_steps: str = "sameshift", markovian_forward: bool = False, **kwargs):
        return DDPMInverseScheduler(DDIMScheduler.from_config({**scheduler.config, **kwargs}), inv_steps=inv_steps,
                                    markovian_forward=markovian_forward)

    def set_timesteps(self, num_inference_steps: int) -> None:
        self.scheduler.set_timesteps(num_inference_steps)
        self.t_to_idx = {int(v): k for k, v in enumerate(self.scheduler.timesteps)}
        self.etas = [1.0] * num_inference_steps

    @property
    def timesteps(self):
        return list(reversed(self.scheduler.timesteps))

    def _prev(self, t: int) -> int:
        return int(t) - self.scheduler.config.num_train_timesteps // self.scheduler.num_inference_steps

    def get_variance(self, timestep) -> float:
        return _f(self.scheduler._get_variance(int(timestep), self._prev(timestep)))

    def sample_latents(self, latent: torch.Tensor, generator: Optional[torch.Generator] = None) -> torch.Tensor:
        """[steps + 1, 4, h, w]: x_t for every inference timestep (index = position in the DESCENDING timestep list) and
        z0 last.  One N(0,1) draw per timestep, ascending in t, from ``generator`` (ddpm_inverse_scheduler.py:104-128)."""
        n = len(self.timesteps)
        gdev = generator.device if generator is not None else latent.device
        xts = torch.zeros((n,) + tuple(latent.shape[1:]), dtype=latent.dtype, device=latent.device)
        cur = latent
        for t in reversed(self.scheduler.timesteps):
            t = int(t)
            r = torch.randn(latent.shape, device=gdev, generator=generator).to(latent.device, latent.dtype)
            a_t = self.scheduler.alpha(t)
            if not self.markovian_forward:
                x = latent * (a_t ** 0.5) + r * ((1 - a_t) ** 0.5)
            else:
                tp = self._prev(t)
                ratio = a_t / (self.scheduler.alpha(tp) if tp >= 0 else 1.0)
                cur = cur * (ratio ** 0.5) + r * ((1 - ratio) ** 0.5)
                x = cur
            xts[self.t_to_idx[t]] = x[0]
        return torch.cat([xts, latent], dim=0)

    def get_sampled_latent_by_t(self, xts: torch.Tensor, t) -> torch.Tensor:
        return xts[self.t_to_idx[int(t)]][None]

    def get_eta_by_t(self, t) -> float:
        return self.etas[self.t_to_idx[int(t)]]

    def step(self, noise_pred: torch.Tensor, t, latent: torch.Tensor, xts: torch.Tensor):
        t = int(t)
        idx = self.t_to_idx[t]
        eta = self.etas[idx]
        a_t = self.scheduler.alpha(t)
        a_p = self.scheduler.alpha(self._prev(t))  # alpha(-k) = final_alpha_cumprod
        var = self.get_variance(t)
        xt, xtm1 = xts[idx][None], xts[idx + 1][None]
        x0 = (xt - (1 - a_t) ** 0.5 * noise_pred) / a_t ** 0.5
        mu = a_p ** 0.5 * x0 + (1 - a_p - eta * var) ** 0.5 * noise_pred
        z = (xtm1 - mu) / (eta * var ** 0.5)
        return DDPMInverseScheduler.Output(mu + (eta * var ** 0.5) * z, z)
