"""DDPM inversion ("An Edit Friendly DDPM Noise Space") and, with ``markovian_forward=True``, CycleDiffusion.

Mirrors modules/inversion/ddpm_inversion.py:10-177: the forward pass does not walk a deterministic trajectory; it samples
x_1..x_T from z0 and extracts, per step, the noise map that reproduces the sampled trajectory under the eta = 1 DDIM
update.  The backward pass replays those maps with a different prompt (first ``skip_steps`` of the steps skipped).
The UNet forwards go through the engine like every other inverter; the scheduler arithmetic on the 16 K-element latents
is a handful of fp32 element-wise ops on the device (the denoise step itself is the fused ``etai_cfg_ddim_step``)."""
from __future__ import annotations

from typing import Any, Dict, List, Optional, Tuple, Union

import torch

from ..inverse_schedulers import DDIMScheduler, DDPMInverseScheduler
from .diffusion_inversion import DiffusionInversion


class DDPMInversion(DiffusionInversion):
    dft_skip_steps = 0.36
    dft_forward_seed = 0

    def __init__(self, model, scheduler: Optional[str] = None, num_inference_steps: Optional[int] = None,
                 guidance_scale_bwd: Optional[float] = None, guidance_scale_fwd: Optional[float] = None,
                 verbose: bool = False, forward_seed: Optional[int] = 0, skip_steps: Optional[float] = None,
                 markovian_forward: bool = False, noise_device: Optional[str] = None) -> None:
        """Reference keywords (ddpm_inversion.py:17-43) plus ``noise_device`` (see EtaInversion): where the seeded
        generator of the forward trajectory lives; None = the model's device like the reference."""
        scheduler = scheduler or "ddpm"
        guidance_scale_fwd = guidance_scale_fwd or 3.5
        guidance_scale_bwd = guidance_scale_bwd or 9
        self.skip_steps = skip_steps or 0.36
        self.forward_seed = forward_seed if forward_seed >= 0 else None
        self.markovian_forward = markovian_forward
        self.noise_device = noise_device
        super().__init__(model, scheduler, num_inference_steps, guidance_scale_bwd, guidance_scale_fwd, verbose)

    def create_schedulers(self, model, scheduler, num_inference_steps: int, scheduler_inv_kwargs=None):
        # "ddpm" is simulated with the DDIM scheduler at eta = 1 (diffusion_inversion.py:141,158-160)
        bwd = DDIMScheduler.from_config({**model.scheduler.config})
        bwd.set_timesteps(num_inference_steps)
        fwd = DDPMInverseScheduler.from_scheduler(bwd, markovian_forward=self.markovian_forward)
        fwd.set_timesteps(num_inference_steps)
        assert fwd.timesteps[0] < fwd.timesteps[1], "wrong timestamp order, not increasing"
        return bwd, bwd, fwd

    # ---- noise prediction with a per-row guidance scale (ddpm_inversion.py:158-164) ---------------------------------
    def predict_noise(self, latent, t, context, guidance_scale, is_fwd: bool = False, **kwargs) -> torch.Tensor:
        if not torch.is_tensor(guidance_scale):
            return super().predict_noise(latent, t, context, guidance_scale, is_fwd=is_fwd, **kwargs)
        eps_raw, _ = self._unet_eps(latent, t, context, -1.0)  # -1: any scale that takes the full [uncond, cond] batch
        n = eps_raw.shape[0] // 2
        return eps_raw[:n] + guidance_scale * (eps_raw[n:] - eps_raw[:n])

    def predict_step_forward(self, latent, t, context, guidance_scale_fwd, xts):
        guidance_scale_fwd = guidance_scale_fwd or self.guidance_scale_fwd
        noise_pred = self.predict_noise(latent, t, context, guidance_scale_fwd, is_fwd=False)
        res = self.step_forward(noise_pred, t, latent, xts)
        return res.prev_sample, noise_pred, res.variance_noise

    def diffusion_forward(self, latent, context, guidance_scale_fwd: Optional[float] = None) -> Dict[str, Any]:
        gen = None
        if self.forward_seed is not None:
            gdev = torch.device(self.noise_device) if self.noise_device is not None else self.device
            gen = torch.Generator(gdev).manual_seed(self.forward_seed)
        xts = self.scheduler_fwd.sample_latents(latent, generator=gen)
        guidance_scale_fwd = guidance_scale_fwd or self.guidance_scale_fwd
        latents, noise_preds, variance_noises, etas = [], [], [], []
        for t in self.pbar(self.scheduler_fwd.timesteps, desc="forward"):
            latent = self.scheduler_fwd.get_sampled_latent_by_t(xts, t)
            latent, noise_pred, variance_noise = self.predict_step_forward(latent, t, context, guidance_scale_fwd, xts)
            noise_preds.append(noise_pred)
            latents.append(latent)
            variance_noises.append(variance_noise)
            etas.append(self.scheduler_fwd.get_eta_by_t(t))
        latents.append(xts[0][None])  # the final inverse latent is the sampled x_T itself
        variance_noises[0] = torch.zeros_like(variance_noises[0])
        return {"latents": latents, "noise_preds": noise_preds, "etas": etas, "variance_noises": variance_noises}

    # ---- backward ---------------------------------------------------------------------------------------------------
    def get_bwd_skip(self) -> int:
        return int(self.skip_steps * len(self.scheduler_bwd.timesteps))

    def skip_inv_result(self, inv_result: Dict[str, Any]) -> Dict[str, Any]:
        skip = self.get_bwd_skip()
        cut = {k: (inv_result[k][:-skip] if skip > 0 else inv_result[k]) for k in ("latents", "noise_preds", "variance_noises", "etas")}
        return {**inv_result, **cut}

    def sample(self, inv_result, prompt=None, context=None):
        if self.skip_steps is not None:
            inv_result = self.skip_inv_result(inv_result)
        return super().sample(inv_result, prompt=prompt, context=context)

    def predict_step_backward(self, latent, t, context, eta: float, variance_noise: torch.Tensor,
                              guidance_scale_bwd: Optional[float] = None) -> Tuple[torch.Tensor, torch.Tensor]:
        latent = self.controller.begin_step(latent=latent)
        if latent.shape[0] == 2:  # row 0 reconstructs the source with the forward scale, row 1 edits
            guidance_scale = torch.tensor([self.guidance_scale_fwd, self.guidance_scale_bwd], dtype=torch.float32)
            guidance_scale = guidance_scale.pin_memory().to(self.device, non_blocking=True)[:, None, None, None]
        else:
            assert latent.shape[0] == 1
            guidance_scale = self.guidance_scale_bwd
        noise_pred = self.predict_noise(latent, t, context, guidance_scale)
        latent = self.step_backward(noise_pred, t, latent, eta=eta, variance_noise=variance_noise).prev_sample
        latent = self.controller.end_step(latent=latent, noise_pred=noise_pred, t=t)
        return latent, noise_pred

    def diffusion_backward(self, latent, context, inv_result) -> torch.Tensor:
        etas = list(reversed(inv_result["etas"]))
        variance_noises = list(reversed(inv_result["variance_noises"]))
        timesteps = self.scheduler_bwd.timesteps[self.get_bwd_skip():]
        for i, t in enumerate(self.pbar(timesteps, desc="backward")):
            latent, _ = self.predict_step_backward(latent, t, context, etas[i], variance_noises[i], self.guidance_scale_bwd)
        return latent
