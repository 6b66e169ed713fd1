"""DDIM inversion with the noise regularisation of pix2pix-zero (modules/inversion/regularized_diffusion_inversion.py:9-137).

Every forward step nudges the predicted noise towards white Gaussian noise: `num_reg_steps` rounds of `num_ac_rolls`
gradient steps on a multi-scale auto-correlation loss plus one step on a KL term.  The gradients are taken with respect
to the 16 K-element noise prediction only (never through the UNet), so they run as a few tiny fp32 autograd graphs on
the device between two engine forwards; the roll offsets come from a host generator seeded per step like the reference."""
from __future__ import annotations

from typing import Optional, Tuple

import numpy as np
import torch
import torch.nn.functional as F

from .diffusion_inversion import DiffusionInversion


class RegularizedDiffusionInversion(DiffusionInversion):
    def __init__(self, model, scheduler: Optional[str] = None, num_inference_steps: Optional[int] = None,
                 guidance_scale_bwd: Optional[float] = None, guidance_scale_fwd: Optional[float] = None,
                 verbose: bool = False, lambda_ac: float = 20.0, lambda_kl: float = 20.0, num_reg_steps: int = 5,
                 num_ac_rolls: int = 5) -> None:
        super().__init__(model, scheduler, num_inference_steps, guidance_scale_bwd, guidance_scale_fwd, verbose)
        self.lambda_ac, self.lambda_kl = lambda_ac, lambda_kl
        self.num_reg_steps, self.num_ac_rolls = num_reg_steps, num_ac_rolls

    def auto_corr_loss(self, x: torch.Tensor, random_shift: bool = True,
                       generator: Optional[torch.Generator] = None) -> torch.Tensor:
        """Squared mean of x * roll(x) along both image axes, per channel, over an average-pooling pyramid down to 8x8."""
        assert x.shape[0] == 1
        reg_loss = 0.0
        for ch in range(x.shape[1]):
            noise = x[0, ch][None, None]
            while True:
                roll = int(torch.randint(0, noise.shape[2] // 2, (), generator=generator)) if random_shift else 1
                reg_loss = reg_loss + (noise * torch.roll(noise, shifts=roll, dims=2)).mean() ** 2
                reg_loss = reg_loss + (noise * torch.roll(noise, shifts=roll, dims=3)).mean() ** 2
                if noise.shape[2] <= 8:
                    break
                noise = F.avg_pool2d(noise, kernel_size=2)
        return reg_loss

    def kl_divergence(self, x: torch.Tensor) -> torch.Tensor:
        mu, var = x.mean(), x.var()
        return var + mu ** 2 - 1 - torch.log(var + 1e-7)

    @torch.enable_grad()
    def regularize_noise_pred(self, noise_pred: torch.Tensor, generator: Optional[torch.Generator] = None) -> torch.Tensor:
        e_t = noise_pred
        for _ in range(self.num_reg_steps):
            if self.lambda_ac > 0:
                for _ in range(self.num_ac_rolls):
                    var = e_t.detach().clone().requires_grad_(True)
                    self.auto_corr_loss(var, generator=generator).backward()
                    e_t = e_t - self.lambda_ac * (var.grad.detach() / self.num_ac_rolls)
            if self.lambda_kl > 0:
                var = e_t.detach().clone().requires_grad_(True)
                self.kl_divergence(var).backward()
                e_t = e_t - self.lambda_kl * var.grad.detach()
            e_t = e_t.detach()
        return e_t

    def predict_step_forward(self, latent, t, context, guidance_scale_fwd: Optional[float] = None
                             ) -> Tuple[torch.Tensor, torch.Tensor]:
        generator = torch.Generator().manual_seed(0)  # host generator, re-seeded every step (reference: line 119)
        guidance_scale_fwd = float(np.linspace(2, 1, 1000)[int(t)])  # the reference overrides the configured scale
        latent = self.controller.begin_step(latent=latent)
        noise_pred = self.predict_noise(latent, t, context, guidance_scale_fwd, is_fwd=True)
        noise_pred = self.regularize_noise_pred(noise_pred, generator=generator)
        new_latent = self.step_forward(noise_pred, t, latent).prev_sample
        new_latent = self.controller.end_step(latent=new_latent, noise_pred=noise_pred, t=t)
        return new_latent, noise_pred
