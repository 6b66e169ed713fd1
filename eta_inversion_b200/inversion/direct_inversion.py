"""Direct inversion (reference: modules/inversion/direct_inversion.py:8-64): plain DDIM denoising where the source
row is pinned to the stored inversion latent after every step (fused into the scheduler kernel as `pin_src`)."""
from __future__ import annotations

from typing import Any, Dict, Optional

import torch

from .diffusion_inversion import DiffusionInversion


class DirectInversion(DiffusionInversion):
    def predict_step_backward(self, latent, t, context, guidance_scale_bwd: Optional[float] = None,
                              source_latent_prev: Optional[torch.Tensor] = None):
        guidance_scale_bwd = guidance_scale_bwd or self.guidance_scale_bwd
        latent = self.controller.begin_step(latent=latent, t=t)
        eps_raw, g = self._unet_eps(latent, t, context, guidance_scale_bwd)
        pin = None if source_latent_prev is None else source_latent_prev.float().contiguous()
        new_latent, noise_pred = self.scheduler_bwd.fused_step(eps_raw, t, latent.float().contiguous(), g, pin_src=pin)
        new_latent = self.controller.end_step(latent=new_latent, noise_pred=noise_pred, t=t)
        return new_latent, noise_pred

    def diffusion_backward(self, latent, context, inv_result: Dict[str, Any]) -> torch.Tensor:
        for i, t in enumerate(self.pbar(self.scheduler_bwd.timesteps, desc="backward")):
            latent, noise_pred = self.predict_step_backward(latent, t, context,
                                                            source_latent_prev=inv_result["latents"][-(i + 2)])
        return latent

    def invert(self, image, prompt: Optional[str] = None, context: Optional[torch.Tensor] = None,
               guidance_scale_fwd: Optional[float] = None, inv_cfg=None) -> Dict[str, Any]:
        return super().invert(image, prompt, context, guidance_scale_fwd)
