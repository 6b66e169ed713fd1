"""Negative-prompt inversion (reference: modules/inversion/negative_prompt_inversion.py:9-31): during denoising the
unconditional rows of the context are the source prompt's conditional embedding."""
from __future__ import annotations

from typing import Any, Dict, Optional

import torch

from .diffusion_inversion import DiffusionInversion


class NegativePromptInversion(DiffusionInversion):
    def diffusion_backward(self, latent, context, inv_result: Dict[str, Any]) -> torch.Tensor:
        half = context.shape[0] // 2
        last = None
        for i, t in enumerate(self.pbar(self.scheduler_bwd.timesteps, desc="backward")):
            new_uncond = inv_result["uncond_embeddings"][i]
            if new_uncond is not last:  # in-place like the reference; a write bumps the version -> K/V re-projection
                context[:half] = new_uncond
                last = new_uncond
            latent, noise_pred = self.predict_step_backward(latent, t, context)
        return latent

    def invert(self, image, prompt: Optional[str] = None, context: Optional[torch.Tensor] = None,
               guidance_scale_fwd: Optional[float] = None, inv_cfg=None) -> Dict[str, Any]:
        fwd_result = super().invert(image, prompt, context, guidance_scale_fwd)
        uncond_embeddings, cond_embeddings = fwd_result["context"].chunk(2)
        fwd_result["uncond_embeddings"] = [cond_embeddings] * self.num_inference_steps
        return fwd_result
