"""Proximal negative-prompt inversion (reference: modules/inversion/proximal_negative_prompt_inversion.py:23-151).

Negative-prompt inversion whose denoising CFG uses a proximally-thresholded score difference:
  delta = eps_cond - eps_uncond ;  thr = quantile_q(|delta|) over the whole tensor (or -q if q < 0)
  l0: delta -= clamp(delta, -thr, thr)            l1: additionally shrink the survivors by thr
  eps = eps_uncond + g * delta
The UNet forward and the DDIM step are the native kernels, and so is the proximal CFG itself: ``etai_prox_guidance`` finds
the global quantile (over 2x16384 values) by radix select and applies threshold + guidance in the same launch (no sort, no
host sync; csrc/elementwise.cu).  The reconstruction-guidance branch of the
upstream ProxNPI is dead code in the reference (it asserts ref_image is None) and is not reproduced."""
from __future__ import annotations

from typing import Optional

import torch

from .. import engine as E
from .negative_prompt_inversion import NegativePromptInversion


class ProximalNegativePromptInversion(NegativePromptInversion):
    dft_prox, dft_quantile, dft_recon_lr, dft_recon_t, dft_dilate_mask = "l0", 0.7, 1, 400, 1

    def __init__(self, model, scheduler: Optional[str] = None, num_inference_steps: Optional[int] = None,
                 guidance_scale_bwd: Optional[float] = None, guidance_scale_fwd: Optional[float] = None,
                 verbose: bool = False, prox: str = "l0", quantile: float = 0.7, recon_lr: int = 1, recon_t: int = 400,
                 dilate_mask: int = 1) -> None:
        super().__init__(model, scheduler, num_inference_steps, guidance_scale_bwd, guidance_scale_fwd, verbose)
        self.prox, self.quantile, self.recon_t, self.recon_lr, self.dilate_mask = prox, quantile, recon_t, recon_lr, dilate_mask

    def proximal_guidance(self, noise_pred_uncond, noise_prediction_text, t, guidance_scale: float) -> torch.Tensor:
        if self.prox is None:
            return noise_pred_uncond + guidance_scale * (noise_prediction_text - noise_pred_uncond)
        if self.prox not in ("l0", "l1"):
            raise NotImplementedError
        return E.prox_guidance(noise_pred_uncond.float().contiguous(), noise_prediction_text.float().contiguous(),
                               float(guidance_scale), float(self.quantile), l1=self.prox == "l1")

    def _unet_eps(self, latent, t, context, guidance_scale, is_fwd: bool = False):
        # always the full [uncond, cond] batch (reference :130-150); the proximal CFG is applied here for denoising steps
        if guidance_scale is None:
            return self._forward_unet(latent, t, context), None
        if latent.shape[0] * 2 == context.shape[0]:
            latent = torch.cat([latent] * 2)
        else:
            assert latent.shape[0] == context.shape[0]
        eps = self._forward_unet(latent, t, context)
        if is_fwd:
            return eps, float(guidance_scale)
        u, c = eps.chunk(2)
        return self.proximal_guidance(u, c, t, float(guidance_scale)).contiguous(), None
