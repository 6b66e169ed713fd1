"""Null-text inversion (reference: modules/inversion/null_text_inversion.py:13-115).

After a plain DDIM inversion, every denoising step optimises the unconditional ("null") text embedding with Adam so that
the CFG denoising step reproduces the stored inversion latent; editing then denoises with the optimised embedding of
each step.  The reference back-propagates ``mse(prev_step(eps), latent_prev)`` through the UNet with autograd
(null_text_inversion.py:72-80); here the loss tail is differentiated in closed form and the UNet part is the engine's
data-gradient pass (``etai_unet_forward_train`` + ``etai_unet_backward_ctx``: conv / linear dgrad on the GEMM kernels with
re-packed weights, flash-style attention backward, norm backward; no weight gradients exist):

    rec   = sqrt(a_p) (x - sqrt(1-a_t) e) / sqrt(a_t) + sqrt(1-a_p) e,        e = e_u + g (e_c - e_u)        (eta = 0)
    dL/de_u = (1 - g) * (sqrt(1-a_p) - sqrt(a_p) sqrt(1-a_t) / sqrt(a_t)) * 2 (rec - latent_prev) / numel

The early stop keeps the reference's ``loss.item()`` per inner step: one B=1 forward + backward is tens of milliseconds of
GPU work, so the host read-back costs nothing next to the forward/backward pairs a device-side flag would waste after the
stop condition is met.
"""
from __future__ import annotations

from typing import Any, Dict, List, Optional

import torch
from torch.optim.adam import Adam

from .. import engine as E
from .diffusion_inversion import DiffusionInversion


class NullTextInversion(DiffusionInversion):
    dft_num_inner_steps = 10
    dft_early_stop_epsilon = 1e-5

    def __init__(self, model, scheduler: Optional[str] = None, num_inference_steps: Optional[int] = None,
                 guidance_scale_bwd: Optional[float] = None, guidance_scale_fwd: Optional[float] = None,
                 verbose: bool = False, num_inner_steps: Optional[int] = None,
                 early_stop_epsilon: Optional[float] = None) -> None:
        super().__init__(model, scheduler, num_inference_steps, guidance_scale_bwd, guidance_scale_fwd, verbose)
        self.num_inner_steps = num_inner_steps or NullTextInversion.dft_num_inner_steps
        self.early_stop_epsilon = early_stop_epsilon or NullTextInversion.dft_early_stop_epsilon
        if not isinstance(self.unet, E.UNetEngine):
            raise RuntimeError("etai: null-text inversion differentiates the UNet and cannot run inside a lock-step group; "
                               "run it one edit at a time (eval.py does)")
        if scheduler not in (None, "ddim"):
            raise NotImplementedError("null-text inversion is built for the deterministic DDIM scheduler only")
        self.unet.enable_backward(1)
        self.inner_steps_taken: List[int] = []  # diagnostics: inner steps actually run per DDIM step of the last call

    def null_optimization(self, latents: List[torch.Tensor], context: torch.Tensor, num_inner_steps: int, epsilon: float
                          ) -> List[torch.Tensor]:
        uncond_embeddings, cond_embeddings = context.chunk(2)
        cond_embeddings = cond_embeddings.contiguous()
        uncond_embeddings_list = []
        latent_cur = latents[-1]
        sch = self.scheduler_bwd
        g = float(self.guidance_scale_bwd)
        self.inner_steps_taken = []
        for i in range(self.num_inference_steps):
            uncond_embeddings = uncond_embeddings.clone().detach().float().contiguous()
            optimizer = Adam([uncond_embeddings.requires_grad_(True)], lr=1e-2 * (1. - i / 100.))
            latent_prev = latents[len(latents) - i - 2].float()
            t = sch.timesteps[i]
            ti, pi = int(t), sch.prev_timestep(t)
            a_t, a_p = sch.alpha(ti), sch.alpha(pi)
            coef = (1.0 - g) * ((1.0 - a_p) ** 0.5 - (a_p ** 0.5) * ((1.0 - a_t) ** 0.5) / (a_t ** 0.5))
            x = latent_cur.float().contiguous()
            with torch.no_grad():
                eps_cond = self.predict_noise(x, t, cond_embeddings, guidance_scale=None)
                j = 0
                for j in range(num_inner_steps):
                    eps_uncond = self.unet.forward_train(x, t, uncond_embeddings)
                    rec = E.cfg_ddim_step(torch.cat([eps_uncond, eps_cond]).contiguous(), x, a_t, a_p, g, 0.0, 0.0)
                    diff = rec - latent_prev
                    loss = (diff * diff).mean()
                    uncond_embeddings.grad = self.unet.backward_ctx(diff * (2.0 * coef / diff.numel()))
                    optimizer.step()
                    if loss.item() < epsilon + i * 2e-5:
                        break
                self.inner_steps_taken.append(j + 1)
                uncond_embeddings_list.append(uncond_embeddings[:1].detach())
                ctx = torch.cat([uncond_embeddings.detach(), cond_embeddings]).contiguous()
                latent_cur, _ = self.predict_step_backward(latent_cur, t, ctx)
        return uncond_embeddings_list

    def diffusion_backward(self, latent, context, inv_result: Dict[str, Any]) -> torch.Tensor:
        half = context.shape[0] // 2
        for i, t in enumerate(self.pbar(self.scheduler_bwd.timesteps, desc="backward")):
            context[:half] = inv_result["uncond_embeddings"][i]  # in place like the reference; bumps the version
            latent, noise_pred = self.predict_step_backward(latent, t, context)
        return latent

    def invert(self, image, prompt: Optional[str] = None, context: Optional[torch.Tensor] = None,
               guidance_scale_fwd: Optional[float] = None, inv_cfg=None) -> Dict[str, Any]:
        fwd_result = super().invert(image, prompt, context, guidance_scale_fwd)
        fwd_result["uncond_embeddings"] = self.null_optimization(
            fwd_result["latents"], fwd_result["context"], self.num_inner_steps, self.early_stop_epsilon)
        return fwd_result
