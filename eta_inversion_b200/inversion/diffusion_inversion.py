"""DDIM inversion + denoising loops over the native UNet (reference: modules/inversion/diffusion_inversion.py:12-542).

Same class surface as the reference.  Differences are internal to a step: instead of
``unet -> chunk -> cfg -> scheduler.step`` (about ten tiny torch kernels plus the un-fused UNet), one step is
``etai_unet_forward`` (with the controller's AttnControl) followed by ONE fused CFG+DDIM kernel; the latent
trajectory stays fp32 on the device whatever the UNet storage dtype is.
"""
from __future__ import annotations

import contextlib
from typing import Any, Dict, Iterable, Iterator, List, Optional, Tuple, Union

import torch

from ..editing.controller import ControllerBase, ControllerEmpty
from ..engine import AttnControl, h2d
from ..inverse_schedulers import DDIMInverseScheduler, DDIMScheduler


def merge_controls(a: Optional[AttnControl], b: Optional[AttnControl]) -> Optional[AttnControl]:
    """Union of two per-forward descriptors (e.g. a controller's store + an editor's self remap)."""
    if a is None or b is None:
        return a if b is None else b
    out = AttnControl(**{k: getattr(a, k) for k in a.__dataclass_fields__ if k != "_keep"})
    if b.self_rows is not None:
        if a.self_rows is not None:
            raise RuntimeError("two self-attention remaps for one forward")
        out.self_rows, out.self_layer_mask, out.self_max_tokens = b.self_rows, b.self_layer_mask, b.self_max_tokens
    if b.edit_pairs is not None:
        if a.edit_pairs is not None:
            raise RuntimeError("two cross-attention edits for one forward")
        for k in ("edit_pairs", "mapper", "blend_a", "equalizer", "alpha_step"):
            setattr(out, k, getattr(b, k))
    if b.store_rows is not None:
        if a.store_rows is not None:
            raise RuntimeError("two attention stores for one forward")
        for k in ("store_rows", "store_res", "store_down", "store_mid", "store_up"):
            setattr(out, k, getattr(b, k))
    out.conv_inject_rows = max(a.conv_inject_rows, b.conv_inject_rows)
    return out


class DiffusionInversion:
    def __init__(self, model, scheduler: Optional[str] = None, num_inference_steps: Optional[int] = None,
                 guidance_scale_bwd: Optional[float] = None, guidance_scale_fwd: Optional[float] = None,
                 verbose: bool = False) -> None:
        scheduler = scheduler or "ddim"
        self.num_inference_steps = num_inference_steps or 50
        self.guidance_scale_bwd = guidance_scale_bwd if guidance_scale_bwd is not None else 7.5
        self.guidance_scale_fwd = guidance_scale_fwd if guidance_scale_fwd is not None else 1
        self.model = model
        self.unet = model.unet
        self.device = model.device
        self.verbose = verbose
        self.controller: ControllerBase = ControllerEmpty()
        self.attn_hooks: List[Any] = []   # loop-long hooks (MasaCtrl): begin_forward(unet,B) / end_forward()
        self.unet_wrapper = None          # loop-long forward wrapper (PnP)
        model.scheduler, self.scheduler_bwd, self.scheduler_fwd = self.create_schedulers(model, scheduler, self.num_inference_steps)
        self.bwd_t_to_i = {t.item(): i for i, t in enumerate(self.scheduler_bwd.timesteps)}
        self.fwd_t_to_i = {t.item(): i for i, t in enumerate(self.scheduler_fwd.timesteps)}
        with self.use_controller(None):
            pass

    # ---- controller plumbing ---------------------------------------------------------------------
    @contextlib.contextmanager
    def use_controller(self, controller: Optional[ControllerBase]) -> Iterator[None]:
        self.controller = ControllerEmpty() if controller is None else controller
        self.controller.begin()
        yield
        self.controller.end()
        self.controller = ControllerEmpty()

    def pbar(self, it: Iterable, **kwargs) -> Iterable:
        if self.verbose:
            from tqdm import tqdm
            return tqdm(it, **kwargs)
        return it

    def create_schedulers(self, model, scheduler: Union[str, Dict[str, Any]], num_inference_steps: int,
                          scheduler_inv_kwargs: Optional[Dict[str, Any]] = None):
        scheduler_inv_kwargs = {} if scheduler_inv_kwargs is None else scheduler_inv_kwargs
        if isinstance(scheduler, str):
            name, kwargs = scheduler, {}
        elif isinstance(scheduler, dict):
            kwargs = {**scheduler}
            name = kwargs.pop("type")
            if "inv_steps" in kwargs:
                scheduler_inv_kwargs["inv_steps"] = kwargs.pop("inv_steps")
        else:
            raise Exception(type(scheduler))
        if name not in self.get_available_schedulers():
            raise NotImplementedError(f"scheduler '{name}': the DPM-Solver multistep (inverse) scheduler is not built in this "
                                      f"engine (diffusion_inversion.py:141-142 of the reference); available: "
                                      f"{self.get_available_schedulers()}")
        if name == "ddim":
            kwargs = {"clip_sample": False, "set_alpha_to_one": False, **kwargs}
        bwd = DDIMScheduler.from_config({**model.scheduler.config, **kwargs})
        bwd.set_timesteps(num_inference_steps)
        if name == "ddpm":  # "simulate ddpm with ddim and eta=1" + the DDPM inverse scheduler (diffusion_inversion.py:141,158-160)
            from ..inverse_schedulers import DDPMInverseScheduler
            fwd = DDPMInverseScheduler.from_scheduler(bwd, **scheduler_inv_kwargs)
        else:
            fwd = DDIMInverseScheduler.from_scheduler(bwd, **scheduler_inv_kwargs)
        fwd.set_timesteps(num_inference_steps)
        assert fwd.timesteps[0] < fwd.timesteps[1], "wrong timestamp order, not increasing"
        return bwd, bwd, fwd

    @staticmethod
    def get_available_schedulers() -> List[str]:
        """The reference lists ["ddim", "ddpm", "dpm"] (diffusion_inversion.py:174-176); "dpm" (diffusers' DPM-Solver
        multistep + the reference's inverse of it) is not built here, so it is not advertised: the CLIs reject it at
        argument parsing instead of failing inside the constructor."""
        return ["ddim", "ddpm"]

    # ---- VAE / text ------------------------------------------------------------------------------
    def decode(self, latent: torch.Tensor) -> torch.Tensor:
        latent = 1 / 0.18215 * latent
        return self.model.vae.decode(latent.to(self.model.vae.dtype))['sample']

    def encode(self, image: torch.Tensor) -> torch.Tensor:
        image = image.to(self.model.device, self.model.vae.dtype)
        latent = self.model.vae.encode(image)['latent_dist'].mean
        return (latent * 0.18215).float()  # latent trajectory is kept in fp32

    def _embed(self, text: str) -> torch.Tensor:
        if not getattr(self.model, "cache_text_embeddings", True):  # bench.py: every edit pays for its own CLIP passes
            return self._embed_uncached(text)
        cache = self.model.__dict__.setdefault("_text_embedding_cache", {})  # weights are frozen: same text, same embedding
        if text in cache:
            return cache[text]
        emb = self._embed_uncached(text)
        if len(cache) < 256:
            cache[text] = emb
        return emb

    def _embed_uncached(self, text: str) -> torch.Tensor:
        tok = self.model.tokenizer([text], padding="max_length", max_length=self.model.tokenizer.model_max_length,
                                   truncation=True, return_tensors="pt")
        return self.model.text_encoder(tok.input_ids)[0].float()  # ids are host data; the native tower uploads them

    def create_context(self, prompt: str, negative_prompt: str = "") -> torch.Tensor:
        if negative_prompt is not None and not getattr(self.model, "cache_text_embeddings", True):
            # no memo (bench.py): both prompts in ONE pass of the text tower instead of two (rows are independent: same
            # numbers as two B = 1 passes, half the launches)
            tok = self.model.tokenizer([negative_prompt, prompt], padding="max_length",
                                       max_length=self.model.tokenizer.model_max_length, truncation=True, return_tensors="pt")
            return self.model.text_encoder(tok.input_ids)[0].float().contiguous()
        text_embeddings = self._embed(prompt)
        if negative_prompt is not None:
            return torch.cat([self._embed(negative_prompt), text_embeddings]).contiguous()
        return text_embeddings.contiguous()

    # ---- UNet call -------------------------------------------------------------------------------
    def _forward_unet(self, latent: torch.Tensor, t, context: torch.Tensor, zero_weight_uncond: bool = False) -> torch.Tensor:
        """One UNet call with the step's attention control.  ``zero_weight_uncond``: ``latent`` holds the n latents of a
        CFG batch whose unconditional half would be combined with weight exactly 0 (``context`` is the full [2n] context);
        only the conditional rows are computed when the attention control does not touch the others -- the controllers
        see the usual 2n-row batch -- and the result is the n conditional predictions.  Otherwise the full batch runs and
        all 2n rows are returned."""
        B = latent.shape[0]
        Bc = 2 * B if zero_weight_uncond else B
        ctrl = self.controller.attn_control(self.unet, Bc)
        for h in self.attn_hooks:
            ctrl = merge_controls(ctrl, h.begin_forward(self.unet, Bc))
        if zero_weight_uncond:
            half = ctrl.drop_leading_rows(B) if ctrl is not None else None
            if ctrl is None or half is not None:
                ctrl, context = half, self._ctx_half(context, 1)
            else:
                latent = torch.cat([latent] * 2)
        latent = latent.float().contiguous()
        if self.unet_wrapper is not None:
            eps = self.unet_wrapper(self.unet, latent, t, context, control=ctrl)
        else:
            eps = self.unet(latent, t, encoder_hidden_states=context, control=ctrl)["sample"]
        self.controller.after_forward()
        for h in self.attn_hooks:
            h.end_forward()
        return eps

    def _unet_eps(self, latent: torch.Tensor, t, context: torch.Tensor, guidance_scale, is_fwd: bool = False
                  ) -> Tuple[torch.Tensor, Optional[float]]:
        """Raw UNet output and the CFG scale still to be applied (None = rows are final).  Implements the batch
        shortcuts of diffusion_inversion.py:263-284 (scale 0 -> uncond rows only, 1 -> cond rows only)."""
        if guidance_scale is None:
            return self._forward_unet(latent, t, context), None
        if latent.shape[0] * 2 == context.shape[0]:
            latent = torch.cat([latent] * 2)
        else:
            assert latent.shape[0] == context.shape[0]
        n = latent.shape[0] // 2
        if isinstance(guidance_scale, (int, float)) and guidance_scale == 0:
            return self._forward_unet(latent[:n], t, self._ctx_half(context, 0)), None
        if isinstance(guidance_scale, (int, float)) and guidance_scale == 1:
            return self._forward_unet(latent[n:], t, self._ctx_half(context, 1)), None
        return self._forward_unet(latent, t, context), float(guidance_scale)

    def _ctx_half(self, context: torch.Tensor, half: int) -> torch.Tensor:
        # stable slice object per (context, half) so the engine's projected-K/V cache survives the loop
        k = getattr(self, "_ctx_half_key", None)
        if k is None or k[0] is not context or k[1] != context._version or k[2] != half:
            n = context.shape[0] // 2
            self._ctx_half_key = (context, context._version, half)  # strong ref: addresses get reused across edits
            self._ctx_half_val = context[half * n:(half + 1) * n].contiguous()
        return self._ctx_half_val

    def predict_noise(self, latent: torch.Tensor, t, context: torch.Tensor, guidance_scale, is_fwd: bool = False,
                      **kwargs) -> torch.Tensor:
        """CFG-combined noise prediction (API parity; the loops use the fused step instead)."""
        eps_raw, g = self._unet_eps(latent, t, context, guidance_scale, is_fwd=is_fwd)
        if g is None:
            return eps_raw
        n = eps_raw.shape[0] // 2
        from .. import engine as E
        _, eps = E.cfg_ddim_step(eps_raw, torch.zeros_like(eps_raw[:n]), 1.0, 1.0, g, want_eps=True)
        return eps

    def step_forward(self, noise_pred, t, latent, *args, **kwargs) -> Any:
        return self.scheduler_fwd.step(noise_pred, t, latent, *args, **kwargs)

    def step_backward(self, noise_pred, t, latent, *args, **kwargs) -> Any:
        return self.scheduler_bwd.step(noise_pred, t, latent, *args, **kwargs)

    # ---- one step --------------------------------------------------------------------------------
    def predict_step_forward(self, latent, t, context, guidance_scale_fwd: Optional[float] = None):
        guidance_scale_fwd = guidance_scale_fwd or self.guidance_scale_fwd
        latent = self.controller.begin_step(latent=latent)
        eps_raw, g = self._unet_eps(latent, t, context, guidance_scale_fwd, is_fwd=True)
        new_latent, noise_pred = self.scheduler_fwd.fused_step(eps_raw, t, latent.float().contiguous(), g)
        new_latent = self.controller.end_step(latent=new_latent, noise_pred=noise_pred, t=t)
        return new_latent, noise_pred

    def predict_step_backward(self, latent, t, context, guidance_scale_bwd: Optional[float] = None):
        guidance_scale_bwd = guidance_scale_bwd or self.guidance_scale_bwd
        latent = self.controller.begin_step(latent=latent, t=t)
        eps_raw, g = self._unet_eps(latent, t, context, guidance_scale_bwd)
        new_latent, noise_pred = self.scheduler_bwd.fused_step(eps_raw, t, latent.float().contiguous(), g)
        new_latent = self.controller.end_step(latent=new_latent, noise_pred=noise_pred, t=t)
        return new_latent, noise_pred

    def get_timesteps_forward(self):
        return self.scheduler_fwd.timesteps

    def get_timesteps_backward(self):
        return self.scheduler_bwd.timesteps

    # ---- loops -----------------------------------------------------------------------------------
    def diffusion_forward(self, latent, context, guidance_scale_fwd: Optional[float] = None) -> Dict[str, Any]:
        guidance_scale_fwd = guidance_scale_fwd or self.guidance_scale_fwd
        latents, noise_preds = [latent], []
        if isinstance(latent, torch.Tensor):
            latent = latent.clone().detach()
        for i, t in enumerate(self.pbar(self.get_timesteps_forward(), desc="forward")):
            latent, noise_pred = self.predict_step_forward(latent, t, context, guidance_scale_fwd)
            noise_preds.append(noise_pred)
            latents.append(latent)
        return {"latents": latents, "noise_preds": noise_preds, "zT_inv": latents[-1]}

    def diffusion_backward(self, latent, context, inv_result: Dict[str, Any]) -> torch.Tensor:
        for i, t in enumerate(self.pbar(self.get_timesteps_backward(), desc="backward")):
            latent, noise_pred = self.predict_step_backward(latent, t, context)
        return latent

    def invert(self, image, prompt: Optional[str] = None, context: Optional[torch.Tensor] = None,
               guidance_scale_fwd: Optional[float] = None, **kwargs) -> Dict[str, Any]:
        context = context if context is not None else self.create_context(prompt)
        latent = self.encode(image)
        fwd_result = self.diffusion_forward(latent, context, guidance_scale_fwd=guidance_scale_fwd)
        fwd_result["context"] = context
        return {**kwargs, **fwd_result}

    def cat_context(self, contexts: List[torch.Tensor]) -> torch.Tensor:
        n, b = len(contexts), contexts[0].shape[0]
        assert b == 2, "Cfg should have batch dimension 2"
        x = torch.stack(contexts, 1)
        return x.reshape(b * n, *x.shape[2:]).contiguous()

    def cat_latent(self, latents: List[torch.Tensor]) -> torch.Tensor:
        return torch.cat(latents)

    def sample(self, inv_result: Dict[str, Any], prompt=None, context=None) -> Dict[str, Any]:
        if inv_result is None:
            return None
        latent = inv_result["latents"][-1]
        context = context if context is not None else self.create_context(prompt)
        if isinstance(context, list):
            num_prompts = len(context)
            context = self.cat_context(context)
            latent = self.cat_latent([latent] * num_prompts)
        z0 = self.diffusion_backward(latent, context, inv_result)
        if z0 is None:
            return None
        return {"image": self.decode(z0), "latent": z0}

    def invert_sample(self, image, prompt: str) -> Dict[str, Any]:
        context = self.create_context(prompt)
        return self.sample(self.invert(image, context=context), context=context)
