"""Eta inversion (reference: modules/inversion/eta_inversion.py:36-403; ECCV 2024 "Eta Inversion").

Loop A (inversion): DDIM inversion at CFG scale `guidance_scale_fwd` while the cross-attention maps of the source
prompt are accumulated (ControllerAttentionStorePerStep) -> per-word 64x64 maps, averaged over steps.
Loop B (edit): DDIM denoising with eta(t) > 0 only inside the thresholded map of the edited word; the injected
variance noise is the best of `noise_sample_count` candidates (closest to the noise that would reproduce the
inversion trajectory); the source row is pinned to the stored trajectory.

Per edit step the device work is: 1 UNet forward (B=4) + `etai_eta_noise_losses` + `etai_cfg_ddim_step`
(CFG, eta map, on-device argmin pick, noise injection, source pin) -- no host sync (the reference does
``argmin().item()`` per step, eta_inversion.py:363).  Noise candidates are pre-generated for the whole loop from the
same seeded CPU generator stream the reference consumes on CPU (RNG quirk).
"""
from __future__ import annotations

from typing import Any, Dict, List, Optional, Tuple, Union

import numpy as np
import torch
import torch.nn.functional as F

from .. import engine as E
from ..editing.ptp_editor import PromptToPromptControllerAttentionStore
from .diffusion_inversion import DiffusionInversion


class ControllerAttentionStorePerStep(PromptToPromptControllerAttentionStore):
    def __init__(self, model, prompt, res, from_where, callback) -> None:
        super().__init__(model, max_size=res)
        self.callback, self.prompt, self.res, self.from_where = callback, prompt, res, from_where

    def end_step(self, latent, noise_pred=None, t=None):
        words = self.prompt.split(" ")
        # one token per word; a repeated word resolves to its FIRST occurrence (ptp_editor.py:72, list.index)
        idx = [words.index(w) + 1 for w in words]
        maps = self.get_attention_maps(idx, res=self.res, from_where=self.from_where, resize=64)
        self.callback(maps, t)  # [words,1,64,64]: maps[w] is the reference's per-word [1,64,64] entry (no per-word slicing ops)
        return super().end_step(latent, noise_pred, t)


def make_eta_schedule(eta, num_train_steps: int = 1000) -> np.ndarray:
    """etas[t] for t in [0,1000): scalar | (start,end) linear | ((x1,y1),(x2,y2)[,p]) clipped power law over t/1000
    (eta_inversion.py:52-58,115-137); no eval()."""
    if not isinstance(eta, (tuple, list)):
        eta = eta, eta
    if len(eta) == 3 or isinstance(eta[0], (tuple, list)):
        (x1, y1), (x2, y2) = eta[0], eta[1]
        p = eta[2] if len(eta) == 3 else 1
        a = (y2 - y1) / (x2 - x1) ** p
        ts = np.linspace(0, 1, num_train_steps)
        etas = a * (np.clip(ts, x1, x2) - x1) ** p + y1
    else:
        etas = np.linspace(eta[0], eta[1], num_train_steps)
    return np.clip(etas, 0, None)


class EtaInversion(DiffusionInversion):
    def __init__(self, model, scheduler: Optional[str] = None, num_inference_steps: Optional[int] = None,
                 guidance_scale_bwd: Optional[float] = None, guidance_scale_fwd: Optional[float] = None,
                 verbose: bool = False, eta=(0.0, 0.4), noise_sample_count: int = 10, seed: int = 0,
                 eta_start: Optional[float] = None, eta_end: Optional[float] = None, use_mask=True,
                 mask_mode_cfg=None, noise_device: Optional[str] = None) -> None:
        """Same keywords as the reference (eta_inversion.py:26-37) plus ``noise_device``: where the seeded generator of the
        candidate noise lives.  None = the model's device, which is what the reference does
        (``torch.Generator(device=self.model.device)``, eta_inversion.py:276); "cpu" = a host generator, whose stream does
        not depend on the device (the golden fixtures were written by the reference running on the CPU)."""
        if use_mask:
            dft = dict(attn_from_where=["up", "down"], attn_res=16, mask_dirinv=None, mask_eta="fwd_mean", pow=None,
                       target_dirinv=None, thres=0.2)
            mask_mode_cfg = {**dft, **(mask_mode_cfg or {})}
        else:
            mask_mode_cfg = None
        self.mask_mode_cfg = mask_mode_cfg
        if isinstance(guidance_scale_fwd, (tuple, list)):
            assert len(guidance_scale_fwd) == 2
            guidance_scale_fwd = np.linspace(guidance_scale_fwd[0], guidance_scale_fwd[1], 1000)
        super().__init__(model, scheduler, num_inference_steps, guidance_scale_bwd, guidance_scale_fwd, verbose)
        if eta_start is not None:
            assert eta_end is not None
            eta = (eta_start, eta_end)
        self.etas = make_eta_schedule(eta)
        self.attn_maps_forward: Dict[Any, Any] = {}
        self.noise_sample_count = noise_sample_count
        self.seed = seed if seed >= 0 else None
        self.noise_device = noise_device
        self.noise_provider = None  # optional callable(step_index) -> [K,1,4,64,64]; default = seeded generator
        self.picks: list = []       # picked candidate index of every denoise step of the last loop (device tensors)
        self._loop_cache = None     # per-loop cache of step-independent tensors (set by diffusion_backward)

    # ---- noise -----------------------------------------------------------------------------------
    def sample_variance_noise(self, n: int, generator: Optional[torch.Generator] = None) -> torch.Tensor:
        hw = self.unet.latent_hw
        dev = generator.device if generator is not None else self.model.device
        return torch.randn((n, 1, 4, hw, hw), generator=generator, device=dev).to(self.model.device)

    def _noise_for_loop(self, steps: int) -> torch.Tensor:
        if self.noise_provider is not None:
            return torch.stack([self.noise_provider(i) for i in range(steps)]).to(self.model.device, torch.float32)
        hw = self.unet.latent_hw
        dev = torch.device(self.noise_device) if self.noise_device is not None else self.model.device
        g = torch.Generator(device=dev)
        if self.seed is not None:
            g.manual_seed(self.seed)
        # One randn per step, all drawn before the loop (still no host<->device sync inside it): exactly the reference's
        # per-step ``torch.randn((K,1,4,64,64), generator=g)`` stream (eta_inversion.py:156,276,348) on either kind of
        # generator.  (One big draw would only equal it for a CPU generator; the CUDA Philox element mapping depends on
        # the tensor size.)
        table = torch.stack([torch.randn((self.noise_sample_count, 1, 4, hw, hw), generator=g, device=dev)
                             for _ in range(steps)])
        return table.to(self.model.device)

    # ---- API parity: the reference's noise-selection methods as thin wrappers over the kernels ----------------------
    def _step_coef(self, t):
        sch = self.scheduler_bwd
        ti, pi = int(t), sch.prev_timestep(t)
        return sch.alpha(ti), sch.alpha(pi), float(sch._get_variance(ti, pi))

    def compute_optimal_variance_noise(self, latent_prev, latent, t, eta: float, noise_pred) -> torch.Tensor:
        """z* = (x_prev^inv - step(eps, eta, z=0)) / (eta * sqrt(var))  (eta_inversion.py:296-317).  ``noise_pred`` is the
        CFG-combined prediction, like in the reference."""
        a_t, a_p, var = self._step_coef(t)
        x, e = latent.float().contiguous(), noise_pred.float().contiguous()
        rec = E.cfg_ddim_step(e, x, a_t, a_p, None, float(eta), var)  # no candidates: z = 0
        return (latent_prev.float() - rec) / (float(eta) * var ** 0.5)

    def get_eta_variance_noise(self, latent_prev, latent, t, noise_pred, generator: Optional[torch.Generator] = None
                               ) -> Dict[str, Any]:
        """Best-of-K variance noise for the source row (eta_inversion.py:330-375); same result keys.  The edit loop itself
        does NOT call this (it keeps the pick on the device inside ``etai_cfg_ddim_step``); this wrapper reads the picked
        index back, like the reference's ``argmin().item()``."""
        eta = float(self.etas[int(t)])
        cand = self.sample_variance_noise(self.noise_sample_count, generator).float()
        a_t, a_p, var = self._step_coef(t)
        x, e = latent.float().contiguous(), noise_pred.float().contiguous()
        prev = latent_prev.float().contiguous()
        flat = cand.reshape(cand.shape[0], -1).contiguous()
        losses, best = E.eta_noise_losses(e, x, prev, a_t, a_p, None, eta, var, flat)
        best_idx = int(best.item())
        variance_noise = cand[best_idx]
        rec = E.cfg_ddim_step(e, x, a_t, a_p, None, eta, var, None, flat[best_idx:best_idx + 1].contiguous(), None)
        return {"eta": eta, "variance_noise": variance_noise, "delta": prev - rec, "latent_prev": latent_prev,
                "latent_prev_rec": rec, "loss": losses[best_idx]}

    # ---- masks -----------------------------------------------------------------------------------
    def get_mask(self, key, mask, t, edit_word_idx):
        if self.mask_mode_cfg is None:
            return None
        res, from_where, mode = self.mask_mode_cfg["attn_res"], self.mask_mode_cfg["attn_from_where"], self.mask_mode_cfg[key]
        if mode == "gt":
            pass
        elif mode == "fwd":
            mask = self.attn_maps_forward[t.item()][edit_word_idx[0]]
        elif mode == "fwd_mean":
            mask = self.attn_maps_forward["mean"][edit_word_idx[0]]
        elif mode in ("bwd_source", "bwd_target", "bwd_source_target"):
            ms = self.controller.get_attention_map(mask_idx=edit_word_idx[0], res=res, from_where=from_where, prompt_idx=0, num_prompts=2, resize=64)
            mt = self.controller.get_attention_map(mask_idx=edit_word_idx[1], res=res, from_where=from_where, prompt_idx=1, num_prompts=2, resize=64)
            mask = ms if mode == "bwd_source" else mt if mode == "bwd_target" else torch.maximum(ms, mt)
        elif mode is None:
            return None
        else:
            assert False
        if self.mask_mode_cfg["thres"] is not None:
            mask = (mask > self.mask_mode_cfg["thres"]).to(mask.dtype)
        if self.mask_mode_cfg["pow"] is not None:
            mask = torch.pow(mask, self.mask_mode_cfg["pow"])
        return mask

    # ---- UNet call ------------------------------------------------------------------------------------------------
    # The reference always runs the full [uncond, cond] batch here (eta_inversion.py:319-328), also in the inversion loop
    # where guidance_scale_fwd is 1 by default and the combine ``u + 1 * (c - u)`` gives the unconditional half weight 0:
    # a third of an edit's UNet rows (50 x 1 of 50 x 2 + 50 x 4) is computed and discarded.  With ``skip_zero_weight_uncond``
    # those rows are not computed (same result up to the rounding of ``u + (c - u)`` vs ``c``, ~1 ulp; the attention store
    # only ever sees the conditional half, ptp.py:112-113).  Set it to False for the reference's exact row count.
    skip_zero_weight_uncond = True

    def _unet_eps(self, latent, t, context, guidance_scale, is_fwd: bool = False):
        if is_fwd:
            guidance_scale = self.guidance_scale_fwd
        if isinstance(guidance_scale, (tuple, list, dict, np.ndarray)):
            guidance_scale = guidance_scale[t.item()]
        if (is_fwd and self.skip_zero_weight_uncond and float(guidance_scale) == 1.0
                and context.shape[0] == 2 * latent.shape[0]):
            eps = self._forward_unet(latent, t, context, zero_weight_uncond=True)
            if eps.shape[0] == latent.shape[0]:
                return eps, None  # conditional rows only: final
            return eps, 1.0
        latent_input = torch.cat([latent] * 2) if latent.shape[0] != context.shape[0] else latent
        return self._forward_unet(latent_input, t, context), float(guidance_scale)

    # ---- loop B step -----------------------------------------------------------------------------
    def predict_step_backward(self, latent, t, context, guidance_scale_bwd: Optional[float] = None,
                              source_latent_prev=None, generator=None, mask=None, edit_word_idx=None,
                              noise_cand: Optional[torch.Tensor] = None):
        guidance_scale_bwd = guidance_scale_bwd or self.guidance_scale_bwd
        latent = self.controller.begin_step(latent=latent, t=t)
        latent = latent.float().contiguous()
        eps_raw, g = self._unet_eps(latent, t, context, guidance_scale_bwd)
        if noise_cand is None:
            noise_cand = self.sample_variance_noise(self.noise_sample_count, generator)
        cand = noise_cand.reshape(noise_cand.shape[0], -1).float().contiguous()
        sch = self.scheduler_bwd
        ti, pi = int(t), sch.prev_timestep(t)
        a_t, a_p, var = sch.alpha(ti), sch.alpha(pi), float(sch._get_variance(ti, pi))
        eta = float(self.etas[ti])
        src_prev = source_latent_prev.float().contiguous()
        n = latent.shape[0]
        losses = None
        if eta > 0 and cand.shape[0] > 1:
            losses, best = E.eta_noise_losses(eps_raw, latent, src_prev, a_t, a_p, g, eta, var, cand)
            self.picks.append(best)  # device int32 [1] per step, never read inside the loop (tests / diagnostics)
        else:
            # eta == 0 (or one candidate): every candidate reproduces the same latent, the reference's argmin over equal
            # losses returns 0 (eta_inversion.py:363)
            self.picks.append(torch.zeros(1, dtype=torch.int32))
        eta_map = None
        delta_mask = None
        if self.mask_mode_cfg is not None:
            # "gt" / "fwd_mean" masks do not depend on the step: thresholded and broadcast once per loop, not once per step
            static = self.mask_mode_cfg["mask_eta"] in ("gt", "fwd_mean") and self._loop_cache is not None
            if static and "eta_map" in self._loop_cache:
                eta_map = self._loop_cache["eta_map"]
            else:
                m = self.get_mask("mask_eta", mask, t, edit_word_idx)
                if m is not None:
                    eta_map = torch.broadcast_to(m.float(), (1,) + tuple(latent.shape[1:])).contiguous()
                if static:
                    self._loop_cache["eta_map"] = eta_map
            delta_mask = self.get_mask("mask_dirinv", mask, t, edit_word_idx)
        leak = self.mask_mode_cfg["target_dirinv"] if self.mask_mode_cfg is not None else None
        if leak is None:
            new_latent, noise_pred = E.cfg_ddim_step(eps_raw, latent, a_t, a_p, g, eta, var, eta_map, cand, losses,
                                                     pin_src=src_prev, want_eps=True)
        else:  # rarely used variant (eta_inversion.py:251-256): leak the source delta into the target rows
            new_latent, noise_pred = E.cfg_ddim_step(eps_raw, latent, a_t, a_p, g, eta, var, eta_map, cand, losses,
                                                     want_eps=True)
            delta = src_prev[:1] - new_latent[:1]
            new_latent[:1] = new_latent[:1] + delta
            if delta_mask is not None:
                delta = (1 - delta_mask) * delta
            new_latent[1:] = new_latent[1:] + leak * delta
        new_latent = self.controller.end_step(latent=new_latent, noise_pred=noise_pred, t=t)
        return new_latent, noise_pred

    def diffusion_backward(self, latent, context, inv_result: Dict[str, Any]) -> torch.Tensor:
        inv_cfg = inv_result["inv_cfg"] or {}
        mask = inv_cfg.get("mask", None)
        edit_word_idx = inv_cfg.get("edit_word_idx", None)
        if mask is not None:
            mask = F.interpolate(mask[None, None].float(), (64, 64), mode="bilinear")[0].to(self.model.device)
        steps = self.scheduler_bwd.timesteps
        noise = self._noise_for_loop(len(steps))
        self.picks = []
        self._loop_cache = {}
        try:
            for i, t in enumerate(self.pbar(steps, desc="backward")):
                latent, noise_pred = self.predict_step_backward(
                    latent, t, context, source_latent_prev=inv_result["latents"][-(i + 2)], mask=mask,
                    edit_word_idx=edit_word_idx, noise_cand=noise[i])
        finally:
            self._loop_cache = None
        return latent

    # ---- loop A ----------------------------------------------------------------------------------
    def invert(self, image, prompt: Optional[str] = None, context: Optional[torch.Tensor] = None,
               guidance_scale_fwd: Optional[float] = None, inv_cfg: Optional[Dict[str, Any]] = None) -> Dict[str, Any]:
        if self.mask_mode_cfg is None:
            return super().invert(image, prompt, context, guidance_scale_fwd, inv_cfg=inv_cfg)
        if inv_cfg["edit_word_idx"][0] is None or inv_cfg["edit_word_idx"][1] is None:
            return None
        self.attn_maps_forward = {}
        store = ControllerAttentionStorePerStep(
            self.model, prompt, res=self.mask_mode_cfg["attn_res"], from_where=self.mask_mode_cfg["attn_from_where"],
            callback=(lambda attn, t: self.attn_maps_forward.update({t.item(): attn})))
        with self.use_controller(store):
            fwd_result = super().invert(image, prompt, context, guidance_scale_fwd, inv_cfg=inv_cfg)
        per_step = list(self.attn_maps_forward.values())
        # mean over the steps for every word at once ([words,1,64,64]; indexing [w] gives the reference's per-word list entry)
        self.attn_maps_forward["mean"] = torch.mean(torch.stack(per_step), dim=0)
        return fwd_result
