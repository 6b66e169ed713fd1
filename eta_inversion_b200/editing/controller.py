"""Controller protocol of the diffusion loops (reference: modules/editing/controller.py:5-110).

A controller is told when a loop begins/ends and before/after every step; it may rewrite the latent and, in this
engine, install an ``AttnControl`` descriptor on the UNet for the step's forward (the reference monkey-patches
``Attention.forward`` at the same two hook points, ptp_editor.py:87-98)."""
from __future__ import annotations

from typing import Optional

import torch


class ControllerBase:
    def begin(self) -> None:
        pass

    def end(self) -> None:
        pass

    def begin_step(self, latent: torch.Tensor, *args, **kwargs) -> torch.Tensor:
        return latent

    def end_step(self, latent: torch.Tensor, noise_pred: Optional[torch.Tensor] = None, t=None) -> torch.Tensor:
        return latent

    def attn_control(self, unet, batch_rows: int):
        """Descriptor for the UNet forward of the current step (None = plain attention)."""
        return None

    def after_forward(self) -> None:
        """Called right after the step's UNet forward (the reference's hooks advance their counters there)."""
        pass

    def copy(self, **kwargs) -> "ControllerBase":
        raise NotImplementedError


class ControllerEmpty(ControllerBase):
    def copy(self, **kwargs) -> "ControllerEmpty":
        return self


class EdictController(ControllerBase):
    """One copy of the wrapped controller per latent of EDICT's coupled pair (modules/editing/controller.py:71-110); the
    copy that belongs to the latent being updated receives the step callbacks and supplies the attention control."""

    def __init__(self, controller: ControllerBase) -> None:
        self.controllers = [controller.copy(latent_idx=i) for i in range(2)]
        self.cur_latent_idx = None

    def begin(self) -> None:
        self.cur_latent_idx = None
        for c in self.controllers:
            c.begin()

    def end(self) -> None:
        for c in self.controllers:
            c.end()

    def begin_step(self, latent_idx: int, latent_base: torch.Tensor, latent_model_input: torch.Tensor) -> None:
        self.cur_latent_idx = latent_idx
        self.controllers[latent_idx].begin_step(latent_base)

    def end_step(self, latent: torch.Tensor, **kwargs) -> torch.Tensor:
        return self.controllers[self.cur_latent_idx].end_step(latent=latent, **kwargs)

    def attn_control(self, unet, batch_rows: int):
        # the inversion pass runs without step callbacks (no controller is current): plain attention, like the reference,
        # whose hooks only see the forwards of the latent selected in begin_step
        return None if self.cur_latent_idx is None else self.controllers[self.cur_latent_idx].attn_control(unet, batch_rows)

    def after_forward(self) -> None:
        if self.cur_latent_idx is not None:
            self.controllers[self.cur_latent_idx].after_forward()
