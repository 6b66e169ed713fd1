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
