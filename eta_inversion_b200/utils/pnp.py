"""Plug-and-Play feature / self-attention injection in data form (reference: modules/utils/pnp.py:86-150,
modules/utils/pnp_utils.py:44-195).  UNet rows are [uncond_src, uncond_tgt, cond_tgt] (B=3, taken from rows
[0,1,3] of the 4-row edit batch); while t is in the first 50% of the timesteps the self-attention q,k of rows 1,2
come from row 0 in up_blocks[1].attentions[1,2] and up_blocks[2,3].attentions[0..2]; while t is in the first 80% the
conv2 output of up_blocks[1].resnets[1] of rows 1,2 is overwritten by row 0."""
from __future__ import annotations

from typing import Optional

import torch

from ..engine import AttnControl

# transformer call-order indices: up1.a0-2 = 7,8,9 ; up2 = 10,11,12 ; up3 = 13,14,15
_QK_LAYER_MASK = sum(1 << i for i in (8, 9, 10, 11, 12, 13, 14, 15))


class PnPForward:
    """Replaces the monkey-patched ``unet.forward`` (pnp.py:101-150)."""

    def __init__(self, model, pnp_f_t: float = 0.8, pnp_attn_t: float = 0.5) -> None:
        n = model.scheduler.num_inference_steps
        ts = model.scheduler.timesteps
        self.qk_timesteps = set(int(t) for t in ts[: int(n * pnp_attn_t)])
        self.conv_timesteps = set(int(t) for t in ts[: int(n * pnp_f_t)])
        self._ctx_src = None
        self._ctx3 = None

    def control(self, t: int) -> Optional[AttnControl]:
        qk = t in self.qk_timesteps or t == 1000
        conv = t in self.conv_timesteps or t == 1000
        if not (qk or conv):
            return None
        c = AttnControl(conv_inject_rows=1 if conv else 0)
        if qk:
            c.self_rows = ([0, 0, 0], [0, 0, 0], [0, 1, 2])
            c.self_layer_mask = _QK_LAYER_MASK
        return c

    def __call__(self, unet, sample: torch.Tensor, timestep, encoder_hidden_states: torch.Tensor, control=None):
        assert sample.shape[0] == 4 and encoder_hidden_states.shape[0] == 4
        src = self._ctx_src
        if src is None or src[0] is not encoder_hidden_states or src[1] != encoder_hidden_states._version:
            # slice the context once per loop so the engine keeps its K/V projection cache
            self._ctx_src = (encoder_hidden_states, encoder_hidden_states._version)
            self._ctx3 = encoder_hidden_states[[0, 1, 3]].contiguous()
        eps3 = unet(sample[[0, 1, 3]].contiguous(), timestep, encoder_hidden_states=self._ctx3,
                    control=self.control(int(timestep)))["sample"]
        return eps3[[0, 1, 0, 2]]
