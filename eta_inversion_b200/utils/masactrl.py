"""MasaCtrl mutual self-attention in data form (reference: modules/utils/masactrl.py:14-72,
modules/utils/masactrl_utils.py:9-35).  Inside the step/layer window every row's queries attend to the K,V of the
source row of its classifier-free-guidance half; expressed as a (q,k,v) row remap for the fused attention kernel,
so the second attention pass of the reference (masactrl.py:41-54) and its materialised sim/attn disappear."""
from __future__ import annotations

from typing import List, Optional

from ..engine import AttnControl


class AttentionBase:
    def __init__(self) -> None:
        self.cur_step = 0
        self.num_att_layers = 32
        self.cur_att_layer = 0

    def after_step(self) -> None:
        pass

    def reset(self):
        self.cur_step = 0
        self.cur_att_layer = 0

    def begin_forward(self, unet, batch_rows: int) -> Optional[AttnControl]:
        return None

    def end_forward(self) -> None:
        self.cur_step += 1
        self.after_step()


class MutualSelfAttentionControl(AttentionBase):
    MODEL_TYPE = {"SD": 16, "SDXL": 70}

    def __init__(self, start_step: int = 4, start_layer: int = 10, layer_idx: Optional[List[int]] = None,
                 step_idx: Optional[List[int]] = None, total_steps: int = 50, model_type: str = "SD") -> None:
        super().__init__()
        self.total_steps = total_steps
        self.total_layers = self.MODEL_TYPE.get(model_type, 16)
        self.start_step, self.start_layer = start_step, start_layer
        self.layer_idx = layer_idx if layer_idx is not None else list(range(start_layer, self.total_layers))
        self.step_idx = step_idx if step_idx is not None else list(range(start_step, total_steps))
        print("MasaCtrl at denoising steps: ", self.step_idx)
        print("MasaCtrl at U-Net layers: ", self.layer_idx)

    def begin_forward(self, unet, batch_rows: int) -> Optional[AttnControl]:
        if self.cur_step not in self.step_idx:
            return None
        half = batch_rows // 2
        rows = list(range(batch_rows))
        src = [0 if r < half else half for r in rows]  # ku[:num_heads] / kc[:num_heads]: first row of each half
        mask = 0
        for i in self.layer_idx:
            if 0 <= i < 16:
                mask |= 1 << i
        return AttnControl(self_rows=(rows, src, list(src)), self_layer_mask=mask, self_max_tokens=1 << 30)
