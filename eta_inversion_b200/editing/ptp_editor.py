"""Prompt-to-prompt editor and controllers (reference: modules/editing/ptp_editor.py:14-159)."""
from __future__ import annotations

from typing import Any, Dict, List, Optional

import torch
import torch.nn.functional as F

from ..utils import ptp
from .controller import ControllerBase
from .editor import ControllerBasedEditor


_WORD_IDX_CACHE: Dict[Any, torch.Tensor] = {}  # (token indices, device) -> index tensor on the device


class PromptToPromptControllerBase(ControllerBase):
    """Wraps a ptp attention controller; installs its per-step ``AttnControl`` instead of patching 32 modules."""

    def __init__(self, model, controller: ptp.AttentionControl) -> None:
        self.model = model
        self.controller = controller
        self.step_idx = None

    def begin(self) -> None:
        self.step_idx = 0

    def end(self) -> None:
        pass

    def attn_control(self, unet, batch_rows: int):
        return self.controller.begin_forward(unet, batch_rows)

    def after_forward(self) -> None:
        self.controller.end_forward()

    def get_attention_maps(self, word_indices: List[int], res: int = 16, from_where=("up", "down"),
                           resize: Optional[int] = None, prompt_idx: int = 0) -> torch.Tensor:
        """Batched form of get_attention_map: [len(words),1,R,R] maps, each max-normalised (and bicubic-resized)."""
        maps = ptp.aggregate_attention([None], self.controller, res, list(from_where), True, select=prompt_idx)
        # index tensor cached on the device: indexing with a Python list uploads it every call, and that pageable
        # host-to-device copy synchronises the host with the GPU once per inversion step
        key = (tuple(word_indices), maps.device)
        idx = _WORD_IDX_CACHE.get(key)
        if idx is None:
            idx = torch.tensor(list(word_indices), dtype=torch.long, device=maps.device)
            torch.cuda.current_stream(maps.device).synchronize()  # published to other threads / streams below
            if len(_WORD_IDX_CACHE) >= 256:
                _WORD_IDX_CACHE.clear()
            _WORD_IDX_CACHE[key] = idx
        m = maps.index_select(2, idx).permute(2, 0, 1)[:, None]  # [W,1,res,res]
        m = m / m.amax(dim=(1, 2, 3), keepdim=True)
        if resize is not None and m.shape[-2:] != (resize, resize):
            m = F.interpolate(m, (resize, resize), mode="bicubic").clamp(0, 1)
        return m

    def get_attention_map(self, prompt: str = None, word: str = None, res: int = 16, from_where=("up", "down"),
                          resize: Optional[int] = None, prompt_idx=None, num_prompts=None, mask_idx=None) -> torch.Tensor:
        if prompt_idx is None:
            assert num_prompts is None
            prompt_idx = 0
        else:
            assert num_prompts is not None
        if mask_idx is None:
            try:
                mask_idx = prompt.split(' ').index(word) + 1  # +1 for the start token; one token per word (App. D)
            except ValueError:
                raise Exception(f"Cannot get attention map. Word {word} not in {prompt}")
        else:
            mask_idx = mask_idx + 1
        return self.get_attention_maps([mask_idx], res, from_where, resize, prompt_idx)[0]

    def begin_step(self, latent: torch.Tensor, *args, **kwargs) -> torch.Tensor:
        return latent

    def end_step(self, latent: torch.Tensor, noise_pred: Optional[torch.Tensor] = None, t=None) -> torch.Tensor:
        latent = self.controller.step_callback(latent)
        self.step_idx += 1
        return latent


class PromptToPromptController(PromptToPromptControllerBase):
    def __init__(self, model, source_prompt: str, target_prompt: str, inv_res: Optional[Dict[str, Any]] = None,
                 **kwargs) -> None:
        self.source_prompt, self.target_prompt = source_prompt, target_prompt
        self.ptp_cfg = {**kwargs}
        if "prompts" in self.ptp_cfg:
            assert list(self.ptp_cfg["prompts"]) == [source_prompt, target_prompt]
            self.ptp_cfg.pop("prompts")
        super().__init__(model, ptp.make_controller(model, prompts=[source_prompt, target_prompt], **self.ptp_cfg))

    def copy(self, **kwargs) -> "PromptToPromptController":
        return PromptToPromptController(self.model, self.source_prompt, self.target_prompt, **self.ptp_cfg)


class PromptToPromptControllerAttentionStore(PromptToPromptControllerBase):
    def __init__(self, model, max_size=32) -> None:
        super().__init__(model, ptp.AttentionStore(max_size=max_size))


class PromptToPromptEditor(ControllerBasedEditor):
    def __init__(self, inverter, no_source_backward: bool = False, dft_cfg: Dict[Any, str] = None, **kwargs) -> None:
        super().__init__(inverter, no_source_backward, dft_cfg, **kwargs)

    def make_controller(self, image, source_prompt: str, target_prompt: str, **kwargs) -> PromptToPromptController:
        return PromptToPromptController(model=self.inverter.model, source_prompt=source_prompt,
                                        target_prompt=target_prompt, **kwargs)
