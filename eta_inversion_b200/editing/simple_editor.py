"""Plain re-denoising with the target prompt (reference: modules/editing/simple_editor.py:8-51)."""
from __future__ import annotations

from typing import Any, Dict, Optional

from .editor import Editor, _pair_result


class SimpleEditor(Editor):
    def __init__(self, inverter, no_source_backward: bool = False) -> None:
        self.inverter = inverter
        self.model = inverter.model
        self.no_source_backward = no_source_backward

    def edit(self, image, source_prompt: str, target_prompt: str, cfg: Optional[Dict[str, Any]] = None,
             inv_cfg=None) -> Dict[str, Any]:
        assert cfg is None
        src_context = self.inverter.create_context(source_prompt)
        target_context = self.inverter.create_context(target_prompt)
        inv_res = self.inverter.invert(image, prompt=source_prompt, context=src_context, guidance_scale_fwd=1, inv_cfg=inv_cfg)
        if not self.no_source_backward:
            return _pair_result(self.inverter.sample(inv_res, context=[src_context, target_context]))
        edit_res = self.inverter.sample(inv_res, context=[target_context])
        return {"image": edit_res["image"], "latent": edit_res["latent"]}
