"""MasaCtrl editor (reference: modules/editing/masactrl_editor.py:12-69)."""
from __future__ import annotations

import contextlib
from typing import Any, Dict, Iterator, Optional

from ..utils.masactrl import MutualSelfAttentionControl
from .editor import Editor, _pair_result


class MasactrlEditor(Editor):
    def __init__(self, inverter, no_null_source_prompt: bool = True, step: int = 4, layer: int = 10) -> None:
        self.inverter = inverter
        self.model = inverter.model
        self.no_null_source_prompt = no_null_source_prompt
        self.step, self.layer = step, layer

    @contextlib.contextmanager
    def register_editor(self) -> Iterator[None]:
        editor = MutualSelfAttentionControl(self.step, self.layer)
        self.inverter.attn_hooks.append(editor)
        try:
            yield
        finally:
            self.inverter.attn_hooks.remove(editor)

    def edit(self, image, source_prompt: str, target_prompt: str, cfg: Optional[Dict[str, Any]] = None,
             inv_cfg=None) -> Dict[str, Any]:
        assert cfg is None, f"{cfg}"
        inv_cfg = {} if inv_cfg is None else inv_cfg
        src_context = self.inverter.create_context("" if not self.no_null_source_prompt else source_prompt)
        target_context = self.inverter.create_context(target_prompt)
        inv_res = self.inverter.invert(image, context=src_context, prompt=source_prompt, inv_cfg=inv_cfg)
        with self.register_editor():
            edit_res = self.inverter.sample(inv_res, context=[src_context, target_context])
        return None if edit_res is None else _pair_result(edit_res)
