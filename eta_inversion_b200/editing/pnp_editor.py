"""Plug-and-Play editor (reference: modules/editing/pnp_editor.py:12-71)."""
from __future__ import annotations

import contextlib
from typing import Any, Dict, Iterator, Optional

from ..utils.pnp import PnPForward
from .editor import Editor, _pair_result


class PlugAndPlayEditor(Editor):
    def __init__(self, inverter, no_null_source_prompt: bool = True) -> None:
        self.inverter = inverter
        self.model = inverter.model
        self.negative_prompt = "ugly, blurry, black, low res, unrealistic"
        self.no_null_source_prompt = no_null_source_prompt

    @contextlib.contextmanager
    def register_editor(self) -> Iterator[None]:
        self.inverter.unet_wrapper = PnPForward(self.model)
        try:
            yield
        finally:
            self.inverter.unet_wrapper = None

    def edit(self, image, source_prompt: str, target_prompt: str, cfg: Optional[Dict[str, Any]] = None,
             inv_cfg=None) -> Dict[str, Any]:
        assert cfg is None
        inv_cfg = {} if inv_cfg is None else inv_cfg
        src_context = self.inverter.create_context("" if not self.no_null_source_prompt else source_prompt)
        target_context = self.inverter.create_context(target_prompt)
        inv_res = self.inverter.invert(image, prompt=source_prompt, context=src_context, inv_cfg=inv_cfg)
        with self.register_editor():
            if self.negative_prompt is not None and self.negative_prompt != "":
                target_context = self.inverter.create_context(target_prompt, negative_prompt=self.negative_prompt)
            edit_res = self.inverter.sample(inv_res, context=[src_context, target_context])
        return None if edit_res is None else _pair_result(edit_res)
