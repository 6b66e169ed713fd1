"""Inversion-only "editor" (reference: modules/editing/inv_editor.py:8-53)."""
from __future__ import annotations

from typing import Any, Dict, Optional

from .editor import Editor


class InversionEditor(Editor):
    def __init__(self, inverter, no_source_backward: bool = False, vae_rec: bool = False,
                 no_null_source_prompt: bool = True) -> None:
        self.inverter = inverter
        self.model = inverter.model
        self.no_source_backward = no_source_backward
        self.vae_rec = vae_rec
        self.no_null_source_prompt = no_null_source_prompt

    def edit(self, image, source_prompt: str, target_prompt: str, cfg: Optional[Dict[str, Any]] = None) -> Dict[str, Any]:
        assert cfg is None
        if self.vae_rec:
            latent = self.inverter.encode(image)
            return {"image": self.inverter.decode(latent), "latent": latent}
        src_context = self.inverter.create_context(source_prompt if self.no_null_source_prompt else "")
        inv_res = self.inverter.invert(image, context=src_context)
        edit_res = self.inverter.sample(inv_res, context=[src_context])
        return {"image": edit_res["image"], "latent": edit_res["latent"]}
