"""Editor base classes (reference: modules/editing/editor.py:6-135). Orchestration only; no arithmetic."""
from __future__ import annotations

from typing import Any, Callable, Dict, Optional

import torch

from .controller import ControllerBase


class Editor:
    def edit(self, image: torch.Tensor, source_prompt: str, target_prompt: str, cfg: Optional[Dict[str, Any]] = None,
             **kwargs) -> Dict[str, Any]:
        raise NotImplementedError


def _pair_result(edit_res):
    return {"image_inv": edit_res["image"][0:1], "image": edit_res["image"][1:2],
            "latent_inv": edit_res["latent"][0:1], "latent": edit_res["latent"][1:2]}


class ControllerBasedEditor(Editor):
    def __init__(self, inverter, no_source_backward: bool = False, dft_cfg: Optional[Dict[Any, str]] = None,
                 fake_edit: bool = False) -> None:
        self.inverter = inverter
        self.no_source_backward = no_source_backward
        self.dft_cfg = dft_cfg if dft_cfg is not None else {}
        self.fake_edit = fake_edit

    def make_controller(self, image, source_prompt: str, target_prompt: str, inv_res, **kwargs) -> ControllerBase:
        raise NotImplementedError

    def edit(self, image, source_prompt: str, target_prompt: str, cfg: Optional[Dict[str, Any]] = None, inv_cfg=None,
             **kwargs) -> Dict[str, Any]:
        cfg = {**self.dft_cfg} if cfg is None else cfg
        inv_cfg = {} if inv_cfg is None else inv_cfg
        src_context = self.inverter.create_context(source_prompt)
        target_context = self.inverter.create_context(target_prompt)
        zT_gt = cfg.pop("zT_gt", None)
        if self.fake_edit:
            image = None
            inv_res = {"latents": [zT_gt.to(self.inverter.model.device)]}
        else:
            inv_res = self.inverter.invert(image, prompt=source_prompt, context=src_context, inv_cfg=inv_cfg)
        controller = self.make_controller(image=image, source_prompt=source_prompt, target_prompt=target_prompt,
                                          inv_res=inv_res, **cfg, **kwargs)
        with self.inverter.use_controller(controller):
            if not self.no_source_backward:
                edit_res = self.inverter.sample(inv_res, context=[src_context, target_context])
                return None if edit_res is None else _pair_result(edit_res)
            edit_res = self.inverter.sample(inv_res, context=[target_context])
            return {"image": edit_res["image"], "latent": edit_res["latent"]}


class ControllerBasedEditorLambda(ControllerBasedEditor):
    def __init__(self, inverter, controller_cls: Optional[Callable] = None, no_source_backward: bool = False, **kwargs):
        super().__init__(inverter, no_source_backward=no_source_backward)
        self.controller_cls = controller_cls
        self.controller_kwargs = kwargs

    def make_controller(self, image, source_prompt: str, target_prompt: str, **kwargs) -> ControllerBase:
        return self.controller_cls(editor=self, image=image, source_prompt=source_prompt, target_prompt=target_prompt,
                                   **kwargs, **self.controller_kwargs)
