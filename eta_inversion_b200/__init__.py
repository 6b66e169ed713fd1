"""eta_inversion_b200 -- B200-native engine for the hot path of furiosa-ai/eta-inversion.

Public surface mirrors the reference's ``modules`` package (modules/__init__.py:31-111): the inverter / editor
registries with identical names and constructor keywords, ``load_inverter`` / ``load_editor`` /
``load_diffusion_model``.  The UNet, the scheduler arithmetic and all attention control run in libetai.so
(hand-written sm_100a CUDA behind the C ABI of include/etai.h); importing this package never touches the GPU.
"""
from __future__ import annotations

from functools import partial
from typing import Callable, List

from .editing.editor import Editor
from .editing.inv_editor import InversionEditor
from .editing.masactrl_editor import MasactrlEditor
from .editing.pnp_editor import PlugAndPlayEditor
from .editing.ptp_editor import PromptToPromptEditor
from .editing.simple_editor import SimpleEditor
from .inversion.diffusion_inversion import DiffusionInversion
from .inversion.direct_inversion import DirectInversion
from .inversion.ddpm_inversion import DDPMInversion
from .inversion.edict_inversion import EdictInversion
from .inversion.eta_inversion import EtaInversion
from .inversion.negative_prompt_inversion import NegativePromptInversion
from .inversion.null_text_inversion import NullTextInversion
from .inversion.proximal_negative_prompt_inversion import ProximalNegativePromptInversion
from .inversion.regularized_diffusion_inversion import RegularizedDiffusionInversion
from .models import StablePostProc, StablePreprocess, load_diffusion_model  # noqa: F401

__version__ = "0.1.0"


def _out_of_scope(name: str, why: str) -> Callable:
    def ctor(*a, **k):
        raise NotImplementedError(f"'{name}' is not built in this engine yet: {why} (SURVEY.md section 2 / 8f)")
    return ctor


_inverters = {
    "diffinv": DiffusionInversion,
    "npi": NegativePromptInversion,
    "dirinv": DirectInversion,
    "etainv": EtaInversion,
    "nti": NullTextInversion,
    "proxnpi": ProximalNegativePromptInversion,
    "edict": EdictInversion,
    "ddpminv": DDPMInversion,
    "cyclediff": partial(DDPMInversion, markovian_forward=True),
    "regdiffinv": RegularizedDiffusionInversion,
}

_editors = {
    "simple": SimpleEditor,
    "ptp": PromptToPromptEditor,
    "masactrl": MasactrlEditor,
    "pnp": PlugAndPlayEditor,
    "invedit": InversionEditor,
    "pix2pix_zero": _out_of_scope("pix2pix_zero", "needs BLIP captions and autograd through the UNet"),
}


def register_editor(name: str, editor_cls: Callable) -> None:
    print(f"Registering editor {name}")
    _editors[name] = editor_cls


def get_inversion_methods() -> List[str]:
    return list(_inverters.keys())


def get_edit_methods() -> List[str]:
    return list(_editors.keys())


def load_inverter(type: str, **kwargs) -> DiffusionInversion:
    return _inverters[type](**kwargs)


def load_editor(type: str, **kwargs) -> Editor:
    return _editors[type](**kwargs)
