"""CPU checks of the drop-in boundary: the C-ABI library loads, exports every symbol include/etai.h declares, and fails
loudly (error code + message, no crash, no fallback) when there is no CUDA device."""
import ctypes
import re

import pytest
import torch

from conftest import ROOT


def _declared():
    text = (ROOT / "include" / "etai.h").read_text()
    return sorted(set(re.findall(r"ETAI_EXPORT\s+[\w\s\*]+?\b(etai_\w+)\s*\(", text)))


def test_header_symbols_are_exported_and_bound():
    from eta_inversion_b200 import _lib
    lib = ctypes.CDLL(str(_lib.lib_path()))
    declared = _declared()
    assert len(declared) >= 16
    for name in declared:
        assert hasattr(lib, name), f"{name} declared in include/etai.h but not exported by libetai.so"
        assert name in _lib.SYMBOLS, f"{name} has no ctypes signature in eta_inversion_b200/_lib.py"
    assert set(_lib.SYMBOLS) == set(declared)


def test_abi_version_and_struct_sizes():
    from eta_inversion_b200 import _lib
    lib = _lib.load()
    assert lib.etai_abi_version() == _lib.ABI_VERSION
    assert ctypes.sizeof(_lib.EtaiTensor) == 64 and ctypes.sizeof(_lib.EtaiUnetCfg) == 44


@pytest.mark.skipif(torch.cuda.is_available(), reason="checks the no-GPU failure path")
def test_no_gpu_is_a_loud_error_not_a_fallback():
    from eta_inversion_b200 import _lib, engine
    lib = _lib.load()
    cfg = _lib.EtaiUnetCfg()
    cfg.dtype, cfg.heads, cfg.cross_dim, cfg.ctx_len, cfg.latent_hw, cfg.max_batch = 0, 8, 768, 77, 64, 4
    for i, c in enumerate((320, 640, 1280, 1280)):
        cfg.block_out_channels[i] = c
    t = _lib.EtaiTensor()
    t.name = b"conv_in.weight"
    h = ctypes.c_void_p()
    rc = lib.etai_unet_create(ctypes.byref(h), ctypes.byref(cfg), ctypes.pointer(t), 1, 0)
    assert rc < 0 and not h.value
    assert len(lib.etai_last_error()) > 0
    with pytest.raises(RuntimeError, match="CUDA"):
        engine.UNetEngine({}, dtype=torch.float32)
    with pytest.raises(RuntimeError):
        engine.gemm(torch.zeros(4, 64), torch.zeros(4, 64))
    with pytest.raises(RuntimeError, match="CUDA"):  # the proximal CFG has no torch fallback either
        engine.prox_guidance(torch.zeros(1, 4, 8, 8), torch.zeros(1, 4, 8, 8), 7.5, 0.7)
    # argument errors come back as codes + text, never as a crash: null pointers, ranks outside the tensor
    assert lib.etai_prox_guidance(None, None, None, 16, 3, 4, 0.5, 0.0, 0, 7.5, None, None) < 0
    assert b"prox_guidance" in lib.etai_last_error()
    assert lib.etai_unet_forward_rows(None, None, None, 0, 2, None, None, None) < 0
    import eta_inversion_b200 as etai
    with pytest.raises(RuntimeError):
        etai.load_diffusion_model("synthetic-sd15", "cpu")


def test_registry_surface_matches_reference():
    """Same registry names as modules/__init__.py:31-54 of the reference."""
    import eta_inversion_b200 as etai
    assert etai.get_inversion_methods() == ["diffinv", "npi", "dirinv", "etainv", "nti", "proxnpi", "edict", "ddpminv",
                                            "cyclediff", "regdiffinv"] or set(etai.get_inversion_methods()) == {
        "diffinv", "nti", "npi", "proxnpi", "edict", "ddpminv", "cyclediff", "dirinv", "etainv", "regdiffinv"}
    assert set(etai.get_edit_methods()) == {"simple", "ptp", "masactrl", "pnp", "pix2pix_zero", "invedit"}
    with pytest.raises(NotImplementedError):
        etai.load_editor(type="pix2pix_zero", inverter=None)
    from eta_inversion_b200.inversion.null_text_inversion import NullTextInversion
    assert etai._inverters["nti"] is NullTextInversion
