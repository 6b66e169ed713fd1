"""ctypes binding of libetai.so (the C ABI in include/etai.h).

There is no fallback: if the shared library is missing or a call fails, a RuntimeError is raised
(`etai_last_error()` text).  PyTorch only supplies device memory (``tensor.data_ptr()``) and the
current CUDA stream.
"""
from __future__ import annotations

import ctypes as C
import os
from pathlib import Path

import torch

ETAI_F32, ETAI_F16, ETAI_BF16 = 0, 1, 2
MATH_AUTO, MATH_SIMT = 0, 1
CTRL_SELF_REMAP, CTRL_CROSS_EDIT, CTRL_CROSS_STORE = 1, 2, 4
ABI_VERSION = 1

_DTYPES = {torch.float32: ETAI_F32, torch.float16: ETAI_F16, torch.bfloat16: ETAI_BF16}


def dtype_code(dt: torch.dtype) -> int:
    try:
        return _DTYPES[dt]
    except KeyError:
        raise RuntimeError(f"etai: unsupported dtype {dt}") from None


class EtaiTensor(C.Structure):
    _fields_ = [("name", C.c_char_p), ("data", C.c_void_p), ("dtype", C.c_int32), ("ndim", C.c_int32),
                ("shape", C.c_int64 * 4), ("on_device", C.c_int32)]


class EtaiUnetCfg(C.Structure):
    _fields_ = [("dtype", C.c_int32), ("math_mode", C.c_int32), ("block_out_channels", C.c_int32 * 4),
                ("heads", C.c_int32), ("cross_dim", C.c_int32), ("ctx_len", C.c_int32), ("latent_hw", C.c_int32),
                ("max_batch", C.c_int32)]


class EtaiVaeCfg(C.Structure):
    _fields_ = [("dtype", C.c_int32), ("math_mode", C.c_int32), ("block_out_channels", C.c_int32 * 4),
                ("image_hw", C.c_int32), ("max_batch", C.c_int32)]


class EtaiClipCfg(C.Structure):
    _fields_ = [("dtype", C.c_int32), ("math_mode", C.c_int32), ("vocab", C.c_int32), ("hidden", C.c_int32),
                ("layers", C.c_int32), ("heads", C.c_int32), ("ffn", C.c_int32), ("max_len", C.c_int32),
                ("max_batch", C.c_int32)]


class EtaiAttnCtrl(C.Structure):
    _fields_ = [("flags", C.c_int32),
                ("self_q_row", C.POINTER(C.c_int32)), ("self_k_row", C.POINTER(C.c_int32)),
                ("self_v_row", C.POINTER(C.c_int32)), ("self_layer_mask", C.c_uint32), ("self_max_tokens", C.c_int32),
                ("n_pairs", C.c_int32), ("edit_base_row", C.POINTER(C.c_int32)), ("edit_tgt_row", C.POINTER(C.c_int32)),
                ("mapper", C.c_void_p), ("blend_a", C.c_void_p), ("equalizer", C.c_void_p), ("alpha_step", C.c_void_p),
                ("store_res", C.c_int32), ("n_store_rows", C.c_int32), ("store_row", C.POINTER(C.c_int32)),
                ("store_down", C.c_void_p), ("store_mid", C.c_void_p), ("store_up", C.c_void_p),
                ("conv_inject_rows", C.c_int32)]


# every symbol include/etai.h declares: name -> (restype, argtypes)
_vp, _i32, _i64, _f = C.c_void_p, C.c_int32, C.c_int64, C.c_float
SYMBOLS = {
    "etai_abi_version": (C.c_int, []),
    "etai_last_error": (C.c_char_p, []),
    "etai_unet_create": (C.c_int, [C.POINTER(_vp), C.POINTER(EtaiUnetCfg), C.POINTER(EtaiTensor), _i32, _i32]),
    "etai_unet_destroy": (C.c_int, [_vp]),
    "etai_unet_clone": (C.c_int, [C.POINTER(_vp), _vp, _i32]),
    "etai_unet_set_context": (C.c_int, [_vp, _vp, _i32, _i32, _vp]),
    "etai_unet_forward": (C.c_int, [_vp, _vp, _f, _i32, _i32, C.POINTER(EtaiAttnCtrl), _vp, _vp]),
    "etai_unet_forward_rows": (C.c_int, [_vp, _vp, C.POINTER(C.c_float), _i32, _i32, C.POINTER(EtaiAttnCtrl), _vp, _vp]),
    "etai_unet_enable_backward": (C.c_int, [_vp, _i32]),
    "etai_unet_forward_train": (C.c_int, [_vp, _vp, _f, _i32, _i32, _vp, _vp]),
    "etai_unet_backward_ctx": (C.c_int, [_vp, _vp, _i32, _vp, _vp]),
    "etai_unet_device_bytes": (_i64, [_vp]),
    "etai_unet_launch_count": (_i64, [_vp]),
    "etai_unet_profile": (C.c_int, [_vp, _i32, C.POINTER(C.c_float), C.POINTER(_i32)]),
    "etai_vae_create": (C.c_int, [C.POINTER(_vp), C.POINTER(EtaiVaeCfg), C.POINTER(EtaiTensor), _i32, _i32]),
    "etai_vae_destroy": (C.c_int, [_vp]),
    "etai_vae_encode": (C.c_int, [_vp, _vp, _i32, _i32, _vp, _vp]),
    "etai_vae_decode": (C.c_int, [_vp, _vp, _i32, _i32, _vp, _vp]),
    "etai_vae_launch_count": (_i64, [_vp]),
    "etai_vae_device_bytes": (_i64, [_vp]),
    "etai_clip_create": (C.c_int, [C.POINTER(_vp), C.POINTER(EtaiClipCfg), C.POINTER(EtaiTensor), _i32, _i32]),
    "etai_clip_destroy": (C.c_int, [_vp]),
    "etai_clip_encode": (C.c_int, [_vp, C.POINTER(_i32), _i32, _vp, _i32, _vp]),
    "etai_clip_launch_count": (_i64, [_vp]),
    "etai_cfg_ddim_step": (C.c_int, [_vp, _i32, _i32, _f, _vp, _vp, _vp, _f, _f, _f, _f, _vp, _vp, _vp, _i32, _vp, _i64, _vp]),
    "etai_eta_noise_losses": (C.c_int, [_vp, _i32, _i32, _f, _vp, _vp, _f, _f, _f, _f, _vp, _i32, _i64, _vp, _vp, _vp]),
    "etai_prox_guidance": (C.c_int, [_vp, _vp, _vp, _i64, _i64, _i64, _f, _f, _i32, _f, _vp, _vp]),
    "etai_groupnorm": (C.c_int, [_vp, _vp, _vp, _vp, _i32, _i64, _i32, _i32, _f, _i32, _i32, _vp, _i64, _vp]),
    "etai_layernorm": (C.c_int, [_vp, _vp, _vp, _vp, _i64, _i32, _f, _i32, _vp]),
    "etai_gemm": (C.c_int, [_vp, _vp, _vp, _vp, _vp, _i64, _i32, _i32, _i32, _i32, _i32, _vp]),
    "etai_conv3x3": (C.c_int, [_vp, _vp, _vp, _vp, _vp, _i32, _i32, _i32, _i32, _i32, _i32, _i32, _i32, _vp, _i64, _vp]),
    "etai_attention": (C.c_int, [_vp, _vp, _vp, _vp, _i32, _i32, _i32, _i32, _i32, _i32, _i32, _i32, _i32, _f,
                                 C.POINTER(_i32), C.POINTER(_i32), C.POINTER(_i32), _i32, _i32, _vp]),
    "etai_cross_attention": (C.c_int, [_vp, _vp, _vp, _i32, _i32, _i32, _i32, _i32, _i32, _i32, _i32, _i32, _i32, _f, _i32,
                                       C.POINTER(_i32), C.POINTER(_i32), _vp, _vp, _vp, _vp, _i32, C.POINTER(_i32), _vp,
                                       _i32, _i32, _vp, _i64, _vp]),
}

_lib = None


def lib_path() -> Path:
    return Path(os.environ.get("ETAI_LIB", Path(__file__).resolve().parent / "libetai.so"))


def load() -> C.CDLL:
    """Load libetai.so once; raise loudly if it is not built (run ``python -c 'import __graft_entry__ as g; g.build()'``)."""
    global _lib
    if _lib is not None:
        return _lib
    p = lib_path()
    if not p.exists():
        raise RuntimeError(f"etai: native library {p} is missing; build it with eta_inversion_b200/csrc/build.sh "
                           "(there is no CPU or PyTorch fallback)")
    lib = C.CDLL(str(p))
    for name, (res, args) in SYMBOLS.items():
        fn = getattr(lib, name)  # AttributeError here means header/library drift
        fn.restype, fn.argtypes = res, args
    if lib.etai_abi_version() != ABI_VERSION:
        raise RuntimeError(f"etai: ABI mismatch, library {lib.etai_abi_version()} vs binding {ABI_VERSION}")
    _lib = lib
    return lib


def tensor_table(state_dict, storage_dtype=None):
    """state_dict -> (ctypes array of etai_tensor, keep-alive list): the weight table every *_create entry point takes.
    ``storage_dtype``: plain tensors (vectors, matrices, 1x1 convolutions) are converted to the engine's storage dtype here, on
    the host (same round-to-nearest as the device kernel), so the library only copies them; 3x3 filters and the GEGLU
    projection are re-packed on the device from fp32 and stay as they are."""
    keep, arr = [], (EtaiTensor * len(state_dict))()
    for i, (name, t) in enumerate(state_dict.items()):
        t = t.detach()
        if t.dtype not in (torch.float32, torch.float16, torch.bfloat16):
            t = t.float()
        plain = t.ndim <= 2 or tuple(t.shape[2:]) == (1, 1)
        if storage_dtype is not None and plain and ".ff.net.0.proj." not in name and t.dtype != storage_dtype:
            t = t.to(storage_dtype)
        t = t.contiguous()
        keep.append(t)
        arr[i].name = name.encode()
        arr[i].data = t.data_ptr()
        arr[i].dtype = dtype_code(t.dtype)
        arr[i].ndim = t.ndim
        for d in range(t.ndim):
            arr[i].shape[d] = t.shape[d]
        arr[i].on_device = 1 if t.is_cuda else 0
    return arr, keep


def check(rc: int) -> None:
    if rc != 0:
        msg = load().etai_last_error().decode(errors="replace")
        raise RuntimeError(f"etai error {rc}: {msg}")


def ptr(t) -> C.c_void_p:
    return C.c_void_p(0 if t is None else t.data_ptr())


def stream_ptr() -> C.c_void_p:
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


def i32_array(vals):
    arr = (C.c_int32 * len(vals))(*[int(v) for v in vals])
    return arr
