"""Python faces of the native AutoencoderKL and CLIP text tower (libetai.so: csrc/vae.cu, csrc/clip.cu).

Both sit around the two diffusion loops of an edit (1 encode + 2 decodes + 4 text encodes, reference
modules/inversion/diffusion_inversion.py:183-247) and keep the call shapes the reference's loops use:
``vae.encode(x)['latent_dist'].mean``, ``vae.decode(z)['sample']``, ``vae.dtype``, ``text_encoder(input_ids)[0]``.
Weights are diffusers / transformers state dicts (``AutoencoderKL``, ``CLIPTextModel``).  No torch.nn / cuDNN / cuBLAS
compute is involved and there is no fallback: without libetai.so or a CUDA device the constructors raise."""
from __future__ import annotations

import ctypes as C
from types import SimpleNamespace
from typing import Dict, Optional, Sequence

import torch

from . import _lib
from ._lib import MATH_AUTO, EtaiClipCfg, EtaiVaeCfg, check, dtype_code, ptr, stream_ptr, tensor_table

SD15_VAE_CHANNELS = (128, 256, 512, 512)


def _device(device) -> torch.device:
    if not torch.cuda.is_available():
        raise RuntimeError("etai: a CUDA device is required (no CPU fallback exists)")
    dev = device if isinstance(device, torch.device) else torch.device(device if isinstance(device, str) else f"cuda:{device}")
    if dev.index is None:
        dev = torch.device("cuda", torch.cuda.current_device())
    return dev


class VAEEngine:
    """Handle on the native SD-1.x AutoencoderKL; drop-in for ``pipe.vae`` in the reference's encode/decode."""

    def __init__(self, state_dict: Dict[str, torch.Tensor], dtype: torch.dtype = torch.float32, device=0, max_batch: int = 1,
                 channels: Sequence[int] = SD15_VAE_CHANNELS, image_hw: int = 512, math_mode: int = MATH_AUTO):
        lib = _lib.load()
        self.device, self.dtype, self.max_batch, self.image_hw = _device(device), dtype, max_batch, image_hw
        cfg = EtaiVaeCfg()
        cfg.dtype, cfg.math_mode, cfg.image_hw, cfg.max_batch = dtype_code(dtype), math_mode, image_hw, max_batch
        for i, c in enumerate(channels):
            cfg.block_out_channels[i] = c
        arr, keep = tensor_table(state_dict, dtype)
        h = C.c_void_p()
        with torch.cuda.device(self.device):
            check(lib.etai_vae_create(C.byref(h), C.byref(cfg), arr, len(state_dict), self.device.index or 0))
        del keep
        self._h, self._lib = h, lib

    def close(self):
        if getattr(self, "_h", None):
            self._lib.etai_vae_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    @property
    def launch_count(self) -> int:
        return int(self._lib.etai_vae_launch_count(self._h))

    @property
    def device_bytes(self) -> int:
        return int(self._lib.etai_vae_device_bytes(self._h))

    def _io(self, x: torch.Tensor, ch: int, hw: int, name: str) -> torch.Tensor:
        if not (isinstance(x, torch.Tensor) and x.is_cuda):
            raise RuntimeError(f"etai: {name} must be a CUDA tensor (no CPU fallback exists)")
        if x.ndim != 4 or x.shape[1] != ch or x.shape[2] != hw or x.shape[3] != hw:
            raise RuntimeError(f"etai: {name} must be [B,{ch},{hw},{hw}], got {tuple(x.shape)}")
        if x.dtype not in (torch.float32, torch.float16, torch.bfloat16):
            x = x.float()
        return x.contiguous()

    def encode(self, image: torch.Tensor):
        """[B,3,hw,hw] in [-1,1] -> ``{'latent_dist': ns(mean=[B,4,hw/8,hw/8])}`` in the input's dtype."""
        x = self._io(image, 3, self.image_hw, "image")
        out = torch.empty((x.shape[0], 4, self.image_hw // 8, self.image_hw // 8), dtype=x.dtype, device=x.device)
        with torch.cuda.device(self.device):
            for i in range(0, x.shape[0], self.max_batch):
                xb, ob = x[i:i + self.max_batch], out[i:i + self.max_batch]
                check(self._lib.etai_vae_encode(self._h, ptr(xb), dtype_code(x.dtype), xb.shape[0], ptr(ob), stream_ptr()))
        return {"latent_dist": SimpleNamespace(mean=out)}

    def decode(self, latent: torch.Tensor):
        """[B,4,hw/8,hw/8] (already divided by 0.18215) -> ``{'sample': [B,3,hw,hw]}`` in the input's dtype."""
        z = self._io(latent, 4, self.image_hw // 8, "latent")
        out = torch.empty((z.shape[0], 3, self.image_hw, self.image_hw), dtype=z.dtype, device=z.device)
        with torch.cuda.device(self.device):
            for i in range(0, z.shape[0], self.max_batch):
                zb, ob = z[i:i + self.max_batch], out[i:i + self.max_batch]
                check(self._lib.etai_vae_decode(self._h, ptr(zb), dtype_code(z.dtype), zb.shape[0], ptr(ob), stream_ptr()))
        return {"sample": out}


class CLIPTextEngine:
    """Handle on the native CLIP ViT-L/14 text tower; ``text_encoder(input_ids)[0]`` like transformers' CLIPTextModel.
    ``input_ids``: integer tensor [B,77], preferably on the CPU (the ids are host data: the tokenizer produced them)."""

    def __init__(self, state_dict: Dict[str, torch.Tensor], dtype: torch.dtype = torch.float32, device=0, max_batch: int = 8,
                 vocab: int = 49408, hidden: int = 768, layers: int = 12, heads: int = 12, ffn: int = 3072, max_len: int = 77,
                 math_mode: int = MATH_AUTO):
        lib = _lib.load()
        self.device, self.dtype, self.max_batch, self.max_len, self.hidden = _device(device), dtype, max_batch, max_len, hidden
        cfg = EtaiClipCfg()
        cfg.dtype, cfg.math_mode = dtype_code(dtype), math_mode
        cfg.vocab, cfg.hidden, cfg.layers, cfg.heads, cfg.ffn, cfg.max_len, cfg.max_batch = (vocab, hidden, layers, heads, ffn,
                                                                                            max_len, max_batch)
        sd = {k: v for k, v in state_dict.items() if not k.endswith("position_ids")}
        arr, keep = tensor_table(sd, dtype)
        h = C.c_void_p()
        with torch.cuda.device(self.device):
            check(lib.etai_clip_create(C.byref(h), C.byref(cfg), arr, len(sd), self.device.index or 0))
        del keep
        self._h, self._lib = h, lib

    @classmethod
    def from_transformers(cls, model, dtype: torch.dtype, device, max_batch: int = 8) -> "CLIPTextEngine":
        """Bind a ``transformers.CLIPTextModel`` (weights + config); compute is native afterwards."""
        c = model.config
        if getattr(c, "hidden_act", "quick_gelu") != "quick_gelu":
            raise RuntimeError(f"etai: CLIP activation '{c.hidden_act}' is not built (SD-1.x uses quick_gelu)")
        return cls(model.state_dict(), dtype=dtype, device=device, max_batch=max_batch, vocab=c.vocab_size,
                   hidden=c.hidden_size, layers=c.num_hidden_layers, heads=c.num_attention_heads, ffn=c.intermediate_size,
                   max_len=c.max_position_embeddings)

    def close(self):
        if getattr(self, "_h", None):
            self._lib.etai_clip_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    @property
    def launch_count(self) -> int:
        return int(self._lib.etai_clip_launch_count(self._h))

    def __call__(self, input_ids: torch.Tensor, attention_mask: Optional[torch.Tensor] = None):
        if attention_mask is not None:
            raise RuntimeError("etai: the text tower applies the causal mask only (the reference passes no attention mask)")
        ids = input_ids.detach().to("cpu", torch.int32).contiguous()  # a CUDA tensor here costs a device->host sync
        if ids.ndim != 2 or ids.shape[1] != self.max_len:
            raise RuntimeError(f"etai: input_ids must be [B,{self.max_len}], got {tuple(ids.shape)}")
        out = torch.empty((ids.shape[0], self.max_len, self.hidden), dtype=self.dtype, device=self.device)
        with torch.cuda.device(self.device):
            for i in range(0, ids.shape[0], self.max_batch):
                ib, ob = ids[i:i + self.max_batch], out[i:i + self.max_batch]
                check(self._lib.etai_clip_encode(self._h, C.cast(C.c_void_p(ib.data_ptr()), C.POINTER(C.c_int32)), ib.shape[0],
                                                 ptr(ob), dtype_code(self.dtype), stream_ptr()))
        return (out,)
