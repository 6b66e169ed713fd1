"""Model container, pre/post-processing and loader; mirrors modules/models/__init__.py of the reference
(StablePreprocess :11-76, StablePostProc :79-101, load_diffusion_model :104-138).

The pipeline object exposes exactly what the reference's loops touch:
``.unet(sample, t, encoder_hidden_states=)["sample"]`` (the native engine), ``.unet.dtype``, ``.vae.encode/.decode``,
``.vae.dtype``, ``.tokenizer``, ``.text_encoder(ids)[0]``, ``.scheduler``, ``.device``.
"""
from __future__ import annotations

import zlib
from pathlib import Path
from types import SimpleNamespace
from typing import Any, Dict, Optional, Tuple, Union

import numpy as np
import torch

from . import synthetic
from .engine import UNetEngine
from .inverse_schedulers import DDIMScheduler
from .vae import CLIPTextEngine, VAEEngine


class StablePreprocess:
    """Image file / uint8 array -> [1,3,size,size] float tensor in [-1,1] on `device`."""

    def __init__(self, device: str, size: int = 512, return_np: bool = False, center_crop: bool = False,
                 pil_resize: bool = False) -> None:
        self.device, self.size, self.return_np = device, size, return_np
        self.center_crop, self.pil_resize = center_crop, pil_resize

    def __call__(self, image: Union[str, Path, np.ndarray]):
        import cv2
        if isinstance(image, (str, Path)):
            image = cv2.cvtColor(cv2.imread(str(image)), cv2.COLOR_BGR2RGB)
        if self.center_crop:
            h, w = image.shape[:2]
            if w > h:
                lo = (w - h) // 2
                hi = w - h - lo
                if hi > 0:
                    image = image[:, lo:-hi]
            else:
                lo = (h - w) // 2
                hi = h - w - lo
                if hi > 0:
                    image = image[lo:-hi]
        if self.pil_resize:
            from PIL import Image
            image = np.array(Image.fromarray(image).resize((self.size, self.size)))
        else:
            image = cv2.resize(image, (self.size, self.size))
        image_pt = (torch.from_numpy(image).float() / 127.5 - 1).permute(2, 0, 1).unsqueeze(0).to(self.device)
        return (image_pt, image) if self.return_np else image_pt


class StablePostProc:
    """VAE output -> uint8 HWC array of the first image."""

    def __call__(self, image: torch.Tensor) -> np.ndarray:
        # same arithmetic and truncation as the reference ((x/2+0.5).clamp(0,1) * 255 -> uint8), done on the image's device so
        # that only the first image's HWC uint8 bytes cross PCIe (the reference moves the fp32 NCHW batch and converts on the host)
        image = ((image[:1].float() / 2 + 0.5).clamp(0, 1) * 255).to(torch.uint8)
        return image.permute(0, 2, 3, 1).contiguous().cpu().numpy()[0]


class SyntheticTokenizer:
    """Deterministic whitespace tokenizer standing in for CLIPTokenizer (no vocab files on the box):
    one id per word (crc32), BOS 49406, EOS/pad 49407, so ``get_word_inds`` == word index + 1."""

    bos, eos = 49406, 49407
    model_max_length = 77

    def __init__(self):
        self._words: Dict[int, str] = {}

    def _id(self, w: str) -> int:
        i = zlib.crc32(w.encode()) % 49000 + 1
        self._words[i] = w
        return i

    def encode(self, text: str):
        return [self.bos] + [self._id(w) for w in text.split(" ") if w != ""] + [self.eos]

    def decode(self, ids) -> str:
        names = {self.bos: "<|startoftext|>", self.eos: "<|endoftext|>"}
        return " ".join(names.get(int(i), self._words.get(int(i), "?")) for i in ids)

    def __call__(self, texts, padding="max_length", max_length=77, truncation=True, return_tensors="pt"):
        if isinstance(texts, str):
            texts = [texts]
        rows = []
        for t in texts:
            ids = self.encode(t)[:max_length]
            if len(ids) == max_length:
                ids[-1] = self.eos
            rows.append(ids + [self.eos] * (max_length - len(ids)))
        return SimpleNamespace(input_ids=torch.tensor(rows, dtype=torch.int64))


class EtaiPipeline:
    """Stand-in for diffusers' StableDiffusionPipeline with the native UNet inside."""

    def __init__(self, unet: UNetEngine, vae, text_encoder, tokenizer, scheduler, device):
        self.unet, self.vae, self.text_encoder, self.tokenizer, self.scheduler = unet, vae, text_encoder, tokenizer, scheduler
        self.device = torch.device(device)


def clone_pipeline(pipe: "EtaiPipeline", max_batch: Optional[int] = None) -> "EtaiPipeline":
    """A second pipeline on the same device for ``batching.run_pipelined``: its own UNet engine handle (activation arena,
    CUDA graphs, stream) on the SAME packed device weights as ``pipe`` (``etai_unet_clone``: nothing is re-uploaded, both
    groups stream one copy of the weights), sharing the read-only VAE / text encoder / tokenizer of ``pipe``."""
    clone = EtaiPipeline(pipe.unet.clone(max_batch), pipe.vae, pipe.text_encoder, pipe.tokenizer, sd_scheduler(), pipe.device)
    if hasattr(pipe, "cache_text_embeddings"):
        clone.cache_text_embeddings = pipe.cache_text_embeddings
    return clone


def sd_scheduler() -> DDIMScheduler:
    """modules/models/__init__.py:134 plus steps_offset=1 from the SD-1.x pipeline config."""
    return DDIMScheduler(beta_start=0.00085, beta_end=0.012, beta_schedule="scaled_linear", clip_sample=False,
                         set_alpha_to_one=False, steps_offset=1)


REAL_MODEL_NAMES = ("sd14", "CompVis/stable-diffusion-v1-4", "runwayml/stable-diffusion-v1-5")


def load_diffusion_model(model: str = "synthetic-sd15", device: str = "cuda", preproc_args: Optional[Dict[str, Any]] = None,
                         variant: Optional[str] = None, max_batch: int = 4, seed: int = 0,
                         unet_state_dict: Optional[Dict[str, torch.Tensor]] = None,
                         vae_state_dict: Optional[Dict[str, torch.Tensor]] = None,
                         text_encoder: Optional[torch.nn.Module] = None, tokenizer: Any = None, **kwargs
                         ) -> Tuple[EtaiPipeline, Tuple[StablePreprocess, StablePostProc]]:
    """Same signature/return shape as the reference loader (modules/models/__init__.py:104-138).  ``model``:
      * "synthetic-sd15" (default): SD-1.x architecture with seeded random-init UNet / VAE / CLIP text tower and the
        whitespace tokenizer -- the only thing that can be built on a box without checkpoints or network;
      * "sd14" / "CompVis/stable-diffusion-v1-4" (the reference's names): a REAL checkpoint, which this loader cannot
        download; all four parts must be supplied -- ``unet_state_dict`` and ``vae_state_dict`` (diffusers key names),
        ``text_encoder`` (a ``CLIPTextModel``) and ``tokenizer`` (a ``CLIPTokenizer``) -- otherwise it raises instead of
        silently editing with random weights.
    ``variant``: "fp32" (SIMT fp32 parity path) | "fp16" | "bf16" (tcgen05 path)."""
    if model in REAL_MODEL_NAMES:
        missing = [n for n, v in (("unet_state_dict", unet_state_dict), ("vae_state_dict", vae_state_dict),
                                  ("text_encoder", text_encoder), ("tokenizer", tokenizer)) if v is None]
        if missing:
            raise RuntimeError(f"etai: model '{model}' names a pretrained checkpoint, which cannot be downloaded here; pass "
                               f"{', '.join(missing)} (or use model='synthetic-sd15' for the random-init architecture)")
    elif model != "synthetic-sd15":
        raise Exception(model)
    variant = variant or "fp32"
    dtype = {"fp32": torch.float32, "fp16": torch.float16, "bf16": torch.bfloat16}[variant]
    if not str(device).startswith("cuda"):
        raise RuntimeError("etai: the engine runs on CUDA devices only (no CPU fallback)")
    dev = torch.device(device if ":" in str(device) else f"cuda:{torch.cuda.current_device()}")
    print(f"Loading model {model} ({variant}) ...")
    usd = unet_state_dict if unet_state_dict is not None else synthetic.random_state_dict(synthetic.unet_param_spec(), seed)
    unet = UNetEngine(usd, dtype=dtype, device=dev, max_batch=max_batch)
    del usd
    # VAE and CLIP text tower: native handles too (csrc/vae.cu, csrc/clip.cu); 16-bit variants run both in 16-bit as the
    # reference does.  The handles serialise their own calls, so lanes / pipelined groups share them.
    vae = VAEEngine(vae_state_dict if vae_state_dict is not None
                    else synthetic.random_state_dict(synthetic.vae_param_spec(), seed + 1), dtype=dtype, device=dev, max_batch=2)
    if isinstance(text_encoder, CLIPTextEngine):
        pass
    else:  # a transformers CLIPTextModel (real checkpoint) or the seeded random-init tower: weights + config only
        text_encoder = CLIPTextEngine.from_transformers(
            text_encoder if text_encoder is not None else synthetic.make_text_encoder(seed), dtype=dtype, device=dev)
    pipe = EtaiPipeline(unet, vae, text_encoder, tokenizer if tokenizer is not None else SyntheticTokenizer(), sd_scheduler(), dev)
    return pipe, (StablePreprocess(str(dev), size=512, **(preproc_args or {})), StablePostProc())
