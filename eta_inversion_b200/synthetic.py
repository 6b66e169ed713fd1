"""Synthetic SD-1.5-architecture inputs: parameter specs (diffusers key names), seeded random-init
weights, a smooth synthetic 512x512 image and pre-generated variance-noise slabs.

There are no pretrained weights or datasets on the build/GPU boxes, so every
parity test and the benchmark use these.  The specs double as the loader contract for real
checkpoints: a diffusers ``unet.state_dict()`` has exactly these keys and shapes
(reference: modules/models/__init__.py:134-135 loads ``CompVis/stable-diffusion-v1-4``).
"""
from __future__ import annotations

from typing import Dict, List, Tuple

import torch

Spec = List[Tuple[str, Tuple[int, ...]]]

UNET_CHANNELS = (320, 640, 1280, 1280)
CROSS_DIM = 768
HEADS = 8
# Gain on the cross-attention to_q / to_k projections of the synthetic UNet.  Variance-preserving random weights give
# logits of unit variance whose softmax, once averaged over 5 layers x 8 heads, is spatially flat: every eta mask
# (eta_inversion.py:196-198, > 0.2) and LocalBlend mask (ptp.py:18-47, > 0.3) would be all-ones and the masked code paths
# would only ever run in their degenerate form.  With this gain the logits are heavy-tailed enough for the per-word maps
# to have real spatial structure (tests assert the mask fraction is strictly inside (0.05, 0.95)).
ATTN2_QK_GAIN = 2.0
# Gain on the token-embedding table of the synthetic CLIP text tower: with transformers' default init (token and position
# embeddings both std 0.02) the 77 output embeddings of a prompt are ~0.9 cosine-similar, so every word's attention map is
# the same map; with the gain the word identity dominates (cosine ~0.25 between different words).
TOKEN_EMB_GAIN = 32.0
# Gain on the synthetic VAE's quant_conv (weight and bias): SD's 0.18215 scaling factor exists to give latents ~unit
# variance, a variance-preserving random encoder gives std ~0.13 after it; then conv_in's bias and the time embedding swamp
# the image content and every UNet feature map is spatially constant.
VAE_LATENT_GAIN = 16.0
# Per-pixel white noise of the synthetic image (random convolutions amplify it into spatially incoherent features).
IMAGE_NOISE = 0.05


def _resnet(prefix: str, cin: int, cout: int, temb: int | None) -> Spec:
    s: Spec = [(f"{prefix}.norm1.weight", (cin,)), (f"{prefix}.norm1.bias", (cin,)),
               (f"{prefix}.conv1.weight", (cout, cin, 3, 3)), (f"{prefix}.conv1.bias", (cout,))]
    if temb:
        s += [(f"{prefix}.time_emb_proj.weight", (cout, temb)), (f"{prefix}.time_emb_proj.bias", (cout,))]
    s += [(f"{prefix}.norm2.weight", (cout,)), (f"{prefix}.norm2.bias", (cout,)),
          (f"{prefix}.conv2.weight", (cout, cout, 3, 3)), (f"{prefix}.conv2.bias", (cout,))]
    if cin != cout:
        s += [(f"{prefix}.conv_shortcut.weight", (cout, cin, 1, 1)), (f"{prefix}.conv_shortcut.bias", (cout,))]
    return s


def _transformer(prefix: str, c: int, cross: int) -> Spec:
    b = f"{prefix}.transformer_blocks.0"
    s: Spec = [(f"{prefix}.norm.weight", (c,)), (f"{prefix}.norm.bias", (c,)),
               (f"{prefix}.proj_in.weight", (c, c, 1, 1)), (f"{prefix}.proj_in.bias", (c,))]
    for n in ("norm1", "norm2", "norm3"):
        s += [(f"{b}.{n}.weight", (c,)), (f"{b}.{n}.bias", (c,))]
    for a, kv in (("attn1", c), ("attn2", cross)):
        s += [(f"{b}.{a}.to_q.weight", (c, c)), (f"{b}.{a}.to_k.weight", (c, kv)), (f"{b}.{a}.to_v.weight", (c, kv)),
              (f"{b}.{a}.to_out.0.weight", (c, c)), (f"{b}.{a}.to_out.0.bias", (c,))]
    s += [(f"{b}.ff.net.0.proj.weight", (8 * c, c)), (f"{b}.ff.net.0.proj.bias", (8 * c,)),
          (f"{b}.ff.net.2.weight", (c, 4 * c)), (f"{b}.ff.net.2.bias", (c,)),
          (f"{prefix}.proj_out.weight", (c, c, 1, 1)), (f"{prefix}.proj_out.bias", (c,))]
    return s


def unet_param_spec(channels=UNET_CHANNELS, cross=CROSS_DIM) -> Spec:
    """Keys/shapes of diffusers ``UNet2DConditionModel`` for SD-1.x."""
    c = list(channels)
    temb = c[0] * 4
    s: Spec = [("conv_in.weight", (c[0], 4, 3, 3)), ("conv_in.bias", (c[0],)),
               ("time_embedding.linear_1.weight", (temb, c[0])), ("time_embedding.linear_1.bias", (temb,)),
               ("time_embedding.linear_2.weight", (temb, temb)), ("time_embedding.linear_2.bias", (temb,))]
    cout = c[0]
    for i in range(4):
        cin, cout = cout, c[i]
        for j in range(2):
            s += _resnet(f"down_blocks.{i}.resnets.{j}", cin if j == 0 else cout, cout, temb)
            if i < 3:
                s += _transformer(f"down_blocks.{i}.attentions.{j}", cout, cross)
        if i < 3:
            s += [(f"down_blocks.{i}.downsamplers.0.conv.weight", (cout, cout, 3, 3)),
                  (f"down_blocks.{i}.downsamplers.0.conv.bias", (cout,))]
    s += _resnet("mid_block.resnets.0", c[3], c[3], temb)
    s += _transformer("mid_block.attentions.0", c[3], cross)
    s += _resnet("mid_block.resnets.1", c[3], c[3], temb)
    rev = c[::-1]
    cout = rev[0]
    for i in range(4):
        prev, cout = cout, rev[i]
        cin = rev[min(i + 1, 3)]
        for j in range(3):
            skip = cin if j == 2 else cout
            rin = prev if j == 0 else cout
            s += _resnet(f"up_blocks.{i}.resnets.{j}", rin + skip, cout, temb)
            if i > 0:
                s += _transformer(f"up_blocks.{i}.attentions.{j}", cout, cross)
        if i < 3:
            s += [(f"up_blocks.{i}.upsamplers.0.conv.weight", (cout, cout, 3, 3)),
                  (f"up_blocks.{i}.upsamplers.0.conv.bias", (cout,))]
    s += [("conv_norm_out.weight", (c[0],)), ("conv_norm_out.bias", (c[0],)),
          ("conv_out.weight", (4, c[0], 3, 3)), ("conv_out.bias", (4,))]
    return s


def _vae_attn(prefix: str, c: int) -> Spec:
    s: Spec = [(f"{prefix}.group_norm.weight", (c,)), (f"{prefix}.group_norm.bias", (c,))]
    for n in ("to_q", "to_k", "to_v", "to_out.0"):
        s += [(f"{prefix}.{n}.weight", (c, c)), (f"{prefix}.{n}.bias", (c,))]
    return s


def vae_param_spec(ch=(128, 256, 512, 512)) -> Spec:
    """Keys/shapes of diffusers ``AutoencoderKL`` for SD-1.x (0.21.1 attention naming)."""
    s: Spec = [("encoder.conv_in.weight", (ch[0], 3, 3, 3)), ("encoder.conv_in.bias", (ch[0],))]
    cout = ch[0]
    for i, c in enumerate(ch):
        cin, cout = cout, c
        for j in range(2):
            s += _resnet(f"encoder.down_blocks.{i}.resnets.{j}", cin if j == 0 else cout, cout, None)
        if i < 3:
            s += [(f"encoder.down_blocks.{i}.downsamplers.0.conv.weight", (cout, cout, 3, 3)),
                  (f"encoder.down_blocks.{i}.downsamplers.0.conv.bias", (cout,))]
    top = ch[-1]
    s += _resnet("encoder.mid_block.resnets.0", top, top, None) + _resnet("encoder.mid_block.resnets.1", top, top, None)
    s += _vae_attn("encoder.mid_block.attentions.0", top)
    s += [("encoder.conv_norm_out.weight", (top,)), ("encoder.conv_norm_out.bias", (top,)),
          ("encoder.conv_out.weight", (8, top, 3, 3)), ("encoder.conv_out.bias", (8,))]
    s += [("decoder.conv_in.weight", (top, 4, 3, 3)), ("decoder.conv_in.bias", (top,))]
    s += _resnet("decoder.mid_block.resnets.0", top, top, None) + _resnet("decoder.mid_block.resnets.1", top, top, None)
    s += _vae_attn("decoder.mid_block.attentions.0", top)
    rev = list(ch[::-1])
    cout = rev[0]
    for i, c in enumerate(rev):
        cin, cout = cout, c
        for j in range(3):
            s += _resnet(f"decoder.up_blocks.{i}.resnets.{j}", cin if j == 0 else cout, cout, None)
        if i < 3:
            s += [(f"decoder.up_blocks.{i}.upsamplers.0.conv.weight", (cout, cout, 3, 3)),
                  (f"decoder.up_blocks.{i}.upsamplers.0.conv.bias", (cout,))]
    s += [("decoder.conv_norm_out.weight", (ch[0],)), ("decoder.conv_norm_out.bias", (ch[0],)),
          ("decoder.conv_out.weight", (3, ch[0], 3, 3)), ("decoder.conv_out.bias", (3,))]
    s += [("quant_conv.weight", (8, 8, 1, 1)), ("quant_conv.bias", (8,)),
          ("post_quant_conv.weight", (4, 4, 1, 1)), ("post_quant_conv.bias", (4,))]
    return s


def _is_norm(name: str) -> bool:
    leaf = name.rsplit(".", 2)[-2]
    return leaf.startswith("norm") or leaf in ("conv_norm_out", "group_norm")


def random_state_dict(spec: Spec, seed: int = 0, dtype=torch.float32) -> Dict[str, torch.Tensor]:
    """Seeded init, one generator per tensor so any subset can be regenerated independently.

    linear/conv weights ~ N(0, 1/fan_in) (variance preserving so activations stay O(1) for 50 steps and
    in fp16), their biases ~ 0.05 N(0,1); norm scales 1 + 0.1 N(0,1), norm shifts 0.1 N(0,1)."""
    out: Dict[str, torch.Tensor] = {}
    for idx, (name, shape) in enumerate(spec):
        g = torch.Generator().manual_seed(seed * 1_000_003 + idx)
        r = torch.randn(shape, generator=g, dtype=torch.float32)
        if _is_norm(name):
            t = 1.0 + 0.1 * r if name.endswith(".weight") else 0.1 * r
        elif name.endswith(".bias"):
            t = 0.05 * r
        else:
            fan_in = 1
            for d in shape[1:]:
                fan_in *= d
            t = r * (fan_in ** -0.5)
            if ".attn2.to_q." in name or ".attn2.to_k." in name:
                t = t * ATTN2_QK_GAIN
        if name.startswith("quant_conv."):
            t = t * VAE_LATENT_GAIN
        out[name] = t.to(dtype)
    return out


def synthetic_image(seed: int = 0, size: int = 512) -> torch.Tensor:
    """Synthetic image in [-1, 1], shape [1,3,size,size]: a smooth low-frequency background plus
    three sharp-edged, flat-coloured ellipses ("objects").  The objects matter: a purely smooth image gives spatially
    flat attention maps under random-init weights, so the eta mask and the LocalBlend mask would be all-ones."""
    g = torch.Generator().manual_seed(10_000 + seed)
    yy, xx = torch.meshgrid(torch.linspace(0, 1, size), torch.linspace(0, 1, size), indexing="ij")
    img = torch.zeros(3, size, size)
    for c in range(3):
        for _ in range(4):
            fx, fy, ph = (torch.rand(3, generator=g) * torch.tensor([6.0, 6.0, 6.2832])).tolist()
            amp = float(torch.rand(1, generator=g)) * 0.5
            img[c] += amp * torch.sin(6.2832 * (fx * xx + fy * yy) + ph)
    img += IMAGE_NOISE * torch.randn(3, size, size, generator=g)
    img = 0.5 * img / img.abs().max()
    for _ in range(3):
        cx, cy, rx, ry = (torch.rand(4, generator=g) * torch.tensor([0.6, 0.6, 0.15, 0.15]) + torch.tensor([0.2, 0.2, 0.1, 0.1])).tolist()
        colour = (torch.rand(3, generator=g) * 2 - 1) * 0.9
        inside = ((xx - cx) / rx) ** 2 + ((yy - cy) / ry) ** 2 < 1.0
        img = torch.where(inside[None], colour[:, None, None] + 0.1 * img, img)
    return img.clamp(-1, 1)[None].contiguous()


def variance_noise_slabs(steps: int, count: int = 10, seed: int = 0) -> torch.Tensor:
    """Pre-generated [steps, count, 1, 4, 64, 64] noise, one slab per backward step; stands in for
    ``torch.randn((10,1,4,64,64), generator=...)`` at eta_inversion.py:156,276,348 so CPU oracle and
    GPU engine consume identical numbers (RNG quirk)."""
    g = torch.Generator().manual_seed(seed)
    return torch.randn((steps, count, 1, 4, 64, 64), generator=g)


def make_text_encoder(seed: int = 0):
    """Random-init CLIP ViT-L/14 text tower (transformers ``CLIPTextModel``; there are no pretrained weights on the
    box).  Shared input definition of the oracle and the engine (modules/models/__init__.py:135 loads the real one)."""
    from transformers import CLIPTextConfig, CLIPTextModel
    cfg = CLIPTextConfig(vocab_size=49408, hidden_size=768, intermediate_size=3072, num_hidden_layers=12,
                         num_attention_heads=12, max_position_embeddings=77, hidden_act="quick_gelu")
    with torch.random.fork_rng():
        torch.manual_seed(seed)
        m = CLIPTextModel(cfg)
    with torch.no_grad():
        m.text_model.embeddings.token_embedding.weight.mul_(TOKEN_EMB_GAIN)
    return m.eval().requires_grad_(False)
