"""SD-1.x AutoencoderKL and the random-init CLIP text tower used by the synthetic pipeline.

Both sit OUTSIDE the hot loops (1 encode + 2 decodes + 4 text encodes per edit, diffusion_inversion.py:183-247) and
are scheduled as the next native components; today they are plain torch.nn modules executed by
torch/cuDNN on the GPU.  Parameter names follow diffusers so real checkpoints load (synthetic.vae_param_spec)."""
from __future__ import annotations

from types import SimpleNamespace

import torch
import torch.nn as nn
import torch.nn.functional as F


class _GN(nn.GroupNorm):
    """GroupNorm(+SiLU).  On CUDA it runs the engine's fused NHWC kernel (etai_groupnorm) on the channels_last
    storage of the activation, which also keeps cuDNN's convolutions in NHWC; on CPU it is plain torch."""

    def forward(self, x, silu: bool = False):
        if x.is_cuda and x.ndim == 4 and x.shape[1] % 8 == 0:
            from . import engine as E
            b, c, h, w = x.shape
            xh = x.contiguous(memory_format=torch.channels_last).permute(0, 2, 3, 1).reshape(b, h * w, c)
            y = E.groupnorm(xh, self.weight, self.bias, self.num_groups, self.eps, silu)
            return y.reshape(b, h, w, c).permute(0, 3, 1, 2)
        y = super().forward(x)
        return F.silu(y) if silu else y


def _gn(c):
    return _GN(32, c, eps=1e-6)


class _Res(nn.Module):
    def __init__(self, cin, cout):
        super().__init__()
        self.norm1, self.conv1 = _gn(cin), nn.Conv2d(cin, cout, 3, padding=1)
        self.norm2, self.conv2 = _gn(cout), nn.Conv2d(cout, cout, 3, padding=1)
        self.conv_shortcut = nn.Conv2d(cin, cout, 1) if cin != cout else None

    def forward(self, x):
        h = self.conv1(self.norm1(x, silu=True))
        h = self.conv2(self.norm2(h, silu=True))
        return (x if self.conv_shortcut is None else self.conv_shortcut(x)) + h


class _Attn(nn.Module):
    def __init__(self, c):
        super().__init__()
        self.group_norm = _gn(c)
        self.to_q, self.to_k, self.to_v = nn.Linear(c, c), nn.Linear(c, c), nn.Linear(c, c)
        self.to_out = nn.ModuleList([nn.Linear(c, c)])

    def forward(self, x):
        b, c, h, w = x.shape
        t = self.group_norm(x).reshape(b, c, h * w).transpose(1, 2)  # [B, HW, C]
        o = F.scaled_dot_product_attention(self.to_q(t)[:, None], self.to_k(t)[:, None], self.to_v(t)[:, None])[:, 0]
        return x + self.to_out[0](o).transpose(1, 2).reshape(b, c, h, w)


class _Conv(nn.Module):  # holder so keys read "...samplers.0.conv.weight"
    def __init__(self, c, stride):
        super().__init__()
        self.conv = nn.Conv2d(c, c, 3, stride=stride, padding=0 if stride == 2 else 1)


class _Mid(nn.Module):
    def __init__(self, c):
        super().__init__()
        self.resnets = nn.ModuleList([_Res(c, c), _Res(c, c)])
        self.attentions = nn.ModuleList([_Attn(c)])

    def forward(self, x):
        return self.resnets[1](self.attentions[0](self.resnets[0](x)))


class _Level(nn.Module):
    def __init__(self, cin, cout, n, down=False, up=False):
        super().__init__()
        self.resnets = nn.ModuleList([_Res(cin if i == 0 else cout, cout) for i in range(n)])
        if down:
            self.downsamplers = nn.ModuleList([_Conv(cout, 2)])
        if up:
            self.upsamplers = nn.ModuleList([_Conv(cout, 1)])

    def forward(self, x):
        for r in self.resnets:
            x = r(x)
        if hasattr(self, "downsamplers"):
            x = self.downsamplers[0].conv(F.pad(x, (0, 1, 0, 1)))
        if hasattr(self, "upsamplers"):
            x = self.upsamplers[0].conv(F.interpolate(x, scale_factor=2.0, mode="nearest"))
        return x


class _Encoder(nn.Module):
    def __init__(self, ch):
        super().__init__()
        self.conv_in = nn.Conv2d(3, ch[0], 3, padding=1)
        self.down_blocks = nn.ModuleList([_Level(ch[max(i - 1, 0)], c, 2, down=i < len(ch) - 1) for i, c in enumerate(ch)])
        self.mid_block = _Mid(ch[-1])
        self.conv_norm_out, self.conv_out = _gn(ch[-1]), nn.Conv2d(ch[-1], 8, 3, padding=1)

    def forward(self, x):
        x = self.conv_in(x)
        for b in self.down_blocks:
            x = b(x)
        return self.conv_out(self.conv_norm_out(self.mid_block(x), silu=True))


class _Decoder(nn.Module):
    def __init__(self, ch):
        super().__init__()
        rev = list(ch[::-1])
        self.conv_in = nn.Conv2d(4, rev[0], 3, padding=1)
        self.mid_block = _Mid(rev[0])
        self.up_blocks = nn.ModuleList([_Level(rev[max(i - 1, 0)], c, 3, up=i < len(ch) - 1) for i, c in enumerate(rev)])
        self.conv_norm_out, self.conv_out = _gn(ch[0]), nn.Conv2d(ch[0], 3, 3, padding=1)

    def forward(self, z):
        x = self.mid_block(self.conv_in(z))
        for b in self.up_blocks:
            x = b(x)
        return self.conv_out(self.conv_norm_out(x, silu=True))


class AutoencoderKL(nn.Module):
    def __init__(self, ch=(128, 256, 512, 512)):
        super().__init__()
        self.encoder, self.decoder = _Encoder(ch), _Decoder(ch)
        self.quant_conv, self.post_quant_conv = nn.Conv2d(8, 8, 1), nn.Conv2d(4, 4, 1)

    @property
    def dtype(self):
        return self.quant_conv.weight.dtype

    def encode(self, x):
        mean = self.quant_conv(self.encoder(x))[:, :4]
        return {"latent_dist": SimpleNamespace(mean=mean)}

    def decode(self, z):
        return {"sample": self.decoder(self.post_quant_conv(z))}


def make_text_encoder(seed: int = 0):
    """Random-init CLIP ViT-L/14 text tower (see synthetic.make_text_encoder)."""
    from .synthetic import make_text_encoder as mk
    return mk(seed)
