"""ORACLE (test infrastructure, never shipped, never on the product path).

CPU/fp32 restatement, in plain PyTorch, of the third-party arithmetic the reference's hot
path calls into: ``diffusers==0.21.1`` (requirements.txt:1 of the reference; NOT vendored
under /root/reference and not installable here) and the ``transformers`` CLIP text model.

What is restated and which reference call sites it serves:

* ``UNet2DConditionModel`` (SD-1.x config)   <- modules/inversion/diffusion_inversion.py:264-280,
                                                modules/inversion/eta_inversion.py:321
* ``Attention`` module surface                <- modules/utils/ptp_utils.py:205-260,
                                                modules/utils/masactrl_utils.py:78-127,
                                                modules/utils/pnp_utils.py:67-118
* ``ResnetBlock2D`` attribute surface         <- modules/utils/pnp_utils.py:136-185
* ``DDIMScheduler`` (set_timesteps/step/_get_variance, steps_offset=1)
                                              <- diffusion_inversion.py:146,153,312;
                                                eta_inversion.py:245,310-312,369
* ``AutoencoderKL`` (encode().latent_dist.mean / decode().sample)
                                              <- diffusion_inversion.py:183-208
* ``StableDiffusionPipeline`` container       <- modules/models/__init__.py:134-135

PARITY STATUS: the reference's own goldens for this path (test/test_inv.py:44-53,
test/test_edit.py:66-108) need pretrained SD-1.4 weights + CUDA, which do not exist in
this container, so this restatement of *diffusers* is "parity unpinned" against diffusers
itself.  What IS pinned: the reference's own loop / controller / scheduler code
(modules/**) is executed unmodified on top of this restatement (oracle/run_reference.py)
and its outputs are committed under tests/golden/; self-consistency checks live in
tests/test_oracle.py.
"""
from __future__ import annotations

import math
import zlib
from collections import namedtuple
from types import SimpleNamespace
from typing import Optional

import numpy as np
import torch
import torch.nn as nn
import torch.nn.functional as F


# --------------------------------------------------------------------------------------
# config helper (diffusers.configuration_utils.FrozenDict)
# --------------------------------------------------------------------------------------
class FrozenDict(dict):
    """Mapping with attribute access; ``{**cfg}`` works (diffusion_inversion.py:146)."""

    def __getattr__(self, k):
        try:
            return self[k]
        except KeyError as e:  # pragma: no cover
            raise AttributeError(k) from e

    def __setattr__(self, k, v):  # pragma: no cover
        raise TypeError("FrozenDict is immutable")


# --------------------------------------------------------------------------------------
# UNet building blocks
# --------------------------------------------------------------------------------------
def timestep_embedding(timesteps: torch.Tensor, dim: int = 320) -> torch.Tensor:
    """diffusers get_timestep_embedding(flip_sin_to_cos=True, downscale_freq_shift=0)."""
    half = dim // 2
    exponent = -math.log(10000.0) * torch.arange(half, dtype=torch.float32, device=timesteps.device) / half
    emb = timesteps[:, None].float() * torch.exp(exponent)[None, :]
    return torch.cat([torch.cos(emb), torch.sin(emb)], dim=-1)


class TimestepEmbedding(nn.Module):
    def __init__(self, cin, cout):
        super().__init__()
        self.linear_1 = nn.Linear(cin, cout)
        self.act = nn.SiLU()
        self.linear_2 = nn.Linear(cout, cout)

    def forward(self, x):
        return self.linear_2(self.act(self.linear_1(x)))


class Attention(nn.Module):
    """Class name must be ``Attention`` (ptp_utils.py:279, masactrl_utils.py:131)."""

    def __init__(self, query_dim, cross_attention_dim=None, heads=8, dim_head=64, bias=False,
                 out_bias=True, norm_num_groups=None, eps=1e-5, residual_connection=False):
        super().__init__()
        inner = heads * dim_head
        kv = cross_attention_dim if cross_attention_dim is not None else query_dim
        self.heads = heads
        self.scale = dim_head ** -0.5
        self.residual_connection = residual_connection
        self.group_norm = nn.GroupNorm(norm_num_groups, query_dim, eps=eps) if norm_num_groups else None
        self.to_q = nn.Linear(query_dim, inner, bias=bias)
        self.to_k = nn.Linear(kv, inner, bias=bias)
        self.to_v = nn.Linear(kv, inner, bias=bias)
        self.to_out = nn.ModuleList([nn.Linear(inner, query_dim, bias=out_bias), nn.Dropout(0.0)])

    def head_to_batch_dim(self, t):
        b, n, c = t.shape
        h = self.heads
        return t.reshape(b, n, h, c // h).permute(0, 2, 1, 3).reshape(b * h, n, c // h)

    def batch_to_head_dim(self, t):
        bh, n, d = t.shape
        h = self.heads
        return t.reshape(bh // h, h, n, d).permute(0, 2, 1, 3).reshape(bh // h, n, d * h)

    def forward(self, hidden_states, encoder_hidden_states=None, attention_mask=None):
        residual = hidden_states
        spatial = hidden_states.ndim == 4
        if spatial:
            b, c, hh, ww = hidden_states.shape
            hidden_states = hidden_states.view(b, c, hh * ww).transpose(1, 2)
        if self.group_norm is not None:
            hidden_states = self.group_norm(hidden_states.transpose(1, 2)).transpose(1, 2)
        ctx = hidden_states if encoder_hidden_states is None else encoder_hidden_states
        q = self.head_to_batch_dim(self.to_q(hidden_states))
        k = self.head_to_batch_dim(self.to_k(ctx))
        v = self.head_to_batch_dim(self.to_v(ctx))
        out = F.scaled_dot_product_attention(q, k, v, scale=self.scale)
        out = self.to_out[1](self.to_out[0](self.batch_to_head_dim(out)))
        if spatial:
            out = out.transpose(-1, -2).reshape(b, c, hh, ww)
        if self.residual_connection:
            out = out + residual
        return out


class GEGLU(nn.Module):
    def __init__(self, dim_in, dim_out):
        super().__init__()
        self.proj = nn.Linear(dim_in, dim_out * 2)

    def forward(self, x):
        h, gate = self.proj(x).chunk(2, dim=-1)
        return h * F.gelu(gate)


class FeedForward(nn.Module):
    def __init__(self, dim, mult=4):
        super().__init__()
        self.net = nn.ModuleList([GEGLU(dim, dim * mult), nn.Dropout(0.0), nn.Linear(dim * mult, dim)])

    def forward(self, x):
        for m in self.net:
            x = m(x)
        return x


class BasicTransformerBlock(nn.Module):
    def __init__(self, dim, heads, dim_head, cross_dim):
        super().__init__()
        self.norm1 = nn.LayerNorm(dim)
        self.attn1 = Attention(dim, None, heads, dim_head)
        self.norm2 = nn.LayerNorm(dim)
        self.attn2 = Attention(dim, cross_dim, heads, dim_head)
        self.norm3 = nn.LayerNorm(dim)
        self.ff = FeedForward(dim)

    def forward(self, x, encoder_hidden_states=None):
        x = self.attn1(self.norm1(x)) + x
        x = self.attn2(self.norm2(x), encoder_hidden_states=encoder_hidden_states) + x
        x = self.ff(self.norm3(x)) + x
        return x


class Transformer2DModel(nn.Module):
    def __init__(self, heads, dim_head, in_channels, cross_dim, groups=32):
        super().__init__()
        inner = heads * dim_head
        self.norm = nn.GroupNorm(groups, in_channels, eps=1e-6)
        self.proj_in = nn.Conv2d(in_channels, inner, 1)
        self.transformer_blocks = nn.ModuleList([BasicTransformerBlock(inner, heads, dim_head, cross_dim)])
        self.proj_out = nn.Conv2d(inner, in_channels, 1)

    def forward(self, x, encoder_hidden_states=None):
        b, c, h, w = x.shape
        res = x
        x = self.proj_in(self.norm(x))
        x = x.permute(0, 2, 3, 1).reshape(b, h * w, -1)
        for blk in self.transformer_blocks:
            x = blk(x, encoder_hidden_states=encoder_hidden_states)
        x = x.reshape(b, h, w, -1).permute(0, 3, 1, 2).contiguous()
        return self.proj_out(x) + res


class ResnetBlock2D(nn.Module):
    """Attribute surface follows modules/utils/pnp_utils.py:136-185."""

    def __init__(self, cin, cout, temb_channels: Optional[int] = 1280, groups=32, eps=1e-5):
        super().__init__()
        self.norm1 = nn.GroupNorm(groups, cin, eps=eps)
        self.conv1 = nn.Conv2d(cin, cout, 3, padding=1)
        self.time_emb_proj = nn.Linear(temb_channels, cout) if temb_channels else None
        self.norm2 = nn.GroupNorm(groups, cout, eps=eps)
        self.dropout = nn.Dropout(0.0)
        self.conv2 = nn.Conv2d(cout, cout, 3, padding=1)
        self.nonlinearity = nn.SiLU()
        self.upsample = self.downsample = None
        self.time_embedding_norm = "default"
        self.output_scale_factor = 1.0
        self.conv_shortcut = nn.Conv2d(cin, cout, 1) if cin != cout else None

    def forward(self, input_tensor, temb=None, scale=1.0):
        h = self.conv1(self.nonlinearity(self.norm1(input_tensor)))
        if temb is not None and self.time_emb_proj is not None:
            h = h + self.time_emb_proj(self.nonlinearity(temb))[:, :, None, None]
        h = self.conv2(self.dropout(self.nonlinearity(self.norm2(h))))
        if self.conv_shortcut is not None:
            input_tensor = self.conv_shortcut(input_tensor)
        return (input_tensor + h) / self.output_scale_factor


class Downsample2D(nn.Module):
    def __init__(self, c, padding=1):
        super().__init__()
        self.padding = padding
        self.conv = nn.Conv2d(c, c, 3, stride=2, padding=padding)

    def forward(self, x):
        if self.padding == 0:  # VAE encoder: asymmetric pad
            x = F.pad(x, (0, 1, 0, 1))
        return self.conv(x)


class Upsample2D(nn.Module):
    def __init__(self, c):
        super().__init__()
        self.conv = nn.Conv2d(c, c, 3, padding=1)

    def forward(self, x):
        return self.conv(F.interpolate(x, scale_factor=2.0, mode="nearest"))


class CrossAttnDownBlock2D(nn.Module):
    def __init__(self, cin, cout, heads, cross_dim, add_downsample):
        super().__init__()
        self.resnets = nn.ModuleList([ResnetBlock2D(cin if i == 0 else cout, cout) for i in range(2)])
        self.attentions = nn.ModuleList([Transformer2DModel(heads, cout // heads, cout, cross_dim) for _ in range(2)])
        self.downsamplers = nn.ModuleList([Downsample2D(cout)]) if add_downsample else None

    def forward(self, x, temb, ctx):
        outs = ()
        for r, a in zip(self.resnets, self.attentions):
            x = a(r(x, temb), encoder_hidden_states=ctx)
            outs += (x,)
        if self.downsamplers is not None:
            x = self.downsamplers[0](x)
            outs += (x,)
        return x, outs


class DownBlock2D(nn.Module):
    def __init__(self, cin, cout, add_downsample):
        super().__init__()
        self.resnets = nn.ModuleList([ResnetBlock2D(cin if i == 0 else cout, cout) for i in range(2)])
        self.downsamplers = nn.ModuleList([Downsample2D(cout)]) if add_downsample else None

    def forward(self, x, temb, ctx=None):
        outs = ()
        for r in self.resnets:
            x = r(x, temb)
            outs += (x,)
        if self.downsamplers is not None:
            x = self.downsamplers[0](x)
            outs += (x,)
        return x, outs


class UNetMidBlock2DCrossAttn(nn.Module):
    def __init__(self, c, heads, cross_dim):
        super().__init__()
        self.resnets = nn.ModuleList([ResnetBlock2D(c, c), ResnetBlock2D(c, c)])
        self.attentions = nn.ModuleList([Transformer2DModel(heads, c // heads, c, cross_dim)])

    def forward(self, x, temb, ctx):
        x = self.resnets[0](x, temb)
        x = self.attentions[0](x, encoder_hidden_states=ctx)
        return self.resnets[1](x, temb)


class UpBlock2D(nn.Module):
    def __init__(self, cin, cout, prev, add_upsample):
        super().__init__()
        rs = []
        for i in range(3):
            skip = cin if i == 2 else cout
            rin = prev if i == 0 else cout
            rs.append(ResnetBlock2D(rin + skip, cout))
        self.resnets = nn.ModuleList(rs)
        self.upsamplers = nn.ModuleList([Upsample2D(cout)]) if add_upsample else None

    def forward(self, x, skips, temb, ctx=None):
        for r in self.resnets:
            x = r(torch.cat([x, skips.pop()], dim=1), temb)
        if self.upsamplers is not None:
            x = self.upsamplers[0](x)
        return x


class CrossAttnUpBlock2D(nn.Module):
    def __init__(self, cin, cout, prev, heads, cross_dim, add_upsample):
        super().__init__()
        rs = []
        for i in range(3):
            skip = cin if i == 2 else cout
            rin = prev if i == 0 else cout
            rs.append(ResnetBlock2D(rin + skip, cout))
        self.resnets = nn.ModuleList(rs)
        self.attentions = nn.ModuleList([Transformer2DModel(heads, cout // heads, cout, cross_dim) for _ in range(3)])
        self.upsamplers = nn.ModuleList([Upsample2D(cout)]) if add_upsample else None

    def forward(self, x, skips, temb, ctx):
        for r, a in zip(self.resnets, self.attentions):
            x = a(r(torch.cat([x, skips.pop()], dim=1), temb), encoder_hidden_states=ctx)
        if self.upsamplers is not None:
            x = self.upsamplers[0](x)
        return x


class UNet2DConditionOutput(dict):
    """``unet(...)["sample"]`` (diffusion_inversion.py:264) and ``.sample`` both work."""

    def __init__(self, sample):
        super().__init__(sample=sample)
        self.sample = sample


class UNet2DConditionModel(nn.Module):
    """SD-1.x UNet as diffusers 0.21.1 builds it."""

    def __init__(self, block_out_channels=(320, 640, 1280, 1280), heads=8, cross_attention_dim=768,
                 in_channels=4, out_channels=4):
        super().__init__()
        c = block_out_channels
        temb = c[0] * 4
        self.config = FrozenDict(in_channels=in_channels, out_channels=out_channels, sample_size=64,
                                 block_out_channels=tuple(c), attention_head_dim=heads,
                                 cross_attention_dim=cross_attention_dim)
        self.conv_in = nn.Conv2d(in_channels, c[0], 3, padding=1)
        self.time_embedding = TimestepEmbedding(c[0], temb)
        # ResnetBlock2D default temb_channels=1280 is the SD value; rebuild when channels differ
        self.down_blocks = nn.ModuleList()
        cout = c[0]
        for i in range(4):
            cin, cout = cout, c[i]
            last = i == 3
            if i < 3:
                self.down_blocks.append(CrossAttnDownBlock2D(cin, cout, heads, cross_attention_dim, not last))
            else:
                self.down_blocks.append(DownBlock2D(cin, cout, not last))
        self.mid_block = UNetMidBlock2DCrossAttn(c[-1], heads, cross_attention_dim)
        self.up_blocks = nn.ModuleList()
        rev = list(reversed(c))
        cout = rev[0]
        for i in range(4):
            prev, cout = cout, rev[i]
            cin = rev[min(i + 1, 3)]
            last = i == 3
            if i == 0:
                self.up_blocks.append(UpBlock2D(cin, cout, prev, not last))
            else:
                self.up_blocks.append(CrossAttnUpBlock2D(cin, cout, prev, heads, cross_attention_dim, not last))
        self.conv_norm_out = nn.GroupNorm(32, c[0], eps=1e-5)
        self.conv_act = nn.SiLU()
        self.conv_out = nn.Conv2d(c[0], out_channels, 3, padding=1)
        if temb != 1280:
            for m in self.modules():
                if isinstance(m, ResnetBlock2D):
                    m.time_emb_proj = nn.Linear(temb, m.conv1.out_channels)

    @property
    def dtype(self):
        return self.conv_in.weight.dtype

    @property
    def device(self):
        return self.conv_in.weight.device

    def forward(self, sample, timestep, encoder_hidden_states=None, **kwargs):
        t = timestep
        if not torch.is_tensor(t):
            t = torch.tensor([t], dtype=torch.int64, device=sample.device)
        elif t.ndim == 0:
            t = t[None].to(sample.device)
        t = t.expand(sample.shape[0])
        emb = self.time_embedding(timestep_embedding(t, self.conv_in.out_channels).to(sample.dtype))
        x = self.conv_in(sample)
        skips = [x]
        for blk in self.down_blocks:
            x, outs = blk(x, emb, encoder_hidden_states)
            skips += list(outs)
        x = self.mid_block(x, emb, encoder_hidden_states)
        for blk in self.up_blocks:
            x = blk(x, skips, emb, encoder_hidden_states)
        x = self.conv_out(self.conv_act(self.conv_norm_out(x)))
        return UNet2DConditionOutput(x)


# --------------------------------------------------------------------------------------
# AutoencoderKL (SD VAE)
# --------------------------------------------------------------------------------------
class _VaeMid(nn.Module):
    def __init__(self, c):
        super().__init__()
        self.resnets = nn.ModuleList([ResnetBlock2D(c, c, None, eps=1e-6), ResnetBlock2D(c, c, None, eps=1e-6)])
        self.attentions = nn.ModuleList([Attention(c, None, 1, c, bias=True, norm_num_groups=32, eps=1e-6,
                                                   residual_connection=True)])

    def forward(self, x):
        return self.resnets[1](self.attentions[0](self.resnets[0](x)))


class _EncBlock(nn.Module):
    def __init__(self, cin, cout, down):
        super().__init__()
        self.resnets = nn.ModuleList([ResnetBlock2D(cin if i == 0 else cout, cout, None, eps=1e-6) for i in range(2)])
        self.downsamplers = nn.ModuleList([Downsample2D(cout, padding=0)]) if down else None

    def forward(self, x):
        for r in self.resnets:
            x = r(x)
        if self.downsamplers is not None:
            x = self.downsamplers[0](x)
        return x


class _DecBlock(nn.Module):
    def __init__(self, cin, cout, up):
        super().__init__()
        self.resnets = nn.ModuleList([ResnetBlock2D(cin if i == 0 else cout, cout, None, eps=1e-6) for i in range(3)])
        self.upsamplers = nn.ModuleList([Upsample2D(cout)]) if up else None

    def forward(self, x):
        for r in self.resnets:
            x = r(x)
        if self.upsamplers is not None:
            x = self.upsamplers[0](x)
        return x


class Encoder(nn.Module):
    def __init__(self, ch=(128, 256, 512, 512)):
        super().__init__()
        self.conv_in = nn.Conv2d(3, ch[0], 3, padding=1)
        self.down_blocks = nn.ModuleList()
        cout = ch[0]
        for i, c in enumerate(ch):
            cin, cout = cout, c
            self.down_blocks.append(_EncBlock(cin, cout, i < len(ch) - 1))
        self.mid_block = _VaeMid(ch[-1])
        self.conv_norm_out = nn.GroupNorm(32, ch[-1], eps=1e-6)
        self.conv_act = nn.SiLU()
        self.conv_out = nn.Conv2d(ch[-1], 8, 3, padding=1)

    def forward(self, x):
        x = self.conv_in(x)
        for b in self.down_blocks:
            x = b(x)
        x = self.mid_block(x)
        return self.conv_out(self.conv_act(self.conv_norm_out(x)))


class Decoder(nn.Module):
    def __init__(self, ch=(128, 256, 512, 512)):
        super().__init__()
        rev = list(reversed(ch))
        self.conv_in = nn.Conv2d(4, rev[0], 3, padding=1)
        self.mid_block = _VaeMid(rev[0])
        self.up_blocks = nn.ModuleList()
        cout = rev[0]
        for i, c in enumerate(rev):
            cin, cout = cout, c
            self.up_blocks.append(_DecBlock(cin, cout, i < len(ch) - 1))
        self.conv_norm_out = nn.GroupNorm(32, ch[0], eps=1e-6)
        self.conv_act = nn.SiLU()
        self.conv_out = nn.Conv2d(ch[0], 3, 3, padding=1)

    def forward(self, z):
        x = self.mid_block(self.conv_in(z))
        for b in self.up_blocks:
            x = b(x)
        return self.conv_out(self.conv_act(self.conv_norm_out(x)))


class AutoencoderKL(nn.Module):
    def __init__(self, ch=(128, 256, 512, 512)):
        super().__init__()
        self.encoder = Encoder(ch)
        self.decoder = Decoder(ch)
        self.quant_conv = nn.Conv2d(8, 8, 1)
        self.post_quant_conv = nn.Conv2d(4, 4, 1)

    @property
    def dtype(self):
        return self.quant_conv.weight.dtype

    def encode(self, x):
        moments = self.quant_conv(self.encoder(x))
        mean, _ = moments.chunk(2, dim=1)
        return {"latent_dist": SimpleNamespace(mean=mean)}

    def decode(self, z):
        return {"sample": self.decoder(self.post_quant_conv(z))}


# --------------------------------------------------------------------------------------
# DDIM scheduler (diffusers 0.21.1 semantics)
# --------------------------------------------------------------------------------------
# diffusers: a dataclass whose pred_original_sample is optional (the reference's EDICT scheduler passes one field)
DDIMSchedulerOutput = namedtuple("DDIMSchedulerOutput", ("prev_sample", "pred_original_sample"), defaults=(None,))


class DDIMScheduler:
    _defaults = dict(num_train_timesteps=1000, beta_start=0.0001, beta_end=0.02, beta_schedule="linear",
                     trained_betas=None, clip_sample=True, set_alpha_to_one=True, steps_offset=0,
                     prediction_type="epsilon", thresholding=False, dynamic_thresholding_ratio=0.995,
                     clip_sample_range=1.0, sample_max_value=1.0, timestep_spacing="leading",
                     rescale_betas_zero_snr=False)

    def __init__(self, **kwargs):
        cfg = {**self._defaults, **{k: v for k, v in kwargs.items() if k in self._defaults}}
        self.config = FrozenDict(cfg)
        n = cfg["num_train_timesteps"]
        if cfg["beta_schedule"] == "linear":
            self.betas = torch.linspace(cfg["beta_start"], cfg["beta_end"], n, dtype=torch.float32)
        elif cfg["beta_schedule"] == "scaled_linear":
            self.betas = torch.linspace(cfg["beta_start"] ** 0.5, cfg["beta_end"] ** 0.5, n, dtype=torch.float32) ** 2
        else:  # pragma: no cover
            raise NotImplementedError(cfg["beta_schedule"])
        self.alphas = 1.0 - self.betas
        self.alphas_cumprod = torch.cumprod(self.alphas, dim=0)
        self.final_alpha_cumprod = torch.tensor(1.0) if cfg["set_alpha_to_one"] else self.alphas_cumprod[0]
        self.init_noise_sigma = 1.0
        self.num_inference_steps = None
        self.timesteps = torch.from_numpy(np.arange(0, n)[::-1].copy().astype(np.int64))

    @classmethod
    def from_config(cls, config, **kwargs):
        return cls(**{**dict(config), **kwargs})

    def set_timesteps(self, num_inference_steps, device=None):
        self.num_inference_steps = num_inference_steps
        assert self.config.timestep_spacing == "leading"
        ratio = self.config.num_train_timesteps // num_inference_steps
        ts = (np.arange(0, num_inference_steps) * ratio).round()[::-1].copy().astype(np.int64)
        ts += self.config.steps_offset
        self.timesteps = torch.from_numpy(ts)

    def _get_variance(self, timestep, prev_timestep):
        a_t = self.alphas_cumprod[timestep]
        a_p = self.alphas_cumprod[prev_timestep] if prev_timestep >= 0 else self.final_alpha_cumprod
        return ((1 - a_p) / (1 - a_t)) * (1 - a_t / a_p)

    def step(self, model_output, timestep, sample, eta=0.0, use_clipped_model_output=False,
             generator=None, variance_noise=None, return_dict=True):
        prev_timestep = timestep - self.config.num_train_timesteps // self.num_inference_steps
        a_t = self.alphas_cumprod[timestep]
        a_p = self.alphas_cumprod[prev_timestep] if prev_timestep >= 0 else self.final_alpha_cumprod
        beta_t = 1 - a_t
        assert self.config.prediction_type == "epsilon"
        x0 = (sample - beta_t ** 0.5 * model_output) / a_t ** 0.5
        pred_eps = model_output
        if self.config.clip_sample:
            x0 = x0.clamp(-self.config.clip_sample_range, self.config.clip_sample_range)
        variance = self._get_variance(timestep, prev_timestep)
        std_dev_t = eta * variance ** 0.5
        direction = (1 - a_p - std_dev_t ** 2) ** 0.5 * pred_eps
        prev = a_p ** 0.5 * x0 + direction
        if eta > 0:
            if variance_noise is None:
                variance_noise = torch.randn(model_output.shape, generator=generator, dtype=model_output.dtype)
            prev = prev + std_dev_t * variance_noise
        return DDIMSchedulerOutput(prev, x0)


class _Unsupported:
    def __init__(self, *a, **k):  # pragma: no cover
        raise NotImplementedError("out of scope for the hot path (SURVEY.md section 2)")

    @classmethod
    def from_config(cls, *a, **k):  # pragma: no cover
        raise NotImplementedError("out of scope for the hot path (SURVEY.md section 2)")


class DDPMScheduler(_Unsupported):
    pass


class DPMSolverMultistepScheduler(_Unsupported):
    pass


class DPMSolverMultistepInverseScheduler(_Unsupported):
    pass


class Pix2PixZeroL2Loss(_Unsupported):
    pass


# --------------------------------------------------------------------------------------
# tokenizer / text encoder / pipeline container
# --------------------------------------------------------------------------------------
class SyntheticTokenizer:
    """Deterministic whitespace tokenizer: one id per word (crc32), BOS=49406, EOS/pad=49407.

    Surface used by the reference: __call__(...).input_ids, model_max_length, encode, decode
    (diffusion_inversion.py:223-241, ptp_utils.py:305-323, seq_aligner.py:113-166)."""

    bos, eos = 49406, 49407
    model_max_length = 77

    def __init__(self):
        self._words = {}

    def _id(self, w):
        i = zlib.crc32(w.encode()) % 49000 + 1
        self._words[i] = w
        return i

    def encode(self, text):
        words = [w for w in text.split(" ") if w != ""]
        return [self.bos] + [self._id(w) for w in words] + [self.eos]

    def decode(self, ids):
        out = []
        for i in ids:
            i = int(i)
            out.append("<|startoftext|>" if i == self.bos else "<|endoftext|>" if i == self.eos else self._words.get(i, "?"))
        return " ".join(out)

    def __call__(self, texts, padding="max_length", max_length=77, truncation=True, return_tensors="pt"):
        if isinstance(texts, str):  # pnp.py:14-20 passes a bare string
            texts = [texts]
        rows = []
        for t in texts:
            ids = self.encode(t)[:max_length]
            ids[-1] = self.eos if len(ids) == max_length else ids[-1]
            rows.append(ids + [self.eos] * (max_length - len(ids)))
        return SimpleNamespace(input_ids=torch.tensor(rows, dtype=torch.int64))


def make_text_encoder(seed: int = 0):
    """Random-init CLIP ViT-L/14 text tower (transformers); the synthetic init is an INPUT definition shared with the
    product (eta_inversion_b200/synthetic.py), like the random UNet / VAE state dicts."""
    from eta_inversion_b200.synthetic import make_text_encoder as mk
    return mk(seed)


class StableDiffusionPipeline:
    """Container with the attributes the reference touches."""

    def __init__(self, unet, vae, text_encoder, tokenizer, scheduler, device="cpu"):
        self.unet, self.vae, self.text_encoder, self.tokenizer, self.scheduler = unet, vae, text_encoder, tokenizer, scheduler
        self.device = torch.device(device)

    def to(self, device):
        self.device = torch.device(device)
        for m in (self.unet, self.vae, self.text_encoder):
            m.to(device)
        return self

    @classmethod
    def from_pretrained(cls, *a, **k):  # pragma: no cover
        raise RuntimeError("no pretrained weights in this environment; use oracle.build.build_pipeline()")


def sd_scheduler():
    """Scheduler exactly as modules/models/__init__.py:134 builds it, plus steps_offset=1 that the SD-1.x
    pipeline config carries ('Assumption to pin')."""
    return DDIMScheduler(beta_start=0.00085, beta_end=0.012, beta_schedule="scaled_linear", clip_sample=False,
                         set_alpha_to_one=False, steps_offset=1)


def build_pipeline(unet_sd=None, vae_sd=None, seed=0, dtype=torch.float32, with_vae=True, with_text=True):
    """Assemble the oracle pipeline from externally supplied state dicts (diffusers key names)."""
    unet = UNet2DConditionModel().to(dtype).eval().requires_grad_(False)
    if unet_sd is not None:
        unet.load_state_dict(unet_sd, strict=True)
    vae = None
    if with_vae:
        vae = AutoencoderKL().to(dtype).eval().requires_grad_(False)
        if vae_sd is not None:
            vae.load_state_dict(vae_sd, strict=True)
    te = make_text_encoder(seed) if with_text else None
    return StableDiffusionPipeline(unet, vae, te, SyntheticTokenizer(), sd_scheduler())
