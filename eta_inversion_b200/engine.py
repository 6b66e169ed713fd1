"""Thin Python face of the native engine: the UNet handle, the attention-control descriptor and the
individually exported ops.  Everything here forwards to libetai.so through ``_lib`` (ctypes, C ABI);
tensors are torch CUDA tensors used as device memory only.

Call shape kept from the reference:  ``unet(sample, t, encoder_hidden_states=ctx)["sample"]``
(modules/inversion/diffusion_inversion.py:264-280, modules/inversion/eta_inversion.py:321).
"""
from __future__ import annotations

import ctypes as C
import threading
from dataclasses import dataclass, field
from typing import Dict, List, Optional, Sequence

import torch

from . import _lib
from ._lib import (CTRL_CROSS_EDIT, CTRL_CROSS_STORE, CTRL_SELF_REMAP, MATH_AUTO, MATH_SIMT, EtaiAttnCtrl, EtaiTensor,
                   EtaiUnetCfg, check, dtype_code, i32_array, ptr, stream_ptr)

SD15_CHANNELS = (320, 640, 1280, 1280)
LAUNCHES = [0]  # kernels launched through the op-level entry points below (the UNet handle counts its own)
_LAUNCHES_LOCK = threading.Lock()  # lanes of lock-step / pipelined groups call the ops from several threads


def _count_launches(n: int) -> None:
    with _LAUNCHES_LOCK:
        LAUNCHES[0] += n


def _require_cuda(t: torch.Tensor, name: str) -> None:
    if not (isinstance(t, torch.Tensor) and t.is_cuda and t.is_contiguous()):
        raise RuntimeError(f"etai: {name} must be a contiguous CUDA tensor (no CPU fallback exists)")


@dataclass
class AttnControl:
    """Data form of the reference's attention hooks for ONE UNet forward (see include/etai.h, etai_attn_ctrl).

    self remap    <- ptp.py:194-200 (self replace), masactrl.py:56-72, pnp_utils.py:76-88
    cross edit    <- ptp.py:205-211,234-274 (replace / refine / reweight)
    cross store   <- ptp.py:150-171 (AttentionStore), consumed by LocalBlend and get_attention_map
    conv inject   <- pnp_utils.py:172-177
    """
    self_rows: Optional[Sequence[Sequence[int]]] = None  # (q_row[B], k_row[B], v_row[B])
    self_layer_mask: int = 0xFFFF
    self_max_tokens: int = 1 << 30
    edit_pairs: Optional[Sequence[Sequence[int]]] = None  # [(base_row, tgt_row), ...]
    mapper: Optional[torch.Tensor] = None       # [P,77,77] fp32 cuda
    blend_a: Optional[torch.Tensor] = None      # [P,77]
    equalizer: Optional[torch.Tensor] = None    # [P,77]
    alpha_step: Optional[torch.Tensor] = None   # [P,77]
    store_rows: Optional[Sequence[int]] = None
    store_res: int = 16
    store_down: Optional[torch.Tensor] = None   # [len(store_rows), res*res, 77] fp32 cuda, accumulated in place
    store_mid: Optional[torch.Tensor] = None
    store_up: Optional[torch.Tensor] = None
    conv_inject_rows: int = 0
    _keep: list = field(default_factory=list, repr=False)

    def drop_leading_rows(self, n: int) -> Optional["AttnControl"]:
        """The same control for a batch whose first ``n`` rows were removed (rows renumbered), or None if the control
        reads or writes one of those rows.  Used when the unconditional half of a CFG batch has weight exactly 0 and is
        not computed (EtaInversion, guidance_scale_fwd == 1)."""
        if self.conv_inject_rows:
            return None
        out = AttnControl(self_layer_mask=self.self_layer_mask, self_max_tokens=self.self_max_tokens, store_res=self.store_res,
                          mapper=self.mapper, blend_a=self.blend_a, equalizer=self.equalizer, alpha_step=self.alpha_step,
                          store_down=self.store_down, store_mid=self.store_mid, store_up=self.store_up)
        if self.self_rows is not None:
            kept = [r[n:] for r in self.self_rows]
            if any(v < n for r in kept for v in r):
                return None
            out.self_rows = tuple([v - n for v in r] for r in kept)
        if self.edit_pairs is not None:
            if any(b < n or t < n for b, t in self.edit_pairs):
                return None
            out.edit_pairs = [(b - n, t - n) for b, t in self.edit_pairs]
        if self.store_rows is not None:
            if any(r < n for r in self.store_rows):
                return None
            out.store_rows = [r - n for r in self.store_rows]
        return out

    def to_struct(self) -> EtaiAttnCtrl:
        s = EtaiAttnCtrl()
        self._keep = []
        flags = 0
        if self.self_rows is not None:
            flags |= CTRL_SELF_REMAP
            arrs = [i32_array(r) for r in self.self_rows]
            self._keep += arrs
            s.self_q_row, s.self_k_row, s.self_v_row = [C.cast(a, C.POINTER(C.c_int32)) for a in arrs]
            s.self_layer_mask = self.self_layer_mask & 0xFFFFFFFF
            s.self_max_tokens = int(self.self_max_tokens)
        if self.edit_pairs is not None:
            flags |= CTRL_CROSS_EDIT
            P = len(self.edit_pairs)
            for name in ("mapper", "blend_a", "equalizer", "alpha_step"):
                t = getattr(self, name)
                _require_cuda(t, name)
                if t.dtype != torch.float32 or t.shape[0] != P:
                    raise RuntimeError(f"etai: {name} must be fp32 with leading dim {P}")
            b, t_ = i32_array([p[0] for p in self.edit_pairs]), i32_array([p[1] for p in self.edit_pairs])
            self._keep += [b, t_]
            s.n_pairs = P
            s.edit_base_row, s.edit_tgt_row = C.cast(b, C.POINTER(C.c_int32)), C.cast(t_, C.POINTER(C.c_int32))
            s.mapper, s.blend_a, s.equalizer, s.alpha_step = (ptr(self.mapper), ptr(self.blend_a), ptr(self.equalizer),
                                                              ptr(self.alpha_step))
        if self.store_rows is not None:
            flags |= CTRL_CROSS_STORE
            r = i32_array(self.store_rows)
            self._keep.append(r)
            s.store_res, s.n_store_rows, s.store_row = int(self.store_res), len(self.store_rows), C.cast(r, C.POINTER(C.c_int32))
            for name in ("store_down", "store_mid", "store_up"):
                t = getattr(self, name)
                if t is not None:
                    _require_cuda(t, name)
                    if t.dtype != torch.float32 or tuple(t.shape) != (len(self.store_rows), self.store_res ** 2, 77):
                        raise RuntimeError(f"etai: {name} must be fp32 [{len(self.store_rows)},{self.store_res ** 2},77]")
                setattr(s, name, ptr(t))
        s.conv_inject_rows = int(self.conv_inject_rows)
        s.flags = flags
        return s


class UNetOutput(dict):
    """``unet(...)["sample"]`` and ``.sample`` (mirrors diffusers' UNet2DConditionOutput)."""

    def __init__(self, sample):
        super().__init__(sample=sample)
        self.sample = sample


class UNetEngine:
    """Handle on the native SD-1.x UNet.  Drop-in for ``pipe.unet`` in the reference's loops."""

    def __init__(self, state_dict: Dict[str, torch.Tensor], dtype: torch.dtype = torch.float32, device: int | str = 0,
                 max_batch: int = 4, channels: Sequence[int] = SD15_CHANNELS, heads: int = 8, cross_dim: int = 768,
                 ctx_len: int = 77, latent_hw: int = 64, math_mode: int = MATH_AUTO):
        if not torch.cuda.is_available():
            raise RuntimeError("etai: a CUDA device is required (no CPU fallback exists)")
        lib = _lib.load()
        dev = device if isinstance(device, torch.device) else torch.device(device if isinstance(device, str) else f"cuda:{device}")
        if dev.index is None:
            dev = torch.device("cuda", torch.cuda.current_device())
        self.device = dev
        self.dtype = dtype
        self.max_batch = max_batch
        self.latent_hw = latent_hw
        self.ctx_len, self.cross_dim = ctx_len, cross_dim
        self.control: Optional[AttnControl] = None  # installed by controllers for the next forward(s)
        cfg = EtaiUnetCfg()
        cfg.dtype, cfg.math_mode = dtype_code(dtype), math_mode
        for i, c in enumerate(channels):
            cfg.block_out_channels[i] = c
        cfg.heads, cfg.cross_dim, cfg.ctx_len, cfg.latent_hw, cfg.max_batch = heads, cross_dim, ctx_len, latent_hw, max_batch
        arr, keep = _lib.tensor_table(state_dict, dtype)  # plain tensors converted to the storage dtype on the host
        h = C.c_void_p()
        with torch.cuda.device(dev):
            check(lib.etai_unet_create(C.byref(h), C.byref(cfg), arr, len(state_dict), dev.index or 0))
        self._h = h
        self._lib = lib
        self._ctx_key = None
        self._timing = None  # list of (start, end) torch events when forward timing is switched on

    def clone(self, max_batch: Optional[int] = None) -> "UNetEngine":
        """A second engine on the SAME packed device weights (reference counted in the library) with its own activation
        arena, CUDA graphs, staging buffers and stream: what ``batching.run_pipelined`` runs its second group on."""
        other = object.__new__(UNetEngine)
        other.__dict__.update({k: v for k, v in self.__dict__.items() if k not in ("_h", "_ctx_key", "_timing", "control")})
        other.max_batch = max_batch or self.max_batch
        other.control, other._ctx_key, other._timing = None, None, None
        h = C.c_void_p()
        with torch.cuda.device(self.device):
            check(self._lib.etai_unet_clone(C.byref(h), self._h, int(other.max_batch)))
        other._h = h
        return other

    def close(self):
        if getattr(self, "_h", None):
            self._lib.etai_unet_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    @property
    def device_bytes(self) -> int:
        return int(self._lib.etai_unet_device_bytes(self._h))

    PROF_CATEGORIES = ("conv3x3", "gemm", "self_attn", "cross_attn", "groupnorm", "layernorm", "other")

    @property
    def launch_count(self) -> int:
        return int(self._lib.etai_unet_launch_count(self._h))

    def profile(self, enable: bool) -> Dict[str, Dict[str, float]]:
        """Read+clear the per-category device time recorded since the last call, then switch recording on/off."""
        ms, cnt = (C.c_float * 7)(), (C.c_int32 * 7)()
        check(self._lib.etai_unet_profile(self._h, int(enable), ms, cnt))
        return {n: {"ms": float(ms[i]), "launches": int(cnt[i])} for i, n in enumerate(self.PROF_CATEGORIES)}

    def time_forwards(self, enable: bool) -> float:
        """Device time (ms) spent in forward() calls since timing was switched on (CUDA events on the caller's stream,
        works with graph replay); then switch timing on/off."""
        ms = 0.0
        if self._timing:
            torch.cuda.synchronize(self.device)
            ms = sum(a.elapsed_time(b) for a, b in self._timing)
        self._timing = [] if enable else None
        return ms

    def set_context(self, ctx: torch.Tensor) -> None:
        _require_cuda(ctx, "encoder_hidden_states")
        if ctx.ndim != 3 or ctx.shape[1] != self.ctx_len or ctx.shape[2] != self.cross_dim:
            raise RuntimeError(f"etai: context must be [B,{self.ctx_len},{self.cross_dim}], got {tuple(ctx.shape)}")
        with torch.cuda.device(self.device):
            check(self._lib.etai_unet_set_context(self._h, ptr(ctx), dtype_code(ctx.dtype), ctx.shape[0], stream_ptr()))
        # keep a strong reference: comparing (data_ptr, version) alone is unsafe, a freed context's address can be handed
        # to the next prompt's context by the caching allocator
        self._ctx_key = (ctx, ctx._version)

    def forward(self, sample: torch.Tensor, timestep, encoder_hidden_states: Optional[torch.Tensor] = None,
                control: Optional[AttnControl] = None, **kwargs) -> UNetOutput:
        _require_cuda(sample, "sample")
        B = sample.shape[0]
        if tuple(sample.shape[1:]) != (4, self.latent_hw, self.latent_hw):
            raise RuntimeError(f"etai: sample must be [B,4,{self.latent_hw},{self.latent_hw}], got {tuple(sample.shape)}")
        if encoder_hidden_states is not None:
            ctx = encoder_hidden_states if encoder_hidden_states.is_contiguous() else encoder_hidden_states.contiguous()
            k = self._ctx_key
            if k is None or k[0] is not ctx or k[1] != ctx._version:  # constant over a loop: re-project only on change
                self.set_context(ctx)
        # diffusers takes a scalar or a [B] tensor of timesteps; the reference's loops only ever pass scalars
        t_rows = None
        if (torch.is_tensor(timestep) and timestep.numel() > 1) or isinstance(timestep, (list, tuple)):
            t_rows = [float(v) for v in (timestep.flatten().tolist() if torch.is_tensor(timestep) else timestep)]
            if len(t_rows) != B:
                raise RuntimeError(f"etai: {len(t_rows)} timesteps for {B} rows")
            t = t_rows[0]
        else:
            t = float(timestep.item() if torch.is_tensor(timestep) else timestep)
        ctrl = control if control is not None else self.control
        out = torch.empty_like(sample)
        cs = ctrl.to_struct() if ctrl is not None else None
        with torch.cuda.device(self.device):
            if self._timing is not None:
                ev = (torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True))
                ev[0].record()
            if t_rows is not None:
                check(self._lib.etai_unet_forward_rows(self._h, ptr(sample), (C.c_float * B)(*t_rows), dtype_code(sample.dtype), B,
                                                       C.byref(cs) if cs is not None else None, ptr(out), stream_ptr()))
            else:
                check(self._lib.etai_unet_forward(self._h, ptr(sample), t, dtype_code(sample.dtype), B,
                                                  C.byref(cs) if cs is not None else None, ptr(out), stream_ptr()))
            if self._timing is not None:
                ev[1].record()
                self._timing.append(ev)
        return UNetOutput(out)

    __call__ = forward

    # ---- gradient w.r.t. the text context (null-text inversion) -------------------------------------------------------
    def enable_backward(self, max_batch: int = 1) -> None:
        with torch.cuda.device(self.device):
            check(self._lib.etai_unet_enable_backward(self._h, int(max_batch)))

    def forward_train(self, sample: torch.Tensor, timestep, encoder_hidden_states: torch.Tensor) -> torch.Tensor:
        """``unet(sample, t, ctx)["sample"]`` keeping the activations for :meth:`backward_ctx` (no attention control)."""
        _require_cuda(sample, "sample")
        ctx = encoder_hidden_states.detach()
        ctx = ctx if ctx.is_contiguous() else ctx.contiguous()
        self.set_context(ctx)  # the context changes every optimisation step: always re-project K/V
        self._ctx_key = None
        out = torch.empty_like(sample)
        t = float(timestep.item() if torch.is_tensor(timestep) else timestep)
        with torch.cuda.device(self.device):
            check(self._lib.etai_unet_forward_train(self._h, ptr(sample), t, dtype_code(sample.dtype), sample.shape[0], ptr(out),
                                                    stream_ptr()))
        return out

    def backward_ctx(self, d_eps: torch.Tensor) -> torch.Tensor:
        """d(ctx) [B,77,768] fp32 for the train-mode forward made just before; ``d_eps``: dL/d(eps), [B,4,hw,hw]."""
        d = d_eps.detach().float().contiguous()
        _require_cuda(d, "d_eps")
        out = torch.empty((d.shape[0], self.ctx_len, self.cross_dim), dtype=torch.float32, device=d.device)
        with torch.cuda.device(self.device):
            check(self._lib.etai_unet_backward_ctx(self._h, ptr(d), d.shape[0], ptr(out), stream_ptr()))
        return out


# ------------------------------------------------------------------------------------------------
# single ops (unit parity + roofline runs)
# ------------------------------------------------------------------------------------------------
def gemm(A: torch.Tensor, W: torch.Tensor, bias=None, residual=None, geglu: bool = False, math_mode: int = MATH_AUTO):
    for n, t in (("A", A), ("W", W)):
        _require_cuda(t, n)
    M, K = A.shape
    N = W.shape[0]
    out = torch.empty((M, N // 2 if geglu else N), dtype=A.dtype, device=A.device)
    check(_lib.load().etai_gemm(ptr(A), ptr(W), ptr(bias), ptr(residual), ptr(out), M, N, K, int(geglu),
                                dtype_code(A.dtype), math_mode, stream_ptr()))
    return out


def conv3x3(x: torch.Tensor, w_packed: torch.Tensor, bias=None, residual=None, stride: int = 1,
            math_mode: int = MATH_AUTO):
    """x: NHWC [B,H,W,Ci]; w_packed: [Co,3,3,Ci]."""
    _require_cuda(x, "x")
    _require_cuda(w_packed, "w")
    B, H, Wd, Ci = x.shape
    Co = w_packed.shape[0]
    Ho, Wo = (H - 1) // stride + 1, (Wd - 1) // stride + 1
    out = torch.empty((B, Ho, Wo, Co), dtype=x.dtype, device=x.device)
    # workspace: im2col for stride 2 plus room for split-K partials of small problems
    ws_elems = (B * Ho * Wo * 9 * Ci if stride == 2 else 0) + (16 * B * Ho * Wo * Co * 2 if B * Ho * Wo <= 2048 else 0)
    ws = torch.empty((max(ws_elems, 1) + 256,), dtype=x.dtype, device=x.device)
    check(_lib.load().etai_conv3x3(ptr(x), ptr(w_packed), ptr(bias), ptr(residual), ptr(out), B, H, Wd, Ci, Co, stride,
                                   dtype_code(x.dtype), math_mode, ptr(ws), 0 if ws is None else ws.numel() * ws.element_size(),
                                   stream_ptr()))
    return out


def cross_attention(q: torch.Tensor, kv: torch.Tensor, heads: int, koff: int, voff: int, scale: Optional[float] = None,
                    edit_pairs=None, mapper=None, blend_a=None, equalizer=None, alpha_step=None, store_rows=None,
                    store_acc=None, math_mode: int = MATH_AUTO) -> torch.Tensor:
    """One control-aware cross-attention layer (include/etai.h, etai_cross_attention).
    q: [B,N,heads*d]; kv: [B,L,ldkv] with K of head h at column koff+h*d and V at voff+h*d."""
    _require_cuda(q, "q")
    _require_cuda(kv, "kv")
    B, N, Cc = q.shape
    L, d = kv.shape[1], Cc // heads
    out = torch.empty((B, N, Cc), dtype=q.dtype, device=q.device)
    P = 0 if edit_pairs is None else len(edit_pairs)
    if P:
        for n, t in (("mapper", mapper), ("blend_a", blend_a), ("equalizer", equalizer), ("alpha_step", alpha_step)):
            _require_cuda(t, n)
            if t.dtype != torch.float32 or t.shape[0] != P:
                raise RuntimeError(f"etai: {n} must be fp32 with leading dim {P}")
    b = i32_array([p[0] for p in edit_pairs]) if P else None
    t_ = i32_array([p[1] for p in edit_pairs]) if P else None
    S = 0 if store_rows is None else len(store_rows)
    sr = i32_array(store_rows) if S else None
    if S:
        _require_cuda(store_acc, "store_acc")
        if store_acc.dtype != torch.float32 or tuple(store_acc.shape) != (S, N, L):
            raise RuntimeError(f"etai: store_acc must be fp32 [{S},{N},{L}]")
    ws = torch.empty((327680 + heads * max(S, 1) * N * L * 4,), dtype=torch.uint8, device=q.device)
    i32p = C.POINTER(C.c_int32)
    check(_lib.load().etai_cross_attention(
        ptr(q), ptr(kv), ptr(out), B, N, L, heads, d, q.stride(1), kv.stride(1), Cc, koff, voff,
        float(scale if scale is not None else d ** -0.5), P, C.cast(b, i32p) if P else None, C.cast(t_, i32p) if P else None,
        ptr(mapper), ptr(blend_a), ptr(equalizer), ptr(alpha_step), S, C.cast(sr, i32p) if S else None, ptr(store_acc),
        dtype_code(q.dtype), math_mode, ptr(ws), ws.numel(), stream_ptr()))
    return out


def h2d(t: torch.Tensor, device) -> torch.Tensor:
    """Small host tensor -> device without stalling the host: a pageable `.to(device)` waits until the stream has
    drained (every queued UNet forward), which leaves the GPU idle afterwards while the host catches up; a pinned
    staging copy is asynchronous (the caching host allocator keeps the staging block alive until the copy has run)."""
    if t.device.type != "cpu":
        return t.to(device)
    return t.pin_memory().to(device, non_blocking=True)


def groupnorm(x: torch.Tensor, gamma, beta, groups: int = 32, eps: float = 1e-5, silu: bool = False):
    """x: NHWC [B,HW,C]."""
    _require_cuda(x, "x")
    B, HW, Cc = x.shape
    out = torch.empty_like(x)
    ws = torch.empty((B * 256 * groups * 2 + 4200,), dtype=torch.float64, device=x.device)
    check(_lib.load().etai_groupnorm(ptr(x), ptr(out), ptr(gamma), ptr(beta), B, HW, Cc, groups, eps, int(silu),
                                     dtype_code(x.dtype), ptr(ws), ws.numel() * 8, stream_ptr()))
    return out


def layernorm(x: torch.Tensor, gamma, beta, eps: float = 1e-5):
    _require_cuda(x, "x")
    M, Cc = x.shape
    out = torch.empty_like(x)
    check(_lib.load().etai_layernorm(ptr(x), ptr(out), ptr(gamma), ptr(beta), M, Cc, eps, dtype_code(x.dtype), stream_ptr()))
    return out


def attention(q, k, v, heads: int, scale: Optional[float] = None, rows=None, math_mode: int = MATH_AUTO):
    """q: [B,Nq,heads*d], k,v: [B,Nk,heads*d] (may be strided views of a fused qkv buffer)."""
    B, Nq, Cc = q.shape
    Nk = k.shape[1]
    d = Cc // heads
    for t in (q, k, v):
        if not t.is_cuda or t.stride(2) != 1 or t.stride(0) != t.shape[1] * t.stride(1):
            raise RuntimeError("etai: attention operands must be CUDA tensors with contiguous heads and dense batch stride")
    out = torch.empty((B, Nq, Cc), dtype=q.dtype, device=q.device)
    maps = [None, None, None] if rows is None else [i32_array(r) for r in rows]
    cast = [None if m is None else C.cast(m, C.POINTER(C.c_int32)) for m in maps]
    check(_lib.load().etai_attention(ptr(q), ptr(k), ptr(v), ptr(out), B, Nq, Nk, heads, d, q.stride(1), k.stride(1),
                                     v.stride(1), Cc, float(scale if scale is not None else d ** -0.5), cast[0], cast[1],
                                     cast[2], dtype_code(q.dtype), math_mode, stream_ptr()))
    return out


def cfg_ddim_step(eps, x, a_from: float, a_to: float, guidance: Optional[float] = None, eta: float = 0.0,
                  variance: float = 0.0, eta_map=None, noise_cand=None, losses=None, pin_src=None, want_eps=False):
    """Fused CFG + (eta-)DDIM step; see include/etai.h etai_cfg_ddim_step."""
    for n, t in (("eps", eps), ("x", x)):
        _require_cuda(t, n)
        if t.dtype != torch.float32:
            raise RuntimeError("etai: scheduler tensors are fp32")
    n = x.shape[0]
    E = x[0].numel()
    has_cfg = guidance is not None
    if eps.shape[0] != (2 * n if has_cfg else n):
        raise RuntimeError(f"etai: eps rows {eps.shape[0]} do not match {n} latents (cfg={has_cfg})")
    out = torch.empty_like(x)
    eps_out = torch.empty_like(x) if want_eps else None
    K = 0 if noise_cand is None else noise_cand.shape[0]
    check(_lib.load().etai_cfg_ddim_step(ptr(eps), n, int(has_cfg), float(guidance or 0.0), ptr(x), ptr(out), ptr(eps_out),
                                         float(a_from), float(a_to), float(eta), float(variance), ptr(eta_map),
                                         ptr(noise_cand), ptr(losses), K, ptr(pin_src), E, stream_ptr()))
    _count_launches(1)
    return (out, eps_out) if want_eps else out


def prox_guidance(eps_u: torch.Tensor, eps_c: torch.Tensor, guidance: float, quantile: float, l1: bool = False,
                  want_thr: bool = False):
    """eps_u + guidance * prox(eps_c - eps_u), the soft threshold taken at ``quantile`` of |eps_c - eps_u| over the whole
    tensor (torch.quantile's 'linear' rule) or, for ``quantile <= 0``, at the fixed value ``-quantile``; one launch
    (include/etai.h etai_prox_guidance; reference: proximal_negative_prompt_inversion.py:61-128)."""
    for n_, t in (("eps_u", eps_u), ("eps_c", eps_c)):
        _require_cuda(t, n_)
        if t.dtype != torch.float32 or not t.is_contiguous():
            raise RuntimeError("etai: prox_guidance takes contiguous fp32 tensors")
    if eps_u.shape != eps_c.shape:
        raise RuntimeError("etai: prox_guidance: shape mismatch")
    n = eps_u.numel()
    lo, hi, w, fixed = -1, -1, 0.0, 0.0
    if quantile > 0:
        lo, hi, w = quantile_ranks(quantile, n)
    else:
        fixed = -float(quantile)
    out = torch.empty_like(eps_u)
    thr = torch.empty((1,), dtype=torch.float32, device=eps_u.device) if want_thr else None
    check(_lib.load().etai_prox_guidance(ptr(eps_u), ptr(eps_c), ptr(out), n, lo, hi, w, fixed, int(bool(l1)), float(guidance),
                                         ptr(thr), stream_ptr()))
    _count_launches(1)
    return (out, thr) if want_thr else out


def quantile_ranks(q: float, n: int):
    """(rank_lo, rank_hi, weight) of torch.quantile(x, q) with 'linear' interpolation over n elements: torch multiplies a
    float32 q by the last index in float32 and interpolates between floor and ceil of that rank (ATen quantile_compute)."""
    import numpy as np
    if not 0.0 <= q <= 1.0:
        raise RuntimeError("etai: quantile must be in [0, 1]")
    r = np.float32(q) * np.float32(n - 1)
    lo, hi = int(np.floor(r)), int(np.ceil(r))
    return lo, hi, float(np.float32(r) - np.float32(lo))


def eta_noise_losses(eps, x, x_prev_inv, a_from: float, a_to: float, guidance: Optional[float], eta: float,
                     variance: float, noise_cand: torch.Tensor):
    n = x.shape[0]
    E = x[0].numel()
    K = noise_cand.shape[0]
    losses = torch.empty((K,), dtype=torch.float32, device=x.device)
    best = torch.empty((1,), dtype=torch.int32, device=x.device)
    check(_lib.load().etai_eta_noise_losses(ptr(eps), n, int(guidance is not None), float(guidance or 0.0), ptr(x),
                                            ptr(x_prev_inv), float(a_from), float(a_to), float(eta), float(variance),
                                            ptr(noise_cand), K, E, ptr(losses), ptr(best), stream_ptr()))
    _count_launches(2)
    return losses, best
