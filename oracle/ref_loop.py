"""ORACLE (test infrastructure, never shipped, never on the product path).

Self-contained CPU/fp32 restatement, in plain PyTorch, of the reference's hot-path LOOPS so that the checker and the
CPU baseline can run where /root/reference does not exist (the GPU box).  It materialises every attention
probability tensor and calls the controller on it exactly like the reference's monkey-patched forward does.

Follows (reference file:line):
  attention hook            modules/utils/ptp_utils.py:205-260  (explicit softmax(QK^T*scale), controller, P.V)
  AttentionControl counters modules/utils/ptp.py:107-119        (cond half only, cur_att_layer / cur_step)
  AttentionStore            modules/utils/ptp.py:150-171
  AttentionControlEdit      modules/utils/ptp.py:194-218 ; Replace/Refine/Reweight :234-274
  LocalBlend                modules/utils/ptp.py:18-47
  aggregate / word maps     modules/utils/ptp.py:288-303, modules/editing/ptp_editor.py:43-85
  token alignment           modules/utils/seq_aligner.py:52-201
  DDIM inverse step         modules/inverse_schedulers/scheduling_ddim_inverse.py:71-143 (mode "sameshift")
  base loops / CFG          modules/inversion/diffusion_inversion.py:249-286,314-436,462-528
  eta inversion             modules/inversion/eta_inversion.py:145-403
  MasaCtrl                  modules/utils/masactrl.py:41-72, modules/utils/masactrl_utils.py:18-31
  editors                   modules/editing/editor.py:67-118, simple_editor.py:27-51, masactrl_editor.py:44-69

PINNED: tests/test_oracle.py checks this file against tests/golden/*.npz, which were produced by running the
reference's own unmodified modules/** over oracle/sd15.py (oracle/run_reference.py).  The diffusers arithmetic
underneath (oracle/sd15.py) remains "parity unpinned" against diffusers itself (see its header).
"""
from __future__ import annotations

import math
from typing import Dict, List, Optional, Sequence

import numpy as np
import torch
import torch.nn.functional as F

L = 77


# ------------------------------------------------------------------------------------------------
# token helpers (one token per word with the synthetic tokenizer, but written for the general case)
# ------------------------------------------------------------------------------------------------
def word_inds(text: str, word, tok) -> np.ndarray:
    words = text.split(" ")
    want = [i for i, w in enumerate(words) if w == word] if isinstance(word, str) else [word]
    out = []
    if want:
        pieces = [tok.decode([i]).strip("#") for i in tok.encode(text)][1:-1]
        cur, ptr = 0, 0
        for i, p in enumerate(pieces):
            cur += len(p)
            if ptr in want:
                out.append(i + 1)
            if cur >= len(words[ptr]):
                ptr, cur = ptr + 1, 0
    return np.array(out)


def replacement_mapper(src: str, tgt: str, tok) -> torch.Tensor:
    ws, wt = src.split(" "), tgt.split(" ")
    assert len(ws) == len(wt)
    diff = [i for i in range(len(wt)) if wt[i] != ws[i]]
    s_inds = [word_inds(src, i, tok) for i in diff]
    t_inds = [word_inds(tgt, i, tok) for i in diff]
    m = np.zeros((L, L))
    i = j = c = 0
    while i < L and j < L:
        if c < len(s_inds) and s_inds[c][0] == i:
            a, b = s_inds[c], t_inds[c]
            if len(a) == len(b):
                m[a, b] = 1
            else:
                for t in b:
                    m[a, t] = 1 / len(b)
            c += 1
            i += len(a)
            j += len(b)
        elif c < len(s_inds):
            m[i, j] = 1
            i += 1
            j += 1
        else:
            m[j, j] = 1
            i += 1
            j += 1
    return torch.from_numpy(m).float()[None]


def refinement_mapper(src: str, tgt: str, tok):
    x, y = tok.encode(src), tok.encode(tgt)
    sc = np.zeros((len(x) + 1, len(y) + 1), dtype=np.int32)  # gap = 0
    tb = np.zeros_like(sc)
    tb[0, 1:], tb[1:, 0], tb[0, 0] = 1, 2, 4
    for i in range(1, len(x) + 1):
        for j in range(1, len(y) + 1):
            left, up, diag = sc[i, j - 1], sc[i - 1, j], sc[i - 1, j - 1] + (1 if x[i - 1] == y[j - 1] else -1)
            sc[i, j] = max(left, up, diag)
            tb[i, j] = 1 if sc[i, j] == left else 2 if sc[i, j] == up else 3
    i, j, pairs = len(x), len(y), []
    while i > 0 or j > 0:
        if tb[i, j] == 3:
            i, j = i - 1, j - 1
            pairs.append((j, i))
        elif tb[i, j] == 1:
            j -= 1
            pairs.append((j, -1))
        elif tb[i, j] == 2:
            i -= 1
        else:
            break
    base = torch.tensor(pairs[::-1], dtype=torch.int64)
    alphas = torch.ones(L)
    alphas[: len(base)] = base[:, 1].ne(-1).float()
    mapper = torch.zeros(L, dtype=torch.int64)
    mapper[: len(base)] = base[:, 1]
    mapper[len(base):] = len(y) + torch.arange(L - len(y))
    return mapper[None], alphas[None]


def cross_alpha(prompts, steps: int, cross_replace_steps, tok) -> torch.Tensor:
    crs = dict(cross_replace_steps) if isinstance(cross_replace_steps, dict) else {"default_": cross_replace_steps}
    crs.setdefault("default_", (0., 1.))
    a = torch.zeros(steps + 1, 1, L)

    def window(bounds, inds=None):
        if type(bounds) is float:
            bounds = (0, bounds)
        s, e = int(bounds[0] * a.shape[0]), int(bounds[1] * a.shape[0])
        inds_ = torch.arange(L) if inds is None else inds
        a[:s, 0, inds_] = 0
        a[s:e, 0, inds_] = 1
        a[e:, 0, inds_] = 0
    window(crs["default_"])
    for k, v in crs.items():
        if k != "default_":
            ind = word_inds(prompts[1], k, tok)
            if len(ind) > 0:
                window(v, ind)
    return a.reshape(steps + 1, 1, 1, 1, L)


# ------------------------------------------------------------------------------------------------
# attention controllers operating on materialised probabilities
# ------------------------------------------------------------------------------------------------
class Store:
    """AttentionStore (+ counters of AttentionControl)."""

    def __init__(self, max_size=32):
        self.cur_step = self.cur_att_layer = 0
        self.num_att_layers = 32
        self.max_size = max_size
        self.step_store = self._empty()
        self.attention_store: Dict[str, List[torch.Tensor]] = {}

    @staticmethod
    def _empty():
        return {f"{p}_{k}": [] for p in ("down", "mid", "up") for k in ("cross", "self")}

    def forward(self, attn, is_cross, place):
        key = f"{place}_{'cross' if is_cross else 'self'}"
        if attn.shape[1] <= 32 ** 2:
            self.step_store[key].append(attn)
        return attn

    def __call__(self, attn, is_cross, place):
        h = attn.shape[0]
        attn[h // 2:] = self.forward(attn[h // 2:], is_cross, place)
        self.cur_att_layer += 1
        if self.cur_att_layer == self.num_att_layers:
            self.cur_att_layer = 0
            self.cur_step += 1
            if not self.attention_store:
                self.attention_store = self.step_store
            else:
                for k in self.attention_store:
                    for i in range(len(self.attention_store[k])):
                        self.attention_store[k][i] += self.step_store[k][i]
            self.step_store = self._empty()
        return attn

    def step_callback(self, x_t):
        return x_t

    def aggregate(self, n_prompts, res, from_where, select):
        out = []
        for loc in (["mid"] if res == 8 else from_where):
            for item in self.attention_store[f"{loc}_cross"]:
                if item.shape[1] == res * res:
                    out.append((item / self.cur_step).reshape(n_prompts, -1, res, res, item.shape[-1])[select])
        out = torch.cat(out, 0)
        return out.sum(0) / out.shape[0]

    def word_map(self, mask_idx, res=16, from_where=("up", "down"), resize=64, prompt_idx=0, n_prompts=1):
        m = self.aggregate(n_prompts, res, list(from_where), prompt_idx)[:, :, mask_idx][None]
        m = m / m.max()
        if resize is not None and m.shape[-2:] != (resize, resize):
            m = F.interpolate(m[None], (resize, resize), mode="bicubic")[0].clamp(0, 1)
        return m


class LocalBlend:
    def __init__(self, prompts, words, tok, num_steps, start_blend=0.2, th=0.3):
        self.alpha = torch.zeros(len(prompts), 1, 1, 1, 1, L)
        for i, (p, ws) in enumerate(zip(prompts, words)):
            for w in ([ws] if isinstance(ws, str) else ws):
                self.alpha[i, :, :, :, :, word_inds(p, w, tok)] = 1
        self.start, self.counter, self.th = int(start_blend * num_steps), 0, th

    def __call__(self, x_t, store):
        self.counter += 1
        if self.counter > self.start:
            maps = store["down_cross"][2:4] + store["up_cross"][:3]
            maps = torch.cat([m.reshape(self.alpha.shape[0], -1, 1, 16, 16, L) for m in maps], dim=1)
            maps = (maps * self.alpha).sum(-1).mean(1)
            maps = F.max_pool2d(maps, (3, 3), (1, 1), padding=(1, 1))
            mask = F.interpolate(maps, size=x_t.shape[2:])
            mask = mask / mask.max(2, keepdims=True)[0].max(3, keepdims=True)[0]
            mask = mask.gt(self.th)
            mask = (mask[:1] + mask).to(x_t.dtype)
            x_t = x_t[:1] + mask * (x_t - x_t[:1])
        return x_t


class Edit(Store):
    """AttentionControlEdit with Replace / Refine and optional Reweight on top."""

    def __init__(self, prompts, tok, num_steps, is_replace_controller, cross_replace_steps, self_replace_steps,
                 blend_words=None, equilizer_params=None):
        super().__init__()
        self.batch_size = 2
        self.alpha = cross_alpha(prompts, num_steps, cross_replace_steps, tok)
        srs = (0, self_replace_steps) if type(self_replace_steps) is float else self_replace_steps
        self.self_window = int(num_steps * srs[0]), int(num_steps * srs[1])
        self.replace = is_replace_controller
        if self.replace:
            self.mapper = replacement_mapper(prompts[0], prompts[1], tok)
        else:
            self.mapper, al = refinement_mapper(prompts[0], prompts[1], tok)
            self.ref_alphas = al.reshape(1, 1, 1, L)
        self.eq = None
        if equilizer_params is not None:
            self.eq = torch.ones(1, L)
            for w, v in zip(equilizer_params["words"], equilizer_params["values"]):
                self.eq[:, word_inds(prompts[1], w, tok)] = v
        self.lb = None if blend_words is None else LocalBlend(prompts, blend_words, tok, num_steps)

    def _cross(self, base, rep):
        if self.replace:
            out = torch.einsum('hpw,bwn->bhpn', base, self.mapper)
        else:
            out = base[:, :, self.mapper].permute(2, 0, 1, 3) * self.ref_alphas + rep * (1 - self.ref_alphas)
        if self.eq is not None:  # Reweight wraps the previous controller: its output becomes the new base
            out = out[0][None] * self.eq[:, None, None, :]
        return out

    def forward(self, attn, is_cross, place):
        super().forward(attn, is_cross, place)
        if is_cross or (self.self_window[0] <= self.cur_step < self.self_window[1]):
            h = attn.shape[0] // self.batch_size
            attn = attn.reshape(self.batch_size, h, *attn.shape[1:])
            base, rep = attn[0], attn[1:]
            if is_cross:
                a = self.alpha[self.cur_step]
                attn[1:] = self._cross(base, rep) * a + (1 - a) * rep
            elif rep.shape[2] <= 32 ** 2:
                attn[1:] = base.unsqueeze(0).expand(rep.shape[0], *base.shape)
            attn = attn.reshape(self.batch_size * h, *attn.shape[2:])
        return attn

    def step_callback(self, x_t):
        return x_t if self.lb is None else self.lb(x_t, self.attention_store)


class Masa:
    """MutualSelfAttentionControl(start_step=4, start_layer=10, total_steps=50)."""

    def __init__(self, step=4, layer=10):
        self.cur_step = self.cur_att_layer = 0
        self.steps, self.layers = range(step, 50), range(layer, 16)

    def __call__(self, q, k, v, attn, is_cross, heads, scale):
        if is_cross or self.cur_step not in self.steps or self.cur_att_layer // 2 not in self.layers:
            out = torch.einsum('b i j, b j d -> b i d', attn, v)
        else:
            outs = []
            for qh, kh, vh in zip(q.chunk(2), k.chunk(2), v.chunk(2)):
                ks, vs = kh[:heads], vh[:heads]
                b = qh.shape[0] // heads
                qq = qh.reshape(b, heads, *qh.shape[1:]).permute(1, 0, 2, 3).reshape(heads, -1, qh.shape[-1])
                p = (torch.einsum("h i d, h j d -> h i j", qq, ks) * scale).softmax(-1)
                o = torch.einsum("h i j, h j d -> h i d", p, vs)
                outs.append(o.reshape(heads, b, -1, o.shape[-1]).permute(1, 0, 2, 3).reshape(b * heads, -1, o.shape[-1]))
            out = torch.cat(outs, 0)
        self.cur_att_layer += 1
        if self.cur_att_layer == 32:
            self.cur_att_layer = 0
            self.cur_step += 1
        return out


def _attention_modules(unet):
    for name, child in unet.named_children():
        place = "down" if "down" in name else "up" if "up" in name else "mid" if "mid" in name else None
        if place is None:
            continue
        for m in child.modules():
            if m.__class__.__name__ == "Attention":
                yield m, place


class hooked:
    """Context manager: route every Attention of the UNet through explicit attention + controller."""

    def __init__(self, unet, ptp: Optional[Store] = None, masa: Optional[Masa] = None):
        self.unet, self.ptp, self.masa = unet, ptp, masa

    def __enter__(self):
        self.saved = []
        n = 0
        for m, place in _attention_modules(self.unet):
            self.saved.append((m, m.forward))
            m.forward = self._make(m, place)
            n += 1
        assert n == 32
        return self

    def __exit__(self, *a):
        for m, f in self.saved:
            m.forward = f

    def _make(self, mod, place):
        def fwd(x, encoder_hidden_states=None, attention_mask=None):
            is_cross = encoder_hidden_states is not None
            ctx = encoder_hidden_states if is_cross else x
            q, k, v = (mod.head_to_batch_dim(t) for t in (mod.to_q(x), mod.to_k(ctx), mod.to_v(ctx)))
            attn = (torch.einsum("b i d, b j d -> b i j", q, k) * mod.scale).softmax(dim=-1)
            if self.ptp is not None:
                attn = self.ptp(attn, is_cross, place)
            if self.masa is not None:
                out = self.masa(q, k, v, attn, is_cross, mod.heads, mod.scale)
            else:
                out = torch.einsum("b i j, b j d -> b i d", attn, v)
            return mod.to_out[0](mod.batch_to_head_dim(out))
        return fwd


# ------------------------------------------------------------------------------------------------
# scheduler algebra
# ------------------------------------------------------------------------------------------------
class Ddim:
    def __init__(self, steps: int):
        betas = torch.linspace(0.00085 ** 0.5, 0.012 ** 0.5, 1000, dtype=torch.float32) ** 2
        self.ac = torch.cumprod(1 - betas, 0)
        self.final = self.ac[0]
        self.steps = steps
        self.delta = 1000 // steps
        self.timesteps = torch.from_numpy((np.arange(steps) * self.delta)[::-1].copy().astype(np.int64) + 1)

    def a(self, t):
        t = min(int(t), 999)
        return self.ac[t] if t >= 0 else self.final

    def inverse(self, eps, t, x):  # sameshift: from t-delta to t
        af, at = self.a(int(t) - self.delta), self.a(t)
        x0 = (x - (1 - af) ** 0.5 * eps) / af ** 0.5
        return at ** 0.5 * x0 + (1 - at) ** 0.5 * eps

    def variance(self, t):
        at, ap = self.a(t), self.a(int(t) - self.delta)
        return ((1 - ap) / (1 - at)) * (1 - at / ap)

    def step(self, eps, t, x, eta=0.0, noise=None, add_noise=False):
        at, ap = self.a(t), self.a(int(t) - self.delta)
        x0 = (x - (1 - at) ** 0.5 * eps) / at ** 0.5
        sd = eta * self.variance(t) ** 0.5
        prev = ap ** 0.5 * x0 + (1 - ap - sd ** 2) ** 0.5 * eps
        if add_noise:
            prev = prev + sd * noise
        return prev


# ------------------------------------------------------------------------------------------------
# end-to-end edits
# ------------------------------------------------------------------------------------------------
def context(pipe, prompt: str, negative: str = "") -> torch.Tensor:
    def emb(t):
        ids = pipe.tokenizer([t], padding="max_length", max_length=77, truncation=True, return_tensors="pt").input_ids
        return pipe.text_encoder(ids)[0]
    return torch.cat([emb(negative), emb(prompt)])


def cat_context(ctxs):
    x = torch.stack(ctxs, 1)
    return x.reshape(2 * len(ctxs), *x.shape[2:])


def _cfg(unet, latent, t, ctx, g):
    u, c = unet(latent, t, encoder_hidden_states=ctx)["sample"].chunk(2)
    return u + g * (c - u)


def eta_table(eta):
    if not isinstance(eta, (tuple, list)):
        eta = (eta, eta)
    if len(eta) == 3 or isinstance(eta[0], (tuple, list)):
        (x1, y1), (x2, y2) = eta[0], eta[1]
        p = eta[2] if len(eta) == 3 else 1
        a = (y2 - y1) / (x2 - x1) ** p
        e = a * (np.clip(np.linspace(0, 1, 1000), x1, x2) - x1) ** p + y1
    else:
        e = np.linspace(eta[0], eta[1], 1000)
    return np.clip(e, 0, None)


@torch.no_grad()
def edit(pipe, image, src: str, tgt: str, inverter: str = "etainv", editor: str = "ptp", steps: int = 50,
         ptp_cfg: Optional[dict] = None, edit_word_idx=(1, 1), eta=(0.0, 0.4), g_bwd: float = 7.5, g_fwd: float = 1.0,
         seed: int = 0, n_noise: int = 10, thres: float = 0.2, decode: bool = True, on_unet=None) -> dict:
    """One edit exactly as `Editor.edit` of the reference executes it (see module docstring for the line map).
    inverter in {diffinv, dirinv, etainv}; editor in {simple, ptp, masactrl}."""
    unet, tok, sch = pipe.unet, pipe.tokenizer, Ddim(steps)
    if on_unet is not None:  # timing hook for the CPU baseline
        orig = unet.forward
        unet.forward = lambda *a, **k: on_unet(orig, *a, **k)
    try:
        src_ctx, tgt_ctx = context(pipe, src), context(pipe, tgt)
        z = pipe.vae.encode(image)["latent_dist"].mean * 0.18215
        # ---- loop A: inversion ----
        lat = [z]
        maps_per_step = []
        use_store = inverter == "etainv"
        gf = 1.0 if editor == "simple" else g_fwd
        store = Store(max_size=16) if use_store else None
        for t in reversed(sch.timesteps):
            if use_store:
                with hooked(unet, ptp=store):
                    eps = _cfg(unet, torch.cat([z] * 2), t, src_ctx, gf)
            elif gf == 1.0:
                eps = unet(z, t, encoder_hidden_states=src_ctx[1:])["sample"]
            else:
                eps = _cfg(unet, torch.cat([z] * 2), t, src_ctx, gf)
            z = sch.inverse(eps, t, z)
            if use_store:
                words = src.split(" ")
                maps_per_step.append([store.word_map(words.index(w) + 1) for w in words])
            lat.append(z)
        fwd_mean = None
        if use_store:
            fwd_mean = [torch.stack([m[w] for m in maps_per_step]).mean(0) for w in range(len(maps_per_step[0]))]
        # ---- loop B: denoise ----
        ctx = cat_context([src_ctx, tgt_ctx])
        x = torch.cat([lat[-1]] * 2)
        ctrl = Edit([src, tgt], tok, steps, **{k: v for k, v in ptp_cfg.items() if k != "prompts"}) if editor == "ptp" else None
        masa = Masa() if editor == "masactrl" else None
        etas = eta_table(eta)
        gen = torch.Generator().manual_seed(seed)
        bwd, picks = [], []
        for i, t in enumerate(sch.timesteps):
            if ctrl is not None or masa is not None:
                with hooked(unet, ptp=ctrl, masa=masa):
                    eps = _cfg(unet, torch.cat([x] * 2), t, ctx, g_bwd)
            else:
                eps = _cfg(unet, torch.cat([x] * 2), t, ctx, g_bwd)
            src_prev = lat[-(i + 2)]
            if inverter == "etainv":
                e = float(etas[int(t)])
                cand = torch.randn((n_noise, 1, 4, 64, 64), generator=gen)
                rec0 = sch.step(eps[:1], t, x[:1], eta=e, noise=torch.zeros_like(eps[:1]), add_noise=e > 0)
                z_opt = (src_prev - rec0) / (e * sch.variance(t) ** 0.5)
                losses = torch.square(cand - z_opt).reshape(n_noise, -1).mean(1)
                k = int(torch.argmin(losses))
                picks.append(k)
                mask = (fwd_mean[edit_word_idx[0]] > thres).to(x.dtype)
                x = sch.step(eps, t, x, eta=mask * torch.full_like(cand[k], e), noise=cand[k], add_noise=True)
                x[:1] = x[:1] + (src_prev[:1] - x[:1])
            else:
                x = sch.step(eps, t, x)
                if inverter == "dirinv":
                    x = torch.cat((x[:1] + (src_prev - x[:1]), x[1:]))
            x = x.clone()
            if ctrl is not None:
                x = ctrl.step_callback(x)
            bwd.append(x)
        out = dict(inv_latents=torch.stack(lat), bwd_latents=torch.stack(bwd), latent_inv=x[:1], latent=x[1:],
                   picks=picks, fwd_mean=fwd_mean)
        if decode:
            img = pipe.vae.decode(1 / 0.18215 * x)["sample"]
            out["image_inv"], out["image"] = img[:1], img[1:]
        return out
    finally:
        if on_unet is not None:
            unet.forward = orig
