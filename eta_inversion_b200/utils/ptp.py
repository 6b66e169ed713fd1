"""Prompt-to-prompt attention controllers in *data* form.

The reference (modules/utils/ptp.py, a port of google/prompt-to-prompt) monkey-patches all 32 ``Attention.forward``
methods, materialises every probability tensor and calls ``controller(attn, is_cross, place)`` on it
(modules/utils/ptp_utils.py:196-302).  Here a controller instead *describes* the edit for the next UNet forward as an
``AttnControl`` (engine.py / include/etai.h); the attention kernels apply it between softmax and P.V and accumulate
the only statistics any consumer reads.  Class names, constructor arguments, counters (``cur_step``) and the public
methods are kept so ``make_controller(**cfg)`` configs written for the reference work unchanged.

Semantics reproduced:
  AttentionStore          ptp.py:133-171   cumulative sum over steps of cond-row cross maps (post-edit)
  AttentionControlEdit    ptp.py:174-231   cross: P' = a_step*F(P_src,P_tgt) + (1-a_step)*P_tgt ; self: P_tgt := P_src
                                           inside the step window for layers with <= 32^2 tokens
  AttentionReplace/Refine/Reweight ptp.py:234-274
  LocalBlend              ptp.py:18-72
  aggregate_attention     ptp.py:288-303
"""
from __future__ import annotations

from typing import Dict, List, Optional, Sequence, Tuple, Union

import torch
import torch.nn.functional as nnf

from ..engine import AttnControl, h2d
from . import ptp_utils, seq_aligner

MAX_NUM_WORDS = 77
_PLACES = ("down", "mid", "up")
# cross-attention layers per place at each query resolution of the SD-1.x UNet at 64x64 latents (App. A)
_LAYERS_AT = {64: {"down": 2, "mid": 0, "up": 3}, 32: {"down": 2, "mid": 0, "up": 3},
              16: {"down": 2, "mid": 0, "up": 3}, 8: {"down": 0, "mid": 1, "up": 0}}
_HEADS = 8


class LocalBlend:
    def __init__(self, model, prompts: List[str], words, substruct_words=None, start_blend: float = 0.2,
                 th: Tuple[float, float] = (.3, .3)) -> None:
        if substruct_words is not None:
            raise NotImplementedError("substruct_words is unusable in the reference too (ptp.py:41 passes 3 of 4 args)")
        self.model = model
        alpha_layers = torch.zeros(len(prompts), MAX_NUM_WORDS)
        for i, (prompt, words_) in enumerate(zip(prompts, words)):
            if type(words_) is str:
                words_ = [words_]
            for word in words_:
                ind = ptp_utils.get_word_inds(prompt, word, self.model.tokenizer)
                alpha_layers[i, ind] = 1
        self.alpha_layers = h2d(alpha_layers, self.model.device)  # [prompts, 77]
        self.start_blend = int(start_blend * self.model.scheduler.num_inference_steps)
        self.counter = 0
        self.th = th

    def get_mask(self, x_t: torch.Tensor, maps: torch.Tensor, use_pool: bool = True) -> torch.Tensor:
        # maps: [prompts, 1, 16, 16] = mean over the 40 (layer, head) maps of sum_w P*alpha
        if use_pool:
            maps = nnf.max_pool2d(maps, (3, 3), (1, 1), padding=(1, 1))
        mask = nnf.interpolate(maps, size=(x_t.shape[2:]))
        mask = mask / mask.max(2, keepdims=True)[0].max(3, keepdims=True)[0]
        mask = mask.gt(self.th[1 - int(use_pool)])
        return mask[:1] + mask

    def __call__(self, x_t: torch.Tensor, store: "AttentionStore") -> torch.Tensor:
        self.counter += 1
        if self.counter > self.start_blend:
            acc = store.accumulated(16, ("down", "up"))  # [prompts, 256, 77], sum over 5 layers x 8 heads x steps
            maps = (acc * self.alpha_layers[:, None, :]).sum(-1) / (5 * _HEADS)
            maps = maps.reshape(-1, 1, 16, 16)
            mask = self.get_mask(x_t, maps, True).to(x_t.dtype)
            x_t = x_t[:1] + mask * (x_t - x_t[:1])
        return x_t


class EmptyControl:
    def step_callback(self, x_t):
        return x_t

    def between_steps(self):
        return

    def begin_forward(self, unet, batch_rows: int) -> Optional[AttnControl]:
        return None

    def end_forward(self) -> None:
        return


class AttentionControl:
    """Counter semantics of ptp.py:91-131: ``cur_step`` counts UNet forwards since creation/reset."""

    def __init__(self) -> None:
        self.cur_step = 0
        self.num_att_layers = 32
        self.cur_att_layer = 0

    def step_callback(self, x_t):
        return x_t

    def between_steps(self):
        return

    def reset(self):
        self.cur_step = 0
        self.cur_att_layer = 0

    def begin_forward(self, unet, batch_rows: int) -> Optional[AttnControl]:
        raise NotImplementedError

    def end_forward(self) -> None:
        self.cur_step += 1
        self.between_steps()


class AttentionStore(AttentionControl):
    """Keeps, per place, the running sum over (steps, layers, heads) of the cond-row cross-attention maps at
    ``store_res`` x ``store_res`` query tokens.  That is every map the reference's consumers read: LocalBlend
    (ptp.py:37) and get_attention_map / aggregate_attention with res=16 (ptp_editor.py:43-85, eta_inversion.py:87-96).
    The self-attention maps the reference also hoards are never consumed on this path and are not
    kept."""

    def __init__(self, max_size: int = 32, store_res: int = 16) -> None:
        super().__init__()
        self.max_size = max_size
        self.store_res = store_res
        self._acc: Dict[str, torch.Tensor] = {}
        self._rows = 0

    def reset(self):
        super().reset()
        self._acc = {}
        self._rows = 0

    def _ensure(self, device, rows: int) -> None:
        if self._rows != rows or not self._acc:
            self._rows = rows
            self._acc = {p: torch.zeros((rows, self.store_res ** 2, MAX_NUM_WORDS), dtype=torch.float32, device=device)
                         for p in _PLACES}

    def begin_forward(self, unet, batch_rows: int) -> AttnControl:
        rows = batch_rows // 2  # only the conditional half is seen by the controller (ptp.py:112-113)
        self._ensure(unet.device, rows)
        return AttnControl(store_rows=list(range(batch_rows - rows, batch_rows)), store_res=self.store_res,
                           store_down=self._acc["down"], store_mid=self._acc["mid"], store_up=self._acc["up"])

    def accumulated(self, res: int, from_where: Sequence[str]) -> torch.Tensor:
        if res != self.store_res:
            raise NotImplementedError(f"attention maps are kept at res {self.store_res} only (asked for {res})")
        if not self._acc:
            raise RuntimeError("no attention stored yet")
        return sum(self._acc[p] for p in from_where if p in self._acc)

    def num_maps(self, res: int, from_where: Sequence[str]) -> int:
        return _HEADS * sum(_LAYERS_AT[res][p] for p in from_where)


class AttentionControlEdit(AttentionStore):
    def __init__(self, model, prompts: List[str], num_steps: int,
                 cross_replace_steps: Union[float, Tuple[float, float], Dict[str, Tuple[float, float]]],
                 self_replace_steps: Union[float, Tuple[float, float]], local_blend: Optional[LocalBlend],
                 attn_replace_thres=None) -> None:
        super().__init__()
        if len(prompts) != 2:
            raise NotImplementedError("one (source, target) prompt pair per controller")
        self.model = model
        self.prompts = prompts
        self.attn_replace_thres = attn_replace_thres or 32 ** 2
        self.batch_size = len(prompts)
        alpha_host = ptp_utils.get_time_words_attention_alpha(prompts, num_steps, cross_replace_steps, model.tokenizer)
        self._alpha_active = [bool(a.any()) for a in alpha_host]  # host copy: skip the edit when a step's alpha is all 0
        self.cross_replace_alpha = h2d(alpha_host, model.device)
        # per-step [1,77] fp32 rows, sliced once (begin_forward runs under the lanes' shared GIL every step)
        self._alpha_rows = list(self.cross_replace_alpha.reshape(len(alpha_host), 1, MAX_NUM_WORDS).float().contiguous().unbind(0))
        if type(self_replace_steps) is float:
            self_replace_steps = 0, self_replace_steps
        self.num_self_replace = int(num_steps * self_replace_steps[0]), int(num_steps * self_replace_steps[1])
        self.local_blend = local_blend
        ident = torch.eye(MAX_NUM_WORDS, device=model.device)[None]
        one = torch.ones((1, MAX_NUM_WORDS), device=model.device)
        self._mapper, self._blend_a, self._equalizer = ident, one, one.clone()

    def step_callback(self, x_t: torch.Tensor) -> torch.Tensor:
        assert tuple(x_t.shape) == (2, 4, 64, 64)
        if self.local_blend is not None:
            x_t = self.local_blend(x_t, self)
        return x_t

    def edit_tables(self) -> Tuple[torch.Tensor, torch.Tensor, torch.Tensor]:
        """(mapper [1,77,77], blend_a [1,77], equalizer [1,77]) of F = eq*(a*(P_src M) + (1-a)*P_tgt)."""
        return self._mapper, self._blend_a, self._equalizer

    def begin_forward(self, unet, batch_rows: int) -> AttnControl:
        if batch_rows != 2 * self.batch_size:
            raise RuntimeError(f"prompt-to-prompt edit expects {2 * self.batch_size} UNet rows, got {batch_rows}")
        ctrl = super().begin_forward(unet, batch_rows)
        src, tgt = batch_rows // 2, batch_rows // 2 + 1  # rows [u_src, u_tgt, c_src, c_tgt]
        mapper, blend_a, eq = self.edit_tables()
        if self._alpha_active[self.cur_step]:  # alpha == 0 for every word leaves P_tgt untouched (ptp.py:209-210)
            ctrl.edit_pairs = [(src, tgt)]
            ctrl.mapper, ctrl.blend_a, ctrl.equalizer = mapper.contiguous(), blend_a.contiguous(), eq.contiguous()
            ctrl.alpha_step = self._alpha_rows[self.cur_step]
        if self.num_self_replace[0] <= self.cur_step < self.num_self_replace[1]:
            rows = list(range(batch_rows))
            qk = list(rows)
            qk[tgt] = src  # target cond row reuses softmax(Q_src K_src^T), keeps its own V
            ctrl.self_rows = (qk, list(qk), rows)
            ctrl.self_max_tokens = int(self.attn_replace_thres)
            ctrl.self_layer_mask = 0xFFFF
        return ctrl


class AttentionReplace(AttentionControlEdit):
    def __init__(self, model, prompts, num_steps: int, cross_replace_steps, self_replace_steps,
                 local_blend: Optional[LocalBlend] = None, attn_replace_thres=None) -> None:
        super().__init__(model, prompts, num_steps, cross_replace_steps, self_replace_steps, local_blend,
                         attn_replace_thres=attn_replace_thres)
        self.mapper = h2d(seq_aligner.get_replacement_mapper(prompts, model.tokenizer), model.device)
        self._mapper = self.mapper.float()


class AttentionRefine(AttentionControlEdit):
    def __init__(self, model, prompts, num_steps: int, cross_replace_steps, self_replace_steps,
                 local_blend: Optional[LocalBlend] = None) -> None:
        super().__init__(model, prompts, num_steps, cross_replace_steps, self_replace_steps, local_blend)
        self.mapper, alphas = seq_aligner.get_refinement_mapper(prompts, model.tokenizer)
        self.mapper, alphas = h2d(self.mapper, model.device), h2d(alphas, model.device)
        self.alphas = alphas.reshape(alphas.shape[0], 1, 1, alphas.shape[1])
        # gather base[..., mapper[n]] as a one-hot 77x77 matrix (index -1 wraps like torch indexing; its alpha is 0)
        onehot = torch.zeros((1, MAX_NUM_WORDS, MAX_NUM_WORDS), device=model.device)
        idx = self.mapper[0] % MAX_NUM_WORDS
        onehot[0, idx, torch.arange(MAX_NUM_WORDS, device=model.device)] = 1.0
        self._mapper, self._blend_a = onehot, alphas.float()


class AttentionReweight(AttentionControlEdit):
    def __init__(self, model, prompts, num_steps: int, cross_replace_steps, self_replace_steps, equalizer: torch.Tensor,
                 local_blend: Optional[LocalBlend] = None, controller: Optional[AttentionControlEdit] = None) -> None:
        super().__init__(model, prompts, num_steps, cross_replace_steps, self_replace_steps, local_blend)
        self.equalizer = h2d(equalizer, model.device)
        self.prev_controller = controller
        if controller is not None:
            self._mapper, self._blend_a, _ = controller.edit_tables()
        self._equalizer = self.equalizer.float().reshape(1, MAX_NUM_WORDS)


def get_equalizer(model, text: str, word_select, values) -> torch.Tensor:
    if type(word_select) is int or type(word_select) is str:
        word_select = (word_select,)
    equalizer = torch.ones(1, 77)
    for word, val in zip(word_select, values):
        inds = ptp_utils.get_word_inds(text, word, model.tokenizer)
        equalizer[:, inds] = val
    return equalizer


def aggregate_attention(prompts, attention_store: AttentionStore, res: int, from_where: List[str], is_cross: bool,
                        select: int) -> torch.Tensor:
    """Average cross-attention map [res,res,77] of prompt `select` (ptp.py:288-303)."""
    if not is_cross:
        raise NotImplementedError("self-attention maps are not kept (never consumed on the hot path)")
    if res == 8:
        from_where = ["mid"]
    acc = attention_store.accumulated(res, from_where)[select]
    avg = acc / max(attention_store.cur_step, 1) / attention_store.num_maps(res, from_where)
    return avg.reshape(res, res, MAX_NUM_WORDS)


def make_controller(model, prompts: List[str], is_replace_controller: bool, cross_replace_steps, self_replace_steps,
                    blend_words=None, equilizer_params=None, **kwargs) -> AttentionControlEdit:
    num_steps = model.scheduler.num_inference_steps
    lb = None if blend_words is None else LocalBlend(model, prompts, blend_words)
    if is_replace_controller:
        controller = AttentionReplace(model, prompts, num_steps, cross_replace_steps=cross_replace_steps,
                                      self_replace_steps=self_replace_steps, local_blend=lb, **kwargs)
    else:
        controller = AttentionRefine(model, prompts, num_steps, cross_replace_steps=cross_replace_steps,
                                     self_replace_steps=self_replace_steps, local_blend=lb, **kwargs)
    if equilizer_params is not None:
        eq = get_equalizer(model, prompts[1], equilizer_params["words"], equilizer_params["values"])
        controller = AttentionReweight(model, prompts, num_steps, cross_replace_steps=cross_replace_steps,
                                       self_replace_steps=self_replace_steps, equalizer=eq, local_blend=lb,
                                       controller=controller, **kwargs)
    return controller
