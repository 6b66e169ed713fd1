"""Host-side prompt-to-prompt schedule tables (reference: modules/utils/ptp_utils.py:305-357)."""
from __future__ import annotations

from typing import Dict, List, Optional, Tuple, Union

import torch

from .seq_aligner import get_word_inds  # noqa: F401  (same symbol is exported by the reference's ptp_utils)


def _set_window(alpha: torch.Tensor, bounds: Union[float, Tuple[float, float]], prompt_ind: int,
                word_inds: Optional[torch.Tensor] = None) -> torch.Tensor:
    if type(bounds) is float:
        bounds = 0, bounds
    start, end = int(bounds[0] * alpha.shape[0]), int(bounds[1] * alpha.shape[0])
    if word_inds is None:
        word_inds = torch.arange(alpha.shape[2])
    alpha[:start, prompt_ind, word_inds] = 0
    alpha[start:end, prompt_ind, word_inds] = 1
    alpha[end:, prompt_ind, word_inds] = 0
    return alpha


def get_time_words_attention_alpha(prompts: List[str], num_steps: int,
                                   cross_replace_steps: Union[float, Dict[str, Tuple[float, float]]], tokenizer,
                                   max_num_words: int = 77) -> torch.Tensor:
    """alpha[step, prompt-1, 0, 0, word] in {0,1}: is the cross-attention edit active for this word at this step."""
    if type(cross_replace_steps) is not dict:
        cross_replace_steps = {"default_": cross_replace_steps}
    if "default_" not in cross_replace_steps:
        cross_replace_steps["default_"] = (0., 1.)
    alpha = torch.zeros(num_steps + 1, len(prompts) - 1, max_num_words)
    for i in range(len(prompts) - 1):
        alpha = _set_window(alpha, cross_replace_steps["default_"], i)
    for key, item in cross_replace_steps.items():
        if key != "default_":
            inds = [get_word_inds(prompts[i], key, tokenizer) for i in range(1, len(prompts))]
            for i, ind in enumerate(inds):
                if len(ind) > 0:
                    alpha = _set_window(alpha, item, i, ind)
    return alpha.reshape(num_steps + 1, len(prompts) - 1, 1, 1, max_num_words)
