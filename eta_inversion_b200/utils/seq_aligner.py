"""Host-side token alignment for prompt-to-prompt (tiny, CPU; runs once per edit).

Behavioural spec: modules/utils/seq_aligner.py of the reference (itself derived from google/prompt-to-prompt):
  get_word_inds            :113-131   word -> token indices (+1 for BOS)
  get_replacement_mapper   :134-201   word-swap prompts of equal word count -> [P-1,77,77] row-stochastic mapper
  get_refinement_mapper    :100-110   Needleman-Wunsch (gap 0, match 1, mismatch -1) -> gather index + alpha
Written from the algorithm description, including the tie-break order of the traceback (left, up, diagonal) that
decides which of several optimal alignments is used.
"""
from __future__ import annotations

from typing import List, Sequence, Tuple, Union

import numpy as np
import torch

MAX_LEN = 77


def get_word_inds(text: str, word_place: Union[int, str, Sequence[int]], tokenizer) -> np.ndarray:
    """Token positions (1-based, BOS at 0) covered by the word(s) `word_place` (index or literal) of `text`."""
    words = text.split(" ")
    if isinstance(word_place, str):
        wanted = [i for i, w in enumerate(words) if w == word_place]
    elif isinstance(word_place, int):
        wanted = [word_place]
    else:
        wanted = list(word_place)
    hits: List[int] = []
    if wanted:
        pieces = [tokenizer.decode([tok]).strip("#") for tok in tokenizer.encode(text)][1:-1]
        word_i, consumed = 0, 0
        for tok_i, piece in enumerate(pieces):
            consumed += len(piece)
            if word_i in wanted:
                hits.append(tok_i + 1)
            if consumed >= len(words[word_i]):
                word_i, consumed = word_i + 1, 0
    return np.array(hits)


def _replacement_mapper_pair(src: str, tgt: str, tokenizer, max_len: int = MAX_LEN) -> torch.Tensor:
    ws, wt = src.split(" "), tgt.split(" ")
    if len(ws) != len(wt):
        raise ValueError("attention replacement edit can only be applied on prompts with the same length"
                         f" but prompt A has {len(ws)} words and prompt B has {len(wt)} words.")
    changed = [i for i in range(len(wt)) if wt[i] != ws[i]]
    src_tok = [get_word_inds(src, i, tokenizer) for i in changed]
    tgt_tok = [get_word_inds(tgt, i, tokenizer) for i in changed]
    m = np.zeros((max_len, max_len))
    i = j = 0
    nxt = 0
    while i < max_len and j < max_len:
        if nxt < len(src_tok) and src_tok[nxt][0] == i:
            s_, t_ = src_tok[nxt], tgt_tok[nxt]
            if len(s_) == len(t_):
                m[s_, t_] = 1
            else:
                for t in t_:
                    m[s_, t] = 1 / len(t_)
            nxt += 1
            i += len(s_)
            j += len(t_)
        elif nxt < len(src_tok):
            m[i, j] = 1
            i, j = i + 1, j + 1
        else:
            m[j, j] = 1
            i, j = i + 1, j + 1
    return torch.from_numpy(m).float()


def get_replacement_mapper(prompts: List[str], tokenizer, max_len: int = MAX_LEN) -> torch.Tensor:
    return torch.stack([_replacement_mapper_pair(prompts[0], p, tokenizer, max_len) for p in prompts[1:]])


def _global_alignment_trace(x: Sequence[int], y: Sequence[int]) -> np.ndarray:
    """Needleman-Wunsch traceback codes: 1 = from left (gap in x), 2 = from above (gap in y), 3 = diagonal, 4 = origin."""
    gap, match, mismatch = 0, 1, -1
    nx, ny = len(x), len(y)
    score = np.zeros((nx + 1, ny + 1), dtype=np.int32)
    score[0, 1:] = (np.arange(ny) + 1) * gap
    score[1:, 0] = (np.arange(nx) + 1) * gap
    trace = np.zeros((nx + 1, ny + 1), dtype=np.int32)
    trace[0, 1:], trace[1:, 0], trace[0, 0] = 1, 2, 4
    for i in range(1, nx + 1):
        for j in range(1, ny + 1):
            left, up = score[i, j - 1] + gap, score[i - 1, j] + gap
            diag = score[i - 1, j - 1] + (match if x[i - 1] == y[j - 1] else mismatch)
            best = max(left, up, diag)
            score[i, j] = best
            trace[i, j] = 1 if best == left else 2 if best == up else 3
    return trace


def _mapper_pair(src: str, tgt: str, tokenizer, max_len: int = MAX_LEN) -> Tuple[torch.Tensor, torch.Tensor]:
    x, y = tokenizer.encode(src), tokenizer.encode(tgt)
    trace = _global_alignment_trace(x, y)
    pairs = []  # (target position, source position or -1), built back to front
    i, j = len(x), len(y)
    while i > 0 or j > 0:
        code = trace[i, j]
        if code == 3:
            i, j = i - 1, j - 1
            pairs.append((j, i))
        elif code == 1:
            j -= 1
            pairs.append((j, -1))
        elif code == 2:
            i -= 1
        else:
            break
    pairs.reverse()
    base = torch.tensor(pairs, dtype=torch.int64)
    alphas = torch.ones(max_len)
    alphas[: base.shape[0]] = base[:, 1].ne(-1).float()
    mapper = torch.zeros(max_len, dtype=torch.int64)
    mapper[: base.shape[0]] = base[:, 1]
    mapper[base.shape[0]:] = len(y) + torch.arange(max_len - len(y))
    return mapper, alphas


def get_refinement_mapper(prompts: List[str], tokenizer, max_len: int = MAX_LEN) -> Tuple[torch.Tensor, torch.Tensor]:
    out = [_mapper_pair(prompts[0], p, tokenizer, max_len) for p in prompts[1:]]
    return torch.stack([m for m, _ in out]), torch.stack([a for _, a in out])
