"""Host-side token alignment tables (integer / 0-1 valued, must be bit-exact with the reference).

Mirrors the *behaviour* of `models/p2p/seq_aligner.py` (get_refinement_mapper :121-128, get_mapper :107-118,
global_align :61-76, get_replacement_mapper :189-195) with a fresh implementation: a gap-0 / match+1 / mismatch-1
global alignment whose tie-breaking is (left, up, diagonal) in that order, traced back from the bottom-right corner.
Parity is checked in tests/test_host_tables.py against tables produced by the reference itself.
"""
from __future__ import annotations

from typing import List, Sequence, Tuple

import numpy as np
import torch

GAP, MATCH, MISMATCH = 0, 1, -1
_LEFT, _UP, _DIAG = 1, 2, 3


def _align(src: Sequence[int], tgt: Sequence[int]) -> List[Tuple[int, int]]:
    """Returns, for every target position j, (j, i) with i the aligned source position or -1 (inserted token)."""
    n, m = len(src), len(tgt)
    score = np.zeros((n + 1, m + 1), dtype=np.int64)
    move = np.zeros((n + 1, m + 1), dtype=np.int8)
    score[0, 1:] = GAP * np.arange(1, m + 1)
    score[1:, 0] = GAP * np.arange(1, n + 1)
    move[0, 1:] = _LEFT
    move[1:, 0] = _UP
    for i in range(1, n + 1):
        for j in range(1, m + 1):
            left = score[i, j - 1] + GAP
            up = score[i - 1, j] + GAP
            diag = score[i - 1, j - 1] + (MATCH if src[i - 1] == tgt[j - 1] else MISMATCH)
            best = max(left, up, diag)
            score[i, j] = best
            move[i, j] = _LEFT if best == left else (_UP if best == up else _DIAG)
    pairs: List[Tuple[int, int]] = []
    i, j = n, m
    while i > 0 or j > 0:
        mv = move[i, j]
        if mv == _DIAG:
            i, j = i - 1, j - 1
            pairs.append((j, i))
        elif mv == _LEFT:
            j -= 1
            pairs.append((j, -1))
        else:  # _UP: a source token with no counterpart
            i -= 1
    pairs.reverse()
    return pairs


def get_mapper(x: str, y: str, tokenizer, max_len: int = 77):
    xs, ys = tokenizer.encode(x), tokenizer.encode(y)
    pairs = _align(xs, ys)
    k = len(pairs)
    src_pos = torch.tensor([p[1] for p in pairs], dtype=torch.int64)
    alphas = torch.ones(max_len)
    alphas[:k] = (src_pos != -1).float()
    mapper = torch.zeros(max_len, dtype=torch.int64)
    mapper[:k] = src_pos
    mapper[k:] = len(ys) + torch.arange(max_len - len(ys))
    return mapper, alphas


def get_refinement_mapper(prompts: Sequence[str], tokenizer, max_len: int = 77):
    ms, als = zip(*(get_mapper(prompts[0], p, tokenizer, max_len) for p in prompts[1:]))
    return torch.stack(ms), torch.stack(als)


def get_word_inds(text: str, word_place, tokenizer) -> np.ndarray:
    """Token positions (1-based after BOS) of a word given by index or by string; utils/utils.py:84-102."""
    words = text.split(" ")
    if isinstance(word_place, str):
        wanted = [i for i, w in enumerate(words) if w == word_place]
    elif isinstance(word_place, int):
        wanted = [word_place]
    else:
        wanted = list(word_place)
    hits: List[int] = []
    if wanted:
        pieces = [tokenizer.decode([tid]).strip("#") for tid in tokenizer.encode(text)][1:-1]
        w_idx, consumed = 0, 0
        for tok_idx, piece in enumerate(pieces):
            consumed += len(piece)
            if w_idx in wanted:
                hits.append(tok_idx + 1)
            if consumed >= len(words[w_idx]):
                w_idx += 1
                consumed = 0
    return np.array(hits)


def get_replacement_mapper_(x: str, y: str, tokenizer, max_len: int = 77) -> torch.Tensor:
    wx, wy = x.split(" "), y.split(" ")
    if len(wx) != len(wy):
        raise ValueError(
            "attention replacement edit can only be applied on prompts with the same length"
            f" but prompt A has {len(wx)} words and prompt B has {len(wy)} words.")
    changed = [i for i in range(len(wy)) if wy[i] != wx[i]]
    src_spans = [get_word_inds(x, i, tokenizer) for i in changed]
    tgt_spans = [get_word_inds(y, i, tokenizer) for i in changed]
    m = np.zeros((max_len, max_len))
    i = j = span = 0
    while i < max_len and j < max_len:
        if span < len(src_spans) and src_spans[span][0] == i:
            s, t = src_spans[span], tgt_spans[span]
            if len(s) == len(t):
                m[s, t] = 1
            else:
                for tt in t:
                    m[s, tt] = 1 / len(t)
            span += 1
            i += len(s)
            j += len(t)
        elif span < len(src_spans):
            m[i, j] = 1
            i += 1
            j += 1
        else:
            m[j, j] = 1
            i += 1
            j += 1
    return torch.from_numpy(m).float()


def get_replacement_mapper(prompts: Sequence[str], tokenizer, max_len: int = 77) -> torch.Tensor:
    return torch.stack([get_replacement_mapper_(prompts[0], p, tokenizer, max_len) for p in prompts[1:]])
