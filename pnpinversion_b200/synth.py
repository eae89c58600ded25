"""Seeded synthetic stand-ins for everything the offline box lacks (weights, tokenizer, text encoder, latents).

There is no network, hence no SD-1.x checkpoint, CLIP weights or PIE-Bench images.  BASELINE.json asks for
"synthetic 512x512 latents" and a "random-init UNet"; this module defines them *once* so that the oracle
(build container, fp64 CPU), the committed golden fixtures and the CUDA engine (GPU box) all see bit-identical
inputs.  Everything is generated on the CPU with explicit `torch.Generator` seeds and rounded to fp16, i.e. the
fp64 oracle and the fp16 engine share the *same rounded parameters* ("hard parts").

The gains are chosen so the random network is numerically meaningful: branch outputs are O(1) against the
residual stream and attention logits have std ~3 (peaked softmax rows), otherwise attention injection would be a
no-op on near-uniform maps and parity tests would be blind to it.
"""
from __future__ import annotations

import math
import re
from typing import Dict, List, Sequence

import torch

from .arch import CROSS_DIM, MAX_TOKENS, unet_param_specs

SEED_IMAGE_BASE = 1234  # the reference's per-image seed, run_editing_p2p.py:30-36,118


def _gain_for(name: str) -> float:
    if re.search(r"attn[12]\.to_[qk]\.weight$", name):
        return 3.0
    if name.startswith("conv_out"):
        return 1.0
    return 2.0


def synth_unet_state_dict(seed: int = 0, dtype=torch.float16) -> Dict[str, torch.Tensor]:
    """name -> CPU tensor (fp16-rounded).  U(-b,b), b = gain/sqrt(fan_in); norms 1+0.1n / 0.1n; biases 0.05n."""
    out: Dict[str, torch.Tensor] = {}
    for idx, (name, shape) in enumerate(unet_param_specs()):
        g = torch.Generator().manual_seed(seed * 1_000_003 + idx)
        if "norm" in name:
            t = torch.randn(shape, generator=g) * 0.1
            if name.endswith("weight"):
                t = t + 1.0
        elif name.endswith("bias"):
            t = torch.randn(shape, generator=g) * 0.05
        else:
            fan_in = 1
            for d in shape[1:]:
                fan_in *= d
            b = _gain_for(name) / math.sqrt(fan_in)
            t = (torch.rand(shape, generator=g) * 2.0 - 1.0) * b
        out[name] = t.to(dtype)
    return out


def synth_vae_state_dict(seed: int = 0, dtype=torch.float16) -> Dict[str, torch.Tensor]:
    """Seeded stand-in for the SD-1.x VAE weights (same recipe as the UNet: U(-b,b) with b = 2/sqrt(fan_in), norms near
    identity, small biases).  Used by the VAE oracle / fixtures."""
    from .arch import vae_param_specs

    out: Dict[str, torch.Tensor] = {}
    for idx, (name, shape) in enumerate(vae_param_specs()):
        g = torch.Generator().manual_seed(900_001 + seed * 1_000_003 + idx)
        if "norm" in name:
            t = torch.randn(shape, generator=g) * 0.1
            if name.endswith("weight"):
                t = t + 1.0
        elif name.endswith("bias"):
            t = torch.randn(shape, generator=g) * 0.05
        else:
            fan_in = 1
            for d in shape[1:]:
                fan_in *= d
            t = (torch.rand(shape, generator=g) * 2.0 - 1.0) * (2.0 / math.sqrt(fan_in))
        out[name] = t.to(dtype)
    return out


def synth_clip_state_dict(seed: int = 0, layers: int = 12, vocab: int = 49408, dtype=torch.float16) -> Dict[str, torch.Tensor]:
    """Seeded stand-in for the SD-1.x CLIP text encoder weights (names of `transformers.CLIPTextModel.state_dict()`):
    embeddings N(0,1) (token) / 0.3 N(0,1) (position), projections U(-b,b) with b = gain/sqrt(fan_in) (gain 3 on q / k so the
    causal attention rows are peaked, 2 elsewhere), LayerNorms near identity, small biases."""
    from .clip import clip_text_param_specs

    out: Dict[str, torch.Tensor] = {}
    for idx, (name, shape) in enumerate(clip_text_param_specs(layers, vocab)):
        g = torch.Generator().manual_seed(700_001 + seed * 1_000_003 + idx)
        if "token_embedding" in name:
            t = torch.randn(shape, generator=g)
        elif "position_embedding" in name:
            t = torch.randn(shape, generator=g) * 0.3
        elif "layer_norm" in name:
            t = torch.randn(shape, generator=g) * 0.1
            if name.endswith("weight"):
                t = t + 1.0
        elif name.endswith("bias"):
            t = torch.randn(shape, generator=g) * 0.05
        else:
            gain = 3.0 if ("q_proj" in name or "k_proj" in name) else 2.0
            t = (torch.rand(shape, generator=g) * 2.0 - 1.0) * (gain / math.sqrt(shape[1]))
        out[name] = t.to(dtype)
    return out


class FakeTokenizer:
    """Whitespace tokenizer with the CLIP tokenizer's call surface (what the reference touches:
    `__call__(..., padding="max_length", max_length=77, truncation=True, return_tensors="pt").input_ids`,
    `.encode`, `.decode`, `.model_max_length`; models/p2p/inversion.py:290-303, utils/utils.py:84-102)."""

    BOS, EOS = 49406, 49407
    model_max_length = MAX_TOKENS

    def __init__(self):
        self._vocab: Dict[str, int] = {}
        self._inv: Dict[int, str] = {self.BOS: "<|startoftext|>", self.EOS: "<|endoftext|>"}

    def _tok(self, w: str) -> int:
        if w not in self._vocab:
            # stable id independent of insertion order
            h = 0
            for ch in w:
                h = (h * 131 + ord(ch)) % 40000
            i = 1000 + h
            while i in self._inv and self._inv[i] != w:
                i += 1
            self._vocab[w] = i
            self._inv[i] = w
        return self._vocab[w]

    def encode(self, text: str) -> List[int]:
        return [self.BOS] + [self._tok(w) for w in text.split()] + [self.EOS]

    def decode(self, ids) -> str:
        if isinstance(ids, int):
            ids = [ids]
        return " ".join(self._inv.get(int(i), "?") for i in ids)

    def __call__(self, text, padding="max_length", max_length=MAX_TOKENS, truncation=True, return_tensors="pt"):
        if isinstance(text, str):
            text = [text]
        rows = []
        for t in text:
            ids = self.encode(t)[:max_length]
            ids = ids + [self.EOS] * (max_length - len(ids))
            rows.append(ids)
        ids = torch.tensor(rows, dtype=torch.long)

        class _Out:
            pass

        o = _Out()
        o.input_ids = ids
        return o


class PieceTokenizer(FakeTokenizer):
    """Splits every word longer than 4 characters into 3-character pieces (a stand-in for CLIP's BPE splitting rare
    words), so that swapped words can have different token counts: the fractional branch of the replacement mapper
    (seq_aligner.py:168-174)."""

    def encode(self, text):
        ids = [self.BOS]
        for w in text.split():
            pieces = [w] if len(w) <= 4 else [w[i:i + 3] for i in range(0, len(w), 3)]
            ids += [self._tok(p) for p in pieces]
        return ids + [self.EOS]

    def decode(self, ids):
        if isinstance(ids, int):
            ids = [ids]
        return "".join(self._inv.get(int(i), "?") for i in ids)


class SynthTextEncoder:
    """Seeded embedding-table "text encoder": token id + position -> N(0,1) row of 768 (fp16-rounded).
    Satisfies `model.text_encoder(ids)[0]` (models/p2p/inversion.py:296,305)."""

    def __init__(self, seed: int = 7, dtype=torch.float32):
        self.seed = seed
        self.dtype = dtype
        self._cache: Dict[tuple, torch.Tensor] = {}

    def _row(self, tok: int, pos: int) -> torch.Tensor:
        key = (tok, pos)
        r = self._cache.get(key)
        if r is None:
            g = torch.Generator().manual_seed((self.seed * 77_003 + tok) * 131 + pos)
            r = torch.randn(CROSS_DIM, generator=g)
            self._cache[key] = r
        return r

    def __call__(self, input_ids: torch.Tensor):
        ids = input_ids.cpu()
        out = torch.empty(ids.shape[0], ids.shape[1], CROSS_DIM)
        for b in range(ids.shape[0]):
            for p in range(ids.shape[1]):
                out[b, p] = self._row(int(ids[b, p]), p)
        out = out.to(torch.float16).to(self.dtype).to(input_ids.device)
        return (out,)

    def to(self, *a, **k):
        return self


def synth_latent(index: int = 0, batch: int = 1) -> torch.Tensor:
    """z0 ~ N(0,1) (batch,4,64,64) fp32, seeded 1234+index."""
    g = torch.Generator().manual_seed(SEED_IMAGE_BASE + index)
    return torch.randn(batch, 4, 64, 64, generator=g)


# the notebook's example pair, run_editing_p2p_one_image.ipynb cell 5
CAT_PROMPTS = ("a cat sitting on a table with a green eyes", "a watercolor of a cat sitting on a table with a green eyes")
