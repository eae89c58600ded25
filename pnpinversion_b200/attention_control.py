"""Seam C of the reference: Prompt-to-Prompt attention controllers (`models/p2p/attention_control.py`).

The reference invokes `controller(attn, is_cross, place_in_unet)` inside each of the 32 attention layers on a
*materialised* probability tensor.  Here the same classes (same names, constructor arguments, counters, windows) hold
the host-side tables and lower themselves, once per UNet call, into a `pnp_attn_ctrl` descriptor that selects the
kernel modes of csrc/attention.cu.  There is no Python-callback path: unknown controller subclasses are rejected.

Deliberate, disclosed difference: `AttentionStore` keeps only what the hot path consumes -- the five 16x16 cross
maps `down_cross[2:4] + up_cross[:3]` LocalBlend reads (attention_control.py:112) -- not every <=32^2 map.
"""
from __future__ import annotations

import abc
import ctypes as C
from typing import Optional, Sequence

import torch

from . import _lib, seq_aligner
from .ptp_utils import get_time_words_attention_alpha, get_word_inds

MAX_NUM_WORDS = 77
LATENT_SIZE = (64, 64)
NUM_ATT_LAYERS = 32
_SELF_REPLACE_MAX_TOKENS = 32 ** 2  # attention_control.py:259


def register_attention_control(model, controller):
    """Replaces the monkey-patching of attention_control.py:12-81: binds the controller to the fused UNet."""
    model.unet.set_controller(controller)
    if controller is not None:
        controller.num_att_layers = NUM_ATT_LAYERS
        if hasattr(controller, "_bind"):
            controller._bind(model.unet.handle)


def get_equalizer(text, word_select, values, tokenizer=None):
    if isinstance(word_select, (int, str)):
        word_select = (word_select,)
    equalizer = torch.ones(1, MAX_NUM_WORDS)
    for word, val in zip(word_select, values):
        inds = get_word_inds(text, word, tokenizer)
        equalizer[:, inds] = val
    return equalizer


def _words_of(alpha_row):
    nz = torch.nonzero(alpha_row).flatten().tolist()
    if len(nz) > 8:
        raise NotImplementedError("at most 8 blend tokens per prompt")
    return nz


class LocalBlend:
    """attention_control.py:95-147.  Works on the maps accumulated by the cross-attention kernel (store slots of the
    source / target row, 0 and 1 unless a batched loop places the pair elsewhere)."""

    def __init__(self, prompts, words, substruct_words=None, start_blend=0.2, th=(.3, .3), tokenizer=None, device="cuda",
                 num_ddim_steps=50):
        if len(prompts) != 2:
            raise NotImplementedError("LocalBlend is implemented for one (source, target) prompt pair")
        self.alpha_layers = self._layers(prompts, words, tokenizer)
        self.substruct_layers = None if substruct_words is None else self._layers(prompts, substruct_words, tokenizer)
        self.start_blend = int(start_blend * num_ddim_steps)
        self.counter = 0
        self.th = th
        self._engine = None
        self.last_mask = None

    @staticmethod
    def _layers(prompts, words, tokenizer):
        layers = torch.zeros(len(prompts), MAX_NUM_WORDS)
        for i, (prompt, words_) in enumerate(zip(prompts, words)):
            if isinstance(words_, str):
                words_ = [words_]
            for word in words_:
                layers[i, get_word_inds(prompt, word, tokenizer)] = 1
        return layers

    def blend_desc(self, src_row=0, tgt_row=1, src_slot=0, tgt_slot=1) -> "_lib.BlendDesc":
        """The `pnp_blend_desc` of this pair (attention_control.py:97-121 incl. the substruct branch :116-118)."""
        d = _lib.BlendDesc()
        d.src_row, d.tgt_row, d.src_slot, d.tgt_slot = src_row, tgt_row, src_slot, tgt_slot
        d.th_pool, d.th_sub = float(self.th[0]), float(self.th[1])
        for p in range(2):
            nz = _words_of(self.alpha_layers[p])
            d.nwords[p] = len(nz)
            for j, w in enumerate(nz):
                d.words[p][j] = w
                d.alpha[p][j] = float(self.alpha_layers[p, w])
            if self.substruct_layers is not None:
                nz = _words_of(self.substruct_layers[p])
                d.nsub[p] = len(nz)
                for j, w in enumerate(nz):
                    d.sub_words[p][j] = w
                    d.sub_alpha[p][j] = float(self.substruct_layers[p, w])
        return d

    def __call__(self, x_t, attention_store=None):
        self.counter += 1
        if self.counter > self.start_blend:
            if self._engine is None:
                raise _lib.PnpError("LocalBlend is not bound to an engine (register_attention_control first)")
            if not (x_t.is_cuda and x_t.dtype == torch.float32 and x_t.shape[0] == 2):
                raise _lib.PnpError("LocalBlend expects CUDA float32 latents of shape (2,4,64,64)")
            x_t = x_t.contiguous()
            d = self.blend_desc()
            _lib.check(_lib.load().pnp_local_blend_batch(self._engine, C.c_void_p(x_t.data_ptr()), 2, C.byref(d), 1, None,
                                                         _lib.current_stream_ptr()))
        return x_t


class EmptyControl:
    def step_callback(self, x_t):
        return x_t

    def between_steps(self):
        return

    def descriptor(self, batch):
        return None

    def after_unet_call(self):
        return


class AttentionControl(abc.ABC):
    def __init__(self):
        self.cur_step = 0
        self.num_att_layers = -1
        self.cur_att_layer = 0
        self._engine = None

    def _bind(self, engine):
        self._engine = engine
        if self.cur_step == 0:
            _lib.check(_lib.load().pnp_store_reset(engine, _lib.current_stream_ptr()))

    def step_callback(self, x_t):
        return x_t

    def between_steps(self):
        return

    def reset(self):
        self.cur_step = 0
        self.cur_att_layer = 0

    # one UNet call == 32 invocations of the reference's __call__ (attention_control.py:178-190)
    def after_unet_call(self):
        self.cur_att_layer = 0
        self.cur_step += 1
        self.between_steps()

    @abc.abstractmethod
    def descriptor(self, batch) -> Optional[_lib.AttnCtrl]:
        raise NotImplementedError

    def __call__(self, attn, is_cross, place_in_unet):
        raise _lib.PnpError("controllers are compiled into kernel modes; there is no materialised-attention callback")


class AttentionStore(AttentionControl):
    """attention_control.py:214-248 (only the maps LocalBlend consumes are accumulated, on the device)."""

    def __init__(self):
        super().__init__()
        self._want_store = False

    def descriptor(self, batch):
        if not self._want_store:
            return None
        c = _lib.new_ctrl()
        n = batch // 2
        c.store_slot[n] = 0
        if n > 1:
            c.store_slot[n + 1] = 1
        return c

    def get_average_attention(self):
        raise NotImplementedError("only the five 16x16 cross maps LocalBlend reads are kept (see DESIGN.md)")

    def reset(self):
        super().reset()
        if self._engine is not None:
            _lib.check(_lib.load().pnp_store_reset(self._engine, _lib.current_stream_ptr()))


class AttentionControlEdit(AttentionStore, abc.ABC):
    def __init__(self, prompts, num_steps, cross_replace_steps, self_replace_steps, local_blend, tokenizer=None,
                 device="cuda"):
        super().__init__()
        self.batch_size = len(prompts)
        if not 2 <= self.batch_size <= 1 + _lib.PNP_MAX_SLOTS:
            raise ValueError("need one source prompt and 1..8 target prompts")
        self.cross_replace_alpha = get_time_words_attention_alpha(prompts, num_steps, cross_replace_steps, tokenizer)
        if isinstance(self_replace_steps, float):
            self_replace_steps = 0, self_replace_steps
        self.num_self_replace = int(num_steps * self_replace_steps[0]), int(num_steps * self_replace_steps[1])
        self.local_blend = local_blend
        self._want_store = local_blend is not None

    def _bind(self, engine):
        super()._bind(engine)
        if self.local_blend is not None:
            self.local_blend._engine = engine

    def step_callback(self, x_t):
        if self.local_blend is not None:
            x_t = self.local_blend(x_t, None)
        return x_t

    # subclasses fill mapper / alphas / equalizer of target `slot` of THIS controller into table slot `dst` of c
    @abc.abstractmethod
    def _fill_tables(self, c: _lib.AttnCtrl, slot: int, dst: Optional[int] = None):
        raise NotImplementedError

    def self_replace_on(self) -> bool:
        return self.num_self_replace[0] <= self.cur_step < self.num_self_replace[1]

    def fill_pair(self, c: _lib.AttnCtrl, src_row: int, tgt_row: int, slot: int, dst: int, store_src=None,
                  store_tgt=None):
        """Writes the edit of target `slot` (UNet batch rows src_row -> tgt_row, conditional half) into descriptor c,
        using table slot `dst`: attention_control.py:269-282 for the current step."""
        gate = self.cross_replace_alpha[self.cur_step]  # (n-1,1,1,77)
        c.cross_base_row[tgt_row] = src_row
        c.cross_slot[tgt_row] = dst
        self._fill_tables(c, slot, dst)
        _set_row(c.cross_alpha[dst], gate[slot].reshape(-1).tolist())
        if self.self_replace_on():
            c.self_layer_lo, c.self_layer_hi, c.self_max_tokens = 0, 16, _SELF_REPLACE_MAX_TOKENS
            c.self_q_row[tgt_row] = src_row
            c.self_k_row[tgt_row] = src_row
        if self._want_store and store_src is not None:
            c.store_slot[src_row] = store_src
            c.store_slot[tgt_row] = store_tgt

    def descriptor(self, batch):
        n = self.batch_size
        if batch != 2 * n:
            raise _lib.PnpError(f"controller built for {n} prompts expects a UNet batch of {2 * n}, got {batch}")
        c = _lib.new_ctrl()
        src = n  # first cond row; the controller only touches attn[h//2:] (attention_control.py:184)
        for i in range(1, n):
            self.fill_pair(c, src, n + i, i - 1, i - 1, store_src=0 if i == 1 else None, store_tgt=1)
        return c


def _set_row(arr, values):
    arr[:] = list(values)[:MAX_NUM_WORDS]


class AttentionReplace(AttentionControlEdit):
    """attention_control.py:301-314: `einsum('hpw,bwn->bhpn', attn_base, mapper)`.  Every column of the replacement
    mapper (seq_aligner.py:152-185) is zero, a single 1 (token kept or swapped for a word of equal token count), or the
    uniform weight 1/len(target span) on the CONSECUTIVE tokens of the replaced source word (unequal token counts), so the
    einsum is lowered to `weight * sum of count consecutive source probabilities` per target token."""

    def __init__(self, prompts, num_steps, cross_replace_steps, self_replace_steps, local_blend=None, tokenizer=None,
                 device="cuda"):
        super().__init__(prompts, num_steps, cross_replace_steps, self_replace_steps, local_blend, tokenizer, device)
        self.mapper = seq_aligner.get_replacement_mapper(prompts, tokenizer)  # (n-1,77,77)
        self._spans = [self._columns(m) for m in self.mapper]

    @staticmethod
    def _columns(m):
        """(start, count, weight) per target token; raises if a column is not a uniform consecutive run."""
        start, count, weight = [], [], []
        for n in range(m.shape[1]):
            rows = torch.nonzero(m[:, n]).flatten().tolist()
            if not rows:
                start.append(0), count.append(1), weight.append(0.0)
                continue
            vals = m[rows, n]
            if rows != list(range(rows[0], rows[0] + len(rows))) or not bool((vals == vals[0]).all()):
                raise NotImplementedError("replacement mapper column is not a uniform run of consecutive source tokens")
            start.append(rows[0]), count.append(len(rows)), weight.append(float(vals[0]))
        return start, count, weight

    def _fill_tables(self, c, slot, dst=None):
        dst = slot if dst is None else dst
        start, count, weight = self._spans[slot]
        _set_row(c.mapper[dst], start)
        _set_row(c.map_count[dst], count)
        _set_row(c.map_weight[dst], weight)
        _set_row(c.alphas[dst], [1.0] * MAX_NUM_WORDS)
        _set_row(c.equalizer[dst], [1.0] * MAX_NUM_WORDS)


class AttentionRefine(AttentionControlEdit):
    """attention_control.py:317-335."""

    def __init__(self, prompts, num_steps, cross_replace_steps, self_replace_steps, local_blend=None, tokenizer=None,
                 device="cuda"):
        super().__init__(prompts, num_steps, cross_replace_steps, self_replace_steps, local_blend, tokenizer, device)
        self.mapper, alphas = seq_aligner.get_refinement_mapper(prompts, tokenizer)
        self.alphas = alphas.reshape(alphas.shape[0], 1, 1, alphas.shape[1])

    def _fill_tables(self, c, slot, dst=None):
        dst = slot if dst is None else dst
        _set_row(c.mapper[dst], self.mapper[slot].tolist())
        _set_row(c.alphas[dst], self.alphas[slot].reshape(-1).tolist())
        _set_row(c.equalizer[dst], [1.0] * MAX_NUM_WORDS)


class AttentionReweight(AttentionControlEdit):
    """attention_control.py:338-363: scales the (optionally refined/replaced) source probabilities per token."""

    def __init__(self, prompts, num_steps, cross_replace_steps, self_replace_steps, equalizer, local_blend=None,
                 controller=None, device="cuda", tokenizer=None):
        super().__init__(prompts, num_steps, cross_replace_steps, self_replace_steps, local_blend, tokenizer, device)
        self.equalizer = equalizer
        self.prev_controller = controller

    def _fill_tables(self, c, slot, dst=None):
        dst = slot if dst is None else dst
        if self.prev_controller is not None:
            self.prev_controller._fill_tables(c, slot, dst)
        else:
            _set_row(c.mapper[dst], list(range(MAX_NUM_WORDS)))
            _set_row(c.alphas[dst], [1.0] * MAX_NUM_WORDS)
        eq = self.equalizer[slot if self.equalizer.shape[0] > 1 else 0].reshape(-1).tolist()
        _set_row(c.equalizer[dst], eq)


def make_controller(pipeline, prompts, is_replace_controller, cross_replace_steps, self_replace_steps, blend_words=None,
                    equilizer_params=None, num_ddim_steps=50, device="cuda") -> AttentionControlEdit:
    """attention_control.py:366-405 (same argument names, including the reference's `equilizer_params` spelling)."""
    tok = pipeline.tokenizer
    lb = None if blend_words is None else LocalBlend(prompts, blend_words, tokenizer=tok, device=device,
                                                      num_ddim_steps=num_ddim_steps)
    cls = AttentionReplace if is_replace_controller else AttentionRefine
    controller = cls(prompts, num_ddim_steps, cross_replace_steps=cross_replace_steps,
                     self_replace_steps=self_replace_steps, local_blend=lb, tokenizer=tok)
    if equilizer_params is not None:
        eq = get_equalizer(prompts[1], equilizer_params["words"], equilizer_params["values"], tokenizer=tok)
        controller = AttentionReweight(prompts, num_ddim_steps, cross_replace_steps=cross_replace_steps,
                                       self_replace_steps=self_replace_steps, equalizer=eq, local_blend=lb,
                                       controller=controller, tokenizer=tok)
    return controller
