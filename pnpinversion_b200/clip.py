"""The text-encoder handle of seam B: `model.text_encoder(input_ids)[0]` (models/p2p/inversion.py:42,50,296,304,
models/p2p/p2p_guidance_forward.py:43,49,86,92, models/edict/edict_functions.py:818-838), backed by the fused CLIP text
encoder of libpnpinv.so (csrc/clip.cu, C ABI `pnp_clip_*`).

Arithmetic spec: `transformers.CLIPTextModel` with the SD-1.x `text_encoder/` configuration (768 hidden, 12 layers of 12
heads, quick_gelu MLP of 3072, 77 positions, causal mask, final LayerNorm); parameters under that model's state_dict names.
No CPU fallback: the output is a CUDA tensor on the device the handle was created on.
"""
from __future__ import annotations

import ctypes as C
import threading
from typing import Dict, List, Tuple

import torch

from . import _lib

HIDDEN, HEADS, MLP, POSITIONS = 768, 12, 3072, 77
Spec = Tuple[str, Tuple[int, ...]]


def clip_text_param_specs(layers: int = 12, vocab: int = 49408) -> List[Spec]:
    """(name, shape) of the CLIPTextModel parameters the engine reads (`position_ids` buffers are ignored)."""
    t = "text_model."
    s: List[Spec] = [(t + "embeddings.token_embedding.weight", (vocab, HIDDEN)),
                     (t + "embeddings.position_embedding.weight", (POSITIONS, HIDDEN))]
    for i in range(layers):
        p = f"{t}encoder.layers.{i}."
        for proj in ("q_proj", "k_proj", "v_proj", "out_proj"):
            s += [(f"{p}self_attn.{proj}.weight", (HIDDEN, HIDDEN)), (f"{p}self_attn.{proj}.bias", (HIDDEN,))]
        s += [(p + "layer_norm1.weight", (HIDDEN,)), (p + "layer_norm1.bias", (HIDDEN,)),
              (p + "mlp.fc1.weight", (MLP, HIDDEN)), (p + "mlp.fc1.bias", (MLP,)),
              (p + "mlp.fc2.weight", (HIDDEN, MLP)), (p + "mlp.fc2.bias", (HIDDEN,)),
              (p + "layer_norm2.weight", (HIDDEN,)), (p + "layer_norm2.bias", (HIDDEN,))]
    s += [(t + "final_layer_norm.weight", (HIDDEN,)), (t + "final_layer_norm.bias", (HIDDEN,))]
    return s


def count_layers(state_dict: Dict[str, torch.Tensor]) -> int:
    n = 0
    while f"text_model.encoder.layers.{n}.layer_norm1.weight" in state_dict:
        n += 1
    return n


class _Out(tuple):
    """`text_encoder(ids)[0]` and `.last_hidden_state` (edict_functions.py:829 uses the attribute)."""

    @property
    def last_hidden_state(self):
        return self[0]


class FusedCLIPTextEncoder:
    def __init__(self, state_dict: Dict[str, torch.Tensor], device="cuda:0"):
        self._lib = _lib.load()
        dev = torch.device(device)
        if dev.type != "cuda":
            raise _lib.PnpError("FusedCLIPTextEncoder needs a CUDA device (sm_100a); there is no CPU fallback")
        if dev.index is None:
            dev = torch.device("cuda", torch.cuda.current_device())
        self.device = dev
        self._lock = threading.Lock()  # one handle may serve several lanes (parallel.EditLanes): its plan buffers are shared
        layers = count_layers(state_dict)
        if layers == 0:
            raise _lib.PnpError("not a CLIPTextModel state dict: text_model.encoder.layers.0.layer_norm1.weight is missing")
        tok = state_dict.get("text_model.embeddings.token_embedding.weight")
        if tok is None or tok.dim() != 2 or tok.shape[1] != HIDDEN:
            raise _lib.PnpError("CLIP token embedding must be [vocab, 768] (the SD-1.x text encoder)")
        h = C.c_void_p()
        with torch.cuda.device(dev):
            _lib.check(self._lib.pnp_clip_create(dev.index, C.byref(h)))
            self._h = h
            for name, shape in clip_text_param_specs(layers, tok.shape[0]):
                if name not in state_dict:
                    raise _lib.PnpError(f"CLIP parameter missing: {name}")
                t = state_dict[name]
                if tuple(t.shape) != tuple(shape):
                    raise _lib.PnpError(f"CLIP parameter {name}: expected shape {shape}, got {tuple(t.shape)}")
                t16 = t.detach().to("cpu", torch.float16).contiguous()
                _lib.check(self._lib.pnp_clip_load_param(h, name.encode(), C.c_void_p(t16.data_ptr()), t16.numel()))
            _lib.check(self._lib.pnp_clip_finalize(h))
        v, n = C.c_int(), C.c_int()
        _lib.check(self._lib.pnp_clip_vocab_size(self._h, C.byref(v), C.byref(n)))
        self.vocab_size, self.num_layers = v.value, n.value

    @classmethod
    def from_transformers(cls, model, device="cuda:0"):
        """From a loaded `transformers.CLIPTextModel` (e.g. `<checkpoint>/text_encoder`)."""
        return cls(model.state_dict(), device=device)

    @property
    def handle(self):
        return self._h

    def __call__(self, input_ids, attention_mask=None, **kwargs):
        if attention_mask is not None:
            raise _lib.PnpError("FusedCLIPTextEncoder: attention_mask is not supported (the reference never passes one)")
        extra = [k for k, v in kwargs.items() if v not in (None, False) and k != "return_dict"]
        if extra:  # output_hidden_states / output_attentions / position_ids ...: not computed here, so not silently dropped
            raise _lib.PnpError(f"FusedCLIPTextEncoder: unsupported arguments {extra} (only last_hidden_state is produced)")
        ids = torch.as_tensor(input_ids)
        if ids.dim() != 2 or ids.shape[1] != POSITIONS:
            raise _lib.PnpError(f"FusedCLIPTextEncoder: expected input_ids (B,{POSITIONS}), got {tuple(ids.shape)}")
        ids32 = ids.detach().to("cpu", torch.int32).contiguous()
        B = ids32.shape[0]
        out = torch.empty((B, POSITIONS, HIDDEN), device=self.device, dtype=torch.float32)
        with self._lock, torch.cuda.device(self.device):
            for b0 in range(0, B, 64):
                nb = min(64, B - b0)
                _lib.check(self._lib.pnp_clip_encode(self._h, C.c_void_p(ids32[b0:b0 + nb].data_ptr()), nb,
                                                     C.c_void_p(out[b0:b0 + nb].data_ptr()), _lib.current_stream_ptr()))
            # the next caller (possibly another lane on another stream) reuses the plan's buffers: finish before unlocking
            torch.cuda.current_stream().synchronize()
        return _Out((out,))

    def to(self, *a, **k):  # the reference moves its pipeline with .to(device); the handle already lives there
        return self

    def kernel_launches(self) -> int:
        n = C.c_int64()
        _lib.check(self._lib.pnp_clip_kernel_launches(self._h, C.byref(n)))
        return n.value

    def close(self):
        if getattr(self, "_h", None):
            self._lib.pnp_clip_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass
