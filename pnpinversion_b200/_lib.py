"""ctypes binding of libpnpinv.so (C ABI declared in include/pnpinv.h).

There is deliberately no fallback: if the shared library is missing or a CUDA device is absent, every compute call
raises.  torch tensors are only containers here (`tensor.data_ptr()`, `torch.cuda.current_stream().cuda_stream`).
"""
from __future__ import annotations

import ctypes as C
import os
from typing import Optional

PNP_MAX_BATCH = 32
PNP_MAX_SLOTS = 8
PNP_TOKENS = 77
LATENT_ELEMS = 4 * 64 * 64

_LIB_PATH = os.path.join(os.path.dirname(os.path.abspath(__file__)), "libpnpinv.so")


class PnpError(RuntimeError):
    pass


class AttnCtrl(C.Structure):
    """Mirror of `pnp_attn_ctrl` (include/pnpinv.h)."""

    _fields_ = [
        ("self_layer_lo", C.c_int32),
        ("self_layer_hi", C.c_int32),
        ("self_max_tokens", C.c_int32),
        ("self_q_row", C.c_int32 * PNP_MAX_BATCH),
        ("self_k_row", C.c_int32 * PNP_MAX_BATCH),
        ("self_v_row", C.c_int32 * PNP_MAX_BATCH),
        ("cross_base_row", C.c_int32 * PNP_MAX_BATCH),
        ("cross_slot", C.c_int32 * PNP_MAX_BATCH),
        ("mapper", (C.c_int32 * PNP_TOKENS) * PNP_MAX_SLOTS),
        ("alphas", (C.c_float * PNP_TOKENS) * PNP_MAX_SLOTS),
        ("equalizer", (C.c_float * PNP_TOKENS) * PNP_MAX_SLOTS),
        ("cross_alpha", (C.c_float * PNP_TOKENS) * PNP_MAX_SLOTS),
        ("map_count", (C.c_int32 * PNP_TOKENS) * PNP_MAX_SLOTS),
        ("map_weight", (C.c_float * PNP_TOKENS) * PNP_MAX_SLOTS),
        ("conv_src_row", C.c_int32 * PNP_MAX_BATCH),
        ("store_slot", C.c_int32 * PNP_MAX_BATCH),
    ]


class StepArgs(C.Structure):
    """Mirror of `pnp_step_args`."""

    _fields_ = [
        ("x_dev", C.c_void_p),
        ("eps_u_dev", C.c_void_p),
        ("eps_c_dev", C.c_void_p),
        ("x_out_dev", C.c_void_p),
        ("n", C.c_int32),
        ("guidance", C.c_float),
        ("sqrt_a_from", C.c_float),
        ("sqrt_1m_a_from", C.c_float),
        ("sqrt_a_to", C.c_float),
        ("sqrt_1m_a_to", C.c_float),
        ("target_dev", C.c_void_p),
        ("target_rows", C.c_int32),
        ("loss_out_dev", C.c_void_p),
        ("loss_scale", C.c_float),
        ("noise_loss_dev", C.c_void_p),
        ("add_mask", C.c_uint32),
    ]

    def __init__(self, *a, **kw):
        super().__init__(*a, **kw)
        if "loss_scale" not in kw:
            self.loss_scale = 1.0


class BlendDesc(C.Structure):
    """Mirror of `pnp_blend_desc`."""

    _fields_ = [
        ("src_row", C.c_int32), ("tgt_row", C.c_int32), ("src_slot", C.c_int32), ("tgt_slot", C.c_int32),
        ("nwords", C.c_int32 * 2),
        ("words", (C.c_int32 * 8) * 2),
        ("alpha", (C.c_float * 8) * 2),
        ("nsub", C.c_int32 * 2),
        ("sub_words", (C.c_int32 * 8) * 2),
        ("sub_alpha", (C.c_float * 8) * 2),
        ("th_pool", C.c_float), ("th_sub", C.c_float),
    ]


PNP_LOOP_INVERT, PNP_LOOP_OFFSET, PNP_LOOP_FORWARD = 0, 1, 2


class LoopArgs(C.Structure):
    """Mirror of `pnp_loop_args`."""

    _fields_ = [
        ("mode", C.c_int32), ("n_steps", C.c_int32), ("rows", C.c_int32), ("images", C.c_int32),
        ("t_host", C.POINTER(C.c_int32)),
        ("coef_host", C.POINTER(C.c_float)),
        ("guidance", C.c_float),
        ("ctx_dev", C.c_void_p),
        ("x_dev", C.c_void_p),
        ("traj_dev", C.c_void_p),
        ("loss_dev", C.c_void_p),
        ("loss_scale_host", C.POINTER(C.c_float)),
        ("add_mask", C.c_uint32),
        ("ctrl_host", C.POINTER(AttnCtrl)),
        ("blend_host", C.POINTER(BlendDesc)),
        ("n_blend", C.c_int32), ("blend_start", C.c_int32),
    ]


_lib: Optional[C.CDLL] = None

# name -> (restype, argtypes); also the list tests/test_capi_cpu.py checks against include/pnpinv.h
_vp, _i, _i64, _f = C.c_void_p, C.c_int, C.c_int64, C.c_float
SIGNATURES = {
    "pnp_create": (_i, [_i, _i, C.POINTER(_vp)]),
    "pnp_clone": (_i, [_vp, _i, C.POINTER(_vp)]),
    "pnp_destroy": (None, [_vp]),
    "pnp_last_error": (C.c_char_p, []),
    "pnp_version": (C.c_char_p, []),
    "pnp_unet_param_count": (_i, []),
    "pnp_unet_param_spec": (_i, [_i, C.c_char_p, C.POINTER(_i), C.POINTER(_i)]),
    "pnp_load_param": (_i, [_vp, C.c_char_p, _vp, _i64]),
    "pnp_finalize_params": (_i, [_vp]),
    "pnp_set_timesteps": (_i, [_vp, C.POINTER(_i64), _i, _vp]),
    "pnp_set_context": (_i, [_vp, _vp, _i, _vp]),
    "pnp_attn_ctrl_init": (None, [C.POINTER(AttnCtrl)]),
    "pnp_unet_forward": (_i, [_vp, _vp, _i, _i, C.POINTER(AttnCtrl), _vp, _vp]),
    "pnp_step_epilogue": (_i, [_vp, C.POINTER(StepArgs), _vp]),
    "pnp_local_blend": (_i, [_vp, _vp, C.POINTER(C.c_int32), C.POINTER(C.c_int32), C.POINTER(_f), _f, _vp, _vp]),
    "pnp_local_blend_batch": (_i, [_vp, _vp, _i, C.POINTER(BlendDesc), _i, _vp, _vp]),
    "pnp_run_loop": (_i, [_vp, C.POINTER(LoopArgs), _vp]),
    "pnp_edict_mix": (_i, [_vp, _vp, _vp, _i, _f, _i, _vp]),
    "pnp_store_reset": (_i, [_vp, _vp]),
    "pnp_store_read": (_i, [_vp, _vp, _i64, _vp]),
    "pnp_unet_gemm_bytes": (_i, [_vp, _i, C.POINTER(C.c_double), C.POINTER(_i)]),
    "pnp_unet_profile": (_i, [_vp, _i, _i, _i, C.POINTER(_f), C.POINTER(C.c_int32), C.POINTER(C.c_double), _i, C.POINTER(_i)]),
    "pnp_debug_read": (_i, [_vp, _i, _i, _vp, _i64, C.POINTER(_i64), _vp]),
    "pnp_struct_size": (_i, [_i]),
    "pnp_kernel_launches": (_i, [_vp, C.POINTER(_i64)]),
    "pnp_set_use_graph": (_i, [_vp, _i]),
    "pnp_vae_create": (_i, [_i, C.POINTER(_vp)]),
    "pnp_vae_destroy": (None, [_vp]),
    "pnp_vae_load_param": (_i, [_vp, C.c_char_p, _vp, _i64]),
    "pnp_vae_finalize": (_i, [_vp]),
    "pnp_vae_encode": (_i, [_vp, _vp, _i, _i, _i, _vp, _vp]),
    "pnp_vae_decode": (_i, [_vp, _vp, _i, _i, _i, _vp, _vp]),
    "pnp_vae_kernel_launches": (_i, [_vp, C.POINTER(_i64)]),
    "pnp_test_mma_probe": (_i, [_i, _i, _i, _i, _i, _i, _i, _i, C.POINTER(_i64)]),
    "pnp_clip_create": (_i, [_i, C.POINTER(_vp)]),
    "pnp_clip_destroy": (None, [_vp]),
    "pnp_clip_load_param": (_i, [_vp, C.c_char_p, _vp, _i64]),
    "pnp_clip_finalize": (_i, [_vp]),
    "pnp_clip_vocab_size": (_i, [_vp, C.POINTER(_i), C.POINTER(_i)]),
    "pnp_clip_encode": (_i, [_vp, _vp, _i, _vp, _vp]),
    "pnp_clip_kernel_launches": (_i, [_vp, C.POINTER(_i64)]),
    "pnp_test_gemm": (_i, [_vp, _i, _i, _i, _vp, _i, _vp, _vp, _vp, _i, _i, _i, _i, _vp]),
    "pnp_test_gemm2": (_i, [_vp, _i, _vp, _i, _i, _vp, _i, _vp, _i, C.POINTER(_f), _vp]),
    "pnp_test_conv3x3": (_i, [_vp, _i, _i, _i, _i, _vp, _i, _vp, _i, _vp, _i, _vp, _vp, _vp, _i, _i, _vp]),
    "pnp_test_groupnorm": (_i, [_vp, _i, _vp, _i, _i, _i, _vp, _vp, _f, _i, _vp, _vp]),
    "pnp_test_groupnorm_path": (_i, [_i, _i, _i]),
    "pnp_test_layernorm": (_i, [_vp, _i, _i, _vp, _vp, _f, _vp, _vp]),
    "pnp_test_self_attention": (_i, [_vp, _i, _i, _i, _i, _vp, _vp, _vp, _vp, _vp]),
    "pnp_test_self_attention_tc": (_i, [_vp, _i, _i, _vp, _vp, _vp, _vp, _vp]),
    "pnp_test_cross_attention": (_i, [_vp, _vp, _i, _i, _i, _i, _i, C.POINTER(AttnCtrl), _vp, _vp, _vp]),
    "pnp_test_upsample2x": (_i, [_vp, _i, _i, _i, _i, _vp, _vp]),
    "pnp_test_im2col_s2": (_i, [_vp, _i, _i, _i, _i, _vp, _vp]),
}


def lib_path() -> str:
    return _LIB_PATH


def load() -> C.CDLL:
    """Loads libpnpinv.so (built in-tree by `make` / `__graft_entry__.build()`); raises if it is missing."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(_LIB_PATH):
        raise PnpError(
            f"{_LIB_PATH} not found: build it with `make` (or `python -c 'import __graft_entry__ as g; g.build()'`). "
            "pnpinversion_b200 has no CPU/PyTorch fallback."
        )
    lib = C.CDLL(_LIB_PATH)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)
        fn.restype = res
        fn.argtypes = args
    for which, cls in enumerate((AttnCtrl, StepArgs, BlendDesc, LoopArgs)):
        if lib.pnp_struct_size(which) != C.sizeof(cls):
            raise PnpError(f"ABI mismatch: {cls.__name__} is {C.sizeof(cls)} bytes here, "
                           f"{lib.pnp_struct_size(which)} in {_LIB_PATH} (rebuild with `make`)")
    _lib = lib
    return lib


def check(rc: int) -> None:
    if rc != 0:
        msg = load().pnp_last_error()
        raise PnpError(f"libpnpinv error {rc}: {msg.decode() if msg else '?'}")


def current_stream_ptr() -> int:
    import torch

    return int(torch.cuda.current_stream().cuda_stream)


def new_ctrl() -> AttnCtrl:
    c = AttnCtrl()
    load().pnp_attn_ctrl_init(C.byref(c))
    return c
