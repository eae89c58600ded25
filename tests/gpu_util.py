"""Helpers for the `-m gpu` tests: call the C ABI with torch tensors as containers."""
import ctypes as C

import torch

from pnpinversion_b200 import _lib


def ptr(t):
    return None if t is None else C.c_void_p(t.data_ptr())


def stream():
    return C.c_void_p(_lib.current_stream_ptr())


def rel_l2(a, b):
    a = a.double().flatten()
    b = b.double().flatten()
    return float((a - b).norm() / b.norm().clamp_min(1e-30))


def gemm(a, w, bias=None, residual=None, geglu=False, bn=0, split=1):
    lib = _lib.load()
    M, K = a.shape
    N = w.shape[0]
    n_out = N // 2 if geglu else N
    out = torch.empty(M, n_out, dtype=torch.float16, device=a.device)
    _lib.check(lib.pnp_test_gemm(ptr(a), M, K, a.stride(0), ptr(w), N, ptr(bias), ptr(residual), ptr(out), n_out,
                                 1 if geglu else 0, bn, split, stream()))
    torch.cuda.synchronize()
    return out


def pack_conv3(w_oihw, shortcut=None):
    """(Cout,Cin,3,3) -> (Cout, 9*Cin [+Csc]) tap-major, the layout pnp_finalize_params produces."""
    co, ci = w_oihw.shape[:2]
    p = w_oihw.permute(0, 2, 3, 1).reshape(co, 9 * ci)
    if shortcut is not None:
        p = torch.cat([p, shortcut.reshape(co, -1)], dim=1)
    return p.contiguous()


def conv3x3(x_nhwc, w_packed, bias=None, residual=None, sc0=None, sc1=None, bn=0, split=1):
    lib = _lib.load()
    B, H, W, Cc = x_nhwc.shape
    N = w_packed.shape[0]
    out = torch.empty(B, H, W, N, dtype=torch.float16, device=x_nhwc.device)
    _lib.check(lib.pnp_test_conv3x3(ptr(x_nhwc), B, H, W, Cc, ptr(w_packed), N, ptr(sc0),
                                    0 if sc0 is None else sc0.shape[-1], ptr(sc1),
                                    0 if sc1 is None else sc1.shape[-1], ptr(bias), ptr(residual), ptr(out), bn,
                                    split, stream()))
    torch.cuda.synchronize()
    return out
