"""The VAE handle of seam B: `model.vae.encode(img)['latent_dist'].mean` and `model.vae.decode(z)['sample']`
(utils/utils.py:61,78), backed by the fused VAE of libpnpinv.so (csrc/vae.cu, C ABI `pnp_vae_*`).

Arithmetic spec: the reference's vendored `AutoencoderKL` (models/edict/my_diffusers/models/vae.py:480-557).  No CPU
fallback: tensors must live on the GPU the handle was created on.
"""
from __future__ import annotations

import ctypes as C
from types import SimpleNamespace
from typing import Dict, Optional

import torch

from . import _lib, arch


class _Out(dict):
    def __getattr__(self, k):
        try:
            return self[k]
        except KeyError as e:
            raise AttributeError(k) from e


class FusedVAE:
    def __init__(self, state_dict: Dict[str, torch.Tensor], device="cuda:0"):
        self._lib = _lib.load()
        dev = torch.device(device)
        if dev.type != "cuda":
            raise _lib.PnpError("FusedVAE needs a CUDA device (sm_100a); there is no CPU fallback")
        if dev.index is None:
            dev = torch.device("cuda", torch.cuda.current_device())
        self.device = dev
        h = C.c_void_p()
        with torch.cuda.device(dev):
            _lib.check(self._lib.pnp_vae_create(dev.index, C.byref(h)))
            self._h = h
            for name, shape in arch.vae_param_specs():
                if name not in state_dict:
                    raise _lib.PnpError(f"VAE parameter missing: {name}")
                t = state_dict[name]
                if tuple(t.shape) != tuple(shape):
                    # diffusers >= 0.15 stores the attention projections as (C, C, 1, 1) convs or renames them; accept
                    # any layout with the right element count for the four linear maps of the attention block
                    if t.numel() != int(torch.tensor(shape).prod()):
                        raise _lib.PnpError(f"VAE parameter {name}: expected shape {shape}, got {tuple(t.shape)}")
                t16 = t.detach().to("cpu", torch.float16).contiguous()
                _lib.check(self._lib.pnp_vae_load_param(h, name.encode(), C.c_void_p(t16.data_ptr()), t16.numel()))
            _lib.check(self._lib.pnp_vae_finalize(h))

    @property
    def handle(self):
        return self._h

    def _check(self, x, channels):
        if not (isinstance(x, torch.Tensor) and x.is_cuda and x.device == self.device):
            raise _lib.PnpError(f"FusedVAE on {self.device}: the input must be a CUDA tensor on that device")
        if x.dim() != 4 or x.shape[1] != channels:
            raise _lib.PnpError(f"FusedVAE: expected (B,{channels},H,W), got {tuple(x.shape)}")
        return x.detach().to(torch.float32).contiguous()

    def decode(self, z: torch.Tensor):
        """AutoencoderKL.decode (vae.py:535-544): (B,4,h,w) -> {'sample': (B,3,8h,8w)} fp32."""
        z = self._check(z, 4)
        B, _, h, w = z.shape
        out = torch.empty((B, 3, 8 * h, 8 * w), device=self.device, dtype=torch.float32)
        for b0 in range(0, B, 8):
            nb = min(8, B - b0)
            _lib.check(self._lib.pnp_vae_decode(self._h, C.c_void_p(z[b0:b0 + nb].data_ptr()), nb, h, w,
                                                C.c_void_p(out[b0:b0 + nb].data_ptr()), _lib.current_stream_ptr()))
        return _Out(sample=out)

    def encode(self, image: torch.Tensor):
        """AutoencoderKL.encode (vae.py:524-533): (B,3,H,W) in [-1,1] -> {'latent_dist': posterior with .mean, .logvar}."""
        image = self._check(image, 3)
        B, _, H, W = image.shape
        mom = torch.empty((B, 8, H // 8, W // 8), device=self.device, dtype=torch.float32)
        for b0 in range(0, B, 8):
            nb = min(8, B - b0)
            _lib.check(self._lib.pnp_vae_encode(self._h, C.c_void_p(image[b0:b0 + nb].data_ptr()), nb, H, W,
                                                C.c_void_p(mom[b0:b0 + nb].data_ptr()), _lib.current_stream_ptr()))
        dist = SimpleNamespace(mean=mom[:, :4], logvar=mom[:, 4:].clamp(-30.0, 20.0), mode=lambda: mom[:, :4])
        return _Out(latent_dist=dist)

    def kernel_launches(self) -> int:
        n = C.c_int64()
        _lib.check(self._lib.pnp_vae_kernel_launches(self._h, C.byref(n)))
        return n.value

    def close(self):
        if getattr(self, "_h", None):
            self._lib.pnp_vae_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass
