"""Seam B of the reference: the duck-typed `model` handle its loops call, backed by libpnpinv.so.

    model.unet(latents, t, encoder_hidden_states=ctx)["sample"]   models/p2p/inversion.py:273, p2p_guidance_forward.py:109
    model.unet.in_channels                                          utils/utils.py:51
    model.scheduler.{timesteps, alphas_cumprod, final_alpha_cumprod, config, set_timesteps, step}
    model.vae / model.tokenizer / model.text_encoder / model.device

The VAE (vae.py -> csrc/vae.cu) and the CLIP text encoder (clip.py -> csrc/clip.cu) are fused handles of the same
library; the tokenizer is a pluggable attribute (`transformers.CLIPTokenizer` for a checkpoint directory).
"""
from __future__ import annotations

import ctypes as C
from typing import Dict, Optional

import torch

from . import _lib, arch
from .scheduler import DDIMSchedulerDev


class UNetOutput(dict):
    def __getattr__(self, k):
        try:
            return self[k]
        except KeyError as e:
            raise AttributeError(k) from e


class FusedUNet:
    """`UNet2DConditionModel.forward` replacement (my_diffusers/models/unet_2d_condition.py:189-273)."""

    in_channels = 4

    def __init__(self, state_dict: Optional[Dict[str, torch.Tensor]], device="cuda:0", max_batch: int = 4,
                 share_weights_with: "Optional[FusedUNet]" = None):
        self._lib = _lib.load()
        dev = torch.device(device)
        if dev.type != "cuda":
            raise _lib.PnpError("FusedUNet needs a CUDA device (sm_100a); there is no CPU fallback")
        if dev.index is None:  # 'cuda' means the CURRENT device (torchrun ranks set it), not ordinal 0
            dev = torch.device("cuda", torch.cuda.current_device())
        self.device = dev
        self.max_batch = max_batch
        self._ctx_ref = None
        self._ctx_version = None
        self._ctx_batch = None
        self._controller = None
        self.num_calls = 0
        self._parent = share_weights_with
        h = C.c_void_p()
        if share_weights_with is not None:
            # a second handle on the same GPU sharing the read-only weights (pnp_clone): own arenas / graphs / stream
            if share_weights_with.device != dev:
                raise _lib.PnpError("share_weights_with: the parent engine lives on another device")
            with torch.cuda.device(dev):
                _lib.check(self._lib.pnp_clone(share_weights_with.handle, max_batch, C.byref(h)))
            self._h = h
            return
        with torch.cuda.device(dev):
            _lib.check(self._lib.pnp_create(dev.index, max_batch, C.byref(h)))
            self._h = h
            for name, shape in arch.unet_param_specs():
                t = state_dict[name]
                if tuple(t.shape) != tuple(shape):
                    raise _lib.PnpError(f"parameter {name}: expected shape {shape}, got {tuple(t.shape)}")
                t16 = t.detach().to("cpu", torch.float16).contiguous()
                _lib.check(self._lib.pnp_load_param(h, name.encode(), C.c_void_p(t16.data_ptr()), t16.numel()))
            _lib.check(self._lib.pnp_finalize_params(h))
            # time embeddings for every possible timestep value: t_index == t
            ts = (C.c_int64 * 1000)(*range(1000))
            _lib.check(self._lib.pnp_set_timesteps(h, ts, 1000, _lib.current_stream_ptr()))

    # -- controller registration (models/p2p/attention_control.py:12-81 patches 32 CrossAttention.forward; here the
    #    controller is lowered to a pnp_attn_ctrl descriptor per call)
    def set_controller(self, controller):
        self._controller = controller

    @property
    def handle(self):
        return self._h

    def set_context(self, ctx: torch.Tensor):
        ctx32 = ctx.detach().to(self.device, torch.float32).contiguous()
        _lib.check(self._lib.pnp_set_context(self._h, C.c_void_p(ctx32.data_ptr()), ctx32.shape[0],
                                             _lib.current_stream_ptr()))
        self._ctx_ref, self._ctx_version, self._ctx_batch = ctx, ctx._version, ctx32.shape[0]

    def named_children(self):  # the reference walks down/mid/up children to count attention layers (:72-79)
        return iter(())

    def kernel_launches(self) -> int:
        n = C.c_int64()
        _lib.check(self._lib.pnp_kernel_launches(self._h, C.byref(n)))
        return n.value

    def __call__(self, sample, timestep, encoder_hidden_states=None, **kwargs):
        if encoder_hidden_states is None:
            raise _lib.PnpError("encoder_hidden_states is required")
        x = sample
        if not x.is_cuda:
            raise _lib.PnpError("FusedUNet: latents must be CUDA tensors (no CPU fallback)")
        if x.device != self.device or (encoder_hidden_states.is_cuda and encoder_hidden_states.device != self.device):
            raise _lib.PnpError(f"FusedUNet on {self.device}: latents / context live on {x.device}")
        x = x.detach().to(torch.float32).contiguous()
        B = x.shape[0]
        if tuple(x.shape[1:]) != (4, 64, 64):
            raise _lib.PnpError(f"FusedUNet: expected latents (B,4,64,64), got {tuple(x.shape)}")
        ctx = encoder_hidden_states
        if ctx.shape[0] != B:
            raise _lib.PnpError("FusedUNet: context batch must equal latent batch")
        if ctx is not self._ctx_ref or ctx._version != self._ctx_version or self._ctx_batch != B:
            self.set_context(ctx)
        t = int(timestep)
        if not 0 <= t < 1000:
            raise _lib.PnpError(f"timestep {t} outside [0, 1000)")
        ctrl = None
        if self._controller is not None and hasattr(self._controller, "descriptor"):
            ctrl = self._controller.descriptor(B)
        out = torch.empty_like(x)
        _lib.check(self._lib.pnp_unet_forward(self._h, C.c_void_p(x.data_ptr()), B, t,
                                              C.byref(ctrl) if ctrl is not None else None,
                                              C.c_void_p(out.data_ptr()), _lib.current_stream_ptr()))
        if self._controller is not None and hasattr(self._controller, "after_unet_call"):
            self._controller.after_unet_call()
        self.num_calls += 1
        return UNetOutput(sample=out)

    def close(self):
        if getattr(self, "_h", None):
            self._lib.pnp_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class FusedModel:
    """The `ldm_stable` pipeline object of the reference editors (models/p2p_editor.py:23-25)."""

    def __init__(self, unet_state_dict, device="cuda:0", max_batch: int = 4, tokenizer=None, text_encoder=None, vae=None,
                 table_dtype: str = "float32", share_weights_with: "Optional[FusedModel]" = None):
        self.unet = FusedUNet(unet_state_dict, device=device, max_batch=max_batch,
                              share_weights_with=None if share_weights_with is None else share_weights_with.unet)
        self.device = self.unet.device
        self.table_dtype = table_dtype
        self.scheduler = DDIMSchedulerDev(engine=self.unet.handle, table_dtype=table_dtype)
        self.tokenizer = tokenizer
        self.text_encoder = text_encoder
        self.vae = vae

    def clone(self, max_batch: Optional[int] = None) -> "FusedModel":
        """A second model handle on the same GPU that shares this one's weight buffers (lanes of parallel.EditLanes)."""
        return FusedModel(None, device=str(self.device), max_batch=max_batch or self.unet.max_batch,
                          tokenizer=self.tokenizer, text_encoder=self.text_encoder, vae=self.vae,
                          table_dtype=self.table_dtype, share_weights_with=self)

    @classmethod
    def synthetic(cls, device="cuda:0", max_batch: int = 4, seed: int = 0, table_dtype: str = "float32",
                  with_vae: bool = False, with_clip: bool = False):
        """Random-init SD-1.x UNet (+ VAE) + fake tokenizer / text encoder (pnpinversion_b200/synth.py) -- the offline
        stand-in for StableDiffusionPipeline.from_pretrained("CompVis/stable-diffusion-v1-4").  `with_clip`: the fused CLIP
        text encoder (csrc/clip.cu) with random-init weights instead of the embedding-table stand-in."""
        from . import synth

        vae = None
        if with_vae:
            from .vae import FusedVAE

            vae = FusedVAE(synth.synth_vae_state_dict(seed), device=device)
        text_encoder = synth.SynthTextEncoder()
        if with_clip:
            from .clip import FusedCLIPTextEncoder

            text_encoder = FusedCLIPTextEncoder(synth.synth_clip_state_dict(seed), device=device)
        return cls(synth.synth_unet_state_dict(seed), device=device, max_batch=max_batch,
                   tokenizer=synth.FakeTokenizer(), text_encoder=text_encoder, vae=vae, table_dtype=table_dtype)

    @classmethod
    def from_state_dict_file(cls, path: str, **kw):
        """Loads a diffusers-format UNet state_dict: a `.safetensors` / `.bin` file, the `unet/` directory or the whole
        pipeline directory of an SD-1.x checkpoint; validated against the architecture table (checkpoint.py)."""
        from .checkpoint import load_unet_state_dict

        return cls(load_unet_state_dict(path), **kw)

    @classmethod
    def from_pretrained(cls, path: str, **kw):
        """Local-directory counterpart of `StableDiffusionPipeline.from_pretrained` (models/p2p_editor.py:23-25): UNet into
        the fused engine, CLIP tokenizer / text encoder from the same directory when present."""
        from .checkpoint import load_fused_model

        return load_fused_model(path, **kw)
