"""DDIM scheduler with the reference's surface (`DDIMSchedulerDev`, models/p2p/scheduler_dev.py:10-121, built in
models/p2p_editor.py:18-22 with beta 0.00085->0.012 scaled_linear, clip_sample=False, set_alpha_to_one=False).

Integer timestep arithmetic and the alphas_cumprod table live on the host (bit-exact with the reference); the tensor
math of `step` runs in the fused CUDA epilogue (csrc/epilogue.cu) -- there is no PyTorch fallback for CUDA tensors.
"""
from __future__ import annotations

import ctypes as C
from types import SimpleNamespace

import numpy as np
import torch

from . import _lib

NUM_TRAIN_TIMESTEPS = 1000


class StepOutput(dict):
    """Supports both out["prev_sample"] (p2p_guidance_forward.py:112) and out.prev_sample."""

    def __getattr__(self, k):
        try:
            return self[k]
        except KeyError as e:
            raise AttributeError(k) from e


def make_alphas_cumprod(table_dtype: str = "float32") -> torch.Tensor:
    """float32: what diffusers>=0.10 (the P2P / MasaCtrl pin) builds with torch.linspace(dtype=float32);
    float64: what the vendored diffusers 0.3.0 builds with numpy (my_diffusers/schedulers/scheduling_ddim.py:105-113)."""
    if table_dtype == "float32":
        betas = torch.linspace(0.00085 ** 0.5, 0.012 ** 0.5, NUM_TRAIN_TIMESTEPS, dtype=torch.float32) ** 2
        return torch.cumprod(1.0 - betas, dim=0)
    if table_dtype == "float64":
        betas = np.linspace(0.00085 ** 0.5, 0.012 ** 0.5, NUM_TRAIN_TIMESTEPS, dtype=np.float64) ** 2
        return torch.from_numpy(np.cumprod(1.0 - betas, axis=0))
    raise ValueError(table_dtype)


def step_coefficients(alphas_cumprod: torch.Tensor, final_alpha, t_from: int, t_to: int):
    """(sqrt a_from, sqrt(1-a_from), sqrt a_to, sqrt(1-a_to)) as Python floats, each evaluated in the table's dtype the
    way the reference evaluates `alpha ** 0.5` / `(1 - alpha) ** 0.5` on 0-d tensors.  t < 0 selects final_alpha."""
    a_from = alphas_cumprod[t_from] if t_from >= 0 else final_alpha
    a_to = alphas_cumprod[t_to] if t_to >= 0 else final_alpha
    return (float(a_from ** 0.5), float((1 - a_from) ** 0.5), float(a_to ** 0.5), float((1 - a_to) ** 0.5))


def fused_step(engine, x, eps_c, coeffs, eps_u=None, guidance=1.0, target=None, loss_out=None, noise_loss=None,
               add_mask=0, out=None, loss_scale=1.0):
    """Launches the fused CFG + DDIM step (+offset / +rectification) kernel on CUDA fp32 tensors [n,4,64,64]."""
    if not x.is_cuda:
        raise _lib.PnpError("fused_step: tensors must live on the GPU (no CPU fallback)")
    n = x.shape[0]
    if out is None:
        out = torch.empty_like(x)
    # every pointer handed to the kernel is validated here: device, dtype, layout and row count
    for name, t, rows in (("x", x, n), ("eps_c", eps_c, n), ("eps_u", eps_u, n), ("target", target, None),
                          ("loss_out", loss_out, n), ("noise_loss", noise_loss, None), ("out", out, n)):
        if t is None:
            continue
        if not t.is_cuda or t.device != x.device:
            raise _lib.PnpError(f"fused_step: {name} must be a CUDA tensor on {x.device} (got {t.device})")
        if t.dtype != torch.float32 or not t.is_contiguous():
            raise _lib.PnpError(f"fused_step: {name} must be contiguous float32")
        if tuple(t.shape[1:]) != tuple(x.shape[1:]) or (rows is not None and t.shape[0] != rows):
            raise _lib.PnpError(f"fused_step: {name} has shape {tuple(t.shape)}, expected rows x {tuple(x.shape[1:])}")
    if target is not None and loss_out is None:
        raise _lib.PnpError("fused_step: offset mode needs loss_out")
    if noise_loss is not None and int(add_mask) >> noise_loss.shape[0]:
        raise _lib.PnpError("fused_step: add_mask selects a row beyond noise_loss")
    if int(add_mask) >> n:
        raise _lib.PnpError("fused_step: add_mask selects a row beyond the latents")
    a = _lib.StepArgs()
    a.x_dev = x.data_ptr()
    a.eps_u_dev = eps_u.data_ptr() if eps_u is not None else None
    a.eps_c_dev = eps_c.data_ptr()
    a.x_out_dev = out.data_ptr()
    a.n = n
    a.guidance = float(guidance)
    a.sqrt_a_from, a.sqrt_1m_a_from, a.sqrt_a_to, a.sqrt_1m_a_to = coeffs
    a.target_dev = target.data_ptr() if target is not None else None
    a.target_rows = target.shape[0] if target is not None else 0
    a.loss_out_dev = loss_out.data_ptr() if loss_out is not None else None
    a.loss_scale = float(loss_scale)
    a.noise_loss_dev = noise_loss.data_ptr() if noise_loss is not None else None
    a.add_mask = int(add_mask)
    _lib.check(_lib.load().pnp_step_epilogue(engine, C.byref(a), _lib.current_stream_ptr()))
    return out


class DDIMSchedulerDev:
    def __init__(self, engine=None, table_dtype: str = "float32"):
        self._engine = engine
        self.config = SimpleNamespace(num_train_timesteps=NUM_TRAIN_TIMESTEPS, beta_start=0.00085, beta_end=0.012,
                                      beta_schedule="scaled_linear", clip_sample=False, set_alpha_to_one=False)
        self.alphas_cumprod = make_alphas_cumprod(table_dtype)
        self.final_alpha_cumprod = self.alphas_cumprod[0]
        self.num_inference_steps = None
        self.timesteps = torch.arange(NUM_TRAIN_TIMESTEPS - 1, -1, -1, dtype=torch.int64)

    def set_timesteps(self, num_inference_steps: int, device=None):
        """diffusers 0.10 (the P2P pin; DDIMSchedulerDev inherits it): `(arange(0, n) * (1000 // n)).round()[::-1]`,
        steps_offset 0 -- exactly n timesteps also when n does not divide 1000 (n = 30 -> 957 ... 0)."""
        if not 1 <= int(num_inference_steps) <= NUM_TRAIN_TIMESTEPS:
            raise ValueError(f"num_inference_steps must be in [1, {NUM_TRAIN_TIMESTEPS}], got {num_inference_steps}")
        self.num_inference_steps = int(num_inference_steps)
        ratio = NUM_TRAIN_TIMESTEPS // self.num_inference_steps
        ts = (np.arange(0, self.num_inference_steps) * ratio).round()[::-1].copy().astype(np.int64)
        self.timesteps = torch.from_numpy(ts)

    def step(self, model_output, timestep, sample, eta: float = 0.0, **kwargs):
        """eta=0, epsilon prediction, no clipping: scheduler_dev.py:40-51,84,91-94."""
        if self.num_inference_steps is None:
            raise ValueError("Number of inference steps is 'None', you need to run 'set_timesteps' after creating the scheduler")
        if eta != 0.0:
            raise NotImplementedError("only the deterministic (eta=0) DDIM step is on the PnP-inversion hot path")
        t = int(timestep)
        prev_t = t - self.config.num_train_timesteps // self.num_inference_steps
        co = step_coefficients(self.alphas_cumprod, self.final_alpha_cumprod, t, prev_t)
        prev = fused_step(self._engine, sample.contiguous(), model_output.contiguous(), co)
        return StepOutput(prev_sample=prev)
