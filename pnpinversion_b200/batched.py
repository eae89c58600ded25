"""DirectInversion + P2P for L images per call, with the three step loops behind the C ABI (`pnp_run_loop`).

The reference edits one image per call: `DirectInversion.invert` (models/p2p/inversion.py:393-400) followed by two
`direct_inversion_p2p_guidance_forward` passes (models/p2p/p2p_guidance_forward.py:135-173), each a Python loop of 50
steps with `torch.cat` / slicing around every UNet call.  Here the same four loops run inside libpnpinv.so for L
independent images at once: UNet batch L for the inversion, 4 L for the offset / reconstruction / edit passes, no eager
tensor op inside a loop.  The per-image arithmetic is unchanged (`tests/test_gpu_batched.py` asserts bit-identity with the
Python loops of inversion.py / p2p_guidance_forward.py for L = 1, and of every image of a batch with its single-image
run).

Row layout (latent rows, PROMPT-MAJOR): row = p * L + image, p = 0 the source prompt, p = 1 the target prompt; the UNet
batch of the guided passes is [unconditional rows | conditional rows] like the reference's `torch.cat([latents] * 2)`.
"""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass
from typing import List, Optional, Sequence

import torch

from . import _lib
from .attention_control import AttentionControlEdit, make_controller
from .scheduler import step_coefficients


@dataclass
class BatchEditResult:
    x_stars: torch.Tensor             # (n_steps+1, L, 4,64,64)
    noise_loss: torch.Tensor          # (n_steps, 2L, 4,64,64), prompt-major rows
    reconstruct_latents: torch.Tensor  # (2L, 4,64,64)
    latents: torch.Tensor             # (2L, 4,64,64): rows [0,L) source branches, [L,2L) edited

    def image(self, i: int):
        """The per-image view the single-image editor returns: (x_stars list, noise_loss list, recon (2,..), latents (2,..))."""
        L = self.x_stars.shape[1]
        rows = [i, L + i]
        return ([self.x_stars[k, i:i + 1] for k in range(self.x_stars.shape[0])],
                [self.noise_loss[k, rows] for k in range(self.noise_loss.shape[0])],
                self.reconstruct_latents[rows], self.latents[rows])


def _encode_rows(model, prompts: Sequence[str]) -> torch.Tensor:
    tok, enc, dev = model.tokenizer, model.text_encoder, model.device
    ids = tok(list(prompts), padding="max_length", max_length=tok.model_max_length, truncation=True,
              return_tensors="pt").input_ids
    return enc(ids.to(dev))[0].to(dev, torch.float32)


def _schedule(model, n_steps):
    sched = model.scheduler
    sched.set_timesteps(n_steps)
    ts = [int(t) for t in sched.timesteps]
    ratio = sched.config.num_train_timesteps // sched.num_inference_steps
    ac, fa = sched.alphas_cumprod, sched.final_alpha_cumprod
    inv_t = [ts[len(ts) - i - 1] for i in range(n_steps)]                       # inversion.py:315
    inv_co = [step_coefficients(ac, fa, min(t - ratio, 999), t) for t in inv_t]  # next_step, inversion.py:262-270
    fwd_co = [step_coefficients(ac, fa, t, t - ratio) for t in ts]               # prev_step / scheduler.step
    return ts, inv_t, inv_co, fwd_co


def _c_ints(v):
    return (C.c_int32 * len(v))(*[int(x) for x in v])


def _c_coefs(cos):
    flat = [float(x) for co in cos for x in co]
    return (C.c_float * len(flat))(*flat)


def run_loop(model, mode, n_steps, rows, images, t_list, coefs, guidance, ctx, x, traj=None, loss=None,
             loss_scales=None, add_mask=0, ctrls=None, blends=None, blend_start=0):
    """One `pnp_run_loop` call on the current torch stream; x is updated in place."""
    for name, t in (("ctx", ctx), ("x", x), ("traj", traj), ("loss", loss)):
        if t is not None and not (t.is_cuda and t.device == model.device and t.dtype == torch.float32 and t.is_contiguous()):
            raise _lib.PnpError(f"run_loop: {name} must be a contiguous float32 CUDA tensor on {model.device}")
    a = _lib.LoopArgs()
    a.mode, a.n_steps, a.rows, a.images = mode, n_steps, rows, images
    t_arr, co_arr = _c_ints(t_list), _c_coefs(coefs)
    a.t_host, a.coef_host = t_arr, co_arr
    a.guidance = float(guidance)
    a.ctx_dev, a.x_dev = ctx.data_ptr(), x.data_ptr()
    a.traj_dev = traj.data_ptr() if traj is not None else None
    a.loss_dev = loss.data_ptr() if loss is not None else None
    ls_arr = None
    if loss_scales is not None:
        ls_arr = (C.c_float * n_steps)(*[float(v) for v in loss_scales])
        a.loss_scale_host = ls_arr
    a.add_mask = int(add_mask)
    if ctrls is not None:
        a.ctrl_host = ctrls
    if blends is not None:
        a.blend_host = blends
        a.n_blend = len(blends)
        a.blend_start = int(blend_start)
    _lib.check(_lib.load().pnp_run_loop(model.unet.handle, C.byref(a), _lib.current_stream_ptr()))
    model.unet._ctx_ref = None  # the loop installed its own context: the per-call cache of FusedUNet is stale
    return x


class BatchedDirectInversionP2P:
    def __init__(self, model, num_ddim_steps: int = 50):
        self.model = model
        self.num_ddim_steps = num_ddim_steps

    # -- the four loops ------------------------------------------------------------------------------------------
    def invert(self, latents: torch.Tensor, prompts_src: Sequence[str], prompts_tar: Sequence[str], guidance_scale=7.5,
               inverse_guidance_scale=None, loss_scales=None):
        """DirectInversion.invert for L images: (x_stars (n+1,L,..), noise_loss (n,2L,..)); `inverse_guidance_scale`
        not None is invert_with_guidance_scale_vary_guidance (inversion.py:412-419), `loss_scales` the not_full /
        skip_step ablations (:478-526)."""
        m, n = self.model, self.num_ddim_steps
        L = latents.shape[0]
        if 4 * L > m.unet.max_batch:
            raise _lib.PnpError(f"{L} images need a UNet batch of {4 * L}; this model handle was built with max_batch "
                                f"{m.unet.max_batch}")
        ts, inv_t, inv_co, fwd_co = _schedule(m, n)
        m.unet.set_controller(None)
        uncond = _encode_rows(m, [""] * (2 * L))
        cond = _encode_rows(m, list(prompts_src) + list(prompts_tar))
        self._ctx = torch.cat([uncond, cond]).contiguous()
        z = latents.to(m.device, torch.float32).contiguous().clone()
        x_stars = torch.empty((n + 1, L, 4, 64, 64), device=m.device, dtype=torch.float32)
        if inverse_guidance_scale is None:
            ctx_inv, g_inv = cond[:L].contiguous(), 0.0            # ddim_loop: the source prompt alone
        elif float(inverse_guidance_scale) == 0.0:
            ctx_inv, g_inv = uncond[:L].contiguous(), 0.0          # u + 0 * (c - u) is u: one unconditional call
        else:
            ctx_inv, g_inv = torch.cat([uncond[:L], cond[:L]]).contiguous(), float(inverse_guidance_scale)
        run_loop(m, _lib.PNP_LOOP_INVERT, n, L, L, inv_t, inv_co, g_inv, ctx_inv, z, traj=x_stars)
        self._sched = (ts, fwd_co)
        noise_loss = torch.empty((n, 2 * L, 4, 64, 64), device=m.device, dtype=torch.float32)
        cur = torch.cat([x_stars[n]] * 2).contiguous()
        run_loop(m, _lib.PNP_LOOP_OFFSET, n, 2 * L, L, ts, fwd_co, guidance_scale, self._ctx, cur, traj=x_stars,
                 loss=noise_loss, loss_scales=loss_scales)
        return x_stars, noise_loss

    def forward(self, x_T: torch.Tensor, noise_loss: Optional[torch.Tensor], guidance_scale=7.5,
                controllers: Optional[List[AttentionControlEdit]] = None, add_target=False):
        """direct_inversion_p2p_guidance_forward[_add_target] for L images; `controllers` = one (source, target)
        controller per image, or None for the AttentionStore (reconstruction) pass."""
        m, n = self.model, self.num_ddim_steps
        L = x_T.shape[0]
        ts, fwd_co = self._sched
        x = torch.cat([x_T] * 2).contiguous()
        ctrls = blends = None
        blend_start = 0
        if controllers is not None:
            if len(controllers) != L or L > _lib.PNP_MAX_SLOTS:
                raise _lib.PnpError(f"need one controller per image and at most {_lib.PNP_MAX_SLOTS} images per batch")
            lib = _lib.load()
            _lib.check(lib.pnp_store_reset(m.unet.handle, _lib.current_stream_ptr()))
            ctrls = (_lib.AttnCtrl * n)()
            nrow = 2 * L  # first conditional row of the UNet batch
            for i in range(n):
                lib.pnp_attn_ctrl_init(C.byref(ctrls[i]))
                for img, c in enumerate(controllers):
                    c.fill_pair(ctrls[i], nrow + img, nrow + L + img, 0, img, store_src=2 * img, store_tgt=2 * img + 1)
                    c.after_unet_call()
            with_blend = [c.local_blend is not None for c in controllers]
            if any(with_blend):
                if not all(with_blend):
                    raise _lib.PnpError("either every image of a batch uses LocalBlend or none")
                blends = (_lib.BlendDesc * L)(*[c.local_blend.blend_desc(img, L + img, 2 * img, 2 * img + 1)
                                                for img, c in enumerate(controllers)])
                starts = {c.local_blend.start_blend for c in controllers}
                if len(starts) != 1:
                    raise _lib.PnpError("the images of a batch must share LocalBlend.start_blend")
                blend_start = starts.pop()
        mask = 0
        if noise_loss is not None:
            mask = (1 << (2 * L)) - 1 if add_target else (1 << L) - 1
        run_loop(m, _lib.PNP_LOOP_FORWARD, n, 2 * L, L, ts, fwd_co, guidance_scale, self._ctx, x, loss=noise_loss,
                 add_mask=mask, ctrls=ctrls, blends=blends, blend_start=blend_start)
        return x

    # -- the editor-level composition (models/p2p_editor.py:415-479) ----------------------------------------------
    def edit(self, latents, prompts_src, prompts_tar, guidance_scale=7.5, cross_replace_steps=0.4,
             self_replace_steps=0.6, blend_word=None, eq_params=None, is_replace_controller=False, add_target=False,
             inverse_guidance_scale=None, forward_guidance_scale=None, loss_scales=None, add_source=False,
             per_image_params=False, device="cuda", minimal=False) -> BatchEditResult:
        """`blend_word` / `eq_params` apply to every image, or are lists with one entry per image when
        `per_image_params` is set.

        `minimal=True` skips the reconstruction pass: the editor decodes only its SOURCE row, and the rectification pins
        that row to x_stars[0] = the inverted latent (tests assert 2e-5), so the 200 sample-forwards of that pass change
        nothing the method returns: 450 instead of 650 per image.  SURVEY.md section 8d also proposes dropping the target
        rows of the offset pass ("minimal-350"); that is exact only in exact arithmetic.  The invariant holds on this
        engine because the offset pass and the edit pass evaluate the source rows in the SAME batch composition - rows of
        one call are bit-reproducible - while another batch size is another realisation of the fp16 rounding noise
        (DESIGN.md section 2): offsets computed at batch 2L no longer cancel the edit pass's batch-4L predictions and
        the source branch drifts away (measured: O(1) after 4 steps).  So the offsets keep their full batch."""
        L = latents.shape[0]
        blends = list(blend_word) if per_image_params else [blend_word] * L
        eqs = list(eq_params) if per_image_params else [eq_params] * L
        if not (len(prompts_src) == len(prompts_tar) == len(blends) == len(eqs) == L):
            raise ValueError("one source prompt, target prompt, blend_word and eq_params per image")
        fwd_g = guidance_scale if forward_guidance_scale is None else forward_guidance_scale
        x_stars, noise_loss = self.invert(latents, prompts_src, prompts_tar, guidance_scale=fwd_g,
                                          inverse_guidance_scale=inverse_guidance_scale, loss_scales=loss_scales)
        x_T = x_stars[self.num_ddim_steps]
        fwd_loss = noise_loss
        if add_source:  # p2p_editor.py:930-932: the source branch's offset on both branches
            fwd_loss = torch.cat([noise_loss[:, :L]] * 2, dim=1).contiguous()
        if minimal:
            recon = torch.cat([x_stars[0], x_stars[0]]).contiguous()  # source rows = z0 by the invariant; targets not produced
        else:
            recon = self.forward(x_T, fwd_loss, fwd_g, controllers=None, add_target=add_target or add_source)
        controllers = [make_controller(pipeline=self.model, prompts=[prompts_src[i], prompts_tar[i]],
                                       is_replace_controller=is_replace_controller,
                                       cross_replace_steps={"default_": cross_replace_steps},
                                       self_replace_steps=self_replace_steps, blend_words=blends[i],
                                       equilizer_params=eqs[i], num_ddim_steps=self.num_ddim_steps, device=device)
                       for i in range(L)]
        out = self.forward(x_T, fwd_loss, fwd_g, controllers=controllers, add_target=add_target or add_source)
        return BatchEditResult(x_stars, noise_loss, recon, out)
