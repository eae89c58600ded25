"""MasaCtrl (mutual self-attention) path with PnP-Inversion rectification -- SURVEY.md section 8 rows a13, a14.

Mirrors, with the same names and call signatures:
  * `models/masactrl/masactrl_utils.py`: `AttentionBase` (:13-36), `regiter_attention_editor_diffusers` (:79-144, the
    reference's spelling is kept);
  * `models/masactrl/masactrl.py`: `MutualSelfAttentionControl` (:14-72);
  * `models/masactrl/diffuser_utils.py`: `MasaCtrlPipeline.__call__` loop (:162-186, rectification :183-184),
    `.invert` (:195-270), `.step` (:39-57), `.next_step` (:16-37);
  * `run_editing_masactrl.py`: `MasaCtrlEditor` (:58-168).

The editor callback `editor(q, k, v, sim, attn, ...)` of the reference recomputes attention from materialised tensors;
here `MutualSelfAttentionControl` lowers itself to the K/V source-row indirection of the fused self-attention kernels
(for both CFG halves the queries of every image attend to the SOURCE image's keys and values, masactrl.py:63-70).
"""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass
from typing import List, Optional

import torch

from . import _lib
from .inversion import DirectInversion
from .ptp_utils import load_512
from .scheduler import fused_step, step_coefficients

NUM_ATT_LAYERS = 32


class AttentionBase:
    def __init__(self):
        self.cur_step = 0
        self.num_att_layers = -1
        self.cur_att_layer = 0

    def after_step(self):
        pass

    def reset(self):
        self.cur_step = 0
        self.cur_att_layer = 0

    # one fused UNet call == num_att_layers invocations of the reference's __call__ (masactrl_utils.py:23-31)
    def after_unet_call(self):
        self.cur_att_layer = 0
        self.cur_step += 1
        self.after_step()

    def descriptor(self, batch) -> Optional[_lib.AttnCtrl]:
        return None

    def __call__(self, *a, **k):
        raise _lib.PnpError("attention editors are compiled into kernel modes; there is no materialised-attention callback")


class MutualSelfAttentionControl(AttentionBase):
    MODEL_TYPE = {"SD": 16, "SDXL": 70}

    def __init__(self, start_step=4, start_layer=10, layer_idx=None, step_idx=None, total_steps=50, model_type="SD",
                 images=1):
        super().__init__()
        self.images = images  # L images per call, rows prompt-major inside each CFG half (batched.py)
        if model_type != "SD":
            raise NotImplementedError("only the SD-1.x UNet (16 transformer blocks) is implemented")
        self.total_steps = total_steps
        self.total_layers = self.MODEL_TYPE.get(model_type, 16)
        self.start_step = start_step
        self.start_layer = start_layer
        self.layer_idx = layer_idx if layer_idx is not None else list(range(start_layer, self.total_layers))
        self.step_idx = step_idx if step_idx is not None else list(range(start_step, total_steps))
        lo, hi = min(self.layer_idx), max(self.layer_idx) + 1
        if sorted(self.layer_idx) != list(range(lo, hi)):
            raise NotImplementedError("layer_idx must be a contiguous range of transformer blocks")
        self._layers = (lo, hi)

    def descriptor(self, batch):
        if self.cur_step not in self.step_idx:
            return None
        if batch % 2:
            raise _lib.PnpError("MutualSelfAttentionControl expects a CFG batch (uncond half + cond half)")
        n = batch // 2
        c = _lib.new_ctrl()
        c.self_layer_lo, c.self_layer_hi = self._layers
        c.self_max_tokens = 1 << 30
        if n % self.images:
            raise _lib.PnpError(f"a CFG half of {n} rows does not hold {self.images} images")
        for r in range(batch):
            # ku[:num_heads] / kc[:num_heads]: the first prompt (= the source) of each CFG half, for the same image
            src = (0 if r < n else n) + (r % n) % self.images
            c.self_k_row[r] = src
            c.self_v_row[r] = src
        return c


def regiter_attention_editor_diffusers(model, editor: AttentionBase):
    """masactrl_utils.py:79-144 patches every `Attention.forward`; here the editor is bound to the fused UNet."""
    model.unet.set_controller(editor)
    editor.num_att_layers = NUM_ATT_LAYERS


def _encode(model, prompts: List[str], uncond_text: str = ""):
    tok, enc, dev = model.tokenizer, model.text_encoder, model.device
    text = enc(tok(prompts, padding="max_length", max_length=77, return_tensors="pt").input_ids.to(dev))[0]
    un = enc(tok([uncond_text] * len(prompts), padding="max_length", max_length=77, return_tensors="pt").input_ids.to(dev))[0]
    return torch.cat([un, text]).to(dev, torch.float32).contiguous()


@torch.no_grad()
def masactrl_sample(model, prompt, latents, num_inference_steps=50, guidance_scale=7.5, noise_loss_list=None):
    """`MasaCtrlPipeline.__call__` (diffuser_utils.py:90-193) up to the final latents: per step one fused UNet call
    (B = 2 * prompts) and one fused epilogue launch (CFG + own `step` :39-57 + rectification :183-184)."""
    if isinstance(prompt, str):
        prompt = [prompt]
    n = len(prompt)
    assert latents.shape == (n, 4, 64, 64), f"The shape of input latent tensor {latents.shape} should equal to predefined one."
    if not guidance_scale > 1.0:
        raise NotImplementedError("guidance_scale <= 1 (no CFG batch) is not on the PnP-inversion path")
    context = _encode(model, prompt)
    sched = model.scheduler
    sched.set_timesteps(num_inference_steps)
    ratio = sched.config.num_train_timesteps // sched.num_inference_steps
    lat = latents.to(model.device, torch.float32).contiguous()
    for i, t in enumerate(sched.timesteps):
        tt = int(t)
        eps = model.unet(torch.cat([lat] * 2), tt, encoder_hidden_states=context)["sample"]
        prev = tt - ratio
        # `.step` uses `prev_timestep > 0` (strict): prev == 0 selects final_alpha_cumprod == alphas_cumprod[0]
        co = step_coefficients(sched.alphas_cumprod, sched.final_alpha_cumprod, tt, prev if prev > 0 else -1)
        nl = noise_loss_list[i].contiguous() if noise_loss_list is not None else None
        lat = fused_step(model.unet.handle, lat, eps[n:], co, eps_u=eps[:n], guidance=guidance_scale, noise_loss=nl,
                         add_mask=1 if nl is not None else 0)
    return lat


@torch.no_grad()
def masactrl_invert(model, latent, prompt, num_inference_steps=50, guidance_scale=7.5, return_intermediates=False):
    """`MasaCtrlPipeline.invert` (diffuser_utils.py:195-270): DDIM inversion WITH classifier-free guidance."""
    if isinstance(prompt, str):
        prompt = [prompt] * latent.shape[0]
    n = len(prompt)
    context = _encode(model, prompt)
    sched = model.scheduler
    sched.set_timesteps(num_inference_steps)
    ratio = sched.config.num_train_timesteps // sched.num_inference_steps
    lat = latent.to(model.device, torch.float32).expand(n, -1, -1, -1).contiguous()
    start = lat
    lats = [lat]
    for t in reversed(sched.timesteps):
        tt = int(t)
        eps = model.unet(torch.cat([lat] * 2), tt, encoder_hidden_states=context)["sample"]
        co = step_coefficients(sched.alphas_cumprod, sched.final_alpha_cumprod, min(tt - ratio, 999), tt)
        lat = fused_step(model.unet.handle, lat, eps[n:], co, eps_u=eps[:n], guidance=guidance_scale)
        lats.append(lat)
    return (lat, lats) if return_intermediates else (lat, start)


@dataclass
class MasaCtrlResult:
    """Returned instead of the 4-panel PIL strip when the model handle has no VAE."""
    x_stars: Optional[List[torch.Tensor]]
    noise_loss_list: Optional[List[torch.Tensor]]
    latents_fixed: torch.Tensor   # (1,4,64,64) direct synthesis with the target prompt
    latents: torch.Tensor         # (2,4,64,64): [reconstruction of the source, MasaCtrl edit]


@dataclass
class MasaCtrlBatchResult:
    x_stars: torch.Tensor       # (n+1, L, 4,64,64)
    noise_loss: torch.Tensor    # (n, 2L, 4,64,64) prompt-major
    latents_fixed: torch.Tensor  # (L, 4,64,64)
    latents: torch.Tensor       # (2L, 4,64,64): rows [0,L) reconstructions of the sources, [L,2L) MasaCtrl edits


class MasaCtrlEditor:
    def __init__(self, method_list, device, num_ddim_steps=50, model=None) -> None:
        if model is None:
            # the reference downloads CompVis/stable-diffusion-v1-4 here (run_editing_masactrl.py:69-70); offline the
            # checkpoint directory comes from PNP_SD_CHECKPOINT, else the seeded random-init stand-in is used
            import os

            from .model import FusedModel
            ckpt = os.environ.get("PNP_SD_CHECKPOINT")
            model = (FusedModel.from_pretrained(ckpt, device=str(device)) if ckpt
                     else FusedModel.synthetic(device=str(device), with_vae=True))
        self.device = device
        self.method_list = method_list
        self.num_ddim_steps = num_ddim_steps
        self.model = model
        self.scheduler = model.scheduler
        self.model.scheduler.set_timesteps(self.num_ddim_steps)

    def __call__(self, edit_method, image_path, prompt_src, prompt_tar, guidance_scale, step=4, layper=10):
        if edit_method == "ddim+masactrl":
            res = self.edit_image_ddim_MasaCtrl(image_path, prompt_src, prompt_tar, guidance_scale, step=step, layper=layper)
        elif edit_method == "directinversion+masactrl":
            res = self.edit_image_directinversion_MasaCtrl(image_path, prompt_src, prompt_tar, guidance_scale,
                                                           step=step, layper=layper)
        else:
            raise NotImplementedError(f"No edit method named {edit_method}")
        if isinstance(image_path, torch.Tensor) or self.model.vae is None:
            return res  # latents in, latents out
        return self._strip(load_512(image_path), prompt_src, prompt_tar, res.latents)

    def _strip(self, image_gt, prompt_src, prompt_tar, latents2):
        """run_editing_masactrl.py:121-129: [instruction | source | reconstruction (row 0) | MasaCtrl edit (row -1)]."""
        import numpy as np
        from PIL import Image

        from .ptp_utils import latent2image, txt_draw

        imgs = latent2image(self.model.vae, latents2)
        instruct = txt_draw(f"source prompt: {prompt_src}\ntarget prompt: {prompt_tar}")
        return Image.fromarray(np.concatenate((instruct, image_gt, imgs[0], imgs[-1]), axis=1))

    def edit_batch_images(self, image_paths, prompts_src, prompts_tar, guidance_scale=7.5, step=4, layper=10):
        """`directinversion+masactrl` for a list of images in one pass -> list of PIL strips."""
        from .ptp_utils import image2latent

        gts = [load_512(p) for p in image_paths]
        lat = torch.cat([image2latent(self.model.vae, g) for g in gts]).to(self.model.device, torch.float32)
        res = self.edit_batch(lat, list(prompts_tar), guidance_scale=guidance_scale, step=step, layper=layper)
        L = len(gts)
        return [self._strip(g, prompts_src[i], prompts_tar[i], res.latents[[i, L + i]]) for i, g in enumerate(gts)]

    def edit_batch(self, latents, prompts_tar, guidance_scale=7.5, step=4, layper=10):
        """`directinversion+masactrl` (run_editing_masactrl.py:89-129) for L images per call - BASELINE config 4
        (batch 4 -> UNet batch 16) - with every 50-step loop inside libpnpinv.so (`pnp_run_loop`): inversion (B = L),
        offsets with prompts ["", target] (B = 4L), direct synthesis with the target prompt (B = 2L), MasaCtrl pass
        (B = 4L) = 550 UNet sample-forwards per image.  Returns MasaCtrlBatchResult."""
        from . import batched as bt

        m, n = self.model, self.num_ddim_steps
        L = latents.shape[0]
        b = bt.BatchedDirectInversionP2P(m, n)
        x_stars, noise_loss = b.invert(latents, [""] * L, list(prompts_tar), guidance_scale=guidance_scale)
        ts, fwd_co = b._sched
        x_T = x_stars[n]
        ctx = b._ctx  # [uncond 2L | "" x L, target x L]
        ctx_fixed = torch.cat([ctx[:L], ctx[3 * L:]]).contiguous()
        fixed = x_T.clone().contiguous()
        bt.run_loop(m, _lib.PNP_LOOP_FORWARD, n, L, L, ts, fwd_co, guidance_scale, ctx_fixed, fixed)
        editor = MutualSelfAttentionControl(step, layper, total_steps=n, images=L)
        ctrls = (_lib.AttnCtrl * n)()
        lib = _lib.load()
        for i in range(n):
            d = editor.descriptor(4 * L)
            if d is None:
                lib.pnp_attn_ctrl_init(C.byref(ctrls[i]))
            else:
                ctrls[i] = d
            editor.after_unet_call()
        out = torch.cat([x_T] * 2).contiguous()
        bt.run_loop(m, _lib.PNP_LOOP_FORWARD, n, 2 * L, L, ts, fwd_co, guidance_scale, ctx, out, loss=noise_loss,
                    add_mask=(1 << L) - 1, ctrls=ctrls)
        return MasaCtrlBatchResult(x_stars, noise_loss, fixed, out)

    def _latent(self, image_path):
        if isinstance(image_path, torch.Tensor) and image_path.dim() == 4:
            return image_path
        if self.model.vae is None:
            raise _lib.PnpError("an image was given but the model handle has no VAE; pass a (1,4,64,64) latent")
        from .ptp_utils import image2latent

        return image2latent(self.model.vae, load_512(image_path))

    def edit_image_directinversion_MasaCtrl(self, image_path, prompt_src, prompt_tar, guidance_scale, step=4, layper=10):
        """run_editing_masactrl.py:89-129.  NB the source prompt of the inversion is the EMPTY string (:93)."""
        z0 = self._latent(image_path)
        prompts = ["", prompt_tar]
        inv = DirectInversion(model=self.model, num_ddim_steps=self.num_ddim_steps)
        _, _, x_stars, noise_loss_list = inv.invert(image_gt=z0, prompt=prompts, guidance_scale=guidance_scale)
        x_t = x_stars[-1]
        editor = AttentionBase()
        regiter_attention_editor_diffusers(self.model, editor)
        fixed = masactrl_sample(self.model, [prompt_tar], x_t, self.num_ddim_steps, guidance_scale, None)
        editor = MutualSelfAttentionControl(step, layper, total_steps=self.num_ddim_steps)
        regiter_attention_editor_diffusers(self.model, editor)
        out = masactrl_sample(self.model, prompts, x_t.expand(len(prompts), -1, -1, -1), self.num_ddim_steps,
                              guidance_scale, noise_loss_list)
        regiter_attention_editor_diffusers(self.model, AttentionBase())
        return MasaCtrlResult(x_stars, noise_loss_list, fixed, out)

    def edit_image_ddim_MasaCtrl(self, image_path, prompt_src, prompt_tar, guidance_scale, step=4, layper=10):
        """run_editing_masactrl.py:131-168 (plain DDIM inversion with CFG, no rectification)."""
        z0 = self._latent(image_path)
        prompts = ["", prompt_tar]
        regiter_attention_editor_diffusers(self.model, AttentionBase())
        start_code, _ = masactrl_invert(self.model, z0, "", self.num_ddim_steps, guidance_scale, return_intermediates=True)
        start_code = start_code.expand(len(prompts), -1, -1, -1)
        editor = AttentionBase()
        regiter_attention_editor_diffusers(self.model, editor)
        fixed = masactrl_sample(self.model, [prompt_tar], start_code[-1:], self.num_ddim_steps, guidance_scale)
        editor = MutualSelfAttentionControl(step, layper, total_steps=self.num_ddim_steps)
        regiter_attention_editor_diffusers(self.model, editor)
        out = masactrl_sample(self.model, prompts, start_code, self.num_ddim_steps, guidance_scale)
        regiter_attention_editor_diffusers(self.model, AttentionBase())
        return MasaCtrlResult(None, None, fixed, out)
