"""Plug-and-Play diffusion features with the PnP-Inversion source branch -- SURVEY.md section 8f-4.

Mirror of `run_editing_pnp.py`: `Preprocess.ddim_inversion` / `ddim_sample` / `extract_latents` (:88-148),
`register_attention_control_efficient` (:176-242: self-attention Q and K of the source sample injected into the
unconditional and conditional samples in decoder blocks 4-11), `register_conv_control_efficient` (:244-294: the conv2
output of up_blocks[1].resnets[1] injected), `PNP.denoise_step` / `run_pnp` / `sample_loop` (:344-392), and the two
editors `edit_image_ddim_PnP` (:412-432) / `edit_image_directinversion_PnP` (:434-452).

The reference patches module forwards; here both injections are descriptor fields of the fused UNet: the Q/K injection
is the self-attention row indirection the P2P / MasaCtrl controllers already use (transformer blocks 8..15 = up_blocks
1.attentions[1,2], 2.*, 3.*), the feature injection is `conv_src_row` (the implicit-GEMM conv reads its 3x3 taps from the
source row, csrc/gemm_sm100.cu).  Batch rows: [source x L | unconditional x L | conditional x L].
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import List

import numpy as np
import torch

from . import _lib
from .scheduler import fused_step, step_coefficients

NEGATIVE_PROMPT = "ugly, blurry, black, low res, unrealistic"  # run_editing_pnp.py:381
QK_BLOCKS = (8, 16)  # transformer blocks with Q/K injection: res_dict {1: [1, 2], 2: [0, 1, 2], 3: [0, 1, 2]} (:236)


def pnp_timesteps(n_steps: int, steps_offset: int = 1) -> List[int]:
    """The scheduler of runwayml/stable-diffusion-v1-5 (`steps_offset` 1): 981, 961, ..., 1 for 50 steps."""
    ratio = 1000 // n_steps
    return [int(t) + steps_offset for t in (np.arange(0, n_steps) * ratio).round()[::-1]]


class PnPController:
    """Descriptor source for one UNet call at timestep t (register_time :150-174 + the two injection schedules)."""

    def __init__(self, images: int, qk_timesteps, conv_timesteps):
        self.images = images
        self.qk_timesteps = set(int(t) for t in qk_timesteps)
        self.conv_timesteps = set(int(t) for t in conv_timesteps)
        self.t = None

    def descriptor(self, batch):
        L = self.images
        if batch != 3 * L:
            raise _lib.PnpError(f"PnP expects the batch [source, uncond, cond] x {L} images, got {batch}")
        c = _lib.new_ctrl()
        if self.t in self.qk_timesteps or self.t == 1000:
            c.self_layer_lo, c.self_layer_hi = QK_BLOCKS
            c.self_max_tokens = 1 << 30
            for r in range(L, 3 * L):
                c.self_q_row[r] = r % L
                c.self_k_row[r] = r % L
        if self.t in self.conv_timesteps or self.t == 1000:
            for r in range(L, 3 * L):
                c.conv_src_row[r] = r % L
        return c

    def after_unet_call(self):
        pass


@dataclass
class PnPResult:
    inverted_x: List[torch.Tensor]
    latent_reconstruction: List[torch.Tensor]
    latents: torch.Tensor  # (L,4,64,64) edited


class PnPFeaturesEditor:
    def __init__(self, model, num_ddim_steps: int = 50):
        self.model = model
        self.n = num_ddim_steps
        self.timesteps = pnp_timesteps(num_ddim_steps)

    def _embed(self, prompts):
        tok, enc, dev = self.model.tokenizer, self.model.text_encoder, self.model.device
        ids = tok(list(prompts), padding="max_length", max_length=tok.model_max_length, truncation=True,
                  return_tensors="pt").input_ids
        return enc(ids.to(dev))[0].to(dev, torch.float32)

    def _alpha(self, t):
        s = self.model.scheduler
        return s.alphas_cumprod[t] if t >= 0 else s.final_alpha_cumprod

    def _co(self, t_from, t_to):
        s = self.model.scheduler
        return step_coefficients(s.alphas_cumprod, s.final_alpha_cumprod, t_from, t_to)

    @torch.no_grad()
    def ddim_inversion(self, cond, latent):
        """Preprocess.ddim_inversion (:88-110): x_t = mu_t x0 + sigma_t eps with eps = unet(x, t) evaluated at the
        TARGET timestep and x0 from the previous alpha (final_alpha for the first step)."""
        m = self.model
        m.unet.set_controller(None)
        lat = latent.to(m.device, torch.float32).contiguous()
        out = [lat]
        ts = list(reversed(self.timesteps))
        for i, t in enumerate(ts):
            eps = m.unet(lat, t, encoder_hidden_states=cond)["sample"]
            lat = fused_step(m.unet.handle, lat, eps, self._co(ts[i - 1] if i > 0 else -1, t))
            out.append(lat)
        return out

    @torch.no_grad()
    def ddim_sample(self, x, cond):
        """Preprocess.ddim_sample (:112-134)."""
        m = self.model
        m.unet.set_controller(None)
        lat = x.contiguous()
        out = []
        ts = self.timesteps
        for i, t in enumerate(ts):
            eps = m.unet(lat, t, encoder_hidden_states=cond)["sample"]
            lat = fused_step(m.unet.handle, lat, eps, self._co(t, ts[i + 1] if i < len(ts) - 1 else -1))
            out.append(lat)
        return out

    def extract_latents(self, latent, inversion_prompts):
        """Preprocess.extract_latents (:136-148) on latents: (inverted_x, latent_reconstruction reversed)."""
        cond = self._embed(inversion_prompts).contiguous()
        inverted_x = self.ddim_inversion(cond, latent)
        rec = self.ddim_sample(inverted_x[-1], cond)
        rec.reverse()
        return inverted_x, rec

    @torch.no_grad()
    def run_pnp(self, noisy_latent, target_prompts, guidance_scale=7.5, pnp_f_t=0.8, pnp_attn_t=0.5):
        """PNP.run_pnp + sample_loop + denoise_step (:344-392): per step ONE UNet call of batch 3L
        [noisy_latent[-1-i] | x | x] with contexts ["" | negative prompt | target prompt]."""
        m = self.model
        L = noisy_latent[-1].shape[0]
        ts = self.timesteps
        qk = ts[: int(self.n * pnp_attn_t)] if pnp_attn_t >= 0 else []
        conv = ts[: int(self.n * pnp_f_t)] if pnp_f_t >= 0 else []
        ctrl = PnPController(L, qk, conv)
        m.unet.set_controller(ctrl)
        ctx = torch.cat([self._embed([""] * L), self._embed([NEGATIVE_PROMPT] * L), self._embed(list(target_prompts))]).contiguous()
        x = noisy_latent[-1].to(m.device, torch.float32).contiguous()
        ratio = 1000 // self.n
        for i, t in enumerate(ts):
            ctrl.t = t
            src = noisy_latent[-1 - i].to(m.device, torch.float32)
            eps = m.unet(torch.cat([src, x, x]).contiguous(), t, encoder_hidden_states=ctx)["sample"]
            x = fused_step(m.unet.handle, x, eps[2 * L:], self._co(t, t - ratio), eps_u=eps[L:2 * L], guidance=guidance_scale)
        m.unet.set_controller(None)
        return x

    # ---- editors (latents in, latents out; with a VAE on the handle: image path in, the reference's strip out)
    def _edit(self, image_path, prompt_src, prompt_tar, guidance_scale, use_inverted):
        from .ptp_utils import image2latent, latent2image, load_512, txt_draw

        is_latent = isinstance(image_path, torch.Tensor) and image_path.dim() == 4
        image_gt = None if is_latent else load_512(image_path)
        z0 = image_path if is_latent else image2latent(self.model.vae, image_gt)
        L = z0.shape[0]
        srcs = [prompt_src] * L if isinstance(prompt_src, str) else list(prompt_src)
        tars = [prompt_tar] * L if isinstance(prompt_tar, str) else list(prompt_tar)
        inverted_x, rec = self.extract_latents(z0, srcs)
        out = self.run_pnp(inverted_x if use_inverted else rec, tars, guidance_scale)
        if is_latent or self.model.vae is None:
            return PnPResult(inverted_x, rec, out)
        from PIL import Image

        recon_lat = inverted_x[1] if use_inverted else rec[0]  # :447 decodes inverted_x[1], :426 the sampled x_0
        panel = latent2image(self.model.vae, recon_lat)[0]
        edit = latent2image(self.model.vae, out)[0]
        instruct = txt_draw(f"source prompt: {prompt_src}\ntarget prompt: {prompt_tar}")
        return Image.fromarray(np.concatenate((instruct, image_gt, panel, edit), axis=1))

    def edit_image_ddim_PnP(self, image_path, prompt_src, prompt_tar, guidance_scale=7.5):
        """:412-432: the source branch follows the DDIM reconstruction trajectory."""
        return self._edit(image_path, prompt_src, prompt_tar, guidance_scale, use_inverted=False)

    def edit_image_directinversion_PnP(self, image_path, prompt_src, prompt_tar, guidance_scale=7.5):
        """:434-452: PnP-Inversion - the source branch is PINNED to the inversion trajectory itself."""
        return self._edit(image_path, prompt_src, prompt_tar, guidance_scale, use_inverted=True)

    def __call__(self, edit_method, image_path, prompt_src, prompt_tar, guidance_scale=7.5):
        if edit_method == "ddim+pnp":
            return self.edit_image_ddim_PnP(image_path, prompt_src, prompt_tar, guidance_scale)
        if edit_method == "directinversion+pnp":
            return self.edit_image_directinversion_PnP(image_path, prompt_src, prompt_tar, guidance_scale)
        raise NotImplementedError(f"No edit method named {edit_method}")
