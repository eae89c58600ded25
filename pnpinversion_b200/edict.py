"""EDICT coupled dual-latent exact inversion (+ its Prompt-to-Prompt variant) -- SURVEY.md section 8 row a15.

Mirror of `models/edict/edict_functions.py`: `coupled_stablediffusion` (:707-956; loop :851-936, leapfrog order :862-880,
mixing layers :854-859 / :931-936), `forward_step` (:621-650), `reverse_step` (:653-684), `get_alpha_and_beta` (:599-617),
`init_attention_edit` (:225-247, difflib opcodes -> mask / indices), `EDICT_editing` (:56-115) and
`run_editing_edict.py::edit_image_edict_p2p` (:32-61).

The reference runs three sequential B=1 fp64 UNet calls per coupled latent (uncond, cond, cond_edit) and passes the
source pass's attention probabilities to the edit pass through module attributes.  Here the three passes are ONE fused
UNet call of batch 3 on the same latent, and "reuse the saved probabilities" is the kernel mode the P2P controllers
already use: self-attention of the edit row takes Q,K of the source row in every layer (:269-270), cross-attention
becomes P*(1-mask) + P_src[..., indices]*mask (:266-268).  The DDIM algebra is the fused step epilogue with the
alpha-quotient coefficients; the mixing layers are `pnp_edict_mix`.  Arithmetic is fp16/fp32 (the reference is fp64):
tolerances are stated in tests/test_gpu_edict.py.
"""
from __future__ import annotations

import ctypes as C
from difflib import SequenceMatcher
from typing import List, Optional, Sequence

import torch

from . import _lib
from .scheduler import fused_step

MAX_TOKENS = 77


def attention_edit_tables(tokens: Sequence[int], tokens_edit: Sequence[int]):
    """init_attention_edit (edict_functions.py:225-247): mask[j]=1 and indices[j]=source position for target tokens in an
    'equal' block or an equal-length 'replace' block of the difflib alignment."""
    mask = torch.zeros(MAX_TOKENS)
    indices = torch.zeros(MAX_TOKENS, dtype=torch.long)
    target = torch.arange(MAX_TOKENS, dtype=torch.long)
    for name, a0, a1, b0, b1 in SequenceMatcher(None, list(tokens), list(tokens_edit)).get_opcodes():
        if b0 < MAX_TOKENS:
            if name == "equal" or (name == "replace" and a1 - a0 == b1 - b0):
                mask[b0:b1] = 1
                indices[b0:b1] = target[a0:a1]
    return mask, indices


class EdictP2PController:
    """Batch rows [uncond x L | cond(source prompts) x L | cond(edit prompts) x L] on the L latents of one call (the
    reference handles one image: L = 1; config 5 of BASELINE.json batches 8).  `masks` / `indices`: one (77,) table per
    image, or a single table for L = 1."""

    def __init__(self, mask, indices, weights=None):
        self.masks = list(mask) if isinstance(mask, (list, tuple)) else [mask]
        self.indices = list(indices) if isinstance(indices, (list, tuple)) else [indices]
        self.weights = weights if weights is not None else torch.ones(MAX_TOKENS)
        self.num_att_layers = 32
        if len(self.masks) > _lib.PNP_MAX_SLOTS:
            raise _lib.PnpError(f"at most {_lib.PNP_MAX_SLOTS} images per EDICT batch")

    def descriptor(self, batch):
        L = len(self.masks)
        if batch != 3 * L:
            raise _lib.PnpError(f"EDICT P2P expects the batch [uncond, cond, cond_edit] x {L} images, got {batch}")
        c = _lib.new_ctrl()
        c.self_layer_lo, c.self_layer_hi, c.self_max_tokens = 0, 16, 1 << 30  # attn1 reused wholesale in every layer
        for img in range(L):
            src, tgt = L + img, 2 * L + img
            c.self_q_row[tgt] = src
            c.self_k_row[tgt] = src
            c.cross_base_row[tgt] = src
            c.cross_slot[tgt] = img
            c.mapper[img][:] = [int(v) for v in self.indices[img]]
            c.alphas[img][:] = [float(v) for v in self.masks[img]]
            c.equalizer[img][:] = [float(v) for v in self.weights]
            c.cross_alpha[img][:] = [1.0] * MAX_TOKENS
        return c

    def after_unet_call(self):
        pass


def _alpha(sched, t: int):
    """get_alpha_and_beta (:599-617) for integer-valued timesteps; t < 0 selects final_alpha_cumprod."""
    return sched.alphas_cumprod[t] if t >= 0 else sched.final_alpha_cumprod


def step_coeffs(sched, t: int, ratio: int, reverse: bool):
    """forward_step (:646-650): x' = (x - sqrt(1-a_t) e)/q + sqrt(1-a_prev) e ;  reverse_step (:679-684):
    x' = q (x - sqrt(1-a_prev) e) + sqrt(1-a_t) e, with q = sqrt(a_t / a_prev).  Expressed as the fused epilogue's
    (sqrt_a_from, sqrt_1m_a_from, sqrt_a_to, sqrt_1m_a_to)."""
    a_t, a_p = _alpha(sched, t), _alpha(sched, t - ratio)
    q = float((a_t / a_p) ** 0.5)
    if reverse:
        return (1.0, float((1 - a_p) ** 0.5), q, float((1 - a_t) ** 0.5))
    return (q, float((1 - a_t) ** 0.5), 1.0, float((1 - a_p) ** 0.5))


def _embed(model, text):
    """text: one prompt or a list of L prompts -> (ids [L,77], embeddings [L,77,768])"""
    tok, enc = model.tokenizer, model.text_encoder
    ids = tok(text, padding="max_length", max_length=tok.model_max_length, truncation=True, return_tensors="pt")
    return ids, enc(ids.input_ids.to(model.device))[0].to(model.device, torch.float32)


@torch.no_grad()
def coupled_stablediffusion(model, prompt="", prompt_edit=None, null_prompt="", guidance_scale=7.0, steps=50,
                            init_image=None, init_image_strength=1.0, reverse=False, fixed_starting_latent=None,
                            mix_weight=0.93, leapfrog_steps=True, run_baseline=False):
    """Returns the coupled latent pair [x, y] (the reference returns it for reverse=True / return_latents=True; the
    VAE decode of the forward direction is outside this library).  `prompt` / `prompt_edit` may be lists of L prompts
    with (L,4,64,64) latents: L images share every UNet call (BASELINE config 5: batch 8)."""
    if run_baseline:
        raise NotImplementedError("run_baseline=True is plain DDIM, covered by the P2P / MasaCtrl paths")
    dev = model.device
    if init_image is not None:
        assert reverse, "want to be performing deterministic noising"
        lat = [t.clone() for t in init_image] if isinstance(init_image, (list, tuple)) else init_image
        t_limit = steps - int(steps * init_image_strength)
    else:
        assert not reverse, "Need image to reverse from"
        assert fixed_starting_latent is not None, "random starts are not part of the editing path"
        lat = [l.clone() for l in fixed_starting_latent] if isinstance(fixed_starting_latent, (list, tuple)) \
            else fixed_starting_latent.clone()
        t_limit = steps - int(steps * init_image_strength)
    pair = list(lat) if isinstance(lat, list) else [lat.clone(), lat.clone()]
    pair = [p.to(dev, torch.float32).contiguous() for p in pair]
    sched = model.scheduler
    sched.set_timesteps(steps)
    ratio = sched.config.num_train_timesteps // steps

    L = pair[0].shape[0]
    as_list = lambda p: list(p) if isinstance(p, (list, tuple)) else [p] * L
    prompts = as_list(prompt)
    if len(prompts) != L:
        raise ValueError(f"{len(prompts)} prompts for {L} latents")
    ids_c, emb_c = _embed(model, prompts)
    _, emb_u = _embed(model, [null_prompt] * L)
    controller = None
    if prompt_edit is not None:
        ids_e, emb_e = _embed(model, as_list(prompt_edit))
        tabs = [attention_edit_tables(ids_c.input_ids[i].tolist(), ids_e.input_ids[i].tolist()) for i in range(L)]
        controller = EdictP2PController([m for m, _ in tabs], [ix for _, ix in tabs])
        context = torch.cat([emb_u, emb_c, emb_e]).contiguous()
    else:
        context = torch.cat([emb_u, emb_c]).contiguous()
    groups = context.shape[0] // L  # 2 = [uncond, cond], 3 = [uncond, cond, cond_edit]
    model.unet.set_controller(controller)
    lib, h = _lib.load(), model.unet.handle

    timesteps = sched.timesteps[t_limit:]
    if reverse:
        timesteps = timesteps.flip(0)
    n = len(timesteps)
    for i, t in enumerate(timesteps):
        tt = int(t)
        if reverse:
            _lib.check(lib.pnp_edict_mix(h, C.c_void_p(pair[0].data_ptr()), C.c_void_p(pair[1].data_ptr()),
                                         pair[0].shape[0], float(mix_weight), 1, _lib.current_stream_ptr()))
        for k in range(2):
            if reverse:
                latent_i = (k + ((n - (i + 1)) + 1) % 2) % 2 if leapfrog_steps else (k + 1) % 2
            else:
                latent_i = (k + i % 2) % 2 if leapfrog_steps else k
            latent_j = (latent_i + 1) % 2
            x_in = torch.cat([pair[latent_j]] * groups).contiguous()
            eps = model.unet(x_in, tt, encoder_hidden_states=context)["sample"]
            eps_c = eps[(groups - 1) * L:]  # cond (or cond_edit when P2P is on) -- edict_functions.py:913-915
            co = step_coeffs(sched, tt, ratio, reverse)
            pair[latent_i] = fused_step(h, pair[latent_i], eps_c.contiguous(), co, eps_u=eps[0:L].contiguous(),
                                        guidance=guidance_scale)
        if not reverse:
            _lib.check(lib.pnp_edict_mix(h, C.c_void_p(pair[0].data_ptr()), C.c_void_p(pair[1].data_ptr()),
                                         pair[0].shape[0], float(mix_weight), 0, _lib.current_stream_ptr()))
    model.unet.set_controller(None)
    return pair


def EDICT_editing(model, latent, base_prompt, edit_prompt, use_p2p=False, steps=50, mix_weight=0.93,
                  init_image_strength=0.8, guidance_scale=3):
    """edict_functions.py:56-115."""
    latents = coupled_stablediffusion(model, base_prompt, reverse=True, init_image=latent,
                                      init_image_strength=init_image_strength, steps=steps, mix_weight=mix_weight,
                                      guidance_scale=guidance_scale)
    return coupled_stablediffusion(model, edit_prompt if not use_p2p else base_prompt,
                                   None if not use_p2p else edit_prompt, fixed_starting_latent=latents,
                                   init_image_strength=init_image_strength, steps=steps, mix_weight=mix_weight,
                                   guidance_scale=guidance_scale)


def edit_image_edict_p2p(model, image_path, prompt_src, prompt_tar, use_p2p, steps=50):
    """run_editing_edict.py:32-61 on latents: returns (reconstruction pair, edit pair)."""
    if not (isinstance(image_path, torch.Tensor) and image_path.dim() == 4):
        raise _lib.PnpError("pass the (L,4,64,64) image latents (prompts: one string, or a list of L)")
    latents = coupled_stablediffusion(model, prompt_src, reverse=True, init_image=image_path, steps=steps)
    recon = coupled_stablediffusion(model, prompt_src, reverse=False, fixed_starting_latent=latents, steps=steps)
    edit = EDICT_editing(model, image_path, prompt_src, prompt_tar, use_p2p=use_p2p, steps=steps)
    return recon, edit


def edit_image_edict_p2p_strip(model, image_path, prompt_src, prompt_tar, use_p2p, steps=50):
    """run_editing_edict.py:32-61 end to end: image file / HWC uint8 array -> the 2048x512 strip
    [instruction | source | reconstruction | edit].  The reference draws the initial latent from the VAE posterior
    (`latent_dist.sample()`, edict_functions.py:753) under the seed its caller set; so does this."""
    import numpy as np
    from PIL import Image

    from .ptp_utils import latent2image, load_512, txt_draw

    if model.vae is None:
        raise _lib.PnpError("edit_image_edict_p2p_strip needs a VAE on the model handle")
    image_gt = load_512(image_path)
    img = torch.from_numpy(image_gt.astype(np.float32) / 255.0 * 2.0 - 1.0).permute(2, 0, 1).unsqueeze(0).to(model.device)
    dist = model.vae.encode(img)["latent_dist"]
    z = (dist.mean + torch.exp(0.5 * dist.logvar) * torch.randn(dist.mean.shape, device=model.device)) * 0.18215
    recon, edit = edit_image_edict_p2p(model, z, prompt_src, prompt_tar, use_p2p, steps=steps)
    rec_img = latent2image(model.vae, recon[0], rounding=True)[0]
    edit_img = latent2image(model.vae, edit[0], rounding=True)[0]
    instruct = txt_draw(f"source prompt: {prompt_src}\ntarget prompt: {prompt_tar}")
    return Image.fromarray(np.concatenate((instruct, image_gt, rec_img, edit_img), axis=1))
