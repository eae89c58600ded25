"""The guided denoising loops with the source-branch latent rectification -- part 3 of the "3 lines".

Mirror of `models/p2p/p2p_guidance_forward.py` (:103-116 step, :135-173 loop).  Per step: one fused-UNet call
(B = 2 * prompts, controller compiled into the attention kernels) and one fused epilogue launch that performs
    noise_pred = uncond + g * (text - uncond)                       (:111)
    latents    = scheduler.step(noise_pred, t, latents).prev_sample (:112)
    latents    = cat(latents[:1] + noise_loss[:1], latents[1:])     (:113-114, the rectification)
followed by `controller.step_callback` (LocalBlend, one more launch when active).
"""
from __future__ import annotations

import torch

from .attention_control import register_attention_control
from .ptp_utils import init_latent
from .scheduler import fused_step, step_coefficients


def _encode(model, prompt):
    tok, enc, dev = model.tokenizer, model.text_encoder, model.device
    text_input = tok(prompt, padding="max_length", max_length=tok.model_max_length, truncation=True,
                     return_tensors="pt")
    text_embeddings = enc(text_input.input_ids.to(dev))[0]
    max_length = text_input.input_ids.shape[-1]
    uncond_input = tok([""] * len(prompt), padding="max_length", max_length=max_length, return_tensors="pt")
    uncond_embeddings = enc(uncond_input.input_ids.to(dev))[0]
    return torch.cat([uncond_embeddings, text_embeddings]).to(dev, torch.float32).contiguous()


def direct_inversion_p2p_guidance_diffusion_step(model, controller, latents, context, t, guidance_scale, noise_loss,
                                                 low_resource=False, add_offset=True, add_target=False):
    if low_resource:
        raise NotImplementedError("low_resource (two B=n UNet calls) is not on the hot path")
    n = latents.shape[0]
    noise_pred = model.unet(torch.cat([latents] * 2), t, encoder_hidden_states=context)["sample"]
    sched = model.scheduler
    tt = int(t)
    co = step_coefficients(sched.alphas_cumprod, sched.final_alpha_cumprod, tt,
                           tt - sched.config.num_train_timesteps // sched.num_inference_steps)
    mask = 0
    if add_offset:
        mask = (1 << n) - 1 if add_target else 1
    latents = fused_step(model.unet.handle, latents.contiguous(), noise_pred[n:], co, eps_u=noise_pred[:n],
                         guidance=guidance_scale, noise_loss=noise_loss.contiguous() if add_offset else None,
                         add_mask=mask)
    if controller is not None:
        latents = controller.step_callback(latents)
    return latents


@torch.no_grad()
def direct_inversion_p2p_guidance_forward(model, prompt, controller, latent=None, num_inference_steps: int = 50,
                                          guidance_scale=7.5, generator=None, noise_loss_list=None, add_offset=True):
    batch_size = len(prompt)
    register_attention_control(model, controller)
    height = width = 512
    context = _encode(model, prompt)
    latent, latents = init_latent(latent, model, height, width, generator, batch_size)
    latents = latents.to(torch.float32).contiguous()
    model.scheduler.set_timesteps(num_inference_steps)
    for i, t in enumerate(model.scheduler.timesteps):
        latents = direct_inversion_p2p_guidance_diffusion_step(model, controller, latents, context, t, guidance_scale,
                                                               noise_loss_list[i], low_resource=False,
                                                               add_offset=add_offset)
    return latents, latent


@torch.no_grad()
def direct_inversion_p2p_guidance_forward_add_target(model, prompt, controller, latent=None,
                                                     num_inference_steps: int = 50, guidance_scale=7.5, generator=None,
                                                     noise_loss_list=None, add_offset=True):
    """p2p_guidance_forward.py:119-132,175-213: the ablation that also rectifies the target branch."""
    batch_size = len(prompt)
    register_attention_control(model, controller)
    context = _encode(model, prompt)
    latent, latents = init_latent(latent, model, 512, 512, generator, batch_size)
    latents = latents.to(torch.float32).contiguous()
    model.scheduler.set_timesteps(num_inference_steps)
    for i, t in enumerate(model.scheduler.timesteps):
        latents = direct_inversion_p2p_guidance_diffusion_step(model, controller, latents, context, t, guidance_scale,
                                                               noise_loss_list[i], add_offset=add_offset,
                                                               add_target=True)
    return latents, latent


def p2p_guidance_diffusion_step(model, controller, latents, context, t, guidance_scale, low_resource=False):
    """p2p_guidance_forward.py:6-18 (no rectification: the plain DDIM+P2P baseline)."""
    return direct_inversion_p2p_guidance_diffusion_step(model, controller, latents, context, t, guidance_scale, None,
                                                        low_resource=low_resource, add_offset=False)


@torch.no_grad()
def p2p_guidance_forward(model, prompt, controller, num_inference_steps: int = 50, guidance_scale=7.5, generator=None,
                         latent=None, uncond_embeddings=None):
    """p2p_guidance_forward.py:21-62."""
    batch_size = len(prompt)
    register_attention_control(model, controller)
    context_default = _encode(model, prompt)
    text_embeddings = context_default[batch_size:]
    latent, latents = init_latent(latent, model, 512, 512, generator, batch_size)
    latents = latents.to(torch.float32).contiguous()
    model.scheduler.set_timesteps(num_inference_steps)
    for i, t in enumerate(model.scheduler.timesteps):
        if uncond_embeddings is not None:
            context = torch.cat([uncond_embeddings[i].to(text_embeddings).expand(*text_embeddings.shape),
                                 text_embeddings]).contiguous()
        else:
            context = context_default
        latents = p2p_guidance_diffusion_step(model, controller, latents, context, t, guidance_scale)
    return latents, latent


@torch.no_grad()
def p2p_guidance_forward_single_branch(model, prompt, controller, num_inference_steps: int = 50, guidance_scale=7.5,
                                       generator=None, latent=None, uncond_embeddings=None):
    """p2p_guidance_forward.py:65-100: the optimised unconditional embedding replaces only the FIRST unconditional row."""
    batch_size = len(prompt)
    register_attention_control(model, controller)
    context_default = _encode(model, prompt)
    uncond_default, text_embeddings = context_default[:batch_size], context_default[batch_size:]
    latent, latents = init_latent(latent, model, 512, 512, generator, batch_size)
    latents = latents.to(torch.float32).contiguous()
    model.scheduler.set_timesteps(num_inference_steps)
    for i, t in enumerate(model.scheduler.timesteps):
        context = torch.cat([torch.cat([uncond_embeddings[i].to(text_embeddings), uncond_default[1:]]),
                             text_embeddings]).contiguous()
        latents = p2p_guidance_diffusion_step(model, controller, latents, context, t, guidance_scale)
    return latents, latent


def _dilate(image, kernel_size, stride=1, padding=0):
    """proximal_guidance_forward.py:7-17."""
    assert image.max() <= 1 and image.min() >= 0
    return torch.nn.functional.max_pool2d(image, kernel_size, stride, padding)


def proximal_guidance_diffusion_step(model, controller, latents, context, t, guidance_scale, low_resource=False,
                                     edit_stage=True, prox=None, quantile=0.7, image_enc=None, recon_lr=0.1,
                                     recon_t=400, inversion_guidance=False, x_stars=None, i=0, dilate_mask=0):
    """proximal_guidance_forward.py:20-93.  Without `prox` this is exactly the plain guided step (one UNet call + one
    fused epilogue).  With prox = 'l0' / 'l1' the score difference is thresholded at its `quantile` (a sort: torch on the
    device, off the headline path) and the DDIM step is fed the already-combined prediction; the optional reconstruction
    term on the predicted x0 (scheduler_dev.py:61-70) is applied with the same scalar coefficients."""
    if low_resource:
        raise NotImplementedError("low_resource (two B=n UNet calls) is not on the hot path")
    if not (edit_stage and prox is not None):
        latents = direct_inversion_p2p_guidance_diffusion_step(model, controller, latents, context, t, guidance_scale,
                                                               None, add_offset=False)
        return latents
    n = latents.shape[0]
    noise_pred = model.unet(torch.cat([latents] * 2), t, encoder_hidden_states=context)["sample"]
    noise_pred_uncond, noise_prediction_text = noise_pred[:n], noise_pred[n:]
    tt = int(t)
    ref_image, lr, recon_mask, mask_edit = None, 0.0, None, None
    score_delta = noise_prediction_text - noise_pred_uncond
    threshold = score_delta.abs().quantile(quantile) if quantile > 0 else -quantile
    score_delta = score_delta - score_delta.clamp(-threshold, threshold)
    if prox == "l1":
        score_delta = torch.where(score_delta > 0, score_delta - threshold, score_delta)
        score_delta = torch.where(score_delta < 0, score_delta + threshold, score_delta)
    elif prox != "l0":
        raise NotImplementedError
    if (recon_t > 0 and tt < recon_t) or (recon_t < 0 and tt > -recon_t):
        ref_image, lr = image_enc, recon_lr
        mask_edit = (score_delta.abs() > threshold).float()
        if dilate_mask > 0:
            radius = int(dilate_mask)
            mask_edit = _dilate(mask_edit.float(), kernel_size=2 * radius + 1, padding=radius)
        recon_mask = 1 - mask_edit
    noise_pred = (noise_pred_uncond + guidance_scale * score_delta).contiguous()
    sched = model.scheduler
    co = step_coefficients(sched.alphas_cumprod, sched.final_alpha_cumprod, tt,
                           tt - sched.config.num_train_timesteps // sched.num_inference_steps)
    if ref_image is not None and lr > 0.0:
        a_t, b_t, a_p, b_p = co
        x0 = (latents - b_t * noise_pred) / a_t
        x0 = x0 - lr * (x0 - ref_image.expand_as(x0)) * recon_mask.expand_as(x0).float()
        latents = a_p * x0 + b_p * noise_pred
    else:
        latents = fused_step(model.unet.handle, latents.contiguous(), noise_pred, co)
    if (mask_edit is not None and inversion_guidance and (recon_t > 0 and tt < recon_t)) or (recon_t < 0 and tt > -recon_t):
        recon_mask = 1 - mask_edit
        latents = latents - recon_lr * (latents - x_stars[len(x_stars) - i - 2].expand_as(latents)) * recon_mask
    return controller.step_callback(latents.contiguous())


@torch.no_grad()
def proximal_guidance_forward(model, prompt, controller, guidance_scale=7.5, generator=None, latent=None,
                              uncond_embeddings=None, edit_stage=True, prox=None, quantile=0.7, image_enc=None,
                              recon_lr=0.1, recon_t=400, inversion_guidance=False, x_stars=None, dilate_mask=None):
    """proximal_guidance_forward.py:96-170 (note: like the reference it does not call scheduler.set_timesteps)."""
    batch_size = len(prompt)
    register_attention_control(model, controller)
    context_default = _encode(model, prompt)
    text_embeddings = context_default[batch_size:]
    latent, latents = init_latent(latent, model, 512, 512, generator, batch_size)
    latents = latents.to(torch.float32).contiguous()
    for i, t in enumerate(model.scheduler.timesteps):
        if uncond_embeddings is not None:
            context = torch.cat([uncond_embeddings[i].to(text_embeddings).expand(*text_embeddings.shape),
                                 text_embeddings]).contiguous()
        else:
            context = context_default
        latents = proximal_guidance_diffusion_step(model, controller, latents, context, t, guidance_scale,
                                                   edit_stage=edit_stage, prox=prox, quantile=quantile,
                                                   image_enc=image_enc, recon_lr=recon_lr, recon_t=recon_t,
                                                   inversion_guidance=inversion_guidance, x_stars=x_stars, i=i,
                                                   dilate_mask=dilate_mask or 0)
    return latents, latent
