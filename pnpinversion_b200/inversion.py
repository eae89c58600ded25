"""DDIM inversion + the offset ("noise_loss") pass: parts 1 and 2 of PnP-Inversion's "3 lines".

Mirror of `models/p2p/inversion.py::DirectInversion` (:245-537): same class / method names and return values
(`invert -> (image_gt, image_rec, ddim_latents, noise_loss_list)`), with every per-step tensor expression replaced by
one launch of the fused epilogue kernel (csrc/epilogue.cu) around one fused-UNet call.
"""
from __future__ import annotations

import torch

from .attention_control import register_attention_control
from .ptp_utils import image2latent, latent2image
from .scheduler import fused_step, step_coefficients


class DirectInversion:
    def __init__(self, model, num_ddim_steps):
        self.model = model
        self.tokenizer = self.model.tokenizer
        self.prompt = None
        self.context = None
        self.num_ddim_steps = num_ddim_steps

    @property
    def scheduler(self):
        return self.model.scheduler

    @property
    def _engine(self):
        return self.model.unet.handle

    def _ratio(self):
        return self.scheduler.config.num_train_timesteps // self.scheduler.num_inference_steps

    # ---- single steps (inversion.py:247-270); integer arithmetic on the host, tensor math in the fused kernel
    def prev_step(self, model_output, timestep: int, sample):
        t = int(timestep)
        co = step_coefficients(self.scheduler.alphas_cumprod, self.scheduler.final_alpha_cumprod, t, t - self._ratio())
        prev_sample = fused_step(self._engine, sample.contiguous(), model_output.contiguous(), co)
        a_t, b_t, a_p, b_p = co
        difference_scale = a_p * (-b_t / a_t) + b_p
        return prev_sample, difference_scale

    def next_step(self, model_output, timestep: int, sample):
        t = int(timestep)
        cur_t, next_t = min(t - self._ratio(), 999), t
        co = step_coefficients(self.scheduler.alphas_cumprod, self.scheduler.final_alpha_cumprod, cur_t, next_t)
        return fused_step(self._engine, sample.contiguous(), model_output.contiguous(), co)

    def get_noise_pred_single(self, latents, t, context):
        return self.model.unet(latents, t, encoder_hidden_states=context)["sample"]

    @torch.no_grad()
    def init_prompt(self, prompt):
        tok, enc, dev = self.model.tokenizer, self.model.text_encoder, self.model.device
        uncond_input = tok([""] * len(prompt), padding="max_length", max_length=tok.model_max_length,
                           return_tensors="pt")
        uncond_embeddings = enc(uncond_input.input_ids.to(dev))[0]
        text_input = tok(prompt, padding="max_length", max_length=tok.model_max_length, truncation=True,
                         return_tensors="pt")
        text_embeddings = enc(text_input.input_ids.to(dev))[0]
        self.context = torch.cat([uncond_embeddings, text_embeddings]).to(dev, torch.float32)
        self.prompt = prompt

    @torch.no_grad()
    def ddim_loop(self, latent):
        """Hot loop #1 (inversion.py:308-319): 50 x { UNet(B=1, cond_src) ; inverse DDIM step }."""
        _, cond_embeddings = self.context.chunk(2)
        cond_embeddings = cond_embeddings[[0]].contiguous()
        all_latent = [latent]
        latent = latent.clone().detach()
        ts = self.scheduler.timesteps
        for i in range(self.num_ddim_steps):
            t = ts[len(ts) - i - 1]
            noise_pred = self.get_noise_pred_single(latent, t, cond_embeddings)
            latent = self.next_step(noise_pred, t, latent)
            all_latent.append(latent)
        return all_latent

    @torch.no_grad()
    def ddim_null_loop(self, latent):
        """inversion.py:322-333: the inversion conditioned on the EMPTY prompt."""
        uncond_embeddings, _ = self.context.chunk(2)
        uncond_embeddings = uncond_embeddings[[0]].contiguous()
        all_latent = [latent]
        latent = latent.clone().detach()
        ts = self.scheduler.timesteps
        for i in range(self.num_ddim_steps):
            t = ts[len(ts) - i - 1]
            noise_pred = self.get_noise_pred_single(latent, t, uncond_embeddings)
            latent = self.next_step(noise_pred, t, latent)
            all_latent.append(latent)
        return all_latent

    @torch.no_grad()
    def ddim_with_guidance_scale_loop(self, latent, guidance_scale):
        """inversion.py:335-350: inversion under classifier-free guidance.  The reference issues two B=1 UNet calls per
        step (empty prompt, source prompt); here they are the two rows of one B=2 call and the CFG combine is part of the
        fused inverse step."""
        uncond_embeddings, cond_embeddings = self.context.chunk(2)
        ctx = torch.cat([uncond_embeddings[[0]], cond_embeddings[[0]]]).contiguous()
        all_latent = [latent]
        latent = latent.clone().detach()
        ts = self.scheduler.timesteps
        ratio = self._ratio()
        for i in range(self.num_ddim_steps):
            t = int(ts[len(ts) - i - 1])
            noise_pred = self.get_noise_pred_single(torch.cat([latent] * 2), t, ctx)
            co = step_coefficients(self.scheduler.alphas_cumprod, self.scheduler.final_alpha_cumprod,
                                   min(t - ratio, 999), t)
            latent = fused_step(self._engine, latent.contiguous(), noise_pred[1:], co, eps_u=noise_pred[:1],
                                guidance=guidance_scale)
            all_latent.append(latent)
        return all_latent

    def _encode_image(self, image):
        latent = image2latent(self.model.vae, image).to(self.model.device, torch.float32)
        image_rec = latent2image(self.model.vae, latent)[0] if self.model.vae is not None else None
        return latent, image_rec

    @torch.no_grad()
    def ddim_inversion(self, image):
        latent, image_rec = self._encode_image(image)
        return image_rec, self.ddim_loop(latent)

    @torch.no_grad()
    def ddim_null_inversion(self, image):
        latent, image_rec = self._encode_image(image)
        return image_rec, self.ddim_null_loop(latent)

    @torch.no_grad()
    def ddim_with_guidance_scale_inversion(self, image, guidance_scale):
        latent, image_rec = self._encode_image(image)
        return image_rec, self.ddim_with_guidance_scale_loop(latent, guidance_scale)

    @torch.no_grad()
    def offset_calculate(self, latents, num_inner_steps, epsilon, guidance_scale, loss_scale_of_step=None):
        """Hot loop #2 (inversion.py:375-391): UNet(B=2*prompts) then ONE kernel doing CFG + prev_step +
        `loss = latent_prev - rec` + `latent_cur = rec + loss`.  `loss_scale_of_step(i)` serves the two ablations that
        scale (offset_calculate_not_full) or skip (offset_calculate_skip_step) the offset."""
        n = self.context.shape[0] // 2
        noise_loss_list = []
        latent_cur = torch.cat([latents[-1]] * n).contiguous()
        ratio = self._ratio()
        for i in range(self.num_ddim_steps):
            target = latents[len(latents) - i - 2].contiguous()  # (1,4,64,64), broadcast over the prompt rows
            t = int(self.scheduler.timesteps[i])
            noise_pred = self.get_noise_pred_single(torch.cat([latent_cur] * 2), t, self.context)
            eps_u, eps_c = noise_pred[:n], noise_pred[n:]
            co = step_coefficients(self.scheduler.alphas_cumprod, self.scheduler.final_alpha_cumprod, t, t - ratio)
            loss = torch.empty_like(latent_cur)
            latent_cur = fused_step(self._engine, latent_cur, eps_c, co, eps_u=eps_u, guidance=guidance_scale,
                                    target=target, loss_out=loss,
                                    loss_scale=1.0 if loss_scale_of_step is None else loss_scale_of_step(i))
            noise_loss_list.append(loss)
        return noise_loss_list

    def offset_calculate_not_full(self, latents, num_inner_steps, epsilon, guidance_scale, scale):
        """inversion.py:478-492: `loss = loss * scale`."""
        return self.offset_calculate(latents, num_inner_steps, epsilon, guidance_scale, lambda i: float(scale))

    def offset_calculate_skip_step(self, latents, num_inner_steps, epsilon, guidance_scale, skip_step):
        """inversion.py:501-519: the offset is kept on every `skip_step`-th step and zero elsewhere."""
        return self.offset_calculate(latents, num_inner_steps, epsilon, guidance_scale,
                                     lambda i: 1.0 if (i % skip_step) == 0 else 0.0)

    def invert(self, image_gt, prompt, guidance_scale, num_inner_steps=10, early_stop_epsilon=1e-5):
        self.init_prompt(prompt)
        register_attention_control(self.model, None)
        image_rec, ddim_latents = self.ddim_inversion(image_gt)
        noise_loss_list = self.offset_calculate(ddim_latents, num_inner_steps, early_stop_epsilon, guidance_scale)
        return image_gt, image_rec, ddim_latents, noise_loss_list

    def invert_without_attn_controller(self, image_gt, prompt, guidance_scale, num_inner_steps=10,
                                       early_stop_epsilon=1e-5):
        """inversion.py:403-410 (used by the MasaCtrl editor, which registers its own attention editor)."""
        self.init_prompt(prompt)
        image_rec, ddim_latents = self.ddim_inversion(image_gt)
        noise_loss_list = self.offset_calculate(ddim_latents, num_inner_steps, early_stop_epsilon, guidance_scale)
        return image_gt, image_rec, ddim_latents, noise_loss_list

    def invert_with_guidance_scale_vary_guidance(self, image_gt, prompt, inverse_guidance_scale, forward_guidance_scale,
                                                 num_inner_steps=10, early_stop_epsilon=1e-5):
        """inversion.py:412-419."""
        self.init_prompt(prompt)
        register_attention_control(self.model, None)
        image_rec, ddim_latents = self.ddim_with_guidance_scale_inversion(image_gt, inverse_guidance_scale)
        noise_loss_list = self.offset_calculate(ddim_latents, num_inner_steps, early_stop_epsilon,
                                                forward_guidance_scale)
        return image_gt, image_rec, ddim_latents, noise_loss_list

    def invert_not_full(self, image_gt, prompt, guidance_scale, num_inner_steps=10, early_stop_epsilon=1e-5, scale=1.):
        """inversion.py:494-499."""
        self.init_prompt(prompt)
        register_attention_control(self.model, None)
        image_rec, ddim_latents = self.ddim_inversion(image_gt)
        noise_loss_list = self.offset_calculate_not_full(ddim_latents, num_inner_steps, early_stop_epsilon,
                                                         guidance_scale, scale)
        return image_gt, image_rec, ddim_latents, noise_loss_list

    def invert_skip_step(self, image_gt, prompt, guidance_scale, skip_step, num_inner_steps=10,
                         early_stop_epsilon=1e-5, scale=1.):
        """inversion.py:521-526."""
        self.init_prompt(prompt)
        register_attention_control(self.model, None)
        image_rec, ddim_latents = self.ddim_inversion(image_gt)
        noise_loss_list = self.offset_calculate_skip_step(ddim_latents, num_inner_steps, early_stop_epsilon,
                                                          guidance_scale, skip_step)
        return image_gt, image_rec, ddim_latents, noise_loss_list

    def invert_null_latent(self, *args, **kwargs):
        """inversion.py:421-476 optimises the unconditional embedding with Adam through the UNet (needs a UNet
        backward pass); out of scope of the forward-only engine."""
        raise NotImplementedError("invert_null_latent needs gradients through the UNet (no backward pass in this engine)")


def slerp(val, low, high):
    """utils/utils.py:7-16: spherical interpolation per row of (rows, features) tensors."""
    low_norm = low / torch.norm(low, dim=1, keepdim=True)
    high_norm = high / torch.norm(high, dim=1, keepdim=True)
    omega = torch.acos((low_norm * high_norm).sum(1))
    so = torch.sin(omega)
    return (torch.sin((1.0 - val) * omega) / so).unsqueeze(1) * low + (torch.sin(val * omega) / so).unsqueeze(1) * high


def slerp_tensor(val, low, high):
    """utils/utils.py:19-25 (used by negative-prompt inversion)."""
    shape = low.shape
    return slerp(val, low.flatten(1), high.flatten(1)).reshape(shape)


class NegativePromptInversion(DirectInversion):
    """`models/p2p/inversion.py::NegativePromptInversion` (:10-108): DDIM inversion with the source prompt, then the
    source-prompt embedding (optionally slerp-ed towards the empty-prompt embedding) stands in for the unconditional
    embedding of every step.  Forward-only, so it is served by the same fused kernels."""

    @torch.no_grad()
    def init_prompt(self, prompt):
        super().init_prompt([prompt])

    @torch.no_grad()
    def ddim_inversion(self, image):
        latent, image_rec = self._encode_image(image)
        return image_rec, self.ddim_loop(latent), latent

    def invert(self, image_gt, prompt, npi_interp=0.0):
        self.init_prompt(prompt)
        register_attention_control(self.model, None)
        image_rec, ddim_latents, image_rec_latent = self.ddim_inversion(image_gt)
        uncond_embeddings, cond_embeddings = self.context.chunk(2)
        if npi_interp > 0.0:
            cond_embeddings = slerp_tensor(npi_interp, cond_embeddings, uncond_embeddings)
        uncond_embeddings = [cond_embeddings] * self.num_ddim_steps
        return image_rec, image_rec_latent, ddim_latents, uncond_embeddings


class NullInversion(DirectInversion):
    """`models/p2p/inversion.py::NullInversion` (:111-241) for `num_inner_steps == 0`, the only setting the `ddim+p2p`
    method uses (p2p_editor.py:153-154): with no inner optimisation steps `null_optimization` (:196-225) returns the
    unmodified empty-prompt embedding for every step and its CFG `get_noise_pred` calls only advance a latent nobody
    reads, so they are not executed.  The optimisation proper (num_inner_steps > 0, Adam on the embedding through the
    UNet) needs a backward pass and is out of scope of this forward-only engine."""

    @torch.no_grad()
    def init_prompt(self, prompt):
        super().init_prompt([prompt])

    def null_optimization(self, latents, num_inner_steps, epsilon, guidance_scale):
        if num_inner_steps != 0:
            raise NotImplementedError("null-text optimisation needs gradients through the UNet (no backward pass here)")
        uncond_embeddings, _ = self.context.chunk(2)
        return [uncond_embeddings[:1]] * self.num_ddim_steps

    def invert(self, image_gt, prompt, guidance_scale, num_inner_steps=10, early_stop_epsilon=1e-5):
        if num_inner_steps != 0:
            raise NotImplementedError("null-text optimisation needs gradients through the UNet (no backward pass here)")
        self.init_prompt(prompt)
        register_attention_control(self.model, None)
        image_rec, ddim_latents = self.ddim_inversion(image_gt)
        uncond_embeddings = self.null_optimization(ddim_latents, num_inner_steps, early_stop_epsilon, guidance_scale)
        return image_gt, image_rec, ddim_latents, uncond_embeddings
