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
    def ddim_inversion(self, image):
        latent = image2latent(self.model.vae, image).to(self.model.device, torch.float32)
        image_rec = latent2image(self.model.vae, latent)[0] if self.model.vae is not None else None
        return image_rec, self.ddim_loop(latent)

    @torch.no_grad()
    def offset_calculate(self, latents, num_inner_steps, epsilon, guidance_scale):
        """Hot loop #2 (inversion.py:375-391): UNet(B=2*prompts) then ONE kernel doing CFG + prev_step +
        `loss = latent_prev - rec` + `latent_cur = rec + loss`."""
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
                                    target=target, loss_out=loss)
            noise_loss_list.append(loss)
        return noise_loss_list

    def invert(self, image_gt, prompt, guidance_scale, num_inner_steps=10, early_stop_epsilon=1e-5):
        self.init_prompt(prompt)
        register_attention_control(self.model, None)
        image_rec, ddim_latents = self.ddim_inversion(image_gt)
        noise_loss_list = self.offset_calculate(ddim_latents, num_inner_steps, early_stop_epsilon, guidance_scale)
        return image_gt, image_rec, ddim_latents, noise_loss_list
