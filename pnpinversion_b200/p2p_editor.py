"""Seam A: the editor facade.  Same constructor / call signature as `models/p2p_editor.py::P2PEditor`
(:13-25, :28-135); returns the reference's 4-panel 2048x512 PIL strip when a VAE is attached to the model handle,
otherwise an `EditResult` carrying the latents (the VAE is a "next" row).

In scope: "directinversion+p2p" (the north-star path, edit_image_directinversion :415-479), "ddim+p2p" (:137-197) and
the add-target ablation "ablation_directinversion_add_target+p2p".  Any other method name raises NotImplementedError
exactly like the reference does for unknown names (:134-135).
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import List, Optional

import numpy as np
import torch

from .attention_control import AttentionStore, make_controller
from .inversion import DirectInversion
from .p2p_guidance_forward import (direct_inversion_p2p_guidance_forward,
                                   direct_inversion_p2p_guidance_forward_add_target, p2p_guidance_forward)
from .ptp_utils import latent2image, load_512


@dataclass
class EditResult:
    """Returned instead of a PIL image when the model handle has no VAE."""
    x_stars: List[torch.Tensor]
    noise_loss_list: Optional[List[torch.Tensor]]
    reconstruct_latent: torch.Tensor  # (2,4,64,64)
    latents: torch.Tensor             # (2,4,64,64): [source branch, edited]


class P2PEditor:
    def __init__(self, method_list, device, num_ddim_steps=50, model=None) -> None:
        self.device = device
        self.method_list = method_list
        self.num_ddim_steps = num_ddim_steps
        if model is None:
            raise RuntimeError(
                "no SD-1.x checkpoint is available offline: pass model=FusedModel.from_state_dict_file(...) or "
                "FusedModel.synthetic(...) (the reference loads CompVis/stable-diffusion-v1-4 here, p2p_editor.py:23)")
        self.ldm_stable = model
        self.ldm_stable.scheduler.set_timesteps(self.num_ddim_steps)

    def __call__(self, edit_method, image_path, prompt_src, prompt_tar, guidance_scale=7.5, proximal=None,
                 quantile=0.7, use_reconstruction_guidance=False, recon_t=400, recon_lr=0.1, cross_replace_steps=0.4,
                 self_replace_steps=0.6, blend_word=None, eq_params=None, is_replace_controller=False,
                 use_inversion_guidance=False, dilate_mask=1):
        kw = dict(guidance_scale=guidance_scale, cross_replace_steps=cross_replace_steps,
                  self_replace_steps=self_replace_steps, blend_word=blend_word, eq_params=eq_params,
                  is_replace_controller=is_replace_controller)
        if edit_method == "ddim+p2p":
            return self.edit_image_ddim(image_path, prompt_src, prompt_tar, **kw)
        if edit_method == "directinversion+p2p":
            return self.edit_image_directinversion(image_path, prompt_src, prompt_tar, **kw)
        if edit_method == "ablation_directinversion_add_target+p2p":
            return self.edit_image_directinversion(image_path, prompt_src, prompt_tar, add_target=True, **kw)
        raise NotImplementedError(f"No edit method named {edit_method}")

    # ---------------------------------------------------------------------------------------------
    def _load(self, image_path):
        if isinstance(image_path, torch.Tensor) and image_path.dim() == 4:
            return image_path  # synthetic-latent path (image2latent passes 4-D tensors through, utils/utils.py:73-74)
        return load_512(image_path)

    def _panel(self, image_gt, prompt_src, prompt_tar, reconstruct_latent, latents, x_stars, noise_loss_list):
        vae = self.ldm_stable.vae
        if vae is None or isinstance(image_gt, torch.Tensor):
            return EditResult(x_stars, noise_loss_list, reconstruct_latent, latents)
        from PIL import Image

        rec = latent2image(vae, reconstruct_latent)[0]
        out = latent2image(vae, latents)
        instruct = np.full((512, 512, 3), 255, dtype=np.uint8)  # txt_draw needs matplotlib (absent): blank panel
        return Image.fromarray(np.concatenate((instruct, image_gt, rec, out[-1]), axis=1))

    def edit_image_directinversion(self, image_path, prompt_src, prompt_tar, guidance_scale=7.5,
                                   cross_replace_steps=0.4, self_replace_steps=0.6, blend_word=None, eq_params=None,
                                   is_replace_controller=False, add_target=False):
        image_gt = self._load(image_path)
        prompts = [prompt_src, prompt_tar]
        inversion = DirectInversion(model=self.ldm_stable, num_ddim_steps=self.num_ddim_steps)
        _, _, x_stars, noise_loss_list = inversion.invert(image_gt=image_gt, prompt=prompts,
                                                          guidance_scale=guidance_scale)
        x_t = x_stars[-1]
        fwd = direct_inversion_p2p_guidance_forward_add_target if add_target else direct_inversion_p2p_guidance_forward
        controller = AttentionStore()
        reconstruct_latent, x_t = fwd(model=self.ldm_stable, prompt=prompts, controller=controller,
                                      noise_loss_list=noise_loss_list, latent=x_t,
                                      num_inference_steps=self.num_ddim_steps, guidance_scale=guidance_scale,
                                      generator=None)
        controller = make_controller(pipeline=self.ldm_stable, prompts=prompts,
                                     is_replace_controller=is_replace_controller,
                                     cross_replace_steps={"default_": cross_replace_steps},
                                     self_replace_steps=self_replace_steps, blend_words=blend_word,
                                     equilizer_params=eq_params, num_ddim_steps=self.num_ddim_steps, device=self.device)
        latents, _ = fwd(model=self.ldm_stable, prompt=prompts, controller=controller, noise_loss_list=noise_loss_list,
                         latent=x_t, num_inference_steps=self.num_ddim_steps, guidance_scale=guidance_scale,
                         generator=None)
        return self._panel(image_gt, prompt_src, prompt_tar, reconstruct_latent, latents, x_stars, noise_loss_list)

    def edit_image_ddim(self, image_path, prompt_src, prompt_tar, guidance_scale=7.5, cross_replace_steps=0.4,
                        self_replace_steps=0.6, blend_word=None, eq_params=None, is_replace_controller=False):
        """DDIM inversion + P2P without rectification.  The reference routes through NullInversion.invert(...,
        num_inner_steps=0), whose 50-step CFG loop only returns the unmodified unconditional embedding 50 times
        (inversion.py:196-225); that result-neutral loop is not executed here."""
        image_gt = self._load(image_path)
        prompts = [prompt_src, prompt_tar]
        inversion = DirectInversion(model=self.ldm_stable, num_ddim_steps=self.num_ddim_steps)
        inversion.init_prompt([prompt_src])
        from .attention_control import register_attention_control

        register_attention_control(self.ldm_stable, None)
        _, x_stars = inversion.ddim_inversion(image_gt)
        x_t = x_stars[-1]
        controller = AttentionStore()
        reconstruct_latent, x_t = p2p_guidance_forward(model=self.ldm_stable, prompt=[prompt_src],
                                                       controller=controller, latent=x_t,
                                                       num_inference_steps=self.num_ddim_steps,
                                                       guidance_scale=guidance_scale, generator=None)
        controller = make_controller(pipeline=self.ldm_stable, prompts=prompts,
                                     is_replace_controller=is_replace_controller,
                                     cross_replace_steps={"default_": cross_replace_steps},
                                     self_replace_steps=self_replace_steps, blend_words=blend_word,
                                     equilizer_params=eq_params, num_ddim_steps=self.num_ddim_steps, device=self.device)
        latents, _ = p2p_guidance_forward(model=self.ldm_stable, prompt=prompts, controller=controller, latent=x_t,
                                          num_inference_steps=self.num_ddim_steps, guidance_scale=guidance_scale,
                                          generator=None)
        return self._panel(image_gt, prompt_src, prompt_tar, reconstruct_latent, latents, x_stars, None)
