"""Seam A: the editor facade.  Same constructor / call signature as `models/p2p_editor.py::P2PEditor`
(:13-25, :28-135); returns the reference's 4-panel 2048x512 PIL strip when a VAE is attached to the model handle,
otherwise an `EditResult` carrying the latents (the VAE is a "next" row).

Every edit_method name the reference dispatches on (:46-133) is accepted: "directinversion+p2p" (the north-star path,
edit_image_directinversion :415-479), "ddim+p2p" (:137-197), negative-prompt inversion with and without proximal
guidance (:324-413), the 20 inverse/forward guidance-scale pairs (:69-88), the scaled / skipped / add-target / add-source
offset ablations (:107-133).  The null-text / null-latent methods optimise an embedding THROUGH the UNet (Adam, backward
pass): they raise NotImplementedError naming the reason.  Unknown names raise NotImplementedError like the reference
(:134-135).
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


_GUIDANCE_TABLE = {"0": 0, "1": 1, "25": 2.5, "5": 5, "75": 7.5}  # p2p_editor.py:77-82
_GUIDANCE_METHODS = tuple(f"directinversion+p2p_guidance_{a}_{b}" for a in ("0", "1", "25", "5", "75")
                          for b in ("1", "5", "25", "75"))  # p2p_editor.py:69-75
_NEEDS_BACKWARD = ("null-text-inversion+p2p", "null-text-inversion+p2p_a800", "null-text-inversion+p2p_3090",
                   "ablation_null-text-inversion_single_branch+p2p", "null-text-inversion+proximal-guidance",
                   "ablation_null-latent-inversion+p2p")

# every edit_method name models/p2p_editor.py:46-133 dispatches on
SUPPORTED_METHODS = (("ddim+p2p", "directinversion+p2p", "negative-prompt-inversion+p2p",
                      "negative-prompt-inversion+proximal-guidance", "ablation_directinversion_08+p2p",
                      "ablation_directinversion_04+p2p", "ablation_directinversion_add-target+p2p",
                      "ablation_directinversion_add-source+p2p")
                     + tuple(f"ablation_directinversion_interval_{k}+p2p" for k in (2, 5, 10, 24, 49)) + _GUIDANCE_METHODS)


class P2PEditor:
    def __init__(self, method_list, device, num_ddim_steps=50, model=None, fused_loops=None) -> None:
        self.device = device
        self.method_list = method_list
        self.num_ddim_steps = num_ddim_steps
        if model is None:
            # the reference downloads CompVis/stable-diffusion-v1-4 here (p2p_editor.py:23); offline the checkpoint
            # directory comes from PNP_SD_CHECKPOINT, and without one the seeded random-init stand-in is used
            import os

            from .model import FusedModel
            ckpt = os.environ.get("PNP_SD_CHECKPOINT")
            model = (FusedModel.from_pretrained(ckpt, device=str(device)) if ckpt
                     else FusedModel.synthetic(device=str(device), with_vae=True))
        self.ldm_stable = model
        self.ldm_stable.scheduler.set_timesteps(self.num_ddim_steps)
        # the C-side step loops need the fused engine; a test double of the engine keeps the Python loops
        self.fused_loops = hasattr(getattr(model, "unet", None), "handle") if fused_loops is None else bool(fused_loops)

    def __call__(self, edit_method, image_path, prompt_src, prompt_tar, guidance_scale=7.5, proximal=None,
                 quantile=0.7, use_reconstruction_guidance=False, recon_t=400, recon_lr=0.1, cross_replace_steps=0.4,
                 self_replace_steps=0.6, blend_word=None, eq_params=None, is_replace_controller=False,
                 use_inversion_guidance=False, dilate_mask=1):
        kw = dict(cross_replace_steps=cross_replace_steps, self_replace_steps=self_replace_steps,
                  blend_word=blend_word, eq_params=eq_params, is_replace_controller=is_replace_controller)
        prox_kw = dict(quantile=quantile, use_reconstruction_guidance=use_reconstruction_guidance, recon_t=recon_t,
                       recon_lr=recon_lr, use_inversion_guidance=use_inversion_guidance, dilate_mask=dilate_mask)
        if edit_method == "ddim+p2p":
            return self.edit_image_ddim(image_path, prompt_src, prompt_tar, guidance_scale=guidance_scale, **kw)
        if edit_method in _NEEDS_BACKWARD:
            raise NotImplementedError(
                f"{edit_method}: null-text / null-latent optimisation runs Adam through the UNet (models/p2p/inversion.py:"
                "196-225,421-476); this engine is forward-only")
        if edit_method == "negative-prompt-inversion+p2p":
            return self.edit_image_negative_prompt_inversion(image_path, prompt_src, prompt_tar,
                                                             guidance_scale=guidance_scale, proximal=None, **prox_kw, **kw)
        if edit_method == "negative-prompt-inversion+proximal-guidance":
            return self.edit_image_negative_prompt_inversion(image_path, prompt_src, prompt_tar,
                                                             guidance_scale=guidance_scale, proximal=proximal, **prox_kw,
                                                             **kw)
        if edit_method == "directinversion+p2p":
            return self.edit_image_directinversion(image_path, prompt_src, prompt_tar, guidance_scale=guidance_scale, **kw)
        if edit_method in _GUIDANCE_METHODS:
            inv_g = _GUIDANCE_TABLE[edit_method.split("_")[-2]]
            fwd_g = _GUIDANCE_TABLE[edit_method.split("_")[-1]]
            return self.edit_image_directinversion_vary_guidance_scale(image_path, prompt_src, prompt_tar,
                                                                       inverse_guidance_scale=inv_g,
                                                                       forward_guidance_scale=fwd_g, **kw)
        if edit_method in ("ablation_directinversion_08+p2p", "ablation_directinversion_04+p2p"):
            scale = float(edit_method.split("+")[0].split("_")[-1]) / 10
            return self.edit_image_directinversion_not_full(image_path, prompt_src, prompt_tar,
                                                            guidance_scale=guidance_scale, scale=scale, **kw)
        if edit_method in tuple(f"ablation_directinversion_interval_{k}+p2p" for k in (2, 5, 10, 24, 49)):
            skip_step = int(edit_method.split("+")[0].split("_")[-1])
            return self.edit_image_directinversion_skip_step(image_path, prompt_src, prompt_tar, skip_step=skip_step,
                                                             guidance_scale=guidance_scale, **kw)
        if edit_method == "ablation_directinversion_add-target+p2p":
            return self.edit_image_directinversion_add_target(image_path, prompt_src, prompt_tar,
                                                              guidance_scale=guidance_scale, **kw)
        if edit_method == "ablation_directinversion_add-source+p2p":
            return self.edit_image_directinversion_add_source(image_path, prompt_src, prompt_tar,
                                                              guidance_scale=guidance_scale, **kw)
        raise NotImplementedError(f"No edit method named {edit_method}")

    # ---------------------------------------------------------------------------------------------
    def _load(self, image_path):
        if isinstance(image_path, torch.Tensor) and image_path.dim() == 4:
            return image_path  # synthetic-latent path (image2latent passes 4-D tensors through, utils/utils.py:73-74)
        return load_512(image_path)

    def _panel(self, image_gt, prompt_src, prompt_tar, reconstruct_latent, latents, x_stars, noise_loss_list):
        """p2p_editor.py:474-479: [instruction text | source | reconstruction | edit], 2048 x 512."""
        vae = self.ldm_stable.vae
        if vae is None or isinstance(image_gt, torch.Tensor):
            return EditResult(x_stars, noise_loss_list, reconstruct_latent, latents)
        from PIL import Image

        from .ptp_utils import txt_draw

        rec = latent2image(vae, reconstruct_latent)[0]
        out = latent2image(vae, latents)
        instruct = txt_draw(f"source prompt: {prompt_src}\ntarget prompt: {prompt_tar}")
        return Image.fromarray(np.concatenate((instruct, image_gt, rec, out[-1]), axis=1))

    def _controller(self, prompts, cross_replace_steps, self_replace_steps, blend_word, eq_params,
                    is_replace_controller):
        return make_controller(pipeline=self.ldm_stable, prompts=prompts, is_replace_controller=is_replace_controller,
                               cross_replace_steps={"default_": cross_replace_steps},
                               self_replace_steps=self_replace_steps, blend_words=blend_word,
                               equilizer_params=eq_params, num_ddim_steps=self.num_ddim_steps, device=self.device)

    def _direct(self, image_path, prompt_src, prompt_tar, forward_guidance, cross_replace_steps, self_replace_steps,
                blend_word, eq_params, is_replace_controller, add_target=False, add_source=False,
                inverse_guidance=None, scale=None, skip_step=None):
        """Shared body of the DirectInversion methods (p2p_editor.py:415-479 and its ablation copies :481-548,
        :707-976).  With `fused_loops` (default) the four 50-step loops run inside libpnpinv.so (`pnp_run_loop`, through
        batched.BatchedDirectInversionP2P with one image); otherwise through the Python loops that mirror the
        reference's (inversion.py, p2p_guidance_forward.py) - bit-identical results (tests/test_gpu_batched.py)."""
        image_gt = self._load(image_path)
        prompts = [prompt_src, prompt_tar]
        n = self.num_ddim_steps
        loss_scales = None
        if scale is not None:
            loss_scales = [float(scale)] * n
        if skip_step is not None:
            loss_scales = [1.0 if (i % skip_step) == 0 else 0.0 for i in range(n)]
        if self.fused_loops:
            from .batched import BatchedDirectInversionP2P
            from .ptp_utils import image2latent

            latent = image2latent(self.ldm_stable.vae, image_gt).to(self.ldm_stable.device, torch.float32)
            res = BatchedDirectInversionP2P(self.ldm_stable, n).edit(
                latent, [prompt_src], [prompt_tar], guidance_scale=forward_guidance,
                cross_replace_steps=cross_replace_steps, self_replace_steps=self_replace_steps, blend_word=blend_word,
                eq_params=eq_params, is_replace_controller=is_replace_controller, add_target=add_target,
                add_source=add_source, inverse_guidance_scale=inverse_guidance, loss_scales=loss_scales,
                device=self.device)
            x_stars, noise_loss_list, reconstruct_latent, latents = res.image(0)
            return self._panel(image_gt, prompt_src, prompt_tar, reconstruct_latent, latents, x_stars, noise_loss_list)
        inversion = DirectInversion(model=self.ldm_stable, num_ddim_steps=n)
        if inverse_guidance is not None:
            out = inversion.invert_with_guidance_scale_vary_guidance(
                image_gt=image_gt, prompt=prompts, inverse_guidance_scale=inverse_guidance,
                forward_guidance_scale=forward_guidance)
        elif scale is not None:
            out = inversion.invert_not_full(image_gt=image_gt, prompt=prompts, guidance_scale=forward_guidance, scale=scale)
        elif skip_step is not None:
            out = inversion.invert_skip_step(image_gt=image_gt, prompt=prompts, guidance_scale=forward_guidance,
                                             skip_step=skip_step)
        else:
            out = inversion.invert(image_gt=image_gt, prompt=prompts, guidance_scale=forward_guidance)
        _, _, x_stars, noise_loss_list = out
        x_t = x_stars[-1]
        fwd_losses = noise_loss_list
        if add_source:  # p2p_editor.py:930-932
            fwd_losses = [l[[0]].repeat(2, 1, 1, 1) for l in noise_loss_list]
        both = add_target or add_source
        fwd = direct_inversion_p2p_guidance_forward_add_target if both else direct_inversion_p2p_guidance_forward
        controller = AttentionStore()
        reconstruct_latent, x_t = fwd(model=self.ldm_stable, prompt=prompts, controller=controller,
                                      noise_loss_list=fwd_losses, latent=x_t, num_inference_steps=n,
                                      guidance_scale=forward_guidance, generator=None)
        controller = self._controller(prompts, cross_replace_steps, self_replace_steps, blend_word, eq_params,
                                      is_replace_controller)
        latents, _ = fwd(model=self.ldm_stable, prompt=prompts, controller=controller, noise_loss_list=fwd_losses,
                         latent=x_t, num_inference_steps=n, guidance_scale=forward_guidance, generator=None)
        return self._panel(image_gt, prompt_src, prompt_tar, reconstruct_latent, latents, x_stars, noise_loss_list)

    def edit_batch(self, images, prompts_src, prompts_tar, guidance_scale=7.5, cross_replace_steps=0.4,
                   self_replace_steps=0.6, blend_word=None, eq_params=None, is_replace_controller=False,
                   per_image_params=False, minimal=False):
        """`directinversion+p2p` for L images in ONE pass (UNet batch L for the inversion, 4 L afterwards): `images` is
        an (L,4,64,64) latent tensor or a list of L image paths / HWC uint8 arrays.  Returns batched.BatchEditResult
        for latents, a list of L reference-format PIL strips for images."""
        from .batched import BatchedDirectInversionP2P
        from .ptp_utils import image2latent

        is_latent = isinstance(images, torch.Tensor) and images.dim() == 4
        if is_latent:
            gts, lat = None, images
        else:
            gts = [load_512(im) for im in images]
            lat = torch.cat([image2latent(self.ldm_stable.vae, g) for g in gts])
        res = BatchedDirectInversionP2P(self.ldm_stable, self.num_ddim_steps).edit(
            lat.to(self.ldm_stable.device, torch.float32), list(prompts_src), list(prompts_tar),
            guidance_scale=guidance_scale, cross_replace_steps=cross_replace_steps,
            self_replace_steps=self_replace_steps, blend_word=blend_word, eq_params=eq_params,
            is_replace_controller=is_replace_controller, per_image_params=per_image_params, device=self.device,
            minimal=minimal)
        if is_latent or self.ldm_stable.vae is None:
            return res
        out = []
        for i, g in enumerate(gts):
            x_stars, nl, rec, lat_i = res.image(i)
            out.append(self._panel(g, prompts_src[i], prompts_tar[i], rec, lat_i, x_stars, nl))
        return out

    def edit_image_directinversion(self, image_path, prompt_src, prompt_tar, guidance_scale=7.5,
                                   cross_replace_steps=0.4, self_replace_steps=0.6, blend_word=None, eq_params=None,
                                   is_replace_controller=False, add_target=False):
        return self._direct(image_path, prompt_src, prompt_tar, guidance_scale, cross_replace_steps,
                            self_replace_steps, blend_word, eq_params, is_replace_controller, add_target=add_target)

    def edit_image_directinversion_vary_guidance_scale(self, image_path, prompt_src, prompt_tar,
                                                       inverse_guidance_scale=1, forward_guidance_scale=7.5,
                                                       cross_replace_steps=0.4, self_replace_steps=0.6, blend_word=None,
                                                       eq_params=None, is_replace_controller=False):
        """p2p_editor.py:481-548."""
        return self._direct(image_path, prompt_src, prompt_tar, forward_guidance_scale, cross_replace_steps,
                            self_replace_steps, blend_word, eq_params, is_replace_controller,
                            inverse_guidance=inverse_guidance_scale)

    def edit_image_directinversion_not_full(self, image_path, prompt_src, prompt_tar, guidance_scale=7.5,
                                            cross_replace_steps=0.4, self_replace_steps=0.6, blend_word=None,
                                            eq_params=None, is_replace_controller=False, scale=1.):
        """p2p_editor.py:707-773."""
        return self._direct(image_path, prompt_src, prompt_tar, guidance_scale, cross_replace_steps,
                            self_replace_steps, blend_word, eq_params, is_replace_controller, scale=scale)

    def edit_image_directinversion_skip_step(self, image_path, prompt_src, prompt_tar, skip_step, guidance_scale=7.5,
                                             cross_replace_steps=0.4, self_replace_steps=0.6, blend_word=None,
                                             eq_params=None, is_replace_controller=False):
        """p2p_editor.py:775-840."""
        return self._direct(image_path, prompt_src, prompt_tar, guidance_scale, cross_replace_steps,
                            self_replace_steps, blend_word, eq_params, is_replace_controller, skip_step=skip_step)

    def edit_image_directinversion_add_target(self, image_path, prompt_src, prompt_tar, guidance_scale=7.5,
                                              cross_replace_steps=0.4, self_replace_steps=0.6, blend_word=None,
                                              eq_params=None, is_replace_controller=False):
        """p2p_editor.py:842-907: the offsets of both branches are added back."""
        return self._direct(image_path, prompt_src, prompt_tar, guidance_scale, cross_replace_steps,
                            self_replace_steps, blend_word, eq_params, is_replace_controller, add_target=True)

    def edit_image_directinversion_add_source(self, image_path, prompt_src, prompt_tar, guidance_scale=7.5,
                                              cross_replace_steps=0.4, self_replace_steps=0.6, blend_word=None,
                                              eq_params=None, is_replace_controller=False):
        """p2p_editor.py:909-976: the SOURCE branch's offset is added to both branches."""
        return self._direct(image_path, prompt_src, prompt_tar, guidance_scale, cross_replace_steps,
                            self_replace_steps, blend_word, eq_params, is_replace_controller, add_source=True)

    def edit_image_ddim(self, image_path, prompt_src, prompt_tar, guidance_scale=7.5, cross_replace_steps=0.4,
                        self_replace_steps=0.6, blend_word=None, eq_params=None, is_replace_controller=False):
        """p2p_editor.py:137-197: NullInversion with num_inner_steps=0 (see inversion.NullInversion for the one
        result-neutral loop that is not executed), then plain P2P with the per-step unconditional embeddings."""
        from .inversion import NullInversion

        image_gt = self._load(image_path)
        prompts = [prompt_src, prompt_tar]
        null_inversion = NullInversion(model=self.ldm_stable, num_ddim_steps=self.num_ddim_steps)
        _, _, x_stars, uncond_embeddings = null_inversion.invert(image_gt=image_gt, prompt=prompt_src,
                                                                 guidance_scale=guidance_scale, num_inner_steps=0)
        x_t = x_stars[-1]
        controller = AttentionStore()
        reconstruct_latent, x_t = p2p_guidance_forward(model=self.ldm_stable, prompt=[prompt_src],
                                                       controller=controller, latent=x_t,
                                                       num_inference_steps=self.num_ddim_steps,
                                                       guidance_scale=guidance_scale, generator=None,
                                                       uncond_embeddings=uncond_embeddings)
        controller = self._controller(prompts, cross_replace_steps, self_replace_steps, blend_word, eq_params,
                                      is_replace_controller)
        latents, _ = p2p_guidance_forward(model=self.ldm_stable, prompt=prompts, controller=controller, latent=x_t,
                                          num_inference_steps=self.num_ddim_steps, guidance_scale=guidance_scale,
                                          generator=None, uncond_embeddings=uncond_embeddings)
        return self._panel(image_gt, prompt_src, prompt_tar, reconstruct_latent, latents, x_stars, None)

    def edit_image_negative_prompt_inversion(self, image_path, prompt_src, prompt_tar, guidance_scale=7.5, proximal=None,
                                             quantile=0.7, use_reconstruction_guidance=False, recon_t=400, recon_lr=0.1,
                                             npi_interp=0, cross_replace_steps=0.4, self_replace_steps=0.6,
                                             blend_word=None, eq_params=None, is_replace_controller=False,
                                             use_inversion_guidance=False, dilate_mask=1):
        """p2p_editor.py:324-413."""
        from .inversion import NegativePromptInversion
        from .p2p_guidance_forward import proximal_guidance_forward

        image_gt = self._load(image_path)
        prompts = [prompt_src, prompt_tar]
        null_inversion = NegativePromptInversion(model=self.ldm_stable, num_ddim_steps=self.num_ddim_steps)
        _, image_enc_latent, x_stars, uncond_embeddings = null_inversion.invert(image_gt=image_gt, prompt=prompt_src,
                                                                                npi_interp=npi_interp)
        x_t = x_stars[-1]
        controller = AttentionStore()
        reconstruct_latent, x_t = proximal_guidance_forward(
            model=self.ldm_stable, prompt=[prompt_src], controller=controller, latent=x_t,
            guidance_scale=guidance_scale, generator=None, uncond_embeddings=uncond_embeddings, edit_stage=False,
            prox=None, quantile=quantile, image_enc=None, recon_lr=recon_lr, recon_t=recon_t, inversion_guidance=False,
            x_stars=None, dilate_mask=dilate_mask)
        controller = self._controller(prompts, cross_replace_steps, self_replace_steps, blend_word, eq_params,
                                      is_replace_controller)
        guided = use_reconstruction_guidance or use_inversion_guidance
        latents, _ = proximal_guidance_forward(
            model=self.ldm_stable, prompt=prompts, controller=controller, latent=x_t, guidance_scale=guidance_scale,
            generator=None, uncond_embeddings=uncond_embeddings, edit_stage=True, prox=proximal, quantile=quantile,
            image_enc=image_enc_latent if use_reconstruction_guidance else None,
            recon_lr=recon_lr if guided else 0, recon_t=recon_t if guided else 1000, x_stars=x_stars,
            dilate_mask=dilate_mask)
        return self._panel(image_gt, prompt_src, prompt_tar, reconstruct_latent, latents, x_stars, None)
