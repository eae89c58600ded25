"""Loading a real SD-1.x checkpoint into the fused model (the `from_pretrained` side of seam B).

The reference builds its pipeline with `StableDiffusionPipeline.from_pretrained("CompVis/stable-diffusion-v1-4", ...)`
(`models/p2p_editor.py:23-25`, `run_editing_masactrl.py:64-66`).  There is no network here and `diffusers` is not
installed, so this module reads the files of such a checkpoint directly:

  <dir>/unet/diffusion_pytorch_model.safetensors | .bin     the 686 UNet tensors (diffusers key names)
  <dir>/tokenizer/, <dir>/text_encoder/                     CLIP, through `transformers` (local files only)

and hands the UNet state dict to `FusedModel` (which repacks it for the kernels).  Everything is validated against the
architecture table (`arch.unet_param_specs`) before a byte goes to the GPU: a wrong checkpoint fails with the list of
missing / unexpected / mis-shaped tensors instead of a shape error deep inside the engine.
"""
from __future__ import annotations

import os
from typing import Dict, List, Optional, Tuple

import torch

from . import arch

_UNET_FILES = ("diffusion_pytorch_model.safetensors", "diffusion_pytorch_model.fp16.safetensors",
               "diffusion_pytorch_model.bin", "diffusion_pytorch_model.fp16.bin")


def find_unet_file(path: str) -> str:
    """`path` may be the weights file itself, the `unet/` directory or the pipeline directory that contains it."""
    if os.path.isfile(path):
        return path
    for d in (path, os.path.join(path, "unet")):
        for f in _UNET_FILES:
            p = os.path.join(d, f)
            if os.path.isfile(p):
                return p
    raise FileNotFoundError(f"no UNet weights under {path!r} (looked for {', '.join(_UNET_FILES)} in it and in unet/)")


def read_state_dict(path: str) -> Dict[str, torch.Tensor]:
    """.safetensors through the safetensors library (zero-copy mmap), anything else through torch.load (weights only)."""
    if path.endswith(".safetensors"):
        from safetensors.torch import load_file

        return load_file(path, device="cpu")
    sd = torch.load(path, map_location="cpu", weights_only=True)
    if isinstance(sd, dict) and "state_dict" in sd and isinstance(sd["state_dict"], dict):
        sd = sd["state_dict"]
    return sd


def check_unet_state_dict(sd: Dict[str, torch.Tensor]) -> Tuple[List[str], List[str], List[str]]:
    """(missing, unexpected, mis-shaped) relative to the SD-1.x UNet2DConditionModel table."""
    specs = dict(arch.unet_param_specs())
    missing = [k for k in specs if k not in sd]
    unexpected = [k for k in sd if k not in specs]
    bad = [f"{k}: expected {specs[k]}, got {tuple(sd[k].shape)}" for k in specs if k in sd and tuple(sd[k].shape) != specs[k]]
    return missing, unexpected, bad


def load_unet_state_dict(path: str, strict: bool = True) -> Dict[str, torch.Tensor]:
    """The 686 tensors of an SD-1.x diffusers UNet, validated.  Original CompVis `.ckpt` files (keys starting with
    `model.diffusion_model.`) use the LDM naming and are rejected with an explicit message - convert them with diffusers'
    `convert_original_stable_diffusion_to_diffusers.py` first."""
    f = find_unet_file(path)
    sd = read_state_dict(f)
    if any(k.startswith("model.diffusion_model.") for k in sd):
        raise NotImplementedError(f"{f}: original LDM checkpoint layout; a diffusers-format UNet state dict is required")
    missing, unexpected, bad = check_unet_state_dict(sd)
    if missing or bad or (strict and unexpected):
        def head(xs):
            return ", ".join(xs[:5]) + (f", ... (+{len(xs) - 5})" if len(xs) > 5 else "")
        parts = []
        if missing:
            parts.append(f"{len(missing)} missing ({head(missing)})")
        if bad:
            parts.append(f"{len(bad)} with the wrong shape ({head(bad)})")
        if strict and unexpected:
            parts.append(f"{len(unexpected)} unexpected ({head(unexpected)})")
        raise ValueError(f"{f} is not an SD-1.x UNet2DConditionModel state dict: " + "; ".join(parts))
    return {k: sd[k] for k, _ in arch.unet_param_specs()}


def load_text_components(path: str, device="cpu", dtype=torch.float32):
    """(tokenizer, text_encoder) from <path>/tokenizer and <path>/text_encoder via transformers, local files only;
    (None, None) when the directories are absent (the caller then supplies its own, e.g. synth.FakeTokenizer)."""
    tok_dir, enc_dir = os.path.join(path, "tokenizer"), os.path.join(path, "text_encoder")
    if not (os.path.isdir(tok_dir) and os.path.isdir(enc_dir)):
        return None, None
    from transformers import CLIPTextModel, CLIPTokenizer

    tok = CLIPTokenizer.from_pretrained(tok_dir, local_files_only=True)
    enc = CLIPTextModel.from_pretrained(enc_dir, local_files_only=True).to(device=device, dtype=dtype).eval()
    return tok, enc


def load_fused_model(path: str, device="cuda:0", max_batch: int = 4, tokenizer=None, text_encoder=None, vae=None,
                     table_dtype: str = "float32", strict: bool = True):
    """`StableDiffusionPipeline.from_pretrained(path)` for the fused path: UNet -> libpnpinv engine, CLIP from the same
    directory when present.  The VAE is not built yet (section 8f): pass latents, or a `vae` object with the reference's
    `encode(...)['latent_dist'].mean` / `decode(...)['sample']` surface (utils/utils.py:58-80)."""
    from .model import FusedModel

    sd = load_unet_state_dict(path, strict=strict)
    if os.path.isdir(path) and (tokenizer is None or text_encoder is None):
        tok, enc = load_text_components(path, device=device)
        tokenizer = tokenizer if tokenizer is not None else tok
        text_encoder = text_encoder if text_encoder is not None else enc
    return FusedModel(sd, device=device, max_batch=max_batch, tokenizer=tokenizer, text_encoder=text_encoder, vae=vae,
                      table_dtype=table_dtype)
