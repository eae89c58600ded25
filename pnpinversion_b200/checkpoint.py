"""Loading a real SD-1.x checkpoint into the fused model (the `from_pretrained` side of seam B).

The reference builds its pipeline with `StableDiffusionPipeline.from_pretrained("CompVis/stable-diffusion-v1-4", ...)`
(`models/p2p_editor.py:23-25`, `run_editing_masactrl.py:64-66`).  There is no network here and `diffusers` is not
installed, so this module reads the files of such a checkpoint directly:

  <dir>/unet/diffusion_pytorch_model.safetensors | .bin     the 686 UNet tensors (diffusers key names)
  <file>.safetensors | .ckpt (CompVis / LDM single file)     same tensors under model.diffusion_model.* (converted)
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


_LDM_PREFIX = "model.diffusion_model."
_LDM_RESNET = {"in_layers.0": "norm1", "in_layers.2": "conv1", "emb_layers.1": "time_emb_proj", "out_layers.0": "norm2",
               "out_layers.3": "conv2", "skip_connection": "conv_shortcut"}


def ldm_to_diffusers_unet(sd: Dict[str, torch.Tensor]) -> Dict[str, torch.Tensor]:
    """Original CompVis / LDM single-file checkpoints (`v1-5-pruned-emaonly.safetensors`, `sd-v1-4.ckpt`: UNet keys under
    `model.diffusion_model.`) -> diffusers key names.  SD-1.x layout: `input_blocks.0` = conv_in, `input_blocks.{1..11}` =
    three entries per resolution (resnet [+ transformer]) x2 then a stride-2 conv (`op`), `middle_block.{0,1,2}`,
    `output_blocks.{0..11}` = (resnet [+ transformer] [+ upsample conv]) three per resolution, `time_embed.{0,2}`,
    `out.{0,2}`.  Tensors are passed through untouched (the 1x1 proj_in / proj_out convolutions keep their 4-D shape in
    SD-1.x); anything that is not a UNet tensor (VAE, CLIP, EMA bookkeeping) is dropped."""
    def resnet(dst, rest):
        for a, b in _LDM_RESNET.items():
            if rest.startswith(a + "."):
                return f"{dst}.{b}.{rest[len(a) + 1:]}"
        raise ValueError(f"unexpected resnet tensor {rest!r}")

    out: Dict[str, torch.Tensor] = {}
    for key, t in sd.items():
        if not key.startswith(_LDM_PREFIX):
            continue
        k = key[len(_LDM_PREFIX):]
        p = k.split(".")
        if p[0] == "time_embed":
            new = f"time_embedding.linear_{1 if p[1] == '0' else 2}.{p[2]}"
        elif p[0] == "out":
            new = f"{'conv_norm_out' if p[1] == '0' else 'conv_out'}.{p[2]}"
        elif p[0] == "input_blocks":
            i, sub, rest = int(p[1]), p[2], ".".join(p[3:])
            if i == 0:
                new = f"conv_in.{rest}"
            else:
                b, l = (i - 1) // 3, (i - 1) % 3
                if rest.startswith("op."):
                    new = f"down_blocks.{b}.downsamplers.0.conv.{rest[3:]}"
                elif sub == "0":
                    new = resnet(f"down_blocks.{b}.resnets.{l}", rest)
                else:
                    new = f"down_blocks.{b}.attentions.{l}.{rest}"
        elif p[0] == "middle_block":
            sub, rest = p[1], ".".join(p[2:])
            new = f"mid_block.attentions.0.{rest}" if sub == "1" else resnet(f"mid_block.resnets.{0 if sub == '0' else 1}", rest)
        elif p[0] == "output_blocks":
            i, sub, rest = int(p[1]), p[2], ".".join(p[3:])
            b, l = i // 3, i % 3
            if rest.startswith("conv."):  # the upsampler: entry 1 of a block without attention, entry 2 otherwise
                new = f"up_blocks.{b}.upsamplers.0.{rest}"
            elif sub == "0":
                new = resnet(f"up_blocks.{b}.resnets.{l}", rest)
            else:
                new = f"up_blocks.{b}.attentions.{l}.{rest}"
        else:
            raise ValueError(f"unexpected UNet tensor {key!r}")
        out[new] = t
    return out


def load_unet_state_dict(path: str, strict: bool = True) -> Dict[str, torch.Tensor]:
    """The 686 tensors of an SD-1.x UNet, validated: diffusers layout as is, original CompVis / LDM single-file
    checkpoints (keys under `model.diffusion_model.`) through `ldm_to_diffusers_unet`."""
    f = path if os.path.isfile(path) else find_unet_file(path)
    sd = read_state_dict(f)
    if any(k.startswith(_LDM_PREFIX) for k in sd):
        sd = ldm_to_diffusers_unet(sd)
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
    """(tokenizer, text_encoder) from <path>/tokenizer and <path>/text_encoder, local files only; (None, None) when the
    directories are absent (the caller then supplies its own, e.g. synth.FakeTokenizer).  The tokenizer is
    `transformers.CLIPTokenizer`.  On a CUDA device the text encoder is the fused one (clip.py -> csrc/clip.cu), fed
    with the weights of `text_encoder/` read straight from its safetensors / bin file; device="cpu" returns the
    `transformers.CLIPTextModel` itself (used by the CPU tests as the comparison model, never by the editors)."""
    tok_dir, enc_dir = os.path.join(path, "tokenizer"), os.path.join(path, "text_encoder")
    if not (os.path.isdir(tok_dir) and os.path.isdir(enc_dir)):
        return None, None
    from transformers import CLIPTokenizer

    tok = CLIPTokenizer.from_pretrained(tok_dir, local_files_only=True)
    if torch.device(device).type == "cuda":
        from .clip import FusedCLIPTextEncoder

        return tok, FusedCLIPTextEncoder(load_clip_state_dict(enc_dir), device=device)
    from transformers import CLIPTextModel

    enc = CLIPTextModel.from_pretrained(enc_dir, local_files_only=True).to(device=device, dtype=dtype).eval()
    return tok, enc


def load_clip_state_dict(path: str) -> Dict[str, torch.Tensor]:
    """The CLIPTextModel tensors of `<path>` (= `<checkpoint>/text_encoder`, or a weight file): `model.safetensors` /
    `pytorch_model.bin`, validated against clip.clip_text_param_specs for the layer count / vocabulary found in the file."""
    from .clip import clip_text_param_specs, count_layers

    if os.path.isdir(path):
        cands = [os.path.join(path, n) for n in ("model.safetensors", "model.fp16.safetensors", "pytorch_model.bin",
                                                 "pytorch_model.fp16.bin")]
        f = next((c for c in cands if os.path.isfile(c)), None)
        if f is None:
            raise FileNotFoundError(f"no text-encoder weight file under {path}")
    else:
        f = path
    sd = read_state_dict(f)
    layers = count_layers(sd)
    tok = sd.get("text_model.embeddings.token_embedding.weight")
    if layers == 0 or tok is None:
        raise ValueError(f"{f} is not a CLIPTextModel state dict (text_model.* tensors missing)")
    res = {}
    for name, shape in clip_text_param_specs(layers, tok.shape[0]):
        if name not in sd:
            raise ValueError(f"{f}: {name} missing")
        if tuple(sd[name].shape) != tuple(shape):
            raise ValueError(f"{f}: {name} has shape {tuple(sd[name].shape)}, expected {shape}")
        res[name] = sd[name]
    return res


# diffusers >= 0.15 renamed the VAE attention projections; the engine takes the 0.3.0 .. 0.14 names of the reference's pins
_VAE_ATTN_RENAMES = (("to_q.", "query."), ("to_k.", "key."), ("to_v.", "value."), ("to_out.0.", "proj_attn."))


def load_vae_state_dict(path: str) -> Dict[str, torch.Tensor]:
    """The 248 AutoencoderKL tensors of `<path>/vae/` (or a single VAE weight file), diffusers names, validated against
    arch.vae_param_specs (linear attention weights may come as (C,C,1,1) convs)."""
    if os.path.isdir(path):
        d = os.path.join(path, "vae") if os.path.isdir(os.path.join(path, "vae")) else path
        cands = [os.path.join(d, n) for n in ("diffusion_pytorch_model.safetensors", "diffusion_pytorch_model.bin")]
        f = next((c for c in cands if os.path.isfile(c)), None)
        if f is None:
            raise FileNotFoundError(f"no VAE weight file under {d}")
    else:
        f = path
    sd = read_state_dict(f)
    out = {}
    for k, v in sd.items():
        if ".attentions." in k:
            for new, old in _VAE_ATTN_RENAMES:
                k = k.replace("." + new, "." + old)
        out[k] = v
    res = {}
    for name, shape in arch.vae_param_specs():
        if name not in out:
            raise ValueError(f"{f} is not an SD-1.x AutoencoderKL state dict: {name} missing")
        t = out[name]
        n = 1
        for s_ in shape:
            n *= s_
        if t.numel() != n:
            raise ValueError(f"{f}: {name} has shape {tuple(t.shape)}, expected {shape}")
        res[name] = t.reshape(shape)
    return res


def load_fused_model(path: str, device="cuda:0", max_batch: int = 4, tokenizer=None, text_encoder=None, vae=None,
                     table_dtype: str = "float32", strict: bool = True):
    """`StableDiffusionPipeline.from_pretrained(path)` for the fused path: UNet -> libpnpinv engine, VAE -> the fused VAE
    (csrc/vae.cu) when `<path>/vae` exists, CLIP from the same directory when present."""
    from .model import FusedModel

    sd = load_unet_state_dict(path, strict=strict)
    if vae is None and os.path.isdir(path) and os.path.isdir(os.path.join(path, "vae")):
        from .vae import FusedVAE

        vae = FusedVAE(load_vae_state_dict(path), device=device)
    if os.path.isdir(path) and (tokenizer is None or text_encoder is None):
        tok, enc = load_text_components(path, device=device)
        tokenizer = tokenizer if tokenizer is not None else tok
        text_encoder = text_encoder if text_encoder is not None else enc
    return FusedModel(sd, device=device, max_batch=max_batch, tokenizer=tokenizer, text_encoder=text_encoder, vae=vae,
                      table_dtype=table_dtype)
