"""Host-side glue mirroring the reference's `utils/utils.py` (init_latent :48-55, latent2image :58-66,
image2latent :68-80, get_word_inds :84-102, get_time_words_attention_alpha :117-135, load_512 :27-46)."""
from __future__ import annotations

from typing import Dict, Optional, Sequence, Union

import numpy as np
import torch

from .seq_aligner import get_word_inds  # noqa: F401  (re-exported, the reference has it in both modules)


def _read_rgb(image):
    """A file path -> HWC uint8 array (first three channels); arrays pass through."""
    if isinstance(image, str):
        from PIL import Image

        return np.asarray(Image.open(image))[..., :3]
    return image


def load_512(image_path, left=0, right=0, top=0, bottom=0):
    """Path or HWC uint8 array -> 512x512x3 uint8: optional margins, centre crop to a square, resize.

    Same results as `utils/utils.py:27-46`, including its margin clamping - each margin is limited so that at least one
    pixel survives, and the TOP margin is limited by the LEFT one (`h - left - 1`, the reference's arithmetic, kept so
    that callers who pass margins get identical crops)."""
    from PIL import Image

    img = _read_rgb(image_path)
    rows, cols = img.shape[:2]
    left = min(left, cols - 1)
    right = min(right, cols - left - 1)
    top = min(top, rows - left - 1)  # sic: limited by `left`, as in the reference
    bottom = min(bottom, rows - top - 1)
    img = img[top:rows - bottom, left:cols - right]
    rows, cols = img.shape[:2]
    side = min(rows, cols)
    r0, c0 = (rows - side) // 2, (cols - side) // 2  # the longer axis loses (longer - shorter) // 2 at its start
    return np.array(Image.fromarray(img[r0:r0 + side, c0:c0 + side]).resize((512, 512)))


def init_latent(latent, model, height, width, generator, batch_size):
    """utils/utils.py:48-55: the start latent (drawn when absent) and its batch-expanded view on the model's device."""
    shape = (model.unet.in_channels, height // 8, width // 8)
    if latent is None:
        latent = torch.randn((1, *shape), generator=generator)
    return latent, latent.expand(batch_size, *shape).to(model.device)


@torch.no_grad()
def latent2image(model, latents, return_type="np", rounding=False):
    """`model` is the VAE handle.  utils/utils.py:58-66: decode(z / 0.18215), to [0,1], HWC uint8 by TRUNCATION;
    EDICT's prep_image_for_return rounds instead (edict_functions.py:698) -> `rounding=True`."""
    image = model.decode(latents.detach() * (1 / 0.18215))["sample"]
    if return_type != "np":
        return image
    hwc = (image / 2 + 0.5).clamp(0, 1).cpu().permute(0, 2, 3, 1).numpy() * 255
    return (hwc.round() if rounding else hwc).astype(np.uint8)


@torch.no_grad()
def image2latent(model, image):
    """`model` is the VAE handle.  utils/utils.py:68-80: 4-D tensors are latents already; an HWC uint8 image becomes
    encode(x / 127.5 - 1).latent_dist.mean * 0.18215."""
    if isinstance(image, torch.Tensor) and image.dim() == 4:
        return image
    x = torch.from_numpy(np.asarray(image)).float().div(127.5).sub(1).permute(2, 0, 1)[None].to(model.device)
    return model.encode(x)["latent_dist"].mean * 0.18215


def _set_window(alpha, bounds, prompt_ind, word_inds=None):
    if isinstance(bounds, float):
        bounds = (0, bounds)
    start, end = int(bounds[0] * alpha.shape[0]), int(bounds[1] * alpha.shape[0])
    if word_inds is None:
        word_inds = torch.arange(alpha.shape[2])
    alpha[:start, prompt_ind, word_inds] = 0
    alpha[start:end, prompt_ind, word_inds] = 1
    alpha[end:, prompt_ind, word_inds] = 0
    return alpha


def get_time_words_attention_alpha(prompts: Sequence[str], num_steps: int,
                                   cross_replace_steps: Union[float, Dict[str, object]], tokenizer,
                                   max_num_words: int = 77) -> torch.Tensor:
    """(num_steps+1, len(prompts)-1, 1, 1, 77) gate table: NB num_steps+1 rows, so the window is int(f*(n+1))."""
    if not isinstance(cross_replace_steps, dict):
        cross_replace_steps = {"default_": cross_replace_steps}
    if "default_" not in cross_replace_steps:
        cross_replace_steps["default_"] = (0.0, 1.0)
    table = torch.zeros(num_steps + 1, len(prompts) - 1, max_num_words)
    for i in range(len(prompts) - 1):
        table = _set_window(table, cross_replace_steps["default_"], i)
    for word, bounds in cross_replace_steps.items():
        if word == "default_":
            continue
        for i, ind in enumerate(get_word_inds(prompts[k], word, tokenizer) for k in range(1, len(prompts))):
            if len(ind) > 0:
                table = _set_window(table, bounds, i, ind)
    return table.reshape(num_steps + 1, len(prompts) - 1, 1, 1, max_num_words)


def txt_draw(text, target_size=(512, 512)):
    """utils/utils.py:137-155 renders the instruction panel with matplotlib (absent offline); this draws the same text,
    wrapped, black on white, top-left anchored, with PIL.  The panel is cosmetic: the evaluator crops it away
    (evaluation/evaluate.py:271-273)."""
    import textwrap

    from PIL import Image, ImageDraw, ImageFont

    w, h = int(target_size[0]), int(target_size[1])
    img = Image.new("RGB", (w, h), (255, 255, 255))
    draw = ImageDraw.Draw(img)
    try:
        font = ImageFont.load_default(size=20)
    except TypeError:  # Pillow < 10.1
        font = ImageFont.load_default()
    y = 12
    for para in str(text).split("\n"):
        for line in textwrap.wrap(para, width=46) or [""]:
            draw.text((12, y), line, fill=(0, 0, 0), font=font)
            y += 26
    return np.asarray(img)[:, :, :3]
