"""PIE-Bench evaluation of the editors' output strips -- the part of `evaluation/evaluate.py` + `matrics_calculator.py`
that needs no pretrained network: PSNR, MSE and SSIM between the source image and the EDIT panel (the last 512x512 of
the 2048x512 strip, evaluate.py:271-273), on the whole image, on the unedited part (`1 - mask`) and on the edited part
(`mask`), written as the reference's csv (`file_id`, then `<method>|<metric>` columns, evaluate.py:236-246,279-282).

The reference computes them with torchmetrics (absent offline); the formulas are restated here with torch ops:
PSNR = 10 log10(1 / MSE) at data_range 1 (PeakSignalNoiseRatio), MSE = mean squared error, SSIM = the Gaussian
(11x11, sigma 1.5, K = 0.01 / 0.03, data_range 1) structural similarity of StructuralSimilarityIndexMeasure with its
reflect padding.  LPIPS, CLIP similarity and the DINO structure distance need pretrained weights and report "nan".
Masks multiply BOTH images before the metric, exactly like matrics_calculator.py:309-314."""
from __future__ import annotations

import csv
import json
import os
from typing import Dict, Iterable, List, Optional

import numpy as np
import torch
import torch.nn.functional as F
from PIL import Image

from .cli import mask_decode

NETWORK_METRICS = ("lpips", "structure_distance", "clip_similarity")


def _prep(img, mask):
    a = np.array(img).astype(np.float32) / 255
    if mask is not None:
        a = a * np.array(mask).astype(np.float32)
    return torch.from_numpy(a).permute(2, 0, 1).unsqueeze(0).double()


def calculate_mse(img_pred, img_gt, mask_pred=None, mask_gt=None) -> float:
    """matrics_calculator.py:344-362."""
    return float(((_prep(img_pred, mask_pred) - _prep(img_gt, mask_gt)) ** 2).mean())


def calculate_psnr(img_pred, img_gt, mask_pred=None, mask_gt=None) -> float:
    """matrics_calculator.py:304-322 (PeakSignalNoiseRatio(data_range=1.0))."""
    mse = calculate_mse(img_pred, img_gt, mask_pred, mask_gt)
    return float("inf") if mse == 0 else float(10.0 * np.log10(1.0 / mse))


def calculate_ssim(img_pred, img_gt, mask_pred=None, mask_gt=None) -> float:
    """matrics_calculator.py:364-382 (StructuralSimilarityIndexMeasure(data_range=1.0): Gaussian kernel 11, sigma 1.5)."""
    x, y = _prep(img_pred, mask_pred), _prep(img_gt, mask_gt)
    k, sigma, c1, c2 = 11, 1.5, 0.01 ** 2, 0.03 ** 2
    ax = torch.arange(k, dtype=torch.float64) - (k - 1) / 2
    g = torch.exp(-(ax / sigma) ** 2 / 2)
    g = g / g.sum()
    w = (g[:, None] * g[None, :]).expand(x.shape[1], 1, k, k).contiguous()
    pad = (k - 1) // 2
    xp, yp = F.pad(x, (pad,) * 4, mode="reflect"), F.pad(y, (pad,) * 4, mode="reflect")
    ch = x.shape[1]
    mu_x, mu_y = F.conv2d(xp, w, groups=ch), F.conv2d(yp, w, groups=ch)
    sxx = F.conv2d(xp * xp, w, groups=ch) - mu_x ** 2
    syy = F.conv2d(yp * yp, w, groups=ch) - mu_y ** 2
    sxy = F.conv2d(xp * yp, w, groups=ch) - mu_x * mu_y
    ssim = ((2 * mu_x * mu_y + c1) * (2 * sxy + c2)) / ((mu_x ** 2 + mu_y ** 2 + c1) * (sxx + syy + c2))
    return float(ssim[..., pad:-pad, pad:-pad].mean())  # torchmetrics crops the padded border again


_FN = {"psnr": calculate_psnr, "mse": calculate_mse, "ssim": calculate_ssim}


def calculate_metric(metric: str, src_image, tgt_image, src_mask, tgt_mask):
    """evaluate.py:30-114 for the metrics that need no network."""
    base = metric.replace("_unedit_part", "").replace("_edit_part", "")
    if base.startswith(NETWORK_METRICS):
        return "nan"
    fn = _FN[base]
    if metric.endswith("_unedit_part"):
        if (1 - src_mask).sum() == 0 or (1 - tgt_mask).sum() == 0:
            return "nan"
        return fn(src_image, tgt_image, 1 - src_mask, 1 - tgt_mask)
    if metric.endswith("_edit_part"):
        if src_mask.sum() == 0 or tgt_mask.sum() == 0:
            return "nan"
        return fn(src_image, tgt_image, src_mask, tgt_mask)
    return fn(src_image, tgt_image, None, None)


def edit_panel(strip: Image.Image, reconstruction: bool = False) -> Image.Image:
    """evaluate.py:271-275: the last 512x512 of a non-square result is the edit, the one before it the reconstruction."""
    if strip.size[0] == strip.size[1]:
        return strip
    w, h = strip.size
    if reconstruction:
        return strip.crop((w - 512 * 2, h - 512, w - 512, h))
    return strip.crop((w - 512, h - 512, w, h))


def evaluate(annotation_mapping_file: str, src_image_folder: str, tgt_image_folders: Dict[str, str], metrics: Iterable[str],
             result_path: str, edit_category_list: Optional[List[str]] = None, reconstruction: bool = False) -> List[list]:
    """evaluate.py:232-282: one csv row per annotated image, `<method>|<metric>` columns."""
    metrics = list(metrics)
    cats = set(edit_category_list or [str(i) for i in range(10)])
    rows = [["file_id"] + [f"{k}|{m}" for k in tgt_image_folders for m in metrics]]
    with open(annotation_mapping_file, "r") as f:
        annotation = json.load(f)
    for key, item in annotation.items():
        if item["editing_type_id"] not in cats:
            continue
        mask = mask_decode(item["mask"])[:, :, np.newaxis].repeat(3, axis=2)
        src = Image.open(os.path.join(src_image_folder, item["image_path"])).convert("RGB")
        row = [key]
        for _, folder in tgt_image_folders.items():
            tgt = edit_panel(Image.open(os.path.join(folder, item["image_path"])).convert("RGB"), reconstruction)
            row += [calculate_metric(m, src, tgt, mask, mask) for m in metrics]
        rows.append(row)
    with open(result_path, "w", newline="") as f:
        csv.writer(f).writerows(rows)
    return rows
