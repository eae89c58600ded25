"""evaluation.py: the network-free PIE-Bench metrics (PSNR / MSE / SSIM, whole / unedited / edited part) and the csv the
reference's evaluate.py writes."""
import csv
import os

import numpy as np
from PIL import Image

from pnpinversion_b200 import cli, evaluation


def test_metrics_known_answers():
    rng = np.random.RandomState(0)
    a = rng.randint(0, 256, (512, 512, 3)).astype(np.uint8)
    b = a.copy()
    assert evaluation.calculate_mse(a, b) == 0 and evaluation.calculate_psnr(a, b) == float("inf")
    assert abs(evaluation.calculate_ssim(a, b) - 1.0) < 1e-12
    b = np.clip(a.astype(np.int32) + 10, 0, 255).astype(np.uint8)
    mse = np.mean(((a.astype(np.float64) - b) / 255) ** 2)
    assert abs(evaluation.calculate_mse(a, b) - mse) < 1e-9
    assert abs(evaluation.calculate_psnr(a, b) - 10 * np.log10(1 / mse)) < 1e-6
    assert 0.8 < evaluation.calculate_ssim(a, b) < 1.0
    noise = rng.randint(0, 256, a.shape).astype(np.uint8)
    assert evaluation.calculate_ssim(a, noise) < 0.1
    # masks multiply both images: outside the mask the difference does not count
    mask = np.zeros((512, 512, 3))
    mask[:256] = 1
    c = a.copy()
    c[256:] = 0
    assert evaluation.calculate_metric("mse_edit_part", a, c, mask, mask) == 0.0
    assert evaluation.calculate_metric("mse_unedit_part", a, c, mask, mask) > 0.1
    assert evaluation.calculate_metric("lpips", a, c, mask, mask) == "nan"
    assert evaluation.calculate_metric("psnr_unedit_part", a, c, np.ones_like(mask), np.ones_like(mask)) == "nan"


def test_evaluate_writes_the_reference_csv(tmp_path):
    data = str(tmp_path / "data")
    cli.write_synthetic_dataset(data, n_items=3, size=512)
    items = cli.read_items(data, cli.CATEGORIES)
    out = str(tmp_path / "output" / "directinversion+p2p" / "annotation_images")
    for it in items:  # a fake result strip whose edit panel is the source with a brightness shift
        src = np.asarray(Image.open(it["image_path"]).convert("RGB"))
        edit = np.clip(src.astype(np.int32) + 5, 0, 255).astype(np.uint8)
        strip = np.concatenate([np.full_like(src, 255), src, src, edit], axis=1)
        dst = it["image_path"].replace(os.path.join(data, "annotation_images"), out)
        os.makedirs(os.path.dirname(dst), exist_ok=True)
        Image.fromarray(strip).save(dst, quality=100, subsampling=0)
    res = str(tmp_path / "evaluation_result.csv")
    rows = evaluation.evaluate(os.path.join(data, "mapping_file.json"), os.path.join(data, "annotation_images"),
                               {"1_directinversion+p2p": out}, ["psnr", "mse_unedit_part", "ssim", "lpips"], res)
    with open(res) as f:
        got = list(csv.reader(f))
    assert got[0] == ["file_id", "1_directinversion+p2p|psnr", "1_directinversion+p2p|mse_unedit_part",
                      "1_directinversion+p2p|ssim", "1_directinversion+p2p|lpips"]
    assert len(got) == 4 and got[1][0] == items[0]["key"] and got[1][4] == "nan"
    assert 30 < float(got[1][1]) < 40 and float(got[1][3]) > 0.9  # +5/255 brightness: PSNR ~34 dB
    # the reconstruction panel equals the source up to JPEG
    rows = evaluation.evaluate(os.path.join(data, "mapping_file.json"), os.path.join(data, "annotation_images"),
                               {"m": out}, ["psnr"], res, reconstruction=True)
    assert float(rows[1][1]) > 35
