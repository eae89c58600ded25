"""The fused VAE (csrc/vae.cu through the C ABI `pnp_vae_*`) against
  * tests/golden/vae_small.npz: outputs of the REFERENCE's vendored AutoencoderKL itself (fp64; 64x64 image, 8x8 latent),
  * oracle/vae_ref.py (pinned to that fixture by tests/test_oracle_cpu.py) at the full 512x512 size the editors use,
and the editor-level image path: P2PEditor.__call__ on an HWC uint8 image returns the reference's 2048x512 PIL strip.
Tolerance: fp16 operands / fp32 accumulation vs fp64, ~30 conv + GroupNorm stages: 5e-3 rel-L2 (stated, measured below)."""
import os

import numpy as np
import pytest
import torch

from pnpinversion_b200 import synth
from tests import gpu_util as G

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(__file__), "golden", "vae_small.npz")
TOL = 5e-3


@pytest.fixture(scope="module")
def vae(cuda):
    from pnpinversion_b200.vae import FusedVAE

    v = FusedVAE(synth.synth_vae_state_dict(0), device="cuda:0")
    yield v
    v.close()


def test_vae_matches_the_reference_fixture(vae, cuda):
    g = np.load(GOLD)
    img = torch.from_numpy(g["img"]).to(cuda)
    z = torch.from_numpy(g["z"]).to(cuda)
    dist = vae.encode(img)["latent_dist"]
    dec = vae.decode(z)["sample"]
    torch.cuda.synchronize()
    e_mean = G.rel_l2(dist.mean.cpu(), torch.from_numpy(g["mean"]))
    e_lv = G.rel_l2(dist.logvar.cpu(), torch.from_numpy(g["logvar"]))
    e_dec = G.rel_l2(dec.cpu(), torch.from_numpy(g["dec"]))
    print(f"vae vs reference fixture: mean {e_mean:.2e} logvar {e_lv:.2e} decode {e_dec:.2e}")
    assert e_mean < TOL and e_lv < TOL and e_dec < TOL


@pytest.mark.parametrize("size", [128, 512])
def test_vae_full_size_vs_oracle(vae, cuda, size):
    """image2latent / latent2image at the sizes the editors use (512x512 -> 64x64 latent), oracle in fp32 on the host."""
    from oracle import vae_ref

    ref = vae_ref.VaeRef(synth.synth_vae_state_dict(0), dtype=torch.float32)
    g = torch.Generator().manual_seed(77 + size)
    img = torch.rand(1, 3, size, size, generator=g) * 2 - 1
    z = torch.randn(1, 4, size // 8, size // 8, generator=g)
    with torch.no_grad():
        mean_ref, lv_ref = ref.encode_moments(img)
        dec_ref = ref.decode(z)
    dist = vae.encode(img.to(cuda))["latent_dist"]
    dec = vae.decode(z.to(cuda))["sample"]
    torch.cuda.synchronize()
    e_mean, e_dec = G.rel_l2(dist.mean.cpu(), mean_ref), G.rel_l2(dec.cpu(), dec_ref)
    print(f"vae {size}x{size} vs oracle: mean {e_mean:.2e} decode {e_dec:.2e}; launches {vae.kernel_launches()}")
    assert dec.shape == (1, 3, size, size) and dist.mean.shape == (1, 4, size // 8, size // 8)
    assert e_mean < TOL and e_dec < TOL


def test_vae_batch_rows_are_independent(vae, cuda):
    g = torch.Generator().manual_seed(5)
    z = torch.randn(3, 4, 16, 16, generator=g).to(cuda)
    all3 = vae.decode(z)["sample"]
    for i in range(3):
        one = vae.decode(z[i:i + 1])["sample"]
        assert G.rel_l2(all3[i:i + 1], one) < 2e-3


def test_editor_returns_the_reference_image_strip(cuda):
    """P2PEditor.__call__ on an HWC uint8 image (utils/utils.py:28-31 accepts an ndarray): VAE encode -> the four loops
    -> three decodes -> [instruction | source | reconstruction | edit] 2048x512 (models/p2p_editor.py:474-479)."""
    from PIL import Image

    from pnpinversion_b200.model import FusedModel
    from pnpinversion_b200.p2p_editor import P2PEditor

    model = FusedModel.synthetic(device="cuda:0", max_batch=4, with_vae=True)
    rng = np.random.RandomState(3)
    yy, xx = np.mgrid[0:480, 0:640]
    img = np.stack([(yy * 255 // 480), (xx * 255 // 640), ((yy + xx) % 256)], axis=-1).astype(np.uint8)
    img = (img.astype(np.int32) + rng.randint(-8, 8, img.shape)).clip(0, 255).astype(np.uint8)  # non-square input
    ed = P2PEditor(["directinversion+p2p"], "cuda:0", num_ddim_steps=4, model=model)
    src, tgt = synth.CAT_PROMPTS
    out = ed("directinversion+p2p", image_path=img, prompt_src=src, prompt_tar=tgt, blend_word=(("cat",), ("cat",)),
             eq_params={"words": ("watercolor",), "values": (2,)})
    assert isinstance(out, Image.Image) and out.size == (2048, 512)
    a = np.asarray(out)
    from pnpinversion_b200.ptp_utils import load_512

    assert np.array_equal(a[:, 512:1024], load_512(img))  # panel 2 is the (centre-cropped, resized) input
    rec, edit = a[:, 1024:1536].astype(np.float64), a[:, 1536:].astype(np.float64)
    assert rec.std() > 1.0 and edit.std() > 1.0 and np.abs(rec - edit).mean() > 0.1
    # the reconstruction panel decodes the rectified source branch = the VAE round trip of the input latent
    from pnpinversion_b200.ptp_utils import image2latent, latent2image

    z0 = image2latent(model.vae, load_512(img))
    rt = latent2image(model.vae, z0)[0].astype(np.float64)
    assert np.abs(rt - rec).mean() < 1.0  # uint8 levels
    model.unet.close()
    model.vae.close()
