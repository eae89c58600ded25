"""MasaCtrl path: the fused UNet with the MutualSelfAttentionControl descriptor vs the
REFERENCE's own editor (models/masactrl/masactrl.py + masactrl_utils.py registered on the vendored UNet, fp64; fixture
tests/golden/masactrl_forward.npz), plus loop-level invariants of the rectified MasaCtrl sampler."""
import os

import numpy as np
import pytest
import torch

from pnpinversion_b200 import synth
from pnpinversion_b200.masactrl import (AttentionBase, MasaCtrlEditor, MutualSelfAttentionControl,
                                        regiter_attention_editor_diffusers)
from pnpinversion_b200.model import FusedModel
from tests import gpu_util as G

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(__file__), "golden", "masactrl_forward.npz")


@pytest.fixture(scope="module")
def model(cuda):
    m = FusedModel.synthetic(device="cuda:0", max_batch=8)
    yield m
    m.unet.close()


def _ctx(model, prompts):
    tok, te = model.tokenizer, model.text_encoder
    return torch.cat([te(tok([""] * len(prompts)).input_ids)[0], te(tok(prompts).input_ids)[0]]).to(model.device,
                                                                                                    torch.float32)


@pytest.mark.parametrize("name,step", [("on", 10), ("off", 2)])
def test_forward_with_mutual_self_attention_matches_reference(model, name, step):
    g = np.load(GOLD)
    prompts = ["", synth.CAT_PROMPTS[1]]
    ctx = _ctx(model, prompts)
    editor = MutualSelfAttentionControl(4, 10)
    regiter_attention_editor_diffusers(model, editor)
    assert editor.num_att_layers == 32
    editor.cur_step = step
    lat = torch.cat([synth.synth_latent(0), synth.synth_latent(1)]).to(model.device)
    eps = model.unet(torch.cat([lat] * 2), 401, encoder_hidden_states=ctx)["sample"]
    torch.cuda.synchronize()
    ref = torch.from_numpy(g[f"{name}_eps"])
    err = [G.rel_l2(eps[i].cpu(), ref[i]) for i in range(4)]
    print(f"masactrl {name}: per-row rel-L2 {err}")
    assert max(err) < 5e-3
    assert editor.cur_step == step + 1
    # the source rows (first image of each CFG half) are never touched by the editor: bit-identical to a plain call
    regiter_attention_editor_diffusers(model, AttentionBase())
    plain = model.unet(torch.cat([lat] * 2), 401, encoder_hidden_states=ctx)["sample"]
    torch.cuda.synchronize()
    assert torch.equal(plain[0], eps[0]) and torch.equal(plain[2], eps[2])
    if name == "on":
        assert not torch.equal(plain[1], eps[1]) and not torch.equal(plain[3], eps[3])
    else:
        assert torch.equal(plain, eps)


def test_directinversion_masactrl_loop_invariants(model):
    editor = MasaCtrlEditor(["directinversion+masactrl"], "cuda:0", num_ddim_steps=10, model=model)
    z0 = synth.synth_latent(3)
    res = editor("directinversion+masactrl", z0, "", synth.CAT_PROMPTS[1], guidance_scale=7.5, step=4, layper=10)
    torch.cuda.synchronize()
    assert len(res.x_stars) == 11 and len(res.noise_loss_list) == 10
    # rectified source branch lands on the inverted image latent (exactness invariant)
    assert (res.latents[0].cpu() - z0[0]).abs().max() < 2e-5
    # the edit differs from both the source and the direct synthesis
    assert G.rel_l2(res.latents[1].cpu(), z0[0]) > 1e-2
    assert G.rel_l2(res.latents[1].cpu(), res.latents_fixed[0].cpu()) > 1e-3
    with pytest.raises(NotImplementedError):
        editor("no-such-method", z0, "", "x", 7.5)


def test_masactrl_image_batch_through_the_c_loops_matches_single_images(model):
    """BASELINE config 4 in small: two images per pass (UNet batch 2 / 8 / 4 / 8, loops inside pnp_run_loop) against the
    single-image Python-loop editor; source rows prompt-major, every image's queries attend to ITS source's keys/values."""
    editor = MasaCtrlEditor(["directinversion+masactrl"], "cuda:0", num_ddim_steps=6, model=model)
    zs = torch.cat([synth.synth_latent(3), synth.synth_latent(4)]).cuda()
    tars = [synth.CAT_PROMPTS[1], "a photo of a red house on a snowy hill at night"]
    res = editor.edit_batch(zs, tars, guidance_scale=7.5, step=2, layper=10)
    torch.cuda.synchronize()
    assert res.latents.shape == (4, 4, 64, 64) and torch.isfinite(res.latents).all()
    for i in range(2):
        one = editor("directinversion+masactrl", zs[i:i + 1], "", tars[i], guidance_scale=7.5, step=2, layper=10)
        torch.cuda.synchronize()
        assert (res.latents[i] - zs[i]).abs().max() < 2e-5  # rectified source branch = the inverted latent
        e_fixed = G.rel_l2(res.latents_fixed[i], one.latents_fixed[0])
        e_edit = G.rel_l2(res.latents[2 + i], one.latents[1])
        print(f"masactrl image {i}: batched vs single: fixed {e_fixed:.2e} edit {e_edit:.2e}")
        assert e_fixed < 8e-2 and e_edit < 8e-2  # CFG 7.5 on rounding-level differences of two tilings, 6 steps
    assert G.rel_l2(res.latents[2], res.latents[3]) > 1e-2
