"""Pins the CPU oracle (oracle/unet_ref.py, oracle/p2p_ref.py) against fixtures produced by the REFERENCE's own code
(vendored UNet2DConditionModel + models/p2p/*.py, see oracle/make_golden.py)."""
import json
import os

import numpy as np
import pytest
import torch

from oracle import p2p_ref, unet_ref
from pnpinversion_b200 import ptp_utils, seq_aligner, synth
from pnpinversion_b200.attention_control import get_equalizer

GOLD = os.path.join(os.path.dirname(__file__), "golden")


def _rel(a, b):
    a, b = a.double().flatten(), b.double().flatten()
    return float((a - b).norm() / b.norm())


@pytest.fixture(scope="module")
def ref_unet():
    torch.set_grad_enabled(False)
    return unet_ref.UNetRef(synth.synth_unet_state_dict(0))


def _ctx(prompts):
    tok, te = synth.FakeTokenizer(), synth.SynthTextEncoder(dtype=torch.float64)
    return torch.cat([te(tok([""] * len(prompts)).input_ids)[0], te(tok(prompts).input_ids)[0]])


def test_oracle_unet_forward_matches_reference_unet(ref_unet):
    g = np.load(os.path.join(GOLD, "unet_forward.npz"))
    ctx = _ctx(list(synth.CAT_PROMPTS))
    eps = ref_unet(synth.synth_latent(0).double(), 981, ctx[2:3])
    assert _rel(eps, torch.from_numpy(g["a_eps"])) < 1e-6  # fixture stored as float32


def _edit_controller(n_steps):
    tok = synth.FakeTokenizer()
    prompts = list(synth.CAT_PROMPTS)
    mapper, alphas = seq_aligner.get_refinement_mapper(prompts, tok)
    ca = ptp_utils.get_time_words_attention_alpha(prompts, n_steps, {"default_": 0.4}, tok).double()
    eq = get_equalizer(prompts[1], ("watercolor",), (2,), tok).double()
    blend = torch.zeros(2, 1, 1, 1, 1, 77, dtype=torch.float64)
    for i, p in enumerate(prompts):
        blend[i, ..., seq_aligner.get_word_inds(p, "cat", tok)] = 1
    return p2p_ref.EditController(n_steps, ca, 0.6, mapper, alphas.double(), eq, blend)


@pytest.mark.slow
def test_oracle_controller_algebra_matches_reference_controllers(ref_unet):
    g = np.load(os.path.join(GOLD, "unet_forward.npz"))
    ctx = _ctx(list(synth.CAT_PROMPTS))
    lat = torch.cat([synth.synth_latent(0), synth.synth_latent(1)]).double()
    ctrl = _edit_controller(50)
    eps = ref_unet(torch.cat([lat] * 2), 601, ctx, attn_hook=ctrl)
    assert _rel(eps, torch.from_numpy(g["b_eps"])) < 1e-6
    maps = ctrl.attention_store["down_cross"][2:4] + ctrl.attention_store["up_cross"][:3]
    maps = torch.cat([m.reshape(2, -1, 1, 16, 16, 77) for m in maps], dim=1).mean(1).reshape(2, 16, 16, 77)
    assert _rel(maps, torch.from_numpy(g["b_maps_mean"])) < 1e-6


def test_oracle_schedule_identities():
    # SURVEY.md section 4 invariant (b): the first inversion step and the last denoise step are identities
    s = p2p_ref.Schedule(50, "float64")
    x = torch.randn(1, 4, 64, 64, dtype=torch.float64)
    e = torch.randn_like(x)
    assert torch.allclose(s.next_step(e, 0, x), x, atol=1e-12)
    assert torch.allclose(s.prev_step(e, 0, x), x, atol=1e-12)
    assert s.timesteps.tolist() == list(range(980, -1, -20))


def test_oracle_pipeline_matches_reference_pipeline_small():
    """The loops (ddim_loop / offset_calculate / guidance_forward with rectification + LocalBlend), restated, reproduce
    the reference pipeline fixture when driven with a cheap stand-in 'UNet' is not possible (fixture used the real
    UNet), so here we check the loop algebra on the fixture itself: source-branch exactness and offsets."""
    path = os.path.join(GOLD, "pipeline_4steps.npz")
    if not os.path.exists(path):
        pytest.skip("pipeline fixture not generated yet")
    g = np.load(path)
    x_stars, nl, recon, edit = (torch.from_numpy(g[k]) for k in ("x_stars", "noise_loss", "recon", "edit"))
    assert x_stars.shape == (5, 4, 64, 64) and nl.shape == (4, 2, 4, 64, 64)
    # invariant (a): with rectification the source branch lands exactly on x_stars[0] in both passes
    assert _rel(recon[0], x_stars[0]) < 1e-6 and _rel(edit[0], x_stars[0]) < 1e-6
    # offset_calculate feeds both prompt rows the same latents but different cond contexts -> different losses
    assert not torch.equal(nl[:, 0], nl[:, 1])
