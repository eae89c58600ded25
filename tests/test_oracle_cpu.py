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


# ---------------------------------------------------------------------------------------------------- EDICT (row a15)
EDICT_GOLD = os.path.join(GOLD, "edict_2steps.npz")


def _edict_embeddings():
    tok, te = synth.FakeTokenizer(), synth.SynthTextEncoder(dtype=torch.float64)
    src, tgt = synth.CAT_PROMPTS
    return [te(tok(s).input_ids)[0] for s in ("", src, tgt)], tok


def test_oracle_edict_loop_replays_the_reference_call_sequence():
    """tests/golden/edict_2steps.npz was produced by the reference's own `coupled_stablediffusion`
    (models/edict/edict_functions.py:707-956, run unmodified through oracle/ref_shim.load_reference_edict): two coupled
    steps of deterministic noising, then two of Prompt-to-Prompt generation, every UNet call recorded.  Driving the
    restatement with a 'UNet' that replays the recorded predictions checks - without running a UNet - that it makes the
    same calls in the same order (timestep, text embedding, input latent: un-mixing, leapfrog order) and lands on the same
    latents (forward_step / reverse_step / mixing layers)."""
    from oracle import edict_ref

    g = np.load(EDICT_GOLD)
    (emb_u, emb_c, emb_e), _ = _edict_embeddings()
    kind = {id(emb_u): 0, id(emb_c): 1, id(emb_e): 2}
    eps = torch.from_numpy(g["call_eps"]).double()
    pos = [0]

    def replay(x, t, ctx, hook):
        i = pos[0]
        pos[0] += 1
        assert int(t) == int(g["call_t"][i]) and kind[id(ctx)] == int(g["call_ctx"][i]), (i, int(t), kind[id(ctx)])
        assert abs(float(x.sum()) - float(g["call_in_sum"][i])) <= 1e-5 * float(g["call_in_abs"][i]), i
        assert abs(float(x.abs().sum()) - float(g["call_in_abs"][i])) <= 1e-5 * float(g["call_in_abs"][i]), i
        return eps[i:i + 1]

    ac, ts = p2p_ref.alphas_cumprod("float64"), p2p_ref.timesteps(50)
    z = torch.from_numpy(g["z"]).double()
    lat = edict_ref.coupled(replay, ac, ac[0], ts, [z, z.clone()], emb_u, emb_c, guidance=3.0, steps=50, t_limit=48,
                            reverse=True)
    assert pos[0] == int(g["n_reverse_calls"]) == 8
    for i in range(2):
        assert _rel(lat[i], torch.from_numpy(g["lat"][i])) < 1e-6  # call_eps is stored as float32
    mask, idx = torch.from_numpy(g["edit_mask"]).double(), torch.from_numpy(g["edit_indices"])
    out = edict_ref.coupled(replay, ac, ac[0], ts, [torch.from_numpy(g["lat"][i]) for i in range(2)], emb_u, emb_c, emb_e,
                            mask, idx, guidance=3.0, steps=50, t_limit=48, reverse=False)
    assert pos[0] == len(g["call_t"]) == 20
    for i in range(2):
        assert _rel(out[i], torch.from_numpy(g["out"][i])) < 1e-6


def test_edict_attention_edit_tables_match_the_reference():
    """`init_attention_edit` (edict_functions.py:225-247) as left on the reference's CrossAttention modules."""
    from pnpinversion_b200 import edict

    g = np.load(EDICT_GOLD)
    tok = synth.FakeTokenizer()
    src, tgt = synth.CAT_PROMPTS
    mask, idx = edict.attention_edit_tables(tok(src).input_ids[0].tolist(), tok(tgt).input_ids[0].tolist())
    assert torch.equal(mask, torch.from_numpy(g["edit_mask"]).to(mask.dtype))
    assert torch.equal(idx, torch.from_numpy(g["edit_indices"]))


@pytest.mark.slow
def test_oracle_edict_attention_reuse_matches_the_reference(ref_unet):
    """First generation sub-step of the fixture (t = 20, input = the second latent of the noised pair): the conditional
    pass saves every attention map, the edited pass reuses them (self-attention wholesale, cross-attention through mask /
    indices) - edict_functions.py:250-297, 893-917.  Two full-size fp64 UNet forwards of the restatement."""
    from oracle import edict_ref

    g = np.load(EDICT_GOLD)
    (emb_u, emb_c, emb_e), _ = _edict_embeddings()
    n = int(g["n_reverse_calls"])
    assert [int(g["call_ctx"][n + k]) for k in range(3)] == [0, 1, 2] and int(g["call_t"][n]) == 20
    x = torch.from_numpy(g["lat"][1])
    assert abs(float(x.sum()) - float(g["call_in_sum"][n])) <= 1e-9 * float(g["call_in_abs"][n])
    hook = edict_ref._Reuse(torch.from_numpy(g["edit_mask"]).double(), torch.from_numpy(g["edit_indices"]))
    hook.mode = "save"
    e_c = ref_unet(x, 20, emb_c, hook)
    hook.mode = "use"
    e_e = ref_unet(x, 20, emb_e, hook)
    assert _rel(e_c, torch.from_numpy(g["call_eps"][n + 1:n + 2])) < 1e-6
    assert _rel(e_e, torch.from_numpy(g["call_eps"][n + 2:n + 3])) < 1e-6


# ---------------------------------------------------------------------------------------------------- VAE (row a16, next)
def test_oracle_vae_matches_the_reference_autoencoder():
    """oracle/vae_ref.py against the vendored AutoencoderKL (tests/golden/vae_small.npz, oracle/make_golden.py vae): the
    posterior moments of `image2latent` and the `latent2image` decode (utils/utils.py:58-80).  The CUDA VAE is a 'next' row
    (SURVEY.md section 8f); this pins the gate it will have to pass."""
    from oracle import vae_ref

    g = np.load(os.path.join(GOLD, "vae_small.npz"))
    torch.set_grad_enabled(False)
    vae = vae_ref.VaeRef(synth.synth_vae_state_dict(0))
    mean, logvar = vae.encode_moments(torch.from_numpy(g["img"]).double())
    assert _rel(mean, torch.from_numpy(g["mean"])) < 1e-9 and _rel(logvar, torch.from_numpy(g["logvar"])) < 1e-9
    assert _rel(vae.decode(torch.from_numpy(g["z"]).double()), torch.from_numpy(g["dec"])) < 1e-9


def test_vae_parameter_table_has_the_reference_layout():
    from pnpinversion_b200 import arch

    specs = arch.vae_param_specs()
    assert len(specs) == 248 and sum(int(np.prod(s)) for _, s in specs) == 83_653_863  # 83.65 M parameters
    sd = synth.synth_vae_state_dict(0)
    assert list(sd) == [k for k, _ in specs] and all(tuple(sd[k].shape) == s for k, s in specs)


def test_oracle_clip_matches_transformers():
    """oracle/clip_ref.py against the transformers CLIPTextModel the reference calls (tests/golden/clip_text.npz, made by
    oracle/make_golden.py clip from the installed transformers) and, where transformers is importable, against a live model
    on other token ids."""
    from oracle import clip_ref

    g = np.load(os.path.join(GOLD, "clip_text.npz"))
    sd = synth.synth_clip_state_dict(0)
    ref = clip_ref.ClipTextRef(sd)
    out = ref(torch.from_numpy(g["ids"]).long())
    assert _rel(out, torch.from_numpy(g["out"])) < 1e-6  # fixture stored as float32
    try:
        from transformers import CLIPTextConfig, CLIPTextModel
    except Exception:
        return
    cfg = CLIPTextConfig(vocab_size=49408, hidden_size=768, intermediate_size=3072, num_hidden_layers=12,
                         num_attention_heads=12, max_position_embeddings=77, hidden_act="quick_gelu", layer_norm_eps=1e-5)
    m = CLIPTextModel(cfg).double().eval()
    m.load_state_dict({k: v.double() for k, v in sd.items()}, strict=False)
    ids = synth.PieceTokenizer()(["a photo of a dog", "watercolor painting of mountains at dusk"]).input_ids
    with torch.no_grad():
        live = m(ids)[0]
    assert _rel(ref(ids), live) < 1e-12


def test_clip_parameter_table_matches_transformers():
    from pnpinversion_b200.clip import clip_text_param_specs, count_layers

    specs = clip_text_param_specs()
    assert len(specs) == 2 + 12 * 16 + 2 and sum(int(np.prod(s)) for _, s in specs) == 123060480
    sd = synth.synth_clip_state_dict(0, layers=2, vocab=100)
    assert count_layers(sd) == 2 and sd["text_model.embeddings.token_embedding.weight"].shape == (100, 768)
    try:
        from transformers import CLIPTextConfig, CLIPTextModel
    except Exception:
        return
    cfg = CLIPTextConfig(vocab_size=100, hidden_size=768, intermediate_size=3072, num_hidden_layers=2,
                         num_attention_heads=12, max_position_embeddings=77, hidden_act="quick_gelu")
    names = {k: tuple(v.shape) for k, v in CLIPTextModel(cfg).state_dict().items() if "position_ids" not in k}
    assert names == dict(clip_text_param_specs(2, 100))
