"""EDICT.  The reference module cannot run here (it loads hub weights onto 'cuda' at import),
so parity rests on (i) the CPU restatement oracle/edict_ref.py for two coupled steps at full UNet size and (ii) EDICT's
size-independent defining property: reverse followed by forward with the same prompt reproduces the input pair.
Tolerances: the reference is fp64; with fp16 operands one UNet call carries ~3e-3 and CFG 7.0 amplifies it, so two steps
are held to 5e-2 on the latent pair; the round trip is held to 2e-2 (the mixing layers contract the error)."""
import pytest
import torch

from oracle import edict_ref, p2p_ref, unet_ref
from pnpinversion_b200 import edict, synth
from pnpinversion_b200.model import FusedModel
from tests import gpu_util as G

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def model(cuda):
    m = FusedModel.synthetic(device="cuda:0", max_batch=4, table_dtype="float64")
    yield m
    m.unet.close()


def test_attention_edit_tables_follow_difflib():
    tok = synth.FakeTokenizer()
    a = tok(synth.CAT_PROMPTS[0]).input_ids[0].tolist()
    b = tok(synth.CAT_PROMPTS[1]).input_ids[0].tolist()
    mask, idx = edict.attention_edit_tables(a, b)
    # "a watercolor of" is inserted after BOS... "a cat ..." shifts by 3; inserted tokens are unmasked
    assert mask[0] == 1 and idx[0] == 0
    assert mask.sum() < 77 and int(idx[5]) == 2


def test_reverse_then_forward_is_the_identity(model):
    z = synth.synth_latent(5)
    prompt = synth.CAT_PROMPTS[0]
    lat = edict.coupled_stablediffusion(model, prompt, reverse=True, init_image=z, steps=10, guidance_scale=3.0)
    assert G.rel_l2(lat[0].cpu(), z) > 0.05  # it really moved
    back = edict.coupled_stablediffusion(model, prompt, reverse=False, fixed_starting_latent=lat, steps=10,
                                         guidance_scale=3.0)
    torch.cuda.synchronize()
    e0, e1 = G.rel_l2(back[0].cpu(), z), G.rel_l2(back[1].cpu(), z)
    print("EDICT round trip rel-L2:", e0, e1)
    assert e0 < 2e-2 and e1 < 2e-2


def test_two_coupled_steps_with_p2p_match_the_oracle(model):
    torch.set_grad_enabled(False)
    src, tgt = synth.CAT_PROMPTS
    z = synth.synth_latent(6)
    steps, strength = 50, 0.04  # t_limit = 48 -> the last two timesteps (20, 0)
    lat = edict.coupled_stablediffusion(model, src, reverse=True, init_image=z, steps=steps,
                                        init_image_strength=strength, guidance_scale=3.0)
    out = edict.coupled_stablediffusion(model, src, tgt, fixed_starting_latent=lat, steps=steps,
                                        init_image_strength=strength, guidance_scale=3.0)
    torch.cuda.synchronize()
    # oracle (fp64, CPU)
    ref_unet = unet_ref.UNetRef(synth.synth_unet_state_dict(0))
    tok, te = synth.FakeTokenizer(), synth.SynthTextEncoder(dtype=torch.float64)
    emb = lambda s: te(tok(s).input_ids)[0]
    mask, idx = edict.attention_edit_tables(tok(src).input_ids[0].tolist(), tok(tgt).input_ids[0].tolist())
    ac = p2p_ref.alphas_cumprod("float64")
    ts = p2p_ref.timesteps(steps)
    t_limit = steps - int(steps * strength)
    zz = z.double()
    lat_ref = edict_ref.coupled(ref_unet, ac, ac[0], ts, [zz, zz.clone()], emb(""), emb(src), guidance=3.0, steps=steps,
                                t_limit=t_limit, reverse=True)
    out_ref = edict_ref.coupled(ref_unet, ac, ac[0], ts, lat_ref, emb(""), emb(src), emb(tgt), mask.double(), idx,
                                guidance=3.0, steps=steps, t_limit=t_limit, reverse=False)
    errs = [G.rel_l2(lat[i].cpu(), lat_ref[i]) for i in range(2)] + [G.rel_l2(out[i].cpu(), out_ref[i]) for i in range(2)]
    print("EDICT vs oracle (reverse pair, forward-P2P pair):", errs)
    assert max(errs) < 5e-2
