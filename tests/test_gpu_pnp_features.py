"""Plug-and-Play features (`pnpinversion_b200/pnp_features.py`, descriptor fields self_q/k_row + conv_src_row) against the
REFERENCE's own functions run on the vendored fp64 UNet (tests/golden/pnp_features_4steps.npz, produced by
oracle/make_golden.py pnp from run_editing_pnp.py's ddim_inversion / ddim_sample / register_*_control_efficient /
denoise_step).  Steps 0-1 inject Q/K and conv features, step 2 conv features only, step 3 nothing."""
import os

import numpy as np
import pytest
import torch

from pnpinversion_b200 import synth
from pnpinversion_b200.model import FusedModel
from pnpinversion_b200.pnp_features import PnPController, PnPFeaturesEditor, pnp_timesteps
from tests import gpu_util as G

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(__file__), "golden", "pnp_features_4steps.npz")


def test_pnp_features_match_the_reference_run(cuda):
    if not os.path.exists(GOLD):
        pytest.fail("tests/golden/pnp_features_4steps.npz missing (python -m oracle.make_golden pnp 4)")
    g = np.load(GOLD)
    m = FusedModel.synthetic(device="cuda:0", max_batch=3, table_dtype="float64")
    ed = PnPFeaturesEditor(m, 4)
    assert ed.timesteps == [int(t) for t in g["timesteps"]] == pnp_timesteps(4)  # integer schedule, bit-exact
    src, tgt = synth.CAT_PROMPTS
    z0 = synth.synth_latent(8).cuda()
    inv, rec = ed.extract_latents(z0, [src])
    x = ed.run_pnp(inv, [tgt], guidance_scale=7.5, pnp_f_t=0.8, pnp_attn_t=0.5)
    torch.cuda.synchronize()
    e_inv = [G.rel_l2(inv[k].cpu(), torch.from_numpy(g["inverted_x"][k:k + 1])) for k in range(1, 5)]
    e_rec = [G.rel_l2(rec[k].cpu(), torch.from_numpy(g["rec"][3 - k:4 - k])) for k in range(4)]  # ours is reversed like extract_latents
    e_x = G.rel_l2(x.cpu(), torch.from_numpy(g["xs"][3:4]))
    print(f"pnp features vs reference: inversion {e_inv}, reconstruction {e_rec}, edited latent {e_x:.2e}")
    assert max(e_inv) < 5e-3 and max(e_rec) < 1e-2 and e_x < 8e-2
    m.unet.close()


def test_feature_injection_really_copies_the_source_rows(cuda):
    """Descriptor semantics on one UNet call: with conv_src_row / self_q,k_row pointing at the source row and IDENTICAL
    inputs in all three rows the output rows are identical to the un-injected call (injection of equal features is a
    no-op); with different inputs the injected rows change and the source row does not."""
    m = FusedModel.synthetic(device="cuda:0", max_batch=3)
    tok, te = m.tokenizer, m.text_encoder
    ctx = te(tok(["", "ugly", synth.CAT_PROMPTS[1]]).input_ids)[0].cuda().float().contiguous()
    ctrl = PnPController(1, [981], [981])
    ctrl.t = 981
    x = torch.cat([synth.synth_latent(i) for i in range(3)]).cuda()
    m.unet.set_controller(None)
    first = m.unet(x, 981, encoder_hidden_states=ctx)["sample"].clone()
    plain = m.unet(x, 981, encoder_hidden_states=ctx)["sample"].clone()
    m.unet.set_controller(ctrl)
    inj = m.unet(x, 981, encoder_hidden_states=ctx)["sample"].clone()
    torch.cuda.synchronize()
    print(f"pnp rows: first-vs-replay {G.rel_l2(first, plain):.2e}; source row plain-vs-injected {G.rel_l2(inj[0], plain[0]):.2e}; "
          f"injected rows {G.rel_l2(inj[1], plain[1]):.2e} {G.rel_l2(inj[2], plain[2]):.2e}")
    assert torch.equal(first, plain)
    assert torch.equal(plain[0], inj[0])  # the source row is never touched
    assert G.rel_l2(inj[1], plain[1]) > 1e-2 and G.rel_l2(inj[2], plain[2]) > 1e-2
    ctrl.t = 1  # outside both schedules: identity descriptor
    off = m.unet(x, 981, encoder_hidden_states=ctx)["sample"]
    assert torch.equal(off, plain)
    m.unet.close()
