"""Diagnostic: is the source row of a Plug-and-Play call bit-identical to the plain call?  (eager first call vs replays)"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from pnpinversion_b200 import synth
from pnpinversion_b200.model import FusedModel
from pnpinversion_b200.pnp_features import PnPController

rel = lambda a, b: float((a.double() - b.double()).norm() / b.double().norm())
for B in (3, 6):
    L = B // 3
    m = FusedModel.synthetic(device="cuda:0", max_batch=B)
    tok, te = m.tokenizer, m.text_encoder
    ctx = te(tok([""] * L + ["ugly"] * L + [synth.CAT_PROMPTS[1]] * L).input_ids)[0].cuda().float().contiguous()
    x = torch.cat([synth.synth_latent(i) for i in range(B)]).cuda()
    m.unet.set_controller(None)
    p1 = m.unet(x, 981, encoder_hidden_states=ctx)["sample"].clone()
    p2 = m.unet(x, 981, encoder_hidden_states=ctx)["sample"].clone()
    p3 = m.unet(x, 981, encoder_hidden_states=ctx)["sample"].clone()
    print(f"B={B}: plain eager vs replay equal: {torch.equal(p1, p2)} ({rel(p1, p2):.2e}); replay vs replay equal: {torch.equal(p2, p3)}")
    for name, qk, conv in (("qk only", [981], []), ("conv only", [], [981]), ("both", [981], [981])):
        ctrl = PnPController(L, qk, conv)
        ctrl.t = 981
        m.unet.set_controller(ctrl)
        i1 = m.unet(x, 981, encoder_hidden_states=ctx)["sample"].clone()
        i2 = m.unet(x, 981, encoder_hidden_states=ctx)["sample"].clone()
        print(f"   {name}: run-to-run equal {torch.equal(i1, i2)}; source rows vs plain: "
              + " ".join(f"{rel(i1[r], p2[r]):.2e}" for r in range(L)) + " | injected rows vs plain: "
              + " ".join(f"{rel(i1[r], p2[r]):.2e}" for r in range(L, B)))
    m.unet.close()
