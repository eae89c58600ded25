"""Diagnostic: does the result of a UNet row depend on the batch size it travels in?  Case a of tests/golden/unet_forward.npz
(B=1, t=981, source prompt, latent 0; reference fp64) replicated B times."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from pnpinversion_b200 import synth
from pnpinversion_b200.model import FusedModel

gold = torch.from_numpy(np.load(os.path.join(os.path.dirname(__file__), "..", "tests", "golden", "unet_forward.npz"))["a_eps"])
m = FusedModel.synthetic(max_batch=16)
tok, te = m.tokenizer, m.text_encoder
ctx1 = te(tok([synth.CAT_PROMPTS[0]]).input_ids)[0].cuda().float()
x1 = synth.synth_latent(0).cuda()
rel = lambda a, b: float((a.double() - b.double()).norm() / b.double().norm())
outs = {}
for B in (1, 2, 3, 4, 8, 12, 16):
    x = x1.expand(B, -1, -1, -1).contiguous()
    ctx = ctx1.expand(B, -1, -1).contiguous()
    o = m.unet(x, 981, encoder_hidden_states=ctx)["sample"].cpu()
    outs[B] = o
    rows = [rel(o[i:i + 1], gold) for i in range(B)]
    same = all(torch.equal(o[0], o[i]) for i in range(B))
    print(f"B={B:2d}: rel-L2 vs reference fp64 rows min {min(rows):.3e} max {max(rows):.3e}; rows bit-identical: {same}; "
          f"row0 vs B=1 result: {rel(o[:1], outs[1]):.3e}", flush=True)
