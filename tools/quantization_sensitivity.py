"""ANALYSIS TOOL (runs the CPU oracle; never part of the product).  How does one UNet forward respond to an infinitesimal
input perturbation (a) in exact arithmetic, (b) with fp16 rounding at the places the fused engine rounds (tools/
error_budget.py)?  Answer (profiles/r2_quantization_sensitivity.txt): the exact network has a gain of ~3.5; the rounded
pipeline answers a 1e-9 perturbation with a 3.5e-3 change - the full size of its own error against the fp64 reference.
A rounding step turns a perturbation d into ~sqrt(d * q) (q = the fp16 step, the few values that cross a rounding
boundary move by a whole q), so after a handful of the ~200 rounding points the rounding-noise REALISATION is independent
of the unperturbed run.  Consequences: (1) the same row computed in two batch sizes (other GEMM tilings / split-K orders,
1e-7 differences in fp32) differs by ~3.7e-3 although each is 3.3e-3 from the reference (tools/diag_batch.py);
(2) EDICT's exact invertibility cannot survive 16-bit activation storage; (3) the 3.3e-3 is noise of fixed size, not a
bias one could calibrate away."""
import sys, torch, time
import os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'tools'))
from error_budget import UNetEmu, CLASSES
from pnpinversion_b200 import synth
sd = synth.synth_unet_state_dict(0)
tok, te = synth.FakeTokenizer(), synth.SynthTextEncoder()
ctx = te(tok([synth.CAT_PROMPTS[0]]).input_ids)[0]
x = synth.synth_latent(0)
g = torch.Generator().manual_seed(1)
rel = lambda a, b: float((a.double() - b.double()).norm() / b.double().norm())
with torch.no_grad():
    for name, rnd in (("exact", {}), ("all_fp16", {c: "fp16" for c in CLASSES})):
        base = UNetEmu(sd, torch.float64, rnd)(x, 981, ctx)
        for eps in (1e-9, 1e-6):
            xp = x.double() * (1 + eps * torch.randn(x.shape, generator=g, dtype=torch.float64))
            out = UNetEmu(sd, torch.float64, rnd)(xp, 981, ctx)
            print(f"{name}: input perturbed by {eps:g} relative -> output changes by {rel(out, base):.3e}", flush=True)
