"""Early look at the 50-step parity while the reference run is still producing the final fixture: inversion latents and
offsets against gpurun_out/pipeline_50_partial_invert.npz (scratch, not committed)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from pnpinversion_b200 import synth
from pnpinversion_b200.batched import BatchedDirectInversionP2P
from pnpinversion_b200.model import FusedModel

g = np.load(os.path.join(os.path.dirname(__file__), "..", "gpurun_out", "pipeline_50_partial_invert.npz"))
xs_ref, nl_ref = torch.from_numpy(g["x_stars"]), torch.from_numpy(g["noise_loss"])
rel = lambda a, b: float((a.double() - b.double()).norm() / b.double().norm())
for table in ("float64", "float32"):
    m = FusedModel.synthetic(max_batch=4, table_dtype=table)
    b = BatchedDirectInversionP2P(m, 50)
    src, tgt = synth.CAT_PROMPTS
    xs, nl = b.invert(synth.synth_latent(0).cuda(), [src], [tgt], guidance_scale=7.5)
    torch.cuda.synchronize()
    xe = [rel(xs[k, 0].cpu(), xs_ref[k]) for k in range(1, 51)]
    ne = [float((nl[i].cpu() - nl_ref[i]).norm() / xs_ref[50 - i - 1].norm()) for i in range(50)]
    print(f"{table}: x_stars rel-L2 @10/20/30/40/50: " + " ".join(f"{xe[k-1]:.2e}" for k in (10, 20, 30, 40, 50)) +
          f" | noise_loss |diff|/|latent| @0/10/25/49: {ne[0]:.2e} {ne[10]:.2e} {ne[25]:.2e} {ne[49]:.2e} max {max(ne):.2e}"
          f" | |noise_loss|/|latent| ref {float(nl_ref[0].norm() / xs_ref[49].norm()):.2e}", flush=True)
    m.unet.close()
