"""Quick device-side timing of the fused UNet forward (CUDA events), used while optimising."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from pnpinversion_b200 import synth
from pnpinversion_b200.model import FusedModel

iters = int(sys.argv[1]) if len(sys.argv) > 1 else 20
batches = [int(b) for b in sys.argv[2].split(",")] if len(sys.argv) > 2 else [1, 4]
t0 = time.time()
m = FusedModel.synthetic(max_batch=max(batches))
print("model ready in %.1fs" % (time.time() - t0), flush=True)
tok, te = m.tokenizer, m.text_encoder
for B in batches:
    ctx = te(tok(["a cat"] * B).input_ids)[0].to("cuda", torch.float32)
    x = torch.randn(B, 4, 64, 64, device="cuda")
    for _ in range(3):
        m.unet(x, 501, encoder_hidden_states=ctx)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        m.unet(x, 501, encoder_hidden_states=ctx)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / iters
    tf = 803.27e9 * B / (ms * 1e-3) / 1e12
    print(f"B={B}: {ms:.3f} ms / forward  -> {tf:.1f} TFLOP/s algorithmic", flush=True)
