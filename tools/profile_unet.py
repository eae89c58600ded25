"""Target for ncu: builds the B=4 plan (incl. GEMM autotuning) and warms up OUTSIDE the profiled range, then runs
`n` UNet forwards between cudaProfilerStart/Stop.  Use with `ncu --profile-from-start off ...`."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from pnpinversion_b200.model import FusedModel

n = int(sys.argv[1]) if len(sys.argv) > 1 else 1
B = int(sys.argv[2]) if len(sys.argv) > 2 else 4
m = FusedModel.synthetic(max_batch=B)
ctx = m.text_encoder(m.tokenizer(["a cat"] * B).input_ids)[0].to("cuda", torch.float32)
x = torch.randn(B, 4, 64, 64, device="cuda")
for _ in range(3):
    m.unet(x, 501, encoder_hidden_states=ctx)
torch.cuda.synchronize()
torch.cuda.profiler.start()
for _ in range(n):
    m.unet(x, 501, encoder_hidden_states=ctx)
torch.cuda.synchronize()
torch.cuda.profiler.stop()
print("profiled", n, "forward(s) at B =", B)
