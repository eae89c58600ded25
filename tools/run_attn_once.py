"""One launch of the tcgen05 self-attention at the UNet's largest shape (B=4, 4096 tokens) -- target for ncu."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from pnpinversion_b200 import _lib
from tests import gpu_util as G

lib = _lib.load()
B, N, H, d = 4, 4096, 8, 40
g = torch.Generator(device="cpu").manual_seed(5)
qkv = (torch.randn(B, N, 3 * H * d, generator=g)).to(torch.float16).cuda()
out = torch.zeros(B, N, H * d, dtype=torch.float16, device="cuda")
for _ in range(3):
    _lib.check(lib.pnp_test_self_attention_tc(G.ptr(qkv), B, N, None, None, None, G.ptr(out), G.stream()))
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(10):
    _lib.check(lib.pnp_test_self_attention_tc(G.ptr(qkv), B, N, None, None, None, G.ptr(out), G.stream()))
e1.record()
torch.cuda.synchronize()
print("tc attention B=4 N=4096: %.1f us per call (incl. V transpose + plan)" % (e0.elapsed_time(e1) * 100))
