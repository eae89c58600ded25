"""GEMM experiments: which side paces the k-block loop?  PNP_GEMM_EXP=1 removes the TMA copies, =2 removes the MMAs."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
os.environ["PNP_GEMM_PROF"] = "1"
os.environ["PNP_GEMM_CLUSTER"] = "0"
import torch
from tests import gpu_util as G


def mk(shape, seed, scale=1.0):
    g = torch.Generator(device="cpu").manual_seed(seed)
    return (torch.randn(shape, generator=g) * scale).to(torch.float16).cuda()


convs = [(4, 64, 320, 320, 160), (4, 64, 320, 320, 64), (4, 16, 1280, 1280, 256), (4, 16, 1280, 1280, 128)]
for exp in ("0", "1", "2"):
    os.environ["PNP_GEMM_EXP"] = exp
    print("==== PNP_GEMM_EXP=" + exp, file=sys.stderr, flush=True)
    for (B, H, C, N, bn) in convs:
        x = mk((B, H, H, C), 1)
        w = mk((N, 9 * C), 2, (9 * C) ** -0.5)
        G.conv3x3(x, w, bn=bn, split=1)
    a = mk((16384, 1280), 3)
    w = mk((320, 1280), 4, 1280 ** -0.5)
    G.gemm(a, w, bn=160, split=1)
