"""Where do the GEMM's three roles wait?  Runs representative UNet shapes with PNP_GEMM_PROF=1 (in-kernel cycle counters,
printed by the C test entry points) in both tile modes."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
os.environ["PNP_GEMM_PROF"] = "1"
import torch
from tests import gpu_util as G


def mk(shape, seed, scale=1.0):
    g = torch.Generator(device="cpu").manual_seed(seed)
    return (torch.randn(shape, generator=g) * scale).to(torch.float16).cuda()


convs = [(4, 64, 320, 320, 0), (4, 64, 320, 320, 160), (4, 64, 320, 320, 320), (4, 32, 640, 640, 0), (4, 32, 640, 640, 160),
         (4, 16, 1280, 1280, 0), (4, 16, 1280, 1280, 320), (4, 8, 1280, 1280, 0), (4, 8, 1280, 1280, 320)]
lins = [(16384, 320, 320, 0), (16384, 320, 320, 160), (16384, 320, 960, 0), (16384, 320, 960, 320), (16384, 1280, 320, 0),
        (4096, 640, 640, 0), (4096, 640, 1920, 0), (4096, 640, 1920, 320), (1024, 1280, 1280, 0), (1024, 1280, 3840, 0),
        (308, 768, 2560, 0)]
for mode in ("0",):
    os.environ["PNP_GEMM_CLUSTER"] = mode
    print("==== PNP_GEMM_CLUSTER=" + mode, file=sys.stderr, flush=True)
    for (B, H, C, N, bn) in convs:
        x = mk((B, H, H, C), 1)
        w = mk((N, 9 * C), 2, (9 * C) ** -0.5)
        for _ in range(2):
            G.conv3x3(x, w, bn=bn, split=0)
    for (M, K, N, bn) in lins:
        a = mk((M, K), 3)
        w = mk((N, K), 4, K ** -0.5)
        for _ in range(2):
            G.gemm(a, w, bn=bn, split=0)
