"""GroupNorm(+SiLU, +skip concat), LayerNorm, upsample, stride-2 im2col vs PyTorch fp32 references."""
import ctypes as C

import pytest
import torch
import torch.nn.functional as F

from pnpinversion_b200 import _lib
from tests import gpu_util as G

pytestmark = pytest.mark.gpu


def _mk(shape, dev, seed, scale=1.0, shift=0.0):
    g = torch.Generator(device="cpu").manual_seed(seed)
    return (torch.randn(shape, generator=g) * scale + shift).to(torch.float16).to(dev)


@pytest.mark.parametrize("B,HW,C0,C1,silu", [(1, 4096, 320, 0, True), (4, 4096, 640, 320, True), (2, 1024, 1280, 640, True),
                                             (4, 64, 1280, 1280, True), (2, 256, 1280, 0, False), (1, 64, 1280, 0, True),
                                             (3, 1024, 640, 0, True)])
def test_groupnorm(cuda, B, HW, C0, C1, silu):
    lib = _lib.load()
    x0 = _mk((B, HW, C0), cuda, 1, 3.0, 0.7)
    x1 = _mk((B, HW, C1), cuda, 2, 2.0, -0.4) if C1 else None
    Cc = C0 + C1
    gamma = torch.randn(Cc, device=cuda) * 0.1 + 1
    beta = torch.randn(Cc, device=cuda) * 0.1
    out = torch.empty(B, HW, Cc, dtype=torch.float16, device=cuda)
    eps = 1e-5 if silu else 1e-6
    _lib.check(lib.pnp_test_groupnorm(G.ptr(x0), C0, G.ptr(x1), C1, B, HW, G.ptr(gamma), G.ptr(beta), eps,
                                      1 if silu else 0, G.ptr(out), G.stream()))
    torch.cuda.synchronize()
    x = x0 if x1 is None else torch.cat([x0, x1], dim=-1)
    ref = F.group_norm(x.float().permute(0, 2, 1), 32, gamma, beta, eps).permute(0, 2, 1)
    if silu:
        ref = F.silu(ref)
    assert (out.float() - ref).abs().max() < 4e-3
    assert G.rel_l2(out, ref) < 6e-4


@pytest.mark.parametrize("B,HW,C0,C1,silu", [(8, 256, 1280, 0, True), (32, 256, 1280, 1280, True), (16, 256, 1280, 640, True),
                                             (8, 64, 1280, 0, False), (32, 64, 1280, 1280, True), (16, 1024, 640, 0, True),
                                             (8, 1024, 640, 640, True), (32, 1024, 640, 0, False), (6, 256, 1280, 0, True)])
def test_gn_register_resident_path(cuda, B, HW, C0, C1, silu):
    """Image batches: the low-resolution levels take the register-resident kernel (one CTA per image x chunk of whole groups,
    one read, centred variance from registers); it must actually be the path taken and match torch's fp32 GroupNorm."""
    lib = _lib.load()
    assert lib.pnp_test_groupnorm_path(C0 + C1, B, HW) == 2
    assert lib.pnp_test_groupnorm_path(C0 + C1, 2, HW) != 2  # too few CTAs at small batch: cluster / two-kernel paths
    test_groupnorm(cuda, B, HW, C0, C1, silu)


def test_groupnorm_cluster_kernel_path(cuda):
    """By default the single-launch cluster kernel (statistics exchanged through distributed shared memory) takes the small
    tensors (HW <= 256) and the statistics + apply pair the large ones; PNP_GN_CLUSTER=16 forces the cluster kernel for
    every shape, PNP_GN_CLUSTER=0 the pair.  Both must be correct everywhere.  The choice is made once per process, hence
    the subprocesses."""
    import os
    import subprocess
    import sys

    for mode in ("16", "0"):
        env = dict(os.environ, PNP_GN_CLUSTER=mode)
        r = subprocess.run([sys.executable, "-m", "pytest", "-q", "-m", "gpu", "-p", "no:cacheprovider", __file__, "-k",
                            "test_groupnorm and not cluster_kernel"], env=env, capture_output=True, text=True, timeout=600,
                           cwd=os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
        assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
        assert "7 passed" in r.stdout, r.stdout[-500:]


@pytest.mark.parametrize("rows,Cc", [(4096, 320), (1000, 640), (77, 1280)])
def test_layernorm(cuda, rows, Cc):
    lib = _lib.load()
    x = _mk((rows, Cc), cuda, 3, 4.0, 1.0)
    gamma = torch.randn(Cc, device=cuda) * 0.1 + 1
    beta = torch.randn(Cc, device=cuda) * 0.1
    out = torch.empty_like(x)
    _lib.check(lib.pnp_test_layernorm(G.ptr(x), rows, Cc, G.ptr(gamma), G.ptr(beta), 1e-5, G.ptr(out), G.stream()))
    torch.cuda.synchronize()
    ref = F.layer_norm(x.float(), (Cc,), gamma, beta, 1e-5)
    assert G.rel_l2(out, ref) < 6e-4


def test_upsample_and_im2col_are_exact(cuda):
    lib = _lib.load()
    B, H, W, Cc = 2, 16, 16, 64
    x = _mk((B, H, W, Cc), cuda, 4)
    up = torch.empty(B, 2 * H, 2 * W, Cc, dtype=torch.float16, device=cuda)
    _lib.check(lib.pnp_test_upsample2x(G.ptr(x), B, H, W, Cc, G.ptr(up), G.stream()))
    ref = F.interpolate(x.permute(0, 3, 1, 2).float(), scale_factor=2.0, mode="nearest").permute(0, 2, 3, 1)
    torch.cuda.synchronize()
    assert torch.equal(up.float(), ref)
    col = torch.empty(B, H // 2, W // 2, 9 * Cc, dtype=torch.float16, device=cuda)
    _lib.check(lib.pnp_test_im2col_s2(G.ptr(x), B, H, W, Cc, G.ptr(col), G.stream()))
    torch.cuda.synchronize()
    # stride-2 pad-1 conv through the im2col matrix == F.conv2d(stride=2, padding=1)
    w = _mk((32, Cc, 3, 3), cuda, 5, (9 * Cc) ** -0.5)
    got = col.float().reshape(-1, 9 * Cc) @ G.pack_conv3(w).float().t()
    refc = F.conv2d(x.permute(0, 3, 1, 2).float(), w.float(), None, stride=2, padding=1).permute(0, 2, 3, 1)
    assert G.rel_l2(got.reshape(refc.shape), refc) < 1e-5
