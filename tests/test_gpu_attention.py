"""Attention kernels vs the materialised-softmax algebra of the reference (attention_control.py:20-47,269-282;
masactrl.py:41-72), evaluated in fp32 with PyTorch on the same fp16 inputs."""
import ctypes as C

import pytest
import torch

from pnpinversion_b200 import _lib
from tests import gpu_util as G

pytestmark = pytest.mark.gpu
H = 8


def _mk(shape, dev, seed, scale=1.0):
    g = torch.Generator(device="cpu").manual_seed(seed)
    return (torch.randn(shape, generator=g) * scale).to(torch.float16).to(dev)


def _heads(t, d):  # (B,N,H*d) -> (B,H,N,d)
    B, N, _ = t.shape
    return t.reshape(B, N, H, d).permute(0, 2, 1, 3).float()


def _self_ref(qkv, d, q_row, k_row, v_row):
    c = H * d
    q, k, v = _heads(qkv[..., :c], d), _heads(qkv[..., c:2 * c], d), _heads(qkv[..., 2 * c:], d)
    q, k, v = q[q_row], k[k_row], v[v_row]
    p = (q @ k.transpose(-1, -2) * d ** -0.5).softmax(-1)
    o = p @ v
    return o.permute(0, 2, 1, 3).reshape(qkv.shape[0], qkv.shape[1], c)


@pytest.mark.parametrize("d,N", [(40, 4096), (80, 1024), (160, 256), (160, 64), (40, 64)])
def test_self_attention_plain(cuda, d, N):
    lib = _lib.load()
    B = 2
    qkv = _mk((B, N, 3 * H * d), cuda, d + N, 1.0)
    qkv[..., :2 * H * d] *= 1.5  # logits std ~ 2-3
    out = torch.empty(B, N, H * d, dtype=torch.float16, device=cuda)
    _lib.check(lib.pnp_test_self_attention(G.ptr(qkv), B, H, N, d, None, None, None, G.ptr(out), G.stream()))
    torch.cuda.synchronize()
    ident = list(range(B))
    ref = _self_ref(qkv, d, ident, ident, ident)
    assert G.rel_l2(out, ref) < 2e-3


@pytest.mark.parametrize("mode", ["p2p_self_replace", "masactrl"])
def test_self_attention_row_indirection(cuda, mode):
    lib = _lib.load()
    B, N, d = 4, 1024, 80
    qkv = _mk((B, N, 3 * H * d), cuda, 7, 1.2)
    ident = list(range(B))
    if mode == "p2p_self_replace":  # cond target (row 3) uses the cond source's (row 2) Q and K, own V
        q_row, k_row, v_row = [0, 1, 2, 2], [0, 1, 2, 2], ident
    else:  # both rows of each CFG half attend to the source image's K,V
        q_row, k_row, v_row = ident, [0, 0, 2, 2], [0, 0, 2, 2]
    dq, dk, dv = (torch.tensor(r, dtype=torch.int32, device=cuda) for r in (q_row, k_row, v_row))
    out = torch.empty(B, N, H * d, dtype=torch.float16, device=cuda)
    _lib.check(lib.pnp_test_self_attention(G.ptr(qkv), B, H, N, d, G.ptr(dq), G.ptr(dk), G.ptr(dv), G.ptr(out),
                                           G.stream()))
    torch.cuda.synchronize()
    ref = _self_ref(qkv, d, q_row, k_row, v_row)
    assert G.rel_l2(out, ref) < 2e-3
    # known-answer invariant: rows whose indirection is the identity are untouched
    plain = torch.empty_like(out)
    _lib.check(lib.pnp_test_self_attention(G.ptr(qkv), B, H, N, d, None, None, None, G.ptr(plain), G.stream()))
    torch.cuda.synchronize()
    assert torch.equal(out[0], plain[0]) and torch.equal(out[2], plain[2])


def _cross_ref(q, kv, d, ctrl=None):
    """attention_control.py:34-45 + :269-282 (Refine :319-323, Reweight :340-345) on materialised probabilities."""
    c = H * d
    qh, kh, vh = _heads(q, d), _heads(kv[..., :c], d), _heads(kv[..., c:], d)
    p = (qh @ kh.transpose(-1, -2) * d ** -0.5).softmax(-1)  # (B,H,N,77)
    store = {}
    if ctrl is not None:
        p = p.clone()
        for r, (base, mapper, alphas, eq, ca) in ctrl["edit"].items():
            src = p[base][:, :, mapper]
            refine = src * alphas + p[r] * (1 - alphas)
            rew = refine * eq
            p[r] = rew * ca + (1 - ca) * p[r]
        for r, slot in ctrl.get("store", {}).items():
            store[slot] = p[r].clone()
    o = p @ vh
    return o.permute(0, 2, 1, 3).reshape(q.shape[0], q.shape[1], c), store


@pytest.mark.parametrize("d,N", [(40, 4096), (80, 1024), (160, 256), (160, 64)])
def test_cross_attention_plain(cuda, d, N):
    lib = _lib.load()
    B = 2
    q = _mk((B, N, H * d), cuda, 11, 1.5)
    kv = _mk((B, 77, 2 * H * d), cuda, 12, 1.5)
    out = torch.empty(B, N, H * d, dtype=torch.float16, device=cuda)
    _lib.check(lib.pnp_test_cross_attention(G.ptr(q), G.ptr(kv), B, H, N, d, 77, None, None, G.ptr(out), G.stream()))
    torch.cuda.synchronize()
    ref, _ = _cross_ref(q, kv, d)
    assert G.rel_l2(out, ref) < 2e-3


@pytest.mark.parametrize("d,N", [(160, 256), (80, 1024)])
def test_cross_attention_p2p_injection_and_store(cuda, d, N):
    lib = _lib.load()
    B = 4
    q = _mk((B, N, H * d), cuda, 13, 1.5)
    kv = _mk((B, 77, 2 * H * d), cuda, 14, 1.5)
    g = torch.Generator().manual_seed(5)
    mapper = torch.arange(77)
    mapper[3:40] = torch.arange(2, 39)  # an "inserted word" shift like get_refinement_mapper produces
    mapper[2] = -1
    alphas = torch.ones(77)
    alphas[2] = 0.0
    eq = torch.ones(77)
    eq[5] = 2.0
    ca = (torch.rand(77, generator=g) > 0.3).float()
    ctrl = _lib.new_ctrl()
    ctrl.cross_base_row[3] = 2
    ctrl.cross_slot[3] = 0
    for i in range(77):
        ctrl.mapper[0][i] = int(mapper[i])
        ctrl.alphas[0][i] = float(alphas[i])
        ctrl.equalizer[0][i] = float(eq[i])
        ctrl.cross_alpha[0][i] = float(ca[i])
    ctrl.store_slot[2] = 0
    ctrl.store_slot[3] = 1
    store = torch.zeros(2, H, N, 77, device=cuda)
    out = torch.empty(B, N, H * d, dtype=torch.float16, device=cuda)
    _lib.check(lib.pnp_test_cross_attention(G.ptr(q), G.ptr(kv), B, H, N, d, 77, C.byref(ctrl), G.ptr(store),
                                            G.ptr(out), G.stream()))
    torch.cuda.synchronize()
    dev = lambda t: t.to(cuda)
    ref, rstore = _cross_ref(q, kv, d, {"edit": {3: (2, dev(mapper), dev(alphas), dev(eq), dev(ca))},
                                        "store": {2: 0, 3: 1}})
    assert G.rel_l2(out, ref) < 2e-3
    assert G.rel_l2(store[0], rstore[0]) < 1e-4 and G.rel_l2(store[1], rstore[1]) < 1e-4
    # rows without a controller entry are bit-identical to the plain kernel
    plain = torch.empty_like(out)
    _lib.check(lib.pnp_test_cross_attention(G.ptr(q), G.ptr(kv), B, H, N, d, 77, None, None, G.ptr(plain), G.stream()))
    torch.cuda.synchronize()
    assert torch.equal(out[:3], plain[:3])


@pytest.mark.parametrize("N,B", [(128, 1), (256, 2), (4096, 2)])
def test_self_attention_tcgen05_d40(cuda, N, B):
    """tcgen05 flash attention (attention_tc.cu): two-pass softmax on the tensor core, 8 heads of dim 40."""
    lib = _lib.load()
    d = 40
    qkv = _mk((B, N, 3 * H * d), cuda, 31 + N, 1.0)
    qkv[..., :2 * H * d] *= 1.5
    out = torch.zeros(B, N, H * d, dtype=torch.float16, device=cuda)
    _lib.check(lib.pnp_test_self_attention_tc(G.ptr(qkv), B, N, None, None, None, G.ptr(out), G.stream()))
    torch.cuda.synchronize()
    ident = list(range(B))
    ref = _self_ref(qkv, d, ident, ident, ident)
    err = G.rel_l2(out, ref)
    print(f"tc attention N={N} B={B}: rel-L2 {err:.3e}")
    assert err < 2e-3


def test_self_attention_tcgen05_cluster_multicast(cuda, monkeypatch):
    """Opt-in variant: two query tiles of one (batch, head) form a cluster and share the key / value tiles through TMA
    multicast (PNP_ATTN_CLUSTER=2, read when the plan is made)."""
    monkeypatch.setenv("PNP_ATTN_CLUSTER", "2")
    lib = _lib.load()
    B, N, d = 2, 1024, 40
    qkv = _mk((B, N, 3 * H * d), cuda, 55, 1.1)
    out = torch.zeros(B, N, H * d, dtype=torch.float16, device=cuda)
    _lib.check(lib.pnp_test_self_attention_tc(G.ptr(qkv), B, N, None, None, None, G.ptr(out), G.stream()))
    torch.cuda.synchronize()
    ident = list(range(B))
    assert G.rel_l2(out, _self_ref(qkv, d, ident, ident, ident)) < 2e-3


def test_self_attention_tcgen05_row_indirection(cuda):
    lib = _lib.load()
    B, N, d = 4, 1024, 40
    qkv = _mk((B, N, 3 * H * d), cuda, 77, 1.2)
    q_row, k_row, v_row = [0, 1, 2, 2], [0, 0, 2, 2], [0, 0, 2, 3]
    dq, dk, dv = (torch.tensor(r, dtype=torch.int32, device=cuda) for r in (q_row, k_row, v_row))
    out = torch.zeros(B, N, H * d, dtype=torch.float16, device=cuda)
    _lib.check(lib.pnp_test_self_attention_tc(G.ptr(qkv), B, N, G.ptr(dq), G.ptr(dk), G.ptr(dv), G.ptr(out), G.stream()))
    torch.cuda.synchronize()
    assert G.rel_l2(out, _self_ref(qkv, d, q_row, k_row, v_row)) < 2e-3


def test_self_attention_tcgen05_optimistic_pass_falls_back_when_fp16_would_overflow(cuda):
    """The kernel first runs a single pass with the softmax offset of the first key tile; here later keys have far larger
    scores than the first 128, so the probabilities 2^(s - m_first) would overflow fp16 and the exact two-pass schedule
    must take over (attention_tc.cu, attempt loop).  The result must still match the reference."""
    lib = _lib.load()
    B, N, d = 2, 1024, 40
    qkv = _mk((B, N, 3 * H * d), cuda, 91, 1.0)
    c = H * d
    qkv[:, 256:, c:2 * c] *= 12.0  # keys beyond the second tile: logits ~12x larger
    out = torch.zeros(B, N, H * d, dtype=torch.float16, device=cuda)
    _lib.check(lib.pnp_test_self_attention_tc(G.ptr(qkv), B, N, None, None, None, G.ptr(out), G.stream()))
    torch.cuda.synchronize()
    ident = list(range(B))
    ref = _self_ref(qkv, d, ident, ident, ident)
    assert torch.isfinite(out.float()).all()
    err = G.rel_l2(out, ref)
    print("tc attention with forced fallback: rel-L2", err)
    assert err < 3e-3


# ---- pair mode (PNP_ATTN_CLUSTER=3: two CTAs share one stream of tcgen05.mma.cta_group::2 instructions) and the FMA-pipe
# exponentials (PNP_ATTN_POLY): same tolerances as the default variant; with the polynomial off the arithmetic is the same
@pytest.mark.parametrize("poly", [0, 3])
@pytest.mark.parametrize("N,B", [(256, 1), (1024, 2), (4096, 2)])
def test_self_attention_tcgen05_pair(cuda, monkeypatch, N, B, poly):
    lib = _lib.load()
    d = 40
    qkv = _mk((B, N, 3 * H * d), cuda, 31 + N, 1.0)
    qkv[..., :2 * H * d] *= 1.5
    ident = list(range(B))
    ref = _self_ref(qkv, d, ident, ident, ident)
    monkeypatch.setenv("PNP_ATTN_CLUSTER", "1")
    monkeypatch.setenv("PNP_ATTN_POLY", "0")
    base = torch.zeros(B, N, H * d, dtype=torch.float16, device=cuda)
    _lib.check(lib.pnp_test_self_attention_tc(G.ptr(qkv), B, N, None, None, None, G.ptr(base), G.stream()))
    monkeypatch.setenv("PNP_ATTN_CLUSTER", "3")
    monkeypatch.setenv("PNP_ATTN_POLY", str(poly))
    out = torch.zeros(B, N, H * d, dtype=torch.float16, device=cuda)
    _lib.check(lib.pnp_test_self_attention_tc(G.ptr(qkv), B, N, None, None, None, G.ptr(out), G.stream()))
    torch.cuda.synchronize()
    err, err0 = G.rel_l2(out, ref), G.rel_l2(base, ref)
    print(f"pair tc attention N={N} B={B} poly={poly}: rel-L2 {err:.3e} (single-CTA kernel {err0:.3e}), "
          f"identical to it: {torch.equal(out, base)}")
    assert err < 2e-3
    if poly == 0:
        assert G.rel_l2(out, base.float()) < 2e-4  # same products and sums; only the V^T padding (48 -> 64 columns) differs


def test_self_attention_tcgen05_pair_row_indirection_and_fallback(cuda, monkeypatch):
    monkeypatch.setenv("PNP_ATTN_CLUSTER", "3")
    monkeypatch.setenv("PNP_ATTN_POLY", "3")
    lib = _lib.load()
    B, N, d = 4, 1024, 40
    qkv = _mk((B, N, 3 * H * d), cuda, 77, 1.2)
    q_row, k_row, v_row = [0, 1, 2, 2], [0, 0, 2, 2], [0, 0, 2, 3]
    dq, dk, dv = (torch.tensor(r, dtype=torch.int32, device=cuda) for r in (q_row, k_row, v_row))
    out = torch.zeros(B, N, H * d, dtype=torch.float16, device=cuda)
    _lib.check(lib.pnp_test_self_attention_tc(G.ptr(qkv), B, N, G.ptr(dq), G.ptr(dk), G.ptr(dv), G.ptr(out), G.stream()))
    torch.cuda.synchronize()
    assert G.rel_l2(out, _self_ref(qkv, d, q_row, k_row, v_row)) < 2e-3
    # later keys with far larger scores: the optimistic pass overflows fp16 and the two-pass schedule takes over, cluster-wide
    B = 2
    qkv = _mk((B, N, 3 * H * d), cuda, 91, 1.0)
    c = H * d
    qkv[:, 256:, c:2 * c] *= 12.0
    out = torch.zeros(B, N, H * d, dtype=torch.float16, device=cuda)
    _lib.check(lib.pnp_test_self_attention_tc(G.ptr(qkv), B, N, None, None, None, G.ptr(out), G.stream()))
    torch.cuda.synchronize()
    ident = list(range(B))
    assert torch.isfinite(out.float()).all()
    err = G.rel_l2(out, _self_ref(qkv, d, ident, ident, ident))
    print("pair tc attention with forced fallback: rel-L2", err)
    assert err < 3e-3


@pytest.mark.parametrize("mode,poly", [(1, 0), (1, 3), (3, 3)])
def test_self_attention_tcgen05_wide_logits(cuda, monkeypatch, mode, poly):
    """Logits with a standard deviation of 4 nats: most probabilities of a row are below 2^-15 of its maximum - where the
    polynomial path flushes to zero and the MUFU path returns fp16 subnormals.  Both must stay within the tolerance."""
    monkeypatch.setenv("PNP_ATTN_CLUSTER", str(mode))
    monkeypatch.setenv("PNP_ATTN_POLY", str(poly))
    lib = _lib.load()
    B, N, d = 1, 4096, 40
    qkv = _mk((B, N, 3 * H * d), cuda, 123, 1.0)
    qkv[..., :2 * H * d] *= 2.0
    out = torch.zeros(B, N, H * d, dtype=torch.float16, device=cuda)
    _lib.check(lib.pnp_test_self_attention_tc(G.ptr(qkv), B, N, None, None, None, G.ptr(out), G.stream()))
    torch.cuda.synchronize()
    err = G.rel_l2(out, _self_ref(qkv, d, [0], [0], [0]))
    print(f"tc attention, wide logits, mode {mode} poly {poly}: rel-L2 {err:.3e}")
    assert err < 3e-3


@pytest.mark.parametrize("cluster,roles,sched", [(1, 1, 0), (1, 0, 1), (1, 1, 1), (3, 0, 1), (3, 1, 0)])
def test_self_attention_tcgen05_role_layout_and_issue_order_are_bit_identical(cuda, monkeypatch, cluster, roles, sched):
    """PNP_ATTN_ROLES (TMA / MMA roles on the highest warp ids) and PNP_ATTN_SCHED (event-driven MMA issue order, barriers
    probed with mbarrier.test_wait) change when instructions are issued, not what they compute."""
    lib = _lib.load()
    B, N, d = 2, 4096, 40
    qkv = _mk((B, N, 3 * H * d), cuda, 211, 1.0)
    qkv[..., :2 * H * d] *= 1.5
    outs = []
    for env in ((1, 0, 0), (cluster, roles, sched)):
        for k, v in zip(("PNP_ATTN_CLUSTER", "PNP_ATTN_ROLES", "PNP_ATTN_SCHED"), env):
            monkeypatch.setenv(k, str(v))
        monkeypatch.setenv("PNP_ATTN_POLY", "0")
        out = torch.zeros(B, N, H * d, dtype=torch.float16, device=cuda)
        _lib.check(lib.pnp_test_self_attention_tc(G.ptr(qkv), B, N, None, None, None, G.ptr(out), G.stream()))
        torch.cuda.synchronize()
        outs.append(out)
    ident = list(range(B))
    assert G.rel_l2(outs[0], _self_ref(qkv, d, ident, ident, ident)) < 2e-3
    assert torch.equal(outs[0], outs[1])
