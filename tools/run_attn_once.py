"""The tcgen05 self-attention at the UNet's largest shape (4096 tokens, 8 heads of 40) - target for ncu, and with `sweep`
the launch variants side by side (PNP_ATTN_CLUSTER: 1 one CTA per query tile, 2 multicast cluster, 3 cta_group::2 pair;
PNP_ATTN_POLY: packed exponentials per 8 on the FMA pipe; PNP_ATTN_ROLES: roles on the highest warp ids; PNP_ATTN_SCHED:
event-driven MMA issue order).

    python tools/run_attn_once.py            # one variant (the environment's), B = 4
    python tools/run_attn_once.py sweep [B ...]   # all variants (x warp-role layouts, PNP_ATTN_ROLES) at B = 4 and 32
"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from pnpinversion_b200 import _lib  # noqa: E402
from tests import gpu_util as G  # noqa: E402

lib = _lib.load()
N, H, d = 4096, 8, 40


def time_variant(qkv, out, B, reps=10):
    for _ in range(3):
        _lib.check(lib.pnp_test_self_attention_tc(G.ptr(qkv), B, N, None, None, None, G.ptr(out), G.stream()))
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        _lib.check(lib.pnp_test_self_attention_tc(G.ptr(qkv), B, N, None, None, None, G.ptr(out), G.stream()))
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) * 1000.0 / reps


def main():
    sweep = len(sys.argv) > 1 and sys.argv[1] == "sweep"
    batches = [int(a) for a in sys.argv[2:]] or [4, 32]
    for B in (batches if sweep else (4,)):
        g = torch.Generator(device="cpu").manual_seed(5)
        qkv = torch.randn(B, N, 3 * H * d, generator=g).to(torch.float16).cuda()
        out = torch.zeros(B, N, H * d, dtype=torch.float16, device="cuda")
        # (cluster mode, poly, roles_hi, sched)
        variants = [(1, 0, 0, 0), (1, 0, 0, 1), (1, 3, 0, 1), (1, 0, 1, 1), (3, 0, 0, 1), (3, 3, 0, 1)] if sweep else [None]
        base = None
        for v in variants:
            if v is not None:
                for k, val in zip(("PNP_ATTN_CLUSTER", "PNP_ATTN_POLY", "PNP_ATTN_ROLES", "PNP_ATTN_SCHED"), v):
                    os.environ[k] = str(val)
            us = time_variant(qkv, out, B)
            if sweep:  # kernel-only time (V transpose + attention, CUDA events inside the entry point) and the role counters -> stderr
                os.environ["PNP_ATTN_PROF"] = "1"
                for _ in range(2):
                    _lib.check(lib.pnp_test_self_attention_tc(G.ptr(qkv), B, N, None, None, None, G.ptr(out), G.stream()))
                del os.environ["PNP_ATTN_PROF"]
            if base is None:
                base = out.clone()
            diff = float((out.float() - base.float()).norm() / base.float().norm())
            gf = 4.0 * B * H * N * N * d / 1e9
            print(f"tc attention B={B} N=4096 variant (cluster, poly, roles_hi, sched)={v}: {us:8.1f} us per call (incl. V transpose + plan + "
                  f"vt alloc) = {gf / us * 1e-3:6.1f} TFLOP/s; rel diff to the first variant {diff:.2e}", flush=True)


if __name__ == "__main__":
    main()
