"""tcgen05.mma instruction cost by shape (csrc/probe.cu): every configuration runs in its own process (an illegal shape
kills only its own CUDA context), `n` back-to-back MMAs over 1 / 2 / 4 accumulators.

    python tools/mma_probe.py [n]            # table on stdout
    python tools/mma_probe.py group <cg1|cg2_smem|cg2_tmem> n   (internal)
"""
import ctypes as C
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

CONFIGS = [  # (cta_group, M, N, A from TMEM, what)
    (1, 128, 48, 1, "P V of the d=40 attention as it is"),
    (1, 128, 64, 1, "P V padded to N=64"),
    (1, 128, 128, 1, "Q K^T as it is (128 keys)"),
    (1, 128, 256, 1, "Q K^T over two key tiles"),
    (1, 128, 48, 0, "P V with P in shared memory"),
    (1, 128, 160, 0, "GEMM tile 128x160 (reference point)"),
    (1, 128, 256, 0, "GEMM tile 128x256 (reference point)"),
    (1, 64, 128, 0, "transposed product O^T = V^T P^T, 128 queries"),
    (1, 64, 256, 0, "transposed product, 256 queries"),
    (2, 256, 64, 1, "pair: P V for 256 queries on two SMs"),
    (2, 256, 128, 1, "pair: Q K^T for 256 queries"),
    (2, 256, 256, 1, "pair: Q K^T, two key tiles"),
    (2, 256, 64, 0, "pair: P V with P in shared memory"),
    (2, 256, 256, 0, "pair: GEMM tile 256x256 (reference point)"),
]


def group(which, n):
    """All configurations of one group in one process (a CUDA error ends the group: the context is gone)."""
    from pnpinversion_b200 import _lib

    lib = _lib.load()
    out = (C.c_int64 * 2)()
    for cg, M, N, ts, what in CONFIGS:
        if GROUP_OF(cg, ts) != which:
            continue
        for nacc in (1, 2, 4):
            if nacc * ((N + 31) // 32 * 32) > 448:
                continue
            try:
                for _ in range(2):  # second run: instruction cache and clocks warm
                    _lib.check(lib.pnp_test_mma_probe(cg, M, N, ts, n, nacc, 8, 0, out))
            except Exception as e:  # noqa: BLE001
                print(f"{cg:>9} {M:>4} {N:>4} {'tmem' if ts else 'smem':>5} {nacc:>3}   FAILED: {str(e)[:120]}", flush=True)
                return
            math = (M // cg) * N * 16 / 4096
            print(f"{cg:>9} {M:>4} {N:>4} {'tmem' if ts else 'smem':>5} {nacc:>3} {out[1] / n:8.1f} {out[0] / n:8.1f} {math:5.0f}  {what}",
                  flush=True)


def GROUP_OF(cg, ts):
    return "cg1" if cg == 1 else ("cg2_smem" if not ts else "cg2_tmem")


def main():
    if len(sys.argv) > 1 and sys.argv[1] == "group":
        group(sys.argv[2], int(sys.argv[3]))
        return
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
    print(f"# tcgen05.mma.kind::f16, K = 16 per instruction, {n} back-to-back instructions; cycles per instruction until the commit "
          f"arrives (issue-side cycles in brackets); math = M*N*16 / 4096 MAC per cycle and SM")
    print(f"{'cta_group':>9} {'M':>4} {'N':>4} {'A':>5} {'acc':>3} {'cyc/MMA':>8} {'(issue)':>8} {'math':>5}  what")
    for g in ("cg1", "cg2_smem", "cg2_tmem"):
        r = subprocess.run([sys.executable, os.path.abspath(__file__), "group", g, str(n)], capture_output=True, text=True,
                           timeout=300)
        sys.stdout.write(r.stdout)
        if r.returncode != 0:
            print(f"# group {g}: exit code {r.returncode}: {(r.stderr.strip().splitlines() or ['?'])[-1][:200]}")
        sys.stdout.flush()


if __name__ == "__main__":
    main()
