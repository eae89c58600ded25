"""What a tcgen05.commit costs in the issue stream (csrc/probe.cu): the same back-to-back MMAs in elected blocks of 2 / 3 / 4 / 8
instructions, with and without a commit after every block.  ctypes only (no torch import: the whole run takes seconds).

    python tools/mma_commit_probe.py [n]
"""
import ctypes as C
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
lib = C.CDLL(os.path.join(ROOT, "pnpinversion_b200", "libpnpinv.so"))
lib.pnp_test_mma_probe.argtypes = [C.c_int] * 8 + [C.POINTER(C.c_int64)]
lib.pnp_test_mma_probe.restype = C.c_int
lib.pnp_last_error.restype = C.c_char_p

SHAPES = [  # (cta_group, M, N, A from TMEM, what)
    (1, 128, 256, 0, "GEMM 128x256, 4 MMAs per 64-wide K block and commit"),
    (1, 128, 160, 0, "GEMM 128x160"),
    (1, 128, 128, 1, "attention S = Q K^T (3 MMAs per commit pair)"),
    (1, 128, 48, 1, "attention P V (8 MMAs per commit pair)"),
    (2, 256, 256, 0, "pair GEMM 256x256"),
    (2, 256, 64, 1, "pair attention P V"),
]


def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 960  # divisible by 2, 3, 4, 8
    out = (C.c_int64 * 2)()
    print(f"# {n} tcgen05.mma.kind::f16 (K = 16) in elected issue blocks; cycles per MMA until the final commit arrives, without / with a "
          f"tcgen05.commit after every block; last column = cycles added per commit")
    print(f"{'cta_group':>9} {'M':>4} {'N':>4} {'A':>5} {'block':>5} {'plain':>8} {'commits':>8} {'per commit':>10}  what")
    for cg, M, N, ts, what in SHAPES:
        for group in (2, 3, 4, 8):
            res = []
            for commit in (0, 1):
                rc = 0
                for _ in range(2):
                    rc = lib.pnp_test_mma_probe(cg, M, N, ts, n, 1, group, commit, out)
                if rc:
                    print(f"{cg:>9} {M:>4} {N:>4} {'tmem' if ts else 'smem':>5} {group:>5}   FAILED: {lib.pnp_last_error().decode()[:120]}")
                    return
                res.append(out[1] / n)
            print(f"{cg:>9} {M:>4} {N:>4} {'tmem' if ts else 'smem':>5} {group:>5} {res[0]:8.1f} {res[1]:8.1f} {(res[1] - res[0]) * group:10.1f}  {what}",
                  flush=True)


if __name__ == "__main__":
    main()
