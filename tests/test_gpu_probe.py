"""The tcgen05.mma instruction-cost probe (csrc/probe.cu, tools/mma_probe.py) runs and reports plausible cycle counts: the
shapes of the d = 40 attention as it is (M = 128, N = 48 / 128, A in tensor memory) and of its pair-mode variant
(cta_group::2, M = 256 on two SMs).  The numbers themselves are in profiles/r2_mma_probe.txt."""
import ctypes as C

import pytest

from pnpinversion_b200 import _lib



@pytest.mark.gpu
@pytest.mark.parametrize("cg,M,N,ts", [(1, 128, 48, 1), (1, 128, 128, 1), (1, 128, 256, 0), (2, 256, 64, 1), (2, 256, 128, 1)])
def test_mma_probe_reports_cycles(cuda, cg, M, N, ts):
    lib = _lib.load()
    out = (C.c_int64 * 2)()
    n = 256
    _lib.check(lib.pnp_test_mma_probe(cg, M, N, ts, n, 1, 8, 0, out))
    issue, total = out[0] / n, out[1] / n
    print(f"tcgen05.mma cta_group::{cg} M={M} N={N} A from {'tmem' if ts else 'smem'}: {total:.1f} cycles per instruction")
    math_cycles = (M // cg) * N * 16 / 4096
    assert math_cycles * 0.9 <= total < 400 and 0 < issue <= total + 1


def test_mma_probe_rejects_shapes_the_instruction_does_not_have():
    lib = _lib.load()
    out = (C.c_int64 * 2)()
    for args in ((1, 256, 64, 1, 64, 1, 8, 0), (2, 256, 48, 1, 64, 1, 8, 0), (1, 128, 40, 1, 64, 1, 8, 0),
                 (1, 128, 256, 1, 64, 4, 8, 0), (1, 128, 64, 1, 64, 1, 9, 0), (1, 128, 64, 1, 60, 1, 8, 1)):
        assert lib.pnp_test_mma_probe(*args, out) != 0
