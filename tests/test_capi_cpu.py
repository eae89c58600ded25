"""The C-ABI library loads without a GPU and exports every symbol include/pnpinv.h declares; no compute calls."""
import ctypes as C
import os
import re

from pnpinversion_b200 import _lib, arch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared_symbols():
    src = open(os.path.join(ROOT, "include", "pnpinv.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(pnp_[a-z0-9_]+)\s*\(", src)))


def test_library_exports_every_declared_symbol():
    lib = _lib.load()
    names = _declared_symbols()
    assert len(names) >= 25
    for n in names:
        assert hasattr(lib, n), f"{n} declared in include/pnpinv.h but not exported"
    assert set(_lib.SIGNATURES) == set(names)


def test_param_table_matches_python_arch():
    lib = _lib.load()
    specs = arch.unet_param_specs()
    assert lib.pnp_unet_param_count() == len(specs) == 686
    got = {}
    for i in range(len(specs)):
        name = C.create_string_buffer(128)
        nd = C.c_int()
        shp = (C.c_int * 4)()
        assert lib.pnp_unet_param_spec(i, name, C.byref(nd), shp) == 0
        got[name.value.decode()] = tuple(shp[: nd.value])
    assert got == dict(specs)
    assert sum(int.__mul__(1, 1) * _numel(s) for s in got.values()) == 859520964


def _numel(s):
    n = 1
    for d in s:
        n *= d
    return n


def test_struct_layouts_match_header():
    assert C.sizeof(_lib.AttnCtrl) == 4 * (3 + 5 * 32 + 6 * 8 * 77 + 32 + 32)
    assert C.sizeof(_lib.BlendDesc) == 4 * (4 + 2 + 16 + 16 + 2 + 16 + 16 + 2)
    assert list(_lib.new_ctrl().conv_src_row) == list(range(32))
    assert _lib.StepArgs().loss_scale == 1.0 and _lib.new_ctrl().map_count[3][5] == 1 and _lib.new_ctrl().map_weight[7][76] == 1.0
    c = _lib.new_ctrl()
    assert list(c.self_q_row) == list(range(32)) and all(v == -1 for v in c.cross_base_row)
    assert list(c.mapper[0][:5]) == [0, 1, 2, 3, 4] and c.equalizer[7][76] == 1.0


def test_no_cpu_fallback():
    import torch

    if torch.cuda.is_available():
        return
    lib = _lib.load()
    h = C.c_void_p()
    assert lib.pnp_create(0, 4, C.byref(h)) != 0
    assert b"no CPU fallback" in lib.pnp_last_error()
