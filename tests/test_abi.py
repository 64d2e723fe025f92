"""CPU: the C-ABI library loads and exports every symbol include/psoap_b200.h declares (no compute calls)."""
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def lib():
    from psoap_b200 import _build, _lib
    _build.build_library()
    return _lib.load()


def header_symbols():
    text = open(os.path.join(ROOT, "include", "psoap_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(psoap_[a-z0-9_]+)\s*\(", text)))


def test_every_declared_symbol_is_exported_and_bound(lib):
    from psoap_b200 import _lib
    syms = header_symbols()
    assert len(syms) >= 20
    for s in syms:
        assert hasattr(lib, s), f"{s} declared in include/psoap_b200.h but not exported"
        assert s in _lib.SIGNATURES, f"{s} has no ctypes signature"
    for s in _lib.SIGNATURES:
        assert s in syms, f"{s} bound in _lib.py but not declared in the header"


def test_pure_host_entry_points(lib):
    assert lib.psoap_version() >= 100
    assert [lib.psoap_model_ncomp(m) for m in range(1, 6)] == [1, 2, 1, 2, 3]
    assert [lib.psoap_model_norb(m) for m in range(1, 6)] == [6, 7, 11, 12, 13]  # utils.py:14
    assert lib.psoap_lnlike_workspace_bytes(9000) > 8 * 9088 * 9088
    assert lib.psoap_schur_workspace_bytes(1000, 200) > 0
    assert lib.psoap_launch_count() == 0


def test_argument_errors_do_not_touch_the_device(lib):
    from psoap_b200 import _lib
    rc = lib.psoap_fill_v11(4, None, 0, 0, None, None, None, None, None, None)
    assert rc == -1 and b"psoap_fill_v11" in lib.psoap_last_error()
    with pytest.raises(_lib.PsoapError):
        _lib.check(lib.psoap_lnlike(0, 0, None, None, None, None, None, None, None, 1.0, None, 0, None, None))


def test_no_cpu_fallback_without_cuda():
    import torch
    if torch.cuda.is_available():
        pytest.skip("CUDA present")
    import numpy as np
    from psoap_b200 import covariance, matrix_functions
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        matrix_functions.fill_V11_f(np.empty((4, 4)), np.zeros(4), 0.1, 5.0)
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        covariance.lnlike_f(None, np.zeros(4), np.ones(4), np.ones(4), 0.1, 5.0)


def test_product_never_imports_oracle():
    """The oracle is test infrastructure: nothing under psoap_b200/ may import, link or call it."""
    pkg = os.path.join(ROOT, "psoap_b200")
    pat = re.compile(r"^\s*(from|import)\s+[\w.]*oracle|psoap_oracle|liboracle|oracle/_ref|oracle\.py", re.M)
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                assert not pat.search(open(os.path.join(dirpath, f)).read()), f"{f} reaches into oracle/"


@pytest.mark.parametrize("R", [1, 2, 3, 5, 8, 47, 256])
def test_trailing_update_tiles_cover_the_lower_triangle_exactly_once(lib, R):
    """Tile enumeration of the dominant kernel (host replay of SyrkSrc::decode): part 0 covers every 128x64 tile
    of the lower triangle once; for look-ahead, parts 1 and 2 split it without overlap for every group size."""
    import ctypes
    full = {(r, j) for r in range(R) for j in range(2 * r + 2)}
    cap = R * (R + 1) + 16
    rows, cols = (ctypes.c_int * cap)(), (ctypes.c_int * cap)()

    def tiles(part, ncol1):
        n = lib.psoap_debug_syrk_tiles(R, part, ncol1, rows, cols, cap)
        assert n >= 0
        out = [(rows[i], cols[i]) for i in range(n)]
        assert len(set(out)) == n, "a tile is visited twice"
        return set(out)

    assert tiles(0, 2) == full
    for ncol1 in (2, 4, 6, 8):
        p1, p2 = tiles(1, ncol1), tiles(2, ncol1)
        assert p1 == {(r, j) for (r, j) in full if j < ncol1}
        assert p2 == {(r, j) for (r, j) in full if j >= ncol1}
