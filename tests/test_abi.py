"""CPU-side checks of the C-ABI boundary: the library builds for sm_100a, loads, and exports every symbol
include/dm_b200.h declares with the argument count the ctypes binding uses.  No compute calls (no GPU here)."""
import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "dm_b200.h")


def _declared():
    """{name: n_args} parsed from the header."""
    src = open(HEADER).read()
    src = re.sub(r"/\*.*?\*/", " ", src, flags=re.S)
    src = re.sub(r"//[^\n]*", " ", src)
    out = {}
    for m in re.finditer(r"\b(?:int|size_t|const char\s*\*)\s+(dm_\w+)\s*\(([^;{]*?)\)\s*;", src, flags=re.S):
        args = m.group(2).strip()
        out[m.group(1)] = 0 if args in ("", "void") else args.count(",") + 1
    return out


@pytest.fixture(scope="module")
def lib():
    from densematcher_b200 import build, _lib
    build.build()
    return _lib.load()


def test_header_declares_the_hot_path():
    d = _declared()
    for name in ("dm_nn_argmax_f32", "dm_nn_argmax_f64", "dm_project", "dm_fmap_solve", "dm_fm_to_p2p", "dm_p2p_to_fm",
                 "dm_zoomout", "dm_icp", "dm_mapped_indicator", "dm_match_dist_f32", "dm_last_error", "dm_version"):
        assert name in d, name


def test_library_exports_every_declared_symbol(lib):
    from densematcher_b200 import _lib
    decl = _declared()
    raw = ctypes.CDLL(_lib.LIB_PATH)
    for name, nargs in decl.items():
        assert hasattr(raw, name), f"{name} declared in dm_b200.h but not exported"
        assert name in _lib.SIGNATURES, f"{name} has no ctypes signature"
        assert len(_lib.SIGNATURES[name][1]) == nargs, f"{name}: header has {nargs} args, binding {len(_lib.SIGNATURES[name][1])}"
    assert set(_lib.SIGNATURES) == set(decl)


def test_version_and_build_info(lib):
    assert lib.dm_version() == 200
    info = lib.dm_build_info().decode()
    assert "sm_100a" in info


def test_workspace_queries_are_pure(lib):
    n = lib.dm_nn_workspace_bytes(4, 8000, 8000, 2000, 2000, 384, 1, 1, 0)
    assert n > 0 and n % 256 == 0
    assert lib.dm_nn_workspace_bytes(4, 8000, 8000, 2000, 2000, 384, 1, 0, 0) < n
    assert lib.dm_nn_workspace_bytes(-1, 0, 0, 0, 0, 384, 1, 0, 0) == 0
    assert lib.dm_zoomout_workspace_bytes(2, 4000, 4000, 2000, 2000, 30, 30, 10, 1, 1, 0) > 0
    assert lib.dm_fmap_solve_workspace_bytes(3, 100, 100, 384) >= 3 * 2 * 100 * 100 * 8


def test_bad_arguments_are_reported_not_crashed(lib):
    # null operands with non-empty sizes: error code + message, no CUDA call needed to find out
    rc = lib.dm_nn_argmax_f32(None, 384, None, 10, 10, None, 384, None, 10, 10, 1, 384, None, 1, None, 0, 0, None, 0, None)
    assert rc == -1 and b"null" in lib.dm_last_error()
    rc = lib.dm_fmap_solve(None, None, None, None, None, 1.0, 1.0, 1, 1, 1, 8, None, None, 0, None)
    assert rc == -1
    rc = lib.dm_zoomout(None, 30, 30, 5, 1, 1, None, 20, None, 0, 0, None, 20, None, 0, 0, None, 1, None, None, 0, None, 0, None)
    assert rc == -1


def test_python_api_refuses_cpu_tensors():
    import torch
    from densematcher_b200 import nn
    with pytest.raises(ValueError):
        nn.nn_argmax(torch.zeros(4, 8), torch.zeros(4, 8))


def test_no_product_import_of_the_oracle():
    """The product package must never import oracle/ (parity claims depend on it)."""
    pkg = os.path.join(ROOT, "densematcher_b200")
    for dp, _, files in os.walk(pkg):
        for f in files:
            if f.endswith(".py"):
                src = open(os.path.join(dp, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle\b", src, flags=re.M), os.path.join(dp, f)


def test_header_is_plain_c():
    """The boundary is a C ABI: include/dm_b200.h must compile as C99 on its own (no C++, no torch types)."""
    import shutil
    import subprocess
    gcc = shutil.which("gcc")
    if gcc is None:
        pytest.skip("gcc not available")
    r = subprocess.run([gcc, "-std=c99", "-fsyntax-only", "-x", "c", HEADER], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr


def test_bank_state_grows_by_the_missing_meshes_only(monkeypatch):
    """Host logic of the piecewise bank preparation (fm.bank_prepare with a previous state): the prepared range stays
    contiguous, every mesh is prepared once, covered requests do nothing.  The library call is replaced by a recorder."""
    import numpy as np
    import torch
    from densematcher_b200 import fm
    off = np.array([0, 10, 25, 30, 50, 64, 80], dtype=np.int64)
    st = fm.BankState(torch.zeros(80, 8), torch.zeros(80, 4, dtype=torch.float64), torch.zeros(80, dtype=torch.float64),
                      torch.zeros(6, 4, dtype=torch.float64), torch.from_numpy(off), off, 4, torch.zeros(256, dtype=torch.uint8))
    calls = []
    monkeypatch.setattr(fm, "_bank_prepare_range", lambda bank, lo, hi, workspace=None: calls.append((lo, hi)))
    assert not st.covers(0, 1) and st.covers(3, 3)
    fm.bank_prepare(None, None, None, None, None, 4, mesh_range=(2, 4), bank=st)
    assert calls == [(2, 4)] and (st.lo, st.hi) == (2, 4)
    fm.bank_prepare(None, None, None, None, None, 4, mesh_range=(2, 3), bank=st)          # covered: nothing to do
    assert calls == [(2, 4)]
    fm.bank_prepare(None, None, None, None, None, 4, mesh_range=(1, 6), bank=st)          # both sides
    assert calls == [(2, 4), (1, 2), (4, 6)] and (st.lo, st.hi) == (1, 6)
    fm.bank_prepare(None, None, None, None, None, 4, mesh_range=None, bank=st)            # the whole bank
    assert calls[-1] == (0, 1) and (st.lo, st.hi) == (0, 6) and len(calls) == 4
    fm.bank_prepare(None, None, None, None, None, 4, mesh_range=(-3, 99), bank=st)        # clamped, covered
    assert len(calls) == 4
