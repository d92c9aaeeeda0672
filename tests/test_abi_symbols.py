"""CPU-side checks of the drop-in boundary: the C-ABI library loads and exports every symbol
include/sqp_b200_qp.h declares; host-only entry points work; device entry points fail loudly
(no CPU fallback) when there is no GPU."""
import os
import re

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def api():
    from sqp_solver_b200 import api, build

    build.build()
    return api


def declared_symbols():
    src = open(os.path.join(ROOT, "include", "sqp_b200_qp.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(sqpb200_[a-z0-9_]+)\s*\(", src)))


def test_header_symbols_are_exported(api):
    L = api.load_library()
    decl = declared_symbols()
    assert len(decl) >= 20
    missing = [s for s in decl if not hasattr(L, s)]
    assert not missing, missing
    assert sorted(api.ABI_SYMBOLS) == decl
    assert L.sqpb200_abi_version() == 1


def test_default_settings_match_reference(api):
    s = api.default_settings()  # include/solvers/qp.hpp:38-53
    assert (s.rho, s.sigma, s.alpha, s.eps_rel, s.eps_abs) == (1e-1, 1e-6, 1.0, 1e-3, 1e-3)
    assert (s.max_iter, s.check_termination, s.warm_start, s.adaptive_rho) == (1000, 25, 0, 0)
    assert (s.adaptive_rho_tolerance, s.adaptive_rho_interval, s.verbose) == (5.0, 25, 0)
    assert (api.SOLVED, api.MAX_ITER_EXCEEDED, api.UNSOLVED, api.NUMERICAL_ISSUES, api.UNINITIALIZED) == (0, 1, 2, 3, 4)


def test_constr_type_init_host(api, golden, oracle):  # tests/qp_solver_test.cpp:127-156
    g = golden["test_constraint"]
    got = api.constr_type_init(g["l"], g["u"])
    assert got.tolist() == g["type_expect"]
    rng = np.random.default_rng(0)
    l = rng.standard_normal(200)
    u = l + np.where(rng.uniform(size=200) < 0.3, 5e-5, 1.0)
    l[::7] = -1e20
    u[::14] = 1e20
    l[3] = -np.inf
    u[3] = np.inf
    assert got.dtype == np.int32
    np.testing.assert_array_equal(api.constr_type_init(l, u), oracle.QPSolver.constr_type_init(l, u))


def test_staging_chunk_bounds(api):
    """The chunk schedule of a staged HOST_PTRS call (host-side logic, no GPU): strictly increasing boundaries from 0 to count, a fine
    prefix (the first CTAs of the persistent launch start after a fraction of a per cent of the transfer), never more flags than fit."""
    import ctypes as C

    L = api.load_library()
    L.sqpb200_staging_chunk_bounds.argtypes = [C.c_int, C.c_int, C.c_int, C.POINTER(C.c_int), C.c_int]
    L.sqpb200_staging_chunk_bounds.restype = C.c_int
    cap = 96
    buf = (C.c_int * cap)()
    rng = np.random.default_rng(5)
    cases = [(8192, 16, 296), (8192, 64, 296), (1024, 16, 296), (100, 16, 296), (2048, 16, 37), (5, 5, 296), (1, 1, 1), (700, 7, 296)]
    cases += [(int(rng.integers(1, 20000)), int(rng.integers(1, 65)), int(rng.integers(1, 600))) for _ in range(300)]
    for count, chunks, wave in cases:
        nb = L.sqpb200_staging_chunk_bounds(count, chunks, wave, buf, cap)
        assert 1 <= nb <= cap - 1, (count, chunks, wave, nb)
        b = np.array(buf[:nb + 1])
        assert b[0] == 0 and b[-1] == count and (np.diff(b) > 0).all(), (count, chunks, wave, b)
        assert np.diff(b).max() <= max(-(-count // min(chunks, count)), 1), (count, chunks, wave, b)  # never coarser than equal chunks
    nb = L.sqpb200_staging_chunk_bounds(8192, 16, 296, buf, cap)
    b = np.array(buf[:nb + 1])
    assert b[1] <= 296 // 8 and (np.diff(b[b <= 8 * 296]) <= 296 // 2).all()  # ramp, then half-wave steps for eight waves
    assert L.sqpb200_staging_chunk_bounds(0, 16, 296, buf, cap) < 0 and L.sqpb200_staging_chunk_bounds(100, 64, 296, buf, 10) < 0


def test_no_cpu_fallback(api):
    import torch

    if torch.cuda.is_available():
        pytest.skip("GPU present")
    with pytest.raises(api.SolverError, match="no CUDA device|no CPU fallback"):
        api.Context(0)


def test_missing_library_fails_loudly(api, tmp_path):
    with pytest.raises(api.SolverError, match="missing"):
        api.load_library(str(tmp_path / "libsqp_b200.so"))


def test_product_never_imports_oracle():
    """The oracle is test infrastructure: nothing under sqp_solver_b200/ or include/ may reference it."""
    bad = []
    for base in ("sqp_solver_b200", "include"):
        for dp, _, files in os.walk(os.path.join(ROOT, base)):
            if "build" in dp.split(os.sep) or "__pycache__" in dp:
                continue
            for f in files:
                if f.endswith((".py", ".cu", ".cuh", ".h", ".hpp", ".cpp")):
                    txt = open(os.path.join(dp, f), errors="ignore").read()
                    if re.search(r"(from|import)\s+oracle|oracle/|liboracle", txt):
                        bad.append(os.path.join(dp, f))
    assert not bad, bad
