"""CPU tests of the sparse synthetic workload (BASELINE config 5's generator) and of the host-side sharding arithmetic."""
import numpy as np

from sqp_solver_b200 import sharding
from sqp_solver_b200.synth import densify, make_batch, make_sparse_batch


def test_sparse_batch_structure_and_feasibility():
    d = make_sparse_batch(5, 40, 70, density=0.1, seed0=3)
    m, n, nnz = 70, 40, d["nnz"]
    assert d["outer"].dtype == np.int32 and d["inner"].dtype == np.int32
    assert d["outer"][0] == 0 and d["outer"][-1] == nnz and (np.diff(d["outer"]) >= 1).all()  # no empty rows
    assert (d["inner"] >= 0).all() and (d["inner"] < n).all()
    for i in range(m):  # column indices strictly increasing inside a row: no duplicate entries
        seg = d["inner"][d["outer"][i]:d["outer"][i + 1]]
        assert (np.diff(seg) > 0).all()
    A = densify(d).reshape(5, n, m).transpose(0, 2, 1)  # [B, m, n]
    assert ((A != 0).sum(axis=(1, 2)) == nnz).all()
    # the same seeds give the same instances; P is symmetric positive definite; l <= u with equalities and loose rows present
    d2 = make_sparse_batch(5, 40, 70, density=0.1, seed0=3)
    assert np.array_equal(d["vals"], d2["vals"]) and np.array_equal(d["P"], d2["P"])
    P0 = d["P"][0].reshape(n, n, order="F")
    assert np.allclose(P0, P0.T) and np.linalg.eigvalsh(P0).min() > 0
    assert (d["l"] <= d["u"]).all() and (d["l"] == d["u"]).any() and (d["u"] > 1e19).any()


def test_sparse_oracle_solution_satisfies_the_constraints(oracle):
    d = make_sparse_batch(3, 30, 50, density=0.12, seed0=11)
    A = densify(d)
    r = oracle.solve_batch(d["P"], d["q"], A, d["l"], d["u"], oracle.default_settings(eps_abs=1e-5, eps_rel=1e-5, max_iter=4000, alpha=1.6, adaptive_rho=1))
    assert (r["status"] == 0).all()
    for i in range(3):
        Ai = A[i].reshape(50, 30, order="F")
        ax = Ai @ r["x"][i]
        assert (ax >= d["l"][i] - 1e-4).all() and (ax <= d["u"][i] + 1e-4).all()


def test_shard_ranges_tile_the_batch():
    for batch in (0, 1, 7, 8192, 2048):
        for world in (1, 2, 3, 8):
            edges = [sharding.shard_range(batch, r, world) for r in range(world)]
            assert edges[0][0] == 0 and edges[-1][1] == batch
            assert all(edges[r][1] == edges[r + 1][0] for r in range(world - 1))
            sizes = [hi - lo for lo, hi in edges]
            assert max(sizes) - min(sizes) <= 1


def test_dense_generator_is_reproducible():
    a, b = make_batch(3, 6, 9, seed0=5), make_batch(3, 6, 9, seed0=5)
    assert all(np.array_equal(a[k], b[k]) for k in ("P", "q", "A", "l", "u"))
