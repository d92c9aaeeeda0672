"""An INDEPENDENT restatement of src/qp.cpp in numpy -- the KKT system of qp.cpp:90 solved by LAPACK's LU with partial pivoting
(scipy.linalg.lu_factor), not by a diagonal-pivoted LDL^T -- against the C oracle (oracle/qp_oracle_impl.h, Eigen-style LDLT).

Why: the reference cannot be built here (Eigen absent), so the oracle's Eigen::LDLT restatement is unpinned below the reference's own
1e-2 assertions. This test shows that the ADMM trajectory does not depend on which backward-stable factorisation solves the KKT system:
two different factorisations give the same termination status, the same iteration count and iterates equal to ~1e-9 on the benchmark's
synthetic QPs (default settings AND adaptive rho). A real Eigen build is a third backward-stable LDL^T of the same matrix."""
import numpy as np
import pytest
import scipy.linalg as sla

RHO_MIN, RHO_MAX, RHO_TOL, RHO_EQ_FACTOR, LOOSE = 1e-6, 1e6, 1e-4, 1e3, 1e16
EPS = np.finfo(np.float64).eps


def solve_numpy(P, q, A, l, u, rho=0.1, sigma=1e-6, alpha=1.0, eps_rel=1e-3, eps_abs=1e-3, max_iter=1000, check_termination=25,
                adaptive_rho=False, adaptive_rho_tolerance=5.0, adaptive_rho_interval=25):
    n, m = P.shape[0], A.shape[0]
    ctype = np.where((l < -LOOSE) & (u > LOOSE), 2, np.where(u - l < RHO_TOL, 1, 0))  # qp.cpp:283-294

    def rho_vec(r0):  # qp.cpp:296-314
        return np.where(ctype == 2, RHO_MIN, np.where(ctype == 1, RHO_EQ_FACTOR * r0, r0))

    Plow = np.tril(P) + np.tril(P, -1).T  # LDLT<Lower> reads the lower triangle only (qp.hpp:129, qp.cpp:185-187)

    def factor(rv):
        K = np.block([[Plow + sigma * np.eye(n), A.T], [A, -np.diag(1.0 / rv)]])
        if not np.all(np.isfinite(K)):
            return None
        return sla.lu_factor(K)

    rv = rho_vec(rho)
    rho_updates = 1
    lu = factor(rv)
    if lu is None:
        return dict(status=3, iter=0, x=np.zeros(n), y=np.zeros(m), rho_updates=rho_updates)
    x, z, y = np.zeros(n), np.zeros(m), np.zeros(m)
    status, it = 2, 0
    for it in range(1, max_iter + 1):
        z_prev = z
        sol = sla.lu_solve(lu, np.concatenate([sigma * x - q, z - y / rv]))  # qp.cpp:272-276, :90
        x_t, nu = sol[:n], sol[n:]
        z_t = z_prev + (nu - y) / rv  # qp.cpp:93
        x = alpha * x_t + (1 - alpha) * x
        zh = alpha * z_t + (1 - alpha) * z_prev
        z = np.minimum(np.maximum(zh + y / rv, l), u)  # qp.cpp:278-281
        y = y + rv * (zh - z)
        chk = check_termination and it % check_termination == 0
        adapt = adaptive_rho and it % adaptive_rho_interval == 0
        if chk or adapt:  # update_state, qp.cpp:316-331
            Ax, Px, Aty = A @ x, P @ x, A.T @ y
            sc_p = max(np.abs(Ax).max(), np.abs(z).max())
            sc_d = max(np.abs(Px).max(), np.abs(Aty).max(), np.abs(q).max())
            rp, rd = np.abs(Ax - z).max(), np.abs(Px + q + Aty).max()
            if chk and rp <= eps_abs + eps_rel * sc_p and rd <= eps_abs + eps_rel * sc_d:
                status = 0
                break
            if adapt:  # qp.cpp:125-144, :333-341
                new_rho = rho * np.sqrt((rp / (sc_p + EPS)) / (rd / (sc_d + EPS) + EPS))
                new_rho = max(RHO_MIN, min(new_rho, RHO_MAX))
                if new_rho < rho / adaptive_rho_tolerance or new_rho > rho * adaptive_rho_tolerance:
                    rho = new_rho
                    rv = rho_vec(rho)
                    rho_updates += 1
                    lu = factor(rv)
    else:
        status, it = 1, max_iter + 1  # qp.cpp:147-150
    return dict(status=status, iter=it, x=x, y=y, rho_updates=rho_updates)


@pytest.mark.parametrize("n,m,batch,kw", [
    (8, 12, 6, {}), (8, 12, 6, dict(alpha=1.6, adaptive_rho=True)),
    (32, 64, 4, {}), (32, 64, 4, dict(alpha=1.6, adaptive_rho=True)),
    (64, 128, 2, {}), (64, 128, 2, dict(alpha=1.6, adaptive_rho=True)),
])
def test_trajectory_is_independent_of_the_kkt_factorisation(oracle, n, m, batch, kw):
    from sqp_solver_b200.synth import make_batch

    d = make_batch(batch, n, m, seed0=4200 + n)
    okw = {k: (1 if v is True else v) for k, v in kw.items()}
    ref = oracle.solve_batch(d["P"], d["q"], d["A"], d["l"], d["u"], oracle.default_settings(**okw), nthreads=2)
    for i in range(batch):
        r = solve_numpy(d["P"][i].reshape(n, n, order="F"), d["q"][i], d["A"][i].reshape(m, n, order="F"), d["l"][i], d["u"][i], **kw)
        assert r["status"] == ref["status"][i] and r["iter"] == ref["iter"][i] and r["rho_updates"] == ref["rho_updates"][i], (i, r["status"], r["iter"])
        assert np.linalg.norm(r["x"] - ref["x"][i]) <= 1e-8 * np.linalg.norm(ref["x"][i]), i
        assert np.abs(r["y"] - ref["y"][i]).max() <= 1e-7 * max(1.0, np.abs(ref["y"][i]).max()), i


def test_simple_qp_cases_of_the_reference(oracle, golden):
    """tests/qp_solver_test.cpp:43-125 through the numpy restatement: the same iteration counts as the C oracle."""
    g = golden["simple_qp"]
    P, A = np.array(g["P"], dtype=float), np.array(g["A"], dtype=float)
    q, l, u = (np.array(g[k], dtype=float) for k in ("q", "l", "u"))
    for kw, okw in (({}, {}), (dict(eps_abs=float(np.float32(1e-4)), eps_rel=float(np.float32(1e-4))),) * 2,
                    (dict(adaptive_rho=True, adaptive_rho_interval=10), dict(adaptive_rho=1, adaptive_rho_interval=10))):
        r = solve_numpy(P, q, A, l, u, **kw)
        ref = oracle.solve_batch(P.reshape(1, -1, order="F"), q[None], A.reshape(1, -1, order="F"), l[None], u[None], oracle.default_settings(**okw))
        assert r["status"] == ref["status"][0] == 0 and r["iter"] == ref["iter"][0]
        np.testing.assert_allclose(r["x"], ref["x"][0], rtol=0, atol=1e-10)
