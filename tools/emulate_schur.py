"""Numerics study on the CPU (no GPU involved): the device formulation of the KKT solve against the oracle.

The kernels eliminate the (2,2) block of the KKT matrix and apply H^-1 = (P_lowsym + sigma I + A^T diag(rho) A)^-1
per iteration. This script restates that iteration in numpy with three ways of applying
the solve -- explicit inverse ("inv"), explicit inverse plus ONE step of iterative refinement against the formed
H ("refine"), Cholesky-free LDL^T substitution ("ldl") -- and reports how far each ends from the C oracle
(tests only: imports oracle/) on the SQP-generated subproblems and on the fuzz-sweep instances.

    python tools/emulate_schur.py
"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))

RHO_MIN, RHO_MAX, RHO_TOL, RHO_EQ = 1e-6, 1e6, 1e-4, 1e3
EPS = np.finfo(np.float64).eps


def classify(l, u):
    t = np.zeros(l.shape[0], np.int32)
    loose = (l < -1e16) & (u > 1e16)
    eq = ~loose & (u - l < RHO_TOL)
    t[loose] = 2
    t[eq] = 1
    return t


def rho_vec(t, rho):
    return np.where(t == 2, RHO_MIN, np.where(t == 1, RHO_EQ * rho, rho))


def ldl_nopiv(H):
    n = H.shape[0]
    L = np.eye(n)
    d = np.zeros(n)
    S = H.copy()
    for k in range(n):
        d[k] = S[k, k]
        L[k + 1:, k] = S[k + 1:, k] / d[k]
        S[k + 1:, k + 1:] -= np.outer(L[k + 1:, k], S[k, k + 1:])
    return L, d


def sweep_inverse(H):
    """symmetric pivot-by-pivot elimination carried through to the inverse (the tile kernel's arithmetic)"""
    S = H.copy()
    n = S.shape[0]
    for k in range(n):
        d = S[k, k]
        inv = 1.0 / d
        c = S[:, k].copy()
        t = c * inv
        S -= np.outer(t, c)
        S[k, :] = c * inv
        S[:, k] = t
        S[k, k] = -inv
    return -S


def solve_emulated(P, q, A, l, u, s, mode):
    n, m = q.shape[0], l.shape[0]
    P = P.reshape(n, n, order="F")
    A = A.reshape(m, n, order="F")
    Pl = np.tril(P) + np.tril(P, -1).T
    t = classify(l, u)
    rho = s["rho"]
    rv = rho_vec(t, rho)
    x, z, y = np.zeros(n), np.zeros(m), np.zeros(m)
    alpha, sigma = s["alpha"], s["sigma"]

    def factor(rv):
        H = Pl + sigma * np.eye(n) + A.T @ (rv[:, None] * A)
        if mode == "ldl":
            return H, ldl_nopiv(H)
        return H, sweep_inverse(H)

    def apply(H, F, b):
        if mode == "ldl":
            L, d = F
            w = np.linalg.solve(L, b) if False else b.copy()
            for k in range(n):
                w[k + 1:] -= L[k + 1:, k] * w[k]
            w /= d
            for k in range(n - 1, -1, -1):
                w[k] -= L[k + 1:, k] @ w[k + 1:]
            return w
        xt = F @ b
        if mode == "refine":
            r = b - H @ xt
            xt = xt + F @ r
        return xt

    H, F = factor(rv)
    status, it = 2, 0
    rho_updates = 1
    for it in range(1, s["max_iter"] + 1):
        b = sigma * x - q + A.T @ (rv * z - y)
        xt = apply(H, F, b)
        zt = A @ xt
        x = alpha * xt + (1 - alpha) * x
        zh = alpha * zt + (1 - alpha) * z
        zn = np.minimum(np.maximum(zh + y / rv, l), u)
        y = y + rv * (zh - zn)
        z = zn
        chk = s["check_termination"] and it % s["check_termination"] == 0
        adapt = s["adaptive_rho"] and it % s["adaptive_rho_interval"] == 0
        if chk or adapt:
            Ax, Px, Aty = A @ x, P @ x, A.T @ y
            nrm = lambda v: np.abs(v).max() if v.size else 0.0
            sc_p = max(nrm(Ax), nrm(z))
            sc_d = max(nrm(Px), nrm(Aty), nrm(q))
            rp, rd = nrm(Ax - z), nrm(Px + q + Aty)
            if chk and rp <= s["eps_abs"] + s["eps_rel"] * sc_p and rd <= s["eps_abs"] + s["eps_rel"] * sc_d:
                status = 0
                break
            if adapt:
                nr = rho * np.sqrt((rp / (sc_p + EPS)) / (rd / (sc_d + EPS) + EPS))
                nr = max(RHO_MIN, min(nr, RHO_MAX))
                if nr < rho / s["adaptive_rho_tolerance"] or nr > rho * s["adaptive_rho_tolerance"]:
                    rho = nr
                    rho_updates += 1
                    rv = rho_vec(t, rho)
                    H, F = factor(rv)
    else:
        status, it = 1, s["max_iter"] + 1
    return dict(x=x, y=y, status=status, iter=it, rho_updates=rho_updates)


def sqp_cases():
    from oracle import sqp_oracle as S

    st = dict(rho=0.1, sigma=1e-6, alpha=1.6, eps_rel=1e-3, eps_abs=1e-3, max_iter=100, check_termination=25,  # sqp.cpp:16-23 ... see api.sqp_ctor_settings
              adaptive_rho=1, adaptive_rho_tolerance=5, adaptive_rho_interval=25)
    for pid, x0, l0, soc in ((S.CONSTRAINED_ROSENBROCK_2D, [0, 0], [0, 0], 0), (S.SIMPLE_NLP, [2, -1], [1, 1, 1], 1),
                             (S.SIMPLE_NLP, [1.2, 0.1], [0, 0, 0], 0), (S.SIMPLE_QP, [0, 0], [0, 0, 0], 1)):
        tr = S.solve(pid, x0, l0, S.default_settings(second_order_correction=soc), trace_cap=512)["qps"]
        yield "sqp problem %d" % pid, tr, st


def report(name, tr, st, modes=("inv", "refine", "ldl")):
    k = tr["count"]
    for mode in modes:
        worst, flips, worst_i = 0.0, 0, -1
        for i in range(k):
            out = solve_emulated(tr["P"][i], tr["q"][i], tr["A"][i], tr["l"][i], tr["u"][i], st, mode)
            if (out["status"], out["iter"]) != (int(tr["status"][i]), int(tr["iter"][i])):
                flips += 1
                continue
            nx = np.linalg.norm(tr["x"][i])
            e = np.linalg.norm(out["x"] - tr["x"][i]) / max(nx, 1e-300)
            if e > worst:
                worst, worst_i = e, i
        print("%-28s %-7s QPs %3d  status/iter flips %2d  worst x rel err %.2e (QP %d)" % (name, mode, k, flips, worst, worst_i))


if __name__ == "__main__":
    import ctypes

    from oracle import qp_oracle as O
    from sqp_solver_b200 import api  # settings helper only; nothing is launched

    for name, tr, st in sqp_cases():
        s2 = api.sqp_ctor_settings()
        st = {k: getattr(s2, k) for k, _ in s2._fields_}
        report(name, tr, st)


def fuzz_cases():
    """the instances of tests/test_gpu_parity.py::test_randomised_shapes_and_settings (same seeds, same draws)"""
    from sqp_solver_b200.synth import make_batch

    rng = np.random.default_rng(2024)
    shapes = [(1, 0), (3, 0), (64, 128), (63, 127), (33, 65), (17, 33), (9, 17), (8, 16), (64, 1), (1, 128), (2, 3)]
    shapes += [(int(rng.integers(1, 65)), int(rng.integers(0, 129))) for _ in range(14)]
    for case, (n, m) in enumerate(shapes):
        batch = int(rng.integers(2, 7))
        d = make_batch(batch, n, m, seed0=16000 + 10 * case)
        kw = dict(alpha=float(rng.choice([1.0, 1.6, 1.8])), adaptive_rho=int(rng.integers(0, 2)),
                  adaptive_rho_interval=int(rng.choice([7, 25, 50])), check_termination=int(rng.choice([1, 10, 25])),
                  max_iter=int(rng.choice([60, 300])), rho=float(rng.choice([0.05, 0.1, 1.0])),
                  sigma=float(rng.choice([1e-6, 1e-4])), eps_abs=float(rng.choice([1e-3, 1e-5])),
                  eps_rel=float(rng.choice([1e-3, 1e-5])), adaptive_rho_tolerance=float(rng.choice([2.0, 5.0])))
        yield case, n, m, d, kw


if __name__ == "__main__" and "--fuzz" in sys.argv:
    from oracle import qp_oracle as O

    for case, n, m, d, kw in fuzz_cases():
        ref = O.solve_batch(d["P"], d["q"], d["A"], d["l"], d["u"], O.default_settings(**kw))
        stable = ref["diag_min_norms"].min(axis=1) > 1e-9
        st = dict(rho=0.1, sigma=1e-6, alpha=1.0, eps_rel=1e-3, eps_abs=1e-3, max_iter=1000, check_termination=25,
                  adaptive_rho=0, adaptive_rho_tolerance=5, adaptive_rho_interval=25)
        st.update(kw)
        for mode in ("inv", "refine", "ldl"):
            msgs = []
            for i in range(d["batch"]):
                out = solve_emulated(d["P"][i], d["q"][i], d["A"][i], d["l"][i], d["u"][i], st, mode)
                same = (out["status"], out["iter"]) == (int(ref["status"][i]), int(ref["iter"][i]))
                e = np.linalg.norm(out["x"] - ref["x"][i]) / max(np.linalg.norm(ref["x"][i]), 1e-300)
                if not same or e > 1e-6:
                    msgs.append("QP%d%s st=%d %s err=%.1e |x|=%.1e" % (i, "" if stable[i] else "(unstable)", ref["status"][i], "" if same else "FLIP", e, np.linalg.norm(ref["x"][i])))
            if msgs:
                print("case %2d n=%2d m=%3d %-6s %s" % (case, n, m, mode, "; ".join(msgs)))
