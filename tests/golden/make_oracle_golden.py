"""Regenerates tests/golden/oracle_synthetic.json: outputs of the CPU oracle (oracle/qp_oracle.c, gcc -O2,
-ffp-contract=off) on seeded synthetic QPs (sqp_solver_b200/synth.py). The inputs are reproducible from the seeds, so only
the outputs are stored. These are regression pins for the oracle itself (a different compiler or libm on another box must
reproduce them) and committed expected values for the GPU parity tests; they are NOT outputs of the reference binary, which
cannot be built here (Eigen absent).

    python tests/golden/make_oracle_golden.py
"""
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))

from oracle import qp_oracle as O  # noqa: E402
from sqp_solver_b200.synth import densify, make_batch, make_sparse_batch  # noqa: E402

CASES = [
    dict(name="n8_m12_defaults", n=8, m=12, batch=8, seed0=700, settings={}),
    dict(name="n8_m12_sqp_ctor", n=8, m=12, batch=8, seed0=700,
         settings=dict(warm_start=1, check_termination=10, eps_abs=1e-4, eps_rel=1e-4, max_iter=100, adaptive_rho=1,
                       adaptive_rho_interval=50, alpha=1.6)),
    dict(name="n32_m64_defaults", n=32, m=64, batch=6, seed0=0, settings={}),
    dict(name="n64_m128_defaults", n=64, m=128, batch=4, seed0=0, settings={}),
    dict(name="n64_m128_S2", n=64, m=128, batch=4, seed0=0, settings=dict(alpha=1.6, adaptive_rho=1)),
]


# BASELINE config 5's path (sparse A, one pattern per batch: sqp_solver_b200.synth.make_sparse_batch) -- the oracle solves the
# densified problem -- and the float instantiation (oracle QPSolver<float>, instance by instance)
SPARSE_CASES = [
    dict(name="sparse_n100_m150_S2", n=100, m=150, batch=4, density=0.08, seed0=40, settings=dict(alpha=1.6, adaptive_rho=1)),
    dict(name="sparse_n256_m512_S2", n=256, m=512, batch=2, density=0.03, seed0=0, settings=dict(alpha=1.6, adaptive_rho=1)),
    dict(name="sparse_n256_m512_defaults_300", n=256, m=512, batch=2, density=0.03, seed0=0, settings=dict(max_iter=300)),
]
F32_CASES = [
    dict(name="f32_n32_m64_defaults", n=32, m=64, batch=4, seed0=0, settings={}),
    dict(name="f32_n64_m128_defaults", n=64, m=128, batch=3, seed0=0, settings={}),
]


def solve_f32(d, settings):
    n, m, B = d["n"], d["m"], d["batch"]
    r = dict(status=[], iter=[], x=[])
    for i in range(B):
        qp = O.QuadraticProblem(d["P"][i].reshape(n, n, order="F"), d["q"][i], d["A"][i].reshape(m, n, order="F"), d["l"][i], d["u"][i],
                                dtype=np.float32)
        s = O.QPSolver(dtype=np.float32)
        for k, v in settings.items():
            setattr(s.settings(), k, v)
        s.setup(qp)
        s.solve(qp)
        r["status"].append(int(s.info().status))
        r["iter"].append(int(s.info().iter))
        r["x"].append([float(v) for v in s.primal_solution()])
    return r


def main_more():
    out = {"_comment": "oracle outputs for the sparse-A path (densified) and the float instantiation; see make_oracle_golden.py",
           "sparse": [], "f32": []}
    for c in SPARSE_CASES:
        d = make_sparse_batch(c["batch"], c["n"], c["m"], density=c["density"], seed0=c["seed0"])
        r = O.solve_batch(d["P"], d["q"], densify(d), d["l"], d["u"], O.default_settings(**c["settings"]), nthreads=1)
        out["sparse"].append(dict(c, nnz=d["nnz"], status=r["status"].tolist(), iter=r["iter"].tolist(), rho_updates=r["rho_updates"].tolist(),
                                  x=[[float(v) for v in row] for row in r["x"]], y=[[float(v) for v in row] for row in r["y"]]))
    for c in F32_CASES:
        d = make_batch(c["batch"], c["n"], c["m"], seed0=c["seed0"])
        out["f32"].append(dict(c, **solve_f32(d, c["settings"])))
    with open(os.path.join(HERE, "oracle_sparse_f32.json"), "w") as f:
        json.dump(out, f)
    print("wrote", len(out["sparse"]), "sparse and", len(out["f32"]), "f32 cases")


def main():
    out = {"_comment": "oracle outputs on seeded synthetic QPs; see make_oracle_golden.py", "cases": []}
    for c in CASES:
        d = make_batch(c["batch"], c["n"], c["m"], seed0=c["seed0"])
        r = O.solve_batch(d["P"], d["q"], d["A"], d["l"], d["u"], O.default_settings(**c["settings"]), nthreads=1)
        out["cases"].append(dict(c, status=r["status"].tolist(), iter=r["iter"].tolist(), rho_updates=r["rho_updates"].tolist(),
                                 x=[[float(v) for v in row] for row in r["x"]], y=[[float(v) for v in row] for row in r["y"]],
                                 res_prim=r["res_prim"].tolist(), res_dual=r["res_dual"].tolist()))
    with open(os.path.join(HERE, "oracle_synthetic.json"), "w") as f:
        json.dump(out, f)
    print("wrote", len(out["cases"]), "cases")


if __name__ == "__main__":
    main()
    main_more()
