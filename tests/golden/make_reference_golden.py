"""Regenerates tests/golden/reference_outputs.json: outputs of THE REFERENCE ITSELF run in the development container --
oracle/_ref = /root/reference/src/qp.cpp + src/sqp.cpp compiled unmodified against oracle/eigen_lite (the stand-in for its absent
Eigen dependency; `make -C oracle ref`). Inputs are reproducible from the seeds (sqp_solver_b200/synth.py), so only outputs are
stored. Used by
  * tests/test_reference_build.py   (CPU): the C oracle restatement must reproduce every number BIT FOR BIT,
  * tests/test_reference_golden_gpu.py (GPU box, where /root/reference does not exist): the CUDA path against the reference's outputs.

    python tests/golden/make_reference_golden.py
"""
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))

from oracle import qp_oracle as O, ref_build as R, sqp_oracle as S  # noqa: E402
from sqp_solver_b200.synth import make_batch  # noqa: E402

SQP_CTOR = dict(warm_start=1, check_termination=10, eps_abs=1e-4, eps_rel=1e-4, max_iter=100, adaptive_rho=1, adaptive_rho_interval=50, alpha=1.6)
QP_CASES = [
    dict(name="n2_m3_defaults", n=2, m=3, batch=16, seed0=900, settings={}),
    dict(name="n2_m2_sqp_ctor", n=2, m=2, batch=16, seed0=910, settings=SQP_CTOR),
    dict(name="n5_m7_S2", n=5, m=7, batch=12, seed0=920, settings=dict(alpha=1.6, adaptive_rho=1)),
    dict(name="n12_m20_sqp_ctor", n=12, m=20, batch=8, seed0=930, settings=SQP_CTOR),
    dict(name="n32_m64_defaults", n=32, m=64, batch=6, seed0=0, settings={}),
    dict(name="n32_m64_S2", n=32, m=64, batch=6, seed0=0, settings=dict(alpha=1.6, adaptive_rho=1)),
    dict(name="n64_m128_defaults", n=64, m=128, batch=4, seed0=0, settings={}),
    dict(name="n64_m128_S2", n=64, m=128, batch=4, seed0=0, settings=dict(alpha=1.6, adaptive_rho=1)),
    dict(name="n64_m128_odd", n=64, m=128, batch=3, seed0=5,
         settings=dict(alpha=1.8, adaptive_rho=1, adaptive_rho_interval=7, check_termination=3, max_iter=300, adaptive_rho_tolerance=2.0)),
    dict(name="n80_m100_S2", n=80, m=100, batch=2, seed0=940, settings=dict(alpha=1.6, adaptive_rho=1)),
]
SQP_CASES = [
    dict(name="ConstrainedRosenbrock2D", pid=S.CONSTRAINED_ROSENBROCK_2D, x0=[0, 0], l0=[0, 0], soc=0),
    dict(name="SimpleNLP_feasible_SOC", pid=S.SIMPLE_NLP, x0=[1.2, 0.1], l0=[0, 0, 0], soc=1),
    dict(name="SimpleNLP_infeasible_SOC", pid=S.SIMPLE_NLP, x0=[2, -1], l0=[1, 1, 1], soc=1),
    dict(name="SimpleQP_as_NLP_SOC", pid=S.SIMPLE_QP, x0=[0, 0], l0=[0, 0, 0], soc=1),
    dict(name="SimpleNLP2", pid=S.SIMPLE_NLP2, x0=[1.2, 0.1], l0=[0], soc=0),
    dict(name="RosenbrockBox2", pid=S.ROSENBROCK_BOX, x0=[0, 0], l0=[0, 0], soc=0),
    dict(name="RosenbrockBox3", pid=S.ROSENBROCK_BOX, x0=[0, 0, 0], l0=[0, 0, 0], soc=0),
]


def main():
    R.build()
    out = dict(how=R.lib().ref_describe().decode(), qp=[], qp_f32=[], sqp=[])
    for c in QP_CASES:
        d = make_batch(c["batch"], c["n"], c["m"], seed0=c["seed0"])
        r = R.solve_batch(d["P"], d["q"], d["A"], d["l"], d["u"], O.default_settings(**c["settings"]), nthreads=1)
        out["qp"].append(dict(c, **{k: r[k].tolist() for k in ("status", "iter", "rho_updates", "x", "y", "res_prim", "res_dual", "rho_estimate")}))
    for c in (QP_CASES[0], QP_CASES[4], QP_CASES[6]):
        d = make_batch(c["batch"], c["n"], c["m"], seed0=c["seed0"])
        r = R.solve_batch(d["P"], d["q"], d["A"], d["l"], d["u"], O.default_settings(dtype=np.float32, **c["settings"]), nthreads=1, dtype=np.float32)
        out["qp_f32"].append(dict(c, **{k: r[k].astype(np.float64).tolist() if r[k].dtype == np.float32 else r[k].tolist()
                                        for k in ("status", "iter", "rho_updates", "x", "y")}))
    for c in SQP_CASES:
        r = R.sqp_solve(c["pid"], c["x0"], c["l0"], S.default_settings(second_order_correction=c["soc"]), n=len(c["x0"]), trace_cap=128)
        out["sqp"].append(dict(c, iter=r["iter"], qp_solver_iter=r["qp_solver_iter"], status=r["status"], x=r["x"].tolist(), lam=r["lam"].tolist(),
                               trace_qp_solver_iter=r["trace"]["qp_solver_iter"].tolist(), trace_x=r["trace"]["x"].tolist()))
    path = os.path.join(HERE, "reference_outputs.json")
    with open(path, "w") as f:
        json.dump(out, f)
    print("wrote", path, os.path.getsize(path), "bytes")


if __name__ == "__main__":
    main()
