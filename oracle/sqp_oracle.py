"""ctypes loader for the SQP outer-loop oracle (oracle/sqp_oracle.c) -- test infrastructure only."""
import ctypes as C

import numpy as np

from . import qp_oracle

SOLVED, MAX_ITER_EXCEEDED, INVALID_SETTINGS = range(3)  # sqp.hpp:33
CONSTRAINED_ROSENBROCK_2D, SIMPLE_NLP, SIMPLE_QP, SIMPLE_NLP2, ROSENBROCK_BOX = range(5)
PROBLEM_DIMS = {CONSTRAINED_ROSENBROCK_2D: (2, 2), SIMPLE_NLP: (2, 3), SIMPLE_QP: (2, 3), SIMPLE_NLP2: (2, 1)}


class Settings(C.Structure):  # sqp_settings_t, sqp.hpp:13-31
    _fields_ = [("tau", C.c_double), ("eta", C.c_double), ("rho", C.c_double), ("eps_prim", C.c_double),
                ("eps_dual", C.c_double), ("max_iter", C.c_int), ("line_search_max_iter", C.c_int),
                ("second_order_correction", C.c_int)]


class Info(C.Structure):  # sqp::Info, sqp.hpp:35-60
    _fields_ = [("iter", C.c_int), ("qp_solver_iter", C.c_int), ("status", C.c_int)]


class Trace(C.Structure):
    _fields_ = [("cap", C.c_int), ("count", C.c_int), ("nx", C.c_int), ("nc", C.c_int)] + \
               [(k, C.POINTER(C.c_double)) for k in ("P", "q", "A", "l", "u", "x", "y")] + \
               [("status", C.POINTER(C.c_int)), ("iter", C.POINTER(C.c_int))]


def default_settings(**kw):
    s = Settings()
    qp_oracle.lib().oracle_sqp_default_settings(C.byref(s))
    for k, v in kw.items():
        if not hasattr(s, k):
            raise AttributeError(k)
        setattr(s, k, v)
    return s


def solve(prob_id, x0, lambda0, settings=None, n=None, trace_cap=0):
    """SQP<double>::solve(prob, x0, lambda0) on a built-in test problem. Returns dict(x, lambda, iter,
    qp_solver_iter, status[, qps]) where qps lists every QP subproblem solved (inputs and oracle outputs)."""
    L = qp_oracle.lib()
    nx, nc = PROBLEM_DIMS.get(prob_id, (n, n))
    x0 = np.ascontiguousarray(x0, dtype=np.float64)
    lambda0 = np.ascontiguousarray(lambda0, dtype=np.float64)
    assert x0.shape == (nx,) and lambda0.shape == (nc,)
    s = settings or default_settings()
    x = np.zeros(nx)
    lam = np.zeros(nc)
    info = Info()
    dp = C.POINTER(C.c_double)
    tr = None
    bufs = {}
    if trace_cap:
        tr = Trace(cap=trace_cap)
        shapes = dict(P=nx * nx, q=nx, A=nc * nx, l=nc, u=nc, x=nx, y=nc)
        for k, w in shapes.items():
            bufs[k] = np.zeros((trace_cap, w))
            setattr(tr, k, bufs[k].ctypes.data_as(dp))
        bufs["status"] = np.zeros(trace_cap, np.int32)
        bufs["iter"] = np.zeros(trace_cap, np.int32)
        tr.status = bufs["status"].ctypes.data_as(C.POINTER(C.c_int))
        tr.iter = bufs["iter"].ctypes.data_as(C.POINTER(C.c_int))
    L.oracle_sqp_solve_builtin.restype = C.c_int
    rc = L.oracle_sqp_solve_builtin(C.c_int(prob_id), C.c_int(nx), C.byref(s), x0.ctypes.data_as(dp), lambda0.ctypes.data_as(dp),
                                    x.ctypes.data_as(dp), lam.ctypes.data_as(dp), C.byref(info), C.byref(tr) if tr else None)
    assert rc == 0
    out = dict(x=x, lam=lam, iter=info.iter, qp_solver_iter=info.qp_solver_iter, status=info.status)
    if tr:
        k = tr.count
        out["qps"] = {key: v[:k].copy() for key, v in bufs.items()}
        out["qps"]["count"] = k
    return out


def bfgs_update(B, s, y):
    B = np.asfortranarray(np.array(B, dtype=np.float64))
    s = np.ascontiguousarray(s, dtype=np.float64)
    y = np.ascontiguousarray(y, dtype=np.float64)
    dp = C.POINTER(C.c_double)
    qp_oracle.lib().oracle_bfgs_update(B.ctypes.data_as(dp), C.c_int(B.shape[0]), s.ctypes.data_as(dp), y.ctypes.data_as(dp))
    return B


def is_posdef(H):
    H = np.asfortranarray(np.array(H, dtype=np.float64))
    f = qp_oracle.lib().oracle_is_posdef
    f.restype = C.c_int
    return bool(f(H.ctypes.data_as(C.POINTER(C.c_double)), C.c_int(H.shape[0])))


def solve_batch(prob_id, x0, lambda0, settings=None, nthreads=0):
    """B independent SQP<double>::solve runs (one built-in problem, B starts) over the host cores. x0: [B, nx], lambda0: [B, nc].
    Returns dict(x, lam, iter, qp_solver_iter, status, threads)."""
    L = qp_oracle.lib()
    x0 = np.ascontiguousarray(x0, dtype=np.float64)
    lambda0 = np.ascontiguousarray(lambda0, dtype=np.float64)
    B, nx = x0.shape
    nc = lambda0.shape[1]
    s = settings or default_settings()
    x, lam = np.zeros((B, nx)), np.zeros((B, nc))
    info = (Info * B)()
    dp = C.POINTER(C.c_double)
    L.oracle_sqp_solve_batch.restype = C.c_int
    used = L.oracle_sqp_solve_batch(C.c_int(prob_id), C.c_int(nx), C.byref(s), C.c_int(B), x0.ctypes.data_as(dp), lambda0.ctypes.data_as(dp),
                                    x.ctypes.data_as(dp), lam.ctypes.data_as(dp), info, C.c_int(nthreads))
    assert used >= 1
    return dict(x=x, lam=lam, iter=np.array([i.iter for i in info]), qp_solver_iter=np.array([i.qp_solver_iter for i in info]),
                status=np.array([i.status for i in info]), threads=used)
