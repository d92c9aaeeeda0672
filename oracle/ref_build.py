"""ctypes loader for oracle/_ref -- test infrastructure, NOT product code.

oracle/_ref/libsqp_ref.so is the REFERENCE'S OWN src/qp.cpp + src/sqp.cpp, compiled unmodified from /root/reference against
oracle/eigen_lite (a stand-in for the absent Eigen dependency; see its header for exactly what is and is not the reference) with
the C entry points of oracle/ref_shim.cpp. It exists where /root/reference exists (the development container, `make -C oracle ref`)
and travels to the GPU box as a prebuilt file; nothing here reads /root/reference at run time.
"""
import ctypes as C
import os
import subprocess

import numpy as np

from . import qp_oracle, sqp_oracle

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB = os.path.join(_HERE, "_ref", "libsqp_ref.so")
TESTS_BIN = os.path.join(_HERE, "_ref", "reference_tests")
REFERENCE = "/root/reference"
_lib = None


def available():
    return os.path.exists(LIB)


def build():
    """Compile oracle/_ref when the reference tree is present (idempotent); a no-op elsewhere (the prebuilt files are used)."""
    if os.path.isdir(os.path.join(REFERENCE, "src")):
        subprocess.check_call(["make", "-s", "-C", _HERE, "ref", "REFERENCE=" + REFERENCE])
    return LIB if available() else None


class Trace(C.Structure):
    _fields_ = [("cap", C.c_int), ("count", C.c_int), ("nx", C.c_int), ("nc", C.c_int), ("x", C.POINTER(C.c_double)),
                ("lam", C.POINTER(C.c_double)), ("qp_x", C.POINTER(C.c_double)), ("qp_solver_iter", C.POINTER(C.c_int)),
                ("qp_status", C.POINTER(C.c_int)), ("qp_iter", C.POINTER(C.c_int))]


def lib():
    global _lib
    if _lib is None:
        if not available():
            raise RuntimeError("oracle/_ref is not built (needs /root/reference: make -C oracle ref)")
        _lib = C.CDLL(LIB)
        _lib.ref_describe.restype = C.c_char_p
        for suf in ("_f64", "_f32"):
            getattr(_lib, "ref_qp_new" + suf).restype = C.c_void_p
            getattr(_lib, "ref_qp_solve_batch" + suf).restype = C.c_int
    return _lib


def _p(a, ct):
    return a.ctypes.data_as(C.POINTER(ct))


def solve_batch(P, q, A, l, u, settings=None, nthreads=0, dtype=np.float64):
    """Fresh `QPSolver<Scalar>` + setup() + solve() per QP with the reference's own code (same arguments and result dict as
    qp_oracle.solve_batch; z is private in the reference and is not returned)."""
    dt = np.dtype(dtype)
    ct = C.c_double if dt == np.float64 else C.c_float
    suf = "_f64" if dt == np.float64 else "_f32"
    P = np.ascontiguousarray(P, dtype=dt)
    B = P.shape[0]
    q = np.ascontiguousarray(q, dtype=dt).reshape(B, -1)
    l = np.ascontiguousarray(l, dtype=dt).reshape(B, -1)
    u = np.ascontiguousarray(u, dtype=dt).reshape(B, -1)
    n, m = q.shape[1], l.shape[1]
    A = np.ascontiguousarray(A, dtype=dt).reshape(B, -1)
    s = settings if settings is not None else qp_oracle.default_settings(dtype=dt)
    out = dict(x=np.zeros((B, n), dt), y=np.zeros((B, m), dt), status=np.zeros(B, np.int32), iter=np.zeros(B, np.int32),
               res_prim=np.zeros(B, dt), res_dual=np.zeros(B, dt), rho_updates=np.zeros(B, np.int32), rho_estimate=np.zeros(B, dt))
    ip = C.POINTER(C.c_int)
    used = getattr(lib(), "ref_qp_solve_batch" + suf)(
        C.byref(s), B, n, m, _p(P, ct), _p(q, ct), _p(A, ct), _p(l, ct), _p(u, ct), _p(out["x"], ct), _p(out["y"], ct),
        out["status"].ctypes.data_as(ip), out["iter"].ctypes.data_as(ip), _p(out["res_prim"], ct), _p(out["res_dual"], ct),
        out["rho_updates"].ctypes.data_as(ip), _p(out["rho_estimate"], ct), int(nthreads))
    out["threads"] = used
    return out


class QPSolver:
    """The reference's qp_solver::QPSolver<double> object (setup / update_qp / solve / primal_solution / dual_solution / info)."""

    def __init__(self):
        self._L = lib()
        self._h = C.c_void_p(self._L.ref_qp_new_f64())
        self._s = qp_oracle.default_settings()
        self.n = self.m = 0

    def __del__(self):
        try:
            self._L.ref_qp_free_f64(self._h)
        except Exception:
            pass

    def settings(self):
        return self._s

    def _call(self, name, qp):
        self.n, self.m = qp.n, qp.m
        d = C.c_double
        self._L.ref_qp_set_settings_f64(self._h, C.byref(self._s))
        getattr(self._L, name)(self._h, qp.n, qp.m, _p(qp.P, d), _p(qp.q, d), _p(qp.A, d), _p(qp.l, d), _p(qp.u, d))

    def setup(self, qp):
        self._call("ref_qp_setup_f64", qp)

    def update_qp(self, qp):
        self._call("ref_qp_update_qp_f64", qp)

    def solve(self, qp):
        self._call("ref_qp_solve_f64", qp)

    def get(self):
        x, y, info = np.zeros(self.n), np.zeros(self.m), qp_oracle.InfoF64()
        self._L.ref_qp_get_f64(self._h, _p(x, C.c_double), _p(y, C.c_double), C.byref(info))
        return x, y, info


def constr_type_init(l, u):
    l = np.ascontiguousarray(l, dtype=np.float64)
    u = np.ascontiguousarray(u, dtype=np.float64)
    out = np.zeros(l.shape[0], np.int32)
    lib().ref_constr_type_init_f64(_p(l, C.c_double), _p(u, C.c_double), l.shape[0], out.ctypes.data_as(C.POINTER(C.c_int)))
    return out


def sqp_solve(prob_id, x0, lambda0, settings=None, n=None, trace_cap=0):
    """sqp::SQP<double>::solve(prob, x0, lambda0) with the reference's own outer loop on a built-in test problem."""
    nx, nc = sqp_oracle.PROBLEM_DIMS.get(prob_id, (n, n))
    x0 = np.ascontiguousarray(x0, dtype=np.float64)
    lambda0 = np.ascontiguousarray(lambda0, dtype=np.float64)
    s = settings or sqp_oracle.default_settings()
    x, lam, info = np.zeros(nx), np.zeros(nc), sqp_oracle.Info()
    dp, ip = C.POINTER(C.c_double), C.POINTER(C.c_int)
    tr, bufs = None, {}
    if trace_cap:
        bufs = dict(x=np.zeros((trace_cap, nx)), lam=np.zeros((trace_cap, nc)), qp_x=np.zeros((trace_cap, nx)),
                    qp_solver_iter=np.zeros(trace_cap, np.int32), qp_status=np.zeros(trace_cap, np.int32), qp_iter=np.zeros(trace_cap, np.int32))
        tr = Trace(cap=trace_cap)
        for k in ("x", "lam", "qp_x"):
            setattr(tr, k, bufs[k].ctypes.data_as(dp))
        for k in ("qp_solver_iter", "qp_status", "qp_iter"):
            setattr(tr, k, bufs[k].ctypes.data_as(ip))
    rc = lib().ref_sqp_solve_builtin(C.c_int(prob_id), C.c_int(nx), C.byref(s), x0.ctypes.data_as(dp), lambda0.ctypes.data_as(dp),
                                     x.ctypes.data_as(dp), lam.ctypes.data_as(dp), C.byref(info), C.byref(tr) if tr else None)
    assert rc == 0
    out = dict(x=x, lam=lam, iter=info.iter, qp_solver_iter=info.qp_solver_iter, status=info.status)
    if tr:
        out["trace"] = {k: v[:tr.count].copy() for k, v in bufs.items()}
    return out


def run_reference_tests():
    """Run the reference's own gtest files (qp_solver_test, sqp_test, bfgs_test) built against eigen_lite + gtest_lite."""
    return subprocess.run([TESTS_BIN], capture_output=True, text=True, timeout=300)
