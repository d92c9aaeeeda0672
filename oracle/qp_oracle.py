"""ctypes loader for the C oracle (oracle/qp_oracle.c) -- test infrastructure, NOT product code.

Mirrors the reference's `qp_solver::QPSolver<Scalar>` object API (include/solvers/qp.hpp:148-173):
setup / update_qp / solve / primal_solution / dual_solution / settings / info, plus the static
`constr_type_init`, so the tests read like /root/reference/tests/qp_solver_test.cpp.
"""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = os.path.join(_HERE, "_build", "liboracle.so")

SOLVED, MAX_ITER_EXCEEDED, UNSOLVED, NUMERICAL_ISSUES, UNINITIALIZED = range(5)  # qp.hpp:70
INEQUALITY_CONSTRAINT, EQUALITY_CONSTRAINT, LOOSE_BOUNDS = range(3)  # qp.hpp:134


def _native_path():
    import tempfile

    return os.path.join(tempfile.gettempdir(), "sqpb200_liboracle_native_%d.so" % os.getuid())


def build(native=False):
    """Compile the oracle with the committed recipe (oracle/Makefile). Idempotent.
    native=True builds the -O3 -march=native variant on THIS machine into the temp dir (never into the tree: a
    -march=native binary must not travel to a box with a different CPU)."""
    if native:
        out = _native_path()
        subprocess.check_call(["make", "-s", "-B", "-C", _HERE, "native", "NATIVE_OUT=" + out])
        return out
    subprocess.check_call(["make", "-s", "-C", _HERE])
    return _LIB


def _mk_structs(ct):
    class Settings(C.Structure):
        _fields_ = [("rho", ct), ("sigma", ct), ("alpha", ct), ("eps_rel", ct), ("eps_abs", ct),
                    ("max_iter", C.c_int), ("check_termination", C.c_int), ("warm_start", C.c_int),
                    ("adaptive_rho", C.c_int), ("adaptive_rho_tolerance", ct),
                    ("adaptive_rho_interval", C.c_int), ("verbose", C.c_int)]

    class Info(C.Structure):
        _fields_ = [("status", C.c_int), ("iter", C.c_int), ("rho_updates", C.c_int),
                    ("rho_estimate", ct), ("res_prim", ct), ("res_dual", ct)]

    return Settings, Info


SettingsF64, InfoF64 = _mk_structs(C.c_double)
SettingsF32, InfoF32 = _mk_structs(C.c_float)

_libs = {}


def lib(native=False):
    key = bool(native)
    if key not in _libs:
        path = _native_path() if native else _LIB
        if native or not os.path.exists(path):
            build(native)
        L = C.CDLL(path)
        for suf, ct, S, I in (("_f64", C.c_double, SettingsF64, InfoF64), ("_f32", C.c_float, SettingsF32, InfoF32)):
            P = C.POINTER(ct)
            getattr(L, "oracle_qp_new" + suf).restype = C.c_void_p
            getattr(L, "oracle_qp_free" + suf).argtypes = [C.c_void_p]
            getattr(L, "oracle_qp_settings" + suf).restype = C.POINTER(S)
            getattr(L, "oracle_qp_settings" + suf).argtypes = [C.c_void_p]
            getattr(L, "oracle_qp_info" + suf).restype = C.POINTER(I)
            getattr(L, "oracle_qp_info" + suf).argtypes = [C.c_void_p]
            for nm in ("primal", "dual", "z"):
                getattr(L, "oracle_qp_%s%s" % (nm, suf)).restype = P
                getattr(L, "oracle_qp_%s%s" % (nm, suf)).argtypes = [C.c_void_p]
            getattr(L, "oracle_qp_rho" + suf).restype = ct
            getattr(L, "oracle_qp_rho" + suf).argtypes = [C.c_void_p]
            getattr(L, "oracle_qp_setup" + suf).argtypes = [C.c_void_p, C.c_int, C.c_int, P, P, P, P, P]
            getattr(L, "oracle_qp_update_qp" + suf).argtypes = [C.c_void_p, P, P, P, P, P]
            getattr(L, "oracle_qp_solve" + suf).argtypes = [C.c_void_p, P, P, P, P, P]
            getattr(L, "oracle_constr_type_init" + suf).argtypes = [P, P, C.c_int, C.POINTER(C.c_int)]
            getattr(L, "oracle_qp_kkt_solve" + suf).argtypes = [C.c_void_p, P, P]
            getattr(L, "oracle_qp_ldlt_dump" + suf).argtypes = [C.c_void_p, P, C.POINTER(C.c_int)]
            fb = getattr(L, "oracle_qp_solve_batch" + suf)
            fb.restype = C.c_int
            IP = C.POINTER(C.c_int)
            fb.argtypes = [C.POINTER(S), C.c_int, C.c_int, C.c_int, P, P, P, P, P, P, P, P, IP, IP, P, P, IP, P, C.c_int, C.POINTER(C.c_double)]
        L.oracle_num_procs.restype = C.c_int
        _libs[key] = L
    return _libs[key]


def _ptr(a, ct):
    return a.ctypes.data_as(C.POINTER(ct))


class QuadraticProblem:
    """Counterpart of qp_solver::QuadraticProblem (qp.hpp:19-34). Matrices are stored
    column-major, bit-compatible with Eigen::MatrixXd::data()."""

    def __init__(self, P, q, A, l, u, dtype=np.float64):
        self.dtype = np.dtype(dtype)
        self.P = np.asfortranarray(np.asarray(P, dtype=dtype))
        self.A = np.asfortranarray(np.asarray(A, dtype=dtype))
        self.q = np.ascontiguousarray(np.asarray(q, dtype=dtype))
        self.l = np.ascontiguousarray(np.asarray(l, dtype=dtype))
        self.u = np.ascontiguousarray(np.asarray(u, dtype=dtype))
        self.n = self.P.shape[0]
        self.m = self.A.shape[0]


class QPSolver:
    """Counterpart of qp_solver::QPSolver<Scalar> (qp.hpp:113-250) backed by the C oracle."""

    def __init__(self, dtype=np.float64):
        self.dtype = np.dtype(dtype)
        self._suf = "_f64" if self.dtype == np.float64 else "_f32"
        self._ct = C.c_double if self.dtype == np.float64 else C.c_float
        self._L = lib()
        self._h = C.c_void_p(getattr(self._L, "oracle_qp_new" + self._suf)())
        self.n = self.m = 0

    def __del__(self):
        try:
            getattr(self._L, "oracle_qp_free" + self._suf)(self._h)
        except Exception:
            pass

    def _f(self, name):
        return getattr(self._L, name + self._suf)

    def settings(self):
        return self._f("oracle_qp_settings")(self._h).contents

    def info(self):
        return self._f("oracle_qp_info")(self._h).contents

    def _args(self, qp):
        ct = self._ct
        return [_ptr(qp.P, ct), _ptr(qp.q, ct), _ptr(qp.A, ct), _ptr(qp.l, ct), _ptr(qp.u, ct)]

    def setup(self, qp):
        self.n, self.m = qp.n, qp.m
        self._f("oracle_qp_setup")(self._h, qp.n, qp.m, *self._args(qp))

    def update_qp(self, qp):
        self._f("oracle_qp_update_qp")(self._h, *self._args(qp))

    def solve(self, qp):
        self._f("oracle_qp_solve")(self._h, *self._args(qp))

    def primal_solution(self):
        return np.ctypeslib.as_array(self._f("oracle_qp_primal")(self._h), shape=(self.n,)).copy()

    def dual_solution(self):
        return np.ctypeslib.as_array(self._f("oracle_qp_dual")(self._h), shape=(self.m,)).copy()

    def set_iterates(self, x=None, y=None, z=None):
        """Overwrite the solver's iterates in place. The reference exposes x and y through the non-const
        `Vector &primal_solution()` / `dual_solution()` accessors (qp.hpp:160,163); z is written here only to
        mirror sqpb200_qp_batch_set_iterates in the warm-start tests."""
        for name, val, ln in (("primal", x, self.n), ("dual", y, self.m), ("z", z, self.m)):
            if val is not None:
                dst = np.ctypeslib.as_array(self._f("oracle_qp_" + name)(self._h), shape=(ln,))
                dst[:] = np.asarray(val, dtype=self.dtype)

    def z(self):
        return np.ctypeslib.as_array(self._f("oracle_qp_z")(self._h), shape=(self.m,)).copy()

    def rho(self):
        return float(self._f("oracle_qp_rho")(self._h))

    def kkt_solve(self, rhs):
        rhs = np.ascontiguousarray(rhs, dtype=self.dtype)
        out = np.empty_like(rhs)
        self._f("oracle_qp_kkt_solve")(self._h, _ptr(rhs, self._ct), _ptr(out, self._ct))
        return out

    def ldlt_dump(self):
        N = self.n + self.m
        D = np.empty(N, dtype=self.dtype)
        T = np.empty(N, dtype=np.int32)
        self._f("oracle_qp_ldlt_dump")(self._h, _ptr(D, self._ct), T.ctypes.data_as(C.POINTER(C.c_int)))
        return D, T

    @staticmethod
    def constr_type_init(l, u, dtype=np.float64):
        dt = np.dtype(dtype)
        ct = C.c_double if dt == np.float64 else C.c_float
        l = np.ascontiguousarray(l, dtype=dt)
        u = np.ascontiguousarray(u, dtype=dt)
        out = np.empty(l.shape[0], dtype=np.int32)
        f = getattr(lib(), "oracle_constr_type_init" + ("_f64" if dt == np.float64 else "_f32"))
        f(_ptr(l, ct), _ptr(u, ct), l.shape[0], out.ctypes.data_as(C.POINTER(C.c_int)))
        return out


def default_settings(dtype=np.float64, **kw):
    S = SettingsF64 if np.dtype(dtype) == np.float64 else SettingsF32
    s = S(rho=1e-1, sigma=1e-6, alpha=1.0, eps_rel=1e-3, eps_abs=1e-3, max_iter=1000, check_termination=25,
          warm_start=0, adaptive_rho=0, adaptive_rho_tolerance=5, adaptive_rho_interval=25, verbose=0)
    for k, v in kw.items():
        if not hasattr(s, k):
            raise AttributeError(k)
        setattr(s, k, v)
    return s


def solve_batch(P, q, A, l, u, settings=None, nthreads=0, native=False):
    """Fresh setup()+solve() per QP over a batch (reference pattern at src/sqp.cpp:221-222).

    P: [B, n*n] (each column-major), q: [B, n], A: [B, m*n] (column-major), l, u: [B, m].
    Returns a dict of arrays. `nthreads` <= 0 uses every host core (OpenMP dynamic schedule)."""
    P = np.ascontiguousarray(P, dtype=np.float64)
    B = P.shape[0]
    q = np.ascontiguousarray(q, dtype=np.float64).reshape(B, -1)
    n = q.shape[1]
    l = np.ascontiguousarray(l, dtype=np.float64).reshape(B, -1)
    u = np.ascontiguousarray(u, dtype=np.float64).reshape(B, -1)
    m = l.shape[1]
    A = np.ascontiguousarray(A, dtype=np.float64).reshape(B, -1)
    assert P.reshape(B, -1).shape[1] == n * n and A.shape[1] == m * n
    s = settings if settings is not None else default_settings()
    out = dict(x=np.zeros((B, n)), y=np.zeros((B, m)), z=np.zeros((B, m)),
               status=np.zeros(B, np.int32), iter=np.zeros(B, np.int32), res_prim=np.zeros(B), res_dual=np.zeros(B),
               rho_updates=np.zeros(B, np.int32), rho_estimate=np.zeros(B), diag_min_norms=np.zeros((B, 2)))
    d, ip = C.c_double, C.POINTER(C.c_int)
    used = lib(native).oracle_qp_solve_batch_f64(
        C.byref(s), B, n, m, _ptr(P, d), _ptr(q, d), _ptr(A, d), _ptr(l, d), _ptr(u, d),
        _ptr(out["x"], d), _ptr(out["y"], d), _ptr(out["z"], d),
        out["status"].ctypes.data_as(ip), out["iter"].ctypes.data_as(ip), _ptr(out["res_prim"], d),
        _ptr(out["res_dual"], d), out["rho_updates"].ctypes.data_as(ip), _ptr(out["rho_estimate"], d), int(nthreads),
        _ptr(out["diag_min_norms"], d))
    out["threads"] = used
    return out


def num_procs():
    return int(lib().oracle_num_procs())
