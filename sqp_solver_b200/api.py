"""ctypes binding of the C-ABI in include/sqp_b200_qp.h -- the Python face of the drop-in boundary.

The classes mirror the reference's `qp_solver::QPSolver<Scalar>` (include/solvers/qp.hpp:148-173)
batched over B instances: setup / update_qp / solve / primal_solution / dual_solution / info.
There is NO CPU fallback: loading fails loudly when libsqp_b200.so is missing, and creating a
Context fails when no sm_100 device is present.

Arrays may be numpy (host pointers; the library stages and synchronises) or torch CUDA tensors
(device pointers; asynchronous on the given / current torch stream). Matrices are batch-major
with each problem column-major: P[B, n*n], A[B, m*n] (see synth.py).
"""
import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libsqp_b200.so")

SOLVED, MAX_ITER_EXCEEDED, UNSOLVED, NUMERICAL_ISSUES, UNINITIALIZED = range(5)  # qp.hpp:70
INEQUALITY_CONSTRAINT, EQUALITY_CONSTRAINT, LOOSE_BOUNDS = range(3)  # qp.hpp:134
STATUS_NAMES = ["SOLVED", "MAX_ITER_EXCEEDED", "UNSOLVED", "NUMERICAL_ISSUES", "UNINITIALIZED"]
HOST_PTRS, DEVICE_PTRS, HOST_ASYNC = 0, 1, 2
OPT_KERNEL, OPT_H2D_CHUNKS, OPT_CTAS_PER_SM, OPT_TILE_WARPS, OPT_SLICE_ITERS = 1, 2, 3, 4, 5
KERNEL_AUTO, KERNEL_GENERIC, KERNEL_TILE, KERNEL_BLOCK, KERNEL_CLUSTER, KERNEL_SMALL = 0, 1, 2, 3, 4, 5
KEEP_FACTOR, REUSE_FACTOR = 1, 2
SPARSE_CSC, SPARSE_CSR = 0, 1

# every symbol include/sqp_b200_qp.h declares (checked by tests/test_abi_symbols.py)
ABI_SYMBOLS = [
    "sqpb200_abi_version", "sqpb200_ctx_create", "sqpb200_ctx_destroy", "sqpb200_ctx_set_option", "sqpb200_last_error",
    "sqpb200_device_query", "sqpb200_launch_count", "sqpb200_last_kernel", "sqpb200_qp_default_settings",
    "sqpb200_constr_type_init", "sqpb200_qp_batch_create", "sqpb200_qp_batch_destroy", "sqpb200_qp_batch_setup",
    "sqpb200_qp_batch_update_qp", "sqpb200_qp_batch_solve", "sqpb200_qp_batch_setup_solve",
    "sqpb200_qp_batch_setup_solve_opts", "sqpb200_qp_batch_setup_solve_sparse", "sqpb200_qp_batch_set_precision", "sqpb200_dev_alloc", "sqpb200_dev_free", "sqpb200_dev_copy",
    "sqpb200_ipc_export", "sqpb200_ipc_import", "sqpb200_ipc_release", "sqpb200_qp_batch_get",
    "sqpb200_qp_batch_set_iterates", "sqpb200_qp_batch_device_view", "sqpb200_qp_batch_total_iters",
    "sqpb200_qp_solve_batch", "sqpb200_measure_fp64_peak", "sqpb200_qp_batch_setup_solve_to", "sqpb200_host_alloc", "sqpb200_host_free", "sqpb200_stream_create", "sqpb200_stream_destroy", "sqpb200_stream_sync", "sqpb200_qp_batch_setup_sparse",
    "sqpb200_qp_batch_update_qp_sparse", "sqpb200_qp_batch_solve_sparse", "sqpb200_staging_chunk_bounds",
]


class Settings(C.Structure):
    """sqpb200_qp_settings == QPSolverSettings<double> (qp.hpp:36-53)."""
    _fields_ = [("rho", C.c_double), ("sigma", C.c_double), ("alpha", C.c_double), ("eps_rel", C.c_double),
                ("eps_abs", C.c_double), ("max_iter", C.c_int), ("check_termination", C.c_int), ("warm_start", C.c_int),
                ("adaptive_rho", C.c_int), ("adaptive_rho_tolerance", C.c_double), ("adaptive_rho_interval", C.c_int),
                ("verbose", C.c_int)]


class DeviceView(C.Structure):
    _fields_ = [(k, C.c_void_p) for k in ("x", "y", "z", "status", "iter", "rho_updates", "rho_estimate", "res_prim",
                                          "res_dual", "rho", "total_iters")]


class SolverError(RuntimeError):
    pass


_lib = None


def load_library(path=None):
    """Load libsqp_b200.so. Raises (never falls back) when the CUDA extension is not built."""
    global _lib
    if _lib is not None and path is None:
        return _lib
    path = path or LIB_PATH
    if not os.path.exists(path):
        raise SolverError("%s is missing: run `python -m sqp_solver_b200.build` (there is no CPU fallback)" % path)
    L = C.CDLL(path)
    dp, ip, vp = C.c_void_p, C.c_void_p, C.c_void_p
    L.sqpb200_abi_version.restype = C.c_int
    L.sqpb200_ctx_create.argtypes = [C.c_int, C.POINTER(vp)]
    L.sqpb200_ctx_destroy.argtypes = [vp]
    L.sqpb200_ctx_set_option.argtypes = [vp, C.c_int, C.c_int]
    L.sqpb200_last_error.argtypes = [vp]
    L.sqpb200_last_error.restype = C.c_char_p
    L.sqpb200_device_query.argtypes = [vp, C.POINTER(C.c_int), C.POINTER(C.c_int), C.POINTER(C.c_int), C.POINTER(C.c_int),
                                       C.POINTER(C.c_size_t)]
    L.sqpb200_measure_fp64_peak.argtypes = [vp, C.POINTER(C.c_double), C.POINTER(C.c_double), C.POINTER(C.c_double)]
    L.sqpb200_launch_count.argtypes = [vp]
    L.sqpb200_launch_count.restype = C.c_longlong
    L.sqpb200_last_kernel.argtypes = [vp]
    L.sqpb200_last_kernel.restype = C.c_char_p
    L.sqpb200_qp_default_settings.argtypes = [C.POINTER(Settings)]
    L.sqpb200_qp_default_settings.restype = None
    L.sqpb200_constr_type_init.argtypes = [dp, dp, C.c_int, ip]
    L.sqpb200_qp_batch_create.argtypes = [vp, C.c_int, C.c_int, C.c_int, C.POINTER(vp)]
    L.sqpb200_qp_batch_destroy.argtypes = [vp]
    for name in ("setup", "update_qp", "solve", "setup_solve"):
        getattr(L, "sqpb200_qp_batch_" + name).argtypes = [vp, C.POINTER(Settings), C.c_int, dp, dp, dp, dp, dp, C.c_uint, vp]
    L.sqpb200_qp_batch_setup_solve_opts.argtypes = [vp, C.POINTER(Settings), C.c_int, dp, dp, dp, dp, dp, C.c_uint, vp, C.c_uint]
    for name in ("setup_solve_sparse", "setup_sparse", "update_qp_sparse", "solve_sparse"):
        getattr(L, "sqpb200_qp_batch_" + name).argtypes = [vp, C.POINTER(Settings), C.c_int, dp, dp, dp, ip, ip, C.c_int, C.c_int, dp, dp,
                                                           C.c_uint, vp]
    L.sqpb200_qp_batch_setup_solve_to.argtypes = [vp, C.POINTER(Settings), C.c_int, dp, dp, dp, dp, dp, dp, dp, dp, ip, ip, ip, dp, dp, dp, vp]
    L.sqpb200_qp_batch_set_precision.argtypes = [vp, C.c_int]
    L.sqpb200_dev_alloc.argtypes = [vp, C.c_size_t, C.POINTER(C.c_void_p)]
    L.sqpb200_dev_free.argtypes = [vp, vp]
    L.sqpb200_stream_create.argtypes = [vp, C.POINTER(C.c_void_p)]
    L.sqpb200_stream_destroy.argtypes = [vp, vp]
    L.sqpb200_stream_sync.argtypes = [vp, vp]
    L.sqpb200_host_alloc.argtypes = [vp, C.c_size_t, C.POINTER(C.c_void_p)]
    L.sqpb200_host_free.argtypes = [vp, vp]
    L.sqpb200_dev_copy.argtypes = [vp, vp, vp, C.c_size_t, vp]
    L.sqpb200_ipc_export.argtypes = [vp, vp, C.c_char_p]
    L.sqpb200_ipc_import.argtypes = [vp, C.c_char_p, C.POINTER(C.c_void_p)]
    L.sqpb200_ipc_release.argtypes = [vp, vp]
    L.sqpb200_qp_batch_get.argtypes = [vp, C.c_int, dp, dp, dp, ip, ip, ip, dp, dp, dp, C.c_uint, vp]
    L.sqpb200_qp_batch_set_iterates.argtypes = [vp, C.c_int, dp, dp, dp, C.c_uint, vp]
    L.sqpb200_qp_batch_device_view.argtypes = [vp, C.POINTER(DeviceView)]
    L.sqpb200_qp_batch_total_iters.argtypes = [vp, C.POINTER(C.c_longlong), vp]
    L.sqpb200_qp_solve_batch.argtypes = [vp, C.POINTER(Settings), C.c_int, C.c_int, C.c_int, dp, dp, dp, dp, dp, dp, dp, dp,
                                         ip, ip, ip, dp, dp, dp, C.c_uint, vp]
    if path == LIB_PATH:
        _lib = L
    return L


def default_settings(**kw):
    s = Settings()
    load_library().sqpb200_qp_default_settings(C.byref(s))
    for k, v in kw.items():
        if not hasattr(s, k):
            raise AttributeError(k)
        setattr(s, k, v)
    return s


def sqp_ctor_settings():
    """The QP settings SQP<T>::SQP() installs (reference src/sqp.cpp:16-23)."""
    return default_settings(warm_start=1, check_termination=10, eps_abs=1e-4, eps_rel=1e-4, max_iter=100, adaptive_rho=1,
                            adaptive_rho_interval=50, alpha=1.6)


def constr_type_init(l, u):
    """static QPSolver::constr_type_init (qp.cpp:283-294); host side, needs no GPU."""
    l = np.ascontiguousarray(l, dtype=np.float64)
    u = np.ascontiguousarray(u, dtype=np.float64)
    out = np.empty(l.shape[0], dtype=np.int32)
    rc = load_library().sqpb200_constr_type_init(l.ctypes.data, u.ctypes.data, l.shape[0], out.ctypes.data)
    if rc:
        raise SolverError("sqpb200_constr_type_init failed (%d)" % rc)
    return out


def _is_torch(a):
    return type(a).__module__.startswith("torch")


class DevPtr:
    """A raw device address (this GPU's memory from Context.dev_alloc, or a peer GPU's memory imported with Context.ipc_import)
    usable wherever the API takes a CUDA tensor. `offset(elements)` advances by whole elements of `dtype`."""

    def __init__(self, ptr, dtype=np.float64):
        self.ptr, self.dtype = int(ptr), np.dtype(dtype)

    def offset(self, elements):
        return DevPtr(self.ptr + int(elements) * self.dtype.itemsize, self.dtype)


def _ptr_and_space(a, dtype, what):
    """(address, DEVICE_PTRS|HOST_PTRS) of a contiguous numpy array or torch tensor."""
    if a is None:
        return None, None
    if isinstance(a, DevPtr):
        if a.dtype != np.dtype(dtype):
            raise SolverError("%s: DevPtr of dtype %s, need %s" % (what, a.dtype, np.dtype(dtype)))
        return a.ptr, DEVICE_PTRS
    if _is_torch(a):
        import torch

        want = torch.float64 if dtype == np.float64 else torch.int32
        if a.dtype != want or not a.is_contiguous():
            raise SolverError("%s: need a contiguous %s tensor" % (what, want))
        return a.data_ptr(), (DEVICE_PTRS if a.is_cuda else HOST_PTRS)
    if not isinstance(a, np.ndarray) or a.dtype != dtype or not a.flags["C_CONTIGUOUS"]:
        raise SolverError("%s: need a C-contiguous numpy array of dtype %s" % (what, np.dtype(dtype)))
    return a.ctypes.data, HOST_PTRS


class Context:
    """One per (process, GPU): sqpb200_ctx."""

    def __init__(self, device=0):
        self._L = load_library()
        h = C.c_void_p()
        rc = self._L.sqpb200_ctx_create(int(device), C.byref(h))
        if rc:
            raise SolverError("sqpb200_ctx_create(%d) failed: %s" % (device, self._L.sqpb200_last_error(None).decode()))
        self._h = h
        self.device = int(device)

    def close(self):
        if getattr(self, "_h", None):
            self._L.sqpb200_ctx_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _check(self, rc, what):
        if rc:
            raise SolverError("%s failed (%d): %s" % (what, rc, self._L.sqpb200_last_error(self._h).decode()))

    def set_option(self, option, value):
        self._check(self._L.sqpb200_ctx_set_option(self._h, option, value), "set_option")

    # ---- raw device buffers and CUDA IPC (peer memory over NVLink; see include/sqp_b200_qp.h) ----
    def dev_alloc(self, nbytes, dtype=np.float64):
        p = C.c_void_p()
        self._check(self._L.sqpb200_dev_alloc(self._h, int(nbytes), C.byref(p)), "dev_alloc")
        return DevPtr(p.value, dtype)

    def dev_free(self, dp):
        self._check(self._L.sqpb200_dev_free(self._h, C.c_void_p(dp.ptr)), "dev_free")

    def dev_copy(self, dst, src, nbytes, stream=None):
        """dst/src: DevPtr, CUDA tensor or contiguous numpy array; asynchronous on `stream`."""
        def addr(a):
            if isinstance(a, DevPtr):
                return a.ptr
            return a.data_ptr() if _is_torch(a) else a.ctypes.data
        self._check(self._L.sqpb200_dev_copy(self._h, C.c_void_p(addr(dst)), C.c_void_p(addr(src)), int(nbytes), C.c_void_p(stream or 0)), "dev_copy")

    def ipc_export(self, dp):
        buf = C.create_string_buffer(64)
        self._check(self._L.sqpb200_ipc_export(self._h, C.c_void_p(dp.ptr), buf), "ipc_export")
        return buf.raw

    def ipc_import(self, handle, dtype=np.float64):
        p = C.c_void_p()
        self._check(self._L.sqpb200_ipc_import(self._h, handle, C.byref(p)), "ipc_import")
        return DevPtr(p.value, dtype)

    def ipc_release(self, dp):
        self._check(self._L.sqpb200_ipc_release(self._h, C.c_void_p(dp.ptr)), "ipc_release")

    def device_query(self):
        dev, sm, ma, mi, sz = C.c_int(), C.c_int(), C.c_int(), C.c_int(), C.c_size_t()
        self._check(self._L.sqpb200_device_query(self._h, C.byref(dev), C.byref(sm), C.byref(ma), C.byref(mi), C.byref(sz)),
                    "device_query")
        return dict(device=dev.value, sm_count=sm.value, cc=(ma.value, mi.value), smem_per_block_optin=sz.value)

    def measure_fp64_peak(self):
        """(TFLOP/s, seconds, DFMA count) of a saturating fp64 FMA kernel on this device, measured now."""
        t, s_, c = C.c_double(), C.c_double(), C.c_double()
        self._check(self._L.sqpb200_measure_fp64_peak(self._h, C.byref(t), C.byref(s_), C.byref(c)), "measure_fp64_peak")
        return t.value, s_.value, c.value

    @property
    def launch_count(self):
        return int(self._L.sqpb200_launch_count(self._h))

    @property
    def last_kernel(self):
        return self._L.sqpb200_last_kernel(self._h).decode()


class QPBatch:
    """B independent `QPSolver<double>` instances of identical (n, m): sqpb200_qp_batch."""

    def __init__(self, ctx, batch, n, m):
        self.ctx, self.batch, self.n, self.m = ctx, int(batch), int(n), int(m)
        self._L = ctx._L
        h = C.c_void_p()
        ctx._check(self._L.sqpb200_qp_batch_create(ctx._h, self.batch, self.n, self.m, C.byref(h)), "qp_batch_create")
        self._h = h
        self.settings = default_settings()  # QPSolver::settings()

    def set_precision(self, fp32):
        """Compute in fp32 (QPSolver<float>, qp.cpp:386) or fp64 (default). The arrays of the interface stay float64."""
        self.ctx._check(self._L.sqpb200_qp_batch_set_precision(self._h, 1 if fp32 else 0), "set_precision")

    def close(self):
        if getattr(self, "_h", None) and getattr(self.ctx, "_h", None):
            self._L.sqpb200_qp_batch_destroy(self._h)
        self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _call(self, name, P, q, A, l, u, count, stream, opts=None):
        count = self.batch if count is None else int(count)
        ptrs, spaces = [], set()
        for nm, a in (("P", P), ("q", q), ("A", A), ("l", l), ("u", u)):
            p, sp = _ptr_and_space(a, np.float64, nm)
            ptrs.append(p)
            spaces.add(sp)
        spaces.discard(None)
        if len(spaces) != 1:
            raise SolverError("P, q, A, l, u must all be host arrays or all be CUDA tensors")
        flags = spaces.pop()
        if flags == DEVICE_PTRS and stream is None:
            import torch

            stream = torch.cuda.current_stream().cuda_stream
        fn = getattr(self._L, "sqpb200_qp_batch_" + name)
        extra = () if opts is None else (int(opts),)
        self.ctx._check(fn(self._h, C.byref(self.settings), count, *ptrs, flags, C.c_void_p(stream or 0), *extra), name)

    def setup(self, P, q, A, l, u, count=None, stream=None):  # QPSolver::setup, qp.cpp:11-44
        self._call("setup", P, q, A, l, u, count, stream)

    def update_qp(self, P, q, A, l, u, count=None, stream=None):  # QPSolver::update_qp, qp.cpp:46-62
        self._call("update_qp", P, q, A, l, u, count, stream)

    def solve(self, P, q, A, l, u, count=None, stream=None):  # QPSolver::solve, qp.cpp:64-157
        self._call("solve", P, q, A, l, u, count, stream)

    def setup_solve(self, P, q, A, l, u, count=None, stream=None, opts=0):  # sqp.cpp:221-222 fused
        """opts: KEEP_FACTOR / REUSE_FACTOR for re-solves with the same P and A (sqp.cpp:244-276)."""
        if opts:
            self._call("setup_solve_opts", P, q, A, l, u, count, stream, opts=opts)
        else:
            self._call("setup_solve", P, q, A, l, u, count, stream)

    def setup_solve_to(self, P, q, A, l, u, out, count=None, stream=None):
        """Fused setup + solve of fresh instances; the kernel writes x, y, z, status, iter, rho_updates, rho_estimate, res_prim,
        res_dual straight into the device arrays of `out` (CUDA tensors or DevPtr, possibly peer memory). Asynchronous on `stream`."""
        count = self.batch if count is None else int(count)
        ins = []
        for nm, a in (("P", P), ("q", q), ("A", A), ("l", l), ("u", u)):
            p_, sp = _ptr_and_space(a, np.float64, nm)
            if sp != DEVICE_PTRS and p_ is not None:
                raise SolverError("setup_solve_to takes device arrays only")
            ins.append(p_)
        order = ["x", "y", "z", "status", "iter", "rho_updates", "rho_estimate", "res_prim", "res_dual"]
        ints = {"status", "iter", "rho_updates"}
        outs = []
        for k in order:
            p_, sp = _ptr_and_space(out[k], np.int32 if k in ints else np.float64, k)
            if sp != DEVICE_PTRS:
                raise SolverError("setup_solve_to takes device arrays only")
            outs.append(p_)
        if stream is None:
            import torch

            stream = torch.cuda.current_stream().cuda_stream
        self.ctx._check(self._L.sqpb200_qp_batch_setup_solve_to(self._h, C.byref(self.settings), count, *ins, *outs, C.c_void_p(stream or 0)),
                        "setup_solve_to")

    def setup_sparse(self, P, q, A_values, A_outer, A_inner, l, u, layout=SPARSE_CSC, count=None, stream=None):
        self.setup_solve_sparse(P, q, A_values, A_outer, A_inner, l, u, layout, count, stream, _fn="setup_sparse")

    def update_qp_sparse(self, P, q, A_values, A_outer, A_inner, l, u, layout=SPARSE_CSC, count=None, stream=None):
        self.setup_solve_sparse(P, q, A_values, A_outer, A_inner, l, u, layout, count, stream, _fn="update_qp_sparse")

    def solve_sparse(self, P, q, A_values, A_outer, A_inner, l, u, layout=SPARSE_CSC, count=None, stream=None):
        self.setup_solve_sparse(P, q, A_values, A_outer, A_inner, l, u, layout, count, stream, _fn="solve_sparse")

    def setup_solve_sparse(self, P, q, A_values, A_outer, A_inner, l, u, layout=SPARSE_CSC, count=None, stream=None, _fn="setup_solve_sparse"):
        """setup + solve with A given as CSC (Eigen::SparseMatrix layout) or CSR: one pattern for the whole batch
        (A_outer, A_inner: int32), per-instance values A_values[B, nnz]. setup_sparse / update_qp_sparse / solve_sparse are the
        separate calls of the object API with the same arguments."""
        count = self.batch if count is None else int(count)
        ptrs, spaces = [], set()
        for nm, a, dt in (("P", P, np.float64), ("q", q, np.float64), ("A_values", A_values, np.float64), ("A_outer", A_outer, np.int32),
                          ("A_inner", A_inner, np.int32), ("l", l, np.float64), ("u", u, np.float64)):
            p_, sp = _ptr_and_space(a, dt, nm)
            ptrs.append(p_)
            spaces.add(sp)
        spaces.discard(None)
        if len(spaces) != 1:
            raise SolverError("all arrays must be host arrays or all CUDA tensors")
        flags = spaces.pop()
        nnz = int(A_inner.shape[0])
        if flags == DEVICE_PTRS and stream is None:
            import torch

            stream = torch.cuda.current_stream().cuda_stream
        P_, q_, Av, Ao, Ai, l_, u_ = ptrs
        self.ctx._check(getattr(self._L, "sqpb200_qp_batch_" + _fn)(self._h, C.byref(self.settings), count, P_, q_, Av, Ao, Ai, nnz,
                                                                    int(layout), l_, u_, flags, C.c_void_p(stream or 0)), _fn)

    def get(self, count=None, fields=("x", "y", "z", "status", "iter", "rho_updates", "rho_estimate", "res_prim", "res_dual")):
        """Copy results to fresh host arrays (synchronises; ordered behind the last launch on this object whatever stream that used)."""
        count = self.batch if count is None else int(count)
        shapes = dict(x=((count, self.n), np.float64), y=((count, self.m), np.float64), z=((count, self.m), np.float64),
                      status=((count,), np.int32), iter=((count,), np.int32), rho_updates=((count,), np.int32),
                      rho_estimate=((count,), np.float64), res_prim=((count,), np.float64), res_dual=((count,), np.float64))
        out = {k: np.empty(*shapes[k]) for k in fields}
        order = ["x", "y", "z", "status", "iter", "rho_updates", "rho_estimate", "res_prim", "res_dual"]
        args = [out[k].ctypes.data if k in out else None for k in order]
        self.ctx._check(self._L.sqpb200_qp_batch_get(self._h, count, *args, HOST_PTRS, None), "qp_batch_get")
        return out

    def get_into(self, count=None, stream=None, **tensors):
        """Asynchronous copy into caller-provided arrays/tensors (all host or all CUDA)."""
        count = self.batch if count is None else int(count)
        order = ["x", "y", "z", "status", "iter", "rho_updates", "rho_estimate", "res_prim", "res_dual"]
        ints = {"status", "iter", "rho_updates"}
        args, spaces = [], set()
        for k in order:
            p, sp = _ptr_and_space(tensors.get(k), np.int32 if k in ints else np.float64, k)
            args.append(p)
            spaces.add(sp)
        spaces.discard(None)
        flags = spaces.pop() if spaces else HOST_PTRS
        if spaces:
            raise SolverError("outputs must all be host arrays or all be CUDA tensors")
        if flags == DEVICE_PTRS and stream is None:
            import torch

            stream = torch.cuda.current_stream().cuda_stream
        self.ctx._check(self._L.sqpb200_qp_batch_get(self._h, count, *args, flags, C.c_void_p(stream or 0)), "qp_batch_get")

    def set_iterates(self, x=None, y=None, z=None, count=None, stream=None):
        count = self.batch if count is None else int(count)
        args, spaces = [], set()
        for nm, a in (("x", x), ("y", y), ("z", z)):
            p, sp = _ptr_and_space(a, np.float64, nm)
            args.append(p)
            spaces.add(sp)
        spaces.discard(None)
        flags = spaces.pop() if spaces else HOST_PTRS
        self.ctx._check(self._L.sqpb200_qp_batch_set_iterates(self._h, count, *args, flags, C.c_void_p(stream or 0)),
                        "qp_batch_set_iterates")

    def device_view(self):
        v = DeviceView()
        self.ctx._check(self._L.sqpb200_qp_batch_device_view(self._h, C.byref(v)), "device_view")
        return v

    def total_iters(self, stream=None):
        """Sum of ADMM iterations executed by the last call (synchronises its stream)."""
        if stream is None:
            try:
                import torch

                stream = torch.cuda.current_stream().cuda_stream if torch.cuda.is_available() else 0
            except ImportError:
                stream = 0
        v = C.c_longlong()
        self.ctx._check(self._L.sqpb200_qp_batch_total_iters(self._h, C.byref(v), C.c_void_p(stream or 0)), "total_iters")
        return int(v.value)

    # QPSolver-style accessors (qp.hpp:159-169)
    def primal_solution(self, count=None):
        return self.get(count, fields=("x",))["x"]

    def dual_solution(self, count=None):
        return self.get(count, fields=("y",))["y"]

    def info(self, count=None):
        return self.get(count, fields=("status", "iter", "rho_updates", "rho_estimate", "res_prim", "res_dual"))


def solve_batch(ctx, P, q, A, l, u, n, m, settings=None):
    """One-shot fresh solvers + setup + solve + read-back through sqpb200_qp_solve_batch (host arrays)."""
    B = P.shape[0]
    b = QPBatch(ctx, B, n, m)
    try:
        if settings is not None:
            b.settings = settings
        b.setup_solve(P, q, A, l, u)
        out = b.get()
    finally:
        b.close()
    return out
