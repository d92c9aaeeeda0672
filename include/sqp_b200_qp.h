/*
 * sqp_b200_qp.h -- C-ABI of the B200-native batched QP-subproblem solver.
 *
 * This is the drop-in boundary for ONE path of msplr/sqp_solver: the OSQP-style ADMM solver
 * `qp_solver::QPSolver<Scalar>` (reference include/solvers/qp.hpp:113-250, src/qp.cpp:11-157).
 * The reference has no FFI of its own (it is a header+static-lib C++ library); the entry
 * points below are exactly what a binding of that class would need, batched over B independent
 * problem instances of identical (n, m):
 *
 *   sqpb200_qp_batch_create      <->  B x `QPSolver<double> solver;`            qp.hpp:148
 *   sqpb200_qp_batch_setup       <->  QPSolver::setup(qp)                        qp.hpp:151, qp.cpp:11-44
 *   sqpb200_qp_batch_update_qp   <->  QPSolver::update_qp(qp)                    qp.hpp:154, qp.cpp:46-62
 *   sqpb200_qp_batch_solve       <->  QPSolver::solve(qp)                        qp.hpp:157, qp.cpp:64-157
 *   sqpb200_qp_batch_setup_solve <->  setup(qp); solve(qp);  (the only in-library call site,
 *                                     SQP<T>::run_solve_qp)                      sqp.cpp:221-222
 *   sqpb200_qp_batch_get         <->  primal_solution() / dual_solution() / info()  qp.hpp:159-169
 *   sqpb200_qp_settings          <->  QPSolverSettings<double>                   qp.hpp:36-53
 *   sqpb200_qp_status            <->  QPSolverStatus                             qp.hpp:70
 *   sqpb200_constr_type_init     <->  static QPSolver::constr_type_init(l,u,out) qp.hpp:173, qp.cpp:283-294
 *
 * Data layout (caller-owned buffers, batch-major, each matrix COLUMN-major so one problem is
 * bit-compatible with Eigen::MatrixXd::data() / VectorXd::data() of QuadraticProblem, qp.hpp:19-34):
 *   P[B][n*n]  q[B][n]  A[B][m*n]  l[B][m]  u[B][m]   (fp64; +-inf and |bound| > 1e16 pass through)
 *   x[B][n]    y[B][m]  z[B][m]    status[B] iter[B] rho_updates[B] rho_estimate[B] res_prim[B] res_dual[B]
 *
 * Plain pointers and sizes only; no C++/torch types; nothing throws across this boundary.
 * Every function returns 0 on success or an sqpb200_error code; sqpb200_last_error() gives text.
 */
#ifndef SQP_B200_QP_H
#define SQP_B200_QP_H

#include <stddef.h>

#ifdef __cplusplus
extern "C" {
#endif

#define SQPB200_ABI_VERSION 1

/* mirrors QPSolverSettings<double>, include/solvers/qp.hpp:36-53 (same names, same defaults) */
typedef struct sqpb200_qp_settings {
    double rho;                    /* 1e-1 */
    double sigma;                  /* 1e-6 */
    double alpha;                  /* 1.0  */
    double eps_rel;                /* 1e-3 */
    double eps_abs;                /* 1e-3 */
    int max_iter;                  /* 1000 */
    int check_termination;         /* 25, 0 disables */
    int warm_start;                /* false; a no-op in the reference (qp.cpp:78-82) and here */
    int adaptive_rho;              /* false */
    double adaptive_rho_tolerance; /* 5 */
    int adaptive_rho_interval;     /* 25 */
    int verbose;                   /* false */
} sqpb200_qp_settings;

/* QPSolverStatus, include/solvers/qp.hpp:70 */
typedef enum sqpb200_qp_status {
    SQPB200_SOLVED = 0,
    SQPB200_MAX_ITER_EXCEEDED = 1,
    SQPB200_UNSOLVED = 2,
    SQPB200_NUMERICAL_ISSUES = 3,
    SQPB200_UNINITIALIZED = 4
} sqpb200_qp_status;

/* QPSolver::ConstraintType, include/solvers/qp.hpp:134 */
typedef enum sqpb200_constr_type {
    SQPB200_INEQUALITY_CONSTRAINT = 0,
    SQPB200_EQUALITY_CONSTRAINT = 1,
    SQPB200_LOOSE_BOUNDS = 2
} sqpb200_constr_type;

typedef enum sqpb200_error {
    SQPB200_OK = 0,
    SQPB200_ERR_INVALID = 1,     /* bad argument / size */
    SQPB200_ERR_CUDA = 2,        /* CUDA runtime failure (no device, launch error, ...) */
    SQPB200_ERR_UNSUPPORTED = 3, /* shape outside what the kernels cover */
    SQPB200_ERR_NOMEM = 4
} sqpb200_error;

/* pointer-space flags for the data arguments of one call */
#define SQPB200_HOST_PTRS 0u   /* caller passes host buffers; the call copies and synchronises */
#define SQPB200_DEVICE_PTRS 1u /* caller passes device buffers; the call is asynchronous on `stream` */
#define SQPB200_HOST_ASYNC 2u  /* caller passes PAGE-LOCKED host buffers (sqpb200_host_alloc) that stay untouched until `stream` has been
                                  synchronised: copies and launch are enqueued on `stream` and the call returns at once. For callers that
                                  pipeline host work against the GPU (sqp::BatchSQP advances two groups of instances alternately) */
/* `stream` is a cudaStream_t passed as void*; NULL is the CUDA legacy default stream. */

/* context options for sqpb200_ctx_set_option */
#define SQPB200_OPT_KERNEL 1       /* 0 = auto (default), 1 = force the generic kernel, 2 = force the register-tiled kernel,
                                      3 = force the blocked kernel, 4 = force the cluster kernel (sparse A only), 5 = force the
                                      thread-per-QP literal KKT kernel (n + m <= 16) (tests / tuning) */
#define SQPB200_OPT_H2D_CHUNKS 2   /* number of staging chunks for HOST_PTRS calls (default 16) */
#define SQPB200_OPT_CTAS_PER_SM 3  /* 0 = auto */
#define SQPB200_OPT_TILE_WARPS 4   /* warps per QP of the register-tiled kernel: 0 = default; 4 (default) or 8 for the 64x128 class, 1 (default) or 2 for the 32x64 class (tuning/tests) */

#define SQPB200_OPT_SLICE_ITERS 5  /* time slicing of the register-tiled kernel (n <= 64, m <= 128 class of 64 x 128): a QP is suspended after this
                                      many ADMM iterations and re-queued, so that a batch of only a few QPs per resident CTA is scheduled in small
                                      units (strong scaling over several GPUs). -1 = automatic (default: a first slice of max_iter / 2 and slices
                                      of max_iter / 8 after it when a launch on local data has at most four QPs per CTA slot), 0 = off; value > 0: iterations per
                                      slice, + 65536 x the length of a QP's first slice if that is to be longer. Results are bit-identical to an
                                      unsliced solve */

typedef struct sqpb200_ctx sqpb200_ctx;           /* one per (host thread, GPU) */
typedef struct sqpb200_qp_batch sqpb200_qp_batch; /* B solver instances: state x,z,y, info, factor */

/* ---- context ------------------------------------------------------------------------------- */
int sqpb200_abi_version(void);
int sqpb200_ctx_create(int device, sqpb200_ctx **out);
int sqpb200_ctx_destroy(sqpb200_ctx *ctx);
int sqpb200_ctx_set_option(sqpb200_ctx *ctx, int option, int value);
/* text of the last failure on this context (ctx == NULL: last failure of a create call on this thread) */
const char *sqpb200_last_error(const sqpb200_ctx *ctx);
int sqpb200_device_query(const sqpb200_ctx *ctx, int *device, int *sm_count, int *cc_major, int *cc_minor,
                         size_t *smem_per_block_optin);
/* number of kernels this context has launched so far (bench.py's "gpu_launches") */
long long sqpb200_launch_count(const sqpb200_ctx *ctx);
/* name of the kernel the last solve dispatched to ("generic", "tile<64,128,8>", ...) */
const char *sqpb200_last_kernel(const sqpb200_ctx *ctx);

/* fp64 FMA peak of this device, measured now with a saturating DFMA kernel (a few ms): *tflops = 2 x FMA / s. Used by bench.py for
 * the roofline of the compute-bound kernels (the driver's MEASURED_PEAKS.json carries HBM and bf16 figures only). */
int sqpb200_measure_fp64_peak(sqpb200_ctx *ctx, double *tflops, double *seconds, double *fma_count);

void sqpb200_qp_default_settings(sqpb200_qp_settings *s);

/* host-side, no GPU: static QPSolver::constr_type_init (qp.cpp:283-294) */
int sqpb200_constr_type_init(const double *l, const double *u, int m, int *constr_type);

/* host-side, no GPU: the chunk boundaries a staged HOST_PTRS call of `count` instances uses (SQPB200_OPT_H2D_CHUNKS = `chunks`, `wave` =
 * resident CTAs of the kernel): bound[0] = 0 < bound[1] < ... < bound[return value] = count, a geometric ramp and half-wave steps first
 * (the first CTAs start after a fraction of a per cent of the transfer), equal chunks after eight waves; -1 on invalid arguments
 * (cap must be at least chunks + 3). Exposed for the tests. */
int sqpb200_staging_chunk_bounds(int count, int chunks, int wave, int *bound, int cap);

/* ---- batched solver object ---------------------------------------------------------------- */
/* `batch` is the capacity; every call below processes the first `count` instances. A new batch
 * object is B default-constructed solvers: status UNINITIALIZED, iter 0, rho_updates 0 (qp.hpp:72-79). */
int sqpb200_qp_batch_create(sqpb200_ctx *ctx, int batch, int n, int m, sqpb200_qp_batch **out);
int sqpb200_qp_batch_destroy(sqpb200_qp_batch *b);

/* Compute precision of one batch object: 0 = fp64 (default), 1 = fp32 -- the reference's second explicit instantiation,
 * `template class QPSolver<float>` (src/qp.cpp:386, tests/qp_solver_test.cpp:58-69). fp32 changes the ARITHMETIC (registers, shared
 * memory, every operation; DIV_BY_ZERO_REGUL becomes numeric_limits<float>::epsilon()); the arrays of this ABI stay `double`
 * (inputs are rounded to float on load, results are float values widened on store: lossless, so warm starts and stored factors
 * round-trip exactly). Available for the shapes of the register-tiled kernel (n <= 64, m <= 128); larger shapes keep computing in
 * fp64. Meets the reference's own 1e-2 float test, not the 1e-6 parity bar of the fp64 path. */
int sqpb200_qp_batch_set_precision(sqpb200_qp_batch *b, int fp32);

/* ---- peer memory (multi-GPU, one process per GPU) --------------------------------------------------------------------------
 * The batch axis shards with no exchange during the solve. For the strong-scaling flow of the north star (one batch
 * owned by one rank, split over the GPUs, results gathered back) these helpers let every rank's solve kernel READ its slice of the
 * inputs straight out of the owner's memory over NVLink (the kernel's TMA bulk copies and loads take the peer address) and WRITE its
 * results straight into the owner's arrays (sqpb200_qp_batch_get with device pointers into the imported mapping): no separate split
 * or gather step, the transfer overlaps the solve QP by QP. Buffers come from sqpb200_dev_alloc (so that a handle maps the whole
 * allocation), are exported as 64-byte CUDA IPC handles, and imported by the other processes of the node. */
/* streams for SQPB200_HOST_ASYNC / SQPB200_DEVICE_PTRS callers that have no CUDA runtime of their own (non-blocking streams) */
int sqpb200_stream_create(sqpb200_ctx *ctx, void **stream);
int sqpb200_stream_destroy(sqpb200_ctx *ctx, void *stream);
int sqpb200_stream_sync(sqpb200_ctx *ctx, void *stream);
/* page-locked host buffers for HOST_PTRS callers that call every few hundred microseconds (the SQP outer loop, src/sqp.cpp:210-242):
 * copies from pinned memory are asynchronous and need no intermediate staging by the driver */
int sqpb200_host_alloc(sqpb200_ctx *ctx, size_t bytes, void **host_ptr);
int sqpb200_host_free(sqpb200_ctx *ctx, void *host_ptr);
#define SQPB200_IPC_HANDLE_BYTES 64
int sqpb200_dev_alloc(sqpb200_ctx *ctx, size_t bytes, void **dev_ptr);
int sqpb200_dev_free(sqpb200_ctx *ctx, void *dev_ptr);
/* copy between any two of: host memory, this device's memory, imported peer memory (cudaMemcpyDefault), asynchronous on `stream` */
int sqpb200_dev_copy(sqpb200_ctx *ctx, void *dst, const void *src, size_t bytes, void *stream);
int sqpb200_ipc_export(sqpb200_ctx *ctx, const void *dev_ptr, unsigned char handle[SQPB200_IPC_HANDLE_BYTES]);
int sqpb200_ipc_import(sqpb200_ctx *ctx, const unsigned char handle[SQPB200_IPC_HANDLE_BYTES], void **dev_ptr);
int sqpb200_ipc_release(sqpb200_ctx *ctx, void *dev_ptr);

/* setup: zero x,z,y; classify constraints; rho vector from settings->rho (rho_updates += 1);
 * build and factor the KKT system; status = UNSOLVED or NUMERICAL_ISSUES. qp.cpp:11-44 */
int sqpb200_qp_batch_setup(sqpb200_qp_batch *b, const sqpb200_qp_settings *settings, int count, const double *P,
                           const double *q, const double *A, const double *l, const double *u, unsigned flags,
                           void *stream);
/* update_qp: same without the x,z,y reset. qp.cpp:46-62 */
int sqpb200_qp_batch_update_qp(sqpb200_qp_batch *b, const sqpb200_qp_settings *settings, int count, const double *P,
                               const double *q, const double *A, const double *l, const double *u, unsigned flags,
                               void *stream);
/* solve: the ADMM loop from the current x,z,y (always a warm start, see qp.cpp:78-82), using the
 * factor and constraint classes of the last setup/update_qp. Instances whose status is
 * UNINITIALIZED or NUMERICAL_ISSUES are left untouched (qp.cpp:68-71). qp.cpp:64-157 */
int sqpb200_qp_batch_solve(sqpb200_qp_batch *b, const sqpb200_qp_settings *settings, int count, const double *P,
                           const double *q, const double *A, const double *l, const double *u, unsigned flags,
                           void *stream);
/* setup immediately followed by solve in ONE kernel launch (the factor never leaves the SM).
 * This is the hot path: SQP<T>::run_solve_qp, sqp.cpp:221-222. */
int sqpb200_qp_batch_setup_solve(sqpb200_qp_batch *b, const sqpb200_qp_settings *settings, int count, const double *P,
                                 const double *q, const double *A, const double *l, const double *u, unsigned flags,
                                 void *stream);

/* setup_solve of `count` FRESH solver instances (default-constructed: rho_updates starts at 0) whose results are written by the
 * kernel's epilogue straight into caller-provided DEVICE arrays -- which may live in another GPU's memory (imported with
 * sqpb200_ipc_import): with the inputs read from the owner's memory as well, a rank solves its slice of a batch another GPU owns
 * with no split, no gather and no copy at all (SURVEY.md 8e "the gather can be folded into the solve kernel's epilogue").
 * Device pointers only; asynchronous on `stream`. The object's own x, y, z and info are NOT advanced by this call. Shapes
 * outside the register-tiled kernel fall back to a solve into the object followed by device-to-device copies. */
int sqpb200_qp_batch_setup_solve_to(sqpb200_qp_batch *b, const sqpb200_qp_settings *settings, int count, const double *P,
                                    const double *q, const double *A, const double *l, const double *u, double *x, double *y,
                                    double *z, int *status, int *iter, int *rho_updates, double *rho_estimate, double *res_prim,
                                    double *res_dual, void *stream);

/* setup_solve with factor bookkeeping for callers that re-solve the SAME P and A with new q, l, u -- the
 * second-order-correction QP of SQP<T>::second_order_correction (sqp.cpp:244-276, "TODO: only l and u change").
 *   SQPB200_KEEP_FACTOR   remember each instance's setup factor (one n*n store per QP)
 *   SQPB200_REUSE_FACTOR  the caller guarantees P and A are those of the call that kept the factor; every instance whose
 *                         constraint classes (from the new l, u) and settings->rho are unchanged skips the factorisation,
 *                         the others factor as usual. Results are bit-identical to a plain setup_solve. */
#define SQPB200_KEEP_FACTOR 1u
#define SQPB200_REUSE_FACTOR 2u
int sqpb200_qp_batch_setup_solve_opts(sqpb200_qp_batch *b, const sqpb200_qp_settings *settings, int count, const double *P,
                                      const double *q, const double *A, const double *l, const double *u, unsigned flags,
                                      void *stream, unsigned opts);

/* setup + solve with a SPARSE constraint matrix: one sparsity pattern shared by the batch, per-instance values
 * A_values[B][nnz]. layout SQPB200_SPARSE_CSC is Eigen::SparseMatrix's compressed column storage (the reference's intended
 * sparse variant, qp.hpp:22-25; A_outer has n+1 entries, A_inner holds row indices); SQPB200_SPARSE_CSR has m+1 outer
 * entries and column indices. P stays dense. Shapes of the thread-per-QP and register-tiled kernels (n <= 64, m <= 128) densify on
 * the device (A lives in registers there anyway; results are identical to the dense entry points); 64 < n <= 256 run the
 * thread-block-cluster kernel (A compressed in shared memory, H^-1 distributed over the cluster) or, when the instance does not
 * fit on chip, the blocked kernel walking the compressed pattern. */
#define SQPB200_SPARSE_CSC 0
#define SQPB200_SPARSE_CSR 1
int sqpb200_qp_batch_setup_solve_sparse(sqpb200_qp_batch *b, const sqpb200_qp_settings *settings, int count, const double *P,
                                        const double *q, const double *A_values, const int *A_outer, const int *A_inner, int nnz,
                                        int layout, const double *l, const double *u, unsigned flags, void *stream);

/* The object API with a sparse A (same arguments): setup / update_qp / solve as separate calls on persistent solver state, the
 * call sequence of the reference's sparse tests (tests/qp_solver_sparse_test.cpp:68-98: setup, solve, solve; setup, solve, update_qp,
 * solve). With the cluster kernel the distributed factor is not kept between launches: solve() rebuilds it from the stored constraint
 * classes and rho (results identical to a kept factor). */
int sqpb200_qp_batch_setup_sparse(sqpb200_qp_batch *b, const sqpb200_qp_settings *settings, int count, const double *P,
                                  const double *q, const double *A_values, const int *A_outer, const int *A_inner, int nnz, int layout,
                                  const double *l, const double *u, unsigned flags, void *stream);
int sqpb200_qp_batch_update_qp_sparse(sqpb200_qp_batch *b, const sqpb200_qp_settings *settings, int count, const double *P,
                                      const double *q, const double *A_values, const int *A_outer, const int *A_inner, int nnz,
                                      int layout, const double *l, const double *u, unsigned flags, void *stream);
int sqpb200_qp_batch_solve_sparse(sqpb200_qp_batch *b, const sqpb200_qp_settings *settings, int count, const double *P,
                                  const double *q, const double *A_values, const int *A_outer, const int *A_inner, int nnz, int layout,
                                  const double *l, const double *u, unsigned flags, void *stream);

/* Read back solutions and info (any pointer may be NULL). primal_solution()/dual_solution()/info(),
 * qp.hpp:159-169; z is exposed in addition so a caller can checkpoint a warm start. */
int sqpb200_qp_batch_get(sqpb200_qp_batch *b, int count, double *x, double *y, double *z, int *status, int *iter,
                         int *rho_updates, double *rho_estimate, double *res_prim, double *res_dual, unsigned flags,
                         void *stream);
/* Overwrite the iterates (warm start from a checkpoint). Any pointer may be NULL. */
int sqpb200_qp_batch_set_iterates(sqpb200_qp_batch *b, int count, const double *x, const double *y, const double *z,
                                  unsigned flags, void *stream);

/* Zero-copy view of the device-resident state for device-side consumers. */
typedef struct sqpb200_qp_device_view {
    double *x, *y, *z;
    int *status, *iter, *rho_updates;
    double *rho_estimate, *res_prim, *res_dual, *rho;
    long long *total_iters; /* device scalar: sum of ADMM iterations executed by the last solve launch */
} sqpb200_qp_device_view;
int sqpb200_qp_batch_device_view(sqpb200_qp_batch *b, sqpb200_qp_device_view *view);

/* Sum of ADMM iterations the last setup/solve call executed over all its instances
 * (synchronises the call's stream). bench.py converts QP/s into ADMM iterations/s with it. */
int sqpb200_qp_batch_total_iters(sqpb200_qp_batch *b, long long *total, void *stream);

/* Convenience one-shot: fresh solvers, setup+solve, read back. Equivalent to the five calls
 * create / setup_solve / get / destroy with HOST or DEVICE pointers. */
int sqpb200_qp_solve_batch(sqpb200_ctx *ctx, const sqpb200_qp_settings *settings, int batch, int n, int m,
                           const double *P, const double *q, const double *A, const double *l, const double *u,
                           double *x, double *y, double *z, int *status, int *iter, int *rho_updates,
                           double *rho_estimate, double *res_prim, double *res_dual, unsigned flags, void *stream);

#ifdef __cplusplus
}
#endif
#endif /* SQP_B200_QP_H */
