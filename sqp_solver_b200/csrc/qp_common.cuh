// Shared host/device definitions of the batched QP-subproblem solver (sm_100a).
//
// Algorithm (per QP, fp64) -- the OSQP-style ADMM of reference src/qp.cpp:64-157, with the KKT
// solve of qp.cpp:90 carried out by eliminating the diagonal (2,2) block of
//     K = [[P + sigma I, A^T], [A, -diag(1/rho)]]                       (qp.hpp:177-188)
// first.  That is an exact block LDL^T of a symmetric permutation of K:
//     H = P_sym + sigma I + A^T diag(rho) A          (n x n Schur complement, SPD for convex QPs)
//     H x~ = sigma x - q + A^T (rho .* z - y)        (rhs of qp.cpp:272-276 pushed through the block)
//     nu   = rho .* (A x~ - z) + y   =>   z~ = z_prev + (nu - y) ./ rho = A x~      (qp.cpp:93)
// H is factored as L D L^T in shared memory (no pivoting; NUMERICAL_ISSUES on a zero/NaN pivot,
// which is when Eigen::LDLT::info() != Success), the unit factor is inverted once and
// H^-1 = L^-T D^-1 L^-1 is applied per iteration as one dense symmetric mat-vec, so the ADMM
// iteration has no serial substitution chain.  Like Eigen::LDLT<Lower> only the LOWER triangle
// of P enters the factor, while P*x in the residuals (qp.cpp:323, :359) uses all of P.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/sqp_b200_qp.h"

namespace sqpb200 {

// include/solvers/qp.hpp:136-141
constexpr double RHO_MIN = 1e-6;
constexpr double RHO_MAX = 1e+6;
constexpr double RHO_TOL = 1e-4;
constexpr double RHO_EQ_FACTOR = 1e+3;
constexpr double LOOSE_BOUNDS_THRESH = 1e+16;
constexpr double DIV_BY_ZERO_REGUL = 2.220446049250313e-16;  // numeric_limits<double>::epsilon()

// what one launch does for every QP it processes
enum : unsigned {
    MODE_RESET = 1u,         // x = z = y = 0                                   (qp.cpp:16-18)
    MODE_FACTOR = 2u,        // classify, rho vec from settings.rho, factor      (qp.cpp:31-43 / 48-61)
    MODE_SOLVE = 4u,         // ADMM loop                                        (qp.cpp:64-157)
    MODE_STORE_FACTOR = 8u,  // keep H^-1, rho, constraint classes for a later solve() launch
    MODE_LOAD_FACTOR = 16u,  // solve() after a separate setup()/update_qp()
    MODE_KEEP_INITIAL = 32u, // with FACTOR: remember the setup factor (for settings.rho) in the slab, tagged with its rho
    MODE_FRESH = 128u,       // the instances are default-constructed solvers (status UNINITIALIZED, counters 0): do not read the
                             // status / info arrays, only write them (they may be caller arrays in peer memory, never initialised)
    MODE_REUSE = 64u,        // with FACTOR: same P, A as the launch that kept the factor; skip the factorisation of every
                             // instance whose constraint classes and rho are unchanged (the TODO at reference sqp.cpp:273)
};

// Sparse constraint matrix with one pattern shared by the batch (blocked kernel). Both a row-compressed and a
// column-compressed view of the pattern are given; `vals` is stored in the order of ONE of them (perm == nullptr) and the
// other view reaches it through its perm array. row_outer == nullptr means "A is dense".
// packed sparse entries (cluster kernel): inner index in the low PACK_BITS bits, position of the value above them. The inner
// index is a row (< 2048 = 8 CTAs x 256 rows) or a column (< 256); the value position must stay below 2^(32 - PACK_BITS).
constexpr int PACK_BITS = 11;
constexpr unsigned PACK_MASK = (1u << PACK_BITS) - 1u;
constexpr int PACK_MAX_NNZ = 1 << (32 - PACK_BITS);

struct SparseA {
    const int *row_outer, *row_inner, *row_perm;  // CSR view: row_outer[m+1], row_inner[nnz] = column indices
    const int *col_outer, *col_inner, *col_perm;  // CSC view: col_outer[n+1], col_inner[nnz] = row indices
    const double *vals;                           // [B][nnz]
    const unsigned *col_pack, *row_pack;          // per stored entry of the CSC / CSR view: inner index | (position in vals << PACK_BITS)
    int nnz;
    int col_slice_cap;  // cluster kernel: most stored entries in the columns one CTA owns
    int cluster_size;   // cluster kernel: CTAs per QP (4 or 8); 0 = not planned for the cluster kernel
};

struct KernelParams {
    int first;  // index of the first QP of this launch inside the batch arrays
    int count;  // QPs in this launch
    int n, m;
    const double *P, *q, *A, *l, *u;  // batch-major inputs (already offset-free; kernels add first)
    double *x, *z, *y;
    int *status, *iter, *rho_updates;
    double *rho_estimate, *res_prim, *res_dual, *rho;
    signed char *ctype;  // [B][m] constraint classes of the last setup/update_qp
    double *fact;        // [B][n*n] H^-1 (column-major, full symmetric); may be null when fused
    double *fact_rho;    // [B] scalar rho the stored H^-1 belongs to (NaN: none stored)
    double *scratch;     // generic kernel: per-CTA n*n workspace
    int *work_counter;   // persistent-CTA work queue
    const int *ready;    // optional: number of leading QPs whose inputs have landed in device memory (host-staged calls);
                         // a CTA that draws QP i waits until *ready > i. nullptr: everything is resident
    unsigned long long *total_iters;
    // Time slicing (register-tiled kernel, fused launches of fresh instances): a QP is suspended after `slice_iters` iterations of one
    // slice -- x, z, y, info and H^-1 go to the object's arrays -- and re-queued, so that a batch of only a few QPs per resident CTA
    // (one batch split over 8 GPUs) is list-scheduled in small units instead of whole 500...1001-iteration solves. 0: off.
    int slice_iters;
    int slice_first;  // length of a QP's FIRST slice (>= slice_iters: every solve needs its first few hundred iterations anyway)
    int *rq;        // re-queue ring: entry e holds the local index of the e-th suspended QP (-1: not published yet)
    int *rq_alloc;  // ring slots handed out so far
    int *done;      // QPs of this launch that have finished
    int rq_cap;
    double *sus_x, *sus_z, *sus_y;  // suspended state (the batch object's own arrays; the final results go to x, z, y, status, ... above)
    int *sus_status, *sus_iter, *sus_rho_updates;
    double *sus_rho_estimate, *sus_res_prim, *sus_res_dual;
    double *loc_P, *loc_A;  // local copies of P and A for the resumes when the inputs live in another GPU's memory (else nullptr)
    double *loc_q, *loc_l, *loc_u;  // likewise q, l, u (set together with loc_P / loc_A): a resume then touches local memory only
    unsigned mode;
    sqpb200_qp_settings s;
    SparseA sp;  // all-zero for dense A
};

#ifdef __CUDACC__

__device__ __forceinline__ int classify(double l, double u) {  // qp.cpp:283-294
    if (l < -LOOSE_BOUNDS_THRESH && u > LOOSE_BOUNDS_THRESH) return SQPB200_LOOSE_BOUNDS;
    if (u - l < RHO_TOL) return SQPB200_EQUALITY_CONSTRAINT;
    return SQPB200_INEQUALITY_CONSTRAINT;
}
__device__ __forceinline__ double rho_of(int type, double rho0) {  // qp.cpp:296-310
    return type == SQPB200_LOOSE_BOUNDS ? RHO_MIN : (type == SQPB200_EQUALITY_CONSTRAINT ? RHO_EQ_FACTOR * rho0 : rho0);
}
// Eigen's z.cwiseMax(l).cwiseMin(u): std::max then std::min comparison forms (qp.cpp:278-281)
__device__ __forceinline__ double box_project(double z, double l, double u) {
    z = (z < l) ? l : z;
    z = (u < z) ? u : z;
    return z;
}
__device__ __forceinline__ double absmax(double r, double v) {
    v = fabs(v);
    return v > r ? v : r;
}
// Draw the next QP of this launch from the work queue (one thread per CTA calls this). For host-staged
// calls the inputs arrive chunk by chunk on a copy stream while the kernel is already running: wait
// until the chunk holding this QP has landed (flag written by a stream-ordered 4-byte copy after the data).
__device__ __forceinline__ int draw_qp(const KernelParams &p) {
    const int v = atomicAdd(p.work_counter, 1);
    if (p.ready != nullptr && v < p.count) {
        // bounded wait: if the staging copies never arrive (a failed copy on the host side) the launch aborts with an error after
        // ~8 s instead of hanging the device
        const long long t0 = clock64();
        while (*reinterpret_cast<const volatile int *>(p.ready) <= v) {
            __nanosleep(500);
            if (clock64() - t0 > (1LL << 34)) __trap();
        }
        __threadfence();
    }
    return v;
}
__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
__device__ __forceinline__ double warp_max(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        double t = __shfl_xor_sync(0xffffffffu, v, o);
        v = t > v ? t : v;
    }
    return v;
}
// rho_estimate + clamp, qp.cpp:130-132 and :333-341
__device__ __forceinline__ double rho_estimate_clamped(double rho, double rp, double rd, double sc_p, double sc_d) {
    double rp_norm = rp / (sc_p + DIV_BY_ZERO_REGUL);
    double rd_norm = rd / (sc_d + DIV_BY_ZERO_REGUL);
    double r = rho * sqrt(rp_norm / (rd_norm + DIV_BY_ZERO_REGUL));
    return fmax(RHO_MIN, fmin(r, RHO_MAX));
}

// ---- scalar-generic versions (the register-tiled kernel is instantiated for fp64 and fp32, the reference's QPSolver<double>
// and QPSolver<float>; the constants are the reference's `Scalar` constexprs, qp.hpp:136-141) ---------------------------------
template <typename S>
__device__ __forceinline__ int classify_t(S l, S u) {  // qp.cpp:283-294
    if (l < -S(LOOSE_BOUNDS_THRESH) && u > S(LOOSE_BOUNDS_THRESH)) return SQPB200_LOOSE_BOUNDS;
    if (u - l < S(RHO_TOL)) return SQPB200_EQUALITY_CONSTRAINT;
    return SQPB200_INEQUALITY_CONSTRAINT;
}
template <typename S>
__device__ __forceinline__ S rho_of_t(int type, S rho0) {  // qp.cpp:296-310
    return type == SQPB200_LOOSE_BOUNDS ? S(RHO_MIN) : (type == SQPB200_EQUALITY_CONSTRAINT ? S(RHO_EQ_FACTOR) * rho0 : rho0);
}
__device__ __forceinline__ float box_project(float z, float l, float u) {
    z = (z < l) ? l : z;
    z = (u < z) ? u : z;
    return z;
}
__device__ __forceinline__ float absmax(float r, float v) {
    v = fabsf(v);
    return v > r ? v : r;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        float t = __shfl_xor_sync(0xffffffffu, v, o);
        v = t > v ? t : v;
    }
    return v;
}
template <typename S>
__device__ __forceinline__ S rho_estimate_clamped_t(S rho, S rp, S rd, S sc_p, S sc_d) {  // qp.cpp:130-132 and :333-341
    const S eps = sizeof(S) == 8 ? S(DIV_BY_ZERO_REGUL) : S(1.1920928955078125e-7);  // numeric_limits<Scalar>::epsilon()
    S rp_norm = rp / (sc_p + eps);
    S rd_norm = rd / (sc_d + eps);
    // `Scalar rho_new = rho0 * sqrt(...)` (qp.cpp:338): sqrt is the double overload for a float argument too, so the product is
    // formed in double and rounded to Scalar once (the CPU path's exact semantics; matters for the float instantiation only)
    S r = (S)((double)rho * sqrt((double)(rp_norm / (rd_norm + eps))));
    return fmax(S(RHO_MIN), fmin(r, S(RHO_MAX)));
}

#endif  // __CUDACC__

// launchers implemented in the kernel translation units
cudaError_t launch_generic(const KernelParams &p, int sm_count, size_t smem_optin, int f32, cudaStream_t stream, int *grid_out);
size_t generic_scratch_bytes(int n, int grid);
int generic_grid(int count, int sm_count);
bool generic_supported(int n, int m, size_t smem_optin);
cudaError_t measure_dfma_peak(int sm_count, cudaStream_t stream, double *seconds, double *fma_count);  // peak_fp64.cu
// thread-per-QP literal KKT kernel for n + m <= 16 (qp_small.cu)
bool small_supported(int n, int m);
cudaError_t launch_small(const KernelParams &p, int f32, cudaStream_t stream, char *name, size_t name_len);
// blocked kernel for 64 < n <= 256 (qp_block.cu)
bool block_supported(int n, int m, size_t smem_optin);
bool block_sparse_supported(int n, int m, int nnz, size_t smem_optin);
size_t block_fact_doubles(int n);
cudaError_t launch_block(const KernelParams &p, int sm_count, size_t smem_optin, cudaStream_t stream, char *name, size_t name_len);
// cluster kernel for sparse A, 64 < n <= 256 (qp_cluster.cu)
int cluster_plan(int n, int m, int nnz, const int *colcount, size_t smem_optin, int *col_slice_cap);
int cluster_max_clusters(int n, int m, int nnz, int col_slice_cap, int cluster_size);
size_t cluster_scratch_bytes(int clusters);
cudaError_t launch_cluster(const KernelParams &p, int clusters, double *scratch, cudaStream_t stream, char *name, size_t name_len);
cudaError_t launch_densify(const double *vals, const int *outer, const int *inner, int nnz, int m, int n, int csr, int count,
                           double *dst, cudaStream_t stream);

}  // namespace sqpb200
