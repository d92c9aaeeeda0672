// Thread-block-cluster batched ADMM QP kernel for sparse constraint matrices with 64 < n <= 256 -- BASELINE.json config 5
// (n = 256, m = 512, one sparsity pattern shared by the batch).
//
// One CLUSTER of CS = 4 CTAs (4 SMs) per QP (8 when the pattern is too dense for four at n > 128), persistent clusters on the atomic work queue. H^-1 (n x n fp64, 512 KB at
// n = 256) does not fit in one SM's shared memory, but it fits in the cluster's: CTA r keeps rows [r n/4, (r+1) n/4) of it
// (all columns) in its own shared memory for the whole solve, so the per-iteration KKT solve (reference qp.cpp:90) is ONE
// dense mat-vec out of shared memory -- no substitution chain, no L2 traffic. Per iteration the CTAs exchange only vectors
// over distributed shared memory (w = rho.*z - y, m doubles, and x~, n doubles: two all-gathers, two cluster barriers).
//   * A stays compressed: the values of this instance (nnz doubles) sit in every CTA's shared memory, the batch-shared
//     pattern (CSR and CSC views) is read through L1/L2. Every CTA forms b = sigma x - q + A^T w redundantly (one column
//     per thread) and z~ = A x~ for the constraint rows it owns.
//   * H = P_lowsym + sigma I + A^T diag(rho) A is formed slice by slice straight from the pattern and inverted in place by a
//     BLOCKED symmetric sweep (Gauss-Jordan on 32-pivot blocks): per block the owner CTA inverts the 32 x 32 pivot block (its
//     pivots are the D of the unpivoted L D L^T of H: a zero/NaN pivot reports NUMERICAL_ISSUES like Eigen::LDLT::info()) and
//     publishes it with its 32 pivot rows through a small global (L2) scratch; every CTA then applies the rank-32 update to
//     its own row slice on the fp64 tensor cores (mma.sync m8n8k4 DMMA, operands from shared memory).
// Reference functions covered: the same list as qp_generic.cu (all of src/qp.cpp:11-371), fused setup + solve only (the sparse
// entry point has no separate setup/solve launches).
#include <cooperative_groups.h>

#include <cstdio>
#include <type_traits>

#include "qp_common.cuh"

namespace cg = cooperative_groups;

namespace sqpb200 {

#ifdef SQPB200_CLUSTER_TIMING
#define TCK(i) { long long t_ = clock64(); tph[i] += t_ - tlast; tlast = t_; }
#else
#define TCK(i)
#endif

constexpr int CT = 256;  // threads per CTA
constexpr int CNW = CT / 32;
constexpr int KB = 32;   // pivot block
constexpr int RCH = 64;  // columns of the pivot row panel staged per chunk
constexpr int LDR = KB + 4;  // leading dimension of a staged chunk (k contiguous): conflict-free B fragments
constexpr int LDE = KB + 1;

__host__ __device__ inline int cluster_np(int n) { return n <= 128 ? 128 : 256; }
__host__ __device__ inline size_t cluster_scratch_doubles() { return 2 * ((size_t)KB * 256 + KB * KB + 8); }  // per cluster

struct ClusterSmem {
    double *S;     // (RS + 2) x np: this CTA's row slice, column-major
    double *X;     // aliased region: sparse data of this instance | sweep buffers (R chunks, T, E)
    double *sw;    // m (+1)
    double *sb, *sxt, *sx, *sq;  // np each
    double *xp;    // CS * np: partial x~ of every CTA of the cluster (also scratch for P x at the checks)
    double *red;   // CS * 8
};
// ccap = the largest number of stored entries in the n/CS columns one CTA owns
__host__ __device__ inline size_t cluster_x_doubles(int np, int m, int nnz, int ccap, int CS) {
    const size_t RS = np / CS;
    const size_t sweep = 2 * (size_t)LDR * RCH + (RS + 4) * KB + 2 * (size_t)LDE * KB;
    // sparse data of the instance: values | packed CSC entries of the own columns | packed CSR entries | own column pointers | row pointers
    size_t v = (size_t)(nnz + 2 - (nnz & 1)) + ((size_t)ccap + (size_t)nnz + RS + 1 + (size_t)m + 1 + 1) / 2;
    v += v & 1;  // keeps every region behind it 16-byte aligned
    return v > sweep ? v : sweep;
}
__host__ __device__ inline size_t cluster_smem_doubles(int np, int m, int nnz, int ccap, int CS) {
    const size_t RS = np / CS;
    return (RS + 2) * np + cluster_x_doubles(np, m, nnz, ccap, CS) + (size_t)(m + (m & 1)) + 4 * (size_t)np + (size_t)CS * np + CS * 8;
}
__device__ __forceinline__ ClusterSmem carve_cluster(double *base, int np, int m, int nnz, int ccap, int CS) {
    ClusterSmem s;
    const int RS = np / CS;
    s.S = base;
    s.X = s.S + (size_t)(RS + 2) * np;
    s.sw = s.X + cluster_x_doubles(np, m, nnz, ccap, CS);
    s.sb = s.sw + (m + (m & 1));
    s.sxt = s.sb + np;
    s.sx = s.sxt + np;
    s.sq = s.sx + np;
    s.xp = s.sq + np;
    s.red = s.xp + (size_t)CS * np;
    return s;
}

// Cluster size for a sparse problem: 4 CTAs when the instance fits their shared memory, else 8 (n > 128 only: a pivot block must
// not straddle two CTAs), else 0 (the blocked kernel takes it). colcount[j] = stored entries of column j; *ccap = the most entries
// in the columns one CTA owns.
int cluster_plan(int n, int m, int nnz, const int *colcount, size_t smem_optin, int *ccap) {
    if (!(n > 64 && n <= 256 && m >= 1 && nnz >= 0)) return 0;
    // packed entries: a row index must fit PACK_BITS bits and a value position (nnz itself for the padding entry) the rest
    if (m > (int)PACK_MASK + 1 || nnz >= PACK_MAX_NNZ - 1) return 0;
    const int np = cluster_np(n);
    for (int cs = 4; cs <= 8; cs *= 2) {
        const int rs = np / cs;
        if (rs < KB || m > cs * CT) continue;
        int cap = 0;
        for (int j0 = 0; j0 < n; j0 += rs) {
            int cnt = 0;
            for (int j = j0; j < j0 + rs && j < n; ++j) cnt += colcount[j];
            if (cnt > cap) cap = cnt;
        }
        if (sizeof(double) * cluster_smem_doubles(np, m, nnz, cap, cs) + 1024 <= smem_optin) {
            *ccap = cap;
            return cs;
        }
    }
    return 0;
}

__device__ __forceinline__ double ldcg(const double *p) { return __ldcg(p); }
// branch-free reciprocal (MUFU.RCP64H seed + two Newton steps; ~1 ulp): keeps the pivot chain of the sweep free of the
// slow-path branch of an IEEE division so the compiler can interleave it with the row update
__device__ __forceinline__ double fast_rcp(double d) {
    double x;
    asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(x) : "d"(d));
    double e = fma(-d, x, 1.0);
    x = fma(x, e, x);
    e = fma(-d, x, 1.0);
    x = fma(x, e, x);
    e = fma(-d, x, 1.0);
    return fma(x, e, x);
}

// sparse dot product of one compressed row / column (entries [p0, p1) of a packed view: inner index | value position << PACK_BITS)
// with a shared-memory vector; everything it touches is in shared memory
// `dummy` is a packed entry whose value slot holds 0.0: the last group of four is padded with it instead of a serial tail loop.
__device__ __forceinline__ double packed_dot(const unsigned *pack, int p0, int p1, const double *vals, const double *vec, unsigned dummy) {
    double a0 = 0.0, a1 = 0.0, a2 = 0.0, a3 = 0.0;
    for (int p = p0; p < p1; p += 4) {
        const unsigned e0 = pack[p];
        const unsigned e1 = p + 1 < p1 ? pack[p + 1] : dummy;
        const unsigned e2 = p + 2 < p1 ? pack[p + 2] : dummy;
        const unsigned e3 = p + 3 < p1 ? pack[p + 3] : dummy;
        a0 = fma(vals[e0 >> PACK_BITS], vec[e0 & PACK_MASK], a0);
        a1 = fma(vals[e1 >> PACK_BITS], vec[e1 & PACK_MASK], a1);
        a2 = fma(vals[e2 >> PACK_BITS], vec[e2 & PACK_MASK], a2);
        a3 = fma(vals[e3 >> PACK_BITS], vec[e3 & PACK_MASK], a3);
    }
    return (a0 + a1) + (a2 + a3);
}

// ---- per-iteration exchanges without a cluster barrier: st.async into the peers' shared memory, completion counted on THEIR mbarrier ----
// (cluster.sync costs ~400-900 cycles per use here: barrier.cluster with release/acquire semantics plus the wait for the slowest CTA;
// the consumer of an all-gather only needs the bytes to have landed)
__device__ __forceinline__ unsigned cl_smem_u32(const void *p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ unsigned cl_mapa(unsigned local_addr, unsigned rank) {
    unsigned r;
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(local_addr), "r"(rank));
    return r;
}
__device__ __forceinline__ void cl_st_async(unsigned remote_addr, double v, unsigned remote_mbar) {
    asm volatile("st.async.weak.shared::cluster.mbarrier::complete_tx::bytes.f64 [%0], %1, [%2];" ::"r"(remote_addr), "d"(v), "r"(remote_mbar)
                 : "memory");
}
__device__ __forceinline__ void cl_mbar_init(unsigned bar, unsigned count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void cl_mbar_arm(unsigned bar, unsigned bytes) {  // the one arrival of a phase + its transaction bytes
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void cl_mbar_wait(unsigned bar, unsigned parity) {
    const long long t0 = clock64();
    for (;;) {
        unsigned done;
        asm volatile(
            "{\n\t"
            ".reg .pred p;\n\t"
            "mbarrier.try_wait.parity.acquire.cluster.shared::cta.b64 p, [%1], %2;\n\t"
            "selp.u32 %0, 1, 0, p;\n\t"
            "}"
            : "=r"(done)
            : "r"(bar), "r"(parity)
            : "memory");
        if (done) break;
        if (clock64() - t0 > (1LL << 32)) __trap();  // ~2 s: a lost transfer must not hang the device
    }
}

// the same dot product over the ELL copy of a thread's entries: slot k of thread t at [k * CT + t]; `trips` groups of 4 slots
__device__ __forceinline__ double ell_dot(const double *ev, const int *ei, int trips, const double *vec, int tid) {
    double a0 = 0.0, a1 = 0.0, a2 = 0.0, a3 = 0.0;
    ev += tid;
    ei += tid;
    for (int g = 0; g < trips; ++g) {
        const int j0 = ei[0], j1 = ei[CT], j2 = ei[2 * CT], j3 = ei[3 * CT];
        const double v0 = ev[0], v1 = ev[CT], v2 = ev[2 * CT], v3 = ev[3 * CT];
        a0 = fma(v0, vec[j0], a0);
        a1 = fma(v1, vec[j1], a1);
        a2 = fma(v2, vec[j2], a2);
        a3 = fma(v3, vec[j3], a3);
        ev += 4 * CT;
        ei += 4 * CT;
    }
    return (a0 + a1) + (a2 + a3);
}

// FUSED: the launch is setup + solve of fresh instances (MODE_RESET | MODE_FACTOR | MODE_SOLVE, the hot path): the mode tests fold away at
// compile time and the kernel is the round-1 one, free of spills at its 255 registers. !FUSED: the object API's separate launches.
template <int CS, bool FUSED>
__global__ void __launch_bounds__(CT, 1) qp_cluster_kernel(KernelParams p, double *scratch) {
    extern __shared__ __align__(16) double smem_raw[];
    __shared__ int s_qp;
    __shared__ double s_wred[7][CNW];
    cg::cluster_group cluster = cg::this_cluster();
    const int rank = (int)cluster.block_rank();
    const int cid = blockIdx.x / CS;
    const int n = p.n, m = p.m, np = cluster_np(p.n);
    const int RS = np / CS, LD = RS + 2, LDT = RS + 4;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int fr = lane >> 2, fk = lane & 3;
    const SparseA sp = p.sp;
    const int nnz = sp.nnz;
    const int ccap = sp.col_slice_cap;
    ClusterSmem s = carve_cluster(smem_raw, np, m, nnz, ccap, CS);
    double *vals = s.X;
    unsigned *cpack = reinterpret_cast<unsigned *>(s.X + (nnz + 2 - (nnz & 1)));  // vals has one extra slot holding 0.0
    const unsigned dummy = (unsigned)nnz << PACK_BITS;
    unsigned *rpack = cpack + ccap;
    int *couter = reinterpret_cast<int *>(rpack + nnz), *router = couter + RS + 1;  // couter: own columns, relative to the slice
    double *Rb = s.X, *Tm = s.X + 2 * LDR * RCH, *Eb = Tm + (size_t)LDT * KB;
    const sqpb200_qp_settings st = p.s;
    const double sigma = st.sigma, alpha = st.alpha;
    // constraint rows owned by this CTA: one per thread
    const int RM = (m + CS - 1) / CS;
    const int row_lo = rank * RM;
    const int rows_own = max(0, min(m, row_lo + RM) - row_lo);
    const int TPR = (2 * RM <= CT) ? 2 : 1;  // threads per owned row (they split its stored entries)
    const bool has_row = tid / TPR < rows_own;
    const int my_row = row_lo + tid / TPR, row_half = tid % TPR;
    const bool row_writer = has_row && row_half == 0;
    // peers' copies of the exchanged vectors
    double *peer_sw[CS], *peer_xp[CS], *peer_red[CS];
    int *peer_qp[CS];
#pragma unroll
    for (int r = 0; r < CS; ++r) {
        peer_sw[r] = cluster.map_shared_rank(s.sw, r);
        peer_xp[r] = cluster.map_shared_rank(s.xp, r);
        peer_red[r] = cluster.map_shared_rank(s.red, r);
        peer_qp[r] = cluster.map_shared_rank(&s_qp, r);
    }
    double *scr_base = scratch + (size_t)cid * cluster_scratch_doubles();

    // ---- ELL copies of the two per-iteration sparse dot products --------------------------------------------------------------
    // The compressed views the factorisation works on make poor operands for the iteration: a thread walks its few entries through
    // three dependent shared-memory loads each (packed index -> value, vector operand), and neighbouring threads start 4-8 words apart
    // (4- to 8-way bank conflicts on the first two). While no factorisation is running, the aliased region X therefore holds the
    // entries thread by thread instead: entry k of thread t at [k * CT + t] (value and inner index in separate arrays), conflict-free
    // and with no index -> value dependency. Same entries per thread, same order, same accumulators as packed_dot: identical sums.
    // The split of a row / column over its threads is a property of the batch-shared pattern: computed once per launch.
    const int TPC = CT / RS, bcol = tid / TPC, bpart = tid % TPC;
    int ell_rlo = 0, ell_rcnt = 0, ell_clo = 0, ell_ccnt = 0;
    if (has_row) {  // the TPR threads of a row split its entries; the first part is a multiple of 4 entries
        const int p0 = __ldg(sp.row_outer + my_row), p1 = __ldg(sp.row_outer + my_row + 1);
        const int pm = TPR == 2 ? p0 + (((p1 - p0 + 1) / 2 + 3) & ~3) : p1;
        ell_rlo = row_half == 0 ? p0 : min(pm, p1);
        ell_rcnt = (row_half == 0 ? min(pm, p1) : p1) - ell_rlo;
    }
    {
        const int p0 = __ldg(sp.col_outer + min(n, RS * rank + bcol)), p1 = __ldg(sp.col_outer + min(n, RS * rank + bcol + 1));
        const int chunk = (((p1 - p0) + TPC - 1) / TPC + 3) & ~3;
        ell_clo = min(p0 + bpart * chunk, p1);
        ell_ccnt = min(ell_clo + chunk, p1) - ell_clo;
    }
    __shared__ __align__(8) unsigned long long s_xbar[2];  // mbarriers of the two per-iteration exchanges (w all-gather, x~ partials)
    const unsigned barA = cl_smem_u32(&s_xbar[0]), barB = cl_smem_u32(&s_xbar[1]);
    unsigned parA = 0, parB = 0;  // phase parity of the next use (every CTA of the cluster runs the same sequence of exchanges)
    if (tid == 0) {
        cl_mbar_init(barA, 1);
        cl_mbar_init(barB, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __shared__ int s_ellk[2];
    if (tid < 2) s_ellk[tid] = 0;
    __syncthreads();
    atomicMax(&s_ellk[0], ell_rcnt);
    atomicMax(&s_ellk[1], ell_ccnt);
    __syncthreads();
    const int KR = (s_ellk[0] + 3) & ~3, KC = (s_ellk[1] + 3) & ~3;  // slots per thread (CTA-wide maxima, multiples of 4)
    const bool use_ell = (size_t)(KR + KC) * CT * 12 <= sizeof(double) * cluster_x_doubles(np, m, nnz, ccap, CS);
    double *ellv_r = s.X, *ellv_c = s.X + (size_t)KR * CT;
    int *elli_r = reinterpret_cast<int *>(ellv_c + (size_t)KC * CT), *elli_c = elli_r + (size_t)KR * CT;
    // loop trips of 4 entries: uniform per warp (the padding slots hold value 0)
    const int trips_r = (__reduce_max_sync(0xffffffffu, ell_rcnt) + 3) >> 2, trips_c = (__reduce_max_sync(0xffffffffu, ell_ccnt) + 3) >> 2;

#ifdef SQPB200_CLUSTER_TIMING
    long long tph[16] = {0}, tlast = 0;
#endif
    cluster.sync();  // every CTA of the cluster has started: its shared memory may be written remotely from here on
    for (;;) {
        if (rank == 0 && tid == 0) {
            const int v = draw_qp(p);
#pragma unroll
            for (int r = 0; r < CS; ++r) *peer_qp[r] = v;
        }
        cluster.sync();
        const int local = s_qp;
        if (local >= p.count) break;
        const size_t b = (size_t)p.first + local;
        const double *P = p.P + b * n * n, *q = p.q + b * n, *l = p.l + b * m, *u = p.u + b * m;
        const double *gvals = sp.vals + b * (size_t)nnz;

        // Launch modes (qp_common.cuh): setup = RESET|FACTOR, update_qp = FACTOR, solve = LOAD_FACTOR|SOLVE, fused = RESET|FACTOR|SOLVE.
        // The distributed H^-1 never leaves the cluster's shared memory: a solve() after a separate setup()/update_qp() launch
        // rebuilds it from the stored constraint classes and rho (deterministic: the same factor setup computed), at the cost of
        // ~40 iterations' worth of work per instance.
        // (the mode bits are re-read from the kernel parameters where needed: this kernel sits at the 255-register limit)
#define m_reset (FUSED || (p.mode & MODE_RESET) != 0)
#define m_factor (FUSED || (p.mode & MODE_FACTOR) != 0)
#define m_solve (FUSED || (p.mode & MODE_SOLVE) != 0)
        int status = FUSED ? (int)SQPB200_UNSOLVED : p.status[b];
        int rho_updates = p.rho_updates[b] + (m_factor ? 1 : 0);  // rho_vec_update, qp.cpp:313
        double rho_est = p.rho_estimate[b], res_prim = p.res_prim[b], res_dual = p.res_dual[b];
        double rho = m_factor ? st.rho : p.rho[b];
        int iter_out = p.iter[b];
        // constraint class of row i: classified from the bounds by setup/update_qp (qp.cpp:31, :48), read back by solve
        auto row_type = [&](int i) -> int { return m_factor ? classify(__ldg(l + i), __ldg(u + i)) : (int)p.ctype[b * m + i]; };

        // values of this instance + the batch-shared pattern -> shared memory (the region is reused by the sweep: restaged after it)
        auto stage_sparse = [&]() {
            for (int e = tid; e < nnz; e += CT) {
                vals[e] = __ldg(gvals + e);
                rpack[e] = __ldg(sp.row_pack + e);
            }
            if (tid == 0) vals[nnz] = 0.0;
            const int cb = __ldg(sp.col_outer + min(n, RS * rank)), ce = __ldg(sp.col_outer + min(n, RS * rank + RS));
            for (int e = tid; e < ce - cb; e += CT) cpack[e] = __ldg(sp.col_pack + cb + e);
            for (int j = tid; j <= RS; j += CT) couter[j] = __ldg(sp.col_outer + min(n, RS * rank + j)) - cb;
            for (int i = tid; i <= m; i += CT) router[i] = __ldg(sp.row_outer + i);
        };
        // the ELL copies (see the kernel prologue): slot k of this thread <- its k-th entry of the row / column view, zero-padded
        // (a padding slot repeats the thread's first inner index, so that it multiplies 0 with an operand the dot product reads anyway)
        auto stage_ell = [&]() {
            {
                const int jpad = ell_rcnt > 0 ? (int)(__ldg(sp.row_pack + ell_rlo) & PACK_MASK) : 0;
#pragma unroll 4
                for (int k = 0; k < KR; ++k) {
                    double v = 0.0;
                    int j = jpad;
                    if (k < ell_rcnt) {
                        const unsigned e = __ldg(sp.row_pack + ell_rlo + k);
                        v = __ldg(gvals + (e >> PACK_BITS));
                        j = (int)(e & PACK_MASK);
                    }
                    ellv_r[k * CT + tid] = v;
                    elli_r[k * CT + tid] = j;
                }
            }
            {
                const int jpad = ell_ccnt > 0 ? (int)(__ldg(sp.col_pack + ell_clo) & PACK_MASK) : 0;
#pragma unroll 4
                for (int k = 0; k < KC; ++k) {
                    double v = 0.0;
                    int j = jpad;
                    if (k < ell_ccnt) {
                        const unsigned e = __ldg(sp.col_pack + ell_clo + k);
                        v = __ldg(gvals + (e >> PACK_BITS));
                        j = (int)(e & PACK_MASK);
                    }
                    ellv_c[k * CT + tid] = v;
                    elli_c[k * CT + tid] = j;
                }
            }
        };
        bool ell_live = false;  // X holds the ELL copies (the compressed views must be restaged before a factorisation)
        stage_sparse();
        // (A v)_i for the owned row of this thread: the TPR threads of a row split its entries and add the halves
        auto own_rowdot = [&](const double *vec) -> double {
            double acc = 0.0;
            if (use_ell) {
                acc = ell_dot(ellv_r, elli_r, trips_r, vec, tid);  // (threads without a row hold zero slots)
            } else if (has_row) {
                const int p0 = router[my_row], p1 = router[my_row + 1];
                const int pm = TPR == 2 ? p0 + (((p1 - p0 + 1) / 2 + 3) & ~3) : p1;  // first half: a multiple of 4 entries
                const int lo_ = row_half == 0 ? p0 : min(pm, p1), hi_ = row_half == 0 ? min(pm, p1) : p1;
                acc = packed_dot(rpack, lo_, hi_, vals, vec, dummy);
            }
            if (TPR == 2) acc += __shfl_xor_sync(0xffffffffu, acc, 1);
            return acc;
        };
        for (int j = tid; j < np; j += CT) {
            s.sq[j] = j < n ? __ldg(q + j) : 0.0;
            s.sx[j] = (j < n && !m_reset) ? p.x[b * n + j] : 0.0;
            s.sxt[j] = 0.0;
        }
        // per-row state of the owned constraint rows lives in registers (qp.cpp:16-18, 31-32)
        double zr = 0.0, yr = 0.0, lo = 0.0, up = 0.0, rhor = 1.0, rinv = 1.0;
        int typ = SQPB200_LOOSE_BOUNDS;
        if (has_row) {
            lo = l[my_row];
            up = u[my_row];
            typ = row_type(my_row);
            if (!m_reset) {
                zr = p.z[b * m + my_row];
                yr = p.y[b * m + my_row];
            }
            rhor = rho_of(typ, rho);
            rinv = 1.0 / rhor;
        }
        __syncthreads();

        // ---- (re)factorisation: S <- -(P_lowsym + sigma I + A^T diag(rho) A)^-1, row slice by row slice -----------------
        // Needs vals staged; leaves vals staged. Returns the same value in every thread of the cluster.
        auto factorize = [&]() -> bool {
#ifdef SQPB200_CLUSTER_TIMING
            tlast = clock64();
#endif
            if (ell_live) {  // refactorisation: the compressed views come back first
                __syncthreads();
                stage_sparse();
                __syncthreads();
                ell_live = false;
            }
            // rho of every row (form H touches all rows of a column): recomputed from the bounds, as classified at setup
            for (int i = tid; i < m; i += CT) s.sw[i] = rho_of(row_type(i), rho);
            // P_lowsym + sigma I (LDLT<Lower> reads the lower triangle only); padded variables get a unit diagonal
            // (two passes so that both triangles are read along P's columns: lanes over rows for j <= i, lanes over columns above)
#pragma unroll 8
            for (int e = tid; e < RS * np; e += CT) {
                const int r = e % RS, j = e / RS, i = RS * rank + r;
                if (i < n && j < n) {
                    if (i >= j) s.S[r + LD * j] = __ldg(P + i + (size_t)n * j) + (i == j ? sigma : 0.0);
                } else {
                    s.S[r + LD * j] = (i == j) ? 1.0 : 0.0;
                }
            }
#pragma unroll 8
            for (int e = tid; e < RS * np; e += CT) {
                const int j = e % np, r = e / np, i = RS * rank + r;
                if (i < n && j < n && j > i) s.S[r + LD * j] = __ldg(P + j + (size_t)n * i);
            }
            __syncthreads();
            // + A^T diag(rho) A: row i of H gathers, for every stored (k, i), rho_k A_ki times row k of A. A warp works on FOUR rows of
            // the slice at once, eight lanes per row (a row of A holds ~8 entries here: a whole warp per row left three quarters of the
            // lanes idle): the lanes of a group cover the entries of row k (distinct columns: no conflicts), the stored entries of
            // column i are fetched by the group in batches of eight and handed round by shuffles. Every H entry receives its
            // contributions in the order of the column's entries, whatever the grouping: the sums are those of a warp per row.
            {
                const int g8 = lane & 24, l8 = lane & 7;
                for (int rb = 4 * warp; rb < RS; rb += 4 * CNW) {
                    const int r = rb + (lane >> 3), i = RS * rank + r;
                    const bool rvalid = r < RS && i < n;
                    const int c0 = rvalid ? couter[r] : 0, c1 = rvalid ? couter[r + 1] : 0;
                    const int len_max = __reduce_max_sync(0xffffffffu, c1 - c0);
                    for (int base = 0; base < len_max; base += 8) {
                        const int cnt = min(8, max(0, c1 - c0 - base));  // entries of this group's batch
                        double f_l = 0.0;
                        int r0_l = 0, r1_l = 0;
                        if (l8 < cnt) {
                            const unsigned ec = cpack[c0 + base + l8];
                            const int k = (int)(ec & PACK_MASK);
                            f_l = s.sw[k] * vals[ec >> PACK_BITS];
                            r0_l = router[k];
                            r1_l = router[k + 1];
                        }
                        const int cnt_max = __reduce_max_sync(0xffffffffu, cnt);
                        for (int t = 0; t < cnt_max; ++t) {
                            const double f = __shfl_sync(0xffffffffu, f_l, g8 | t);
                            const int r0 = __shfl_sync(0xffffffffu, r0_l, g8 | t), r1 = __shfl_sync(0xffffffffu, r1_l, g8 | t);
                            if (t < cnt) {
                                for (int pr = r0 + l8; pr < r1; pr += 8) {
                                    const unsigned er = rpack[pr];
                                    double *dst = s.S + r + LD * (int)(er & PACK_MASK);
                                    *dst = fma(f, vals[er >> PACK_BITS], *dst);
                                }
                            }
                            __syncwarp();
                        }
                    }
                }
            }
            __syncthreads();  // the sweep buffers alias vals from here on
            TCK(0)

            bool ok = true;
            const int nblk = np / KB;
            const size_t SCR1 = (size_t)KB * 256 + KB * KB + 8;
            // ---- pieces of one block step -------------------------------------------------------------------------------------
            // publish the pivot rows S[K, :] of block kb (owned by this CTA) with `nthr` threads (t = index among them; a multiple of 32)
            auto publish_R = [&](int kb, int t, int nthr) {
                const int k0l = kb * KB - (kb * KB / RS) * RS;
                double *scrR = scr_base + (size_t)(kb & 1) * SCR1;
                for (int e = t; e < KB * np; e += nthr) scrR[e] = s.S[(k0l + (e % KB)) + LD * (e / KB)];
                __threadfence();  // belt and braces: the cluster barrier that ends the step is already a release/acquire at cluster scope
            };
            // Symmetric sweep of the 32 x 32 pivot block of block kb by ONE warp, a row per lane in registers (the pivots are a
            // serial chain: a CTA-wide version pays a barrier per pivot), published as E^-1 with the failure flag. Per step the lanes
            // exchange the pivot columns through shared memory (details at the loop below).
            auto sweep_block = [&](int kb) {
                const int k0 = kb * KB, k0l = k0 - (k0 / RS) * RS;
                double *scrE = scr_base + (size_t)(kb & 1) * SCR1 + (size_t)KB * np, *scrF = scrE + KB * KB;
                double row[KB];
#pragma unroll
                for (int j = 0; j < KB; ++j) row[j] = s.S[(k0l + lane) + LD * (k0 + j)];
                // TWO pivots per step (16 steps instead of 32: the step is a serial chain of exchange -> multipliers -> row update). With
                // K = {pv, pv + 1} and E = S[K][K]: S <- S - S[:,K] E^-1 S[K,:], S[:,K] <- S[:,K] E^-1, S[K,:] <- E^-1 S[K,:], S[K][K] <- -E^-1.
                // The pivots of the unpivoted LDL^T are d_pv = E00 and d_pv+1 = det(E) / E00: a zero or NaN in either is a failure
                // (Eigen::LDLT::info() != Success). The inverse of the NEXT pivot block is formed by every lane from the exchanged
                // values with the very operations the owner lanes' update performs, so it overlaps the row update.
                double *cv = Eb;  // 2 x (2 KB + 4) doubles of exchange space (E^-1 of the current step is no longer needed)
                bool bad = false;
                double e00 = 0.0, e01 = 0.0, e11 = 0.0;
                {
                    const double a = __shfl_sync(0xffffffffu, row[0], 0), bq = __shfl_sync(0xffffffffu, row[0], 1);
                    const double cq = __shfl_sync(0xffffffffu, row[1], 1);
                    const double det = fma(a, cq, -(bq * bq));
                    if (!(fabs(a) > 0.0) || !(fabs(det) > 0.0)) {
                        bad = true;
                    } else {
                        const double id = fast_rcp(det);
                        e00 = cq * id;
                        e01 = -bq * id;
                        e11 = a * id;
                    }
                }
#pragma unroll
                for (int pv = 0; pv < KB; pv += 2) {
                    if (bad) break;
                    double *ca = cv + ((pv >> 1) & 1) * (2 * KB + 4), *cb = ca + KB;  // exchanged columns pv and pv + 1
                    ca[lane] = row[pv];
                    cb[lane] = row[pv + 1];
                    if (pv + 2 < KB) {  // the next pivot block, before this update
                        if (lane == pv + 2) cb[KB] = row[pv + 2];
                        if (lane == pv + 3) {
                            cb[KB + 1] = row[pv + 2];
                            cb[KB + 2] = row[pv + 3];
                        }
                    }
                    __syncwarp();
                    const bool pa = lane == pv, pb = lane == pv + 1, piv = pa || pb;
                    const double xa = row[pv], xb = row[pv + 1];
                    const double wA = fma(xa, e00, xb * e01), wB = fma(xa, e01, xb * e11);  // this row of S[:,K] E^-1
                    const double mA = piv ? (pa ? e00 : e01) : -wA, mB = piv ? (pa ? e01 : e11) : -wB;
                    double n00 = 0.0, n01 = 0.0, n11 = 0.0;
                    bool bad_next = false;
                    if (pv + 2 < KB) {
                        const double a2 = ca[pv + 2], b2 = cb[pv + 2], a3 = ca[pv + 3], b3 = cb[pv + 3];
                        const double wA2 = fma(a2, e00, b2 * e01), wB2 = fma(a2, e01, b2 * e11);
                        const double wA3 = fma(a3, e00, b3 * e01), wB3 = fma(a3, e01, b3 * e11);
                        const double f22 = fma(-wB2, b2, fma(-wA2, a2, cb[KB]));
                        const double f32 = fma(-wB3, b2, fma(-wA3, a2, cb[KB + 1]));
                        const double f33 = fma(-wB3, b3, fma(-wA3, a3, cb[KB + 2]));
                        const double det = fma(f22, f33, -(f32 * f32));
                        if (!(fabs(f22) > 0.0) || !(fabs(det) > 0.0)) {
                            bad_next = true;
                        } else {
                            const double id = fast_rcp(det);
                            n00 = f33 * id;
                            n01 = -f32 * id;
                            n11 = f22 * id;
                        }
                    }
#pragma unroll
                    for (int j = 0; j < KB; j += 2) {
                        const double2 a2 = *reinterpret_cast<const double2 *>(ca + j);
                        const double2 b2 = *reinterpret_cast<const double2 *>(cb + j);
                        const double b0 = piv ? 0.0 : row[j], b1 = piv ? 0.0 : row[j + 1];
                        row[j] = fma(mB, b2.x, fma(mA, a2.x, b0));
                        row[j + 1] = fma(mB, b2.y, fma(mA, a2.y, b1));
                    }
                    row[pv] = piv ? (pa ? -e00 : -e01) : wA;
                    row[pv + 1] = piv ? (pa ? -e01 : -e11) : wB;
                    if (pv + 2 < KB) {
                        if (bad_next) bad = true;
                        e00 = n00;
                        e01 = n01;
                        e11 = n11;
                    }
                }
                // row holds row `lane` of -(E^-1): publish E^-1 (column-major, coalesced over the lanes)
#pragma unroll
                for (int j = 0; j < KB; ++j) scrE[lane + KB * j] = bad ? 0.0 : -row[j];
                if (lane == 0) scrF[0] = bad ? 1.0 : 0.0;
                __threadfence();
            };
            // rank-32 update S[rows of block rb, j] <- S - T R[:, j] on the fp64 tensor cores for the column groups g0, g0 + gstep, ...
            // (a group = four 8-column tiles = four independent DMMA chains); the pivot rows R of step kb are staged chunk by chunk
            // by the NTHR participating threads (index st among them), the next chunk's L2 loads in flight while one is consumed.
            auto update_pass = [&](auto nthr_c, int kb, int rb, int g0, int gstep, int st, bool owner_cta, int k0l) {
                constexpr int NTHR = decltype(nthr_c)::value, PER = KB * RCH / NTHR;
                const int k0 = kb * KB;
                const double *scrR = scr_base + (size_t)(kb & 1) * SCR1;
                const bool pivot_rb = owner_cta && 8 * rb >= k0l && 8 * rb < k0l + KB;
                double af[KB / 4];
#pragma unroll
                for (int ks = 0; ks < KB / 4; ++ks) af[ks] = rb >= 0 ? -Tm[(8 * rb + fr) + LDT * (4 * ks + fk)] : 0.0;
                double rpre[PER];
#pragma unroll
                for (int t = 0; t < PER; ++t) rpre[t] = ldcg(scrR + st + NTHR * t);
                const int nch = np / RCH;
                for (int ch = 0; ch < nch; ++ch) {
                    double *Rc = Rb + (size_t)(ch & 1) * LDR * RCH;
#pragma unroll
                    for (int t = 0; t < PER; ++t) {
                        const int e = st + NTHR * t;
                        Rc[(e % KB) + LDR * (e / KB)] = rpre[t];
                    }
                    if (ch + 1 < nch) {
#pragma unroll
                        for (int t = 0; t < PER; ++t) rpre[t] = ldcg(scrR + (size_t)KB * RCH * (ch + 1) + st + NTHR * t);
                    }
                    // chunk ch staged; chunk ch-1's readers finished before they staged ch (two buffers)
                    if constexpr (NTHR == CT) __syncthreads();
                    else asm volatile("bar.sync 1, %0;" ::"n"(NTHR) : "memory");
                    if (rb < 0) continue;
                    if (gstep == 1) {
                        // both column groups of the chunk at once: eight independent DMMA chains per warp (the chains are latency bound)
                        const bool live0 = ch * RCH != k0, live1 = ch * RCH + KB != k0;
                        double *cp = s.S + (8 * rb + fr) + LD * (ch * RCH + 2 * fk);
                        double c0[8], c1[8];
#pragma unroll
                        for (int t = 0; t < 8; ++t) {  // (the pivot columns are neither read nor written here: other warps may already be turning them into T)
                            const bool skip = pivot_rb || !(t < 4 ? live0 : live1);
                            c0[t] = skip ? 0.0 : cp[LD * 8 * t];
                            c1[t] = skip ? 0.0 : cp[LD * (8 * t + 1)];
                        }
                        const double *bp = Rc + fk + LDR * fr;
#pragma unroll
                        for (int ks = 0; ks < KB / 4; ++ks) {
#pragma unroll
                            for (int t = 0; t < 8; ++t) {
                                const double bf = bp[4 * ks + LDR * 8 * t];
                                asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0, %1}, {%2}, {%3}, {%0, %1};"
                                             : "+d"(c0[t]), "+d"(c1[t])
                                             : "d"(af[ks]), "d"(bf));
                            }
                        }
#pragma unroll
                        for (int t = 0; t < 8; ++t) {
                            if (t < 4 ? live0 : live1) {  // the pivot columns become T
                                cp[LD * 8 * t] = c0[t];
                                cp[LD * (8 * t + 1)] = c1[t];
                            }
                        }
                        continue;
                    }
                    for (int g = g0; g < RCH / KB; g += gstep) {
                        const int col0 = ch * RCH + KB * g;
                        if (col0 == k0) continue;  // the pivot columns become T
                        double *cp = s.S + (8 * rb + fr) + LD * (col0 + 2 * fk);
                        double c0[4], c1[4];
#pragma unroll
                        for (int t = 0; t < 4; ++t) {
                            c0[t] = pivot_rb ? 0.0 : cp[LD * 8 * t];
                            c1[t] = pivot_rb ? 0.0 : cp[LD * (8 * t + 1)];
                        }
                        const double *bp = Rc + fk + LDR * (KB * g + fr);
#pragma unroll
                        for (int ks = 0; ks < KB / 4; ++ks) {
#pragma unroll
                            for (int t = 0; t < 4; ++t) {
                                const double bf = bp[4 * ks + LDR * 8 * t];
                                asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0, %1}, {%2}, {%3}, {%0, %1};"
                                             : "+d"(c0[t]), "+d"(c1[t])
                                             : "d"(af[ks]), "d"(bf));
                            }
                        }
#pragma unroll
                        for (int t = 0; t < 4; ++t) {
                            cp[LD * 8 * t] = c0[t];
                            cp[LD * (8 * t + 1)] = c1[t];
                        }
                    }
                }
            };

            // block 0 is published up front; every later block is published by its owner DURING the previous step (look-ahead),
            // so the serial pivot chain of the 32 x 32 sweep hides behind the tensor-core update of the other row blocks
            if (rank == 0) {
                publish_R(0, tid, CT);
                if (warp == 0) sweep_block(0);
            }
            for (int kb = 0; kb < nblk; ++kb) {
                const int k0 = kb * KB;
                const int owner = k0 / RS, k0l = k0 - owner * RS;
                const double *scrE = scr_base + (size_t)(kb & 1) * SCR1 + (size_t)KB * np, *scrF = scrE + KB * KB;
                TCK(1)
                cluster.sync();  // block kb is published; every CTA has finished step kb - 1
                TCK(2)
                if (ldcg(scrF) != 0.0) {
                    ok = false;
                    break;
                }
                // E^-1 -> Eb (every CTA)
                {
                    double ev[KB * KB / CT];
#pragma unroll
                    for (int t = 0; t < KB * KB / CT; ++t) ev[t] = ldcg(scrE + tid + CT * t);
#pragma unroll
                    for (int t = 0; t < KB * KB / CT; ++t) {
                        const int e = tid + CT * t;
                        Eb[(e % KB) * LDE + (e / KB)] = ev[t];
                    }
                }
                __syncthreads();
                // T = S[own rows, K] E^-1 on the tensor cores (8-row block x four 8-column tiles per warp, K = 32); the owner's
                // pivot rows carry -E^-1 instead (their update then yields E^-1 S[K, :])
                {
                    const int RBN = RS / 8, rb = warp % RBN, tsplit = CNW / RBN, tpart = warp / RBN;
                    const bool pivot_rb = rank == owner && 8 * rb >= k0l && 8 * rb < k0l + KB;
                    double t0[4] = {0, 0, 0, 0}, t1[4] = {0, 0, 0, 0};
#pragma unroll
                    for (int ks = 0; ks < KB / 4; ++ks) {
                        const double av = s.S[(8 * rb + fr) + LD * (k0 + 4 * ks + fk)];
#pragma unroll
                        for (int jt = 0; jt < 4; ++jt) {
                            if (jt % tsplit != tpart) continue;
                            const double bv = Eb[(4 * ks + fk) * LDE + 8 * jt + fr];
                            asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0, %1}, {%2}, {%3}, {%0, %1};"
                                         : "+d"(t0[jt]), "+d"(t1[jt])
                                         : "d"(av), "d"(bv));
                        }
                    }
#pragma unroll
                    for (int jt = 0; jt < 4; ++jt) {
                        if (jt % tsplit != tpart) continue;
                        const int r = 8 * rb + fr, c = 8 * jt + 2 * fk;
                        Tm[r + LDT * c] = pivot_rb ? -Eb[(r - k0l) * LDE + c] : t0[jt];
                        Tm[r + LDT * (c + 1)] = pivot_rb ? -Eb[(r - k0l) * LDE + c + 1] : t1[jt];
                    }
                }
                __syncthreads();
                TCK(3)
                const int RBN = RS / 8;
                const bool next_owner = kb + 1 < nblk && rank == (k0 + KB) / RS;
                if (!next_owner) {
                    update_pass(std::integral_constant<int, CT>{}, kb, warp % RBN, warp / RBN, CNW / RBN, tid, rank == owner, k0l);
                    // pivot columns of the slice <- T (nobody reads S[own, K] any more: T was formed before the last barrier)
                    for (int e = tid; e < RS * KB; e += CT) s.S[(e % RS) + LD * (k0 + e / RS)] = Tm[(e % RS) + LDT * (e / RS)];
                    __syncthreads();
                } else {
                    // look-ahead: first the four row blocks that hold the NEXT pivot rows, then publish them while the rest is updated
                    const int k1l = (k0 + KB) - ((k0 + KB) / RS) * RS, rbn0 = k1l / 8;
                    update_pass(std::integral_constant<int, CT>{}, kb, rbn0 + (warp & 3), warp >> 2, 2, tid, rank == owner, k0l);
                    __syncthreads();
                    for (int e = tid; e < RS * KB; e += CT) s.S[(e % RS) + LD * (k0 + e / RS)] = Tm[(e % RS) + LDT * (e / RS)];
                    __syncthreads();
                    if (warp == CNW - 1) {
                        sweep_block(kb + 1);
                    } else if (warp >= 4) {
                        publish_R(kb + 1, tid - 128, 32 * (CNW - 5));
                    } else {
                        // the four row blocks that are not the next pivot rows (none when the slice is a single 32-row block)
                        const int rbo = RBN > 4 ? (rbn0 == 0 ? 4 : 0) + warp : -1;
                        update_pass(std::integral_constant<int, 128>{}, kb, rbo, 0, 1, tid, rank == owner, k0l);
                    }
                    __syncthreads();
                }
                TCK(4)
            }
            TCK(5)
            // the instance's sparse data comes back into the aliased region: as ELL copies for the iterations when they fit
            __syncthreads();
            if (use_ell) {
                stage_ell();
                ell_live = true;
            } else {
                stage_sparse();
            }
            __syncthreads();
            TCK(6)
            return ok;
        };

#ifdef SQPB200_CLUSTER_TIMING
        long long tq0 = clock64();
#endif
        // The mat-vec operands of this thread (2 rows x CPG columns of the slice) are copied into registers after every
        // factorisation: the iteration then reads only b from shared memory.
        // thread j keeps column j of the slice (RS <= 64 values): x~ = -(S^T-slice) b_slice summed over the cluster (S is symmetric)
        double sreg[64];
        auto load_slice = [&]() {
            const double *col = s.S + (size_t)LD * min(tid, np - 1);
#pragma unroll
            for (int r = 0; r < 64; r += 2) {
                if (r < RS) {
                    const double2 v = *reinterpret_cast<const double2 *>(col + r);
                    sreg[r] = v.x;
                    sreg[r + 1] = v.y;
                } else {
                    sreg[r] = sreg[r + 1] = 0.0;
                }
            }
        };
        // b (own slice) = sigma x - q + A^T w for the own columns: TPC threads per column split its stored entries
        auto own_coldot = [&](const double *vec) -> double {
            double acc;
            if (use_ell) {
                acc = ell_dot(ellv_c, elli_c, trips_c, vec, tid);
            } else {
                const int p0 = couter[bcol], p1 = couter[bcol + 1];
                const int chunk = (((p1 - p0) + TPC - 1) / TPC + 3) & ~3;
                const int lo_ = min(p0 + bpart * chunk, p1), hi_ = min(lo_ + chunk, p1);
                acc = packed_dot(cpack, lo_, hi_, vals, vec, dummy);
            }
            for (int o = 1; o < TPC; o <<= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
            return acc;
        };
        // setup / update_qp factor (qp.cpp:34-43, :51-61); a separate solve() rebuilds the factor unless it is a no-op (qp.cpp:68-71)
        bool ok = true;
        if (m_factor || (m_solve && status != SQPB200_UNINITIALIZED && status != SQPB200_NUMERICAL_ISSUES)) {
            ok = factorize();
            load_slice();
        }
#ifdef SQPB200_CLUSTER_TIMING
        long long tq1 = clock64();
#endif
        if (m_factor) status = ok ? SQPB200_UNSOLVED : SQPB200_NUMERICAL_ISSUES;  // qp.cpp:39-43

        long long executed = 0;
        if (m_solve && ok && status != SQPB200_UNINITIALIZED && status != SQPB200_NUMERICAL_ISSUES) {
            int iter;
            for (iter = 1; iter <= st.max_iter; ++iter) {
#ifdef SQPB200_CLUSTER_TIMING
                tlast = clock64();
#endif
                // w = rho .* z - y for the owned rows -> every CTA's copy
#ifdef SQPB200_CLUSTER_SYNC_EXCHANGE
                if (row_writer) {
                    const double wv = rhor * zr - yr;
#pragma unroll
                    for (int r = 0; r < CS; ++r) peer_sw[r][my_row] = wv;
                }
                cluster.sync();
#else
                // (a peer overwrites sw only after it has received this CTA's x~ partials of the previous iteration, which were sent
                // after the last read of sw: no second buffer, no barrier)
                if (tid == 0) cl_mbar_arm(barA, (unsigned)m * 8u);
                if (row_writer) {
                    const double wv = rhor * zr - yr;
                    const unsigned dst = cl_smem_u32(s.sw + my_row);
#pragma unroll
                    for (int r = 0; r < CS; ++r) cl_st_async(cl_mapa(dst, r), wv, cl_mapa(barA, r));
                }
                cl_mbar_wait(barA, parA);
                parA ^= 1;
#endif
                TCK(8)
                // b (own columns) = sigma x - q + A^T w; padded entries stay 0
                {
                    const int j = RS * rank + bcol;
                    const double g = own_coldot(s.sw);
                    if (bpart == 0) s.sb[bcol] = j < n ? fma(sigma, s.sx[j], g - s.sq[j]) : 0.0;
                }
                __syncthreads();
                TCK(9)
                // partial x~ = (S slice)^T b_slice, one column per thread out of registers -> every CTA's xp[rank]
                if (tid < np) {
                    double a0 = 0.0, a1 = 0.0, a2 = 0.0, a3 = 0.0;
#pragma unroll
                    for (int r = 0; r < 64; r += 4) {
                        if (r < RS) {
                            const double2 b0 = *reinterpret_cast<const double2 *>(s.sb + r);
                            const double2 b1 = *reinterpret_cast<const double2 *>(s.sb + r + 2);
                            a0 = fma(sreg[r], b0.x, a0);
                            a1 = fma(sreg[r + 1], b0.y, a1);
                            a2 = fma(sreg[r + 2], b1.x, a2);
                            a3 = fma(sreg[r + 3], b1.y, a3);
                        }
                    }
                    const double xpart = (a0 + a1) + (a2 + a3);
#ifdef SQPB200_CLUSTER_SYNC_EXCHANGE
#pragma unroll
                    for (int r = 0; r < CS; ++r) peer_xp[r][rank * np + tid] = xpart;
#else
                    const unsigned dst = cl_smem_u32(s.xp + rank * np + tid);
#pragma unroll
                    for (int r = 0; r < CS; ++r) cl_st_async(cl_mapa(dst, r), xpart, cl_mapa(barB, r));
#endif
                }
                TCK(10)
#ifdef SQPB200_CLUSTER_SYNC_EXCHANGE
                cluster.sync();
#else
                // (the x~ slots are overwritten by a peer only after it has received this CTA's w of the next iteration, sent after
                // the reads below)
                if (tid == 0) cl_mbar_arm(barB, (unsigned)(CS * np) * 8u);
                cl_mbar_wait(barB, parB);
                parB ^= 1;
#endif
                TCK(11)
                // x = alpha x~ + (1 - alpha) x (every CTA keeps all of x); z~ = A x~ and the z, y updates for the owned rows
                if (tid < np) {
                    double xt = s.xp[tid];
#pragma unroll
                    for (int r = 1; r < CS; ++r) xt += s.xp[r * np + tid];
                    xt = -xt;  // S = -(H^-1)
                    s.sxt[tid] = xt;
                    s.sx[tid] = alpha * xt + (1.0 - alpha) * s.sx[tid];
                }
                __syncthreads();
                const double zt = own_rowdot(s.sxt);
                if (has_row) {
                    const double zh = alpha * zt + (1.0 - alpha) * zr;
                    const double zn = box_project(zh + rinv * yr, lo, up);
                    yr = yr + rhor * (zh - zn);
                    zr = zn;
                }

                TCK(12)
                const bool chk = st.check_termination != 0 && iter % st.check_termination == 0;
                const bool adapt = st.adaptive_rho && st.adaptive_rho_interval > 0 && iter % st.adaptive_rho_interval == 0;
                if (chk || adapt) {
                    double mx[7] = {0, 0, 0, 0, 0, 0, 0};  // |Ax| |z| |Px| |A^T y| |q| |Ax - z| |Px + q + A^T y|
                    const double ax = own_rowdot(s.sx);
                    if (has_row) {
                        mx[0] = fabs(ax);
                        mx[1] = fabs(zr);
                        mx[5] = fabs(ax - zr);
                        if (row_half == 0) {
#pragma unroll
                            for (int r = 0; r < CS; ++r) peer_sw[r][my_row] = yr;  // all-gather y (w is rebuilt next iteration)
                        }
                    }
                    // P x for the rows of the own slice: partial sums over CT / RS column groups
                    {
                        const int KG = CT / RS, r = tid % RS, kg = tid / RS, j = RS * rank + r;
                        double acc = 0.0;
                        if (j < n) {
                            double a4[4] = {0.0, 0.0, 0.0, 0.0};
                            int k = kg;
                            for (; k + 15 * KG < n; k += 16 * KG) {  // sixteen independent L2 loads in flight
                                double pv[16];
#pragma unroll
                                for (int t = 0; t < 16; ++t) pv[t] = __ldg(P + j + (size_t)n * (k + t * KG));
#pragma unroll
                                for (int t = 0; t < 16; ++t) a4[t & 3] = fma(pv[t], s.sx[k + t * KG], a4[t & 3]);
                            }
                            for (; k < n; k += KG) a4[0] = fma(P[j + (size_t)n * k], s.sx[k], a4[0]);
                            acc = (a4[0] + a4[1]) + (a4[2] + a4[3]);
                        }
                        s.xp[kg * RS + r] = acc;
                    }
                    cluster.sync();
                    if (use_ell) {  // A^T y of the own columns through the column dot of the iteration (s.sb is free here)
                        const double g = own_coldot(s.sw);
                        if (bpart == 0) s.sb[bcol] = g;
                        __syncthreads();
                    }
                    if (tid < RS) {
                        const int KG = CT / RS, j = RS * rank + tid;
                        if (j < n) {
                            double px = 0.0;
                            for (int g = 0; g < KG; ++g) px += s.xp[g * RS + tid];
                            const double aty = use_ell ? s.sb[tid] : packed_dot(cpack, couter[tid], couter[tid + 1], vals, s.sw, dummy);
                            const double qv = s.sq[j];
                            mx[2] = fabs(px);
                            mx[3] = fabs(aty);
                            mx[4] = fabs(qv);
                            mx[6] = fabs(px + qv + aty);
                        }
                    }
#pragma unroll
                    for (int k = 0; k < 7; ++k) {
                        const double v = warp_max(mx[k]);
                        if (lane == 0) s_wred[k][warp] = v;
                    }
                    __syncthreads();
                    if (tid < 7) {
                        double v = s_wred[tid][0];
                        for (int w2 = 1; w2 < CNW; ++w2) v = s_wred[tid][w2] > v ? s_wred[tid][w2] : v;
#pragma unroll
                        for (int r = 0; r < CS; ++r) peer_red[r][rank * 8 + tid] = v;
                    }
                    cluster.sync();
#pragma unroll
                    for (int k = 0; k < 7; ++k) {
                        double v = s.red[k];
#pragma unroll
                        for (int r = 1; r < CS; ++r) v = s.red[r * 8 + k] > v ? s.red[r * 8 + k] : v;
                        mx[k] = v;
                    }
                    const double sc_p = fmax(mx[0], mx[1]);
                    const double sc_d = fmax(mx[2], fmax(mx[3], mx[4]));
                    res_prim = mx[5];
                    res_dual = mx[6];
                    if (chk && res_prim <= st.eps_abs + st.eps_rel * sc_p && res_dual <= st.eps_abs + st.eps_rel * sc_d) {
                        status = SQPB200_SOLVED;  // termination_criteria, qp.cpp:363-371
                        break;
                    }
                    if (adapt) {  // qp.cpp:125-144
                        const double new_rho = rho_estimate_clamped(rho, res_prim, res_dual, sc_p, sc_d);
                        rho_est = new_rho;
                        if (new_rho < rho / st.adaptive_rho_tolerance || new_rho > rho * st.adaptive_rho_tolerance) {
                            rho = new_rho;
                            rho_updates += 1;
                            rhor = rho_of(typ, rho);
                            rinv = 1.0 / rhor;
                            cluster.sync();  // every CTA is done with s.sw (y) before it is reused for the rho vector
                            if (!factorize()) {
                                status = SQPB200_NUMERICAL_ISSUES;  // qp.cpp:139-142
                                break;
                            }
                            load_slice();
                        }
                    }
                }
            }
            executed = iter <= st.max_iter ? iter : st.max_iter;
            if (iter > st.max_iter) status = SQPB200_MAX_ITER_EXCEEDED;  // qp.cpp:147-149
            iter_out = iter;                                            // qp.cpp:150
        }

#ifdef SQPB200_CLUSTER_TIMING
        if (tid == 0 && blockIdx.x < CS) { printf("qp %d rank %d: factor %lld cyc, solve %lld cyc, iters %d, rho_updates %d | formH %lld owner %lld sync %lld T %lld dmma %lld tail %lld restage %lld | it: sync1 %lld coldot %lld matvec %lld sync2 %lld rowdot %lld | own: publishR %lld sweep %lld\n", local, rank, tq1 - tq0, clock64() - tq1, iter_out, rho_updates, tph[0], tph[1], tph[2], tph[3], tph[4], tph[5], tph[6], tph[8], tph[9], tph[10], tph[11], tph[12], tph[13], tph[14]); for (int i_ = 0; i_ < 16; ++i_) tph[i_] = 0; }
#endif
        __syncthreads();
        if (tid < RS && RS * rank + tid < n) p.x[b * n + RS * rank + tid] = s.sx[RS * rank + tid];
        if (row_writer) {
            p.z[b * m + my_row] = zr;
            p.y[b * m + my_row] = yr;
            // the classes are written LAST: every CTA of the cluster read the old ones (row_type in a solve launch) before this point
            if (m_factor) p.ctype[b * m + my_row] = (signed char)typ;
        }
        if (rank == 0 && tid == 0) {
            p.status[b] = status;
            p.iter[b] = iter_out;
            p.rho_updates[b] = rho_updates;
            p.rho_estimate[b] = rho_est;
            p.res_prim[b] = res_prim;
            p.res_dual[b] = res_dual;
            p.rho[b] = rho;
            if (executed) atomicAdd(p.total_iters, (unsigned long long)executed);
        }
        cluster.sync();  // nobody is still reading this QP's exchanged vectors (or s_qp) when the next one starts
    }
    cluster.sync();  // no CTA exits while a peer may still touch its shared memory
#undef m_reset
#undef m_factor
#undef m_solve
}

template <int CS>
static cudaError_t cluster_config(size_t smem, int *max_clusters) {
    cudaError_t e = cudaFuncSetAttribute(qp_cluster_kernel<CS, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(CS * 16, 1, 1);
    cfg.blockDim = dim3(CT, 1, 1);
    cfg.dynamicSmemBytes = smem;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = CS;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    return cudaOccupancyMaxActiveClusters(max_clusters, qp_cluster_kernel<CS, true>, &cfg);
}

int cluster_max_clusters(int n, int m, int nnz, int ccap, int cs) {
    int mc = 0;
    const size_t smem = sizeof(double) * cluster_smem_doubles(cluster_np(n), m, nnz, ccap, cs);
    if ((cs == 8 ? cluster_config<8>(smem, &mc) : cluster_config<4>(smem, &mc)) != cudaSuccess) {
        cudaGetLastError();
        return 0;
    }
    return mc;
}
size_t cluster_scratch_bytes(int clusters) { return sizeof(double) * cluster_scratch_doubles() * (size_t)clusters; }

template <int CS>
static cudaError_t launch_cluster_cs(const KernelParams &p, int clusters, double *scratch, cudaStream_t stream, char *name, size_t name_len) {
    const size_t smem = sizeof(double) * cluster_smem_doubles(cluster_np(p.n), p.m, p.sp.nnz, p.sp.col_slice_cap, CS);
    const bool fused = p.mode == (MODE_RESET | MODE_FACTOR | MODE_SOLVE);
    auto kernel = fused ? qp_cluster_kernel<CS, true> : qp_cluster_kernel<CS, false>;
    cudaError_t e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    if (clusters > p.count) clusters = p.count;
    if (clusters < 1) return cudaErrorLaunchOutOfResources;
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(CS * clusters, 1, 1);
    cfg.blockDim = dim3(CT, 1, 1);
    cfg.dynamicSmemBytes = smem;
    cfg.stream = stream;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = CS;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    if (name) snprintf(name, name_len, "cluster<%d>/sparse x%d", CS, clusters);
    return cudaLaunchKernelEx(&cfg, kernel, p, scratch);
}

cudaError_t launch_cluster(const KernelParams &p, int clusters, double *scratch, cudaStream_t stream, char *name, size_t name_len) {
    return p.sp.cluster_size == 8 ? launch_cluster_cs<8>(p, clusters, scratch, stream, name, name_len)
                                  : launch_cluster_cs<4>(p, clusters, scratch, stream, name, name_len);
}

}  // namespace sqpb200
