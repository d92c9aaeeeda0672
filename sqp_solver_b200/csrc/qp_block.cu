// Blocked batched ADMM QP kernel for the medium shapes the register-tiled kernel cannot hold:
// 64 < n <= 256 (or m > 128), m <= 1024 -- BASELINE.json config 5's shape (n = 256, m = 512).
//
// One 256-thread CTA per QP, persistent CTAs on the atomic work queue. H = P_lowsym + sigma I + A^T diag(rho) A
// (n x n; see qp_common.cuh for why this is the reference's KKT solve) no longer fits on chip, so:
//   * H is formed on the fp64 tensor cores (DMMA m8n8k4, fragments straight from global/L2) into a per-QP global slab;
//   * it is factored IN PLACE by a right-looking BLOCKED L D L^T (panels of 32 columns staged in shared memory;
//     a zero/NaN pivot reports NUMERICAL_ISSUES like Eigen::LDLT::info()); every 32x32 diagonal block's unit factor
//     is inverted once (W_k = L_kk^-1);
//   * the per-iteration solve is a blocked substitution: 8 forward and 8 backward block steps for n = 256, each
//     one small triangular mat-vec with W_k plus one coalesced panel mat-vec streamed from L2 -- no explicit inverse
//     (the generic kernel's explicit n x n inverse cost ~20 ms per factorisation at n = 256);
//   * A may be dense (read through L2) or sparse with a batch-shared pattern: CSR and CSC copies of one instance's
//     values live in shared memory and both A-products become register-free gathers.
// Reference functions covered: the same list as qp_generic.cu (all of src/qp.cpp:11-371).
#include <cstdio>

#include "qp_common.cuh"

namespace sqpb200 {

constexpr int BT = 256;  // threads per CTA
constexpr int BNW = BT / 32;
constexpr int NB = 32;  // panel width
constexpr int WBLK = NB * NB + NB;  // doubles per diagonal block record: W (32x32, column-major) + 1/d (32)

__host__ __device__ inline int block_np(int n) { return (n + NB - 1) / NB * NB; }
__host__ __device__ inline size_t block_fact_doubles_hd(int n) {
    const size_t np = block_np(n);
    return np * np + (np / NB) * WBLK;
}
size_t block_fact_doubles(int n) { return block_fact_doubles_hd(n); }

struct BlockSmem {
    double *x, *xt, *b, *q, *t, *dinv;             // np each
    double *z, *y, *w, *l, *u, *rho;               // m each (1/rho is recomputed where it is used: one division per row and iteration buys
                                                   // the 4 KB that let a second CTA share the SM at n = 256, m = 512)
    double *panel;                                 // np x 32 (factorisation only)
    double *wsm;                                   // 32 x 32 + 32
    double *vals;                                  // nnz (sparse A: this instance's values), else unused
    signed char *type;                             // m
};
__device__ __forceinline__ BlockSmem carve_block(double *base, int np, int m, int nnz) {
    BlockSmem s;
    s.x = base; s.xt = s.x + np; s.b = s.xt + np; s.q = s.b + np; s.t = s.q + np; s.dinv = s.t + np;
    s.z = s.dinv + np; s.y = s.z + m; s.w = s.y + m; s.l = s.w + m; s.u = s.l + m; s.rho = s.u + m;
    s.panel = s.rho + m + (m & 1);
    s.wsm = s.panel + (size_t)np * NB;
    s.vals = s.wsm + WBLK;
    s.type = reinterpret_cast<signed char *>(s.vals + nnz);
    return s;
}
static size_t block_smem_bytes(int n, int m, int nnz = 0) {
    const size_t np = block_np(n);
    return sizeof(double) * (6 * np + 6 * (size_t)m + 1 + np * NB + WBLK + (size_t)nnz) + (size_t)m + 16;
}
bool block_supported(int n, int m, size_t smem_optin) { return n >= 1 && n <= 256 && m >= 0 && m <= 1024 && block_smem_bytes(n, m) <= smem_optin; }
// sparse A: this instance's nnz values are staged in shared memory next to the vectors
bool block_sparse_supported(int n, int m, int nnz, size_t smem_optin) {
    return n >= 1 && n <= 256 && m >= 1 && m <= 1024 && nnz >= 0 && block_smem_bytes(n, m, nnz) <= smem_optin;
}

// sparse row / column dot products against a shared-memory vector
__device__ __forceinline__ double sp_rowdot(const SparseA &sp, const double *vals, int i, const double *vec) {
    double acc = 0.0;
    const int e = sp.row_outer[i + 1];
    if (sp.row_perm) for (int p = sp.row_outer[i]; p < e; ++p) acc = fma(vals[sp.row_perm[p]], vec[sp.row_inner[p]], acc);
    else for (int p = sp.row_outer[i]; p < e; ++p) acc = fma(vals[p], vec[sp.row_inner[p]], acc);
    return acc;
}
__device__ __forceinline__ double sp_coldot(const SparseA &sp, const double *vals, int j, const double *vec) {
    double acc = 0.0;
    const int e = sp.col_outer[j + 1];
    if (sp.col_perm) for (int p = sp.col_outer[j]; p < e; ++p) acc = fma(vals[sp.col_perm[p]], vec[sp.col_inner[p]], acc);
    else for (int p = sp.col_outer[j]; p < e; ++p) acc = fma(vals[p], vec[sp.col_inner[p]], acc);
    return acc;
}

// H (lower triangle) for sparse A: column j of A^T diag(rho) A is sum over the stored (k, j) of rho_k A_kj * (row k of A).
// One warp per column, the entries of row k spread over the lanes (distinct i per lane: deterministic, no atomics);
// the column is accumulated in a per-warp shared buffer (carved from the panel region, idle at this point).
__device__ void form_H_sparse(const double *__restrict__ P, const SparseA &sp, const double *vals, const BlockSmem &s, int n, int np,
                              double sigma, double *Hw) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    double *col = s.panel + (size_t)warp * np;
    for (int j = warp; j < np; j += BNW) {
        for (int i = lane; i < np; i += 32) col[i] = 0.0;
        __syncwarp();
        if (j < n) {
            for (int pc = sp.col_outer[j]; pc < sp.col_outer[j + 1]; ++pc) {
                const int k = sp.col_inner[pc];
                const double f = s.rho[k] * vals[sp.col_perm ? sp.col_perm[pc] : pc];
                for (int pr = sp.row_outer[k] + lane; pr < sp.row_outer[k + 1]; pr += 32) {
                    const int i = sp.row_inner[pr];
                    if (i >= j) col[i] = fma(f, vals[sp.row_perm ? sp.row_perm[pr] : pr], col[i]);
                }
                __syncwarp();
            }
        }
        for (int i = j + lane; i < np; i += 32) {
            double v;
            if (i < n && j < n) v = col[i] + P[i + (size_t)n * j] + (i == j ? sigma : 0.0);
            else v = (i == j) ? 1.0 : 0.0;
            Hw[i + (size_t)np * j] = v;
        }
        __syncwarp();
    }
}

// ---- H = P_lowsym + sigma I + A^T diag(rho) A, lower triangle, into Hw (np x np, column-major) --------------------
__device__ void form_H(const double *__restrict__ P, const double *__restrict__ A, const BlockSmem &s, int n, int np, int m,
                       double sigma, double *Hw) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int fr = lane >> 2, fk = lane & 3;
    const int nt = np / 8;                 // 8x8 tiles per side
    const int ntiles = nt * (nt + 1) / 2;  // lower tiles ib >= jb
    const int ksteps = (m + 3) / 4;
    // each warp takes a row block ib and sweeps its jb <= ib tiles four at a time, so one A-fragment feeds four DMMAs
    for (int ib = warp; ib < nt; ib += BNW) {
        const int ci = 8 * ib + fr;
        for (int jb0 = 0; jb0 <= ib; jb0 += 4) {
            double c0[4] = {0, 0, 0, 0}, c1[4] = {0, 0, 0, 0};
            for (int ks = 0; ks < ksteps; ++ks) {
                const int k = 4 * ks + fk;
                const bool kv = k < m;
                const double af = (kv && ci < n) ? A[k + (size_t)m * ci] * s.rho[k] : 0.0;
#pragma unroll
                for (int t = 0; t < 4; ++t) {
                    const int cj = 8 * (jb0 + t) + fr;
                    const double bf = (kv && jb0 + t <= ib && cj < n) ? A[k + (size_t)m * cj] : 0.0;
                    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0, %1}, {%2}, {%3}, {%0, %1};"
                                 : "+d"(c0[t]), "+d"(c1[t])
                                 : "d"(af), "d"(bf));
                }
            }
#pragma unroll
            for (int t = 0; t < 4; ++t) {
                if (jb0 + t > ib) continue;
                const int i = 8 * ib + fr;
#pragma unroll
                for (int e = 0; e < 2; ++e) {
                    const int j = 8 * (jb0 + t) + 2 * fk + e;
                    if (i < j) continue;
                    double v = e ? c1[t] : c0[t];
                    if (i < n && j < n) v += P[i + (size_t)n * j] + (i == j ? sigma : 0.0);  // LDLT<Lower>: lower triangle of P only
                    else v = (i == j) ? 1.0 : 0.0;                                               // padded variables: unit block
                    Hw[i + (size_t)np * j] = v;
                }
            }
        }
    }
    (void)ntiles;
}

// ---- blocked L D L^T in place + inverted diagonal blocks. Returns false (uniformly) on a zero/NaN pivot ---------
__device__ bool factor_block(const BlockSmem &s, int np, double *Hw, double *Wd, int *s_fail) {
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int nb = np / NB;
    double *panel = s.panel, *Wsm = s.wsm, *dinv = s.wsm + NB * NB;
    if (tid == 0) *s_fail = 0;
    __syncthreads();
    for (int kb = 0; kb < nb; ++kb) {
        const int r0 = kb * NB, rows = np - r0;
        // a. stage the panel: rows r0.. of columns r0..r0+31 (lower part of the trailing matrix)
        for (int e = tid; e < rows * NB; e += BT) {
            const int r = e % rows, c = e / rows;
            panel[r + np * c] = (r >= c) ? Hw[(r0 + r) + (size_t)np * (r0 + c)] : 0.0;
        }
        __syncthreads();
        // b. 32x32 diagonal block: unblocked L D L^T by warp 0, lane i owns row i
        if (warp == 0) {
            bool bad = false;
            for (int j = 0; j < NB; ++j) {
                const double dj = panel[j + np * j];
                if (!(fabs(dj) > 0.0)) {
                    bad = true;  // same value in every lane: uniform
                    break;
                }
                double lij = 0.0;
                if (lane > j) {
                    lij = panel[lane + np * j] / dj;
                    panel[lane + np * j] = lij;
                }
                __syncwarp();
                if (lane > j) {
                    const double t = lij * dj;
                    for (int k = j + 1; k <= lane; ++k) panel[lane + np * k] -= t * panel[k + np * j];
                }
                __syncwarp();
            }
            if (bad) {
                if (lane == 0) *s_fail = 1;
            } else {
                // c. W = L11^-1 (unit lower), lane c computes column c by forward substitution; 1/d
                double wcol[NB];
#pragma unroll
                for (int i = 0; i < NB; ++i) wcol[i] = 0.0;
#pragma unroll
                for (int i = 0; i < NB; ++i) {
                    double v = (i == lane) ? 1.0 : 0.0;
                    if (i > lane) {
                        double acc = 0.0;
#pragma unroll
                        for (int k = 0; k < NB; ++k)
                            if (k < i) acc += panel[i + np * k] * wcol[k];
                        v = -acc;
                    }
                    wcol[i] = v;
                }
#pragma unroll
                for (int i = 0; i < NB; ++i) Wsm[i + NB * lane] = wcol[i];
                dinv[lane] = 1.0 / panel[lane + np * lane];
            }
        }
        __syncthreads();
        if (*s_fail) return false;
        // record W and 1/d for the substitutions; write the factored diagonal block back
        double *wrec = Wd + (size_t)kb * WBLK;
        for (int e = tid; e < WBLK; e += BT) wrec[e] = Wsm[e];
        // d. L21 = A21 L11^-T D^-1 = (A21 W^T) .* dinv, one thread per row below the block
        for (int r = NB + tid; r < rows; r += BT) {
            double a[NB];
#pragma unroll
            for (int c = 0; c < NB; ++c) a[c] = panel[r + np * c];
#pragma unroll
            for (int c = 0; c < NB; ++c) {
                double acc = 0.0;
#pragma unroll
                for (int k = 0; k < NB; ++k)
                    if (k <= c) acc += a[k] * Wsm[c + NB * k];
                const double lrc = acc * dinv[c];
                panel[r + np * c] = lrc;
                Hw[(r0 + r) + (size_t)np * (r0 + c)] = lrc;
            }
        }
        __syncthreads();
        // f. trailing update  H22[i][j] -= sum_c L[i][c] d_c L[j][c]  (i >= j), one warp per column j, lanes over i
        for (int j = NB + warp; j < rows; j += BNW) {
            double tj[NB];
#pragma unroll
            for (int c = 0; c < NB; ++c) tj[c] = panel[j + np * c] / dinv[c];
            for (int i = j + lane; i < rows; i += 32) {
                double acc = 0.0;
#pragma unroll
                for (int c = 0; c < NB; ++c) acc += panel[i + np * c] * tj[c];
                Hw[(r0 + i) + (size_t)np * (r0 + j)] -= acc;
            }
        }
        __syncthreads();
    }
    return true;
}

// ---- the inverted diagonal blocks W_k and 1/d of the factor in the slab -> shared memory (the panel region is idle
// outside the factorisation): the substitutions below read them every iteration ----------------------------------------
__device__ void stage_W(const BlockSmem &s, int np, const double *Wd) {
    const int nb = np / NB;
    for (int e = threadIdx.x; e < nb * NB * NB; e += BT) s.panel[e] = Wd[(size_t)(e / (NB * NB)) * WBLK + (e % (NB * NB))];
    for (int i = threadIdx.x; i < np; i += BT) s.dinv[i] = Wd[(size_t)(i / NB) * WBLK + NB * NB + (i % NB)];
    __syncthreads();
}

// ---- xt = H^-1 b by blocked substitution; b (in s.b) is destroyed. W_k, 1/d from shared memory (stage_W), the
// off-diagonal panels of L streamed from the slab (L2) with every load of a step in flight at once -----------------------
__device__ void solve_block(const BlockSmem &s, int np, const double *Hw) {
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int nb = np / NB;
    // forward: t = D^-1 L^-1 b
    for (int kb = 0; kb < nb; ++kb) {
        const int r0 = kb * NB;
        const double *W = s.panel + (size_t)kb * NB * NB;
        if (tid < NB) {
            double a0 = 0.0, a1 = 0.0, a2 = 0.0, a3 = 0.0;
            int c = 0;
            for (; c + 3 <= tid; c += 4) {
                a0 = fma(W[tid + NB * c], s.b[r0 + c], a0);
                a1 = fma(W[tid + NB * (c + 1)], s.b[r0 + c + 1], a1);
                a2 = fma(W[tid + NB * (c + 2)], s.b[r0 + c + 2], a2);
                a3 = fma(W[tid + NB * (c + 3)], s.b[r0 + c + 3], a3);
            }
            for (; c <= tid; ++c) a0 = fma(W[tid + NB * c], s.b[r0 + c], a0);
            s.t[r0 + tid] = (a0 + a1) + (a2 + a3);
        }
        __syncthreads();
        const int r = r0 + NB + tid;  // np - r0 - NB <= 224 < BT rows remain: one per thread
        if (r < np) {
            const double *col = Hw + r + (size_t)np * r0;
            double v[NB];
#pragma unroll
            for (int c = 0; c < NB; ++c) v[c] = col[(size_t)np * c];
            double a0 = 0.0, a1 = 0.0, a2 = 0.0, a3 = 0.0;
#pragma unroll
            for (int c = 0; c < NB; c += 4) {
                a0 = fma(v[c], s.t[r0 + c], a0);
                a1 = fma(v[c + 1], s.t[r0 + c + 1], a1);
                a2 = fma(v[c + 2], s.t[r0 + c + 2], a2);
                a3 = fma(v[c + 3], s.t[r0 + c + 3], a3);
            }
            s.b[r] -= (a0 + a1) + (a2 + a3);
        }
        __syncthreads();
    }
    for (int i = tid; i < np; i += BT) s.t[i] *= s.dinv[i];
    __syncthreads();
    // backward: xt = L^-T t
    for (int kb = nb - 1; kb >= 0; --kb) {
        const int r0 = kb * NB;
        const double *W = s.panel + (size_t)kb * NB * NB;
        const int chunks = (np - r0 - NB) / 32;  // <= 7 chunks of 32 rows below the block
        if (chunks > 0) {
            // warp w owns columns r0 + w + 8 q (q = 0..3) of the panel; all 4 x chunks loads are issued before the first use
            double acc[4] = {0.0, 0.0, 0.0, 0.0};
            const double *base = Hw + (size_t)np * (r0 + warp);
#pragma unroll
            for (int ch = 0; ch < 7; ++ch) {
                const bool live = ch < chunks;
                const int r = live ? r0 + NB + 32 * ch + lane : r0 + NB + lane;
                const double xr = live ? s.xt[r] : 0.0;
#pragma unroll
                for (int q = 0; q < 4; ++q) acc[q] = fma(base[r + (size_t)np * 8 * q], xr, acc[q]);
            }
#pragma unroll
            for (int o = 16; o > 0; o >>= 1)
#pragma unroll
                for (int q = 0; q < 4; ++q) acc[q] += __shfl_xor_sync(0xffffffffu, acc[q], o);
            if (lane < 4) {
                const double a = lane == 0 ? acc[0] : (lane == 1 ? acc[1] : (lane == 2 ? acc[2] : acc[3]));
                s.b[r0 + warp + 8 * lane] = s.t[r0 + warp + 8 * lane] - a;  // b is free: reuse as the block right-hand side
            }
        } else if (tid < NB) {
            s.b[r0 + tid] = s.t[r0 + tid];
        }
        __syncthreads();
        if (tid < NB) {
            double a0 = 0.0, a1 = 0.0, a2 = 0.0, a3 = 0.0;
            int c = tid;
            for (; c + 3 < NB; c += 4) {
                a0 = fma(W[c + NB * tid], s.b[r0 + c], a0);
                a1 = fma(W[c + 1 + NB * tid], s.b[r0 + c + 1], a1);
                a2 = fma(W[c + 2 + NB * tid], s.b[r0 + c + 2], a2);
                a3 = fma(W[c + 3 + NB * tid], s.b[r0 + c + 3], a3);
            }
            for (; c < NB; ++c) a0 = fma(W[c + NB * tid], s.b[r0 + c], a0);
            s.xt[r0 + tid] = (a0 + a1) + (a2 + a3);
        }
        __syncthreads();
    }
}

__global__ void __launch_bounds__(BT, 2) qp_block_kernel(KernelParams p) {
    extern __shared__ __align__(16) double smem_raw[];
    __shared__ int s_qp, s_fail;
    __shared__ double s_red[7][BNW];
    const int n = p.n, m = p.m, np = block_np(p.n);
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const bool sparse = p.sp.row_outer != nullptr;
    const int nnz = sparse ? p.sp.nnz : 0;
    BlockSmem s = carve_block(smem_raw, np, m, nnz);
    const SparseA sp = p.sp;
    const sqpb200_qp_settings st = p.s;
    const size_t fstride = block_fact_doubles_hd(n);

    for (;;) {
        __syncthreads();
        if (tid == 0) s_qp = draw_qp(p);
        __syncthreads();
        const int local = s_qp;
        if (local >= p.count) break;
        const size_t b = (size_t)p.first + local;
        const double *P = p.P + b * n * n, *A = sparse ? nullptr : p.A + b * m * n;
        if (sparse) {
            const double *gv = sp.vals + b * (size_t)nnz;
            for (int e = tid; e < nnz; e += BT) s.vals[e] = gv[e];
        }
        const double *vals = s.vals;
        const double *q = p.q + b * n, *l = p.l + b * m, *u = p.u + b * m;
        double *Hw = p.fact + b * fstride, *Wd = Hw + (size_t)np * np;

        int status = p.status[b];
        int rho_updates = p.rho_updates[b];
        double rho_est = p.rho_estimate[b], res_prim = p.res_prim[b], res_dual = p.res_dual[b];
        double rho = p.rho[b];
        int iter_out = p.iter[b];
        bool same_classes = true;

        for (int i = tid; i < np; i += BT) {
            s.q[i] = i < n ? q[i] : 0.0;
            s.x[i] = (i < n && !(p.mode & MODE_RESET)) ? p.x[b * n + i] : 0.0;
            s.xt[i] = 0.0;
        }
        for (int i = tid; i < m; i += BT) {
            s.l[i] = l[i];
            s.u[i] = u[i];
            s.z[i] = (p.mode & MODE_RESET) ? 0.0 : p.z[b * m + i];
            s.y[i] = (p.mode & MODE_RESET) ? 0.0 : p.y[b * m + i];
        }
        if (p.mode & MODE_FACTOR) {
            rho = st.rho;
            rho_updates += 1;  // rho_vec_update, qp.cpp:313
            for (int i = tid; i < m; i += BT) {
                const int t = classify(l[i], u[i]);
                s.type[i] = (signed char)t;
                if ((p.mode & MODE_REUSE) && p.ctype[b * m + i] != (signed char)t) same_classes = false;
                p.ctype[b * m + i] = (signed char)t;
            }
        } else {
            for (int i = tid; i < m; i += BT) s.type[i] = p.ctype[b * m + i];
        }
        __syncthreads();
        for (int i = tid; i < m; i += BT) {
            const double r = rho_of(s.type[i], rho);
            s.rho[i] = r;
        }
        __syncthreads();
        if (p.mode & MODE_FACTOR) {
            const bool reuse = (p.mode & MODE_REUSE) && __syncthreads_and(same_classes && p.fact_rho[b] == st.rho);
            if (reuse) {
                status = SQPB200_UNSOLVED;
            } else {
                if (sparse) form_H_sparse(P, sp, vals, s, n, np, st.sigma, Hw);
                else form_H(P, A, s, n, np, m, st.sigma, Hw);
                __syncthreads();
                const bool ok = factor_block(s, np, Hw, Wd, &s_fail);
                status = ok ? SQPB200_UNSOLVED : SQPB200_NUMERICAL_ISSUES;  // qp.cpp:39-43
                if (tid == 0) p.fact_rho[b] = ok ? rho : nan("");
            }
        }

        long long executed = 0;
        if ((p.mode & MODE_SOLVE) && status != SQPB200_UNINITIALIZED && status != SQPB200_NUMERICAL_ISSUES) {
            const double alpha = st.alpha, sigma = st.sigma;
            stage_W(s, np, Wd);
            int iter;
            for (iter = 1; iter <= st.max_iter; ++iter) {
                for (int i = tid; i < m; i += BT) s.w[i] = s.rho[i] * s.z[i] - s.y[i];
                __syncthreads();
                // b = sigma x - q + A^T w   (padded entries stay 0)
                if (sparse) {
                    for (int j = tid; j < np; j += BT) s.b[j] = (j < n) ? sigma * s.x[j] - s.q[j] + sp_coldot(sp, vals, j, s.w) : 0.0;
                } else
                for (int j = warp; j < np; j += BNW) {
                    double acc = 0.0;
                    if (j < n) {
                        const double *cj = A + (size_t)j * m;
                        double a1 = 0.0;
                        int i = lane;
#pragma unroll 4
                        for (; i + 32 < m; i += 64) {
                            acc = fma(cj[i], s.w[i], acc);
                            a1 = fma(cj[i + 32], s.w[i + 32], a1);
                        }
                        for (; i < m; i += 32) acc = fma(cj[i], s.w[i], acc);
                        acc = warp_sum(acc + a1);
                    }
                    if (lane == 0) s.b[j] = (j < n) ? sigma * s.x[j] - s.q[j] + acc : 0.0;
                }
                __syncthreads();
                solve_block(s, np, Hw);  // x~ in s.xt (qp.cpp:90)
                for (int i = tid; i < n; i += BT) s.x[i] = alpha * s.xt[i] + (1.0 - alpha) * s.x[i];
                // z~ = A x~ ; z, y updates (qp.cpp:93-103)
                for (int i = tid; i < m; i += BT) {
                    double acc;
                    if (sparse) {
                        acc = sp_rowdot(sp, vals, i, s.xt);
                    } else {
                        double acc0 = 0.0, acc1 = 0.0, acc2 = 0.0, acc3 = 0.0;  // independent chains: the loads are L2/HBM latency bound
                        int j = 0;
#pragma unroll 2
                        for (; j + 3 < n; j += 4) {
                            acc0 = fma(A[i + (size_t)m * j], s.xt[j], acc0);
                            acc1 = fma(A[i + (size_t)m * (j + 1)], s.xt[j + 1], acc1);
                            acc2 = fma(A[i + (size_t)m * (j + 2)], s.xt[j + 2], acc2);
                            acc3 = fma(A[i + (size_t)m * (j + 3)], s.xt[j + 3], acc3);
                        }
                        for (; j < n; ++j) acc0 = fma(A[i + (size_t)m * j], s.xt[j], acc0);
                        acc = (acc0 + acc1) + (acc2 + acc3);
                    }
                    const double zh = alpha * acc + (1.0 - alpha) * s.z[i];
                    const double zn = box_project(zh + (1.0 / s.rho[i]) * s.y[i], s.l[i], s.u[i]);
                    s.y[i] = s.y[i] + s.rho[i] * (zh - zn);
                    s.z[i] = zn;
                }
                __syncthreads();

                const bool chk = st.check_termination != 0 && iter % st.check_termination == 0;
                const bool adapt = st.adaptive_rho && st.adaptive_rho_interval > 0 && iter % st.adaptive_rho_interval == 0;
                if (chk || adapt) {
                    double mx[7] = {0, 0, 0, 0, 0, 0, 0};  // |Ax| |z| |Px| |A^T y| |q| |Ax - z| |Px + q + A^T y|
                    for (int i = tid; i < m; i += BT) {
                        double ax = 0.0;
                        if (sparse) ax = sp_rowdot(sp, vals, i, s.x);
                        else for (int j = 0; j < n; ++j) ax += A[i + (size_t)m * j] * s.x[j];
                        mx[0] = absmax(mx[0], ax);
                        mx[1] = absmax(mx[1], s.z[i]);
                        mx[5] = absmax(mx[5], ax - s.z[i]);
                    }
                    for (int j = warp; j < n; j += BNW) {
                        double aty = 0.0, px = 0.0;
                        if (sparse) {
                            if (lane == 0) aty = sp_coldot(sp, vals, j, s.y);
                        } else {
                            const double *cj = A + (size_t)j * m;
                            for (int i = lane; i < m; i += 32) aty += cj[i] * s.y[i];
                        }
                        for (int k = lane; k < n; k += 32) px += P[j + (size_t)n * k] * s.x[k];
                        aty = warp_sum(aty);
                        px = warp_sum(px);
                        mx[2] = absmax(mx[2], px);
                        mx[3] = absmax(mx[3], aty);
                        mx[4] = absmax(mx[4], s.q[j]);
                        mx[6] = absmax(mx[6], px + s.q[j] + aty);
                    }
#pragma unroll
                    for (int k = 0; k < 7; ++k) {
                        const double v = warp_max(mx[k]);
                        if (lane == 0) s_red[k][warp] = v;
                    }
                    __syncthreads();
#pragma unroll
                    for (int k = 0; k < 7; ++k) {
                        double v = s_red[k][0];
                        for (int w2 = 1; w2 < BNW; ++w2) v = s_red[k][w2] > v ? s_red[k][w2] : v;
                        mx[k] = v;
                    }
                    __syncthreads();
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
                            for (int i = tid; i < m; i += BT) {
                                const double r = rho_of(s.type[i], rho);
                                s.rho[i] = r;
                            }
                            __syncthreads();
                            if (sparse) form_H_sparse(P, sp, vals, s, n, np, sigma, Hw);
                            else form_H(P, A, s, n, np, m, sigma, Hw);
                            __syncthreads();
                            const bool ok2 = factor_block(s, np, Hw, Wd, &s_fail);
                            if (tid == 0) p.fact_rho[b] = ok2 ? rho : nan("");
                            if (!ok2) {
                                status = SQPB200_NUMERICAL_ISSUES;
                                break;
                            }
                            stage_W(s, np, Wd);
                        }
                    }
                }
            }
            executed = iter <= st.max_iter ? iter : st.max_iter;
            if (iter > st.max_iter) status = SQPB200_MAX_ITER_EXCEEDED;  // qp.cpp:147-149
            iter_out = iter;                                            // qp.cpp:150
        }

        for (int i = tid; i < n; i += BT) p.x[b * n + i] = s.x[i];
        for (int i = tid; i < m; i += BT) {
            p.z[b * m + i] = s.z[i];
            p.y[b * m + i] = s.y[i];
        }
        if (tid == 0) {
            p.status[b] = status;
            p.iter[b] = iter_out;
            p.rho_updates[b] = rho_updates;
            p.rho_estimate[b] = rho_est;
            p.res_prim[b] = res_prim;
            p.res_dual[b] = res_dual;
            p.rho[b] = rho;
            if (executed) atomicAdd(p.total_iters, (unsigned long long)executed);
        }
    }
}

cudaError_t launch_block(const KernelParams &p, int sm_count, size_t smem_optin, cudaStream_t stream, char *name, size_t name_len) {
    const size_t smem = block_smem_bytes(p.n, p.m, p.sp.row_outer ? p.sp.nnz : 0);
    if (smem > smem_optin) return cudaErrorInvalidValue;
    cudaError_t e = cudaFuncSetAttribute(qp_block_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    int occ = 0;
    e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, qp_block_kernel, BT, smem);
    if (e != cudaSuccess) return e;
    if (occ < 1) return cudaErrorLaunchOutOfResources;
    long long grid = (long long)sm_count * occ;
    if (grid > p.count) grid = p.count;
    if (name) snprintf(name, name_len, "block<%d>%sx%d", NB, p.sp.row_outer ? "/sparse" : "", occ);
    qp_block_kernel<<<(int)grid, BT, smem, stream>>>(p);
    return cudaGetLastError();
}

}  // namespace sqpb200
