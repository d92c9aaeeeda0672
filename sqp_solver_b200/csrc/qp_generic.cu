// Generic batched ADMM QP kernel: any (n, m), one CTA per QP, persistent CTAs pulling QP
// indices from an atomic work queue.  Vectors live in shared memory; A and P are read
// through L1/L2 from the caller's arrays; H^-1 lives in a global n*n slab per QP.
// This is the shape-agnostic fallback (and the first correct path); the register-tiled
// kernel in qp_tile.cu is the fast path for n <= 64, m <= 128.
//
// Reference functions covered (all src/qp.cpp): setup :11-44, update_qp :46-62, solve :64-157,
// form_KKT_rhs :272-276, box_projection :278-281, constr_type_init :283-294,
// rho_vec_update :296-314, update_state :316-331, rho_estimate :333-341,
// eps_prim/eps_dual :343-351, residual_prim/dual :353-361, termination_criteria :363-371.
#include <cstdio>

#include "qp_common.cuh"

namespace sqpb200 {

constexpr int GT = 256;  // threads per CTA
constexpr int GNW = GT / 32;

// S is the compute scalar: double, or float for QPSolver<float> (qp.cpp:386) at the shapes the register-tiled kernel does not cover.
// The interface arrays stay double (converted at the loads and stores); the float factor is packed into the double slab.
template <typename S>
struct GenericSmem {
    S *x, *xt, *b, *q, *d;             // n each
    S *z, *y, *w, *l, *u, *rho, *rhoinv;  // m each
    int *type;                               // m
};
__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

template <typename S>
__device__ __forceinline__ GenericSmem<S> carve(S *base, int n, int m) {
    GenericSmem<S> s;
    s.x = base;
    s.xt = s.x + n;
    s.b = s.xt + n;
    s.q = s.b + n;
    s.d = s.q + n;
    s.z = s.d + n;
    s.y = s.z + m;
    s.w = s.y + m;
    s.l = s.w + m;
    s.u = s.l + m;
    s.rho = s.u + m;
    s.rhoinv = s.rho + m;
    s.type = reinterpret_cast<int *>(s.rhoinv + m);
    return s;
}
static size_t generic_smem_bytes(int n, int m) { return sizeof(double) * (5 * (size_t)n + 7 * (size_t)m) + sizeof(int) * (size_t)m + 16; }  // sized for double

// H^-1 = (P_lowsym + sigma I + A^T diag(rho) A)^-1 into H (n x n column-major, both triangles).
// W is an n x n scratch slab.  Returns false (uniformly over the CTA) on a zero or NaN pivot.
template <typename S>
__device__ bool factor_generic(const double *__restrict__ P, const double *__restrict__ A, const GenericSmem<S> &s, int n,
                               int m, S sigma, S *H, S *W, int *s_fail) {
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    // 1. lower triangle of H, one warp per entry, lanes over the constraint index (coalesced columns of A)
    for (int e = warp; e < n * n; e += GNW) {
        int i = e % n, j = e / n;
        if (i < j) continue;
        const double *ci = A + (size_t)i * m, *cj = A + (size_t)j * m;
        S acc = 0;
        for (int k = lane; k < m; k += 32) acc += s.rho[k] * (S)ci[k] * (S)cj[k];
        acc = warp_sum(acc);
        if (lane == 0) H[i + (size_t)n * j] = (S)P[i + (size_t)n * j] + (i == j ? sigma : S(0.0)) + acc;
    }
    if (tid == 0) *s_fail = 0;
    __syncthreads();
    // 2. right-looking LDL^T in place: column k of L below the diagonal, D on the diagonal
    for (int k = 0; k < n; ++k) {
        S dk = H[k + (size_t)n * k];
        if (!(fabs(dk) > S(0.0))) {  // zero or NaN pivot
            if (tid == 0) *s_fail = 1;
            break;  // dk is the same value for every thread: uniform exit
        }
        for (int i = k + 1 + tid; i < n; i += GT) H[i + (size_t)n * k] /= dk;
        __syncthreads();
        for (int j = k + 1 + warp; j < n; j += GNW) {
            S t = dk * H[j + (size_t)n * k];
            for (int i = j + lane; i < n; i += 32) H[i + (size_t)n * j] -= H[i + (size_t)n * k] * t;
        }
        __syncthreads();
    }
    __syncthreads();
    if (*s_fail) return false;
    // 3. W = L^-1 (unit lower), one warp per column, column-sweep forward substitution
    for (int c = warp; c < n; c += GNW) {
        S *wc = W + (size_t)n * c;
        for (int i = lane; i < n; i += 32) wc[i] = (i == c) ? S(1.0) : S(0.0);
        __syncwarp();
        for (int k = c; k < n - 1; ++k) {
            S wk = wc[k];
            const S *lk = H + (size_t)n * k;
            for (int i = k + 1 + lane; i < n; i += 32) wc[i] -= lk[i] * wk;
            __syncwarp();
        }
    }
    for (int k = tid; k < n; k += GT) s.d[k] = H[k + (size_t)n * k];
    __syncthreads();
    // 4. H^-1[i][j] = sum_{k >= i} W[k][i] W[k][j] / d_k  for i >= j, mirrored
    for (int e = warp; e < n * n; e += GNW) {
        int i = e % n, j = e / n;
        if (i < j) continue;
        const S *wi = W + (size_t)n * i, *wj = W + (size_t)n * j;
        S acc = 0;
        for (int k = i + lane; k < n; k += 32) acc += wi[k] * wj[k] / s.d[k];
        acc = warp_sum(acc);
        if (lane == 0) {
            H[i + (size_t)n * j] = acc;
            H[j + (size_t)n * i] = acc;
        }
    }
    __syncthreads();
    return true;
}

template <typename S>
__global__ void __launch_bounds__(GT) qp_generic_kernel(KernelParams p) {
    extern __shared__ double smem_raw[];
    __shared__ int s_qp, s_fail;
    __shared__ S s_red[7][GNW];
    const int n = p.n, m = p.m;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    GenericSmem<S> s = carve<S>(reinterpret_cast<S *>(smem_raw), n, m);
    const sqpb200_qp_settings st = p.s;
    S *W = reinterpret_cast<S *>(p.scratch + (size_t)blockIdx.x * n * n);

    for (;;) {
        __syncthreads();
        if (tid == 0) s_qp = draw_qp(p);
        __syncthreads();
        const int local = s_qp;
        if (local >= p.count) break;
        const size_t b = (size_t)p.first + local;
        const double *P = p.P + b * n * n, *A = p.A + b * m * n;
        const double *q = p.q + b * n, *l = p.l + b * m, *u = p.u + b * m;
        S *H = reinterpret_cast<S *>(p.fact + b * n * n);

        int status = p.status[b];
        int rho_updates = p.rho_updates[b];
        S rho_est = (S)p.rho_estimate[b], res_prim = (S)p.res_prim[b], res_dual = (S)p.res_dual[b];
        S rho = (S)p.rho[b];
        int iter_out = p.iter[b];
        bool same_classes = true;

        for (int i = tid; i < n; i += GT) {
            s.q[i] = (S)q[i];
            s.x[i] = (p.mode & MODE_RESET) ? S(0.0) : (S)p.x[b * n + i];
        }
        for (int i = tid; i < m; i += GT) {
            s.l[i] = (S)l[i];
            s.u[i] = (S)u[i];
            s.z[i] = (p.mode & MODE_RESET) ? S(0.0) : (S)p.z[b * m + i];
            s.y[i] = (p.mode & MODE_RESET) ? S(0.0) : (S)p.y[b * m + i];
        }
        if (p.mode & MODE_FACTOR) {
            rho = (S)st.rho;
            rho_updates += 1;  // rho_vec_update, qp.cpp:313
            for (int i = tid; i < m; i += GT) {
                int t = classify_t<S>((S)l[i], (S)u[i]);
                s.type[i] = t;
                if ((p.mode & MODE_REUSE) && p.ctype[b * m + i] != (signed char)t) same_classes = false;
                p.ctype[b * m + i] = (signed char)t;
            }
        } else {
            for (int i = tid; i < m; i += GT) s.type[i] = p.ctype[b * m + i];
        }
        __syncthreads();
        for (int i = tid; i < m; i += GT) {
            S r = rho_of_t<S>(s.type[i], rho);
            s.rho[i] = r;
            s.rhoinv[i] = S(1.0) / r;
        }
        __syncthreads();
        if (p.mode & MODE_FACTOR) {
            // MODE_REUSE: same P, A as the launch whose factor sits in the slab; skip when classes and rho are unchanged
            const bool reuse = (p.mode & MODE_REUSE) && __syncthreads_and(same_classes && (S)p.fact_rho[b] == (S)st.rho);
            if (reuse) {
                status = SQPB200_UNSOLVED;
            } else {
                bool ok = factor_generic<S>(P, A, s, n, m, (S)st.sigma, H, W, &s_fail);
                status = ok ? SQPB200_UNSOLVED : SQPB200_NUMERICAL_ISSUES;  // qp.cpp:39-43
                if (tid == 0) p.fact_rho[b] = ok ? st.rho : nan("");
            }
        }

        long long executed = 0;
        if ((p.mode & MODE_SOLVE) && status != SQPB200_UNINITIALIZED && status != SQPB200_NUMERICAL_ISSUES) {
            const S alpha = (S)st.alpha, sigma = (S)st.sigma;
            int iter;
            for (iter = 1; iter <= st.max_iter; ++iter) {
                // w = rho .* z - y   (tail of the KKT rhs, qp.cpp:275, times rho)
                for (int i = tid; i < m; i += GT) s.w[i] = s.rho[i] * s.z[i] - s.y[i];
                __syncthreads();
                // b = sigma x - q + A^T w
                for (int j = warp; j < n; j += GNW) {
                    const double *cj = A + (size_t)j * m;
                    S acc = 0;
                    for (int i = lane; i < m; i += 32) acc += (S)cj[i] * s.w[i];
                    acc = warp_sum(acc);
                    if (lane == 0) s.b[j] = sigma * s.x[j] - s.q[j] + acc;
                }
                __syncthreads();
                // x~ = H^-1 b ; x = alpha x~ + (1 - alpha) x   (qp.cpp:90-96)
                for (int i = warp; i < n; i += GNW) {
                    const S *hi = H + (size_t)n * i;  // row i == column i (symmetric)
                    S acc = 0;
                    for (int j = lane; j < n; j += 32) acc += hi[j] * s.b[j];
                    acc = warp_sum(acc);
                    if (lane == 0) {
                        s.xt[i] = acc;
                        s.x[i] = alpha * acc + (S(1.0) - alpha) * s.x[i];
                    }
                }
                __syncthreads();
                // z~ = A x~ ; z, y updates (qp.cpp:93-103)
                for (int i = tid; i < m; i += GT) {
                    S acc = 0;
                    for (int j = 0; j < n; ++j) acc += (S)A[i + (size_t)m * j] * s.xt[j];
                    S zh = alpha * acc + (S(1.0) - alpha) * s.z[i];
                    S zn = box_project(zh + s.rhoinv[i] * s.y[i], s.l[i], s.u[i]);
                    s.y[i] = s.y[i] + s.rho[i] * (zh - zn);
                    s.z[i] = zn;
                }
                __syncthreads();

                const bool chk = st.check_termination != 0 && iter % st.check_termination == 0;
                const bool adapt = st.adaptive_rho && st.adaptive_rho_interval > 0 && iter % st.adaptive_rho_interval == 0;
                if (chk || adapt) {
                    // update_state, qp.cpp:316-331
                    S mx[7] = {0, 0, 0, 0, 0, 0, 0};  // |Ax| |z| |Px| |A^T y| |q| |Ax - z| |Px + q + A^T y|
                    for (int i = tid; i < m; i += GT) {
                        S ax = 0;
                        for (int j = 0; j < n; ++j) ax += (S)A[i + (size_t)m * j] * s.x[j];
                        mx[0] = absmax(mx[0], ax);
                        mx[1] = absmax(mx[1], s.z[i]);
                        mx[5] = absmax(mx[5], ax - s.z[i]);
                    }
                    for (int j = warp; j < n; j += GNW) {
                        const double *cj = A + (size_t)j * m;
                        S aty = 0, px = 0;
                        for (int i = lane; i < m; i += 32) aty += (S)cj[i] * s.y[i];
                        for (int k = lane; k < n; k += 32) px += (S)P[j + (size_t)n * k] * s.x[k];
                        aty = warp_sum(aty);
                        px = warp_sum(px);
                        mx[2] = absmax(mx[2], px);
                        mx[3] = absmax(mx[3], aty);
                        mx[4] = absmax(mx[4], s.q[j]);
                        mx[6] = absmax(mx[6], px + s.q[j] + aty);
                    }
#pragma unroll
                    for (int k = 0; k < 7; ++k) {
                        S v = warp_max(mx[k]);
                        if (lane == 0) s_red[k][warp] = v;
                    }
                    __syncthreads();
#pragma unroll
                    for (int k = 0; k < 7; ++k) {
                        S v = s_red[k][0];
                        for (int w2 = 1; w2 < GNW; ++w2) v = s_red[k][w2] > v ? s_red[k][w2] : v;
                        mx[k] = v;
                    }
                    __syncthreads();
                    const S sc_p = fmax(mx[0], mx[1]);
                    const S sc_d = fmax(mx[2], fmax(mx[3], mx[4]));
                    res_prim = mx[5];
                    res_dual = mx[6];
                    if (chk && st.verbose && local == 0) {  // print_status, qp.cpp:114-118 and :375-382 (first instance of the launch only)
                        if (tid == 0) {
                            double obj = 0.0;
                            for (int j = 0; j < n; ++j) {
                                double pxj = 0.0;
                                for (int k = 0; k < n; ++k) pxj += P[j + (size_t)n * k] * (double)s.x[k];
                                obj += (double)s.x[j] * (0.5 * pxj + (double)s.q[j]);
                            }
                            if (iter == st.check_termination) printf("iter   obj       rp        rd\n");
                            printf("%4d  %.2e  %.2e  %.2e\n", iter, obj, (double)res_prim, (double)res_dual);
                        }
                    }
                    if (chk) {  // termination_criteria, qp.cpp:363-371
                        if (res_prim <= (S)st.eps_abs + (S)st.eps_rel * sc_p && res_dual <= (S)st.eps_abs + (S)st.eps_rel * sc_d) {
                            status = SQPB200_SOLVED;
                            break;
                        }
                    }
                    if (adapt) {  // qp.cpp:125-144
                        S new_rho = rho_estimate_clamped_t<S>(rho, res_prim, res_dual, sc_p, sc_d);
                        rho_est = new_rho;
                        if (new_rho < rho / (S)st.adaptive_rho_tolerance || new_rho > rho * (S)st.adaptive_rho_tolerance) {
                            rho = new_rho;
                            rho_updates += 1;
                            for (int i = tid; i < m; i += GT) {
                                S r = rho_of_t<S>(s.type[i], rho);
                                s.rho[i] = r;
                                s.rhoinv[i] = S(1.0) / r;
                            }
                            __syncthreads();
                            const bool ok2 = factor_generic<S>(P, A, s, n, m, sigma, H, W, &s_fail);
                            if (tid == 0) p.fact_rho[b] = ok2 ? (double)rho : nan("");  // the slab now holds the factor for the new rho
                            if (!ok2) {
                                status = SQPB200_NUMERICAL_ISSUES;
                                break;
                            }
                        }
                    }
                }
            }
            executed = iter <= st.max_iter ? iter : st.max_iter;
            if (iter > st.max_iter) status = SQPB200_MAX_ITER_EXCEEDED;  // qp.cpp:147-149
            iter_out = iter;                                            // qp.cpp:150
        }

        for (int i = tid; i < n; i += GT) p.x[b * n + i] = (double)s.x[i];
        for (int i = tid; i < m; i += GT) {
            p.z[b * m + i] = (double)s.z[i];
            p.y[b * m + i] = (double)s.y[i];
        }
        if (tid == 0) {
            p.status[b] = status;
            p.iter[b] = iter_out;
            p.rho_updates[b] = rho_updates;
            p.rho_estimate[b] = (double)rho_est;
            p.res_prim[b] = (double)res_prim;
            p.res_dual[b] = (double)res_dual;
            p.rho[b] = (double)rho;
            if (executed) atomicAdd(p.total_iters, (unsigned long long)executed);
        }
    }
}

bool generic_supported(int n, int m, size_t smem_optin) { return n >= 1 && m >= 0 && generic_smem_bytes(n, m) <= smem_optin; }
int generic_grid(int count, int sm_count) {
    int g = sm_count * 4;
    return count < g ? count : g;
}
size_t generic_scratch_bytes(int n, int grid) { return sizeof(double) * (size_t)n * n * grid; }

cudaError_t launch_generic(const KernelParams &p, int sm_count, size_t smem_optin, int f32, cudaStream_t stream, int *grid_out) {
    size_t smem = generic_smem_bytes(p.n, p.m);
    if (smem > smem_optin) return cudaErrorInvalidValue;
    auto kernel = f32 ? qp_generic_kernel<float> : qp_generic_kernel<double>;
    cudaError_t e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    int grid = generic_grid(p.count, sm_count);
    if (grid_out) *grid_out = grid;
    kernel<<<grid, GT, smem, stream>>>(p);
    return cudaGetLastError();
}

}  // namespace sqpb200
