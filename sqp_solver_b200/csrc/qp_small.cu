// Thread-per-QP kernel for tiny QPs (n + m <= 16): the SQP regime (reference src/sqp.cpp:210-242 hands the QP solver
// n = 2..3 variables and m = 1..3 constraints per subproblem, thousands of instances per batch in BASELINE config 4).
//
// This kernel is the reference's formulation LITERALLY, not the Schur-complement one of the larger kernels: the full
// (n+m) x (n+m) KKT matrix [[P + sigma I, .], [A, -diag(1/rho)]] (qp.cpp:185-187, lower blocks only) is factored by the
// diagonally pivoted unblocked LDL^T of Eigen::LDLT<MatrixXd, Lower> (qp.hpp:129; qp.cpp:242, :253) and every ADMM
// iteration SOLVES with that factor (qp.cpp:90: transpositions, unit-lower forward substitution, pseudo-inverse of D, backward
// substitution, inverse transpositions) -- the same operations in the same order as the CPU path, compiled without FMA
// contraction (-fmad=false for this translation unit), so late SQP subproblems whose BFGS Hessians reach cond(P) ~ 1e14
// and adaptive-rho steps decided by rounding noise come out the same as on the CPU. (At these sizes the (n+m)^2 factor is
// a few hundred bytes; eliminating the (2,2) block first would save nothing and costs the digits that
// P + A^T diag(rho) A loses when rho |A|^2 >> |P|, see tools/emulate_schur.py.)
//
// Mapping: one THREAD per QP, 32 QPs per CTA. All per-QP state -- the factor, the transpositions, x, z, y, rho, bounds -- lives
// in shared memory laid out [entry][thread]: a lane's bank depends only on its thread index, so the data-dependent pivot
// indices (different in every lane) never cause a bank conflict. No barrier, no shuffle: lanes of a warp run
// independent recursions and diverge freely (iteration counts, pivot swaps).
//
// Reference functions covered: all of src/qp.cpp:11-371 (same list as qp_generic.cu).
#include <cfloat>
#include <cstdio>

#include "qp_common.cuh"

namespace sqpb200 {

constexpr int ST = 32;  // QPs (= threads) per CTA

template <typename S> struct SmallNum;
template <> struct SmallNum<double> {
    static __device__ __forceinline__ double eps() { return DBL_EPSILON; }  // DIV_BY_ZERO_REGUL, qp.hpp:141
    static __device__ __forceinline__ double minpos() { return DBL_MIN; }
};
template <> struct SmallNum<float> {
    static __device__ __forceinline__ float eps() { return FLT_EPSILON; }
    static __device__ __forceinline__ float minpos() { return FLT_MIN; }
};

template <int NMAX>
constexpr size_t small_smem_bytes(size_t scalar) {
    // M[NMAX*NMAX] + 10 vectors of NMAX (x, q, z, y, rho, rho_inv, l, u, sol, tmp) + transpositions + classes
    return (size_t)ST * ((size_t)(NMAX * NMAX + 10 * NMAX) * scalar + (size_t)NMAX * (sizeof(int) + 1));
}

template <int NMAX, typename S>
__global__ void __launch_bounds__(ST) qp_small_kernel(KernelParams p) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int tid = threadIdx.x;
    const int local = blockIdx.x * ST + tid;
    if (local >= p.count) return;
    if (p.ready != nullptr) {  // host-staged call: wait until the chunk holding this QP has landed (draw_qp's protocol)
        const long long t0 = clock64();
        while (*reinterpret_cast<const volatile int *>(p.ready) <= local) {
            __nanosleep(500);
            if (clock64() - t0 > (1LL << 34)) __trap();  // the staging copies never arrived: fail instead of hanging
        }
        __threadfence();
    }
    const int n = p.n, m = p.m, N = n + m;
    const size_t b = (size_t)p.first + local;
    const double *gP = p.P + b * n * n, *gA = p.A + b * m * n, *gq = p.q + b * n, *gl = p.l + b * m, *gu = p.u + b * m;

    S *base = reinterpret_cast<S *>(smem_raw);
#define AT(arr, i) arr[(i) * ST + tid]
#define M_(i, j) AT(sM, (i) + (j) * N)
    S *sM = base;
    S *sx = sM + NMAX * NMAX * ST, *sq = sx + NMAX * ST, *sz = sq + NMAX * ST, *sy = sz + NMAX * ST;
    S *srv = sy + NMAX * ST, *sri = srv + NMAX * ST, *sl = sri + NMAX * ST, *su = sl + NMAX * ST;
    S *ssol = su + NMAX * ST, *stmp = ssol + NMAX * ST;
    int *stransp = reinterpret_cast<int *>(stmp + NMAX * ST);
    signed char *styp = reinterpret_cast<signed char *>(stransp + NMAX * ST);

    const sqpb200_qp_settings st = p.s;
    const S sigma = (S)st.sigma, alpha = (S)st.alpha;
    int status = p.status[b];
    int rho_updates = p.rho_updates[b];
    S rho_est = (S)p.rho_estimate[b], res_prim = (S)p.res_prim[b], res_dual = (S)p.res_dual[b];
    S rho = (S)p.rho[b];
    int iter_out = p.iter[b];
    const bool reset = (p.mode & MODE_RESET) != 0;

    for (int i = 0; i < n; ++i) {
        AT(sq, i) = (S)gq[i];
        AT(sx, i) = reset ? S(0) : (S)p.x[b * n + i];  // qp.cpp:16-18 (setup is the only real cold start)
    }
    for (int i = 0; i < m; ++i) {
        AT(sl, i) = (S)gl[i];
        AT(su, i) = (S)gu[i];
        AT(sz, i) = reset ? S(0) : (S)p.z[b * m + i];
        AT(sy, i) = reset ? S(0) : (S)p.y[b * m + i];
    }

    // rho_vec_update, qp.cpp:296-314
    auto rho_vec_update = [&](S rho0) {
        for (int i = 0; i < m; ++i) AT(srv, i) = rho_of_t<S>(AT(styp, i), rho0);
        for (int i = 0; i < m; ++i) AT(sri, i) = S(1) / AT(srv, i);
        rho = rho0;
    };
    // construct_KKT_mat (qp.cpp:185-187) straight into the matrix LDLT::compute works on, then Eigen's
    // ldlt_inplace<Lower>::unblocked. Returns LDLT::info() == Success.
    auto factorize = [&]() -> bool {
        for (int j = 0; j < n; ++j) {
            for (int i = 0; i < n; ++i) M_(i, j) = (S)gP[i + (size_t)n * j] + (i == j ? sigma : S(0));
            for (int i = 0; i < m; ++i) M_(n + i, j) = (S)gA[i + (size_t)m * j];
        }
        for (int j = 0; j < m; ++j)
            for (int i = j; i < m; ++i) M_(n + i, n + j) = (i == j) ? S(-1.0) * AT(sri, i) : S(0);
        bool found_zero_pivot = false, ret = true;
        if (N <= 1) {
            for (int i = 0; i < N; ++i) AT(stransp, i) = i;
            return true;
        }
        for (int k = 0; k < N; ++k) {
            int big = k;
            S best = fabs(M_(k, k));
            for (int i = k + 1; i < N; ++i) {
                const S v = fabs(M_(i, i));
                if (v > best) {
                    best = v;
                    big = i;
                }
            }
            AT(stransp, k) = big;
            if (k != big) {
                const int s = N - big - 1;
                for (int j = 0; j < k; ++j) {
                    const S t = M_(k, j);
                    M_(k, j) = M_(big, j);
                    M_(big, j) = t;
                }
                for (int i = 0; i < s; ++i) {
                    const S t = M_(big + 1 + i, k);
                    M_(big + 1 + i, k) = M_(big + 1 + i, big);
                    M_(big + 1 + i, big) = t;
                }
                {
                    const S t = M_(k, k);
                    M_(k, k) = M_(big, big);
                    M_(big, big) = t;
                }
                for (int i = k + 1; i < big; ++i) {
                    const S t = M_(i, k);
                    M_(i, k) = M_(big, i);
                    M_(big, i) = t;
                }
            }
            const int rs = N - k - 1;
            if (k > 0) {
                for (int j = 0; j < k; ++j) AT(stmp, j) = M_(j, j) * M_(k, j);
                S acc = 0;
                for (int j = 0; j < k; ++j) acc += M_(k, j) * AT(stmp, j);
                M_(k, k) -= acc;
                for (int j = 0; j < k; ++j) {
                    const S tj = AT(stmp, j);
                    for (int i = 0; i < rs; ++i) M_(k + 1 + i, k) -= M_(k + 1 + i, j) * tj;
                }
            }
            const S akk = M_(k, k);
            const bool pivot_is_valid = fabs(akk) > S(0);
            if (k == 0 && !pivot_is_valid) {
                for (int j = 0; j < N; ++j) {
                    AT(stransp, j) = j;
                    for (int i = j + 1; i < N; ++i) ret = ret && (M_(i, j) == S(0));
                }
                return ret;
            }
            if (rs > 0 && pivot_is_valid) {
                for (int i = 0; i < rs; ++i) M_(k + 1 + i, k) /= akk;
            } else if (rs > 0) {
                for (int i = 0; i < rs; ++i) ret = ret && (M_(k + 1 + i, k) == S(0));
            }
            if (found_zero_pivot && pivot_is_valid) ret = false;
            else if (!pivot_is_valid) found_zero_pivot = true;
        }
        return ret;
    };
    // LDLT::_solve_impl in place on ssol: P^T L^-T D^+ L^-1 P rhs
    auto kkt_solve = [&]() {
        for (int k = 0; k < N; ++k) {
            const int j = AT(stransp, k);
            if (j != k) {
                const S t = AT(ssol, k);
                AT(ssol, k) = AT(ssol, j);
                AT(ssol, j) = t;
            }
        }
        for (int k = 0; k < N; ++k) {
            const S v = AT(ssol, k);
            if (v != S(0))
                for (int i = k + 1; i < N; ++i) AT(ssol, i) -= M_(i, k) * v;
        }
        for (int i = 0; i < N; ++i) {
            const S d = M_(i, i);
            if (fabs(d) > SmallNum<S>::minpos()) AT(ssol, i) /= d;
            else AT(ssol, i) = S(0);
        }
        for (int i = N - 1; i >= 0; --i) {
            S acc = AT(ssol, i);
            for (int r = i + 1; r < N; ++r) acc -= M_(r, i) * AT(ssol, r);
            AT(ssol, i) = acc;
        }
        for (int k = N - 1; k >= 0; --k) {
            const int j = AT(stransp, k);
            if (j != k) {
                const S t = AT(ssol, k);
                AT(ssol, k) = AT(ssol, j);
                AT(ssol, j) = t;
            }
        }
    };

    if (p.mode & MODE_FACTOR) {  // setup / update_qp: qp.cpp:31-43, :48-61
        for (int i = 0; i < m; ++i) {
            const int t = classify_t<S>(AT(sl, i), AT(su, i));
            AT(styp, i) = (signed char)t;
            p.ctype[b * m + i] = (signed char)t;
        }
        rho_vec_update((S)st.rho);
        rho_updates += 1;
        status = factorize() ? SQPB200_UNSOLVED : SQPB200_NUMERICAL_ISSUES;
        p.fact_rho[b] = __longlong_as_double(0x7ff8000000000000LL);  // this kernel keeps no factor in the slab
    } else {
        // solve() after a separate setup()/update_qp() launch: the factor is a few hundred bytes, so it is rebuilt from the
        // stored classes and rho (bit-identical to the one setup computed) instead of travelling through HBM
        for (int i = 0; i < m; ++i) AT(styp, i) = p.ctype[b * m + i];
        rho_vec_update(rho);
        if (status != SQPB200_UNINITIALIZED && status != SQPB200_NUMERICAL_ISSUES && (p.mode & MODE_SOLVE)) factorize();
    }

    long long executed = 0;
    if ((p.mode & MODE_SOLVE) && status != SQPB200_UNINITIALIZED && status != SQPB200_NUMERICAL_ISSUES) {  // qp.cpp:68-71
        int iter;
        for (iter = 1; iter <= st.max_iter; ++iter) {
            // form_KKT_rhs, qp.cpp:272-276
            for (int i = 0; i < n; ++i) AT(ssol, i) = sigma * AT(sx, i) - AT(sq, i);
            for (int i = 0; i < m; ++i) AT(ssol, n + i) = AT(sz, i) - AT(sri, i) * AT(sy, i);
            kkt_solve();  // qp.cpp:90
            for (int i = 0; i < n; ++i) AT(sx, i) = alpha * AT(ssol, i) + (S(1) - alpha) * AT(sx, i);  // qp.cpp:96
            for (int i = 0; i < m; ++i) {
                const S zp = AT(sz, i), yi = AT(sy, i), ri = AT(sri, i);
                const S zt = zp + ri * (AT(ssol, n + i) - yi);                          // qp.cpp:93
                const S zh = alpha * zt + (S(1) - alpha) * zp;
                const S zn = box_project(zh + ri * yi, AT(sl, i), AT(su, i));            // qp.cpp:99-100
                AT(sy, i) = yi + AT(srv, i) * (zh - zn);                                // qp.cpp:103
                AT(sz, i) = zn;
            }
            const bool chk = st.check_termination != 0 && iter % st.check_termination == 0;
            const bool adapt = st.adaptive_rho && st.adaptive_rho_interval > 0 && iter % st.adaptive_rho_interval == 0;
            if (chk || adapt) {
                // update_state, qp.cpp:316-331; residual_prim/dual :353-361 (column-wise accumulation like a col-major gemv)
                S nAx = 0, nz = 0, nPx = 0, nATy = 0, nq = 0, rp = 0, rd = 0;
                for (int i = 0; i < m; ++i) {
                    S ax = 0;
                    for (int j = 0; j < n; ++j) ax += (S)gA[i + (size_t)m * j] * AT(sx, j);
                    nAx = absmax(nAx, ax);
                    nz = absmax(nz, AT(sz, i));
                    rp = absmax(rp, ax - AT(sz, i));
                }
                for (int j = 0; j < n; ++j) {
                    S px = 0, aty = 0;
                    for (int k = 0; k < n; ++k) px += (S)gP[j + (size_t)n * k] * AT(sx, k);
                    for (int i = 0; i < m; ++i) aty += (S)gA[i + (size_t)m * j] * AT(sy, i);
                    nPx = absmax(nPx, px);
                    nATy = absmax(nATy, aty);
                    nq = absmax(nq, AT(sq, j));
                    rd = absmax(rd, px + AT(sq, j) + aty);
                }
                const S sc_p = fmax(nAx, nz), sc_d = fmax(nPx, fmax(nATy, nq));
                res_prim = rp;
                res_dual = rd;
                if (chk && rp <= (S)st.eps_abs + (S)st.eps_rel * sc_p && rd <= (S)st.eps_abs + (S)st.eps_rel * sc_d) {
                    status = SQPB200_SOLVED;  // qp.cpp:119-122
                    break;
                }
                if (adapt) {  // qp.cpp:125-144
                    const S new_rho = rho_estimate_clamped_t<S>(rho, rp, rd, sc_p, sc_d);
                    rho_est = new_rho;
                    if (new_rho < rho / (S)st.adaptive_rho_tolerance || new_rho > rho * (S)st.adaptive_rho_tolerance) {
                        rho_vec_update(new_rho);
                        rho_updates += 1;
                        if (!factorize()) {
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

    for (int i = 0; i < n; ++i) p.x[b * n + i] = (double)AT(sx, i);
    for (int i = 0; i < m; ++i) {
        p.z[b * m + i] = (double)AT(sz, i);
        p.y[b * m + i] = (double)AT(sy, i);
    }
    p.status[b] = status;
    p.iter[b] = iter_out;
    p.rho_updates[b] = rho_updates;
    p.rho_estimate[b] = (double)rho_est;
    p.res_prim[b] = (double)res_prim;
    p.res_dual[b] = (double)res_dual;
    p.rho[b] = (double)rho;
    if (executed) atomicAdd(p.total_iters, (unsigned long long)executed);
#undef AT
#undef M_
}


// ---- register-resident variant for the sizes the SQP loop actually produces (n <= 4, n + m <= 8) ----------------------------------
// Same arithmetic in the same order (bit-identical results), but the ADMM state x, z, y, the right-hand side and the FACTOR's lower
// triangle live in registers: every loop of the iteration is unrolled over compile-time bounds (N_ variables exactly, up to
// NMAX - N_ constraints guarded by `i < m`), the row transpositions are a packed 4-bit-per-entry word applied with predicated swaps,
// so an iteration is a register-only dependency chain (~2 (n+m) multiply-subtract pairs) instead of ~30 dependent shared-memory round
// trips. The (rare) factorisation keeps its data-dependent pivot indices in shared memory, as above, and hands the factor over.
template <int N_, int NMAX, typename S>
__global__ void __launch_bounds__(ST) qp_small_reg_kernel(KernelParams p) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    constexpr int MMAX = NMAX - N_;
    constexpr int NTRI = NMAX * (NMAX + 1) / 2;
    const int tid = threadIdx.x;
    const int local = blockIdx.x * ST + tid;
    if (local >= p.count) return;
    if (p.ready != nullptr) {
        const long long t0 = clock64();
        while (*reinterpret_cast<const volatile int *>(p.ready) <= local) {
            __nanosleep(500);
            if (clock64() - t0 > (1LL << 34)) __trap();
        }
        __threadfence();
    }
    constexpr int n = N_;
    const int m = p.m, N = n + m;
    const size_t b = (size_t)p.first + local;
    const double *gP = p.P + b * n * n, *gA = p.A + b * m * n, *gq = p.q + b * n, *gl = p.l + b * m, *gu = p.u + b * m;

    S *base = reinterpret_cast<S *>(smem_raw);
#define AT(arr, i) arr[(i) * ST + tid]
#define M_(i, j) AT(sM, (i) + (j) * N)
    S *sM = base;
    S *stmp = sM + NMAX * NMAX * ST;
    int *stransp = reinterpret_cast<int *>(stmp + NMAX * ST);

    const sqpb200_qp_settings st = p.s;
    const S sigma = (S)st.sigma, alpha = (S)st.alpha;
    int status = p.status[b];
    int rho_updates = p.rho_updates[b];
    S rho_est = (S)p.rho_estimate[b], res_prim = (S)p.res_prim[b], res_dual = (S)p.res_dual[b];
    S rho = (S)p.rho[b];
    int iter_out = p.iter[b];
    const bool reset = (p.mode & MODE_RESET) != 0;

    S x[N_], q[N_], z[MMAX > 0 ? MMAX : 1], y[MMAX > 0 ? MMAX : 1], lo[MMAX > 0 ? MMAX : 1], up[MMAX > 0 ? MMAX : 1];
    S rv[MMAX > 0 ? MMAX : 1], ri[MMAX > 0 ? MMAX : 1];
    int typ[MMAX > 0 ? MMAX : 1];
    S sol[NMAX], Lr[NTRI];  // Lr: lower triangle of the factor incl. D on the diagonal, column by column
    unsigned tp = 0;        // transpositions, 4 bits each
#pragma unroll
    for (int i = 0; i < N_; ++i) {
        q[i] = (S)gq[i];
        x[i] = reset ? S(0) : (S)p.x[b * n + i];
    }
#pragma unroll
    for (int i = 0; i < MMAX; ++i) {
        const bool real = i < m;
        lo[i] = real ? (S)gl[i] : S(0);
        up[i] = real ? (S)gu[i] : S(0);
        z[i] = (real && !reset) ? (S)p.z[b * m + i] : S(0);
        y[i] = (real && !reset) ? (S)p.y[b * m + i] : S(0);
        rv[i] = ri[i] = S(1);
        typ[i] = SQPB200_LOOSE_BOUNDS;
    }
    auto rho_vec_update = [&](S rho0) {  // qp.cpp:296-314
#pragma unroll
        for (int i = 0; i < MMAX; ++i)
            if (i < m) rv[i] = rho_of_t<S>(typ[i], rho0);
#pragma unroll
        for (int i = 0; i < MMAX; ++i)
            if (i < m) ri[i] = S(1) / rv[i];
        rho = rho0;
    };
    // identical to the shared-memory kernel's factorisation (data-dependent pivot indices); afterwards the factor moves to registers
    auto factorize = [&]() -> bool {
#pragma unroll
        for (int j = 0; j < N_; ++j) {
#pragma unroll
            for (int i = 0; i < N_; ++i) M_(i, j) = (S)gP[i + (size_t)n * j] + (i == j ? sigma : S(0));
            for (int i = 0; i < m; ++i) M_(n + i, j) = (S)gA[i + (size_t)m * j];
        }
#pragma unroll
        for (int j = 0; j < MMAX; ++j)
            if (j < m)
                for (int i = j; i < m; ++i) M_(n + i, n + j) = (i == j) ? S(-1.0) * ri[j] : S(0);
        bool found_zero_pivot = false, ret = true, done = false;
        if (N <= 1) {
            for (int i = 0; i < N; ++i) AT(stransp, i) = i;
            done = true;
        }
        for (int k = 0; k < N && !done; ++k) {
            int big = k;
            S best = fabs(M_(k, k));
            for (int i = k + 1; i < N; ++i) {
                const S v = fabs(M_(i, i));
                if (v > best) {
                    best = v;
                    big = i;
                }
            }
            AT(stransp, k) = big;
            if (k != big) {
                const int s = N - big - 1;
                for (int j = 0; j < k; ++j) {
                    const S t = M_(k, j);
                    M_(k, j) = M_(big, j);
                    M_(big, j) = t;
                }
                for (int i = 0; i < s; ++i) {
                    const S t = M_(big + 1 + i, k);
                    M_(big + 1 + i, k) = M_(big + 1 + i, big);
                    M_(big + 1 + i, big) = t;
                }
                {
                    const S t = M_(k, k);
                    M_(k, k) = M_(big, big);
                    M_(big, big) = t;
                }
                for (int i = k + 1; i < big; ++i) {
                    const S t = M_(i, k);
                    M_(i, k) = M_(big, i);
                    M_(big, i) = t;
                }
            }
            const int rs = N - k - 1;
            if (k > 0) {
                for (int j = 0; j < k; ++j) AT(stmp, j) = M_(j, j) * M_(k, j);
                S acc = 0;
                for (int j = 0; j < k; ++j) acc += M_(k, j) * AT(stmp, j);
                M_(k, k) -= acc;
                for (int j = 0; j < k; ++j) {
                    const S tj = AT(stmp, j);
                    for (int i = 0; i < rs; ++i) M_(k + 1 + i, k) -= M_(k + 1 + i, j) * tj;
                }
            }
            const S akk = M_(k, k);
            const bool pivot_is_valid = fabs(akk) > S(0);
            if (k == 0 && !pivot_is_valid) {
                for (int j = 0; j < N; ++j) {
                    AT(stransp, j) = j;
                    for (int i = j + 1; i < N; ++i) ret = ret && (M_(i, j) == S(0));
                }
                done = true;
                break;
            }
            if (rs > 0 && pivot_is_valid) {
                for (int i = 0; i < rs; ++i) M_(k + 1 + i, k) /= akk;
            } else if (rs > 0) {
                for (int i = 0; i < rs; ++i) ret = ret && (M_(k + 1 + i, k) == S(0));
            }
            if (found_zero_pivot && pivot_is_valid) ret = false;
            else if (!pivot_is_valid) found_zero_pivot = true;
        }
        // hand the factor over: lower triangle and transpositions into registers (static indices only)
        tp = 0;
#pragma unroll
        for (int j = 0; j < NMAX; ++j) {
            tp |= (unsigned)((j < N) ? AT(stransp, j) : j) << (4 * j);
#pragma unroll
            for (int i = j; i < NMAX; ++i) Lr[j * NMAX - j * (j - 1) / 2 + (i - j)] = (i < N) ? M_(i, j) : S(0);
        }
        return ret;
    };
#define LR(i, j) Lr[(j) * NMAX - (j) * ((j) - 1) / 2 + ((i) - (j))]
    // LDLT::_solve_impl on the register copy: P^T L^-T D^+ L^-1 P rhs, entries beyond N are skipped by predicates
    auto kkt_solve = [&]() {
#pragma unroll
        for (int k = 0; k < NMAX; ++k) {
            const int j = (int)((tp >> (4 * k)) & 15u);
#pragma unroll
            for (int jj = k + 1; jj < NMAX; ++jj)
                if (jj == j) {
                    const S t = sol[k];
                    sol[k] = sol[jj];
                    sol[jj] = t;
                }
        }
#pragma unroll
        for (int k = 0; k < NMAX; ++k) {
            const S v = sol[k];
            const bool nz = v != S(0);
#pragma unroll
            for (int i = k + 1; i < NMAX; ++i)
                if (i < N && nz) sol[i] -= LR(i, k) * v;
        }
#pragma unroll
        for (int i = 0; i < NMAX; ++i)
            if (i < N) {
                const S d = LR(i, i);
                if (fabs(d) > SmallNum<S>::minpos()) sol[i] /= d;
                else sol[i] = S(0);
            }
#pragma unroll
        for (int i = NMAX - 1; i >= 0; --i) {
            S acc = sol[i];
#pragma unroll
            for (int r = i + 1; r < NMAX; ++r)
                if (r < N) acc -= LR(r, i) * sol[r];
            sol[i] = acc;
        }
#pragma unroll
        for (int k = NMAX - 1; k >= 0; --k) {
            const int j = (int)((tp >> (4 * k)) & 15u);
#pragma unroll
            for (int jj = k + 1; jj < NMAX; ++jj)
                if (jj == j) {
                    const S t = sol[k];
                    sol[k] = sol[jj];
                    sol[jj] = t;
                }
        }
    };

    if (p.mode & MODE_FACTOR) {
#pragma unroll
        for (int i = 0; i < MMAX; ++i)
            if (i < m) {
                typ[i] = classify_t<S>(lo[i], up[i]);
                p.ctype[b * m + i] = (signed char)typ[i];
            }
        rho_vec_update((S)st.rho);
        rho_updates += 1;
        status = factorize() ? SQPB200_UNSOLVED : SQPB200_NUMERICAL_ISSUES;
        p.fact_rho[b] = __longlong_as_double(0x7ff8000000000000LL);
    } else {
#pragma unroll
        for (int i = 0; i < MMAX; ++i)
            if (i < m) typ[i] = p.ctype[b * m + i];
        rho_vec_update(rho);
        if (status != SQPB200_UNINITIALIZED && status != SQPB200_NUMERICAL_ISSUES && (p.mode & MODE_SOLVE)) factorize();
    }

    long long executed = 0;
    if ((p.mode & MODE_SOLVE) && status != SQPB200_UNINITIALIZED && status != SQPB200_NUMERICAL_ISSUES) {
        int iter;
        for (iter = 1; iter <= st.max_iter; ++iter) {
#pragma unroll
            for (int i = 0; i < NMAX; ++i) sol[i] = S(0);
#pragma unroll
            for (int i = 0; i < N_; ++i) sol[i] = sigma * x[i] - q[i];
#pragma unroll
            for (int i = 0; i < MMAX; ++i)
                if (i < m) sol[N_ + i] = z[i] - ri[i] * y[i];
            kkt_solve();
#pragma unroll
            for (int i = 0; i < N_; ++i) x[i] = alpha * sol[i] + (S(1) - alpha) * x[i];
#pragma unroll
            for (int i = 0; i < MMAX; ++i)
                if (i < m) {
                    const S zp = z[i], yi = y[i];
                    const S zt = zp + ri[i] * (sol[N_ + i] - yi);
                    const S zh = alpha * zt + (S(1) - alpha) * zp;
                    const S zn = box_project(zh + ri[i] * yi, lo[i], up[i]);
                    y[i] = yi + rv[i] * (zh - zn);
                    z[i] = zn;
                }
            const bool chk = st.check_termination != 0 && iter % st.check_termination == 0;
            const bool adapt = st.adaptive_rho && st.adaptive_rho_interval > 0 && iter % st.adaptive_rho_interval == 0;
            if (chk || adapt) {
                S nAx = 0, nz = 0, nPx = 0, nATy = 0, nq = 0, rp = 0, rd = 0;
#pragma unroll
                for (int i = 0; i < MMAX; ++i)
                    if (i < m) {
                        S ax = 0;
#pragma unroll
                        for (int j = 0; j < N_; ++j) ax += (S)gA[i + (size_t)m * j] * x[j];
                        nAx = absmax(nAx, ax);
                        nz = absmax(nz, z[i]);
                        rp = absmax(rp, ax - z[i]);
                    }
#pragma unroll
                for (int j = 0; j < N_; ++j) {
                    S px = 0, aty = 0;
#pragma unroll
                    for (int k = 0; k < N_; ++k) px += (S)gP[j + (size_t)n * k] * x[k];
#pragma unroll
                    for (int i = 0; i < MMAX; ++i)
                        if (i < m) aty += (S)gA[i + (size_t)m * j] * y[i];
                    nPx = absmax(nPx, px);
                    nATy = absmax(nATy, aty);
                    nq = absmax(nq, q[j]);
                    rd = absmax(rd, px + q[j] + aty);
                }
                const S sc_p = fmax(nAx, nz), sc_d = fmax(nPx, fmax(nATy, nq));
                res_prim = rp;
                res_dual = rd;
                if (chk && rp <= (S)st.eps_abs + (S)st.eps_rel * sc_p && rd <= (S)st.eps_abs + (S)st.eps_rel * sc_d) {
                    status = SQPB200_SOLVED;
                    break;
                }
                if (adapt) {
                    const S new_rho = rho_estimate_clamped_t<S>(rho, rp, rd, sc_p, sc_d);
                    rho_est = new_rho;
                    if (new_rho < rho / (S)st.adaptive_rho_tolerance || new_rho > rho * (S)st.adaptive_rho_tolerance) {
                        rho_vec_update(new_rho);
                        rho_updates += 1;
                        if (!factorize()) {
                            status = SQPB200_NUMERICAL_ISSUES;
                            break;
                        }
                    }
                }
            }
        }
        executed = iter <= st.max_iter ? iter : st.max_iter;
        if (iter > st.max_iter) status = SQPB200_MAX_ITER_EXCEEDED;
        iter_out = iter;
    }

#pragma unroll
    for (int i = 0; i < N_; ++i) p.x[b * n + i] = (double)x[i];
#pragma unroll
    for (int i = 0; i < MMAX; ++i)
        if (i < m) {
            p.z[b * m + i] = (double)z[i];
            p.y[b * m + i] = (double)y[i];
        }
    p.status[b] = status;
    p.iter[b] = iter_out;
    p.rho_updates[b] = rho_updates;
    p.rho_estimate[b] = (double)rho_est;
    p.res_prim[b] = (double)res_prim;
    p.res_dual[b] = (double)res_dual;
    p.rho[b] = (double)rho;
    if (executed) atomicAdd(p.total_iters, (unsigned long long)executed);
#undef AT
#undef M_
#undef LR
}

bool small_supported(int n, int m) { return n >= 1 && m >= 0 && n + m <= 16; }

template <int NMAX, typename S>
static cudaError_t launch_small_cfg(const KernelParams &p, cudaStream_t stream, char *name, size_t name_len) {
    auto kernel = qp_small_kernel<NMAX, S>;
    const size_t smem = small_smem_bytes<NMAX>(sizeof(S));
    cudaError_t e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    if (name) snprintf(name, name_len, "small<%d%s>", NMAX, sizeof(S) == 8 ? "" : ",f32");
    kernel<<<(p.count + ST - 1) / ST, ST, smem, stream>>>(p);
    return cudaGetLastError();
}

template <int N_, int NMAX, typename S>
static cudaError_t launch_small_reg_cfg(const KernelParams &p, cudaStream_t stream, char *name, size_t name_len) {
    auto kernel = qp_small_reg_kernel<N_, NMAX, S>;
    const size_t smem = (size_t)ST * ((size_t)(NMAX * NMAX + NMAX) * sizeof(S) + (size_t)NMAX * sizeof(int));
    cudaError_t e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    if (name) snprintf(name, name_len, "small<%d%s>/reg n=%d", NMAX, sizeof(S) == 8 ? "" : ",f32", N_);
    kernel<<<(p.count + ST - 1) / ST, ST, smem, stream>>>(p);
    return cudaGetLastError();
}
template <int N_, typename S>
static cudaError_t launch_small_reg_n(const KernelParams &p, cudaStream_t stream, char *name, size_t name_len) {
    const int N = N_ + p.m;
    if constexpr (N_ <= 4) {
        if (N <= 4) return launch_small_reg_cfg<N_, 4, S>(p, stream, name, name_len);
    }
    if (N <= 6) return launch_small_reg_cfg<N_, 6, S>(p, stream, name, name_len);
    return launch_small_reg_cfg<N_, 8, S>(p, stream, name, name_len);
}
template <typename S>
static cudaError_t launch_small_reg(const KernelParams &p, cudaStream_t stream, char *name, size_t name_len) {
    switch (p.n) {
        case 1: return launch_small_reg_n<1, S>(p, stream, name, name_len);
        case 2: return launch_small_reg_n<2, S>(p, stream, name, name_len);
        case 3: return launch_small_reg_n<3, S>(p, stream, name, name_len);
        default: return launch_small_reg_n<4, S>(p, stream, name, name_len);
    }
}

cudaError_t launch_small(const KernelParams &p, int f32, cudaStream_t stream, char *name, size_t name_len) {
    const int N = p.n + p.m;
    if (p.n <= 4 && N <= 8) return f32 ? launch_small_reg<float>(p, stream, name, name_len) : launch_small_reg<double>(p, stream, name, name_len);
    if (f32) return N <= 8 ? launch_small_cfg<8, float>(p, stream, name, name_len) : launch_small_cfg<16, float>(p, stream, name, name_len);
    return N <= 8 ? launch_small_cfg<8, double>(p, stream, name, name_len) : launch_small_cfg<16, double>(p, stream, name, name_len);
}

}  // namespace sqpb200
