// Register-tiled batched ADMM QP kernel for sm_100a: the fast path for n <= 64, m <= 128.
//
// One CTA (NW warps) per QP, persistent CTAs on an atomic work queue.  Per QP:
//   * A (m x n) lives in REGISTERS for the whole solve: warp w owns MP/NW consecutive rows, a lane
//     owns an R x C tile (R rows, C interleaved column pairs).  Both per-iteration mat-vecs,
//     g = A^T w (reduce over rows) and z~ = A x~ (reduce over columns), run from registers;
//     partial sums are combined with warp-shuffle reduce-scatter trees, so after A x~ each row's
//     z, y, l, u, rho live in the registers of the lane that owns the row. The trees are select-free:
//     the tiles sit in the registers in a lane-dependent permutation (HalveSF / sf_perm below), so a
//     tree step is v[k] += shfl_xor(v[k + V/2]) with no send/keep selects.
//   * H^-1 = (P_sym + sigma I + A^T diag(rho) A)^-1 lives in REGISTERS too (HR x HC entries per
//     lane); x~ = H^-1 b is one dense symmetric mat-vec with a short shuffle tree.  (The first
//     version kept H^-1 in shared memory; ncu showed the LSU/shared pipe, not fp64, was the bound --
//     profiles/r01a_*.json -- so every per-iteration matrix now sits in the register file.)
//     H is formed by a register-blocked SYRK from a shared-memory staging copy of A and is
//     factored in registers by symmetric pivot-by-pivot elimination (the pivots are exactly the D of
//     the LDL^T of H; a zero/NaN pivot reports NUMERICAL_ISSUES) that is carried through to the
//     inverse, so the ADMM loop has no serial substitution chain.
//   * P lives in shared memory for the residual checks (P*x uses all of P, qp.cpp:323).
// Reference functions covered: identical list to qp_generic.cu (all of src/qp.cpp:11-371).
#include <cstdint>
#include <cstdio>

#include "qp_tile.cuh"

namespace sqpb200 {

constexpr unsigned FULL = 0xffffffffu;

// S_ is the COMPUTE scalar: registers, shared memory and all arithmetic. The arrays in HBM are fp64 either way (the fp32
// instantiation converts at its loads and stores: every fp32 value is exactly representable, so x, z, y and the stored factor
// round-trip losslessly between launches). fp32 is the reference's QPSolver<float> (qp.cpp:386).
template <typename T> struct Vec2;
template <> struct Vec2<double> { using type = double2; };
template <> struct Vec2<float> { using type = float2; };
__device__ __forceinline__ double2 mk2(double a, double b) { return make_double2(a, b); }
__device__ __forceinline__ float2 mk2(float a, float b) { return make_float2(a, b); }

template <int NP_, int MP_, int NW_, int LC_, int HR_, int MINB_, typename S_ = double, bool SF_ = true, bool SEPP_ = false>
struct TileCfg {
    // SEPP: P keeps a shared-memory region of its own next to the staging copy of A (instead of replacing it after every
    // factorisation): A and P are staged together, a refactorisation reads P_lowsym from shared memory and re-stages A only
    static constexpr bool SEPP = SEPP_;
    using S = S_;
    // SF: select-free reduce-scatter trees -- the A and H^-1 tiles sit in the registers in a lane-dependent permutation (HalveSF below)
    static constexpr bool SF = SF_;
    static constexpr bool F64 = sizeof(S_) == 8;
    static constexpr int NP = NP_, MP = MP_, NW = NW_, LC = LC_, HR = HR_, MINB = MINB_;
    static constexpr int T = 32 * NW;
    static constexpr int LR = 32 / LC;   // lanes along the row direction inside a warp
    static constexpr int RW = MP / NW;   // rows per warp
    static constexpr int R = RW / LR;    // rows per lane
    static constexpr int C = NP / LC;    // columns per lane
    static constexpr int RO = (R >= LC) ? R / LC : 1;  // rows a lane owns after the A x~ reduce-scatter
    static constexpr int CO = (C >= LR) ? C / LR : 1;  // columns a lane holds after the A^T w reduce-scatter
    // symmetric n x n tiles (H^-1 in registers, P in smem): a lane holds HR consecutive rows x HC strided columns
    static constexpr int CG = T * HR / NP;              // lanes sharing a row group
    static constexpr int HC = NP / CG;                  // columns per lane
    static constexpr int HS = NP + 16 / CG;             // padded column stride of P in smem
    static constexpr int LS = MP + 2;                   // padded column stride of the A staging copy
    static_assert(MP % NW == 0 && RW % LR == 0 && R >= 1, "row tiling");
    static_assert(NP % LC == 0 && C >= 2 && C % 2 == 0, "column tiling");
    static_assert(CG >= 2 && CG <= 16 && (CG & (CG - 1)) == 0 && NP % CG == 0, "symmetric tile: column groups");
    static_assert(HR >= 2 && HR % 2 == 0 && HR <= CG && NP % HR == 0 && (NP / HR) * CG == T, "symmetric tile: rows");
    static_assert(R == 1 || R % 2 == 0, "row tile must be vectorisable");
    static_assert(T >= NP, "b stage needs one thread per variable");
    // A^T diag(rho) A on the fp64 tensor cores (mma.sync m8n8k4): warp w owns IB 8-row blocks of H x all JB 8-column blocks
    static constexpr bool DMMA = F64 && (NP % (8 * NW) == 0);
    static constexpr int IB = DMMA ? NP / (8 * NW) : 1, JB = NP / 8;
    // shared memory carve-up (doubles)
    // The padded staging copy of A (needed only while H is formed) and the padded copy of P (needed afterwards, for
    // P*x at the checks) share one region unless SEPP: a second 34 KB region measurably slows iteration-dominated launches
    // (round 1: 26.7 vs 24.9 ms on config 3; round 2: 23.04 vs 22.56 ms), while launches that refactor often win by it.
    static_assert(NP * HS <= NP * LS, "P must fit in the A staging region it aliases");
    static constexpr int OFF_STAGE = SEPP ? NP * HS : 0;
    static constexpr int OFF_P = 0;
    // Regions that are only live during a (re)factorisation share storage with regions that are only live during iterations:
    // the pivot columns with the per-warp partial sums, the rho vector (SYRK) with w
    static constexpr int PIVS = 2 * NP;  // two pivot columns per elimination step
    static constexpr int OFF_PART = OFF_STAGE + NP * LS;
    static constexpr int OFF_PIV = OFF_PART;
    static constexpr int PART_DOUBLES = (NW * NP > 2 * PIVS) ? NW * NP : 2 * PIVS;
    static constexpr int OFF_X = OFF_PART + PART_DOUBLES;
    static constexpr int OFF_XT = OFF_X + NP;
    static constexpr int OFF_B = OFF_XT + NP;
    static constexpr int OFF_Q = OFF_B + NP;
    static constexpr int OFF_PX = OFF_Q + NP;
    static constexpr int OFF_W = OFF_PX + NP;
    static constexpr int OFF_RHO = OFF_W;
    static_assert(CG % 2 == 0, "the rank-2 elimination pairs adjacent pivot columns");
    static constexpr int OFF_BND = OFF_W + MP;  // (l, u) pairs per row: read once per iteration by the row owner
    static constexpr int OFF_RED = OFF_BND + 2 * MP;
    static constexpr int SMEM_DOUBLES = OFF_RED + 8 * NW;
    static constexpr size_t SMEM_BYTES = sizeof(S) * SMEM_DOUBLES;
};

// ---- TMA bulk copy (cp.async.bulk, SASS UBLKCP) + mbarrier helpers ---------------------------------
__device__ __forceinline__ unsigned smem_u32(const void *p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(unsigned long long *bar, unsigned count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_expect_tx(unsigned long long *bar, unsigned bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(unsigned long long *bar, unsigned parity) {
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "WAIT_%=:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
        "@p bra DONE_%=;\n\t"
        "bra WAIT_%=;\n\t"
        "DONE_%=:\n\t"
        "}" ::"r"(smem_u32(bar)), "r"(parity) : "memory");
}
// order earlier generic-proxy accesses to shared memory before later async-proxy (TMA) writes
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
// global -> shared bulk copy of `bytes` (multiple of 16, both addresses 16-byte aligned), completion on `bar`
__device__ __forceinline__ void tma_bulk_g2s(void *dst, const void *src, unsigned bytes, unsigned long long *bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(dst)),
                 "l"(src), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}

template <int NW>
__device__ __forceinline__ bool cta_all(bool pred) {
    if constexpr (NW == 1) return __all_sync(0xffffffffu, pred);
    else return __syncthreads_and(pred) != 0;
}
template <int NW>
__device__ __forceinline__ void cta_sync() {
    if constexpr (NW == 1) __syncwarp();
    else __syncthreads();
}

// Reduce-scatter over the lanes whose ids differ in the bits HI..LO (powers of two): every step
// halves the number of live values per lane; once one value is left the remaining steps are a
// plain butterfly (the result is then duplicated over those lanes).
template <int V, int HI, int LO>
struct Halve {
    template <typename T>
    static __device__ __forceinline__ void run(T *v, int lane) {
        if constexpr (HI >= LO && HI >= 1) {
            if constexpr (V > 1) {
                const bool up = (lane & HI) != 0;
#pragma unroll
                for (int k = 0; k < V / 2; ++k) {
                    const T send = up ? v[k] : v[k + V / 2];
                    const T keep = up ? v[k + V / 2] : v[k];
                    v[k] = keep + __shfl_xor_sync(FULL, send, HI);
                }
                Halve<V / 2, HI / 2, LO>::run(v, lane);
            } else {
                v[0] += __shfl_xor_sync(FULL, v[0], HI);
                Halve<1, HI / 2, LO>::run(v, lane);
            }
        }
    }
};
// index (into the original V values) of the first value this lane ends up with, and whether this
// lane is the primary copy among duplicates
template <int V, int HI, int LO>
__device__ __forceinline__ int halve_base(int lane, bool &primary) {
    int base = 0, v = V;
    primary = true;
#pragma unroll
    for (int mask = HI; mask >= LO && mask >= 1; mask >>= 1) {
        if (v > 1) {
            v >>= 1;
            if (lane & mask) base += v;
        } else if (lane & mask) {
            primary = false;
        }
    }
    return base;
}
// Select-free reduce-scatter. A lane that keeps the logical value (slot ^ sf_perm(lane)) in physical slot `slot` holds, at every halving
// step, the values it must keep in the low half of its slots and the values its partner wants in the high half: v[k] += shfl(v[k + V/2])
// with no selects (the select form costs 4 FSEL per fp64 value and step, a sixth of the instructions of an ADMM iteration). The sums are
// formed exactly as in Halve (keep + received), and a lane ends up owning the same logical values. Steps whose halves are smaller than
// MINHALF values fall back to Halve (MINHALF = 2 keeps adjacent pairs adjacent, so 16-byte loads of the operands stay possible).
template <int V, int HI, int LO, int MINHALF>
struct HalveSF {
    template <typename T>
    static __device__ __forceinline__ void run(T *v, int lane) {
        if constexpr (HI >= LO && HI >= 1) {
            if constexpr (V > 1 && V / 2 >= MINHALF) {
#pragma unroll
                for (int k = 0; k < V / 2; ++k) v[k] = v[k] + __shfl_xor_sync(FULL, v[k + V / 2], HI);
                HalveSF<V / 2, HI / 2, LO, MINHALF>::run(v, lane);
            } else {
                Halve<V, HI, LO>::run(v, lane);
            }
        }
    }
};
// the permutation that goes with HalveSF<V, HI, LO, MINHALF>: logical index = physical slot ^ sf_perm(lane)
template <int V, int HI, int LO, int MINHALF>
__device__ __forceinline__ int sf_perm(int lane) {
    int perm = 0, v = V;
#pragma unroll
    for (int mask = HI; mask >= LO && mask >= 1; mask >>= 1) {
        if (v > 1 && v / 2 >= MINHALF) {
            v >>= 1;
            if (lane & mask) perm += v;
        } else {
            break;
        }
    }
    return perm;
}
template <class Cfg>
struct Tile {
    using S = typename Cfg::S;
    using V2 = typename Vec2<S>::type;
    static constexpr int NP = Cfg::NP, MP = Cfg::MP, NW = Cfg::NW, LC = Cfg::LC, LR = Cfg::LR, T = Cfg::T;
    static constexpr int R = Cfg::R, C = Cfg::C, RO = Cfg::RO, CO = Cfg::CO, CG = Cfg::CG, HC = Cfg::HC, HR = Cfg::HR;
    static constexpr int HS = Cfg::HS, LS = Cfg::LS, RW = Cfg::RW;

    // column (0..NP) of the kk-th entry of this lane's A tile: interleaved pairs so that the x~ loads
    // of one warp form contiguous 16-byte chunks
    static __device__ __forceinline__ int colA(int lc, int kk) { return 2 * lc + 2 * LC * (kk >> 1) + (kk & 1); }
    // Register permutations of the select-free trees (0 when Cfg::SF is off): slot (kr, kk) of the A tile holds row (kr ^ px) and
    // column slot (kk ^ py) of the lane's logical tile, row slot r of the H^-1 tile holds logical row (r ^ ph). px and py are even.
    static __device__ __forceinline__ int perm_rows(int lane) { return Cfg::SF ? sf_perm<R, LC / 2, 1, 2>(lane) : 0; }
    static __device__ __forceinline__ int perm_cols(int lane) { return Cfg::SF ? sf_perm<C, 16, LC, 2>(lane) : 0; }
    static __device__ __forceinline__ int perm_sym(int lane) { return Cfg::SF ? sf_perm<HR, CG / 2, 1, 1>(lane) : 0; }

    // ---- z~ = A v (v in shared memory); result: RO fully reduced rows per lane -----------------
    static __device__ __forceinline__ void mv_A(const S (&a)[R][C], const S *sv, int lc, int lane, S (&out)[RO]) {
        S acc[R];
#pragma unroll
        for (int kr = 0; kr < R; ++kr) acc[kr] = S(0);
#pragma unroll
        for (int t = 0; t < C / 2; ++t) {
            const V2 xv = *reinterpret_cast<const V2 *>(sv + 2 * lc + 2 * LC * (t ^ (perm_cols(lane) >> 1)));
#pragma unroll
            for (int kr = 0; kr < R; ++kr) {
                acc[kr] = fma(a[kr][2 * t], xv.x, acc[kr]);
                acc[kr] = fma(a[kr][2 * t + 1], xv.y, acc[kr]);
            }
        }
        if constexpr (Cfg::SF) HalveSF<R, LC / 2, 1, 2>::run(acc, lane);
        else Halve<R, LC / 2, 1>::run(acc, lane);
#pragma unroll
        for (int t = 0; t < RO; ++t) out[t] = acc[t];
    }

    // ---- partial g = A^T v over this warp's rows -> part[warp][*] -------------------------------
    // v for the lane's R rows is read from sw (written by the row owners just before)
    static __device__ __forceinline__ void mv_At(const S (&a)[R][C], const S *sw, S *part_w, int row0,
                                                 int lc, int lane) {
        // w is consumed in chunks of at most 4 rows so that only 4 of its values are live next to the C accumulators
        // and the R*C tile: the hot loop of the 64x128 configuration sits at the 255-register limit
        S acc[C];
#pragma unroll
        for (int kk = 0; kk < C; ++kk) acc[kk] = S(0);
        constexpr int WCH = (R >= 4) ? 4 : R;
#pragma unroll
        for (int h = 0; h < R; h += WCH) {
            S wv[WCH];
            if constexpr (WCH >= 2) {
#pragma unroll
                for (int kr = 0; kr < WCH; kr += 2) {
                    const V2 t2 = *reinterpret_cast<const V2 *>(sw + row0 + ((h + kr) ^ perm_rows(lane)));
                    wv[kr] = t2.x;
                    wv[kr + 1] = t2.y;
                }
            } else {
                wv[0] = sw[row0 + h];
            }
#pragma unroll
            for (int kk = 0; kk < C; ++kk)
#pragma unroll
                for (int kr = 0; kr < WCH; ++kr) acc[kk] = fma(a[h + kr][kk], wv[kr], acc[kk]);
        }
        if constexpr (Cfg::SF) HalveSF<C, 16, LC, 2>::run(acc, lane);
        else Halve<C, 16, LC>::run(acc, lane);
        bool primary;
        const int kb = halve_base<C, 16, LC>(lane, primary);
        if (primary) {
            if constexpr (CO >= 2) {
#pragma unroll
                for (int t = 0; t < CO; t += 2)
                    *reinterpret_cast<V2 *>(part_w + colA(lc, kb + t)) = mk2(acc[t], acc[t + 1]);
            } else {
                part_w[colA(lc, kb)] = acc[0];
            }
        }
    }

    // ---- y = M v for the symmetric-tile layout; each row ends fully reduced on CG/HR lanes --------
    // M in registers (H^-1)
    static __device__ __forceinline__ S mv_sym_reg(const S (&hv)[HR][HC], const S *sv, int rg, int cg, int lane,
                                                        int &row, bool &primary) {
        S acc[HR], acc2[HR];  // two chains per row: the mat-vec is latency bound, not issue bound
#pragma unroll
        for (int r = 0; r < HR; ++r) acc[r] = acc2[r] = S(0);
#pragma unroll
        for (int s = 0; s < HC; ++s) {
            const S bv = sv[cg + CG * s];
#pragma unroll
            for (int r = 0; r < HR; ++r) {
                if (s & 1) acc2[r] = fma(hv[r][s], bv, acc2[r]);
                else acc[r] = fma(hv[r][s], bv, acc[r]);
            }
        }
        if constexpr (HC > 1) {
#pragma unroll
            for (int r = 0; r < HR; ++r) acc[r] += acc2[r];
        }
        if constexpr (Cfg::SF) HalveSF<HR, CG / 2, 1, 1>::run(acc, lane);
        else Halve<HR, CG / 2, 1>::run(acc, lane);
        row = HR * rg + halve_base<HR, CG / 2, 1>(lane, primary);
        return acc[0];
    }
    // Bring a freshly built H^-1 tile (logical row r in slot r) into the slot order mv_sym_reg expects: slot r <- logical row r ^ ph
    static __device__ __forceinline__ void permute_sym(S (&hv)[HR][HC], int lane) {
        if constexpr (Cfg::SF) {
            const int ph = perm_sym(lane);
#pragma unroll
            for (int bit = 1; bit < HR; bit <<= 1) {
                const bool sw = (ph & bit) != 0;
#pragma unroll
                for (int r = 0; r < HR; ++r) {
                    if (r & bit) continue;
#pragma unroll
                    for (int s = 0; s < HC; ++s) {
                        const S lo = hv[r][s], hi = hv[r | bit][s];
                        hv[r][s] = sw ? hi : lo;
                        hv[r | bit][s] = sw ? lo : hi;
                    }
                }
            }
        }
    }
    // M in shared memory (P at the residual checks): padded column-major, column stride HS
    static __device__ __forceinline__ S mv_sym_smem(const S *sM, const S *sv, int rg, int cg, int lane, int &row,
                                                         bool &primary) {
        S acc[HR];
#pragma unroll
        for (int r = 0; r < HR; ++r) acc[r] = S(0);
#pragma unroll
        for (int s = 0; s < HC; ++s) {
            const int k = cg + CG * s;
            const S bv = sv[k];
#pragma unroll
            for (int r = 0; r < HR; r += 2) {
                const V2 mv = *reinterpret_cast<const V2 *>(sM + HR * rg + r + HS * k);
                acc[r] = fma(mv.x, bv, acc[r]);
                acc[r + 1] = fma(mv.y, bv, acc[r + 1]);
            }
        }
        Halve<HR, CG / 2, 1>::run(acc, lane);
        row = HR * rg + halve_base<HR, CG / 2, 1>(lane, primary);
        return acc[0];
    }
};

// SWEEP2 selects the elimination variant of the (re)factorisation: two pivots per barrier step (fewer, heavier steps: the
// better choice when adaptive rho refactors often) or one pivot per step (leaves the compiler the leaner hot loop: the better
// choice when the launch is iteration-dominated). Measured on config 3: S1 24.6 vs 25.1 ms, S2 5.58 vs 5.32 ms.
constexpr int RESUME_FLAG = 1 << 30;
// Next unit of work of a time-sliced launch (one thread per CTA): a fresh QP while there are any, then suspended QPs from the re-queue
// ring in the order they were published; -1 once every QP of the launch has finished.
__device__ __forceinline__ int draw_sliced(const KernelParams &p) {
    const int v = atomicAdd(p.work_counter, 1);
    if (v < p.count) {
        if (p.ready != nullptr) {
            const long long t0 = clock64();
            while (*reinterpret_cast<const volatile int *>(p.ready) <= v) {
                __nanosleep(500);
                if (clock64() - t0 > (1LL << 34)) __trap();
            }
            __threadfence();
        }
        return v;
    }
    const int e = v - p.count;
    if (e >= p.rq_cap) return -1;
    const long long t0 = clock64();
    for (;;) {
        const int q = *reinterpret_cast<const volatile int *>(p.rq + e);
        if (q >= 0) {
            __threadfence();
            return q | RESUME_FLAG;
        }
        if (*reinterpret_cast<const volatile int *>(p.done) >= p.count) return -1;
        __nanosleep(200);
        if (clock64() - t0 > (1LL << 35)) __trap();  // ~17 s without any progress: fail instead of hanging
    }
}

// SLICED: time-sliced launch (KernelParams::slice_iters > 0); the unsliced instantiation is the one the benchmarks of one GPU run and is
// untouched by the slicing logic (the kernel sits at the 255-register limit).
template <class Cfg, bool SWEEP2, bool SLICED = false>
__global__ void __launch_bounds__(Cfg::T, Cfg::MINB) qp_tile_kernel(KernelParams p) {
    using TL = Tile<Cfg>;
    using S = typename Cfg::S;
    using V2 = typename Vec2<S>::type;
    constexpr int NP = Cfg::NP, MP = Cfg::MP, NW = Cfg::NW, LC = Cfg::LC, T = Cfg::T;
    constexpr int R = Cfg::R, C = Cfg::C, RO = Cfg::RO, CG = Cfg::CG, HC = Cfg::HC, HR = Cfg::HR, HS = Cfg::HS, LS = Cfg::LS, RW = Cfg::RW;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    S *smem = reinterpret_cast<S *>(smem_raw);
    __shared__ int s_qp, s_fail;
    __shared__ __align__(8) unsigned long long s_mbar;  // completion barrier of the TMA bulk copies
    __shared__ S s_info[4];  // rho_estimate, res_prim, res_dual, rho (CTA-uniform scalars that only change at checks)
    __shared__ int s_cnt[1];      // rho_updates
    unsigned mbar_parity = 0;
    S *sA = smem + Cfg::OFF_STAGE, *sP = smem + Cfg::OFF_P;
    S *part = smem + Cfg::OFF_PART, *sx = smem + Cfg::OFF_X, *sxt = smem + Cfg::OFF_XT, *sb = smem + Cfg::OFF_B;
    S *sq = smem + Cfg::OFF_Q, *spx = smem + Cfg::OFF_PX, *sw = smem + Cfg::OFF_W, *srho = smem + Cfg::OFF_RHO;
    S *piv = smem + Cfg::OFF_PIV, *red = smem + Cfg::OFF_RED, *sbnd = smem + Cfg::OFF_BND;

    const int n = p.n, m = p.m;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int lc = lane % LC, lr = lane / LC;
    const int row0 = warp * RW + lr * R;  // first row of this lane's A tile
    bool row_primary;
    const int own0 = row0 + halve_base<R, LC / 2, 1>(lane, row_primary);  // first owned row after A x~
    const int rg = tid / CG, cg = tid % CG;                                // symmetric-tile coordinates
    const int i0 = HR * rg;
    const sqpb200_qp_settings st = p.s;
    const S sigma = st.sigma, alpha = st.alpha;
    const S INF = __longlong_as_double(0x7ff0000000000000LL);
    if (tid == 0) mbar_init(&s_mbar, 1);

    for (;;) {
        cta_sync<NW>();
        if (tid == 0) s_qp = SLICED ? draw_sliced(p) : draw_qp(p);
        cta_sync<NW>();
        // SLICED: a resumed QP continues like a solve() after a stored setup (state and H^-1 from the object's arrays), with the
        // iteration counter carried on -- the arithmetic sequence is the one of an unsliced solve, bit for bit
        const bool resumed = SLICED && s_qp >= 0 && (s_qp & RESUME_FLAG) != 0;
        const int local = SLICED ? (s_qp < 0 ? p.count : (s_qp & ~RESUME_FLAG)) : s_qp;
        if (local >= p.count) break;
        // Only the batch index stays live across the solve; problem pointers are re-derived from the kernel
        // parameters (constant bank) where needed -- the hot loop is register-limited.
        const int bi = p.first + local;
        const size_t b = (size_t)bi;
#define PMODE ((SLICED && resumed) ? (unsigned)(MODE_LOAD_FACTOR | MODE_SOLVE) : p.mode)
#define gP (((SLICED && resumed && p.loc_P) ? p.loc_P : p.P) + (size_t)bi * n * n)
#define gA (((SLICED && resumed && p.loc_A) ? p.loc_A : p.A) + (size_t)bi * m * n)
#define gq (((SLICED && resumed && p.loc_A) ? p.loc_q : p.q) + (size_t)bi * n)
#define gl (((SLICED && resumed && p.loc_A) ? p.loc_l : p.l) + (size_t)bi * m)
#define gu (((SLICED && resumed && p.loc_A) ? p.loc_u : p.u) + (size_t)bi * m)
        // where the iterates and the info come from: the caller-visible arrays, or (resumed) the arrays holding suspended state
#define ST_(field) ((SLICED && resumed) ? p.sus_##field : p.field)

        const bool fresh = !resumed && (p.mode & MODE_FRESH) != 0;  // default-constructed solvers: the info arrays are write-only
        int status = fresh ? (int)SQPB200_UNINITIALIZED : ST_(status)[b];
        // rho_estimate / res_prim / res_dual (QPSolverInfo, qp.hpp:76-78) only change at checks: kept in shared memory
        if (tid == 0) {
            s_info[0] = fresh ? 0.0 : ST_(rho_estimate)[b];
            s_info[1] = fresh ? 0.0 : ST_(res_prim)[b];
            s_info[2] = fresh ? 0.0 : ST_(res_dual)[b];
            s_info[3] = (PMODE & MODE_FACTOR) ? st.rho : p.rho[b];
            s_cnt[0] = (fresh ? 0 : ST_(rho_updates)[b]) + ((PMODE & MODE_FACTOR) ? 1 : 0);  // rho_vec_update, qp.cpp:313
        }
        const S rho0 = (PMODE & MODE_FACTOR) ? st.rho : p.rho[b];
        const bool reset = (PMODE & MODE_RESET) != 0;

        // ---- per-row state in the owner lanes' registers ----------------------------------------
        S zr[RO], yr[RO], rhor[RO], rinv[RO];
        bool same_classes = true;
#pragma unroll
        for (int t = 0; t < RO; ++t) {
            const int i = own0 + t;
            const bool real = i < m;
            const S lo = real ? gl[i] : -INF, up = real ? gu[i] : INF;
            if (row_primary) *reinterpret_cast<V2 *>(sbnd + 2 * i) = mk2(lo, up);
            zr[t] = (real && !reset) ? ST_(z)[b * m + i] : S(0.0);
            yr[t] = (real && !reset) ? ST_(y)[b * m + i] : S(0.0);
            int typ;
            if (PMODE & MODE_FACTOR) {
                typ = classify_t<S>(lo, up);
                if (real) {
                    if ((PMODE & MODE_REUSE) && p.ctype[b * m + i] != (signed char)typ) same_classes = false;
                    if (row_primary) p.ctype[b * m + i] = (signed char)typ;
                }
            } else {
                typ = real ? p.ctype[b * m + i] : SQPB200_LOOSE_BOUNDS;
            }
            rhor[t] = rho_of_t<S>(typ, rho0);
            rinv[t] = S(1.0) / rhor[t];
        }
        if (tid < NP) {
            sq[tid] = tid < n ? gq[tid] : S(0.0);
            sx[tid] = (tid < n && !reset) ? ST_(x)[b * n + tid] : S(0.0);
        }

        // ---- stage A (zero padded) and pull this lane's tile into registers -----------------------
        S a[R][C];
        // Full-size, 16-byte aligned problems are staged by TMA bulk copies (one per matrix column, into the
        // padded shared-memory columns) issued by warp 0 and awaited on an mbarrier; ragged or unaligned ones
        // fall back to guarded, coalesced loads with zero padding. Callers synchronise the CTA before (no reader
        // of the region is left) and after.
        const bool bulk = Cfg::F64 && n == NP && m == MP && ((reinterpret_cast<uintptr_t>(gA) | reinterpret_cast<uintptr_t>(gP)) & 15) == 0;
        // stage(A?, P?): both land on the same mbarrier phase, so the first staging of a QP overlaps the two loads
        auto stage = [&](bool want_A, bool want_P) {
            if (bulk) {
                if (warp == 0) {
                    fence_proxy_async();
                    if (lane == 0)
                        mbar_expect_tx(&s_mbar, (unsigned)(((want_A ? NP * MP : 0) + (want_P ? NP * NP : 0)) * sizeof(S)));
                    __syncwarp();
                    if (want_A)
                        for (int j = lane; j < NP; j += 32) tma_bulk_g2s(sA + LS * j, gA + (size_t)MP * j, (unsigned)(MP * sizeof(S)), &s_mbar);
                    if (want_P)
                        for (int j = lane; j < NP; j += 32) tma_bulk_g2s(sP + HS * j, gP + (size_t)NP * j, (unsigned)(NP * sizeof(S)), &s_mbar);
                }
                mbar_wait(&s_mbar, mbar_parity);
                mbar_parity ^= 1;
                return;
            }
            if (want_A)
                for (int e = tid; e < NP * MP; e += T) {
                    const int i = e % MP, j = e / MP;
                    sA[i + LS * j] = (i < m && j < n) ? gA[i + (size_t)m * j] : S(0.0);
                }
            if (want_P)
                for (int e = tid; e < NP * NP; e += T) {
                    const int i = e % NP, j = e / NP;
                    sP[i + HS * j] = (i < n && j < n) ? gP[i + (size_t)n * j] : S(0.0);
                }
        };
        // H^-1 from the staged A, kept in registers hv[HR][HC]: SYRK, then symmetric elimination.
        // Requires sA staged; ends with sP valid (it replaces the staging copy) and the CTA synchronised.
        S hv[HR][HC];
        auto factorize = [&]() -> bool {
            // lower triangle of P mirrored (LDLT<Lower> reads nothing else), + sigma on the diagonal;
            // padded variables get a unit diagonal block
            auto h_init = [&](int i, int j) -> S {
                if (i >= n || j >= n) return (i == j) ? S(1.0) : S(0.0);
                S v;
                if constexpr (Cfg::SEPP) v = (i >= j) ? sP[i + HS * j] : sP[j + HS * i];  // P is resident in shared memory
                else v = (i >= j) ? gP[i + (size_t)n * j] : gP[j + (size_t)n * i];
                return (i == j) ? v + sigma : v;
            };
            if constexpr (Cfg::DMMA) {
                // The one genuine dense contraction of the path, H += A^T diag(rho) A (n x n x m), runs on the fp64 tensor
                // cores: D(8x8) += A(8x4) B(4x8) with A(i,k) = rho_k A[k][i] and B(k,j) = A[k][j]; both fragments are the same
                // one-LDS.64-per-lane read of the staged A. Accumulators start at 0; P_lowsym + sigma I is added afterwards.
                constexpr int IB = Cfg::IB, JB = Cfg::JB;
                S c0[IB][JB], c1[IB][JB];
#pragma unroll
                for (int ib = 0; ib < IB; ++ib)
#pragma unroll
                    for (int jb = 0; jb < JB; ++jb) c0[ib][jb] = c1[ib][jb] = S(0.0);
                const int fr = lane >> 2, fk = lane & 3;
                const int ksteps = (m + 3) / 4;
                for (int ks = 0; ks < ksteps; ++ks) {
                    const int k = 4 * ks + fk;
                    const S rk = srho[k];
                    S af[IB], bf[JB];
#pragma unroll
                    for (int ib = 0; ib < IB; ++ib) af[ib] = sA[k + LS * (8 * (IB * warp + ib) + fr)] * rk;
#pragma unroll
                    for (int jb = 0; jb < JB; ++jb) bf[jb] = sA[k + LS * (8 * jb + fr)];
#pragma unroll
                    for (int ib = 0; ib < IB; ++ib)
#pragma unroll
                        for (int jb = 0; jb < JB; ++jb)
                            asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0, %1}, {%2}, {%3}, {%0, %1};"
                                         : "+d"(c0[ib][jb]), "+d"(c1[ib][jb])
                                         : "d"(af[ib]), "d"(bf[jb]));
                }
                cta_sync<NW>();  // every warp is done with the staged A: the region now carries the accumulators across lanes
#pragma unroll
                for (int ib = 0; ib < IB; ++ib)
#pragma unroll
                    for (int jb = 0; jb < JB; ++jb) {
                        const int i = 8 * (IB * warp + ib) + fr, j = 8 * jb + 2 * fk;
                        sA[i + HS * j] = c0[ib][jb];
                        sA[i + HS * (j + 1)] = c1[ib][jb];
                    }
                cta_sync<NW>();
#pragma unroll
                for (int s = 0; s < HC; ++s)
#pragma unroll
                    for (int r = 0; r < HR; r += 2) {
                        const V2 h2 = *reinterpret_cast<const V2 *>(sA + i0 + r + HS * (cg + CG * s));
                        hv[r][s] = h_init(i0 + r, cg + CG * s) + h2.x;
                        hv[r + 1][s] = h_init(i0 + r + 1, cg + CG * s) + h2.y;
                    }
            } else {
#pragma unroll
                for (int s = 0; s < HC; ++s)
#pragma unroll
                    for (int r = 0; r < HR; ++r) hv[r][s] = h_init(i0 + r, cg + CG * s);
                const int mloop = (m + 1) / 2;
                for (int kp = 0; kp < mloop; ++kp) {
                    const V2 rr = *reinterpret_cast<const V2 *>(srho + 2 * kp);
                    V2 ar[HR];
    #pragma unroll
                    for (int r = 0; r < HR; ++r) {
                        ar[r] = *reinterpret_cast<const V2 *>(sA + 2 * kp + LS * (i0 + r));
                        ar[r].x *= rr.x;
                        ar[r].y *= rr.y;
                    }
    #pragma unroll
                    for (int s = 0; s < HC; ++s) {
                        const V2 cj = *reinterpret_cast<const V2 *>(sA + 2 * kp + LS * (cg + CG * s));
    #pragma unroll
                        for (int r = 0; r < HR; ++r) {
                            hv[r][s] = fma(ar[r].x, cj.x, hv[r][s]);
                            hv[r][s] = fma(ar[r].y, cj.y, hv[r][s]);
                        }
                    }
                }
            }
            if (tid == 0) s_fail = 0;
            cta_sync<NW>();  // everyone is done with sA; the staging region may be overwritten from here on
            bool ok = true;
            if constexpr (SWEEP2) {
                // Symmetric elimination, TWO pivots per step (NP/2 steps, one barrier each): with K = {k, k+1},
                // E = S[K][K], S <- S - S[:,K] E^-1 S[K,:],  S[:,K] <- S[:,K] E^-1,  S[K,:] <- E^-1 S[K,:],  S[K][K] <- -E^-1.
                // After all steps hv holds -(H^-1). The pivots of the unpivoted LDL^T are d_k = E00 and d_k+1 = det(E)/E00: a zero
                // or NaN in either reports NUMERICAL_ISSUES (Eigen::LDLT::info() != Success). Every lane inverts the 2x2 block
                // itself after the barrier, so no lane waits on a pivot owner's reciprocal.
    #pragma unroll
                for (int s = 0; s < HC; ++s) {
    #pragma unroll 1
                    for (int cgk = 0; cgk < CG; cgk += 2) {
                        const int k = cgk + CG * s;
                        S *pvA = piv + ((k >> 1) & 1) * Cfg::PIVS, *pvB = pvA + NP;
                        if (cg == cgk || cg == cgk + 1) {
                            S *dst = (cg == cgk) ? pvA : pvB;
    #pragma unroll
                            for (int r = 0; r < HR; r += 2) *reinterpret_cast<V2 *>(dst + i0 + r) = mk2(hv[r][s], hv[r + 1][s]);
                        }
                        cta_sync<NW>();
                        const V2 ab = *reinterpret_cast<const V2 *>(pvA + k);  // E00, E10
                        const S e_c = pvB[k + 1];                                     // E11
                        const S det = fma(ab.x, e_c, -ab.y * ab.y);
                        if (!(fabs(ab.x) > S(0.0)) || !(fabs(det) > S(0.0))) {
                            ok = false;
                            break;
                        }
                        const S inv_det = S(1.0) / det;
                        const S e00 = e_c * inv_det, e01 = -ab.y * inv_det, e11 = ab.x * inv_det;  // E^-1
                        S wA[HR], wB[HR];  // rows of S[:,K] E^-1
    #pragma unroll
                        for (int r = 0; r < HR; r += 2) {
                            const V2 ca = *reinterpret_cast<const V2 *>(pvA + i0 + r);
                            const V2 cb = *reinterpret_cast<const V2 *>(pvB + i0 + r);
                            wA[r] = fma(ca.x, e00, cb.x * e01);
                            wB[r] = fma(ca.x, e01, cb.x * e11);
                            wA[r + 1] = fma(ca.y, e00, cb.y * e01);
                            wB[r + 1] = fma(ca.y, e01, cb.y * e11);
                        }
                        const int krow = k - i0;  // pivot rows k, k+1 are rows krow, krow+1 of this lane's tile if 0 <= krow < HR
    #pragma unroll
                        for (int s2 = 0; s2 < HC; ++s2) {
                            const S cja = pvA[cg + CG * s2], cjb = pvB[cg + CG * s2];
    #pragma unroll
                            for (int r = 0; r < HR; ++r) hv[r][s2] = fma(-wB[r], cjb, fma(-wA[r], cja, hv[r][s2]));
                        }
                        if ((unsigned)krow < (unsigned)HR) {  // the row group that holds the two pivot rows
    #pragma unroll
                            for (int s2 = 0; s2 < HC; ++s2) {
                                const S cja = pvA[cg + CG * s2], cjb = pvB[cg + CG * s2];
                                const S ra = fma(e00, cja, e01 * cjb), rb = fma(e01, cja, e11 * cjb);
    #pragma unroll
                                for (int r = 0; r < HR; ++r) hv[r][s2] = (r == krow) ? ra : ((r == krow + 1) ? rb : hv[r][s2]);
                            }
                        }
                        if (cg == cgk) {
    #pragma unroll
                            for (int r = 0; r < HR; ++r) hv[r][s] = (r == krow) ? -e00 : ((r == krow + 1) ? -e01 : wA[r]);
                        } else if (cg == cgk + 1) {
    #pragma unroll
                            for (int r = 0; r < HR; ++r) hv[r][s] = (r == krow) ? -e01 : ((r == krow + 1) ? -e11 : wB[r]);
                        }
                    }
                    if (!ok) break;
                }
            } else {
                // symmetric elimination, pivot by pivot: after all NP steps hv holds -(H^-1)
    #pragma unroll
                for (int s = 0; s < HC; ++s) {
    #pragma unroll 1
                    for (int cgk = 0; cgk < CG; ++cgk) {
                        const int k = cgk + CG * s;
                        S *pv = piv + (k & 1) * (NP + 2);
                        const int krow = k - i0;  // row of this lane's tile that is the pivot row (if 0 <= krow < HR)
                        if (cg == cgk) {
    #pragma unroll
                            for (int r = 0; r < HR; r += 2) *reinterpret_cast<V2 *>(pv + i0 + r) = mk2(hv[r][s], hv[r + 1][s]);
                            if ((unsigned)krow < (unsigned)HR) {  // the one lane that owns the pivot publishes d and 1/d
                                S d = hv[0][s];
    #pragma unroll
                                for (int r = 1; r < HR; ++r) d = (r == krow) ? hv[r][s] : d;
                                pv[NP] = S(1.0) / d;
                                pv[NP + 1] = d;
                            }
                        }
                        cta_sync<NW>();
                        const V2 dd = *reinterpret_cast<const V2 *>(pv + NP);
                        const S inv_d = dd.x;
                        if (!(fabs(dd.y) > S(0.0))) {  // zero or NaN pivot: Eigen::LDLT::info() != Success
                            ok = false;
                            break;
                        }
                        S t[HR];
    #pragma unroll
                        for (int r = 0; r < HR; r += 2) {
                            const V2 ci = *reinterpret_cast<const V2 *>(pv + i0 + r);
                            t[r] = ci.x * inv_d;
                            t[r + 1] = ci.y * inv_d;
                        }
    #pragma unroll
                        for (int s2 = 0; s2 < HC; ++s2) {
                            const S cj = pv[cg + CG * s2];
    #pragma unroll
                            for (int r = 0; r < HR; ++r) hv[r][s2] = fma(-t[r], cj, hv[r][s2]);
                        }
                        if ((unsigned)krow < (unsigned)HR) {  // the row group holding pivot row k: that row becomes c_j / d
    #pragma unroll
                            for (int s2 = 0; s2 < HC; ++s2) {
                                const S rowv = pv[cg + CG * s2] * inv_d;
    #pragma unroll
                                for (int r = 0; r < HR; ++r) hv[r][s2] = (r == krow) ? rowv : hv[r][s2];
                            }
                        }
                        if (cg == cgk) {
    #pragma unroll
                            for (int r = 0; r < HR; ++r) hv[r][s] = (r == krow) ? -inv_d : t[r];
                        }
                    }
                    if (!ok) break;
                }
            }
#pragma unroll
            for (int s = 0; s < HC; ++s)
#pragma unroll
                for (int r = 0; r < HR; ++r) hv[r][s] = -hv[r][s];
            TL::permute_sym(hv, lane);
            if constexpr (!Cfg::SEPP) {
                cta_sync<NW>();
                stage(false, true);  // P replaces the staging copy of A
            }
            cta_sync<NW>();
            return ok;
        };

        cta_sync<NW>();
        stage(true, Cfg::SEPP);
        cta_sync<NW>();
#pragma unroll
        for (int kk = 0; kk < C; ++kk) {
            const S *colp = sA + row0 + LS * TL::colA(lc, kk ^ TL::perm_cols(lane));  // slot (kr, kk) <- logical (kr ^ px, kk ^ py)
            if constexpr (R >= 2) {
#pragma unroll
                for (int kr = 0; kr < R; kr += 2) {
                    const V2 t2 = *reinterpret_cast<const V2 *>(colp + (kr ^ TL::perm_rows(lane)));
                    a[kr][kk] = t2.x;
                    a[kr + 1][kk] = t2.y;
                }
            } else {
                a[0][kk] = colp[0];
            }
        }
        if constexpr (SLICED) {
            // the first slice of a QP whose inputs live in another GPU's memory leaves a local copy of A for the resumes
            if (!resumed && p.loc_A) {
                double *dst = p.loc_A + (size_t)bi * m * n;
                for (int e = tid; e < n * m; e += T) dst[e] = sA[(e % m) + LS * (e / m)];
                // q, l, u as well (out of the shared-memory copies made above): a resume then reads nothing over NVLink
                if (tid < n) p.loc_q[(size_t)bi * n + tid] = sq[tid];
                if (row_primary) {
#pragma unroll
                    for (int t = 0; t < RO; ++t) {
                        const int i = own0 + t;
                        if (i < m) {
                            p.loc_l[(size_t)bi * m + i] = sbnd[2 * i];
                            p.loc_u[(size_t)bi * m + i] = sbnd[2 * i + 1];
                        }
                    }
                }
            }
        }
        S xown = S(0);  // x entry of the row this lane ends up with after the H^-1 b tree (sx was written before the staging barriers)
        if constexpr (NW > 1) {
            bool prim;
            xown = sx[HR * rg + halve_base<HR, CG / 2, 1>(lane, prim)];
        }
        bool do_factor = (PMODE & MODE_FACTOR) != 0, first_factor = true, in_solve = false;
        if (do_factor && (PMODE & MODE_REUSE)) {
            // same P and A as the launch that kept the factor: reuse it where classes and rho are unchanged
            // (duplicate owner lanes read the old classes before the primary wrote the same row: benign, values equal or both differ)
            // (compared in the compute scalar: the fp32 instantiation tags the slab with the rounded rho)
            const bool reuse = cta_all<NW>(same_classes && (S)p.fact_rho[b] == (S)st.rho);
            if (reuse) {
                do_factor = false;
                status = SQPB200_UNSOLVED;
            }
        }
        if (!do_factor) {
            const double *gF = p.fact + b * n * n;
#pragma unroll
            for (int s = 0; s < HC; ++s) {
                const int j = cg + CG * s;
#pragma unroll
                for (int r = 0; r < HR; ++r) {
                    const int i = i0 + (r ^ TL::perm_sym(lane));
                    hv[r][s] = (i < n && j < n) ? gF[i + (size_t)n * j] : (i == j ? S(1.0) : S(0.0));
                }
            }
            if constexpr (!Cfg::SEPP) {
                cta_sync<NW>();  // tiles are in registers; the staging area may be overwritten
                stage(false, true);
                cta_sync<NW>();
                first_factor = false;  // the staging copy of A is gone
            }
        }

        long long executed = 0;
        int iter = 0;
        bool suspended = false;
        int iter_lim = st.max_iter;  // SLICED: last iteration of this slice
        // Outer loop: one (re)factorisation, then ADMM iterations until convergence, max_iter, or the
        // next adaptive-rho refactorisation.  A single factorisation call site keeps the code compact.
        for (;;) {
            if (do_factor) {
                cta_sync<NW>();
                if (!first_factor) stage(true, false);  // adaptive-rho refactorisation: bring A back (P is reloaded after)
                if (row_primary) {
#pragma unroll
                    for (int t = 0; t < RO; ++t) srho[own0 + t] = rhor[t];
                }
                cta_sync<NW>();
                const bool ok = factorize();
                if constexpr (SLICED) {
                    if (first_factor && !resumed && p.loc_P) {  // P was just staged: local copy for the resumes
                        double *dst = p.loc_P + (size_t)bi * n * n;
                        for (int e = tid; e < n * n; e += T) dst[e] = sP[(e % n) + HS * (e / n)];
                    }
                }
                first_factor = false;
                do_factor = false;
                if (!in_solve) {
                    status = ok ? SQPB200_UNSOLVED : SQPB200_NUMERICAL_ISSUES;  // qp.cpp:39-43
                    // A factorisation that is not written to the slab leaves an older kept factor behind whose classes were just
                    // overwritten: forget it (also on failure), or a later REUSE would match the new classes against the old factor.
                    if (tid == 0 && !(ok && (PMODE & MODE_KEEP_INITIAL))) p.fact_rho[b] = __longlong_as_double(0x7ff8000000000000LL);
                    if (ok && (PMODE & MODE_KEEP_INITIAL)) {
                        double *gF = p.fact + b * n * n;
#pragma unroll
                        for (int s = 0; s < HC; ++s) {
                            const int j = cg + CG * s;
#pragma unroll
                            for (int r = 0; r < HR; ++r) {
                                const int i = i0 + (r ^ TL::perm_sym(lane));
                                if (i < n && j < n) gF[i + (size_t)n * j] = hv[r][s];
                            }
                        }
                        if (tid == 0) p.fact_rho[b] = st.rho;
                    }
                } else if (!ok) {
                    status = SQPB200_NUMERICAL_ISSUES;  // qp.cpp:139-142 (break before the loop increment)
                    iter -= 1;
                    break;
                }
            }
            if (!in_solve) {
                if (!((PMODE & MODE_SOLVE) && status != SQPB200_UNINITIALIZED && status != SQPB200_NUMERICAL_ISSUES)) break;
                in_solve = true;
                iter = 1;
                if constexpr (SLICED) {
                    if (resumed) iter = p.sus_iter[b];  // the next iteration of the suspended solve
                    iter_lim = min(st.max_iter, iter + (resumed ? p.slice_iters : p.slice_first) - 1);
                }
            }
            bool refactor = false;
            // iterations until the next termination check / adaptive-rho step (replaces iter % N, qp.cpp:105,125)
            int to_chk = st.check_termination > 0 ? st.check_termination - (iter - 1) % st.check_termination : 0x7fffffff;
            int to_adapt = (st.adaptive_rho && st.adaptive_rho_interval > 0)
                               ? st.adaptive_rho_interval - (iter - 1) % st.adaptive_rho_interval : 0x7fffffff;
            bool stopped = false;  // left the loop by a break (converged / refactorisation), not by running into the limit
            for (; iter <= (SLICED ? iter_lim : st.max_iter); ++iter) {
                // P1: w = rho .* z - y (owner lanes) -> sw; partial A^T w over this warp's rows -> part[warp]
                if (row_primary) {
#pragma unroll
                    for (int t = 0; t < RO; ++t) sw[own0 + t] = rhor[t] * zr[t] - yr[t];
                }
                __syncwarp();
                TL::mv_At(a, sw, part + warp * NP, row0, lc, lane);
                cta_sync<NW>();
                // P2: b = sigma x - q + A^T w   (rhs of qp.cpp:272-276 pushed through the (2,2) block)
                if (tid < NP) {
                    S pw[NW];
#pragma unroll
                    for (int w2 = 0; w2 < NW; ++w2) pw[w2] = part[w2 * NP + tid];
#pragma unroll
                    for (int h = NW / 2; h >= 1; h /= 2)
#pragma unroll
                        for (int w2 = 0; w2 < h; ++w2) pw[w2] += pw[w2 + h];
                    sb[tid] = fma(sigma, sx[tid], pw[0] - sq[tid]);
                }
                cta_sync<NW>();
                // P3: x~ = H^-1 b ; x = alpha x~ + (1 - alpha) x   (qp.cpp:90-96)
                {
                    int row;
                    bool prim;
                    const S xt = TL::mv_sym_reg(hv, sb, rg, cg, lane, row, prim);
                    // the owner lanes carry their x entry in a register: x~ reaches shared memory (and the barrier) straight from the
                    // tree, the x update and its store (first read by the b stage of the NEXT iteration) leave the critical path
                    if constexpr (NW > 1) {
                        if (prim) sxt[row] = xt;
                        cta_sync<NW>();
                        // (sx is next read by the b stage of the following iteration, two barriers on; the check block synchronises first)
                        xown = alpha * xt + (S(1.0) - alpha) * xown;
                        if (prim) sx[row] = xown;
                    } else {  // one warp per QP: the "barrier" is a __syncwarp, nothing to hide
                        if (prim) {
                            sxt[row] = xt;
                            sx[row] = alpha * xt + (S(1.0) - alpha) * sx[row];
                        }
                        cta_sync<NW>();
                    }
                }
                // P4: z~ = A x~ ; z, y updates in the owner lanes (qp.cpp:93-103)
                {
                    S zt[RO];
                    TL::mv_A(a, sxt, lc, lane, zt);
#pragma unroll
                    for (int t = 0; t < RO; ++t) {
                        const S zh = alpha * zt[t] + (S(1.0) - alpha) * zr[t];
                        const V2 bd = *reinterpret_cast<const V2 *>(sbnd + 2 * (own0 + t));
                        const S zn = box_project(zh + rinv[t] * yr[t], bd.x, bd.y);
                        yr[t] = yr[t] + rhor[t] * (zh - zn);
                        zr[t] = zn;
                    }
                }
                const bool chk = --to_chk == 0;
                const bool adapt = --to_adapt == 0;
                if (chk) to_chk = st.check_termination;
                if (adapt) to_adapt = st.adaptive_rho_interval;
                if (chk || adapt) {
                    // update_state, qp.cpp:316-331
                    if constexpr (NW > 1) cta_sync<NW>();  // the x entries stored after the last barrier of the iteration are read below
                    S mx[7] = {0, 0, 0, 0, 0, 0, 0};  // |Ax| |z| |Px| |A^T y| |q| |Ax - z| |Px + q + A^T y|
                    {
                        S ax[RO];  // A x (x, not x~: they differ when alpha != 1), qp.cpp:319
                        TL::mv_A(a, sx, lc, lane, ax);
#pragma unroll
                        for (int t = 0; t < RO; ++t) {
                            mx[0] = absmax(mx[0], ax[t]);
                            mx[1] = absmax(mx[1], zr[t]);
                            mx[5] = absmax(mx[5], ax[t] - zr[t]);
                        }
                    }
                    if (row_primary) {
#pragma unroll
                        for (int t = 0; t < RO; ++t) sw[own0 + t] = yr[t];
                    }
                    __syncwarp();
                    TL::mv_At(a, sw, part + warp * NP, row0, lc, lane);
                    {
                        int row;
                        bool prim;
                        const S px = TL::mv_sym_smem(sP, sx, rg, cg, lane, row, prim);
                        if (prim) spx[row] = px;
                    }
                    cta_sync<NW>();
                    if (tid < NP) {
                        S aty = part[tid];
#pragma unroll
                        for (int w2 = 1; w2 < NW; ++w2) aty += part[w2 * NP + tid];
                        const S px = spx[tid], qv = sq[tid];
                        mx[2] = absmax(mx[2], px);
                        mx[3] = absmax(mx[3], aty);
                        mx[4] = absmax(mx[4], qv);
                        mx[6] = absmax(mx[6], px + qv + aty);
                    }
#pragma unroll
                    for (int k = 0; k < 7; ++k) {
                        const S v = warp_max(mx[k]);
                        if (lane == 0) red[k * NW + warp] = v;
                    }
                    const S rho = s_info[3];  // read before the barrier: thread 0 rewrites it after the decision below
                    cta_sync<NW>();
#pragma unroll
                    for (int k = 0; k < 7; ++k) {
                        S v = red[k * NW];
#pragma unroll
                        for (int w2 = 1; w2 < NW; ++w2) v = red[k * NW + w2] > v ? red[k * NW + w2] : v;
                        mx[k] = v;
                    }
                    const S sc_p = fmax(mx[0], mx[1]);
                    const S sc_d = fmax(mx[2], fmax(mx[3], mx[4]));
                    const S res_prim = mx[5], res_dual = mx[6];
                    if (tid == 0) {
                        s_info[1] = res_prim;
                        s_info[2] = res_dual;
                    }
                    if (chk) {  // termination_criteria, qp.cpp:363-371
                        if (res_prim <= st.eps_abs + st.eps_rel * sc_p && res_dual <= st.eps_abs + st.eps_rel * sc_d) {
                            status = SQPB200_SOLVED;
                            stopped = true;
                            break;
                        }
                    }
                    if (adapt) {  // qp.cpp:125-144
                        const S new_rho = rho_estimate_clamped_t<S>(rho, res_prim, res_dual, sc_p, sc_d);
                        if (tid == 0) s_info[0] = new_rho;
                        if (new_rho < rho / st.adaptive_rho_tolerance || new_rho > rho * st.adaptive_rho_tolerance) {
                            if (tid == 0) {
                                s_info[3] = new_rho;
                                s_cnt[0] += 1;
                            }
#pragma unroll
                            for (int t = 0; t < RO; ++t) {
                                const int i = own0 + t;  // constraint classes were fixed by setup (qp.cpp:31): read them back
                                const int typ = i < m ? p.ctype[b * m + i] : SQPB200_LOOSE_BOUNDS;
                                rhor[t] = rho_of_t<S>(typ, new_rho);
                                rinv[t] = S(1.0) / rhor[t];
                            }
                            refactor = true;
                            stopped = true;
                            break;
                        }
                    }
                }
            }
            if constexpr (SLICED) {
                // ran into the end of the slice (not of the solve): suspend; `iter` is the iteration the resume starts with
                if (!stopped && iter <= st.max_iter) suspended = true;
            }
            if (!refactor) break;
            ++iter;  // the reference finishes the iteration (loop increment) after refactoring
            do_factor = true;
        }
        if (in_solve && !suspended) {
            executed = iter <= st.max_iter ? iter : st.max_iter;
            if (iter > st.max_iter) status = SQPB200_MAX_ITER_EXCEEDED;  // qp.cpp:147-149
        }

        // ---- write back -----------------------------------------------------------------------------
        cta_sync<NW>();
        if constexpr (SLICED) {
            if (suspended) {
                // park the solve: iterates, info and H^-1 into the object's arrays, then publish the QP in the re-queue ring
                if (tid < n) p.sus_x[b * n + tid] = sx[tid];
                if (row_primary) {
#pragma unroll
                    for (int t = 0; t < RO; ++t) {
                        const int i = own0 + t;
                        if (i < m) {
                            p.sus_z[b * m + i] = zr[t];
                            p.sus_y[b * m + i] = yr[t];
                        }
                    }
                }
                double *gF = p.fact + b * n * n;
#pragma unroll
                for (int s = 0; s < HC; ++s) {
                    const int j = cg + CG * s;
#pragma unroll
                    for (int r = 0; r < HR; ++r) {
                        const int i = i0 + (r ^ TL::perm_sym(lane));
                        if (i < n && j < n) gF[i + (size_t)n * j] = hv[r][s];
                    }
                }
                if (tid == 0) {
                    p.sus_status[b] = status;
                    p.sus_iter[b] = iter;
                    p.sus_rho_updates[b] = s_cnt[0];
                    p.sus_rho_estimate[b] = s_info[0];
                    p.sus_res_prim[b] = s_info[1];
                    p.sus_res_dual[b] = s_info[2];
                    p.rho[b] = s_info[3];
                }
                __threadfence();
                cta_sync<NW>();
                if (tid == 0) {
                    const int e = atomicAdd(p.rq_alloc, 1);
                    *reinterpret_cast<volatile int *>(p.rq + e) = local;
                }
                continue;
            }
        }
        if (tid < n) p.x[b * n + tid] = sx[tid];
        if (row_primary) {
#pragma unroll
            for (int t = 0; t < RO; ++t) {
                const int i = own0 + t;
                if (i < m) {
                    p.z[b * m + i] = zr[t];
                    p.y[b * m + i] = yr[t];
                }
            }
        }
        if ((PMODE & MODE_STORE_FACTOR) && status != SQPB200_NUMERICAL_ISSUES && status != SQPB200_UNINITIALIZED) {
            double *gF = p.fact + b * n * n;
#pragma unroll
            for (int s = 0; s < HC; ++s) {
                const int j = cg + CG * s;
#pragma unroll
                for (int r = 0; r < HR; ++r) {
                    const int i = i0 + (r ^ TL::perm_sym(lane));
                    if (i < n && j < n) gF[i + (size_t)n * j] = hv[r][s];
                }
            }
            if (tid == 0) p.fact_rho[b] = s_info[3];
        }
        if (tid == 0) {
            p.status[b] = status;
            if (in_solve) p.iter[b] = iter;  // qp.cpp:150
            p.rho_updates[b] = s_cnt[0];
            p.rho_estimate[b] = s_info[0];
            p.res_prim[b] = s_info[1];
            p.res_dual[b] = s_info[2];
            p.rho[b] = s_info[3];
            if (executed) atomicAdd(p.total_iters, (unsigned long long)executed);
        }
        if constexpr (SLICED) {  // this QP is finished: the launch ends when all are
            __threadfence();
            cta_sync<NW>();
            if (tid == 0) atomicAdd(p.done, 1);
        }
    }
}

#undef PMODE
#undef ST_
#undef gP
#undef gA
#undef gq
#undef gl
#undef gu

// ---- dispatch ------------------------------------------------------------------------------------
using Cfg64x128w4 = TileCfg<64, 128, 4, 8, 4, 2>;  // 128 threads: 8x8 A tile + 4x8 H^-1 tile per lane
// the same with P in a region of its own: adaptive-rho launches (S2 on config 3: 5.11 -> 5.02 ms; without refactorisations the larger
// shared-memory footprint costs 2 %: 23.04 against 22.56 ms)
using Cfg64x128w4P = TileCfg<64, 128, 4, 8, 4, 2, double, true, true>;
using Cfg64x128w8 = TileCfg<64, 128, 8, 8, 2, 2>;  // 256 threads: 4x8 A tile + 2x8 H^-1 tile per lane
using Cfg32x64 = TileCfg<32, 64, 2, 4, 2, 8>;      // two warps per QP
using Cfg32x64w1 = TileCfg<32, 64, 1, 4, 4, 8>;    // ONE warp per QP: no CTA barrier anywhere (BASELINE config 2's mapping)
using Cfg16x32 = TileCfg<16, 32, 1, 4, 2, 16>;
using Cfg8x16 = TileCfg<8, 16, 1, 4, 2, 16>;
// fp32 compute (QPSolver<float>): half the registers and shared memory per QP, so more resident CTAs per SM. (Four CTAs per SM at 128
// registers with a few spills and three at 168 registers without any measure the same: 13.36 / 13.43 ms on config 3.)
using Cfg64x128w4f = TileCfg<64, 128, 4, 8, 4, 4, float>;
using Cfg32x64w1f = TileCfg<32, 64, 1, 4, 4, 16, float>;
using Cfg16x32f = TileCfg<16, 32, 1, 4, 2, 16, float>;
using Cfg8x16f = TileCfg<8, 16, 1, 4, 2, 16, float>;

bool tile_supported(int n, int m) { return n >= 1 && m >= 0 && n <= 64 && m <= 128; }
// time slicing is instantiated for the 64 x 128 class in fp64 with four warps per QP (the headline configuration)
bool tile_sliceable(int n, int m, int tile_warps, int f32) { return tile_supported(n, m) && (n > 32 || m > 64) && !f32 && tile_warps != 8; }
int tile_slots(int sm_count) { return 2 * sm_count; }  // resident CTAs of that configuration

template <class Cfg, bool SWEEP2, bool SLICED>
static cudaError_t launch_one(const KernelParams &p, int sm_count, int ctas_per_sm, cudaStream_t stream, char *name, size_t name_len) {
    auto kernel = qp_tile_kernel<Cfg, SWEEP2, SLICED>;
    cudaError_t e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)Cfg::SMEM_BYTES);
    if (e != cudaSuccess) return e;
    int occ = 0;
    e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kernel, Cfg::T, Cfg::SMEM_BYTES);
    if (e != cudaSuccess) return e;
    if (occ < 1) return cudaErrorLaunchOutOfResources;
    if (ctas_per_sm > 0 && ctas_per_sm < occ) occ = ctas_per_sm;
    long long grid = (long long)sm_count * occ;  // persistent: a multiple of the SM count
    if (grid > p.count) grid = p.count;
    if (name) snprintf(name, name_len, "tile<%d,%d,%d%s%s>x%d%s", Cfg::NP, Cfg::MP, Cfg::NW, Cfg::F64 ? "" : ",f32", Cfg::SEPP ? ",sepP" : "", occ, SLICED ? "/sliced" : "");  // (the sweep variant is not part of the name)
    kernel<<<(int)grid, Cfg::T, Cfg::SMEM_BYTES, stream>>>(p);
    return cudaGetLastError();
}
// CfgA: the configuration of launches with adaptive rho (frequent refactorisations: the rank-2 sweep, and for the headline class P in a
// shared-memory region of its own); Cfg: iteration-dominated launches
template <class Cfg, bool SLICED = false, class CfgA = Cfg>
static cudaError_t launch_cfg(const KernelParams &p, int sm_count, int ctas_per_sm, cudaStream_t stream, char *name, size_t name_len) {
    if (p.s.adaptive_rho) return launch_one<CfgA, true, SLICED>(p, sm_count, ctas_per_sm, stream, name, name_len);
    return launch_one<Cfg, false, SLICED>(p, sm_count, ctas_per_sm, stream, name, name_len);
}

cudaError_t launch_tile(const KernelParams &p, int sm_count, int ctas_per_sm, int tile_warps, int f32, cudaStream_t stream, char *name,
                        size_t name_len) {
    if (f32) {  // fp32 compute: one configuration per size class
        if (p.n <= 8 && p.m <= 16) return launch_cfg<Cfg8x16f>(p, sm_count, ctas_per_sm, stream, name, name_len);
        if (p.n <= 16 && p.m <= 32) return launch_cfg<Cfg16x32f>(p, sm_count, ctas_per_sm, stream, name, name_len);
        if (p.n <= 32 && p.m <= 64) return launch_cfg<Cfg32x64w1f>(p, sm_count, ctas_per_sm, stream, name, name_len);
        return launch_cfg<Cfg64x128w4f>(p, sm_count, ctas_per_sm, stream, name, name_len);
    }
    if (p.n <= 8 && p.m <= 16) return launch_cfg<Cfg8x16>(p, sm_count, ctas_per_sm, stream, name, name_len);
    if (p.n <= 16 && p.m <= 32) return launch_cfg<Cfg16x32>(p, sm_count, ctas_per_sm, stream, name, name_len);
    if (p.n <= 32 && p.m <= 64) {
        // one warp per QP is the faster mapping here (5.8 vs 6.9 ms on 8192 QPs); the two-warp variant stays selectable
        if (tile_warps == 2) return launch_cfg<Cfg32x64>(p, sm_count, ctas_per_sm, stream, name, name_len);
        return launch_cfg<Cfg32x64w1>(p, sm_count, ctas_per_sm, stream, name, name_len);
    }
    if (tile_warps == 8) return launch_cfg<Cfg64x128w8>(p, sm_count, ctas_per_sm, stream, name, name_len);
    if (p.slice_iters > 0) return launch_cfg<Cfg64x128w4, true>(p, sm_count, ctas_per_sm, stream, name, name_len);
    return launch_cfg<Cfg64x128w4, false, Cfg64x128w4P>(p, sm_count, ctas_per_sm, stream, name, name_len);
}

}  // namespace sqpb200
