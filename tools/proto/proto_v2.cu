// Prototype of the two-stage ADMM iteration (stage 1: b = sigma x - q + A^T w from registers; stage 2: [x~; z~] = [H^-1; A H^-1] b
// from registers). Measures cycles per iteration of ONE 256-thread CTA per SM. Not product code.
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>
constexpr int N = 64, M = 128, T = 256;
__device__ __forceinline__ double clipd(double z, double l, double u) { z = z < l ? l : z; return u < z ? u : z; }
__global__ void __launch_bounds__(T, 1) proto(const double *gA, const double *gG, const double *gv, double *out, long long *cyc, int iters,
                                              double sigma, double alpha) {
    __shared__ __align__(16) double sw[4 * 34], sb[4 * 18];
#define SW(i) sw[((i) >> 5) * 34 + ((i) & 31)]
#define SB(i) sb[((i) >> 4) * 18 + ((i) & 15)]
    const int tid = threadIdx.x, lane = tid & 31;
    const int j = tid >> 2, sub = tid & 3;
    const size_t qp = blockIdx.x;
    double a[32], g[3][16];
#pragma unroll
    for (int r = 0; r < 32; ++r) a[r] = gA[qp * N * M + (size_t)j * M + 32 * sub + r];
    // rows of the stage-2 group j: x row j, z rows 2j, 2j+1 (G rows 64 + 2j, 64 + 2j + 1)
#pragma unroll
    for (int c = 0; c < 16; ++c) {
        g[0][c] = gG[qp * 192 * N + (size_t)j * N + 16 * sub + c];
        g[1][c] = gG[qp * 192 * N + (size_t)(64 + 2 * j) * N + 16 * sub + c];
        g[2][c] = gG[qp * 192 * N + (size_t)(64 + 2 * j + 1) * N + 16 * sub + c];
    }
    double x = 0, qv = gv[j], z = 0, y = 0;
    const int zrow = 2 * j + (sub - 1);
    const bool zown = sub == 1 || sub == 2;
    double lo = zown ? gv[64 + zrow] - 1.0 : 0, up = zown ? gv[64 + zrow] + 1.0 : 0;
    const double rho = 0.1, rinv = 10.0;
    if (tid < M) SW(tid) = 0.0;
    __syncthreads();
    long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
        // stage 1
        double s0 = 0, s1 = 0, s2 = 0, s3 = 0;
#pragma unroll
        for (int r = 0; r < 32; r += 4) {
            const double2 w0 = *reinterpret_cast<const double2 *>(sw + 34 * sub + r);
            const double2 w1 = *reinterpret_cast<const double2 *>(sw + 34 * sub + r + 2);
            s0 = fma(a[r], w0.x, s0);
            s1 = fma(a[r + 1], w0.y, s1);
            s2 = fma(a[r + 2], w1.x, s2);
            s3 = fma(a[r + 3], w1.y, s3);
        }
        double gsum = (s0 + s1) + (s2 + s3);
        gsum += __shfl_xor_sync(0xffffffffu, gsum, 1);
        gsum += __shfl_xor_sync(0xffffffffu, gsum, 2);
        if (sub == 0) SB(j) = fma(sigma, x, gsum - qv);
        __syncthreads();
        // stage 2
        double r0 = 0, r1 = 0, r2 = 0, r0b = 0, r1b = 0, r2b = 0;
#pragma unroll
        for (int c = 0; c < 16; c += 2) {
            const double2 b2 = *reinterpret_cast<const double2 *>(sb + 18 * sub + c);
            r0 = fma(g[0][c], b2.x, r0);
            r1 = fma(g[1][c], b2.x, r1);
            r2 = fma(g[2][c], b2.x, r2);
            r0b = fma(g[0][c + 1], b2.y, r0b);
            r1b = fma(g[1][c + 1], b2.y, r1b);
            r2b = fma(g[2][c + 1], b2.y, r2b);
        }
        r0 += r0b; r1 += r1b; r2 += r2b;
        r0 += __shfl_xor_sync(0xffffffffu, r0, 1);
        r1 += __shfl_xor_sync(0xffffffffu, r1, 1);
        r2 += __shfl_xor_sync(0xffffffffu, r2, 1);
        r0 += __shfl_xor_sync(0xffffffffu, r0, 2);
        r1 += __shfl_xor_sync(0xffffffffu, r1, 2);
        r2 += __shfl_xor_sync(0xffffffffu, r2, 2);
        if (zown) {
            const double zt = sub == 1 ? r1 : r2;
            const double zh = alpha * zt + (1.0 - alpha) * z;
            const double zn = clipd(zh + rinv * y, lo, up);
            y = y + rho * (zh - zn);
            z = zn;
            SW(zrow) = rho * z - y;
        } else if (sub == 0) {
            x = alpha * r0 + (1.0 - alpha) * x;
        }
        __syncthreads();
    }
    long long t1 = clock64();
    if (tid == 0) cyc[qp] = t1 - t0;
    out[qp * T + tid] = x + z + y;
}
int main(int argc, char **argv) {
    int iters = argc > 1 ? atoi(argv[1]) : 2000;
    int grid = argc > 2 ? atoi(argv[2]) : 148;
    double *A, *G, *v, *out; long long *cyc;
    size_t nA = (size_t)grid * N * M, nG = (size_t)grid * 192 * N;
    cudaMalloc(&A, nA * 8); cudaMalloc(&G, nG * 8); cudaMalloc(&v, 256 * 8); cudaMalloc(&out, (size_t)grid * T * 8); cudaMalloc(&cyc, grid * 8);
    double *h = (double *)malloc(nG * 8);
    srand(1);
    for (size_t i = 0; i < nG; ++i) h[i] = (rand() / (double)RAND_MAX - 0.5) * 0.05;
    cudaMemcpy(A, h, nA * 8, cudaMemcpyHostToDevice); cudaMemcpy(G, h, nG * 8, cudaMemcpyHostToDevice); cudaMemcpy(v, h, 256 * 8, cudaMemcpyHostToDevice);
    for (int rep = 0; rep < 3; ++rep) {
        cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
        cudaEventRecord(e0);
        proto<<<grid, T>>>(A, G, v, out, cyc, iters, 1e-6, rep == 2 ? 1.6 : 1.0);
        cudaEventRecord(e1); cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1);
        long long c[4]; cudaMemcpy(c, cyc, 32, cudaMemcpyDeviceToHost);
        printf("rep %d: %.3f ms, %d iters, cycles/iter (cta0) %.1f, ns/iter %.1f, err=%s\n", rep, ms, iters, (double)c[0] / iters, ms * 1e6 / iters,
               cudaGetErrorString(cudaGetLastError()));
    }
    return 0;
}
