// Microbenchmarks: fp64 pipe latency/throughput, shuffle, LDS, barrier on sm_100a. Not product code.
#include <cstdio>
#include <cuda_runtime.h>
__global__ void dfma_chain(double *out, long long *cyc, double a, double b, int n) {
    double x = threadIdx.x;
    long long t0 = clock64();
    for (int i = 0; i < n; ++i) {
#pragma unroll
        for (int k = 0; k < 16; ++k) x = fma(x, a, b);
    }
    long long t1 = clock64();
    if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
    out[blockIdx.x * blockDim.x + threadIdx.x] = x;
}
template <int CH>
__global__ void dfma_tput(double *out, long long *cyc, double a, double b, int n) {
    double x[CH];
#pragma unroll
    for (int k = 0; k < CH; ++k) x[k] = threadIdx.x + k;
    long long t0 = clock64();
    for (int i = 0; i < n; ++i) {
#pragma unroll
        for (int r = 0; r < 4; ++r)
#pragma unroll
            for (int k = 0; k < CH; ++k) x[k] = fma(x[k], a, b);
    }
    long long t1 = clock64();
    if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
    double s = 0;
#pragma unroll
    for (int k = 0; k < CH; ++k) s += x[k];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
__global__ void shfl_chain(double *out, long long *cyc, int n) {
    double x = threadIdx.x;
    long long t0 = clock64();
    for (int i = 0; i < n; ++i) {
#pragma unroll
        for (int k = 0; k < 8; ++k) x += __shfl_xor_sync(0xffffffffu, x, 1 + (k & 3));
    }
    long long t1 = clock64();
    if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
    out[blockIdx.x * blockDim.x + threadIdx.x] = x;
}
__global__ void lds_chain(double *out, long long *cyc, int n) {
    __shared__ int idx[1024];
    for (int i = threadIdx.x; i < 1024; i += blockDim.x) idx[i] = (i * 17 + 5) & 1023;
    __syncthreads();
    int x = threadIdx.x;
    long long t0 = clock64();
    for (int i = 0; i < n; ++i) {
#pragma unroll
        for (int k = 0; k < 8; ++k) x = idx[x];
    }
    long long t1 = clock64();
    if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
    out[blockIdx.x * blockDim.x + threadIdx.x] = x;
}
__global__ void bar_chain(double *out, long long *cyc, int n) {
    long long t0 = clock64();
    for (int i = 0; i < n; ++i) {
#pragma unroll
        for (int k = 0; k < 8; ++k) __syncthreads();
    }
    long long t1 = clock64();
    if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
    out[blockIdx.x * blockDim.x + threadIdx.x] = 1;
}
__global__ void sts_bar_lds(double *out, long long *cyc, int n) {  // smem round trip: STS -> BAR -> LDS dependent
    __shared__ double s[256];
    double x = threadIdx.x;
    long long t0 = clock64();
    for (int i = 0; i < n; ++i) {
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            s[threadIdx.x] = x;
            __syncthreads();
            x = s[(threadIdx.x + 33) & 255] + 1.0;
            __syncthreads();
        }
    }
    long long t1 = clock64();
    if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
    out[blockIdx.x * blockDim.x + threadIdx.x] = x;
}
int main() {
    double *out; long long *cyc;
    cudaMalloc(&out, 148 * 1024 * 8 * 4); cudaMalloc(&cyc, 148 * 8 * 4);
    long long c;
    const int n = 1000;
    dfma_chain<<<1, 32>>>(out, cyc, 1.0000001, 1e-9, n); cudaMemcpy(&c, cyc, 8, cudaMemcpyDeviceToHost);
    printf("DFMA dependent latency: %.2f cycles\n", (double)c / (16.0 * n));
    for (int warps = 4; warps <= 32; warps *= 2) {
        dfma_tput<8><<<148, 32 * warps>>>(out, cyc, 1.0000001, 1e-9, n); cudaMemcpy(&c, cyc, 8, cudaMemcpyDeviceToHost);
        printf("DFMA tput %2d warps/SM x8 chains: %.2f DFMA/clk/SM\n", warps, 32.0 * warps * 32.0 * n / c);
    }
    dfma_tput<4><<<148, 256>>>(out, cyc, 1.0000001, 1e-9, n); cudaMemcpy(&c, cyc, 8, cudaMemcpyDeviceToHost);
    printf("DFMA tput 8 warps/SM x4 chains: %.2f DFMA/clk/SM\n", 32.0 * 8 * 16.0 * n / c);
    dfma_tput<2><<<148, 256>>>(out, cyc, 1.0000001, 1e-9, n); cudaMemcpy(&c, cyc, 8, cudaMemcpyDeviceToHost);
    printf("DFMA tput 8 warps/SM x2 chains: %.2f DFMA/clk/SM\n", 32.0 * 8 * 8.0 * n / c);
    dfma_tput<8><<<148, 128>>>(out, cyc, 1.0000001, 1e-9, n); cudaMemcpy(&c, cyc, 8, cudaMemcpyDeviceToHost);
    printf("DFMA tput 4 warps/SM x8 chains: %.2f DFMA/clk/SM\n", 32.0 * 4 * 32.0 * n / c);
    shfl_chain<<<1, 32>>>(out, cyc, n); cudaMemcpy(&c, cyc, 8, cudaMemcpyDeviceToHost);
    printf("fp64 SHFL+DADD dependent: %.2f cycles\n", (double)c / (8.0 * n));
    lds_chain<<<1, 32>>>(out, cyc, n); cudaMemcpy(&c, cyc, 8, cudaMemcpyDeviceToHost);
    printf("LDS dependent: %.2f cycles\n", (double)c / (8.0 * n));
    bar_chain<<<1, 256>>>(out, cyc, n); cudaMemcpy(&c, cyc, 8, cudaMemcpyDeviceToHost);
    printf("BAR.SYNC 256 threads: %.2f cycles\n", (double)c / (8.0 * n));
    bar_chain<<<1, 128>>>(out, cyc, n); cudaMemcpy(&c, cyc, 8, cudaMemcpyDeviceToHost);
    printf("BAR.SYNC 128 threads: %.2f cycles\n", (double)c / (8.0 * n));
    sts_bar_lds<<<1, 256>>>(out, cyc, n); cudaMemcpy(&c, cyc, 8, cudaMemcpyDeviceToHost);
    printf("STS->BAR->LDS->DADD->BAR round trip (256 thr): %.2f cycles\n", (double)c / (4.0 * n));
    printf("%s\n", cudaGetErrorString(cudaDeviceSynchronize()));
    return 0;
}
