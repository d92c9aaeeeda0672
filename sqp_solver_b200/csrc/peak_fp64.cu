// fp64 FMA peak of the device, measured in the run that quotes it (bench.py's `roofline.peak`): MEASURED_PEAKS.json has no
// fp64 entry, and the register-tiled kernel is bound by the fp64 pipe and the latencies around it, not by HBM.
// Eight independent DFMA chains per thread, 8 warps per scheduler: enough to saturate the pipe (tools/proto/README.md).
#include <cuda_runtime.h>

#include "../../include/sqp_b200_qp.h"

namespace sqpb200 {

__global__ void __launch_bounds__(256) dfma_peak_kernel(double *out, int iters, double seed) {
    double a0 = seed, a1 = seed + 1, a2 = seed + 2, a3 = seed + 3, a4 = seed + 4, a5 = seed + 5, a6 = seed + 6, a7 = seed + 7;
    const double m = 1.0000001, c = 1e-9 * threadIdx.x;
#pragma unroll 1
    for (int i = 0; i < iters; ++i) {
#pragma unroll
        for (int u = 0; u < 16; ++u) {
            a0 = fma(a0, m, c); a1 = fma(a1, m, c); a2 = fma(a2, m, c); a3 = fma(a3, m, c);
            a4 = fma(a4, m, c); a5 = fma(a5, m, c); a6 = fma(a6, m, c); a7 = fma(a7, m, c);
        }
    }
    const double s = ((a0 + a1) + (a2 + a3)) + ((a4 + a5) + (a6 + a7));
    if (s == 12345.678) out[0] = s;  // never true: keeps the chains alive
}

// runs the kernel `reps` times on `stream`; *seconds = best elapsed time, *fma_count = DFMA instructions (per lane) x lanes of one run
cudaError_t measure_dfma_peak(int sm_count, cudaStream_t stream, double *seconds, double *fma_count) {
    double *out = nullptr;
    cudaError_t e = cudaMalloc(&out, sizeof(double));
    if (e != cudaSuccess) return e;
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    const int iters = 4096, grid = sm_count * 4;
    float best = 1e30f;
    for (int r = 0; r < 5 && e == cudaSuccess; ++r) {
        cudaEventRecord(e0, stream);
        dfma_peak_kernel<<<grid, 256, 0, stream>>>(out, iters, 1.0 + r);
        cudaEventRecord(e1, stream);
        e = cudaEventSynchronize(e1);
        float ms = 0;
        if (e == cudaSuccess) e = cudaEventElapsedTime(&ms, e0, e1);
        if (r > 0 && ms < best) best = ms;  // first run warms up
    }
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    cudaFree(out);
    if (e != cudaSuccess) return e;
    *seconds = best * 1e-3;
    *fma_count = (double)iters * 16 * 8 * 256 * grid;
    return cudaGetLastError();
}

}  // namespace sqpb200
