// Sparse-A input format: the reference's intended sparse variant stores A as an
// Eigen::SparseMatrix (compressed column storage; include/solvers/qp.hpp:22-25, include/unsupported/qp_solver.hpp:363-394),
// BASELINE.json config 5 names CSR. Both are accepted here with ONE sparsity pattern shared by the batch (the Jacobian
// pattern of a batch of same-structure NLPs) and per-instance values. This file is the densify step for the shapes whose kernels keep A
// in registers anyway (n <= 64, m <= 128: thread-per-QP and register-tiled kernels); 64 < n <= 256 stay compressed in the cluster
// kernel (qp_cluster.cu) or the blocked kernel (qp_block.cu).
#include "qp_common.cuh"

namespace sqpb200 {

// dst[b][row + m*col] = vals[b][p] for every stored entry p of the shared pattern; dst is zero-filled beforehand.
__global__ void densify_kernel(const double *__restrict__ vals, const int *__restrict__ outer, const int *__restrict__ inner,
                               int n_outer, int nnz, int m, int n, int csr, int count, double *__restrict__ dst) {
    const size_t total = (size_t)count * nnz;
    for (size_t e = blockIdx.x * (size_t)blockDim.x + threadIdx.x; e < total; e += (size_t)gridDim.x * blockDim.x) {
        const int b = (int)(e / nnz), p = (int)(e % nnz);
        // outer index of entry p: last o with outer[o] <= p (binary search; the pattern is tiny and cache resident)
        int lo = 0, hi = n_outer;
        while (hi - lo > 1) {
            const int mid = (lo + hi) >> 1;
            if (outer[mid] <= p) lo = mid;
            else hi = mid;
        }
        const int o = lo, i = inner[p];
        const int row = csr ? o : i, col = csr ? i : o;
        if (row < m && col < n) dst[(size_t)b * m * n + row + (size_t)m * col] = vals[e];
    }
}

cudaError_t launch_densify(const double *vals, const int *outer, const int *inner, int nnz, int m, int n, int csr, int count,
                           double *dst, cudaStream_t stream) {
    cudaError_t e = cudaMemsetAsync(dst, 0, sizeof(double) * (size_t)count * m * n, stream);
    if (e != cudaSuccess || nnz == 0) return e;
    const size_t total = (size_t)count * nnz;
    int grid = (int)((total + 255) / 256);
    if (grid > 148 * 16) grid = 148 * 16;
    densify_kernel<<<grid, 256, 0, stream>>>(vals, outer, inner, csr ? m : n, nnz, m, n, csr, count, dst);
    return cudaGetLastError();
}

}  // namespace sqpb200
