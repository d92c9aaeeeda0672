// Register-tiled fast-path kernel interface (implemented in qp_tile.cu).
#pragma once
#include "qp_common.cuh"

namespace sqpb200 {
bool tile_supported(int n, int m);
cudaError_t launch_tile(const KernelParams &p, int sm_count, int ctas_per_sm, cudaStream_t stream, char *name, size_t name_len);
}  // namespace sqpb200
