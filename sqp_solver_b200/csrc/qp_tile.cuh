// Register-tiled fast-path kernel interface (implemented in qp_tile.cu).
#pragma once
#include "qp_common.cuh"

namespace sqpb200 {
bool tile_supported(int n, int m);
bool tile_sliceable(int n, int m, int tile_warps, int f32);  // time slicing (KernelParams::slice_iters) is available for this shape
int tile_slots(int sm_count);                                // resident CTAs of the sliceable configuration
// tile_warps: 0 = default, 4 or 8 selects the warps-per-QP variant of the 64x128 configuration; f32: compute in fp32
// (QPSolver<float>; the arrays in HBM stay fp64)
cudaError_t launch_tile(const KernelParams &p, int sm_count, int ctas_per_sm, int tile_warps, int f32, cudaStream_t stream, char *name,
                        size_t name_len);
}  // namespace sqpb200
