#include "qp_tile.cuh"
namespace sqpb200 {
bool tile_supported(int, int) { return false; }
cudaError_t launch_tile(const KernelParams &, int, int, cudaStream_t, char *, size_t) { return cudaErrorNotSupported; }
}  // namespace sqpb200
