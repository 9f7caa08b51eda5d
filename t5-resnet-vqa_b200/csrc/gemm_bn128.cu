// Instantiations of the tcgen05 GEMM kernel for the 128 x 128 output tile (see gemm_kernel.cuh).
#include "gemm_kernel.cuh"

namespace vqa {
int launch_gemm_bn128(const CUtensorMap& tmA, const CUtensorMap& tmB, const CUtensorMap& tmOut, const CUtensorMap& tmRes,
                     const GemmParams& p, int tiles_m, int tiles_n, int splits, int ctas, cudaStream_t stream) {
  if (ctas == 2) return launch_bn<128, 5, 2>(tmA, tmB, tmOut, tmRes, p, tiles_m, tiles_n, splits, stream);
  return launch_bn<128, 4, 1>(tmA, tmB, tmOut, tmRes, p, tiles_m, tiles_n, splits, stream);
}
}  // namespace vqa
