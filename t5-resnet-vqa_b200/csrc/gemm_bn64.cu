// Instantiations of the tcgen05 GEMM kernel for the 128 x 64 output tile (see gemm_kernel.cuh).
#include "gemm_kernel.cuh"

namespace vqa {
int launch_gemm_bn64(const CUtensorMap& tmA, const CUtensorMap& tmB, const CUtensorMap& tmOut, const CUtensorMap& tmRes,
                     const GemmParams& p, int tiles_m, int tiles_n, int splits, int ctas, cudaStream_t stream) {
  if (ctas == 2) return -3;   // a CTA pair needs a tile at least 128 columns wide
  return launch_bn<64, 6, 1>(tmA, tmB, tmOut, tmRes, p, tiles_m, tiles_n, splits, stream);
}
}  // namespace vqa
