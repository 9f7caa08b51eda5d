// Host dispatch of the tcgen05 GEMM / implicit-GEMM kernel (kernel: gemm_kernel.cuh; one translation unit per
// tile width: gemm_bn64.cu, gemm_bn128.cu, gemm_bn256.cu).
#include "gemm.cuh"

namespace vqa {

int launch_gemm_bn64(const CUtensorMap&, const CUtensorMap&, const CUtensorMap&, const CUtensorMap&, const GemmParams&, int, int, int,
                     int, cudaStream_t);
int launch_gemm_bn128(const CUtensorMap&, const CUtensorMap&, const CUtensorMap&, const CUtensorMap&, const GemmParams&, int, int, int,
                     int, cudaStream_t);
int launch_gemm_bn256(const CUtensorMap&, const CUtensorMap&, const CUtensorMap&, const CUtensorMap&, const GemmParams&, int, int, int,
                     int, cudaStream_t);

int gemm_out_tiles(const GemmParams& p, int bn) {
  const int tiles_m = p.a_mode == LOAD_CONV ? p.tiles_w * p.tiles_h * ((p.Nimg + p.bx_n - 1) / p.bx_n) : (p.M + 127) / 128;
  return tiles_m * ((p.N + bn - 1) / bn);
}

int launch_gemm(const CUtensorMap& tmA, const CUtensorMap& tmB, const CUtensorMap& tmOut, const CUtensorMap& tmRes,
                const GemmParams& p, int bn, int split_k, int ctas, cudaStream_t stream) {
  int tiles_m;
  if (p.a_mode == LOAD_CONV) {
    tiles_m = p.tiles_w * p.tiles_h * ((p.Nimg + p.bx_n - 1) / p.bx_n);
  } else {
    tiles_m = (p.M + 127) / 128;
  }
  if (split_k < 1) split_k = 1;
  if (split_k > p.kb_total) split_k = p.kb_total > 0 ? p.kb_total : 1;
  const int tiles_n = (p.N + bn - 1) / bn;
  if (tiles_m <= 0 || tiles_n <= 0) return 0;
  switch (bn) {
    case 64:  return launch_gemm_bn64(tmA, tmB, tmOut, tmRes, p, tiles_m, tiles_n, split_k, ctas, stream);
    case 128: return launch_gemm_bn128(tmA, tmB, tmOut, tmRes, p, tiles_m, tiles_n, split_k, ctas, stream);
    case 256: return launch_gemm_bn256(tmA, tmB, tmOut, tmRes, p, tiles_m, tiles_n, split_k, ctas, stream);
    default:  return static_cast<int>(cudaErrorInvalidValue);
  }
}

}  // namespace vqa
