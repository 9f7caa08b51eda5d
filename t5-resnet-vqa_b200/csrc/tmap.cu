#include "tmap.cuh"
#include "common.cuh"

#include <cudaTypedefs.h>
#include <stdarg.h>
#include <stdio.h>
#include <stdlib.h>

namespace vqa {

namespace {
thread_local char g_err[512] = "";
PFN_cuTensorMapEncodeTiled_v12000 g_encode = nullptr;

int load_encode() {
  if (g_encode) return 0;
  void* fn = nullptr;
  cudaDriverEntryPointQueryResult qres;
  cudaError_t e = cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres);
  if (e != cudaSuccess || qres != cudaDriverEntryPointSuccess || fn == nullptr) {
    set_last_error("cuTensorMapEncodeTiled unavailable (cudaGetDriverEntryPoint: %s)",
                   cudaGetErrorString(e));
    return -1;
  }
  g_encode = reinterpret_cast<PFN_cuTensorMapEncodeTiled_v12000>(fn);
  return 0;
}
}  // namespace

bool pdl_enabled() {
  static int v = -1;
  if (v < 0) {
    const char* e = getenv("VQA_B200_PDL");
    v = (e != nullptr && e[0] == '0') ? 0 : 1;
  }
  return v == 1;
}

void set_last_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}
const char* get_last_error() { return g_err; }

static int make_tmap_any(CUtensorMap* out, CUtensorMapDataType dtype, const void* base, int rank, const uint64_t* dims,
                         const uint64_t* strides_bytes, const uint32_t* box, const uint32_t* elem_strides);

int make_tmap_bf16(CUtensorMap* out, const void* base, int rank, const uint64_t* dims,
                   const uint64_t* strides_bytes, const uint32_t* box, const uint32_t* elem_strides) {
  return make_tmap_any(out, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, base, rank, dims, strides_bytes, box, elem_strides);
}
int make_tmap_f32(CUtensorMap* out, const void* base, int rank, const uint64_t* dims,
                  const uint64_t* strides_bytes, const uint32_t* box, const uint32_t* elem_strides) {
  return make_tmap_any(out, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, base, rank, dims, strides_bytes, box, elem_strides);
}

static int make_tmap_any(CUtensorMap* out, CUtensorMapDataType dtype, const void* base, int rank, const uint64_t* dims,
                         const uint64_t* strides_bytes, const uint32_t* box, const uint32_t* elem_strides) {
  if (load_encode() != 0) return -1;
  cuuint64_t gdim[5];
  cuuint64_t gstr[4];
  cuuint32_t bdim[5];
  cuuint32_t estr[5];
  for (int i = 0; i < rank; ++i) {
    gdim[i] = dims[i];
    bdim[i] = box[i];
    estr[i] = elem_strides ? elem_strides[i] : 1;
  }
  for (int i = 0; i + 1 < rank; ++i) gstr[i] = strides_bytes[i];
  CUresult r = g_encode(out, dtype, static_cast<cuuint32_t>(rank),
                        const_cast<void*>(base), gdim, gstr, bdim, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                        CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                        CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    set_last_error(
        "cuTensorMapEncodeTiled failed: CUresult %d (rank %d base %p dims %llu,%llu,%llu,%llu box "
        "%u,%u,%u,%u stride0 %llu)",
        static_cast<int>(r), rank, base, (unsigned long long)gdim[0],
        (unsigned long long)(rank > 1 ? gdim[1] : 0), (unsigned long long)(rank > 2 ? gdim[2] : 0),
        (unsigned long long)(rank > 3 ? gdim[3] : 0), bdim[0], rank > 1 ? bdim[1] : 0,
        rank > 2 ? bdim[2] : 0, rank > 3 ? bdim[3] : 0, (unsigned long long)(rank > 1 ? gstr[0] : 0));
    return static_cast<int>(r);
  }
  return 0;
}

int make_tmap_2d(CUtensorMap* out, const void* base, uint64_t rows, uint64_t cols, uint64_t ld,
                 uint32_t box_cols, uint32_t box_rows) {
  const uint64_t dims[2] = {cols, rows};
  const uint64_t strides[1] = {ld * 2};
  const uint32_t box[2] = {box_cols, box_rows};
  return make_tmap_bf16(out, base, 2, dims, strides, box, nullptr);
}

}  // namespace vqa
