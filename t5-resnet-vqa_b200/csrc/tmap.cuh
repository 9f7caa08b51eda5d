// Host-side CUtensorMap construction (cuTensorMapEncodeTiled fetched through the runtime's driver
// entry point, so the library has no link-time dependency on libcuda).
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>

namespace vqa {

// bf16, 128-byte swizzle, zero fill for out-of-bounds.  dims/box are innermost-first; strides_bytes
// has rank-1 entries (stride of dims 1..rank-1).  Returns 0 on success, else a CUresult/-1.
int make_tmap_bf16(CUtensorMap* out, const void* base, int rank, const uint64_t* dims,
                   const uint64_t* strides_bytes, const uint32_t* box, const uint32_t* elem_strides);

// Same for fp32 elements (epilogue output tiles of fp32 tensors: 32 columns = one 128-byte swizzle row).
int make_tmap_f32(CUtensorMap* out, const void* base, int rank, const uint64_t* dims,
                  const uint64_t* strides_bytes, const uint32_t* box, const uint32_t* elem_strides);

// Row-major [rows, cols] bf16 matrix with row stride ld (elements); box = (box_cols, box_rows).
int make_tmap_2d(CUtensorMap* out, const void* base, uint64_t rows, uint64_t cols, uint64_t ld,
                 uint32_t box_cols, uint32_t box_rows);

void set_last_error(const char* fmt, ...);
const char* get_last_error();

}  // namespace vqa
