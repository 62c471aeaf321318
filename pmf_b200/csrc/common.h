// Host-side helpers shared by the C-ABI translation units.
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdarg.h>
#include <stdio.h>

#include "../../include/pmfb.h"

namespace pmfb {

void set_error(const char* fmt, ...);
// Multiprocessor count of the current device (cached; 148 on a B200): every grid is sized from it.
int sm_count();
int fail(int code, const char* fmt, ...);

// cuTensorMapEncodeTiled resolved through cudaGetDriverEntryPoint (no link-time libcuda dependency).
typedef CUresult (*encode_tiled_fn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*,
                                    const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                                    const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                    CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
encode_tiled_fn get_encode_tiled();

// fp32 tensor map with 128B swizzle and zero OOB fill. strides_bytes has rank-1 entries.
// swizzle32b_atom=false: SWIZZLE_128B (16B chunks; K-major UMMA operands);
// swizzle32b_atom=true : SWIZZLE_128B_ATOM_32B (32B chunks; the only layout tcgen05 accepts for
//                        MN-major tf32 operands, i.e. the wgrad kernel).
int make_tmap_f32(CUtensorMap* out, const void* ptr, int rank, const uint64_t* dims,
                  const uint64_t* strides_bytes, const uint32_t* box, bool swizzle32b_atom = false);
// 16-bit (fp16 / bf16) tensor map, 128B swizzle, zero OOB fill: the K-major UMMA operands of the kind::f16 path.
int make_tmap_16(CUtensorMap* out, const void* ptr, int rank, const uint64_t* dims, const uint64_t* strides_bytes,
                 const uint32_t* box, bool bf16, bool swizzle64 = false);

#define PMFB_CUDA_CHECK(expr)                                                            \
  do {                                                                                   \
    cudaError_t _e = (expr);                                                             \
    if (_e != cudaSuccess)                                                               \
      return ::pmfb::fail(PMFB_ERR_CUDA, "%s failed: %s", #expr, cudaGetErrorString(_e)); \
  } while (0)

#define PMFB_LAUNCH_CHECK(name)                                                            \
  do {                                                                                     \
    cudaError_t _e = cudaGetLastError();                                                   \
    if (_e != cudaSuccess)                                                                 \
      return ::pmfb::fail(PMFB_ERR_CUDA, "launch of %s failed: %s", name,                  \
                          cudaGetErrorString(_e));                                         \
  } while (0)

}  // namespace pmfb
