// Error plumbing, device probe and tensor-map creation for libpmf_b200.so.
#include "common.h"

#include <string.h>

namespace pmfb {

static thread_local char g_err[512] = "";

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

int fail(int code, const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
  return code;
}

encode_tiled_fn get_encode_tiled() {
  static encode_tiled_fn fn = nullptr;
  if (fn) return fn;
  void* p = nullptr;
  cudaDriverEntryPointQueryResult q;
  cudaError_t e = cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q);
  if (e != cudaSuccess || q != cudaDriverEntryPointSuccess || !p) {
    cudaGetLastError();
    return nullptr;
  }
  fn = reinterpret_cast<encode_tiled_fn>(p);
  return fn;
}

int make_tmap_f32(CUtensorMap* out, const void* ptr, int rank, const uint64_t* dims,
                  const uint64_t* strides_bytes, const uint32_t* box, bool swizzle32b_atom) {
  encode_tiled_fn enc = get_encode_tiled();
  if (!enc) return fail(PMFB_ERR_NO_DEVICE, "cuTensorMapEncodeTiled not available (no CUDA driver)");
  if ((reinterpret_cast<uintptr_t>(ptr) & 15) != 0)
    return fail(PMFB_ERR_INVALID, "tensor-map base %p not 16-byte aligned", ptr);
  cuuint64_t gdim[5];
  cuuint64_t gstr[4];
  cuuint32_t bx[5];
  cuuint32_t es[5];
  for (int i = 0; i < rank; ++i) {
    gdim[i] = dims[i];
    bx[i] = box[i];
    es[i] = 1;
    if (box[i] == 0 || box[i] > 256) return fail(PMFB_ERR_INVALID, "box[%d]=%u out of range", i, box[i]);
  }
  for (int i = 0; i + 1 < rank; ++i) {
    gstr[i] = strides_bytes[i];
    if (gstr[i] % 16 != 0)
      return fail(PMFB_ERR_INVALID, "tensor-map stride[%d]=%llu not a multiple of 16 bytes", i,
                  (unsigned long long)gstr[i]);
  }
  CUresult r = enc(out, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, (cuuint32_t)rank, const_cast<void*>(ptr),
                   gdim, gstr, bx, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                   swizzle32b_atom ? CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B : CU_TENSOR_MAP_SWIZZLE_128B,
                   CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS)
    return fail(PMFB_ERR_CUDA,
                "cuTensorMapEncodeTiled failed (%d) rank=%d dims=[%llu,%llu,%llu,%llu,%llu] box=[%u,%u,%u,%u,%u]",
                (int)r, rank, (unsigned long long)dims[0], (unsigned long long)(rank > 1 ? dims[1] : 0),
                (unsigned long long)(rank > 2 ? dims[2] : 0), (unsigned long long)(rank > 3 ? dims[3] : 0),
                (unsigned long long)(rank > 4 ? dims[4] : 0), box[0], rank > 1 ? box[1] : 0,
                rank > 2 ? box[2] : 0, rank > 3 ? box[3] : 0, rank > 4 ? box[4] : 0);
  return PMFB_OK;
}

int make_tmap_16(CUtensorMap* out, const void* ptr, int rank, const uint64_t* dims, const uint64_t* strides_bytes,
                 const uint32_t* box, bool bf16, bool swizzle64) {
  encode_tiled_fn enc = get_encode_tiled();
  if (!enc) return fail(PMFB_ERR_NO_DEVICE, "cuTensorMapEncodeTiled not available (no CUDA driver)");
  if ((reinterpret_cast<uintptr_t>(ptr) & 15) != 0) return fail(PMFB_ERR_INVALID, "tensor-map base %p not 16-byte aligned", ptr);
  cuuint64_t gdim[5];
  cuuint64_t gstr[4];
  cuuint32_t bx[5];
  cuuint32_t es[5];
  for (int i = 0; i < rank; ++i) {
    gdim[i] = dims[i];
    bx[i] = box[i];
    es[i] = 1;
    if (box[i] == 0 || box[i] > 256) return fail(PMFB_ERR_INVALID, "box[%d]=%u out of range", i, box[i]);
  }
  for (int i = 0; i + 1 < rank; ++i) {
    gstr[i] = strides_bytes[i];
    if (gstr[i] % 16 != 0)
      return fail(PMFB_ERR_INVALID, "tensor-map stride[%d]=%llu not a multiple of 16 bytes", i, (unsigned long long)gstr[i]);
  }
  CUresult r = enc(out, bf16 ? CU_TENSOR_MAP_DATA_TYPE_BFLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT16, (cuuint32_t)rank,
                   const_cast<void*>(ptr), gdim, gstr, bx, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                   swizzle64 ? CU_TENSOR_MAP_SWIZZLE_64B : CU_TENSOR_MAP_SWIZZLE_128B,
                   CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) return fail(PMFB_ERR_CUDA, "cuTensorMapEncodeTiled (16-bit) failed (%d) rank=%d", (int)r, rank);
  return PMFB_OK;
}

}  // namespace pmfb

namespace pmfb {
int sm_count() {
  static int n = 0;
  if (n == 0) {
    int dev = 0, v = 0;
    if (cudaGetDevice(&dev) == cudaSuccess && cudaDeviceGetAttribute(&v, cudaDevAttrMultiProcessorCount, dev) == cudaSuccess && v > 0) n = v;
    else return 148;
  }
  return n;
}
}  // namespace pmfb

extern "C" {

int pmfb_sm_count(void) { return pmfb::sm_count(); }

int pmfb_abi_version(void) { return PMFB_ABI_VERSION; }

const char* pmfb_last_error(void) { return pmfb::g_err; }

int pmfb_init(void) {
  int n = 0;
  cudaError_t e = cudaGetDeviceCount(&n);
  if (e != cudaSuccess || n == 0) {
    cudaGetLastError();
    return pmfb::fail(PMFB_ERR_NO_DEVICE, "no CUDA device: %s", cudaGetErrorString(e));
  }
  int dev = 0;
  cudaGetDevice(&dev);
  cudaDeviceProp p;
  e = cudaGetDeviceProperties(&p, dev);
  if (e != cudaSuccess) return pmfb::fail(PMFB_ERR_CUDA, "cudaGetDeviceProperties: %s", cudaGetErrorString(e));
  if (p.major != 10)
    return pmfb::fail(PMFB_ERR_NO_DEVICE, "device %d is sm_%d%d; this library is sm_100a only", dev, p.major,
                      p.minor);
  if (!pmfb::get_encode_tiled())
    return pmfb::fail(PMFB_ERR_NO_DEVICE, "cuTensorMapEncodeTiled entry point missing");
  return PMFB_OK;
}

}  // extern "C"
