#include "host.h"

#include <stdarg.h>
#include <stdlib.h>
#include <stdio.h>
#include <string.h>

#include <mutex>

#include "../../include/idf_b200.h"

namespace idf {

static thread_local char g_err[512] = "";

int fail(int code, const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
  return code;
}

int check_cuda(cudaError_t e, const char* what) {
  if (e == cudaSuccess) return IDF_OK;
  return fail(IDF_ERR_CUDA, "%s: %s", what, cudaGetErrorString(e));
}

typedef CUresult (*encode_tiled_fn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                    const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                    CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static encode_tiled_fn g_encode = nullptr;
static std::once_flag g_encode_once;

static void resolve_encode() {
  void* fn = nullptr;
  cudaDriverEntryPointQueryResult q;
  cudaError_t e = cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q);
  if (e == cudaSuccess && q == cudaDriverEntryPointSuccess) g_encode = reinterpret_cast<encode_tiled_fn>(fn);
}

int encode_tmap(CUtensorMap* out, CUtensorMapDataType dtype, const void* ptr, int rank, const uint64_t* dims,
                const uint64_t* strides_bytes, const uint32_t* box, CUtensorMapSwizzle swizzle,
                const uint32_t* elem_strides) {
  std::call_once(g_encode_once, resolve_encode);
  if (!g_encode) return fail(IDF_ERR_CUDA, "cuTensorMapEncodeTiled is not available from this driver");
  if ((reinterpret_cast<uintptr_t>(ptr) & 15) != 0) return fail(IDF_ERR_ARG, "TMA base pointer not 16-byte aligned");
  cuuint64_t gdim[5];
  cuuint64_t gstr[4];
  cuuint32_t gbox[5];
  cuuint32_t estr[5];
  for (int i = 0; i < rank; ++i) {
    gdim[i] = dims[i];
    gbox[i] = box[i];
    estr[i] = elem_strides ? elem_strides[i] : 1;
    if (box[i] == 0 || box[i] > 256) return fail(IDF_ERR_ARG, "TMA box dim %d = %u out of range", i, box[i]);
  }
  for (int i = 0; i + 1 < rank; ++i) {
    gstr[i] = strides_bytes[i];
    if (strides_bytes[i] % 16 != 0) return fail(IDF_ERR_ARG, "TMA stride %d = %llu not a multiple of 16 bytes", i,
                                                (unsigned long long)strides_bytes[i]);
  }
  CUresult r = g_encode(out, dtype, (cuuint32_t)rank, const_cast<void*>(ptr), gdim, gstr, gbox, estr,
                        CU_TENSOR_MAP_INTERLEAVE_NONE, swizzle, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                        CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) return fail(IDF_ERR_CUDA, "cuTensorMapEncodeTiled failed with CUresult %d", (int)r);
  return IDF_OK;
}

int sm_count() {
  static int n = 0;
  if (n == 0) {
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev);
  }
  return n;
}

}  // namespace idf

extern "C" const char* idf_last_error(void) { return idf::g_err; }

extern "C" int idf_abi_version(void) { return IDF_B200_ABI_VERSION; }

extern "C" int idf_struct_size(int which) {
  switch (which) {
    case 0: return (int)sizeof(idf_nhwc_t);
    case 1: return (int)sizeof(idf_igemm_args);
    case 2: return (int)sizeof(idf_wgrad_args);
    case 3: return (int)sizeof(idf_pack_job);
    default: return -1;
  }
}
