// Host-side helpers shared by the C-ABI entry points: error reporting and TMA descriptor encoding.
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>

namespace idf {

enum : int {
  IDF_OK = 0,
  IDF_ERR_ARG = 1,      // shape / alignment / dtype violation
  IDF_ERR_CUDA = 2,     // CUDA runtime or driver error
  IDF_ERR_UNSUPPORTED = 3,
};

// Records a thread-local message retrievable through idf_last_error(); returns `code`.
int fail(int code, const char* fmt, ...);
int check_cuda(cudaError_t e, const char* what);

// Encode a tiled bf16/f32 tensor map. dims[0] is the contiguous dimension. strides_bytes has rank-1 entries
// (dimension 0 is dense). Returns IDF_OK or an error code with the message set.
int encode_tmap(CUtensorMap* out, CUtensorMapDataType dtype, const void* ptr, int rank, const uint64_t* dims,
                const uint64_t* strides_bytes, const uint32_t* box, CUtensorMapSwizzle swizzle,
                const uint32_t* elem_strides = nullptr);

int sm_count();

// cudaLaunchKernelEx wrapper: typed arguments, error returned instead of left in the runtime's sticky state.
template <typename... KArgs, typename... Args>
inline cudaError_t launch_kernel(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t stream,
                                 Args&&... args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid;
  cfg.blockDim = block;
  cfg.dynamicSmemBytes = smem;
  cfg.stream = stream;
  return cudaLaunchKernelEx(&cfg, kernel, static_cast<KArgs>(args)...);
}

}  // namespace idf
