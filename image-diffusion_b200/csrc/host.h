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

// IDF_PDL=1 enables programmatic dependent launch. Default off: measured no gain on the CUDA-graph replay of the
// sampling step (4.28 vs 4.25 ms), the kernel-to-kernel gaps inside the graph are already short.
bool pdl_enabled();

// Launch with the programmatic-stream-serialization attribute: only for kernels that call pdl_wait() before their
// first global-memory access.
template <typename... KArgs, typename... Args>
inline cudaError_t launch_pdl(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t stream,
                              Args&&... args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid;
  cfg.blockDim = block;
  cfg.dynamicSmemBytes = smem;
  cfg.stream = stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = pdl_enabled() ? 1 : 0;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  return cudaLaunchKernelEx(&cfg, kernel, static_cast<KArgs>(args)...);
}

}  // namespace idf
