// GroupNorm (+ SiLU) over channels-last bf16 activations, and the row softmax used by the VAE attention.
// Both are memory-bound: 16-byte vector accesses, fp32 statistics, second pass served from L2.
#include <cooperative_groups.h>

#include "common.cuh"
#include "host.h"
#include "../../include/idf_b200.h"

namespace idf {

constexpr int GN_MAX_GPS = 32;

__device__ __forceinline__ void unpack8(const uint4& v, float (&f)[8]) {
  f[0] = bf16_lo(v.x); f[1] = bf16_hi(v.x);
  f[2] = bf16_lo(v.y); f[3] = bf16_hi(v.y);
  f[4] = bf16_lo(v.z); f[5] = bf16_hi(v.z);
  f[6] = bf16_lo(v.w); f[7] = bf16_hi(v.w);
}

constexpr double GN_FIX = 268435456.0;  // 2^28
__device__ __forceinline__ unsigned long long to_fixed(float v) {
  return static_cast<unsigned long long>(__double2ll_rn(static_cast<double>(v) * GN_FIX));
}
__device__ __forceinline__ float from_fixed(unsigned long long v) {
  return static_cast<float>(static_cast<double>(static_cast<long long>(v)) * (1.0 / GN_FIX));
}

// grid = (B, slabs). A slab is `gps` consecutive groups = gps*cpg channels = V 16-byte vectors per pixel.
// Thread t owns vector (t % V) of pixels (t / V), (t / V) + rows_per_iter, ...
template <bool SILU>
__global__ void __launch_bounds__(512) groupnorm_kernel(const __nv_bfloat16* __restrict__ x, long long ldx,
                                                        __nv_bfloat16* __restrict__ y, long long ldy,
                                                        const float* __restrict__ gamma,
                                                        const float* __restrict__ beta, int HW, int cpg, int gps,
                                                        int V, float eps) {
  // Group sums are accumulated as Q36.28 fixed point: integer addition is associative, so the shared-memory
  // atomics give bit-identical statistics from run to run (and across batch sizes) in any arrival order.
  __shared__ unsigned long long s_sum[GN_MAX_GPS];
  __shared__ unsigned long long s_sq[GN_MAX_GPS];
  const int b = blockIdx.x;
  const int c0 = blockIdx.y * gps * cpg;
  const int v = threadIdx.x % V;
  const int prow = threadIdx.x / V;
  const int rows_per_iter = blockDim.x / V;
  if (threadIdx.x < GN_MAX_GPS) {
    s_sum[threadIdx.x] = 0ull;
    s_sq[threadIdx.x] = 0ull;
  }
  __syncthreads();

  const __nv_bfloat16* xb = x + (long long)b * HW * ldx + c0 + v * 8;
  float s[8], q[8];
#pragma unroll
  for (int e = 0; e < 8; ++e) { s[e] = 0.f; q[e] = 0.f; }
  int pix = prow;
  for (; pix + 3 * rows_per_iter < HW; pix += 4 * rows_per_iter) {
    uint4 r0 = *reinterpret_cast<const uint4*>(xb + (long long)pix * ldx);
    uint4 r1 = *reinterpret_cast<const uint4*>(xb + (long long)(pix + rows_per_iter) * ldx);
    uint4 r2 = *reinterpret_cast<const uint4*>(xb + (long long)(pix + 2 * rows_per_iter) * ldx);
    uint4 r3 = *reinterpret_cast<const uint4*>(xb + (long long)(pix + 3 * rows_per_iter) * ldx);
    float f[8];
    unpack8(r0, f);
#pragma unroll
    for (int e = 0; e < 8; ++e) { s[e] += f[e]; q[e] = fmaf(f[e], f[e], q[e]); }
    unpack8(r1, f);
#pragma unroll
    for (int e = 0; e < 8; ++e) { s[e] += f[e]; q[e] = fmaf(f[e], f[e], q[e]); }
    unpack8(r2, f);
#pragma unroll
    for (int e = 0; e < 8; ++e) { s[e] += f[e]; q[e] = fmaf(f[e], f[e], q[e]); }
    unpack8(r3, f);
#pragma unroll
    for (int e = 0; e < 8; ++e) { s[e] += f[e]; q[e] = fmaf(f[e], f[e], q[e]); }
  }
  for (; pix < HW; pix += rows_per_iter) {
    uint4 r0 = *reinterpret_cast<const uint4*>(xb + (long long)pix * ldx);
    float f[8];
    unpack8(r0, f);
#pragma unroll
    for (int e = 0; e < 8; ++e) { s[e] += f[e]; q[e] = fmaf(f[e], f[e], q[e]); }
  }
  // fold the 8 channel lanes into their groups, then one shared atomic per (thread, group)
  {
    int g_prev = (v * 8) / cpg;
    float as = 0.f, aq = 0.f;
#pragma unroll
    for (int e = 0; e < 8; ++e) {
      const int g = (v * 8 + e) / cpg;
      if (g != g_prev) {
        atomicAdd(&s_sum[g_prev], to_fixed(as));
        atomicAdd(&s_sq[g_prev], to_fixed(aq));
        as = 0.f; aq = 0.f; g_prev = g;
      }
      as += s[e];
      aq += q[e];
    }
    atomicAdd(&s_sum[g_prev], to_fixed(as));
    atomicAdd(&s_sq[g_prev], to_fixed(aq));
  }
  __syncthreads();

  const float inv_cnt = 1.f / ((float)HW * (float)cpg);
  float sc[8], sh[8];
#pragma unroll
  for (int e = 0; e < 8; ++e) {
    const int cl = v * 8 + e;
    const int g = cl / cpg;
    const float mean = from_fixed(s_sum[g]) * inv_cnt;
    const float var = fmaxf(from_fixed(s_sq[g]) * inv_cnt - mean * mean, 0.f);
    const float rstd = rsqrtf(var + eps);
    const float ga = gamma[c0 + cl], be = beta[c0 + cl];
    sc[e] = rstd * ga;
    sh[e] = be - mean * rstd * ga;
  }
  __nv_bfloat16* yb = y + (long long)b * HW * ldy + c0 + v * 8;
  for (pix = prow; pix < HW; pix += rows_per_iter) {
    uint4 r0 = *reinterpret_cast<const uint4*>(xb + (long long)pix * ldx);
    float f[8];
    unpack8(r0, f);
#pragma unroll
    for (int e = 0; e < 8; ++e) {
      float t = fmaf(f[e], sc[e], sh[e]);
      if (SILU) t = t / (1.f + __expf(-t));
      f[e] = t;
    }
    uint4 o;
    o.x = pack_bf16x2(f[0], f[1]);
    o.y = pack_bf16x2(f[2], f[3]);
    o.z = pack_bf16x2(f[4], f[5]);
    o.w = pack_bf16x2(f[6], f[7]);
    *reinterpret_cast<uint4*>(yb + (long long)pix * ldy) = o;
  }
}


// Single-read variant: the pixels of one (sample, slab) are split over a thread-block cluster. Every CTA keeps its
// pixels in registers, publishes fixed-point partial sums in its shared memory, the cluster exchanges them through
// distributed shared memory, and each CTA normalises its own registers: one HBM read + one write per element.
constexpr int GN_MAXP = 8;
template <bool SILU>
__global__ void __launch_bounds__(512) groupnorm_cluster_kernel(const __nv_bfloat16* __restrict__ x, long long ldx,
                                                                __nv_bfloat16* __restrict__ y, long long ldy,
                                                                const float* __restrict__ gamma,
                                                                const float* __restrict__ beta, int HW, int cpg,
                                                                int gps, int V, float eps, int cs) {
  namespace cg = cooperative_groups;
  cg::cluster_group cluster = cg::this_cluster();
  __shared__ unsigned long long s_sum[GN_MAX_GPS];
  __shared__ unsigned long long s_sq[GN_MAX_GPS];
  const int rank = (int)cluster.block_rank();
  const int b = blockIdx.x / cs;
  const int c0 = blockIdx.y * gps * cpg;
  const int v = threadIdx.x % V;
  const int prow = threadIdx.x / V;
  const int rows_per_iter = blockDim.x / V;
  const int ppc = HW / cs;  // pixels per CTA
  if (threadIdx.x < GN_MAX_GPS) {
    s_sum[threadIdx.x] = 0ull;
    s_sq[threadIdx.x] = 0ull;
  }
  __syncthreads();
  const long long row0 = (long long)b * HW + (long long)rank * ppc;
  const __nv_bfloat16* xb = x + row0 * ldx + c0 + v * 8;
  uint4 held[GN_MAXP];
  float s[8], q[8];
#pragma unroll
  for (int e = 0; e < 8; ++e) { s[e] = 0.f; q[e] = 0.f; }
#pragma unroll
  for (int i = 0; i < GN_MAXP; ++i) {
    const int pix = prow + i * rows_per_iter;
    held[i] = make_uint4(0u, 0u, 0u, 0u);
    if (pix < ppc) held[i] = *reinterpret_cast<const uint4*>(xb + (long long)pix * ldx);
  }
#pragma unroll
  for (int i = 0; i < GN_MAXP; ++i) {
    float f[8];
    unpack8(held[i], f);  // out-of-range slots hold zeros and add nothing
#pragma unroll
    for (int e = 0; e < 8; ++e) { s[e] += f[e]; q[e] = fmaf(f[e], f[e], q[e]); }
  }
  {
    int g_prev = (v * 8) / cpg;
    float as = 0.f, aq = 0.f;
#pragma unroll
    for (int e = 0; e < 8; ++e) {
      const int g = (v * 8 + e) / cpg;
      if (g != g_prev) {
        atomicAdd(&s_sum[g_prev], to_fixed(as));
        atomicAdd(&s_sq[g_prev], to_fixed(aq));
        as = 0.f; aq = 0.f; g_prev = g;
      }
      as += s[e];
      aq += q[e];
    }
    atomicAdd(&s_sum[g_prev], to_fixed(as));
    atomicAdd(&s_sq[g_prev], to_fixed(aq));
  }
  cluster.sync();
  // every CTA sums the cluster's partials itself (integer addition: any order gives the same bits)
  __shared__ unsigned long long t_sum[GN_MAX_GPS];
  __shared__ unsigned long long t_sq[GN_MAX_GPS];
  if (threadIdx.x < gps) {
    unsigned long long a = 0ull, c = 0ull;
    for (int r = 0; r < cs; ++r) {
      a += *cluster.map_shared_rank(&s_sum[threadIdx.x], r);
      c += *cluster.map_shared_rank(&s_sq[threadIdx.x], r);
    }
    t_sum[threadIdx.x] = a;
    t_sq[threadIdx.x] = c;
  }
  cluster.sync();  // nobody may exit (or proceed) while a peer still reads its partials

  const float inv_cnt = 1.f / ((float)HW * (float)cpg);
  float sc[8], sh[8];
#pragma unroll
  for (int e = 0; e < 8; ++e) {
    const int cl = v * 8 + e;
    const int g = cl / cpg;
    const float mean = from_fixed(t_sum[g]) * inv_cnt;
    const float var = fmaxf(from_fixed(t_sq[g]) * inv_cnt - mean * mean, 0.f);
    const float rstd = rsqrtf(var + eps);
    const float ga = gamma[c0 + cl], be = beta[c0 + cl];
    sc[e] = rstd * ga;
    sh[e] = be - mean * rstd * ga;
  }
  __nv_bfloat16* yb = y + row0 * ldy + c0 + v * 8;
#pragma unroll
  for (int i = 0; i < GN_MAXP; ++i) {
    const int pix = prow + i * rows_per_iter;
    if (pix < ppc) {
      float f[8];
      unpack8(held[i], f);
#pragma unroll
      for (int e = 0; e < 8; ++e) {
        float t = fmaf(f[e], sc[e], sh[e]);
        if (SILU) t = t / (1.f + __expf(-t));
        f[e] = t;
      }
      uint4 o;
      o.x = pack_bf16x2(f[0], f[1]);
      o.y = pack_bf16x2(f[2], f[3]);
      o.z = pack_bf16x2(f[4], f[5]);
      o.w = pack_bf16x2(f[6], f[7]);
      *reinterpret_cast<uint4*>(yb + (long long)pix * ldy) = o;
    }
  }
}

// one CTA per row; cols <= 8 * blockDim * 4
__global__ void __launch_bounds__(256) softmax_rows_kernel(const float* __restrict__ in, long long ld_in,
                                                           __nv_bfloat16* __restrict__ out, long long ld_out,
                                                           int cols, float scale) {
  __shared__ float red[8];
  const float* row = in + (long long)blockIdx.x * ld_in;
  __nv_bfloat16* orow = out + (long long)blockIdx.x * ld_out;
  constexpr int PER = 8;
  float v[PER];
  float mx = -INFINITY;
#pragma unroll
  for (int i = 0; i < PER; ++i) {
    const int c = threadIdx.x + i * 256;
    v[i] = c < cols ? row[c] * scale : -INFINITY;
    mx = fmaxf(mx, v[i]);
  }
  for (int o = 16; o > 0; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = mx;
  __syncthreads();
  mx = red[0];
#pragma unroll
  for (int i = 1; i < 8; ++i) mx = fmaxf(mx, red[i]);
  __syncthreads();
  float sum = 0.f;
#pragma unroll
  for (int i = 0; i < PER; ++i) {
    v[i] = (threadIdx.x + i * 256 < cols) ? __expf(v[i] - mx) : 0.f;
    sum += v[i];
  }
  for (int o = 16; o > 0; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = sum;
  __syncthreads();
  sum = 0.f;
#pragma unroll
  for (int i = 0; i < 8; ++i) sum += red[i];
  const float inv = 1.f / sum;
#pragma unroll
  for (int i = 0; i < PER; ++i) {
    const int c = threadIdx.x + i * 256;
    if (c < cols) orow[c] = __float2bfloat16_rn(v[i] * inv);
  }
}

}  // namespace idf

using namespace idf;

extern "C" int idf_groupnorm_silu(const void* x, int64_t ldx, void* y, int64_t ldy, const float* gamma,
                                  const float* beta, int32_t B, int32_t HW, int32_t C, int32_t groups, float eps,
                                  int32_t apply_silu, idf_stream_t stream) {
  if (!x || !y || !gamma || !beta) return fail(IDF_ERR_ARG, "groupnorm: null pointer");
  if (B <= 0 || HW <= 0 || C <= 0 || groups <= 0 || C % groups != 0) return fail(IDF_ERR_ARG, "groupnorm: bad shape");
  const int cpg = C / groups;
  int gps = (groups % 8 == 0) ? 8 : groups;
  if (gps > GN_MAX_GPS) return fail(IDF_ERR_UNSUPPORTED, "groupnorm: %d groups per slab", gps);
  if ((gps * cpg) % 8 != 0) return fail(IDF_ERR_UNSUPPORTED, "groupnorm: slab of %d channels not a multiple of 8", gps * cpg);
  const int V = gps * cpg / 8;
  if (V > 512) return fail(IDF_ERR_UNSUPPORTED, "groupnorm: slab too wide");
  if (ldx % 8 != 0 || ldy % 8 != 0 || (reinterpret_cast<uintptr_t>(x) & 15) || (reinterpret_cast<uintptr_t>(y) & 15))
    return fail(IDF_ERR_ARG, "groupnorm: 16-byte alignment required");
  cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
  const __nv_bfloat16* xp = reinterpret_cast<const __nv_bfloat16*>(x);
  __nv_bfloat16* yp = reinterpret_cast<__nv_bfloat16*>(y);
  int threads = (384 / V) * V;
  if (threads == 0) threads = V;
  {
    // cluster path: split the pixels of a sample over up to 8 CTAs so each thread holds <= GN_MAXP vectors
    int cs = 8;
    while (cs > 1 && (HW % cs != 0 || HW / cs < 16)) cs >>= 1;
    const int ppc = HW / cs;
    int th = threads;
    while (th / V > ppc && th > V) th -= V;
    const int per_thread = (ppc + th / V - 1) / (th / V);
    if (per_thread <= GN_MAXP) {
      cudaLaunchConfig_t cfg = {};
      cfg.gridDim = dim3(B * cs, groups / gps);
      cfg.blockDim = dim3(th);
      cfg.dynamicSmemBytes = 0;
      cfg.stream = s;
      cudaLaunchAttribute attr[1];
      attr[0].id = cudaLaunchAttributeClusterDimension;
      attr[0].val.clusterDim.x = cs;
      attr[0].val.clusterDim.y = 1;
      attr[0].val.clusterDim.z = 1;
      cfg.attrs = attr;
      cfg.numAttrs = 1;
      cudaError_t e;
      if (apply_silu)
        e = cudaLaunchKernelEx(&cfg, groupnorm_cluster_kernel<true>, xp, (long long)ldx, yp, (long long)ldy, gamma, beta,
                               (int)HW, cpg, gps, V, eps, cs);
      else
        e = cudaLaunchKernelEx(&cfg, groupnorm_cluster_kernel<false>, xp, (long long)ldx, yp, (long long)ldy, gamma, beta,
                               (int)HW, cpg, gps, V, eps, cs);
      return check_cuda(e, "groupnorm cluster launch");
    }
  }
  // small images: do not launch more pixel rows than exist
  while (threads / V > HW && threads > V) threads -= V;
  dim3 grid(B, groups / gps);
  if (apply_silu)
    groupnorm_kernel<true><<<grid, threads, 0, s>>>(xp, ldx, yp, ldy, gamma, beta, HW, cpg, gps, V, eps);
  else
    groupnorm_kernel<false><<<grid, threads, 0, s>>>(xp, ldx, yp, ldy, gamma, beta, HW, cpg, gps, V, eps);
  return check_cuda(cudaGetLastError(), "groupnorm launch");
}

extern "C" int idf_softmax_rows(const float* in, int64_t ld_in, void* out, int64_t ld_out, int32_t rows, int32_t cols,
                                float scale, idf_stream_t stream) {
  if (!in || !out) return fail(IDF_ERR_ARG, "softmax_rows: null pointer");
  if (cols <= 0 || cols > 2048) return fail(IDF_ERR_UNSUPPORTED, "softmax_rows: cols = %d > 2048", cols);
  softmax_rows_kernel<<<rows, 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(
      in, ld_in, reinterpret_cast<__nv_bfloat16*>(out), ld_out, cols, scale);
  return check_cuda(cudaGetLastError(), "softmax_rows launch");
}
