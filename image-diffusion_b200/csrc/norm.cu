// GroupNorm (+ SiLU) over channels-last bf16 activations, and the row softmax used by the VAE attention.
// Both are memory-bound: 16-byte vector accesses, fp32 statistics, second pass served from L2.
#include <stdlib.h>

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

// ---------------------------------------------------------------------------------------------------------------
// Standalone GroupNorm(+SiLU). grid = (B, slabs); a slab is `gps` consecutive groups = V (<= VP) 16-byte vectors
// per pixel. Lane l of a warp owns vector (l % VP) of pixel row (l / VP), so a warp reads 32/VP whole pixel slabs
// per sweep (coalesced) and lanes holding the same vector are reduced with a fixed xor-shuffle tree. Per-warp
// per-channel partials are then summed over warps and over the channels of a group in a fixed order: the
// statistics are bit-identical from run to run and independent of the batch size (no atomics anywhere).
// Pass 2 re-reads the slab (L2) and writes the normalised, activated bf16 output.
// ---------------------------------------------------------------------------------------------------------------
constexpr int GN_MAX_WARPS = 12;
constexpr int GN_UNROLL = 8;  // independent 16-byte loads in flight per thread

// silu(t) = t * sigmoid(t) = h + h * tanh(h) with h = t / 2: one MUFU op instead of two (ex2 + rcp); the
// approximation error (~2^-11 relative) is far below the bf16 rounding of the stored result.
__device__ __forceinline__ float silu_tanh(float t) {
  const float h = 0.5f * t;
  float th;
  asm("tanh.approx.f32 %0, %1;" : "=f"(th) : "f"(h));
  return fmaf(h, th, h);
}

template <bool SILU, int VP>
__global__ void __launch_bounds__(GN_MAX_WARPS * 32) groupnorm_kernel(const __nv_bfloat16* __restrict__ x,
                                                                      long long ldx, __nv_bfloat16* __restrict__ y,
                                                                      long long ldy, const float* __restrict__ gamma,
                                                                      const float* __restrict__ beta, int HW, int cpg,
                                                                      int gps, int V, float eps,
                                                                      float* __restrict__ stats, int groups) {
  __shared__ float ch_s[GN_MAX_WARPS][VP * 8];
  __shared__ float ch_q[GN_MAX_WARPS][VP * 8];
  __shared__ float ct_s[VP * 8];
  __shared__ float ct_q[VP * 8];
  __shared__ float g_mean[GN_MAX_GPS];
  __shared__ float g_rstd[GN_MAX_GPS];
  constexpr int RPW = 32 / VP;  // pixel rows per warp per sweep
  const int b = blockIdx.x;
  const int c0 = blockIdx.y * gps * cpg;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarps = blockDim.x >> 5;
  const int v = lane % VP, prl = lane / VP;
  const bool active = v < V;
  const int rows_per_iter = nwarps * RPW;

  const __nv_bfloat16* xb = x + (long long)b * HW * ldx + c0 + v * 8;
  float s[8], q[8];
#pragma unroll
  for (int e = 0; e < 8; ++e) { s[e] = 0.f; q[e] = 0.f; }
  if (active) {
    int pix = warp * RPW + prl;
    for (; pix + (GN_UNROLL - 1) * rows_per_iter < HW; pix += GN_UNROLL * rows_per_iter) {
      uint4 r[GN_UNROLL];
#pragma unroll
      for (int i = 0; i < GN_UNROLL; ++i) r[i] = *reinterpret_cast<const uint4*>(xb + (long long)(pix + i * rows_per_iter) * ldx);
#pragma unroll
      for (int i = 0; i < GN_UNROLL; ++i) {
        float f[8];
        unpack8(r[i], f);
#pragma unroll
        for (int e = 0; e < 8; ++e) { s[e] += f[e]; q[e] = fmaf(f[e], f[e], q[e]); }
      }
    }
    for (; pix < HW; pix += rows_per_iter) {
      const uint4 r0 = *reinterpret_cast<const uint4*>(xb + (long long)pix * ldx);
      float f[8];
      unpack8(r0, f);
#pragma unroll
      for (int e = 0; e < 8; ++e) { s[e] += f[e]; q[e] = fmaf(f[e], f[e], q[e]); }
    }
  }
#pragma unroll
  for (int off = VP; off < 32; off <<= 1) {
#pragma unroll
    for (int e = 0; e < 8; ++e) {
      s[e] += __shfl_xor_sync(0xffffffffu, s[e], off);
      q[e] += __shfl_xor_sync(0xffffffffu, q[e], off);
    }
  }
  if (prl == 0) {
#pragma unroll
    for (int e = 0; e < 8; ++e) { ch_s[warp][v * 8 + e] = s[e]; ch_q[warp][v * 8 + e] = q[e]; }
  }
  __syncthreads();
  if (threadIdx.x < V * 8) {
    float a = 0.f, c = 0.f;
    for (int w = 0; w < nwarps; ++w) { a += ch_s[w][threadIdx.x]; c += ch_q[w][threadIdx.x]; }
    ct_s[threadIdx.x] = a;
    ct_q[threadIdx.x] = c;
  }
  __syncthreads();
  if (threadIdx.x < gps) {
    float a = 0.f, c = 0.f;
    for (int k = 0; k < cpg; ++k) { a += ct_s[threadIdx.x * cpg + k]; c += ct_q[threadIdx.x * cpg + k]; }
    const float inv_cnt = 1.f / ((float)HW * (float)cpg);
    const float mean = a * inv_cnt;
    const float var = fmaxf(c * inv_cnt - mean * mean, 0.f);
    const float rstd = rsqrtf(var + eps);
    g_mean[threadIdx.x] = mean;
    g_rstd[threadIdx.x] = rstd;
    if (stats != nullptr) {  // training: (mean, rstd) per (sample, group) for the backward pass
      float* st = stats + ((long long)b * groups + blockIdx.y * gps + threadIdx.x) * 2;
      st[0] = mean;
      st[1] = rstd;
    }
  }
  __syncthreads();
  if (!active) return;

  float sc[8], sh[8];
#pragma unroll
  for (int e = 0; e < 8; ++e) {
    const int cl = v * 8 + e;
    const int g = cl / cpg;
    const float ga = gamma[c0 + cl], be = beta[c0 + cl];
    sc[e] = g_rstd[g] * ga;
    sh[e] = be - g_mean[g] * g_rstd[g] * ga;
  }
  __nv_bfloat16* yb = y + (long long)b * HW * ldy + c0 + v * 8;
  auto apply_store = [&](const uint4& r0, int pix) {
    float f[8];
    unpack8(r0, f);
#pragma unroll
    for (int e = 0; e < 8; ++e) {
      float t = fmaf(f[e], sc[e], sh[e]);
      if (SILU) t = silu_tanh(t);
      f[e] = t;
    }
    uint4 o;
    o.x = pack_bf16x2(f[0], f[1]);
    o.y = pack_bf16x2(f[2], f[3]);
    o.z = pack_bf16x2(f[4], f[5]);
    o.w = pack_bf16x2(f[6], f[7]);
    *reinterpret_cast<uint4*>(yb + (long long)pix * ldy) = o;
  };
  // four independent 16-byte loads in flight per thread (a one-load-per-iteration loop is latency-bound: ncu showed
  // 26-31 % DRAM and 34 % SM throughput with every CTA resident for the whole kernel)
  int pix = warp * RPW + prl;
  for (; pix + (GN_UNROLL - 1) * rows_per_iter < HW; pix += GN_UNROLL * rows_per_iter) {
    uint4 r[GN_UNROLL];
#pragma unroll
    for (int i = 0; i < GN_UNROLL; ++i) r[i] = *reinterpret_cast<const uint4*>(xb + (long long)(pix + i * rows_per_iter) * ldx);
#pragma unroll
    for (int i = 0; i < GN_UNROLL; ++i) apply_store(r[i], pix + i * rows_per_iter);
  }
  for (; pix < HW; pix += rows_per_iter) apply_store(*reinterpret_cast<const uint4*>(xb + (long long)pix * ldx), pix);
}

// ---------------------------------------------------------------------------------------------------------------
// Large images (the VAE at 64x64 .. 128x128 pixels: 1 .. 8 MB per sample). The slab kernel above would hand each CTA
// a 32-byte piece of every pixel row and re-read its 0.5 MB slab from HBM (hundreds of slabs in flight overflow the
// 126 MB L2): measured 24 % of HBM peak over the VQ-VAE's 40 GroupNorms. Here a CTA owns a CHUNK OF WHOLE PIXEL ROWS
// (fully coalesced 16-byte vectors):
//   gn_rows_stats_kernel   per-(sample, chunk, group) partial sum / sum of squares (fp32, fixed order),
//   gn_rows_apply_kernel   sums the sample's partials in chunk order, normalises (+ SiLU), writes bf16.
// The host launches the pair for a few samples at a time, sized so that the group's input is still in L2 when the
// apply kernel reads it again: HBM traffic = one read + one write. No atomics: bit-deterministic, batch invariant.
// Thread t: vector v = t % V of row t / V (V = C / 8 vectors per row, R = blockDim / V rows per sweep).
// ---------------------------------------------------------------------------------------------------------------
constexpr int GNR_THREADS = 256;
constexpr int GNR_UNROLL = 4;

__global__ void __launch_bounds__(GNR_THREADS) gn_rows_stats_kernel(const __nv_bfloat16* __restrict__ x, long long ldx,
                                                                    int HW, int C, int cpg, int rows_per_chunk,
                                                                    float* __restrict__ part, int groups) {
  __shared__ float red_s[GNR_THREADS * 8];
  __shared__ float red_q[GNR_THREADS * 8];
  __shared__ float ch_s[512], ch_q[512];
  const int V = C / 8, R = blockDim.x / V;
  const int v = threadIdx.x % V, r = threadIdx.x / V;
  const int chunk = blockIdx.x, b = blockIdx.y, chunks = gridDim.x;
  const int row0 = chunk * rows_per_chunk;
  const int row1 = min(row0 + rows_per_chunk, HW);
  const __nv_bfloat16* xb = x + ((long long)b * HW) * ldx + v * 8;
  float s[8], q[8];
#pragma unroll
  for (int e = 0; e < 8; ++e) { s[e] = 0.f; q[e] = 0.f; }
  int row = row0 + r;
  for (; row + (GNR_UNROLL - 1) * R < row1; row += GNR_UNROLL * R) {
    uint4 t[GNR_UNROLL];
#pragma unroll
    for (int i = 0; i < GNR_UNROLL; ++i) t[i] = *reinterpret_cast<const uint4*>(xb + (long long)(row + i * R) * ldx);
#pragma unroll
    for (int i = 0; i < GNR_UNROLL; ++i) {
      float f[8];
      unpack8(t[i], f);
#pragma unroll
      for (int e = 0; e < 8; ++e) { s[e] += f[e]; q[e] = fmaf(f[e], f[e], q[e]); }
    }
  }
  for (; row < row1; row += R) {
    float f[8];
    unpack8(*reinterpret_cast<const uint4*>(xb + (long long)row * ldx), f);
#pragma unroll
    for (int e = 0; e < 8; ++e) { s[e] += f[e]; q[e] = fmaf(f[e], f[e], q[e]); }
  }
#pragma unroll
  for (int e = 0; e < 8; ++e) { red_s[r * C + v * 8 + e] = s[e]; red_q[r * C + v * 8 + e] = q[e]; }
  __syncthreads();
  for (int c = threadIdx.x; c < C; c += blockDim.x) {
    float a = 0.f, d = 0.f;
    for (int k = 0; k < R; ++k) { a += red_s[k * C + c]; d += red_q[k * C + c]; }
    ch_s[c] = a; ch_q[c] = d;
  }
  __syncthreads();
  if (threadIdx.x < groups) {
    float a = 0.f, d = 0.f;
    for (int k = 0; k < cpg; ++k) { a += ch_s[threadIdx.x * cpg + k]; d += ch_q[threadIdx.x * cpg + k]; }
    float* dst = part + (((long long)b * chunks + chunk) * groups + threadIdx.x) * 2;
    dst[0] = a; dst[1] = d;
  }
}

template <bool SILU>
__global__ void __launch_bounds__(GNR_THREADS) gn_rows_apply_kernel(const __nv_bfloat16* __restrict__ x, long long ldx,
                                                                    __nv_bfloat16* __restrict__ y, long long ldy,
                                                                    const float* __restrict__ gamma,
                                                                    const float* __restrict__ beta, int HW, int C,
                                                                    int cpg, int rows_per_chunk, float eps,
                                                                    const float* __restrict__ part, int groups) {
  __shared__ float g_mean[GN_MAX_GPS * 2], g_rstd[GN_MAX_GPS * 2];
  const int V = C / 8, R = blockDim.x / V;
  const int v = threadIdx.x % V, r = threadIdx.x / V;
  const int chunk = blockIdx.x, b = blockIdx.y, chunks = gridDim.x;
  float ga[8], be[8];
#pragma unroll
  for (int e = 0; e < 8; ++e) { ga[e] = gamma[v * 8 + e]; be[e] = beta[v * 8 + e]; }
  if (threadIdx.x < groups) {
    const float* src = part + ((long long)b * chunks * groups + threadIdx.x) * 2;
    float a = 0.f, d = 0.f;
    for (int k = 0; k < chunks; ++k) { a += src[(long long)k * groups * 2]; d += src[(long long)k * groups * 2 + 1]; }
    const float inv_cnt = 1.f / ((float)HW * (float)cpg);
    const float mean = a * inv_cnt;
    const float var = fmaxf(d * inv_cnt - mean * mean, 0.f);
    g_mean[threadIdx.x] = mean;
    g_rstd[threadIdx.x] = rsqrtf(var + eps);
  }
  __syncthreads();
  float sc[8], sh[8];
#pragma unroll
  for (int e = 0; e < 8; ++e) {
    const int g = (v * 8 + e) / cpg;
    sc[e] = g_rstd[g] * ga[e];
    sh[e] = be[e] - g_mean[g] * g_rstd[g] * ga[e];
  }
  const int row0 = chunk * rows_per_chunk;
  const int row1 = min(row0 + rows_per_chunk, HW);
  const __nv_bfloat16* xb = x + ((long long)b * HW) * ldx + v * 8;
  __nv_bfloat16* yb = y + ((long long)b * HW) * ldy + v * 8;
  auto apply_store = [&](const uint4& t, int row) {
    float f[8];
    unpack8(t, f);
#pragma unroll
    for (int e = 0; e < 8; ++e) {
      float u = fmaf(f[e], sc[e], sh[e]);
      if (SILU) u = silu_tanh(u);
      f[e] = u;
    }
    uint4 o;
    o.x = pack_bf16x2(f[0], f[1]);
    o.y = pack_bf16x2(f[2], f[3]);
    o.z = pack_bf16x2(f[4], f[5]);
    o.w = pack_bf16x2(f[6], f[7]);
    *reinterpret_cast<uint4*>(yb + (long long)row * ldy) = o;
  };
  int row = row0 + r;
  for (; row + (GNR_UNROLL - 1) * R < row1; row += GNR_UNROLL * R) {
    uint4 t[GNR_UNROLL];
#pragma unroll
    for (int i = 0; i < GNR_UNROLL; ++i) t[i] = *reinterpret_cast<const uint4*>(xb + (long long)(row + i * R) * ldx);
#pragma unroll
    for (int i = 0; i < GNR_UNROLL; ++i) apply_store(t[i], row + i * R);
  }
  for (; row < row1; row += R) apply_store(*reinterpret_cast<const uint4*>(xb + (long long)row * ldx), row);
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

static int groupnorm_impl(const void* x, int64_t ldx, void* y, int64_t ldy, const float* gamma, const float* beta,
                          int32_t B, int32_t HW, int32_t C, int32_t groups, float eps, int32_t apply_silu,
                          float* stats, idf_stream_t stream) {
  if (!x || !y || !gamma || !beta) return fail(IDF_ERR_ARG, "groupnorm: null pointer");
  if (B <= 0 || HW <= 0 || C <= 0 || groups <= 0 || C % groups != 0) return fail(IDF_ERR_ARG, "groupnorm: bad shape");
  const int cpg = C / groups;
  // groups per slab: as many as possible (<= 8) while a slab stays <= 8 sixteen-byte vectors wide
  // ... and small enough that a CTA's slab (HW x V x 16 bytes) stays L1-resident between the two passes
  static const int l1_kb = [] { const char* e = getenv("IDF_GN_SLAB_KB"); return e ? atoi(e) : 512; }();  // (64 KiB slabs measured no faster)
  int vmax = 8;
  while (vmax > 1 && (long long)HW * vmax * 16 > (long long)l1_kb * 1024) vmax >>= 1;
  int gps = 0;
  for (int d = 8; d >= 1; --d)
    if (groups % d == 0 && (d * cpg) % 8 == 0 && d * cpg / 8 <= vmax) { gps = d; break; }
  if (gps == 0)
    for (int d = 1; d <= 8; ++d)  // no slab that small: take the narrowest legal one
      if (groups % d == 0 && (d * cpg) % 8 == 0 && d * cpg / 8 <= 8) { gps = d; break; }
  if (gps == 0) return fail(IDF_ERR_UNSUPPORTED, "groupnorm: %d channels in %d groups does not split into slabs", C, groups);
  if (ldx % 8 != 0 || ldy % 8 != 0 || (reinterpret_cast<uintptr_t>(x) & 15) || (reinterpret_cast<uintptr_t>(y) & 15))
    return fail(IDF_ERR_ARG, "groupnorm: 16-byte alignment required");
  cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
  const __nv_bfloat16* xp = reinterpret_cast<const __nv_bfloat16*>(x);
  __nv_bfloat16* yp = reinterpret_cast<__nv_bfloat16*>(y);

  const int V = gps * cpg / 8;
  const int VP = V <= 4 ? 4 : 8;
  // CTA size by slab size (measured per shape, profiles/r01_groupnorm_warps_sweep.txt): 12 warps for the 128 KB slabs
  // of the 32x32 stage, 8 for ~32 KB, 6 below that - small slabs on 12-warp CTAs run a few CTAs over one wave of
  // resident CTAs (768 CTAs on 148 x 5 slots) and pay for two.
  const long long slab_bytes = (long long)HW * V * 16;
  int warps = slab_bytes >= 100 * 1024 ? GN_MAX_WARPS : (slab_bytes >= 30 * 1024 ? 8 : 6);
  while (warps > 1 && (warps - 1) * (32 / VP) >= HW) --warps;  // tiny images: no idle warps
  dim3 grid(B, groups / gps);
  const int threads = warps * 32;
  if (VP == 4) {
    if (apply_silu) launch_kernel(groupnorm_kernel<true, 4>, grid, dim3(threads), 0, s, xp, (long long)ldx, yp, (long long)ldy, gamma, beta, HW, cpg, gps, V, eps, stats, groups);
    else launch_kernel(groupnorm_kernel<false, 4>, grid, dim3(threads), 0, s, xp, (long long)ldx, yp, (long long)ldy, gamma, beta, HW, cpg, gps, V, eps, stats, groups);
  } else {
    if (apply_silu) launch_kernel(groupnorm_kernel<true, 8>, grid, dim3(threads), 0, s, xp, (long long)ldx, yp, (long long)ldy, gamma, beta, HW, cpg, gps, V, eps, stats, groups);
    else launch_kernel(groupnorm_kernel<false, 8>, grid, dim3(threads), 0, s, xp, (long long)ldx, yp, (long long)ldy, gamma, beta, HW, cpg, gps, V, eps, stats, groups);
  }
  return check_cuda(cudaGetLastError(), "groupnorm launch");
}

extern "C" int idf_groupnorm_silu(const void* x, int64_t ldx, void* y, int64_t ldy, const float* gamma,
                                  const float* beta, int32_t B, int32_t HW, int32_t C, int32_t groups, float eps,
                                  int32_t apply_silu, idf_stream_t stream) {
  return groupnorm_impl(x, ldx, y, ldy, gamma, beta, B, HW, C, groups, eps, apply_silu, nullptr, stream);
}

extern "C" int idf_groupnorm_silu_rows(const void* x, int64_t ldx, void* y, int64_t ldy, const float* gamma,
                                       const float* beta, int32_t B, int32_t HW, int32_t C, int32_t groups, float eps,
                                       int32_t apply_silu, float* ws, int64_t ws_bytes, int64_t l2_bytes,
                                       idf_stream_t stream) {
  if (!x || !y || !gamma || !beta || !ws) return fail(IDF_ERR_ARG, "groupnorm_rows: null pointer");
  if (B <= 0 || HW <= 0 || C <= 0 || groups <= 0 || C % groups != 0 || C % 8 != 0 || C > 512 || groups > 2 * GN_MAX_GPS ||
      GNR_THREADS / (C / 8) < 1)
    return fail(IDF_ERR_UNSUPPORTED, "groupnorm_rows: C = %d (multiple of 8, <= 512) in %d groups unsupported", C, groups);
  if (ldx % 8 != 0 || ldy % 8 != 0 || (reinterpret_cast<uintptr_t>(x) & 15) || (reinterpret_cast<uintptr_t>(y) & 15))
    return fail(IDF_ERR_ARG, "groupnorm_rows: 16-byte alignment required");
  const int V = C / 8, R = GNR_THREADS / V, threads = V * R;
  int rows_per_chunk = 4 * GNR_UNROLL * R;  // four unrolled sweeps per CTA ...
  while (rows_per_chunk < 256) rows_per_chunk *= 2;  // ... and at least 256 rows
  const int chunks = (HW + rows_per_chunk - 1) / rows_per_chunk;
  if ((long long)B * chunks * groups * 2 * 4 > ws_bytes)
    return fail(IDF_ERR_ARG, "groupnorm_rows: workspace needs %lld bytes", (long long)B * chunks * groups * 2 * 4);
  // samples per launch pair: their input (read twice) should still be in L2 for the second read
  const long long sample_bytes = (long long)HW * C * 2;
  long long gsz = l2_bytes > 0 ? l2_bytes / sample_bytes : B;
  if (gsz < 1) gsz = 1;
  if (gsz > B) gsz = B;
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  const __nv_bfloat16* xp = reinterpret_cast<const __nv_bfloat16*>(x);
  __nv_bfloat16* yp = reinterpret_cast<__nv_bfloat16*>(y);
  const int cpg = C / groups;
  for (int b0 = 0; b0 < B; b0 += (int)gsz) {
    const int nb = (int)(B - b0 < gsz ? B - b0 : gsz);
    dim3 grid(chunks, nb);
    const __nv_bfloat16* xs = xp + (long long)b0 * HW * ldx;
    __nv_bfloat16* ys = yp + (long long)b0 * HW * ldy;
    float* part = ws + (long long)b0 * chunks * groups * 2;
    gn_rows_stats_kernel<<<grid, threads, 0, st>>>(xs, (long long)ldx, HW, C, cpg, rows_per_chunk, part, groups);
    if (apply_silu)
      gn_rows_apply_kernel<true><<<grid, threads, 0, st>>>(xs, (long long)ldx, ys, (long long)ldy, gamma, beta, HW, C, cpg,
                                                           rows_per_chunk, eps, part, groups);
    else
      gn_rows_apply_kernel<false><<<grid, threads, 0, st>>>(xs, (long long)ldx, ys, (long long)ldy, gamma, beta, HW, C,
                                                            cpg, rows_per_chunk, eps, part, groups);
  }
  return check_cuda(cudaGetLastError(), "groupnorm_rows launch");
}

extern "C" int idf_groupnorm_silu_train(const void* x, int64_t ldx, void* y, int64_t ldy, const float* gamma,
                                        const float* beta, int32_t B, int32_t HW, int32_t C, int32_t groups, float eps,
                                        int32_t apply_silu, float* stats, idf_stream_t stream) {
  if (!stats) return fail(IDF_ERR_ARG, "groupnorm_train: stats is null");
  return groupnorm_impl(x, ldx, y, ldy, gamma, beta, B, HW, C, groups, eps, apply_silu, stats, stream);
}

extern "C" int idf_softmax_rows(const float* in, int64_t ld_in, void* out, int64_t ld_out, int32_t rows, int32_t cols,
                                float scale, idf_stream_t stream) {
  if (!in || !out) return fail(IDF_ERR_ARG, "softmax_rows: null pointer");
  if (cols <= 0 || cols > 2048) return fail(IDF_ERR_UNSUPPORTED, "softmax_rows: cols = %d > 2048", cols);
  softmax_rows_kernel<<<rows, 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(
      in, ld_in, reinterpret_cast<__nv_bfloat16*>(out), ld_out, cols, scale);
  return check_cuda(cudaGetLastError(), "softmax_rows launch");
}
