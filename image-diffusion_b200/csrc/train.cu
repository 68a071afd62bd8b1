// Memory-bound kernels of the UNet training step (trainers/diffusion_trainer.py:141-187): GroupNorm(+SiLU)
// backward, bias / time-bias gradients (column sums), the layout movers' adjoints, the edge convolutions' backward,
// the embedding MLP backward, the MSE loss, gradient-norm clipping and Adam. Every reduction runs in a fixed order
// (no float atomics): gradients are bit-reproducible from run to run.
#include <stdlib.h>

#include "common.cuh"
#include "host.h"
#include "../../include/idf_b200.h"

namespace idf {

__device__ __forceinline__ void unpack8t(const uint4& v, float (&f)[8]) {
  f[0] = bf16_lo(v.x); f[1] = bf16_hi(v.x);
  f[2] = bf16_lo(v.y); f[3] = bf16_hi(v.y);
  f[4] = bf16_lo(v.z); f[5] = bf16_hi(v.z);
  f[6] = bf16_lo(v.w); f[7] = bf16_hi(v.w);
}
__device__ __forceinline__ uint4 pack8t(const float (&f)[8]) {
  uint4 o;
  o.x = pack_bf16x2(f[0], f[1]);
  o.y = pack_bf16x2(f[2], f[3]);
  o.z = pack_bf16x2(f[4], f[5]);
  o.w = pack_bf16x2(f[6], f[7]);
  return o;
}
__device__ __forceinline__ float sigmoid_tanh(float z) {
  float th;
  asm("tanh.approx.f32 %0, %1;" : "=f"(th) : "f"(0.5f * z));
  return fmaf(0.5f, th, 0.5f);
}
// d/dz [z * sigmoid(z)]
__device__ __forceinline__ float silu_grad(float z) {
  const float s = sigmoid_tanh(z);
  return s * fmaf(z, 1.f - s, 1.f);
}
__device__ __forceinline__ float silu_grad_exact(float z) {
  const float s = 1.f / (1.f + __expf(-z));
  return s * (1.f + z * (1.f - s));
}

// ---------------------------------------------------------------------------------------------------------------
// GroupNorm (+SiLU) backward. Same CTA / lane mapping as the forward kernel (norm.cu): grid = (B, slabs), lane l owns
// the 16-byte vector (l % VP) of pixel row (l / VP). Pass 1 accumulates per channel  s1 = sum dz,  s2 = sum dz*xhat
// (dz = dy * silu'(gamma*xhat + beta), or dy without SiLU); these are the per-sample contributions to dbeta / dgamma
// and, folded over a group's channels with gamma, the two group means of the GroupNorm backward formula. Pass 2
// re-reads x and dy (L2) and writes  dx = rstd * (gamma*dz - mean(gamma*dz) - xhat * mean(gamma*dz*xhat)) (+ add).
// ---------------------------------------------------------------------------------------------------------------
constexpr int GNB_MAX_WARPS = 12;
constexpr int GNB_MAX_GPS = 32;

template <bool SILU, int VP>
__global__ void __launch_bounds__(GNB_MAX_WARPS * 32) groupnorm_bwd_kernel(
    const __nv_bfloat16* __restrict__ x, long long ldx, const __nv_bfloat16* __restrict__ dy, long long lddy,
    const __nv_bfloat16* __restrict__ add, long long ldadd, __nv_bfloat16* __restrict__ dx, long long lddx,
    const float* __restrict__ gamma, const float* __restrict__ beta, const float* __restrict__ stats,
    float* __restrict__ dgamma_part, float* __restrict__ dbeta_part, float* __restrict__ colsum_part, long long ld_cs,
    int HW, int C, int groups, int cpg, int gps, int V) {
  __shared__ float ch_a[GNB_MAX_WARPS][VP * 8];
  __shared__ float ch_b[GNB_MAX_WARPS][VP * 8];
  __shared__ float ct_a[VP * 8];
  __shared__ float ct_b[VP * 8];
  __shared__ float g_a[GNB_MAX_GPS];
  __shared__ float g_b[GNB_MAX_GPS];
  constexpr int RPW = 32 / VP;
  const int b = blockIdx.x;
  const int c0 = blockIdx.y * gps * cpg;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarps = blockDim.x >> 5;
  const int v = lane % VP, prl = lane / VP;
  const bool active = v < V;
  const int rows_per_iter = nwarps * RPW;

  float rs[8], mr[8], ga[8], be[8];
#pragma unroll
  for (int e = 0; e < 8; ++e) {
    const int cl = active ? v * 8 + e : 0;
    const int g = blockIdx.y * gps + cl / cpg;
    const float mean = stats[((long long)b * groups + g) * 2], rstd = stats[((long long)b * groups + g) * 2 + 1];
    rs[e] = rstd;
    mr[e] = mean * rstd;
    ga[e] = gamma[c0 + cl];
    be[e] = beta[c0 + cl];
  }
  const __nv_bfloat16* xb = x + (long long)b * HW * ldx + c0 + v * 8;
  const __nv_bfloat16* dyb = dy + (long long)b * HW * lddy + c0 + v * 8;
  float s1[8], s2[8];
#pragma unroll
  for (int e = 0; e < 8; ++e) { s1[e] = 0.f; s2[e] = 0.f; }
  if (active) {
    auto accum = [&](const uint4& xr, const uint4& dr) {
      float xf[8], df[8];
      unpack8t(xr, xf);
      unpack8t(dr, df);
#pragma unroll
      for (int e = 0; e < 8; ++e) {
        const float xh = fmaf(xf[e], rs[e], -mr[e]);
        float dz = df[e];
        if (SILU) dz *= silu_grad(fmaf(ga[e], xh, be[e]));
        s1[e] += dz;
        s2[e] = fmaf(dz, xh, s2[e]);
      }
    };
    // batches of two pixel rows: four independent 16-byte loads in flight per thread
    int pix = warp * RPW + prl;
    for (; pix + rows_per_iter < HW; pix += 2 * rows_per_iter) {
      uint4 xr[2], dr[2];
#pragma unroll
      for (int i = 0; i < 2; ++i) {
        xr[i] = *reinterpret_cast<const uint4*>(xb + (long long)(pix + i * rows_per_iter) * ldx);
        dr[i] = *reinterpret_cast<const uint4*>(dyb + (long long)(pix + i * rows_per_iter) * lddy);
      }
#pragma unroll
      for (int i = 0; i < 2; ++i) accum(xr[i], dr[i]);
    }
    for (; pix < HW; pix += rows_per_iter)
      accum(*reinterpret_cast<const uint4*>(xb + (long long)pix * ldx),
            *reinterpret_cast<const uint4*>(dyb + (long long)pix * lddy));
  }
#pragma unroll
  for (int off = VP; off < 32; off <<= 1) {
#pragma unroll
    for (int e = 0; e < 8; ++e) {
      s1[e] += __shfl_xor_sync(0xffffffffu, s1[e], off);
      s2[e] += __shfl_xor_sync(0xffffffffu, s2[e], off);
    }
  }
  if (prl == 0) {
#pragma unroll
    for (int e = 0; e < 8; ++e) { ch_a[warp][v * 8 + e] = s1[e]; ch_b[warp][v * 8 + e] = s2[e]; }
  }
  __syncthreads();
  if (threadIdx.x < V * 8) {
    float a = 0.f, c = 0.f;
    for (int w = 0; w < nwarps; ++w) { a += ch_a[w][threadIdx.x]; c += ch_b[w][threadIdx.x]; }
    dbeta_part[(long long)b * C + c0 + threadIdx.x] = a;
    dgamma_part[(long long)b * C + c0 + threadIdx.x] = c;
    const float gm = gamma[c0 + threadIdx.x];
    ct_a[threadIdx.x] = a * gm;
    ct_b[threadIdx.x] = c * gm;
  }
  __syncthreads();
  if (threadIdx.x < gps) {
    float a = 0.f, c = 0.f;
    for (int k = 0; k < cpg; ++k) { a += ct_a[threadIdx.x * cpg + k]; c += ct_b[threadIdx.x * cpg + k]; }
    const float inv_cnt = 1.f / ((float)HW * (float)cpg);
    g_a[threadIdx.x] = a * inv_cnt;
    g_b[threadIdx.x] = c * inv_cnt;
  }
  __syncthreads();
  float ma[8], mb[8], cs[8];
#pragma unroll
  for (int e = 0; e < 8; ++e) {
    const int g = active ? (v * 8 + e) / cpg : 0;
    ma[e] = g_a[g];
    mb[e] = g_b[g];
    cs[e] = 0.f;
  }
  if (active) {
    __nv_bfloat16* dxb = dx + (long long)b * HW * lddx + c0 + v * 8;
    const __nv_bfloat16* ab = add ? add + (long long)b * HW * ldadd + c0 + v * 8 : nullptr;
    auto apply_store = [&](const uint4& xr, const uint4& dr, const uint4& ar, int pix) {
      float xf[8], df[8], af[8];
      unpack8t(xr, xf);
      unpack8t(dr, df);
      if (ab) unpack8t(ar, af);
      float o[8];
#pragma unroll
      for (int e = 0; e < 8; ++e) {
        const float xh = fmaf(xf[e], rs[e], -mr[e]);
        float dz = df[e];
        if (SILU) dz *= silu_grad(fmaf(ga[e], xh, be[e]));
        float d = rs[e] * (fmaf(ga[e], dz, -ma[e]) - xh * mb[e]);
        if (ab) d += af[e];
        o[e] = d;
        cs[e] += d;
      }
      *reinterpret_cast<uint4*>(dxb + (long long)pix * lddx) = pack8t(o);
    };
    const uint4 zero4 = make_uint4(0u, 0u, 0u, 0u);
    int pix = warp * RPW + prl;
    for (; pix + rows_per_iter < HW; pix += 2 * rows_per_iter) {
      uint4 xr[2], dr[2], ar[2];
#pragma unroll
      for (int i = 0; i < 2; ++i) {
        const long long pp = pix + i * rows_per_iter;
        xr[i] = *reinterpret_cast<const uint4*>(xb + pp * ldx);
        dr[i] = *reinterpret_cast<const uint4*>(dyb + pp * lddy);
        ar[i] = ab ? *reinterpret_cast<const uint4*>(ab + pp * ldadd) : zero4;
      }
#pragma unroll
      for (int i = 0; i < 2; ++i) apply_store(xr[i], dr[i], ar[i], pix + i * rows_per_iter);
    }
    for (; pix < HW; pix += rows_per_iter)
      apply_store(*reinterpret_cast<const uint4*>(xb + (long long)pix * ldx),
                  *reinterpret_cast<const uint4*>(dyb + (long long)pix * lddy),
                  ab ? *reinterpret_cast<const uint4*>(ab + (long long)pix * ldadd) : zero4, pix);
  }
  if (colsum_part == nullptr) return;
  // per-sample column sums of dx: the bias gradient of the layer that produced x (and, for a first-half conv, the
  // per-sample gradient of its time-projection bias)
#pragma unroll
  for (int off = VP; off < 32; off <<= 1) {
#pragma unroll
    for (int e = 0; e < 8; ++e) cs[e] += __shfl_xor_sync(0xffffffffu, cs[e], off);
  }
  if (prl == 0) {
#pragma unroll
    for (int e = 0; e < 8; ++e) ch_a[warp][v * 8 + e] = cs[e];
  }
  __syncthreads();
  if (threadIdx.x < V * 8) {
    float a = 0.f;
    for (int w = 0; w < nwarps; ++w) a += ch_a[w][threadIdx.x];
    colsum_part[(long long)b * ld_cs + c0 + threadIdx.x] = a;
  }
}

// Finalisation of one GroupNorm backward: dgamma / dbeta (and up to two copies of the fused bias gradient) as the
// fixed-order sums over the batch of the per-sample partials.
__global__ void __launch_bounds__(256) gn_bwd_finalize_kernel(
    const float* __restrict__ dgamma_part, const float* __restrict__ dbeta_part, const float* __restrict__ colsum_part,
    long long ld_cs, int B, int C, float* __restrict__ g_gamma, float* __restrict__ g_beta, float* __restrict__ g_bias1,
    float* __restrict__ g_bias2) {
  // 32 channels x 8 row groups per CTA: row group k sums samples k, k+8, ...; the eight partials are then added in
  // order (fixed summation order, independent loads in flight)
  __shared__ float red[3][8][33];
  const int cl = threadIdx.x & 31, rg = threadIdx.x >> 5;
  const int c = blockIdx.x * 32 + cl;
  float a = 0.f, bb = 0.f, cc = 0.f;
  if (c < C) {
    for (int r = rg; r < B; r += 8) {
      a += dgamma_part[(long long)r * C + c];
      bb += dbeta_part[(long long)r * C + c];
      if (colsum_part != nullptr) cc += colsum_part[(long long)r * ld_cs + c];
    }
  }
  red[0][rg][cl] = a;
  red[1][rg][cl] = bb;
  red[2][rg][cl] = cc;
  __syncthreads();
  if (rg == 0 && c < C) {
    float s0 = 0.f, s1 = 0.f, s2 = 0.f;
#pragma unroll
    for (int k = 0; k < 8; ++k) { s0 += red[0][k][cl]; s1 += red[1][k][cl]; s2 += red[2][k][cl]; }
    g_gamma[c] = s0;
    g_beta[c] = s1;
    if (g_bias1 != nullptr) g_bias1[c] = s2;
    if (g_bias2 != nullptr) g_bias2[c] = s2;
  }
}

// out[c] (+)= sum_r in[r, c]. 32 columns x 8 row groups per CTA: row group k sums rows k, k+8, ... (independent loads in
// flight instead of one dependent chain over all rows), the eight partials are then added in order: fixed summation
// order, bit-deterministic.
__global__ void __launch_bounds__(256) reduce_rows_kernel(const float* __restrict__ in, long long ld, int rows, int cols,
                                                          float* __restrict__ out, int accumulate) {
  __shared__ float red[8][33];
  const int cl = threadIdx.x & 31, rg = threadIdx.x >> 5;
  const int c = blockIdx.x * 32 + cl;
  float a = 0.f;
  if (c < cols)
    for (int r = rg; r < rows; r += 8) a += in[(long long)r * ld + c];
  red[rg][cl] = a;
  __syncthreads();
  if (rg == 0 && c < cols) {
    float s0 = 0.f;
#pragma unroll
    for (int k = 0; k < 8; ++k) s0 += red[k][cl];
    out[c] = accumulate ? out[c] + s0 : s0;
  }
}

// per_sample[b, c] = sum over the HW pixel rows of sample b of x[., c]. CTA = (sample, 64 channels), 256 threads =
// 32 pixel rows x 8 sixteen-byte vectors.
__global__ void __launch_bounds__(256) colsum_kernel(const __nv_bfloat16* __restrict__ x, long long ldx, int HW, int C,
                                                     float* __restrict__ per_sample, long long ld_ps) {
  __shared__ float part[32][65];
  const int b = blockIdx.x, c0 = blockIdx.y * 64;
  const int v = threadIdx.x & 7, pr = threadIdx.x >> 3;
  float s[8];
#pragma unroll
  for (int e = 0; e < 8; ++e) s[e] = 0.f;
  if (c0 + v * 8 < C) {
    const __nv_bfloat16* xb = x + (long long)b * HW * ldx + c0 + v * 8;
    for (int pix = pr; pix < HW; pix += 32) {
      float f[8];
      unpack8t(*reinterpret_cast<const uint4*>(xb + (long long)pix * ldx), f);
#pragma unroll
      for (int e = 0; e < 8; ++e) s[e] += f[e];
    }
  }
#pragma unroll
  for (int e = 0; e < 8; ++e) part[pr][v * 8 + e] = s[e];
  __syncthreads();
  if (threadIdx.x < 64 && c0 + threadIdx.x < C) {
    float a = 0.f;
    for (int r = 0; r < 32; ++r) a += part[r][threadIdx.x];
    per_sample[(long long)b * ld_ps + c0 + threadIdx.x] = a;
  }
}

// ---------------------------------------------------------------------------------------------------------------
// adjoints of the layout movers
// ---------------------------------------------------------------------------------------------------------------
// nearest-2x upsample backward: y[b, h, w, :] = sum of the 2x2 block of x (fp32 sum, bf16 out)
__global__ void sum2x2_kernel(const __nv_bfloat16* __restrict__ x, long long ldx, __nv_bfloat16* __restrict__ y,
                              long long ldy, int B, int H, int W, int C) {
  const int vec = C / 8;
  const long long total = (long long)B * H * W * vec;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int v = (int)(i % vec);
    long long pix = i / vec;
    const int w = (int)(pix % W);
    pix /= W;
    const int h = (int)(pix % H);
    const int b = (int)(pix / H);
    float acc[8];
#pragma unroll
    for (int e = 0; e < 8; ++e) acc[e] = 0.f;
#pragma unroll
    for (int dh = 0; dh < 2; ++dh)
#pragma unroll
      for (int dw = 0; dw < 2; ++dw) {
        const long long row = ((long long)b * 2 * H + 2 * h + dh) * 2 * W + 2 * w + dw;
        float f[8];
        unpack8t(*reinterpret_cast<const uint4*>(x + row * ldx + v * 8), f);
#pragma unroll
        for (int e = 0; e < 8; ++e) acc[e] += f[e];
      }
    *reinterpret_cast<uint4*>(y + (((long long)b * H + h) * W + w) * ldy + v * 8) = pack8t(acc);
  }
}

// inverse of space_to_depth2 with an optional addend: y[b, 2h+ph, 2w+pw, :] = planes[(ph*2+pw)*B + b, h, w, :] (+ add)
__global__ void depth_to_space2_kernel(const __nv_bfloat16* __restrict__ planes, __nv_bfloat16* __restrict__ y,
                                       long long ldy, const __nv_bfloat16* __restrict__ add, long long ldadd, int B,
                                       int H, int W, int C) {
  const int vec = C / 8, H2 = H / 2, W2 = W / 2;
  const long long total = (long long)B * H * W * vec;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int v = (int)(i % vec);
    long long pix = i / vec;
    const int w = (int)(pix % W);
    pix /= W;
    const int h = (int)(pix % H);
    const int b = (int)(pix / H);
    const int plane = (h & 1) * 2 + (w & 1);
    const long long prow = (((long long)plane * B + b) * H2 + (h >> 1)) * W2 + (w >> 1);
    uint4 val = *reinterpret_cast<const uint4*>(planes + prow * C + v * 8);
    const long long orow = ((long long)b * H + h) * W + w;
    if (add) {
      float f[8], a[8];
      unpack8t(val, f);
      unpack8t(*reinterpret_cast<const uint4*>(add + orow * ldadd + v * 8), a);
#pragma unroll
      for (int e = 0; e < 8; ++e) f[e] += a[e];
      val = pack8t(f);
    }
    *reinterpret_cast<uint4*>(y + orow * ldy + v * 8) = val;
  }
}

// zero the last row and column of every image (the gradient of Downsample's output padding is dropped)
__global__ void zero_last_rowcol_kernel(__nv_bfloat16* __restrict__ x, long long ldx, int B, int H, int W, int C) {
  const int vec = C / 8;
  const int edge = H + W - 1;
  const long long total = (long long)B * edge * vec;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int v = (int)(i % vec);
    const long long q = i / vec;
    const int k = (int)(q % edge);
    const int b = (int)(q / edge);
    const int h = k < W ? H - 1 : k - W;
    const int w = k < W ? k : W - 1;
    *reinterpret_cast<uint4*>(x + (((long long)b * H + h) * W + w) * ldx + v * 8) = make_uint4(0u, 0u, 0u, 0u);
  }
}

// ---------------------------------------------------------------------------------------------------------------
// edge convolutions, backward
// ---------------------------------------------------------------------------------------------------------------
// in_conv weight gradient: part[(b*nb + band)][co][ci*9 + tap] = sum over the band's pixels of dy[m, co] * x[b, ci, m+tap].
// CTA = (band of RB rows of one image, 128 output channels); thread = output channel, 9*CIN accumulators.
template <int CIN, int RB>
__global__ void __launch_bounds__(128) small_cin_wgrad_kernel(const float* __restrict__ x,
                                                              const __nv_bfloat16* __restrict__ dy, long long lddy,
                                                              float* __restrict__ part, int H, int W, int Cout) {
  extern __shared__ float patch[];  // [CIN][RB + 2][W + 2]
  const int nb = H / RB;
  const int b = blockIdx.x / nb, band = blockIdx.x % nb;
  const int h0 = band * RB;
  const int co = blockIdx.y * 128 + threadIdx.x;
  const int PW = W + 2;
  for (int i = threadIdx.x; i < CIN * (RB + 2) * PW; i += 128) {
    const int ci = i / ((RB + 2) * PW), rem = i % ((RB + 2) * PW);
    const int hh = h0 + rem / PW - 1, ww = rem % PW - 1;
    patch[i] = (hh >= 0 && hh < H && ww >= 0 && ww < W) ? x[(((long long)b * CIN + ci) * H + hh) * W + ww] : 0.f;
  }
  __syncthreads();
  float acc[CIN * 9];
#pragma unroll
  for (int k = 0; k < CIN * 9; ++k) acc[k] = 0.f;
  for (int r = 0; r < RB; ++r)
    for (int w = 0; w < W; ++w) {
      const float g = __bfloat162float(dy[(((long long)b * H + h0 + r) * W + w) * lddy + co]);
#pragma unroll
      for (int ci = 0; ci < CIN; ++ci)
#pragma unroll
        for (int t = 0; t < 9; ++t)
          acc[ci * 9 + t] = fmaf(g, patch[(ci * (RB + 2) + r + t / 3) * PW + w + t % 3], acc[ci * 9 + t]);
    }
  float* dst = part + ((long long)blockIdx.x * Cout + co) * (CIN * 9);
#pragma unroll
  for (int k = 0; k < CIN * 9; ++k) dst[k] = acc[k];
}

// out_conv data gradient: dh[m, c] = sum_tap sum_co dout[b, co, m - off(tap)] * W[co, c, tap]; CTA = (band, 128
// channels), thread = channel with its COUT*9 weights in registers.
template <int COUT, int RB>
__global__ void __launch_bounds__(128) small_cout_dgrad_kernel(const float* __restrict__ dout,
                                                               const float* __restrict__ wgt,
                                                               __nv_bfloat16* __restrict__ dh, long long lddh, int H,
                                                               int W, int C) {
  extern __shared__ float patch[];  // [COUT][RB + 2][W + 2]
  const int nb = H / RB;
  const int b = blockIdx.x / nb, band = blockIdx.x % nb;
  const int h0 = band * RB;
  const int c = blockIdx.y * 128 + threadIdx.x;
  const int PW = W + 2;
  for (int i = threadIdx.x; i < COUT * (RB + 2) * PW; i += 128) {
    const int co = i / ((RB + 2) * PW), rem = i % ((RB + 2) * PW);
    const int hh = h0 + rem / PW - 1, ww = rem % PW - 1;
    patch[i] = (hh >= 0 && hh < H && ww >= 0 && ww < W) ? dout[(((long long)b * COUT + co) * H + hh) * W + ww] : 0.f;
  }
  float wr[COUT * 9];
#pragma unroll
  for (int co = 0; co < COUT; ++co)
#pragma unroll
    for (int t = 0; t < 9; ++t) wr[co * 9 + t] = wgt[((long long)co * C + c) * 9 + t];
  __syncthreads();
  for (int r = 0; r < RB; ++r)
    for (int w = 0; w < W; ++w) {
      float a = 0.f;
#pragma unroll
      for (int co = 0; co < COUT; ++co)
#pragma unroll
        for (int t = 0; t < 9; ++t)  // forward tap t reads h[p + (t/3-1, t%3-1)]: its adjoint reads dout[p - off]
          a = fmaf(patch[(co * (RB + 2) + r + 2 - t / 3) * PW + w + 2 - t % 3], wr[co * 9 + t], a);
      dh[(((long long)b * H + h0 + r) * W + w) * lddh + c] = __float2bfloat16_rn(a);
    }
}

// out_conv weight gradient: part[(b*nb+band)][(co*C + c)*9 + tap] = sum over the band of dout[b, co, m] * h[m + off(tap), c]
template <int COUT, int RB>
__global__ void __launch_bounds__(128) small_cout_wgrad_kernel(const __nv_bfloat16* __restrict__ h, long long ldh,
                                                               const float* __restrict__ dout,
                                                               float* __restrict__ part, int H, int W, int C) {
  extern __shared__ float dsm[];  // [COUT][RB][W]
  const int nb = H / RB;
  const int b = blockIdx.x / nb, band = blockIdx.x % nb;
  const int h0 = band * RB;
  const int c = blockIdx.y * 128 + threadIdx.x;
  for (int i = threadIdx.x; i < COUT * RB * W; i += 128) {
    const int co = i / (RB * W), rem = i % (RB * W);
    dsm[i] = dout[(((long long)b * COUT + co) * H + h0 + rem / W) * W + rem % W];
  }
  __syncthreads();
  float acc[COUT * 9];
#pragma unroll
  for (int k = 0; k < COUT * 9; ++k) acc[k] = 0.f;
  for (int r = 0; r < RB; ++r)
    for (int w = 0; w < W; ++w) {
      float hv[9];
#pragma unroll
      for (int t = 0; t < 9; ++t) {
        const int hh = h0 + r + t / 3 - 1, ww = w + t % 3 - 1;
        hv[t] = (hh >= 0 && hh < H && ww >= 0 && ww < W)
                    ? __bfloat162float(h[(((long long)b * H + hh) * W + ww) * ldh + c]) : 0.f;
      }
#pragma unroll
      for (int co = 0; co < COUT; ++co) {
        const float g = dsm[(co * RB + r) * W + w];
#pragma unroll
        for (int t = 0; t < 9; ++t) acc[co * 9 + t] = fmaf(g, hv[t], acc[co * 9 + t]);
      }
    }
  float* dst = part + (long long)blockIdx.x * COUT * C * 9;
#pragma unroll
  for (int co = 0; co < COUT; ++co)
#pragma unroll
    for (int t = 0; t < 9; ++t) dst[((long long)co * C + c) * 9 + t] = acc[co * 9 + t];
}

// sum over (b, h, w) of an fp32 NCHW tensor per channel: out[c]; one CTA per channel
__global__ void __launch_bounds__(256) nchw_channel_sum_kernel(const float* __restrict__ x, int B, int C, int HW,
                                                               float* __restrict__ out) {
  __shared__ float red[256];
  const int c = blockIdx.x;
  float a = 0.f;
  for (int i = threadIdx.x; i < B * HW; i += 256) a += x[((long long)(i / HW) * C + c) * HW + i % HW];
  red[threadIdx.x] = a;
  __syncthreads();
  for (int s = 128; s > 0; s >>= 1) {
    if (threadIdx.x < s) red[threadIdx.x] += red[threadIdx.x + s];
    __syncthreads();
  }
  if (threadIdx.x == 0) out[c] = red[0];
}

// ---------------------------------------------------------------------------------------------------------------
// small fp32 GEMMs of the embedding MLP backward (R <= a few hundred rows)
// ---------------------------------------------------------------------------------------------------------------
// out[j, k] = sum_r A[r, j] * Bm[r, k]   (dW = dY^T X); optional bias_out[j] = sum_r A[r, j].
// CTA = (32 rows j, 256 columns k): the (R x 32) slice of A is staged in shared memory once, every thread keeps 32
// accumulators for its column, so Bm is read once per 32 output rows instead of once per output row.
constexpr int OR_JT = 32;
__global__ void __launch_bounds__(256) outer_rows_kernel(const float* __restrict__ A, int lda,
                                                         const float* __restrict__ Bm, int ldb, int R, int J, int K,
                                                         float* __restrict__ out, float* __restrict__ bias_out) {
  extern __shared__ float or_sm[];  // [R][OR_JT]
  const int k = blockIdx.x * 256 + threadIdx.x;
  const int j0 = blockIdx.y * OR_JT;
  for (int i = threadIdx.x; i < R * OR_JT; i += 256) {
    const int r = i / OR_JT, jj = i % OR_JT;
    or_sm[i] = (j0 + jj < J) ? A[(long long)r * lda + j0 + jj] : 0.f;
  }
  __syncthreads();
  float acc[OR_JT];
#pragma unroll
  for (int jj = 0; jj < OR_JT; ++jj) acc[jj] = 0.f;
  if (k < K) {
    for (int r = 0; r < R; ++r) {
      const float b = Bm[(long long)r * ldb + k];
#pragma unroll
      for (int jj = 0; jj < OR_JT; ++jj) acc[jj] = fmaf(or_sm[r * OR_JT + jj], b, acc[jj]);
    }
#pragma unroll
    for (int jj = 0; jj < OR_JT; ++jj)
      if (j0 + jj < J) out[(long long)(j0 + jj) * K + k] = acc[jj];
  }
  if (bias_out != nullptr && blockIdx.x == 0 && threadIdx.x < OR_JT && j0 + threadIdx.x < J) {
    float bs = 0.f;
    for (int r = 0; r < R; ++r) bs += or_sm[r * OR_JT + threadIdx.x];
    bias_out[j0 + threadIdx.x] = bs;
  }
}

// part[s][r, k] = sum_{j in slice s} X[r, j] * W[j, k]; thread = column k, 8 rows per CTA (blockIdx.y), slice = blockIdx.z
constexpr int XW_SLICE = 256;
__global__ void __launch_bounds__(128) xw_slice_kernel(const float* __restrict__ X, int ldx,
                                                       const float* __restrict__ W, int R, int J, int K,
                                                       float* __restrict__ part) {
  const int k = blockIdx.x * 128 + threadIdx.x;
  const int r0 = blockIdx.y * 8;
  const int j0 = blockIdx.z * XW_SLICE;
  const int j1 = j0 + XW_SLICE < J ? j0 + XW_SLICE : J;
  __shared__ float xs[8][XW_SLICE];
  for (int i = threadIdx.x; i < 8 * XW_SLICE; i += 128) {
    const int rr = i / XW_SLICE, jj = i % XW_SLICE;
    xs[rr][jj] = (r0 + rr < R && j0 + jj < J) ? X[(long long)(r0 + rr) * ldx + j0 + jj] : 0.f;
  }
  __syncthreads();
  if (k >= K) return;
  float acc[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) acc[i] = 0.f;
  for (int j = j0; j < j1; ++j) {
    const float w = W[(long long)j * K + k];
#pragma unroll
    for (int i = 0; i < 8; ++i) acc[i] = fmaf(xs[i][j - j0], w, acc[i]);
  }
#pragma unroll
  for (int i = 0; i < 8; ++i)
    if (r0 + i < R) part[((long long)blockIdx.z * R + r0 + i) * K + k] = acc[i];
}

// out[r, k] = (sum_s part[s][r, k]) * silu'(pre[r, k])
__global__ void xw_finish_kernel(const float* __restrict__ part, int slices, int R, int K,
                                 const float* __restrict__ pre, float* __restrict__ out) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (long long)R * K) return;
  float a = 0.f;
  for (int s = 0; s < slices; ++s) a += part[(long long)s * R * K + i];
  if (pre != nullptr) a *= silu_grad_exact(pre[i]);
  out[i] = a;
}

// dcls[c, k] = sum over rows r with ctx[r] == c of mask[r] * dtemb[r, k]
__global__ void class_grad_kernel(const float* __restrict__ dtemb, const int64_t* __restrict__ ctx,
                                  const float* __restrict__ mask, int R, int D, int num_classes,
                                  float* __restrict__ dcls) {
  const int k = blockIdx.x * blockDim.x + threadIdx.x;
  const int c = blockIdx.y;
  if (k >= D) return;
  float a = 0.f;
  for (int r = 0; r < R; ++r)
    if (ctx[r] == c) a = fmaf(mask ? mask[r] : 1.f, dtemb[(long long)r * D + k], a);
  dcls[(long long)c * D + k] = a;
}

// ---------------------------------------------------------------------------------------------------------------
// loss, gradient norm, Adam
// ---------------------------------------------------------------------------------------------------------------
// loss = mean((pred - target)^2); dpred = grad_scale * 2 * (pred - target) / n. Single CTA, fixed-order reduction.
__global__ void __launch_bounds__(1024) mse_loss_grad_kernel(const float* __restrict__ pred,
                                                             const float* __restrict__ target, long long n,
                                                             float grad_scale, float* __restrict__ dpred,
                                                             float* __restrict__ loss) {
  __shared__ float red[1024];
  const float k = 2.f * grad_scale / (float)n;
  float a = 0.f;
  for (long long i = threadIdx.x; i < n; i += 1024) {
    const float d = pred[i] - target[i];
    a = fmaf(d, d, a);
    if (dpred) dpred[i] = k * d;
  }
  red[threadIdx.x] = a;
  __syncthreads();
  for (int s = 512; s > 0; s >>= 1) {
    if (threadIdx.x < s) red[threadIdx.x] += red[threadIdx.x + s];
    __syncthreads();
  }
  if (threadIdx.x == 0 && loss) *loss = red[0] / (float)n;
}

constexpr int SQ_BLOCKS = 1184;  // 148 SMs x 8
__global__ void __launch_bounds__(256) sumsq_part_kernel(const float* __restrict__ g, long long n,
                                                         float* __restrict__ part) {
  __shared__ float red[256];
  float a = 0.f;
  const long long n4 = n / 4;
  const float4* g4 = reinterpret_cast<const float4*>(g);
  for (long long i = (long long)blockIdx.x * 256 + threadIdx.x; i < n4; i += (long long)gridDim.x * 256) {
    const float4 v = g4[i];
    a = fmaf(v.x, v.x, a); a = fmaf(v.y, v.y, a); a = fmaf(v.z, v.z, a); a = fmaf(v.w, v.w, a);
  }
  if (blockIdx.x == 0)
    for (long long i = n4 * 4 + threadIdx.x; i < n; i += 256) a = fmaf(g[i], g[i], a);
  red[threadIdx.x] = a;
  __syncthreads();
  for (int s = 128; s > 0; s >>= 1) {
    if (threadIdx.x < s) red[threadIdx.x] += red[threadIdx.x + s];
    __syncthreads();
  }
  if (threadIdx.x == 0) part[blockIdx.x] = red[0];
}
// out[0] = ||g||_2 / inv_scale_div (the unscaled norm), out[1] = clip coefficient min(1, max_norm / (norm + 1e-6))
__global__ void __launch_bounds__(256) sumsq_finish_kernel(const float* __restrict__ part, int nparts, float grad_div,
                                                           float max_norm, float* __restrict__ out) {
  __shared__ float red[256];
  float a = 0.f;
  for (int i = threadIdx.x; i < nparts; i += 256) a += part[i];
  red[threadIdx.x] = a;
  __syncthreads();
  for (int s = 128; s > 0; s >>= 1) {
    if (threadIdx.x < s) red[threadIdx.x] += red[threadIdx.x + s];
    __syncthreads();
  }
  if (threadIdx.x == 0) {
    const float norm = sqrtf(red[0]) / grad_div;
    out[0] = norm;
    float coef = max_norm > 0.f ? max_norm / (norm + 1e-6f) : 1.f;
    out[1] = coef < 1.f ? coef : 1.f;
  }
}

// torch.optim.Adam (no weight decay, no amsgrad): g = grad * clip2[1] / grad_div; m, v updated in place;
// p -= lr / bc1 * m / (sqrt(v) / sqrt(bc2) + eps). hyper = {lr, bc1 = 1 - beta1^step, sqrt(bc2)} lives in device memory
// so that a captured CUDA graph can be replayed with a new learning rate / step count.
__global__ void __launch_bounds__(256) adam_kernel(float* __restrict__ p, const float* __restrict__ g,
                                                   float* __restrict__ m, float* __restrict__ v, long long n,
                                                   const float* __restrict__ hyper, float beta1, float beta2, float eps,
                                                   float grad_div, const float* __restrict__ gscale) {
  const float gs = (gscale ? gscale[1] : 1.f) / grad_div;
  const float step = hyper[0] / hyper[1];
  const float bc2_sqrt = hyper[2];
  auto upd = [&](float& pi, float gi, float& mi, float& vi) {
    gi *= gs;
    mi = fmaf(beta1, mi, (1.f - beta1) * gi);
    vi = fmaf(beta2, vi, (1.f - beta2) * gi * gi);
    pi -= step * mi / (sqrtf(vi) / bc2_sqrt + eps);
  };
  const long long n4 = n / 4;  // (the flat buffers are 16-byte aligned)
  float4* p4 = reinterpret_cast<float4*>(p);
  const float4* g4 = reinterpret_cast<const float4*>(g);
  float4* m4 = reinterpret_cast<float4*>(m);
  float4* v4 = reinterpret_cast<float4*>(v);
  for (long long i = (long long)blockIdx.x * 256 + threadIdx.x; i < n4; i += (long long)gridDim.x * 256) {
    float4 pp = p4[i], mm = m4[i], vv = v4[i];
    const float4 gg = g4[i];
    upd(pp.x, gg.x, mm.x, vv.x);
    upd(pp.y, gg.y, mm.y, vv.y);
    upd(pp.z, gg.z, mm.z, vv.z);
    upd(pp.w, gg.w, mm.w, vv.w);
    p4[i] = pp;
    m4[i] = mm;
    v4[i] = vv;
  }
  if (blockIdx.x == 0)
    for (long long i = n4 * 4 + threadIdx.x; i < n; i += 256) upd(p[i], g[i], m[i], v[i]);
}

// KL reparametrisation of the stored (mean || logvar) latents + forward diffusion (diffusion_trainer.py:149-164,
// components.py:399-403): x0 = mean + rn * exp(0.5 * clamp(logvar, -30, 20)); x_t = sqrt(acp[t]) x0 + sqrt(1-acp[t]) noise.
// lat is (N, 2*chw) fp32 when rn != nullptr, else (N, chw) and x0 = lat.
__global__ void reparam_add_noise_kernel(const float* __restrict__ lat, const float* __restrict__ rn,
                                         const float* __restrict__ noise, const int64_t* __restrict__ t,
                                         const float* __restrict__ sa, const float* __restrict__ s1,
                                         float* __restrict__ out, int N, int chw) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (long long)N * chw) return;
  const int n = (int)(i / chw), k = (int)(i % chw);
  float x0;
  if (rn != nullptr) {
    const float mean = lat[(long long)n * 2 * chw + k];
    const float lv = fminf(fmaxf(lat[(long long)n * 2 * chw + chw + k], -30.f), 20.f);
    x0 = fmaf(rn[i], expf(0.5f * lv), mean);
  } else {
    x0 = lat[i];
  }
  const int64_t tt = t[n];
  out[i] = sa[tt] * x0 + s1[tt] * noise[i];
}

// row sums of dO * O per (token, head): D[m, h] = sum_d dO[m, h*hd + d] * O[m, h*hd + d]
__global__ void attn_delta_kernel(const __nv_bfloat16* __restrict__ dO, long long ld_do,
                                  const __nv_bfloat16* __restrict__ O, long long ld_o, int M, int heads, int hd,
                                  float* __restrict__ delta) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (long long)M * heads) return;
  const int h = (int)(i % heads);
  const long long m = i / heads;
  const __nv_bfloat16* a = dO + m * ld_do + h * hd;
  const __nv_bfloat16* b = O + m * ld_o + h * hd;
  float s = 0.f;
  for (int d = 0; d < hd; d += 8) {
    float fa[8], fb[8];
    unpack8t(*reinterpret_cast<const uint4*>(a + d), fa);
    unpack8t(*reinterpret_cast<const uint4*>(b + d), fb);
#pragma unroll
    for (int e = 0; e < 8; ++e) s = fmaf(fa[e], fb[e], s);
  }
  delta[i] = s;
}

// fp32 (M, C) -> bf16 (M, ld) (the atomically accumulated dQ)
__global__ void f32_to_bf16_rows_kernel(const float* __restrict__ x, __nv_bfloat16* __restrict__ y, long long ldy,
                                        long long M, int C) {
  const int vec = C / 8;
  const long long total = M * vec;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int v = (int)(i % vec);
    const long long m = i / vec;
    const float4 a = *reinterpret_cast<const float4*>(x + m * C + v * 8);
    const float4 b = *reinterpret_cast<const float4*>(x + m * C + v * 8 + 4);
    const float f[8] = {a.x, a.y, a.z, a.w, b.x, b.y, b.z, b.w};
    *reinterpret_cast<uint4*>(y + m * ldy + v * 8) = pack8t(f);
  }
}

// ---------------------------------------------------------------------------------------------------------------
// Weight re-pack after an optimizer step: every bf16 operand copy (forward and data-gradient layouts of all conv /
// linear weights) and the few fused fp32 vectors, in ONE launch driven by a job table. CTA = (job, outer index): the
// source elements of one output row are gathered into shared memory (contiguous for forward layouts, 36-byte runs for
// data-gradient layouts) and written out as coalesced bf16 rows  dst[outer*dldo + t*dldt + inner].
// ---------------------------------------------------------------------------------------------------------------
constexpr int PACK_CHUNK = 1024;  // inner elements staged per pass (x taps <= 9 -> 36 KiB of shared memory)

__global__ void __launch_bounds__(256) pack_weights_kernel(const idf_pack_job* __restrict__ jobs,
                                                           const int32_t* __restrict__ prefix, int njobs) {
  extern __shared__ float pk_sm[];
  int lo = 0, hi = njobs - 1;
  while (lo < hi) {
    const int mid = (lo + hi + 1) >> 1;
    if (prefix[mid] <= (int)blockIdx.x) lo = mid; else hi = mid - 1;
  }
  const idf_pack_job jb = jobs[lo];
  const long long outer = (int)blockIdx.x - prefix[lo];
  const float* src = jb.src + outer * jb.so;
  if (jb.out_f32) {
    float* dst = reinterpret_cast<float*>(jb.dst) + outer * jb.dldo;
    const float* src2 = jb.src2 ? jb.src2 + outer * jb.so : nullptr;
    for (int k = threadIdx.x; k < jb.n_inner; k += blockDim.x)
      dst[k] = src[(long long)k * jb.si] + (src2 ? src2[(long long)k * jb.si] : 0.f);
    return;
  }
  __nv_bfloat16* dst = reinterpret_cast<__nv_bfloat16*>(jb.dst) + outer * jb.dldo;
  const int T = jb.n_taps;
  for (int i0 = 0; i0 < jb.n_inner; i0 += PACK_CHUNK) {
    const int ni = jb.n_inner - i0 < PACK_CHUNK ? jb.n_inner - i0 : PACK_CHUNK;
    const int n = ni * T;
    for (int f = threadIdx.x; f < n; f += blockDim.x) {
      const int inner = f / T, t = f - inner * T;
      pk_sm[f] = src[(long long)(i0 + inner) * jb.si + (long long)t * jb.st];
    }
    __syncthreads();
    for (int g = threadIdx.x; g < n; g += blockDim.x) {
      const int t = g / ni, inner = g - t * ni;
      dst[(long long)t * jb.dldt + i0 + inner] = __float2bfloat16_rn(pk_sm[inner * T + t]);
    }
    __syncthreads();
  }
}

static inline unsigned grid_for(long long work, int block, int cap = 148 * 16) {
  long long g = (work + block - 1) / block;
  if (g > cap) g = cap;
  if (g < 1) g = 1;
  return (unsigned)g;
}

}  // namespace idf

using namespace idf;

static inline cudaStream_t S(idf_stream_t s) { return reinterpret_cast<cudaStream_t>(s); }
static inline const __nv_bfloat16* BF(const void* p) { return reinterpret_cast<const __nv_bfloat16*>(p); }
static inline __nv_bfloat16* BF(void* p) { return reinterpret_cast<__nv_bfloat16*>(p); }

extern "C" int idf_groupnorm_silu_bwd(const void* x, int64_t ldx, const void* dy, int64_t lddy, const void* add,
                                      int64_t ldadd, void* dx, int64_t lddx, const float* gamma, const float* beta,
                                      const float* stats, float* dgamma_part, float* dbeta_part, float* colsum_part,
                                      int64_t ld_cs, int32_t B, int32_t HW, int32_t C, int32_t groups,
                                      int32_t apply_silu, idf_stream_t stream) {
  if (!x || !dy || !dx || !gamma || !beta || !stats || !dgamma_part || !dbeta_part)
    return fail(IDF_ERR_ARG, "groupnorm_bwd: null pointer");
  if (B <= 0 || HW <= 0 || C <= 0 || groups <= 0 || C % groups != 0) return fail(IDF_ERR_ARG, "groupnorm_bwd: bad shape");
  const int cpg = C / groups;
  int gps = 0;
  for (int d = 8; d >= 1; --d)
    if (groups % d == 0 && (d * cpg) % 8 == 0 && d * cpg / 8 <= 8) { gps = d; break; }
  if (gps == 0) return fail(IDF_ERR_UNSUPPORTED, "groupnorm_bwd: %d channels in %d groups does not split into slabs", C, groups);
  const int V = gps * cpg / 8;
  if (ldx % 8 || lddy % 8 || lddx % 8 || (add && ldadd % 8) || (reinterpret_cast<uintptr_t>(x) & 15) ||
      (reinterpret_cast<uintptr_t>(dy) & 15) || (reinterpret_cast<uintptr_t>(dx) & 15) ||
      (add && (reinterpret_cast<uintptr_t>(add) & 15)))
    return fail(IDF_ERR_ARG, "groupnorm_bwd: 16-byte alignment required");
  const int VP = V <= 4 ? 4 : 8;
  // CTA size by slab size, as in the forward kernel (norm.cu); IDF_GNB_WARPS_BY_SLAB=0: always 12 warps
  static const int by_slab = [] { const char* e = getenv("IDF_GNB_WARPS_BY_SLAB"); return e ? atoi(e) : 1; }();
  const long long slab_bytes = (long long)HW * V * 16;
  int warps = (!by_slab || slab_bytes >= 100 * 1024) ? GNB_MAX_WARPS : (slab_bytes >= 30 * 1024 ? 8 : 6);
  while (warps > 1 && (warps - 1) * (32 / VP) >= HW) --warps;
  dim3 grid(B, groups / gps);
  const int threads = warps * 32;
#define GNB_LAUNCH(SILU_, VP_)                                                                                        \
  groupnorm_bwd_kernel<SILU_, VP_><<<grid, threads, 0, S(stream)>>>(BF(x), ldx, BF(dy), lddy, BF(add), ldadd, BF(dx),  \
                                                                    lddx, gamma, beta, stats, dgamma_part, dbeta_part, \
                                                                    colsum_part, ld_cs, HW, C, groups, cpg, gps, V)
  if (VP == 4) { if (apply_silu) GNB_LAUNCH(true, 4); else GNB_LAUNCH(false, 4); }
  else { if (apply_silu) GNB_LAUNCH(true, 8); else GNB_LAUNCH(false, 8); }
#undef GNB_LAUNCH
  return check_cuda(cudaGetLastError(), "groupnorm_bwd launch");
}

extern "C" int idf_groupnorm_bwd_finalize(const float* dgamma_part, const float* dbeta_part, const float* colsum_part,
                                          int64_t ld_cs, int32_t B, int32_t C, float* g_gamma, float* g_beta,
                                          float* g_bias1, float* g_bias2, idf_stream_t stream) {
  if (!dgamma_part || !dbeta_part || !g_gamma || !g_beta || B <= 0 || C <= 0)
    return fail(IDF_ERR_ARG, "groupnorm_bwd_finalize: bad argument");
  if ((g_bias1 || g_bias2) && !colsum_part) return fail(IDF_ERR_ARG, "groupnorm_bwd_finalize: bias without colsum");
  gn_bwd_finalize_kernel<<<(C + 31) / 32, 256, 0, S(stream)>>>(dgamma_part, dbeta_part, colsum_part, ld_cs, B, C, g_gamma,
                                                               g_beta, g_bias1, g_bias2);
  return check_cuda(cudaGetLastError(), "groupnorm_bwd_finalize launch");
}

extern "C" int idf_reduce_rows_f32(const float* in, int64_t ld, int32_t rows, int32_t cols, float* out,
                                   int32_t accumulate, idf_stream_t stream) {
  if (!in || !out || rows <= 0 || cols <= 0) return fail(IDF_ERR_ARG, "reduce_rows: bad argument");
  reduce_rows_kernel<<<(cols + 31) / 32, 256, 0, S(stream)>>>(in, ld, rows, cols, out, accumulate);
  return check_cuda(cudaGetLastError(), "reduce_rows launch");
}

extern "C" int idf_colsum_bf16(const void* x, int64_t ldx, int32_t B, int32_t HW, int32_t C, float* per_sample,
                               int64_t ld_ps, float* total, int32_t accumulate_total, idf_stream_t stream) {
  if (!x || !per_sample) return fail(IDF_ERR_ARG, "colsum: null pointer");
  if (C % 8 != 0 || ldx % 8 != 0 || (reinterpret_cast<uintptr_t>(x) & 15)) return fail(IDF_ERR_ARG, "colsum: alignment");
  colsum_kernel<<<dim3(B, (C + 63) / 64), 256, 0, S(stream)>>>(BF(x), ldx, HW, C, per_sample, ld_ps);
  if (total != nullptr)
    reduce_rows_kernel<<<(C + 31) / 32, 256, 0, S(stream)>>>(per_sample, ld_ps, B, C, total, accumulate_total);
  return check_cuda(cudaGetLastError(), "colsum launch");
}

extern "C" int idf_sum2x2_bf16(const void* x, int64_t ldx, void* y, int64_t ldy, int32_t B, int32_t H, int32_t W,
                               int32_t C, idf_stream_t stream) {
  if (!x || !y || C % 8 != 0) return fail(IDF_ERR_ARG, "sum2x2: bad argument");
  sum2x2_kernel<<<grid_for((long long)B * H * W * (C / 8), 256), 256, 0, S(stream)>>>(BF(x), ldx, BF(y), ldy, B, H, W, C);
  return check_cuda(cudaGetLastError(), "sum2x2 launch");
}

extern "C" int idf_depth_to_space2(const void* planes, void* y, int64_t ldy, const void* add, int64_t ldadd, int32_t B,
                                   int32_t H, int32_t W, int32_t C, idf_stream_t stream) {
  if (!planes || !y || C % 8 != 0 || (H & 1) || (W & 1)) return fail(IDF_ERR_ARG, "depth_to_space2: bad argument");
  depth_to_space2_kernel<<<grid_for((long long)B * H * W * (C / 8), 256), 256, 0, S(stream)>>>(BF(planes), BF(y), ldy,
                                                                                               BF(add), ldadd, B, H, W, C);
  return check_cuda(cudaGetLastError(), "depth_to_space2 launch");
}

extern "C" int idf_zero_last_rowcol(void* x, int64_t ldx, int32_t B, int32_t H, int32_t W, int32_t C,
                                    idf_stream_t stream) {
  if (!x || C % 8 != 0) return fail(IDF_ERR_ARG, "zero_last_rowcol: bad argument");
  zero_last_rowcol_kernel<<<grid_for((long long)B * (H + W - 1) * (C / 8), 256), 256, 0, S(stream)>>>(BF(x), ldx, B, H, W, C);
  return check_cuda(cudaGetLastError(), "zero_last_rowcol launch");
}

static int pick_band(int H) { return H % 2 == 0 ? 2 : 0; }

extern "C" int idf_conv3x3_small_cin_wgrad(const float* x, const void* dy, int64_t lddy, float* grad_w, float* part,
                                           int64_t part_bytes, int32_t B, int32_t Cin, int32_t H, int32_t W,
                                           int32_t Cout, idf_stream_t stream) {
  if (!x || !dy || !grad_w || !part) return fail(IDF_ERR_ARG, "small_cin_wgrad: null pointer");
  if (Cin != 3 || Cout % 128 != 0) return fail(IDF_ERR_UNSUPPORTED, "small_cin_wgrad: Cin must be 3, Cout %% 128 == 0");
  const int RB = pick_band(H);
  if (RB == 0) return fail(IDF_ERR_UNSUPPORTED, "small_cin_wgrad: H %% 2 != 0");
  const int nparts = B * (H / RB);
  if ((long long)nparts * Cout * 27 * 4 > part_bytes) return fail(IDF_ERR_ARG, "small_cin_wgrad: scratch too small");
  const int smem = 3 * (RB + 2) * (W + 2) * 4;
  small_cin_wgrad_kernel<3, 2><<<dim3(nparts, Cout / 128), 128, smem, S(stream)>>>(x, BF(dy), lddy, part, H, W, Cout);
  reduce_rows_kernel<<<(Cout * 27 + 31) / 32, 256, 0, S(stream)>>>(part, (long long)Cout * 27, nparts, Cout * 27, grad_w, 0);
  return check_cuda(cudaGetLastError(), "small_cin_wgrad launch");
}

extern "C" int idf_conv3x3_small_cout_bwd(const void* h, int64_t ldh, const float* dout, const float* w, void* dh,
                                          int64_t lddh, float* grad_w, float* grad_b, float* part, int64_t part_bytes,
                                          int32_t B, int32_t C, int32_t H, int32_t W, int32_t Cout,
                                          idf_stream_t stream) {
  if (!h || !dout || !w || !dh || !grad_w || !grad_b || !part) return fail(IDF_ERR_ARG, "small_cout_bwd: null pointer");
  if (Cout != 3 || C % 128 != 0) return fail(IDF_ERR_UNSUPPORTED, "small_cout_bwd: Cout must be 3, C %% 128 == 0");
  const int RB = pick_band(H);
  if (RB == 0) return fail(IDF_ERR_UNSUPPORTED, "small_cout_bwd: H %% 2 != 0");
  const int nparts = B * (H / RB);
  if ((long long)nparts * Cout * C * 9 * 4 > part_bytes) return fail(IDF_ERR_ARG, "small_cout_bwd: scratch too small");
  small_cout_dgrad_kernel<3, 2><<<dim3(nparts, C / 128), 128, 3 * (RB + 2) * (W + 2) * 4, S(stream)>>>(dout, w, BF(dh), lddh, H, W, C);
  small_cout_wgrad_kernel<3, 2><<<dim3(nparts, C / 128), 128, 3 * RB * W * 4, S(stream)>>>(BF(h), ldh, dout, part, H, W, C);
  reduce_rows_kernel<<<(Cout * C * 9 + 31) / 32, 256, 0, S(stream)>>>(part, (long long)Cout * C * 9, nparts, Cout * C * 9, grad_w, 0);
  nchw_channel_sum_kernel<<<Cout, 256, 0, S(stream)>>>(dout, B, Cout, H * W, grad_b);
  return check_cuda(cudaGetLastError(), "small_cout_bwd launch");
}

extern "C" int idf_embed_time_class_bwd(const float* dtable, const int64_t* ctx, const float* ctx_mask, int32_t R,
                                        int32_t D, int32_t P, int32_t num_classes, const float* w2, const float* wp,
                                        const float* saved, float* g_w1, float* g_b1, float* g_w2, float* g_b2,
                                        float* g_cls, float* g_wp, float* g_bp, float* scratch, int64_t scratch_bytes,
                                        idf_stream_t stream) {
  if (!dtable || !w2 || !wp || !saved || !g_w1 || !g_b1 || !g_w2 || !g_b2 || !g_wp || !g_bp || !scratch)
    return fail(IDF_ERR_ARG, "embed_bwd: null pointer");
  // saved (from idf_embed_time_class_train): e [R, D], z1 [R, 4D], a1 [R, 4D], temb [R, D], s [R, D]
  const float* e = saved;
  const float* z1 = e + (long long)R * D;
  const float* a1 = z1 + (long long)R * 4 * D;
  const float* temb = a1 + (long long)R * 4 * D;
  const float* sv = temb + (long long)R * D;
  const int slices_p = (P + XW_SLICE - 1) / XW_SLICE, slices_d = (D + XW_SLICE - 1) / XW_SLICE;
  const int max_slices = slices_p > slices_d ? slices_p : slices_d;
  const long long need = ((long long)max_slices * R * 4 * D + (long long)R * D + (long long)R * 4 * D) * 4;
  if (need > scratch_bytes) return fail(IDF_ERR_ARG, "embed_bwd: scratch needs %lld bytes", need);
  float* part = scratch;
  float* dtemb = part + (long long)max_slices * R * 4 * D;
  float* dz1 = dtemb + (long long)R * D;
  cudaStream_t s = S(stream);
  // time projections: g_wp = dtable^T s, g_bp = colsum(dtable); ds = dtable wp; dtemb = ds * silu'(temb)
  const int or_smem = R * OR_JT * (int)sizeof(float);
  if (or_smem > 48 * 1024) return fail(IDF_ERR_UNSUPPORTED, "embed_bwd: batch %d too large for the staged outer product", R);
  outer_rows_kernel<<<dim3((D + 255) / 256, (P + OR_JT - 1) / OR_JT), 256, or_smem, s>>>(dtable, P, sv, D, R, P, D, g_wp, g_bp);
  xw_slice_kernel<<<dim3((D + 127) / 128, (R + 7) / 8, slices_p), 128, 0, s>>>(dtable, P, wp, R, P, D, part);
  xw_finish_kernel<<<(R * D + 255) / 256, 256, 0, s>>>(part, slices_p, R, D, temb, dtemb);
  if (ctx != nullptr && g_cls != nullptr)
    class_grad_kernel<<<dim3((D + 127) / 128, num_classes), 128, 0, s>>>(dtemb, ctx, ctx_mask, R, D, num_classes, g_cls);
  // second Linear: g_w2 = dtemb^T a1, g_b2 = colsum(dtemb); da1 = dtemb w2; dz1 = da1 * silu'(z1)
  outer_rows_kernel<<<dim3((4 * D + 255) / 256, (D + OR_JT - 1) / OR_JT), 256, or_smem, s>>>(dtemb, D, a1, 4 * D, R, D, 4 * D, g_w2, g_b2);
  xw_slice_kernel<<<dim3((4 * D + 127) / 128, (R + 7) / 8, slices_d), 128, 0, s>>>(dtemb, D, w2, R, D, 4 * D, part);
  xw_finish_kernel<<<(R * 4 * D + 255) / 256, 256, 0, s>>>(part, slices_d, R, 4 * D, z1, dz1);
  // first Linear: g_w1 = dz1^T e, g_b1 = colsum(dz1)
  outer_rows_kernel<<<dim3((D + 255) / 256, (4 * D + OR_JT - 1) / OR_JT), 256, or_smem, s>>>(dz1, 4 * D, e, D, R, 4 * D, D, g_w1, g_b1);
  return check_cuda(cudaGetLastError(), "embed_bwd launch");
}

extern "C" int idf_mse_loss_grad(const float* pred, const float* target, int64_t n, float grad_scale, float* dpred,
                                 float* loss, idf_stream_t stream) {
  if (!pred || !target || n <= 0) return fail(IDF_ERR_ARG, "mse_loss_grad: bad argument");
  mse_loss_grad_kernel<<<1, 1024, 0, S(stream)>>>(pred, target, n, grad_scale, dpred, loss);
  return check_cuda(cudaGetLastError(), "mse_loss_grad launch");
}

extern "C" int idf_grad_norm_clip(const float* grad, int64_t n, float grad_div, float max_norm, float* out2,
                                  float* scratch, int64_t scratch_bytes, idf_stream_t stream) {
  if (!grad || !out2 || !scratch || n <= 0) return fail(IDF_ERR_ARG, "grad_norm_clip: bad argument");
  if (scratch_bytes < SQ_BLOCKS * 4 || (reinterpret_cast<uintptr_t>(grad) & 15))
    return fail(IDF_ERR_ARG, "grad_norm_clip: scratch / alignment");
  sumsq_part_kernel<<<SQ_BLOCKS, 256, 0, S(stream)>>>(grad, n, scratch);
  sumsq_finish_kernel<<<1, 256, 0, S(stream)>>>(scratch, SQ_BLOCKS, grad_div, max_norm, out2);
  return check_cuda(cudaGetLastError(), "grad_norm_clip launch");
}

extern "C" int idf_adam_step(float* param, const float* grad, float* exp_avg, float* exp_avg_sq, int64_t n,
                             const float* hyper, float beta1, float beta2, float eps, float grad_div,
                             const float* clip2, idf_stream_t stream) {
  if (!param || !grad || !exp_avg || !exp_avg_sq || !hyper || n <= 0) return fail(IDF_ERR_ARG, "adam_step: bad argument");
  if ((reinterpret_cast<uintptr_t>(param) | reinterpret_cast<uintptr_t>(grad) | reinterpret_cast<uintptr_t>(exp_avg) |
       reinterpret_cast<uintptr_t>(exp_avg_sq)) & 15)
    return fail(IDF_ERR_ARG, "adam_step: buffers must be 16-byte aligned");
  adam_kernel<<<grid_for(n, 256, 148 * 8), 256, 0, S(stream)>>>(param, grad, exp_avg, exp_avg_sq, n, hyper, beta1, beta2,
                                                               eps, grad_div, clip2);
  return check_cuda(cudaGetLastError(), "adam_step launch");
}

extern "C" int idf_reparam_add_noise(const float* latents, const float* reparam_noise, const float* noise,
                                     const int64_t* t, const float* sqrt_alpha_cum_prod,
                                     const float* sqrt_one_minus_alpha_cum_prod, float* out, int32_t N, int32_t chw,
                                     idf_stream_t stream) {
  if (!latents || !noise || !t || !sqrt_alpha_cum_prod || !sqrt_one_minus_alpha_cum_prod || !out)
    return fail(IDF_ERR_ARG, "reparam_add_noise: null pointer");
  reparam_add_noise_kernel<<<(unsigned)(((long long)N * chw + 255) / 256), 256, 0, S(stream)>>>(
      latents, reparam_noise, noise, t, sqrt_alpha_cum_prod, sqrt_one_minus_alpha_cum_prod, out, N, chw);
  return check_cuda(cudaGetLastError(), "reparam_add_noise launch");
}

extern "C" int idf_pack_weights(const idf_pack_job* jobs_dev, const int32_t* cta_prefix_dev, int32_t njobs,
                                int32_t total_ctas, idf_stream_t stream) {
  if (!jobs_dev || !cta_prefix_dev || njobs <= 0 || total_ctas <= 0) return fail(IDF_ERR_ARG, "pack_weights: bad argument");
  pack_weights_kernel<<<total_ctas, 256, PACK_CHUNK * 9 * sizeof(float), S(stream)>>>(jobs_dev, cta_prefix_dev, njobs);
  return check_cuda(cudaGetLastError(), "pack_weights launch");
}

extern "C" int idf_attention_delta(const void* d_out, int64_t ld_do, const void* out, int64_t ld_o, int32_t M,
                                   int32_t heads, int32_t head_dim, float* delta, idf_stream_t stream) {
  if (!d_out || !out || !delta || head_dim % 8 != 0) return fail(IDF_ERR_ARG, "attention_delta: bad argument");
  attn_delta_kernel<<<(unsigned)(((long long)M * heads + 255) / 256), 256, 0, S(stream)>>>(BF(d_out), ld_do, BF(out), ld_o,
                                                                                           M, heads, head_dim, delta);
  return check_cuda(cudaGetLastError(), "attention_delta launch");
}

extern "C" int idf_f32_to_bf16_rows(const float* x, void* y, int64_t ldy, int64_t M, int32_t C, idf_stream_t stream) {
  if (!x || !y || C % 8 != 0) return fail(IDF_ERR_ARG, "f32_to_bf16_rows: bad argument");
  f32_to_bf16_rows_kernel<<<grid_for(M * (C / 8), 256), 256, 0, S(stream)>>>(x, BF(y), ldy, M, C);
  return check_cuda(cudaGetLastError(), "f32_to_bf16_rows launch");
}
