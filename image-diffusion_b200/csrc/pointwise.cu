// Bandwidth- and latency-bound pieces of the path: timestep/class embedding, the fused CFG + DDPM posterior
// step, the VQ nearest-code search, the tiny-channel edge convolutions and a few layout movers.
#include "common.cuh"
#include "host.h"
#include "../../include/idf_b200.h"

namespace idf {

__device__ __forceinline__ float silu_f(float x) { return x / (1.f + expf(-x)); }

// ---------------------------------------------------------------------------------------------
// embedding path
// ---------------------------------------------------------------------------------------------
// e[r, j] = sin(t_r / factor_j) for j < D/2, cos(t_r / factor_{j - D/2}) otherwise  (components.py:442-443)
__global__ void sincos_kernel(const int64_t* __restrict__ t, const float* __restrict__ factor,
                              float* __restrict__ e, int R, int D) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= R * D) return;
  const int r = i / D, j = i % D, half = D / 2;
  const float arg = (float)t[r] / factor[j < half ? j : j - half];
  e[i] = j < half ? sinf(arg) : cosf(arg);
}

// Y[r, j] = act_out( b[j] + sum_k W[j, k] * X[r, k] + mask[r] * cls[ctx[r], j] ), one warp per output column,
// rows in register chunks of 8 so each weight row is streamed once per chunk.
template <bool SILU_OUT>
__global__ void __launch_bounds__(256) linear_rows_kernel(const float* __restrict__ X, int ldx,
                                                          const float* __restrict__ W, const float* __restrict__ b,
                                                          float* __restrict__ Y, int ldy, int R, int K, int J,
                                                          const float* __restrict__ cls,
                                                          const int64_t* __restrict__ ctx,
                                                          const float* __restrict__ mask,
                                                          float* __restrict__ Ypre = nullptr) {
  const int j = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (j >= J) return;
  const float4* w4 = reinterpret_cast<const float4*>(W + (long long)j * K);
  // blockIdx.y walks the 8-row chunks in parallel (gridDim.y == 1: one CTA loops over all of them)
  for (int r0 = blockIdx.y * 8; r0 < R; r0 += gridDim.y * 8) {
    float acc[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) acc[i] = 0.f;
    for (int k4 = lane; k4 < K / 4; k4 += 32) {
      const float4 w = __ldg(w4 + k4);
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        if (r0 + i < R) {
          const float4 xv = *reinterpret_cast<const float4*>(X + (long long)(r0 + i) * ldx + k4 * 4);
          acc[i] = fmaf(w.x, xv.x, acc[i]);
          acc[i] = fmaf(w.y, xv.y, acc[i]);
          acc[i] = fmaf(w.z, xv.z, acc[i]);
          acc[i] = fmaf(w.w, xv.w, acc[i]);
        }
      }
    }
#pragma unroll
    for (int i = 0; i < 8; ++i)
      for (int o = 16; o > 0; o >>= 1) acc[i] += __shfl_xor_sync(0xffffffffu, acc[i], o);
    if (lane == 0) {
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const int r = r0 + i;
        if (r < R) {
          float v = acc[i] + b[j];
          if (cls != nullptr && ctx != nullptr) {
            const float mk = mask ? mask[r] : 1.f;
            v += mk * cls[(long long)ctx[r] * J + j];
          }
          if (Ypre != nullptr) Ypre[(long long)r * ldy + j] = v;
          Y[(long long)r * ldy + j] = SILU_OUT ? silu_f(v) : v;
        }
      }
    }
  }
}

// The same linear layer for MANY rows (the sampler's per-timestep table: R = num_steps * (classes + 1) = 4000 rows, 41 GFLOP
// in all): a shared-memory tiled fp32 GEMM, 64 x 64 outputs per CTA, 4 x 4 per thread. The warp-per-column kernel above
// streams each weight row once per 8 input rows - fine for a batch of rows, 2.4 TFLOP/s (15.5 ms per weight version) for
// the table. fp32 throughout: the time bias keeps the accuracy of the reference's fp32 embedding path.
constexpr int LT_BM = 64, LT_BN = 64, LT_BK = 16;
template <bool SILU_OUT>
__global__ void __launch_bounds__(256) linear_tiled_kernel(const float* __restrict__ X, int ldx,
                                                           const float* __restrict__ W, const float* __restrict__ b,
                                                           float* __restrict__ Y, int ldy, int R, int K, int J,
                                                           const float* __restrict__ cls,
                                                           const int64_t* __restrict__ ctx,
                                                           const float* __restrict__ mask) {
  __shared__ float As[LT_BK][LT_BM + 4];
  __shared__ float Bs[LT_BK][LT_BN + 4];
  const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;
  const int r0 = blockIdx.y * LT_BM, j0 = blockIdx.x * LT_BN;
  const int lrow = threadIdx.x >> 2, lk = (threadIdx.x & 3) * 4;  // this thread's float4 of each operand tile
  float acc[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;
  for (int k0 = 0; k0 < K; k0 += LT_BK) {
    float4 a = make_float4(0.f, 0.f, 0.f, 0.f), w = make_float4(0.f, 0.f, 0.f, 0.f);
    if (r0 + lrow < R) a = *reinterpret_cast<const float4*>(X + (long long)(r0 + lrow) * ldx + k0 + lk);
    if (j0 + lrow < J) w = __ldg(reinterpret_cast<const float4*>(W + (long long)(j0 + lrow) * K + k0 + lk));
    As[lk + 0][lrow] = a.x; As[lk + 1][lrow] = a.y; As[lk + 2][lrow] = a.z; As[lk + 3][lrow] = a.w;
    Bs[lk + 0][lrow] = w.x; Bs[lk + 1][lrow] = w.y; Bs[lk + 2][lrow] = w.z; Bs[lk + 3][lrow] = w.w;
    __syncthreads();
#pragma unroll
    for (int k = 0; k < LT_BK; ++k) {
      const float4 av = *reinterpret_cast<const float4*>(&As[k][ty * 4]);
      const float4 bv = *reinterpret_cast<const float4*>(&Bs[k][tx * 4]);
      const float aa[4] = {av.x, av.y, av.z, av.w}, bb[4] = {bv.x, bv.y, bv.z, bv.w};
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(aa[i], bb[j], acc[i][j]);
    }
    __syncthreads();
  }
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int r = r0 + ty * 4 + i;
    if (r >= R) continue;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int col = j0 + tx * 4 + j;
      if (col >= J) continue;
      float v = acc[i][j] + b[col];
      if (cls != nullptr && ctx != nullptr) v += (mask ? mask[r] : 1.f) * cls[(long long)ctx[r] * J + col];
      Y[(long long)r * ldy + col] = SILU_OUT ? silu_f(v) : v;
    }
  }
}

template <bool SILU_OUT>
static void launch_linear(const float* X, int ldx, const float* W, const float* b, float* Y, int ldy, int R, int K, int J,
                          const float* cls, const int64_t* ctx, const float* mask, cudaStream_t s) {
  if (R >= 256 && K % LT_BK == 0 && ldx % 4 == 0) {
    dim3 grid((J + LT_BN - 1) / LT_BN, (R + LT_BM - 1) / LT_BM);
    linear_tiled_kernel<SILU_OUT><<<grid, 256, 0, s>>>(X, ldx, W, b, Y, ldy, R, K, J, cls, ctx, mask);
  } else {
    const int wpb = 8;  // warps per block
    linear_rows_kernel<SILU_OUT><<<(J + wpb - 1) / wpb, 256, 0, s>>>(X, ldx, W, b, Y, ldy, R, K, J, cls, ctx, mask);
  }
}

// ---------------------------------------------------------------------------------------------
// CFG mix + DDPM ancestral step (diffusion.py:55, components.py:405-424)
// ---------------------------------------------------------------------------------------------
__global__ void cfg_posterior_kernel(const float* __restrict__ xt, const float* __restrict__ ec,
                                     const float* __restrict__ eu, const float* __restrict__ z,
                                     const float* __restrict__ cfg, const int64_t* __restrict__ t, int t_stride,
                                     const float* __restrict__ betas, const float* __restrict__ alphas,
                                     const float* __restrict__ acp, const float* __restrict__ sacp,
                                     const float* __restrict__ somacp, float* __restrict__ x_prev,
                                     float* __restrict__ x_prev_dup, float* __restrict__ x0_out, int N, int chw,
                                     int num_steps) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (long long)N * chw) return;
  const int n = (int)(i / chw);
  const int tn = (int)t[(long long)n * t_stride];
  const int t_first = (int)t[0];
  const float x = xt[i];
  const float u = eu[i];
  const float eps = u + cfg[n] * (ec[i] - u);
  const float b = betas[tn];
  const float so = somacp[tn];
  if (x0_out != nullptr) {
    float x0 = (x - so * eps) / sacp[tn];
    x0_out[i] = fminf(fmaxf(x0, -1.f), 1.f);
  }
  float mean = x - (b * eps) / so;
  mean = mean / sqrtf(alphas[tn]);
  float out = mean;
  if (t_first != 0) {
    // per-sample t with t[0] != 0 but t[n] == 0: the reference indexes alpha_cum_prod[t - 1] = [-1], which torch
    // wraps to the LAST table entry (components.py:419); same here instead of reading out of bounds
    const int tp = tn > 0 ? tn - 1 : num_steps - 1;
    float var = (1.f - acp[tp]) / (1.f - acp[tn]);
    var = var * b;
    out = mean + sqrtf(var) * z[i];
  }
  x_prev[i] = out;  // may alias xt: every thread reads its own element before writing it
  if (x_prev_dup != nullptr) x_prev_dup[i] = out;
}

// CFG mix + one step of the generalised (strided) sampler of Song et al., "Denoising Diffusion Implicit Models"
// (ICLR 2021), eq. 12, from timestep t to an arbitrary earlier timestep t_prev (t_prev < 0: the final step, abar = 1):
//   x0     = (x_t - sqrt(1 - abar_t) eps) / sqrt(abar_t)                     (optionally clamped to [-1, 1])
//   sigma  = eta * sqrt((1 - abar_prev) / (1 - abar_t)) * sqrt(1 - abar_t / abar_prev)
//   x_prev = sqrt(abar_prev) x0 + sqrt(1 - abar_prev - sigma^2) eps + sigma z
// eta = 0: deterministic DDIM; eta = 1 with t_prev = t - 1 and no clamp: the reference's ancestral DDPM step
// (components.py:405-424) in another algebraic form. Both timesteps are read on the device (graph replay safe).
__global__ void cfg_ddim_kernel(const float* __restrict__ xt, const float* __restrict__ ec,
                                const float* __restrict__ eu, const float* __restrict__ z,
                                const float* __restrict__ cfg, const int64_t* __restrict__ t,
                                const int64_t* __restrict__ t_prev, const float* __restrict__ acp, float eta,
                                int clamp_x0, float* __restrict__ x_prev, float* __restrict__ x0_out, int N, int chw) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (long long)N * chw) return;
  const int n = (int)(i / chw);
  const int tn = (int)t[0];
  const int tp = (int)t_prev[0];
  const float x = xt[i];
  const float u = eu[i];
  const float eps = u + cfg[n] * (ec[i] - u);
  const float a_t = acp[tn];
  const float a_p = tp >= 0 ? acp[tp] : 1.f;
  float x0 = (x - sqrtf(1.f - a_t) * eps) / sqrtf(a_t);
  if (clamp_x0) x0 = fminf(fmaxf(x0, -1.f), 1.f);
  if (x0_out != nullptr) x0_out[i] = x0;
  const float sigma = eta * sqrtf((1.f - a_p) / (1.f - a_t)) * sqrtf(fmaxf(1.f - a_t / a_p, 0.f));
  float out = sqrtf(a_p) * x0 + sqrtf(fmaxf(1.f - a_p - sigma * sigma, 0.f)) * eps;
  if (sigma > 0.f) out += sigma * z[i];
  x_prev[i] = out;  // may alias xt
}

__global__ void add_noise_kernel(const float* __restrict__ x, const float* __restrict__ noise,
                                 const int64_t* __restrict__ t, const float* __restrict__ sacp,
                                 const float* __restrict__ somacp, float* __restrict__ out, int N, int chw) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (long long)N * chw) return;
  const int tn = (int)t[i / chw];
  out[i] = sacp[tn] * x[i] + somacp[tn] * noise[i];
}

// ---------------------------------------------------------------------------------------------
// VQ nearest code: one warp per latent vector, codebook (augmented with its norms) in shared memory,
// lanes stride over the codes, warp-shuffle (distance, index) min-reduction with first-index tie break.
// ---------------------------------------------------------------------------------------------
// |v|^2 exactly as ATen's CUDA reduce kernel evaluates v.pow(2).sum(-1) for a tiny last dimension: squares are
// rounded separately (no FMA), lane i of a power-of-two-wide group sums elements i, i+w, ... and the group is
// folded with a halving tree. For DIM = 3 this is (v0^2 + v2^2) + v1^2, verified bit-for-bit on B200
// (tools/diag_cdist.py); other DIMs follow the same scheme but are unverified.
template <int DIM>
__device__ __forceinline__ float aten_sumsq(const float (&v)[DIM]) {
  constexpr int WIDTH = DIM >= 8 ? 8 : (DIM >= 4 ? 4 : (DIM >= 2 ? 2 : 1));
  float part[WIDTH];
#pragma unroll
  for (int i = 0; i < WIDTH; ++i) {
    part[i] = __fmul_rn(v[i], v[i]);
#pragma unroll
    for (int j = i + WIDTH; j < DIM; j += WIDTH) part[i] = __fadd_rn(part[i], __fmul_rn(v[j], v[j]));
  }
#pragma unroll
  for (int off = WIDTH / 2; off > 0; off >>= 1)
#pragma unroll
    for (int i = 0; i < off; ++i) part[i] = __fadd_rn(part[i], part[i + off]);
  return part[0];
}

template <int DIM>
__global__ void __launch_bounds__(256) vq_argmin_kernel(const float* __restrict__ z, const float* __restrict__ cb,
                                                        int64_t* __restrict__ idx_out, float* __restrict__ zq_out,
                                                        int rows, int size, int hw) {
  extern __shared__ float s_cb[];  // [size][DIM + 1]: code, |code|^2
  for (int i = threadIdx.x; i < size; i += blockDim.x) {
    float ev[DIM];
#pragma unroll
    for (int d = 0; d < DIM; ++d) {
      ev[d] = cb[(long long)i * DIM + d];
      s_cb[i * (DIM + 1) + d] = ev[d];
    }
    s_cb[i * (DIM + 1) + DIM] = aten_sumsq<DIM>(ev);
  }
  __syncthreads();
  const int lane = threadIdx.x & 31;
  const int warps_per_block = blockDim.x >> 5;
  for (int row = blockIdx.x * warps_per_block + (threadIdx.x >> 5); row < rows; row += gridDim.x * warps_per_block) {
    float xv[DIM];
    // hw > 0: z and zq are NCHW tensors with hw pixels per image (row = image * hw + pixel); else row-major rows
    const long long base = hw > 0 ? ((long long)(row / hw) * DIM * hw + row % hw) : (long long)row * DIM;
    const long long dstride = hw > 0 ? hw : 1;
#pragma unroll
    for (int d = 0; d < DIM; ++d) xv[d] = z[base + d * dstride];
    const float xn = aten_sumsq<DIM>(xv);
    float best = INFINITY;
    int best_i = 0x7fffffff;
    for (int i = lane; i < size; i += 32) {
      // x1_ . x2_ with x1_ = [-2x, |x|^2, 1], x2_ = [e, 1, |e|^2]; K = DIM + 2 products accumulated in order
      float acc = 0.f;
#pragma unroll
      for (int d = 0; d < DIM; ++d) acc = fmaf(-2.f * xv[d], s_cb[i * (DIM + 1) + d], acc);
      acc = fmaf(xn, 1.f, acc);
      acc = fmaf(1.f, s_cb[i * (DIM + 1) + DIM], acc);
      const float dist = sqrtf(fmaxf(acc, 0.f));
      if (dist < best) {  // strict: keeps the first minimal index within the lane's ascending walk
        best = dist;
        best_i = i;
      }
    }
    for (int o = 16; o > 0; o >>= 1) {
      const float ob = __shfl_xor_sync(0xffffffffu, best, o);
      const int oi = __shfl_xor_sync(0xffffffffu, best_i, o);
      if (ob < best || (ob == best && oi < best_i)) {
        best = ob;
        best_i = oi;
      }
    }
    if (lane == 0) idx_out[row] = (int64_t)best_i;
    if (zq_out != nullptr && lane < DIM) zq_out[base + lane * dstride] = s_cb[best_i * (DIM + 1) + lane];
  }
}

// ---------------------------------------------------------------------------------------------
// edge convolutions with a handful of channels on one side
// ---------------------------------------------------------------------------------------------
// fp32 NCHW (Cin <= 4) -> bf16 NHWC, 3x3 s1 p1. Thread (lane) owns 4 output channels and keeps their 4*9*CIN
// weights in registers; a warp covers 128 output channels of one pixel at a time: the 9*CIN patch values are
// warp-wide broadcast loads and the warp's store is one contiguous 256-byte row segment. With dup_rows > 0 the
// result is also written `dup_rows` rows further down (the unconditional half of a batch-doubled CFG input).
template <int CIN>
__global__ void __launch_bounds__(256) conv3x3_small_cin_kernel(const float* __restrict__ x,
                                                                const float* __restrict__ w,
                                                                const float* __restrict__ bias,
                                                                __nv_bfloat16* __restrict__ y, long long ldy, int B,
                                                                int H, int W, int Cout, long long dup_rows, int RB) {
  // CTA = (row block of RB <= 8 rows, image). The fp32 input rows (+halo, zero padded) are staged once in shared
  // memory; warp w walks output row h0 + w pixel by pixel with warp-wide broadcast reads of the 9*CIN patch.
  extern __shared__ float s_x[];  // [CIN][RB + 2][W + 2]
  constexpr int K = CIN * 9;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int b = blockIdx.y, h0 = blockIdx.x * RB;
  const int pw = W + 2;
  for (int i = threadIdx.x; i < CIN * (RB + 2) * pw; i += blockDim.x) {
    const int c = i / ((RB + 2) * pw), rem = i % ((RB + 2) * pw);
    const int hh = h0 + rem / pw - 1, ww = rem % pw - 1;
    s_x[i] = (hh >= 0 && hh < H && ww >= 0 && ww < W) ? x[(((long long)b * CIN + c) * H + hh) * W + ww] : 0.f;
  }
  __syncthreads();
  if (warp >= RB) return;
  const int hq = h0 + warp;
  for (int cg = 0; cg < Cout / 128; ++cg) {
    const int co = cg * 128 + lane * 4;
    float wr[4][K];
    float bs[4];
#pragma unroll
    for (int o = 0; o < 4; ++o) {
      bs[o] = bias ? bias[co + o] : 0.f;
#pragma unroll
      for (int k = 0; k < K; ++k) wr[o][k] = w[(long long)(co + o) * K + k];
    }
    const long long row_pix = ((long long)b * H + hq) * W;
#pragma unroll 2
    for (int wq = 0; wq < W; ++wq) {
      float acc[4] = {bs[0], bs[1], bs[2], bs[3]};
#pragma unroll
      for (int c = 0; c < CIN; ++c)
#pragma unroll
        for (int kh = 0; kh < 3; ++kh)
#pragma unroll
          for (int kw = 0; kw < 3; ++kw) {
            const float v = s_x[(c * (RB + 2) + warp + kh) * pw + wq + kw];
#pragma unroll
            for (int o = 0; o < 4; ++o) acc[o] = fmaf(v, wr[o][c * 9 + kh * 3 + kw], acc[o]);
          }
      uint2 o2;
      o2.x = pack_bf16x2(acc[0], acc[1]);
      o2.y = pack_bf16x2(acc[2], acc[3]);
      *reinterpret_cast<uint2*>(y + (row_pix + wq) * ldy + co) = o2;
      if (dup_rows > 0) *reinterpret_cast<uint2*>(y + (row_pix + wq + dup_rows) * ldy + co) = o2;
    }
  }
}

// bf16 NHWC -> fp32 NCHW (Cout <= 8), 3x3 s1 p1. One warp = one pixel; lanes stride over (tap, 8-channel vector).
template <int COUT>
__global__ void __launch_bounds__(256) conv3x3_small_cout_kernel(const __nv_bfloat16* __restrict__ x, long long ldx,
                                                                 const float* __restrict__ w,
                                                                 const float* __restrict__ bias,
                                                                 float* __restrict__ y, int B, int Cin, int H,
                                                                 int W) {
  extern __shared__ float s_w[];  // [9][Cin][COUT]
  for (int i = threadIdx.x; i < COUT * Cin * 9; i += blockDim.x) {
    const int o = i / (Cin * 9), rem = i % (Cin * 9), c = rem / 9, tap = rem % 9;  // OIHW
    s_w[(tap * Cin + c) * COUT + o] = w[i];
  }
  __syncthreads();
  const int lane = threadIdx.x & 31;
  const int vec_per_tap = Cin / 8;
  const int items = 9 * vec_per_tap;
  const long long total = (long long)B * H * W;
  const int warps_per_block = blockDim.x >> 5;
  for (long long pix = (long long)blockIdx.x * warps_per_block + (threadIdx.x >> 5); pix < total;
       pix += (long long)gridDim.x * warps_per_block) {
    const int wq = (int)(pix % W);
    const int hq = (int)((pix / W) % H);
    const int b = (int)(pix / ((long long)W * H));
    float acc[COUT];
#pragma unroll
    for (int o = 0; o < COUT; ++o) acc[o] = 0.f;
    for (int it = lane; it < items; it += 32) {
      const int tap = it / vec_per_tap, vv = it % vec_per_tap;
      const int hh = hq + tap / 3 - 1, ww = wq + tap % 3 - 1;
      if (hh < 0 || hh >= H || ww < 0 || ww >= W) continue;
      const uint4 raw =
          *reinterpret_cast<const uint4*>(x + (((long long)b * H + hh) * W + ww) * ldx + vv * 8);
      float f[8];
      f[0] = bf16_lo(raw.x); f[1] = bf16_hi(raw.x); f[2] = bf16_lo(raw.y); f[3] = bf16_hi(raw.y);
      f[4] = bf16_lo(raw.z); f[5] = bf16_hi(raw.z); f[6] = bf16_lo(raw.w); f[7] = bf16_hi(raw.w);
      const float* wp = &s_w[(tap * Cin + vv * 8) * COUT];
#pragma unroll
      for (int e = 0; e < 8; ++e)
#pragma unroll
        for (int o = 0; o < COUT; ++o) acc[o] = fmaf(f[e], wp[e * COUT + o], acc[o]);
    }
#pragma unroll
    for (int o = 0; o < COUT; ++o)
      for (int s = 16; s > 0; s >>= 1) acc[o] += __shfl_xor_sync(0xffffffffu, acc[o], s);
    if (lane == 0) {
#pragma unroll
      for (int o = 0; o < COUT; ++o)
        y[(((long long)b * COUT + o) * H + hq) * W + wq] = acc[o] + (bias ? bias[o] : 0.f);
    }
  }
}


// Tiled variant: one CTA = a (TH x TW) pixel tile of one image (TH*TW = 128 threads, one output pixel each). The
// (TH+2) x (TW+2) input halo tile is staged once in shared memory (rows padded by 16 bytes so that the per-pixel
// 16-byte reads of a warp fall in distinct banks); weights are read as warp-wide broadcasts.
template <int COUT>
__global__ void __launch_bounds__(128) conv3x3_small_cout_tiled_kernel(const __nv_bfloat16* __restrict__ x,
                                                                       long long ldx, const float* __restrict__ w,
                                                                       const float* __restrict__ bias,
                                                                       float* __restrict__ y, int Cin, int H, int W,
                                                                       int TH, int TW) {
  extern __shared__ __align__(16) uint8_t s_raw[];
  const int pix_bytes = Cin * 2 + 16;
  float* s_w = reinterpret_cast<float*>(s_raw);                       // [9][Cin][COUT]
  uint8_t* s_in = s_raw + ((9 * Cin * COUT * 4 + 15) & ~15);          // [(TH+2)*(TW+2)][pix_bytes]
  for (int i = threadIdx.x; i < COUT * Cin * 9; i += blockDim.x) {
    const int o = i / (Cin * 9), rem = i % (Cin * 9), c = rem / 9, tap = rem % 9;  // OIHW
    s_w[(tap * Cin + c) * COUT + o] = w[i];
  }
  const int tiles_w = W / TW;
  const int h0 = (blockIdx.x / tiles_w) * TH, w0 = (blockIdx.x % tiles_w) * TW;
  const int b = blockIdx.y;
  const int vec = Cin / 8;
  const int halo_w = TW + 2;
  const int nvec = (TH + 2) * halo_w * vec;
  for (int i0 = threadIdx.x; i0 < nvec; i0 += 8 * blockDim.x) {  // 8 independent 16-byte loads in flight per thread
    uint4 val[8];
#pragma unroll
    for (int u = 0; u < 8; ++u) {
      const int i = i0 + u * blockDim.x;
      val[u] = make_uint4(0u, 0u, 0u, 0u);
      if (i < nvec) {
        const int v = i % vec, pp = i / vec;
        const int hh = h0 + pp / halo_w - 1, ww = w0 + pp % halo_w - 1;
        if (hh >= 0 && hh < H && ww >= 0 && ww < W)
          val[u] = *reinterpret_cast<const uint4*>(x + (((long long)b * H + hh) * W + ww) * ldx + v * 8);
      }
    }
#pragma unroll
    for (int u = 0; u < 8; ++u) {
      const int i = i0 + u * blockDim.x;
      if (i < nvec) *reinterpret_cast<uint4*>(s_in + (i / vec) * pix_bytes + (i % vec) * 16) = val[u];
    }
  }
  __syncthreads();
  const int r = threadIdx.x / TW, c = threadIdx.x % TW;
  float acc[COUT];
#pragma unroll
  for (int o = 0; o < COUT; ++o) acc[o] = 0.f;
#pragma unroll 1
  for (int tap = 0; tap < 9; ++tap) {
    const uint8_t* src = s_in + ((r + tap / 3) * halo_w + (c + tap % 3)) * pix_bytes;
    const float* wt = s_w + tap * Cin * COUT;
    for (int v = 0; v < vec; ++v) {
      const uint4 raw = *reinterpret_cast<const uint4*>(src + v * 16);
      float f[8];
      f[0] = bf16_lo(raw.x); f[1] = bf16_hi(raw.x); f[2] = bf16_lo(raw.y); f[3] = bf16_hi(raw.y);
      f[4] = bf16_lo(raw.z); f[5] = bf16_hi(raw.z); f[6] = bf16_lo(raw.w); f[7] = bf16_hi(raw.w);
      const float* wp = wt + v * 8 * COUT;
#pragma unroll
      for (int e = 0; e < 8; ++e)
#pragma unroll
        for (int o = 0; o < COUT; ++o) acc[o] = fmaf(f[e], wp[e * COUT + o], acc[o]);
    }
  }
#pragma unroll
  for (int o = 0; o < COUT; ++o)
    y[(((long long)b * COUT + o) * H + (h0 + r)) * W + (w0 + c)] = acc[o] + (bias ? bias[o] : 0.f);
}

// Cin == 128 fast path: CTA = (TH x TW) pixel tile, warp w walks tile row w; lane l owns input channels 4l..4l+3
// and keeps their 9 * 4 * COUT weights in registers. Per pixel: nine 8-byte shared-memory reads (one contiguous
// 256-byte pixel per tap across the warp), 36 * COUT FMAs, and a 5-step shuffle reduction of the COUT partials.
template <int COUT>
__global__ void __launch_bounds__(128) conv3x3_small_cout_c128_kernel(const __nv_bfloat16* __restrict__ x,
                                                                      long long ldx, const float* __restrict__ w,
                                                                      const float* __restrict__ bias,
                                                                      float* __restrict__ y, int H, int W, int TH,
                                                                      int TW) {
  constexpr int CIN = 128;
  extern __shared__ __align__(16) uint8_t s_in[];  // [(TH+2)*(TW+2)][CIN*2 + 16]
  constexpr int pix_bytes = CIN * 2 + 16;
  const int tiles_w = W / TW;
  const int h0 = (blockIdx.x / tiles_w) * TH, w0 = (blockIdx.x % tiles_w) * TW;
  const int b = blockIdx.y;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  constexpr int vec = CIN / 8;
  const int halo_w = TW + 2;
  const int nvec = (TH + 2) * halo_w * vec;
  for (int i0 = threadIdx.x; i0 < nvec; i0 += 8 * blockDim.x) {
    uint4 val[8];
#pragma unroll
    for (int u = 0; u < 8; ++u) {
      const int i = i0 + u * blockDim.x;
      val[u] = make_uint4(0u, 0u, 0u, 0u);
      if (i < nvec) {
        const int v = i % vec, pp = i / vec;
        const int hh = h0 + pp / halo_w - 1, ww = w0 + pp % halo_w - 1;
        if (hh >= 0 && hh < H && ww >= 0 && ww < W)
          val[u] = *reinterpret_cast<const uint4*>(x + (((long long)b * H + hh) * W + ww) * ldx + v * 8);
      }
    }
#pragma unroll
    for (int u = 0; u < 8; ++u) {
      const int i = i0 + u * blockDim.x;
      if (i < nvec) *reinterpret_cast<uint4*>(s_in + (i / vec) * pix_bytes + (i % vec) * 16) = val[u];
    }
  }
  float wr[COUT][9][4];  // OIHW: w[(o*CIN + c)*9 + tap]
#pragma unroll
  for (int o = 0; o < COUT; ++o)
#pragma unroll
    for (int tap = 0; tap < 9; ++tap)
#pragma unroll
      for (int e = 0; e < 4; ++e) wr[o][tap][e] = w[((long long)o * CIN + lane * 4 + e) * 9 + tap];
  __syncthreads();
  // two pixels per iteration: independent FMA chains and shuffle trees overlap (one pixel at a time left the warp
  // waiting on its own dependent chain: ncu 0.49 issued instructions per scheduler-cycle with 2.6 warps each)
  for (int r = warp; r < TH; r += 4) {
    for (int c = 0; c < TW; c += 2) {
      float acc[2][COUT];
#pragma unroll
      for (int o = 0; o < COUT; ++o) { acc[0][o] = 0.f; acc[1][o] = 0.f; }
#pragma unroll
      for (int tap = 0; tap < 9; ++tap) {
        const uint8_t* src = s_in + ((r + tap / 3) * halo_w + (c + tap % 3)) * pix_bytes + lane * 8;
        const uint2 raw0 = *reinterpret_cast<const uint2*>(src);
        const uint2 raw1 = *reinterpret_cast<const uint2*>(src + pix_bytes);
        const float f0 = bf16_lo(raw0.x), f1 = bf16_hi(raw0.x), f2 = bf16_lo(raw0.y), f3 = bf16_hi(raw0.y);
        const float g0 = bf16_lo(raw1.x), g1 = bf16_hi(raw1.x), g2 = bf16_lo(raw1.y), g3 = bf16_hi(raw1.y);
#pragma unroll
        for (int o = 0; o < COUT; ++o) {
          acc[0][o] = fmaf(f0, wr[o][tap][0], acc[0][o]);
          acc[1][o] = fmaf(g0, wr[o][tap][0], acc[1][o]);
          acc[0][o] = fmaf(f1, wr[o][tap][1], acc[0][o]);
          acc[1][o] = fmaf(g1, wr[o][tap][1], acc[1][o]);
          acc[0][o] = fmaf(f2, wr[o][tap][2], acc[0][o]);
          acc[1][o] = fmaf(g2, wr[o][tap][2], acc[1][o]);
          acc[0][o] = fmaf(f3, wr[o][tap][3], acc[0][o]);
          acc[1][o] = fmaf(g3, wr[o][tap][3], acc[1][o]);
        }
      }
#pragma unroll
      for (int px = 0; px < 2; ++px)
#pragma unroll
        for (int o = 0; o < COUT; ++o)
#pragma unroll
          for (int sft = 16; sft > 0; sft >>= 1) acc[px][o] += __shfl_xor_sync(0xffffffffu, acc[px][o], sft);
      if (lane < 2 * COUT) {  // lanes [0, COUT): pixel c, lanes [COUT, 2 COUT): pixel c + 1
        const int px = lane >= COUT ? 1 : 0, o_sel = lane - px * COUT;
        float v = 0.f;
#pragma unroll
        for (int o = 0; o < COUT; ++o) v = (o_sel == o) ? (px ? acc[1][o] : acc[0][o]) : v;
        if (c + px < TW)
          y[(((long long)b * COUT + o_sel) * H + (h0 + r)) * W + (w0 + c + px)] = v + (bias ? bias[o_sel] : 0.f);
      }
    }
  }
}

// fp32 NCHW 1x1 convolution with tiny channel counts (decoder's first conv, encoder's last conv).
__global__ void conv1x1_small_kernel(const float* __restrict__ x, const float* __restrict__ w,
                                     const float* __restrict__ bias, float* __restrict__ y, int B, int Cin, int Cout,
                                     int HW) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (long long)B * Cout * HW) return;
  const int p = (int)(i % HW);
  const int o = (int)((i / HW) % Cout);
  const int b = (int)(i / ((long long)HW * Cout));
  float acc = 0.f;
  for (int c = 0; c < Cin; ++c) acc = fmaf(x[((long long)b * Cin + c) * HW + p], w[o * Cin + c], acc);
  y[i] = acc + (bias ? bias[o] : 0.f);
}

// ---------------------------------------------------------------------------------------------
// layout movers (16-byte vectors over channels-last bf16)
// ---------------------------------------------------------------------------------------------
__global__ void upsample2x_kernel(const __nv_bfloat16* __restrict__ x, long long ldx, __nv_bfloat16* __restrict__ y,
                                  long long ldy, int B, int H, int W, int C) {
  const int vec = C / 8;
  const long long total = (long long)B * (2 * H) * (2 * W) * vec;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const int v = (int)(i % vec);
    const long long opix = i / vec;
    const int ow = (int)(opix % (2 * W));
    const int oh = (int)((opix / (2 * W)) % (2 * H));
    const int b = (int)(opix / ((long long)4 * W * H));
    const long long ipix = ((long long)b * H + (oh >> 1)) * W + (ow >> 1);
    *reinterpret_cast<uint4*>(y + opix * ldy + v * 8) = *reinterpret_cast<const uint4*>(x + ipix * ldx + v * 8);
  }
}

// y[(ph*2+pw)*B + b][h][w][:] = x[b][2h+ph][2w+pw][:]  (16-byte vectors, four independent copies per thread)
__global__ void __launch_bounds__(256) space_to_depth2_kernel(const __nv_bfloat16* __restrict__ x, long long ldx,
                                                              __nv_bfloat16* __restrict__ y, int B, int H, int W,
                                                              int C) {
  const int OH = H / 2, OW = W / 2, vec = C / 8;
  const long long total = (long long)B * OH * OW * vec;  // one thread = one output pixel vector of all 4 planes
  const long long plane = (long long)B * OH * OW * C;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const int v = (int)(i % vec);
    const long long opix = i / vec;
    const int ow = (int)(opix % OW);
    const int oh = (int)((opix / OW) % OH);
    const int b = (int)(opix / ((long long)OW * OH));
    const __nv_bfloat16* src = x + (((long long)b * H + 2 * oh) * W + 2 * ow) * ldx + v * 8;
    const uint4 v00 = *reinterpret_cast<const uint4*>(src);
    const uint4 v01 = *reinterpret_cast<const uint4*>(src + ldx);
    const uint4 v10 = *reinterpret_cast<const uint4*>(src + (long long)W * ldx);
    const uint4 v11 = *reinterpret_cast<const uint4*>(src + (long long)W * ldx + ldx);
    __nv_bfloat16* dst = y + opix * C + v * 8;
    *reinterpret_cast<uint4*>(dst) = v00;
    *reinterpret_cast<uint4*>(dst + plane) = v01;
    *reinterpret_cast<uint4*>(dst + 2 * plane) = v10;
    *reinterpret_cast<uint4*>(dst + 3 * plane) = v11;
  }
}

__global__ void nchw_to_nhwc_kernel(const float* __restrict__ x, __nv_bfloat16* __restrict__ y, long long ldy, int B,
                                    int C, int HW) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (long long)B * C * HW) return;
  const int c = (int)(i % C);
  const long long pix = i / C;
  const int p = (int)(pix % HW);
  const int b = (int)(pix / HW);
  y[pix * ldy + c] = __float2bfloat16_rn(x[((long long)b * C + c) * HW + p]);
}

__global__ void nhwc_to_nchw_kernel(const __nv_bfloat16* __restrict__ x, long long ldx, float* __restrict__ y, int B,
                                    int C, int HW) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (long long)B * C * HW) return;
  const int p = (int)(i % HW);
  const int c = (int)((i / HW) % C);
  const int b = (int)(i / ((long long)HW * C));
  y[i] = __bfloat162float(x[((long long)b * HW + p) * ldx + c]);
}

// ---------------------------------------------------------------------------------------------------------------
// Codebook.forward, eval mode, after the argmin (components.py:301-313): commitment loss beta * mean((z_q - z)^2), the
// straight-through output z + (z_q - z) (same arithmetic as the reference, so the same bits), code usage histogram and
// perplexity exp(-sum p log(p + 1e-6)). Integer histogram (exact, order independent); the squared-error sum goes
// through per-CTA partials added in CTA order: bit-deterministic.
// ---------------------------------------------------------------------------------------------------------------
constexpr int VQS_MAX_CODES = 4096;
__global__ void __launch_bounds__(256) vq_stats_kernel(const float* __restrict__ z, const float* __restrict__ zq,
                                                       const int64_t* __restrict__ idx, float* __restrict__ st_out,
                                                       long long elems, int rows, int size, int* __restrict__ counts,
                                                       float* __restrict__ part) {
  extern __shared__ int s_hist[];  // [size]
  __shared__ float s_red[256];
  for (int i = threadIdx.x; i < size; i += blockDim.x) s_hist[i] = 0;
  __syncthreads();
  for (long long r = (long long)blockIdx.x * blockDim.x + threadIdx.x; r < rows; r += (long long)gridDim.x * blockDim.x)
    atomicAdd(&s_hist[(int)idx[r]], 1);
  float acc = 0.f;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < elems; i += (long long)gridDim.x * blockDim.x) {
    const float a = z[i], d = zq[i] - a;
    acc = fmaf(d, d, acc);
    st_out[i] = a + d;
  }
  s_red[threadIdx.x] = acc;
  __syncthreads();
  for (int o = 128; o > 0; o >>= 1) {
    if ((int)threadIdx.x < o) s_red[threadIdx.x] += s_red[threadIdx.x + o];
    __syncthreads();
  }
  if (threadIdx.x == 0) part[blockIdx.x] = s_red[0];
  for (int i = threadIdx.x; i < size; i += blockDim.x)
    if (s_hist[i] != 0) atomicAdd(&counts[i], s_hist[i]);
}

__global__ void __launch_bounds__(1024) vq_stats_finish_kernel(const int* __restrict__ counts, const float* __restrict__ part,
                                                               int nparts, int size, int rows, long long elems, float beta,
                                                               float* __restrict__ loss, float* __restrict__ perplexity) {
  __shared__ float s_red[1024];
  float e = 0.f;
  for (int i = threadIdx.x; i < size; i += blockDim.x) {
    const float pr = (float)counts[i] / (float)rows;
    e += pr * logf(pr + 1e-6f);
  }
  s_red[threadIdx.x] = e;
  __syncthreads();
  for (int o = 512; o > 0; o >>= 1) {
    if ((int)threadIdx.x < o) s_red[threadIdx.x] += s_red[threadIdx.x + o];
    __syncthreads();
  }
  if (threadIdx.x == 0) {
    *perplexity = expf(-s_red[0]);
    float sse = 0.f;
    for (int i = 0; i < nparts; ++i) sse += part[i];
    *loss = beta * (sse / (float)elems);
  }
}

// ---------------------------------------------------------------------------------------------------------------
// KL bottleneck of VAE.encode (vae.py:99-113): z6 = (mean || log_var) per sample; log_var clamped to [-30, 20];
// kl[b] = -0.5 * sum(1 + lv - mean^2 - exp(lv)) over the sample, loss = mean_b kl[b]; optional reparametrised
// sample z = mean + noise * exp(0.5 lv). One CTA per sample, fixed-order reductions (deterministic).
// ---------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) kl_stats_kernel(const float* __restrict__ z6, const float* __restrict__ noise,
                                                       float* __restrict__ z_out, float* __restrict__ kl_per_sample,
                                                       int half) {
  __shared__ float s_red[256];
  const int b = blockIdx.x;
  const float* mean = z6 + (long long)b * 2 * half;
  const float* lvp = mean + half;
  float acc = 0.f;
  for (int i = threadIdx.x; i < half; i += blockDim.x) {
    const float m = mean[i];
    const float lv = fminf(fmaxf(lvp[i], -30.f), 20.f);
    acc += 1.f + lv - m * m - expf(lv);
    if (z_out != nullptr) z_out[(long long)b * half + i] = m + noise[(long long)b * half + i] * expf(0.5f * lv);
  }
  s_red[threadIdx.x] = acc;
  __syncthreads();
  for (int o = 128; o > 0; o >>= 1) {
    if ((int)threadIdx.x < o) s_red[threadIdx.x] += s_red[threadIdx.x + o];
    __syncthreads();
  }
  if (threadIdx.x == 0) kl_per_sample[b] = -0.5f * s_red[0];
}

__global__ void kl_mean_kernel(const float* __restrict__ kl_per_sample, int B, float* __restrict__ loss) {
  if (threadIdx.x == 0 && blockIdx.x == 0) {
    float a = 0.f;
    for (int b = 0; b < B; ++b) a += kl_per_sample[b];
    *loss = a / (float)B;
  }
}

static inline unsigned blocks_for(long long n, int threads, int cap = 148 * 16) {
  long long b = (n + threads - 1) / threads;
  if (b > cap) b = cap;
  if (b < 1) b = 1;
  return (unsigned)b;
}

}  // namespace idf

using namespace idf;

extern "C" int idf_embed_time_class(const int64_t* t, const int64_t* ctx, const float* ctx_mask, int32_t R,
                                    int32_t D, const float* factor, const float* w1, const float* b1, const float* w2,
                                    const float* b2, const float* class_w, const float* wp, const float* bp,
                                    int32_t P, float* out, float* scratch, idf_stream_t stream) {
  if (!t || !factor || !w1 || !b1 || !w2 || !b2 || !wp || !bp || !out || !scratch)
    return fail(IDF_ERR_ARG, "embed: null pointer");
  if (R <= 0 || D <= 0 || D % 8 != 0 || P <= 0) return fail(IDF_ERR_ARG, "embed: bad shape");
  if (ctx != nullptr && class_w == nullptr) return fail(IDF_ERR_ARG, "embed: ctx without class_w");
  cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
  float* e = scratch;               // (R, D)    sin/cos, later silu(temb)
  float* h1 = scratch + (long long)R * D;  // (R, 4D)
  sincos_kernel<<<(R * D + 255) / 256, 256, 0, s>>>(t, factor, e, R, D);
  launch_linear<true>(e, D, w1, b1, h1, 4 * D, R, D, 4 * D, nullptr, nullptr, nullptr, s);
  // temb = Linear2(h1) + mask * class_w[ctx]; every consumer applies SiLU first (components.py:486), so store that
  launch_linear<true>(h1, 4 * D, w2, b2, e, D, R, 4 * D, D, class_w, ctx, ctx_mask, s);
  launch_linear<false>(e, D, wp, bp, out, P, R, D, P, nullptr, nullptr, nullptr, s);
  return check_cuda(cudaGetLastError(), "embed launch");
}

// Training variant: same arithmetic, but every intermediate the backward pass needs is kept:
// saved = e [R, D] | z1 [R, 4D] (pre-activation) | a1 = silu(z1) [R, 4D] | temb [R, D] (pre-activation) | s = silu(temb) [R, D]
extern "C" int idf_embed_time_class_train(const int64_t* t, const int64_t* ctx, const float* ctx_mask, int32_t R,
                                          int32_t D, const float* factor, const float* w1, const float* b1,
                                          const float* w2, const float* b2, const float* class_w, const float* wp,
                                          const float* bp, int32_t P, float* out, float* saved, idf_stream_t stream) {
  if (!t || !factor || !w1 || !b1 || !w2 || !b2 || !wp || !bp || !out || !saved)
    return fail(IDF_ERR_ARG, "embed_train: null pointer");
  if (R <= 0 || D <= 0 || D % 8 != 0 || P <= 0) return fail(IDF_ERR_ARG, "embed_train: bad shape");
  if (ctx != nullptr && class_w == nullptr) return fail(IDF_ERR_ARG, "embed_train: ctx without class_w");
  cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
  float* e = saved;
  float* z1 = e + (long long)R * D;
  float* a1 = z1 + (long long)R * 4 * D;
  float* temb = a1 + (long long)R * 4 * D;
  float* sv = temb + (long long)R * D;
  sincos_kernel<<<(R * D + 255) / 256, 256, 0, s>>>(t, factor, e, R, D);
  const int wpb = 8;
  const int rc = (R + 7) / 8;  // row chunks run as separate CTAs (R = batch size here, not a handful of rows)
  linear_rows_kernel<true><<<dim3((4 * D + wpb - 1) / wpb, rc), 256, 0, s>>>(e, D, w1, b1, a1, 4 * D, R, D, 4 * D, nullptr,
                                                                             nullptr, nullptr, z1);
  linear_rows_kernel<true><<<dim3((D + wpb - 1) / wpb, rc), 256, 0, s>>>(a1, 4 * D, w2, b2, sv, D, R, 4 * D, D, class_w, ctx,
                                                                         ctx_mask, temb);
  linear_rows_kernel<false><<<dim3((P + wpb - 1) / wpb, rc), 256, 0, s>>>(sv, D, wp, bp, out, P, R, D, P, nullptr, nullptr,
                                                                          nullptr);
  return check_cuda(cudaGetLastError(), "embed_train launch");
}

extern "C" int idf_cfg_posterior_step(const float* xt, const float* eps_cond, const float* eps_uncond,
                                      const float* noise, const float* cfg, const int64_t* t, int32_t t_stride,
                                      const float* betas, const float* alphas, const float* alpha_cum_prod,
                                      const float* sqrt_alpha_cum_prod, const float* sqrt_one_minus_alpha_cum_prod,
                                      float* x_prev, float* x_prev_dup, float* x0_out, int32_t N, int32_t chw,
                                      int32_t num_steps, idf_stream_t stream) {
  if (!xt || !eps_cond || !eps_uncond || !noise || !cfg || !t || !betas || !alphas || !alpha_cum_prod ||
      !sqrt_alpha_cum_prod || !sqrt_one_minus_alpha_cum_prod || !x_prev)
    return fail(IDF_ERR_ARG, "cfg_posterior: null pointer");
  if (num_steps < 1) return fail(IDF_ERR_ARG, "cfg_posterior: num_steps must be the schedule length");
  const long long n = (long long)N * chw;
  cfg_posterior_kernel<<<(unsigned)((n + 255) / 256), 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(
      xt, eps_cond, eps_uncond, noise, cfg, t, t_stride, betas, alphas, alpha_cum_prod, sqrt_alpha_cum_prod,
      sqrt_one_minus_alpha_cum_prod, x_prev, x_prev_dup, x0_out, N, chw, num_steps);
  return check_cuda(cudaGetLastError(), "cfg_posterior launch");
}

extern "C" int idf_cfg_ddim_step(const float* xt, const float* eps_cond, const float* eps_uncond, const float* noise,
                                 const float* cfg, const int64_t* t, const int64_t* t_prev,
                                 const float* alpha_cum_prod, float eta, int32_t clamp_x0, float* x_prev,
                                 float* x0_out, int32_t N, int32_t chw, idf_stream_t stream) {
  if (!xt || !eps_cond || !eps_uncond || !cfg || !t || !t_prev || !alpha_cum_prod || !x_prev)
    return fail(IDF_ERR_ARG, "cfg_ddim: null pointer");
  if (eta < 0.f || (eta > 0.f && !noise)) return fail(IDF_ERR_ARG, "cfg_ddim: eta > 0 needs a noise tensor");
  const long long n = (long long)N * chw;
  cfg_ddim_kernel<<<(unsigned)((n + 255) / 256), 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(
      xt, eps_cond, eps_uncond, noise, cfg, t, t_prev, alpha_cum_prod, eta, clamp_x0, x_prev, x0_out, N, chw);
  return check_cuda(cudaGetLastError(), "cfg_ddim launch");
}

extern "C" int idf_add_noise(const float* x, const float* noise, const int64_t* t, const float* sqrt_alpha_cum_prod,
                             const float* sqrt_one_minus_alpha_cum_prod, float* out, int32_t N, int32_t chw,
                             idf_stream_t stream) {
  if (!x || !noise || !t || !out) return fail(IDF_ERR_ARG, "add_noise: null pointer");
  const long long n = (long long)N * chw;
  add_noise_kernel<<<(unsigned)((n + 255) / 256), 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(
      x, noise, t, sqrt_alpha_cum_prod, sqrt_one_minus_alpha_cum_prod, out, N, chw);
  return check_cuda(cudaGetLastError(), "add_noise launch");
}

extern "C" int idf_vq_argmin(const float* z, const float* codebook, int64_t* idx_out, float* zq_out, int32_t rows,
                             int32_t dim, int32_t size, int32_t nchw_hw, idf_stream_t stream) {
  if (!z || !codebook || !idx_out) return fail(IDF_ERR_ARG, "vq_argmin: null pointer");
  if (rows <= 0 || size <= 0) return fail(IDF_ERR_ARG, "vq_argmin: bad shape");
  const int smem = size * (dim + 1) * 4;
  if (smem > 48 * 1024) return fail(IDF_ERR_UNSUPPORTED, "vq_argmin: codebook does not fit shared memory");
  const unsigned grid = blocks_for((long long)rows, 8, 148 * 8);
  cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
  switch (dim) {
    case 3: vq_argmin_kernel<3><<<grid, 256, smem, s>>>(z, codebook, idx_out, zq_out, rows, size, nchw_hw); break;
    case 4: vq_argmin_kernel<4><<<grid, 256, smem, s>>>(z, codebook, idx_out, zq_out, rows, size, nchw_hw); break;
    case 8: vq_argmin_kernel<8><<<grid, 256, smem, s>>>(z, codebook, idx_out, zq_out, rows, size, nchw_hw); break;
    default: return fail(IDF_ERR_UNSUPPORTED, "vq_argmin: dim %d not in {3,4,8}", dim);
  }
  return check_cuda(cudaGetLastError(), "vq_argmin launch");
}

extern "C" int idf_vq_loss_perplexity(const float* z, const float* zq, const int64_t* idx, float* quant_out,
                                      int32_t rows, int32_t dim, int32_t size, float beta, float* loss,
                                      float* perplexity, void* ws, int64_t ws_bytes, idf_stream_t stream) {
  if (!z || !zq || !idx || !quant_out || !loss || !perplexity || !ws) return fail(IDF_ERR_ARG, "vq_loss_perplexity: null pointer");
  if (rows <= 0 || dim <= 0 || size <= 0 || size > VQS_MAX_CODES) return fail(IDF_ERR_ARG, "vq_loss_perplexity: bad shape");
  const long long elems = (long long)rows * dim;
  const int grid = (int)blocks_for(elems, 256, 148 * 4);
  if ((long long)size * 4 + (long long)grid * 4 > ws_bytes)
    return fail(IDF_ERR_ARG, "vq_loss_perplexity: workspace needs %lld bytes", (long long)size * 4 + (long long)grid * 4);
  int* counts = reinterpret_cast<int*>(ws);
  float* part = reinterpret_cast<float*>(counts + size);
  cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
  int rc = check_cuda(cudaMemsetAsync(counts, 0, (size_t)size * 4, s), "vq_loss_perplexity: memset");
  if (rc != IDF_OK) return rc;
  vq_stats_kernel<<<grid, 256, size * 4, s>>>(z, zq, idx, quant_out, elems, rows, size, counts, part);
  vq_stats_finish_kernel<<<1, 1024, 0, s>>>(counts, part, grid, size, rows, elems, beta, loss, perplexity);
  return check_cuda(cudaGetLastError(), "vq_loss_perplexity launch");
}

extern "C" int idf_kl_loss_reparam(const float* z6, const float* noise, float* z_out, float* kl_per_sample, float* loss,
                                   int32_t B, int32_t half, idf_stream_t stream) {
  if (!z6 || !kl_per_sample || !loss || B <= 0 || half <= 0) return fail(IDF_ERR_ARG, "kl_loss_reparam: bad argument");
  if ((z_out != nullptr) != (noise != nullptr)) return fail(IDF_ERR_ARG, "kl_loss_reparam: z_out and noise go together");
  cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
  kl_stats_kernel<<<B, 256, 0, s>>>(z6, noise, z_out, kl_per_sample, half);
  kl_mean_kernel<<<1, 32, 0, s>>>(kl_per_sample, B, loss);
  return check_cuda(cudaGetLastError(), "kl_loss_reparam launch");
}

extern "C" int idf_conv3x3_small_cin(const float* x, const float* w, const float* bias, void* y, int64_t ldy,
                                     int32_t B, int32_t Cin, int32_t H, int32_t W, int32_t Cout, int32_t dup,
                                     idf_stream_t stream) {
  if (!x || !w || !y) return fail(IDF_ERR_ARG, "conv_small_cin: null pointer");
  if (Cout % 128 != 0 || ldy % 4 != 0) return fail(IDF_ERR_ARG, "conv_small_cin: Cout must be a multiple of 128");
  const long long pixels = (long long)B * H * W;
  int RB = 8;
  while (RB > 1 && H % RB != 0) RB >>= 1;
  const int smem = Cin * (RB + 2) * (W + 2) * 4;
  if (smem > 48 * 1024) return fail(IDF_ERR_UNSUPPORTED, "conv_small_cin: image row block exceeds 48 KiB");
  dim3 grid(H / RB, B);
  cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
  __nv_bfloat16* yp = reinterpret_cast<__nv_bfloat16*>(y);
  const long long dup_rows = dup ? pixels : 0;
  switch (Cin) {
    case 3: conv3x3_small_cin_kernel<3><<<grid, 256, smem, s>>>(x, w, bias, yp, ldy, B, H, W, Cout, dup_rows, RB); break;
    case 4: conv3x3_small_cin_kernel<4><<<grid, 256, smem, s>>>(x, w, bias, yp, ldy, B, H, W, Cout, dup_rows, RB); break;
    default: return fail(IDF_ERR_UNSUPPORTED, "conv_small_cin: Cin %d not in {3,4}", Cin);
  }
  return check_cuda(cudaGetLastError(), "conv_small_cin launch");
}

template <int COUT>
static int launch_small_cout(const __nv_bfloat16* x, long long ldx, const float* w, const float* bias, float* y, int B,
                             int Cin, int H, int W, cudaStream_t s) {
  {
    // tiled path whenever the image splits into 128-pixel tiles and the halo tile fits shared memory
    const int TW = W >= 32 ? 32 : W;
    const int TH = 128 / (TW > 0 ? TW : 1);
    if (Cin == 128 && COUT <= 3 && TW * TH == 128 && W % TW == 0 && H % TH == 0) {
      const int rsmem = (TH + 2) * (TW + 2) * (128 * 2 + 16);
      static int rsmem_set = 0;
      if (rsmem > 48 * 1024 && rsmem > rsmem_set) {
        int rc = check_cuda(cudaFuncSetAttribute(conv3x3_small_cout_c128_kernel<COUT>,
                                                 cudaFuncAttributeMaxDynamicSharedMemorySize, rsmem),
                            "conv_small_cout c128: cudaFuncSetAttribute");
        if (rc != IDF_OK) return rc;
        rsmem_set = rsmem;
      }
      dim3 grid((H / TH) * (W / TW), B);
      conv3x3_small_cout_c128_kernel<COUT><<<grid, 128, rsmem, s>>>(x, ldx, w, bias, y, H, W, TH, TW);
      return check_cuda(cudaGetLastError(), "conv_small_cout c128 launch");
    }
    const int tsmem = ((9 * Cin * COUT * 4 + 15) & ~15) + (TH + 2) * (TW + 2) * (Cin * 2 + 16);
    if (TW * TH == 128 && W % TW == 0 && H % TH == 0 && tsmem <= 100 * 1024) {
      static int tsmem_set = 0;
      if (tsmem > 48 * 1024 && tsmem > tsmem_set) {
        int rc = check_cuda(cudaFuncSetAttribute(conv3x3_small_cout_tiled_kernel<COUT>,
                                                 cudaFuncAttributeMaxDynamicSharedMemorySize, tsmem),
                            "conv_small_cout tiled: cudaFuncSetAttribute");
        if (rc != IDF_OK) return rc;
        tsmem_set = tsmem;
      }
      dim3 grid((H / TH) * (W / TW), B);
      conv3x3_small_cout_tiled_kernel<COUT><<<grid, 128, tsmem, s>>>(x, ldx, w, bias, y, Cin, H, W, TH, TW);
      return check_cuda(cudaGetLastError(), "conv_small_cout tiled launch");
    }
  }
  const int smem = 9 * Cin * COUT * 4;
  static int smem_set = 0;
  if (smem > 48 * 1024 && smem > smem_set) {
    int rc = check_cuda(cudaFuncSetAttribute(conv3x3_small_cout_kernel<COUT>,
                                             cudaFuncAttributeMaxDynamicSharedMemorySize, smem),
                        "conv_small_cout: cudaFuncSetAttribute");
    if (rc != IDF_OK) return rc;
    smem_set = smem;
  }
  const unsigned grid = blocks_for((long long)B * H * W, 8, 148 * 8);
  conv3x3_small_cout_kernel<COUT><<<grid, 256, smem, s>>>(x, ldx, w, bias, y, B, Cin, H, W);
  return check_cuda(cudaGetLastError(), "conv_small_cout launch");
}

extern "C" int idf_conv3x3_small_cout(const void* x, int64_t ldx, const float* w, const float* bias, float* y,
                                      int32_t B, int32_t Cin, int32_t H, int32_t W, int32_t Cout,
                                      idf_stream_t stream) {
  if (!x || !w || !y) return fail(IDF_ERR_ARG, "conv_small_cout: null pointer");
  if (Cin % 8 != 0 || ldx % 8 != 0) return fail(IDF_ERR_ARG, "conv_small_cout: Cin and ldx must be multiples of 8");
  if (9 * Cin * Cout * 4 > 200 * 1024) return fail(IDF_ERR_UNSUPPORTED, "conv_small_cout: weights exceed shared memory");
  cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
  const __nv_bfloat16* xp = reinterpret_cast<const __nv_bfloat16*>(x);
  switch (Cout) {
    case 3: return launch_small_cout<3>(xp, ldx, w, bias, y, B, Cin, H, W, s);
    case 6: return launch_small_cout<6>(xp, ldx, w, bias, y, B, Cin, H, W, s);
    default: return fail(IDF_ERR_UNSUPPORTED, "conv_small_cout: Cout %d not in {3,6}", Cout);
  }
}

extern "C" int idf_conv1x1_small_f32(const float* x, const float* w, const float* bias, float* y, int32_t B,
                                     int32_t Cin, int32_t Cout, int32_t HW, idf_stream_t stream) {
  if (!x || !w || !y) return fail(IDF_ERR_ARG, "conv1x1_small: null pointer");
  const long long n = (long long)B * Cout * HW;
  conv1x1_small_kernel<<<(unsigned)((n + 255) / 256), 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(
      x, w, bias, y, B, Cin, Cout, HW);
  return check_cuda(cudaGetLastError(), "conv1x1_small launch");
}

extern "C" int idf_upsample_nearest2x(const void* x, int64_t ldx, void* y, int64_t ldy, int32_t B, int32_t H,
                                      int32_t W, int32_t C, idf_stream_t stream) {
  if (!x || !y) return fail(IDF_ERR_ARG, "upsample: null pointer");
  if (C % 8 != 0 || ldx % 8 != 0 || ldy % 8 != 0) return fail(IDF_ERR_ARG, "upsample: C/ld must be multiples of 8");
  const long long total = (long long)B * 4 * H * W * (C / 8);
  upsample2x_kernel<<<blocks_for(total, 256), 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(
      reinterpret_cast<const __nv_bfloat16*>(x), ldx, reinterpret_cast<__nv_bfloat16*>(y), ldy, B, H, W, C);
  return check_cuda(cudaGetLastError(), "upsample launch");
}

extern "C" int idf_space_to_depth2(const void* x, int64_t ldx, void* y, int32_t B, int32_t H, int32_t W, int32_t C,
                                   idf_stream_t stream) {
  if (!x || !y) return fail(IDF_ERR_ARG, "space_to_depth2: null pointer");
  if (C % 8 != 0 || ldx % 8 != 0 || H % 2 != 0 || W % 2 != 0) return fail(IDF_ERR_ARG, "space_to_depth2: bad shape");
  const long long total = (long long)B * (H / 2) * (W / 2) * (C / 8);
  space_to_depth2_kernel<<<blocks_for(total, 256), 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(
      reinterpret_cast<const __nv_bfloat16*>(x), ldx, reinterpret_cast<__nv_bfloat16*>(y), B, H, W, C);
  return check_cuda(cudaGetLastError(), "space_to_depth2 launch");
}

// out[i] = base[i] + t[0] * rows_per_t: row of sample i in a per-timestep embedding table (timestep read on the device)
__global__ void rowidx_from_timestep_kernel(const int32_t* __restrict__ base, const int64_t* __restrict__ t,
                                            int rows_per_t, int32_t* __restrict__ out, int n) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) out[i] = base[i] + (int32_t)t[0] * rows_per_t;
}

extern "C" int idf_rowidx_from_timestep(const int32_t* base, const int64_t* t, int32_t rows_per_t, int32_t* out,
                                        int32_t n, idf_stream_t stream) {
  if (!base || !t || !out || n <= 0) return fail(IDF_ERR_ARG, "rowidx_from_timestep: bad argument");
  rowidx_from_timestep_kernel<<<(n + 255) / 256, 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(base, t, rows_per_t, out, n);
  return check_cuda(cudaGetLastError(), "rowidx_from_timestep launch");
}

// uint8 NHWC image batch -> fp32 NCHW, y = x * scale + shift (prepare_dataset.py:104-105: / 127.5 - 1.0 and permute)
__global__ void u8_nhwc_to_f32_nchw_kernel(const uint8_t* __restrict__ x, float* __restrict__ y, int B, int H, int W,
                                           int C, float scale, float shift) {
  const long long total = (long long)B * C * H * W;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int w = (int)(i % W);
    long long r = i / W;
    const int h = (int)(r % H);
    r /= H;
    const int c = (int)(r % C);
    const long long b = r / C;
    y[i] = fmaf((float)x[((b * H + h) * W + w) * C + c], scale, shift);
  }
}

extern "C" int idf_u8_nhwc_to_f32_nchw(const uint8_t* x, float* y, int32_t B, int32_t H, int32_t W, int32_t C,
                                       float scale, float shift, idf_stream_t stream) {
  if (!x || !y || B <= 0) return fail(IDF_ERR_ARG, "u8_nhwc_to_f32_nchw: bad argument");
  const long long total = (long long)B * C * H * W;
  u8_nhwc_to_f32_nchw_kernel<<<blocks_for(total, 256), 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(x, y, B, H, W, C,
                                                                                                        scale, shift);
  return check_cuda(cudaGetLastError(), "u8_nhwc_to_f32_nchw launch");
}

extern "C" int idf_nchw_f32_to_nhwc_bf16(const float* x, void* y, int64_t ldy, int32_t B, int32_t C, int32_t HW,
                                         idf_stream_t stream) {
  if (!x || !y) return fail(IDF_ERR_ARG, "nchw_to_nhwc: null pointer");
  const long long n = (long long)B * C * HW;
  nchw_to_nhwc_kernel<<<(unsigned)((n + 255) / 256), 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(
      x, reinterpret_cast<__nv_bfloat16*>(y), ldy, B, C, HW);
  return check_cuda(cudaGetLastError(), "nchw_to_nhwc launch");
}

extern "C" int idf_nhwc_bf16_to_nchw_f32(const void* x, int64_t ldx, float* y, int32_t B, int32_t C, int32_t HW,
                                         idf_stream_t stream) {
  if (!x || !y) return fail(IDF_ERR_ARG, "nhwc_to_nchw: null pointer");
  const long long n = (long long)B * C * HW;
  nhwc_to_nchw_kernel<<<(unsigned)((n + 255) / 256), 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(
      reinterpret_cast<const __nv_bfloat16*>(x), ldx, y, B, C, HW);
  return check_cuda(cudaGetLastError(), "nhwc_to_nchw launch");
}
