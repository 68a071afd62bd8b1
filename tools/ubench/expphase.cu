// Micro-benchmark of the attention "exp phase": 128 independent scores per thread in registers ->
// p = exp2(s*c - mc), row sum, bf16 pack. Variants: all MUFU, or a fraction computed by a polynomial on the FMA pipe.
#include <cuda_bf16.h>
#include <cstdio>
#include <cstdint>

__device__ __forceinline__ float ex2(float x) { float y; asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
// Cody-Waite style exp2 on the FMA/ALU pipes: x <= 0. 2^x = 2^floor-ish(x) * p(frac), degree-3 polynomial.
__device__ __forceinline__ float ex2_poly(float x) {
  x = fmaxf(x, -125.f);
  const float t = x + 12582912.f;             // 1.5 * 2^23: rounds x to nearest integer in the low mantissa bits
  const float r = t - 12582912.f;
  const float f = x - r;                      // in [-0.5, 0.5]
  float p = fmaf(f, 0.0555041086f, 0.2402265069f);
  p = fmaf(p, f, 0.6931471806f);
  p = fmaf(p, f, 1.0f);
  return __int_as_float(__float_as_int(p) + (__float_as_int(t) << 23));
}
__device__ __forceinline__ uint32_t pack(float a, float b) { __nv_bfloat162 v = __floats2bfloat162_rn(a, b); return *reinterpret_cast<uint32_t*>(&v); }

template <int POLY_PER_8>
__global__ void __launch_bounds__(128, 1) k(const float* in, uint4* out, float* sums, int iters, float c, float mc) {
  float s[128];
#pragma unroll
  for (int i = 0; i < 128; ++i) s[i] = in[(i * 128 + threadIdx.x)];
  float tot = 0.f;
  long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
    float ps[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
    for (int q = 0; q < 16; ++q) {
      float e[8];
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const float x = fmaf(s[8 * q + i], c, -mc);
        e[i] = (i < POLY_PER_8) ? ex2_poly(x) : ex2(x);
        ps[i & 3] += e[i];
      }
      uint4 o; o.x = pack(e[0], e[1]); o.y = pack(e[2], e[3]); o.z = pack(e[4], e[5]); o.w = pack(e[6], e[7]);
      out[(blockIdx.x * 16 + q) * 128 + threadIdx.x] = o;
    }
    tot += (ps[0] + ps[1]) + (ps[2] + ps[3]);
    mc += 1e-6f;
  }
  long long t1 = clock64();
  sums[blockIdx.x * 128 + threadIdx.x] = tot;
  if (blockIdx.x == 0 && threadIdx.x == 0) printf("  poly %d/8: %.1f cycles per 128-element row-block (%.2f cyc/elt)\n", POLY_PER_8, double(t1 - t0) / iters, double(t1 - t0) / iters / 128);
}

int main() {
  float* in; uint4* out; float* sums;
  cudaMalloc(&in, 128 * 128 * 4); cudaMalloc(&out, 148 * 16 * 128 * 16); cudaMalloc(&sums, 148 * 128 * 4);
  cudaMemset(in, 0, 128 * 128 * 4);
  printf("1 warp per scheduler (128 threads/SM):\n");
  k<0><<<148, 128>>>(in, out, sums, 200, 0.25f, 1.f); cudaDeviceSynchronize();
  k<2><<<148, 128>>>(in, out, sums, 200, 0.25f, 1.f); cudaDeviceSynchronize();
  k<3><<<148, 128>>>(in, out, sums, 200, 0.25f, 1.f); cudaDeviceSynchronize();
  k<4><<<148, 128>>>(in, out, sums, 200, 0.25f, 1.f); cudaDeviceSynchronize();
  k<8><<<148, 128>>>(in, out, sums, 200, 0.25f, 1.f); cudaDeviceSynchronize();
  printf("2 blocks per SM (2 warps per scheduler):\n");
  k<0><<<296, 128>>>(in, out, sums, 200, 0.25f, 1.f); cudaDeviceSynchronize();
  k<2><<<296, 128>>>(in, out, sums, 200, 0.25f, 1.f); cudaDeviceSynchronize();
  k<4><<<296, 128>>>(in, out, sums, 200, 0.25f, 1.f); cudaDeviceSynchronize();
  printf("%s\n", cudaGetErrorString(cudaGetLastError()));
  return 0;
}
