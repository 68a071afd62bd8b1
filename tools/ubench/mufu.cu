// Micro-benchmark: per-SM throughput of the exp2 variants that could feed the attention softmax.
#include <cuda_fp16.h>
#include <cuda_bf16.h>
#include <cstdio>
#include <cstdint>

template <int MODE>
__global__ void k(float* out, int iters) {
  float a0 = threadIdx.x * 1e-3f, a1 = a0 + 0.1f, a2 = a0 + 0.2f, a3 = a0 + 0.3f;
  uint32_t h0 = 0x3c003800u + threadIdx.x, h1 = h0 + 1, h2 = h0 + 2, h3 = h0 + 3;
  for (int i = 0; i < iters; ++i) {
    if (MODE == 0) {  // ex2.approx.ftz.f32
      asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(a0));
      asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(a1));
      asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(a2));
      asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(a3));
    } else if (MODE == 1) {  // ex2.approx.f16x2
      asm volatile("ex2.approx.f16x2 %0, %0;" : "+r"(h0));
      asm volatile("ex2.approx.f16x2 %0, %0;" : "+r"(h1));
      asm volatile("ex2.approx.f16x2 %0, %0;" : "+r"(h2));
      asm volatile("ex2.approx.f16x2 %0, %0;" : "+r"(h3));
    } else if (MODE == 2) {  // ex2.approx.ftz.bf16x2
      asm volatile("ex2.approx.ftz.bf16x2 %0, %0;" : "+r"(h0));
      asm volatile("ex2.approx.ftz.bf16x2 %0, %0;" : "+r"(h1));
      asm volatile("ex2.approx.ftz.bf16x2 %0, %0;" : "+r"(h2));
      asm volatile("ex2.approx.ftz.bf16x2 %0, %0;" : "+r"(h3));
    } else if (MODE == 3) {  // tanh.approx.f32
      asm volatile("tanh.approx.f32 %0, %0;" : "+f"(a0));
      asm volatile("tanh.approx.f32 %0, %0;" : "+f"(a1));
      asm volatile("tanh.approx.f32 %0, %0;" : "+f"(a2));
      asm volatile("tanh.approx.f32 %0, %0;" : "+f"(a3));
    } else if (MODE == 4) {  // FFMA reference
      asm volatile("fma.rn.f32 %0, %0, %0, %0;" : "+f"(a0));
      asm volatile("fma.rn.f32 %0, %0, %0, %0;" : "+f"(a1));
      asm volatile("fma.rn.f32 %0, %0, %0, %0;" : "+f"(a2));
      asm volatile("fma.rn.f32 %0, %0, %0, %0;" : "+f"(a3));
    } else if (MODE == 5) {  // ex2.approx.f16 (scalar half)
      unsigned short s0 = h0, s1 = h1, s2 = h2, s3 = h3;
      asm volatile("ex2.approx.f16 %0, %0;" : "+h"(s0));
      asm volatile("ex2.approx.f16 %0, %0;" : "+h"(s1));
      asm volatile("ex2.approx.f16 %0, %0;" : "+h"(s2));
      asm volatile("ex2.approx.f16 %0, %0;" : "+h"(s3));
      h0 = s0; h1 = s1; h2 = s2; h3 = s3;
    } else if (MODE == 7) {  // cvt.rn.bf16x2.f32 pack
      asm volatile("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(h0) : "f"(a0), "f"(a1));
      asm volatile("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(h1) : "f"(a1), "f"(a2));
      asm volatile("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(h2) : "f"(a2), "f"(a3));
      asm volatile("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(h3) : "f"(a3), "f"(a0));
      a0 += __uint_as_float(h0); a1 += __uint_as_float(h1);
    } else if (MODE == 8) {  // fmnmx
      asm volatile("max.f32 %0, %0, %1;" : "+f"(a0) : "f"(a1));
      asm volatile("max.f32 %0, %0, %1;" : "+f"(a1) : "f"(a2));
      asm volatile("max.f32 %0, %0, %1;" : "+f"(a2) : "f"(a3));
      asm volatile("max.f32 %0, %0, %1;" : "+f"(a3) : "f"(a0));
    } else if (MODE == 9) {  // 2 ex2 + 1 bf16 pack + 2 ffma + 2 fadd (softmax inner mix)
      float e0, e1;
      asm volatile("fma.rn.f32 %0, %1, %2, %3;" : "=f"(e0) : "f"(a0), "f"(a2), "f"(a3));
      asm volatile("fma.rn.f32 %0, %1, %2, %3;" : "=f"(e1) : "f"(a1), "f"(a2), "f"(a3));
      asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(e0));
      asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(e1));
      asm volatile("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(h0) : "f"(e0), "f"(e1));
      a0 += e0; a1 += e1; h1 ^= h0;
    } else if (MODE == 6) {  // cvt.rn.f16x2.f32 pack
      asm volatile("cvt.rn.f16x2.f32 %0, %1, %2;" : "=r"(h0) : "f"(a0), "f"(a1));
      asm volatile("cvt.rn.f16x2.f32 %0, %1, %2;" : "=r"(h1) : "f"(a1), "f"(a2));
      asm volatile("cvt.rn.f16x2.f32 %0, %1, %2;" : "=r"(h2) : "f"(a2), "f"(a3));
      asm volatile("cvt.rn.f16x2.f32 %0, %1, %2;" : "=r"(h3) : "f"(a3), "f"(a0));
      a0 += __uint_as_float(h0); a1 += __uint_as_float(h1);
    }
  }
  out[blockIdx.x * blockDim.x + threadIdx.x] = a0 + a1 + a2 + a3 + __uint_as_float(h0 ^ h1 ^ h2 ^ h3);
}

template <int MODE>
void run_occ(const char* name, int threads) {
  float* out;
  cudaMalloc(&out, 148 * 8 * 1024 * 4);
  const int iters = 20000;
  k<MODE><<<148, threads>>>(out, 100);
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0); cudaEventCreate(&e1);
  cudaEventRecord(e0);
  k<MODE><<<148, threads>>>(out, iters);
  cudaEventRecord(e1);
  cudaEventSynchronize(e1);
  float ms; cudaEventElapsedTime(&ms, e0, e1);
  int clk_khz; cudaDeviceGetAttribute(&clk_khz, cudaDevAttrClockRate, 0);
  double instr = 148.0 * threads * iters * 4;
  printf("%-28s threads/SM=%4d  %6.2f thread-instr/clk/SM\n", name, threads, instr / (ms * 1e-3) / 148 / (clk_khz * 1e3));
  cudaFree(out);
}

template <int MODE>
void run(const char* name, int per_iter_elems) {
  float* out;
  cudaMalloc(&out, 148 * 8 * 1024 * 4);
  const int iters = 20000;
  k<MODE><<<148 * 2, 1024>>>(out, 100);
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0); cudaEventCreate(&e1);
  cudaEventRecord(e0);
  k<MODE><<<148 * 2, 1024>>>(out, iters);
  cudaEventRecord(e1);
  cudaEventSynchronize(e1);
  float ms; cudaEventElapsedTime(&ms, e0, e1);
  int clk_khz; cudaDeviceGetAttribute(&clk_khz, cudaDevAttrClockRate, 0);
  double instr = 148.0 * 2 * 1024 * iters * 4;          // thread-instructions
  double per_clk_sm = instr / (ms * 1e-3) / 148 / (clk_khz * 1e3);
  printf("%-28s %8.3f ms  %6.2f thread-instr/clk/SM (at %d MHz nominal)  -> %6.2f elems/clk/SM\n", name, ms, per_clk_sm,
         clk_khz / 1000, per_clk_sm * per_iter_elems);
  cudaFree(out);
}

int main() {
  for (int t : {128, 256, 512, 1024}) run_occ<0>("ex2 f32 (4 indep chains)", t);
  for (int t : {128, 256, 512}) run_occ<9>("softmax mix", t);
  for (int t : {128, 256}) run_occ<4>("ffma", t);
  run<0>("ex2.approx.ftz.f32", 1);
  run<1>("ex2.approx.f16x2", 2);
  run<2>("ex2.approx.ftz.bf16x2", 2);
  run<3>("tanh.approx.f32", 1);
  run<5>("ex2.approx.f16", 1);
  run<6>("cvt.rn.f16x2.f32 (+2 fadd)", 2);
  run<7>("cvt.rn.bf16x2.f32 (+2 fadd)", 2);
  run<8>("max.f32", 1);
  run<9>("softmax mix (x4 per it = 2 exps)", 0);
  run<4>("fma.rn.f32", 1);
  return 0;
}
