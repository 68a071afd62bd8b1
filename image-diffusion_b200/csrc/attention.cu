// Fused multi-head self-attention forward on tcgen05 (sm_100a): softmax(Q K^T * scale) V with the score tile
// held in tensor memory and never written to HBM.
//
// One CTA owns 128 consecutive token rows (a "query tile") of one head. For every 128-token block of keys of the
// same sample:   S = Q K^T  (UMMA 128x128xhd, fp32 in TMEM)  ->  four softmax warps read S, one row per thread,
// keep a running max / sum (online softmax, fp32) and write P = exp2(..) as bf16 into shared memory in the
// swizzled K-major UMMA layout  ->  O_blk = P V (UMMA 128xhdx128, V supplied pre-transposed so both operands
// are K-major)  ->  the softmax warps fold O_blk into per-thread fp32 output rows.
// Samples with fewer than 128 tokens share a tile; a block-diagonal mask keeps them independent.
//
// Warp roles (192 threads): warps 0..3 = softmax / output rows, warp 4 = TMA producer, warp 5 = TMEM + MMA issuer.
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "common.cuh"
#include "host.h"
#include "../../include/idf_b200.h"

namespace idf {

constexpr int ATT_THREADS = 192;
constexpr int ATT_TMEM_COLS = 256;  // S: columns [0,128), O_blk: columns [128, 128+hd)

struct AttnParams {
  CUtensorMap tmQK;  // (M, 2C) bf16, box (SWZ/2, 128)
  CUtensorMap tmVT;  // (C, M)  bf16, box (64, HD)
  __nv_bfloat16* out;
  long long ld_out;
  int M, T, C;
  int t_shift;     // log2(T)
  int nblk;        // key blocks per query tile
  int kv_stages;   // 1 or 2
  float scale_log2e;
  int heads;
  float* lse;  // optional (M, heads): log2-domain log-sum-exp of the scaled scores (training: consumed by the backward)
  int v_tok;   // 1: tmVT maps the token-major (M, 3C) QKV matrix, box (64, 128); V tiles are MN-major UMMA operands
#ifdef IDF_ATTN_TRACE
  long long* trace;  // debug build only (csrc/build.py --trace): clock64 stamps of CTA 0, [role][block][event]
#endif
};

#ifdef IDF_ATTN_TRACE
#define ATT_TRACE_BLOCKS 32
#define ATT_TRACE_EVENTS 8
#define ATT_STAMP(role, blk, evt)                                                                              \
  do {                                                                                                         \
    if (blockIdx.x == 0 && (blk) < ATT_TRACE_BLOCKS)                                                           \
      p.trace[((role) * ATT_TRACE_BLOCKS + (blk)) * ATT_TRACE_EVENTS + (evt)] = clock64();                     \
  } while (0)
#else
#define ATT_STAMP(role, blk, evt) do { } while (0)
#endif

__device__ __forceinline__ float fast_exp2(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

// 2^x on the FMA pipe (Cody-Waite range reduction + degree-3 minimax polynomial, relative error 7.5e-5: far below the
// bf16 rounding of the probabilities): x = n + f, f in [-0.5, 0.5]; 2^f by Horner; n goes into the exponent field with
// one integer multiply-add. 9 FMA-pipe instructions, no MUFU. IDF_ATTN_POLY (compile time) = how many of every 8
// exponentials of the pipelined kernel's softmax take this path instead of MUFU.EX2.
#ifndef IDF_ATTN_POLY
#define IDF_ATTN_POLY 0
#endif
__device__ __forceinline__ float poly_exp2(float x) {
  x = fmaxf(x, -126.f);
  const float t = x + 12582912.f;  // 1.5 * 2^23: the integer part of x lands in the low mantissa bits
  const float f = x - (t - 12582912.f);
  float p = fmaf(0.055171460f, f, 0.24261086f);
  p = fmaf(p, f, 0.69326097f);
  p = fmaf(p, f, 0.99992812f);
  return __uint_as_float(__float_as_uint(t) * 8388608u + __float_as_uint(p));  // p * 2^n
}

template <int HD>
__global__ void __launch_bounds__(ATT_THREADS, 2) attention_kernel(const __grid_constant__ AttnParams p) {
  constexpr int SWZ = HD <= 16 ? 32 : (HD <= 32 ? 64 : 128);  // bytes per Q/K row in shared memory
  constexpr int QK_BYTES = 128 * SWZ;
  constexpr int V_BYTES = 2 * HD * 128;  // two boxes of (HD rows x 64 keys)
  constexpr int P_BYTES = 2 * 128 * 128; // two blocks of (128 rows x 64 keys)
  // V tile: two (HD rows x 64 keys) boxes of V^T (K-major operand), or - token-major V - one (128 keys x 64 columns)
  // box of the QKV matrix used as an MN-major operand (no transposed copy of V exists then)
  const int KV_BYTES = QK_BYTES + (p.v_tok ? 128 * 128 : V_BYTES);

  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw_addr = smem_u32(smem_raw);
  uint8_t* smem = smem_raw + ((1024u - (raw_addr & 1023u)) & 1023u);
  uint8_t* smem_q = smem;
  uint8_t* smem_p = smem + QK_BYTES;
  uint8_t* smem_kv = smem_p + P_BYTES;  // [kv_stages][K | V]
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem_kv + p.kv_stages * KV_BYTES);
  uint64_t* q_full = bars;
  uint64_t* kv_full = bars + 1;   // [2]
  uint64_t* kv_empty = bars + 3;  // [2]
  uint64_t* s_full = bars + 5;
  uint64_t* s_empty = bars + 6;
  uint64_t* p_full = bars + 7;
  uint64_t* o_full = bars + 8;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 9);

  const int warp = __shfl_sync(0xffffffffu, threadIdx.x >> 5, 0);
  const int lane = threadIdx.x & 31;
  const int tile_m = blockIdx.x;
  const int head = blockIdx.y;
  const int row0 = tile_m * 128;
  // first key row of this tile's key range
  const int kv_base = (p.T >= 128) ? (row0 >> p.t_shift) << p.t_shift : row0;

  if (warp == 4 && lane == 0) {
    tma_prefetch_desc(&p.tmQK);
    tma_prefetch_desc(&p.tmVT);
    mbar_init(q_full, 1);
    for (int s = 0; s < 2; ++s) {
      mbar_init(&kv_full[s], 1);
      mbar_init(&kv_empty[s], 1);
    }
    mbar_init(s_full, 1);
    mbar_init(s_empty, 4);
    mbar_init(p_full, 4);
    mbar_init(o_full, 1);
    fence_mbar_init();
  }
  if (warp == 5) {
    tmem_alloc(tmem_slot, ATT_TMEM_COLS);
    tmem_relinquish();
  }
  tc_fence_before_sync();
  __syncthreads();
  tc_fence_after_sync();
  const uint32_t tmem_base = *tmem_slot;
  const uint32_t tmem_s = tmem_base;
  const uint32_t tmem_o = tmem_base + 128;

  if (warp == 4) {
    // ------------------------------------------------------------------ TMA producer
    if (elect_one()) {
      mbar_expect_tx(q_full, QK_BYTES);
      tma_load_2d(smem_q, &p.tmQK, q_full, head * HD, row0);
      for (int j = 0; j < p.nblk; ++j) {
        const int st = j % p.kv_stages;
        const uint32_t ph = (j / p.kv_stages) & 1;
        mbar_wait(&kv_empty[st], ph ^ 1);
        mbar_expect_tx(&kv_full[st], KV_BYTES);
        uint8_t* k_dst = smem_kv + st * KV_BYTES;
        uint8_t* v_dst = k_dst + QK_BYTES;
        const int kv0 = kv_base + j * 128;
        tma_load_2d(k_dst, &p.tmQK, &kv_full[st], p.C + head * HD, kv0);
        if (p.v_tok) {
          tma_load_2d(v_dst, &p.tmVT, &kv_full[st], 2 * p.C + head * HD, kv0);
        } else {
          tma_load_2d(v_dst, &p.tmVT, &kv_full[st], kv0, head * HD);
          tma_load_2d(v_dst + HD * 128, &p.tmVT, &kv_full[st], kv0 + 64, head * HD);
        }
      }
    }
  } else if (warp == 5) {
    // ------------------------------------------------------------------ MMA issuer
    if (elect_one()) {
      constexpr uint32_t idesc_s = umma_idesc_bf16(128, 128);
      constexpr uint32_t idesc_o = umma_idesc_bf16(128, HD);
      constexpr uint32_t idesc_o_mn = umma_idesc_bf16(128, HD, 0, 1);
      const uint64_t dq = umma_desc_kmajor(smem_u32(smem_q), SWZ);
      const uint64_t dp0 = umma_desc_kmajor(smem_u32(smem_p), 128);
      const uint64_t dp1 = umma_desc_kmajor(smem_u32(smem_p + 128 * 128), 128);
      mbar_wait(q_full, 0);
      for (int j = 0; j < p.nblk; ++j) {
        const int st = j % p.kv_stages;
        const uint32_t ph = (j / p.kv_stages) & 1;
        uint8_t* k_src = smem_kv + st * KV_BYTES;
        uint8_t* v_src = k_src + QK_BYTES;
        mbar_wait(&kv_full[st], ph);
        if (j > 0) mbar_wait(s_empty, (j - 1) & 1);
        tc_fence_after_sync();
        const uint64_t dk = umma_desc_kmajor(smem_u32(k_src), SWZ);
#pragma unroll
        for (int k = 0; k < HD / 16; ++k) umma_bf16(tmem_s, dq + 2 * k, dk + 2 * k, idesc_s, k != 0);
        umma_commit(s_full);

        mbar_wait(p_full, j & 1);
        tc_fence_after_sync();
        if (p.v_tok) {
          const uint64_t dvm = umma_desc_mnmajor(smem_u32(v_src), 8192, 1024);
#pragma unroll
          for (int k = 0; k < 8; ++k)  // 16 keys per step = 16 rows of 128 bytes
            umma_bf16(tmem_o, (k < 4 ? dp0 : dp1) + 2 * (k & 3), dvm + 128 * k, idesc_o_mn, k != 0);
        } else {
          const uint64_t dv0 = umma_desc_kmajor(smem_u32(v_src), 128);
          const uint64_t dv1 = umma_desc_kmajor(smem_u32(v_src + HD * 128), 128);
#pragma unroll
          for (int k = 0; k < 8; ++k) {
            const uint64_t da = (k < 4 ? dp0 : dp1) + 2 * (k & 3);
            const uint64_t db = (k < 4 ? dv0 : dv1) + 2 * (k & 3);
            umma_bf16(tmem_o, da, db, idesc_o, k != 0);
          }
        }
        umma_commit(&kv_empty[st]);
        umma_commit(o_full);
      }
    }
  } else {
    // ------------------------------------------------------------------ softmax + output (warps 0..3)
    const int r = warp * 32 + lane;  // row of the tile == TMEM lane
    const uint32_t lane_addr = (uint32_t)(warp * 32) << 16;
    const long long m = (long long)row0 + r;
    const bool masked = p.T < 128;
    const int row_seg = r >> p.t_shift;
    const float c = p.scale_log2e;

    float o_acc[HD];
#pragma unroll
    for (int d = 0; d < HD; ++d) o_acc[d] = 0.f;
    float m_run = -INFINITY;
    float l_run = 0.f;

    for (int j = 0; j <= p.nblk; ++j) {
      if (j > 0) {
        // fold the previous block's P V into the running output
        mbar_wait(o_full, (j - 1) & 1);
        tc_fence_after_sync();
#pragma unroll
        for (int d0 = 0; d0 < HD; d0 += 16) {
          uint32_t v[16];
          tmem_ld_32x16(tmem_o + lane_addr + d0, v);
          tmem_ld_wait();
#pragma unroll
          for (int d = 0; d < 16; ++d) o_acc[d0 + d] += __uint_as_float(v[d]);
        }
      }
      if (j == p.nblk) break;

      mbar_wait(s_full, j & 1);
      tc_fence_after_sync();

      // pass A: row maximum over the valid keys
      float mx = -INFINITY;
#pragma unroll 1
      for (int ch = 0; ch < 4; ++ch) {
        if (masked && p.T >= 32 && ((ch * 32) >> p.t_shift) != row_seg) continue;  // warp-uniform
        uint32_t v[32];
        tmem_ld_32x32(tmem_s + lane_addr + ch * 32, v);
        tmem_ld_wait();
#pragma unroll
        for (int i = 0; i < 32; ++i) {
          float s = __uint_as_float(v[i]);
          if (masked && ((ch * 32 + i) >> p.t_shift) != row_seg) s = -INFINITY;
          mx = fmaxf(mx, s);
        }
      }
      const float m_new = fmaxf(m_run, mx);
      const float alpha = fast_exp2((m_run - m_new) * c);  // first block: exp2(-inf) = 0
      const float mc = m_new * c;
      l_run *= alpha;
#pragma unroll
      for (int d = 0; d < HD; ++d) o_acc[d] *= alpha;
      m_run = m_new;

      // pass B: P = exp2(S*c - m*c) -> bf16 -> swizzled shared memory; running sum in fp32
      float psum = 0.f;
#pragma unroll 1
      for (int ch = 0; ch < 4; ++ch) {
        uint8_t* prow = smem_p + (ch >> 1) * (128 * 128) + r * 128;
        const bool skip = masked && p.T >= 32 && ((ch * 32) >> p.t_shift) != row_seg;
        uint32_t v[32];
        if (!skip) {
          tmem_ld_32x32(tmem_s + lane_addr + ch * 32, v);
          tmem_ld_wait();
        }
        float pv[32];
#pragma unroll
        for (int i = 0; i < 32; ++i) {
          float e = 0.f;
          if (!skip) {
            const bool ok = !masked || (((ch * 32 + i) >> p.t_shift) == row_seg);
            e = ok ? fast_exp2(fmaf(__uint_as_float(v[i]), c, -mc)) : 0.f;
          }
          pv[i] = e;
          psum += e;
        }
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          const int chunk = ((ch & 1) * 4 + q) ^ (r & 7);
          uint4 o;
          o.x = pack_bf16x2(pv[8 * q + 0], pv[8 * q + 1]);
          o.y = pack_bf16x2(pv[8 * q + 2], pv[8 * q + 3]);
          o.z = pack_bf16x2(pv[8 * q + 4], pv[8 * q + 5]);
          o.w = pack_bf16x2(pv[8 * q + 6], pv[8 * q + 7]);
          *reinterpret_cast<uint4*>(prow + chunk * 16) = o;
        }
      }
      l_run += psum;

      // S has been consumed: the next Q K^T may overwrite it; P is ready for the P V product.
      tc_fence_before_sync();
      fence_proxy_async_smem();
      __syncwarp();
      if (lane == 0) {
        mbar_arrive(s_empty);
        mbar_arrive(p_full);
      }
    }

    if (m < p.M) {
      const float inv = 1.f / l_run;
      if (p.lse != nullptr) p.lse[m * p.heads + head] = fmaf(m_run, c, __log2f(l_run));
      __nv_bfloat16* dst = p.out + m * p.ld_out + head * HD;
#pragma unroll
      for (int d0 = 0; d0 < HD; d0 += 8) {
        uint4 o;
        o.x = pack_bf16x2(o_acc[d0 + 0] * inv, o_acc[d0 + 1] * inv);
        o.y = pack_bf16x2(o_acc[d0 + 2] * inv, o_acc[d0 + 3] * inv);
        o.z = pack_bf16x2(o_acc[d0 + 4] * inv, o_acc[d0 + 5] * inv);
        o.w = pack_bf16x2(o_acc[d0 + 6] * inv, o_acc[d0 + 7] * inv);
        *reinterpret_cast<uint4*>(dst + d0) = o;
      }
    }
  }

  tc_fence_before_sync();
  __syncthreads();
  if (warp == 5) {
    __syncwarp();
    tc_fence_after_sync();
    tmem_dealloc(tmem_base, ATT_TMEM_COLS);
  }
}


// ---------------------------------------------------------------------------------------------------------------
// Pipelined variant for samples with >= 256 tokens. One CTA (384 threads) per SM owns TWO query tiles (256 rows)
// of one head; every K / V block is fetched once and used for both tiles.
//   * warpgroup g (warps 4g..4g+3, one row per thread) runs the online softmax of query tile g; the two groups
//     are started half a period apart so that one group's MUFU-bound exp phase overlaps the other group's TMEM
//     reads, max reduction and stores (a single warp per scheduler reaches only ~60% of the MUFU rate);
//   * the MMA thread issues Q_g K_{j+1}^T as soon as group g has pulled S_g(j) into registers, so the next score
//     tile is waiting in TMEM when the group comes back for it;
//   * the output rows accumulate in TMEM over all key blocks (lazy rescaling, see below): no per-block read-back.
// TMEM columns: S_0 [0,128) S_1 [128,256) O_0 [256,256+hd) O_1 [320,320+hd) -> 512 allocated.
// ---------------------------------------------------------------------------------------------------------------
constexpr int ATTP_TMEM_COLS = 512;
template <int HD> struct AttpK { static constexpr int STAGES = HD > 32 ? 3 : 4; };  // K ring depth (smem budget)
constexpr int ATTP_VSTAGES = 3;
constexpr int ATTP_THREADS = 384;  // warps 0..7 softmax (2 warpgroups), warp 8 TMA producer, warp 9 TMEM + MMA, 10-11 idle

template <int N>
__device__ __forceinline__ void tmem_ld_cols(uint32_t taddr, float* dst) {
  static_assert(N == 8 || N == 16 || N == 24 || N == 32, "unsupported column count");
  if constexpr (N == 32) {
    uint32_t v[32];
    tmem_ld_32x32(taddr, v);
    tmem_ld_wait();
#pragma unroll
    for (int i = 0; i < 32; ++i) dst[i] = __uint_as_float(v[i]);
  } else {
    if constexpr (N >= 16) {
      uint32_t v[16];
      tmem_ld_32x16(taddr, v);
      tmem_ld_wait();
#pragma unroll
      for (int i = 0; i < 16; ++i) dst[i] = __uint_as_float(v[i]);
    }
    if constexpr (N == 8 || N == 24) {
      constexpr int base = N - 8;
      uint32_t v[8];
      asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
                   : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7])
                   : "r"(taddr + base)
                   : "memory");
      tmem_ld_wait();
#pragma unroll
      for (int i = 0; i < 8; ++i) dst[base + i] = __uint_as_float(v[i]);
    }
  }
}

template <int HD>
__global__ void __launch_bounds__(ATTP_THREADS, 1) attention_pipe_kernel(const __grid_constant__ AttnParams p) {
  constexpr int SWZ = HD <= 16 ? 32 : (HD <= 32 ? 64 : 128);
  constexpr int QK_BYTES = 128 * SWZ;
  const int V_BYTES = p.v_tok ? 128 * 128 : 2 * HD * 128;
  constexpr int ATTP_KSTAGES = AttpK<HD>::STAGES;
  // Row sums on the tensor core: for head_dim <= 48 the 64-column output accumulator has 16 spare columns, which
  // receive P x ones (a second, tiny product per key block). The softmax threads then do no additions at all (128
  // FADD per row and key block saved; the loop is issue-bound, not MUFU-bound), and the denominator is the sum of
  // the SAME bf16-rounded probabilities that multiply V.
  constexpr bool LSUM = HD <= 48;

  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw_addr = smem_u32(smem_raw);
  uint8_t* smem = smem_raw + ((1024u - (raw_addr & 1023u)) & 1023u);
  uint8_t* smem_q = smem;                                 // [2 items][2 tiles][QK_BYTES]
  uint8_t* smem_k = smem_q + 4 * QK_BYTES;                // [ATTP_KSTAGES][QK_BYTES]
  uint8_t* smem_v = smem_k + ATTP_KSTAGES * QK_BYTES;     // [ATTP_VSTAGES][V_BYTES]
  uint8_t* smem_ones = smem_v + ATTP_VSTAGES * V_BYTES;   // 16 rows x 128 bytes of bf16 1.0 (any layout: all ones)
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem_ones + 2048);
  uint64_t* q_full = bars;                           // [2]
  uint64_t* q_empty = q_full + 2;                    // [2]
  uint64_t* k_full = q_empty + 2;                    // [4]
  uint64_t* k_empty = k_full + ATTP_KSTAGES;         // [4]
  uint64_t* v_full = k_empty + ATTP_KSTAGES;         // [3]
  uint64_t* v_empty = v_full + ATTP_VSTAGES;         // [3]
  uint64_t* s_full = v_empty + ATTP_VSTAGES;         // [2] per softmax group
  uint64_t* s_empty = s_full + 2;                    // [2]
  uint64_t* p_full = s_empty + 2;                    // [2]
  uint64_t* o_full = p_full + 2;                     // [2]
  uint64_t* o_empty = o_full + 2;                    // [2]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(o_empty + 2);

  const int warp = __shfl_sync(0xffffffffu, threadIdx.x >> 5, 0);
  const int lane = threadIdx.x & 31;
  const int n = p.nblk;
  // work items: (pair of query tiles, head), dealt round-robin to the persistent CTAs
  const int pairs = (p.M + 255) / 256;
  const int total_items = pairs * p.heads;
  const int my_items = (total_items - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x;
  const int E = my_items * n;  // key blocks this CTA walks through (per softmax group)
  auto item_head = [&](int item) { return (int)((blockIdx.x + item * gridDim.x) % p.heads); };
  auto item_row0 = [&](int item) { return (int)((blockIdx.x + item * gridDim.x) / p.heads) * 256; };

  if (warp == 8 && lane == 0) {
    tma_prefetch_desc(&p.tmQK);
    tma_prefetch_desc(&p.tmVT);
    for (int i = 0; i < 2; ++i) { mbar_init(&q_full[i], 1); mbar_init(&q_empty[i], 1); }
    for (int i = 0; i < ATTP_KSTAGES; ++i) { mbar_init(&k_full[i], 1); mbar_init(&k_empty[i], 1); }
    for (int i = 0; i < ATTP_VSTAGES; ++i) { mbar_init(&v_full[i], 1); mbar_init(&v_empty[i], 1); }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&s_full[i], 1);
      mbar_init(&s_empty[i], 4);
      mbar_init(&p_full[i], 4);
      mbar_init(&o_full[i], 1);
      mbar_init(&o_empty[i], 4);
    }
    fence_mbar_init();
  }
  if (warp == 9) {
    tmem_alloc(tmem_slot, ATTP_TMEM_COLS);
    tmem_relinquish();
  }
  if (threadIdx.x < 128) reinterpret_cast<uint4*>(smem_ones)[threadIdx.x] = make_uint4(0x3f803f80u, 0x3f803f80u, 0x3f803f80u, 0x3f803f80u);
  fence_proxy_async_smem();
  tc_fence_before_sync();
  __syncthreads();
  tc_fence_after_sync();
  const uint32_t tmem_base = *tmem_slot;

  // register re-balancing between warpgroups: the data-movement warpgroup gives its registers to the two softmax
  // warpgroups, which keep a whole 128-wide score row per thread
  if (warp >= 8) {
    asm volatile("setmaxnreg.dec.sync.aligned.u32 56;");
    if (warp == 8) {
      // ------------------------------------------------------------------ TMA producer
      if (elect_one()) {
        auto load_q = [&](int item) {
          const int st = item & 1;
          mbar_wait(&q_empty[st], ((item >> 1) & 1) ^ 1);
          mbar_expect_tx(&q_full[st], 2 * QK_BYTES);
          uint8_t* dst = smem_q + st * 2 * QK_BYTES;
          tma_load_2d(dst, &p.tmQK, &q_full[st], item_head(item) * HD, item_row0(item));
          tma_load_2d(dst + QK_BYTES, &p.tmQK, &q_full[st], item_head(item) * HD, item_row0(item) + 128);
        };
        auto load_k = [&](int e) {
          const int item = e / n, j = e % n;
          if (j == 0) load_q(item);
          const int st = e % ATTP_KSTAGES;
          const int kv0 = ((item_row0(item) >> p.t_shift) << p.t_shift) + j * 128;
          mbar_wait(&k_empty[st], ((e / ATTP_KSTAGES) & 1) ^ 1);
          mbar_expect_tx(&k_full[st], QK_BYTES);
          tma_load_2d(smem_k + st * QK_BYTES, &p.tmQK, &k_full[st], p.C + item_head(item) * HD, kv0);
        };
        auto load_v = [&](int e) {
          const int item = e / n, j = e % n;
          const int st = e % ATTP_VSTAGES;
          const int kv0 = ((item_row0(item) >> p.t_shift) << p.t_shift) + j * 128;
          mbar_wait(&v_empty[st], ((e / ATTP_VSTAGES) & 1) ^ 1);
          mbar_expect_tx(&v_full[st], V_BYTES);
          uint8_t* dst = smem_v + st * V_BYTES;
          if (p.v_tok) {
            tma_load_2d(dst, &p.tmVT, &v_full[st], 2 * p.C + item_head(item) * HD, kv0);
          } else {
            tma_load_2d(dst, &p.tmVT, &v_full[st], kv0, item_head(item) * HD);
            tma_load_2d(dst + HD * 128, &p.tmVT, &v_full[st], kv0 + 64, item_head(item) * HD);
          }
        };
        if (E > 0) load_k(0);
        if (E > 1) load_k(1);
        for (int e = 0; e < E; ++e) {  // K runs two blocks ahead of V (S(e+1) is issued during the softmax of e)
          load_v(e);
          if (e + 2 < E) load_k(e + 2);
        }
      }
    } else if (warp == 9) {
      // ------------------------------------------------------------------ MMA issuer
      if (elect_one()) {
        constexpr uint32_t idesc_s = umma_idesc_bf16(128, 128);
        constexpr uint32_t idesc_o = umma_idesc_bf16(128, HD);
        constexpr uint32_t idesc_o_mn = umma_idesc_bf16(128, HD, 0, 1);
        auto issue_s = [&](int g, int e) {  // S_g = Q_g K_e^T
          const int item = e / n, j = e % n;
          const int st = e % ATTP_KSTAGES;
          if (g == 0) {
            if (j == 0) mbar_wait(&q_full[item & 1], (item >> 1) & 1);
            mbar_wait(&k_full[st], (e / ATTP_KSTAGES) & 1);
          }
          if (e > 0) mbar_wait(&s_empty[g], (e - 1) & 1);  // the group has pulled its previous S into registers
          ATT_STAMP(0, e, g);  // S_g(e) issue
          tc_fence_after_sync();
          const uint64_t dq = umma_desc_kmajor(smem_u32(smem_q + ((item & 1) * 2 + g) * QK_BYTES), SWZ);
          const uint64_t dk = umma_desc_kmajor(smem_u32(smem_k + st * QK_BYTES), SWZ);
#pragma unroll
          for (int k = 0; k < HD / 16; ++k) umma_bf16(tmem_base + g * 128, dq + 2 * k, dk + 2 * k, idesc_s, k != 0);
          umma_commit(&s_full[g]);
          if (g == 1) {
            umma_commit(&k_empty[st]);                         // both tiles have consumed K_e
            if (j == n - 1) umma_commit(&q_empty[item & 1]);   // ... and this item's Q tiles
          }
        };
        auto issue_pv = [&](int g, int e) {  // O_g (+)= P_g V_e
          const int item = e / n, j = e % n;
          const int vs = e % ATTP_VSTAGES;
          ATT_STAMP(0, e, 2 + g);  // start waiting for P_g(e)
          mbar_wait(&p_full[g], e & 1);
          ATT_STAMP(0, e, 4 + g);  // P_g(e) ready
          if (g == 0) mbar_wait(&v_full[vs], (e / ATTP_VSTAGES) & 1);
          if (j == 0 && item > 0) mbar_wait(&o_empty[g], (item - 1) & 1);  // previous item's output rows were read
          tc_fence_after_sync();
          // A = P_g from tensor memory (packed bf16 pairs, 64 columns for 128 keys): the probabilities never touch
          // shared memory, whose port is left to TMA writes and the K / V / Q operand reads
          const uint8_t* vb = smem_v + vs * V_BYTES;
          const uint32_t tp = tmem_base + 384 + g * 64;
          if (p.v_tok) {
            const uint64_t dvm = umma_desc_mnmajor(smem_u32(vb), 8192, 1024);
#pragma unroll
            for (int k = 0; k < 8; ++k)
              umma_bf16_ts(tmem_base + 256 + g * 64, tp + 8 * k, dvm + 128 * k, idesc_o_mn, (j > 0) || (k != 0));
          } else {
            const uint64_t dv0 = umma_desc_kmajor(smem_u32(vb), 128);
            const uint64_t dv1 = umma_desc_kmajor(smem_u32(vb + HD * 128), 128);
#pragma unroll
            for (int k = 0; k < 8; ++k)
              umma_bf16_ts(tmem_base + 256 + g * 64, tp + 8 * k, (k < 4 ? dv0 : dv1) + 2 * (k & 3), idesc_o,
                           (j > 0) || (k != 0));  // accumulate over the item's key blocks
          }
          if constexpr (LSUM) {
            constexpr uint32_t idesc_l = umma_idesc_bf16(128, 16);
            const uint64_t d1 = umma_desc_kmajor(smem_u32(smem_ones), 128);
#pragma unroll
            for (int k = 0; k < 8; ++k)  // l_g (+)= P_g x 1 into the accumulator's spare columns [48, 64)
              umma_bf16_ts(tmem_base + 256 + g * 64 + 48, tp + 8 * k, d1, idesc_l, (j > 0) || (k != 0));
          }
          umma_commit(&o_full[g]);
          ATT_STAMP(0, e, 6 + g);  // P_g V(e) issued + committed
          if (g == 1) umma_commit(&v_empty[vs]);
        };
        if (E > 0) {
          // Group 1 starts about half a period after group 0 (its first S is issued only once group 0 has taken
          // S_0(0) into registers): with the groups out of phase, one group's MUFU-bound exp phase overlaps the
          // other's TMEM reads / max reduction / stores. The offset persists across work items.
          issue_s(0, 0);
          if (E > 1) issue_s(0, 1);  // waits for s_empty[0](0)
          else mbar_wait(&s_empty[0], 0);
          issue_s(1, 0);
          for (int e = 0; e < E; ++e) {
            issue_pv(0, e);
            if (e + 1 < E) issue_s(1, e + 1);
            if (e + 2 < E) issue_s(0, e + 2);
            issue_pv(1, e);
          }
        }
      }
    }
  } else {
    asm volatile("setmaxnreg.inc.sync.aligned.u32 224;");
    // ------------------------------------------------------------------ softmax + output rows (warps 0..7)
    // Warpgroup g (warps 4g..4g+3, one row per thread) owns query tile g of every work item: own S / P / O buffers,
    // own running (max, sum). The output rows accumulate in TMEM across the item's key blocks (P V with
    // accumulate). The softmax uses a possibly stale row maximum m_run: P = exp2((S - m_run) * c) is exact
    // arithmetic as long as it cannot overflow, so the accumulator is only rescaled when the true maximum outgrows
    // m_run by more than 2^8 (rare after the first block). No per-block read-back of the output rows.
    const int quad = warp & 3;
    const int g = warp >> 2;
    const int r = quad * 32 + lane;
    const uint32_t lane_addr = (uint32_t)(quad * 32) << 16;
    const float c = p.scale_log2e;
    const uint32_t tmem_o = tmem_base + 256 + g * 64 + lane_addr;
    const uint32_t tmem_p = tmem_base + 384 + g * 64 + lane_addr;

    int e = 0;
    for (int item = 0; item < my_items; ++item) {
      float m_run = -INFINITY, l_run = 0.f;
      for (int j = 0; j < n; ++j, ++e) {
#ifdef IDF_ATTN_TRACE
        const bool tr = quad == 0 && lane == 0;
#define SM_STAMP(evt) do { if (tr) ATT_STAMP(1 + g, e, evt); } while (0)
#else
#define SM_STAMP(evt) do { } while (0)
#endif
        SM_STAMP(0);  // waiting for S(e)
        mbar_wait(&s_full[g], e & 1);
        SM_STAMP(1);  // S(e) in TMEM
        tc_fence_after_sync();
        uint32_t sv[4][32];
#pragma unroll
        for (int ch = 0; ch < 4; ++ch) tmem_ld_32x32(tmem_base + g * 128 + lane_addr + ch * 32, sv[ch]);
        tmem_ld_wait();
        tc_fence_before_sync();
        __syncwarp();
        if (lane == 0) mbar_arrive(&s_empty[g]);  // S now lives in registers
        SM_STAMP(2);  // S(e) in registers

        // rescale of the output accumulator (and of the row sums kept next to it) by alpha = 2^((m_old - m_new) c)
        auto rescale = [&](const float alpha) {
          tc_fence_after_sync();
#pragma unroll
          for (int d0 = 0; d0 < HD; d0 += 16) {
            uint32_t v[16];
            tmem_ld_32x16(tmem_o + d0, v);
            tmem_ld_wait();
#pragma unroll
            for (int d = 0; d < 16; ++d) v[d] = __float_as_uint(__uint_as_float(v[d]) * alpha);
            tmem_st_32x16(tmem_o + d0, v);
          }
          if constexpr (LSUM) {
            uint32_t v[16];
            tmem_ld_32x16(tmem_o + 48, v);
            tmem_ld_wait();
#pragma unroll
            for (int d = 0; d < 16; ++d) v[d] = __float_as_uint(__uint_as_float(v[d]) * alpha);
            tmem_st_32x16(tmem_o + 48, v);
          }
          tmem_st_wait();
          tc_fence_before_sync();
        };
        float ps[4];
        // P = 2^(S c - mc) as packed bf16 pairs into tensor memory
        auto exp_store = [&](const float mc) {
          ps[0] = ps[1] = ps[2] = ps[3] = 0.f;
#pragma unroll
          for (int ch = 0; ch < 4; ++ch) {
            uint32_t pk[16];  // 32 probabilities of this row as 16 packed bf16 pairs -> 16 TMEM columns
#pragma unroll
            for (int q = 0; q < 4; ++q) {
              float ev[8];
#pragma unroll
              for (int i = 0; i < 8; ++i) {
                const float arg = fmaf(__uint_as_float(sv[ch][8 * q + i]), c, -mc);
                // (interleaved: every MUFU.EX2 of the pair that feeds one bf16 pack has a polynomial neighbour)
                ev[i] = ((i * IDF_ATTN_POLY) % 8 < IDF_ATTN_POLY) ? poly_exp2(arg) : fast_exp2(arg);
                if constexpr (!LSUM) ps[i & 3] += ev[i];
              }
              pk[4 * q + 0] = pack_bf16x2(ev[0], ev[1]);
              pk[4 * q + 1] = pack_bf16x2(ev[2], ev[3]);
              pk[4 * q + 2] = pack_bf16x2(ev[4], ev[5]);
              pk[4 * q + 3] = pack_bf16x2(ev[6], ev[7]);
            }
            tmem_st_32x16(tmem_p + ch * 16, pk);
          }
        };
        float mx[4] = {-INFINITY, -INFINITY, -INFINITY, -INFINITY};
#pragma unroll
        for (int ch = 0; ch < 4; ++ch)
#pragma unroll
          for (int i = 0; i < 32; i += 2)
            mx[(i >> 1) & 3] = fmaxf(mx[(i >> 1) & 3],
                                     fmaxf(__uint_as_float(sv[ch][i]), __uint_as_float(sv[ch][i + 1])));  // FMNMX3
        const float mrow = fmaxf(fmaxf(mx[0], mx[1]), fmaxf(mx[2], mx[3]));

        // P_{j-1} V has to be complete before P is overwritten (and before the accumulator may be rescaled)
        SM_STAMP(3);  // row max done, waiting for P V(e-1)
        if (j > 0) mbar_wait(&o_full[g], (e - 1) & 1);
        SM_STAMP(4);  // P V(e-1) complete
        const bool grow = (mrow - m_run) * c > 8.0f;  // first block: m_run = -inf -> true
        if (__any_sync(0xffffffffu, grow)) {
          const float m_new = grow ? mrow : m_run;
          const float alpha = fast_exp2((m_run - m_new) * c);  // 1 for rows that keep their maximum
          m_run = m_new;
          l_run *= alpha;
          if (j > 0) rescale(alpha);
        }
        exp_store(m_run * c);
        SM_STAMP(5);  // exp phase issued
        tmem_st_wait();
        tc_fence_before_sync();
        __syncwarp();
        if (lane == 0) mbar_arrive(&p_full[g]);
        SM_STAMP(6);  // P(e) handed to the MMA thread
        l_run += (ps[0] + ps[1]) + (ps[2] + ps[3]);
      }
      // all of this item's P V products for tile g are complete: fetch the output rows, free the accumulator
      mbar_wait(&o_full[g], (e - 1) & 1);
      tc_fence_after_sync();
      float o_acc[HD];
#pragma unroll
      for (int d0 = 0; d0 < HD; d0 += 16) {
        uint32_t v[16];
        tmem_ld_32x16(tmem_o + d0, v);
        tmem_ld_wait();
#pragma unroll
        for (int d = 0; d < 16; ++d) o_acc[d0 + d] = __uint_as_float(v[d]);
      }
      if constexpr (LSUM) {
        uint32_t v[16];
        tmem_ld_32x16(tmem_o + 48, v);
        tmem_ld_wait();
        l_run = __uint_as_float(v[0]);
      }
      tc_fence_before_sync();
      __syncwarp();
      if (lane == 0) mbar_arrive(&o_empty[g]);
      const long long m = (long long)item_row0(item) + g * 128 + r;
      if (m < p.M) {
        const float inv = 1.f / l_run;
        if (p.lse != nullptr) p.lse[m * p.heads + item_head(item)] = fmaf(m_run, c, __log2f(l_run));
        __nv_bfloat16* dst = p.out + m * p.ld_out + item_head(item) * HD;
#pragma unroll
        for (int d0 = 0; d0 < HD; d0 += 8) {
          uint4 o;
          o.x = pack_bf16x2(o_acc[d0 + 0] * inv, o_acc[d0 + 1] * inv);
          o.y = pack_bf16x2(o_acc[d0 + 2] * inv, o_acc[d0 + 3] * inv);
          o.z = pack_bf16x2(o_acc[d0 + 4] * inv, o_acc[d0 + 5] * inv);
          o.w = pack_bf16x2(o_acc[d0 + 6] * inv, o_acc[d0 + 7] * inv);
          *reinterpret_cast<uint4*>(dst + d0) = o;
        }
      }
    }
  }

  tc_fence_before_sync();
  __syncthreads();
  if (warp == 9) {
    __syncwarp();
    tc_fence_after_sync();
    tmem_dealloc(tmem_base, ATTP_TMEM_COLS);
  }
}

template <int HD>
static int launch_attention_pipe(const AttnParams& p, int tiles, int heads, cudaStream_t stream) {
  constexpr int SWZ = HD <= 16 ? 32 : (HD <= 32 ? 64 : 128);
  // (sized for the larger token-major V tile so that one attribute setting serves both operand layouts)
  const int smem = 4 * 128 * SWZ + AttpK<HD>::STAGES * 128 * SWZ + ATTP_VSTAGES * 128 * 128 + 2048 + 1024 + 512;
  static bool attr_set = false;
  if (!attr_set) {
    int rc = check_cuda(cudaFuncSetAttribute(attention_pipe_kernel<HD>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem),
                        "attention_pipe: cudaFuncSetAttribute");
    if (rc != IDF_OK) return rc;
    attr_set = true;
  }
  const int items = ((tiles + 1) / 2) * heads;
  const int grid = items < sm_count() ? items : sm_count();
  return check_cuda(launch_kernel(attention_pipe_kernel<HD>, dim3(grid), dim3(ATTP_THREADS), smem, stream, p),
                    "attention_pipe launch");
}

template <int HD>
static int launch_attention(const AttnParams& p, int tiles, int heads, cudaStream_t stream) {
  constexpr int SWZ = HD <= 16 ? 32 : (HD <= 32 ? 64 : 128);
  const int smem = 128 * SWZ + 2 * 128 * 128 + p.kv_stages * (128 * SWZ + 128 * 128) + 1024 + 128;
  static int smem_set = 0;
  if (smem > smem_set) {
    int rc = check_cuda(cudaFuncSetAttribute(attention_kernel<HD>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem),
                        "attention: cudaFuncSetAttribute");
    if (rc != IDF_OK) return rc;
    smem_set = smem;
  }
  return check_cuda(launch_kernel(attention_kernel<HD>, dim3(tiles, heads), dim3(ATT_THREADS), smem, stream, p),
                    "attention launch");
}

}  // namespace idf

using namespace idf;

static int attention_fwd_impl(const void* qk, int64_t ld_qk, const void* vt, int64_t ld_vt, void* out,
                              int64_t ld_out, int32_t M, int32_t T, int32_t heads, int32_t head_dim, float scale,
                              float* lse, int v_tok, idf_stream_t stream) {
  if (!qk || !vt || !out) return fail(IDF_ERR_ARG, "attention: null pointer");
  if (T < 16 || (T & (T - 1)) != 0) return fail(IDF_ERR_UNSUPPORTED, "attention: T = %d must be a power of two >= 16", T);
  if (M <= 0 || M % T != 0) return fail(IDF_ERR_ARG, "attention: M = %d not a multiple of T = %d", M, T);
  const int C = heads * head_dim;
  if ((reinterpret_cast<uintptr_t>(out) & 15) || ld_out % 8 != 0 || head_dim % 8 != 0)
    return fail(IDF_ERR_ARG, "attention: output alignment");
  AttnParams p;
  memset(&p, 0, sizeof(p));
  p.out = reinterpret_cast<__nv_bfloat16*>(out);
  p.ld_out = ld_out;
  p.M = M; p.T = T; p.C = C;
  p.t_shift = 0;
  while ((1 << p.t_shift) < T) ++p.t_shift;
  p.nblk = T >= 128 ? T / 128 : 1;
  p.kv_stages = p.nblk > 1 ? 2 : 1;
  p.scale_log2e = scale * 1.4426950408889634f;
  p.heads = heads;
  p.lse = lse;
  p.v_tok = v_tok;
#ifdef IDF_ATTN_TRACE
  {
    static long long* trace_buf = nullptr;
    if (trace_buf == nullptr) {
      cudaMalloc(&trace_buf, 3 * ATT_TRACE_BLOCKS * ATT_TRACE_EVENTS * sizeof(long long));
      FILE* fh = fopen("/tmp/idf_attn_trace_ptr", "w");
      if (fh) { fprintf(fh, "%llu", (unsigned long long)(uintptr_t)trace_buf); fclose(fh); }
    }
    cudaMemsetAsync(trace_buf, 0, 3 * ATT_TRACE_BLOCKS * ATT_TRACE_EVENTS * sizeof(long long), reinterpret_cast<cudaStream_t>(stream));
    p.trace = trace_buf;
  }
#endif

  const int swz = head_dim <= 16 ? 32 : (head_dim <= 32 ? 64 : 128);
  const CUtensorMapSwizzle swz_enum = swz == 32 ? CU_TENSOR_MAP_SWIZZLE_32B
                                                : (swz == 64 ? CU_TENSOR_MAP_SWIZZLE_64B : CU_TENSOR_MAP_SWIZZLE_128B);
  int rc;
  {
    const uint64_t dims[2] = {(uint64_t)(2 * C), (uint64_t)M};
    const uint64_t strides[1] = {(uint64_t)ld_qk * 2};
    const uint32_t box[2] = {(uint32_t)(swz / 2), 128u};
    if ((rc = encode_tmap(&p.tmQK, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, qk, 2, dims, strides, box, swz_enum)) != IDF_OK)
      return rc;
  }
  if (v_tok) {  // vt == the (M, 3C) QKV matrix itself
    const uint64_t dims[2] = {(uint64_t)(3 * C), (uint64_t)M};
    const uint64_t strides[1] = {(uint64_t)ld_vt * 2};
    const uint32_t box[2] = {64u, 128u};
    if ((rc = encode_tmap(&p.tmVT, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, vt, 2, dims, strides, box,
                          CU_TENSOR_MAP_SWIZZLE_128B)) != IDF_OK)
      return rc;
  } else {
    const uint64_t dims[2] = {(uint64_t)M, (uint64_t)C};
    const uint64_t strides[1] = {(uint64_t)ld_vt * 2};
    const uint32_t box[2] = {64u, (uint32_t)head_dim};
    if ((rc = encode_tmap(&p.tmVT, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, vt, 2, dims, strides, box,
                          CU_TENSOR_MAP_SWIZZLE_128B)) != IDF_OK)
      return rc;
  }
  const int tiles = (M + 127) / 128;
  cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
  static const int pipe_min_blocks = [] {
    const char* e = getenv("IDF_ATTN_PIPE_MIN_BLOCKS");
    return e ? atoi(e) : 2;
  }();
  if (p.nblk >= pipe_min_blocks && T >= 256) {
    switch (head_dim) {
      case 16: return launch_attention_pipe<16>(p, tiles, heads, s);
      case 32: return launch_attention_pipe<32>(p, tiles, heads, s);
      case 48: return launch_attention_pipe<48>(p, tiles, heads, s);
      case 64: return launch_attention_pipe<64>(p, tiles, heads, s);
      default: break;
    }
  }
  switch (head_dim) {
    case 16: return launch_attention<16>(p, tiles, heads, s);
    case 32: return launch_attention<32>(p, tiles, heads, s);
    case 48: return launch_attention<48>(p, tiles, heads, s);
    case 64: return launch_attention<64>(p, tiles, heads, s);
    default: return fail(IDF_ERR_UNSUPPORTED, "attention: head_dim %d not in {16,32,48,64}", head_dim);
  }
}

extern "C" int idf_attention_fwd(const void* qk, int64_t ld_qk, const void* vt, int64_t ld_vt, void* out,
                                 int64_t ld_out, int32_t M, int32_t T, int32_t heads, int32_t head_dim, float scale,
                                 idf_stream_t stream) {
  return attention_fwd_impl(qk, ld_qk, vt, ld_vt, out, ld_out, M, T, heads, head_dim, scale, nullptr, 0, stream);
}

extern "C" int idf_attention_fwd_train(const void* qk, int64_t ld_qk, const void* vt, int64_t ld_vt, void* out,
                                       int64_t ld_out, int32_t M, int32_t T, int32_t heads, int32_t head_dim,
                                       float scale, float* lse, idf_stream_t stream) {
  if (!lse) return fail(IDF_ERR_ARG, "attention_fwd_train: lse is null");
  return attention_fwd_impl(qk, ld_qk, vt, ld_vt, out, ld_out, M, T, heads, head_dim, scale, lse, 0, stream);
}

extern "C" int idf_attention_fwd_qkv(const void* qkv, int64_t ld_qkv, void* out, int64_t ld_out, int32_t M, int32_t T,
                                     int32_t heads, int32_t head_dim, float scale, float* lse, idf_stream_t stream) {
  return attention_fwd_impl(qkv, ld_qkv, qkv, ld_qkv, out, ld_out, M, T, heads, head_dim, scale, lse, 1, stream);
}
