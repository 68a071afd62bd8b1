// Fused multi-head self-attention backward on tcgen05 (sm_100a): dQ, dK, dV from Q, K, V^T, dO, the forward's
// log-sum-exp and delta = rowsum(dO * O), with the probability tile recomputed on chip (never in HBM).
//
// One CTA owns one 128-key tile j of one head and walks over the query tiles i of the same sample(s). Everything is
// computed TRANSPOSED (keys along TMEM lanes, one key row per thread) so that P^T and dS^T come out of the softmax
// threads row-major in the layout the next three products need:
//   S^T  = K_j Q_i^T                    (A = K_j  K-major,  B = Q_i  K-major)          128 x 128 x hd
//   dP^T = V_j dO_i^T                   (A = V_j  K-major,  B = dO_i K-major)          128 x 128 x hd
//   P^T  = exp2(S^T * c - lse[q]),  dS^T = P^T * (dP^T - delta[q]) * scale             softmax threads -> bf16 smem
//   dV_j += P^T  dO_i                   (A = P^T  in TMEM,  B = dO_i MN-major)         128 x hd x 128   (TMEM, all i)
//   dK_j += dS^T Q_i                    (A = dS^T K-major,  B = Q_i  MN-major)         128 x hd x 128   (TMEM, all i)
//   dQ_i  = dS   K_j                    (A = dS^T MN-major, B = K_j  MN-major)         128 x hd x 128   per i
// Q, K, V are read straight from the token-major (M, 3C) QKV matrix. The MN-major descriptors let Q_i, dO_i and K_j
// be used in both operand roles from ONE shared-memory copy as TMA wrote it. dQ_i partials of different key tiles are added with vector fp32 reductions (red.global.add.v4.f32) when
// a sample has more than one key tile, and stored directly otherwise.
// Samples with fewer than 128 tokens share a tile under a block-diagonal mask, as in the forward kernel.
//
// Pipelining: every 128-query tile is processed as two 64-column halves ("mini-iterations"). S^T / dP^T live in two
// TMEM buffers and dS^T in four shared-memory blocks (tile parity x half), so the MMA thread issues the score products
// of mini-iteration m+1 BEFORE it waits for the softmax threads of m, and the accumulate products of m run while the
// softmax threads already work on m+1. dQ_i needs all 128 queries and is issued after the second half; its rows are
// read back one mini-iteration later.
//
// Warp roles (320 threads): warps 0..7 = softmax / output rows (warp w owns TMEM lanes 32*(w%4).. and the 32-column
// group w/4 of every 64-column half), warp 8 = TMA producer, warp 9 = TMEM + MMA issuer.
// TMEM columns: [S^T | dP^T] x 2 buffers [0,256), dV [256,256+hd), dK [320,320+hd), dQ [384,384+hd), P^T as packed bf16
// pairs [448,512) (32 columns per 64-query half): P^T never touches shared memory, it is the A operand of the dV
// product straight from tensor memory. dS^T has to live in shared memory because dQ needs it transposed (MN-major).
#include <stdlib.h>
#include <string.h>

#include "common.cuh"
#include "host.h"
#include "../../include/idf_b200.h"

namespace idf {

constexpr int ATB_THREADS = 320;
constexpr int ATB_TILE_BYTES = 128 * 128;  // 128 rows x 64 bf16 columns
constexpr int ATB_SMEM = 2 * ATB_TILE_BYTES /*K, V*/ + 4 * ATB_TILE_BYTES /*Q, dO x2 stages*/ +
                         4 * ATB_TILE_BYTES /*dS^T: tile parity x half*/ + 4 * 128 * 4 /*lse, delta x2*/ + 1024 + 256;

struct AttnBwdParams {
  CUtensorMap tmQK;  // (M, 3C) bf16 QKV matrix, box (64, 128)
  CUtensorMap tmDO;  // (M, C)  bf16, box (64, 128)
  const float* lse;
  const float* delta;
  __nv_bfloat16* dqkv;
  long long ld_dqkv;
  float* dq32;
  int M, T, C, t_shift, nblk, heads;
  float scale, scale_log2e;
};

__device__ __forceinline__ float atb_exp2(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

template <int HD>
__global__ void __launch_bounds__(ATB_THREADS, 1) attention_bwd_kernel(const __grid_constant__ AttnBwdParams p) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw_addr = smem_u32(smem_raw);
  uint8_t* smem = smem_raw + ((1024u - (raw_addr & 1023u)) & 1023u);
  uint8_t* smem_k = smem;
  uint8_t* smem_vt = smem_k + ATB_TILE_BYTES;
  uint8_t* smem_q = smem_vt + ATB_TILE_BYTES;       // [2 stages]
  uint8_t* smem_do = smem_q + 2 * ATB_TILE_BYTES;   // [2 stages]
  uint8_t* smem_ds = smem_do + 2 * ATB_TILE_BYTES;  // [2 tile parities][2 halves]
  float* s_lse = reinterpret_cast<float*>(smem_ds + 4 * ATB_TILE_BYTES);  // [2][128]
  float* s_delta = s_lse + 256;                                          // [2][128]
  uint64_t* bars = reinterpret_cast<uint64_t*>(s_delta + 256);
  uint64_t* kv_full = bars;
  uint64_t* q_full = bars + 1;      // [2] per Q / dO stage
  uint64_t* q_empty = bars + 3;     // [2]
  uint64_t* sdp_full = bars + 5;    // [2] per S / dP TMEM buffer
  uint64_t* sdp_empty = bars + 7;   // [2]
  uint64_t* pds_full = bars + 9;    // [2] per half
  uint64_t* p_empty = bars + 11;    // [2] per half: P block read by the dV products
  uint64_t* ds_empty = bars + 13;   // [2] per tile parity: dS blocks read by the dK / dQ products
  uint64_t* dq_full = bars + 15;
  uint64_t* dq_empty = bars + 16;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 17);

  const int warp = __shfl_sync(0xffffffffu, threadIdx.x >> 5, 0);
  const int lane = threadIdx.x & 31;
  const int head = blockIdx.y;
  const int row0 = blockIdx.x * 128;  // first key row of this tile
  const int q_base = (p.T >= 128) ? (row0 >> p.t_shift) << p.t_shift : row0;
  const int n = p.nblk;
  const int nm = 2 * n;  // mini-iterations

  if (warp == 8 && lane == 0) {
    tma_prefetch_desc(&p.tmQK);
    tma_prefetch_desc(&p.tmDO);
    mbar_init(kv_full, 1);
    for (int s = 0; s < 2; ++s) {
      mbar_init(&q_full[s], 1);
      mbar_init(&q_empty[s], 1);
      mbar_init(&sdp_full[s], 1);
      mbar_init(&sdp_empty[s], 8);
      mbar_init(&pds_full[s], 8);
      mbar_init(&p_empty[s], 1);
      mbar_init(&ds_empty[s], 1);
    }
    mbar_init(dq_full, 1);
    mbar_init(dq_empty, 8);
    fence_mbar_init();
  }
  if (warp == 9) {
    tmem_alloc(tmem_slot, 512);
    tmem_relinquish();
  }
  tc_fence_before_sync();
  __syncthreads();
  tc_fence_after_sync();
  const uint32_t tmem_base = *tmem_slot;
  const uint32_t tmem_dv = tmem_base + 256, tmem_dk = tmem_base + 320, tmem_dq = tmem_base + 384,
                 tmem_p = tmem_base + 448;

  if (warp == 8) {
    if (elect_one()) {
      mbar_expect_tx(kv_full, 2 * ATB_TILE_BYTES);
      tma_load_2d(smem_k, &p.tmQK, kv_full, p.C + head * HD, row0);
      tma_load_2d(smem_vt, &p.tmQK, kv_full, 2 * p.C + head * HD, row0);
      for (int i = 0; i < n; ++i) {
        const int st = i & 1;
        mbar_wait(&q_empty[st], ((i >> 1) & 1) ^ 1);
        mbar_expect_tx(&q_full[st], 2 * ATB_TILE_BYTES);
        tma_load_2d(smem_q + st * ATB_TILE_BYTES, &p.tmQK, &q_full[st], head * HD, q_base + i * 128);
        tma_load_2d(smem_do + st * ATB_TILE_BYTES, &p.tmDO, &q_full[st], head * HD, q_base + i * 128);
      }
    }
  } else if (warp == 9) {
    if (elect_one()) {
      constexpr uint32_t idesc_s = umma_idesc_bf16(128, 64, 0, 0);     // S^T / dP^T halves: 128 keys x 64 queries
      constexpr uint32_t idesc_acc = umma_idesc_bf16(128, HD, 0, 1);   // dV / dK: A K-major, B MN-major
      constexpr uint32_t idesc_dq = umma_idesc_bf16(128, HD, 1, 1);    // dQ: both MN-major
      const uint64_t dk_k = umma_desc_kmajor(smem_u32(smem_k), 128);
      const uint64_t dk_mn = umma_desc_mnmajor(smem_u32(smem_k), 8192, 1024);
      const uint64_t dv_k = umma_desc_kmajor(smem_u32(smem_vt), 128);
      mbar_wait(kv_full, 0);
      auto issue_sdp = [&](int m) {
        const int i = m >> 1, h = m & 1, st = i & 1, buf = m & 1;
        if (h == 0) mbar_wait(&q_full[st], (i >> 1) & 1);
        mbar_wait(&sdp_empty[buf], ((m >> 1) & 1) ^ 1);  // softmax threads have pulled the buffer's previous contents
        tc_fence_after_sync();
        const uint64_t dq_k = umma_desc_kmajor(smem_u32(smem_q + st * ATB_TILE_BYTES + h * 8192), 128);
        const uint64_t ddo_k = umma_desc_kmajor(smem_u32(smem_do + st * ATB_TILE_BYTES + h * 8192), 128);
        const uint32_t ts = tmem_base + buf * 128;
#pragma unroll
        for (int k = 0; k < HD / 16; ++k) umma_bf16(ts, dk_k + 2 * k, dq_k + 2 * k, idesc_s, k != 0);
#pragma unroll
        for (int k = 0; k < HD / 16; ++k) umma_bf16(ts + 64, dv_k + 2 * k, ddo_k + 2 * k, idesc_s, k != 0);
        umma_commit(&sdp_full[buf]);
      };
      issue_sdp(0);
      for (int m = 0; m < nm; ++m) {
        const int i = m >> 1, h = m & 1, st = i & 1;
        if (m + 1 < nm) issue_sdp(m + 1);  // scores of the next half run under the softmax of this one
        mbar_wait(&pds_full[h], i & 1);
        tc_fence_after_sync();
        const uint64_t dds_h = umma_desc_kmajor(smem_u32(smem_ds + ((i & 1) * 2 + h) * ATB_TILE_BYTES), 128);
        const uint64_t ddo_mn = umma_desc_mnmajor(smem_u32(smem_do + st * ATB_TILE_BYTES), 8192, 1024);
        const uint64_t dq_mn = umma_desc_mnmajor(smem_u32(smem_q + st * ATB_TILE_BYTES), 8192, 1024);
#pragma unroll
        for (int k = 0; k < 4; ++k)  // dV_j += P^T[:, half] dO_i[half]   (K = 64 queries)
          umma_bf16_ts(tmem_dv, tmem_p + h * 32 + 8 * k, ddo_mn + 128 * (4 * h + k), idesc_acc, (m > 0) || (k != 0));
        umma_commit(&p_empty[h]);
#pragma unroll
        for (int k = 0; k < 4; ++k)  // dK_j += dS^T[:, half] Q_i[half]
          umma_bf16(tmem_dk, dds_h + 2 * k, dq_mn + 128 * (4 * h + k), idesc_acc, (m > 0) || (k != 0));
        if (h == 1) {
          mbar_wait(dq_empty, (i & 1) ^ 1);  // rows of dQ_{i-1} have been read back
          tc_fence_after_sync();
          const uint64_t dds_mn = umma_desc_mnmajor(smem_u32(smem_ds + (i & 1) * 2 * ATB_TILE_BYTES), ATB_TILE_BYTES, 1024);
#pragma unroll
          for (int k = 0; k < 8; ++k)  // dQ_i = dS K_j     (K = 128 keys)
            umma_bf16(tmem_dq, dds_mn + 128 * k, dk_mn + 128 * k, idesc_dq, k != 0);
          umma_commit(&ds_empty[i & 1]);
          umma_commit(&q_empty[st]);
          umma_commit(dq_full);
        }
      }
    }
  } else {
    const int quad = warp & 3, cg = warp >> 2;
    const int r = quad * 32 + lane;  // key row of the tile == TMEM lane; also the query row when reading dQ
    const uint32_t lane_addr = (uint32_t)(quad * 32) << 16;
    const bool masked = p.T < 128;
    const int row_seg = r >> p.t_shift;
    const float c = p.scale_log2e;
    constexpr int HH = HD / 2;
    // per-query row terms of tile i: column group 0 stages lse, group 1 stages delta (one value per thread)
    auto fetch = [&](int i) {
      const long long m = (long long)q_base + i * 128 + r;
      if (m >= p.M) return 0.f;
      return cg == 0 ? p.lse[m * p.heads + head] : p.delta[m * p.heads + head];
    };
    auto dq_readout = [&](int t) {  // rows of dQ_t (lane = query row); each column group takes HD/2 columns
      mbar_wait(dq_full, t & 1);
      tc_fence_after_sync();
      float dqv[HH];
#pragma unroll
      for (int d0 = 0; d0 < HH; d0 += 8) {
        uint32_t v[8];
        asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
                     : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7])
                     : "r"(tmem_dq + lane_addr + cg * HH + d0)
                     : "memory");
        tmem_ld_wait();
#pragma unroll
        for (int d = 0; d < 8; ++d) dqv[d0 + d] = __uint_as_float(v[d]);
      }
      tc_fence_before_sync();
      __syncwarp();
      if (lane == 0) mbar_arrive(dq_empty);
      const long long m = (long long)q_base + t * 128 + r;
      if (m < p.M) {
        if (n == 1) {
          __nv_bfloat16* dst = p.dqkv + m * p.ld_dqkv + head * HD + cg * HH;
#pragma unroll
          for (int d0 = 0; d0 < HH; d0 += 8) {
            uint4 o;
            o.x = pack_bf16x2(dqv[d0 + 0], dqv[d0 + 1]);
            o.y = pack_bf16x2(dqv[d0 + 2], dqv[d0 + 3]);
            o.z = pack_bf16x2(dqv[d0 + 4], dqv[d0 + 5]);
            o.w = pack_bf16x2(dqv[d0 + 6], dqv[d0 + 7]);
            *reinterpret_cast<uint4*>(dst + d0) = o;
          }
        } else {
          float* dst = p.dq32 + m * p.C + head * HD + cg * HH;
#pragma unroll
          for (int d0 = 0; d0 < HH; d0 += 4)
            asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(dst + d0), "f"(dqv[d0]), "f"(dqv[d0 + 1]),
                         "f"(dqv[d0 + 2]), "f"(dqv[d0 + 3])
                         : "memory");
        }
      }
    };
    (cg == 0 ? s_lse : s_delta)[r] = fetch(0);
    named_bar_sync(1, 256);
    float next = 0.f;
    for (int m = 0; m < nm; ++m) {
      const int i = m >> 1, h = m & 1, buf = m & 1;
      if (h == 0 && i + 1 < n) next = fetch(i + 1);
      const float* lq = s_lse + (i & 1) * 128 + h * 64 + cg * 32;
      const float* dq = s_delta + (i & 1) * 128 + h * 64 + cg * 32;
      mbar_wait(&sdp_full[buf], (m >> 1) & 1);
      tc_fence_after_sync();
      uint32_t sv[32], dv[32];
      tmem_ld_32x32(tmem_base + buf * 128 + lane_addr + cg * 32, sv);
      tmem_ld_32x32(tmem_base + buf * 128 + 64 + lane_addr + cg * 32, dv);
      tmem_ld_wait();
      tc_fence_before_sync();
      __syncwarp();
      if (lane == 0) mbar_arrive(&sdp_empty[buf]);
      float pv[32], gv[32];
#pragma unroll
      for (int j = 0; j < 32; ++j) {
        float e = atb_exp2(fmaf(__uint_as_float(sv[j]), c, -lq[j]));
        if (masked && ((h * 64 + cg * 32 + j) >> p.t_shift) != row_seg) e = 0.f;
        pv[j] = e;
        gv[j] = e * (__uint_as_float(dv[j]) - dq[j]) * p.scale;
      }
      // the blocks about to be overwritten must have been consumed: P block h by the dV products of m-2, the dS
      // blocks of this tile parity by the dK / dQ products of tile i-2
      mbar_wait(&p_empty[h], ((m >> 1) & 1) ^ 1);
      mbar_wait(&ds_empty[i & 1], ((i >> 1) & 1) ^ 1);
      uint8_t* grow = smem_ds + ((i & 1) * 2 + h) * ATB_TILE_BYTES + r * 128;
      uint32_t pk[16];
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        pk[4 * q + 0] = pack_bf16x2(pv[8 * q + 0], pv[8 * q + 1]);
        pk[4 * q + 1] = pack_bf16x2(pv[8 * q + 2], pv[8 * q + 3]);
        pk[4 * q + 2] = pack_bf16x2(pv[8 * q + 4], pv[8 * q + 5]);
        pk[4 * q + 3] = pack_bf16x2(pv[8 * q + 6], pv[8 * q + 7]);
        const int chunk = (cg * 4 + q) ^ (r & 7);
        uint4 o;
        o.x = pack_bf16x2(gv[8 * q + 0], gv[8 * q + 1]);
        o.y = pack_bf16x2(gv[8 * q + 2], gv[8 * q + 3]);
        o.z = pack_bf16x2(gv[8 * q + 4], gv[8 * q + 5]);
        o.w = pack_bf16x2(gv[8 * q + 6], gv[8 * q + 7]);
        *reinterpret_cast<uint4*>(grow + chunk * 16) = o;
      }
      tmem_st_32x16(tmem_p + lane_addr + h * 32 + cg * 16, pk);  // P^T[key r][32 queries] as 16 packed columns
      tmem_st_wait();
      tc_fence_before_sync();
      fence_proxy_async_smem();
      __syncwarp();
      if (lane == 0) mbar_arrive(&pds_full[h]);
      if (h == 0) {
        if (i > 0) dq_readout(i - 1);  // one half-tile late: the dQ products have had time to finish
      } else {
        if (i + 1 < n) (cg == 0 ? s_lse : s_delta)[((i + 1) & 1) * 128 + r] = next;
        named_bar_sync(1, 256);
      }
    }
    dq_readout(n - 1);
    // dK_j rows (column group 0) / dV_j rows (group 1), lane = key row; the last dq_full covers every product
    tc_fence_after_sync();
    const long long m = (long long)row0 + r;
    {
      float acc[HD];
#pragma unroll
      for (int d0 = 0; d0 < HD; d0 += 16) {
        uint32_t v[16];
        tmem_ld_32x16((cg == 0 ? tmem_dk : tmem_dv) + lane_addr + d0, v);
        tmem_ld_wait();
#pragma unroll
        for (int d = 0; d < 16; ++d) acc[d0 + d] = __uint_as_float(v[d]);
      }
      if (m < p.M) {
        __nv_bfloat16* dst = p.dqkv + m * p.ld_dqkv + (cg == 0 ? p.C : 2 * p.C) + head * HD;
#pragma unroll
        for (int d0 = 0; d0 < HD; d0 += 8) {
          uint4 o;
          o.x = pack_bf16x2(acc[d0 + 0], acc[d0 + 1]);
          o.y = pack_bf16x2(acc[d0 + 2], acc[d0 + 3]);
          o.z = pack_bf16x2(acc[d0 + 4], acc[d0 + 5]);
          o.w = pack_bf16x2(acc[d0 + 6], acc[d0 + 7]);
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
    tmem_dealloc(tmem_base, 512);
  }
}

template <int HD>
static int launch_attention_bwd(const AttnBwdParams& p, int tiles, cudaStream_t stream) {
  static bool attr_set = false;
  if (!attr_set) {
    int rc = check_cuda(cudaFuncSetAttribute(attention_bwd_kernel<HD>, cudaFuncAttributeMaxDynamicSharedMemorySize, ATB_SMEM),
                        "attention_bwd: cudaFuncSetAttribute");
    if (rc != IDF_OK) return rc;
    attr_set = true;
  }
  attention_bwd_kernel<HD><<<dim3(tiles, p.heads), ATB_THREADS, ATB_SMEM, stream>>>(p);
  return check_cuda(cudaGetLastError(), "attention_bwd launch");
}

}  // namespace idf

using namespace idf;

extern "C" int idf_attention_bwd(const void* qkv, int64_t ld_qkv, const void* d_out, int64_t ld_do, const float* lse,
                                 const float* delta, void* dqkv, int64_t ld_dqkv, float* dq32, int32_t M, int32_t T,
                                 int32_t heads, int32_t head_dim, float scale, idf_stream_t stream) {
  if (!qkv || !d_out || !lse || !delta || !dqkv) return fail(IDF_ERR_ARG, "attention_bwd: null pointer");
  if (T < 16 || (T & (T - 1)) != 0) return fail(IDF_ERR_UNSUPPORTED, "attention_bwd: T = %d must be a power of two >= 16", T);
  if (M <= 0 || M % T != 0) return fail(IDF_ERR_ARG, "attention_bwd: M = %d not a multiple of T = %d", M, T);
  const int C = heads * head_dim;
  if ((reinterpret_cast<uintptr_t>(dqkv) & 15) || ld_dqkv % 8 != 0) return fail(IDF_ERR_ARG, "attention_bwd: output alignment");
  AttnBwdParams p;
  memset(&p, 0, sizeof(p));
  p.lse = lse; p.delta = delta;
  p.dqkv = reinterpret_cast<__nv_bfloat16*>(dqkv);
  p.ld_dqkv = ld_dqkv;
  p.dq32 = dq32;
  p.M = M; p.T = T; p.C = C; p.heads = heads;
  while ((1 << p.t_shift) < T) ++p.t_shift;
  p.nblk = T >= 128 ? T / 128 : 1;
  if (p.nblk > 1 && (dq32 == nullptr || (reinterpret_cast<uintptr_t>(dq32) & 15)))
    return fail(IDF_ERR_ARG, "attention_bwd: T > 128 needs a zeroed 16-byte aligned fp32 (M, C) dq32 accumulator");
  p.scale = scale;
  p.scale_log2e = scale * 1.4426950408889634f;
  int rc;
  {
    const uint64_t dims[2] = {(uint64_t)(3 * C), (uint64_t)M};
    const uint64_t strides[1] = {(uint64_t)ld_qkv * 2};
    const uint32_t box[2] = {64u, 128u};
    if ((rc = encode_tmap(&p.tmQK, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, qkv, 2, dims, strides, box,
                          CU_TENSOR_MAP_SWIZZLE_128B)) != IDF_OK)
      return rc;
  }
  {
    const uint64_t dims[2] = {(uint64_t)C, (uint64_t)M};
    const uint64_t strides[1] = {(uint64_t)ld_do * 2};
    const uint32_t box[2] = {64u, 128u};
    if ((rc = encode_tmap(&p.tmDO, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, d_out, 2, dims, strides, box,
                          CU_TENSOR_MAP_SWIZZLE_128B)) != IDF_OK)
      return rc;
  }
  const int tiles = (M + 127) / 128;
  cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
  switch (head_dim) {
    case 16: return launch_attention_bwd<16>(p, tiles, s);
    case 32: return launch_attention_bwd<32>(p, tiles, s);
    case 48: return launch_attention_bwd<48>(p, tiles, s);
    case 64: return launch_attention_bwd<64>(p, tiles, s);
    default: return fail(IDF_ERR_UNSUPPORTED, "attention_bwd: head_dim %d not in {16,32,48,64}", head_dim);
  }
}
