// Weight gradient of a convolution / linear layer as an implicit GEMM on tcgen05 (sm_100a).
//
//   dW[co, (tap, ci)] = sum_m dY[m, co] * X[pixel(m) + tap, ci]        (reduction over all B*H*W output pixels)
//
// Both operands are consumed exactly as the forward pass stored them (channels-last, pixels along rows): the
// reduction index m is the ROW index of both, so they are "MN-major" UMMA operands. A TMA box {64 channels, 64
// pixels} lands in shared memory as 64 rows of 128 bytes (128-byte swizzle) and is described to the tensor core
// with an MN-major descriptor; no transposed copy of any activation is ever made. The tap shift of the X operand
// is applied by the TMA coordinates (out-of-bounds = zero = the convolution's padding), one 64-channel chunk at a
// time, so every 64-column chunk of an output tile may belong to a different filter tap.
//
// Output tile 128 (co) x BN ((tap, ci) columns); K = pixels, split over `splits` work units per tile so that the
// ~10..100 output tiles of a layer fill 148 SMs. Partials go to an fp32 workspace; wgrad_finish_kernel adds them
// in split order (deterministic) and scatters into PyTorch's OIHW parameter layout.
//
// Warp roles (224 threads): warp 0 = TMA producer (dY chunks + first half of the X chunks), warp 6 = second TMA
// producer (the other X chunks: six small boxes per k-block are too many for one issuing thread), warp 1 = TMEM
// allocator + MMA issuer, warps 2..5 = epilogue.
#include <stdlib.h>
#include <string.h>

#include "common.cuh"
#include "host.h"
#include "../../include/idf_b200.h"

namespace idf {

constexpr int WG_BM = 128;
constexpr int WG_BK = 64;                 // pixels per k-block
constexpr int WG_CHUNK_BYTES = 64 * 128;  // one {64 channels x 64 pixels} box
constexpr int WG_THREADS = 224;

struct WgradParams {
  CUtensorMap tmDY;  // 2-D (M, Cout), box {64, 64}
  CUtensorMap tmX;   // 4-D NHWC, box {64, tile_w, tile_h, tile_n} = 64 pixels
  int taps, cb;      // filter taps, 64-channel blocks per tap
  signed char dh[9], dw[9];
  int dn[9];
  int tile_h, tile_n, tiles_per_img, matrix;
  int kb_total;      // ceil(M / 64)
  int m_tiles;       // Cout / 128
  int N;             // taps * Cin
  int splits;
  float* ws;         // [splits][Cout][N]
  long long split_stride;
};

template <int BN>
struct WgCfg {
  static constexpr int A_BYTES = 2 * WG_CHUNK_BYTES;
  static constexpr int B_BYTES = (BN / 64) * WG_CHUNK_BYTES;
  static constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
  static constexpr int STAGES = BN == 256 ? 4 : (BN == 192 ? 5 : 6);
  static constexpr int TMEM_COLS = BN == 128 ? 256 : 512;
  static constexpr int SMEM = STAGES * STAGE_BYTES + 1024 + 256;
};

template <int BN>
__global__ void __launch_bounds__(WG_THREADS, 1) wgrad_kernel(const __grid_constant__ WgradParams p) {
  using Cfg = WgCfg<BN>;
  constexpr int STAGES = Cfg::STAGES;
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw_addr = smem_u32(smem_raw);
  uint8_t* smem = smem_raw + ((1024u - (raw_addr & 1023u)) & 1023u);
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + STAGES * Cfg::STAGE_BYTES);
  uint64_t* full_bar = bars;
  uint64_t* empty_bar = bars + STAGES;
  uint64_t* tmem_full = bars + 2 * STAGES;   // [2]
  uint64_t* tmem_empty = tmem_full + 2;      // [2]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tmem_empty + 2);

  const int warp = __shfl_sync(0xffffffffu, threadIdx.x >> 5, 0);
  const int lane = threadIdx.x & 31;
  const int n_tiles = p.N / BN;
  const int splits = p.splits;
  const int total = p.m_tiles * n_tiles * splits;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&p.tmDY);
    tma_prefetch_desc(&p.tmX);
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(&full_bar[s], 2);  // one arrive.expect_tx per producer warp
      mbar_init(&empty_bar[s], 1);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&tmem_full[i], 1);
      mbar_init(&tmem_empty[i], 4);
    }
    fence_mbar_init();
  }
  if (warp == 1) {
    tmem_alloc(tmem_slot, Cfg::TMEM_COLS);
    tmem_relinquish();
  }
  tc_fence_before_sync();
  __syncthreads();
  tc_fence_after_sync();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0 || warp == 6) {
    if (elect_one()) {
      constexpr int NCH = BN / 64;        // X chunks per stage
      constexpr int C_SPLIT = NCH / 2;    // warp 0 loads dY + chunks [0, C_SPLIT), warp 6 chunks [C_SPLIT, NCH)
      const bool first = warp == 0;
      const int c_lo = first ? 0 : C_SPLIT, c_hi = first ? C_SPLIT : NCH;
      const uint32_t my_bytes = (uint32_t)(c_hi - c_lo) * WG_CHUNK_BYTES + (first ? Cfg::A_BYTES : 0);
      // The issuing thread is the kernel's critical path (six small boxes per 512-clock k-block): everything that
      // does not change from one k-block to the next - the chunks' tap offsets and channel origins - is computed once
      // per tile, and the pixel origin / ring position advance incrementally (no divisions inside the K loop).
      int stage = 0;
      uint32_t phase = 0;
      for (int u = blockIdx.x; u < total; u += gridDim.x) {
        const int t = u / splits, sp = u % splits;
        const int co0 = (t / n_tiles) * WG_BM;
        const int g0 = (t % n_tiles) * NCH;  // first 64-column chunk of this tile
        const int kb0 = (int)((long long)sp * p.kb_total / splits), kb1 = (int)((long long)(sp + 1) * p.kb_total / splits);
        int c_ci[NCH], c_dw[NCH], c_dh[NCH], c_dn[NCH];
#pragma unroll
        for (int c = 0; c < NCH; ++c) {
          const int g = g0 + c;
          const int tap = g / p.cb;
          c_ci[c] = (g - tap * p.cb) * 64;
          c_dw[c] = p.dw[tap]; c_dh[c] = p.dh[tap]; c_dn[c] = p.dn[tap];
        }
        int img0 = 0, h0 = 0, w0 = 0, tin = 0;
        if (p.matrix) w0 = kb0 * WG_BK;
        else if (p.tiles_per_img > 0) { img0 = kb0 / p.tiles_per_img; tin = kb0 - img0 * p.tiles_per_img; h0 = tin * p.tile_h; }
        else img0 = kb0 * p.tile_n;
        for (int kb = kb0; kb < kb1; ++kb) {
          mbar_wait(&empty_bar[stage], phase ^ 1);
          mbar_expect_tx(&full_bar[stage], my_bytes);
          uint8_t* sa = smem + stage * Cfg::STAGE_BYTES;
          uint8_t* sb = sa + Cfg::A_BYTES;
          if (first) {
            tma_load_2d(sa, &p.tmDY, &full_bar[stage], co0, kb * WG_BK);
            tma_load_2d(sa + WG_CHUNK_BYTES, &p.tmDY, &full_bar[stage], co0 + 64, kb * WG_BK);
          }
#pragma unroll
          for (int c = 0; c < NCH; ++c)
            if (c >= c_lo && c < c_hi)
              tma_load_4d(sb + c * WG_CHUNK_BYTES, &p.tmX, &full_bar[stage], c_ci[c], w0 + c_dw[c], h0 + c_dh[c],
                          img0 + c_dn[c]);
          if (++stage == STAGES) { stage = 0; phase ^= 1; }
          if (p.matrix) w0 += WG_BK;
          else if (p.tiles_per_img > 0) {
            h0 += p.tile_h;
            if (++tin == p.tiles_per_img) { tin = 0; h0 = 0; ++img0; }
          } else img0 += p.tile_n;
        }
      }
    }
  } else if (warp == 1) {
    if (elect_one()) {
      constexpr uint32_t idesc = umma_idesc_bf16(WG_BM, BN, 1, 1);
      int it = 0, stage = 0;
      uint32_t phase = 0;
      for (int u = blockIdx.x; u < total; u += gridDim.x, ++it) {
        const int sp = u % splits;
        const int kb0 = (int)((long long)sp * p.kb_total / splits), kb1 = (int)((long long)(sp + 1) * p.kb_total / splits);
        const int acc = it & 1;
        mbar_wait(&tmem_empty[acc], ((it >> 1) & 1) ^ 1);
        tc_fence_after_sync();
        const uint32_t tmem_d = tmem_base + acc * BN;
        for (int kb = kb0; kb < kb1; ++kb) {
          const int s = stage;
          mbar_wait(&full_bar[s], phase);
          if (++stage == STAGES) { stage = 0; phase ^= 1; }
          tc_fence_after_sync();
          const uint32_t sa = smem_u32(smem + s * Cfg::STAGE_BYTES);
          const uint64_t da = umma_desc_mnmajor(sa, WG_CHUNK_BYTES, 1024);
          const uint64_t db = umma_desc_mnmajor(sa + Cfg::A_BYTES, WG_CHUNK_BYTES, 1024);
#pragma unroll
          for (int k = 0; k < WG_BK / 16; ++k)  // 16 pixels = 16 rows of 128 bytes = +128 in the 16-byte address field
            umma_bf16(tmem_d, da + 128 * k, db + 128 * k, idesc, (kb > kb0) || (k != 0));
          umma_commit(&empty_bar[s]);
        }
        umma_commit(&tmem_full[acc]);
      }
    }
  } else if (warp < 6) {
    const int quad = warp & 3;
    const int r = quad * 32 + lane;
    const uint32_t lane_addr = (uint32_t)(quad * 32) << 16;
    int it = 0;
    for (int u = blockIdx.x; u < total; u += gridDim.x, ++it) {
      const int t = u / splits, sp = u % splits;
      const int co = (t / n_tiles) * WG_BM + r;
      const int n0 = (t % n_tiles) * BN;
      const int acc = it & 1;
      float* dst = p.ws + (long long)sp * p.split_stride + (long long)co * p.N + n0;
      mbar_wait(&tmem_full[acc], (it >> 1) & 1);
      tc_fence_after_sync();
      const uint32_t tmem_d = tmem_base + acc * BN + lane_addr;
      const int kb0 = (int)((long long)sp * p.kb_total / splits), kb1 = (int)((long long)(sp + 1) * p.kb_total / splits);
#pragma unroll 1
      for (int c = 0; c < BN / 32; ++c) {
        uint32_t v[32];
        if (kb1 > kb0) {
          tmem_ld_32x32(tmem_d + c * 32, v);
          tmem_ld_wait();
        } else {
#pragma unroll
          for (int j = 0; j < 32; ++j) v[j] = 0u;
        }
        if (c == BN / 32 - 1) {
          tc_fence_before_sync();
          __syncwarp();
          if (lane == 0) mbar_arrive(&tmem_empty[acc]);
        }
        float4* d4 = reinterpret_cast<float4*>(dst + c * 32);
#pragma unroll
        for (int j = 0; j < 8; ++j)
          d4[j] = make_float4(__uint_as_float(v[4 * j]), __uint_as_float(v[4 * j + 1]), __uint_as_float(v[4 * j + 2]),
                              __uint_as_float(v[4 * j + 3]));
      }
    }
  }

  tc_fence_before_sync();
  __syncthreads();
  if (warp == 1) {
    __syncwarp();
    tc_fence_after_sync();
    tmem_dealloc(tmem_base, Cfg::TMEM_COLS);
  }
}

// grad[(co * Cin + ci) * taps + tap] (+)= sum_s ws[s][co][tap * Cin + ci]. One thread per FOUR consecutive ci of one
// (co, tap): the partials are read as float4 in the workspace's own order (perfectly coalesced), four splits in flight
// at a time, and summed in split order (same bits as a sequential sum). The earlier version gave a thread all taps and
// all splits of one (co, ci): 216 dependent scalar loads per thread and only Cout * Cin / 256 blocks - 25 us of pure
// latency for the small layers (Cout = Cin = 128: 64 blocks).
__global__ void __launch_bounds__(256) wgrad_finish_kernel(const float* __restrict__ ws, long long split_stride,
                                                           int splits, float* __restrict__ grad, int Cout, int Cin,
                                                           int taps, int accumulate) {
  const long long N = (long long)taps * Cin;
  const long long j = ((long long)blockIdx.x * blockDim.x + threadIdx.x) * 4;
  if (j >= (long long)Cout * N) return;
  const float* src = ws + j;
  float4 acc = *reinterpret_cast<const float4*>(src);
  int s = 1;
  for (; s + 3 < splits; s += 4) {
    const float4 p0 = *reinterpret_cast<const float4*>(src + (long long)s * split_stride);
    const float4 p1 = *reinterpret_cast<const float4*>(src + (long long)(s + 1) * split_stride);
    const float4 p2 = *reinterpret_cast<const float4*>(src + (long long)(s + 2) * split_stride);
    const float4 p3 = *reinterpret_cast<const float4*>(src + (long long)(s + 3) * split_stride);
    acc.x = ((acc.x + p0.x) + p1.x) + p2.x + p3.x;
    acc.y = ((acc.y + p0.y) + p1.y) + p2.y + p3.y;
    acc.z = ((acc.z + p0.z) + p1.z) + p2.z + p3.z;
    acc.w = ((acc.w + p0.w) + p1.w) + p2.w + p3.w;
  }
  for (; s < splits; ++s) {
    const float4 p0 = *reinterpret_cast<const float4*>(src + (long long)s * split_stride);
    acc.x += p0.x; acc.y += p0.y; acc.z += p0.z; acc.w += p0.w;
  }
  const long long co = j / N;
  const int r = (int)(j - co * N);
  const int t = r / Cin, ci = r - t * Cin;  // Cin % 64 == 0: the four elements share (co, t)
  float* dst = grad + (co * Cin + ci) * taps + t;
  if (accumulate) {
    dst[0] += acc.x; dst[taps] += acc.y; dst[2 * taps] += acc.z; dst[3 * taps] += acc.w;
  } else {
    dst[0] = acc.x; dst[taps] = acc.y; dst[2 * taps] = acc.z; dst[3 * taps] = acc.w;
  }
}

template <int BN>
static int launch_wgrad(const WgradParams& p, cudaStream_t stream) {
  using Cfg = WgCfg<BN>;
  static bool attr_set = false;
  if (!attr_set) {
    int rc = check_cuda(cudaFuncSetAttribute(wgrad_kernel<BN>, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::SMEM),
                        "wgrad: cudaFuncSetAttribute");
    if (rc != IDF_OK) return rc;
    attr_set = true;
  }
  const int total = p.m_tiles * (p.N / BN) * p.splits;
  const int grid = total < sm_count() ? total : sm_count();
  wgrad_kernel<BN><<<grid, WG_THREADS, Cfg::SMEM, stream>>>(p);
  return check_cuda(cudaGetLastError(), "wgrad launch");
}

}  // namespace idf

using namespace idf;

extern "C" int idf_conv2d_wgrad(const idf_wgrad_args* a, idf_stream_t stream) {
  if (a == nullptr || a->x.ptr == nullptr || a->dy == nullptr || a->grad == nullptr || a->ws == nullptr)
    return fail(IDF_ERR_ARG, "wgrad: null argument");
  const idf_nhwc_t& x = a->x;
  if (x.c <= 0 || x.c % 64 != 0) return fail(IDF_ERR_UNSUPPORTED, "wgrad: Cin = %d not a multiple of 64", x.c);
  if (a->cout <= 0 || a->cout % WG_BM != 0) return fail(IDF_ERR_UNSUPPORTED, "wgrad: Cout = %d not a multiple of 128", a->cout);
  if (a->taps != 1 && a->taps != 9) return fail(IDF_ERR_ARG, "wgrad: taps must be 1 or 9");
  if (a->s2_batch > 0 && (x.n != 4 * a->s2_batch || a->taps != 9))
    return fail(IDF_ERR_ARG, "wgrad: s2_batch needs a 9-tap input holding 4*s2_batch parity planes");
  WgradParams p;
  memset(&p, 0, sizeof(p));
  const int H = x.h, W = x.w, HW = H * W;
  const long long M = (long long)(a->s2_batch > 0 ? a->s2_batch : x.n) * HW;
  if (M <= 0 || M > 0x7fffffffLL) return fail(IDF_ERR_ARG, "wgrad: bad M");
  int tile_w;
  const bool is_matrix = x.n == 1 && x.h == 1 && a->taps == 1;
  if (is_matrix) {
    p.matrix = 1; tile_w = WG_BK; p.tile_h = 1; p.tile_n = 1;
  } else if (W > WG_BK) {
    return fail(IDF_ERR_UNSUPPORTED, "wgrad: image width %d > 64", W);
  } else if (HW >= WG_BK) {
    if (WG_BK % W != 0 || HW % WG_BK != 0) return fail(IDF_ERR_UNSUPPORTED, "wgrad: %dx%d image does not tile", H, W);
    tile_w = W; p.tile_h = WG_BK / W; p.tile_n = 1; p.tiles_per_img = HW / WG_BK;
  } else {
    if (WG_BK % HW != 0) return fail(IDF_ERR_UNSUPPORTED, "wgrad: %dx%d image does not tile", H, W);
    tile_w = W; p.tile_h = H; p.tile_n = WG_BK / HW;
  }
  int rc;
  {
    const uint64_t dims[4] = {(uint64_t)x.c, (uint64_t)x.w, (uint64_t)x.h, (uint64_t)x.n};
    const uint64_t strides[3] = {(uint64_t)x.sw * 2, (uint64_t)x.sh * 2, (uint64_t)x.sn * 2};
    const uint32_t box[4] = {64u, (uint32_t)tile_w, (uint32_t)p.tile_h, (uint32_t)p.tile_n};
    if ((rc = encode_tmap(&p.tmX, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, x.ptr, 4, dims, strides, box,
                          CU_TENSOR_MAP_SWIZZLE_128B)) != IDF_OK)
      return rc;
  }
  {
    const uint64_t dims[2] = {(uint64_t)a->cout, (uint64_t)M};
    const uint64_t strides[1] = {(uint64_t)a->ld_dy * 2};
    const uint32_t box[2] = {64u, 64u};
    if ((rc = encode_tmap(&p.tmDY, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, a->dy, 2, dims, strides, box,
                          CU_TENSOR_MAP_SWIZZLE_128B)) != IDF_OK)
      return rc;
  }
  p.taps = a->taps;
  p.cb = x.c / 64;
  for (int t = 0; t < a->taps; ++t) {
    if (a->taps == 1) { p.dh[t] = 0; p.dw[t] = 0; p.dn[t] = 0; }
    else if (a->s2_batch > 0) {
      const int kh = t / 3, kw = t % 3;
      p.dn[t] = ((kh & 1) * 2 + (kw & 1)) * a->s2_batch;
      p.dh[t] = (signed char)(kh >> 1);
      p.dw[t] = (signed char)(kw >> 1);
    } else { p.dh[t] = (signed char)(t / 3 - 1); p.dw[t] = (signed char)(t % 3 - 1); p.dn[t] = 0; }
  }
  p.kb_total = (int)((M + WG_BK - 1) / WG_BK);
  p.m_tiles = a->cout / WG_BM;
  p.N = a->taps * x.c;
  int bn = 0;
  {
    static const int force_bn = [] { const char* e = getenv("IDF_WGRAD_BN"); return e ? atoi(e) : 0; }();
    const int cands[3] = {256, 192, 128};
    for (int i = 0; i < 3 && bn == 0; ++i)
      if (p.N % cands[i] == 0 && (force_bn == 0 || cands[i] <= force_bn)) bn = cands[i];
  }
  if (bn == 0) return fail(IDF_ERR_UNSUPPORTED, "wgrad: N = %d has no legal tile width", p.N);
  const long long per_split = (long long)a->cout * p.N * 4;
  const int units = p.m_tiles * (p.N / bn);
  // one wave: the largest split count whose work units still fit the SM count (a second, partly filled wave would
  // cost a full pass of the slowest unit)
  int splits = units >= sm_count() ? 1 : sm_count() / units;
  if (splits > p.kb_total / 2) splits = p.kb_total / 2;
  if (splits > 64) splits = 64;
  {
    static const int force_splits = [] { const char* e = getenv("IDF_WGRAD_SPLITS"); return e ? atoi(e) : 0; }();
    if (force_splits > 0) splits = force_splits;
  }
  if (splits < 1) splits = 1;
  if (per_split * splits > a->ws_bytes) splits = (int)(a->ws_bytes / per_split);
  if (splits < 1) return fail(IDF_ERR_ARG, "wgrad: workspace of %lld bytes is smaller than one partial (%lld)",
                              (long long)a->ws_bytes, per_split);
  if (reinterpret_cast<uintptr_t>(a->ws) & 15) return fail(IDF_ERR_ARG, "wgrad: workspace must be 16-byte aligned");
  p.splits = splits;
  p.ws = reinterpret_cast<float*>(a->ws);
  p.split_stride = (long long)a->cout * p.N;
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  switch (bn) {
    case 256: rc = launch_wgrad<256>(p, st); break;
    case 192: rc = launch_wgrad<192>(p, st); break;
    default: rc = launch_wgrad<128>(p, st); break;
  }
  if (rc != IDF_OK) return rc;
  const long long cells = (long long)a->cout * x.c * a->taps / 4;  // one thread per float4 of the (Cout, taps * Cin) partial
  wgrad_finish_kernel<<<(unsigned)((cells + 255) / 256), 256, 0, st>>>(p.ws, p.split_stride, splits, a->grad, a->cout,
                                                                       x.c, a->taps, a->accumulate ? 1 : 0);
  return check_cuda(cudaGetLastError(), "wgrad_finish launch");
}
