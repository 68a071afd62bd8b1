// Implicit-GEMM convolution / linear layer on tcgen05 tensor cores (sm_100a).
//
//   D[m, n] = sum_k A[m, k] * W[n, k]      m = pixel (n, h, w) of a channels-last activation, k = (tap, cin)
//
// A tiles are never materialised: for every filter tap the TMA engine fetches a (tile_n, tile_h, tile_w, 64ch)
// box of the NHWC activation at the tap's (dh, dw) offset; coordinates that fall outside the image are
// zero-filled by the TMA unit, which is exactly the convolution's zero padding. The box lands in shared memory
// as 128 rows x 128 bytes with the 128-byte swizzle, i.e. the canonical K-major UMMA operand. Weights are a plain
// 2-D (N, K) bf16 matrix. One elected thread issues tcgen05.mma (128 x BLOCK_N x 16) with the fp32 accumulator
// in tensor memory; four epilogue warps read it back (one accumulator row per thread), fuse bias / time bias /
// residual / padding mask, and hand a bf16 tile to a TMA store.
//
// Warp roles (192 threads): warp 0 = TMA producer, warp 1 = TMEM allocator + MMA issuer, warps 2..5 = epilogue.
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "common.cuh"
#include "host.h"
#include "../../include/idf_b200.h"

namespace idf {

constexpr int BLOCK_M = 128;
constexpr int BLOCK_N = 128;
constexpr int BLOCK_K = 64;  // 64 bf16 = 128 bytes = one swizzle row
constexpr int A_STAGE_BYTES = BLOCK_M * BLOCK_K * 2;  // 16 KiB

enum : int { F_RES = 1, F_ZERO_PAD = 2, F_VT = 4, F_OUT_F32 = 8, F_OUT_UP2 = 16 };

// A partial-sum record of the GroupNorm-fused epilogue (see gn_record_load / gn_record_store)
struct GnRecord { unsigned long long lo, hi; };

struct IgemmParams {
  CUtensorMap tmA[2];
  CUtensorMap tmB;
  CUtensorMap tmC;
  CUtensorMap tmCx[3];  // up2_all: output maps of parities 1..3 (tmC is parity 0)
  CUtensorMap tmR;
  int kb_seg0;   // k-blocks (of 64) in segment 0
  int kb_total;  // k-blocks in both segments
  int cb[2];     // channel blocks per tap, per segment
  int taps[2];
  int H, W, HW;
  int tile_w, tile_h, tile_n;
  int tiles_per_img;  // HW / 128 when HW >= 128, else 0
  int tpi_shift;      // log2(tiles_per_img) when it is a power of two, else -1 (the per-tile divisions sit on the
  int hw_shift;       // single-thread TMA producer's critical path: shifts where possible); same for HW and W
  int w_shift;
  int matrix;         // 1: A is a plain (rows, cols) matrix walked 128 rows at a time along the w axis
  int s2_batch;       // > 0: segment 0 holds the 4 parity planes of a stride-2 conv input, stacked along n
  int w_mn;           // 1: W is stored (K rows, N columns) per tap - a FORWARD-packed weight used for the data gradient:
                      //    B tiles are fetched as 64x64 boxes and consumed as MN-major operands (no transposed copy)
  signed char wtap[9];  // tap column block of W for the i-th A tap (identity unless a tap subset is used)
  int s2_direct;      // 1: segment 0 is the FULL-resolution input read with TMA element strides (2, 2): tap (kh, kw)
                      //    of output tile origin (h0, w0) starts at source pixel (2*h0 + kh, 2*w0 + kw)
  signed char tdh[2][9], tdw[2][9];  // per-segment tap offsets (rows, columns)
  int tdn[9];                        // segment-0 image offset per tap (parity plane of a stride-2 conv)
  int up2_all;        // 1: ONE launch covers the four sub-pixel convolutions of an Upsample layer: the N-tile index
                      //    carries the output parity (row parity * 2 + column parity) in its two low bits; parity
                      //    selects the weight block (rows parity * N .. of the stacked weight matrix), the tap
                      //    offsets (dh, dw) = ((tap >> 1) - 1 + row parity, (tap & 1) - 1 + column parity) and
                      //    the output map
  int bb_row, bb_col; // batched B operand: image i reads W rows + i * bb_row, columns + i * bb_col
  int splits;         // split-K factor (persistent kernel): partial sums go to out_f32 + split * split_stride
  long long split_stride;
  int M, N;
  int flags;
  const float* bias;
  const float* rowbias;
  const int* rowbias_idx;
  int rowbias_ld;
  __nv_bfloat16* vt;
  int vt_col0;
  long long vt_ld;
  float* out_f32;
  long long out_f32_ld;
  float* out_nchw;    // narrow tile (BN = 16): columns [0, nchw_c) written as fp32 NCHW planes (the network's last conv)
  int nchw_c;
  // GroupNorm-fused epilogue (GN template instances): see the kernel's GN branch
  int gn_mode;        // 1: tmC receives GroupNorm(+SiLU)(result); 2: tmC receives the raw result and tmG the normalised one
  int gn_silu;
  int gn_qpg;         // quad-columns (4 channels) per group
  int gn_groups;
  int gn_ipt;         // images per tile: 1 (H*W % 128 == 0) or 2 (8x8 images, 128-wide tiles only)
  int gn_tpi;         // tiles per image (1 when gn_ipt == 2)
  int gn_images;      // images in the batch (gn_ipt == 2 with an odd batch: the last tile's second half is empty)
  float gn_inv_cnt;   // 1 / (HW * channels per group)
  float gn_eps;
  const float* gn_gamma;
  const float* gn_beta;
  GnRecord* gn_part;  // [images][groups][tiles per image][gn_qpg] per-tile {sum, tag}, {sum of squares, tag} of every
                      // quad-column: one 16-byte record
  unsigned* gn_epoch; // [0] launch epoch of this workspace (records of this launch carry epoch + 1), [1] finished CTAs
  CUtensorMap tmG;
#ifdef IDF_GN_TRACE
  long long* trace;   // debug build only (csrc/build.py --variant gntrace -DIDF_GN_TRACE): clock64 stamps of CTA 0
#endif
};

#ifdef IDF_GN_TRACE
#define GN_TRACE_TILES 8
#define GN_TRACE_EVENTS 8
#define GN_STAMP(role, tile, evt)                                                                     \
  do {                                                                                                \
    if (blockIdx.x == 0 && (tile) < GN_TRACE_TILES)                                                   \
      p.trace[((role) * GN_TRACE_TILES + (tile)) * GN_TRACE_EVENTS + (evt)] = clock64();              \
  } while (0)
#else
#define GN_STAMP(role, tile, evt) do { } while (0)
#endif


// silu(t) = t * sigmoid(t) = h + h * tanh(h), h = t / 2 (one MUFU op; same form as the standalone GroupNorm kernel)
__device__ __forceinline__ float igemm_silu(float t) {
  const float h = 0.5f * t;
  float th;
  asm("tanh.approx.f32 %0, %1;" : "=f"(th) : "f"(h));
  return fmaf(h, th, h);
}

// A partial-sum record: two 64-bit words {sum, tag} and {sum of squares, tag}. Each word is one scalar 64-bit access
// (single-copy atomic in the PTX memory model), so a value can never be seen without the tag it was written with, however
// the two words of the 16-byte vector access are ordered.
__device__ __forceinline__ GnRecord gn_record_load(const GnRecord* p) {
  GnRecord v;
  asm volatile("ld.relaxed.gpu.global.v2.b64 {%0, %1}, [%2];" : "=l"(v.lo), "=l"(v.hi) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void gn_record_store(GnRecord* p, float sum, float sumsq, unsigned tag) {
  const unsigned long long lo = ((unsigned long long)tag << 32) | __float_as_uint(sum);
  const unsigned long long hi = ((unsigned long long)tag << 32) | __float_as_uint(sumsq);
  asm volatile("st.relaxed.gpu.global.v2.b64 [%0], {%1, %2};" ::"l"(p), "l"(lo), "l"(hi) : "memory");
}
__device__ __forceinline__ bool gn_record_valid(const GnRecord& v, unsigned tag) {
  return (unsigned)(v.lo >> 32) == tag && (unsigned)(v.hi >> 32) == tag;
}
__device__ __forceinline__ unsigned ld_relaxed_gpu(const unsigned* p) {
  unsigned v;
  asm volatile("ld.relaxed.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}

// ---------------------------------------------------------------------------------------------------------------
// Persistent variant: one CTA per SM walks a static list of output tiles (128 x BN, BN in {128, 192, 256}).
// The fp32 accumulator is double-buffered in tensor memory, so the epilogue of tile i (TMEM -> registers -> bias /
// time bias / residual -> bf16 -> swizzled staging -> TMA store) overlaps the TMA + tcgen05.mma mainloop of tile
// i + 1, and the smem ring keeps streaming across tile boundaries. BN = 256 halves the A re-reads and takes the
// per-MMA shared-memory traffic below the 128 B/clk limit that bounds 128 x 128 tiles.
// ---------------------------------------------------------------------------------------------------------------
constexpr int PG_STG_BYTES = 2 * BLOCK_M * 128;  // one staging buffer: two (128 rows x 64 cols) swizzled boxes

// Work-unit walk shared by the three roles of the persistent kernel: unit u = (tile_m * n_tiles + n_idx) * splits + sp
// advances by gridDim.x per iteration. The stride is decomposed once, so a step costs three compare-and-carry adds
// instead of four integer divisions (the producer is ONE thread: with 2-8 k-blocks per tile its per-tile scalar
// work, not the tensor pipe or the loads in flight, bounded the short-K GEMMs).
struct TileWalk {
  int sp, n_idx, tile_m;
  int dsp, dn, dm, splits, n_tiles;
  __host__ __device__ __forceinline__ void init(int u0, int stride, int splits_, int n_tiles_) {
    splits = splits_; n_tiles = n_tiles_;
    int t = u0, g1 = stride;
    sp = 0; dsp = 0;
    if (splits > 1) { sp = u0 % splits; t = u0 / splits; dsp = stride % splits; g1 = stride / splits; }
    n_idx = t % n_tiles; tile_m = t / n_tiles;
    dn = g1 % n_tiles; dm = g1 / n_tiles;
  }
  __host__ __device__ __forceinline__ void next() {
    int c = 0;
    sp += dsp;
    if (sp >= splits) { sp -= splits; c = 1; }
    n_idx += dn + c;
    if (n_idx >= n_tiles) { n_idx -= n_tiles; tile_m += 1; }
    tile_m += dm;
  }
};

// BN: tile width; NSTG: epilogue staging buffers (1: long-K tiles whose epilogue hides under the next mainloop;
// 3: short-K, epilogue-bound tiles - store of group g-1, fill of group g and residual prefetch of group g+1 overlap)
// PAIR: two CTAs of a cluster share one 256-row MMA (cta_group::2); each stages half of the B tile
// GN: the GroupNorm-fused epilogue keeps per-column tables of the current 128-column group behind the barriers: bias
// (512 B) and (scale, shift) (1 KiB) per image of the tile; GN = 2: two 8x8 images per tile (one ring stage less)
template <int BN, int NSTG, bool PAIR = false, int GN = 0>
struct PgCfg {
  static constexpr int B_BYTES = (PAIR ? BN / 2 : BN) * 128;
  static constexpr int STAGE_BYTES = A_STAGE_BYTES + B_BYTES;
  static constexpr int GN_IPT = GN == 2 ? 2 : 1;  // images per tile
  static constexpr int GN_BYTES = GN ? GN_IPT * 1536 : 0;
  // as many ring stages as fit beside the staging buffers in the 227 KiB of shared memory
  static constexpr int STAGES = (232448 - 1024 - 256 - GN_BYTES - NSTG * PG_STG_BYTES) / STAGE_BYTES;
  static constexpr int TMEM_COLS = (BN == 128) ? 256 : 512;
  static constexpr int SMEM = STAGES * STAGE_BYTES + NSTG * PG_STG_BYTES + 1024 + 256 + GN_BYTES;
};

// EW: epilogue warpgroups (warps 2 .. 2 + 4 EW - 1). Warp w reads TMEM lanes 32 (w % 4) .. +31, so every warpgroup
// covers all 128 rows; the warpgroups split the 32-column chunks of each column group between them. With one warp
// per scheduler (EW = 1) nothing hides the TMEM / L1 / shared-memory latencies of the epilogue, which is what bounds
// the short-K GEMMs and the single-tile-per-CTA launches of the 8x8 / 4x4 stages.
//
// GN != 0: GroupNorm(+SiLU) of the convolution result is applied in the epilogue, so the standalone GroupNorm pass
// (read + write of the whole tensor and a launch) between two convolutions disappears (components.py:448-460: the
// second ConvBlock's GroupNorm reads only what the first one's conv wrote). A tile covers 128 pixels of ONE image
// (HW % 128 == 0) and BN of its channels, the statistics need the whole image: every tile reduces its accumulator
// (+ bias + time bias) to per-quad-column (sum, sum of squares) - fixed shuffle tree over the 32 rows of a warp, the
// four warps added in order - and publishes them to global memory as records whose 64-bit words carry the launch's tag; a
// tile then polls the records of its groups until all of the image's tiles have written theirs (no counters, fences
// or atomics on the path); the tiles of an image are consecutive work units, i.e. they run on neighbouring
// CTAs at the same time, and an image never has two units on one CTA (host-checked), so the wait cannot deadlock:
// by induction over the image index every image's tiles get their accumulators. Each tile then sums the partials of
// its groups in a fixed order (bit-reproducible, independent of the batch size), reads its accumulator from tensor
// memory a second time and stores the normalised tile (mode 2: the raw tile as well, stored between the arrival and the
// wait so that the exchange latency is hidden). The tag is the workspace's launch epoch + 1; the last CTA to finish
// advances the epoch, so stale records of earlier launches (or of other layers sharing the workspace) never match. Per-column constants (bias + time bias, gamma, beta; scale and shift once the statistics are known)
// live in one register of "their" thread and are handed to the row-per-thread loops through small shared tables.
template <int BN, int NSTG, int EW, bool PAIR, int GN = 0>
__global__ void __launch_bounds__(64 + 128 * EW, 1) igemm_persist_kernel(const __grid_constant__ IgemmParams p) {
  using Cfg = PgCfg<BN, NSTG, PAIR, GN>;
  constexpr int EPI_THREADS = 128 * EW;
  constexpr int STAGES = Cfg::STAGES;
  constexpr int GPT = (BN + 127) / 128;  // column groups per tile
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw_addr = smem_u32(smem_raw);
  uint8_t* smem = smem_raw + ((1024u - (raw_addr & 1023u)) & 1023u);
  uint8_t* smem_a = smem;                                // [STAGES][16 KiB]
  uint8_t* smem_b = smem + STAGES * A_STAGE_BYTES;       // [STAGES][BN * 128]
  uint8_t* stage_base = smem + STAGES * Cfg::STAGE_BYTES;  // [NSTG][32 KiB] dedicated epilogue staging
  uint64_t* bars = reinterpret_cast<uint64_t*>(stage_base + NSTG * PG_STG_BYTES);
  uint64_t* full_bar = bars;                    // [STAGES]
  uint64_t* empty_bar = bars + STAGES;          // [STAGES]
  uint64_t* tmem_full = bars + 2 * STAGES;      // [2]
  uint64_t* tmem_empty = tmem_full + 2;         // [2]
  uint64_t* res_bar = tmem_empty + 2;           // [NSTG]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(res_bar + NSTG);

  const int warp = __shfl_sync(0xffffffffu, threadIdx.x >> 5, 0);
  const int lane = threadIdx.x & 31;
  const int n_tiles = (p.N / BN) * (p.up2_all ? 4 : 1);
  // PAIR: a work unit covers the two consecutive M tiles 2 * tile_mp + rank; the cluster (not the CTA) walks the list
  const int rank = PAIR ? (int)cluster_ctarank() : 0;
  const int walker = PAIR ? (int)(blockIdx.x >> 1) : (int)blockIdx.x;
  const int walkers = PAIR ? (int)(gridDim.x >> 1) : (int)gridDim.x;
  const int m_tiles = PAIR ? ((p.M + BLOCK_M - 1) / BLOCK_M + 1) / 2 : (p.M + BLOCK_M - 1) / BLOCK_M;
  // work unit = (output tile, K split); "total_tiles" counts units, the split index varies fastest
  const int splits = p.splits;
  const int total_tiles = m_tiles * n_tiles * splits;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&p.tmA[0]);
    tma_prefetch_desc(&p.tmB);
    if (p.kb_total > p.kb_seg0) tma_prefetch_desc(&p.tmA[1]);
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], 1);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&tmem_full[i], 1);
      mbar_init(&tmem_empty[i], (PAIR ? 2 : 1) * 4 * EW);  // PAIR: the epilogue warps of both CTAs arrive on the even CTA's
    }
    for (int i = 0; i < NSTG; ++i) mbar_init(&res_bar[i], 1);
    fence_mbar_init();
  }
  if (warp == 1) {
    if constexpr (PAIR) { tmem_alloc_pair(tmem_slot, Cfg::TMEM_COLS); tmem_relinquish_pair(); }
    else { tmem_alloc(tmem_slot, Cfg::TMEM_COLS); tmem_relinquish(); }
  }
  tc_fence_before_sync();
  if constexpr (PAIR) cluster_sync_all();  // the peer's barriers are initialised before anything arrives on them
  else __syncthreads();
  tc_fence_after_sync();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    // ------------------------------------------------------------------ TMA producer
    if (elect_one()) {
      TileWalk tw;
      tw.init(walker, walkers, splits, n_tiles);
      int stage = 0;
      uint32_t phase = 0;  // ring position, carried across tiles
      const bool wmn = p.w_mn != 0;
      for (int u = walker; u < total_tiles; u += walkers, tw.next()) {
        const int par = p.up2_all ? (tw.n_idx & 3) : 0;
        const int tile_m = PAIR ? 2 * tw.tile_m + rank : tw.tile_m, n0 = (p.up2_all ? (tw.n_idx >> 2) : tw.n_idx) * BN;
        int brow0 = par * p.N + n0;  // row of the (stacked) weight matrix
        int img0, h0, w0 = 0;
        if (p.matrix) {
          img0 = 0; h0 = 0; w0 = tile_m * BLOCK_M;
        } else if (p.tiles_per_img > 0) {
          if (p.tpi_shift >= 0) {
            img0 = tile_m >> p.tpi_shift;
            h0 = (tile_m & (p.tiles_per_img - 1)) * p.tile_h;
          } else {
            img0 = tile_m / p.tiles_per_img;
            h0 = (tile_m % p.tiles_per_img) * p.tile_h;
          }
        } else {
          img0 = tile_m * p.tile_n; h0 = 0;
        }
        brow0 += img0 * p.bb_row;
        const int bcol0 = img0 * p.bb_col;
        int kb0 = 0, kb1 = p.kb_total;
        int seg = 0, tap = 0, cbk = 0;
        if (splits > 1) {
          kb0 = tw.sp * p.kb_total / splits; kb1 = (tw.sp + 1) * p.kb_total / splits;
          if (kb0 < p.kb_seg0) { tap = kb0 / p.cb[0]; cbk = kb0 % p.cb[0]; }
          else { seg = 1; tap = (kb0 - p.kb_seg0) / p.cb[1]; cbk = (kb0 - p.kb_seg0) % p.cb[1]; }
        }
        // per-tap values live in registers and are refreshed only when the tap changes
        int cb_cur = p.cb[seg], taps_cur = p.taps[seg];
        int ax = 0, ay = 0, an = 0, bcol = 0;
        auto load_tap = [&]() {
          const int sc = (p.s2_direct && seg == 0) ? 2 : 1;
          if (p.up2_all) {
            ax = w0 + (tap & 1) - 1 + (par & 1);
            ay = h0 + (tap >> 1) - 1 + (par >> 1);
          } else {
            ax = sc * w0 + p.tdw[seg][tap];
            ay = sc * h0 + p.tdh[seg][tap];
          }
          an = img0 + (seg == 0 ? p.tdn[tap] : 0);
          if (wmn) bcol = p.wtap[tap] * p.N + n0 + (PAIR ? rank * (BN / 2) : 0);
        };
        load_tap();
        for (int kb = kb0; kb < kb1; ++kb) {
          mbar_wait(&empty_bar[stage], phase ^ 1);
          if constexpr (PAIR) {
            // the even CTA's barrier counts the bytes of both CTAs' loads (its own arrival keeps the phase open)
            if (rank == 0) mbar_expect_tx(&full_bar[stage], 2 * Cfg::STAGE_BYTES);
            tma_load_4d_pair(smem_a + stage * A_STAGE_BYTES, &p.tmA[seg], &full_bar[stage], cbk * BLOCK_K, ax, ay, an);
            if (wmn) {
#pragma unroll
              for (int c = 0; c < BN / 128; ++c)
                tma_load_2d_pair(smem_b + stage * Cfg::B_BYTES + c * 8192, &p.tmB, &full_bar[stage], bcol + c * 64,
                                 cbk * BLOCK_K);
            } else {
              tma_load_2d_pair(smem_b + stage * Cfg::B_BYTES, &p.tmB, &full_bar[stage], bcol0 + kb * BLOCK_K,
                               brow0 + rank * (BN / 2));
            }
          } else {
            mbar_expect_tx(&full_bar[stage], Cfg::STAGE_BYTES);
            tma_load_4d(smem_a + stage * A_STAGE_BYTES, &p.tmA[seg], &full_bar[stage], cbk * BLOCK_K, ax, ay, an);
            if (wmn) {
#pragma unroll
              for (int c = 0; c < BN / 64; ++c)
                tma_load_2d(smem_b + stage * Cfg::B_BYTES + c * 8192, &p.tmB, &full_bar[stage], bcol + c * 64, cbk * BLOCK_K);
            } else {
              tma_load_2d(smem_b + stage * Cfg::B_BYTES, &p.tmB, &full_bar[stage], bcol0 + kb * BLOCK_K, brow0);
            }
          }
          if (++stage == STAGES) { stage = 0; phase ^= 1; }
          if (++cbk == cb_cur) {
            cbk = 0;
            if (++tap == taps_cur) { tap = 0; seg = 1; cb_cur = p.cb[1]; taps_cur = p.taps[1]; }
            if (kb + 1 < kb1) load_tap();
          }
        }
      }
    }
  } else if (warp == 1) {
    // ------------------------------------------------------------------ MMA issuer
    if (rank == 0 && elect_one()) {  // PAIR: the even CTA issues for both
      constexpr uint32_t idesc = umma_idesc_bf16(PAIR ? 2 * BLOCK_M : BLOCK_M, BN);
      int it = 0, stage = 0, sp = walker % splits;
      const int dsp = walkers % splits;
      uint32_t phase = 0;
      for (int u = walker; u < total_tiles; u += walkers, ++it) {
        int kb0 = 0, kb1 = p.kb_total;
        if (splits > 1) {
          kb0 = sp * p.kb_total / splits; kb1 = (sp + 1) * p.kb_total / splits;
          sp += dsp;
          if (sp >= splits) sp -= splits;
        }
        const int acc = it & 1;
        if constexpr (GN != 0) GN_STAMP(0, it, 0);
        mbar_wait(&tmem_empty[acc], ((it >> 1) & 1) ^ 1);  // epilogue has drained this accumulator buffer
        if constexpr (GN != 0) GN_STAMP(0, it, 1);
        tc_fence_after_sync();
        const uint32_t tmem_d = tmem_base + acc * BN;
        for (int kb = kb0; kb < kb1; ++kb) {
          mbar_wait(&full_bar[stage], phase);
          tc_fence_after_sync();
          const uint64_t da = umma_desc_kmajor(smem_u32(smem_a + stage * A_STAGE_BYTES), 128);
          if (p.w_mn) {
            constexpr uint32_t idesc_mn = umma_idesc_bf16(PAIR ? 2 * BLOCK_M : BLOCK_M, BN, 0, 1);
            const uint64_t db = umma_desc_mnmajor(smem_u32(smem_b + stage * Cfg::B_BYTES), 8192, 1024);
#pragma unroll
            for (int k = 0; k < BLOCK_K / 16; ++k) {
              if constexpr (PAIR) umma_bf16_pair(tmem_d, da + 2 * k, db + 128 * k, idesc_mn, (kb > kb0) || (k != 0));
              else umma_bf16(tmem_d, da + 2 * k, db + 128 * k, idesc_mn, (kb > kb0) || (k != 0));
            }
          } else {
            const uint64_t db = umma_desc_kmajor(smem_u32(smem_b + stage * Cfg::B_BYTES), 128);
#pragma unroll
            for (int k = 0; k < BLOCK_K / 16; ++k) {
              if constexpr (PAIR) umma_bf16_pair(tmem_d, da + 2 * k, db + 2 * k, idesc, (kb > kb0) || (k != 0));
              else umma_bf16(tmem_d, da + 2 * k, db + 2 * k, idesc, (kb > kb0) || (k != 0));
            }
          }
          if constexpr (PAIR) umma_commit_pair(&empty_bar[stage]);
          else umma_commit(&empty_bar[stage]);
          if (++stage == STAGES) { stage = 0; phase ^= 1; }
        }
        if constexpr (PAIR) umma_commit_pair(&tmem_full[acc]);
        else umma_commit(&tmem_full[acc]);
        if constexpr (GN != 0) GN_STAMP(0, it, 2);
      }
    }
  } else {
    // ------------------------------------------------------------------ epilogue (warps 2 .. 2 + 4 EW - 1)
    // Work unit = "column group": up to 128 output columns of a tile = one staging buffer. Groups are numbered gc
    // across tiles; group gc owns staging buffer gc % NSTG.
    const int quad = warp & 3;
    const int wg = (warp - 2) >> 2;  // which share of each column group's chunks this warp takes
    const int r = quad * 32 + lane;
    const int et = threadIdx.x - 64;
    const uint32_t lane_addr = (uint32_t)(quad * 32) << 16;
    const bool has_res = (p.flags & F_RES) != 0;
    const bool to_f32 = (p.flags & F_OUT_F32) != 0;
    const bool issuer = (warp == 2) && elect_one();
    auto group_cols = [&](int cg) { return (BN - cg * 128) < 128 ? (BN - cg * 128) : 128; };
    // residual boxes of group (tile (tm, ni), cg) -> staging buffer gc % NSTG
    auto load_res = [&](int tm, int ni, int cg, int gc) {
      const int bufi = gc % NSTG;
      const int gcols = group_cols(cg);
      const int nc0 = ni * BN + cg * 128;
      uint8_t* dst = stage_base + bufi * PG_STG_BYTES;
      mbar_expect_tx(&res_bar[bufi], (gcols / 64) * BLOCK_M * 128);
      for (int bx = 0; bx < gcols / 64; ++bx)
        tma_load_2d(dst + bx * (BLOCK_M * 128), &p.tmR, &res_bar[bufi], nc0 + bx * 64, tm * BLOCK_M);
    };
    int it = 0, gc = 0;
    unsigned tag = 0;
    if constexpr (GN != 0) {  // records of this launch carry epoch + 1 (never 0: fresh workspaces are zero-filled)
      tag = ld_relaxed_gpu(p.gn_epoch) + 1u;
      if (tag == 0u) tag = 1u;
    }
    if (has_res && issuer && walker < total_tiles && !((p.flags & F_VT) || to_f32))
      load_res((PAIR ? 2 : 1) * (walker / n_tiles) + rank, walker % n_tiles, 0, 0);  // (residual path: splits == 1)
    TileWalk tw;
    tw.init(walker, walkers, splits, n_tiles);
    for (int u = walker; u < total_tiles; u += walkers, ++it, tw.next()) {
      if constexpr (BN == 16) {
        // Narrow tile: the <= 16 output channels of the network's last convolution (128 -> 3, 384 -> z), one row per
        // thread, written straight as fp32 NCHW planes (consecutive lanes = consecutive pixels: coalesced per plane).
        const int acc = it & 1;
        const int tile_m = tw.tile_m;
        const long long m = (long long)tile_m * BLOCK_M + r;
        mbar_wait(&tmem_full[acc], (it >> 1) & 1);
        tc_fence_after_sync();
        uint32_t v[16];
        if (wg == 0) {
          tmem_ld_32x16(tmem_base + acc * BN + lane_addr, v);
          tmem_ld_wait();
        }
        tc_fence_before_sync();
        __syncwarp();
        if (lane == 0) mbar_arrive(&tmem_empty[acc]);
        if (wg == 0 && m < p.M) {
          int sample, pix;
          if (p.hw_shift >= 0) { sample = (int)(m >> p.hw_shift); pix = (int)m & (p.HW - 1); }
          else { sample = (int)(m / p.HW); pix = (int)(m % p.HW); }
          float* dst = p.out_nchw + ((long long)sample * p.nchw_c) * p.HW + pix;
#pragma unroll
          for (int c = 0; c < 16; ++c)
            if (c < p.nchw_c) dst[(long long)c * p.HW] = __uint_as_float(v[c]) + (p.bias ? __ldg(p.bias + c) : 0.f);
        }
        continue;
      }
      if constexpr (GN != 0) {
        constexpr int QC = BN / 4;
        // tables of the current column group, per image of the tile: [ipt][128] bias + time bias, [ipt][128] (scale, shift)
        float* cbt = reinterpret_cast<float*>(stage_base + NSTG * PG_STG_BYTES + 256);
        float2* sct = reinterpret_cast<float2*>(cbt + Cfg::GN_IPT * 128);
        float* red = reinterpret_cast<float*>(stage_base);  // [4 warps][BN / 4][2]: staging buffer 0 during pass 1
        const int tile_m = PAIR ? 2 * tw.tile_m + rank : tw.tile_m, n0 = tw.n_idx * BN;
        const int acc = it & 1;
        // ipt = 2: the tile holds two 8x8 images (rows 0..63 / 64..127 = TMEM quadrants 0,1 / 2,3); statistics stay
        // per image, the exchange is between the N tiles of the same 128 rows
        constexpr int ipt = Cfg::GN_IPT;
        const int tpi = p.gn_tpi, qpg = p.gn_qpg;
        const int img = ipt == 2 ? 2 * tile_m : (p.tpi_shift >= 0 ? (tile_m >> p.tpi_shift) : (tile_m / tpi));  // first image
        // column role: thread et <-> (image tsel of the tile, column n0 + colr); row role: this thread's row is in image rsel
        const int colr = ipt == 2 ? (et & 127) : et;
        const int tsel = ipt == 2 ? (et >> 7) : 0;
        const int rsel = ipt == 2 ? (quad >> 1) : 0;
        const bool col_thread = et < BN * ipt && img + tsel < p.gn_images;
        // bias + time bias, gamma, beta of this thread's column - fetched while the MMAs run
        float cb_r = 0.f, ga_r = 0.f, be_r = 0.f;
        if (col_thread) {
          const int col = n0 + colr;
          if (p.bias != nullptr) cb_r = __ldg(p.bias + col);
          if (p.rowbias != nullptr)
            cb_r += __ldg(p.rowbias + (long long)(p.rowbias_idx ? p.rowbias_idx[img + tsel] : img + tsel) * p.rowbias_ld + col);
          ga_r = __ldg(p.gn_gamma + col);
          be_r = __ldg(p.gn_beta + col);
        }
        if (issuer) tma_store_wait_read_all();  // the previous tile's stores have finished reading the staging buffer
        if (issuer) GN_STAMP(1, it, 0);
        mbar_wait(&tmem_full[acc], (it >> 1) & 1);
        tc_fence_after_sync();
        if (issuer) GN_STAMP(1, it, 1);
        const uint32_t tmem_d = tmem_base + acc * BN + lane_addr;
        // ---- pass 1: per-quad-column (sum, sum of squares) over this tile's 128 rows
#pragma unroll 1
        for (int cg = 0; cg < GPT; ++cg) {
          const int gcols = group_cols(cg);
          const int my_nch = (gcols / 32) / EW;
          named_bar_sync(1, EPI_THREADS);  // staging buffer (red) and cbt free
          if (col_thread && colr >= cg * 128 && colr < cg * 128 + gcols) cbt[tsel * 128 + colr - cg * 128] = cb_r;
          named_bar_sync(1, EPI_THREADS);
#pragma unroll 1
          for (int i0 = 0; i0 < my_nch; i0 += 2) {
            const bool two = i0 + 1 < my_nch;
            const int c0 = wg * my_nch + i0;
            uint32_t v[2][32];
            tmem_ld_32x32(tmem_d + cg * 128 + c0 * 32, v[0]);
            if (two) tmem_ld_32x32(tmem_d + cg * 128 + (c0 + 1) * 32, v[1]);
            tmem_ld_wait();
#pragma unroll
            for (int h = 0; h < 2; ++h) {
              if (h == 1 && !two) break;
              const int c = c0 + h;
              float w16[16];
#pragma unroll
              for (int q = 0; q < 8; ++q) {
                const float4 cb4 = *reinterpret_cast<const float4*>(cbt + rsel * 128 + c * 32 + 4 * q);
                const float a0 = __uint_as_float(v[h][4 * q + 0]) + cb4.x, a1 = __uint_as_float(v[h][4 * q + 1]) + cb4.y;
                const float a2 = __uint_as_float(v[h][4 * q + 2]) + cb4.z, a3 = __uint_as_float(v[h][4 * q + 3]) + cb4.w;
                w16[q] = (a0 + a1) + (a2 + a3);
                w16[8 + q] = fmaf(a3, a3, fmaf(a2, a2, fmaf(a1, a1, a0 * a0)));
              }
              // halving butterfly over the 32 rows of the warp: 16 values -> lane l ends with the total of value l >> 1
#pragma unroll
              for (int k = 0; k < 8; ++k) {
                const bool up = (lane & 16) != 0;
                const float send = up ? w16[k] : w16[k + 8], keep = up ? w16[k + 8] : w16[k];
                w16[k] = keep + __shfl_xor_sync(0xffffffffu, send, 16);
              }
#pragma unroll
              for (int k = 0; k < 4; ++k) {
                const bool up = (lane & 8) != 0;
                const float send = up ? w16[k] : w16[k + 4], keep = up ? w16[k + 4] : w16[k];
                w16[k] = keep + __shfl_xor_sync(0xffffffffu, send, 8);
              }
#pragma unroll
              for (int k = 0; k < 2; ++k) {
                const bool up = (lane & 4) != 0;
                const float send = up ? w16[k] : w16[k + 2], keep = up ? w16[k + 2] : w16[k];
                w16[k] = keep + __shfl_xor_sync(0xffffffffu, send, 4);
              }
              {
                const bool up = (lane & 2) != 0;
                const float send = up ? w16[0] : w16[1], keep = up ? w16[1] : w16[0];
                w16[0] = keep + __shfl_xor_sync(0xffffffffu, send, 2);
              }
              w16[0] += __shfl_xor_sync(0xffffffffu, w16[0], 1);
              if ((lane & 1) == 0) {
                const int idx = lane >> 1;  // 0..7: sum of quad idx, 8..15: sum of squares of quad idx - 8
                red[((quad * QC) + ((cg * 128 + c * 32) >> 2) + (idx & 7)) * 2 + (idx >> 3)] = w16[0];
              }
            }
          }
        }
        named_bar_sync(1, EPI_THREADS);
        if (issuer) GN_STAMP(1, it, 2);
        if (et < QC * ipt && img + (ipt == 2 ? et / QC : 0) < p.gn_images) {
          // records laid out [image][group][tile of the image][quad-column of the group]
          const int isel = ipt == 2 ? et / QC : 0, ql = et - isel * QC;
          float s_ = 0.f, q_ = 0.f;
#pragma unroll
          for (int w4 = 0; w4 < 4; ++w4)  // the warps holding this image's rows, in order
            if (ipt == 1 || (w4 >> 1) == isel) { s_ += red[(w4 * QC + ql) * 2]; q_ += red[(w4 * QC + ql) * 2 + 1]; }
          const int qc = (n0 >> 2) + ql;
          const int g = qc / qpg;
          const int tin = ipt == 2 ? 0 : tile_m - img * tpi;
          gn_record_store(p.gn_part + (((long long)(img + isel) * p.gn_groups + g) * tpi + tin) * qpg + (qc - g * qpg), s_, q_, tag);
        }
        if (issuer) GN_STAMP(1, it, 3);
        // one pass over the accumulator through the staging buffer: raw (+ bias) tile or normalised tile
        float sc_r = 0.f, sh_r = 0.f;
        auto store_pass = [&](const bool norm, const CUtensorMap* tc) {
#pragma unroll 1
          for (int cg = 0; cg < GPT; ++cg) {
            const int gcols = group_cols(cg);
            const int my_nch = (gcols / 32) / EW;
            if (issuer) tma_store_wait_read_all();
            named_bar_sync(1, EPI_THREADS);  // staging buffer and tables free
            if (col_thread && colr >= cg * 128 && colr < cg * 128 + gcols) {
              if (norm) sct[tsel * 128 + colr - cg * 128] = make_float2(sc_r, sh_r);
              else cbt[tsel * 128 + colr - cg * 128] = cb_r;
            }
            named_bar_sync(1, EPI_THREADS);
#pragma unroll 1
            for (int i0 = 0; i0 < my_nch; i0 += 2) {
              const bool two = i0 + 1 < my_nch;
              const int c0 = wg * my_nch + i0;
              uint32_t v[2][32];
              tmem_ld_32x32(tmem_d + cg * 128 + c0 * 32, v[0]);
              if (two) tmem_ld_32x32(tmem_d + cg * 128 + (c0 + 1) * 32, v[1]);
              tmem_ld_wait();
              if (norm && cg == GPT - 1 && i0 + 2 >= my_nch) {  // this warp's last TMEM read of the tile
                tc_fence_before_sync();
                __syncwarp();
                if (lane == 0) {
                  if constexpr (PAIR) mbar_arrive_even_cta(&tmem_empty[acc]);
                  else mbar_arrive(&tmem_empty[acc]);
                }
              }
#pragma unroll
              for (int h = 0; h < 2; ++h) {
                if (h == 1 && !two) break;
                const int c = c0 + h;
                float a[32];
                if (norm) {
#pragma unroll
                  for (int j = 0; j < 32; j += 2) {
                    const float4 t4 = *reinterpret_cast<const float4*>(sct + rsel * 128 + c * 32 + j);  // (scale, shift) of two columns
                    a[j] = fmaf(__uint_as_float(v[h][j]), t4.x, t4.y);
                    a[j + 1] = fmaf(__uint_as_float(v[h][j + 1]), t4.z, t4.w);
                  }
                  if (p.gn_silu) {
#pragma unroll
                    for (int j = 0; j < 32; ++j) a[j] = igemm_silu(a[j]);
                  }
                } else {
#pragma unroll
                  for (int q = 0; q < 8; ++q) {
                    const float4 cb4 = *reinterpret_cast<const float4*>(cbt + rsel * 128 + c * 32 + 4 * q);
                    a[4 * q + 0] = __uint_as_float(v[h][4 * q + 0]) + cb4.x; a[4 * q + 1] = __uint_as_float(v[h][4 * q + 1]) + cb4.y;
                    a[4 * q + 2] = __uint_as_float(v[h][4 * q + 2]) + cb4.z; a[4 * q + 3] = __uint_as_float(v[h][4 * q + 3]) + cb4.w;
                  }
                }
                uint8_t* box = stage_base + (c >> 1) * (BLOCK_M * 128) + r * 128;
#pragma unroll
                for (int q = 0; q < 4; ++q) {
                  const int chunk = ((c & 1) * 4 + q) ^ (r & 7);
                  uint4 o;
                  o.x = pack_bf16x2(a[8 * q + 0], a[8 * q + 1]);
                  o.y = pack_bf16x2(a[8 * q + 2], a[8 * q + 3]);
                  o.z = pack_bf16x2(a[8 * q + 4], a[8 * q + 5]);
                  o.w = pack_bf16x2(a[8 * q + 6], a[8 * q + 7]);
                  *reinterpret_cast<uint4*>(box + chunk * 16) = o;
                }
              }
            }
            fence_proxy_async_smem();
            named_bar_sync(1, EPI_THREADS);
            if (issuer) {
              for (int bx = 0; bx < gcols / 64; ++bx)
                tma_store_2d(tc, stage_base + bx * (BLOCK_M * 128), n0 + cg * 128 + bx * 64, tile_m * BLOCK_M);
              tma_store_commit();
            }
          }
        };
        if (p.gn_mode == 2) store_pass(false, &p.tmC);  // the raw tile goes out while the other tiles' records arrive
        if (issuer) GN_STAMP(1, it, 4);
        // statistics of this thread's column's group: the records of all of the image's tiles, summed in a fixed order.
        // A record is valid once both of its 64-bit words carry this launch's tag (value and tag share a word, so no
        // separate flag, fence or counter is needed). One thread waits for the records of its own group first - the
        // other tiles write all of theirs at the same moment - so that a waiting CTA polls with one thread, not 256
        // (measured: all threads polling slowed the TMA loads of every CTA); then each thread loads and checks its own.
        const int n = tpi * qpg;
        float s_ = 0.f, q_ = 0.f;
        // the group's n records, 16 per batch (one round trip): sums them in order once every tag matches
        auto sum_group = [&](const GnRecord* src, const unsigned backoff_ns) {
          s_ = 0.f; q_ = 0.f;
          const GnRecord none = {(unsigned long long)tag << 32, (unsigned long long)tag << 32};  // (0, 0), valid
          for (int i0 = 0; i0 < n; i0 += 16) {
            GnRecord v4[16];
            bool ok;
            unsigned ns = backoff_ns, spins = 0;
            do {
              ok = true;
#pragma unroll
              for (int k = 0; k < 16; ++k) v4[k] = (i0 + k < n) ? gn_record_load(src + i0 + k) : none;
#pragma unroll
              for (int k = 0; k < 16; ++k) ok = ok && gn_record_valid(v4[k], tag);
              if (!ok) {
                __nanosleep(ns);
                if (ns < 256) ns *= 2;
                // ~4 s without the other tiles' records: their CTAs are not running (the launch shares the GPU with
                // another resident kernel that waits in the same way) - fail loudly instead of hanging
                if (++spins > (1u << 24)) __trap();
              }
            } while (!ok);
#pragma unroll
            for (int k = 0; k < 16; ++k) { s_ += __uint_as_float((unsigned)v4[k].lo); q_ += __uint_as_float((unsigned)v4[k].hi); }
          }
        };
        const GnRecord* gsrc = p.gn_part + ((long long)(img + tsel) * p.gn_groups + (n0 + colr) / (4 * qpg)) * n;
        if (issuer) sum_group(gsrc, 32);
        named_bar_sync(1, EPI_THREADS);
        if (col_thread) {
          sum_group(gsrc, 256);
          const float mean = s_ * p.gn_inv_cnt;
          const float var = fmaxf(q_ * p.gn_inv_cnt - mean * mean, 0.f);
          sc_r = rsqrtf(var + p.gn_eps) * ga_r;
          sh_r = fmaf(cb_r, sc_r, be_r - mean * sc_r);  // (acc + cb - mean) * rstd * gamma + beta = acc * sc + sh
        }
        if (issuer) GN_STAMP(1, it, 6);
        store_pass(true, p.gn_mode == 2 ? &p.tmG : &p.tmC);
        if (issuer) GN_STAMP(1, it, 7);
        continue;
      }
      const int t = u;  // (the residual prefetch below is only used with splits == 1, where the unit is the tile)
      float* out_f32 = p.out_f32 + (long long)tw.sp * p.split_stride;
      const int par = p.up2_all ? (tw.n_idx & 3) : 0;
      const int tile_m = PAIR ? 2 * tw.tile_m + rank : tw.tile_m, n0 = (p.up2_all ? (tw.n_idx >> 2) : tw.n_idx) * BN;
      const int acc = it & 1;
      const long long m = (long long)tile_m * BLOCK_M + r;
      const bool row_ok = m < p.M;
      const bool to_vt = (p.flags & F_VT) && n0 >= p.vt_col0;
      const bool staged = !to_vt && !to_f32;

      int sample = 0;
      bool zero_row = false;
      if (p.rowbias != nullptr || (p.flags & F_ZERO_PAD)) {
        const int mm = row_ok ? (int)m : 0;
        int pix;
        if (p.hw_shift >= 0) { sample = mm >> p.hw_shift; pix = mm & (p.HW - 1); }
        else { sample = mm / p.HW; pix = mm % p.HW; }
        if (p.flags & F_ZERO_PAD) {
          int ph, pw;
          if (p.w_shift >= 0) { ph = pix >> p.w_shift; pw = pix & (p.W - 1); }
          else { ph = pix / p.W; pw = pix % p.W; }
          zero_row = (ph == p.H - 1) || (pw == p.W - 1);
        }
      }
      const float* rb = nullptr;
      if (p.rowbias != nullptr) {
        const int rrow = p.rowbias_idx ? p.rowbias_idx[sample] : sample;
        rb = p.rowbias + (long long)rrow * p.rowbias_ld + n0;
      }

      mbar_wait(&tmem_full[acc], (it >> 1) & 1);
      tc_fence_after_sync();
      const uint32_t tmem_d = tmem_base + acc * BN + lane_addr;

#pragma unroll 1
      for (int cg = 0; cg < GPT; ++cg, ++gc) {
        const int gcols = group_cols(cg);
        const int nc0 = n0 + cg * 128;
        uint8_t* stage_c = stage_base + (gc % NSTG) * PG_STG_BYTES;
        // Staging buffer reuse: the TMA store issued NSTG groups ago must have finished READING this buffer; with a
        // residual prefetch one group ahead, the buffer of group gc+1 (stored NSTG-1 groups ago) must be free too.
        if (issuer) {
          if (NSTG >= 3 && has_res) asm volatile("cp.async.bulk.wait_group.read %0;" ::"n"(NSTG >= 3 ? NSTG - 2 : 0) : "memory");
          else asm volatile("cp.async.bulk.wait_group.read %0;" ::"n"(NSTG - 1) : "memory");
        }
        named_bar_sync(1, EPI_THREADS);
        if (has_res && staged) {
          if (NSTG >= 3) {
            if (issuer) {  // prefetch the next group's residual
              int nt = t, ncg = cg + 1;
              TileWalk nx = tw;
              if (ncg == GPT) { nt = t + walkers; ncg = 0; nx.next(); }  // splits == 1 here
              if (nt < total_tiles) load_res(PAIR ? 2 * nx.tile_m + rank : nx.tile_m, nx.n_idx, ncg, gc + 1);
            }
          } else if (gc > 0) {
            if (issuer) load_res(tile_m, tw.n_idx, cg, gc);
          }
          mbar_wait(&res_bar[gc % NSTG], (gc / NSTG) & 1);
        }
        // one 32-column chunk: + bias terms, + residual, mask, convert, store (staging / transposed staging / fp32)
        auto finish_chunk = [&](const int c, const uint32_t (&v)[32], const float (&bb)[32]) {
          float a[32];
#pragma unroll
          for (int j = 0; j < 32; ++j) a[j] = __uint_as_float(v[j]) + bb[j];
          if (staged) {
            uint8_t* box = stage_c + (c >> 1) * (BLOCK_M * 128) + r * 128;
#pragma unroll
            for (int q = 0; q < 4; ++q) {
              const int chunk = ((c & 1) * 4 + q) ^ (r & 7);
              uint4* slot = reinterpret_cast<uint4*>(box + chunk * 16);
              if (has_res) {
                const uint4 rv = *slot;
                a[8 * q + 0] += bf16_lo(rv.x); a[8 * q + 1] += bf16_hi(rv.x);
                a[8 * q + 2] += bf16_lo(rv.y); a[8 * q + 3] += bf16_hi(rv.y);
                a[8 * q + 4] += bf16_lo(rv.z); a[8 * q + 5] += bf16_hi(rv.z);
                a[8 * q + 6] += bf16_lo(rv.w); a[8 * q + 7] += bf16_hi(rv.w);
              }
              uint4 o;
              if (zero_row) {
                o = make_uint4(0u, 0u, 0u, 0u);
              } else {
                o.x = pack_bf16x2(a[8 * q + 0], a[8 * q + 1]);
                o.y = pack_bf16x2(a[8 * q + 2], a[8 * q + 3]);
                o.z = pack_bf16x2(a[8 * q + 4], a[8 * q + 5]);
                o.w = pack_bf16x2(a[8 * q + 6], a[8 * q + 7]);
              }
              *slot = o;
            }
          } else if (to_vt) {
            __nv_bfloat16* tcol = reinterpret_cast<__nv_bfloat16*>(stage_c) + (c * 32) * BLOCK_M + r;
#pragma unroll
            for (int j = 0; j < 32; ++j) tcol[j * BLOCK_M] = __float2bfloat16_rn(a[j]);
          } else if (row_ok) {
            float4* dst = reinterpret_cast<float4*>(out_f32 + m * p.out_f32_ld + nc0 + c * 32);
#pragma unroll
            for (int j = 0; j < 8; ++j) dst[j] = make_float4(a[4 * j + 0], a[4 * j + 1], a[4 * j + 2], a[4 * j + 3]);
          }
        };
        // bias + per-sample time bias of one chunk, fetched while the TMEM loads are in flight
        auto load_bias = [&](const int c, float (&bb)[32]) {
#pragma unroll
          for (int j = 0; j < 32; ++j) bb[j] = 0.f;
          if (p.bias != nullptr) {
            const float4* b4 = reinterpret_cast<const float4*>(p.bias + nc0 + c * 32);
#pragma unroll
            for (int j = 0; j < 8; ++j) {
              const float4 b = __ldg(b4 + j);
              bb[4 * j + 0] = b.x; bb[4 * j + 1] = b.y; bb[4 * j + 2] = b.z; bb[4 * j + 3] = b.w;
            }
          }
          if (rb != nullptr) {
            const float4* b4 = reinterpret_cast<const float4*>(rb + cg * 128 + c * 32);
#pragma unroll
            for (int j = 0; j < 8; ++j) {
              const float4 b = __ldg(b4 + j);
              bb[4 * j + 0] += b.x; bb[4 * j + 1] += b.y; bb[4 * j + 2] += b.z; bb[4 * j + 3] += b.w;
            }
          }
        };
        // this warpgroup's chunks of the group, two at a time: both TMEM loads and the bias loads are issued before
        // the single wait
        const int my_nch = (gcols / 32) / EW;
        const int cbase = wg * my_nch;
#pragma unroll 1
        for (int i0 = 0; i0 < my_nch; i0 += 2) {
          const int c0 = cbase + i0;
          const bool two = i0 + 1 < my_nch;
          uint32_t v0[32], v1[32];
          float bb0[32], bb1[32];
          tmem_ld_32x32(tmem_d + cg * 128 + c0 * 32, v0);
          if (two) tmem_ld_32x32(tmem_d + cg * 128 + (c0 + 1) * 32, v1);
          load_bias(c0, bb0);
          if (two) load_bias(c0 + 1, bb1);
          tmem_ld_wait();
          if (cg == GPT - 1 && i0 + 2 >= my_nch) {  // this warp's last TMEM read of the tile: release the accumulator
            tc_fence_before_sync();
            __syncwarp();
            if (lane == 0) {
              if constexpr (PAIR) mbar_arrive_even_cta(&tmem_empty[acc]);
              else mbar_arrive(&tmem_empty[acc]);
            }
          }
          finish_chunk(c0, v0, bb0);
          if (two) finish_chunk(c0 + 1, v1, bb1);
        }
        if (to_vt) {
          named_bar_sync(1, EPI_THREADS);
          const long long m0 = (long long)tile_m * BLOCK_M;
          __nv_bfloat16* gbase = p.vt + (long long)(nc0 - p.vt_col0) * p.vt_ld + m0;
          for (int q = et; q < gcols * 16; q += EPI_THREADS) {
            const int col = q >> 4, part = q & 15;
            if (m0 + part * 8 < p.M) {
              const uint4 vv = *reinterpret_cast<const uint4*>(stage_c + col * (BLOCK_M * 2) + part * 16);
              *reinterpret_cast<uint4*>(gbase + (long long)col * p.vt_ld + part * 8) = vv;
            }
          }
          if (issuer) tma_store_commit();  // empty bulk group: keeps "groups pending" == "staging buffers in use"
        } else if (staged) {
          fence_proxy_async_smem();
          named_bar_sync(1, EPI_THREADS);
          if (issuer) {
            if (p.flags & F_OUT_UP2) {
              // tile rows = low-resolution pixels (img, h, w); tmC is a 4-D map over the 2x-resolution output whose
              // (w, h) strides are doubled and whose base sits at this launch's (row, column) parity
              int img0, h0;
              if (p.tiles_per_img > 0) { img0 = tile_m / p.tiles_per_img; h0 = (tile_m % p.tiles_per_img) * p.tile_h; }
              else { img0 = tile_m * p.tile_n; h0 = 0; }
              const CUtensorMap* tc = par == 0 ? &p.tmC : &p.tmCx[par - 1];
              for (int bx = 0; bx < gcols / 64; ++bx)
                tma_store_4d(tc, stage_c + bx * (BLOCK_M * 128), nc0 + bx * 64, 0, h0, img0);
            } else {
              for (int bx = 0; bx < gcols / 64; ++bx)
                tma_store_2d(&p.tmC, stage_c + bx * (BLOCK_M * 128), nc0 + bx * 64, tile_m * BLOCK_M);
            }
            tma_store_commit();
          }
        } else if (issuer) {
          tma_store_commit();
        }
      }
    }
    if (issuer) tma_store_wait_read_all();
    if constexpr (GN != 0) {
      // the last CTA to finish advances the workspace's epoch: every CTA has read it by then, and the next launch on
      // this workspace (stream order) sees the new value
      if (issuer && atomicAdd(p.gn_epoch + 1, 1u) == gridDim.x - 1u) {
        p.gn_epoch[1] = 0u;
        p.gn_epoch[0] = tag;
      }
    }
  }

  tc_fence_before_sync();
  if constexpr (PAIR) cluster_sync_all();  // neither CTA may leave while the pair's MMAs / remote arrivals can touch it
  else __syncthreads();
  if (warp == 1) {
    __syncwarp();
    tc_fence_after_sync();
    if constexpr (PAIR) tmem_dealloc_pair(tmem_base, Cfg::TMEM_COLS);
    else tmem_dealloc(tmem_base, Cfg::TMEM_COLS);
  }
}

// GN-fused launches: a walker count that is a whole number of images per wave, so that the tiles of an image always
// run in the same wave (an image split over two waves makes its first tiles wait a whole tile time for the others, and
// the delay then spreads through the images those CTAs share later: measured +17 us on a 48 us conv)
static int gn_walkers(int walkers, int per_img) {
  return per_img > 0 && walkers >= per_img ? walkers / per_img * per_img : walkers;
}

template <int BN, int NSTG, int EW, bool PAIR, int GN = 0>
static int launch_persist_ew(const IgemmParams& p, cudaStream_t stream) {
  using Cfg = PgCfg<BN, NSTG, PAIR, GN>;
  constexpr int SMEM = Cfg::SMEM;
  static_assert(SMEM <= 232448, "shared memory");
  static bool attr_set = false;
  if (!attr_set) {
    int rc = check_cuda(cudaFuncSetAttribute(igemm_persist_kernel<BN, NSTG, EW, PAIR, GN>,
                                             cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM),
                        "igemm_persist: cudaFuncSetAttribute");
    if (rc != IDF_OK) return rc;
    attr_set = true;
  }
  const int m_tiles = (p.M + BLOCK_M - 1) / BLOCK_M;
  if constexpr (PAIR) {
    // one cluster of two CTAs per TPC; the cluster walks the list of (M-tile pair, N tile, K split) units
    const int units = ((m_tiles + 1) / 2) * (p.N / BN) * p.splits * (p.up2_all ? 4 : 1);
    int clusters = units < sm_count() / 2 ? units : sm_count() / 2;
    if constexpr (GN != 0) clusters = gn_walkers(clusters, (p.gn_tpi / 2) * (p.N / BN));
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(2 * clusters);
    cfg.blockDim = dim3(64 + 128 * EW);
    cfg.dynamicSmemBytes = SMEM;
    cfg.stream = stream;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = 2; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    return check_cuda(cudaLaunchKernelEx(&cfg, igemm_persist_kernel<BN, NSTG, EW, PAIR, GN>, p), "igemm_persist pair launch");
  } else {
    const int tiles = m_tiles * (p.N / BN) * p.splits * (p.up2_all ? 4 : 1);
    int grid = tiles < sm_count() ? tiles : sm_count();
    if constexpr (GN != 0) grid = gn_walkers(grid, p.gn_tpi * (p.N / BN));
    return check_cuda(launch_kernel(igemm_persist_kernel<BN, NSTG, EW, PAIR, GN>, dim3(grid), dim3(64 + 128 * EW), SMEM,
                                 stream, p),
                      "igemm_persist launch");
  }
}

template <int BN, int NSTG>
static int launch_persist(const IgemmParams& p, cudaStream_t stream, bool pair) {
  // IDF_EPI_WG=1 selects the single-warpgroup epilogue (kept for A/B measurements; single-CTA kernels only)
  static const int ew = [] { const char* e = getenv("IDF_EPI_WG"); return e ? atoi(e) : 2; }();
  if (pair) return launch_persist_ew<BN, NSTG, 2, true>(p, stream);
  return ew == 1 ? launch_persist_ew<BN, NSTG, 1, false>(p, stream) : launch_persist_ew<BN, NSTG, 2, false>(p, stream);
}

// Split-K finish: out[m, n] = bf16( sum_s partial[s][m][n] + bias[n] + rowbias[row(sample(m))][n] ), partials summed in
// split order (deterministic), padded rows written as exact zeros.
__global__ void __launch_bounds__(256) splitk_finish_kernel(const float* __restrict__ ws, long long split_stride,
                                                            int splits, __nv_bfloat16* __restrict__ out, long long ldo,
                                                            int M, int N, const float* __restrict__ bias,
                                                            const float* __restrict__ rowbias,
                                                            const int* __restrict__ rowbias_idx, int rowbias_ld,
                                                            int H, int W, int zero_pad) {
  const int vec = N / 8;
  const long long total = (long long)M * vec;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const int v = (int)(i % vec);
    const long long m = i / vec;
    const int n = v * 8;
    float acc[8];
    {
      const float4 a = *reinterpret_cast<const float4*>(ws + m * N + n);
      const float4 b = *reinterpret_cast<const float4*>(ws + m * N + n + 4);
      acc[0] = a.x; acc[1] = a.y; acc[2] = a.z; acc[3] = a.w; acc[4] = b.x; acc[5] = b.y; acc[6] = b.z; acc[7] = b.w;
    }
    for (int s = 1; s < splits; ++s) {
      const float4 a = *reinterpret_cast<const float4*>(ws + s * split_stride + m * N + n);
      const float4 b = *reinterpret_cast<const float4*>(ws + s * split_stride + m * N + n + 4);
      acc[0] += a.x; acc[1] += a.y; acc[2] += a.z; acc[3] += a.w; acc[4] += b.x; acc[5] += b.y; acc[6] += b.z; acc[7] += b.w;
    }
    if (bias != nullptr) {
#pragma unroll
      for (int e = 0; e < 8; ++e) acc[e] += bias[n + e];
    }
    const int HW = H * W;
    if (rowbias != nullptr) {
      const int sample = (int)(m / HW);
      const float* rb = rowbias + (long long)(rowbias_idx ? rowbias_idx[sample] : sample) * rowbias_ld + n;
#pragma unroll
      for (int e = 0; e < 8; ++e) acc[e] += rb[e];
    }
    uint4 o;
    const int pix = (int)(m % HW);
    if (zero_pad && ((pix / W == H - 1) || (pix % W == W - 1))) {
      o = make_uint4(0u, 0u, 0u, 0u);
    } else {
      o.x = pack_bf16x2(acc[0], acc[1]);
      o.y = pack_bf16x2(acc[2], acc[3]);
      o.z = pack_bf16x2(acc[4], acc[5]);
      o.w = pack_bf16x2(acc[6], acc[7]);
    }
    *reinterpret_cast<uint4*>(out + m * ldo + n) = o;
  }
}

static int make_act_map(CUtensorMap* tm, const idf_nhwc_t& a, int tile_w, int tile_h, int tile_n, int stride = 1) {
  const uint64_t dims[4] = {(uint64_t)a.c, (uint64_t)a.w, (uint64_t)a.h, (uint64_t)a.n};
  const uint64_t strides[3] = {(uint64_t)a.sw * 2, (uint64_t)a.sh * 2, (uint64_t)a.sn * 2};
  // with element strides the box is the SOURCE span: tile_w * stride pixels of which every stride-th is fetched
  const uint32_t box[4] = {(uint32_t)BLOCK_K, (uint32_t)(tile_w * stride), (uint32_t)(tile_h * stride), (uint32_t)tile_n};
  const uint32_t estr[4] = {1u, (uint32_t)stride, (uint32_t)stride, 1u};
  return encode_tmap(tm, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, a.ptr, 4, dims, strides, box,
                     CU_TENSOR_MAP_SWIZZLE_128B, estr);
}

static int make_mat_map(CUtensorMap* tm, const void* ptr, uint64_t rows, uint64_t cols, uint64_t ld,
                        uint32_t box_cols, uint32_t box_rows) {
  const uint64_t dims[2] = {cols, rows};
  const uint64_t strides[1] = {ld * 2};
  const uint32_t box[2] = {box_cols, box_rows};
  return encode_tmap(tm, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, ptr, 2, dims, strides, box, CU_TENSOR_MAP_SWIZZLE_128B);
}

}  // namespace idf

using namespace idf;

extern "C" int idf_conv2d_igemm(const idf_igemm_args* a, idf_stream_t stream) {
  if (a == nullptr || a->a[0].ptr == nullptr || a->w == nullptr)
    return fail(IDF_ERR_ARG, "idf_conv2d_igemm: null argument");
  const idf_nhwc_t& x0 = a->a[0];
  const int nseg = a->a[1].ptr != nullptr ? 2 : 1;
  for (int s = 0; s < nseg; ++s) {
    const idf_nhwc_t& x = a->a[s];
    if (x.n != x0.n || x.h != x0.h || x.w != x0.w) return fail(IDF_ERR_ARG, "igemm: segments disagree on n/h/w");
    if (x.c <= 0 || x.c % BLOCK_K != 0) return fail(IDF_ERR_UNSUPPORTED, "igemm: channels %d not a multiple of 64", x.c);
    const bool custom = s == 0 && a->custom_taps != 0;
    const bool up2_all_seg = s == 0 && a->out_up2 == 2 && !a->custom_taps && a->taps[0] == 4;
    if (!up2_all_seg && (custom ? (a->taps[s] < 1 || a->taps[s] > 9) : (a->taps[s] != 1 && a->taps[s] != 9)))
      return fail(IDF_ERR_ARG, "igemm: taps must be 1 or 9 (1..9 with custom_taps, 4 with out_up2 == 2)");
  }
  const bool narrow = a->out_nchw != nullptr;
  if (narrow) {
    if (a->N != 16 || a->out_nchw_c < 1 || a->out_nchw_c > 16 || a->res || a->vt || a->ws || a->out_up2 || a->w_mn ||
        a->s2_batch || a->s2_direct || a->rowbias || a->zero_pad_last || a->a[1].ptr || a->w_batch_row || a->w_batch_col)
      return fail(IDF_ERR_ARG, "igemm: out_nchw takes one plain segment, N == 16 (zero-padded weights) and 1..16 channels");
  } else if (a->out == nullptr) {
    return fail(IDF_ERR_ARG, "idf_conv2d_igemm: null output");
  }
  if (!narrow && (a->N <= 0 || a->N % BLOCK_N != 0))
    return fail(IDF_ERR_UNSUPPORTED, "igemm: N = %d not a multiple of %d", a->N, BLOCK_N);

  IgemmParams p;
  memset(&p, 0, sizeof(p));
  const bool up2_all = a->out_up2 == 2;
  if (up2_all && (a->custom_taps || a->taps[0] != 4 || a->a[1].ptr != nullptr || a->w_mn || a->ws != nullptr))
    return fail(IDF_ERR_ARG, "igemm: out_up2 == 2 takes one 4-tap segment, stacked K-major weights and no workspace");
  p.up2_all = up2_all ? 1 : 0;
  const int par_tiles = up2_all ? 4 : 1;
  const int s2d = a->s2_direct ? 1 : 0;
  if (s2d && (a->s2_batch > 0 || a->taps[0] != 9 || nseg != 1 || a->custom_taps || (x0.h & 1) || (x0.w & 1)))
    return fail(IDF_ERR_ARG, "igemm: s2_direct needs one 9-tap segment over an even-sized full-resolution input");
  const int H = s2d ? x0.h / 2 : x0.h, W = s2d ? x0.w / 2 : x0.w, HW = H * W;  // OUTPUT grid
  if (a->s2_batch > 0 && (x0.n != 4 * a->s2_batch || a->taps[0] != 9 || nseg != 1 || a->custom_taps))
    return fail(IDF_ERR_ARG, "igemm: s2_batch needs one 9-tap segment holding 4*s2_batch parity planes");
  const long long M = (long long)(a->s2_batch > 0 ? a->s2_batch : x0.n) * HW;
  p.s2_batch = a->s2_batch;
  p.s2_direct = s2d;
  p.splits = 1;
  if (M <= 0 || M > 0x7fffffffLL) return fail(IDF_ERR_ARG, "igemm: bad M");
  // M-tile geometry: 128 consecutive NHWC pixels = tile_n images x tile_h rows x tile_w columns.
  const bool is_matrix = x0.n == 1 && x0.h == 1 && a->taps[0] == 1 && !a->custom_taps && (nseg == 1 || a->taps[1] == 1);
  if (is_matrix) {
    p.matrix = 1; p.tile_w = BLOCK_M; p.tile_h = 1; p.tile_n = 1; p.tiles_per_img = 0;
  } else if (W > BLOCK_M) {
    return fail(IDF_ERR_UNSUPPORTED, "igemm: image width %d > 128", W);
  } else if (HW >= BLOCK_M) {
    if (BLOCK_M % W != 0 || HW % BLOCK_M != 0) return fail(IDF_ERR_UNSUPPORTED, "igemm: %dx%d image does not tile", H, W);
    p.tile_w = W; p.tile_h = BLOCK_M / W; p.tile_n = 1; p.tiles_per_img = HW / BLOCK_M;
  } else {
    if (BLOCK_M % HW != 0) return fail(IDF_ERR_UNSUPPORTED, "igemm: %dx%d image does not tile", H, W);
    p.tile_w = W; p.tile_h = H; p.tile_n = BLOCK_M / HW; p.tiles_per_img = 0;
  }
  // tap offsets: 3x3 stride 1 pad 1 -> (kh-1, kw-1); stride-2 pad-0 over parity planes -> plane (kh&1, kw&1)
  // shifted by (kh>>1, kw>>1); custom_taps -> the caller's list (segment 0 only)
  for (int s = 0; s < nseg; ++s)
    for (int t = 0; t < a->taps[s]; ++t) {
      if (s == 0 && a->custom_taps) {
        p.tdh[s][t] = a->tap_dh[t]; p.tdw[s][t] = a->tap_dw[t];
      } else if (a->taps[s] == 9 && s2d && s == 0) {
        p.tdh[s][t] = (signed char)(t / 3); p.tdw[s][t] = (signed char)(t % 3);
      } else if (a->taps[s] == 9 && a->s2_batch > 0 && s == 0) {
        const int kh = t / 3, kw = t % 3;
        p.tdn[t] = ((kh & 1) * 2 + (kw & 1)) * a->s2_batch;
        p.tdh[s][t] = (signed char)(kh >> 1); p.tdw[s][t] = (signed char)(kw >> 1);
      } else if (a->taps[s] == 9) {
        p.tdh[s][t] = (signed char)(t / 3 - 1); p.tdw[s][t] = (signed char)(t % 3 - 1);
      }
    }
  int rc;
  int ktot = 0;
  for (int s = 0; s < nseg; ++s) {
    if ((rc = make_act_map(&p.tmA[s], a->a[s], p.tile_w, p.tile_h, p.tile_n, (s2d && s == 0) ? 2 : 1)) != IDF_OK) return rc;
    p.cb[s] = a->a[s].c / BLOCK_K;
    p.taps[s] = a->taps[s];
    ktot += a->taps[s] * a->a[s].c;
  }
  if (nseg == 1) { p.cb[1] = 1; p.taps[1] = 1; }
  p.kb_seg0 = a->taps[0] * p.cb[0];
  p.kb_total = ktot / BLOCK_K;
  if (a->w_mn) {
    if (nseg != 1) return fail(IDF_ERR_UNSUPPORTED, "igemm: w_mn takes one input segment");
    int maxt = 0;
    for (int t = 0; t < a->taps[0]; ++t) {
      p.wtap[t] = a->custom_taps ? a->w_tap_ids[t] : (signed char)t;
      if (p.wtap[t] < 0 || p.wtap[t] > 8) return fail(IDF_ERR_ARG, "igemm: bad w_tap_ids");
      if (p.wtap[t] > maxt) maxt = p.wtap[t];
    }
    if (a->ldw < (long long)(maxt + 1) * a->N) return fail(IDF_ERR_ARG, "igemm: ldw %lld too small for w_mn", (long long)a->ldw);
    p.w_mn = 1;
  } else if (a->ldw < ktot) return fail(IDF_ERR_ARG, "igemm: ldw %lld < K %d", (long long)a->ldw, ktot);
  const bool batched = a->w_batch_row != 0 || a->w_batch_col != 0;
  if (batched) {
    if (a->w_mn || up2_all || nseg != 1 || a->taps[0] != 1 || a->custom_taps || p.tiles_per_img <= 0 || a->ws != nullptr ||
        a->w_batch_row < 0 || a->w_batch_col < 0 || a->w_batch_row > 0x3fffffff || a->w_batch_col > 0x3fffffff)
      return fail(IDF_ERR_ARG, "igemm: w_batch_* needs one 1-tap K-major segment whose images are multiples of 128 pixels");
    p.bb_row = (int)a->w_batch_row;
    p.bb_col = (int)a->w_batch_col;
    if (a->ldw < ktot + (long long)(x0.n - 1) * p.bb_col) return fail(IDF_ERR_ARG, "igemm: ldw too small for w_batch_col");
  }
  // extent of the weight tensor map (all images' blocks when the second operand is batched)
  const uint64_t w_rows = (uint64_t)a->N * par_tiles + (uint64_t)(x0.n - 1) * p.bb_row;
  const uint64_t w_cols = (uint64_t)ktot + (uint64_t)(x0.n - 1) * p.bb_col;
  // tile width: the persistent kernel takes 256 / 192 / 128 columns per tile; pick the widest that divides N (and
  // the V^T split point) unless that would leave SMs without a tile
  int bn = BLOCK_N;
  // short-K GEMMs (K <= 1024: QKV / out_proj / skip projections) are epilogue- and memory-bound: narrow tiles with
  // triple-buffered staging keep loads, residual prefetch and stores in flight together
  const bool gn = a->gn_mode != 0;
  const bool short_k = !narrow && !gn && ktot <= 1024 && a->taps[0] == 1 && (nseg == 1 || a->taps[1] == 1);
  if (narrow) {
    bn = 16;
  } else if (!short_k) {
    // Tile width by a small cost model, calibrated on B200 with the per-launch table of one sampling step
    // (profiles/r01_sample_step_launches_events.txt): the busiest CTA works through waves(c) = ceil(units / SMs)
    // tiles of K / 64 k-blocks; one k-block takes ~250 / 375 / 425 ns for 128- / 192- / 256-wide tiles (~395 ns on
    // CTA pairs), i.e. the wide tiles are only ~15 % cheaper per FLOP, so they lose whenever they need more waves per
    // FLOP (192 tiles of 256 columns on 148 SMs: two waves where 384 tiles of 128 columns need three half-size ones).
    // The last tile's epilogue is exposed. IDF_IGEMM_TILE_MODEL=0: the former rule (widest width that fills the GPU).
    static const int tile_model = [] { const char* e = getenv("IDF_IGEMM_TILE_MODEL"); return e ? atoi(e) : 1; }();
    const long long m_tiles = (M + BLOCK_M - 1) / BLOCK_M;
    const int cands[3] = {256, 192, 128};
    const double t_kb[3] = {425.0, 375.0, 250.0}, t_epi[3] = {3000.0, 2400.0, 1700.0};
    const int fs = a->force_splits > 1 ? a->force_splits : 1;
    double best = 0.0;
    bn = 0;
    for (int i = 0; i < 3; ++i) {
      const int c = cands[i];
      if (a->N % c != 0) continue;
      if (a->vt != nullptr && a->vt_col0 % c != 0) continue;
      if (!tile_model) {
        if (bn == 0) bn = c;                                                    // widest legal
        if (m_tiles * (a->N / c) * par_tiles >= sm_count()) { bn = c; break; }  // widest that still fills the GPU
        bn = c;                                                                 // otherwise keep narrowing
        continue;
      }
      const long long units = m_tiles * (a->N / c) * par_tiles * fs;
      const long long waves = (units + sm_count() - 1) / sm_count();
      const bool pairs = c == 256 && m_tiles * par_tiles >= sm_count();
      const double t = (double)waves * ((double)p.kb_total / fs) * (pairs ? 395.0 : t_kb[i]) + t_epi[i];
      if (bn == 0 || t < best) { bn = c; best = t; }
    }
    if (bn == 0) return fail(IDF_ERR_UNSUPPORTED, "igemm: N = %d has no legal tile width", a->N);
    if (gn && p.tiles_per_img == 0) bn = BLOCK_N;  // two 8x8 images per tile: 128-wide tiles (table layout of the epilogue)
  }
  // split-K: when the tile list leaves a large part of the GPU idle (8x8 / 4x4 stages), split the K range of each
  // tile over up to 4 work units; fp32 partials go to the caller's workspace and a finish kernel adds them in a
  // fixed order (deterministic) together with the epilogue terms.
  int splits = 1;
  if (!short_k && a->ws != nullptr && a->res == nullptr && a->vt == nullptr && !a->out_f32 && !a->out_up2) {
    const long long units = ((M + BLOCK_M - 1) / BLOCK_M) * (a->N / bn);
    auto eff = [&](int sfac) {
      const long long u = units * sfac;
      return (double)u / (double)(((u + sm_count() - 1) / sm_count()) * sm_count());
    };
    if (a->force_splits > 1) {
      if ((long long)a->force_splits * M * a->N * 4 > a->ws_bytes || p.kb_total < a->force_splits)
        return fail(IDF_ERR_ARG, "igemm: force_splits = %d does not fit the workspace / K", a->force_splits);
      splits = a->force_splits;
    } else {
      double best = eff(1);
      for (int sfac = 2; sfac <= 4; ++sfac) {
        if (p.kb_total / sfac < 8) break;
        if ((long long)sfac * M * a->N * 4 > a->ws_bytes) break;
        if (eff(sfac) > best * 1.15) { best = eff(sfac); splits = sfac; }
      }
    }
  }
  p.splits = splits;
  // CTA pairs (cta_group::2). Measured on B200 (profiles/r01_igemm_pair_vs_single.txt): 256-wide tiles of the
  // 32x32 / 16x16 stages gain 5-9 %, 192-wide tiles lose 13-17 %, everything else is neutral or pays for the
  // cluster start-up: IDF_IGEMM_PAIR = 1 (default) pairs only the former, 2 = wherever legal, 0 = never.
  static const int pair_mode = [] { const char* e = getenv("IDF_IGEMM_PAIR"); return e ? atoi(e) : 1; }();
  const bool pair_legal = !narrow && !batched && M > BLOCK_M && (!a->w_mn || bn != 192) && sm_count() >= 2 &&
                          !(gn && ((p.tiles_per_img & 1) || p.tiles_per_img == 0));
  const bool pair = pair_legal && (pair_mode == 2 || (pair_mode == 1 && bn == 256 && !short_k &&
                                                      (M + BLOCK_M - 1) / BLOCK_M * par_tiles >= sm_count()));
  if (gn) {
    // GroupNorm-fused epilogue: whole images of 128-pixel tiles, plain bf16 output, no split-K
    const bool two_img = !is_matrix && p.tiles_per_img == 0 && p.tile_n == 2;
    if ((a->gn_mode != 1 && a->gn_mode != 2) || narrow || is_matrix || (p.tiles_per_img <= 0 && !two_img) || a->res || a->vt ||
        a->out_f32 || a->out_up2 || a->w_mn || a->ws || a->zero_pad_last || batched || a->s2_batch || a->epi_h || a->epi_w)
      return fail(IDF_ERR_UNSUPPORTED, "igemm: gn_mode needs an image-shaped conv (H*W %% 128 == 0 or H*W == 64) with a plain "
                                       "bf16 output");
    const int ipt = two_img ? 2 : 1, tpi = two_img ? 1 : p.tiles_per_img;
    if (a->gn_groups <= 0 || a->N % a->gn_groups != 0 || (a->N / a->gn_groups) % 4 != 0)
      return fail(IDF_ERR_UNSUPPORTED, "igemm: gn_mode needs channels per group %% 4 == 0 (N = %d, groups = %d)", a->N, a->gn_groups);
    if (!a->gn_gamma || !a->gn_beta || (reinterpret_cast<uintptr_t>(a->gn_gamma) & 15) || (reinterpret_cast<uintptr_t>(a->gn_beta) & 15))
      return fail(IDF_ERR_ARG, "igemm: gn_gamma / gn_beta must be 16-byte aligned fp32 vectors");
    if (a->gn_mode == 2 && (a->gn_out == nullptr || a->gn_ldo < a->N)) return fail(IDF_ERR_ARG, "igemm: gn_mode 2 needs gn_out");
    const long long m_tiles = (M + BLOCK_M - 1) / BLOCK_M;
    const long long cnt_bytes = 256;
    const long long need = cnt_bytes + (m_tiles > x0.n ? m_tiles : (long long)x0.n) * (a->N / 4) * 16;
    if (a->gn_ws == nullptr || (reinterpret_cast<uintptr_t>(a->gn_ws) & 255) || a->gn_ws_bytes < need)
      return fail(IDF_ERR_ARG, "igemm: gn_ws must be 256-byte aligned and hold %lld bytes", need);
    const long long n_tiles = a->N / bn;
    const long long units = (pair ? m_tiles / 2 : m_tiles) * n_tiles;
    const long long walkers = pair ? (units < sm_count() / 2 ? units : sm_count() / 2) : (units < sm_count() ? units : sm_count());
    const long long per_img = (pair ? tpi / 2 : tpi) * n_tiles;
    if (per_img > walkers)
      return fail(IDF_ERR_UNSUPPORTED, "igemm: gn_mode: an image's %lld tiles do not fit one wave of %lld walkers", per_img, walkers);
    p.gn_mode = a->gn_mode;
    p.gn_silu = a->gn_silu ? 1 : 0;
    p.gn_qpg = a->N / a->gn_groups / 4;
    p.gn_groups = a->gn_groups;
    p.gn_ipt = ipt;
    p.gn_tpi = tpi;
    p.gn_images = x0.n;
    p.gn_inv_cnt = 1.0f / ((float)HW * (float)(a->N / a->gn_groups));
    p.gn_eps = a->gn_eps;
    p.gn_gamma = a->gn_gamma;
    p.gn_beta = a->gn_beta;
    p.gn_epoch = reinterpret_cast<unsigned*>(a->gn_ws);
    p.gn_part = reinterpret_cast<GnRecord*>(reinterpret_cast<char*>(a->gn_ws) + cnt_bytes);
    if (a->gn_mode == 2 &&
        (rc = make_mat_map(&p.tmG, a->gn_out, (uint64_t)M, (uint64_t)a->N, (uint64_t)a->gn_ldo, 64, BLOCK_M)) != IDF_OK)
      return rc;
#ifdef IDF_GN_TRACE
    {
      static long long* trace_buf = nullptr;
      if (trace_buf == nullptr) {
        cudaMalloc(&trace_buf, 2 * GN_TRACE_TILES * GN_TRACE_EVENTS * sizeof(long long));
        FILE* fh = fopen("/tmp/idf_gn_trace_ptr", "w");
        if (fh) { fprintf(fh, "%llu", (unsigned long long)(uintptr_t)trace_buf); fclose(fh); }
      }
      cudaMemsetAsync(trace_buf, 0, 2 * GN_TRACE_TILES * GN_TRACE_EVENTS * sizeof(long long), reinterpret_cast<cudaStream_t>(stream));
      p.trace = trace_buf;
    }
#endif
  }
  if (a->w_mn) {  // rows = the A operand's channels (K per tap), columns = (weight tap, output column)
    int maxt = 0;
    for (int t = 0; t < a->taps[0]; ++t) maxt = p.wtap[t] > maxt ? p.wtap[t] : maxt;
    if ((rc = make_mat_map(&p.tmB, a->w, (uint64_t)x0.c, (uint64_t)(maxt + 1) * a->N, (uint64_t)a->ldw, 64, 64)) != IDF_OK)
      return rc;
  } else if ((rc = make_mat_map(&p.tmB, a->w, w_rows, w_cols, (uint64_t)a->ldw, BLOCK_K,
                                (uint32_t)(pair ? bn / 2 : bn))) != IDF_OK)
    return rc;

  p.H = a->epi_h > 0 ? a->epi_h : H;
  p.W = a->epi_w > 0 ? a->epi_w : W;
  p.HW = p.H * p.W;
  auto log2_or_neg = [](int v) {
    if (v <= 0 || (v & (v - 1)) != 0) return -1;
    int sh = 0;
    while ((1 << sh) < v) ++sh;
    return sh;
  };
  p.tpi_shift = log2_or_neg(p.tiles_per_img);
  p.hw_shift = log2_or_neg(p.HW);
  p.w_shift = log2_or_neg(p.W);
  p.M = (int)M; p.N = a->N;
  p.bias = a->bias;
  p.rowbias = a->rowbias;
  p.rowbias_idx = a->rowbias_idx;
  p.rowbias_ld = a->rowbias_ld;
  if (a->rowbias != nullptr && (a->rowbias_ld % 4 != 0 || (reinterpret_cast<uintptr_t>(a->rowbias) & 15)))
    return fail(IDF_ERR_ARG, "igemm: rowbias must be 16-byte aligned with ld %% 4 == 0");
  if (a->bias != nullptr && (reinterpret_cast<uintptr_t>(a->bias) & 15))
    return fail(IDF_ERR_ARG, "igemm: bias must be 16-byte aligned");
  if (a->zero_pad_last) p.flags |= F_ZERO_PAD;
  if (narrow) {
    if (a->bias != nullptr && is_matrix) return fail(IDF_ERR_UNSUPPORTED, "igemm: out_nchw needs an image-shaped input");
    p.out_nchw = a->out_nchw;
    p.nchw_c = a->out_nchw_c;
  } else if (splits > 1) {
    // the GEMM writes raw fp32 partials; bias / time bias / padding mask move to the finish kernel
    if ((reinterpret_cast<uintptr_t>(a->ws) & 15) || a->N % 8 != 0 || a->ldo % 8 != 0 || (reinterpret_cast<uintptr_t>(a->out) & 15))
      return fail(IDF_ERR_ARG, "igemm: split-K alignment");
    p.bias = nullptr; p.rowbias = nullptr; p.flags = F_OUT_F32;
    p.out_f32 = reinterpret_cast<float*>(a->ws);
    p.out_f32_ld = a->N;
    p.split_stride = M * a->N;
  } else if (a->out_f32) {
    if (a->res != nullptr || a->vt != nullptr) return fail(IDF_ERR_UNSUPPORTED, "igemm: fp32 output excludes res/vt");
    if (a->ldo % 4 != 0 || (reinterpret_cast<uintptr_t>(a->out) & 15)) return fail(IDF_ERR_ARG, "igemm: fp32 out alignment");
    p.flags |= F_OUT_F32;
    p.out_f32 = reinterpret_cast<float*>(a->out);
    p.out_f32_ld = a->ldo;
  } else {
    int out_cols = a->N;
    if (a->vt != nullptr) {
      if (a->vt_col0 % BLOCK_N != 0 || a->vt_col0 <= 0 || a->vt_col0 >= a->N)
        return fail(IDF_ERR_ARG, "igemm: vt_col0 must be a positive multiple of %d below N", BLOCK_N);
      p.flags |= F_VT;
      p.vt = reinterpret_cast<__nv_bfloat16*>(a->vt);
      p.vt_col0 = a->vt_col0;
      if (a->vt_ld % 8 != 0 || (reinterpret_cast<uintptr_t>(a->vt) & 15) || M % 8 != 0)
        return fail(IDF_ERR_ARG, "igemm: vt needs 16-byte alignment, vt_ld %% 8 == 0 and M %% 8 == 0");
      p.vt_ld = a->vt_ld;
      out_cols = a->vt_col0;
    }
    if (a->out_up2) {
      // sub-pixel store: row (img, h, w) of this GEMM lands at pixel (img, 2h + ph, 2w + pw) of the (n, 2h, 2w) output
      if (a->res != nullptr || a->vt != nullptr || is_matrix || a->s2_batch > 0 || s2d)
        return fail(IDF_ERR_UNSUPPORTED, "igemm: out_up2 excludes res / vt / matrix / stride-2 inputs");
      if (!up2_all && (a->out_ph < 0 || a->out_ph > 1 || a->out_pw < 0 || a->out_pw > 1))
        return fail(IDF_ERR_ARG, "igemm: bad output parity");
      p.flags |= F_OUT_UP2;
      const long long ldo = a->ldo;
      const uint64_t dims[4] = {(uint64_t)out_cols, (uint64_t)W, (uint64_t)H, (uint64_t)x0.n};
      const uint64_t strides[3] = {(uint64_t)(2 * ldo) * 2, (uint64_t)(4 * W * ldo) * 2, (uint64_t)(4LL * HW * ldo) * 2};
      const uint32_t box[4] = {64u, (uint32_t)p.tile_w, (uint32_t)p.tile_h, (uint32_t)p.tile_n};
      for (int par = 0; par < par_tiles; ++par) {  // one output map per parity (row parity * 2 + column parity)
        const int ph = up2_all ? (par >> 1) : a->out_ph, pw = up2_all ? (par & 1) : a->out_pw;
        const __nv_bfloat16* base = reinterpret_cast<const __nv_bfloat16*>(a->out) + ((long long)ph * 2 * W + pw) * ldo;
        if ((rc = encode_tmap(par == 0 ? &p.tmC : &p.tmCx[par - 1], CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, base, 4, dims,
                              strides, box, CU_TENSOR_MAP_SWIZZLE_128B)) != IDF_OK)
          return rc;
      }
    } else if ((rc = make_mat_map(&p.tmC, a->out, (uint64_t)M, (uint64_t)out_cols, (uint64_t)a->ldo, 64, BLOCK_M)) != IDF_OK)
      return rc;
    if (a->res != nullptr) {
      p.flags |= F_RES;
      if ((rc = make_mat_map(&p.tmR, a->res, (uint64_t)M, (uint64_t)out_cols, (uint64_t)a->ldres, 64, BLOCK_M)) != IDF_OK)
        return rc;
    }
  }

  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  if (short_k) {
    // Without a residual to prefetch, K >= 256 and at least two waves of them, wide tiles with two staging buffers
    // (3- / 4-stage ring) win 7-13 % (QKV projections of the 32x32 / 16x16 stages: fewer per-tile fixed costs in
    // the epilogue-bound regime); with K = 128 or few tiles they lose as much, so those keep 128 x 128 tiles.
    const int wbn = a->N % 256 == 0 ? 256 : (a->N % 192 == 0 ? 192 : 128);
    const long long wtiles = ((M + BLOCK_M - 1) / BLOCK_M) * (a->N / wbn);
    if (wbn != 128 && a->res == nullptr && a->vt == nullptr && !a->w_mn && ktot >= 256 && wtiles >= 2 * sm_count()) {
      if ((rc = make_mat_map(&p.tmB, a->w, w_rows, w_cols, (uint64_t)a->ldw, BLOCK_K, (uint32_t)wbn)) != IDF_OK)
        return rc;
      return wbn == 256 ? launch_persist<256, 2>(p, st, false) : launch_persist<192, 2>(p, st, false);
    }
    return launch_persist<128, 3>(p, st, pair);
  }
  if (gn) {
    if (p.gn_ipt == 2) return launch_persist_ew<128, 1, 2, false, 2>(p, st);
    switch (bn) {
      case 256: return pair ? launch_persist_ew<256, 1, 2, true, 1>(p, st) : launch_persist_ew<256, 1, 2, false, 1>(p, st);
      case 192: return pair ? launch_persist_ew<192, 1, 2, true, 1>(p, st) : launch_persist_ew<192, 1, 2, false, 1>(p, st);
      default: return pair ? launch_persist_ew<128, 1, 2, true, 1>(p, st) : launch_persist_ew<128, 1, 2, false, 1>(p, st);
    }
  }
  switch (bn) {
    case 16: return launch_persist_ew<16, 1, 2, false>(p, st);
    case 256: rc = launch_persist<256, 1>(p, st, pair); break;
    case 192: rc = launch_persist<192, 1>(p, st, pair); break;
    default: rc = launch_persist<128, 1>(p, st, pair); break;
  }
  if (rc != IDF_OK || splits == 1) return rc;
  const long long vecs = M * (a->N / 8);
  long long blocks = (vecs + 255) / 256;
  if (blocks > 148 * 8) blocks = 148 * 8;
  return check_cuda(launch_kernel(splitk_finish_kernel, dim3((unsigned)blocks), dim3(256), 0, st,
                               reinterpret_cast<const float*>(a->ws), p.split_stride, splits,
                               reinterpret_cast<__nv_bfloat16*>(a->out), (long long)a->ldo, (int)M, a->N, a->bias,
                               a->rowbias, a->rowbias_idx, a->rowbias_ld, p.H, p.W, a->zero_pad_last ? 1 : 0),
                    "splitk_finish launch");
}

// Host-side test hook: the launch plan of a GroupNorm-fused launch and the two properties its wait-for-the-image's-tiles
// step relies on, checked by walking every walker's unit list exactly as the kernel does (TileWalk): (1) no walker ever
// holds two units of the same image (a CTA waiting for a tile it has not started yet would wait forever); (2) all units
// of an image fall into the same wave (same position in their walkers' lists), so no tile waits a whole tile time.
// out = {walkers, units per image, waves, violations of (1), images split over two waves}.
extern "C" int idf_gn_plan_check(int32_t m_tiles, int32_t tiles_per_img, int32_t n_tiles, int32_t pair, int32_t sms,
                                 int32_t* out, idf_stream_t /*stream*/) {
  if (out == nullptr || m_tiles <= 0 || tiles_per_img <= 0 || n_tiles <= 0 || sms <= 0 || m_tiles % tiles_per_img != 0 ||
      (pair && ((tiles_per_img & 1) || sms < 2)))
    return fail(IDF_ERR_ARG, "gn_plan_check: bad argument");
  const int units = (pair ? m_tiles / 2 : m_tiles) * n_tiles;
  const int per_img = (pair ? tiles_per_img / 2 : tiles_per_img) * n_tiles;
  int walkers = pair ? sms / 2 : sms;
  if (units < walkers) walkers = units;
  if (per_img > walkers) return fail(IDF_ERR_UNSUPPORTED, "gn_plan_check: an image's %d units exceed %d walkers", per_img, walkers);
  walkers = gn_walkers(walkers, per_img);
  const int images = m_tiles / tiles_per_img;
  int* wave_of = static_cast<int*>(malloc(sizeof(int) * (size_t)images));
  if (wave_of == nullptr) return fail(IDF_ERR_ARG, "gn_plan_check: out of memory");
  for (int i = 0; i < images; ++i) wave_of[i] = -1;
  int same_walker = 0, split = 0, waves = 0;
  for (int w = 0; w < walkers; ++w) {
    TileWalk tw;
    tw.init(w, walkers, 1, n_tiles);
    int prev_img = -1, it = 0;
    for (int u = w; u < units; u += walkers, tw.next(), ++it) {
      const int tile_m = pair ? 2 * tw.tile_m : tw.tile_m;
      const int img = tile_m / tiles_per_img;
      if (img == prev_img) ++same_walker;
      prev_img = img;
      if (wave_of[img] < 0) wave_of[img] = it;
      else if (wave_of[img] != it && wave_of[img] != -2) { ++split; wave_of[img] = -2; }
      if (it + 1 > waves) waves = it + 1;
    }
  }
  free(wave_of);
  out[0] = walkers; out[1] = per_img; out[2] = waves; out[3] = same_walker; out[4] = split;
  return IDF_OK;
}

// Host-side test hook: the work-unit walk of the persistent kernel (TileWalk) for `steps` iterations of the walker that
// starts at unit u0 and advances by `stride` units: out[3 i .. 3 i + 2] = (tile_m, n_idx, split) of its i-th unit.
extern "C" int idf_tile_walk_trace(int32_t u0, int32_t stride, int32_t splits, int32_t n_tiles, int32_t steps,
                                   int32_t* out, idf_stream_t /*stream*/) {
  if (out == nullptr || u0 < 0 || stride <= 0 || splits <= 0 || n_tiles <= 0 || steps < 0)
    return fail(IDF_ERR_ARG, "tile_walk_trace: bad argument");
  TileWalk tw;
  tw.init(u0, stride, splits, n_tiles);
  for (int i = 0; i < steps; ++i, tw.next()) {
    out[3 * i] = tw.tile_m; out[3 * i + 1] = tw.n_idx; out[3 * i + 2] = tw.sp;
  }
  return IDF_OK;
}
