/*
 * idf_b200.h — C ABI of libidf_b200.so, the sm_100a kernels behind the drop-in `modules.*` classes.
 *
 * The reference (jklimmek/image-diffusion) has no FFI: its boundary is the Python class surface
 * (modules/unet.py, modules/components.py, modules/diffusion.py, modules/vae.py) and every FLOP is a stock
 * torch op. Each entry point below names the reference call site(s) (file:line under the reference tree)
 * whose arithmetic it replaces. The Python host code (image-diffusion_b200/idf_b200) binds these with ctypes.
 *
 * Conventions
 *   - All pointers are DEVICE pointers borrowed from the caller (torch tensors). The library never allocates
 *     or frees device memory and keeps no pointer past return.
 *   - Activations are channels-last (N, H, W, C) bf16 unless stated otherwise; "ld" arguments are row strides
 *     in ELEMENTS of a (rows, channels) matrix whose rows are pixels/tokens.
 *   - Every call is asynchronous on `stream` and safe to capture into a CUDA graph (no sync, no allocation).
 *   - Return value: 0 = OK, non-zero = error (1 bad argument, 2 CUDA error, 3 unsupported shape); the message
 *     is available from idf_last_error() on the calling thread. Nothing is ever routed to a CPU fallback.
 */
#ifndef IDF_B200_H_
#define IDF_B200_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define IDF_B200_ABI_VERSION 4

typedef struct CUstream_st* idf_stream_t;

const char* idf_last_error(void);
int idf_abi_version(void);

/*
 * idf_struct_size — sizeof() of the argument structs as this library was compiled, for bindings that re-declare them
 * (ctypes / cgo / JNI): which = 0 idf_nhwc_t, 1 idf_igemm_args, 2 idf_wgrad_args, 3 idf_pack_job; -1 for anything else.
 * A binding whose own layout gives a different size must refuse to call the library.
 */
int idf_struct_size(int which);

/* A channels-last activation view: element (n, h, w, c) lives at ptr[n*sn + h*sh + w*sw + c]. A plain
 * (rows, cols) matrix is the view n = 1, h = 1, w = rows, c = cols, sw = row stride. */
typedef struct idf_nhwc {
  const void* ptr;
  int32_t n, h, w, c;
  int64_t sn, sh, sw;
} idf_nhwc_t;

/*
 * idf_conv2d_igemm — implicit-GEMM convolution / linear layer on tcgen05 tensor cores.
 *
 *   out[m, :] = sum_seg sum_tap A_seg[pixel(m) + tap] . W[:, k-range(seg, tap)]^T  (+ epilogue terms)
 *
 * Replaces nn.Conv2d 3x3 s1 p1 (components.py:455, 33, 36, 125; unet.py:45,100), nn.Conv2d 1x1
 * (components.py:501, 43), nn.Linear over tokens (components.py:81-83, 97) and the adds that follow them
 * (components.py:527 time bias, :533 skip conv, :101 and :48 residual).
 *
 *   a[0], a[1]   one or two input segments (a[1].ptr == NULL for one). taps[i] is 9 (3x3, stride 1, zero
 *                padding 1) or 1. Both segments share n/h/w; channel counts must be multiples of 64.
 *   w            bf16 (N, ldw) row-major; column order is segment 0 taps (kh, kw) major then channel, followed
 *                by segment 1 in the same order.
 *   out          bf16 (M, ldo) written for columns [0, N) when out_f32 == 0, fp32 otherwise. M = n*h*w.
 *   bias         fp32 [N] or NULL.
 *   rowbias      fp32 (R, rowbias_ld) or NULL: row rowbias_idx[sample] (or `sample` when idx is NULL) is added
 *                to every pixel of that sample (the per-sample time-projection bias).
 *   res          bf16 (M, ldres) residual added in the epilogue, or NULL.
 *   vt           when non-NULL, output columns >= vt_col0 are NOT written to `out` but transposed into
 *                vt[(col - vt_col0) * vt_ld + m] (the V^T operand of idf_attention_fwd).
 *   zero_pad_last  when 1, rows whose pixel is in the last output row or column are written as exact zeros
 *                (Downsample's ConstantPad2d on the conv OUTPUT, components.py:110-117).
 *   epi_h/epi_w  image geometry for rowbias / zero_pad_last when the A operand is a flattened matrix
 *                (e.g. a pre-gathered patch matrix); 0 means "use a[0].h / a[0].w".
 */
typedef struct idf_igemm_args {
  idf_nhwc_t a[2];
  int32_t taps[2];
  const void* w;
  int64_t ldw;
  int32_t N;
  void* out;
  int64_t ldo;
  int32_t out_f32;
  const float* bias;
  const float* rowbias;
  const int32_t* rowbias_idx;
  int32_t rowbias_ld;
  const void* res;
  int64_t ldres;
  void* vt;
  int32_t vt_col0;
  int64_t vt_ld;
  int32_t zero_pad_last;
  int32_t epi_h, epi_w; /* output image geometry seen by the epilogue (sample = m / (epi_h*epi_w)); 0 = a[0].h/w */
  int32_t s2_batch;     /* > 0: stride-2 pad-0 3x3 conv (Downsample, components.py:110): a[0] holds the four parity
                           planes written by idf_space_to_depth2, stacked along n (n = 4*s2_batch); h/w are the
                           OUTPUT grid. Tap (kh, kw) reads plane (kh&1, kw&1) shifted by (kh>>1, kw>>1). */
  void* ws;             /* optional caller-owned fp32 scratch (16-byte aligned) enabling split-K for GEMMs whose tile
                           list underfills the GPU; NULL = never split */
  int64_t ws_bytes;
  int32_t custom_taps;  /* != 0: segment 0 has taps[0] in 1..9 taps at the explicit offsets below (rows, columns) instead
                           of the 3x3 pattern; w's columns follow the list order. Used for the data gradient of the
                           stride-2 Downsample conv, where each input parity plane sees 4, 2, 2 or 1 taps. */
  int8_t tap_dh[9], tap_dw[9];
  int32_t force_splits; /* > 1 (needs ws): split K into exactly this many work units per tile instead of the occupancy
                           heuristic. A fixed count keeps the summation order - and so the output bits - independent
                           of the batch size. */
  int32_t out_up2;      /* != 0: sub-pixel output. Row (img, h, w) of the GEMM is stored at pixel (img, 2h + out_ph,
                           2w + out_pw) of an (n, 2h, 2w) image in `out` (row stride ldo). Four such launches, one per
                           parity, with 2x2 custom taps and pre-summed weights, ARE nearest-2x upsampling followed by a
                           3x3 conv (Upsample, components.py:124-130) at 4/9 of the FLOPs and without the upsampled
                           tensor ever existing. out_up2 == 2: ONE launch does all four parities: taps[0] = 4 (no custom
                           taps: the 2x2 offsets follow from the parity), `w` holds the four pre-summed (N, 4*Cin) weight
                           matrices stacked along rows in parity order (0,0) (0,1) (1,0) (1,1); out_ph / out_pw unused. */
  int32_t out_ph, out_pw;
  int32_t w_mn;         /* != 0: data-gradient mode. `w` is the FORWARD weight matrix of the layer, (a[0].c rows, ldw), whose
                           column block [t*N, (t+1)*N) holds tap t: out[m, n] = sum_t sum_k A[pixel(m)+tap_t, k] w[k, t*N+n].
                           The tensor core reads it as an MN-major operand, so no transposed weight copy is needed.
                           With custom_taps, w_tap_ids[i] names the column block of the i-th tap. One segment only. */
  int8_t w_tap_ids[9];
  int32_t s2_direct;    /* != 0: stride-2 pad-0 3x3 conv (Downsample, components.py:110) read straight from the
                           FULL-resolution input a[0] (n, 2h, 2w) through a TMA map with element strides (2, 2): no
                           parity-plane copy. The output grid is (n, h, w); use with zero_pad_last. */
  int64_t w_batch_row;  /* batched second operand: image i of a[0] multiplies the (N, K) block of `w` that starts at row */
  int64_t w_batch_col;  /* i * w_batch_row and column i * w_batch_col (both 0: one weight matrix for all images). One 1-tap
                           segment whose images are whole multiples of 128 pixels. Used for the single 384-wide attention
                           head of the VAE (components.py:87-95): S_i = Q_i K_i^T with w = K (w_batch_row = tokens per
                           image) and O_i = P_i V_i with w = V^T (w_batch_col = tokens per image), one launch each for the
                           whole batch. */
  float* out_nchw;      /* narrow output (the network's last convolutions: 128 -> 3 of unet.py:100 / components.py:244, 384 -> z
                           of components.py:183): N must be 16 (weight rows beyond the real channel count zero, bias padded
                           to 16 floats) and columns [0, out_nchw_c) are written as fp32 NCHW planes
                           out_nchw[(img * out_nchw_c + c) * h * w + pixel]; `out` is unused and may be NULL. One plain
                           9- or 1-tap segment. */
  int32_t out_nchw_c;
  int32_t gn_mode;      /* != 0: GroupNorm (+ SiLU) of the convolution result (+ bias + rowbias) applied in the epilogue, i.e.
                           the GroupNorm that opens the NEXT ConvBlock (components.py:448-460) without its own pass over the
                           tensor. 1: `out` receives the normalised tensor only; 2: `out` receives the raw result and `gn_out`
                           the normalised one (the raw tensor is also a residual / skip input). Needs an image-shaped input
                           with h*w a multiple of 128 (or h*w == 64: two images per tile), N / gn_groups a multiple of 4, a plain bf16 output (no res / vt / ws /
                           up2 / zero_pad_last). Statistics are taken from the fp32 accumulators over the whole (sample,
                           group), summed in a fixed order: results do not depend on the batch size. The tiles of a sample
                           exchange their partial sums through gn_ws while the kernel runs, which relies on all of the
                           launch's CTAs being resident (one per SM; the launch never has more CTAs than SMs). */
  int32_t gn_groups;
  int32_t gn_silu;
  float gn_eps;
  const float* gn_gamma; /* fp32 (N), 16-byte aligned */
  const float* gn_beta;
  void* gn_out;         /* gn_mode 2: bf16 (M, N) matrix, row stride gn_ldo */
  int64_t gn_ldo;
  void* gn_ws;          /* 256-byte aligned scratch of 256 + 16 * max(M / 128, images) * (N / 4) bytes, ZERO-filled before its first use:
                           a launch epoch (advanced by every launch) and one 16-byte record of partial sums per (tile,
                           4 channels), each 64-bit word of it tagged with the epoch. One launch at a time per workspace; launches of different
                           shapes may share one. */
  int64_t gn_ws_bytes;
} idf_igemm_args;

int idf_conv2d_igemm(const idf_igemm_args* args, idf_stream_t stream);

/*
 * idf_tile_walk_trace — HOST-side test hook (no GPU work, `out` is a host pointer, `stream` unused): the work-unit walk
 * of the persistent implicit-GEMM kernel. A CTA (or CTA pair) starts at unit u0 = its index and advances by `stride` =
 * the number of walkers; unit u = (tile_m * n_tiles + n_idx) * splits + split. The kernel decomposes the stride once
 * and steps with compare-and-carry adds (the TMA-issuing thread cannot afford integer divisions per tile); this entry
 * point returns what that walk yields so that CPU tests can hold it to the closed form:
 * out[3 i .. 3 i + 2] = (tile_m, n_idx, split) of the walker's i-th unit.
 */
int idf_tile_walk_trace(int32_t u0, int32_t stride, int32_t splits, int32_t n_tiles, int32_t steps, int32_t* out,
                        idf_stream_t stream);

/*
 * idf_gn_plan_check — HOST-side test hook (no GPU work, `out` is a host pointer to 5 ints, `stream` unused): the launch
 * plan of a gn_mode launch over m_tiles M tiles (tiles_per_img per image), n_tiles N tiles, on CTA pairs or not, on a GPU
 * with `sms` SMs, and the two properties the in-kernel wait for an image's tiles relies on, verified by walking every
 * walker's unit list as the kernel does: out = {walkers, units per image, waves, times a walker holds two consecutive
 * units of one image (must be 0: such a CTA would wait for a tile it has not started), images whose units fall into two
 * waves (0 when the walker count is a whole number of images)}.
 */
int idf_gn_plan_check(int32_t m_tiles, int32_t tiles_per_img, int32_t n_tiles, int32_t pair, int32_t sms, int32_t* out,
                      idf_stream_t stream);

/*
 * idf_groupnorm_silu — GroupNorm(groups, C, eps, affine) optionally followed by SiLU, over a channels-last
 * (B, HW, C) bf16 tensor; fp32 statistics. Replaces nn.GroupNorm + nn.SiLU (components.py:31-35, 58, 453-454;
 * unet.py:98-99). x and y are (B*HW, ld) matrices; only C channels are read/written (lets the caller normalise
 * into / out of a concatenated buffer).
 */
int idf_groupnorm_silu(const void* x, int64_t ldx, void* y, int64_t ldy, const float* gamma, const float* beta,
                       int32_t B, int32_t HW, int32_t C, int32_t groups, float eps, int32_t apply_silu,
                       idf_stream_t stream);

/*
 * idf_groupnorm_silu_rows — idf_groupnorm_silu for LARGE images (>= ~1 MB per sample: the VAE's 64x64 .. 128x128
 * stages, components.py:31-35 inside Residual): CTAs own chunks of whole pixel rows (fully coalesced), statistics go
 * through deterministic per-chunk partials in `ws` (fp32, at least B * ceil(HW / 256) * groups * 2 elements), and the
 * work is issued a few samples at a time so that the second read of a sample is served by L2 (l2_bytes = budget for
 * one group of samples, e.g. 48 MB; <= 0: all samples in one pair of launches). Same result contract as
 * idf_groupnorm_silu (fp32 statistics, bit-deterministic, batch invariant). C <= 512, C % 8 == 0.
 */
int idf_groupnorm_silu_rows(const void* x, int64_t ldx, void* y, int64_t ldy, const float* gamma, const float* beta,
                            int32_t B, int32_t HW, int32_t C, int32_t groups, float eps, int32_t apply_silu, float* ws,
                            int64_t ws_bytes, int64_t l2_bytes, idf_stream_t stream);

/*
 * idf_attention_fwd — fused softmax(Q K^T * scale) V per (sample, head), flash style on tcgen05 with the score
 * tile in TMEM. Replaces components.py:86-94 (head split, QK^T / sqrt(hd), softmax, PV, head merge).
 *   qk   bf16 (M, ld_qk): columns [0, C) are Q, [C, 2C) are K, C = heads*head_dim, head-major channel blocks.
 *   vt   bf16 (C, ld_vt): V transposed (channel-major), as written by idf_conv2d_igemm's vt path.
 *   out  bf16 (M, ld_out), columns [0, C).
 *   M = B*T tokens; T tokens per sample (power of two, 16 <= T, T | 128 or 128 | T); head_dim in {16,32,48,64}.
 */
int idf_attention_fwd(const void* qk, int64_t ld_qk, const void* vt, int64_t ld_vt, void* out, int64_t ld_out,
                      int32_t M, int32_t T, int32_t heads, int32_t head_dim, float scale, idf_stream_t stream);

/*
 * idf_softmax_rows — out[r, :] = softmax(in[r, :] * scale), fp32 in, bf16 out. Used by the single-head
 * head_dim = 384 attention of the VAE (components.py:91-92) where the score matrix is formed by two GEMMs.
 */
int idf_softmax_rows(const float* in, int64_t ld_in, void* out, int64_t ld_out, int32_t rows, int32_t cols,
                     float scale, idf_stream_t stream);

/*
 * idf_embed_time_class — TimeEmbedding + class embedding + all per-layer time projections in one call.
 * Replaces components.py:441-445, unet.py:106-114 and every `time_projs[i](timestep)` (components.py:526).
 *   t[r]           int64 timestep of row r (R rows).
 *   ctx[r]         int64 class id or NULL (unconditional); ctx_mask[r] fp32 multiplier or NULL (= 1).
 *   factor         fp32 [time_dim/2] (the `time_embedding.factor` buffer).
 *   w1,b1 / w2,b2  fp32 Linear(time_dim, 4*time_dim) / Linear(4*time_dim, time_dim).
 *   class_w        fp32 (num_classes, time_dim).
 *   wp, bp         fp32 (P, time_dim) / [P]: every layer's time_projs Linear stacked along rows.
 *   out            fp32 (R, P): out[r] = wp @ silu(temb[r]) + bp, temb = MLP(sincos(t/factor)) + mask*class_w[ctx].
 *   scratch        fp32, at least R * 5 * time_dim elements.
 */
int idf_embed_time_class(const int64_t* t, const int64_t* ctx, const float* ctx_mask, int32_t R, int32_t time_dim,
                         const float* factor, const float* w1, const float* b1, const float* w2, const float* b2,
                         const float* class_w, const float* wp, const float* bp, int32_t P, float* out,
                         float* scratch, idf_stream_t stream);

/*
 * idf_cfg_posterior_step — classifier-free-guidance mix fused with the DDPM ancestral step: reads x_t, both
 * epsilon predictions and the noise once, writes x_{t-1} once. Replaces diffusion.py:55 and
 * components.py:405-424 (Scheduler.sample_prev_timestep), including the "no noise at t == 0" branch, with the
 * timestep read on the device (no host sync).
 *   eps_cond/eps_uncond  fp32 (N, chw); cfg[N] fp32 guidance scale per sample; sample n uses timestep
 *   t[n * t_stride] (t_stride = 0: one device scalar for the batch) and, like the reference, the no-noise branch
 *   is decided by t[0]; tables are the Scheduler's fp32 [num_steps] vectors.
 *   x_prev may alias xt (in-place step). x_prev_dup (optional) receives a second copy of x_{t-1}: the
 *   unconditional half of the batch-doubled UNet input. x0_out may be NULL (the reference computes x0 and
 *   drops it). num_steps = length of the schedule tables: a sample with t[n] == 0 in a batch whose t[0] != 0 reads
 *   alpha_cum_prod[num_steps - 1], as the reference's negative index does (components.py:419).
 */
int idf_cfg_posterior_step(const float* xt, const float* eps_cond, const float* eps_uncond, const float* noise,
                           const float* cfg, const int64_t* t, int32_t t_stride, const float* betas,
                           const float* alphas,
                           const float* alpha_cum_prod, const float* sqrt_alpha_cum_prod,
                           const float* sqrt_one_minus_alpha_cum_prod, float* x_prev, float* x_prev_dup,
                           float* x0_out, int32_t N, int32_t chw, int32_t num_steps, idf_stream_t stream);

/*
 * idf_cfg_ddim_step — guidance mix fused with one step of the generalised (strided) sampler from timestep t[0] to
 * an arbitrary earlier timestep t_prev[0] (< 0: final step, abar_prev = 1); SURVEY.md §8 row f4: step-skipping
 * samplers on top of Scheduler's tables (components.py:364-397). Song et al., DDIM (ICLR 2021) eq. 12:
 *   x0 = (x_t - sqrt(1 - abar_t) eps) / sqrt(abar_t), clamped to [-1, 1] when clamp_x0 != 0;
 *   sigma = eta sqrt((1 - abar_prev) / (1 - abar_t)) sqrt(1 - abar_t / abar_prev);
 *   x_prev = sqrt(abar_prev) x0 + sqrt(1 - abar_prev - sigma^2) eps + sigma * noise.
 * eta = 0: deterministic (noise may be NULL); eta = 1, t_prev = t - 1, clamp_x0 = 0: the ancestral step of
 * components.py:405-424. Both timesteps are device scalars (no host sync; CUDA-graph replay safe). x_prev may alias xt.
 */
int idf_cfg_ddim_step(const float* xt, const float* eps_cond, const float* eps_uncond, const float* noise,
                      const float* cfg, const int64_t* t, const int64_t* t_prev, const float* alpha_cum_prod,
                      float eta, int32_t clamp_x0, float* x_prev, float* x0_out, int32_t N, int32_t chw,
                      idf_stream_t stream);

/* idf_add_noise — sqrt(acp[t_n]) * x + sqrt(1 - acp[t_n]) * noise with per-sample t (components.py:399-403). */
int idf_add_noise(const float* x, const float* noise, const int64_t* t, const float* sqrt_alpha_cum_prod,
                  const float* sqrt_one_minus_alpha_cum_prod, float* out, int32_t N, int32_t chw,
                  idf_stream_t stream);

/*
 * idf_vq_argmin — nearest codebook entry per latent vector, reproducing torch.cdist's matmul formulation
 * (components.py:272-275): d2 = x1_ . x2_ with x1_ = [-2x, |x|^2, 1], x2_ = [e, 1, |e|^2], clamp_min(0), sqrt,
 * first minimal index. codebook is fp32 (size, dim); idx_out int64 [rows]. With nchw_hw == 0, z (and the optional
 * zq_out = codebook[idx]) are fp32 (rows, dim) row-major; with nchw_hw = H*W they are NCHW tensors and
 * row = image * H*W + pixel (the "B C H W -> B (H W) C" rearrange of components.py:269 folded into the addressing).
 */
int idf_vq_argmin(const float* z, const float* codebook, int64_t* idx_out, float* zq_out, int32_t rows,
                  int32_t dim, int32_t size, int32_t nchw_hw, idf_stream_t stream);

/*
 * idf_vq_loss_perplexity — the rest of Codebook.forward in eval mode (components.py:301-313), after idf_vq_argmin:
 * loss = beta * mean((zq - z)^2), quant_out = z + (zq - z) (the straight-through output, same arithmetic and bits as
 * the reference), perplexity = exp(-sum_k p_k log(p_k + 1e-6)) with p = code usage histogram / rows. z, zq, quant_out
 * are fp32 tensors of rows * dim elements in the SAME layout (elementwise); idx int64 [rows]; loss / perplexity are
 * device scalars; ws = scratch of at least (size + 592) * 4 bytes. Integer histogram + ordered partial sums:
 * deterministic.
 */
int idf_vq_loss_perplexity(const float* z, const float* zq, const int64_t* idx, float* quant_out, int32_t rows,
                           int32_t dim, int32_t size, float beta, float* loss, float* perplexity, void* ws,
                           int64_t ws_bytes, idf_stream_t stream);

/*
 * idf_kl_loss_reparam — the KL bottleneck of VAE.encode (vae.py:99-113). z6 fp32 (B, 2*half): per sample the mean
 * block followed by the log-variance block ((B, 2*z_dim, H, W) NCHW, half = z_dim*H*W). log_var is clamped to [-30, 20];
 * kl_per_sample[b] = -0.5 * sum(1 + lv - mean^2 - exp(lv)); *loss = mean over the batch. With noise / z_out (both or
 * neither; fp32 (B, half)) also z_out = mean + noise * exp(0.5 * lv) (reparametrised sample). Deterministic.
 */
int idf_kl_loss_reparam(const float* z6, const float* noise, float* z_out, float* kl_per_sample, float* loss, int32_t B,
                        int32_t half, idf_stream_t stream);

/*
 * idf_conv3x3_small_cin — direct 3x3 s1 p1 convolution for tiny Cin (the 3-channel latent): fp32 NCHW in,
 * bf16 NHWC out. Replaces unet.py:45,116 (in_conv) and components.py:207-208 (decoder 1x1 folded by the
 * caller + 3x3). w is fp32 OIHW, bias fp32 [Cout]; Cout % 128 == 0, Cin in {3, 4}. With dup != 0 the B*H*W output
 * rows are also written to rows [B*H*W, 2*B*H*W): cond and uncond halves of a batch-doubled CFG pass share x_t.
 */
int idf_conv3x3_small_cin(const float* x, const float* w, const float* bias, void* y, int64_t ldy, int32_t B,
                          int32_t Cin, int32_t H, int32_t W, int32_t Cout, int32_t dup, idf_stream_t stream);

/*
 * idf_conv3x3_small_cout — direct 3x3 s1 p1 convolution for tiny Cout: bf16 NHWC in (already normalised and
 * activated), fp32 NCHW out. Replaces unet.py:100 (out_conv's Conv2d) and components.py:240,178. w fp32 OIHW.
 */
int idf_conv3x3_small_cout(const void* x, int64_t ldx, const float* w, const float* bias, float* y, int32_t B,
                           int32_t Cin, int32_t H, int32_t W, int32_t Cout, idf_stream_t stream);

/* idf_conv1x1_small_f32 — fp32 NCHW 1x1 convolution with a handful of channels (components.py:207 decoder's
 * first conv, components.py:179 encoder's last conv). w fp32 (Cout, Cin). */
int idf_conv1x1_small_f32(const float* x, const float* w, const float* bias, float* y, int32_t B, int32_t Cin,
                          int32_t Cout, int32_t HW, idf_stream_t stream);

/* idf_upsample_nearest2x — nn.Upsample(scale_factor=2) (components.py:124,128) on channels-last bf16. */
int idf_upsample_nearest2x(const void* x, int64_t ldx, void* y, int64_t ldy, int32_t B, int32_t H, int32_t W,
                           int32_t C, idf_stream_t stream);

/*
 * idf_space_to_depth2 — splits a channels-last bf16 image into its four (row parity, column parity) planes:
 * y[(ph*2+pw)*B + b, h, w, :] = x[b, 2h+ph, 2w+pw, :]. Feeds the stride-2 tap addressing of idf_conv2d_igemm (the
 * Downsample conv of components.py:110) without an im2col matrix. y is (4*B*(H/2)*(W/2), C) dense.
 */
int idf_space_to_depth2(const void* x, int64_t ldx, void* y, int32_t B, int32_t H, int32_t W, int32_t C,
                        idf_stream_t stream);

/* idf_nchw_f32_to_nhwc_bf16 / idf_nhwc_bf16_to_nchw_f32 — layout + dtype conversion at the module boundary. */
int idf_nchw_f32_to_nhwc_bf16(const float* x, void* y, int64_t ldy, int32_t B, int32_t C, int32_t HW,
                              idf_stream_t stream);
int idf_nhwc_bf16_to_nchw_f32(const void* x, int64_t ldx, float* y, int32_t B, int32_t C, int32_t HW,
                              idf_stream_t stream);

/* =====================================================================================================
 * Training step (trainers/diffusion_trainer.py:141-187): backward kernels of the UNet path. The data gradient of
 * every convolution / linear layer is idf_conv2d_igemm itself, run on the output gradient with a re-packed weight
 * (taps mirrored, channel roles swapped); everything else is below.
 * ===================================================================================================== */

/*
 * idf_conv2d_wgrad — weight gradient of nn.Conv2d 3x3 (s1 p1, or the stride-2 Downsample conv) / 1x1 / nn.Linear:
 *   grad[co, ci, kh, kw] (+)= sum over output pixels m of dy[m, co] * x[pixel(m) + (kh, kw), ci]
 * (what autograd computes for components.py:455,501,110,125 and the Linear layers of components.py:81-83,97).
 *   x        the layer's forward input, channels-last bf16 (the same view that was passed to idf_conv2d_igemm).
 *   dy       bf16 (M, ld_dy) gradient of the layer output, M = n*h*w output pixels (s2_batch*h*w for stride 2).
 *   grad     fp32, PyTorch parameter layout (cout, cin, kh, kw) contiguous; accumulate != 0 adds to it.
 *   ws       caller-owned fp32 scratch for split-K partials, at least cout*taps*cin*4 bytes (more = more splits).
 * Deterministic: partials are summed in a fixed order.
 */
typedef struct idf_wgrad_args {
  idf_nhwc_t x;
  int32_t taps;
  int32_t s2_batch;
  const void* dy;
  int64_t ld_dy;
  int32_t cout;
  float* grad;
  int32_t accumulate;
  void* ws;
  int64_t ws_bytes;
} idf_wgrad_args;

int idf_conv2d_wgrad(const idf_wgrad_args* args, idf_stream_t stream);


/* idf_groupnorm_silu_train — idf_groupnorm_silu that also stores (mean, rstd) per (sample, group) into
 * stats[(b*groups + g)*2 + {0,1}] for idf_groupnorm_silu_bwd. */
int idf_groupnorm_silu_train(const void* x, int64_t ldx, void* y, int64_t ldy, const float* gamma, const float* beta,
                             int32_t B, int32_t HW, int32_t C, int32_t groups, float eps, int32_t apply_silu,
                             float* stats, idf_stream_t stream);

/*
 * idf_groupnorm_silu_bwd — backward of GroupNorm (+SiLU) (components.py:453-454, 58; unet.py:98-99):
 *   dx = rstd * (gamma*dz - mean_g(gamma*dz) - xhat * mean_g(gamma*dz*xhat)) (+ add),  dz = dy * silu'(gamma*xhat+beta)
 * x is the forward INPUT, dy the gradient of the forward output; add (optional, bf16) is summed into dx (second
 * gradient path into the same tensor: the residual / skip branch). dgamma_part / dbeta_part are fp32 (B, C)
 * per-sample contributions. colsum_part (optional, fp32, row stride ld_cs) receives the per-sample column sums of dx:
 * x is the output of a conv, so these are that conv's bias gradient per sample (and the per-sample gradient of its
 * time-projection bias, components.py:526-527). idf_groupnorm_bwd_finalize sums all three over the batch.
 */
int idf_groupnorm_silu_bwd(const void* x, int64_t ldx, const void* dy, int64_t lddy, const void* add, int64_t ldadd,
                           void* dx, int64_t lddx, const float* gamma, const float* beta, const float* stats,
                           float* dgamma_part, float* dbeta_part, float* colsum_part, int64_t ld_cs, int32_t B,
                           int32_t HW, int32_t C, int32_t groups, int32_t apply_silu, idf_stream_t stream);

/* idf_groupnorm_bwd_finalize — g_gamma[c] = sum_b dgamma_part[b, c], g_beta likewise, and (optional) g_bias1 / g_bias2
 * = sum_b colsum_part[b, c] (two destinations: the second-half conv and the 1x1 skip conv share their bias gradient). */
int idf_groupnorm_bwd_finalize(const float* dgamma_part, const float* dbeta_part, const float* colsum_part,
                               int64_t ld_cs, int32_t B, int32_t C, float* g_gamma, float* g_beta, float* g_bias1,
                               float* g_bias2, idf_stream_t stream);

/* idf_reduce_rows_f32 — out[c] (+)= sum_r in[r*ld + c], rows added in order (deterministic). */
int idf_reduce_rows_f32(const float* in, int64_t ld, int32_t rows, int32_t cols, float* out, int32_t accumulate,
                        idf_stream_t stream);

/* idf_colsum_bf16 — per_sample[b*ld_ps + c] = sum over the HW pixel rows of sample b of x[., c] (bias gradient of a
 * conv / linear layer, and the per-sample gradient of the time-projection bias, components.py:526-527); total
 * (optional, fp32 [C]) (+)= the sum over samples. */
int idf_colsum_bf16(const void* x, int64_t ldx, int32_t B, int32_t HW, int32_t C, float* per_sample, int64_t ld_ps,
                    float* total, int32_t accumulate_total, idf_stream_t stream);

/* idf_sum2x2_bf16 — adjoint of idf_upsample_nearest2x: y[b,h,w,:] = sum of the 2x2 block of x (B, 2H, 2W, C). */
int idf_sum2x2_bf16(const void* x, int64_t ldx, void* y, int64_t ldy, int32_t B, int32_t H, int32_t W, int32_t C,
                    idf_stream_t stream);

/* idf_depth_to_space2 — adjoint of idf_space_to_depth2 (H, W are the FULL resolution), plus an optional bf16 addend
 * (the skip-connection gradient that meets the Downsample gradient at the same tensor). */
int idf_depth_to_space2(const void* planes, void* y, int64_t ldy, const void* add, int64_t ldadd, int32_t B, int32_t H,
                        int32_t W, int32_t C, idf_stream_t stream);

/* idf_zero_last_rowcol — zeroes the last row and column of every image in place: the gradient that reaches the
 * ConstantPad2d region of Downsample's output (components.py:113) is dropped. */
int idf_zero_last_rowcol(void* x, int64_t ldx, int32_t B, int32_t H, int32_t W, int32_t C, idf_stream_t stream);

/* idf_conv3x3_small_cin_wgrad — weight gradient of in_conv (unet.py:45): x fp32 NCHW (B, 3, H, W), dy bf16 (M, lddy),
 * grad_w fp32 (Cout, 3, 3, 3). part: fp32 scratch of B*(H/2)*Cout*27 floats. (Bias gradient: idf_colsum_bf16.) */
int idf_conv3x3_small_cin_wgrad(const float* x, const void* dy, int64_t lddy, float* grad_w, float* part,
                                int64_t part_bytes, int32_t B, int32_t Cin, int32_t H, int32_t W, int32_t Cout,
                                idf_stream_t stream);

/* idf_conv3x3_small_cout_bwd — backward of out_conv's Conv2d (unet.py:100): h bf16 (M, ldh) is its (normalised,
 * activated) input, dout fp32 NCHW (B, 3, H, W) the loss gradient, w fp32 (3, C, 3, 3). Writes dh bf16 (M, lddh),
 * grad_w (3, C, 3, 3), grad_b (3). part: fp32 scratch of B*(H/2)*3*C*9 floats. */
int idf_conv3x3_small_cout_bwd(const void* h, int64_t ldh, const float* dout, const float* w, void* dh, int64_t lddh,
                               float* grad_w, float* grad_b, float* part, int64_t part_bytes, int32_t B, int32_t C,
                               int32_t H, int32_t W, int32_t Cout, idf_stream_t stream);

/* idf_embed_time_class_train — idf_embed_time_class keeping what the backward needs:
 * saved = e [R,D] | z1 [R,4D] | silu(z1) [R,4D] | temb [R,D] | silu(temb) [R,D]  (11*R*D floats). */
int idf_embed_time_class_train(const int64_t* t, const int64_t* ctx, const float* ctx_mask, int32_t R, int32_t time_dim,
                               const float* factor, const float* w1, const float* b1, const float* w2, const float* b2,
                               const float* class_w, const float* wp, const float* bp, int32_t P, float* out,
                               float* saved, idf_stream_t stream);

/* idf_embed_time_class_bwd — gradients of the time MLP, class embedding and all time projections
 * (components.py:429-445, unet.py:42,109-114, components.py:486,526) from dtable (R, P) = the per-sample column
 * sums of every first-half conv's output gradient. scratch: fp32, >= (ceil(P/256)*R*4D + 5*R*D) * 4 bytes. */
int idf_embed_time_class_bwd(const float* dtable, const int64_t* ctx, const float* ctx_mask, int32_t R,
                             int32_t time_dim, int32_t P, int32_t num_classes, const float* w2, const float* wp,
                             const float* saved, float* g_w1, float* g_b1, float* g_w2, float* g_b2, float* g_cls,
                             float* g_wp, float* g_bp, float* scratch, int64_t scratch_bytes, idf_stream_t stream);

/* idf_mse_loss_grad — nn.MSELoss (diffusion_trainer.py:54,170): loss[0] = mean((pred-target)^2) and, when dpred is
 * non-NULL, dpred = grad_scale * 2 (pred - target) / n. */
int idf_mse_loss_grad(const float* pred, const float* target, int64_t n, float grad_scale, float* dpred, float* loss,
                      idf_stream_t stream);

/* idf_grad_norm_clip — nn.utils.clip_grad_norm_ (diffusion_trainer.py:178-181) over one flat fp32 gradient buffer:
 * out2[0] = ||grad||_2 / grad_div, out2[1] = min(1, max_norm / (out2[0] + 1e-6)) (1 when max_norm <= 0).
 * scratch: >= 1184 floats. */
int idf_grad_norm_clip(const float* grad, int64_t n, float grad_div, float max_norm, float* out2, float* scratch,
                       int64_t scratch_bytes, idf_stream_t stream);

/* idf_adam_step — torch.optim.Adam defaults (diffusion_trainer.py:57-58,185) over flat fp32 buffers; the gradient is
 * used as grad * clip2[1] / grad_div (clip2 from idf_grad_norm_clip, may be NULL). hyper is a DEVICE array
 * {lr, 1 - beta1^step, sqrt(1 - beta2^step)} so a captured graph can be replayed with new values. */
int idf_adam_step(float* param, const float* grad, float* exp_avg, float* exp_avg_sq, int64_t n, const float* hyper,
                  float beta1, float beta2, float eps, float grad_div, const float* clip2, idf_stream_t stream);

/* idf_reparam_add_noise — KL reparametrisation of stored (mean || logvar) latents (diffusion_trainer.py:149-155; skipped
 * when reparam_noise is NULL, latents then being (N, chw)) followed by Scheduler.add_noise (components.py:399-403). */
int idf_reparam_add_noise(const float* latents, const float* reparam_noise, const float* noise, const int64_t* t,
                          const float* sqrt_alpha_cum_prod, const float* sqrt_one_minus_alpha_cum_prod, float* out,
                          int32_t N, int32_t chw, idf_stream_t stream);

/* idf_attention_fwd_train — idf_attention_fwd that also writes lse[m*heads + h] = log2(sum_j exp(s_mj)) + max (log2
 * domain, scaled scores) for idf_attention_bwd. */
int idf_attention_fwd_train(const void* qk, int64_t ld_qk, const void* vt, int64_t ld_vt, void* out, int64_t ld_out,
                            int32_t M, int32_t T, int32_t heads, int32_t head_dim, float scale, float* lse,
                            idf_stream_t stream);

/* idf_attention_delta — delta[m*heads + h] = sum_d dO[m, h*hd+d] * O[m, h*hd+d] (softmax backward row term). */
int idf_attention_delta(const void* d_out, int64_t ld_do, const void* out, int64_t ld_o, int32_t M, int32_t heads,
                        int32_t head_dim, float* delta, idf_stream_t stream);

/* idf_attention_fwd_qkv — idf_attention_fwd reading Q, K AND V straight from the token-major (M, ld) QKV matrix
 * (columns [0,C) Q, [C,2C) K, [2C,3C) V): V tiles are consumed as MN-major tcgen05 operands, so the QKV GEMM needs no
 * transposing epilogue. lse (optional, fp32 (M, heads)) as in idf_attention_fwd_train. */
int idf_attention_fwd_qkv(const void* qkv, int64_t ld_qkv, void* out, int64_t ld_out, int32_t M, int32_t T,
                          int32_t heads, int32_t head_dim, float scale, float* lse, idf_stream_t stream);

/*
 * idf_attention_bwd — backward of components.py:86-94 on tcgen05, probabilities recomputed on chip.
 *   qkv         bf16 (M, ld_qkv) token-major [Q | K | V] as written by the QKV GEMM;  d_out bf16 (M, ld_do).
 *   dqkv        bf16 (M, ld_dqkv): dQ -> columns [0,C), dK -> [C,2C), dV -> [2C,3C) (token-major: the operand layout
 *               of the QKV Linear's data and weight gradients).
 *   dq32        fp32 (M, C), ZEROED by the caller, required when T > 128: dQ partials of different key tiles are
 *               accumulated there (fp32 reductions); convert with idf_f32_to_bf16_rows. Unused for T <= 128.
 */
int idf_attention_bwd(const void* qkv, int64_t ld_qkv, const void* d_out, int64_t ld_do, const float* lse,
                      const float* delta, void* dqkv, int64_t ld_dqkv, float* dq32, int32_t M, int32_t T,
                      int32_t heads, int32_t head_dim, float scale, idf_stream_t stream);

/* idf_rowidx_from_timestep — out[i] = base[i] + t[0] * rows_per_t: per-sample row into a table of idf_embed_time_class
 * outputs precomputed for every timestep of a sampling run (the timestep is read on the device, so a captured step
 * needs no host-side index update). */
int idf_rowidx_from_timestep(const int32_t* base, const int64_t* t, int32_t rows_per_t, int32_t* out, int32_t n,
                             idf_stream_t stream);

/*
 * idf_pack_weights — refreshes the kernels' operand copies of the parameters after an optimizer step, all in one launch.
 * Job j copies an fp32 source tensor into its destination layout: for outer index o (one CTA each), tap t, inner index i
 *     dst[o*dldo + t*dldt + i] = src[o*so + i*si + t*st]        (bf16 destination; fp32 and "+ src2[...]" if out_f32)
 * which covers (O,I,kh,kw) -> (O, taps*I) forward layouts, (I, taps*O) data-gradient layouts, transposes and the fused
 * bias / stacked vectors. jobs_dev and cta_prefix_dev (first CTA of each job, ascending) are DEVICE arrays built by the
 * caller once; total_ctas = sum of n_outer. n_taps <= 9; for out_f32 jobs n_taps must be 1.
 */
typedef struct idf_pack_job {
  const float* src;
  const float* src2; /* optional second addend (out_f32 only) */
  void* dst;
  int32_t n_outer, n_taps, n_inner, out_f32;
  int64_t so, si, st;
  int64_t dldo, dldt;
} idf_pack_job;

int idf_pack_weights(const idf_pack_job* jobs_dev, const int32_t* cta_prefix_dev, int32_t njobs, int32_t total_ctas,
                     idf_stream_t stream);

/* idf_u8_nhwc_to_f32_nchw — dataset-scale latent extraction front end (scripts/prepare_dataset.py:103-106): uint8
 * (B, H, W, C) images -> fp32 (B, C, H, W) with y = x * scale + shift (1/127.5, -1). */
int idf_u8_nhwc_to_f32_nchw(const uint8_t* x, float* y, int32_t B, int32_t H, int32_t W, int32_t C, float scale,
                            float shift, idf_stream_t stream);

/* idf_f32_to_bf16_rows — y[m*ldy + c] = bf16(x[m*C + c]). */
int idf_f32_to_bf16_rows(const float* x, void* y, int64_t ldy, int64_t M, int32_t C, idf_stream_t stream);

#ifdef __cplusplus
}
#endif
#endif /* IDF_B200_H_ */
