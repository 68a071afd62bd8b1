"""Tensor-level wrappers over the C ABI. Activations are channels-last bf16 matrices of shape (B*H*W, C) (possibly
column slices of a wider buffer, so the row stride may exceed C). Nothing here computes in PyTorch."""
from __future__ import annotations

import math
import os

import torch

from . import native
from .native import IgemmArgs, call, matrix_view, nhwc_view, ptr


def _check_bf16_rows(t: torch.Tensor, name: str) -> None:
    if t.dtype != torch.bfloat16 or t.dim() != 2 or t.stride(1) != 1:
        raise ValueError(f"{name}: expected a 2-D bf16 matrix with unit column stride, got {t.dtype} {tuple(t.shape)} "
                         f"strides {t.stride()}")


def pack_conv_weight(w: torch.Tensor) -> torch.Tensor:
    """OIHW fp32 -> (O, KH*KW*I) bf16, tap-major then input channel (the igemm K order)."""
    o, i, kh, kw = w.shape
    return w.detach().permute(0, 2, 3, 1).reshape(o, kh * kw * i).to(torch.bfloat16).contiguous()


def igemm(segs, w: torch.Tensor, n_out: int, out: torch.Tensor, *, bias=None, rowbias=None, rowbias_idx=None,
          res=None, vt=None, vt_col0: int = 0, zero_pad_last: bool = False, epi_hw=None, s2_batch: int = 0, ws=None,
          tap_offsets=None, splits: int = 0, out_up2=None, s2_direct: bool = False, w_mn: bool = False,
          w_tap_ids=None, w_batch=None, out_nchw=None, gn=None) -> torch.Tensor:
    """segs: list of (tensor2d, (n, h, w), channels, taps). `out` is a 2-D bf16 (or fp32) matrix view; with `out_nchw`
    (fp32 (B, C <= 16, H, W), weights / bias zero-padded to n_out = 16) the result goes there instead and `out` is None.
    gn = dict(gamma, beta, groups, silu, ws[, out, eps]): GroupNorm(+SiLU) of the result in the epilogue; without `out`
    the normalised tensor replaces the raw one in `out`, with it `out` keeps the raw result and gn["out"] gets the
    normalised one. ws: zero-initialised byte scratch from gn_workspace_bytes()."""
    a = IgemmArgs()
    if not 1 <= len(segs) <= 2:
        raise ValueError("igemm takes one or two input segments")
    for i, (t, (n, h, wd), c, taps) in enumerate(segs):
        _check_bf16_rows(t, f"igemm segment {i}")
        a.a[i] = nhwc_view(t, n, h, wd, c)
        a.taps[i] = taps
    if len(segs) == 1:
        a.a[1] = nhwc_view(None)
        a.taps[1] = 1
    _check_bf16_rows(w, "igemm weight")
    a.w, a.ldw, a.N = w.data_ptr(), w.stride(0), n_out
    if out_nchw is not None:
        a.out, a.ldo, a.out_f32 = None, 0, 0
        a.out_nchw, a.out_nchw_c = out_nchw.data_ptr(), out_nchw.shape[1]
    else:
        a.out, a.ldo = out.data_ptr(), out.stride(0)
        a.out_f32 = 1 if out.dtype == torch.float32 else 0
    a.bias = ptr(bias)
    a.rowbias = ptr(rowbias)
    a.rowbias_idx = ptr(rowbias_idx)
    a.rowbias_ld = rowbias.stride(0) if rowbias is not None else 0
    a.res = ptr(res)
    a.ldres = res.stride(0) if res is not None else 0
    a.vt = ptr(vt)
    a.vt_col0 = vt_col0
    a.vt_ld = vt.stride(0) if vt is not None else 0
    a.zero_pad_last = 1 if zero_pad_last else 0
    a.epi_h, a.epi_w = epi_hw if epi_hw is not None else (0, 0)
    a.s2_batch = s2_batch
    a.ws = ptr(ws)
    a.ws_bytes = ws.numel() * ws.element_size() if ws is not None else 0
    a.force_splits = splits
    a.s2_direct = 1 if s2_direct else 0
    a.w_mn = 1 if w_mn else 0
    if w_batch is not None:  # per-image (row, column) offset of the second operand: a batch of independent GEMMs
        a.w_batch_row, a.w_batch_col = w_batch
    if w_tap_ids is not None:
        for i, t in enumerate(w_tap_ids):
            a.w_tap_ids[i] = t
    if out_up2 == "all":  # all four parities in one launch (w = the four weight matrices stacked along rows)
        a.out_up2 = 2
    elif out_up2 is not None:  # (row parity, column parity) of the 2x-resolution output this launch fills
        a.out_up2, (a.out_ph, a.out_pw) = 1, out_up2
    if gn is not None:
        g_out = gn.get("out")
        a.gn_mode = 2 if g_out is not None else 1
        a.gn_groups, a.gn_silu, a.gn_eps = gn["groups"], 1 if gn["silu"] else 0, gn.get("eps", 1e-5)
        a.gn_gamma, a.gn_beta = gn["gamma"].data_ptr(), gn["beta"].data_ptr()
        if g_out is not None:
            _check_bf16_rows(g_out, "igemm gn out")
            a.gn_out, a.gn_ldo = g_out.data_ptr(), g_out.stride(0)
        a.gn_ws, a.gn_ws_bytes = gn["ws"].data_ptr(), gn["ws"].numel() * gn["ws"].element_size()
    if tap_offsets is not None:  # explicit (dh, dw) list for segment 0
        a.custom_taps = 1
        for i, (dh, dw) in enumerate(tap_offsets):
            a.tap_dh[i], a.tap_dw[i] = dh, dw
    call("idf_conv2d_igemm", a)
    return out if out_nchw is None else out_nchw


def gn_workspace_bytes(images: int, rows: int, n_out: int) -> int:
    """Scratch of the GroupNorm-fused igemm epilogue (idf_igemm_args.gn_ws): launch epoch + per-tile partial records."""
    return 256 + max((rows + 127) // 128, images) * (n_out // 4) * 16


def groupnorm_silu(x: torch.Tensor, y: torch.Tensor, gamma: torch.Tensor, beta: torch.Tensor, B: int, HW: int, C: int,
                   groups: int, silu: bool, eps: float = 1e-5) -> torch.Tensor:
    _check_bf16_rows(x, "groupnorm x")
    _check_bf16_rows(y, "groupnorm y")
    call("idf_groupnorm_silu", x.data_ptr(), x.stride(0), y.data_ptr(), y.stride(0), gamma.data_ptr(), beta.data_ptr(),
         B, HW, C, groups, eps, 1 if silu else 0)
    return y


# Whole-row GroupNorm kernels: measured on B200 (VQ-VAE forward, batch 256: 75.6 -> 63.9 ms) they win once the tensor is
# far beyond the 126 MB L2 (the slab kernel's second pass then comes from HBM in 32-byte pieces); below that the slab
# kernel's single launch wins (KL decode of batch 48: 4.03 vs 4.23 ms; 403 MB tensor: 129 vs 168 us): threshold 1 GB. Issuing the work in L2-sized sample groups
# (IDF_GN_L2_MB > 0) was measured slower than all samples at once (67.1 vs 63.9 ms): default off.
GN_ROWS_MIN_TOTAL_BYTES = 1 << 30
GN_ROWS_L2_BYTES = int(os.environ.get("IDF_GN_L2_MB", "0")) << 20


def groupnorm_silu_rows(x, y, gamma, beta, B: int, HW: int, C: int, groups: int, silu: bool, ws: torch.Tensor,
                        eps: float = 1e-5):
    """GroupNorm(+SiLU) for large images: whole-row chunks, partial statistics in `ws`, L2-sized sample groups."""
    _check_bf16_rows(x, "groupnorm x")
    _check_bf16_rows(y, "groupnorm y")
    call("idf_groupnorm_silu_rows", x.data_ptr(), x.stride(0), y.data_ptr(), y.stride(0), gamma.data_ptr(),
         beta.data_ptr(), B, HW, C, groups, eps, 1 if silu else 0, ws.data_ptr(), ws.numel() * ws.element_size(),
         GN_ROWS_L2_BYTES)
    return y


def attention(qk: torch.Tensor, vt: torch.Tensor, out: torch.Tensor, M: int, T: int, heads: int, head_dim: int):
    _check_bf16_rows(qk, "attention qk")
    _check_bf16_rows(vt, "attention vt")
    _check_bf16_rows(out, "attention out")
    call("idf_attention_fwd", qk.data_ptr(), qk.stride(0), vt.data_ptr(), vt.stride(0), out.data_ptr(), out.stride(0),
         M, T, heads, head_dim, 1.0 / math.sqrt(head_dim))
    return out


def softmax_rows(s: torch.Tensor, out: torch.Tensor, scale: float) -> torch.Tensor:
    call("idf_softmax_rows", s.data_ptr(), s.stride(0), out.data_ptr(), out.stride(0), s.shape[0], s.shape[1], scale)
    return out


def embed_time_class(t, ctx, ctx_mask, factor, w1, b1, w2, b2, class_w, wp, bp, out, scratch):
    R, D, P = t.shape[0], factor.shape[0] * 2, wp.shape[0]
    call("idf_embed_time_class", t.data_ptr(), ptr(ctx), ptr(ctx_mask), R, D, factor.data_ptr(), w1.data_ptr(),
         b1.data_ptr(), w2.data_ptr(), b2.data_ptr(), ptr(class_w), wp.data_ptr(), bp.data_ptr(), P, out.data_ptr(),
         scratch.data_ptr())
    return out


def cfg_posterior_step(xt, eps_c, eps_u, noise, cfg, t, sched, x_prev, x0_out=None, x_prev_dup=None):
    N = xt.shape[0]
    chw = xt.numel() // N
    t_stride = 0 if t.numel() == 1 else 1
    call("idf_cfg_posterior_step", xt.data_ptr(), eps_c.data_ptr(), eps_u.data_ptr(), noise.data_ptr(), cfg.data_ptr(),
         t.data_ptr(), t_stride, sched.betas.data_ptr(), sched.alphas.data_ptr(), sched.alpha_cum_prod.data_ptr(),
         sched.sqrt_alpha_cum_prod.data_ptr(), sched.sqrt_one_minus_alpha_cum_prod.data_ptr(), x_prev.data_ptr(),
         ptr(x_prev_dup), ptr(x0_out), N, chw, sched.num_steps)
    return x_prev


def cfg_ddim_step(xt, eps_c, eps_u, noise, cfg, t, t_prev, sched, x_prev, eta: float = 0.0, clamp_x0: bool = False,
                  x0_out=None):
    """Guidance mix + one strided step t -> t_prev (device int64 scalars; t_prev < 0 = final step)."""
    N = xt.shape[0]
    call("idf_cfg_ddim_step", xt.data_ptr(), eps_c.data_ptr(), eps_u.data_ptr(), ptr(noise), cfg.data_ptr(),
         t.data_ptr(), t_prev.data_ptr(), sched.alpha_cum_prod.data_ptr(), float(eta), 1 if clamp_x0 else 0,
         x_prev.data_ptr(), ptr(x0_out), N, xt.numel() // N)
    return x_prev


def add_noise(x, noise, t, sched, out):
    N = x.shape[0]
    call("idf_add_noise", x.data_ptr(), noise.data_ptr(), t.data_ptr(), sched.sqrt_alpha_cum_prod.data_ptr(),
         sched.sqrt_one_minus_alpha_cum_prod.data_ptr(), out.data_ptr(), N, x.numel() // N)
    return out


def vq_argmin(z: torch.Tensor, codebook: torch.Tensor, idx_out: torch.Tensor, zq_out=None):
    """z: fp32 (rows, dim) row-major, or an NCHW (B, dim, H, W) tensor (rows are then (image, pixel))."""
    if z.dim() == 4:
        B, dim, H, W = z.shape
        rows, hw = B * H * W, H * W
    else:
        (rows, dim), hw = z.shape, 0
    call("idf_vq_argmin", z.data_ptr(), codebook.data_ptr(), idx_out.data_ptr(), ptr(zq_out), rows, dim,
         codebook.shape[0], hw)
    return idx_out


def vq_loss_perplexity(z: torch.Tensor, zq: torch.Tensor, idx: torch.Tensor, size: int, beta: float):
    """Eval-mode tail of Codebook.forward: (straight-through output, commitment loss, perplexity) - the last two as
    0-dim device tensors."""
    out = torch.empty_like(z)
    stats = torch.empty(2, device=z.device, dtype=torch.float32)
    ws = torch.empty(size + 592, device=z.device, dtype=torch.int32)
    rows = idx.numel()
    call("idf_vq_loss_perplexity", z.data_ptr(), zq.data_ptr(), idx.data_ptr(), out.data_ptr(), rows, z.numel() // rows,
         size, float(beta), stats.data_ptr(), stats.data_ptr() + 4, ws.data_ptr(), ws.numel() * 4)
    return out, stats[0], stats[1]


def kl_loss_reparam(z6: torch.Tensor, noise=None):
    """KL bottleneck (vae.py:99-113) on the encoder's (B, 2*z, H, W) output: returns (mean KL loss as a 0-dim device
    tensor, reparametrised sample (B, z, H, W) or None when `noise` is None)."""
    B = z6.shape[0]
    half = z6.numel() // B // 2
    kl = torch.empty(B + 1, device=z6.device, dtype=torch.float32)
    z_out = None if noise is None else torch.empty(B, z6.shape[1] // 2, *z6.shape[2:], device=z6.device, dtype=torch.float32)
    call("idf_kl_loss_reparam", z6.data_ptr(), ptr(noise), ptr(z_out), kl.data_ptr(), kl.data_ptr() + 4 * B, B, half)
    return kl[B], z_out


def conv3x3_small_cin(x_nchw: torch.Tensor, w: torch.Tensor, bias, y: torch.Tensor, dup: bool = False):
    """dup=True also writes the result to rows [B*H*W, 2*B*H*W) of y (batch-doubled CFG input)."""
    B, Cin, H, W = x_nchw.shape
    _check_bf16_rows(y, "conv_small_cin y")
    call("idf_conv3x3_small_cin", x_nchw.data_ptr(), w.data_ptr(), ptr(bias), y.data_ptr(), y.stride(0), B, Cin, H, W,
         w.shape[0], 1 if dup else 0)
    return y


def conv3x3_small_cout(x: torch.Tensor, w: torch.Tensor, bias, y_nchw: torch.Tensor):
    B, Cout, H, W = y_nchw.shape
    _check_bf16_rows(x, "conv_small_cout x")
    call("idf_conv3x3_small_cout", x.data_ptr(), x.stride(0), w.data_ptr(), ptr(bias), y_nchw.data_ptr(), B,
         w.shape[1], H, W, Cout)
    return y_nchw


def conv1x1_small_f32(x_nchw: torch.Tensor, w: torch.Tensor, bias, y_nchw: torch.Tensor):
    B, Cin, H, W = x_nchw.shape
    call("idf_conv1x1_small_f32", x_nchw.data_ptr(), w.data_ptr(), ptr(bias), y_nchw.data_ptr(), B, Cin, w.shape[0],
         H * W)
    return y_nchw


def upsample_nearest2x(x: torch.Tensor, y: torch.Tensor, B: int, H: int, W: int, C: int):
    call("idf_upsample_nearest2x", x.data_ptr(), x.stride(0), y.data_ptr(), y.stride(0), B, H, W, C)
    return y


def space_to_depth2(x: torch.Tensor, y: torch.Tensor, B: int, H: int, W: int, C: int):
    call("idf_space_to_depth2", x.data_ptr(), x.stride(0), y.data_ptr(), B, H, W, C)
    return y


def nchw_to_rows(x_nchw: torch.Tensor, y: torch.Tensor):
    B, Cc, H, W = x_nchw.shape
    call("idf_nchw_f32_to_nhwc_bf16", x_nchw.data_ptr(), y.data_ptr(), y.stride(0), B, Cc, H * W)
    return y


def rows_to_nchw(x: torch.Tensor, y_nchw: torch.Tensor):
    B, Cc, H, W = y_nchw.shape
    call("idf_nhwc_bf16_to_nchw_f32", x.data_ptr(), x.stride(0), y_nchw.data_ptr(), B, Cc, H * W)
    return y_nchw


# =================================================================================================
# training step (backward kernels)
# =================================================================================================
def conv_wgrad(x: torch.Tensor, grid, cin: int, taps: int, dy: torch.Tensor, cout: int, grad: torch.Tensor,
               ws: torch.Tensor, *, accumulate: bool = False, s2_batch: int = 0) -> torch.Tensor:
    """grad (cout, cin, kh, kw) fp32 (+)= sum_m dy[m, :]^T x[m + tap, :]. x / grid / cin / taps as passed to igemm."""
    _check_bf16_rows(x, "wgrad x")
    _check_bf16_rows(dy, "wgrad dy")
    if grad.dtype != torch.float32 or not grad.is_contiguous() or grad.numel() != cout * cin * taps:
        raise ValueError(f"wgrad: grad must be contiguous fp32 with {cout * cin * taps} elements")
    a = native.WgradArgs()
    n, h, w = grid
    a.x = nhwc_view(x, n, h, w, cin)
    a.taps, a.s2_batch = taps, s2_batch
    a.dy, a.ld_dy, a.cout = dy.data_ptr(), dy.stride(0), cout
    a.grad, a.accumulate = grad.data_ptr(), 1 if accumulate else 0
    a.ws, a.ws_bytes = ws.data_ptr(), ws.numel() * ws.element_size()
    call("idf_conv2d_wgrad", a)
    return grad


def pack_dgrad_weight(w: torch.Tensor) -> torch.Tensor:
    """OIHW fp32 -> (I, KH*KW*O) bf16 such that igemm(dy, packed) is the data gradient of the stride-1 'same' conv:
    dx[p] = sum_t dy[p + off(t)] W[:, :, mirrored t]^T (taps mirrored, channel roles swapped)."""
    o, i, kh, kw = w.shape
    return w.detach().flip(2, 3).permute(1, 2, 3, 0).reshape(i, kh * kw * o).to(torch.bfloat16).contiguous()


_S2_PLANE_TAPS = {(p, q): [(kh, kw) for kh in ((0, 2) if p == 0 else (1,)) for kw in ((0, 2) if q == 0 else (1,))]
                  for p in (0, 1) for q in (0, 1)}


def pack_s2_dgrad_weights(w: torch.Tensor):
    """Downsample conv (3x3, stride 2, pad 0): per input parity plane (p, q) the taps that reach it and the packed
    (I, ntaps*O) bf16 weight; plane (p, q) position (a, b) receives dy[a - (kh>>1), b - (kw>>1)] W[:, :, kh, kw]^T."""
    out = []
    for p in (0, 1):
        for q in (0, 1):
            taps = _S2_PLANE_TAPS[(p, q)]
            wp = torch.cat([w.detach()[:, :, kh, kw].t() for kh, kw in taps], dim=1).to(torch.bfloat16).contiguous()
            out.append(([(-(kh >> 1), -(kw >> 1)) for kh, kw in taps], wp))
    return out


def conv_s2_dgrad(dy: torch.Tensor, B: int, OH: int, OW: int, cout: int, packed, planes_out: torch.Tensor, cin: int):
    """Data gradient of the stride-2 conv into the four parity planes (the layout idf_space_to_depth2 produced in the
    forward pass): four small igemm launches with 4 / 2 / 2 / 1 taps. dy must already have its padded last row /
    column zeroed."""
    rows = B * OH * OW
    for k, (offs, wp) in enumerate(packed):
        igemm([(dy, (B, OH, OW), cout, len(offs))], wp, cin, planes_out[k * rows:(k + 1) * rows], tap_offsets=offs)
    return planes_out


def groupnorm_silu_train(x, y, gamma, beta, B, HW, C, groups, silu, stats, eps: float = 1e-5):
    _check_bf16_rows(x, "groupnorm x")
    _check_bf16_rows(y, "groupnorm y")
    call("idf_groupnorm_silu_train", x.data_ptr(), x.stride(0), y.data_ptr(), y.stride(0), gamma.data_ptr(),
         beta.data_ptr(), B, HW, C, groups, eps, 1 if silu else 0, stats.data_ptr())
    return y


def groupnorm_silu_bwd(x, dy, dx, gamma, beta, stats, dgamma_part, dbeta_part, B, HW, C, groups, silu, add=None,
                       colsum_part=None):
    for n, t in (("x", x), ("dy", dy), ("dx", dx)):
        _check_bf16_rows(t, f"groupnorm_bwd {n}")
    call("idf_groupnorm_silu_bwd", x.data_ptr(), x.stride(0), dy.data_ptr(), dy.stride(0), ptr(add),
         add.stride(0) if add is not None else 0, dx.data_ptr(), dx.stride(0), gamma.data_ptr(), beta.data_ptr(),
         stats.data_ptr(), dgamma_part.data_ptr(), dbeta_part.data_ptr(), ptr(colsum_part),
         colsum_part.stride(0) if colsum_part is not None else 0, B, HW, C, groups, 1 if silu else 0)
    return dx


def groupnorm_bwd_finalize(dgamma_part, dbeta_part, B, C, g_gamma, g_beta, colsum_part=None, g_bias1=None, g_bias2=None):
    call("idf_groupnorm_bwd_finalize", dgamma_part.data_ptr(), dbeta_part.data_ptr(), ptr(colsum_part),
         colsum_part.stride(0) if colsum_part is not None else 0, B, C, g_gamma.data_ptr(), g_beta.data_ptr(),
         ptr(g_bias1), ptr(g_bias2))


def reduce_rows(src: torch.Tensor, rows: int, cols: int, out: torch.Tensor, accumulate: bool = False, ld=None):
    call("idf_reduce_rows_f32", src.data_ptr(), ld if ld is not None else src.stride(0), rows, cols, out.data_ptr(),
         1 if accumulate else 0)
    return out


def colsum(x: torch.Tensor, B: int, HW: int, C: int, per_sample: torch.Tensor, total=None, accumulate=False):
    _check_bf16_rows(x, "colsum x")
    call("idf_colsum_bf16", x.data_ptr(), x.stride(0), B, HW, C, per_sample.data_ptr(), per_sample.stride(0),
         ptr(total), 1 if accumulate else 0)
    return per_sample


def sum2x2(x, y, B, H, W, C):
    call("idf_sum2x2_bf16", x.data_ptr(), x.stride(0), y.data_ptr(), y.stride(0), B, H, W, C)
    return y


def depth_to_space2(planes, y, B, H, W, C, add=None):
    call("idf_depth_to_space2", planes.data_ptr(), y.data_ptr(), y.stride(0), ptr(add),
         add.stride(0) if add is not None else 0, B, H, W, C)
    return y


def zero_last_rowcol(x, B, H, W, C):
    call("idf_zero_last_rowcol", x.data_ptr(), x.stride(0), B, H, W, C)
    return x


def conv3x3_small_cin_wgrad(x_nchw, dy, grad_w, part):
    B, Cin, H, W = x_nchw.shape
    call("idf_conv3x3_small_cin_wgrad", x_nchw.data_ptr(), dy.data_ptr(), dy.stride(0), grad_w.data_ptr(),
         part.data_ptr(), part.numel() * 4, B, Cin, H, W, grad_w.shape[0])
    return grad_w


def conv3x3_small_cout_bwd(h, dout_nchw, w, dh, grad_w, grad_b, part):
    B, Cout, H, W = dout_nchw.shape
    call("idf_conv3x3_small_cout_bwd", h.data_ptr(), h.stride(0), dout_nchw.data_ptr(), w.data_ptr(), dh.data_ptr(),
         dh.stride(0), grad_w.data_ptr(), grad_b.data_ptr(), part.data_ptr(), part.numel() * 4, B, w.shape[1], H, W, Cout)
    return dh


def embed_time_class_train(t, ctx, ctx_mask, factor, w1, b1, w2, b2, class_w, wp, bp, out, saved):
    R, D, P = t.shape[0], factor.shape[0] * 2, wp.shape[0]
    call("idf_embed_time_class_train", t.data_ptr(), ptr(ctx), ptr(ctx_mask), R, D, factor.data_ptr(), w1.data_ptr(),
         b1.data_ptr(), w2.data_ptr(), b2.data_ptr(), ptr(class_w), wp.data_ptr(), bp.data_ptr(), P, out.data_ptr(),
         saved.data_ptr())
    return out


def embed_time_class_bwd(dtable, ctx, ctx_mask, D, num_classes, w2, wp, saved, g_w1, g_b1, g_w2, g_b2, g_cls, g_wp,
                         g_bp, scratch):
    R, P = dtable.shape
    call("idf_embed_time_class_bwd", dtable.data_ptr(), ptr(ctx), ptr(ctx_mask), R, D, P, num_classes, w2.data_ptr(),
         wp.data_ptr(), saved.data_ptr(), g_w1.data_ptr(), g_b1.data_ptr(), g_w2.data_ptr(), g_b2.data_ptr(),
         ptr(g_cls), g_wp.data_ptr(), g_bp.data_ptr(), scratch.data_ptr(), scratch.numel() * 4)


def mse_loss_grad(pred, target, dpred, loss, grad_scale: float = 1.0):
    call("idf_mse_loss_grad", pred.data_ptr(), target.data_ptr(), pred.numel(), grad_scale, ptr(dpred), ptr(loss))
    return loss


def grad_norm_clip(grad, out2, scratch, max_norm: float, grad_div: float = 1.0):
    call("idf_grad_norm_clip", grad.data_ptr(), grad.numel(), grad_div, max_norm, out2.data_ptr(), scratch.data_ptr(),
         scratch.numel() * 4)
    return out2


def adam_hyper(lr: float, step: int, beta1=0.9, beta2=0.999):
    """Host values of the device `hyper` array of idf_adam_step: {lr, 1 - beta1^step, sqrt(1 - beta2^step), 0}."""
    return [lr, 1.0 - beta1 ** step, math.sqrt(1.0 - beta2 ** step), 0.0]


def adam_step(param, grad, exp_avg, exp_avg_sq, hyper, beta1=0.9, beta2=0.999, eps=1e-8, grad_div=1.0, clip2=None):
    call("idf_adam_step", param.data_ptr(), grad.data_ptr(), exp_avg.data_ptr(), exp_avg_sq.data_ptr(), param.numel(),
         hyper.data_ptr(), beta1, beta2, eps, grad_div, ptr(clip2))


def reparam_add_noise(latents, reparam_noise, noise, t, sched, out):
    N = noise.shape[0]
    call("idf_reparam_add_noise", latents.data_ptr(), ptr(reparam_noise), noise.data_ptr(), t.data_ptr(),
         sched.sqrt_alpha_cum_prod.data_ptr(), sched.sqrt_one_minus_alpha_cum_prod.data_ptr(), out.data_ptr(), N,
         noise.numel() // N)
    return out


def attention_train(qk, vt, out, lse, M, T, heads, head_dim):
    call("idf_attention_fwd_train", qk.data_ptr(), qk.stride(0), vt.data_ptr(), vt.stride(0), out.data_ptr(),
         out.stride(0), M, T, heads, head_dim, 1.0 / math.sqrt(head_dim), lse.data_ptr())
    return out


def attention_qkv(qkv, out, M, T, heads, head_dim, lse=None):
    """softmax(Q K^T / sqrt(hd)) V with Q | K | V read from the token-major (M, 3C) QKV GEMM output."""
    _check_bf16_rows(qkv, "attention qkv")
    _check_bf16_rows(out, "attention out")
    call("idf_attention_fwd_qkv", qkv.data_ptr(), qkv.stride(0), out.data_ptr(), out.stride(0), M, T, heads, head_dim,
         1.0 / math.sqrt(head_dim), ptr(lse))
    return out


def attention_bwd(qkv, o, d_out, lse, delta, dqkv, dq32, M, T, heads, head_dim):
    """dqkv (M, 3C) bf16 <- [dQ | dK | dV]. dq32: fp32 (M, C) scratch, required (and zeroed here) when T > 128."""
    C = heads * head_dim
    call("idf_attention_delta", d_out.data_ptr(), d_out.stride(0), o.data_ptr(), o.stride(0), M, heads, head_dim,
         delta.data_ptr())
    if T > 128:
        dq32.zero_()
    call("idf_attention_bwd", qkv.data_ptr(), qkv.stride(0), d_out.data_ptr(), d_out.stride(0), lse.data_ptr(),
         delta.data_ptr(), dqkv.data_ptr(), dqkv.stride(0), ptr(dq32), M, T, heads, head_dim,
         1.0 / math.sqrt(head_dim))
    if T > 128:
        call("idf_f32_to_bf16_rows", dq32.data_ptr(), dqkv.data_ptr(), dqkv.stride(0), M, C)
    return dqkv


# Nearest-2x upsampling followed by a 3x3 'same' conv (Upsample, components.py:124-130) as four sub-pixel convolutions
# on the LOW-resolution input: output pixel (2h+p, 2w+q) only sees 2x2 source pixels, with the 3x3 taps that fall on
# the same source pixel summed. Row parity p: source row offsets and the kernel rows that map onto each of them.
_UP2_ROWS = {0: ((-1, (0,)), (0, (1, 2))), 1: ((0, (0, 1)), (1, (2,)))}


class _UpsamplePack(list):
    """[((p, q), tap offsets, (O, 4*I) bf16)] * 4, plus `.stacked`: the four matrices along rows, (4*O, 4*I)."""
    stacked = None


def pack_upsample_conv_weights(w: torch.Tensor):
    """OIHW fp32 3x3 -> [((p, q), tap offsets [(dh, dw)] * 4, (O, 4*I) bf16)] for the four output parities."""
    out = _UpsamplePack()
    for p in (0, 1):
        for q in (0, 1):
            offs, mats = [], []
            for dh, khs in _UP2_ROWS[p]:
                for dw, kws in _UP2_ROWS[q]:
                    offs.append((dh, dw))
                    mats.append(sum(w.detach()[:, :, kh, kw] for kh in khs for kw in kws))
            out.append(((p, q), offs, torch.cat(mats, dim=1).to(torch.bfloat16).contiguous()))
    out.stacked = torch.cat([wp for _, _, wp in out], dim=0).contiguous()
    return out


def upsample_conv3x3(x: torch.Tensor, grid, c: int, packed, n_out: int, out: torch.Tensor, bias=None):
    """out (B*2H*2W, ld) <- conv3x3(nearest2x(x)) + bias, x (B*H*W, c) at the low resolution `grid` = (B, H, W)."""
    stacked = getattr(packed, "stacked", None)
    if stacked is not None and os.environ.get("IDF_UP2_FUSED", "1") != "0":
        # one launch for the four parities: 4x the tiles per launch (the per-parity launches leave SMs idle: 192 tiles
        # on 148 SMs at the 16x16 -> 32x32 stage) and three launches fewer
        igemm([(x, grid, c, 4)], stacked, n_out, out, bias=bias, out_up2="all")
        return out
    for (p, q), offs, wp in packed:
        igemm([(x, grid, c, 4)], wp, n_out, out, bias=bias, tap_offsets=offs, out_up2=(p, q))
    return out
