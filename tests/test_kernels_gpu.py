"""Kernel-level parity (B200): every C-ABI kernel against the same op in plain PyTorch fp32 on the same inputs.

Inputs are rounded to bf16 first where the kernel consumes bf16, so the only differences left are fp32
accumulation order and the final bf16 rounding of the output: tolerances are a few bf16 ulps of the output scale.
"""
import math

import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu

# the PyTorch side must be true fp32 (cuDNN/cuBLAS default to TF32 for convolutions)
torch.backends.cudnn.allow_tf32 = False
torch.backends.cuda.matmul.allow_tf32 = False

DEV = "cuda"


def _ops():
    from idf_b200 import ops
    return ops


def rows(x_nchw):
    """NCHW fp32 -> (B*H*W, C) bf16 channels-last matrix."""
    B, C, H, W = x_nchw.shape
    return x_nchw.permute(0, 2, 3, 1).reshape(B * H * W, C).to(torch.bfloat16).contiguous()


def unrows(y, B, H, W):
    return y.float().reshape(B, H, W, -1).permute(0, 3, 1, 2).contiguous()


def rel_err(a, b):
    return ((a - b).norm() / b.norm().clamp_min(1e-12)).item()


def bf(x):
    return x.to(torch.bfloat16).float()


@pytest.mark.parametrize("B,Cin,Cout,H", [
    (2, 128, 128, 32), (3, 256, 384, 16), (5, 384, 512, 8), (9, 512, 512, 4), (1, 512, 128, 32),
    (2, 1024, 384, 8), (1, 128, 128, 64), (1, 128, 128, 128), (1, 64, 128, 4),
])
def test_conv3x3_igemm(B, Cin, Cout, H):
    ops = _ops()
    g = torch.Generator(device=DEV).manual_seed(B * 1000 + Cin + Cout + H)
    x = torch.randn(B, Cin, H, H, device=DEV, generator=g)
    w = torch.randn(Cout, Cin, 3, 3, device=DEV, generator=g) / math.sqrt(9 * Cin)
    b = torch.randn(Cout, device=DEV, generator=g)
    xr = rows(x)
    out = torch.empty(B * H * H, Cout, device=DEV, dtype=torch.bfloat16)
    ops.igemm([(xr, (B, H, H), Cin, 9)], ops.pack_conv_weight(w), Cout, out, bias=b)
    ref = F.conv2d(bf(x), bf(w), b, padding=1)
    got = unrows(out, B, H, H)
    assert rel_err(got, ref) < 6e-3, rel_err(got, ref)
    assert (got - ref).abs().max().item() < 0.06


@pytest.mark.parametrize("B,Cin,Cout,H", [(20, 128, 256, 32), (80, 64, 512, 16), (297, 64, 256, 8)])
def test_conv3x3_igemm_cta_pair(B, Cin, Cout, H):
    """Layers with >= 148 M tiles and 256-wide N tiles run on CTA pairs (cta_group::2: one 256-row MMA per two CTAs,
    each staging half of the weight tile). The last shape has an odd number of M tiles, the last one partial: the
    odd CTA of the final pair works on a tile that lies entirely outside the matrix."""
    ops = _ops()
    g = torch.Generator(device=DEV).manual_seed(B * 1000 + Cin + Cout + H)
    x = torch.randn(B, Cin, H, H, device=DEV, generator=g)
    w = torch.randn(Cout, Cin, 3, 3, device=DEV, generator=g) / math.sqrt(9 * Cin)
    b = torch.randn(Cout, device=DEV, generator=g)
    table = torch.randn(B, Cout, device=DEV, generator=g)
    out = torch.empty(B * H * H, Cout, device=DEV, dtype=torch.bfloat16)
    ops.igemm([(rows(x), (B, H, H), Cin, 9)], ops.pack_conv_weight(w), Cout, out, bias=b, rowbias=table)
    ref = F.conv2d(bf(x), bf(w), b, padding=1) + table[:, :, None, None]
    got = unrows(out, B, H, H)
    assert rel_err(got, ref) < 6e-3, rel_err(got, ref)
    assert (got - ref).abs().max().item() < 0.08
    out2 = torch.empty_like(out)
    ops.igemm([(rows(x), (B, H, H), Cin, 9)], ops.pack_conv_weight(w), Cout, out2, bias=b, rowbias=table)
    assert torch.equal(out, out2)  # deterministic


@pytest.mark.parametrize("B,Cin,Cout,H", [(9, 512, 512, 4), (24, 384, 512, 8), (96, 512, 512, 4)])
def test_conv3x3_split_k(B, Cin, Cout, H):
    """Small-M layers: K split over several work units, fp32 partials reduced by the finish kernel (with the time
    bias and the fused 1x1 skip segment); deterministic from run to run."""
    ops = _ops()
    g = torch.Generator(device=DEV).manual_seed(B + Cin)
    h = torch.randn(B, Cin, H, H, device=DEV, generator=g)
    x = torch.randn(B, 128, H, H, device=DEV, generator=g)
    w3 = torch.randn(Cout, Cin, 3, 3, device=DEV, generator=g) / math.sqrt(9 * Cin)
    w1 = torch.randn(Cout, 128, 1, 1, device=DEV, generator=g) / math.sqrt(128)
    b = torch.randn(Cout, device=DEV, generator=g)
    table = torch.randn(B, Cout, device=DEV, generator=g)
    wcat = torch.cat([ops.pack_conv_weight(w3), ops.pack_conv_weight(w1)], dim=1).contiguous()
    ws = torch.empty(16 * 1024 * 1024, device=DEV, dtype=torch.float32)
    outs = []
    for rep in range(2):
        out = torch.empty(B * H * H, Cout, device=DEV, dtype=torch.bfloat16)
        ops.igemm([(rows(h), (B, H, H), Cin, 9), (rows(x), (B, H, H), 128, 1)], wcat, Cout, out, bias=b, rowbias=table,
                  ws=ws)
        outs.append(out)
    assert torch.equal(outs[0], outs[1])
    ref = F.conv2d(bf(h), bf(w3), b, padding=1) + F.conv2d(bf(x), bf(w1)) + table[:, :, None, None]
    assert rel_err(unrows(outs[0], B, H, H), ref) < 6e-3


def test_conv3x3_rowbias_and_fused_skip():
    """second_half conv3x3 + 1x1 skip conv as one two-segment GEMM, plus a per-sample time bias."""
    ops = _ops()
    B, C1, C2, Cout, H = 3, 256, 128, 256, 16
    g = torch.Generator(device=DEV).manual_seed(7)
    h = torch.randn(B, C1, H, H, device=DEV, generator=g)
    x = torch.randn(B, C2, H, H, device=DEV, generator=g)
    w3 = torch.randn(Cout, C1, 3, 3, device=DEV, generator=g) / math.sqrt(9 * C1)
    w1 = torch.randn(Cout, C2, 1, 1, device=DEV, generator=g) / math.sqrt(C2)
    b = torch.randn(Cout, device=DEV, generator=g)
    table = torch.randn(4, 512, device=DEV, generator=g)
    idx = torch.tensor([2, 0, 3], device=DEV, dtype=torch.int32)
    wcat = torch.cat([ops.pack_conv_weight(w3), ops.pack_conv_weight(w1)], dim=1).contiguous()
    out = torch.empty(B * H * H, Cout, device=DEV, dtype=torch.bfloat16)
    ops.igemm([(rows(h), (B, H, H), C1, 9), (rows(x), (B, H, H), C2, 1)], wcat, Cout, out, bias=b,
              rowbias=table[:, 128:128 + Cout], rowbias_idx=idx)
    ref = F.conv2d(bf(h), bf(w3), b, padding=1) + F.conv2d(bf(x), bf(w1))
    ref = ref + table[idx.long(), 128:128 + Cout][:, :, None, None]
    got = unrows(out, B, H, H)
    assert rel_err(got, ref) < 6e-3, rel_err(got, ref)


def test_linear_residual_strided_output():
    """out_proj + residual written into the right half of a concatenated buffer."""
    ops = _ops()
    M, K, N = 3 * 256, 384, 384
    g = torch.Generator(device=DEV).manual_seed(11)
    a = torch.randn(M, K, device=DEV, generator=g).to(torch.bfloat16)
    w = (torch.randn(N, K, device=DEV, generator=g) / math.sqrt(K)).to(torch.bfloat16)
    b = torch.randn(N, device=DEV, generator=g)
    res = torch.randn(M, N, device=DEV, generator=g).to(torch.bfloat16)
    cat = torch.zeros(M, 2 * N, device=DEV, dtype=torch.bfloat16)
    ops.igemm([(a, (1, 1, M), K, 1)], w, N, cat[:, N:], bias=b, res=res)
    ref = a.float() @ w.float().t() + b + res.float()
    assert rel_err(cat[:, N:].float(), ref) < 6e-3
    assert cat[:, :N].abs().max().item() == 0.0


@pytest.mark.parametrize("M,K,N", [(49152, 256, 768), (24576, 384, 1152), (98304, 128, 384), (40000, 256, 256)])
def test_linear_large_m_tile_paths(M, K, N):
    """Projection GEMMs at the sizes of the 32x32 / 16x16 stages: with K >= 256 and two or more waves of tiles the
    short-K kernel takes 256- / 192-wide tiles with two staging buffers, otherwise 128 x 128 tiles with three; the
    last shape has a partial final M tile. Against an fp32 matmul of the same bf16 operands."""
    ops = _ops()
    g = torch.Generator(device=DEV).manual_seed(M + K + N)
    a = torch.randn(M, K, device=DEV, generator=g).to(torch.bfloat16)
    w = (torch.randn(N, K, device=DEV, generator=g) / math.sqrt(K)).to(torch.bfloat16)
    b = torch.randn(N, device=DEV, generator=g)
    out = torch.empty(M, N, device=DEV, dtype=torch.bfloat16)
    ops.igemm([(a, (1, 1, M), K, 1)], w, N, out, bias=b)
    ref = a.float() @ w.float().t() + b
    assert rel_err(out.float(), ref) < 6e-3
    assert (out.float() - ref).abs().max().item() < 0.08


def test_gemm_small_m_and_f32_out():
    ops = _ops()
    M, K, N = 16, 512, 1536
    g = torch.Generator(device=DEV).manual_seed(13)
    a = torch.randn(M, K, device=DEV, generator=g).to(torch.bfloat16)
    w = (torch.randn(N, K, device=DEV, generator=g) / math.sqrt(K)).to(torch.bfloat16)
    out = torch.empty(M, N, device=DEV, dtype=torch.float32)
    ops.igemm([(a, (1, 1, M), K, 1)], w, N, out)
    ref = a.float() @ w.float().t()
    assert rel_err(out, ref) < 1e-4, rel_err(out, ref)


@pytest.mark.parametrize("B,T,heads,hd", [
    (2, 1024, 8, 32), (1, 1024, 8, 16), (3, 256, 8, 48), (2, 256, 8, 32), (4, 64, 8, 64), (3, 64, 8, 48),
    (9, 16, 8, 64), (1, 16, 8, 64),
])
def test_qkv_attention(B, T, heads, hd):
    """QKV projection (V written transposed) followed by the fused attention kernel."""
    ops = _ops()
    Cc = heads * hd
    M = B * T
    g = torch.Generator(device=DEV).manual_seed(B * 100 + T + hd)
    x = torch.randn(M, Cc, device=DEV, generator=g).to(torch.bfloat16)
    wqkv = (torch.randn(3 * Cc, Cc, device=DEV, generator=g) * (1.5 / math.sqrt(Cc))).to(torch.bfloat16)
    bqkv = torch.randn(3 * Cc, device=DEV, generator=g) * 0.1
    qk = torch.empty(M, 2 * Cc, device=DEV, dtype=torch.bfloat16)
    vt = torch.empty(Cc, M, device=DEV, dtype=torch.bfloat16)
    ops.igemm([(x, (1, 1, M), Cc, 1)], wqkv, 3 * Cc, qk, bias=bqkv, vt=vt, vt_col0=2 * Cc)
    qkv_ref = x.float() @ wqkv.float().t() + bqkv
    assert rel_err(qk.float(), qkv_ref[:, :2 * Cc]) < 6e-3
    assert rel_err(vt.float().t(), qkv_ref[:, 2 * Cc:]) < 6e-3

    out = torch.empty(M, Cc, device=DEV, dtype=torch.bfloat16)
    ops.attention(qk, vt, out, M, T, heads, hd)
    q = qk[:, :Cc].float().reshape(B, T, heads, hd).transpose(1, 2)
    k = qk[:, Cc:].float().reshape(B, T, heads, hd).transpose(1, 2)
    v = vt.float().t().reshape(B, T, heads, hd).transpose(1, 2)
    ref = F.scaled_dot_product_attention(q, k, v).transpose(1, 2).reshape(M, Cc)
    err = rel_err(out.float(), ref)
    assert err < 1.5e-2, err


@pytest.mark.parametrize("B,C,H,silu", [(3, 128, 32, True), (2, 384, 16, True), (2, 768, 16, True),
                                        (5, 1024, 8, True), (4, 512, 4, False), (1, 256, 64, True)])
def test_groupnorm_silu(B, C, H, silu):
    ops = _ops()
    g = torch.Generator(device=DEV).manual_seed(C + H)
    x = torch.randn(B, C, H, H, device=DEV, generator=g) * 1.7 + 0.3
    gamma = 1 + 0.2 * torch.randn(C, device=DEV, generator=g)
    beta = 0.2 * torch.randn(C, device=DEV, generator=g)
    xr = rows(x)
    y = torch.empty_like(xr)
    ops.groupnorm_silu(xr, y, gamma, beta, B, H * H, C, 32, silu)
    ref = F.group_norm(bf(x), 32, gamma, beta, 1e-5)
    if silu:
        ref = F.silu(ref)
    got = unrows(y, B, H, H)
    assert (got - ref).abs().max().item() < 0.04
    assert rel_err(got, ref) < 5e-3


def test_embed_time_class():
    ops = _ops()
    D, P, R = 512, 4864, 7
    g = torch.Generator(device=DEV).manual_seed(3)
    factor = 10000 ** (torch.arange(0, D // 2, dtype=torch.float32, device=DEV) / (D // 2))
    w1 = torch.randn(4 * D, D, device=DEV, generator=g) / math.sqrt(D)
    b1 = torch.randn(4 * D, device=DEV, generator=g) * 0.1
    w2 = torch.randn(D, 4 * D, device=DEV, generator=g) / math.sqrt(4 * D)
    b2 = torch.randn(D, device=DEV, generator=g) * 0.1
    cw = torch.randn(3, D, device=DEV, generator=g)
    wp = torch.randn(P, D, device=DEV, generator=g) / math.sqrt(D)
    bp = torch.randn(P, device=DEV, generator=g) * 0.1
    t = torch.tensor([0, 1, 17, 500, 998, 999, 999], device=DEV)
    ctx = torch.tensor([0, 1, 2, 0, 1, 2, 0], device=DEV)
    mask = torch.tensor([1, 1, 0, 1, 0, 1, 1], device=DEV, dtype=torch.float32)
    out = torch.empty(R, P, device=DEV)
    scratch = torch.empty(R * 5 * D, device=DEV)
    ops.embed_time_class(t, ctx, mask, factor, w1, b1, w2, b2, cw, wp, bp, out, scratch)
    with torch.no_grad():
        e = t[:, None] / factor
        e = torch.cat([torch.sin(e), torch.cos(e)], -1)
        temb = F.linear(F.silu(F.linear(e, w1, b1)), w2, b2) + mask[:, None] * cw[ctx]
        ref = F.linear(F.silu(temb), wp, bp)
    assert (out - ref).abs().max().item() < 2e-3, (out - ref).abs().max().item()
    out2 = torch.empty(R, P, device=DEV)
    ops.embed_time_class(t, None, None, factor, w1, b1, w2, b2, cw, wp, bp, out2, scratch)
    temb = F.linear(F.silu(F.linear(e, w1, b1)), w2, b2)
    ref2 = F.linear(F.silu(temb), wp, bp)
    assert (out2 - ref2).abs().max().item() < 2e-3
    # many rows (the sampler's per-timestep table): the shared-memory tiled fp32 GEMM path, ragged row / column counts
    R = 1003
    t = torch.randint(0, 1000, (R,), device=DEV, generator=g)
    ctx = torch.randint(0, 3, (R,), device=DEV, generator=g)
    mask = (torch.rand(R, device=DEV, generator=g) > 0.3).float()
    Pr = 4840
    out3 = torch.empty(R, Pr, device=DEV)
    ops.embed_time_class(t, ctx, mask, factor, w1, b1, w2, b2, cw, wp[:Pr].contiguous(), bp[:Pr].contiguous(), out3,
                         torch.empty(R * 5 * D, device=DEV))
    with torch.no_grad():
        e = t[:, None] / factor
        e = torch.cat([torch.sin(e), torch.cos(e)], -1)
        temb = F.linear(F.silu(F.linear(e, w1, b1)), w2, b2) + mask[:, None] * cw[ctx]
        ref3 = F.linear(F.silu(temb), wp[:Pr], bp[:Pr])
    assert (out3 - ref3).abs().max().item() < 2e-3, (out3 - ref3).abs().max().item()


def test_small_channel_convs_and_movers():
    ops = _ops()
    g = torch.Generator(device=DEV).manual_seed(5)
    B, H = 3, 32
    x = torch.randn(B, 3, H, H, device=DEV, generator=g)
    w = torch.randn(128, 3, 3, 3, device=DEV, generator=g) / math.sqrt(27)
    b = torch.randn(128, device=DEV, generator=g)
    y = torch.empty(B * H * H, 128, device=DEV, dtype=torch.bfloat16)
    ops.conv3x3_small_cin(x, w, b, y)
    ref = F.conv2d(x, w, b, padding=1)
    assert rel_err(unrows(y, B, H, H), ref) < 4e-3

    hcl = torch.randn(B, 128, H, H, device=DEV, generator=g)
    w2 = torch.randn(3, 128, 3, 3, device=DEV, generator=g) / math.sqrt(9 * 128)
    b2 = torch.randn(3, device=DEV, generator=g)
    out = torch.empty(B, 3, H, H, device=DEV)
    ops.conv3x3_small_cout(rows(hcl), w2, b2, out)
    ref = F.conv2d(bf(hcl), w2, b2, padding=1)
    assert (out - ref).abs().max().item() < 1e-4

    w11 = torch.randn(3, 3, device=DEV, generator=g)
    b11 = torch.randn(3, device=DEV, generator=g)
    o11 = torch.empty_like(x)
    ops.conv1x1_small_f32(x, w11, b11, o11)
    assert (o11 - F.conv2d(x, w11[:, :, None, None], b11)).abs().max().item() < 1e-5

    xs = torch.randn(2, 256, 8, 8, device=DEV, generator=g)
    up = torch.empty(2 * 16 * 16, 256, device=DEV, dtype=torch.bfloat16)
    ops.upsample_nearest2x(rows(xs), up, 2, 8, 8, 256)
    assert torch.equal(unrows(up, 2, 16, 16), F.interpolate(bf(xs), scale_factor=2))

    back = torch.empty(2, 256, 8, 8, device=DEV)
    ops.rows_to_nchw(ops.nchw_to_rows(xs, torch.empty(2 * 64, 256, device=DEV, dtype=torch.bfloat16)), back)
    assert torch.equal(back, bf(xs))


@pytest.mark.parametrize("B,C,H", [(2, 256, 32), (3, 384, 16), (5, 512, 8)])
def test_downsample_planes_gemm(B, C, H):
    """Conv2d(C, C, 3, stride 2, pad 0) followed by ConstantPad2d((0,1,0,1)) on the OUTPUT."""
    ops = _ops()
    g = torch.Generator(device=DEV).manual_seed(C)
    x = torch.randn(B, C, H, H, device=DEV, generator=g)
    w = torch.randn(C, C, 3, 3, device=DEV, generator=g) / math.sqrt(9 * C)
    b = torch.randn(C, device=DEV, generator=g)
    OH = H // 2
    ref = F.pad(F.conv2d(bf(x), bf(w), b, stride=2), (0, 1, 0, 1))
    # parity planes + stride-2 tap addressing inside the GEMM (no im2col matrix)
    planes = torch.empty(4 * B * OH * OH, C, device=DEV, dtype=torch.bfloat16)
    ops.space_to_depth2(rows(x), planes, B, H, H, C)
    out2 = torch.empty(B * OH * OH, C, device=DEV, dtype=torch.bfloat16)
    ops.igemm([(planes, (4 * B, OH, OH), C, 9)], ops.pack_conv_weight(w), C, out2, bias=b, zero_pad_last=True,
              s2_batch=B)
    got2 = unrows(out2, B, OH, OH)
    assert rel_err(got2, ref) < 6e-3
    assert got2[:, :, -1, :].abs().max().item() == 0.0 and got2[:, :, :, -1].abs().max().item() == 0.0


def test_cfg_posterior_step():
    ops = _ops()
    from types import SimpleNamespace
    T = 1000
    betas = torch.linspace(1e-4 ** 0.5, 0.02 ** 0.5, T, device=DEV) ** 2
    alphas = 1 - betas
    acp = torch.cumprod(alphas, 0)
    sched = SimpleNamespace(num_steps=T, betas=betas, alphas=alphas, alpha_cum_prod=acp, sqrt_alpha_cum_prod=acp.sqrt(),
                            sqrt_one_minus_alpha_cum_prod=(1 - acp).sqrt())
    g = torch.Generator(device=DEV).manual_seed(9)
    N = 6
    xt, ec, eu, z = (torch.randn(N, 3, 32, 32, device=DEV, generator=g) for _ in range(4))
    cfg = torch.tensor([1, 3, 5, 7, 9, 2], device=DEV, dtype=torch.float32)
    for step in (999, 500, 1, 0):
        t = torch.full((N,), step, device=DEV, dtype=torch.long)
        out = torch.empty_like(xt)
        x0 = torch.empty_like(xt)
        ops.cfg_posterior_step(xt, ec, eu, z, cfg, t, sched, out, x0)
        eps = eu + cfg.long()[:, None, None, None] * (ec - eu)
        so, sa = sched.sqrt_one_minus_alpha_cum_prod[t].view(-1, 1, 1, 1), sched.sqrt_alpha_cum_prod[t].view(-1, 1, 1, 1)
        x0_ref = ((xt - so * eps) / sa).clamp(-1, 1)
        mean = (xt - betas[t].view(-1, 1, 1, 1) * eps / so) / alphas[t].view(-1, 1, 1, 1).sqrt()
        if step > 0:
            var = (1 - acp[t - 1]) / (1 - acp[t]) * betas[t]
            mean = mean + (var ** 0.5).view(-1, 1, 1, 1) * z
        assert (out - mean).abs().max().item() <= 1e-5 * mean.abs().max().item()
        assert (x0 - x0_ref).abs().max().item() <= 1e-5


@pytest.mark.parametrize("spread", ["default_init", "normal"])
def test_vq_argmin_bit_exact(spread):
    """Indices must equal torch.cdist(...).argmin on the same device, including near-ties."""
    ops = _ops()
    g = torch.Generator(device=DEV).manual_seed(21)
    B, size, dim = 16, 1024, 3
    z = torch.randn(B, 1024, dim, device=DEV, generator=g)
    if spread == "default_init":
        cb = (torch.rand(size, dim, device=DEV, generator=g) * 2 - 1) / size
    else:
        cb = torch.randn(size, dim, device=DEV, generator=g)
    ref = torch.cdist(z, cb[None].repeat(B, 1, 1)).argmin(dim=-1).view(-1)
    idx = torch.empty(B * 1024, device=DEV, dtype=torch.long)
    zq = torch.empty(B * 1024, dim, device=DEV)
    ops.vq_argmin(z.reshape(-1, dim).contiguous(), cb, idx, zq)
    mism = (idx != ref).sum().item()
    assert mism == 0, f"{mism} / {idx.numel()} indices differ"
    assert torch.equal(zq, cb[ref])


@pytest.mark.parametrize("B,C,H", [(3, 512, 4), (2, 384, 8), (5, 256, 16), (96, 256, 16)])
def test_upsample_conv_as_subpixel_convs(B, C, H):
    """Upsample (components.py:124-130) = nearest 2x + conv3x3, computed as four sub-pixel 2x2 convolutions on the
    low-resolution input that store straight into the (strided) left half of the concat buffer."""
    ops = _ops()
    g = torch.Generator(device=DEV).manual_seed(B + C + H)
    x = torch.randn(B, C, H, H, device=DEV, generator=g)
    w = torch.randn(C, C, 3, 3, device=DEV, generator=g) / math.sqrt(9 * C)
    b = torch.randn(C, device=DEV, generator=g)
    cat = torch.zeros(B * 4 * H * H, 2 * C, device=DEV, dtype=torch.bfloat16)
    ops.upsample_conv3x3(rows(x), (B, H, H), C, ops.pack_upsample_conv_weights(w), C, cat[:, :C], bias=b)
    ref = F.conv2d(F.interpolate(bf(x), scale_factor=2.0, mode="nearest"), w, b, padding=1)
    got = unrows(cat[:, :C], B, 2 * H, 2 * H)
    assert rel_err(got, ref) < 8e-3, rel_err(got, ref)
    assert cat[:, C:].abs().max().item() == 0.0  # the other half of the buffer is untouched
    # default: ONE launch for the four parities (out_up2 == 2); the four-launch form gives the same bits
    import os
    cat4 = torch.zeros_like(cat)
    os.environ["IDF_UP2_FUSED"] = "0"
    try:
        ops.upsample_conv3x3(rows(x), (B, H, H), C, ops.pack_upsample_conv_weights(w), C, cat4[:, :C], bias=b)
    finally:
        del os.environ["IDF_UP2_FUSED"]
    assert torch.equal(cat, cat4)


@pytest.mark.parametrize("B,C,H", [(3, 256, 32), (5, 384, 16), (7, 512, 8), (96, 512, 8)])
def test_downsample_conv_strided_tma(B, C, H):
    """Downsample (components.py:106-117) read from the full-resolution input through a TMA map with element strides
    (2, 2); the input may be a column slice of the concat buffer."""
    ops = _ops()
    g = torch.Generator(device=DEV).manual_seed(B + C + H)
    x = torch.randn(B, C, H, H, device=DEV, generator=g)
    w = torch.randn(C, C, 3, 3, device=DEV, generator=g) / math.sqrt(9 * C)
    b = torch.randn(C, device=DEV, generator=g)
    cat = torch.zeros(B * H * H, 2 * C, device=DEV, dtype=torch.bfloat16)
    cat[:, C:] = rows(x)
    out = torch.empty(B * H * H // 4, C, device=DEV, dtype=torch.bfloat16)
    ops.igemm([(cat[:, C:], (B, H, H), C, 9)], ops.pack_conv_weight(w), C, out, bias=b, zero_pad_last=True, s2_direct=True)
    ref = F.pad(F.conv2d(bf(x), bf(w), b, stride=2), (0, 1, 0, 1))
    got = unrows(out, B, H // 2, H // 2)
    assert rel_err(got, ref) < 6e-3, rel_err(got, ref)
    assert got[:, :, -1, :].abs().max().item() == 0.0 and got[:, :, :, -1].abs().max().item() == 0.0


@pytest.mark.parametrize("B,T,hd", [(3, 1024, 384), (5, 256, 128), (70, 256, 128)])
def test_igemm_batched_second_operand_is_a_batch_of_gemms(B, T, hd):
    """w_batch_row / w_batch_col: image i of the A operand multiplies its own block of the second operand. The two
    shapes of the VAE's single wide attention head (components.py:87-95): S_i = Q_i K_i^T (fp32 out) and O_i = P_i V_i
    with V stored transposed (hd, B*T)."""
    ops = _ops()
    g = torch.Generator(device=DEV).manual_seed(B + T)
    M = B * T
    qk = torch.randn(M, 2 * hd, device=DEV, generator=g).to(torch.bfloat16)
    vt = torch.randn(hd, M, device=DEV, generator=g).to(torch.bfloat16)
    q, k = qk[:, :hd], qk[:, hd:]
    s = torch.empty(M, T, device=DEV, dtype=torch.float32)
    grid = (B, T // 128, 128)
    ops.igemm([(q, grid, hd, 1)], k, T, s, w_batch=(T, 0))
    ref = torch.bmm(q.float().view(B, T, hd), k.float().view(B, T, hd).transpose(1, 2)).view(M, T)
    assert rel_err(s, ref) < 1e-5
    pm = torch.empty(M, T, device=DEV, dtype=torch.bfloat16)
    ops.softmax_rows(s, pm, 1.0 / math.sqrt(hd))
    pref = torch.softmax(ref / math.sqrt(hd), dim=-1)
    assert rel_err(pm.float(), pref) < 6e-3
    o = torch.empty(M, 2 * hd, device=DEV, dtype=torch.bfloat16)
    ops.igemm([(pm, grid, T, 1)], vt, hd, o[:, hd:], w_batch=(0, T))
    v = vt.float().t().reshape(B, T, hd)
    oref = torch.bmm(pm.float().view(B, T, T), v).view(M, hd)
    assert rel_err(o[:, hd:].float(), oref) < 6e-3


@pytest.mark.parametrize("B,C,H,silu", [(3, 128, 128, True), (2, 256, 64, True), (5, 384, 64, False), (2, 256, 128, True),
                                        (7, 128, 20, True)])
def test_groupnorm_rows_large_images(B, C, H, silu):
    """Whole-row GroupNorm kernels (VAE stages): against F.group_norm, and bit-identical for any L2 grouping of the
    samples (the statistics never mix samples and are summed in a fixed order)."""
    ops = _ops()
    g = torch.Generator(device=DEV).manual_seed(C + H)
    x = torch.randn(B, C, H, H, device=DEV, generator=g) * 1.7 + 0.3
    gamma = 1 + 0.2 * torch.randn(C, device=DEV, generator=g)
    beta = 0.2 * torch.randn(C, device=DEV, generator=g)
    xr = rows(x)
    ws = torch.empty(B * ((H * H + 255) // 256) * 32 * 2, device=DEV)
    y = torch.empty_like(xr)
    ops.groupnorm_silu_rows(xr, y, gamma, beta, B, H * H, C, 32, silu, ws)
    ref = F.group_norm(bf(x), 32, gamma, beta, 1e-5)
    if silu:
        ref = F.silu(ref)
    assert rel_err(unrows(y, B, H, H), ref) < 6e-3
    y1 = torch.empty_like(xr)
    ops.call("idf_groupnorm_silu_rows", xr.data_ptr(), xr.stride(0), y1.data_ptr(), y1.stride(0), gamma.data_ptr(),
             beta.data_ptr(), B, H * H, C, 32, 1e-5, 1 if silu else 0, ws.data_ptr(), ws.numel() * 4, 1)  # one sample per pair
    assert torch.equal(y, y1)
    y2 = torch.empty_like(xr)
    ops.groupnorm_silu(xr, y2, gamma, beta, B, H * H, C, 32, silu)  # the slab kernel: same maths, other summation order
    assert rel_err(y2.float(), y.float()) < 4e-3


@pytest.mark.parametrize("B,Cin,Cout,H", [(3, 128, 3, 32), (2, 384, 6, 32), (1, 128, 3, 128), (5, 128, 3, 8), (96, 128, 3, 32)])
def test_igemm_narrow_tile_writes_fp32_nchw(B, Cin, Cout, H):
    """The network's last convolutions (128 -> 3, 384 -> z) on one 16-wide tcgen05 tile, fp32 NCHW planes out."""
    ops = _ops()
    g = torch.Generator(device=DEV).manual_seed(Cin + H + B)
    x = torch.randn(B, Cin, H, H, device=DEV, generator=g)
    w = torch.randn(Cout, Cin, 3, 3, device=DEV, generator=g) / math.sqrt(9 * Cin)
    b = torch.randn(Cout, device=DEV, generator=g)
    wp = torch.zeros(16, 9 * Cin, device=DEV, dtype=torch.bfloat16)
    wp[:Cout] = ops.pack_conv_weight(w)
    bp = torch.zeros(16, device=DEV)
    bp[:Cout] = b
    out = torch.full((B, Cout, H, H), float("nan"), device=DEV)
    ops.igemm([(rows(x), (B, H, H), Cin, 9)], wp, 16, None, bias=bp, out_nchw=out)
    ref = F.conv2d(bf(x), bf(w), b, padding=1)
    assert torch.isfinite(out).all()
    assert rel_err(out, ref) < 2e-5 * 100, rel_err(out, ref)   # fp32 accumulation order only (inputs are bf16-exact)
    assert rel_err(out, ref) < 1e-4


def test_vq_loss_perplexity_matches_torch():
    """Eval-mode tail of Codebook.forward (components.py:301-313) in one fused kernel pair vs the torch expressions."""
    ops = _ops()
    g = torch.Generator(device=DEV).manual_seed(21)
    B, D, H, size = 7, 3, 32, 1024
    z = torch.randn(B, D, H, H, device=DEV, generator=g)
    cb = torch.randn(size, D, device=DEV, generator=g) * 0.7
    idx = torch.empty(B * H * H, device=DEV, dtype=torch.int64)
    zq = torch.empty_like(z)
    ops.vq_argmin(z, cb, idx, zq)
    out, loss, perp = ops.vq_loss_perplexity(z, zq, idx, size, 0.25)
    assert torch.equal(out, z + (zq - z))
    ref_loss = 0.25 * torch.mean((zq - z) ** 2)
    probs = torch.bincount(idx, minlength=size).float() / idx.numel()
    ref_perp = torch.exp(-torch.sum(probs * torch.log(probs + 1e-6)))
    assert abs(loss.item() - ref_loss.item()) <= 1e-6 * max(1.0, abs(ref_loss.item()))
    assert abs(perp.item() - ref_perp.item()) <= 1e-4 * ref_perp.item()


def test_kl_loss_reparam_matches_torch():
    """KL bottleneck of VAE.encode (vae.py:99-113): clamp, per-sample KL sum, batch mean, reparametrised sample."""
    ops = _ops()
    g = torch.Generator(device=DEV).manual_seed(22)
    B, zd, H = 5, 3, 32
    z6 = torch.randn(B, 2 * zd, H, H, device=DEV, generator=g)
    z6[0, zd, 0, 0] = 35.0      # above the clamp
    z6[1, zd + 1, 3, 3] = -50.0  # below it
    noise = torch.randn(B, zd, H, H, device=DEV, generator=g)
    mean, lv = torch.chunk(z6, 2, dim=1)
    lv = torch.clamp(lv, -30.0, 20.0)
    ref_kl = (-0.5 * torch.sum(1 + lv - mean.pow(2) - lv.exp(), dim=[1, 2, 3])).mean()
    ref_z = mean + noise * torch.exp(0.5 * lv)
    kl, z = ops.kl_loss_reparam(z6, noise)
    assert abs(kl.item() - ref_kl.item()) <= 2e-6 * abs(ref_kl.item())
    assert (z - ref_z).abs().max().item() <= 2e-6 * ref_z.abs().max().item()
    kl2, none = ops.kl_loss_reparam(z6)
    assert none is None and kl2.item() == kl.item()
