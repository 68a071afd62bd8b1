"""Kernel-level parity of the backward (training-step) kernels on B200: every C-ABI kernel against torch autograd in
fp32 on the same (bf16-rounded) inputs. Tolerances are a few bf16 ulps of the output scale; fp32 outputs (weight
gradients) are limited by the bf16 rounding of their inputs only."""
import math

import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu

torch.backends.cudnn.allow_tf32 = False
torch.backends.cuda.matmul.allow_tf32 = False

DEV = "cuda"


@pytest.fixture(autouse=True)
def _grad_enabled():
    """Other test modules switch autograd off globally; the torch references here need it."""
    with torch.enable_grad():
        yield


def _ops():
    from idf_b200 import ops
    return ops


def rows(x_nchw):
    B, C, H, W = x_nchw.shape
    return x_nchw.permute(0, 2, 3, 1).reshape(B * H * W, C).to(torch.bfloat16).contiguous()


def unrows(y, B, H, W):
    return y.float().reshape(B, H, W, -1).permute(0, 3, 1, 2).contiguous()


def rel_err(a, b):
    return ((a - b).norm() / b.norm().clamp_min(1e-12)).item()


def bf(x):
    return x.to(torch.bfloat16).float()


@pytest.mark.parametrize("B,Cin,Cout,H,taps", [
    (2, 128, 128, 32, 9), (3, 256, 384, 16, 9), (5, 384, 512, 8, 9), (9, 512, 512, 4, 9), (3, 1024, 384, 8, 9),
    (2, 128, 256, 32, 1), (3, 768, 256, 16, 1), (48, 256, 256, 32, 9),
])
def test_conv_wgrad(B, Cin, Cout, H, taps):
    ops = _ops()
    g = torch.Generator(device=DEV).manual_seed(B * 1000 + Cin + Cout + H)
    x = bf(torch.randn(B, Cin, H, H, device=DEV, generator=g))
    dy = bf(torch.randn(B, Cout, H, H, device=DEV, generator=g))
    k = 3 if taps == 9 else 1
    w = torch.zeros(Cout, Cin, k, k, device=DEV, requires_grad=True)
    F.conv2d(x, w, padding=k // 2).backward(dy)
    ws = torch.empty(64 * 1024 * 1024 // 4, device=DEV, dtype=torch.float32)
    grads = []
    for rep in range(2):
        grad = torch.full((Cout, Cin, k, k), 7.0, device=DEV)
        ops.conv_wgrad(rows(x), (B, H, H), Cin, taps, rows(dy), Cout, grad, ws)
        grads.append(grad)
    assert torch.equal(grads[0], grads[1])  # deterministic
    assert rel_err(grads[0], w.grad) < 2e-5, rel_err(grads[0], w.grad)
    ops.conv_wgrad(rows(x), (B, H, H), Cin, taps, rows(dy), Cout, grads[0], ws, accumulate=True)
    assert rel_err(grads[0], 2 * w.grad) < 2e-5


def test_linear_wgrad_matrix_view():
    """nn.Linear over tokens (QKV / out_proj): x is a plain (M, K) matrix, dy a column slice of a wider buffer."""
    ops = _ops()
    M, K, N = 5 * 256, 384, 3 * 384
    g = torch.Generator(device=DEV).manual_seed(5)
    x = torch.randn(M, K, device=DEV, generator=g).to(torch.bfloat16)
    wide = torch.randn(M, N + 128, device=DEV, generator=g).to(torch.bfloat16)
    dy = wide[:, 128:]
    ws = torch.empty(16 * 1024 * 1024, device=DEV, dtype=torch.float32)
    grad = torch.empty(N, K, device=DEV)
    ops.conv_wgrad(x, (1, 1, M), K, 1, dy, N, grad, ws)
    ref = dy.float().t() @ x.float()
    assert rel_err(grad, ref) < 2e-5, rel_err(grad, ref)


@pytest.mark.parametrize("B,C,H", [(3, 256, 32), (5, 384, 16), (7, 512, 8)])
def test_downsample_conv_backward(B, C, H):
    """Downsample (components.py:106-117): conv 3x3 stride 2 pad 0 then zero pad on the output. Weight gradient via
    the parity-plane tap addressing, data gradient via four per-plane igemm launches with explicit tap lists."""
    ops = _ops()
    g = torch.Generator(device=DEV).manual_seed(B + C + H)
    x = bf(torch.randn(B, C, H, H, device=DEV, generator=g)).requires_grad_(True)
    w = bf(torch.randn(C, C, 3, 3, device=DEV, generator=g) / math.sqrt(9 * C)).requires_grad_(True)
    OH = H // 2
    dy = bf(torch.randn(B, C, OH, OH, device=DEV, generator=g))
    y = F.pad(F.conv2d(x, w, stride=2), (0, 1, 0, 1))
    y.backward(dy)
    dym = dy.clone()
    dym[:, :, -1, :] = 0
    dym[:, :, :, -1] = 0
    planes = torch.empty(B * H * H, C, device=DEV, dtype=torch.bfloat16)
    ops.space_to_depth2(rows(x.detach()), planes, B, H, H, C)
    ws = torch.empty(16 * 1024 * 1024, device=DEV, dtype=torch.float32)
    grad = torch.empty(C, C, 3, 3, device=DEV)
    ops.conv_wgrad(planes, (4 * B, OH, OH), C, 9, rows(dym), C, grad, ws, s2_batch=B)
    assert rel_err(grad, w.grad) < 2e-5, rel_err(grad, w.grad)
    dplanes = torch.empty(B * H * H, C, device=DEV, dtype=torch.bfloat16)
    ops.conv_s2_dgrad(rows(dym), B, OH, OH, C, ops.pack_s2_dgrad_weights(w), dplanes, C)
    ref_planes = torch.empty_like(dplanes)
    ops.space_to_depth2(rows(x.grad), ref_planes, B, H, H, C)
    assert rel_err(dplanes.float(), ref_planes.float()) < 6e-3, rel_err(dplanes.float(), ref_planes.float())


@pytest.mark.parametrize("B,Cin,Cout,H", [(2, 128, 256, 32), (3, 1024, 384, 8)])
def test_conv3x3_dgrad(B, Cin, Cout, H):
    ops = _ops()
    g = torch.Generator(device=DEV).manual_seed(B + Cin)
    x = torch.zeros(B, Cin, H, H, device=DEV, requires_grad=True)
    w = bf(torch.randn(Cout, Cin, 3, 3, device=DEV, generator=g) / math.sqrt(9 * Cout))
    dy = bf(torch.randn(B, Cout, H, H, device=DEV, generator=g))
    F.conv2d(x, w, padding=1).backward(dy)
    dx = torch.empty(B * H * H, Cin, device=DEV, dtype=torch.bfloat16)
    ops.igemm([(rows(dy), (B, H, H), Cout, 9)], ops.pack_dgrad_weight(w), Cin, dx)
    assert rel_err(unrows(dx, B, H, H), x.grad) < 6e-3
    # same result from the FORWARD-packed weight read as an MN-major operand (what the training engine does)
    dx2 = torch.empty_like(dx)
    taps = [(1 - kh, 1 - kw) for kh in range(3) for kw in range(3)]
    ops.igemm([(rows(dy), (B, H, H), Cout, 9)], ops.pack_conv_weight(w), Cin, dx2, tap_offsets=taps, w_mn=True,
              w_tap_ids=range(9))
    assert rel_err(unrows(dx2, B, H, H), x.grad) < 6e-3
    assert rel_err(dx2.float(), dx.float()) < 2e-3


def test_linear_dgrad_from_forward_weight():
    """dX = dY W for nn.Linear (W stored (out, in) as in the forward pass), column-sliced weight, residual epilogue."""
    ops = _ops()
    M, O, I = 3 * 256, 768, 256
    g = torch.Generator(device=DEV).manual_seed(2)
    dy = torch.randn(M, O, device=DEV, generator=g).to(torch.bfloat16)
    wide = (torch.randn(O, I + 128, device=DEV, generator=g) / math.sqrt(O)).to(torch.bfloat16)
    w = wide[:, 128:]
    res = torch.randn(M, I, device=DEV, generator=g).to(torch.bfloat16)
    dx = torch.empty(M, I, device=DEV, dtype=torch.bfloat16)
    ops.igemm([(dy, (1, 1, M), O, 1)], w, I, dx, res=res, w_mn=True)
    ref = dy.float() @ w.float() + res.float()
    assert rel_err(dx.float(), ref) < 6e-3


@pytest.mark.parametrize("B,C,HW,silu,with_add", [(3, 128, 1024, True, False), (2, 384, 256, True, True),
                                                  (5, 512, 64, False, True), (4, 1024, 64, True, False),
                                                  (3, 512, 16, True, True), (2, 768, 256, True, False)])
def test_groupnorm_silu_backward(B, C, HW, silu, with_add):
    ops = _ops()
    G = 32
    g = torch.Generator(device=DEV).manual_seed(C + HW)
    x = bf(torch.randn(B, HW, C, device=DEV, generator=g) * 1.5 + 0.3).requires_grad_(True)
    gamma = (1 + 0.2 * torch.randn(C, device=DEV, generator=g)).requires_grad_(True)
    beta = (0.2 * torch.randn(C, device=DEV, generator=g)).requires_grad_(True)
    dy = bf(torch.randn(B, HW, C, device=DEV, generator=g))
    add = bf(torch.randn(B, HW, C, device=DEV, generator=g)) if with_add else None
    y = F.group_norm(x.transpose(1, 2), G, gamma, beta, 1e-5).transpose(1, 2)
    if silu:
        y = F.silu(y)
    y.backward(dy)
    xr = x.detach().reshape(B * HW, C).to(torch.bfloat16)
    yk = torch.empty_like(xr)
    stats = torch.empty(B, G, 2, device=DEV)
    ops.groupnorm_silu_train(xr, yk, gamma.detach(), beta.detach(), B, HW, C, G, silu, stats)
    assert rel_err(yk.float(), y.detach().reshape(B * HW, C)) < 6e-3
    dx = torch.empty_like(xr)
    dgp, dbp = torch.empty(B, C, device=DEV), torch.empty(B, C, device=DEV)
    cs = torch.empty(B, C + 64, device=DEV)[:, 64:]  # strided per-sample column sums
    ops.groupnorm_silu_bwd(xr, dy.reshape(B * HW, C).to(torch.bfloat16), dx, gamma.detach(), beta.detach(), stats, dgp,
                           dbp, B, HW, C, G, silu, add=None if add is None else add.reshape(B * HW, C).to(torch.bfloat16),
                           colsum_part=cs)
    ref_dx = x.grad.reshape(B * HW, C) + (add.reshape(B * HW, C) if with_add else 0)
    assert rel_err(dx.float(), ref_dx) < 8e-3, rel_err(dx.float(), ref_dx)
    dg, db, b1, b2 = (torch.empty(C, device=DEV) for _ in range(4))
    ops.groupnorm_bwd_finalize(dgp, dbp, B, C, dg, db, colsum_part=cs, g_bias1=b1, g_bias2=b2)
    assert rel_err(dg, gamma.grad) < 2e-3, rel_err(dg, gamma.grad)
    assert rel_err(db, beta.grad) < 2e-3, rel_err(db, beta.grad)
    ref_cs = ref_dx.reshape(B, HW, C).sum(1)
    assert rel_err(cs, ref_cs) < 5e-3 and torch.equal(b1, b2) and rel_err(b1, ref_cs.sum(0)) < 5e-3


def test_colsum_and_movers():
    ops = _ops()
    B, H, C = 5, 16, 384
    g = torch.Generator(device=DEV).manual_seed(3)
    wide = torch.randn(B * H * H, C + 64, device=DEV, generator=g).to(torch.bfloat16)
    x = wide[:, 64:]
    ps = torch.empty(B, 2 * C, device=DEV)
    tot = torch.zeros(C, device=DEV)
    ops.colsum(x, B, H * H, C, ps[:, C:], total=tot)
    ref = x.float().reshape(B, H * H, C).sum(1)
    assert rel_err(ps[:, C:], ref) < 1e-5
    assert rel_err(tot, ref.sum(0)) < 1e-5
    # nearest-2x adjoint
    y = torch.empty(B * H * H // 4, C, device=DEV, dtype=torch.bfloat16)
    ops.sum2x2(x, y, B, H // 2, H // 2, C)
    refy = F.avg_pool2d(unrows(x, B, H, H), 2) * 4
    assert rel_err(unrows(y, B, H // 2, H // 2), refy) < 4e-3
    # space_to_depth2 round trip (+ addend)
    planes = torch.empty(B * H * H, C, device=DEV, dtype=torch.bfloat16)
    xc = x.contiguous()
    ops.space_to_depth2(xc, planes, B, H, H, C)
    back = torch.empty_like(xc)
    ops.depth_to_space2(planes, back, B, H, H, C)
    assert torch.equal(back, xc)
    ops.depth_to_space2(planes, back, B, H, H, C, add=xc)
    assert rel_err(back.float(), 2 * xc.float()) < 4e-3
    z = xc.clone()
    ops.zero_last_rowcol(z, B, H, H, C)
    zr = unrows(xc, B, H, H)
    zr[:, :, -1, :] = 0
    zr[:, :, :, -1] = 0
    assert torch.equal(unrows(z, B, H, H), zr)


@pytest.mark.parametrize("B,T,heads,hd", [(3, 1024, 8, 32), (2, 1024, 8, 16), (5, 256, 8, 48), (6, 64, 8, 64),
                                          (9, 16, 8, 64), (4, 256, 8, 32), (2, 128, 4, 64)])
def test_attention_backward(B, T, heads, hd):
    ops = _ops()
    C, M = heads * hd, B * T
    g = torch.Generator(device=DEV).manual_seed(T + hd)
    q = bf(torch.randn(B, heads, T, hd, device=DEV, generator=g)).requires_grad_(True)
    k = bf(torch.randn(B, heads, T, hd, device=DEV, generator=g)).requires_grad_(True)
    v = bf(torch.randn(B, heads, T, hd, device=DEV, generator=g)).requires_grad_(True)
    do = bf(torch.randn(B, heads, T, hd, device=DEV, generator=g))
    o_ref = torch.softmax(q @ k.transpose(-1, -2) / math.sqrt(hd), -1) @ v
    o_ref.backward(do)
    tok = lambda t: t.detach().transpose(1, 2).reshape(M, C)  # (B, heads, T, hd) -> (M, C) head-major channels
    qkv = torch.cat([tok(q), tok(k), tok(v)], dim=1).to(torch.bfloat16).contiguous()
    o = torch.empty(M, C, device=DEV, dtype=torch.bfloat16)
    lse = torch.empty(M, heads, device=DEV)
    ops.attention_qkv(qkv, o, M, T, heads, hd, lse=lse)
    assert rel_err(o.float(), tok(o_ref)) < 8e-3
    # the V^T-operand variant of the forward kernel gives the same result
    o2 = torch.empty_like(o)
    ops.attention(qkv[:, :2 * C], qkv[:, 2 * C:].t().contiguous(), o2, M, T, heads, hd)
    assert rel_err(o2.float(), o.float()) < 2e-3
    s = (q @ k.transpose(-1, -2) / math.sqrt(hd)).detach()
    lse_ref = (torch.logsumexp(s, -1) * 1.4426950408889634).transpose(1, 2).reshape(M, heads)
    assert (lse - lse_ref).abs().max().item() < 2e-2
    dqkv = torch.empty(M, 3 * C, device=DEV, dtype=torch.bfloat16)
    delta = torch.empty(M, heads, device=DEV)
    dq32 = torch.empty(M, C, device=DEV)
    ops.attention_bwd(qkv, o, tok(do).to(torch.bfloat16).contiguous(), lse, delta, dqkv, dq32, M, T, heads, hd)
    for name, got, ref in (("dq", dqkv[:, :C], tok(q.grad)), ("dk", dqkv[:, C:2 * C], tok(k.grad)),
                           ("dv", dqkv[:, 2 * C:], tok(v.grad))):
        e = rel_err(got.float(), ref)
        assert e < 2e-2, (name, e)


def test_edge_conv_backward():
    ops = _ops()
    B, H, C = 3, 32, 128
    g = torch.Generator(device=DEV).manual_seed(9)
    # in_conv weight gradient
    x = torch.randn(B, 3, H, H, device=DEV, generator=g)
    w = torch.zeros(C, 3, 3, 3, device=DEV, requires_grad=True)
    dy = bf(torch.randn(B, C, H, H, device=DEV, generator=g))
    F.conv2d(x, w, padding=1).backward(dy)
    gw = torch.empty(C, 3, 3, 3, device=DEV)
    part = torch.empty(B * (H // 2) * C * 27 * 4, device=DEV)
    ops.conv3x3_small_cin_wgrad(x, rows(dy), gw, part)
    assert rel_err(gw, w.grad) < 1e-4, rel_err(gw, w.grad)
    # out_conv backward
    h = bf(torch.randn(B, C, H, H, device=DEV, generator=g)).requires_grad_(True)
    wo = (torch.randn(3, C, 3, 3, device=DEV, generator=g) / 30).requires_grad_(True)
    bo = torch.zeros(3, device=DEV, requires_grad=True)
    dout = torch.randn(B, 3, H, H, device=DEV, generator=g)
    F.conv2d(h, wo, bo, padding=1).backward(dout)
    dh = torch.empty(B * H * H, C, device=DEV, dtype=torch.bfloat16)
    gwo, gbo = torch.empty(3, C, 3, 3, device=DEV), torch.empty(3, device=DEV)
    ops.conv3x3_small_cout_bwd(rows(h.detach()), dout, wo.detach(), dh, gwo, gbo, part)
    assert rel_err(unrows(dh, B, H, H), h.grad) < 6e-3
    assert rel_err(gwo, wo.grad) < 1e-4
    assert rel_err(gbo, bo.grad) < 1e-4


def test_embedding_backward():
    ops = _ops()
    R, D, P, NC = 7, 128, 640, 3
    g = torch.Generator(device=DEV).manual_seed(21)
    rnd = lambda *s, sc=1.0: (torch.randn(*s, device=DEV, generator=g) * sc)
    factor = 10000 ** (torch.arange(D // 2, device=DEV) / (D // 2))
    w1, b1 = rnd(4 * D, D, sc=D ** -0.5).requires_grad_(True), rnd(4 * D, sc=0.1).requires_grad_(True)
    w2, b2 = rnd(D, 4 * D, sc=(4 * D) ** -0.5).requires_grad_(True), rnd(D, sc=0.1).requires_grad_(True)
    cls = rnd(NC, D).requires_grad_(True)
    wp, bp = rnd(P, D, sc=D ** -0.5).requires_grad_(True), rnd(P, sc=0.1).requires_grad_(True)
    t = torch.tensor([0, 5, 999, 250, 31, 700, 1], device=DEV)
    ctx = torch.tensor([0, 1, 2, 2, 1, 0, 0], device=DEV)
    mask = torch.tensor([1., 0, 1, 1, 0, 1, 1], device=DEV)
    a = t[:, None] / factor
    e = torch.cat([torch.sin(a), torch.cos(a)], -1)
    temb = F.linear(F.silu(F.linear(e, w1, b1)), w2, b2) + (F.one_hot(ctx, NC).float() @ cls) * mask[:, None]
    table_ref = F.linear(F.silu(temb), wp, bp)
    dtable = rnd(R, P)
    table_ref.backward(dtable)
    table = torch.empty(R, P, device=DEV)
    saved = torch.empty(R * 11 * D, device=DEV)
    d = lambda x: x.detach().contiguous()
    ops.embed_time_class_train(t, ctx, mask, factor.float().contiguous(), d(w1), d(b1), d(w2), d(b2), d(cls), d(wp),
                               d(bp), table, saved)
    assert rel_err(table, table_ref.detach()) < 1e-5
    gs = {n: torch.empty_like(p) for n, p in (("w1", w1), ("b1", b1), ("w2", w2), ("b2", b2), ("cls", cls),
                                               ("wp", wp), ("bp", bp))}
    scratch = torch.empty(((P + 255) // 256) * R * 4 * D + 5 * R * D, device=DEV)
    ops.embed_time_class_bwd(dtable, ctx, mask, D, NC, d(w2), d(wp), saved, gs["w1"], gs["b1"], gs["w2"], gs["b2"],
                             gs["cls"], gs["wp"], gs["bp"], scratch)
    for n, p in (("w1", w1), ("b1", b1), ("w2", w2), ("b2", b2), ("cls", cls), ("wp", wp), ("bp", bp)):
        assert rel_err(gs[n], p.grad) < 1e-4, (n, rel_err(gs[n], p.grad))


def test_loss_clip_adam():
    ops = _ops()
    g = torch.Generator(device=DEV).manual_seed(2)
    pred, tgt = torch.randn(48, 3, 32, 32, device=DEV, generator=g), torch.randn(48, 3, 32, 32, device=DEV, generator=g)
    dp, loss = torch.empty_like(pred), torch.empty(1, device=DEV)
    ops.mse_loss_grad(pred, tgt, dp, loss)
    pr = pred.clone().requires_grad_(True)
    lr_ = F.mse_loss(pr, tgt)
    lr_.backward()
    assert abs(loss.item() - lr_.item()) < 1e-6 * abs(lr_.item()) + 1e-7
    assert rel_err(dp, pr.grad) < 1e-6
    n = 1_000_003
    p0 = torch.randn(n, device=DEV, generator=g)
    grad = torch.randn(n, device=DEV, generator=g) * 0.01
    ref_p = torch.nn.Parameter(p0.clone())
    opt = torch.optim.Adam([ref_p], lr=1e-3)
    p, m, v = p0.clone(), torch.zeros(n, device=DEV), torch.zeros(n, device=DEV)
    out2, scratch = torch.empty(2, device=DEV), torch.empty(2048, device=DEV)
    for step in (1, 2, 3):
        ref_p.grad = grad.clone() * step
        tn = torch.nn.utils.clip_grad_norm_([ref_p], 1.0)
        opt.step()
        gk = grad * step
        ops.grad_norm_clip(gk, out2, scratch, 1.0)
        assert abs(out2[0].item() - tn.item()) < 1e-4 * tn.item()
        hyper = torch.tensor(ops.adam_hyper(1e-3, step), device=DEV)
        ops.adam_step(p, gk, m, v, hyper, clip2=out2)
        assert (p - ref_p.detach()).abs().max().item() < 2e-6
