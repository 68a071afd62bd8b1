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
