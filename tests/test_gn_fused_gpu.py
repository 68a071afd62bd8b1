"""GroupNorm fused into the implicit-GEMM epilogue (idf_igemm_args.gn_mode) on the B200: conv3x3 (+ bias + time bias)
-> GroupNorm (+ SiLU) against conv2d + F.group_norm in fp32 on the same bf16-rounded inputs. Covers single tiles and CTA
pairs, all three tile widths, 8x8 images (two per tile), groups that straddle tiles (384 channels: 12 per group), both
output modes, workspace reuse between launches and batch invariance of the statistics (components.py:448-460)."""
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


def bf(x):
    return x.to(torch.bfloat16).float()


def rel_err(a, b):
    return ((a - b).norm() / b.norm().clamp_min(1e-12)).item()


def run_case(B, Cin, Cout, H, silu, dual, seed=0, skip_c=0):
    ops = _ops()
    g = torch.Generator(device=DEV).manual_seed(seed + B * 1000 + Cin + Cout + H)
    x = torch.randn(B, Cin, H, H, device=DEV, generator=g)
    w = torch.randn(Cout, Cin, 3, 3, device=DEV, generator=g) / math.sqrt(9 * Cin)
    b = torch.randn(Cout, device=DEV, generator=g)
    table = torch.randn(B, Cout, device=DEV, generator=g)
    gamma = 1.0 + 0.3 * torch.randn(Cout, device=DEV, generator=g)
    beta = 0.3 * torch.randn(Cout, device=DEV, generator=g)
    segs = [(rows(x), (B, H, H), Cin, 9)]
    wp = ops.pack_conv_weight(w)
    ref_raw = F.conv2d(bf(x), bf(w), b, padding=1) + table[:, :, None, None]
    if skip_c:  # second segment: the block's 1x1 skip projection (two-segment GEMM)
        xs = torch.randn(B, skip_c, H, H, device=DEV, generator=g)
        w1 = torch.randn(Cout, skip_c, 1, 1, device=DEV, generator=g) / math.sqrt(skip_c)
        segs.append((rows(xs), (B, H, H), skip_c, 1))
        wp = torch.cat([wp, ops.pack_conv_weight(w1)], dim=1).contiguous()
        ref_raw = ref_raw + F.conv2d(bf(xs), bf(w1))
    ws = torch.zeros(ops.gn_workspace_bytes(B, B * H * H, Cout), device=DEV, dtype=torch.uint8)
    out = torch.empty(B * H * H, Cout, device=DEV, dtype=torch.bfloat16)
    gn = dict(gamma=gamma, beta=beta, groups=32, silu=silu, ws=ws)
    nrm = out
    if dual:
        nrm = torch.empty_like(out)
        gn["out"] = nrm
    ops.igemm(segs, wp, Cout, out, bias=b, rowbias=table, gn=gn)
    ref = F.group_norm(ref_raw, 32, gamma, beta, eps=1e-5)
    if silu:
        ref = F.silu(ref)
    got = unrows(nrm, B, H, H)
    assert torch.isfinite(got).all()
    e = rel_err(got, ref)
    assert e < 6e-3, e
    assert (got - ref).abs().max().item() < 0.06
    if dual:
        assert rel_err(unrows(out, B, H, H), ref_raw) < 6e-3
    # workspace header after one launch: epoch 1, no CTA still counted as running; a second launch on the same
    # workspace (stale records of the first one in every slot) gives the same bits
    hdr = ws[:8].view(torch.int32).tolist()
    assert hdr == [1, 0], hdr
    again = torch.empty_like(nrm)
    gn2 = dict(gn)
    if dual:
        gn2["out"] = again
        ops.igemm(segs, wp, Cout, torch.empty_like(out), bias=b, rowbias=table, gn=gn2)
    else:
        ops.igemm(segs, wp, Cout, again, bias=b, rowbias=table, gn=gn2)
    assert torch.equal(again, nrm)
    return nrm


@pytest.mark.parametrize("B,Cin,Cout,H,silu,dual", [
    (3, 128, 256, 32, True, False),     # 24 M tiles, 8 per image
    (5, 128, 128, 32, True, False),     # 4 channels per group
    (6, 256, 384, 16, True, False),     # 12 channels per group: groups straddle chunks and 128-wide tiles
    (7, 128, 512, 16, False, False),    # 16 channels per group, no SiLU
    (20, 128, 256, 32, True, False),    # 160 M tiles: CTA pairs, 256-wide tiles; images straddle the wave boundary
    (40, 64, 128, 32, True, False),     # 320 tiles of 128 columns: several waves per CTA
    (96, 128, 256, 32, True, False),    # the bench's batch: 768 M tiles, 256-wide tiles on CTA pairs
    (96, 64, 384, 16, False, True),     # both outputs, the bench's batch
    (20, 128, 256, 32, False, True),    # both outputs on CTA pairs
    (4, 64, 1024, 16, True, True),      # 32 channels per group
    (96, 128, 512, 8, True, False),     # 8x8: two images per 128-pixel tile, statistics per half tile
    (96, 128, 384, 8, False, True),     # 8x8, 12 channels per group: groups straddle the 128-wide tiles; both outputs
    (2, 64, 128, 8, True, True),        # 8x8, one tile
    (5, 128, 384, 8, True, True),       # 8x8, odd batch: the last tile holds one image and an empty half
    (1, 64, 128, 8, False, False),      # 8x8, a single image
])
def test_conv_groupnorm_fused(B, Cin, Cout, H, silu, dual):
    run_case(B, Cin, Cout, H, silu, dual)


def test_conv_groupnorm_fused_two_segments():
    """conv3x3 (+) 1x1 skip projection as one two-segment GEMM, raw and normalised outputs (the block's conv2)."""
    run_case(12, 256, 256, 32, False, True, skip_c=128)
    run_case(24, 384, 384, 16, False, True, skip_c=256)
    run_case(30, 512, 512, 8, False, True, skip_c=384)


def test_conv_groupnorm_fused_batch_invariance():
    """The statistics of a sample are summed in an order that does not depend on the batch: the first samples of a
    large batch (CTA pairs, several waves) equal the same samples run alone, bit for bit."""
    ops = _ops()
    g = torch.Generator(device=DEV).manual_seed(5)
    B, Cin, Cout, H = 40, 128, 256, 32
    x = torch.randn(B, Cin, H, H, device=DEV, generator=g)
    w = ops.pack_conv_weight(torch.randn(Cout, Cin, 3, 3, device=DEV, generator=g) / math.sqrt(9 * Cin))
    b = torch.randn(Cout, device=DEV, generator=g)
    gamma = 1.0 + 0.3 * torch.randn(Cout, device=DEV, generator=g)
    beta = 0.3 * torch.randn(Cout, device=DEV, generator=g)

    def run(n):
        ws = torch.zeros(ops.gn_workspace_bytes(n, n * H * H, Cout), device=DEV, dtype=torch.uint8)
        out = torch.empty(n * H * H, Cout, device=DEV, dtype=torch.bfloat16)
        ops.igemm([(rows(x[:n]), (n, H, H), Cin, 9)], w, Cout, out, bias=b,
                  gn=dict(gamma=gamma, beta=beta, groups=32, silu=True, ws=ws))
        return out

    big, small = run(B), run(2)
    assert torch.equal(big[:2 * H * H], small)
    # 8x8 images (two per tile)
    H = 8
    x = torch.randn(B, Cin, H, H, device=DEV, generator=g)
    big, small = run(B), run(3)
    assert torch.equal(big[:3 * H * H], small)


def test_conv_groupnorm_fused_rejects_unsupported_shapes():
    ops = _ops()
    x = torch.zeros(2 * 64, 128, device=DEV, dtype=torch.bfloat16)
    w = torch.zeros(128, 9 * 128, device=DEV, dtype=torch.bfloat16)
    out = torch.empty(2 * 64, 128, device=DEV, dtype=torch.bfloat16)
    ws = torch.zeros(1 << 16, device=DEV, dtype=torch.uint8)
    gam = torch.ones(128, device=DEV)
    with pytest.raises(RuntimeError):   # 4x4 images: eight samples per 128-pixel tile
        ops.igemm([(x, (8, 4, 4), 128, 9)], w, 128, out, gn=dict(gamma=gam, beta=gam, groups=32, silu=True, ws=ws))
    x = torch.zeros(256, 128, device=DEV, dtype=torch.bfloat16)
    out = torch.empty(256, 128, device=DEV, dtype=torch.bfloat16)
    with pytest.raises(RuntimeError):   # 2 channels per group
        ops.igemm([(x, (1, 16, 16), 128, 9)], w, 128, out, gn=dict(gamma=gam, beta=gam, groups=64, silu=True, ws=ws))
