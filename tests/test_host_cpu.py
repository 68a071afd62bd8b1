"""CPU-side checks: the C ABI (header <-> shared library <-> ctypes table), the drop-in modules' parameter trees and
checkpoint formats, error behaviour without a GPU, and the batch-sharding logic (gloo, world_size 2)."""
import os
import re
import tempfile

import pytest
import torch

from oracle import ref_path as O

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_exports_every_symbol_of_the_header():
    from idf_b200 import native
    header = open(os.path.join(ROOT, "include", "idf_b200.h")).read()
    declared = set(re.findall(r"\b(idf_[a-z0-9_]+)\s*\(", header))
    declared -= {"idf_nhwc", "idf_igemm_args"}
    lib = native.load()
    assert lib.idf_abi_version() == 4 == native.ABI_VERSION
    # the ctypes re-declarations of the argument structs have the library's sizes (native.load refuses to run otherwise)
    import ctypes
    for which, struct in enumerate((native.NHWC, native.IgemmArgs, native.WgradArgs, native.PackJob)):
        assert lib.idf_struct_size(which) == ctypes.sizeof(struct), struct.__name__
    assert lib.idf_struct_size(99) == -1
    for name in sorted(declared):
        assert hasattr(lib, name), f"{name} declared in include/idf_b200.h but not exported"
    assert declared == set(native.EXPORTS), declared ^ set(native.EXPORTS)
    assert lib.idf_last_error() is not None


def test_missing_library_fails_loudly(monkeypatch):
    from idf_b200 import native
    monkeypatch.setattr(native, "_lib", None)
    monkeypatch.setattr(native, "LIB_PATH", "/nonexistent/libidf_b200.so")
    with pytest.raises(native.NativeError, match="no fallback"):
        native.load()


def test_parameter_trees_match_reference_schema():
    from modules.unet import Unet
    from modules.vae import VAE
    u = Unet(**O.UNET_ARCH)
    assert {k: tuple(v.shape) for k, v in u.state_dict().items()} == \
        {k: tuple(v) for k, v in O.unet_param_shapes(O.UNET_ARCH).items()}
    assert sum(p.numel() for p in u.parameters()) == 60_475_523
    assert u.time_dim == 512 and u.num_classes == 3 and set(u.architecture) == set(O.UNET_ARCH)
    for arch, n in ((O.VAE_KL_ARCH, 36_319_935), (O.VAE_VQ_ARCH, 36_315_678)):
        v = VAE(**arch)
        assert {k: tuple(t.shape) for k, t in v.state_dict().items()} == \
            {k: tuple(s) for k, s in O.vae_param_shapes(arch).items()}
        assert sum(p.numel() for p in v.parameters()) == n
        assert (v.codebook is None) == (arch["bottleneck"] == "kl")
    # default init follows torch's Conv2d/Linear/GroupNorm/Embedding conventions
    sd = u.state_dict()
    assert torch.all(sd["out_conv.0.weight"] == 1) and torch.all(sd["out_conv.0.bias"] == 0)
    w = sd["down_blocks.0.first_halfs.0.layers.2.weight"]
    assert w.abs().max().item() <= 1 / (128 * 9) ** 0.5 + 1e-6
    assert torch.allclose(sd["time_embedding.factor"], 10000 ** (torch.arange(0, 256, dtype=torch.float32) / 256))


def test_checkpoint_round_trips():
    from modules.components import Scheduler
    from modules.diffusion import Diffusion
    from modules.unet import Unet
    from modules.vae import VAE
    small_u = dict(O.UNET_ARCH, channels=[128, 256], mid_channels=[256, 256], time_dim=64)
    small_v = dict(O.VAE_VQ_ARCH, channels=[128, 256], codebook_size=32)
    with tempfile.TemporaryDirectory() as d:
        u = Unet(**small_u)
        path = os.path.join(d, "sub", "unet.pt")
        u.to_checkpoint(path)
        ck = torch.load(path, weights_only=False)
        assert set(ck) == {"unet", "architecture"}
        ck["unet"] = {"_orig_mod." + k: v for k, v in ck["unet"].items()}  # what torch.compile leaves behind
        u2 = Unet.from_checkpoint(checkpoint=ck)
        assert all(torch.equal(a, b) for a, b in zip(u.state_dict().values(), u2.state_dict().values()))
        with pytest.raises(ValueError):
            Unet.from_checkpoint()
        v = VAE(**small_v)
        dfn = Diffusion(v, u, Scheduler(1000), "cat,dog,bird", "cuda")
        assert dfn.classes == ["cat", "dog", "bird"] and dfn.latent_shape == (3, 64, 64)
        bundle = os.path.join(d, "bundle.pt")
        dfn.to_checkpoint(bundle)
        ck = torch.load(bundle, weights_only=False)
        assert set(ck) == {"v", "u", "scheduler", "classes"} and ck["scheduler"]["type"] == "linear"
        v2 = VAE.from_checkpoint(checkpoint=ck["v"])
        assert all(torch.equal(a, b) for a, b in zip(v.state_dict().values(), v2.state_dict().values()))
        with pytest.raises(ValueError):
            VAE.from_checkpoint()


def test_error_behaviour_without_gpu():
    from modules.components import Codebook, Scheduler
    from modules.diffusion import Diffusion
    from modules.unet import Unet
    from modules.vae import VAE
    u = Unet(**dict(O.UNET_ARCH, channels=[128, 256], mid_channels=[256, 256], time_dim=64)).eval()
    with torch.no_grad(), pytest.raises(RuntimeError, match="no CPU path"):
        u(torch.zeros(1, 3, 8, 8), torch.zeros(1, dtype=torch.long))
    vq = VAE(**dict(O.VAE_VQ_ARCH, channels=[128, 256], codebook_size=32)).eval()
    kl = VAE(**dict(O.VAE_KL_ARCH, channels=[128, 256])).eval()
    with pytest.raises(ValueError, match="Cannot sample"):
        vq.encode(torch.zeros(1, 3, 16, 16), sample=True)
    with pytest.raises(ValueError, match="Cannot quantize"):
        kl.decode(torch.zeros(1, 3, 8, 8), quantize=True)
    with torch.no_grad(), pytest.raises(RuntimeError, match="no CPU path"):
        kl.decode(torch.zeros(1, 3, 8, 8))
    s = Scheduler(1000)
    with pytest.raises(RuntimeError, match="no CPU path"):
        s.add_noise(torch.zeros(1, 3, 4, 4), torch.zeros(1, 3, 4, 4), torch.zeros(1, dtype=torch.long))
    with pytest.raises(ValueError):
        Scheduler(10, type="quadratic")
    if not torch.cuda.is_available():
        with pytest.raises(AssertionError, match="GPU"):
            Diffusion(kl, u, s, "a,b,c", "cuda").sample(3, num_images=1)
    cb = Codebook(16, 3, 0.25, 0.99)
    assert set(dict(cb.named_parameters())) == {"embeddings.weight", "ema_w"} and "ema_cluster_size" in dict(cb.named_buffers())


def test_scheduler_tables_match_reference_golden():
    from modules.components import Scheduler
    g = torch.load(os.path.join(ROOT, "tests", "golden", "scheduler.pt"), weights_only=False)
    for typ in ("linear", "cosine"):
        s = Scheduler(1000, 1e-4, 0.02, typ)
        assert torch.equal(s.betas, g[typ]["betas"]) and torch.equal(s.alpha_cum_prod, g[typ]["alpha_cum_prod"])
        assert torch.equal(s.sqrt_one_minus_alpha_cum_prod, torch.sqrt(1 - s.alpha_cum_prod))


def test_layer_programs_match_oracle():
    from idf_b200.spec import vae_program
    for arch in (O.VAE_KL_ARCH, O.VAE_VQ_ARCH):
        assert vae_program(arch, "decoder") == O.decoder_program(arch)
        assert vae_program(arch, "encoder") == O.encoder_program(arch)


def test_shard_bounds_cover_everything():
    from idf_b200.dist import shard_bounds
    for total, world, align in ((4096, 8, 48 * 0 + 64), (4096, 3, 64), (96, 2, 48), (48, 4, 48), (10, 4, 1)):
        spans = [shard_bounds(total, world, r, align) for r in range(world)]
        assert spans[0][0] == 0 and spans[-1][1] == total
        assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
        assert all(lo % align == 0 and hi % align == 0 for lo, hi in spans)
    with pytest.raises(ValueError):
        shard_bounds(100, 2, 0, 48)


def _gloo_worker(rank, world, port, total, out):
    import torch.distributed as dist
    from idf_b200.dist import gather_shards, shard_bounds
    dist.init_process_group("gloo", init_method=f"tcp://127.0.0.1:{port}", rank=rank, world_size=world)
    lo, hi = shard_bounds(total, world, rank, 2)
    local = torch.arange(lo, hi, dtype=torch.float32)[:, None].repeat(1, 3)
    full = gather_shards(local, total, 2)
    ok = torch.equal(full, torch.arange(total, dtype=torch.float32)[:, None].repeat(1, 3))
    t = torch.tensor([1.0 if ok else 0.0])
    dist.all_reduce(t, op=dist.ReduceOp.MIN)
    if rank == 0:
        out.put(bool(t.item()))
    dist.destroy_process_group()


def test_gather_shards_gloo_world2():
    import socket
    import torch.multiprocessing as mp
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_gloo_worker, args=(r, 2, port, 10, q)) for r in range(2)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(120)
        assert p.exitcode == 0
    assert q.get(timeout=10) is True


# ---------------------------------------------------------------------------------------------
# training step: host-side logic (gradient layout, buckets, data-parallel averaging over gloo)
# ---------------------------------------------------------------------------------------------
def test_train_engine_gradient_layout_covers_every_parameter():
    """The flat gradient buffer holds every reference parameter exactly once, in backward-completion order, with
    q/k/v and the time projections stacked contiguously (they are written by one GEMM each)."""
    from idf_b200.train_engine import UnetTrainEngine
    from modules.unet import Unet
    m = Unet(**O.UNET_ARCH)
    e = UnetTrainEngine(m, m.architecture, "cpu")
    names = dict(m.named_parameters())
    assert set(e.grad_names) == set(names) and len(e.grad_names) == len(names) == 331
    assert "time_embedding.factor" not in e.goff  # a buffer, not a parameter
    spans = sorted((e.goff[n], e.goff[n] + names[n].numel()) for n in e.grad_names)
    assert all(a[1] <= b[0] for a, b in zip(spans, spans[1:])) and spans[-1][1] <= e.flat_numel
    assert e.flat_numel - sum(p.numel() for p in names.values()) < 4 * len(names)
    a = "mid_blocks.0.self_attns.1"
    assert e.goff[a + ".to_k.weight"] == e.goff[a + ".to_q.weight"] + 512 * 512
    assert e.gspan(a + ".to_q.weight", a + ".to_v.weight").numel() == 3 * 512 * 512
    assert e.grad_names[0] == "out_conv.2.weight" and e.grad_names[-1] == "class_embedding.weight"
    ends = [end for _, end in e.bucket_ends]
    assert ends == sorted(ends) and ends[-1] == e.flat_numel and len(ends) == 8
    for n in e.grad_names:  # views alias the flat buffer
        assert e.gv[n].data_ptr() == e.flat_grad.data_ptr() + 4 * e.goff[n] and e.gv[n].shape == names[n].shape


def _ddp_worker(rank, world, port, out):
    import torch.distributed as dist
    from idf_b200.trainer import GradBuckets
    dist.init_process_group("gloo", init_method=f"tcp://127.0.0.1:{port}", rank=rank, world_size=world)
    n = 1000
    ends = [("a", 100), ("b", 450), ("c", 460), ("d", 1000)]
    flat = torch.arange(n, dtype=torch.float32) * (rank + 1)
    gb = GradBuckets(flat, ends, min_bucket_elems=300)
    plan = gb.plan()
    ok = plan == [(0, 450), (450, 1000)]  # stages merged until a bucket holds >= 300 elements; the last closes the rest
    for stage in range(len(ends)):
        gb.on_stage_done(stage)
    gb.works.clear()
    expect = torch.arange(n, dtype=torch.float32) * sum(r + 1 for r in range(world))
    ok = ok and torch.equal(flat, expect)
    t = torch.tensor([1.0 if ok else 0.0])
    dist.all_reduce(t, op=dist.ReduceOp.MIN)
    if rank == 0:
        out.put(bool(t.item()))
    dist.destroy_process_group()


def test_grad_buckets_gloo_world2():
    import socket
    import torch.multiprocessing as mp
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_ddp_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(120)
        assert p.exitcode == 0
    assert q.get(timeout=10) is True


# ---------------------------------------------------------------------------------------------
# host-side pieces either side of the hot path
# ---------------------------------------------------------------------------------------------
def test_warmup_lr_matches_reference_formula():
    """trainers/diffusion_trainer.py:133-138."""
    from idf_b200.pipeline import warmup_lr
    lr, ws = 2e-4, 100
    assert warmup_lr(0, lr, ws) == pytest.approx(lr / 100)
    assert warmup_lr(50, lr, ws) == pytest.approx(lr / 100 + (lr - lr / 100) * 0.5)
    assert warmup_lr(100, lr, ws) == lr and warmup_lr(10**6, lr, ws) == lr
    assert warmup_lr(0, lr, 0) == lr


def test_image_grid_matches_make_grid():
    """scripts/sample_grid.py:44-45: make_grid(images, nrow) then clamp(-1, 1), (x + 1) / 2."""
    from idf_b200.pipeline import image_grid
    from torchvision.utils import make_grid
    imgs = torch.randn(7, 3, 16, 12, generator=torch.Generator().manual_seed(3)) * 1.5
    ref = (make_grid(imgs, nrow=3).permute(1, 2, 0).clamp(-1.0, 1.0).numpy() + 1) / 2
    got = image_grid(imgs, nrow=3)
    assert got.dtype.name == "uint8" and got.shape == ref.shape
    assert abs(got.astype("float32") / 255.0 - ref).max() <= 0.5 / 255 + 1e-6


def test_upsample_subpixel_weight_packing_matches_nearest2x_conv():
    """Host side of the fused Upsample launch (idf_igemm_args.out_up2 == 2): the four pre-summed 2x2 weight matrices,
    their stacking order (parity = row parity * 2 + column parity) and the tap offsets the kernel derives from the
    parity, (dh, dw) = ((tap >> 1) - 1 + p, (tap & 1) - 1 + q), reproduce nearest-2x + conv3x3 (components.py:124-130)
    in plain fp32 torch."""
    import torch.nn.functional as F
    from idf_b200 import ops
    g = torch.Generator().manual_seed(5)
    B, C, Co, H = 2, 64, 128, 6
    x = torch.randn(B, C, H, H, generator=g)
    w = torch.randn(Co, C, 3, 3, generator=g) / (9 * C) ** 0.5
    ref = F.conv2d(F.interpolate(x, scale_factor=2.0, mode="nearest"), w, None, padding=1)
    packed = ops.pack_upsample_conv_weights(w)
    assert packed.stacked.shape == (4 * Co, 4 * C)
    xp = F.pad(x, (1, 1, 1, 1))  # zero padding = TMA out-of-bounds fill
    out = torch.zeros(B, Co, 2 * H, 2 * H)
    for par in range(4):
        p, q = par >> 1, par & 1
        (pp, qq), offs, wp = packed[par]
        assert (pp, qq) == (p, q)
        wpar = packed.stacked[par * Co:(par + 1) * Co].float()  # (Co, 4 * C): tap-major, then channel
        assert torch.equal(wpar, wp.float())
        acc = torch.zeros(B, Co, H, H)
        for tap in range(4):
            dh, dw = (tap >> 1) - 1 + p, (tap & 1) - 1 + q
            assert offs[tap] == (dh, dw)
            xs = xp[:, :, 1 + dh:1 + dh + H, 1 + dw:1 + dw + H]
            acc += torch.einsum("bchw,oc->bohw", xs, wpar[:, tap * C:(tap + 1) * C])
        out[:, :, p::2, q::2] = acc
    # the packed weights are bf16: compare at bf16 resolution of the weights
    assert ((out - ref).norm() / ref.norm()).item() < 5e-3


def test_persistent_kernel_tile_walk_matches_closed_form():
    """The persistent implicit GEMM walks its work units with carry adds instead of divisions (csrc/igemm.cu: TileWalk);
    idf_tile_walk_trace exposes that walk on the host. Unit u = (tile_m * n_tiles + n_idx) * splits + split."""
    import ctypes as C
    from idf_b200 import native
    lib = native.load()
    for stride in (1, 2, 7, 74, 148):
        for splits in (1, 2, 3, 4):
            for n_tiles in (1, 2, 3, 6, 9, 12):
                for u0 in (0, 1, stride - 1):
                    steps = 40
                    out = (C.c_int32 * (3 * steps))()
                    assert lib.idf_tile_walk_trace(u0, stride, splits, n_tiles, steps, out, None) == 0
                    for i in range(steps):
                        u = u0 + i * stride
                        want = (u // splits // n_tiles, u // splits % n_tiles, u % splits)
                        assert tuple(out[3 * i:3 * i + 3]) == want, (stride, splits, n_tiles, u0, i)
    bad = (C.c_int32 * 3)()
    assert lib.idf_tile_walk_trace(0, 0, 1, 1, 1, bad, None) != 0


def test_groupnorm_fused_launch_plan_never_makes_a_cta_wait_for_itself():
    """GroupNorm-fused igemm launches (idf_igemm_args.gn_mode): a tile waits inside the kernel until all tiles of its
    image have published their partial sums. That is deadlock-free only if no CTA (walker) ever holds two units of one
    image, and free of whole-tile stalls only if an image's units share a wave. idf_gn_plan_check walks the unit lists
    as the kernel does; checked for every stage shape of the UNet over batch sizes, tile widths and SM counts."""
    import ctypes as C
    from idf_b200 import native
    lib = native.load()
    out = (C.c_int32 * 5)()
    seen = 0
    for sms in (148, 132, 64):
        for batch in (1, 2, 3, 5, 6, 8, 48, 96, 128, 256):
            # (tiles per image, N tiles): 32x32 / 16x16 stages, and 8x8 (two images per tile: one tile row block per "image")
            shapes = [(8, 1, batch * 8), (8, 2, batch * 8), (2, 1, batch * 2), (2, 2, batch * 2), (2, 3, batch * 2),
                      (1, 3, (batch + 1) // 2), (1, 4, (batch + 1) // 2)]
            for tpi, n_tiles, m_tiles in shapes:
                for pair in (0, 1):
                    if pair and (tpi & 1):
                        continue
                    rc = lib.idf_gn_plan_check(m_tiles, tpi, n_tiles, pair, sms, out, None)
                    assert rc == 0, (sms, batch, tpi, n_tiles, pair, lib.idf_last_error())
                    walkers, per_img, waves, same_walker, split = list(out)
                    units = (m_tiles // 2 if pair else m_tiles) * n_tiles
                    assert same_walker == 0 and split == 0, (sms, batch, tpi, n_tiles, pair, list(out))
                    assert walkers % per_img == 0 and walkers <= (sms // 2 if pair else sms)
                    assert waves == -(-units // walkers)
                    seen += 1
    assert seen > 300
    # an image with more units than the GPU has walkers cannot be fused (128x128 image, two N tiles)
    assert lib.idf_gn_plan_check(256, 128, 2, 0, 148, out, None) != 0
    assert lib.idf_gn_plan_check(16, 3, 1, 0, 148, out, None) != 0   # M tiles not a whole number of images


def test_packed_weights_refresh_in_place_and_follow_the_weights_epoch():
    """Round-1 advisor findings, host side (no GPU needed): the packed operand copies are rewritten IN PLACE when the
    layout is unchanged (captured graphs hold raw pointers to them), and a writer that bypasses torch's version
    counter (the fused Adam kernel) invalidates them by bumping `_weights_epoch`."""
    import torch
    from idf_b200.engine import _Packed

    m = torch.nn.Linear(4, 3)
    pk = _Packed(m)
    assert pk.stale() and not pk.stale()
    first = {"w": m.weight.detach().to(torch.bfloat16).clone(), "b": m.bias.detach().clone()}
    pk.install(first)
    ptrs = {k: v.data_ptr() for k, v in pk.w.items()}
    m.weight.data.add_(1.0)               # raw write: no version bump, same storage -> not detected by itself
    assert not pk.stale()
    m._weights_epoch = getattr(m, "_weights_epoch", 0) + 1
    assert pk.stale()
    pk.install({"w": m.weight.detach().to(torch.bfloat16).clone(), "b": m.bias.detach().clone()})
    assert {k: v.data_ptr() for k, v in pk.w.items()} == ptrs          # same buffers ...
    assert torch.equal(pk.w["w"].float(), m.weight.detach().to(torch.bfloat16).float())  # ... new values
    with torch.no_grad():
        m.weight.mul_(2.0)                # in-place op through torch: version counter moves
    assert pk.stale()
    pk.install({"w": torch.zeros(5, 4, dtype=torch.bfloat16), "b": m.bias.detach().clone()})  # layout changed: replaced
    assert pk.w["w"].shape == (5, 4)
