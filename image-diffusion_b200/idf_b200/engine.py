"""Execution engines: turn a parameter tree (drop-in Unet / VAE module) into a fixed sequence of C-ABI kernel calls
over persistent channels-last bf16 workspaces. One engine instance serves one (batch, resolution); the whole call
sequence is allocation-free after the first run, so it can be captured into a CUDA graph.

Kernel sequence per DiffusionBlock layer (reference components.py:518-536):
  GN+SiLU -> conv3x3 (+bias +time bias) -> GN+SiLU -> [conv3x3 (+) 1x1 skip conv] -> GN -> QKV GEMM (V transposed)
  -> fused attention -> out_proj GEMM (+bias +residual)
"""
from __future__ import annotations

import torch

from . import ops
from .spec import unet_blocks, vae_program

BF16 = torch.bfloat16


class Act:
    """A channels-last activation: 2-D bf16 view of shape (B*H*W, C) (row stride may exceed C)."""
    __slots__ = ("t", "B", "H", "W", "C")

    def __init__(self, t, B, H, W, C):
        self.t, self.B, self.H, self.W, self.C = t, B, H, W, C

    @property
    def M(self):
        return self.B * self.H * self.W

    @property
    def grid(self):
        return (self.B, self.H, self.W)


class Workspace:
    """Named persistent device buffers; a name is allocated once and reused by every later call."""

    def __init__(self, device):
        self.device = device
        self.bufs = {}

    def get(self, name, rows, cols, dtype=BF16):
        key = (name, rows, cols, dtype)
        t = self.bufs.get(key)
        if t is None:
            t = torch.empty(rows, cols, device=self.device, dtype=dtype)
            self.bufs[key] = t
        return t

    def nbytes(self):
        return sum(t.numel() * t.element_size() for t in self.bufs.values())


def _f32(t):
    return t.detach().to(torch.float32).contiguous()


def _narrow_pack(weight, bias):
    """(Cout <= 16, Cin, 3, 3) conv -> bf16 (16, 9*Cin) weight rows (zero beyond Cout) + fp32 bias[16]: the operands of
    the implicit GEMM's 16-wide tile (idf_igemm_args.out_nchw)."""
    cout = weight.shape[0]
    wp = torch.zeros(16, weight.shape[1] * weight.shape[2] * weight.shape[3], device=weight.device, dtype=BF16)
    wp[:cout] = ops.pack_conv_weight(weight)
    bp = torch.zeros(16, device=weight.device, dtype=torch.float32)
    bp[:cout] = bias.detach().float()
    return wp, bp


class _Packed:
    """bf16 / fp32 device copies of a module's parameters in the layouts the kernels consume; rebuilt whenever a
    parameter has been modified in place or replaced (optimizer step, load_state_dict, .to())."""

    def __init__(self, module):
        self.module = module
        self.key = None
        self.w = {}

    def stale(self):
        # `_weights_epoch` is bumped by writers that bypass torch's version counter (the fused Adam kernel updates
        # the flat parameter buffer through a raw pointer: DiffusionTrainStep.step)
        key = (getattr(self.module, "_weights_epoch", 0),) + tuple((p.data_ptr(), p._version)
                                                                  for p in self.module.parameters())
        if key != self.key:
            self.key = key
            return True
        return False

    def install(self, w):
        """Refreshes the packed copies IN PLACE when the layout is unchanged, so that captured CUDA graphs (which hold
        raw pointers to these tensors) keep reading valid, current weights."""
        old = self.w

        def same(a, b):
            if type(a) is not type(b):
                return False
            if isinstance(a, list):  # _UpsamplePack
                return len(a) == len(b) and a.stacked.shape == b.stacked.shape
            return a.shape == b.shape and a.dtype == b.dtype and a.device == b.device

        if old and old.keys() == w.keys() and all(same(old[k], w[k]) for k in w):
            for k, v in w.items():
                if isinstance(v, list):  # _UpsamplePack: per-parity matrices + the stacked matrix
                    for (_, _, dst), (_, _, src) in zip(old[k], v):
                        dst.copy_(src)
                    old[k].stacked.copy_(v.stacked)
                else:
                    old[k].copy_(v)
        else:
            self.w = w

    def sd(self):
        return {k: v for k, v in self.module.state_dict().items()}


# =================================================================================================
# UNet
# =================================================================================================
class UnetEngine:
    def __init__(self, module, arch, device):
        self.m, self.arch, self.device = module, arch, device
        self.packed = _Packed(module)
        self.ws = Workspace(device)
        self.L, self.heads, self.G = arch["num_res_layers"], arch["num_heads"], arch["num_groups"]
        self.taps = None  # debugging aid: set to a dict to collect fp32 copies of intermediate activations
        # split-K: only the 4x4 stage (48 output tiles for 148 SMs at batch 96), with a FIXED split count so that the
        # summation order - and the output bits - do not depend on the batch size. At 8x8 the fp32 partial traffic
        # eats the gain (measured).
        self.splitk = int(__import__("os").environ.get("IDF_SPLITK_4X4", "3"))
        self._B, self._rng = 0, (0, 0)
        self.tc_tail = __import__("os").environ.get("IDF_TC_TAIL", "1") != "0"
        # GroupNorm fused into the producing conv's epilogue where an image is a whole number of 128-pixel tiles or half
        # of one (32x32 / 16x16 / 8x8 stages): 1 = GN2 (y1 -> h2, y1 never stored), 2 = also GN3 (conv2 stores x2 and h3)
        self.gn_fuse = int(__import__("os").environ.get("IDF_GN_FUSE", "2"))
        self._gn_ws = {}
        self.downs, self.mids, self.ups = unet_blocks(arch)
        for _, cin, cout in self.downs + self.mids + self.ups:
            if cin % 64 or cout % 128:
                raise ValueError(f"UnetEngine: block {cin}->{cout}: the tcgen05 path needs Cin % 64 == 0 and "
                                 f"Cout % 128 == 0")
        self.P = 0
        self.tp_off = {}
        for p, _, cout in self.downs + self.mids + self.ups:
            for l in range(self.L):
                self.tp_off[(p, l)] = self.P
                self.P += cout

    # ---------------------------------------------------------------------------------------------
    def prepare(self):
        if not self.packed.stale():
            return
        sd = self.packed.sd()
        w = {}
        pk = ops.pack_conv_weight
        for p, cin, cout in self.downs + self.mids + self.ups:
            for l in range(self.L):
                a = f"{p}.self_attns.{l}"
                w[f"{p}.{l}.w1"] = pk(sd[f"{p}.first_halfs.{l}.layers.2.weight"])
                w[f"{p}.{l}.b1"] = _f32(sd[f"{p}.first_halfs.{l}.layers.2.bias"])
                w[f"{p}.{l}.w2"] = torch.cat([pk(sd[f"{p}.second_halfs.{l}.layers.2.weight"]),
                                             pk(sd[f"{p}.residuals.{l}.weight"])], dim=1).contiguous()
                w[f"{p}.{l}.b2"] = _f32(sd[f"{p}.second_halfs.{l}.layers.2.bias"] + sd[f"{p}.residuals.{l}.bias"])
                w[f"{p}.{l}.wqkv"] = torch.cat([sd[a + ".to_q.weight"], sd[a + ".to_k.weight"],
                                                sd[a + ".to_v.weight"]], dim=0).to(BF16).contiguous()
                w[f"{p}.{l}.bqkv"] = _f32(torch.cat([sd[a + ".to_q.bias"], sd[a + ".to_k.bias"],
                                                     sd[a + ".to_v.bias"]]))
                w[f"{p}.{l}.wo"] = sd[a + ".out_proj.weight"].detach().to(BF16).contiguous()
                w[f"{p}.{l}.bo"] = _f32(sd[a + ".out_proj.bias"])
                for n, key in (("g1", f"{p}.first_halfs.{l}.layers.0"), ("g2", f"{p}.second_halfs.{l}.layers.0"),
                               ("g3", a + ".groupnorm")):
                    w[f"{p}.{l}.{n}w"] = _f32(sd[key + ".weight"])
                    w[f"{p}.{l}.{n}b"] = _f32(sd[key + ".bias"])
        for i in range(len(self.downs)):
            w[f"down.{i}.w"] = pk(sd[f"downsamples.{i}.down.weight"])
            w[f"down.{i}.b"] = _f32(sd[f"downsamples.{i}.down.bias"])
            w[f"up.{i}.w"] = ops.pack_upsample_conv_weights(sd[f"upsamples.{i}.conv.weight"])
            w[f"up.{i}.b"] = _f32(sd[f"upsamples.{i}.conv.bias"])
        w["in.w"], w["in.b"] = _f32(sd["in_conv.weight"]), _f32(sd["in_conv.bias"])
        w["out.gw"], w["out.gb"] = _f32(sd["out_conv.0.weight"]), _f32(sd["out_conv.0.bias"])
        w["out.w"], w["out.b"] = _f32(sd["out_conv.2.weight"]), _f32(sd["out_conv.2.bias"])
        w["out.wtc"], w["out.btc"] = _narrow_pack(sd["out_conv.2.weight"], sd["out_conv.2.bias"])
        w["t.factor"] = _f32(sd["time_embedding.factor"])
        w["t.w1"], w["t.b1"] = _f32(sd["time_embedding.embeddings.0.weight"]), _f32(sd["time_embedding.embeddings.0.bias"])
        w["t.w2"], w["t.b2"] = _f32(sd["time_embedding.embeddings.2.weight"]), _f32(sd["time_embedding.embeddings.2.bias"])
        w["t.cls"] = _f32(sd["class_embedding.weight"])
        order = [(p, l) for p, _, _ in self.downs + self.mids + self.ups for l in range(self.L)]
        w["t.wp"] = _f32(torch.cat([sd[f"{p}.time_projs.{l}.1.weight"] for p, l in order], dim=0))
        w["t.bp"] = _f32(torch.cat([sd[f"{p}.time_projs.{l}.1.bias"] for p, l in order], dim=0))
        self.packed.install(w)

    # ---------------------------------------------------------------------------------------------
    def _tap(self, name, t2d):
        if self.taps is not None:
            self.taps[name] = t2d.float().clone()

    def _gn_fused(self, key, B, M, cout, silu, out=None):
        """Arguments of the GroupNorm-fused igemm epilogue (idf_igemm_args.gn_*) for GroupNorm `key` of the module."""
        w = self.packed.w
        nbytes = ops.gn_workspace_bytes(B, M, cout)
        ws = self._gn_ws.get(nbytes)
        if ws is None:  # counters start at zero; every launch leaves them zero
            ws = self._gn_ws[nbytes] = torch.zeros(nbytes, device=self.device, dtype=torch.uint8)
        g = dict(gamma=w[key + "w"], beta=w[key + "b"], groups=self.G, silu=silu, ws=ws)
        if out is not None:
            g["out"] = out
        return g

    def _block(self, p, x: Act, cout, table, idx, final_dst=None) -> Act:
        w, G = self.packed.w, self.G
        B, H, W = x.grid
        M, HW = x.M, x.H * x.W
        hd = cout // self.heads
        tiles_ok = HW % 128 == 0 or HW == 64   # whole images per 128-pixel tile, or two 8x8 images
        fuse = self.gn_fuse if (self.taps is None and tiles_ok and (cout // G) % 4 == 0 and cout % G == 0) else 0
        if HW == 64 and cout >= 512:
            # 512-channel convs of the 8x8 stage: the fused launch is held to 128-wide tiles; measured per layer
            # (profiles/r02_gn_fused_per_layer.txt) the double-output form gains nothing there, the single one 1-3 us
            fuse = min(fuse, 1)
        sk = {}
        if self.splitk > 1 and HW <= 16:
            sk = dict(ws=self._splitk_ws(M * cout), splits=self.splitk)
        for l in range(self.L):
            cin = x.C
            k = f"{p}.{l}"
            h1 = self._buf("h1", HW, cin)
            ops.groupnorm_silu(x.t, h1, w[k + ".g1w"], w[k + ".g1b"], B, HW, cin, G, True)
            off = self.tp_off[(p, l)]
            h2 = self._buf("h2", HW, cout)
            y1 = None if fuse >= 1 else self._buf("y1", HW, cout)
            if fuse >= 1:  # conv1's epilogue normalises: y1 is never written
                ops.igemm([(h1, x.grid, cin, 9)], w[k + ".w1"], cout, h2, bias=w[k + ".b1"],
                          rowbias=table[:, off:off + cout], rowbias_idx=idx, gn=self._gn_fused(k + ".g2", B, M, cout, True))
            else:
                ops.igemm([(h1, x.grid, cin, 9)], w[k + ".w1"], cout, y1, bias=w[k + ".b1"],
                          rowbias=table[:, off:off + cout], rowbias_idx=idx, **sk)
                ops.groupnorm_silu(y1, h2, w[k + ".g2w"], w[k + ".g2b"], B, HW, cout, G, True)
            x2 = self._buf("x2", HW, cout)
            h3 = self._buf("h3", HW, cout)
            if fuse >= 2:  # conv2 stores the raw x2 (attention residual) and GroupNorm(x2) (QKV input)
                ops.igemm([(h2, x.grid, cout, 9), (x.t, x.grid, cin, 1)], w[k + ".w2"], cout, x2, bias=w[k + ".b2"],
                          gn=self._gn_fused(k + ".g3", B, M, cout, False, out=h3))
            else:
                ops.igemm([(h2, x.grid, cout, 9), (x.t, x.grid, cin, 1)], w[k + ".w2"], cout, x2, bias=w[k + ".b2"], **sk)
                ops.groupnorm_silu(x2, h3, w[k + ".g3w"], w[k + ".g3b"], B, HW, cout, G, False)
            # Q | K | V token-major in one buffer: the attention kernel takes V tiles as MN-major tcgen05 operands, so
            # the QKV GEMM has a plain TMA-store epilogue (no transposed V^T copy)
            qk = self._buf("qkv", HW, 3 * cout)
            ops.igemm([(h3, (1, 1, M), cout, 1)], w[k + ".wqkv"], 3 * cout, qk, bias=w[k + ".bqkv"])
            o = self._buf("o", HW, cout)
            ops.attention_qkv(qk, o, M, HW, self.heads, hd)
            if l == self.L - 1 and final_dst is not None:
                dst = final_dst
            else:
                dst = self._buf("xa" if (l % 2 == 0) else "xb", HW, cout)
            ops.igemm([(o, (1, 1, M), cout, 1)], w[k + ".wo"], cout, dst, bias=w[k + ".bo"], res=x2)
            if self.taps is not None:
                for nm, tt in (("h1", h1), ("y1", y1), ("h2", h2), ("x2", x2), ("h3", h3), ("qk", qk), ("o", o),
                               ("out", dst)):
                    self._tap(f"{k}.{nm}", tt)
            x = Act(dst, B, H, W, cout)
        return x

    def embedding_table(self, t_rows, ctx_rows, mask_rows, out=None):
        """(R, P) fp32 table of every layer's time-projection bias for R (timestep, class, mask) rows
        (TimeEmbedding + class embedding + all time_projs, unet.py:106-114 and components.py:526)."""
        self.prepare()
        w, ws = self.packed.w, self.ws
        R, D = t_rows.shape[0], self.arch["time_dim"]
        table = out if out is not None else ws.get("tp_table", R, self.P, torch.float32)
        scratch = ws.get("tp_scratch", R, 5 * D, torch.float32)
        ops.embed_time_class(t_rows, ctx_rows, mask_rows, w["t.factor"], w["t.w1"], w["t.b1"], w["t.w2"], w["t.b2"],
                             w["t.cls"], w["t.wp"], w["t.bp"], table, scratch)
        return table

    def run(self, x_nchw, t_rows, ctx_rows, mask_rows, row_idx, out_nchw, dup_input=False, table=None):
        """x_nchw fp32 (B, z, H, W) -> out_nchw fp32 (B, z, H, W). With dup_input the network runs at batch 2*B on
        [x ; x] (CFG batch doubling; out_nchw then has 2*B samples) without materialising the doubled input.

        t_rows int64 (R,), ctx_rows int64 (R,) or None, mask_rows fp32 (R,) or None describe R distinct
        (timestep, class) embedding rows; row_idx int32 (B,) maps each sample to its row (None: R == B, identity).
        With `table` (from embedding_table) the embedding kernels are skipped and row_idx indexes that table.
        """
        self.prepare()
        w, ws = self.packed.w, self.ws
        B, _, H, W = x_nchw.shape
        if dup_input:
            B *= 2
        if table is None:  # `table`: precomputed rows (e.g. every timestep of a sampling run), indexed by row_idx
            table = self.embedding_table(t_rows, ctx_rows, mask_rows)
        ch = list(self.arch["channels"])
        self._B, self._rng = B, (0, B)
        a0 = self._buf("in", H * W, ch[0])
        ops.conv3x3_small_cin(x_nchw, w["in.w"], w["in.b"], a0, dup=dup_input)
        self._tap("table", table)
        self._tap("in", a0)
        x = self._level(0, Act(a0, B, H, W, ch[0]), table, row_idx)
        h = self._buf("h1", x.H * x.W, x.C)
        ops.groupnorm_silu(x.t, h, w["out.gw"], w["out.gb"], x.B, x.H * x.W, x.C, self.G, True)
        if self.tc_tail and x.C % 64 == 0 and out_nchw.shape[1] <= 16:
            # 128 -> z_dim conv on the tensor cores: one 16-wide tile, epilogue writes fp32 NCHW planes
            ops.igemm([(h, x.grid, x.C, 9)], w["out.wtc"], 16, None, bias=w["out.btc"], out_nchw=out_nchw)
        else:
            ops.conv3x3_small_cout(h, w["out.w"], w["out.b"], out_nchw)
        return out_nchw

    # ---------------------------------------------------------------------------------------------
    # Every operation of the network is per sample, so a sample range [b0, b0 + nb) of the batch can run as its own
    # kernel chain on row slices of the same full-batch buffers (_rng). (Running the 8x8 / 4x4 interior as two
    # half-batch chains on two streams was measured: 3.724 vs 3.729 ms per step - the one-CTA-per-SM persistent
    # kernels do not overlap; not kept.)
    # ---------------------------------------------------------------------------------------------
    def _buf(self, name, HW, cols, dtype=BF16):
        b0, nb = self._rng
        return self.ws.get(name, self._B * HW, cols, dtype)[b0 * HW:(b0 + nb) * HW]

    def _splitk_ws(self, elems):
        return self.ws.get(f"splitk{self._rng[0]}", 1, self.splitk * elems, torch.float32)

    def _level(self, i, x: Act, table, idx) -> Act:
        """Down block i, everything below it, and the matching up block (unet.py:117-133)."""
        n = len(self.downs)
        p, _, cout = self.downs[i]
        cat = self._buf(f"cat{i}", x.H * x.W, 2 * cout)
        x = self._block(p, x, cout, table, idx, final_dst=cat[:, cout:])
        self._inner(i, x, cat, table, idx)
        pu, _, cu = self.ups[n - 1 - i]
        return self._block(pu, Act(cat, x.B, x.H, x.W, 2 * cout), cu, table, idx)

    def _inner(self, i, x: Act, cat, table, idx):
        """Downsample i -> deeper levels (or the mid blocks) -> Upsample conv into the left half of level i's concat
        buffer."""
        w, n = self.packed.w, len(self.downs)
        cout = x.C
        Hd, Wd = x.H // 2, x.W // 2
        nxt = self._buf("dn", Hd * Wd, cout)
        skd = {}
        if self.splitk > 1 and Hd * Wd <= 16:
            skd = dict(ws=self._splitk_ws(x.B * Hd * Wd * cout), splits=self.splitk)
        # stride-2 pad-0 conv read straight from the block output (TMA element strides): no parity-plane copy
        ops.igemm([(x.t, x.grid, cout, 9)], w[f"down.{i}.w"], cout, nxt, bias=w[f"down.{i}.b"], zero_pad_last=True,
                  s2_direct=True, **skd)
        self._tap(f"down.{i}", nxt)
        y = Act(nxt, x.B, Hd, Wd, cout)
        if i + 1 < n:
            y = self._level(i + 1, y, table, idx)
        else:
            for p, _, cm in self.mids:
                y = self._block(p, y, cm, table, idx)
        j = n - 1 - i
        # nearest-2x + conv3x3 as four sub-pixel convolutions on the low-resolution tensor (4/9 of the FLOPs, no
        # upsampled copy), stored straight into the left half of the concat buffer
        ops.upsample_conv3x3(y.t, y.grid, y.C, w[f"up.{j}.w"], y.C, cat[:, :y.C], bias=w[f"up.{j}.b"])
        self._tap(f"up.{j}", cat)

# =================================================================================================
# VAE encoder / decoder
# =================================================================================================
class VaeEngine:
    """Runs Encoder.down / Decoder.up (components.py:133-246) for one (batch, resolution)."""

    def __init__(self, module, arch, device):
        self.m, self.arch, self.device = module, arch, device
        self.packed = _Packed(module)
        self.ws = Workspace(device)
        self.G, self.heads = arch["num_groups"], arch["num_heads"]
        self.use_graph = __import__("os").environ.get("IDF_VAE_GRAPH", "1") != "0"
        self.gn_rows = __import__("os").environ.get("IDF_GN_ROWS", "1") != "0"
        self.tc_tail = __import__("os").environ.get("IDF_VAE_TC_TAIL", "1") != "0"
        self.graphs = {}

    def prepare(self):
        if not self.packed.stale():
            return
        sd = self.packed.sd()
        w = {}
        pk = ops.pack_conv_weight
        for part, prefix in (("encoder", "encoder.down"), ("decoder", "decoder.up")):
            for kind, i, cin, cout in vae_program(self.arch, part):
                p = f"{prefix}.{i}"
                if kind in ("conv1x1", "conv3x3"):
                    w[p + ".w"], w[p + ".b"] = _f32(sd[p + ".weight"]), _f32(sd[p + ".bias"])
                    if kind == "conv3x3" and cin % 64 == 0 and cout <= 16:
                        # narrow tail conv (128 -> 3, 384 -> z) on the tensor cores: one 16-wide tile (the CUDA-core
                        # kernel runs at ~6 TFLOP/s: 15 % of a batch-48 decode)
                        w[p + ".wtc"], w[p + ".btc"] = _narrow_pack(sd[p + ".weight"], sd[p + ".bias"])
                elif kind == "res":
                    w[p + ".g1w"], w[p + ".g1b"] = _f32(sd[p + ".branch.0.weight"]), _f32(sd[p + ".branch.0.bias"])
                    w[p + ".g2w"], w[p + ".g2b"] = _f32(sd[p + ".branch.3.weight"]), _f32(sd[p + ".branch.3.bias"])
                    w[p + ".w1"], w[p + ".b1"] = pk(sd[p + ".branch.2.weight"]), _f32(sd[p + ".branch.2.bias"])
                    if cin != cout:
                        w[p + ".w2"] = torch.cat([pk(sd[p + ".branch.5.weight"]),
                                                  pk(sd[p + ".residual_wrapper.weight"])], dim=1).contiguous()
                        w[p + ".b2"] = _f32(sd[p + ".branch.5.bias"] + sd[p + ".residual_wrapper.bias"])
                    else:
                        w[p + ".w2"], w[p + ".b2"] = pk(sd[p + ".branch.5.weight"]), _f32(sd[p + ".branch.5.bias"])
                elif kind == "attn":
                    w[p + ".gw"], w[p + ".gb"] = _f32(sd[p + ".groupnorm.weight"]), _f32(sd[p + ".groupnorm.bias"])
                    w[p + ".wqkv"] = torch.cat([sd[p + ".to_q.weight"], sd[p + ".to_k.weight"],
                                                sd[p + ".to_v.weight"]], dim=0).to(BF16).contiguous()
                    w[p + ".bqkv"] = _f32(torch.cat([sd[p + ".to_q.bias"], sd[p + ".to_k.bias"], sd[p + ".to_v.bias"]]))
                    w[p + ".wo"], w[p + ".bo"] = sd[p + ".out_proj.weight"].detach().to(BF16).contiguous(), _f32(sd[p + ".out_proj.bias"])
                elif kind == "up":
                    w[p + ".w"] = ops.pack_upsample_conv_weights(sd[f"{p}.conv.weight"])
                    w[p + ".b"] = _f32(sd[f"{p}.conv.bias"])
                elif kind == "down":
                    w[p + ".w"], w[p + ".b"] = pk(sd[f"{p}.down.weight"]), _f32(sd[f"{p}.down.bias"])
                elif kind == "gn_silu":
                    w[p + ".gw"], w[p + ".gb"] = _f32(sd[p + ".weight"]), _f32(sd[p + ".bias"])
        self.packed.install(w)

    # ---- layer runners -------------------------------------------------------------------------
    def _gn(self, x_t, y_t, gw, gb, B, HW, C, silu):
        """GroupNorm(+SiLU): tensors far beyond L2 (large batches at the 64x64 / 128x128 stages) take the whole-row
        kernels (fully coalesced, two launches), everything else the slab kernel (one launch)."""
        if B * HW * C * 2 >= ops.GN_ROWS_MIN_TOTAL_BYTES and self.gn_rows:
            part = self.ws.get("gn_part", 1, B * ((HW + 255) // 256) * self.G * 2, torch.float32)
            return ops.groupnorm_silu_rows(x_t, y_t, gw, gb, B, HW, C, self.G, silu, part)
        return ops.groupnorm_silu(x_t, y_t, gw, gb, B, HW, C, self.G, silu)

    def _res(self, p, x: Act, cout) -> Act:
        w, ws, G = self.packed.w, self.ws, self.G
        B, H, W = x.grid
        M, HW, cin = x.M, x.H * x.W, x.C
        h1 = ws.get("h1", M, cin)
        self._gn(x.t, h1, w[p + ".g1w"], w[p + ".g1b"], B, HW, cin, True)
        y1 = ws.get("y1", M, cout)
        ops.igemm([(h1, x.grid, cin, 9)], w[p + ".w1"], cout, y1, bias=w[p + ".b1"])
        h2 = ws.get("h2", M, cout)
        self._gn(y1, h2, w[p + ".g2w"], w[p + ".g2b"], B, HW, cout, True)
        dst = ws.get("xa" if x.t is not self.ws.bufs.get(("xa", M, cout, BF16)) else "xb", M, cout)
        if cin != cout:
            ops.igemm([(h2, x.grid, cout, 9), (x.t, x.grid, cin, 1)], w[p + ".w2"], cout, dst, bias=w[p + ".b2"])
        else:
            ops.igemm([(h2, x.grid, cout, 9)], w[p + ".w2"], cout, dst, bias=w[p + ".b2"], res=x.t)
        return Act(dst, B, H, W, cout)

    def _attn(self, p, x: Act) -> Act:
        """GroupNorm -> QKV -> softmax(QK^T/sqrt(hd)) V -> out_proj + x (components.py:64-103). head_dim <= 64 uses
        the fused kernel; the VAE's single 384-wide head forms the score matrix with two GEMMs per sample."""
        w, ws = self.packed.w, self.ws
        B, H, W = x.grid
        M, T, C = x.M, x.H * x.W, x.C
        hd = C // self.heads
        h3 = ws.get("h3", M, C)
        self._gn(x.t, h3, w[p + ".gw"], w[p + ".gb"], B, T, C, False)
        qk = ws.get("qk", M, 2 * C)
        vt = ws.get("vt", C, M)
        ops.igemm([(h3, (1, 1, M), C, 1)], w[p + ".wqkv"], 3 * C, qk, bias=w[p + ".bqkv"], vt=vt, vt_col0=2 * C)
        o = ws.get("o", M, C)
        if hd in (16, 32, 48, 64):
            ops.attention(qk, vt, o, M, T, self.heads, hd)
        else:
            if T % 128 or hd % 128:
                raise ValueError(f"VaeEngine attention: T={T}, head_dim={hd} unsupported by the GEMM-pair path")
            # One launch per stage for a whole chunk of samples (igemm with a per-image second operand): S = Q K^T in
            # fp32, row softmax to bf16, O = P V. Chunks of <= 64 samples bound the score workspace (64 x T x T fp32).
            cb = min(B, 64)
            s = ws.get("scores", cb * T, T, torch.float32)
            pm = ws.get("probs", cb * T, T)
            scale = 1.0 / (hd ** 0.5)
            for b0 in range(0, B, cb):
                nb = min(cb, B - b0)
                rows = slice(b0 * T, (b0 + nb) * T)
                grid = (nb, T // 128, 128)
                for hh in range(self.heads):
                    q = qk[rows, hh * hd:(hh + 1) * hd]
                    kk = qk[rows, C + hh * hd:C + (hh + 1) * hd]
                    ops.igemm([(q, grid, hd, 1)], kk, T, s[:nb * T], w_batch=(T, 0))          # S_i = Q_i K_i^T  (fp32)
                    ops.softmax_rows(s[:nb * T], pm[:nb * T], scale)                           # P = softmax(S / sqrt(hd))
                    ops.igemm([(pm[:nb * T], grid, T, 1)], vt[hh * hd:(hh + 1) * hd, b0 * T:], hd,
                              o[rows, hh * hd:(hh + 1) * hd], w_batch=(0, T))                  # O_i = P_i V_i
        dst = ws.get("xa" if x.t is not self.ws.bufs.get(("xa", M, C, BF16)) else "xb", M, C)
        ops.igemm([(o, (1, 1, M), C, 1)], w[p + ".wo"], C, dst, bias=w[p + ".bo"], res=x.t)
        return Act(dst, B, H, W, C)

    def _run(self, part, prefix, x_nchw, out_nchw):
        self.prepare()
        w, ws = self.packed.w, self.ws
        B, _, H, W = x_nchw.shape
        prog = vae_program(self.arch, part)
        x = None
        cur = x_nchw
        for n, (kind, i, cin, cout) in enumerate(prog):
            p = f"{prefix}.{i}"
            if kind == "conv1x1":  # tiny channel count, stays fp32 NCHW
                dst = out_nchw if n == len(prog) - 1 else ws.get(f"c11_{p}", B * cout, cur.shape[2] * cur.shape[3],
                                                               torch.float32).view(B, cout, cur.shape[2], cur.shape[3])
                if x is not None:  # encoder tail: previous conv3x3 wrote fp32 NCHW into `cur`
                    pass
                ops.conv1x1_small_f32(cur, w[p + ".w"].view(cout, cin), w[p + ".b"], dst)
                cur = dst
            elif kind == "conv3x3" and x is None:  # first wide conv: fp32 NCHW -> bf16 rows
                a0 = ws.get("in", B * H * W, cout)
                ops.conv3x3_small_cin(cur, w[p + ".w"], w[p + ".b"], a0)
                x = Act(a0, B, H, W, cout)
            elif kind == "conv3x3":  # last narrow conv: bf16 rows -> fp32 NCHW
                last = n == len(prog) - 1
                dst = out_nchw if last else ws.get("tail", B * cout, x.H * x.W, torch.float32).view(B, cout, x.H, x.W)
                if self.tc_tail and (p + ".wtc") in w:
                    ops.igemm([(x.t, x.grid, x.C, 9)], w[p + ".wtc"], 16, None, bias=w[p + ".btc"], out_nchw=dst)
                else:
                    ops.conv3x3_small_cout(x.t, w[p + ".w"], w[p + ".b"], dst)
                cur = dst
            elif kind == "res":
                x = self._res(p, x, cout)
            elif kind == "attn":
                x = self._attn(p, x)
            elif kind == "gn_silu":
                h = ws.get("h1", x.M, x.C)
                self._gn(x.t, h, w[p + ".gw"], w[p + ".gb"], x.B, x.H * x.W, x.C, True)
                x = Act(h, x.B, x.H, x.W, x.C)
            elif kind == "up":
                # nearest-2x + conv3x3 as four sub-pixel convolutions on the low-resolution tensor, one launch
                # (4/9 of the FLOPs, the upsampled tensor never exists)
                dst = ws.get("xu", 4 * x.M, cout)
                ops.upsample_conv3x3(x.t, x.grid, x.C, w[p + ".w"], cout, dst, bias=w[p + ".b"])
                x = Act(dst, x.B, 2 * x.H, 2 * x.W, cout)
            elif kind == "down":
                # stride-2 pad-0 conv read straight from the full-resolution tensor (TMA element strides)
                dst = ws.get("xd", x.M // 4, cout)
                ops.igemm([(x.t, x.grid, x.C, 9)], w[p + ".w"], cout, dst, bias=w[p + ".b"], zero_pad_last=True,
                          s2_direct=True)
                x = Act(dst, x.B, x.H // 2, x.W // 2, cout)
        return out_nchw

    def _run_graphed(self, part, prefix, x_nchw, out_nchw):
        """One CUDA-graph replay per call (the program is a fixed, allocation-free kernel sequence over persistent
        workspaces): static input / output buffers, weights re-packed in place before the replay when they changed."""
        if not self.use_graph:
            return self._run(part, prefix, x_nchw, out_nchw)
        self.prepare()
        st = self.graphs.get(part)
        if st is None:
            sin, sout = x_nchw.clone(), torch.empty_like(out_nchw)
            side = torch.cuda.Stream()
            side.wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(side):
                self._run(part, prefix, sin, sout)  # warm-up: allocates workspaces, sets kernel attributes
            torch.cuda.current_stream().wait_stream(side)
            torch.cuda.synchronize()
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g):
                self._run(part, prefix, sin, sout)
            st = self.graphs[part] = (g, sin, sout)
        g, sin, sout = st
        sin.copy_(x_nchw)
        g.replay()
        out_nchw.copy_(sout)
        return out_nchw

    def decode(self, z_nchw, out_nchw):
        return self._run_graphed("decoder", "decoder.up", z_nchw, out_nchw)

    def encode(self, x_nchw, out_nchw):
        return self._run_graphed("encoder", "encoder.down", x_nchw, out_nchw)
