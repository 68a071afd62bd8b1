"""UNet training step engine (reference trainers/diffusion_trainer.py:141-187): forward pass that
keeps what the backward needs, and the backward pass as a fixed sequence of C-ABI kernel calls writing fp32
gradients (PyTorch parameter layout) into ONE flat buffer ordered by backward completion time, so that data-parallel
gradient buckets are contiguous ranges that become final progressively (idf_b200/dist.py).

Per DiffusionBlock layer (components.py:518-536), backward of
  h1 = SiLU(GN1(x)); y1 = conv3x3(h1) + b1 + tproj; h2 = SiLU(GN2(y1)); x2 = conv3x3(h2) + conv1x1(x) + b2;
  h3 = GN3(x2); qkv = Linear(h3); o = attention(q, k, v); out = Linear(o) + x2
is: wgrad/colsum/dgrad of out_proj -> attention backward -> wgrad/colsum/dgrad of QKV -> GN3 backward (+ dout) ->
wgrad x2 / colsum / dgrad of the second conv -> GN2+SiLU backward -> wgrad/colsum (bias + per-sample time-bias
gradient)/dgrad of the first conv -> GN1+SiLU backward -> 1x1 skip dgrad with the GN1 result as epilogue addend.
Data gradients of all convolutions run on the SAME tcgen05 implicit-GEMM kernel as the forward pass, reading the SAME
forward-packed weights as MN-major operands (mirrored tap offsets, no transposed weight copies); weight gradients on
the MN-major tcgen05 kernel (csrc/wgrad.cu).
"""
from __future__ import annotations

import torch

from . import native, ops
from .engine import BF16, Act, Workspace
from .spec import unet_blocks

F32 = torch.float32


def _align(n, a=4):
    return (n + a - 1) // a * a


class UnetTrainEngine:
    def __init__(self, module, arch, device):
        self.m, self.arch, self.device = module, arch, device
        self.ws = Workspace(device)
        self.L, self.heads, self.G = arch["num_res_layers"], arch["num_heads"], arch["num_groups"]
        self.D = arch["time_dim"]
        self.downs, self.mids, self.ups = unet_blocks(arch)
        for _, cin, cout in self.downs + self.mids + self.ups:
            if cin % 64 or cout % 128:
                raise ValueError(f"UnetTrainEngine: block {cin}->{cout}: the tcgen05 path needs Cin % 64 == 0 and "
                                 f"Cout % 128 == 0")
        self.order = [(p, l) for p, _, _ in self.downs + self.mids + self.ups for l in range(self.L)]
        self.P, self.tp_off = 0, {}
        for p, _, cout in self.downs + self.mids + self.ups:
            for l in range(self.L):
                self.tp_off[(p, l)] = self.P
                self.P += cout
        self.params = dict(module.named_parameters())
        self._layout_grads()
        self.pw = {}       # packed weights (persistent buffers)
        self.pkey = None
        self.saved = {}

    # ---------------------------------------------------------------------------------------------
    # flat gradient buffer, backward-completion order
    # ---------------------------------------------------------------------------------------------
    def _layer_names(self, p, l):
        a = f"{p}.self_attns.{l}"
        return [a + ".out_proj.weight", a + ".out_proj.bias",
                a + ".to_q.weight", a + ".to_k.weight", a + ".to_v.weight",
                a + ".to_q.bias", a + ".to_k.bias", a + ".to_v.bias",
                a + ".groupnorm.weight", a + ".groupnorm.bias",
                f"{p}.second_halfs.{l}.layers.2.weight", f"{p}.second_halfs.{l}.layers.2.bias",
                f"{p}.residuals.{l}.weight", f"{p}.residuals.{l}.bias",
                f"{p}.second_halfs.{l}.layers.0.weight", f"{p}.second_halfs.{l}.layers.0.bias",
                f"{p}.first_halfs.{l}.layers.2.weight", f"{p}.first_halfs.{l}.layers.2.bias",
                f"{p}.first_halfs.{l}.layers.0.weight", f"{p}.first_halfs.{l}.layers.0.bias"]

    def _layout_grads(self):
        names = ["out_conv.2.weight", "out_conv.2.bias", "out_conv.0.weight", "out_conv.0.bias"]
        self.bucket_marks = []  # (stage name, number of names completed when that stage's backward has been issued)
        nlev = len(self.downs)
        for i in reversed(range(nlev)):
            p = self.ups[i][0]
            for l in reversed(range(self.L)):
                names += self._layer_names(p, l)
            names += [f"upsamples.{i}.conv.weight", f"upsamples.{i}.conv.bias"]
            self.bucket_marks.append((p, len(names)))
        for p, _, _ in reversed(self.mids):
            for l in reversed(range(self.L)):
                names += self._layer_names(p, l)
            self.bucket_marks.append((p, len(names)))
        for i in reversed(range(nlev)):
            p = self.downs[i][0]
            names += [f"downsamples.{i}.down.weight", f"downsamples.{i}.down.bias"]
            for l in reversed(range(self.L)):
                names += self._layer_names(p, l)
            self.bucket_marks.append((p, len(names)))
        names += ["in_conv.weight", "in_conv.bias"]
        names += [f"{p}.time_projs.{l}.1.weight" for p, l in self.order]
        names += [f"{p}.time_projs.{l}.1.bias" for p, l in self.order]
        names += ["time_embedding.embeddings.2.weight", "time_embedding.embeddings.2.bias",
                  "time_embedding.embeddings.0.weight", "time_embedding.embeddings.0.bias", "class_embedding.weight"]
        self.bucket_marks.append(("embed", len(names)))
        missing = set(self.params) - set(names)
        if missing or len(names) != len(self.params):
            raise RuntimeError(f"UnetTrainEngine: gradient layout does not cover the parameter tree: {sorted(missing)[:5]}")
        self.grad_names = names
        self.goff, off = {}, 0
        for n in names:
            self.goff[n] = off
            # stacked groups (q/k/v, time projections) must stay contiguous: their sizes are multiples of 4 anyway
            off += _align(self.params[n].numel())
        self.flat_numel = off
        self.flat_grad = torch.zeros(off, device=self.device, dtype=F32)
        self.gv = {n: self.flat_grad[self.goff[n]:self.goff[n] + self.params[n].numel()].view(self.params[n].shape)
                   for n in names}
        # bucket boundaries (element offsets) in completion order
        self.bucket_ends = [(s, self.goff[names[k]] if k < len(names) else off) for s, k in self.bucket_marks]

    def g(self, name):
        return self.gv[name]

    def gspan(self, first, last):
        """Contiguous flat-buffer span covering parameters first..last (adjacent in the layout)."""
        a, b = self.goff[first], self.goff[last] + self.params[last].numel()
        return self.flat_grad[a:b]

    # ---------------------------------------------------------------------------------------------
    # packed weights (forward + data-gradient layouts), refreshed in place so that addresses stay fixed
    # ---------------------------------------------------------------------------------------------
    def _buf(self, name, rows, cols, dtype=BF16):
        buf = self.pw.get(name)
        if buf is None:
            buf = self.pw[name] = torch.empty(rows, cols, device=self.device, dtype=dtype)
        return buf

    # tap offsets of the data gradient of a 3x3 stride-1 'same' conv when the weight keeps its (kh, kw) order:
    # dx[p] = sum_t dy[p - off(t)] W_t^T with off(t) = (kh - 1, kw - 1)
    DGRAD_TAPS = [(1 - kh, 1 - kw) for kh in range(3) for kw in range(3)]

    def prepare(self, force=False):
        """Refreshes the bf16 operand copies of the weights IN PLACE (fixed addresses: safe inside a captured graph)
        with ONE launch of idf_pack_weights over a job table; fp32 parameters (biases, norm gains, embedding MLP) are
        used where they live, without a copy. The table is rebuilt only when a parameter's storage moves."""
        key = tuple((p.data_ptr(), p._version) for p in self.params.values())
        if not force and key == self.pkey:
            return
        self.pkey = key
        ptrs = tuple(k[0] for k in key)
        if ptrs != getattr(self, "_job_ptrs", None):
            self._build_pack_jobs()
            self._job_ptrs = ptrs
        native.call("idf_pack_weights", self._jobs_dev.data_ptr(), self._job_prefix.data_ptr(), self._njobs,
                    self._pack_ctas)

    def _build_pack_jobs(self):
        sd = {k: v.detach() for k, v in self.params.items()}
        for k, v in sd.items():
            if v.dtype != F32 or not v.is_contiguous():
                raise RuntimeError(f"UnetTrainEngine: parameter {k} must be contiguous fp32")
        pw = self.pw
        jobs = []

        def job(src, dst, n_outer, n_taps, n_inner, so, si, st, dldo, dldt, out_f32=0, src2=None, src_off=0):
            j = native.PackJob()
            j.src, j.src2 = src.data_ptr() + 4 * src_off, (src2.data_ptr() if src2 is not None else None)
            j.dst = dst.data_ptr()
            j.n_outer, j.n_taps, j.n_inner, j.out_f32 = n_outer, n_taps, n_inner, out_f32
            j.so, j.si, j.st, j.dldo, j.dldt = so, si, st, dldo, dldt
            jobs.append(j)

        def conv_fwd(dst2d, w):      # (O, I, kh, kw) -> rows = output channel, columns (tap, ci)
            o, i, kh, kw = w.shape
            job(w, dst2d, o, kh * kw, i, i * kh * kw, kh * kw, 1, dst2d.stride(0), i)

        def lin(dst2d, w2d):         # (O, I) row-major copy
            o, i = w2d.shape
            job(w2d, dst2d, o, 1, i, i, 1, 0, dst2d.stride(0), 0)

        def vec(dst1d, a, b=None):   # fp32 vector copy / sum
            job(a, dst1d, 1, 1, a.numel(), 0, 1, 0, 0, 0, out_f32=1, src2=b)

        for p, cin0, cout in self.downs + self.mids + self.ups:
            for l in range(self.L):
                cin = cin0 if l == 0 else cout
                k, a = f"{p}.{l}", f"{p}.self_attns.{l}"
                w1, w2 = sd[f"{p}.first_halfs.{l}.layers.2.weight"], sd[f"{p}.second_halfs.{l}.layers.2.weight"]
                wr = sd[f"{p}.residuals.{l}.weight"].view(cout, cin)
                conv_fwd(self._buf(k + ".w1", cout, 9 * cin), w1)
                w2b = self._buf(k + ".w2", cout, 9 * cout + cin)
                conv_fwd(w2b[:, :9 * cout], w2)
                lin(w2b[:, 9 * cout:], wr)
                vec(self._buf(k + ".b2", 1, cout, F32).view(-1), sd[f"{p}.second_halfs.{l}.layers.2.bias"],
                    sd[f"{p}.residuals.{l}.bias"])
                wqkv = self._buf(k + ".wqkv", 3 * cout, cout)
                bqkv = self._buf(k + ".bqkv", 1, 3 * cout, F32).view(-1)
                for j, n in enumerate(("to_q", "to_k", "to_v")):
                    wj = sd[f"{a}.{n}.weight"]
                    lin(wqkv[j * cout:(j + 1) * cout], wj)
                    vec(bqkv[j * cout:(j + 1) * cout], sd[f"{a}.{n}.bias"])
                lin(self._buf(k + ".wo", cout, cout), sd[a + ".out_proj.weight"])
                pw[k + ".b1"], pw[k + ".bo"] = sd[f"{p}.first_halfs.{l}.layers.2.bias"], sd[a + ".out_proj.bias"]
                pw[k + ".b2"], pw[k + ".bqkv"] = pw[k + ".b2"].view(-1), bqkv
                for n, key_ in (("g1", f"{p}.first_halfs.{l}.layers.0"), ("g2", f"{p}.second_halfs.{l}.layers.0"),
                                ("g3", a + ".groupnorm")):
                    pw[f"{k}.{n}w"], pw[f"{k}.{n}b"] = sd[key_ + ".weight"], sd[key_ + ".bias"]
        for i in range(len(self.downs)):
            wd, wu = sd[f"downsamples.{i}.down.weight"], sd[f"upsamples.{i}.conv.weight"]
            c = wd.shape[0]
            conv_fwd(self._buf(f"down.{i}.w", c, 9 * c), wd)
            pw[f"down.{i}.b"] = sd[f"downsamples.{i}.down.bias"]
            cu = wu.shape[0]
            conv_fwd(self._buf(f"up.{i}.w", cu, 9 * cu), wu)
            pw[f"up.{i}.b"] = sd[f"upsamples.{i}.conv.bias"]
        pw["in.w"], pw["in.b"] = sd["in_conv.weight"], sd["in_conv.bias"]
        pw["out.gw"], pw["out.gb"] = sd["out_conv.0.weight"], sd["out_conv.0.bias"]
        pw["out.w"], pw["out.b"] = sd["out_conv.2.weight"], sd["out_conv.2.bias"]
        pw["t.factor"] = self.m.time_embedding.factor.detach().to(F32)
        pw["t.w1"], pw["t.b1"] = sd["time_embedding.embeddings.0.weight"], sd["time_embedding.embeddings.0.bias"]
        pw["t.w2"], pw["t.b2"] = sd["time_embedding.embeddings.2.weight"], sd["time_embedding.embeddings.2.bias"]
        pw["t.cls"] = sd["class_embedding.weight"]
        twp, tbp = self._buf("t.wp", self.P, self.D, F32), self._buf("t.bp", 1, self.P, F32).view(-1)
        pw["t.bp"] = tbp
        for p, l in self.order:
            off, wproj = self.tp_off[(p, l)], sd[f"{p}.time_projs.{l}.1.weight"]
            cout = wproj.shape[0]
            job(wproj, twp[off:off + cout], cout, 1, self.D, self.D, 1, 0, self.D, 0, out_f32=1)
            vec(tbp[off:off + cout], sd[f"{p}.time_projs.{l}.1.bias"])
        # device job table + first-CTA prefix
        import ctypes
        arr = (native.PackJob * len(jobs))(*jobs)
        raw = torch.frombuffer(bytearray(ctypes.string_at(ctypes.addressof(arr), ctypes.sizeof(arr))), dtype=torch.uint8)
        prefix, tot = [], 0
        for j in jobs:
            prefix.append(tot)
            tot += j.n_outer
        self._jobs_dev = raw.to(self.device)
        self._job_prefix = torch.tensor(prefix, dtype=torch.int32, device=self.device)
        self._njobs, self._pack_ctas = len(jobs), tot

    # ---------------------------------------------------------------------------------------------
    # forward (activations kept per layer)
    # ---------------------------------------------------------------------------------------------
    def _fwd_block(self, p, x: Act, cout, table, final_dst=None) -> Act:
        w, ws, G = self.pw, self.ws, self.G
        B, H, W = x.grid
        M, HW = x.M, x.H * x.W
        hd = cout // self.heads
        for l in range(self.L):
            cin, k = x.C, f"{p}.{l}"
            S = self.saved[k] = dict(x=x)
            h1, st1 = ws.get(k + ".h1", M, cin), ws.get(k + ".st1", B, 2 * G, F32)
            ops.groupnorm_silu_train(x.t, h1, w[k + ".g1w"], w[k + ".g1b"], B, HW, cin, G, True, st1)
            y1 = ws.get(k + ".y1", M, cout)
            off = self.tp_off[(p, l)]
            ops.igemm([(h1, x.grid, cin, 9)], w[k + ".w1"], cout, y1, bias=w[k + ".b1"], rowbias=table[:, off:off + cout])
            h2, st2 = ws.get(k + ".h2", M, cout), ws.get(k + ".st2", B, 2 * G, F32)
            ops.groupnorm_silu_train(y1, h2, w[k + ".g2w"], w[k + ".g2b"], B, HW, cout, G, True, st2)
            x2 = ws.get(k + ".x2", M, cout)
            ops.igemm([(h2, x.grid, cout, 9), (x.t, x.grid, cin, 1)], w[k + ".w2"], cout, x2, bias=w[k + ".b2"])
            h3, st3 = ws.get(k + ".h3", M, cout), ws.get(k + ".st3", B, 2 * G, F32)
            ops.groupnorm_silu_train(x2, h3, w[k + ".g3w"], w[k + ".g3b"], B, HW, cout, G, False, st3)
            qkv = ws.get(k + ".qkv", M, 3 * cout)
            ops.igemm([(h3, (1, 1, M), cout, 1)], w[k + ".wqkv"], 3 * cout, qkv, bias=w[k + ".bqkv"])
            o, lse = ws.get(k + ".o", M, cout), ws.get(k + ".lse", M, self.heads, F32)
            ops.attention_qkv(qkv, o, M, HW, self.heads, hd, lse=lse)
            dst = final_dst if (l == self.L - 1 and final_dst is not None) else ws.get(k + ".out", M, cout)
            ops.igemm([(o, (1, 1, M), cout, 1)], w[k + ".wo"], cout, dst, bias=w[k + ".bo"], res=x2)
            S.update(h1=h1, st1=st1, y1=y1, h2=h2, st2=st2, x2=x2, h3=h3, st3=st3, qkv=qkv, o=o, lse=lse)
            x = Act(dst, B, H, W, cout)
        return x

    def forward(self, x_nchw, t, ctx, mask, out_nchw):
        """eps = Unet(x, t, ctx, mask) for a batch where every sample has its own (timestep, class, mask) row."""
        self.prepare()
        w, ws = self.pw, self.ws
        B, _, H, W = x_nchw.shape
        D = self.D
        table = ws.get("tp_table", B, self.P, F32)
        emb_saved = ws.get("emb_saved", B, 11 * D, F32)
        ops.embed_time_class_train(t, ctx, mask, w["t.factor"], w["t.w1"], w["t.b1"], w["t.w2"], w["t.b2"], w["t.cls"],
                                   w["t.wp"], w["t.bp"], table, emb_saved)
        ch = list(self.arch["channels"])
        a0 = ws.get("in", B * H * W, ch[0])
        ops.conv3x3_small_cin(x_nchw, w["in.w"], w["in.b"], a0)
        self.saved["io"] = dict(x_nchw=x_nchw, t=t, ctx=ctx, mask=mask, B=B, H=H, W=W)
        x = Act(a0, B, H, W, ch[0])
        cats = []
        for i, (p, cin, cout) in enumerate(self.downs):
            cat = ws.get(f"cat{i}", x.M, 2 * cout)
            cats.append(cat)
            x = self._fwd_block(p, x, cout, table, final_dst=cat[:, cout:])
            planes = ws.get(f"s2d{i}", x.M, cout)
            ops.space_to_depth2(x.t, planes, x.B, x.H, x.W, cout)
            nxt = ws.get(f"dn{i}", x.M // 4, cout)
            ops.igemm([(planes, (4 * x.B, x.H // 2, x.W // 2), cout, 9)], w[f"down.{i}.w"], cout, nxt,
                      bias=w[f"down.{i}.b"], zero_pad_last=True, s2_batch=x.B)
            self.saved[f"down.{i}"] = dict(planes=planes, grid=x.grid, C=cout)
            x = Act(nxt, x.B, x.H // 2, x.W // 2, cout)
        for p, cin, cout in self.mids:
            x = self._fwd_block(p, x, cout, table)
        for i, (p, cin, cout) in enumerate(self.ups):
            c = x.C
            up = ws.get(f"up{i}", 4 * x.M, c)
            ops.upsample_nearest2x(x.t, up, x.B, x.H, x.W, c)
            cat = cats.pop()
            ops.igemm([(up, (x.B, 2 * x.H, 2 * x.W), c, 9)], w[f"up.{i}.w"], c, cat[:, :c], bias=w[f"up.{i}.b"])
            self.saved[f"up.{i}"] = dict(up=up, grid=(x.B, 2 * x.H, 2 * x.W), C=c)
            x = self._fwd_block(p, Act(cat, x.B, 2 * x.H, 2 * x.W, 2 * c), cout, table)
        h, st = ws.get("out.h", x.M, x.C), ws.get("out.st", x.B, 2 * self.G, F32)
        ops.groupnorm_silu_train(x.t, h, w["out.gw"], w["out.gb"], x.B, x.H * x.W, x.C, self.G, True, st)
        ops.conv3x3_small_cout(h, w["out.w"], w["out.b"], out_nchw)
        self.saved["out"] = dict(x=x, h=h, st=st)
        return out_nchw

    # ---------------------------------------------------------------------------------------------
    # backward
    # ---------------------------------------------------------------------------------------------
    def _scratch(self, Mmax, Cmax, B):
        ws = self.ws
        return dict(wg=ws.get("b.wgws", 1, 24 * 1024 * 1024, F32), ps=ws.get("b.ps", B, 4096, F32),
                    dgp=ws.get("b.dgp", B, 2048, F32), dbp=ws.get("b.dbp", B, 2048, F32))

    def _gn_bwd(self, x2d, dy, dx, gw, gb, st, gname_w, gname_b, B, HW, C, silu, add=None, colsum=None, bias1=None,
                bias2=None):
        """GroupNorm(+SiLU) backward; `colsum` (B, C view, any row stride) additionally receives the per-sample column
        sums of dx, whose batch sum goes to the bias gradients bias1 / bias2."""
        sc = self.sc
        dgp, dbp = sc["dgp"].view(-1)[:B * C].view(B, C), sc["dbp"].view(-1)[:B * C].view(B, C)
        if colsum is None and bias1 is not None:
            colsum = sc["ps"][:, :C]
        ops.groupnorm_silu_bwd(x2d, dy, dx, gw, gb, st, dgp, dbp, B, HW, C, self.G, silu, add=add, colsum_part=colsum)
        ops.groupnorm_bwd_finalize(dgp, dbp, B, C, self.g(gname_w), self.g(gname_b), colsum_part=colsum,
                                   g_bias1=None if bias1 is None else self.g(bias1),
                                   g_bias2=None if bias2 is None else self.g(bias2))

    def _bwd_block(self, p, dout, cout, dtable, final_name=None):
        """dout: (M, cout) bf16 gradient of the block output. Returns the (M, cin_0) gradient of the block input."""
        w, ws, sc = self.pw, self.ws, self.sc
        for l in reversed(range(self.L)):
            k, a = f"{p}.{l}", f"{p}.self_attns.{l}"
            S = self.saved[k]
            x = S["x"]
            B, H, W = x.grid
            M, HW, cin = x.M, x.H * x.W, x.C
            hd = cout // self.heads
            grid = x.grid
            # out_proj: out = o Wo^T + bo + x2
            ops.conv_wgrad(S["o"], (1, 1, M), cout, 1, dout, cout, self.g(a + ".out_proj.weight"), sc["wg"])
            ops.colsum(dout, B, HW, cout, sc["ps"], total=self.g(a + ".out_proj.bias"))
            do = ws.get(f"b.do.{M}", M, cout)
            ops.igemm([(dout, (1, 1, M), cout, 1)], w[k + ".wo"], cout, do, w_mn=True)
            # attention
            dqkv = ws.get(f"b.dqkv.{M}", M, 3 * cout)
            delta = ws.get(f"b.delta.{M}", M, self.heads, F32)
            dq32 = ws.get(f"b.dq32.{M}.{cout}", M, cout, F32) if HW > 128 else None
            ops.attention_bwd(S["qkv"], S["o"], do, S["lse"], delta, dqkv, dq32, M, HW, self.heads, hd)
            # QKV Linear (weights stacked q | k | v: adjacent in the flat gradient buffer)
            ops.conv_wgrad(S["h3"], (1, 1, M), cout, 1, dqkv, 3 * cout,
                           self.gspan(a + ".to_q.weight", a + ".to_v.weight"), sc["wg"])
            ops.colsum(dqkv, B, HW, 3 * cout, sc["ps"], total=self.gspan(a + ".to_q.bias", a + ".to_v.bias"))
            dh3 = ws.get(f"b.dh3.{M}", M, cout)
            ops.igemm([(dqkv, (1, 1, M), 3 * cout, 1)], w[k + ".wqkv"], cout, dh3, w_mn=True)
            # GN3 (no SiLU); the residual path adds dout
            dx2 = ws.get(f"b.dx2.{M}", M, cout)
            self._gn_bwd(S["x2"], dh3, dx2, w[k + ".g3w"], w[k + ".g3b"], S["st3"], a + ".groupnorm.weight",
                         a + ".groupnorm.bias", B, HW, cout, False, add=dout,
                         bias1=f"{p}.second_halfs.{l}.layers.2.bias", bias2=f"{p}.residuals.{l}.bias")
            # second conv3x3 (+) 1x1 skip conv (their shared bias gradient came out of the GN3 backward above)
            ops.conv_wgrad(S["h2"], grid, cout, 9, dx2, cout, self.g(f"{p}.second_halfs.{l}.layers.2.weight"), sc["wg"])
            ops.conv_wgrad(x.t, grid, cin, 1, dx2, cout, self.g(f"{p}.residuals.{l}.weight"), sc["wg"])
            dh2 = ws.get(f"b.dh2.{M}", M, cout)
            ops.igemm([(dx2, grid, cout, 9)], w[k + ".w2"][:, :9 * cout], cout, dh2, tap_offsets=self.DGRAD_TAPS, w_mn=True,
                      w_tap_ids=range(9))
            dy1 = ws.get(f"b.dy1.{M}", M, cout)
            off = self.tp_off[(p, l)]
            self._gn_bwd(S["y1"], dh2, dy1, w[k + ".g2w"], w[k + ".g2b"], S["st2"],
                         f"{p}.second_halfs.{l}.layers.0.weight", f"{p}.second_halfs.{l}.layers.0.bias", B, HW, cout, True,
                         colsum=dtable[:, off:off + cout], bias1=f"{p}.first_halfs.{l}.layers.2.bias")
            # first conv3x3 (bias gradient + per-sample time-bias gradient came out of the GN2 backward above)
            ops.conv_wgrad(S["h1"], grid, cin, 9, dy1, cout, self.g(f"{p}.first_halfs.{l}.layers.2.weight"), sc["wg"])
            dh1 = ws.get(f"b.dh1.{M}.{cin}", M, cin)
            ops.igemm([(dy1, grid, cout, 9)], w[k + ".w1"], cin, dh1, tap_offsets=self.DGRAD_TAPS, w_mn=True, w_tap_ids=range(9))
            dxa = ws.get(f"b.dxa.{M}.{cin}", M, cin)
            self._gn_bwd(x.t, dh1, dxa, w[k + ".g1w"], w[k + ".g1b"], S["st1"], f"{p}.first_halfs.{l}.layers.0.weight",
                         f"{p}.first_halfs.{l}.layers.0.bias", B, HW, cin, True)
            dx = ws.get(final_name if (l == 0 and final_name) else f"b.dx{l % 2}.{M}.{cin}", M, cin)
            ops.igemm([(dx2, grid, cout, 1)], w[k + ".w2"][:, 9 * cout:], cin, dx, res=dxa, w_mn=True)
            dout = dx
        return dout

    def backward(self, dout_nchw, on_stage_done=None):
        """dout_nchw: fp32 (B, z, H, W) gradient of eps. Fills self.flat_grad (every parameter's gradient is
        overwritten). on_stage_done(stage_index) is called after the kernels of each bucket stage were issued."""
        w, ws = self.pw, self.ws
        io = self.saved["io"]
        B, H, W = io["B"], io["H"], io["W"]
        self.sc = self._scratch(B * H * W, 1024, B)
        sc = self.sc
        dtable = ws.get("b.dtable", B, self.P, F32)
        nlev = len(self.downs)
        stage = 0

        def done():
            nonlocal stage
            if on_stage_done is not None:
                on_stage_done(stage)
            stage += 1

        # out_conv: GN + SiLU + conv 128 -> z
        so = self.saved["out"]
        x = so["x"]
        part = ws.get("b.edge_part", 1, max(B * (H // 2) * 3 * x.C * 9, B * (H // 2) * x.C * 27), F32)
        dh = ws.get("b.out.dh", x.M, x.C)
        ops.conv3x3_small_cout_bwd(so["h"], dout_nchw, w["out.w"], dh, self.g("out_conv.2.weight"),
                                   self.g("out_conv.2.bias"), part)
        d = ws.get("b.out.dx", x.M, x.C)
        self._gn_bwd(x.t, dh, d, w["out.gw"], w["out.gb"], so["st"], "out_conv.0.weight", "out_conv.0.bias", x.B,
                     x.H * x.W, x.C, True)
        dskips = {}
        for i in reversed(range(nlev)):
            p, cin, cout = self.ups[i]
            dcat = self._bwd_block(p, d, cout, dtable, final_name=f"b.dcat{i}")  # (M, 2c): [d up-conv out | d skip]
            su = self.saved[f"up.{i}"]
            c, (b_, h_, w_) = su["C"], su["grid"]
            dleft = dcat[:, :c]
            dskips[nlev - 1 - i] = dcat[:, c:]
            ops.conv_wgrad(su["up"], su["grid"], c, 9, dleft, c, self.g(f"upsamples.{i}.conv.weight"), sc["wg"])
            ops.colsum(dleft, b_, h_ * w_, c, sc["ps"], total=self.g(f"upsamples.{i}.conv.bias"))
            dup = ws.get(f"b.dup{i}", b_ * h_ * w_, c)
            ops.igemm([(dleft, su["grid"], c, 9)], w[f"up.{i}.w"], c, dup, tap_offsets=self.DGRAD_TAPS, w_mn=True,
                      w_tap_ids=range(9))
            d = ws.get(f"b.dlow{i}", b_ * h_ * w_ // 4, c)
            ops.sum2x2(dup, d, b_, h_ // 2, w_ // 2, c)
            done()
        for p, cin, cout in reversed(self.mids):
            d = self._bwd_block(p, d, cout, dtable)
            done()
        for i in reversed(range(nlev)):
            p, cin, cout = self.downs[i]
            sd_ = self.saved[f"down.{i}"]
            b_, h_, w_ = sd_["grid"]  # full-resolution grid of the block output
            oh, ow = h_ // 2, w_ // 2
            ops.zero_last_rowcol(d, b_, oh, ow, cout)
            ops.conv_wgrad(sd_["planes"], (4 * b_, oh, ow), cout, 9, d, cout, self.g(f"downsamples.{i}.down.weight"),
                           sc["wg"], s2_batch=b_)
            ops.colsum(d, b_, oh * ow, cout, sc["ps"], total=self.g(f"downsamples.{i}.down.bias"))
            dplanes = ws.get(f"b.dplanes{i}", b_ * h_ * w_, cout)
            rows_ = b_ * oh * ow
            for j, (pq, taps) in enumerate(sorted(ops._S2_PLANE_TAPS.items())):  # one launch per input parity plane
                ops.igemm([(d, (b_, oh, ow), cout, len(taps))], w[f"down.{i}.w"], cout, dplanes[j * rows_:(j + 1) * rows_],
                          tap_offsets=[(-(kh >> 1), -(kw >> 1)) for kh, kw in taps], w_mn=True,
                          w_tap_ids=[kh * 3 + kw for kh, kw in taps])
            dfull = ws.get(f"b.dfull{i}", b_ * h_ * w_, cout)
            ops.depth_to_space2(dplanes, dfull, b_, h_, w_, cout, add=dskips[i])
            d = self._bwd_block(p, dfull, cout, dtable)
            done()
        # in_conv
        ops.conv3x3_small_cin_wgrad(io["x_nchw"], d, self.g("in_conv.weight"), part)
        ops.colsum(d, B, H * W, d.shape[1], sc["ps"], total=self.g("in_conv.bias"))
        # embeddings: time projections (stacked), time MLP, class embedding
        D = self.D
        es = ws.get("b.emb_scratch", 1, ((self.P + 255) // 256) * B * 4 * D + 5 * B * D, F32)
        first_p, last_p = self.order[0], self.order[-1]
        g_wp = self.gspan(f"{first_p[0]}.time_projs.{first_p[1]}.1.weight", f"{last_p[0]}.time_projs.{last_p[1]}.1.weight")
        g_bp = self.gspan(f"{first_p[0]}.time_projs.{first_p[1]}.1.bias", f"{last_p[0]}.time_projs.{last_p[1]}.1.bias")
        g_cls = self.g("class_embedding.weight")
        if io["ctx"] is None:
            g_cls.zero_()
        ops.embed_time_class_bwd(dtable, io["ctx"], io["mask"], D, self.arch["num_classes"], w["t.w2"], w["t.wp"],
                                 ws.get("emb_saved", B, 11 * D, F32), self.g("time_embedding.embeddings.0.weight"),
                                 self.g("time_embedding.embeddings.0.bias"), self.g("time_embedding.embeddings.2.weight"),
                                 self.g("time_embedding.embeddings.2.bias"), g_cls if io["ctx"] is not None else None,
                                 g_wp, g_bp, es)
        done()
        return self.flat_grad
