"""CFG DDPM sampling loop (reference modules/diffusion.py:46-56) as one CUDA-graph replay per step.

A step = one batch-doubled UNet pass (rows [0, N) conditional, rows [N, 2N) unconditional — the reference's two
calls, diffusion.py:53-54) + the fused guidance-mix / posterior kernel that advances x_t in place. The per-step
embedding work is done for the (num_classes + 1) distinct (timestep, class) rows only; samples index into that
table. Nothing inside the step synchronises with the host.
"""
from __future__ import annotations

import torch

from . import native, ops


class CfgSampler:
    def __init__(self, unet, scheduler, labels: torch.Tensor, cfg_scales: torch.Tensor, latent_shape, use_graph=True,
                 kind: str = "ddpm", eta: float = 0.0, clamp_x0: bool = False):
        """kind="ddpm": the reference's 1-step ancestral update (components.py:405-424). kind="ddim": the strided
        update of idf_cfg_ddim_step - `steps` passed to run() may then be any decreasing subset of
        the schedule; eta = 0 is deterministic, eta = 1 ancestral."""
        if kind not in ("ddpm", "ddim"):
            raise ValueError(f"CfgSampler: unknown kind {kind!r}")
        self.kind, self.eta, self.clamp_x0 = kind, float(eta), clamp_x0
        dev = labels.device
        self.unet, self.sched = unet, scheduler._on(dev)
        self.N = N = labels.shape[0]
        self.shape = (N, *latent_shape)
        K = unet.num_classes
        self.engine = unet.engine(2 * N, latent_shape[1], latent_shape[2])
        # embedding rows: classes 0..K-1 (conditional) and one masked row (unconditional)
        self.t_rows = torch.zeros(K + 1, device=dev, dtype=torch.int64)
        self.t_prev = torch.full((1,), -1, device=dev, dtype=torch.int64)
        self.ctx_rows = torch.cat([torch.arange(K, device=dev), torch.zeros(1, device=dev, dtype=torch.int64)])
        self.mask_rows = torch.cat([torch.ones(K, device=dev), torch.zeros(1, device=dev)]).to(torch.float32)
        self.row_idx = torch.cat([labels.to(torch.int32), torch.full((N,), K, device=dev, dtype=torch.int32)])
        # The embedding path depends only on (timestep, class row): its output for EVERY timestep of the schedule is
        # computed once per weight version ((T*(K+1), P) fp32, 78 MB for T = 1000) and the step indexes into it, so no
        # embedding kernel runs inside the step.
        self.rows_per_t = K + 1
        self.base_idx = self.row_idx.clone()
        self.table_all = None
        self._epoch = -1
        self.cfg = cfg_scales.to(device=dev, dtype=torch.float32).contiguous()
        self.xx = torch.zeros(N, *latent_shape, device=dev, dtype=torch.float32)        # x_t (fp32 state)
        self.eps = torch.empty(2 * N, *latent_shape, device=dev, dtype=torch.float32)  # [eps_cond ; eps_uncond]
        self.z = torch.zeros(N, *latent_shape, device=dev, dtype=torch.float32)
        self.use_graph = use_graph
        self.graph = None
        self.launches_per_step = None

    def _ensure_table(self):
        """Cached on the engine (shared by every sampler of the same batch size), refreshed when the weights change.
        Both the packed weights (engine.prepare) and the table are rewritten IN PLACE, so a graph captured earlier
        keeps reading valid, current data."""
        eng = self.engine
        eng.prepare()
        self._epoch = getattr(self.unet, "_weights_epoch", 0)
        T, R = self.sched.num_steps, self.rows_per_t
        cache = eng.__dict__.setdefault("cfg_tables", {})
        entry = cache.get((T, R))
        if entry is not None and entry[0] == eng.packed.key:
            self.table_all = entry[1]
            return
        dev = self.row_idx.device
        table = entry[1] if entry is not None else torch.empty(T * R, eng.P, device=dev, dtype=torch.float32)
        t_all = torch.arange(T, device=dev, dtype=torch.int64).repeat_interleave(R)
        eng.embedding_table(t_all, self.ctx_rows.repeat(T), self.mask_rows.repeat(T), out=table)
        cache[(T, R)] = (eng.packed.key, table)
        self.table_all = table

    def _step(self):
        N = self.N
        native.call("idf_rowidx_from_timestep", self.base_idx.data_ptr(), self.t_rows.data_ptr(), self.rows_per_t,
                    self.row_idx.data_ptr(), 2 * N)
        self.engine.run(self.xx, self.t_rows, self.ctx_rows, self.mask_rows, self.row_idx, self.eps, dup_input=True,
                        table=self.table_all)
        if self.kind == "ddim":
            ops.cfg_ddim_step(self.xx, self.eps[:N], self.eps[N:], self.z, self.cfg, self.t_rows[:1], self.t_prev,
                              self.sched, self.xx, eta=self.eta, clamp_x0=self.clamp_x0)
        else:
            ops.cfg_posterior_step(self.xx, self.eps[:N], self.eps[N:], self.z, self.cfg, self.t_rows[:1], self.sched,
                                   self.xx)

    def _ensure_graph(self):
        if self.graph is not None or not self.use_graph:
            return
        keep = self.xx.clone()
        s = torch.cuda.Stream()
        s.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(s):
            self._step()  # warm-up: allocates workspaces, packs weights, sets kernel attributes
        torch.cuda.current_stream().wait_stream(s)
        torch.cuda.synchronize()
        before = native.launch_count
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g):
            self._step()
        self.launches_per_step = native.launch_count - before
        self.graph = g
        self.xx.copy_(keep)

    def set_conditioning(self, labels: torch.Tensor, cfg_scales: torch.Tensor):
        """New class labels / guidance scales for the same batch size, written IN PLACE: the captured graph (which
        reads these device tensors) serves every micro-batch of a sharded run."""
        if labels.shape[0] != self.N or cfg_scales.shape[0] != self.N:
            raise ValueError(f"CfgSampler.set_conditioning: expected {self.N} labels and scales")
        self.base_idx[: self.N].copy_(labels.to(torch.int32))
        self.cfg.copy_(cfg_scales.to(torch.float32))

    def set_latent(self, x_T: torch.Tensor):
        self._ensure_table()  # start of a sampling run: full staleness check of the weights (load_state_dict, optimizer)
        self.xx.copy_(x_T)

    @property
    def latent(self) -> torch.Tensor:
        return self.xx

    def step(self, i: int, noise: torch.Tensor | None = None, i_prev: int | None = None):
        """Advance x_i -> x_{i-1} (kind="ddim": x_i -> x_{i_prev}; i_prev < 0 or None = the final step). `noise`
        injects the step's N(0,1) draw; None draws it from the global CUDA generator exactly where the reference does
        (components.py:423: randn_like(xt), skipped at i == 0; the deterministic eta = 0 sampler draws nothing)."""
        if self.table_all is None or self._epoch != getattr(self.unet, "_weights_epoch", 0):
            self._ensure_table()
        self._ensure_graph()
        self.t_rows.fill_(i)
        ip = -1 if i_prev is None else int(i_prev)
        if self.kind == "ddim":
            self.t_prev.fill_(ip)
        if (i > 0 and self.kind == "ddpm") or (self.kind == "ddim" and self.eta > 0.0 and ip >= 0):
            if noise is None:
                self.z.normal_()
            else:
                self.z.copy_(noise)
        if self.graph is not None:
            self.graph.replay()
        else:
            self._step()

    def run(self, x_T: torch.Tensor, steps=None, noises=None) -> torch.Tensor:
        self.set_latent(x_T)
        steps = list(reversed(range(self.sched.num_steps))) if steps is None else list(steps)
        for k, i in enumerate(steps):
            self.step(i, None if noises is None else noises[k], steps[k + 1] if k + 1 < len(steps) else -1)
        return self.latent
