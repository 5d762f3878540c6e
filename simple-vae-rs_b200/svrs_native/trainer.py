"""The fused optimisation step: models/base.py:103-107 of the reference
(zero_grad -> train_step -> backward -> clip_grad_norm_(1.0) -> Adam.step) as one kernel chain with no
host synchronisation, optionally replayed from a CUDA graph and data-parallel over NCCL.

Semantics kept from the reference:
  * loss = mse_x + kld_u + mse_y + kld_z, NLL terms are SUMS, KL terms batch MEANS
  * clipping covers module parameters only; gammas are a second Adam group, unclipped
  * Adam: lr 1e-4, betas (0.9, 0.999), eps 1e-8 unless the torch optimizer passed to fit() says otherwise
Data parallel (new capability, SURVEY 8.4 row e): every rank holds B_local samples; gradients are SUM
all-reduced with the KL upstream gradients pre-scaled by 1/world so the result equals the single-process
gradient on the global batch (up to per-rank BatchNorm statistics).
"""
from __future__ import annotations

import contextlib
import os
from typing import Dict, Optional

import numpy as np
import torch

from .elbo import elbo_forward
from .engine import BNOp, CondEngine, ConvOp, PatchBatch, VaeEngine, _dt, _p, _st
from .lib import ACT_SIGMOID, F32, lib
from .parallel import upstream_grad_scales


class _Nvtx:
    """NVTX ranges around the phases of the fused step (SVRS_NVTX=1): forward / elbo / backward / allreduce / optimizer show
    up as named ranges in nsys / ncu --nvtx timelines.  Host-side markers only; no effect on the enqueued work."""
    on = os.environ.get("SVRS_NVTX", "0") == "1"

    def __init__(self, name: str):
        self.name = name

    def __enter__(self):
        if _Nvtx.on:
            torch.cuda.nvtx.range_push("svrs/" + self.name)

    def __exit__(self, *exc):
        if _Nvtx.on:
            torch.cuda.nvtx.range_pop()
        return False


class _AdamCfg:
    def __init__(self, lr=1e-4, betas=(0.9, 0.999), eps=1e-8, max_norm=1.0):
        self.lr, self.b1, self.b2, self.eps, self.max_norm = lr, betas[0], betas[1], eps, max_norm

    def read(self, optimizer):
        if optimizer is None:
            return
        g = optimizer.param_groups[0]
        self.lr = float(g["lr"])
        self.b1, self.b2 = (float(b) for b in g.get("betas", (self.b1, self.b2)))
        self.eps = float(g.get("eps", self.eps))


class _FusedBase:
    n_gammas = 2

    def __init__(self, engine, optimizer=None, max_norm: float = 1.0, process_group=None, overlap: bool = True,
                 sync_bn: bool = False, world_override: Optional[int] = None):
        self.eng = engine
        self.rt = engine.rt
        self.cfg = _AdamCfg(max_norm=max_norm)
        self.cfg.read(optimizer)
        self.optimizer = optimizer
        self.pg = process_group
        self.world = 1
        self.rank = 0
        if process_group is not None or (torch.distributed.is_available() and torch.distributed.is_initialized()):
            self.world = torch.distributed.get_world_size(process_group)
            self.rank = torch.distributed.get_rank(process_group)
        if world_override is not None:          # world_override=1: a single-process step inside a distributed job
            self.world = world_override
        self.overlap = overlap and self.world > 1
        # sync_bn: BatchNorm statistics of the GLOBAL batch (per-layer all-reduce of the 2C sums, forward and backward) -
        # exact single-process parity of the data-parallel step (SURVEY 8.4 row e-ii); default = per-rank statistics
        self.sync_bn = bool(sync_bn) and self.world > 1
        self.keep_grad = False                  # keep a copy of the (all-reduced) flat gradient of the last step (checks)
        self.last_grad = None
        self.fused_tail = os.environ.get("SVRS_FUSED_TAIL", "1") != "0"
        self._adam_jobs = None
        self._norm_stream = None                # side stream of _early_sumsq
        self._norm_segs = []                    # flat-gradient ranges whose squared norm is already in normacc (this step)
        self._norm_zeroed = False
        self.early_norm = os.environ.get("SVRS_EARLY_NORM", "1") != "0"
        self.m = self.v = None
        self._flat_ptr = None
        self._graphs: Dict[tuple, dict] = {}
        self._comm_stream = None
        self.steps_done = 0

    # ---- state ----------------------------------------------------------------------------------
    def _ensure_state(self):
        rt = self.rt
        rt.ensure()
        rt.sync_bn, rt.pg, rt.world = self.sync_bn, self.pg, self.world
        store = rt.store
        fresh = self.m is None or self.m.device != store.flat.device or self.m.numel() != store.flat.numel()
        if not fresh and self._flat_ptr != store.flat.data_ptr():
            # the parameter store was re-flattened on the same device (a stand-alone block forward, load_state_dict(assign=True)
            # ...): captured graphs, job tables and packs point at freed buffers - drop them, keep the optimiser state
            self._flat_ptr = store.flat.data_ptr()
            self._graphs.clear()
            self._adam_jobs = None
            rt.packs_dirty = True
        if fresh:
            self._flat_ptr = store.flat.data_ptr()
            dev = store.flat.device
            self.m = torch.zeros_like(store.flat)
            self.v = torch.zeros_like(store.flat)
            self.step_ptr = torch.zeros(1, device=dev, dtype=torch.int64)
            self.normacc = torch.zeros(1, device=dev, dtype=torch.float64)
            self.gam = torch.ones(2, device=dev, dtype=torch.float32)
            self.gam_m = torch.zeros(2, device=dev, dtype=torch.float32)
            self.gam_v = torch.zeros(2, device=dev, dtype=torch.float32)
            self.dgam = store.tail[:2]       # gamma gradients live behind the flat gradient: one all-reduce carries both
            self.gout = torch.tensor(upstream_grad_scales(self.world), device=dev)
            self._load_gammas_from_model()
            self.eng.rng.step_ptr = self.step_ptr
            self._graphs.clear()
            self._adam_jobs = None
            rt.packs_dirty = True
            if self.world > 1:
                self._setup_comm(dev)

    def _setup_comm(self, dev):
        """Gradient exchange plumbing.  The all-reduce of the nets whose backward finishes first (decoders, prior heads,
        u_to_z: 18.4 M of the 20.6 M parameters) is issued while the encoders' backward pass is still running.  Round 1
        found that harmful at 8 GPUs (6.3 vs 4.1 ms/step): NCCL's CTAs queued behind five compute streams whose persistent
        CTAs hold ~200 KB of shared memory and all 512 TMEM columns per SM, and every rank waited for the slowest one.  Two
        changes make the overlap pay: the reduction runs on its OWN communicator with a bounded CTA count
        (ncclConfig maxCTAs = SVRS_NCCL_MAX_CTAS, default 8: ~5 % of the SMs) on a HIGH-PRIORITY stream, so its few CTAs
        are scheduled ahead of queued compute CTAs as soon as any SM frees up.  SVRS_AR_OVERLAP=0 restores the single
        all-reduce after the backward pass."""
        self._ar_overlap = os.environ.get("SVRS_AR_OVERLAP", "1") == "1"
        self._ar_two_buckets = os.environ.get("SVRS_AR_BUCKETS", "2") != "1"
        self._comm_stream = torch.cuda.Stream(device=dev, priority=-1)
        self._ar_pg = self.pg
        max_ctas = int(os.environ.get("SVRS_NCCL_MAX_CTAS", "8"))
        if self._ar_overlap and max_ctas > 0 and torch.distributed.get_backend(self.pg) == "nccl":
            opts = torch.distributed.ProcessGroupNCCL.Options(is_high_priority_stream=True)
            opts.config.max_ctas = max_ctas
            opts.config.min_ctas = min(max_ctas, int(os.environ.get("SVRS_NCCL_MIN_CTAS", "1")))
            ranks = torch.distributed.get_process_group_ranks(self.pg) if self.pg is not None else None
            self._ar_pg = torch.distributed.new_group(ranks=ranks, backend="nccl", pg_options=opts)

    def _gamma_attrs(self):
        raise NotImplementedError

    def _load_gammas_from_model(self):
        vals = [float(getattr(self.eng.model, a).detach()) for a in self._gamma_attrs()]
        while len(vals) < 2:
            vals.append(1.0)
        self.gam.copy_(torch.tensor(vals))

    def sync_to_model(self):
        """Write the device-resident gammas back to the model's (CPU) attributes; one D2H sync."""
        if self.m is None:
            return
        host = self.gam.detach().cpu()
        for i, a in enumerate(self._gamma_attrs()):
            getattr(self.eng.model, a).data.fill_(float(host[i]))

    # ---- optimizer tail ----------------------------------------------------------------------------
    def _adam_table(self, batch: int):
        """Device job table of svrs_adam_multi: one conv job per conv / convT weight (gradient layout as the wgrad kernel
        that takes the layer at this batch size leaves it) and one plain job per gap between them (biases, BatchNorm)."""
        rt, eng = self.rt, self.eng
        store = rt.store
        key = (store.flat.data_ptr(), rt.dtype, batch)
        if self._adam_jobs is None:
            self._adam_jobs = {}
        hit = self._adam_jobs.get(key)       # one table per batch size: captured graphs keep pointing at theirs
        if hit is not None:
            return hit
        rec = np.dtype([("off", "<i8"), ("p01", "<u8"), ("p10", "<u8"), ("d0", "<i4"), ("d1", "<i4"), ("kk", "<i4"),
                        ("layout", "<i4"), ("tile0", "<i4"), ("tiles_b", "<i4")])
        assert rec.itemsize == lib.adam_job_bytes()
        shapes = eng.conv_input_shapes(batch)            # id(op) -> (N, H, W) of the layer's input at this batch
        convs = sorted(((store.offsets[store._index[id(op.mod.weight)]], op) for net in rt.nets for op in net.ops
                        if isinstance(op, ConvOp)), key=lambda t: t[0])
        rows, tile0, cur = [], 0, 0
        tr = lib.adam_tile_rows()

        def plain(lo, hi):
            nonlocal tile0
            if hi > lo:
                rows.append((lo, 0, 0, hi - lo, 0, 1, 0, tile0, 0))
                tile0 += (hi - lo + 2047) // 2048

        for off, op in convs:
            plain(cur, off)
            w = op.mod.weight
            d0, d1 = w.shape[0], w.shape[1]
            n, h, wd = shapes[id(op)]
            if op.kind == "ct":
                layout = lib.convT2d_wgrad_layout(rt.dt, n, h, wd, op.cin, op.cout)
            else:
                layout = lib.conv2d_wgrad_layout(rt.dt, n, h, wd, op.cin, op.cout, 3 if op.kind == "c3" else 4)
            p01, p10 = (op.pack_f, op.pack_b) if op.kind == "ct" else (op.pack_b, op.pack_f)
            tc = lib.adam_tile_cols(op.kk)
            tiles_b = (d1 + tc - 1) // tc
            rows.append((off, p01.data_ptr(), p10.data_ptr(), d0, d1, op.kk, int(layout), tile0, tiles_b))
            tile0 += ((d0 + tr - 1) // tr) * tiles_b
            cur = off + w.numel()
        plain(cur, store.total)
        jobs = np.array(rows, dtype=rec)
        dev_jobs = torch.from_numpy(jobs.view(np.uint8).copy()).to(store.flat.device)
        self._adam_jobs[key] = (dev_jobs, len(rows), tile0)
        return self._adam_jobs[key]

    def _optim_tail(self, batch: int):
        rt, cfg, st = self.rt, self.cfg, _st()
        store = rt.store
        if self.keep_grad:
            self.last_grad = store.grad.clone()
        if self._norm_segs:
            # ranges already accumulated by _early_sumsq: join its stream and read only the gaps
            torch.cuda.current_stream().wait_stream(self._norm_stream)
            pos = 0
            for lo, hi in sorted(self._norm_segs) + [(store.grad.numel(), store.grad.numel())]:
                if lo > pos:
                    lib.sumsq(store.grad.data_ptr() + 4 * pos, lo - pos, _p(self.normacc), st)
                    rt.launches += 1
                pos = max(pos, hi)
            self._norm_segs = []
        else:
            if not self._norm_zeroed:
                lib.fill_zero(_p(self.normacc), 8, st)
                rt.launches += 1
            lib.sumsq(_p(store.grad), store.grad.numel(), _p(self.normacc), st)
            rt.launches += 1
        self._norm_zeroed = False
        # the two gammas (unclipped group, Q3) next to the big launch instead of behind it
        side = None
        if self.world == 1 and self.early_norm and rt.branch_streams:
            if self._norm_stream is None:
                self._norm_stream = torch.cuda.Stream(device=store.grad.device)
            side = self._norm_stream
            side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side) if side is not None else contextlib.nullcontext():
            lib.clip_adam(_p(self.gam), _p(self.dgam), _p(self.gam_m), _p(self.gam_v), self.n_gammas, None,
                          cfg.max_norm, 1.0, cfg.lr, cfg.b1, cfg.b2, cfg.eps, _p(self.step_ptr), _st())
        if rt.fused_grads:
            jobs, njobs, tiles = self._adam_table(batch)
            lib.adam_multi(_p(jobs), njobs, tiles, 16, _p(store.flat), _p(store.grad), _p(self.m), _p(self.v), rt.dt,
                           _p(self.normacc), cfg.max_norm, 1.0, cfg.lr, cfg.b1, cfg.b2, cfg.eps, _p(self.step_ptr), st)
        else:
            lib.clip_adam(_p(store.flat), _p(store.grad), _p(self.m), _p(self.v), store.flat.numel(), _p(self.normacc),
                          cfg.max_norm, 1.0, cfg.lr, cfg.b1, cfg.b2, cfg.eps, _p(self.step_ptr), st)
        rt.launches += 2
        if side is not None:
            torch.cuda.current_stream().wait_stream(side)
        if not rt.fused_grads:
            rt.packs_dirty = True
            rt.pack_weights()

    def _sync_bn_grad_fix(self):
        """sync_bn: every rank computed dgamma / dbeta from the GLOBAL sums, so the SUM all-reduce would count them `world`
        times - pre-scale them by 1/world."""
        if not (self.sync_bn and self.world > 1):
            return
        store = self.rt.store
        for net in self.rt.nets:
            for op in net.ops:
                if isinstance(op, BNOp):
                    store.grad_view(op.mod.weight).mul_(1.0 / self.world)
                    store.grad_view(op.mod.bias).mul_(1.0 / self.world)

    # ---- data-parallel gradient exchange ---------------------------------------------------------------
    # The flat fp32 gradient is SUM-all-reduced (NCCL).  With SVRS_AR_OVERLAP=1 the part that belongs to the nets whose
    # backward finishes first (decoders, prior heads, u_to_z: CondEngine.backward phase 1) is unpacked and reduced on a
    # communication stream while the encoders' backward pass is still running; the rest follows after the backward pass.
    def _net_segments(self, nets):
        """Merged [lo, hi) element ranges of the flat buffers covered by the parameters of `nets`."""
        store = self.rt.store
        A = store.ALIGN
        spans = []
        for net in nets:
            for op in net.ops:
                for p in op.mod.parameters(recurse=False):
                    i = store._index[id(p)]
                    lo = store.offsets[i]
                    spans.append((lo, lo + (p.numel() + A - 1) // A * A))
        spans.sort()
        merged = []
        for lo, hi in spans:
            if merged and lo <= merged[-1][1]:
                merged[-1][1] = max(merged[-1][1], hi)
            else:
                merged.append([lo, hi])
        return [(lo, min(hi, store.total)) for lo, hi in merged]

    def _early_allreduce(self, nets):
        """Called by CondEngine.backward when a group of sub-networks is complete - first the prior heads + u_to_z (15 M
        parameters, done a few hundred microseconds into the backward pass), then the decoders: each group is ONE contiguous
        range of the flat gradient buffer (ParamStore lays the encoders out last, the heads right before them) and ONE
        all-reduce on the communication stream, overlapped with whatever backward work is still running."""
        rt = self.rt
        store = rt.store
        segs = self._net_segments(nets)
        assert len(segs) == 1 and segs[0][1] <= store.early_end, (segs, store.early_end)
        lo, hi = segs[0]
        comm = self._comm_stream
        comm.wait_stream(torch.cuda.current_stream())
        if rt._side_busy:
            for side in rt.wgrad_streams():            # weight gradients queued so far (a superset of those of `nets`)
                comm.wait_stream(side)
        with torch.cuda.stream(comm):
            if not rt.fused_grads:
                rt.unpack_nets(nets)
            torch.distributed.all_reduce(store.grad_full[lo:hi], group=self._ar_pg)
        self._early_segs = (self._early_segs or []) + segs

    def _early_sumsq(self, nets):
        """Single-process counterpart of _early_allreduce: the squared norm of a finished group's gradients (one contiguous
        range of the flat buffer) is accumulated on a side stream while the rest of the backward pass runs, so the serial
        optimiser tail only has the encoders' range left to read (sumsq over all 82 MB: 20 us of the tail; the late range: ~4)."""
        rt = self.rt
        store = rt.store
        segs = self._net_segments(nets)
        assert len(segs) == 1 and segs[0][1] <= store.early_end, (segs, store.early_end)
        lo, hi = segs[0]
        if self._norm_stream is None:
            self._norm_stream = torch.cuda.Stream(device=store.grad.device)
        side = self._norm_stream
        side.wait_stream(torch.cuda.current_stream())
        if rt._side_busy:
            for ws in rt.wgrad_streams():              # weight gradients queued so far (a superset of those of `nets`)
                side.wait_stream(ws)
        with torch.cuda.stream(side):
            lib.sumsq(store.grad.data_ptr() + 4 * lo, hi - lo, _p(self.normacc), _st())
        rt.launches += 1
        self._norm_segs.append((lo, hi))

    def _allreduce_all(self):
        """The rest of the exchange after the backward pass: the encoders' range plus the gamma gradients in the tail of the
        same buffer (one call), or the whole buffer when nothing was reduced early."""
        if self.world == 1:
            return
        self._sync_bn_grad_fix()
        store = self.rt.store
        early = getattr(self, "_early_segs", None)
        self._early_segs = None
        if not early:
            torch.distributed.all_reduce(store.grad_full, group=self.pg)
        else:
            torch.distributed.all_reduce(store.grad_full[store.early_end:], group=self.pg)
            torch.cuda.current_stream().wait_stream(self._comm_stream)

    # ---- public ------------------------------------------------------------------------------------
    def grad_norm(self) -> torch.Tensor:
        return self.normacc.sqrt().float()

    def _step_impl(self, *inputs):
        raise NotImplementedError

    def step(self, *inputs, use_graph: bool = False) -> torch.Tensor:
        """One optimisation step on device-resident inputs (NCHW tensors as the reference's loaders yield them, or
        PatchBatch objects from dataset.grid_patch_pair).  Returns the device tensor
        [mse_x, kld_u, mse_y, kld_z, loss] (Cond) / [mse, kld, 0, 0, loss] (VAE) of THIS rank's batch."""
        self._ensure_state()
        self.cfg.read(self.optimizer)
        self.steps_done += 1
        if not use_graph:
            return self._step_impl(*inputs)
        flat, rebuild = _flatten_inputs(inputs)
        key = tuple((tuple(t.shape), t.dtype) for t in flat) + (self.cfg.lr, self.rt.dtype, "step")
        g = self._graphs.get(key)
        if g is None:
            if self.steps_done == 1:
                # first step runs eagerly (module loading, allocator warm-up); capture on the next one
                return self._step_impl(*inputs)
            static_in = [torch.empty_like(t) for t in flat]
            for s_, t in zip(static_in, flat):
                s_.copy_(t)
            if self.fused_tail:
                self._adam_table(int(flat[0].shape[0]))     # host -> device table upload must not happen under capture
            graph = torch.cuda.CUDAGraph()
            torch.cuda.synchronize()
            l0 = self.rt.launches
            with torch.cuda.graph(graph):
                out = self._step_impl(*rebuild(static_in))
            g = dict(graph=graph, inputs=static_in, out=out, launches=self.rt.launches - l0)
            self._graphs[key] = g
        else:
            for s_, t in zip(g["inputs"], flat):
                if s_.data_ptr() != t.data_ptr():
                    s_.copy_(t, non_blocking=True)
            self.rt.add_replayed(g["launches"])
        g["graph"].replay()
        return g["out"]

    def step_tiles(self, *tiles, patch_size: int, use_graph: bool = True) -> torch.Tensor:
        """One optimisation step straight from device-resident tiles (grid mode): the patch gather + normalise kernels
        (dataset.grid_patch_pair) are part of the captured graph.  tiles = (hr,) for the VAE, (hr, lr) for Cond_SRVAE."""
        from dataset import grid_patch_pair
        self._ensure_state()
        self.cfg.read(self.optimizer)
        self.steps_done += 1
        sizes = [patch_size, patch_size // 2][:len(tiles)]

        def run(ts):
            return self._step_impl(*[grid_patch_pair(t, p, self.rt.dtype, rt=self.rt) for t, p in zip(ts, sizes)])

        if not use_graph or self.steps_done == 1:
            return run(tiles)
        key = tuple((tuple(t.shape), t.dtype) for t in tiles) + (self.cfg.lr, self.rt.dtype, "tiles", patch_size)
        g = self._graphs.get(key)
        if g is None:
            static_in = [t.clone() for t in tiles]
            if self.fused_tail:
                self._adam_table(int(tiles[0].shape[0]) * (tiles[0].shape[-1] // patch_size) ** 2)
            graph = torch.cuda.CUDAGraph()
            torch.cuda.synchronize()
            l0 = self.rt.launches
            with torch.cuda.graph(graph):
                out = run(static_in)
            g = dict(graph=graph, inputs=static_in, out=out, launches=self.rt.launches - l0)
            self._graphs[key] = g
        else:
            for s_, t in zip(g["inputs"], tiles):
                if s_.data_ptr() != t.data_ptr():
                    s_.copy_(t, non_blocking=True)
            self.rt.add_replayed(g["launches"])
        g["graph"].replay()
        return g["out"]

    def static_inputs(self, *like):
        """Device buffers a caller may fill directly (avoids the extra device copy before a replay)."""
        flat, _ = _flatten_inputs(like)
        key = tuple((tuple(t.shape), t.dtype) for t in flat) + (self.cfg.lr, self.rt.dtype, "step")
        g = self._graphs.get(key)
        return None if g is None else g["inputs"]


def _flatten_inputs(inputs):
    """(tensors | PatchBatch | None ...) -> flat tensor list + a function that rebuilds the argument tuple from a list of
    same-shaped tensors (CUDA-graph static inputs)."""
    flat, spec = [], []
    for a in inputs:
        if isinstance(a, PatchBatch):
            ts = a.tensors()
            spec.append(("pb", len(ts)))
            flat.extend(ts)
        elif a is None:
            spec.append(("none", 0))
        else:
            spec.append(("t", 1))
            flat.append(a)

    def rebuild(ts):
        out, i = [], 0
        for kind, n in spec:
            if kind == "pb":
                out.append(PatchBatch(ts[i], ts[i + n - 1]))
                i += n
            elif kind == "none":
                out.append(None)
            else:
                out.append(ts[i])
                i += 1
        return tuple(out)

    return flat, rebuild


class FusedCondTrainer(_FusedBase):
    def __init__(self, model, optimizer=None, compute_dtype=None, **kw):
        eng = model._engine(compute_dtype)
        super().__init__(eng, optimizer, **kw)

    def _gamma_attrs(self):
        return ("gammax", "gammay")

    def _step_impl(self, x, y, eps_u=None, eps_z=None):
        eng, rt, st = self.eng, self.rt, _st()
        Wz, Wu = eng.Wz, eng.Wu
        lib.step_begin(_p(self.step_ptr), _p(self.normacc), st)   # norm accumulator zeroed off the optimiser tail; _early_sumsq adds into it
        self._norm_zeroed, self._norm_segs = True, []
        rt.fused_grads = self.fused_tail
        rt.zero_grads(with_scratch=True, deferred=True)
        rt.scratch_prezeroed = True
        try:
            with _Nvtx("forward"):
                outs, ctx = eng.forward(x, y, eps_u, eps_z, training=True, save=True, repack=False, fused_io=True)
        finally:
            rt.scratch_prezeroed = False
        xb, yb = outs["xb"], outs["yb"]
        B = xb.B
        enc_u, enc_z = outs["enc_u"], outs["enc_z"]
        x_hat, y_hat, mu3, lv3 = outs["x_hat_nhwc"], outs["y_hat_nhwc"], outs["mu3"], outs["lv3"]
        mu_u, lv_u = enc_u[:, :Wu], enc_u[:, Wu:]
        mu_z, lv_z = enc_z[:, :Wz], enc_z[:, Wz:]
        # ELBO on NHWC operands: reconstruction (fp32, straight from the sigmoid tail) vs the fp32 NHWC target
        terms, acc = elbo_forward(x_hat, xb.f32, y_hat, yb.f32, mu_u, lv_u, mu_z, lv_z, mu3, lv3, self.gam, B)
        # gradients wrt the decoders' PRE-sigmoid outputs, NHWC in the compute dtype (no layout / activation kernels)
        d_xhat = torch.empty(x_hat.shape, device=x_hat.device, dtype=rt.dtype)
        d_yhat = torch.empty(y_hat.shape, device=y_hat.device, dtype=rt.dtype)
        d_enc_u, d_enc_z = torch.empty_like(enc_u), torch.empty_like(enc_z)
        d_mu3, d_lv3 = torch.empty_like(mu3), torch.empty_like(lv3)
        lib.elbo_bwd(_p(x_hat), _p(xb.f32), F32, F32, x_hat.numel(), _p(d_xhat), rt.dt,
                     _p(y_hat), _p(yb.f32), F32, F32, y_hat.numel(), _p(d_yhat), rt.dt,
                     _p(mu_u), _p(lv_u), 2 * Wu, Wu, d_enc_u.data_ptr(), d_enc_u.data_ptr() + 4 * Wu, 2 * Wu,
                     _p(mu_z), _p(lv_z), 2 * Wz, d_enc_z.data_ptr(), d_enc_z.data_ptr() + 4 * Wz, 2 * Wz,
                     _p(mu3), _p(lv3), Wz, Wz, _p(d_mu3), _p(d_lv3), Wz,
                     B, _p(acc), _p(self.gam), _p(self.gout), _p(self.dgam), ACT_SIGMOID, st)
        rt.launches += 3
        rt.scratch_prezeroed = True
        self._early_segs = None
        if self.world > 1 and getattr(self, "_ar_overlap", False) and not self.sync_bn:
            rt.after_phase1 = self._early_allreduce      # only while the fused step's backward runs
            rt.after_heads = self._early_allreduce if self._ar_two_buckets else None
        elif self.world == 1 and self.early_norm and self.fused_tail and rt.branch_streams:     # (off in single-stream profiling passes)
            rt.after_phase1 = self._early_sumsq
            rt.after_heads = self._early_sumsq
        rt.join_zero_grads()
        try:
            with _Nvtx("backward"):
                eng.backward(ctx, d_xhat, d_yhat, d_enc_z, d_enc_u, d_mu3, d_lv3, fused_io=True)
        finally:
            rt.scratch_prezeroed = False
            rt.after_phase1 = None
            rt.after_heads = None
        with _Nvtx("allreduce"):
            self._allreduce_all()
        with _Nvtx("optimizer"):
            self._optim_tail(B)
        rt.fused_grads = False
        return terms


class FusedVaeTrainer(_FusedBase):
    n_gammas = 1

    def __init__(self, model, optimizer=None, compute_dtype=None, **kw):
        eng = model._engine(compute_dtype)
        super().__init__(eng, optimizer, **kw)

    def _gamma_attrs(self):
        return ("gamma",)

    def _step_impl(self, x, eps=None):
        eng, rt, st = self.eng, self.rt, _st()
        Wd = eng.Wd
        lib.step_begin(_p(self.step_ptr), _p(self.normacc), st)
        self._norm_zeroed, self._norm_segs = True, []
        rt.fused_grads = self.fused_tail
        rt.zero_grads(with_scratch=True, deferred=True)
        rt.scratch_prezeroed = True
        try:
            with _Nvtx("forward"):
                outs, ctx = eng.forward(x, eps, training=True, save=True, repack=False, fused_io=True)
        finally:
            rt.scratch_prezeroed = False
        xb = outs["xb"]
        B = xb.B
        enc, x_hat = outs["enc"], outs["x_hat_nhwc"]
        mu, lv = enc[:, :Wd], enc[:, Wd:]
        terms, acc = elbo_forward(x_hat, xb.f32, None, None, mu, lv, None, None, None, None, self.gam, B)
        d_xhat = torch.empty(x_hat.shape, device=x_hat.device, dtype=rt.dtype)
        d_enc = torch.empty_like(enc)
        lib.elbo_bwd(_p(x_hat), _p(xb.f32), F32, F32, x_hat.numel(), _p(d_xhat), rt.dt,
                     None, None, F32, F32, 0, None, F32,
                     _p(mu), _p(lv), 2 * Wd, Wd, d_enc.data_ptr(), d_enc.data_ptr() + 4 * Wd, 2 * Wd,
                     None, None, 0, None, None, 0,
                     None, None, 0, 0, None, None, 0,
                     B, _p(acc), _p(self.gam), _p(self.gout), _p(self.dgam), ACT_SIGMOID, st)
        rt.launches += 3
        rt.scratch_prezeroed = True
        rt.join_zero_grads()
        try:
            with _Nvtx("backward"):
                eng.backward(ctx, d_xhat, d_enc, fused_io=True)
        finally:
            rt.scratch_prezeroed = False
        with _Nvtx("allreduce"):
            self._allreduce_all()
        with _Nvtx("optimizer"):
            self._optim_tail(B)
        rt.fused_grads = False
        return terms
