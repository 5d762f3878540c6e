"""Host-side runtime for the (Cond_)SRVAE training step on B200.

Everything arithmetic happens in libsvrs_b200.so (hand-written sm_100a CUDA, bound through the C ABI in
include/svrs_b200.h).  PyTorch is used for device memory (tensors as buffers), streams and - in
`parallel.py` - torch.distributed.  There is no eager / CPU fallback: a model that is not on a CUDA device
cannot run.

Layout: activations are NHWC in the compute dtype (fp32 or bf16); latents cross sub-network boundaries in
the reference's NCHW-flat fp32 order, which is also what the Python API returns.
"""
from __future__ import annotations

import os
from dataclasses import dataclass, field
from typing import Dict, List, Optional, Sequence, Tuple

import torch
import torch.nn as nn

from .lib import ACT_HARDTANH7, ACT_NONE, ACT_SIGMOID, BF16, F32, SvrsError, SvrsUnsupported, lib

BN_EPS_DEFAULT = 1e-5
FUSE_BN = os.environ.get("SVRS_FUSE_BN", "1") != "0"            # BatchNorm statistics in the producing conv's epilogue
FUSE_BN_MAX_MB = float(os.environ.get("SVRS_FUSE_BN_MAX_MB", "5"))   # fuse only for conv outputs up to this size (0 = no limit): on the big maps the extra epilogue costs what the streaming bn_stats pass does (measured 2.856 -> 2.827 ms/step)


@dataclass
class PatchBatch:
    """A batch of normalised patches as the fused step consumes it: NHWC fp32 (the NLL target, and the conv operand in fp32
    mode) plus the NHWC compute-dtype copy that feeds the first conv layer (the same tensor in fp32 mode).  Produced in one
    launch by dataset.grid_patch_pair / svrs_patch_gather_normalize, or from a plain NCHW tensor by Runtime.patch_batch."""
    f32: torch.Tensor
    op: torch.Tensor

    @property
    def B(self) -> int:
        return self.f32.shape[0]

    def tensors(self):
        return [self.f32] if self.op is self.f32 else [self.f32, self.op]


def _dt(dtype: torch.dtype) -> int:
    if dtype == torch.float32:
        return F32
    if dtype == torch.bfloat16:
        return BF16
    raise SvrsError(f"unsupported compute dtype {dtype}")


def _p(t: Optional[torch.Tensor]):
    return None if t is None else t.data_ptr()


BN_REPLICAS = 8      # SVRS_BN_REPLICAS of include/svrs_b200.h (checked against the header in tests/test_cpu_abi_and_geometry.py)


def _st() -> int:
    return torch.cuda.current_stream().cuda_stream


def _require_cuda(t: torch.Tensor, what: str):
    if not t.is_cuda:
        raise SvrsError(
            f"{what} is on {t.device}: the svrs_b200 path runs hand-written sm_100a kernels only and has no "
            f"CPU fallback - move the model and its inputs to a CUDA device.")


# ------------------------------------------------------------------------------------------------
# parameter storage: one flat fp32 buffer for parameters, one for gradients (clip / Adam / all-reduce
# then run over a single contiguous range).  nn.Parameter objects are kept (state_dict keys, optimizer
# references) and re-pointed at views of the flat buffer.
# ------------------------------------------------------------------------------------------------
class ParamStore:
    ALIGN = 4  # floats -> 16-byte aligned views

    def __init__(self, module: nn.Module, late_prefixes: Tuple[str, ...] = ()):
        """late_prefixes: parameter-name prefixes of the sub-networks whose backward pass finishes LAST (the encoders).  They
        are laid out at the END of the flat buffers, so the gradients that are complete early (everything else) form one
        contiguous range [0, early_end) that the data-parallel trainer can all-reduce while the late nets' backward is still
        running, and the late range [early_end, total) plus the small tail (gamma gradients) is a second single call.
        state_dict / named_parameters order is the module's own and unaffected."""
        self.module = module
        named = list(module.named_parameters())
        late = [i for i, (n, _) in enumerate(named) if n.startswith(tuple(late_prefixes))] if late_prefixes else []
        order = [i for i in range(len(named)) if i not in set(late)] + late
        self.params: List[nn.Parameter] = [named[i][1] for i in order]
        self.names: List[str] = [named[i][0] for i in order]
        self.offsets: List[int] = []
        off = 0
        self.early_end = None
        n_early = len(order) - len(late)
        for j, p in enumerate(self.params):
            if j == n_early:
                self.early_end = off
            self.offsets.append(off)
            off += (p.numel() + self.ALIGN - 1) // self.ALIGN * self.ALIGN
        self.total = off
        if self.early_end is None:
            self.early_end = off
        self.TAIL = 4            # floats after the parameters' gradients: scalar gradients that travel with the all-reduce
        self.flat: Optional[torch.Tensor] = None
        self.grad: Optional[torch.Tensor] = None
        self.gpack: Optional[torch.Tensor] = None
        self._index = {id(p): i for i, p in enumerate(self.params)}

    def valid(self) -> bool:
        if self.flat is None or not self.params:
            return self.flat is not None
        base = self.flat.data_ptr()
        for p, off in zip(self.params, self.offsets):       # every parameter: a stand-alone block forward, .to() or
            if p.data.data_ptr() != base + 4 * off:          # load_state_dict(assign=True) may have re-pointed any of them
                return False
        return True

    def ensure(self) -> bool:
        """(Re)flatten if the module's parameters moved (model.to(), fresh construction). True if rebuilt."""
        if self.valid():
            return False
        dev = self.params[0].device
        _require_cuda(self.params[0], "model")
        flat = torch.zeros(self.total, device=dev, dtype=torch.float32)
        with torch.no_grad():
            for p, off in zip(self.params, self.offsets):
                v = flat[off:off + p.numel()].view(p.shape)
                v.copy_(p.data.to(torch.float32))
                p.data = v
        self.flat = flat
        self.grad_full = torch.zeros(self.total + self.TAIL, device=dev, dtype=torch.float32)
        self.grad = self.grad_full[:self.total]
        self.tail = self.grad_full[self.total:]
        self.gpack = torch.zeros_like(flat)     # per-tap packed scratch for the tensor-core wgrad kernels
        return True

    def grad_view(self, p: nn.Parameter) -> torch.Tensor:
        i = self._index[id(p)]
        off = self.offsets[i]
        return self.grad[off:off + p.numel()].view(p.shape)

    def grad_ptr(self, p: nn.Parameter) -> int:
        return self.grad.data_ptr() + 4 * self.offsets[self._index[id(p)]]

    def gpack_ptr(self, p: nn.Parameter) -> int:
        return self.gpack.data_ptr() + 4 * self.offsets[self._index[id(p)]]


# ------------------------------------------------------------------------------------------------
# network plans
# ------------------------------------------------------------------------------------------------
@dataclass
class ConvOp:
    kind: str                  # "c3" (k3 s1 p1), "c4" (k4 s2 p1), "ct" (transposed k4 s2 p1)
    mod: nn.Module
    cin: int
    cout: int
    act: int = ACT_NONE
    pack_f: Optional[torch.Tensor] = None   # KN pack for fprop
    pack_b: Optional[torch.Tensor] = None   # KN pack for dgrad
    fuse_bn: Optional[bool] = None          # producer epilogue can take the BatchNorm statistics (None = not probed yet)
    fuse_head: Optional[bool] = None        # epilogue can write the fp32 NCHW-flat head output
    fuse_f32: Optional[bool] = None         # kernel can store fp32 from bf16 operands (sigmoid tail)

    @property
    def kk(self) -> int:
        return 9 if self.kind == "c3" else 16

    def out_hw(self, h: int, w: int) -> Tuple[int, int]:
        if self.kind == "c3":
            return h, w
        if self.kind == "c4":
            return h // 2, w // 2
        return 2 * h, 2 * w


@dataclass
class BNOp:
    mod: nn.BatchNorm2d
    relu: bool = True
    sums_f: Optional[torch.Tensor] = None   # double[2C] scratch, forward stats
    sums_b: Optional[torch.Tensor] = None   # double[2C] scratch, backward sums


@dataclass
class Net:
    name: str
    ops: list = field(default_factory=list)

    @property
    def cin(self) -> int:
        return self.ops[0].cin

    @property
    def cout(self) -> int:
        for op in reversed(self.ops):
            if isinstance(op, ConvOp):
                return op.cout
        raise RuntimeError("empty net")


def plan_sequential(name: str, seq: nn.Sequential) -> Net:
    """Translate one of the reference's nn.Sequential sub-networks (cond_vae.py:27-231, vae.py:36-85)
    into a list of kernel ops.  Flatten/Unflatten are geometry handled by the caller."""
    net = Net(name)
    for m in seq:
        cls = m.__class__.__name__
        if cls == "down_block":          # layers.py:217-256
            net.ops.append(ConvOp("c3", m.conv, m.conv.in_channels, m.conv.out_channels))
            net.ops.append(ConvOp("c4", m.downsample, m.downsample.in_channels, m.downsample.out_channels))
            if m.with_bn:
                net.ops.append(BNOp(m.bn, relu=m.with_relu))
            elif m.with_relu:
                raise SvrsError("down_block(with_bn=False, with_relu=True) is not used by any model")
        elif cls == "up_block":          # layers.py:259-297
            net.ops.append(ConvOp("c3", m.conv, m.conv.in_channels, m.conv.out_channels))
            net.ops.append(ConvOp("ct", m.upsample, m.upsample.in_channels, m.upsample.out_channels))
            if m.with_bn:
                net.ops.append(BNOp(m.bn, relu=m.with_relu))
            elif m.with_relu:
                raise SvrsError("up_block(with_bn=False, with_relu=True) is not used by any model")
        elif isinstance(m, nn.Conv2d):
            assert m.kernel_size == (3, 3) and m.stride == (1, 1) and m.padding == (1, 1)
            net.ops.append(ConvOp("c3", m, m.in_channels, m.out_channels))
        elif isinstance(m, nn.Sigmoid):
            net.ops[-1].act = ACT_SIGMOID
        elif isinstance(m, nn.Hardtanh):
            assert m.min_val == -7 and m.max_val == 7
            net.ops[-1].act = ACT_HARDTANH7
        elif isinstance(m, (nn.Flatten, nn.Unflatten)):
            continue
        else:
            raise SvrsError(f"unsupported layer {cls} in {name}")
    return net


class Runtime:
    """Kernel-level forward/backward over Net plans; owns weight packs and scratch."""

    def __init__(self, module: nn.Module, nets: Sequence[Net], compute_dtype: torch.dtype = torch.float32,
                 late_prefixes: Tuple[str, ...] = ()):
        lib.load()
        self.module = module
        self.nets = list(nets)
        self.dtype = compute_dtype
        self.dt = _dt(compute_dtype)
        self.store = ParamStore(module, late_prefixes)
        self._scratch: Optional[torch.Tensor] = None
        self._pack_jobs: Optional[torch.Tensor] = None
        self._pack_key = None
        self._unpack_cache: Dict = {}
        self._unpacked_early: set = set()
        self.after_phase1 = None   # optional hook(list of Nets whose backward is complete), see CondEngine.backward
        self.after_heads = None    # same, called as soon as the prior heads + u_to_z are complete (inside their branch)
        self.scratch_prezeroed = False   # the fused step zeroes all BatchNorm scratch once per step
        self.packs_dirty = True
        self._replayed = 0         # kernels re-issued by CUDA-graph replays (not seen by the library's own counter)
        # fused-step gradient layout: the tcgen05 wgrad kernels accumulate their per-tap packed scratch DIRECTLY in the flat
        # gradient buffer (at the parameter's offset) and svrs_adam_multi reads it there, so no gpack buffer / unpack pass
        self.fused_grads = False
        self.fuse_epilogues = os.environ.get("SVRS_FUSE_EPI", "1") != "0"   # BN statistics / head outputs in conv epilogues
        # sync_bn (SURVEY 8.4 row e-ii): all-reduce the [sum, sum^2] / [sum dy, sum dy*xhat] scratch of every BatchNorm so
        # the statistics are those of the GLOBAL batch (exact single-process parity of the data-parallel step)
        self.sync_bn = False
        self.pg = None
        self.world = 1
        # Weight gradients are leaves of the backward pass: they run on a side stream, concurrently with the
        # dgrad -> BatchNorm chain of the main stream (most of them are small, latency-bound launches that leave SMs
        # idle).  SVRS_WGRAD_STREAM=0 keeps everything on one stream.
        self.wgrad_side = os.environ.get("SVRS_WGRAD_STREAM", "1") != "0"
        self._sides: Dict[tuple, torch.cuda.Stream] = {}    # producer stream handle -> its wgrad stream
        self._busy_sides: List[torch.cuda.Stream] = []
        self.wgrad_ways = int(os.environ.get("SVRS_WGRAD_WAYS", "1"))    # measured: more than one stream per producer does not help
        self._wg_rr = 0
        self._side_busy = False
        self._wg_keep: list = []   # operands of in-flight side-stream wgrads (kept alive until the join)
        # Independent sub-networks (encoder_y | encoder_x | y_to_z, decoder_y | prior heads | decoder_x, and their
        # backward passes) run on parallel branch streams: their small-map layers are latency-bound and leave most SMs
        # idle.  SVRS_BRANCH_STREAMS=0 serialises them.
        self.branch_streams = os.environ.get("SVRS_BRANCH_STREAMS", "1") != "0"
        self._branches: List[torch.cuda.Stream] = []
        self._branch_open: List[bool] = []
        self._zero_stream: Optional[torch.cuda.Stream] = None
        self._zero_pending = False

    # kernels of libsvrs_b200.so enqueued so far (bench.py's gpu_launches): the library counts every launch site itself
    # (svrs_launch_count); graph replays add the number of kernels captured in the graph.  The `+= n` bookkeeping at the
    # call sites is kept as documentation of what each call enqueues but no longer feeds the number.
    @property
    def launches(self) -> int:
        return int(lib.launch_count()) + self._replayed

    @launches.setter
    def launches(self, _v):
        pass

    def add_replayed(self, n: int):
        self._replayed += int(n)

    # -------------------------------------------------------------------------------------- setup
    @property
    def device(self) -> torch.device:
        return self.store.params[0].device

    def set_dtype(self, compute_dtype: torch.dtype):
        if compute_dtype != self.dtype:
            self.dtype = compute_dtype
            self.dt = _dt(compute_dtype)
            for net in self.nets:
                for op in net.ops:
                    if isinstance(op, ConvOp):
                        op.pack_f = op.pack_b = None
                        op.fuse_bn = op.fuse_head = op.fuse_f32 = None
            self.packs_dirty = True

    def ensure(self):
        rebuilt = self.store.ensure()
        dev = self.device
        if rebuilt or self._scratch is None or self._scratch.device != dev:
            n = 0
            for net in self.nets:
                for op in net.ops:
                    if isinstance(op, BNOp):
                        n += 4 * op.mod.num_features * BN_REPLICAS
            self._scratch = torch.zeros(max(n, 1), device=dev, dtype=torch.float64)
            off = 0
            for net in self.nets:
                for op in net.ops:
                    if isinstance(op, BNOp):
                        c = op.mod.num_features * BN_REPLICAS     # SVRS_BN_REPLICAS copies of double[2C] each
                        op.sums_f = self._scratch[off:off + 2 * c]
                        op.sums_b = self._scratch[off + 2 * c:off + 4 * c]
                        off += 4 * c
            self.packs_dirty = True
            for net in self.nets:
                for op in net.ops:
                    if isinstance(op, ConvOp):
                        op.pack_f = op.pack_b = None

    def zero_scratch(self):
        lib.fill_zero(_p(self._scratch), self._scratch.numel() * 8, _st())
        self.launches += 1

    def zero_grads(self, with_scratch: bool = False, deferred: bool = False):
        """deferred: the 82 MB gradient fill is only needed by the first weight-gradient kernel of the backward pass, so the
        fused step runs it on a side stream next to the forward pass; join_zero_grads() is the matching wait."""
        if with_scratch:
            self.zero_scratch()
        cur = torch.cuda.current_stream()
        if deferred:
            if self._zero_stream is None:
                self._zero_stream = torch.cuda.Stream(device=self.device)
            self._zero_stream.wait_stream(cur)
            st = self._zero_stream.cuda_stream
            self._zero_pending = True
        else:
            st = cur.cuda_stream
        lib.fill_zero(_p(self.store.grad), self.store.grad.numel() * 4, st)
        if not self.fused_grads:
            lib.fill_zero(_p(self.store.gpack), self.store.gpack.numel() * 4, st)
        self.launches += 2

    def join_zero_grads(self):
        if self._zero_pending:
            torch.cuda.current_stream().wait_stream(self._zero_stream)
            self._zero_pending = False

    def branch(self, i: int):
        """Context manager: run the enclosed launches on branch stream `i`, forked from the current stream.
        Memory rules (torch caching allocator): tensors allocated inside belong to the branch stream's pool and may be
        consumed by the parent after join(); tensors of the parent that the branch reads must stay referenced until
        join() - every caller keeps them in locals / the tape."""
        rt = self

        class _Branch:
            def __enter__(self_b):
                if not rt.branch_streams:
                    return self_b
                while len(rt._branches) <= i:
                    rt._branches.append(torch.cuda.Stream(device=rt.device))
                    rt._branch_open.append(False)
                s_ = rt._branches[i]
                s_.wait_stream(torch.cuda.current_stream())
                rt._branch_open[i] = True
                self_b.cm = torch.cuda.stream(s_)
                self_b.cm.__enter__()
                return self_b

            def __exit__(self_b, *exc):
                if rt.branch_streams:
                    self_b.cm.__exit__(*exc)
                return False

        return _Branch()

    def join(self, *idx: int):
        """The current stream waits for the given branch streams."""
        for i in idx:
            if i < len(self._branches) and self._branch_open[i]:
                torch.cuda.current_stream().wait_stream(self._branches[i])
                self._branch_open[i] = False

    def _wgrad_stream(self, *operands) -> int:
        """Stream handle for a weight-gradient launch whose operands were produced by work already enqueued on the
        current stream.  Every producer stream (main, branch 0, branch 1) has its OWN wgrad stream, so the weight
        gradients of nets that run in parallel do not serialise behind each other.  The operands are kept alive until
        join_wgrads(): the caching allocator must not hand their memory to a later allocation while a side stream still
        reads them."""
        if not self.wgrad_side:
            return _st()
        cur = torch.cuda.current_stream()
        self._wg_rr = (self._wg_rr + 1) % self.wgrad_ways
        key = (cur.cuda_stream, self._wg_rr)          # round-robin over `wgrad_ways` streams per producer
        side = self._sides.get(key)
        if side is None:
            side = self._sides[key] = torch.cuda.Stream(device=self.device)
        side.wait_stream(cur)
        self._side_busy = True
        if side not in self._busy_sides:
            self._busy_sides.append(side)
        self._wg_keep.append(operands)
        return side.cuda_stream

    def wgrad_streams(self):
        """wgrad streams with work of the current step (only these may be waited on: under CUDA-graph capture the
        streams used by earlier eager steps are not part of the capture)."""
        return list(self._busy_sides)

    def join_wgrads(self):
        if self._side_busy:
            cur = torch.cuda.current_stream()
            for side in self._busy_sides:
                cur.wait_stream(side)
            self._side_busy = False
        self._busy_sides = []
        self._wg_keep.clear()

    def _unpack_table(self, convs):
        """Device job table (cached) for svrs_unpack_grads_multi over the given conv ops."""
        key = (self.store.grad.data_ptr(), self.store.gpack.data_ptr(), tuple(id(op) for op in convs))
        hit = self._unpack_cache.get(key)
        if hit is None:
            import numpy as np
            rec = np.dtype([("w", "<u8"), ("p01", "<u8"), ("p10", "<u8"), ("d0", "<i4"), ("d1", "<i4"), ("kk", "<i4"),
                            ("tile0", "<i4"), ("tiles_b", "<i4"), ("pad", "<i4")])
            jobs = np.zeros(len(convs), dtype=rec)
            tile0 = 0
            for i, op in enumerate(convs):
                w = op.mod.weight
                d0, d1 = w.shape[0], w.shape[1]
                tiles_b = (d1 + 15) // 16
                jobs[i] = (self.store.grad_ptr(w), self.store.gpack_ptr(w), 0, d0, d1, op.kk, tile0, tiles_b, 0)
                tile0 += ((d0 + 31) // 32) * tiles_b
            hit = (torch.from_numpy(jobs.view(np.uint8).copy()).to(self.device), tile0, len(convs))
            self._unpack_cache[key] = hit
        return hit

    def unpack_nets(self, nets):
        """Add the packed weight-gradient scratch of the given nets' conv layers into the flat gradient NOW (current
        stream) and remember them, so that finish_grads() only handles the rest.  Used by the data-parallel trainer to
        start the all-reduce of the decoders' gradients while the encoders' backward pass is still running."""
        convs = [op for net in nets for op in net.ops if isinstance(op, ConvOp)]
        if not convs:
            return
        jobs, tiles, n = self._unpack_table(convs)
        lib.unpack_grads_multi(_p(jobs), n, tiles, 16, _st())
        self.launches += 1
        self._unpacked_early = {id(op) for op in convs}

    def finish_grads(self):
        """Add the per-tap packed weight-gradient scratch of every conv layer into the torch-layout flat gradient
        (one launch).  Must run after the last net_backward of a step and before anything reads store.grad."""
        self.join_wgrads()
        if self.fused_grads:
            return
        convs = [op for net in self.nets for op in net.ops if isinstance(op, ConvOp) and id(op) not in self._unpacked_early]
        self._unpacked_early = set()
        if not convs:
            return
        jobs, tiles, n = self._unpack_table(convs)
        lib.unpack_grads_multi(_p(jobs), n, tiles, 16, _st())
        self.launches += 1

    def pack_weights(self, force: bool = False):
        """fp32 master weights (torch layout) -> per-tap KN/NK packs in the compute dtype, all layers in one launch."""
        if not (self.packs_dirty or force):
            return
        dev = self.device
        convs = [op for net in self.nets for op in net.ops if isinstance(op, ConvOp)]
        rebuild = self._pack_jobs is None or self._pack_jobs.device != dev
        for op in convs:
            if op.pack_f is None or op.pack_f.device != dev or op.pack_f.dtype != self.dtype:
                n = op.mod.weight.numel()
                op.pack_f = torch.empty(n, device=dev, dtype=self.dtype)
                op.pack_b = torch.empty(n, device=dev, dtype=self.dtype)
                rebuild = True
        if rebuild or self._pack_key != self.store.flat.data_ptr():
            import numpy as np
            rec = np.dtype([("w", "<u8"), ("p01", "<u8"), ("p10", "<u8"), ("d0", "<i4"), ("d1", "<i4"), ("kk", "<i4"),
                            ("tile0", "<i4"), ("tiles_b", "<i4"), ("pad", "<i4")])
            assert rec.itemsize == lib.pack_job_bytes()
            jobs = np.zeros(len(convs), dtype=rec)
            tile0 = 0
            for i, op in enumerate(convs):
                w = op.mod.weight
                d0, d1 = w.shape[0], w.shape[1]
                # conv  weight [Cout][Cin][kk]: fprop KN pack [t][Cin][Cout] = p10, dgrad KN pack [t][Cout][Cin] = p01
                # convT weight [Cin][Cout][16]: fprop KN pack [t][Cin][Cout] = p01, dgrad KN pack [t][Cout][Cin] = p10
                p01, p10 = (op.pack_f, op.pack_b) if op.kind == "ct" else (op.pack_b, op.pack_f)
                tiles_b = (d1 + 15) // 16
                jobs[i] = (w.data.data_ptr(), p01.data_ptr(), p10.data_ptr(), d0, d1, op.kk, tile0, tiles_b, 0)
                tile0 += ((d0 + 31) // 32) * tiles_b
            self._pack_jobs = torch.from_numpy(jobs.view(np.uint8).copy()).to(dev)
            self._pack_tiles = tile0
            self._pack_n = len(convs)
            self._pack_key = self.store.flat.data_ptr()
        lib.pack_weights_multi(_p(self._pack_jobs), self._pack_n, self._pack_tiles, 16, self.dt, _st())
        self.launches += 1
        self.packs_dirty = False

    # -------------------------------------------------------------------------------------- layout glue
    def to_nhwc(self, src: torch.Tensor, src_ld: int, n: int, c: int, h: int, w: int, dtype=None) -> torch.Tensor:
        """NCHW-flat rows (fp32, row stride src_ld) -> NHWC tensor [n,h,w,c] in the compute dtype (or `dtype`)."""
        dtype = self.dtype if dtype is None else dtype
        out = torch.empty((n, h, w, c), device=src.device, dtype=dtype)
        lib.nchw_to_nhwc(_p(src), _dt(src.dtype), src_ld, _p(out), _dt(dtype), n, c, h, w, _st())
        self.launches += 1
        return out

    def to_nchw(self, src: torch.Tensor, dst: torch.Tensor, dst_ld: int, accumulate: bool = False):
        """NHWC tensor [n,h,w,c] -> NCHW-flat rows of dst (row stride dst_ld)."""
        n, h, w, c = src.shape
        lib.nhwc_to_nchw(_p(src), _dt(src.dtype), _p(dst), _dt(dst.dtype), dst_ld, n, c, h, w, int(accumulate), _st())
        self.launches += 1

    def copy2d(self, src, src_ld, dst, dst_ld, rows, cols, accumulate=False):
        lib.copy2d(_p(src) if isinstance(src, torch.Tensor) else src, _dt(self._dtype_of(src)), src_ld,
                   _p(dst) if isinstance(dst, torch.Tensor) else dst, _dt(self._dtype_of(dst)), dst_ld,
                   rows, cols, int(accumulate), _st())
        self.launches += 1

    @staticmethod
    def _dtype_of(t):
        return t.dtype

    def cast(self, t: torch.Tensor, dtype) -> torch.Tensor:
        if t.dtype == dtype:
            return t
        out = torch.empty(t.shape, device=t.device, dtype=dtype)
        lib.cast(_p(t), _dt(t.dtype), _p(out), _dt(dtype), t.numel(), _st())
        self.launches += 1
        return out

    def patch_batch(self, t) -> "PatchBatch":
        """Accept either a PatchBatch (dataset.grid_patch_pair) or a plain NCHW tensor [B,C,P,P] (the reference's batch
        layout) and return NHWC fp32 + NHWC compute-dtype operands."""
        if isinstance(t, PatchBatch):
            _require_cuda(t.f32, "patch batch")
            if t.op.dtype != self.dtype:
                return PatchBatch(t.f32, self.cast(t.f32, self.dtype))
            return t
        _require_cuda(t, "input batch")
        t = t.contiguous().float()
        b, c, h, w = t.shape
        f32 = self.to_nhwc(t, c * h * w, b, c, h, w, torch.float32)
        return PatchBatch(f32, self.cast(f32, self.dtype))

    def _bn_allreduce(self, sums: torch.Tensor):
        torch.distributed.all_reduce(sums, group=self.pg)

    # -------------------------------------------------------------------------------------- forward
    def net_forward(self, net: Net, x: torch.Tensor, training: bool, save: bool, bn_updates: int = 1,
                    head: Optional[Tuple[torch.Tensor, int]] = None, need_nhwc: bool = True, f32_out: bool = False):
        """x: NHWC [N,H,W,Cin] in the compute dtype.  Returns (out NHWC, tape).
        head = (dst, ld): the LAST conv also writes its result as fp32 NCHW-flat rows of dst (row stride ld) - from the
        tcgen05 epilogue when that kernel takes the layer, else through svrs_nhwc_to_nchw; with need_nhwc=False the NHWC
        result may be skipped (returned as None).  f32_out: the last conv stores fp32 (sigmoid tail -> NLL)."""
        st = _st()
        n, h, w, c = x.shape
        assert c == net.cin, f"{net.name}: expected {net.cin} channels, got {c}"
        tape = []
        last_conv = max(i for i, op in enumerate(net.ops) if isinstance(op, ConvOp))
        fuse = self.fuse_epilogues and self.dtype == torch.bfloat16
        pending_stats = None      # BNOp whose statistics the producing conv already accumulated
        for i, op in enumerate(net.ops):
            if isinstance(op, ConvOp):
                oh, ow = op.out_hw(h, w)
                bias = op.mod.bias
                is_last = i == last_conv
                nxt = net.ops[i + 1] if i + 1 < len(net.ops) else None
                want_bn = fuse and FUSE_BN and training and isinstance(nxt, BNOp) and op.fuse_bn is not False
                if want_bn and FUSE_BN_MAX_MB and n * oh * ow * op.cout * 2 > FUSE_BN_MAX_MB * 1e6:
                    want_bn = False
                want_head = is_last and head is not None and fuse and op.fuse_head is not False and op.kind != "ct"
                want_f32 = is_last and f32_out and self.dtype != torch.float32 and op.fuse_f32 is not False and op.kind != "ct"
                out_dtype = torch.float32 if want_f32 else self.dtype
                skip_nhwc = want_head and not need_nhwc
                y = None if skip_nhwc else torch.empty((n, oh, ow, op.cout), device=x.device, dtype=out_dtype)
                sums = None
                if want_bn:
                    sums = nxt.sums_f
                    if not self.scratch_prezeroed:
                        lib.fill_zero(_p(sums), 16 * nxt.mod.num_features * BN_REPLICAS, st)
                done = False
                if want_bn or want_head or want_f32:
                    try:
                        if op.kind == "ct":
                            lib.convT2d_fprop_ex(_p(x), _p(op.pack_f), _p(op.pack_b), _p(bias), _p(y), self.dt, _p(sums),
                                                 n, h, w, op.cin, op.cout, op.act, st)
                        else:
                            lib.conv2d_fprop_ex(_p(x), _p(op.pack_f), _p(op.pack_b), _p(bias), _p(y), self.dt, _dt(out_dtype),
                                                _p(head[0]) if want_head else None, head[1] if want_head else 0, _p(sums),
                                                n, h, w, op.cin, op.cout, 3 if op.kind == "c3" else 4, op.act, st)
                        done = True
                        if want_bn:
                            op.fuse_bn = True
                            pending_stats = nxt
                        if want_head:
                            op.fuse_head = True
                        if want_f32:
                            op.fuse_f32 = True
                    except SvrsUnsupported:
                        # remember which extra this layer's kernel lacks and fall through to the separate kernels
                        if want_bn:
                            op.fuse_bn = False
                        if want_head:
                            op.fuse_head = False
                        if want_f32:
                            op.fuse_f32 = False
                        want_head = False
                        if y is None or y.dtype != self.dtype:
                            y = torch.empty((n, oh, ow, op.cout), device=x.device, dtype=self.dtype)
                if not done:
                    if op.kind == "ct":
                        lib.convT2d_fprop(_p(x), _p(op.pack_f), _p(op.pack_b), _p(bias), _p(y), self.dt, n, h, w, op.cin, op.cout, op.act, st)
                    else:
                        lib.conv2d_fprop(_p(x), _p(op.pack_f), _p(op.pack_b), _p(bias), _p(y), self.dt, n, h, w, op.cin, op.cout,
                                         3 if op.kind == "c3" else 4, op.act, st)
                self.launches += 1
                if is_last and head is not None and not (done and want_head):
                    self.to_nchw(y, head[0], head[1])
                if is_last and f32_out and y is not None and y.dtype != torch.float32:
                    y = self.cast(y, torch.float32)
                if save:
                    tape.append((op, x, y if op.act != ACT_NONE else None))
                x, h, w = y, oh, ow
            else:
                bn = op.mod
                cch = bn.num_features
                m = n * h * w
                dev = x.device
                scale = torch.empty(cch, device=dev, dtype=torch.float32)
                shift = torch.empty(cch, device=dev, dtype=torch.float32)
                if training:
                    mean = torch.empty(cch, device=dev, dtype=torch.float32)
                    invstd = torch.empty(cch, device=dev, dtype=torch.float32)
                    if pending_stats is not op:
                        if not self.scratch_prezeroed:
                            lib.fill_zero(_p(op.sums_f), 16 * cch * BN_REPLICAS, st)
                        lib.bn_stats(_p(x), self.dt, m, cch, _p(op.sums_f), st)
                    pending_stats = None
                    m_stat = 0
                    if self.sync_bn and self.world > 1:
                        self._bn_allreduce(op.sums_f)
                        m_stat = m * self.world
                    mom = 0.1 if bn.momentum is None else bn.momentum
                    track = bn.track_running_stats and bn.running_mean is not None
                    y = torch.empty_like(x) if save else x
                    lib.bn_apply_train(_p(x), _p(y), self.dt, m, m_stat, cch, _p(op.sums_f), _p(bn.weight), _p(bn.bias), bn.eps, mom,
                                       _p(bn.running_mean) if track else None,
                                       _p(bn.running_var) if track else None,
                                       _p(bn.num_batches_tracked) if track else None, bn_updates, int(op.relu),
                                       _p(scale), _p(shift), _p(mean), _p(invstd), st)
                    self.launches += 3
                    if save:
                        tape.append((op, x, scale, shift, mean, invstd))
                    x = y
                    continue
                else:
                    mean = invstd = None
                    lib.bn_finalize_eval(cch, _p(bn.weight), _p(bn.bias), bn.eps, _p(bn.running_mean),
                                         _p(bn.running_var), _p(scale), _p(shift), st)
                    self.launches += 1
                y = torch.empty_like(x) if save else x
                lib.bn_apply(_p(x), _p(y), self.dt, m, cch, _p(scale), _p(shift), int(op.relu), st)
                self.launches += 1
                if save:
                    tape.append((op, x, scale, shift, mean, invstd))
                x = y
        return x, tape

    # -------------------------------------------------------------------------------------- backward
    def net_backward(self, net: Net, tape: list, dy: torch.Tensor, need_dx: bool, act_done: bool = False) -> Optional[torch.Tensor]:
        """dy: NHWC grad wrt the net output (compute dtype; MAY be modified in place).  Accumulates parameter
        gradients into the flat fp32 gradient buffer; returns dx (NHWC) if need_dx.  act_done: the backward of the
        output activation (Sigmoid / Hardtanh) has already been applied to dy by the caller."""
        st = _st()
        store = self.store
        for idx in range(len(tape) - 1, -1, -1):
            entry = tape[idx]
            op = entry[0]
            if isinstance(op, ConvOp):
                _, x, yact = entry
                n, h, w, _c = x.shape
                if op.act != ACT_NONE and not (act_done and idx == len(tape) - 1):
                    lib.act_bwd(_p(yact), _p(dy), _p(dy), self.dt, op.act, dy.numel(), st)
                    self.launches += 1
                dw = store.grad_ptr(op.mod.weight)
                dwp = dw if self.fused_grads else store.gpack_ptr(op.mod.weight)
                db = store.grad_ptr(op.mod.bias) if op.mod.bias is not None else None
                wst = self._wgrad_stream(x, dy)
                if op.kind == "ct":
                    lib.convT2d_wgrad(_p(x), _p(dy), dw, dwp, db, self.dt, n, h, w, op.cin, op.cout, 0, wst)
                else:
                    lib.conv2d_wgrad(_p(x), _p(dy), dw, dwp, db, self.dt, n, h, w, op.cin, op.cout,
                                     3 if op.kind == "c3" else 4, 0, wst)
                self.launches += 2
                if idx > 0 or need_dx:
                    dx = torch.empty_like(x)
                    if op.kind == "ct":
                        lib.convT2d_dgrad(_p(dy), _p(op.pack_b), _p(op.pack_f), _p(dx), self.dt, n, h, w, op.cin, op.cout, st)
                    else:
                        lib.conv2d_dgrad(_p(dy), _p(op.pack_b), _p(op.pack_f), _p(dx), self.dt, n, h, w, op.cin, op.cout,
                                         3 if op.kind == "c3" else 4, st)
                    self.launches += 1
                    dy = dx
                else:
                    dy = None
            else:
                _, x, scale, shift, mean, invstd = entry
                if mean is None:
                    raise SvrsError("backward through eval-mode BatchNorm is not part of the training path")
                n, h, w, cch = x.shape
                m = n * h * w
                bn = op.mod
                if not self.scratch_prezeroed:
                    lib.fill_zero(_p(op.sums_b), 16 * cch * BN_REPLICAS, st)
                lib.bn_bwd_reduce(_p(x), _p(dy), self.dt, m, cch, _p(scale), _p(shift), _p(mean), _p(invstd),
                                  int(op.relu), _p(op.sums_b), st)
                m_stat = 0
                if self.sync_bn and self.world > 1:
                    self._bn_allreduce(op.sums_b)
                    m_stat = m * self.world
                lib.bn_bwd_apply(_p(x), _p(dy), _p(dy), self.dt, m, m_stat, cch, _p(scale), _p(shift), _p(mean), _p(invstd),
                                 _p(bn.weight), int(op.relu), _p(op.sums_b),
                                 store.grad_ptr(bn.weight), store.grad_ptr(bn.bias), st)
                self.launches += 3
        return dy if need_dx else None


def _walk_conv_shapes(net: Net, n: int, h: int, w: int, out: dict):
    for op in net.ops:
        if isinstance(op, ConvOp):
            out[id(op)] = (n, h, w)
            h, w = op.out_hw(h, w)


# ------------------------------------------------------------------------------------------------
# reparameterisation config
# ------------------------------------------------------------------------------------------------
@dataclass
class RngState:
    """Philox4x32-10 addressing of the reparameterisation noise: key = seed, counter = (element index of the GLOBAL sample,
    stream id, step).  seed None = derived from torch.initial_seed() at first use (torch.manual_seed then selects the noise,
    as it does for the reference's torch.randn_like)."""
    seed: Optional[int] = None
    sample_offset: int = 0          # global index of this rank's first sample (partition-invariant eps)
    step_ptr: Optional[torch.Tensor] = None   # device int64 step counter (Philox counter word 3)
    sid_base: int = 0               # added to the stream ids (the non-fused path draws from its own streams)

    def key(self) -> int:
        if self.seed is None:
            self.seed = int(torch.initial_seed()) & 0x7FFFFFFFFFFFFFFF
        return self.seed


def reparam_fwd(rt: Runtime, enc, eps, z, z_ld, b, wd, rng: RngState, stream_id: int, eps_out=None):
    lib.reparam_fwd(_p(enc), _p(eps), z if isinstance(z, int) else _p(z), z_ld, _p(eps_out), b, wd,
                    rng.key(), stream_id + rng.sid_base, rng.sample_offset, _p(rng.step_ptr), _st())
    rt.launches += 1


def reparam_bwd(rt: Runtime, enc, eps, dz, dz_ld, denc, b, wd, rng: RngState, stream_id: int):
    lib.reparam_bwd(_p(enc), _p(eps), dz if isinstance(dz, int) else _p(dz), dz_ld, _p(denc), b, wd,
                    rng.key(), stream_id + rng.sid_base, rng.sample_offset, _p(rng.step_ptr), _st())
    rt.launches += 1


class _OwnNoise:
    """Noise of every forward that is NOT driven by the fused trainer (model(x, y) under autograd, validation, evaluate(),
    sample()): the engine keeps its own device counter, advanced once per call that draws eps on the device, so two
    consecutive calls never see the same noise (the reference draws fresh torch.randn_like every call, cond_vae.py:261-265)."""

    def __init__(self):
        self.step = None

    def tick(self, base: RngState, device) -> RngState:
        if self.step is None or self.step.device != device:
            self.step = torch.zeros(1, device=device, dtype=torch.int64)
        lib.step_increment(_p(self.step), _st())
        return RngState(seed=base.key(), sample_offset=base.sample_offset, step_ptr=self.step, sid_base=8)


# ------------------------------------------------------------------------------------------------
# Cond_SRVAE engine
# ------------------------------------------------------------------------------------------------
class CondEngine:
    """Forward/backward of Cond_SRVAE.forward (cond_vae.py:275-286) over the kernel library.

    y_to_z is evaluated ONCE and consumed twice (z_cond :239 and decode_x :271); its BatchNorm running
    statistics are advanced twice and num_batches_tracked by 2 to keep state_dict parity."""

    def __init__(self, model: nn.Module, compute_dtype=torch.float32):
        self.model = model
        self.P = model.patch_size
        self.L = model.latent_size
        self.Lu = model.latent_size_y
        names = ["encoder_y", "decoder_y", "encoder_x", "decoder_x", "y_to_z", "u_to_z", "mu_u_y_to_z", "logvar_u_y_to_z"]
        self.nets = {n: plan_sequential(n, getattr(model, n)) for n in names}
        # backward phase 2 (CondEngine.backward) = the three encoders: their gradients are complete last
        self.rt = Runtime(model, list(self.nets.values()), compute_dtype, late_prefixes=("encoder_y.", "encoder_x.", "y_to_z."))
        P = self.P
        assert P % 16 == 0, "patch_size must be a multiple of 16"
        self.cz = self.L // 64                   # channels of mu_z at (P/8)^2
        self.cu = self.Lu // 64
        self.Wz = self.cz * (P // 8) ** 2        # true latent widths (SURVEY 3.2 caveat)
        self.Wu = self.cu * (P // 8) ** 2
        self.c16 = self.L // 16                  # y_to_z / u_to_z / prior-head channels at (P/16)^2
        self.cu16 = self.Lu // 16
        if self.c16 * (P // 16) ** 2 != self.Wz or self.cu16 * (P // 16) ** 2 != self.Wu:
            raise SvrsError("inconsistent latent geometry (reference would fail in torch.cat/Unflatten)")
        if self.Wz % 4 or self.Wu % 4:
            raise SvrsError("latent widths must be multiples of 4")
        self.rng = RngState()
        self._own = _OwnNoise()

    def conv_input_shapes(self, B: int) -> dict:
        """id(ConvOp) -> (N, H, W) of the layer's input at batch B (the wgrad dispatch depends on the map size)."""
        P, N, out = self.P, self.nets, {}
        h8, h16 = P // 8, P // 16
        for name, hw in (("encoder_y", P // 2), ("y_to_z", P // 2), ("encoder_x", P), ("decoder_y", h8), ("decoder_x", h8),
                         ("u_to_z", h16), ("mu_u_y_to_z", h16), ("logvar_u_y_to_z", h16)):
            _walk_conv_shapes(N[name], B, hw, hw, out)
        return out

    # ---- forward ---------------------------------------------------------------------------------
    def forward(self, x, y, eps_u: Optional[torch.Tensor], eps_z: Optional[torch.Tensor],
                training: bool, save: bool, repack: bool = True, fused_io: bool = False):
        """x [B,4,P,P], y [B,4,P/2,P/2] (NCHW fp32 CUDA tensors, or PatchBatch objects).  Returns (outs, ctx) where
        outs = dict(x_hat_nhwc, y_hat_nhwc (fp32 NHWC), enc_z [B,2Wz], enc_u [B,2Wu], mu3, lv3 (fp32 NCHW-flat)) plus, unless
        fused_io, x_hat / y_hat in the reference's NCHW layout."""
        rt = self.rt
        rt.ensure()
        if repack:
            rt.packs_dirty = True
        rt.pack_weights()
        xb, yb = rt.patch_batch(x), rt.patch_batch(y)
        P, B = self.P, xb.B
        assert tuple(xb.f32.shape[1:]) == (P, P, 4) and tuple(yb.f32.shape[1:]) == (P // 2, P // 2, 4) and yb.B == B
        dev = xb.f32.device
        f32 = dict(device=dev, dtype=torch.float32)
        h8, h16 = P // 8, P // 16
        Wz, Wu = self.Wz, self.Wu
        N = self.nets
        ctx = {}

        y_nhwc, x_nhwc = yb.op, xb.op
        # noise: the fused trainer owns rng.step_ptr (one tick per optimisation step, eps recomputed in backward); every other
        # caller draws from the engine's own counter and, when a backward may follow, keeps the eps it drew
        rng = self.rng
        eps_u_out = eps_z_out = None
        if not fused_io and (eps_u is None or eps_z is None):
            rng = self._own.tick(self.rng, dev)
            if save:
                eps_u_out = torch.empty((B, Wu), **f32) if eps_u is None else None
                eps_z_out = torch.empty((B, Wz), **f32) if eps_z is None else None

        enc_u = torch.empty((B, 2 * Wu), **f32)
        u = torch.empty((B, Wu), **f32)
        enc_z = torch.empty((B, 2 * Wz), **f32)
        stack = torch.empty((B, 2 * Wz), **f32)          # torch.cat((y_enc, z), dim=1)  cond_vae.py:272
        mu3 = torch.empty((B, Wz), **f32)
        lv3 = torch.empty((B, Wz), **f32)

        # ---- phase 1: three independent encoders -------------------------------------------------------------------
        # The posterior heads (mu || logvar) leave the conv stacks as fp32 NCHW-flat rows written by the conv epilogue
        # (net_forward head=...): chunk / Flatten are addressing, mu / logvar are never rounded to bf16.
        with rt.branch(0):
            # q(u|y): encoder_y -> chunk -> reparameterize (RNG draw #1; Philox is counter-based, so the
            # draw order is a naming convention - stream ids 0 / 1 - not an execution order)
            _, ctx["t_ey"] = rt.net_forward(N["encoder_y"], y_nhwc, training, save, head=(enc_u, 2 * Wu), need_nhwc=False)
            reparam_fwd(rt, enc_u, eps_u, u, Wu, B, Wu, rng, 0, eps_out=eps_u_out)
        with rt.branch(1):
            # y_to_z once (two BN running-stat updates); left half of stack = y_enc flat, NHWC copy feeds the prior heads
            yz, ctx["t_yz"] = rt.net_forward(N["y_to_z"], y_nhwc, training, save, bn_updates=2, head=(stack, 2 * Wz))
        # q(z|x): encoder_x -> chunk -> reparameterize (draw #2); z lands in the right half of `stack`
        _, ctx["t_ex"] = rt.net_forward(N["encoder_x"], x_nhwc, training, save, head=(enc_z, 2 * Wz), need_nhwc=False)
        reparam_fwd(rt, enc_z, eps_z, stack.data_ptr() + 4 * Wz, 2 * Wz, B, Wz, rng, 1, eps_out=eps_z_out)
        rt.join(0, 1)

        # ---- phase 2: decoder_y | prior heads | decoder_x -----------------------------------------------------------
        with rt.branch(0):
            # decode_y(u): u viewed (Lu/64, P/8, P/8)
            u8 = rt.to_nhwc(u, Wu, B, self.cu, h8, h8)
            yh, ctx["t_dy"] = rt.net_forward(N["decoder_y"], u8, training, save, f32_out=True)
        with rt.branch(1):
            # u_to_z on u re-viewed as (Lu/16, P/16, P/16)   cond_vae.py:168-175
            u16 = rt.to_nhwc(u, Wu, B, self.cu16, h16, h16)
            uz, ctx["t_uz"] = rt.net_forward(N["u_to_z"], u16, training, save)
            # jointure = cat(y_enc, u_enc) viewed (2L/16, P/16, P/16) == channel concat in NHWC
            c16 = self.c16
            joint = torch.empty((B, h16, h16, 2 * c16), device=dev, dtype=rt.dtype)
            rows = B * h16 * h16
            es = joint.element_size()
            rt.copy2d(yz, c16, joint, 2 * c16, rows, c16)
            lib.copy2d(_p(uz), rt.dt, c16, joint.data_ptr() + es * c16, rt.dt, 2 * c16, rows, c16, 0, _st())
            rt.launches += 1
            with rt.branch(2):                           # the two prior heads only share their input
                _, ctx["t_lv"] = rt.net_forward(N["logvar_u_y_to_z"], joint, training, save, head=(lv3, Wz), need_nhwc=False)
            _, ctx["t_mu"] = rt.net_forward(N["mu_u_y_to_z"], joint, training, save, head=(mu3, Wz), need_nhwc=False)
            rt.join(2)
        # decode_x(z, y): stack viewed (2L/64, P/8, P/8)
        s8 = rt.to_nhwc(stack, 2 * Wz, B, 2 * self.cz, h8, h8)
        xh, ctx["t_dx"] = rt.net_forward(N["decoder_x"], s8, training, save, f32_out=True)
        rt.join(0, 1)

        outs = dict(x_hat_nhwc=xh, y_hat_nhwc=yh, enc_z=enc_z, enc_u=enc_u, mu3=mu3, lv3=lv3, xb=xb, yb=yb)
        if not fused_io:
            x_hat = torch.empty((B, 4, P, P), **f32)
            y_hat = torch.empty((B, 4, P // 2, P // 2), **f32)
            rt.to_nchw(xh, x_hat, 4 * P * P)
            rt.to_nchw(yh, y_hat, 4 * (P // 2) ** 2)
            outs.update(x_hat=x_hat, y_hat=y_hat)
        if save:
            ctx.update(B=B, enc_u=enc_u, enc_z=enc_z, eps_u=eps_u if eps_u_out is None else eps_u_out,
                       eps_z=eps_z if eps_z_out is None else eps_z_out, rng=RngState(**vars(rng)),
                       lv3=lv3, x_hat_nhwc=xh, y_hat_nhwc=yh)
        return outs, ctx

    # ---- backward --------------------------------------------------------------------------------
    def _image_grad(self, d_nchw: torch.Tensor, y_nhwc_f32: torch.Tensor, B: int, hw: int) -> torch.Tensor:
        """Upstream gradient wrt a sigmoid output in the reference's NCHW fp32 layout -> gradient wrt the pre-activation,
        NHWC in the compute dtype (autograd path; the fused step gets this straight from svrs_elbo_bwd)."""
        rt = self.rt
        g = rt.to_nhwc(d_nchw.contiguous().float(), 4 * hw * hw, B, 4, hw, hw, torch.float32)
        lib.act_bwd(_p(y_nhwc_f32), _p(g), _p(g), F32, ACT_SIGMOID, g.numel(), _st())
        rt.launches += 1
        return rt.cast(g, rt.dtype)

    def backward(self, ctx, d_xhat, d_yhat, d_enc_z, d_enc_u, d_mu3, d_lv3, fused_io: bool = False):
        """Gradients wrt the forward outputs (fp32, NCHW(-flat); None = zero; d_enc_* and d_lv3 are MODIFIED in place).
        fused_io: d_xhat / d_yhat are NHWC compute-dtype gradients wrt the decoders' PRE-sigmoid outputs (svrs_elbo_bwd with
        act = SIGMOID).  Parameter gradients are accumulated into rt.store.grad (caller zeroes it)."""
        rt = self.rt
        N = self.nets
        P, B = self.P, ctx["B"]
        h8, h16 = P // 8, P // 16
        Wz, Wu, c16 = self.Wz, self.Wu, self.c16
        dev = ctx["enc_u"].device
        f32 = dict(device=dev, dtype=torch.float32)
        rows = B * h16 * h16
        rng = ctx["rng"]

        def zeros(*shape):
            t = torch.empty(shape, **f32)
            lib.fill_zero(_p(t), t.numel() * 4, _st())
            rt.launches += 1
            return t

        # ---- phase 1: decoder_y | prior heads + u_to_z | decoder_x (independent until the latents) ------------------
        d_stack = d_u = d_yz = du16 = None
        with rt.branch(0):
            if d_yhat is not None:
                g_y = d_yhat if fused_io else self._image_grad(d_yhat, ctx["y_hat_nhwc"], B, P // 2)
                du8 = rt.net_backward(N["decoder_y"], ctx["t_dy"], g_y, True, act_done=True)
                d_u = torch.empty((B, Wu), **f32)
                rt.to_nchw(du8, d_u, Wu)
        with rt.branch(1):
            d_joint = dj_lv = None
            with rt.branch(2):                           # the two prior heads run in parallel
                if d_lv3 is not None:
                    # Hardtanh(-7, 7) backward (cond_vae.py:230) on the NCHW-flat fp32 rows, from the saved OUTPUT
                    d_lv3 = d_lv3.contiguous()
                    lib.act_bwd(_p(ctx["lv3"]), _p(d_lv3), _p(d_lv3), F32, ACT_HARDTANH7, d_lv3.numel(), _st())
                    rt.launches += 1
                    g_l = rt.to_nhwc(d_lv3, Wz, B, c16, h16, h16)
                    dj_lv = rt.net_backward(N["logvar_u_y_to_z"], ctx["t_lv"], g_l, True, act_done=True)
            if d_mu3 is not None:
                g_h = rt.to_nhwc(d_mu3.contiguous(), Wz, B, c16, h16, h16)
                d_joint = rt.net_backward(N["mu_u_y_to_z"], ctx["t_mu"], g_h, True)
            rt.join(2)
            if dj_lv is not None:
                if d_joint is None:
                    d_joint = dj_lv
                else:
                    rt.copy2d(dj_lv, 2 * c16, d_joint, 2 * c16, rows, 2 * c16, accumulate=True)
            # gradient wrt y_to_z output (NHWC) and u_to_z output
            if d_joint is not None:
                es = d_joint.element_size()
                d_yz = torch.empty((B, h16, h16, c16), device=dev, dtype=rt.dtype)
                d_uz = torch.empty((B, h16, h16, c16), device=dev, dtype=rt.dtype)
                rt.copy2d(d_joint, 2 * c16, d_yz, c16, rows, c16)
                lib.copy2d(d_joint.data_ptr() + es * c16, rt.dt, 2 * c16, _p(d_uz), rt.dt, c16, rows, c16, 0, _st())
                rt.launches += 1
                du16 = rt.net_backward(N["u_to_z"], ctx["t_uz"], d_uz, True)
            if rt.after_heads is not None:
                # prior heads + u_to_z are done long before the decoder_x chain: their 15 M parameters (60 MB of gradients) can
                # travel while the rest of the backward pass runs
                rt.after_heads([N[k] for k in ("u_to_z", "mu_u_y_to_z", "logvar_u_y_to_z")])
        if d_xhat is not None:
            g_x = d_xhat if fused_io else self._image_grad(d_xhat, ctx["x_hat_nhwc"], B, P)
            ds8 = rt.net_backward(N["decoder_x"], ctx["t_dx"], g_x, True, act_done=True)
            d_stack = torch.empty((B, 2 * Wz), **f32)
            rt.to_nchw(ds8, d_stack, 2 * Wz)
        rt.join(0, 1)
        if rt.after_phase1 is not None:
            # decoders (and, unless after_heads took them, prior heads and u_to_z) are done; their wgrads are queued on the wgrad streams
            rest = ("decoder_x", "decoder_y") if rt.after_heads is not None else ("decoder_x", "decoder_y", "mu_u_y_to_z", "logvar_u_y_to_z", "u_to_z")
            rt.after_phase1([N[k] for k in rest])
        if du16 is not None:
            if d_u is None:
                d_u = zeros(B, Wu)
            rt.to_nchw(du16, d_u, Wu, accumulate=True)
        if d_stack is not None:
            g_s = rt.to_nhwc(d_stack, 2 * Wz, B, c16, h16, h16)   # left half rows: y_enc as (L/16, P/16, P/16)
            if d_yz is None:
                d_yz = g_s
            else:
                rt.copy2d(g_s, c16, d_yz, c16, rows, c16, accumulate=True)
        # ---- phase 2: y_to_z | encoder_y | encoder_x ------------------------------------------------------------------
        with rt.branch(0):
            if d_yz is not None:
                rt.net_backward(N["y_to_z"], ctx["t_yz"], d_yz, False)
        with rt.branch(1):
            # encoder_y through reparameterize(u)
            if d_enc_u is None and d_u is not None:
                d_enc_u = zeros(B, 2 * Wu)
            if d_enc_u is not None:
                if d_u is not None:
                    reparam_bwd(rt, ctx["enc_u"], ctx["eps_u"], d_u, Wu, d_enc_u, B, Wu, rng, 0)
                g_eu = rt.to_nhwc(d_enc_u, 2 * Wu, B, 2 * self.cu, h8, h8)
                rt.net_backward(N["encoder_y"], ctx["t_ey"], g_eu, False)
        # encoder_x through reparameterize(z)
        if d_enc_z is None and d_stack is not None:
            d_enc_z = zeros(B, 2 * Wz)
        if d_enc_z is not None:
            if d_stack is not None:
                reparam_bwd(rt, ctx["enc_z"], ctx["eps_z"], d_stack.data_ptr() + 4 * Wz, 2 * Wz, d_enc_z, B, Wz, rng, 1)
            g_ez = rt.to_nhwc(d_enc_z, 2 * Wz, B, 2 * self.cz, h8, h8)
            rt.net_backward(N["encoder_x"], ctx["t_ex"], g_ez, False)
        rt.join(0, 1)
        rt.finish_grads()

    # ---- inference: Cond_SRVAE.sample (cond_vae.py:299-318) ---------------------------------------
    def _sample_decoder_input(self, y, samples: int, eps_u, eps_s, training: bool):
        """Everything of sample() up to the decoder_x input, for B LR patches at once: encode_y -> u -> z_cond -> the S
        draws z(b, s) (svrs_sample_latents: one launch, Philox on device unless eps_s [B*S, Wz] is injected) -> NHWC rows
        [y_enc(b) | z(b, s)], row index b*S + s.  y_to_z(y) is identical for every draw of a patch, so it is computed once
        per patch and broadcast (the reference recomputes it on S expanded copies)."""
        rt = self.rt
        rt.ensure()
        rt.packs_dirty = True
        rt.pack_weights()
        P = self.P
        yb = rt.patch_batch(y)
        B = yb.B
        assert tuple(yb.f32.shape[1:]) == (P // 2, P // 2, 4)
        dev = yb.f32.device
        f32 = dict(device=dev, dtype=torch.float32)
        h8, h16 = P // 8, P // 16
        Wz, Wu, c16, S = self.Wz, self.Wu, self.c16, samples
        N = self.nets
        y_nhwc = yb.op
        rng = self._own.tick(self.rng, dev) if (eps_u is None or eps_s is None) else self.rng
        enc_u = torch.empty((B, 2 * Wu), **f32)
        rt.net_forward(N["encoder_y"], y_nhwc, training, False, head=(enc_u, 2 * Wu), need_nhwc=False)
        u = torch.empty((B, Wu), **f32)
        reparam_fwd(rt, enc_u, eps_u, u, Wu, B, Wu, rng, 0)
        yflat = torch.empty((B, Wz), **f32)
        yz, _ = rt.net_forward(N["y_to_z"], y_nhwc, training, False, bn_updates=2, head=(yflat, Wz))
        u16 = rt.to_nhwc(u, Wu, B, self.cu16, h16, h16)
        uz, _ = rt.net_forward(N["u_to_z"], u16, training, False)
        joint = torch.empty((B, h16, h16, 2 * c16), device=dev, dtype=rt.dtype)
        rows = B * h16 * h16
        es = joint.element_size()
        rt.copy2d(yz, c16, joint, 2 * c16, rows, c16)
        lib.copy2d(_p(uz), rt.dt, c16, joint.data_ptr() + es * c16, rt.dt, 2 * c16, rows, c16, 0, _st())
        rt.launches += 1
        mu3 = torch.empty((B, Wz), **f32)
        lv3 = torch.empty((B, Wz), **f32)
        rt.net_forward(N["mu_u_y_to_z"], joint, training, False, head=(mu3, Wz), need_nhwc=False)
        rt.net_forward(N["logvar_u_y_to_z"], joint, training, False, head=(lv3, Wz), need_nhwc=False)
        stack = torch.empty((B * S, 2 * Wz), **f32)
        lib.sample_latents(_p(mu3), _p(lv3), Wz, _p(yflat), Wz, _p(eps_s), _p(stack), B, S, Wz, rng.key(), 2 + rng.sid_base,
                           rng.sample_offset, _p(rng.step_ptr), _st())
        rt.launches += 1
        return rt.to_nhwc(stack, 2 * Wz, B * S, 2 * self.cz, h8, h8), B

    def sample(self, y: torch.Tensor, samples: int, eps_u=None, eps_s=None, training: bool = False):
        """S posterior-predictive decodes of ONE LR patch y [1,4,P/2,P/2] -> [S,4,P,P] (cond_vae.py:299-318)."""
        rt = self.rt
        _require_cuda(y, "y")
        P = self.P
        if y.ndim == 3:
            y = y.unsqueeze(0)
        assert y.shape == (1, 4, P // 2, P // 2), "sample() takes a single LR patch"
        s8, _ = self._sample_decoder_input(y.contiguous().float(), samples, eps_u, eps_s, training)
        xh, _ = rt.net_forward(self.nets["decoder_x"], s8, training, False, f32_out=True)
        out = torch.empty((samples, 4, P, P), device=y.device, dtype=torch.float32)
        rt.to_nchw(xh, out, 4 * P * P)
        return out

    def sample_stats(self, y, samples: int, target=None, eps_u=None, eps_s=None, training: bool = False,
                     splits: int = 0) -> Dict[str, torch.Tensor]:
        """The uncertainty maps of BaseVAE.task (models/base.py:305-313, 341) for B LR patches at once, WITHOUT materialising
        the [S,4,P,P] draws: decoder_x runs up to its last layer on all B*S rows and svrs_sample_tail_stats applies the
        final 16->4 conv + Sigmoid per draw while accumulating streaming Welford statistics over s (also
        BASELINE config 5's kernel).  y: [B,4,P/2,P/2] NCHW tensor or a PatchBatch; target: [B,4,P,P] NCHW, a PatchBatch, or
        None.  Returns fp32 tensors: mean [B,4,P,P], std / mae / mse / mean_bias [B,P,P], sample0 [B,4,P,P] (first draw).
        Train-mode BatchNorm (--test never calls eval()) is batch dependent, so B > 1 then runs patch by patch."""
        rt = self.rt
        P = self.P
        if isinstance(y, torch.Tensor):
            _require_cuda(y, "y")
            if y.ndim == 3:
                y = y.unsqueeze(0)
        B_in = y.B if isinstance(y, PatchBatch) else y.shape[0]
        if training and B_in > 1:
            parts = []
            for b in range(B_in):
                yb = PatchBatch(y.f32[b:b + 1], y.op[b:b + 1]) if isinstance(y, PatchBatch) else y[b:b + 1]
                tb = None
                if target is not None:
                    tb = PatchBatch(target.f32[b:b + 1], target.op[b:b + 1]) if isinstance(target, PatchBatch) else target[b:b + 1]
                parts.append(self.sample_stats(yb, samples, tb, None if eps_u is None else eps_u[b:b + 1],
                                               None if eps_s is None else eps_s[b * samples:(b + 1) * samples], True, splits))
            return {k: torch.cat([p_[k] for p_ in parts]) for k in parts[0]}
        s8, B = self._sample_decoder_input(y, samples, eps_u, eps_s, training)
        dec = self.nets["decoder_x"]
        last = dec.ops[-1]
        body = Net(dec.name, dec.ops[:-1])
        x16, _ = rt.net_forward(body, s8, training, False)
        dev = x16.device
        f32 = dict(device=dev, dtype=torch.float32)
        tgt = None
        if target is not None:
            tgt = rt.patch_batch(target).f32
            assert tuple(tgt.shape) == (B, P, P, 4)
        S = samples
        if splits <= 0:
            splits = int(lib.sample_tail_splits(B, S, P, P))
        scratch = torch.empty(int(lib.sample_tail_scratch_floats(B, P, P, splits)), **f32)
        out = dict(mean=torch.empty((B, 4, P, P), **f32), std=torch.empty((B, P, P), **f32),
                   sample0=torch.empty((B, 4, P, P), **f32))
        if tgt is not None:
            out.update(mae=torch.empty((B, P, P), **f32), mse=torch.empty((B, P, P), **f32),
                       mean_bias=torch.empty((B, P, P), **f32))
        lib.sample_tail_stats(_p(x16), rt.dt, _p(last.pack_f), _p(last.mod.bias), _p(tgt), B, S, P, P, last.cin, last.cout,
                              splits, _p(scratch), _p(out["mean"]), _p(out["std"]), _p(out.get("mae")), _p(out.get("mse")),
                              _p(out.get("mean_bias")), _p(out["sample0"]), _st())
        rt.launches += 2
        return out

    def sample_stats_batch(self, yb, samples: int) -> torch.Tensor:
        """bench.py's config-5 step: the statistics of `samples` draws for every LR patch of a tile batch; returns the
        [B,P,P] std map (the other maps are produced as well)."""
        return self.sample_stats(yb, samples)["std"]


# ------------------------------------------------------------------------------------------------
# VAE engine (vae.py:36-107)
# ------------------------------------------------------------------------------------------------
class VaeEngine:
    def __init__(self, model: nn.Module, compute_dtype=torch.float32):
        self.model = model
        self.P = model.patch_size
        self.L = model.latent_size
        self.nets = {n: plan_sequential(n, getattr(model, n)) for n in ("encoder", "decoder")}
        self.rt = Runtime(model, list(self.nets.values()), compute_dtype)
        assert self.P % 4 == 0
        self.c = self.L // 64
        self.Wd = self.c * (self.P // 4) ** 2
        if self.Wd % 4:
            raise SvrsError("latent width must be a multiple of 4")
        self.rng = RngState()
        self._own = _OwnNoise()

    def conv_input_shapes(self, B: int) -> dict:
        out = {}
        _walk_conv_shapes(self.nets["encoder"], B, self.P, self.P, out)
        _walk_conv_shapes(self.nets["decoder"], B, self.P // 4, self.P // 4, out)
        return out

    def forward(self, x, eps, training: bool, save: bool, repack: bool = True, fused_io: bool = False):
        rt = self.rt
        rt.ensure()
        if repack:
            rt.packs_dirty = True
        rt.pack_weights()
        xb = rt.patch_batch(x)
        P, B, Wd = self.P, xb.B, self.Wd
        assert tuple(xb.f32.shape[1:]) == (P, P, 4)
        f32 = dict(device=xb.f32.device, dtype=torch.float32)
        ctx = {}
        enc = torch.empty((B, 2 * Wd), **f32)
        _, ctx["t_e"] = rt.net_forward(self.nets["encoder"], xb.op, training, save, head=(enc, 2 * Wd), need_nhwc=False)
        z = torch.empty((B, Wd), **f32)
        rng, eps_out = self.rng, None
        if not fused_io and eps is None:
            rng = self._own.tick(self.rng, enc.device)
            eps_out = torch.empty((B, Wd), **f32) if save else None
        reparam_fwd(rt, enc, eps, z, Wd, B, Wd, rng, 0, eps_out=eps_out)
        z4 = rt.to_nhwc(z, Wd, B, self.c, P // 4, P // 4)
        d, ctx["t_d"] = rt.net_forward(self.nets["decoder"], z4, training, save, f32_out=True)
        outs = dict(x_hat_nhwc=d, enc=enc, xb=xb)
        if not fused_io:
            x_hat = torch.empty((B, 4, P, P), **f32)
            rt.to_nchw(d, x_hat, 4 * P * P)
            outs["x_hat"] = x_hat
        if save:
            ctx.update(B=B, enc=enc, eps=eps if eps_out is None else eps_out, rng=RngState(**vars(rng)), x_hat_nhwc=d)
        return outs, ctx

    def backward(self, ctx, d_xhat, d_enc, fused_io: bool = False):
        rt = self.rt
        P, B, Wd = self.P, ctx["B"], self.Wd
        f32 = dict(device=ctx["enc"].device, dtype=torch.float32)
        d_z = None
        if d_xhat is not None:
            if fused_io:
                g = d_xhat
            else:
                g = rt.to_nhwc(d_xhat.contiguous().float(), 4 * P * P, B, 4, P, P, torch.float32)
                lib.act_bwd(_p(ctx["x_hat_nhwc"]), _p(g), _p(g), F32, ACT_SIGMOID, g.numel(), _st())
                rt.launches += 1
                g = rt.cast(g, rt.dtype)
            dz4 = rt.net_backward(self.nets["decoder"], ctx["t_d"], g, True, act_done=True)
            d_z = torch.empty((B, Wd), **f32)
            rt.to_nchw(dz4, d_z, Wd)
        if d_enc is None and d_z is not None:
            d_enc = torch.empty((B, 2 * Wd), **f32)
            lib.fill_zero(_p(d_enc), d_enc.numel() * 4, _st())
            rt.launches += 1
        if d_enc is not None:
            if d_z is not None:
                reparam_bwd(rt, ctx["enc"], ctx["eps"], d_z, Wd, d_enc, B, Wd, ctx["rng"], 0)
            g = rt.to_nhwc(d_enc, 2 * Wd, B, 2 * self.c, P // 4, P // 4)
            rt.net_backward(self.nets["encoder"], ctx["t_e"], g, False)
        rt.finish_grads()
