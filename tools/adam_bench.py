"""Stand-alone timing of the optimiser tail (svrs_adam_multi) on the real job table of the bench model (Cond_SRVAE cr=2 P=64,
128 patches), against the plain streaming svrs_clip_adam over the same flat buffers (no transposes, no packs: 28 B per
parameter) and svrs_sumsq.  SVRS_ADAM_BULK=0 routes every tile to the generic (register-file) path.
    python tools/adam_bench.py"""
import os
import sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "simple-vae-rs_b200"), os.path.join(ROOT, "tests")]
import torch
import models
from svrs_native.lib import lib
from svrs_native.trainer import FusedCondTrainer

dev = "cuda"
torch.manual_seed(0)
model = models.Cond_SRVAE(2, 64).to(dev)
model.set_compute_dtype(torch.bfloat16)
model.train()
tr = FusedCondTrainer(model)
B = 128
x = torch.rand(B, 4, 64, 64, device=dev)
y = torch.rand(B, 4, 32, 32, device=dev)
for _ in range(2):
    tr.step(x, y)
torch.cuda.synchronize()
rt, store, cfg = tr.rt, tr.rt.store, tr.cfg
jobs, njobs, tiles = tr._adam_table(B)
n = store.flat.numel()
st = lambda: torch.cuda.current_stream().cuda_stream
_p = lambda t: t.data_ptr()
print(f"{n} parameters, {njobs} jobs, {tiles} tiles")


def timeit(fn, reps=10):
    for _ in range(3):
        fn()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps * 1e3


def multi():
    lib.adam_multi(_p(jobs), njobs, tiles, 16, _p(store.flat), _p(store.grad), _p(tr.m), _p(tr.v), rt.dt,
                   _p(tr.normacc), cfg.max_norm, 1.0, cfg.lr, cfg.b1, cfg.b2, cfg.eps, _p(tr.step_ptr), st())


def plain():
    lib.clip_adam(_p(store.flat), _p(store.grad), _p(tr.m), _p(tr.v), n, _p(tr.normacc),
                  cfg.max_norm, 1.0, cfg.lr, cfg.b1, cfg.b2, cfg.eps, _p(tr.step_ptr), st())


keep = (store.flat.clone(), tr.m.clone(), tr.v.clone())
us = timeit(plain)
print(f"clip_adam (plain stream, 28 B/param)  {us:7.1f} us  {28 * n / us / 1e6:5.2f} TB/s")
us = timeit(lambda: lib.sumsq(_p(store.grad), n, _p(tr.normacc), st()))
print(f"sumsq                                 {us:7.1f} us  {4 * n / us / 1e6:5.2f} TB/s")
us = timeit(multi)
print(f"adam_multi (SVRS_ADAM_BULK={os.environ.get('SVRS_ADAM_BULK', '1')}, SVRS_ADAM_OCC={os.environ.get('SVRS_ADAM_OCC', '5')})   {us:7.1f} us  {32 * n / us / 1e6:5.2f} TB/s (32 B/param)")
