"""Stand-alone timing of the narrow (4 / 16 channel) layers at the bench shapes (128 patches): fprop, dgrad, wgrad through the
C ABI, CUDA events over back-to-back launches.  SVRS_TC=0 routes them to the CUDA-core kernels (A/B).
    python tools/narrow_bench.py"""
import os
import sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "simple-vae-rs_b200"), os.path.join(ROOT, "tests")]
import torch
from svrs_native.lib import BF16, F32, lib

lib.load()
dev = "cuda"
st = lambda: torch.cuda.current_stream().cuda_stream
N = 128
CASES = [  # name, H, Cin, Cout, ksize, form
    ("enc_x.0.conv   c3 4->4   @64", 64, 4, 4, 3), ("enc_x.0.down   c4 4->16  @64", 64, 4, 16, 4),
    ("enc_y.0.conv   c3 4->4   @32", 32, 4, 4, 3), ("enc_y.0.down   c4 4->16  @32", 32, 4, 16, 4),
    ("dec_x.7        c3 16->4  @64", 64, 16, 4, 3), ("dec_y.6        c3 16->4  @32", 32, 16, 4, 3),
    ("dec_x.6        c3 16->16 @64", 64, 16, 16, 3), ("dec_x.5        c3 64->16 @64", 64, 64, 16, 3),
    ("dec_y.5        c3 16->16 @32", 32, 16, 16, 3), ("enc_y.1.conv   c3 16->16 @16", 16, 16, 16, 3),
]
if os.environ.get("ONLY16"):
    CASES = CASES[-4:]


def timeit(fn, reps=20):
    """Device time per launch: `reps` launches captured in a CUDA graph (no host launch overhead), replayed 5 times."""
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        for _ in range(reps):
            fn()
    g.replay()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    e0.record()
    for _ in range(5):
        g.replay()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / (5 * reps) * 1e3


for tc in ([1, 0] if os.environ.get("AB", "1") == "1" else [1]):
    lib.set_tc_enabled(tc)
    print(f"==== tc_enabled={tc}")
    for name, H, Ci, Co, ks in CASES:
        s = 1 if ks == 3 else 2
        OH = H // s
        x = torch.randn(N, H, H, Ci, device=dev).to(torch.bfloat16)
        w = torch.randn(Co, Ci, ks, ks, device=dev) * 0.1
        b = torch.randn(Co, device=dev)
        p01 = torch.empty(w.numel(), device=dev, dtype=torch.bfloat16)
        p10 = torch.empty(w.numel(), device=dev, dtype=torch.bfloat16)
        lib.pack_weights(w.data_ptr(), Co, Ci, ks * ks, p01.data_ptr(), p10.data_ptr(), BF16, st())
        pf, pb = p10, p01
        y = torch.empty(N, OH, OH, Co, device=dev, dtype=torch.bfloat16)
        dy = torch.randn(N, OH, OH, Co, device=dev).to(torch.bfloat16)
        dx = torch.empty_like(x)
        dw = torch.zeros_like(w)
        db = torch.zeros(Co, device=dev)
        mb_f = (x.numel() + y.numel()) * 2 / 1e6
        t_f = timeit(lambda: lib.conv2d_fprop(x.data_ptr(), pf.data_ptr(), pb.data_ptr(), b.data_ptr(), y.data_ptr(), BF16, N, H, H, Ci, Co, ks, 0, st()))
        t_d = timeit(lambda: lib.conv2d_dgrad(dy.data_ptr(), pb.data_ptr(), pf.data_ptr(), dx.data_ptr(), BF16, N, H, H, Ci, Co, ks, st()))
        t_w = timeit(lambda: lib.conv2d_wgrad(x.data_ptr(), dy.data_ptr(), dw.data_ptr(), None, db.data_ptr(), BF16, N, H, H, Ci, Co, ks, 0, st()))
        print(f"{name}: fprop {t_f:6.1f} us ({mb_f / t_f:5.2f} TB/s)  dgrad {t_d:6.1f} us ({mb_f / t_d:5.2f} TB/s)  "
              f"wgrad+bias {t_w:6.1f} us ({mb_f / t_w:5.2f} TB/s)   [{mb_f:.1f} MB per pass]")
lib.set_tc_enabled(1)
