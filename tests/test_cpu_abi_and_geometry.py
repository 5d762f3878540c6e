"""CPU tests (-m "not gpu"): the C-ABI library loads and exports every symbol the header declares, and the
host-side conv geometry (the multi-tap GEMM tables every conv form is lowered to) is correct - checked by
emulating the tap GEMM in numpy from the tables the library itself produces and comparing with the
first-principles conv definitions of oracle/np_primitives.  No device work is launched."""
import ctypes
import os

import numpy as np
import pytest

from helpers import lib
from svrs_native import lib as libmod
from oracle import np_primitives as npp


def test_header_symbols_exported():
    protos = libmod.parse_header()
    assert len(protos) >= 30
    dll = ctypes.CDLL(libmod.LIB_PATH)
    for name in protos:
        assert hasattr(dll, name), f"{name} declared in include/svrs_b200.h but not exported"
    assert lib.abi_version() == 1


def test_header_constants_match_host_code():
    """Constants the Python host mirrors from the header stay in sync with it."""
    import re
    from svrs_native import engine
    hdr = open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "include", "svrs_b200.h")).read()
    assert int(re.search(r"#define\s+SVRS_BN_REPLICAS\s+(\d+)", hdr).group(1)) == engine.BN_REPLICAS


def test_adam_tile_geometry_is_host_callable():
    """The tile geometry svrs_adam_multi's job table is built with (trainer._adam_table): pure host
    functions, callable without a GPU.  A tile's dense m / v / p rows must fit the kernel's 27 KB shared-memory stage and
    a tile row must be a whole number of 16-byte pieces (bulk-copy granularity) for the 3x3 and 4x4 kernels."""
    assert lib.adam_job_bytes() == 48
    rows = lib.adam_tile_rows()
    assert rows == 16
    for kk in (1, 4, 9, 16):
        cols = lib.adam_tile_cols(kk)
        assert cols in (8, 16)
        assert 3 * rows * cols * kk * 4 <= 27648
    for kk in (9, 16):
        assert (lib.adam_tile_cols(kk) * kk * 4) % 16 == 0


def test_header_cites_reference_call_sites():
    src = open(libmod.HEADER).read()
    for cite in ("layers.py:231-236", "layers.py:275-277", "cond_vae.py:261-265", "loss/cond_vae_loss.py:39-58",
                 "models/base.py:106-107", "dataset.py:220-247", "utils.py:4-23"):
        assert cite in src, f"include/svrs_b200.h lost its citation of {cite}"


def _geom(form, N, H, W, Cr, Cw):
    buf = (ctypes.c_int64 * 512)()
    n = lib.debug_tap_geometry(form, N, H, W, Cr, Cw, ctypes.cast(buf, ctypes.c_void_p), 512)
    assert n > 0, lib.last_error()
    v = list(buf[:n])
    g = dict(zip(("N", "OH", "OW", "IH", "IW", "K", "Nc", "nprob", "o_sn", "o_sy", "o_sx", "i_sn", "i_sy", "i_sx"), v[:14]))
    i, probs = 14, []
    for _ in range(g["nprob"]):
        out_off, ntaps = v[i], v[i + 1]
        i += 2
        taps = []
        for _ in range(ntaps):
            taps.append(tuple(v[i:i + 4]))
            i += 4
        probs.append((out_off, taps))
    g["probs"] = probs
    return g


def _run_taps(g, x_flat, wpack, out_size):
    """numpy emulation of conv_taps_kernel: x_flat = NHWC input flattened, wpack = [tap][K][Nc] flattened."""
    out = np.zeros(out_size)
    K, Nc = g["K"], g["Nc"]
    for out_off, taps in g["probs"]:
        for n in range(g["N"]):
            for oy in range(g["OH"]):
                for ox in range(g["OW"]):
                    acc = np.zeros(Nc)
                    for in_off, w_off, dy, dx in taps:
                        iy, ix = oy + dy, ox + dx
                        if 0 <= iy < g["IH"] and 0 <= ix < g["IW"]:
                            s = in_off + n * g["i_sn"] + iy * g["i_sy"] + ix * g["i_sx"]
                            acc += x_flat[s:s + K] @ wpack[w_off:w_off + K * Nc].reshape(K, Nc)
                    o = out_off + n * g["o_sn"] + oy * g["o_sy"] + ox * g["o_sx"]
                    out[o:o + Nc] = acc
    return out


def _nhwc(a):
    return np.ascontiguousarray(a.transpose(0, 2, 3, 1))


@pytest.mark.parametrize("N,H,W,Ci,Co", [(2, 4, 6, 3, 5), (1, 2, 2, 4, 2)])
def test_tap_geometry_all_conv_forms(N, H, W, Ci, Co):
    rng = np.random.default_rng(0)
    x = rng.standard_normal((N, Ci, H, W))
    zero_b = np.zeros(Co)
    # conv3 fprop: pack [t][ci][co]
    w = rng.standard_normal((Co, Ci, 3, 3))
    ref = npp.conv2d(x, w, zero_b, 1, 1)
    pk = np.ascontiguousarray(w.reshape(Co, Ci, 9).transpose(2, 1, 0)).reshape(-1)
    got = _run_taps(_geom(0, N, H, W, Ci, Co), _nhwc(x).reshape(-1), pk, N * H * W * Co)
    np.testing.assert_allclose(got.reshape(N, H, W, Co), _nhwc(ref), atol=1e-12)
    # conv3 dgrad: dx = sum dy[.., co] w[co, ci] at flipped taps; pack [t][co][ci]
    gy = rng.standard_normal(ref.shape)
    import torch
    xt = torch.from_numpy(x).requires_grad_(True)
    torch.nn.functional.conv2d(xt, torch.from_numpy(w), None, 1, 1).backward(torch.from_numpy(gy))
    pk = np.ascontiguousarray(w.reshape(Co, Ci, 9).transpose(2, 0, 1)).reshape(-1)
    got = _run_taps(_geom(1, N, H, W, Co, Ci), _nhwc(gy).reshape(-1), pk, N * H * W * Ci)
    np.testing.assert_allclose(got.reshape(N, H, W, Ci), _nhwc(xt.grad.numpy()), atol=1e-12)
    # conv4 s2 fprop
    w4 = rng.standard_normal((Co, Ci, 4, 4))
    ref = npp.conv2d(x, w4, zero_b, 2, 1)
    pk = np.ascontiguousarray(w4.reshape(Co, Ci, 16).transpose(2, 1, 0)).reshape(-1)
    got = _run_taps(_geom(2, N, H, W, Ci, Co), _nhwc(x).reshape(-1), pk, N * (H // 2) * (W // 2) * Co)
    np.testing.assert_allclose(got.reshape(N, H // 2, W // 2, Co), _nhwc(ref), atol=1e-12)
    # conv4 s2 dgrad == transposed form reading dy [N,H/2,W/2,Co], pack [t][co][ci]
    gy = rng.standard_normal(ref.shape)
    xt = torch.from_numpy(x).requires_grad_(True)
    torch.nn.functional.conv2d(xt, torch.from_numpy(w4), None, 2, 1).backward(torch.from_numpy(gy))
    pk = np.ascontiguousarray(w4.reshape(Co, Ci, 16).transpose(2, 0, 1)).reshape(-1)
    got = _run_taps(_geom(3, N, H // 2, W // 2, Co, Ci), _nhwc(gy).reshape(-1), pk, N * H * W * Ci)
    np.testing.assert_allclose(got.reshape(N, H, W, Ci), _nhwc(xt.grad.numpy()), atol=1e-12)
    # convT fprop: weight [Ci][Co][16], pack [t][ci][co]
    wt = rng.standard_normal((Ci, Co, 4, 4))
    ref = npp.conv_transpose2d_k4s2p1(x, wt, zero_b)
    pk = np.ascontiguousarray(wt.reshape(Ci, Co, 16).transpose(2, 0, 1)).reshape(-1)
    got = _run_taps(_geom(3, N, H, W, Ci, Co), _nhwc(x).reshape(-1), pk, N * 4 * H * W * Co)
    np.testing.assert_allclose(got.reshape(N, 2 * H, 2 * W, Co), _nhwc(ref), atol=1e-12)
    # convT dgrad == strided form reading dy [N,2H,2W,Co], pack [t][co][ci]
    gy = rng.standard_normal(ref.shape)
    xt = torch.from_numpy(x).requires_grad_(True)
    torch.nn.functional.conv_transpose2d(xt, torch.from_numpy(wt), None, 2, 1).backward(torch.from_numpy(gy))
    pk = np.ascontiguousarray(wt.reshape(Ci, Co, 16).transpose(2, 1, 0)).reshape(-1)
    got = _run_taps(_geom(2, N, 2 * H, 2 * W, Co, Ci), _nhwc(gy).reshape(-1), pk, N * H * W * Ci)
    np.testing.assert_allclose(got.reshape(N, H, W, Ci), _nhwc(xt.grad.numpy()), atol=1e-12)


def test_product_path_has_no_cpu_fallback():
    import torch
    import models
    from svrs_native.lib import SvrsError
    m = models.VAE(2, 32)
    with pytest.raises(SvrsError, match="no CPU fallback|CUDA"):
        m(torch.rand(1, 4, 32, 32))
    from loss import base_loss
    with pytest.raises(SvrsError, match="CUDA"):
        base_loss(torch.rand(1, 4, 8, 8), torch.rand(1, 4, 8, 8), torch.rand(1, 8), torch.rand(1, 8), torch.tensor(1.0))
    with pytest.raises(RuntimeError, match="no CPU path"):
        m.fit(train_loader=[(torch.rand(1, 4, 32, 32),) * 2], val_loader=[], device="cpu",
              optimizer=torch.optim.Adam(m.parameters()), epochs=1)


def test_product_never_imports_oracle():
    """The oracle is test infrastructure: nothing under simple-vae-rs_b200/ may reference it."""
    pkg = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "simple-vae-rs_b200")
    for root, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                src = open(os.path.join(root, f)).read()
                assert "oracle" not in src.replace("oracle/", "ORACLE_DOC/") or "import oracle" not in src and "from oracle" not in src, f
                assert "from oracle" not in src and "import oracle" not in src, f"{f} imports the oracle"
