"""ctypes binding of the C-ABI kernel library (include/svrs_b200.h -> libsvrs_b200.so).

The signatures are parsed from the header itself, so the header is the single source of truth for the
boundary.  There is NO fallback: if the library is missing or a call fails, this raises.
"""
from __future__ import annotations

import ctypes
import os
import re
from typing import Dict, List, Tuple

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
HEADER = os.path.join(ROOT, "include", "svrs_b200.h")
LIB_PATH = os.path.join(HERE, "libsvrs_b200.so")

F32, BF16 = 0, 1
ACT_NONE, ACT_SIGMOID, ACT_HARDTANH7 = 0, 1, 2

_SCALARS = {
    "int": ctypes.c_int,
    "int64_t": ctypes.c_int64,
    "uint64_t": ctypes.c_uint64,
    "uint32_t": ctypes.c_uint32,
    "float": ctypes.c_float,
}


class SvrsError(RuntimeError):
    pass


class SvrsUnsupported(SvrsError):
    """SVRS_E_UNSUPPORTED (-3): the kernel that takes this shape cannot provide a requested optional extra; nothing was
    enqueued.  Callers use the separate kernels instead (never a CPU path)."""


def parse_header(path: str = HEADER) -> Dict[str, Tuple[str, List[Tuple[str, str]]]]:
    """-> {name: (return_type, [(ctype, argname), ...])} for every `svrs_*` prototype."""
    src = open(path).read()
    src = re.sub(r"/\*.*?\*/", " ", src, flags=re.S)
    src = re.sub(r"//[^\n]*", " ", src)
    protos = {}
    for m in re.finditer(r"(const\s+char\s*\*|int64_t|int|void)\s+(svrs_\w+)\s*\(([^;{]*?)\)\s*;", src, flags=re.S):
        ret, name, args = m.group(1), m.group(2), m.group(3)
        arglist = []
        args = " ".join(args.split())
        if args and args != "void":
            for a in args.split(","):
                a = a.strip()
                mm = re.match(r"(.*?)(\w+)$", a)
                ty, an = mm.group(1).strip(), mm.group(2)
                arglist.append((ty.replace(" *", "*"), an))
        protos[name] = ("char*" if "char" in ret else ret.strip(), arglist)
    return protos


def _ctype(ty: str):
    if "*" in ty:
        return ctypes.c_void_p
    ty = ty.replace("const", "").strip()
    return _SCALARS[ty]


class _Lib:
    def __init__(self):
        self._dll = None
        self.protos = parse_header()
        self.timing = None      # when a list: every call is bracketed by CUDA events (svrs_native.profile)
        self.timing_pad_cycles = 0

    def load(self):
        if self._dll is not None:
            return self._dll
        if not os.path.exists(LIB_PATH):
            raise SvrsError(
                f"{LIB_PATH} is missing: the sm_100a kernel library has not been built "
                f"(run `python simple-vae-rs_b200/svrs_native/build.py`). There is no CPU/eager fallback.")
        dll = ctypes.CDLL(LIB_PATH)
        for name, (ret, args) in self.protos.items():
            fn = getattr(dll, name)  # AttributeError => header/library mismatch, fail loudly
            fn.argtypes = [_ctype(t) for t, _ in args]
            fn.restype = {"char*": ctypes.c_char_p, "void": None, "int64_t": ctypes.c_int64}.get(ret, ctypes.c_int)
        self._dll = dll
        return dll

    def last_error(self) -> str:
        return self.load().svrs_last_error().decode()

    def __getattr__(self, name):
        # lib.conv2d_fprop(...) -> svrs_conv2d_fprop(...), raising on a non-zero status
        full = "svrs_" + name
        if full not in self.protos:
            raise AttributeError(name)
        fn = getattr(self.load(), full)
        if self.protos[full][0] != "int" or full in ("svrs_abi_version", "svrs_device_cc", "svrs_debug_tap_geometry", "svrs_tc_would_run", "svrs_pack_job_bytes", "svrs_launch_count", "svrs_adam_job_bytes", "svrs_adam_tile_rows", "svrs_adam_tile_cols",
                                                  "svrs_conv2d_wgrad_layout", "svrs_convT2d_wgrad_layout", "svrs_sample_tail_splits"):
            setattr(self, name, fn)
            return fn

        def call(*a, _fn=fn, _n=full):
            if self.timing is not None:
                import torch
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                if self.timing_pad_cycles:
                    # a short device-side spin ahead of the bracket hides the host launch latency of the timed kernel,
                    # so e0 -> e1 is pure device time even for microsecond kernels
                    torch.cuda._sleep(self.timing_pad_cycles)
                dll = self.load()
                dll.svrs_trace_reset(1)
                e0.record()
                rc = _fn(*a)
                e1.record()
                kernels = dll.svrs_trace().decode()
                dll.svrs_trace_reset(0)
                self.timing.append((_n, a, e0, e1, kernels))
            else:
                rc = _fn(*a)
            if rc == -3:
                raise SvrsUnsupported(f"{_n}: {self.last_error()}")
            if rc != 0:
                raise SvrsError(f"{_n} failed ({rc}): {self.last_error()}")

        setattr(self, name, call)
        return call


lib = _Lib()


def require_cuda_library():
    """Fail loudly unless the native library loads (used by the product path at import of the engine)."""
    lib.load()
    return LIB_PATH
