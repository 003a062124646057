"""Loader for the in-tree CUDA library (citcomcu_b200/csrc/libcitcomcu_b200.so).

There is no CPU fallback: if the library has not been built (`python -c "import
__graft_entry__ as g; g.build()"`), importing the operators raises.
"""
from __future__ import annotations

import ctypes as C
import subprocess
from pathlib import Path

CSRC = Path(__file__).resolve().parent / "csrc"
LIB_PATH = CSRC / "libcitcomcu_b200.so"
MAX_LEVELS = 12

NVCC_FLAGS = ["-O3", "-std=c++17", "-lineinfo", "-gencode", "arch=compute_100a,code=sm_100a",
              "--shared", "-Xcompiler", "-fPIC"]


def build_library(verbose: bool = False) -> Path:
    """Compile every CUDA source of the package for sm_100a into the in-tree shared library."""
    srcs = sorted(str(p) for p in CSRC.glob("*.cu"))
    deps = list(CSRC.glob("*.cu")) + list(CSRC.glob("*.cuh")) + list(CSRC.glob("*.h")) + [CSRC.parent.parent / "include" / "citcomcu_b200.h"]
    if LIB_PATH.exists() and all(LIB_PATH.stat().st_mtime >= d.stat().st_mtime for d in deps):
        return LIB_PATH
    # *_exact.cu reproduce the reference's rounding sequence: no FMA contraction there
    objs = []
    for src in srcs:
        obj = src[:-3] + ".o"
        extra = ["-fmad=false"] if src.endswith("_exact.cu") else []
        cmd = ["nvcc", *[f for f in NVCC_FLAGS if f != "--shared"], *extra, "-c", "-o", obj, src]
        if verbose:
            print(" ".join(cmd))
        subprocess.run(cmd, check=True)
        objs.append(obj)
    cmd = ["nvcc", "--shared", "-gencode", "arch=compute_100a,code=sm_100a", "-o", str(LIB_PATH), *objs]
    if verbose:
        print(" ".join(cmd))
    subprocess.run(cmd, check=True)
    return LIB_PATH


class ccu_config(C.Structure):
    _fields_ = [("levmin", C.c_int), ("levmax", C.c_int),
                ("nox", C.c_int * MAX_LEVELS), ("noy", C.c_int * MAX_LEVELS), ("noz", C.c_int * MAX_LEVELS),
                ("v_steps_low", C.c_int), ("v_steps_high", C.c_int),
                ("down_heavy", C.c_int), ("up_heavy", C.c_int), ("mg_cycle", C.c_int),
                ("p_iterations", C.c_int), ("accuracy", C.c_double), ("device", C.c_int)]


_lib = None


def lib() -> C.CDLL:
    global _lib
    if _lib is None:
        if not LIB_PATH.exists():
            raise RuntimeError(f"{LIB_PATH} is missing: build it with __graft_entry__.build(); "
                               "citcomcu_b200 has no CPU fallback")
        _lib = C.CDLL(str(LIB_PATH))
        _lib.ccu_last_error.restype = C.c_char_p
        _lib.ccu_launch_count.restype = C.c_longlong
        _lib.ccu_launch_count.argtypes = [C.c_void_p]
    return _lib


def check(rc: int) -> None:
    if rc != 0:
        raise RuntimeError(f"libcitcomcu_b200: rc={rc}: {lib().ccu_last_error().decode()}")
