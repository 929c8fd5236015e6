"""ctypes binding of liblss_b200.so (C ABI in include/lss_b200.h).

There is no CPU fallback and no PyTorch fallback: if the CUDA extension is
missing or a call fails, the caller gets an exception.
"""
from __future__ import annotations

import ctypes as C
import os
import threading

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_NAME = "liblss_b200.so"
# LSS_B200_LIB: load another build of the same ABI (kernel-variant experiments, tools/)
LIB_PATH = os.environ.get("LSS_B200_LIB") or os.path.join(_HERE, LIB_NAME)

ABI_VERSION = 3
LSS_F32, LSS_F16, LSS_BF16 = 0, 1, 2


class LssGrid(C.Structure):
    _fields_ = [("dx", C.c_float * 3), ("bx", C.c_float * 3), ("nx", C.c_int32 * 3)]


class LssShape(C.Structure):
    _fields_ = [("B", C.c_int32), ("N", C.c_int32), ("D", C.c_int32),
                ("fH", C.c_int32), ("fW", C.c_int32), ("C", C.c_int32)]


class LssError(RuntimeError):
    def __init__(self, fn: str, status: int, text: str):
        super().__init__("%s failed: status %d (%s)" % (fn, status, text))
        self.status = status


_p = C.c_void_p
_i32 = C.c_int32
_i64 = C.c_int64
_sz = C.c_size_t
_G = C.POINTER(LssGrid)
_S = C.POINTER(LssShape)

# name -> (restype, argtypes); every symbol include/lss_b200.h declares
SIGNATURES = {
    "lss_abi_version": (C.c_int, []),
    "lss_status_string": (C.c_char_p, [C.c_int]),
    "lss_last_cuda_error": (C.c_char_p, []),
    "lss_camera_prep": (C.c_int, [_p, _p, _p, _i32, _p, _p, _p]),
    "lss_quantize_rank": (C.c_int, [_p, _G, _i32, _i64, _p, _p, _p, _p, _p]),
    "lss_geometry_rank": (C.c_int, [_p, _p, _p, _p, _p, _p, _p, _G, _S, _p, _p, _p, _p, _p, _p]),
    "lss_sort_workspace_bytes": (_sz, [_i64, _i32]),
    "lss_sort_ranks": (C.c_int, [_p, _i64, _i32, _p, _p, _p, _sz, _p]),
    "lss_intervals": (C.c_int, [_p, _i64, _G, _i32, _p, _p, _p, _p, _p]),
    "lss_pool_dense_fwd": (C.c_int, [_p, _p, _p, _G, _i32, _i32, _i64, _p, _p]),
    "lss_pool_dense_bwd": (C.c_int, [_p, _p, _G, _i32, _i32, _i64, _p, _p]),
    "lss_feat_stage": (C.c_int, [_p, _i64, _S, _i32, _p, _p]),
    "lss_depth_softmax": (C.c_int, [_p, _i64, _S, _i32, _p, _p]),
    "lss_liftsplat_fwd": (C.c_int, [_p, _i64, _i32, _p, _p, _p, _G, _S, _p, _p]),
    "lss_liftsplat_bwd": (C.c_int, [_p, _p, _i64, _i32, _p, _p, _G, _S, _i32, _i32, _p, _i64, _p, _i64, _p]),
    "lss_plan_workspace_bytes": (_sz, [_S, _G]),
    "lss_plan_workspace_control_bytes": (_sz, [_S, _G]),
    "lss_plan_key_count": (_i64, [_G, _i32]),
    "lss_plan_key_tile": (C.c_int, []),
    "lss_build_plan": (C.c_int, [_p] * 8 + [_G, _S, _p, _p, _p, _p, _p, _sz, _p]),
    "lss_plan_from_geom_workspace_bytes": (_sz, [_i64, _G, _i32]),
    "lss_build_plan_from_geom": (C.c_int, [_p, _G, _i32, _i64, _p, _p, _p, _p, _p, _sz, _p]),
}

_lock = threading.Lock()
_lib = None


def load() -> C.CDLL:
    """Load the extension (once).  Raises if it has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    with _lock:
        if _lib is not None:
            return _lib
        if not os.path.exists(LIB_PATH):
            raise RuntimeError(
                "%s not found: the CUDA extension is not built. Run "
                "`python -c 'import __graft_entry__ as g; g.build()'` (nvcc, sm_100a). "
                "There is no CPU or PyTorch fallback for this path." % LIB_PATH)
        lib = C.CDLL(LIB_PATH)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(lib, name)  # AttributeError if the .so lacks a declared symbol
            fn.restype = res
            fn.argtypes = args
        if lib.lss_abi_version() != ABI_VERSION:
            raise RuntimeError("liblss_b200.so ABI version %d, expected %d" % (lib.lss_abi_version(), ABI_VERSION))
        _lib = lib
    return _lib


def check(fn: str, status: int) -> None:
    if status == 0:
        return
    lib = load()
    text = lib.lss_status_string(status).decode()
    if status == -6:
        text += ": " + lib.lss_last_cuda_error().decode()
    raise LssError(fn, status, text)


def call(name: str, *args) -> None:
    """Invoke a status-returning entry point and raise on failure."""
    check(name, getattr(load(), name)(*args))


def make_grid(dx, bx, nx) -> LssGrid:
    g = LssGrid()
    for i in range(3):
        g.dx[i] = float(dx[i])
        g.bx[i] = float(bx[i])
        g.nx[i] = int(nx[i])
    return g


def make_shape(B, N, D, fH, fW, Cc) -> LssShape:
    return LssShape(int(B), int(N), int(D), int(fH), int(fW), int(Cc))
