"""ctypes loader for oracle/_build/liblss_oracle.so (C restatement; TEST INFRASTRUCTURE)."""
import ctypes as C
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIB = os.path.join(HERE, "_build", "liblss_oracle.so")
_lib = None


def load():
    global _lib
    if _lib is None:
        if not os.path.exists(LIB) or os.path.getmtime(LIB) < os.path.getmtime(os.path.join(HERE, "lss_oracle.c")):
            subprocess.run(["make", "-s", "-C", HERE], check=True)
        _lib = C.CDLL(LIB)
        _lib.lss_oracle_quantize.restype = C.c_int64
    return _lib


def _p(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


def threads():
    return load().lss_oracle_threads()


def set_threads(n):
    load().lss_oracle_set_threads(int(n))


def inverse3x3(A):
    A = np.ascontiguousarray(A, np.float32)
    X = np.empty_like(A)
    load().lss_oracle_inverse3x3(_p(A), C.c_int(A.size // 9), _p(X))
    return X


def camera_prep(rots, intrins, post_rots):
    rots, intrins, post_rots = (np.ascontiguousarray(a, np.float32) for a in (rots, intrins, post_rots))
    ipr = np.empty_like(post_rots); comb = np.empty_like(rots)
    load().lss_oracle_camera_prep(_p(rots), _p(intrins), _p(post_rots), C.c_int(rots.size // 9), _p(ipr), _p(comb))
    return ipr, comb


def geometry(us, vs, ds, rots, trans, intrins, post_rots, post_trans):
    B, N = trans.shape[:2]
    ipr, comb = camera_prep(rots, intrins, post_rots)
    us, vs, ds = (np.ascontiguousarray(a, np.float32) for a in (us, vs, ds))
    geom = np.empty((B, N, len(ds), len(vs), len(us), 3), np.float32)
    load().lss_oracle_geometry(_p(us), _p(vs), _p(ds), C.c_int(len(ds)), C.c_int(len(vs)), C.c_int(len(us)),
                               _p(ipr), _p(np.ascontiguousarray(post_trans, np.float32)), _p(comb),
                               _p(np.ascontiguousarray(trans, np.float32)), C.c_int(B * N), _p(geom))
    return geom


def index(geom, dx, bx, nx, B):
    geom = np.ascontiguousarray(geom, np.float32).reshape(-1, 3)
    P = geom.shape[0]
    dx = np.ascontiguousarray(dx, np.float32); bx = np.ascontiguousarray(bx, np.float32)
    nx = np.ascontiguousarray(nx, np.int64)
    coords = np.empty((P, 3), np.int64); kept = np.empty(P, np.uint8); ranks_all = np.empty(P, np.int64)
    K = load().lss_oracle_quantize(_p(geom), C.c_int64(P), C.c_int(B), _p(dx), _p(bx), _p(nx), _p(coords),
                                   _p(kept), _p(ranks_all))
    kept_idx = np.empty(K + 1, np.int64); ranks = np.empty(K + 1, np.int64); sorts = np.empty(K + 1, np.int64)
    load().lss_oracle_sort(_p(ranks_all), C.c_int64(P), C.c_int64(int(nx[0] * nx[1] * nx[2]) * B), _p(kept_idx),
                           _p(ranks), _p(sorts))
    return {"coords": coords, "kept": kept.astype(bool), "kept_idx": kept_idx[:K], "ranks": ranks[:K],
            "sorts": sorts[:K]}


def step(depth, feat, geom, dbev, dx, bx, nx, B, N, mode=0, backward=True):
    """One fwd(+bwd) of lift + voxel_pooling as the reference does it; returns (bev, ddepth, dfeat, K, V)."""
    depth = np.ascontiguousarray(depth, np.float32); feat = np.ascontiguousarray(feat, np.float32)
    geom = np.ascontiguousarray(geom, np.float32)
    BN, D, fH, fW = depth.shape
    Cc = feat.shape[1]
    dx = np.ascontiguousarray(dx, np.float32); bx = np.ascontiguousarray(bx, np.float32)
    nx = np.ascontiguousarray(nx, np.int64)
    dt = np.float64 if mode else np.float32
    bev = np.empty((B, Cc * int(nx[2]), int(nx[0]), int(nx[1])), dt)
    dd = np.empty(depth.shape, dt) if backward else None
    df = np.empty(feat.shape, dt) if backward else None
    dbev_c = np.ascontiguousarray(dbev, np.float32) if backward else None
    counts = np.zeros(2, np.int64)
    load().lss_oracle_step(_p(depth), _p(feat), _p(geom), _p(dbev_c), C.c_int(B), C.c_int(N), C.c_int(D),
                           C.c_int(fH), C.c_int(fW), C.c_int(Cc), _p(dx), _p(bx), _p(nx), C.c_int(mode),
                           _p(bev), _p(dd), _p(df), _p(counts))
    return bev, dd, df, int(counts[0]), int(counts[1])
