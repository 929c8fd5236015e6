"""Host side of the Lift-Splat hot path: torch tensors in, C-ABI calls out.

Every function here enqueues hand-written sm_100a kernels from liblss_b200.so on
torch's current CUDA stream.  PyTorch supplies device memory, streams and
autograd plumbing only; there is no CPU path and no PyTorch fallback -- a
non-CUDA tensor raises.

Mapping to the reference (paths relative to the reference root):
  GridSpec            gen_dx_bx                     src/tools.py:172-178
  camera_prep         torch.inverse / matmul        src/model_baseline.py:60,66
  geometry            get_geometry                  src/model_baseline.py:50-70
  quantize_rank       quantise, kept mask, ranks    src/model_baseline.py:92-109
  sort_ranks          ranks.argsort()               src/model_baseline.py:110
  intervals           QuickCumsum boundary mask     src/tools.py:196-197
  build_plan          geometry half of get_voxels   src/model_baseline.py:128-131, :110,
                                                    src/tools.py:196-197
  lift_splat          lift + voxel_pooling fwd/bwd  src/modules.py:84,
                                                    src/model_baseline.py:84-126,
                                                    src/tools.py:192-218
  pool_dense          voxel_pooling on a dense x    src/model_baseline.py:84-126
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import Dict, Optional, Sequence, Tuple

import torch

from . import _abi


# --------------------------------------------------------------------------
# small helpers
# --------------------------------------------------------------------------
def _ptr(t: Optional[torch.Tensor]):
    return None if t is None else t.data_ptr()


_raw_stream = getattr(torch._C, "_cuda_getCurrentRawStream", None)


def _stream(device) -> int:
    """cudaStream_t of torch's current stream on ``device`` (the raw getter is ~10x cheaper than building a
    torch.cuda.Stream object; the eager module path asks a dozen times per step)."""
    if _raw_stream is not None and device.index is not None:
        return _raw_stream(device.index)
    return torch.cuda.current_stream(device).cuda_stream


class _NoGuard:
    def __enter__(self):
        return None

    def __exit__(self, *a):
        return False


_NO_GUARD = _NoGuard()


def _device_guard(dev):
    """torch.cuda.device(dev) only when dev is not already the current device."""
    return _NO_GUARD if torch.cuda.current_device() == dev.index else torch.cuda.device(dev)


def _call(dev, name: str, *args) -> None:
    """One C-ABI call with ``dev`` as the current CUDA device.  The library launches on the current
    device (d_* pointers are "device pointers on the current CUDA device", include/lss_b200.h); the
    reference scripts move the model with ``.to(f'cuda:{gpuid}')`` and never call set_device
    (train.py:34, predict.py:35), so the tensors' device is what counts, not torch's current one."""
    if torch.cuda.current_device() == dev.index:
        _abi.call(name, *args)
    else:
        with torch.cuda.device(dev):
            _abi.call(name, *args)


def _need_cuda(*tensors: torch.Tensor) -> torch.device:
    dev = None
    for t in tensors:
        if t is None:
            continue
        if not t.is_cuda:
            raise RuntimeError(
                "lss2_multimodal_nu_b200: tensor on %s -- this path runs only on CUDA "
                "(sm_100a kernels); there is no CPU fallback" % t.device)
        if dev is None:
            dev = t.device
        elif t.device != dev:
            raise RuntimeError("tensors on different devices: %s vs %s" % (dev, t.device))
    return dev


_DTYPES = {torch.float32: _abi.LSS_F32, torch.float16: _abi.LSS_F16, torch.bfloat16: _abi.LSS_BF16}


def _batch_dense(t: torch.Tensor) -> bool:
    """True when every batch item of a (BN, R, H, W) tensor is dense (only the batch stride is free),
    i.e. the tensor is contiguous or a channel slice of a contiguous conv output."""
    BN, R, H, W = t.shape
    return t.stride(3) == 1 and t.stride(2) == W and t.stride(1) == H * W and t.stride(0) >= R * H * W


def _feature_input(t: torch.Tensor) -> torch.Tensor:
    """A feature tensor the staging kernel can read in place: float32 / float16 / bfloat16, dense per
    batch item.  Anything else is converted / copied."""
    if t.dtype not in _DTYPES:
        t = t.float()
    return t if _batch_dense(t) else t.contiguous()


def _f32c(t: torch.Tensor) -> torch.Tensor:
    """float32 + contiguous (no copy when already so)."""
    if t.dtype == torch.float32 and t.is_contiguous():
        return t
    if t.dtype != torch.float32:
        t = t.float()
    return t.contiguous()


# ctypes structs and size queries are pure functions of small tuples: built once (the eager module path
# issues a dozen C calls per step; their Python-side preparation is what bounds it)
_GRID_C: Dict[Tuple, "_abi.LssGrid"] = {}
_SHAPE_C: Dict[Tuple, "_abi.LssShape"] = {}
_PLAN_SIZES: Dict[Tuple, Tuple[int, int, int]] = {}


def _shape_c(B, N, D, fH, fW, C) -> "_abi.LssShape":
    key = (B, N, D, fH, fW, C)
    s = _SHAPE_C.get(key)
    if s is None:
        s = _SHAPE_C[key] = _abi.make_shape(*key)
    return s


def _plan_sizes(grid: "GridSpec", B, N, D, fH, fW) -> Tuple[int, int, int]:
    """(workspace bytes, control bytes, n_keys) of a plan."""
    key = (grid.dx, grid.bx, grid.nx, B, N, D, fH, fW)
    v = _PLAN_SIZES.get(key)
    if v is None:
        lib = _abi.load()
        shape, g = _shape_c(B, N, D, fH, fW, 4), grid.c()
        v = _PLAN_SIZES[key] = (int(lib.lss_plan_workspace_bytes(shape, g)),
                                int(lib.lss_plan_workspace_control_bytes(shape, g)),
                                int(lib.lss_plan_key_count(g, B)))
    return v


@dataclass(frozen=True)
class GridSpec:
    """Host copy of the grid constants gen_dx_bx produces."""
    dx: Tuple[float, float, float]
    bx: Tuple[float, float, float]
    nx: Tuple[int, int, int]

    @staticmethod
    def from_bounds(xbound, ybound, zbound) -> "GridSpec":
        # same arithmetic as reference src/tools.py:173-175; torch.Tensor(...) rounds
        # the python floats to float32
        rows = [xbound, ybound, zbound]
        dx = torch.tensor([r[2] for r in rows], dtype=torch.float64).float().tolist()
        bx = torch.tensor([r[0] + r[2] / 2.0 for r in rows], dtype=torch.float64).float().tolist()
        nx = [int((r[1] - r[0]) / r[2]) for r in rows]
        return GridSpec(tuple(dx), tuple(bx), tuple(nx))

    @staticmethod
    def from_tensors(dx, bx, nx) -> "GridSpec":
        """From the module's dx/bx/nx parameters (one device->host read; cache the result)."""
        return GridSpec(tuple(float(v) for v in dx.detach().cpu().tolist()),
                        tuple(float(v) for v in bx.detach().cpu().tolist()),
                        tuple(int(v) for v in nx.detach().cpu().tolist()))

    def c(self) -> _abi.LssGrid:
        key = (self.dx, self.bx, self.nx)
        g = _GRID_C.get(key)
        if g is None:
            g = _GRID_C[key] = _abi.make_grid(self.dx, self.bx, self.nx)
        return g

    def n_cells(self, B: int) -> int:
        return self.nx[0] * self.nx[1] * self.nx[2] * B


def make_frustum(final_dim, downsample: int, dbound) -> torch.Tensor:
    """The (D, fH, fW, 3) table of (u, v, depth) the models keep as their ``frustum`` parameter: the same
    torch calls as the reference's create_frustum (src/model_baseline.py:37-48), for callers that have no
    model instance at hand (benchmarks, tests)."""
    ogfH, ogfW = final_dim
    fH, fW = ogfH // downsample, ogfW // downsample
    ds = torch.arange(*dbound, dtype=torch.float).view(-1, 1, 1).expand(-1, fH, fW)
    D = ds.shape[0]
    xs = torch.linspace(0, ogfW - 1, fW, dtype=torch.float).view(1, 1, fW).expand(D, fH, fW)
    ys = torch.linspace(0, ogfH - 1, fH, dtype=torch.float).view(1, fH, 1).expand(D, fH, fW)
    return torch.stack((xs, ys, ds), -1)


def frustum_axes(frustum: torch.Tensor):
    """(us[fW], vs[fH], ds[D]) from the module's (D,fH,fW,3) frustum parameter
    (reference src/model_baseline.py:41-47): the frustum is the outer product of
    these three tables, so they carry its exact bits."""
    us = frustum[0, 0, :, 0].detach().float().contiguous()
    vs = frustum[0, :, 0, 1].detach().float().contiguous()
    ds = frustum[:, 0, 0, 2].detach().float().contiguous()
    return us, vs, ds


# --------------------------------------------------------------------------
# K0 / K1 / K1' / K2 / K3 as individual operators (parity surface)
# --------------------------------------------------------------------------
def camera_prep(rots, intrins, post_rots):
    """inverse(post_rots), rots @ inverse(intrins) -- (..., 3, 3) float32 each."""
    dev = _need_cuda(rots, intrins, post_rots)
    rots, intrins, post_rots = _f32c(rots), _f32c(intrins), _f32c(post_rots)
    n = rots.numel() // 9
    ipr = torch.empty_like(post_rots)
    comb = torch.empty_like(rots)
    _call(dev, "lss_camera_prep", _ptr(rots), _ptr(intrins), _ptr(post_rots), n, _ptr(ipr),
              _ptr(comb), _stream(dev))
    return ipr, comb


def geometry(us, vs, ds, rots, trans, intrins, post_rots, post_trans, grid: GridSpec,
             inv_post_rots=None, combine=None, want_geom=True, want_coords=False,
             want_kept=False) -> Dict[str, torch.Tensor]:
    """Fused frustum geometry -> rank.  Returns a dict with 'ranks' (P) int32,
    'cells' (P) int32 and, on request, 'geom' (B,N,D,fH,fW,3), 'coords' (P,3)
    int32, 'kept' (P) uint8.  ``inv_post_rots`` / ``combine`` override K0."""
    dev = _need_cuda(us, vs, ds, rots, trans, intrins, post_rots, post_trans)
    B, N = trans.shape[0], trans.shape[1]
    D, fH, fW = ds.numel(), vs.numel(), us.numel()
    if inv_post_rots is None or combine is None:
        ipr, comb = camera_prep(rots, intrins, post_rots)
        inv_post_rots = ipr if inv_post_rots is None else inv_post_rots
        combine = comb if combine is None else combine
    P = B * N * D * fH * fW
    shape = _abi.make_shape(B, N, D, fH, fW, 4)
    out = {"ranks": torch.empty(P, dtype=torch.int32, device=dev),
           "cells": torch.empty(P, dtype=torch.int32, device=dev)}
    if want_geom:
        out["geom"] = torch.empty((B, N, D, fH, fW, 3), dtype=torch.float32, device=dev)
    if want_coords:
        out["coords"] = torch.empty((P, 3), dtype=torch.int32, device=dev)
    if want_kept:
        out["kept"] = torch.empty(P, dtype=torch.uint8, device=dev)
    g = grid.c()
    _call(dev, "lss_geometry_rank", _ptr(_f32c(us)), _ptr(_f32c(vs)), _ptr(_f32c(ds)),
              _ptr(_f32c(inv_post_rots)), _ptr(_f32c(post_trans)), _ptr(_f32c(combine)),
              _ptr(_f32c(trans)), g, shape, _ptr(out.get("geom")), _ptr(out.get("coords")),
              _ptr(out.get("kept")), _ptr(out["ranks"]), _ptr(out["cells"]), _stream(dev))
    return out


def quantize_rank(geom: torch.Tensor, grid: GridSpec, B: int, want_coords=False,
                  want_kept=False) -> Dict[str, torch.Tensor]:
    """Quantise a dense geometry tensor (..., 3) whose leading dim is the batch."""
    dev = _need_cuda(geom)
    geom = _f32c(geom)
    P = geom.numel() // 3
    out = {"ranks": torch.empty(P, dtype=torch.int32, device=dev),
           "cells": torch.empty(P, dtype=torch.int32, device=dev)}
    if want_coords:
        out["coords"] = torch.empty((P, 3), dtype=torch.int32, device=dev)
    if want_kept:
        out["kept"] = torch.empty(P, dtype=torch.uint8, device=dev)
    _call(dev, "lss_quantize_rank", _ptr(geom), grid.c(), B, P, _ptr(out.get("coords")),
              _ptr(out.get("kept")), _ptr(out["ranks"]), _ptr(out["cells"]), _stream(dev))
    return out


def sort_ranks(ranks: torch.Tensor, n_cells: int):
    """Stable sort of int32 ranks -> (sorted_ranks, sorted_points)."""
    dev = _need_cuda(ranks)
    assert ranks.dtype == torch.int32 and ranks.is_contiguous()
    P = ranks.numel()
    nbytes = _abi.load().lss_sort_workspace_bytes(P, n_cells)
    ws = torch.empty(nbytes, dtype=torch.uint8, device=dev)
    sk = torch.empty_like(ranks)
    sp = torch.empty_like(ranks)
    _call(dev, "lss_sort_ranks", _ptr(ranks), P, n_cells, _ptr(sk), _ptr(sp), _ptr(ws), nbytes,
              _stream(dev))
    return sk, sp


def intervals(sorted_ranks: torch.Tensor, grid: GridSpec, B: int, want_last_mask=False):
    """Run detection -> (cell_range (n_cells,2) int32, counts (2) int32 [K, V], last_mask|None,
    sorted_cells (P) int32)."""
    dev = _need_cuda(sorted_ranks)
    P = sorted_ranks.numel()
    cell_range = torch.zeros((grid.n_cells(B), 2), dtype=torch.int32, device=dev)
    counts = torch.zeros(2, dtype=torch.int32, device=dev)
    last = torch.empty(P, dtype=torch.uint8, device=dev) if want_last_mask else None
    sorted_cells = torch.empty(P, dtype=torch.int32, device=dev)
    _call(dev, "lss_intervals", _ptr(sorted_ranks), P, grid.c(), B, _ptr(last), _ptr(sorted_cells),
              _ptr(cell_range), _ptr(counts), _stream(dev))
    return cell_range, counts, last, sorted_cells


# --------------------------------------------------------------------------
# the plan: everything that depends only on the calibration
# --------------------------------------------------------------------------
@dataclass
class Plan:
    """Per-batch index tables: depends on calibration + grid only, never on features.

    The point list is sorted stably by the output cell's TILE-MAJOR key (keys_of_cells below;
    include/lss_b200.h): the digits (b, x, y, z) of the reference's rank
    (src/model_baseline.py:106-109) regrouped as (b, x//T, y//T, x%T, y%T, z), a bijection of the
    rank, so every voxel's run holds the reference's points in the reference's order (see
    reference_order()) and neighbours in the list are neighbours on the map."""
    grid: GridSpec
    B: int
    N: int
    D: int
    fH: int
    fW: int
    cells: torch.Tensor          # (P) int32 output cell of each point, -1 if dropped
    key_start: torch.Tensor      # (n_keys + 1) int32: key k owns sorted_rec[key_start[k]:key_start[k+1]]
    sorted_rec: torch.Tensor     # (P, 2) int32 {output cell, point id} ordered by (key, point id); {-1, 0} beyond K
    counts: torch.Tensor         # (2) int32 {K, V}

    @property
    def P(self) -> int:
        return self.B * self.N * self.D * self.fH * self.fW

    @property
    def sorted_points(self) -> torch.Tensor:
        """(P) point ids ordered by (key, point id); the first K entries are valid."""
        return self.sorted_rec[:, 1]

    @property
    def sorted_cells(self) -> torch.Tensor:
        """(P) output cell of each sorted point, -1 beyond the K kept points."""
        return self.sorted_rec[:, 0]

    def shape(self, C: int) -> _abi.LssShape:
        return _shape_c(self.B, self.N, self.D, self.fH, self.fW, C)

    def keys_of_cells(self, cells: torch.Tensor) -> torch.Tensor:
        """Tile-major sort key of output cells ((b*X + x)*Y + y)*Z + z (int64 tensor in, int64 out)."""
        return keys_of_cells(cells, self.grid)

    def reference_order(self) -> torch.Tensor:
        """The kept points in the order of the reference's ``ranks.argsort()``
        (src/model_baseline.py:110): the plan's order with the batch digit moved back to the
        least-significant place, via lss_sort_ranks (K2) on the reference's ranks."""
        K = int(self.counts[0])
        X, Y, Z = self.grid.nx
        c = self.cells.long()
        b, rest = c // (X * Y * Z), c % (X * Y * Z)
        n_cells = self.grid.n_cells(self.B)
        ranks = torch.where(c >= 0, rest * self.B + b, torch.full_like(c, n_cells)).int()
        _, sp = sort_ranks(ranks.contiguous(), n_cells)
        return sp[:K]


def key_tile() -> int:
    return int(_abi.load().lss_plan_key_tile())


def keys_of_cells(cells, grid: GridSpec):
    """key = ((((b*XT + x//T)*YT + y//T)*T + x%T)*T + y%T)*Z + z for cell = ((b*X + x)*Y + y)*Z + z.
    Works on torch tensors and numpy arrays (integer arithmetic only)."""
    X, Y, Z = grid.nx
    T = key_tile()
    XT, YT = (X + T - 1) // T, (Y + T - 1) // T
    z = cells % Z; t = cells // Z
    y = t % Y; t = t // Y
    x = t % X; b = t // X
    return ((((b * XT + x // T) * YT + y // T) * T + x % T) * T + y % T) * Z + z


def n_keys(grid: GridSpec, B: int) -> int:
    return int(_abi.load().lss_plan_key_count(grid.c(), B))     # (cached per plan shape in _plan_sizes)


class _Workspace:
    """Scratch of the plan kernels.  Its control part (histogram, per-run counters, tile totals,
    ticket) must be zero on entry and a successful call leaves it zero again, so a cached
    workspace is zeroed once; calls on different streams are ordered through an event and a
    failed call re-zeroes it."""

    def __init__(self, dev, nbytes: int, control: int, cached: bool):
        self.buf = torch.empty(nbytes, dtype=torch.uint8, device=dev)
        self.control = control
        self.buf[:control].zero_()
        self.cached = cached
        if cached:
            self.event = torch.cuda.Event()
            self.event.record(torch.cuda.current_stream(dev))
            self.stream = _stream(dev)

    def acquire(self, dev) -> torch.Tensor:
        if self.cached:
            st = _stream(dev)
            if st != self.stream:
                torch.cuda.current_stream(dev).wait_event(self.event)
                self.stream = st
        return self.buf

    def release(self, dev) -> None:
        if self.cached:
            self.event.record(torch.cuda.current_stream(dev))

    def reset(self) -> None:
        self.buf[:self.control].zero_()


_WORKSPACES: Dict[Tuple, _Workspace] = {}


def _workspace(dev, nbytes: int, control: int) -> _Workspace:
    """One workspace per (device, size), reused across calls.  While the stream is capturing a CUDA
    graph a fresh one is taken instead (memory from the graph's pool, zeroed by a captured memset):
    events and a buffer shared with eager calls have no place inside a graph."""
    if nbytes == 0:
        raise RuntimeError("the plan workspace query rejected the shape/grid")
    if torch.cuda.is_current_stream_capturing():
        return _Workspace(dev, nbytes, control, cached=False)
    key = (dev.index, nbytes)
    ws = _WORKSPACES.get(key)
    if ws is None:
        ws = _Workspace(dev, nbytes, control, cached=True)
        _WORKSPACES[key] = ws
    return ws


def _plan_outputs(P: int, nkeys: int, dev):
    return (torch.empty(P, dtype=torch.int32, device=dev),
            torch.empty(nkeys + 1, dtype=torch.int32, device=dev),
            torch.empty((P, 2), dtype=torch.int32, device=dev),
            torch.empty(2, dtype=torch.int32, device=dev))


def build_plan(us, vs, ds, rots, trans, intrins, post_rots, post_trans, grid: GridSpec) -> Plan:
    """K0 -> K1' -> counting sort -> intervals in one C call (lss_build_plan)."""
    dev = _need_cuda(us, vs, ds, rots, trans, intrins, post_rots, post_trans)
    B, N = trans.shape[0], trans.shape[1]
    D, fH, fW = ds.numel(), vs.numel(), us.numel()
    nbytes, control, nkeys = _plan_sizes(grid, B, N, D, fH, fW)
    P = B * N * D * fH * fW
    with _device_guard(dev):
        ws = _workspace(dev, nbytes, control)
        buf = ws.acquire(dev)
        cells, key_start, sorted_rec, counts = _plan_outputs(P, nkeys, dev)
        try:
            _abi.call("lss_build_plan", _ptr(_f32c(us)), _ptr(_f32c(vs)), _ptr(_f32c(ds)),
                      _ptr(_f32c(rots)), _ptr(_f32c(trans)), _ptr(_f32c(intrins)),
                      _ptr(_f32c(post_rots)), _ptr(_f32c(post_trans)), grid.c(), _shape_c(B, N, D, fH, fW, 4),
                      _ptr(cells), _ptr(key_start), _ptr(sorted_rec), _ptr(counts), _ptr(buf), nbytes, _stream(dev))
        except _abi.LssError:
            ws.reset()  # a failed call may leave the control words dirty
            raise
        ws.release(dev)
    return Plan(grid, B, N, D, fH, fW, cells, key_start, sorted_rec, counts)


def plan_from_geom(geom: torch.Tensor, grid: GridSpec) -> Plan:
    """Plan from a dense (B,N,D,fH,fW,3) geometry tensor (the literal
    voxel_pooling(geom_feats, x) signature): lss_build_plan_from_geom."""
    dev = _need_cuda(geom)
    B, N, D, fH, fW, _ = geom.shape
    geom = _f32c(geom)
    P = B * N * D * fH * fW
    g = grid.c()
    shape = _abi.make_shape(B, N, D, fH, fW, 4)
    lib = _abi.load()
    with torch.cuda.device(dev):
        ws = _workspace(dev, lib.lss_plan_from_geom_workspace_bytes(P, g, B),
                        lib.lss_plan_workspace_control_bytes(shape, g))
        buf = ws.acquire(dev)
        cells, key_start, sorted_rec, counts = _plan_outputs(P, n_keys(grid, B), dev)
        try:
            _abi.call("lss_build_plan_from_geom", _ptr(geom), g, B, P, _ptr(cells), _ptr(key_start),
                      _ptr(sorted_rec), _ptr(counts), _ptr(buf), buf.numel(), _stream(dev))
        except _abi.LssError:
            ws.reset()
            raise
        ws.release(dev)
    return Plan(grid, B, N, D, fH, fW, cells, key_start, sorted_rec, counts)


# --------------------------------------------------------------------------
# BEV tensor layout helpers
# --------------------------------------------------------------------------
def _alloc_bev(plan: Plan, C: int, dev) -> torch.Tensor:
    """(B, X, Y, Z*C) storage; .permute(0,3,1,2) is the logical (B, C*Z, X, Y) result
    with channels_last strides."""
    X, Y, Z = plan.grid.nx
    return torch.empty((plan.B, X, Y, Z * C), dtype=torch.float32, device=dev)


# how often an upstream gradient arrived in another layout than channels_last and had to be transposed
# (82 MB at the headline config: as much traffic as the backward itself).  A consumer that runs in
# channels_last -- cuDNN does as soon as its input is, e.g. bevencode.conv1 (reference src/modules.py:99)
# fed with our output -- hands the gradient back in that layout and this stays 0 (tests assert it).
NHWC_TRANSPOSES = 0


def _as_nhwc(grad: torch.Tensor) -> torch.Tensor:
    """View/copy of a logical (B, C, X, Y) tensor as contiguous (B, X, Y, C)."""
    global NHWC_TRANSPOSES
    g = grad.permute(0, 2, 3, 1)
    if g.dtype != torch.float32:
        g = g.float()
    if not g.is_contiguous():
        NHWC_TRANSPOSES += 1
        g = g.contiguous()
    return g


# --------------------------------------------------------------------------
# fused lift + splat (K4 / K5)
# --------------------------------------------------------------------------
def feat_stage(feat: torch.Tensor, plan: Plan) -> torch.Tensor:
    """Pixel-major float32 copy (B*N*fH*fW, C) of the context features.  feat may be float32, float16
    or bfloat16 (the AMP scripts hand over half tensors) and a channel slice of a conv output."""
    dev = _need_cuda(feat)
    feat = _feature_input(feat)
    BN, C = feat.shape[0], feat.shape[1]
    feat_t = torch.empty((BN * plan.fH * plan.fW, C), dtype=torch.float32, device=dev)
    _call(dev, "lss_feat_stage", _ptr(feat), feat.stride(0), plan.shape(C), _DTYPES[feat.dtype], _ptr(feat_t),
          _stream(dev))
    return feat_t


def depth_softmax(logits: torch.Tensor, plan: Plan) -> torch.Tensor:
    """softmax over the D logit channels (reference src/modules.py:76-77) -> float32 (B*N, D, fH, fW)."""
    dev = _need_cuda(logits)
    logits = _feature_input(logits)
    BN = logits.shape[0]
    out = torch.empty((BN, plan.D, plan.fH, plan.fW), dtype=torch.float32, device=dev)
    _call(dev, "lss_depth_softmax", _ptr(logits), logits.stride(0), plan.shape(4), _DTYPES[logits.dtype], _ptr(out),
          _stream(dev))
    return out


def _fwd(dev, depth, feat_t, plan: Plan, C: int) -> torch.Tensor:
    bev = _alloc_bev(plan, C, dev)
    _call(dev, "lss_liftsplat_fwd", _ptr(depth), depth.stride(0), _DTYPES[depth.dtype], _ptr(feat_t),
          _ptr(plan.sorted_rec), _ptr(plan.key_start), plan.grid.c(), plan.shape(C), _ptr(bev), _stream(dev))
    return bev


class _LiftSplat(torch.autograd.Function):
    """Counterpart of QuickCumsum (reference src/tools.py:192-218) for the fused op:
    forward saves the depth distribution (the caller's tensor, read in place), the staged
    features and the per-point cell table, index tensors are non-differentiable, backward is
    one kernel."""

    @staticmethod
    def forward(ctx, depth, feat, plan: Plan):
        dev = _need_cuda(depth, feat)
        BN, C = feat.shape[0], feat.shape[1]
        if C % 4 != 0:
            raise RuntimeError("C must be a multiple of 4 (128-bit channel vectors), got %d" % C)
        if tuple(depth.shape) != (plan.B * plan.N, plan.D, plan.fH, plan.fW) or \
                tuple(feat.shape[2:]) != (plan.fH, plan.fW) or BN != plan.B * plan.N:
            raise RuntimeError("depth %s / feat %s do not match the plan (B=%d N=%d D=%d fH=%d fW=%d)"
                               % (tuple(depth.shape), tuple(feat.shape), plan.B, plan.N, plan.D,
                                  plan.fH, plan.fW))
        # depth and feat may differ in dtype (under autocast the softmax output is float32, the conv
        # output half: train_vovnet_transformer.py:196): each is read in its own type, no promotion copy
        ctx.in_dtypes = (depth.dtype, feat.dtype)
        depth_in = _feature_input(depth.detach())          # read in place: indexed by point id
        feat_t = feat_stage(feat.detach(), plan)
        bev = _fwd(dev, depth_in, feat_t, plan, C)
        ctx.plan = plan
        ctx.C = C
        ctx.save_for_backward(depth_in, feat_t)
        return bev.permute(0, 3, 1, 2)

    @staticmethod
    def backward(ctx, grad_bev):
        depth, feat_t = ctx.saved_tensors
        plan, C = ctx.plan, ctx.C
        dev = grad_bev.device
        g = _as_nhwc(grad_bev)
        dt_d, dt_f = ctx.in_dtypes
        out_dt = dt_d if (dt_d == dt_f and dt_d in _DTYPES) else torch.float32
        BN, HW = plan.B * plan.N, plan.fH * plan.fW
        ddepth = torch.empty((BN, plan.D, plan.fH, plan.fW), dtype=out_dt, device=dev)
        dfeat = torch.empty((BN, C, plan.fH, plan.fW), dtype=out_dt, device=dev)
        _call(dev, "lss_liftsplat_bwd", _ptr(g), _ptr(depth), depth.stride(0), _DTYPES[depth.dtype], _ptr(feat_t),
              _ptr(plan.cells), plan.grid.c(), plan.shape(C), 0, _DTYPES[out_dt], _ptr(ddepth), plan.D * HW,
              _ptr(dfeat), C * HW, _stream(dev))
        return ddepth.to(dt_d), dfeat.to(dt_f), None


def lift_splat(depth: torch.Tensor, feat: torch.Tensor, plan: Plan,
               memory_format: torch.memory_format = torch.channels_last) -> torch.Tensor:
    """BEV (B, C*Z, X, Y) float32 = splat(lift(depth, feat)).

    depth (B*N, D, fH, fW) is the per-pixel depth distribution, feat
    (B*N, C, fH, fW) the context features; the (B*N, C, D, fH, fW) product of
    reference src/modules.py:84 is never materialised.  The result has
    channels_last strides by default (each voxel's C values are one contiguous
    line); pass torch.contiguous_format for the reference's dense NCHW layout
    (one extra transpose pass)."""
    out = _LiftSplat.apply(depth, feat, plan)
    if memory_format == torch.contiguous_format:
        out = out.contiguous()
    return out


class _LiftSplatLogits(torch.autograd.Function):
    """lift + splat straight from the CamEncode conv output y = depthnet(x) (reference
    src/modules.py:82-84): channels [0, D) are the depth LOGITS, [D, D+C) the context features,
    both read as channel slices of y (no .contiguous() copies).  The softmax of src/modules.py:77
    is one small kernel (lss_depth_softmax), its backward is fused into K5 (lss_liftsplat_bwd,
    softmax=1), and the gradient comes back as ONE tensor shaped like y."""

    @staticmethod
    def forward(ctx, y, D: int, C: int, plan: Plan):
        dev = _need_cuda(y)
        if C % 4 != 0:
            raise RuntimeError("C must be a multiple of 4 (128-bit channel vectors), got %d" % C)
        BN, R, fH, fW = y.shape
        if R < D + C or (BN, D, fH, fW) != (plan.B * plan.N, plan.D, plan.fH, plan.fW):
            raise RuntimeError("conv output %s does not match D=%d C=%d and the plan (B=%d N=%d D=%d fH=%d fW=%d)"
                               % (tuple(y.shape), D, C, plan.B, plan.N, plan.D, plan.fH, plan.fW))
        ctx.in_dtype = y.dtype
        yy = _feature_input(y.detach())
        depth_p = depth_softmax(yy[:, :D], plan)
        feat_t = feat_stage(yy[:, D:D + C], plan)
        bev = _fwd(dev, depth_p, feat_t, plan, C)
        ctx.plan, ctx.D, ctx.C, ctx.R = plan, D, C, R
        ctx.save_for_backward(depth_p, feat_t)
        return bev.permute(0, 3, 1, 2)

    @staticmethod
    def backward(ctx, grad_bev):
        depth_p, feat_t = ctx.saved_tensors
        plan, D, C, R = ctx.plan, ctx.D, ctx.C, ctx.R
        dev = grad_bev.device
        g = _as_nhwc(grad_bev)
        BN, HW = plan.B * plan.N, plan.fH * plan.fW
        out_dt = ctx.in_dtype if ctx.in_dtype in _DTYPES else torch.float32
        dy = (torch.zeros if R > D + C else torch.empty)((BN, R, plan.fH, plan.fW), dtype=out_dt, device=dev)
        _call(dev, "lss_liftsplat_bwd", _ptr(g), _ptr(depth_p), D * HW, _abi.LSS_F32, _ptr(feat_t), _ptr(plan.cells),
              plan.grid.c(), plan.shape(C), 1, _DTYPES[out_dt], dy.data_ptr(), R * HW,
              dy.data_ptr() + D * HW * dy.element_size(), R * HW, _stream(dev))
        return dy.to(ctx.in_dtype), None, None, None


def lift_splat_logits(y: torch.Tensor, D: int, C: int, plan: Plan,
                      memory_format: torch.memory_format = torch.channels_last) -> torch.Tensor:
    """BEV (B, C*Z, X, Y) = splat(lift(softmax(y[:, :D], dim=1), y[:, D:D+C])) for the CamEncode conv
    output y (B*N, >= D+C, fH, fW); see _LiftSplatLogits."""
    out = _LiftSplatLogits.apply(y, D, C, plan)
    if memory_format == torch.contiguous_format:
        out = out.contiguous()
    return out


# --------------------------------------------------------------------------
# dense pooling (K4a / K5a): voxel_pooling on a materialised frustum tensor
# --------------------------------------------------------------------------
class _PoolDense(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, plan: Plan):
        dev = _need_cuda(x)
        C = x.shape[-1]
        if C % 4 != 0:
            raise RuntimeError("C must be a multiple of 4, got %d" % C)
        x2 = _f32c(x.reshape(-1, C))
        if x2.shape[0] != plan.P:
            raise RuntimeError("x has %d points, plan has %d" % (x2.shape[0], plan.P))
        bev = _alloc_bev(plan, C, dev)
        _call(dev, "lss_pool_dense_fwd", _ptr(x2), _ptr(plan.sorted_rec), _ptr(plan.key_start), plan.grid.c(),
              plan.B, C, plan.P, _ptr(bev), _stream(dev))
        ctx.plan = plan
        ctx.x_shape = tuple(x.shape)
        ctx.x_dtype = x.dtype
        return bev.permute(0, 3, 1, 2)

    @staticmethod
    def backward(ctx, grad_bev):
        plan = ctx.plan
        C = ctx.x_shape[-1]
        g = _as_nhwc(grad_bev)
        dx = torch.empty((plan.P, C), dtype=torch.float32, device=grad_bev.device)
        _call(grad_bev.device, "lss_pool_dense_bwd", _ptr(g), _ptr(plan.cells), plan.grid.c(), plan.B, C,
              plan.P, _ptr(dx), _stream(grad_bev.device))
        return dx.view(ctx.x_shape).to(ctx.x_dtype), None


def pool_dense(x: torch.Tensor, plan: Plan,
               memory_format: torch.memory_format = torch.channels_last) -> torch.Tensor:
    """voxel_pooling for a materialised (B,N,D,fH,fW,C) tensor (any strides)."""
    out = _PoolDense.apply(x, plan)
    if memory_format == torch.contiguous_format:
        out = out.contiguous()
    return out
