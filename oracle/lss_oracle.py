"""CPU oracle for the Lift-Splat (camera -> BEV) hot path.  TEST INFRASTRUCTURE.

This is a numpy restatement of the reference's algorithm, written from the
behaviour of the reference functions cited below.  It exists only to check the
CUDA path: only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
--impl reference legs may import it.  The product package never does.

Parity status: PINNED.  The reference ships no golden vectors of its own
(SURVEY.md section 4), so the pins are outputs of the unmodified reference
executed in the build container (oracle/make_golden.py -> tests/golden/*.npz)
and tests/test_oracle_vs_reference.py, which runs the reference side by side
whenever /root/reference is present.

Reference functions restated here (paths relative to /root/reference):
  gen_dx_bx            src/tools.py:172-178
  create_frustum       src/model_baseline.py:37-48
  get_geometry         src/model_baseline.py:50-70
  lift (outer product) src/modules.py:79-86, src/model_vovnet_transformer.py:100-122
  get_cam_feats        src/model_baseline.py:72-82
  voxel_pooling        src/model_baseline.py:84-126
  QuickCumsum          src/tools.py:192-218 (forward) / :211-218 (backward)

All index arithmetic is done exactly as torch-on-CPU does it for the reference
(float32 IEEE ops, no FMA contraction in the per-point transforms, truncation
toward zero in ``.long()``); the two 3x3 inverses follow the LU sequence that
torch.inverse executes on this CPU build (MKL sgetrf/sgetrs, probed bit-exact
on 10^5 matrices, see ``inverse3x3``).
"""
from __future__ import annotations

import numpy as np

f32 = np.float32


# --------------------------------------------------------------------------
# helpers: exactly-rounded float32 fused multiply-add
# --------------------------------------------------------------------------
def _fma32(a, b, c):
    """round_to_f32(a*b + c) for float32 arrays.

    a*b is exact in float64 (24+24 <= 53 bits); the float64 add rounds once to 53
    bits and the cast once more to 24.  The double rounding can only differ from
    a true fmaf when the 53-bit result lands exactly on a float32 tie, which for
    these inputs has probability ~2^-29 per operation; the C oracle
    (oracle/lss_oracle.c) uses fmaf() and the two are cross-checked.
    """
    return (np.asarray(a, np.float64) * np.asarray(b, np.float64)
            + np.asarray(c, np.float64)).astype(f32)


# --------------------------------------------------------------------------
# grid constants and frustum
# --------------------------------------------------------------------------
def gen_dx_bx(xbound, ybound, zbound):
    """reference src/tools.py:172-178.  dx = step, bx = lo + step/2 (python
    float math, then rounded to float32 by torch.Tensor), nx = truncated
    (hi - lo) / step as int64 (torch.LongTensor of python floats truncates)."""
    rows = [xbound, ybound, zbound]
    dx = np.array([r[2] for r in rows], dtype=np.float64).astype(f32)
    bx = np.array([r[0] + r[2] / 2.0 for r in rows], dtype=np.float64).astype(f32)
    nx = np.array([int((r[1] - r[0]) / r[2]) for r in rows], dtype=np.int64)
    return dx, bx, nx


def _linspace_f32(start: float, end: float, steps: int) -> np.ndarray:
    """torch.linspace(start, end, steps, dtype=float32) on CPU: float32 step;
    element i is fma(step, i, start) in the first half and fma(-step,
    steps-1-i, end) in the second (the compiler contracts ATen's
    ``start + step*i`` / ``end - step*(steps-i-1)``; probed bit-exact for
    steps 2..299)."""
    if steps == 1:
        return np.array([start], dtype=f32)
    start32, end32 = f32(start), f32(end)
    step = f32(f32(end32 - start32) / f32(steps - 1))
    i = np.arange(steps)
    up = _fma32(step, i.astype(f32), start32)
    down = _fma32(-step, (steps - 1 - i).astype(f32), end32)
    return np.where(i < steps // 2, up, down).astype(f32)


def _arange_f32(lo: float, hi: float, st: float) -> np.ndarray:
    """torch.arange(lo, hi, st, dtype=float32): size = ceil((hi-lo)/st) in
    double, value_i = lo + i*st evaluated in double then rounded."""
    n = int(np.ceil((float(hi) - float(lo)) / float(st)))
    return (float(lo) + np.arange(n, dtype=np.float64) * float(st)).astype(f32)


def frustum_axes(final_dim, downsample, dbound):
    """The three 1-D tables the frustum is the outer product of
    (reference src/model_baseline.py:39-44): us (fW,), vs (fH,), ds (D,)."""
    ogfH, ogfW = final_dim
    fH, fW = ogfH // downsample, ogfW // downsample
    ds = _arange_f32(*dbound)
    us = _linspace_f32(0, ogfW - 1, fW)
    vs = _linspace_f32(0, ogfH - 1, fH)
    return us, vs, ds


def create_frustum(final_dim, downsample, dbound):
    """(D, fH, fW, 3) float32 table of (u, v, d); reference
    src/model_baseline.py:37-48."""
    us, vs, ds = frustum_axes(final_dim, downsample, dbound)
    D, fH, fW = len(ds), len(vs), len(us)
    fr = np.empty((D, fH, fW, 3), dtype=f32)
    fr[..., 0] = us[None, None, :]
    fr[..., 1] = vs[None, :, None]
    fr[..., 2] = ds[:, None, None]
    return fr


# --------------------------------------------------------------------------
# 3x3 inverse / camera preparation
# --------------------------------------------------------------------------
def inverse3x3(A: np.ndarray) -> np.ndarray:
    """float32 inverse of a batch of 3x3 matrices, op for op what
    ``torch.inverse`` does on the CPU build the fixtures were made with
    (reference src/model_baseline.py:60,66 call torch.inverse; ATen routes it to
    LAPACK getrf + getrs on a column-major copy of A with an identity RHS).

    Sequence (probed bit-exact against torch 2.11 / MKL 2024.2, AVX-512):
      LU, partial pivoting (first maximum of |column|):
        l10 = a10 * (1/a00); l20 = a20 * (1/a00)
        a11 -= l10*a01 (fma); a21 -= l20*a01 (fma); pivot among rows 1,2
        l21 = a21 / a11  (true division)
        u12 = fma(-l10, u02, a12); u22 = fma(-l21, u12, fma(-l20, u02, a22))
      solve L U X = P^T I column by column:
        y1 = b1 - l10*y0;  y2 = (b2 - l20*y0) - l21*y1      (separate mul, sub)
        x2 = y2 * (1/u22)  [columns 0,1]   or  y2 / u22      [column 2]
        x1 = fma(-u12, x2, y1) scaled the same way by u11
        x0 = (y0 - fma(u02, x2, u01*x1)) scaled the same way by u00
    """
    A = np.asarray(A, dtype=f32)
    batch = A.shape[:-2]
    M = A.reshape(-1, 3, 3).copy()
    n = M.shape[0]
    idx = np.arange(n)
    one = f32(1.0)
    with np.errstate(all="ignore"):
        B = np.broadcast_to(np.eye(3, dtype=f32), (n, 3, 3)).copy()

        def swap_rows(T, r, p):
            # swap row r with row p[i] for each batch element
            a = T[idx, r, :].copy()
            b = T[idx, p, :].copy()
            T[idx, r, :] = b
            T[idx, p, :] = a

        # column 0
        p0 = np.argmax(np.abs(M[:, :, 0]), axis=1)
        swap_rows(M, 0, p0); swap_rows(B, 0, p0)
        r0 = (one / M[:, 0, 0]).astype(f32)
        M[:, 1, 0] = M[:, 1, 0] * r0
        M[:, 2, 0] = M[:, 2, 0] * r0
        M[:, 1, 1] = _fma32(-M[:, 1, 0], M[:, 0, 1], M[:, 1, 1])
        M[:, 2, 1] = _fma32(-M[:, 2, 0], M[:, 0, 1], M[:, 2, 1])
        # column 1
        p1 = 1 + np.argmax(np.abs(M[:, 1:, 1]), axis=1)
        swap_rows(M, 1, p1); swap_rows(B, 1, p1)
        M[:, 2, 1] = (M[:, 2, 1] / M[:, 1, 1]).astype(f32)
        M[:, 1, 2] = _fma32(-M[:, 1, 0], M[:, 0, 2], M[:, 1, 2])
        t = _fma32(-M[:, 2, 0], M[:, 0, 2], M[:, 2, 2])
        M[:, 2, 2] = _fma32(-M[:, 2, 1], M[:, 1, 2], t)

        l10, l20, l21 = M[:, 1, 0], M[:, 2, 0], M[:, 2, 1]
        u00, u01, u02 = M[:, 0, 0], M[:, 0, 1], M[:, 0, 2]
        u11, u12, u22 = M[:, 1, 1], M[:, 1, 2], M[:, 2, 2]
        rd0, rd1, rd2 = (one / u00).astype(f32), (one / u11).astype(f32), (one / u22).astype(f32)
        X = np.empty_like(M)
        for c in range(3):
            y0 = B[:, 0, c]
            y1 = (B[:, 1, c] - (l10 * y0).astype(f32)).astype(f32)
            y2 = ((B[:, 2, c] - (l20 * y0).astype(f32)).astype(f32)
                  - (l21 * y1).astype(f32)).astype(f32)
            if c < 2:
                x2 = (y2 * rd2).astype(f32)
                x1 = (_fma32(-u12, x2, y1) * rd1).astype(f32)
                s = _fma32(u02, x2, (u01 * x1).astype(f32))
                x0 = ((y0 - s).astype(f32) * rd0).astype(f32)
            else:
                x2 = (y2 / u22).astype(f32)
                x1 = (_fma32(-u12, x2, y1) / u11).astype(f32)
                s = _fma32(u02, x2, (u01 * x1).astype(f32))
                x0 = ((y0 - s).astype(f32) / u00).astype(f32)
            X[:, 0, c], X[:, 1, c], X[:, 2, c] = x0, x1, x2
    return X.reshape(batch + (3, 3))


def matmul3x3(A: np.ndarray, Bm: np.ndarray) -> np.ndarray:
    """float32 3x3 @ 3x3 as torch.matmul evaluates it on CPU for these sizes:
    out[i,j] = (a[i,0]*b[0,j] + a[i,1]*b[1,j]) + a[i,2]*b[2,j], no FMA."""
    A = np.asarray(A, f32); Bm = np.asarray(Bm, f32)
    out = np.empty(np.broadcast_shapes(A.shape, Bm.shape), dtype=f32)
    for i in range(3):
        for j in range(3):
            t = ((A[..., i, 0] * Bm[..., 0, j]).astype(f32)
                 + (A[..., i, 1] * Bm[..., 1, j]).astype(f32)).astype(f32)
            out[..., i, j] = (t + (A[..., i, 2] * Bm[..., 2, j]).astype(f32)).astype(f32)
    return out


def camera_prep(rots, intrins, post_rots):
    """inv_post_rots = inverse(post_rots); combine = rots @ inverse(intrins)
    (reference src/model_baseline.py:60,66)."""
    return inverse3x3(post_rots), matmul3x3(rots, inverse3x3(intrins))


# --------------------------------------------------------------------------
# geometry
# --------------------------------------------------------------------------
def _matvec(M, p):
    """(B,N,3,3) x (B,N,D,H,W,3): out_i = (m_i0*p0 + m_i1*p1) + m_i2*p2, each
    product and sum rounded to float32 separately (no FMA) -- the order torch's
    CPU broadcast matmul uses for 3x3 @ 3x1."""
    Mb = M[:, :, None, None, None]
    out = np.empty(p.shape, dtype=f32)
    for i in range(3):
        t = ((Mb[..., i, 0] * p[..., 0]).astype(f32)
             + (Mb[..., i, 1] * p[..., 1]).astype(f32)).astype(f32)
        out[..., i] = (t + (Mb[..., i, 2] * p[..., 2]).astype(f32)).astype(f32)
    return out


def get_geometry(frustum, rots, trans, intrins, post_rots, post_trans,
                 inv_post_rots=None, combine=None):
    """Ego-frame xyz of every frustum point, (B,N,D,fH,fW,3) float32.
    reference src/model_baseline.py:50-70:
        p = frustum - post_trans                     (:59)
        p = inv(post_rots) @ p                       (:60)
        p = (p.x*p.z, p.y*p.z, p.z)                  (:63-65)
        p = (rots @ inv(intrins)) @ p                (:66-67)
        p += trans                                   (:68)
    ``inv_post_rots`` / ``combine`` may be supplied (e.g. the reference's own
    tensors) to pin the per-point arithmetic independently of the inverse."""
    frustum = np.asarray(frustum, f32)
    rots, trans, intrins = (np.asarray(a, f32) for a in (rots, trans, intrins))
    post_rots, post_trans = np.asarray(post_rots, f32), np.asarray(post_trans, f32)
    if inv_post_rots is None or combine is None:
        ipr, comb = camera_prep(rots, intrins, post_rots)
        inv_post_rots = ipr if inv_post_rots is None else inv_post_rots
        combine = comb if combine is None else combine
    with np.errstate(all="ignore"):
        p = (frustum[None, None] - post_trans[:, :, None, None, None, :]).astype(f32)
        p = _matvec(np.asarray(inv_post_rots, f32), p)
        q = np.empty_like(p)
        q[..., 0] = (p[..., 0] * p[..., 2]).astype(f32)
        q[..., 1] = (p[..., 1] * p[..., 2]).astype(f32)
        q[..., 2] = p[..., 2]
        r = _matvec(np.asarray(combine, f32), q)
        r = (r + trans[:, :, None, None, None, :]).astype(f32)
    return r


# --------------------------------------------------------------------------
# quantise -> kept -> rank -> sort -> intervals
# --------------------------------------------------------------------------
_I64_MIN = np.iinfo(np.int64).min


def quantize(geom, dx, bx):
    """coords = ((geom - (bx - dx/2)) / dx).long()   (reference
    src/model_baseline.py:92).  float32 sub, true float32 divide, conversion
    truncates toward zero.  Non-finite / out-of-int64-range values convert to
    INT64_MIN (what the x86 cvttss2si instruction torch uses returns); they are
    never 'kept' either way."""
    geom = np.asarray(geom, f32)
    off = (np.asarray(bx, f32) - (np.asarray(dx, f32) / f32(2.0)).astype(f32)).astype(f32)
    with np.errstate(all="ignore"):
        q = ((geom - off).astype(f32) / np.asarray(dx, f32)).astype(f32)
        bad = ~np.isfinite(q) | (np.abs(q) >= f32(2.0 ** 63))
        coords = np.where(bad, 0, q).astype(np.int64)  # astype truncates toward zero
    coords[bad] = _I64_MIN
    return coords.reshape(-1, 3)


def kept_mask(coords, nx):
    """0 <= ix < nx on all three axes (reference src/model_baseline.py:99-101)."""
    nx = np.asarray(nx, np.int64)
    return ((coords[:, 0] >= 0) & (coords[:, 0] < nx[0])
            & (coords[:, 1] >= 0) & (coords[:, 1] < nx[1])
            & (coords[:, 2] >= 0) & (coords[:, 2] < nx[2]))


def batch_index(P, B):
    """sample id of each flattened point, sample-major (reference
    src/model_baseline.py:94-95; model_vovnet_transformer.py:525-526)."""
    return np.repeat(np.arange(B, dtype=np.int64), P // B)


def ranks_of(coords, batch_ix, nx, B):
    """rank = x*(ny*nz*B) + y*(nz*B) + z*B + b, int64 (reference
    src/model_baseline.py:106-109)."""
    nx = np.asarray(nx, np.int64)
    return (coords[:, 0] * (nx[1] * nx[2] * B) + coords[:, 1] * (nx[2] * B)
            + coords[:, 2] * B + batch_ix)


def argsort_ranks(ranks):
    """ranks.argsort() (reference src/model_baseline.py:110).  torch's CPU sort
    is stable for this input: ties keep ascending compacted
    index."""
    return np.argsort(ranks, kind="stable")


def interval_last_mask(sorted_ranks):
    """QuickCumsum's ``kept``: True at the LAST element of every run of equal
    ranks (reference src/tools.py:196-197)."""
    K = len(sorted_ranks)
    m = np.ones(K, dtype=bool)
    if K > 1:
        m[:-1] = sorted_ranks[1:] != sorted_ranks[:-1]
    return m


def index_pipeline(geom, dx, bx, nx, B):
    """Everything the reference computes between ``geom`` and the cumsum:
    coords (P,3) int64, kept (P,), compacted ranks (K,), sorts (K,), the
    run-last mask over the sorted order (K,), interval starts / lengths (V,)."""
    coords = quantize(geom, dx, bx)
    P = coords.shape[0]
    kept = kept_mask(coords, nx)
    bix = batch_index(P, B)
    kept_idx = np.nonzero(kept)[0]
    ranks = ranks_of(coords[kept], bix[kept], nx, B)
    sorts = argsort_ranks(ranks)
    sranks = ranks[sorts]
    last = interval_last_mask(sranks)
    ends = np.nonzero(last)[0] + 1
    starts = np.concatenate(([0], ends[:-1])) if len(ends) else np.zeros(0, np.int64)
    return {"coords": coords, "kept": kept, "kept_idx": kept_idx, "ranks": ranks,
            "sorts": sorts, "sorted_ranks": sranks, "last_mask": last,
            "interval_start": starts.astype(np.int64),
            "interval_len": (ends - starts).astype(np.int64),
            "sorted_point": kept_idx[sorts]}


# --------------------------------------------------------------------------
# lift and splat
# --------------------------------------------------------------------------
def lift(depth, feat):
    """x[bn, d, h, w, c] = depth[bn, d, h, w] * feat[bn, c, h, w] -- the outer
    product of reference src/modules.py:84 followed by the view/permute of
    get_cam_feats (src/model_baseline.py:79-80), returned already in
    (B*N, D, fH, fW, C) order.  float32 product (or float64 if inputs are)."""
    return depth[:, :, :, :, None] * np.transpose(feat, (0, 2, 3, 1))[:, None]


def voxel_pooling(geom, x, dx, bx, nx, B, mode="exact"):
    """Splat: sum the lifted features of all points that fall in the same
    (b, x, y, z) voxel; returns (B, C*Z, X, Y).  reference
    src/model_baseline.py:84-126 + QuickCumsum src/tools.py:194-208.

    mode="exact"     per-voxel sums accumulated in float64 (the value oracle:
                     equals the reference run under default dtype float64 up to
                     float64 rounding).
    mode="reference" float32 emulation of the reference's cumsum trick: global
                     prefix sum over the sorted points (float64 accumulator
                     rounded to float32 per row, as ATen's CPU cumsum does),
                     then adjacent differences at run ends in float32.
    """
    nx = np.asarray(nx, np.int64)
    C = x.shape[-1]
    ip = index_pipeline(geom, dx, bx, nx, B)
    xs = x.reshape(-1, C)[ip["sorted_point"]]
    Z, X, Y = int(nx[2]), int(nx[0]), int(nx[1])
    V = len(ip["interval_start"])
    if mode == "exact":
        out_dtype = np.float64
        if V:
            pooled = np.add.reduceat(xs.astype(np.float64), ip["interval_start"], axis=0)
        else:
            pooled = np.zeros((0, C), np.float64)
    elif mode == "reference":
        out_dtype = f32
        cs = np.cumsum(xs.astype(np.float64), axis=0).astype(f32)
        cs = cs[ip["last_mask"]]
        pooled = np.concatenate((cs[:1], (cs[1:] - cs[:-1]).astype(f32))) if V else cs
    else:
        raise ValueError(mode)
    final = np.zeros((B, C, Z, X, Y), dtype=out_dtype)
    if V:
        r = ip["sorted_ranks"][ip["interval_start"]]
        b = r % B
        z = (r // B) % Z
        y = (r // (B * Z)) % Y
        xx = r // (B * Z * Y)
        final[b, :, z, xx, y] = pooled
    # collapse Z: cat(final.unbind(2), 1)  -> channel index = z*C + c
    return np.concatenate([final[:, :, zi] for zi in range(Z)], axis=1), ip


def voxel_pooling_backward(dbev, ip, depth, feat, nx, B):
    """Gradients of the fused lift+splat w.r.t. depth and feat (float64).
    QuickCumsum.backward is an exact gather of the voxel gradient to every kept
    point (reference src/tools.py:211-218); the lift's product rule then gives
    d_depth[bn,d,h,w] = sum_c g*feat and d_feat[bn,c,h,w] = sum_d g*depth."""
    nx = np.asarray(nx, np.int64)
    BN, D, fH, fW = depth.shape
    C = feat.shape[1]
    Z, X, Y = int(nx[2]), int(nx[0]), int(nx[1])
    g5 = np.asarray(dbev, np.float64).reshape(B, Z, C, X, Y)
    kept_idx = ip["kept_idx"]
    r = ip["ranks"]
    b = r % B; z = (r // B) % Z; y = (r // (B * Z)) % Y; xx = r // (B * Z * Y)
    g = g5[b, z, :, xx, y]                               # (K, C)
    gpt = np.zeros((BN * D * fH * fW, C), np.float64)
    gpt[kept_idx] = g
    gpt = gpt.reshape(BN, D, fH, fW, C)
    featT = np.transpose(np.asarray(feat, np.float64), (0, 2, 3, 1))  # BN,H,W,C
    d_depth = np.einsum("ndhwc,nhwc->ndhw", gpt, featT)
    d_feat = np.einsum("ndhwc,ndhw->nchw", gpt, np.asarray(depth, np.float64))
    return d_depth, d_feat
