"""Drop-in installation: rebind the reference models' hot-path methods.

The reference has no plugin registry; its boundary is the nn.Module method
surface.  ``install(model_or_class)`` rebinds

    get_geometry(rots, trans, intrins, post_rots, post_trans)   src/model_baseline.py:50
    get_cam_feats(x)                                            src/model_baseline.py:72
    voxel_pooling(geom_feats, x)                                src/model_baseline.py:84
    get_voxels(x, rots, trans, intrins, post_rots, post_trans)  src/model_baseline.py:128

(and ``cam_encode.forward`` for the VoVNet class, src/model_vovnet_transformer.py:100)
on LSS / BEV_TXT / VoVNetBEVTransformer / PreTrainingModel instances or classes.
Signatures, argument meaning, output shapes and values are the reference's; no
parameter or buffer is added, so ``load_state_dict(strict=True)`` keeps working.
The reference source files stay byte-identical: ``python -m
lss2_multimodal_nu_b200.run train.py ...`` patches the classes and then runs the
unmodified script.
"""
from __future__ import annotations

import types
from typing import Optional

import torch

from . import functional as F
from .lazy import LiftedFrustum

_CACHE_ATTR = "_lss_b200_cache"  # plain python attribute: not in state_dict
STATIC_BY_DEFAULT = False         # run.py sets it from LSS_STATIC_CALIB (evaluation with a fixed rig)


# --------------------------------------------------------------------------
# per-module host cache of the grid constants and frustum axes
# --------------------------------------------------------------------------
class _ModuleCache:
    def __init__(self):
        self.key = None
        self.grid: Optional[F.GridSpec] = None
        self.axes = None
        # evaluation-time plan reuse: opt-in, see static_calibration()
        self.static = False
        self.plan: Optional[F.Plan] = None
        self.plan_builds = 0
        # producer fusion: None = not probed yet
        self.fuse_softmax: Optional[bool] = None


def _cache(module) -> _ModuleCache:
    c = module.__dict__.get(_CACHE_ATTR)
    if c is None:
        c = _ModuleCache()
        c.static = STATIC_BY_DEFAULT
        object.__setattr__(module, _CACHE_ATTR, c)
    key = (module.dx.data_ptr(), module.dx._version, module.bx._version, module.nx._version,
           module.frustum.data_ptr(), module.frustum._version, str(module.frustum.device))
    if c.key != key:
        # one device->host read of 9 numbers; repeated only if the parameters move or change
        c.grid = F.GridSpec.from_tensors(module.dx, module.bx, module.nx)
        c.axes = F.frustum_axes(module.frustum)
        c.key = key
    return c


def static_calibration(module, enabled: bool = True) -> None:
    """Opt in to evaluation-time plan reuse: with a fixed camera rig and the deterministic
    validation augmentation (reference src/data.py:104-112) the calibration tensors are the same
    for every frame, so the plan (geometry, sort, intervals) is built once and reused until
    ``static_calibration(module, False)`` or ``invalidate_plan(module)``.  The caller vouches that
    the calibration does not change; shapes are still checked."""
    c = _cache(module)
    c.static = bool(enabled)
    c.plan = None


def invalidate_plan(module) -> None:
    _cache(module).plan = None


def _plan_for(c: _ModuleCache, calib) -> F.Plan:
    if c.static and c.plan is not None:
        rots = calib[0]
        if (c.plan.B, c.plan.N) == (rots.shape[0], rots.shape[1]) and c.plan.cells.device == rots.device:
            return c.plan
    us, vs, ds = c.axes
    plan = F.build_plan(us, vs, ds, *calib, c.grid)
    c.plan_builds += 1
    if c.static:
        c.plan = plan
    return plan


def _channels(module) -> int:
    return int(getattr(module, "camC", getattr(module, "C", 0)))


# --------------------------------------------------------------------------
# replacement methods
# --------------------------------------------------------------------------
def get_geometry(self, rots, trans, intrins, post_rots, post_trans):
    """(B, N, D, fH, fW, 3) ego-frame points; reference src/model_baseline.py:50-70.
    Geometry is always computed in float32 (also under autocast, where the
    reference silently drops to half precision: SURVEY.md 7.3-8)."""
    c = _cache(self)
    us, vs, ds = c.axes
    out = F.geometry(us, vs, ds, rots, trans, intrins, post_rots, post_trans, c.grid, want_geom=True)
    geom = out["geom"]
    # remember where this tensor came from so voxel_pooling can skip re-quantising it
    geom._lss_calib = (rots, trans, intrins, post_rots, post_trans)
    return geom


def _split_depth_feat(self, x):
    """Run the reference's own depthnet conv + softmax (src/modules.py:82-83) and
    return (depth, feat) WITHOUT forming their outer product (:84)."""
    ce = self.camencode
    y = ce.depthnet(x)
    depth = ce.get_depth_dist(y[:, :ce.D])
    feat = y[:, ce.D:(ce.D + ce.C)]
    return depth, feat


def _softmax_is_fusable(c: _ModuleCache, ce, like: torch.Tensor) -> bool:
    """The fused path replaces ``ce.get_depth_dist`` by its own softmax over the D logit channels;
    that is only allowed when get_depth_dist IS that softmax (reference src/modules.py:76-77).
    Probed once per module on a small random tensor."""
    if c.fuse_softmax is None:
        ok = hasattr(ce, "depthnet") and hasattr(ce, "get_depth_dist") and int(ce.D) <= 128
        if ok:
            with torch.no_grad():
                probe = torch.randn(2, int(ce.D), 2, 3, device=like.device, dtype=torch.float32)
                got = ce.get_depth_dist(probe)
                ok = tuple(got.shape) == tuple(probe.shape) and torch.allclose(got, probe.softmax(dim=1), rtol=1e-6, atol=1e-7)
        c.fuse_softmax = bool(ok)
    return c.fuse_softmax


def get_cam_feats(self, x):
    """Lazy B x N x D x fH x fW x C handle; reference src/model_baseline.py:72-82."""
    BN = x.shape[0]
    B = self.bsize
    N = BN // B
    depth, feat = _split_depth_feat(self, x)
    return LiftedFrustum(depth, feat, B, N).to_pooling_layout()


def voxel_pooling(self, geom_feats, x):
    """(B, C*Z, X, Y) BEV; reference src/model_baseline.py:84-126."""
    c = _cache(self)
    calib = getattr(geom_feats, "_lss_calib", None)
    if calib is not None:
        plan = _plan_for(c, calib)
    else:
        plan = F.plan_from_geom(geom_feats, c.grid)
    if isinstance(x, LiftedFrustum):
        if not x.is_pooling_layout():
            x = x.materialize()
        else:
            return F.lift_splat(x.depth, x.feat, plan)
    return F.pool_dense(x, plan)


def get_voxels(self, x, rots, trans, intrins, post_rots, post_trans):
    """Fused entry: geometry -> ranks -> sort -> intervals -> lift+splat; the
    geometry tensor and the frustum feature tensor are never written.
    reference src/model_baseline.py:128-133."""
    c = _cache(self)
    plan = _plan_for(c, (rots, trans, intrins, post_rots, post_trans))
    ce = self.camencode
    if _softmax_is_fusable(c, ce, x):
        # conv output -> (softmax + split + staging) -> lift+splat; softmax backward fused into K5
        y = ce.depthnet(x)
        if y.shape[0] != plan.B * plan.N:
            raise RuntimeError("get_voxels: %d camera images but calibration for %d x %d"
                               % (y.shape[0], plan.B, plan.N))
        return F.lift_splat_logits(y, int(ce.D), int(ce.C), plan)
    depth, feat = _split_depth_feat(self, x)
    if depth.shape[0] != plan.B * plan.N:
        raise RuntimeError("get_voxels: %d camera images but calibration for %d x %d"
                           % (depth.shape[0], plan.B, plan.N))
    return F.lift_splat(depth, feat, plan)


def _cam_encode_v2_forward(self, features, depth):
    """CamEncodeV2.forward (src/model_vovnet_transformer.py:100-122) returning a lazy
    (B*N, C, D, H, W) handle instead of the materialised product."""
    feat = self.feat_proj(features)
    return LiftedFrustum(depth, feat, None, None)


# --------------------------------------------------------------------------
# installation
# --------------------------------------------------------------------------
_METHODS = {"get_geometry": get_geometry, "get_cam_feats": get_cam_feats,
            "voxel_pooling": voxel_pooling, "get_voxels": get_voxels}


def install(target):
    """Patch a model instance or a model class in place and return it."""
    is_class = isinstance(target, type)
    for name, fn in _METHODS.items():
        if not hasattr(target, name):
            continue  # e.g. the VoVNet classes lift inline: no get_cam_feats / get_voxels
        if is_class:
            setattr(target, name, fn)
        else:
            object.__setattr__(target, name, types.MethodType(fn, target))
    if not is_class:
        ce = getattr(target, "cam_encode", None)
        if ce is not None and hasattr(ce, "feat_proj"):
            object.__setattr__(ce, "forward", types.MethodType(_cam_encode_v2_forward, ce))
    return target


def install_reference_classes() -> int:
    """Patch every hot-path class of the reference that is importable (``src`` on
    sys.path).  Returns how many classes were patched."""
    import importlib
    n = 0
    for mod, classes in (("src.model_baseline", ("LSS", "BEV_TXT")),
                         ("src.model_BEV_TXT", ("LSS", "BEV_TXT")),
                         ("src.model_vovnet_transformer", ("VoVNetBEVTransformer",))):
        try:
            m = importlib.import_module(mod)
        except Exception:
            continue
        for cname in classes:
            cls = getattr(m, cname, None)
            if cls is not None:
                install(cls)
                n += 1
        ce = getattr(m, "CamEncodeV2", None)
        if ce is not None:
            ce.forward = _cam_encode_v2_forward
    return n
