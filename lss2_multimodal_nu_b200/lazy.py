"""LiftedFrustum: the lifted camera features, never materialised.

The reference forms new_x[bn, c, d, h, w] = depth[bn, d, h, w] * feat[bn, c, h, w]
(src/modules.py:84, src/model_vovnet_transformer.py:120), views it as
(B, N, C, D, fH, fW) and permutes to (B, N, D, fH, fW, C)
(src/model_baseline.py:79-80, src/model_vovnet_transformer.py:597-599,
pre_train_vovnet.py:150-154) only so that voxel_pooling can read it back point by
point.  This handle carries the two factors and answers exactly those shape
queries / views; the patched voxel_pooling consumes the factors directly.  Any
other use materialises the real tensor with the reference's own expression, so
the handle is always safe to pass to unpatched code.
"""
from __future__ import annotations

from typing import Optional, Tuple

import torch

# logical axes of the product, named
_AXES_5 = ("bn", "c", "d", "h", "w")            # what cam_encode returns
_AXES_6 = ("b", "n", "c", "d", "h", "w")        # after .view(B, N, C, D, H, W)
_POOL_6 = ("b", "n", "d", "h", "w", "c")        # what voxel_pooling wants
_POOL_5 = ("bn", "d", "h", "w", "c")            # pre_train_vovnet permutes first


class LiftedFrustum:
    def __init__(self, depth: torch.Tensor, feat: torch.Tensor, B: Optional[int],
                 N: Optional[int], axes: Tuple[str, ...] = _AXES_5):
        if depth.shape[0] != feat.shape[0] or depth.shape[2:] != feat.shape[2:]:
            raise RuntimeError("depth %s and feat %s disagree" % (tuple(depth.shape), tuple(feat.shape)))
        self.depth, self.feat = depth, feat
        self.B, self.N = B, N
        self.axes = axes

    # ---- tensor-like metadata ------------------------------------------------
    def _extent(self, a: str) -> int:
        BN, D, H, W = self.depth.shape
        C = self.feat.shape[1]
        return {"bn": BN, "b": self.B, "n": self.N, "c": C, "d": D, "h": H, "w": W}[a]

    @property
    def shape(self) -> torch.Size:
        return torch.Size(self._extent(a) for a in self.axes)

    def size(self, dim=None):
        return self.shape if dim is None else self.shape[dim]

    def dim(self) -> int:
        return len(self.axes)

    @property
    def device(self):
        return self.feat.device

    @property
    def dtype(self):
        return torch.result_type(self.depth, self.feat)

    @property
    def requires_grad(self) -> bool:
        return self.depth.requires_grad or self.feat.requires_grad

    # ---- the views the reference performs -------------------------------------
    def view(self, *shape):
        if len(shape) == 1 and isinstance(shape[0], (tuple, list, torch.Size)):
            shape = tuple(shape[0])
        shape = tuple(int(s) for s in shape)
        cur = tuple(self.shape)
        if len(self.axes) == 5 and self.axes[0] == "bn" and len(shape) == 6 and shape[2:] == cur[1:] \
                and shape[0] * shape[1] == cur[0]:
            rest = self.axes[1:]
            return LiftedFrustum(self.depth, self.feat, shape[0], shape[1], ("b", "n") + rest)
        if shape == cur:
            return self
        return self.materialize().view(*shape)

    reshape = view

    def permute(self, *dims):
        if len(dims) == 1 and isinstance(dims[0], (tuple, list)):
            dims = tuple(dims[0])
        if len(dims) != len(self.axes):
            raise RuntimeError("permute: expected %d dims, got %d" % (len(self.axes), len(dims)))
        return LiftedFrustum(self.depth, self.feat, self.B, self.N, tuple(self.axes[d] for d in dims))

    def to_pooling_layout(self) -> "LiftedFrustum":
        if self.B is None:
            raise RuntimeError("batch split unknown: view(B, N, ...) first")
        return LiftedFrustum(self.depth, self.feat, self.B, self.N, _POOL_6)

    def is_pooling_layout(self) -> bool:
        return self.axes == _POOL_6

    # ---- escape hatch -----------------------------------------------------------
    def materialize(self) -> torch.Tensor:
        """The real tensor, built with the reference's expression (src/modules.py:84)."""
        x = self.depth.unsqueeze(1) * self.feat.unsqueeze(2)  # (BN, C, D, H, W)
        names = list(_AXES_5)
        if "b" in self.axes:
            x = x.view(self.B, self.N, *x.shape[1:])
            names = list(_AXES_6)
        return x.permute(*[names.index(a) for a in self.axes])

    def __getattr__(self, name):
        # anything we do not model: behave like the materialised tensor
        if name.startswith("__") or name in ("depth", "feat", "B", "N", "axes"):
            raise AttributeError(name)
        return getattr(self.materialize(), name)

    def __repr__(self):
        return "LiftedFrustum(shape=%s, axes=%s)" % (tuple(self.shape), self.axes)
