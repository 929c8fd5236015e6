"""Deterministic synthetic nuScenes-shaped inputs for the Lift-Splat hot path.

There is no dataset (and no network) in the build or GPU containers, so every
test and benchmark drives the path with a synthetic 6-camera rig whose
calibration/augmentation tensors have the same structure the reference's data
loader produces:

* extrinsics/intrinsics per camera as ``NuscData.get_image_data`` reads them
  (reference src/data.py:115-159): ``rots`` (3x3 cam->ego), ``trans`` (3),
  ``intrins`` (3x3 pinhole);
* augmentation as ``sample_augmentation`` + ``img_transform`` build it
  (reference src/data.py:90-113, src/tools.py:118-142): resize, crop, optional
  flip, rotation folded into a 2x2 ``post_rot`` / 2-vector ``post_tran`` that is
  then embedded in a 3x3 / 3-vector (src/data.py:146-149).

Everything is generated with numpy's RandomState so the bits do not depend on
the torch version; tensors are returned as float32 numpy arrays (use
``to_torch`` for tensors).
"""
from __future__ import annotations

import math
from dataclasses import dataclass, field
from typing import Dict, Tuple

import numpy as np

# nominal nuScenes rig: yaw of each camera's optical axis in the ego frame (deg)
_CAM_YAW_DEG = (55.0, 0.0, -55.0, 110.0, 180.0, -110.0)
_CAM_TRANS = (
    (1.5, 0.5, 1.5), (1.7, 0.0, 1.5), (1.5, -0.5, 1.5),
    (1.0, 0.5, 1.5), (0.0, 0.0, 1.5), (1.0, -0.5, 1.5),
)
# camera frame (x right, y down, z forward) -> ego frame (x fwd, y left, z up)
_CAM2EGO = np.array([[0.0, 0.0, 1.0], [-1.0, 0.0, 0.0], [0.0, -1.0, 0.0]])


@dataclass
class LSSConfig:
    """Shape/grid description of one BASELINE.json configuration."""
    name: str
    B: int
    N: int = 6
    final_dim: Tuple[int, int] = (128, 352)
    downsample: int = 16
    xbound: Tuple[float, float, float] = (-50.0, 50.0, 0.5)
    ybound: Tuple[float, float, float] = (-50.0, 50.0, 0.5)
    zbound: Tuple[float, float, float] = (-10.0, 10.0, 20.0)
    dbound: Tuple[float, float, float] = (4.0, 45.0, 1.0)
    C: int = 64
    H: int = 900
    W: int = 1600
    resize_lim: Tuple[float, float] = (0.193, 0.225)
    bot_pct_lim: Tuple[float, float] = (0.0, 0.22)
    rot_lim: Tuple[float, float] = (-5.4, 5.4)
    rand_flip: bool = True
    extra: Dict = field(default_factory=dict)

    @property
    def fH(self) -> int:
        return self.final_dim[0] // self.downsample

    @property
    def fW(self) -> int:
        return self.final_dim[1] // self.downsample

    @property
    def D(self) -> int:
        # number of elements torch.arange(*dbound) yields
        lo, hi, st = self.dbound
        return int(math.ceil((hi - lo) / st))

    @property
    def nx(self) -> Tuple[int, int, int]:
        # LongTensor of python-float quotients truncates (reference src/tools.py:175)
        return tuple(int((b[1] - b[0]) / b[2]) for b in (self.xbound, self.ybound, self.zbound))

    @property
    def P(self) -> int:
        return self.B * self.N * self.D * self.fH * self.fW

    def grid_conf(self) -> Dict:
        return {"xbound": list(self.xbound), "ybound": list(self.ybound),
                "zbound": list(self.zbound), "dbound": list(self.dbound)}

    def data_aug_conf(self) -> Dict:
        return {"resize_lim": self.resize_lim, "final_dim": self.final_dim,
                "rot_lim": self.rot_lim, "H": self.H, "W": self.W,
                "rand_flip": self.rand_flip, "bot_pct_lim": self.bot_pct_lim,
                "cams": ["CAM_FRONT_LEFT", "CAM_FRONT", "CAM_FRONT_RIGHT",
                         "CAM_BACK_LEFT", "CAM_BACK", "CAM_BACK_RIGHT"],
                "Ncams": self.N}

    def algorithmic_bytes(self) -> Dict[str, int]:
        """fwd = in + bev, bwd = bev + 2*in (fp32)."""
        nx = self.nx
        inp = self.B * self.N * self.fH * self.fW * (self.D + self.C) * 4
        bev = self.B * self.C * nx[2] * nx[0] * nx[1] * 4
        return {"in": inp, "bev": bev, "fwd": inp + bev, "bwd": bev + 2 * inp,
                "total": 2 * bev + 3 * inp}


def config(name: str, B: int | None = None) -> LSSConfig:
    """The BASELINE.json configurations."""
    if name in ("config1", "cfg1"):
        c = LSSConfig("config1", B=1)
    elif name in ("config2", "cfg2", "config3", "cfg3"):
        c = LSSConfig("config2", B=8)
    elif name in ("config4", "cfg4"):
        c = LSSConfig("config4", B=16, final_dim=(256, 704), dbound=(1.0, 60.0, 1.0), C=80,
                      resize_lim=(0.44, 0.50))
    elif name in ("config5", "cfg5"):
        c = LSSConfig("config5", B=32, final_dim=(256, 704), dbound=(1.0, 60.0, 0.5), C=128,
                      xbound=(-51.2, 51.2, 0.2), ybound=(-51.2, 51.2, 0.2),
                      resize_lim=(0.44, 0.50))
    elif name == "tiny":
        c = LSSConfig("tiny", B=2, N=3, final_dim=(64, 96), dbound=(4.0, 20.0, 1.0), C=8,
                      xbound=(-16.0, 16.0, 1.0), ybound=(-16.0, 16.0, 1.0),
                      zbound=(-10.0, 10.0, 5.0), resize_lim=(0.10, 0.12),
                      bot_pct_lim=(0.0, 0.1))
    else:
        raise KeyError(name)
    if B is not None:
        c.B = B
    return c


def _rot2(h: float) -> np.ndarray:
    return np.array([[math.cos(h), math.sin(h)], [-math.sin(h), math.cos(h)]])


def _augment(rs: np.random.RandomState, cfg: LSSConfig):
    """One draw of (post_rot 2x2, post_tran 2): resize -> crop -> flip -> rotate."""
    fH, fW = cfg.final_dim
    resize = rs.uniform(*cfg.resize_lim)
    newW, newH = int(cfg.W * resize), int(cfg.H * resize)
    crop_h = int((1 - rs.uniform(*cfg.bot_pct_lim)) * newH) - fH
    crop_w = int(rs.uniform(0, max(0, newW - fW)))
    flip = bool(cfg.rand_flip and rs.randint(0, 2))
    rotate = rs.uniform(*cfg.rot_lim)

    post_rot = np.eye(2) * resize
    post_tran = -np.array([crop_w, crop_h], dtype=np.float64)
    if flip:
        A = np.array([[-1.0, 0.0], [0.0, 1.0]])
        post_rot = A @ post_rot
        post_tran = A @ post_tran + np.array([fW, 0.0])
    A = _rot2(rotate / 180.0 * math.pi)
    b = np.array([fW, fH], dtype=np.float64) / 2
    b = A @ (-b) + b
    post_rot = A @ post_rot
    post_tran = A @ post_tran + b
    return post_rot, post_tran


def make_calibration(cfg: LSSConfig, seed: int = 1234) -> Dict[str, np.ndarray]:
    """rots (B,N,3,3), trans (B,N,3), intrins (B,N,3,3), post_rots (B,N,3,3),
    post_trans (B,N,3) as float32, one independent augmentation per camera."""
    rs = np.random.RandomState(seed)
    B, N = cfg.B, cfg.N
    rots = np.zeros((B, N, 3, 3)); trans = np.zeros((B, N, 3))
    intrins = np.zeros((B, N, 3, 3)); post_rots = np.zeros((B, N, 3, 3))
    post_trans = np.zeros((B, N, 3))
    for b in range(B):
        for n in range(N):
            yaw = math.radians(_CAM_YAW_DEG[n % 6] + rs.normal(0.0, 1.0))
            Rz = np.array([[math.cos(yaw), -math.sin(yaw), 0.0],
                           [math.sin(yaw), math.cos(yaw), 0.0], [0.0, 0.0, 1.0]])
            rots[b, n] = Rz @ _CAM2EGO
            trans[b, n] = np.array(_CAM_TRANS[n % 6]) + rs.normal(0.0, 0.01, 3)
            f = 1266.0 + rs.normal(0.0, 5.0)
            intrins[b, n] = np.array([[f, 0.0, 816.0 + rs.normal(0.0, 5.0)],
                                      [0.0, f, 491.0 + rs.normal(0.0, 5.0)],
                                      [0.0, 0.0, 1.0]])
            pr, pt = _augment(rs, cfg)
            post_rots[b, n] = np.eye(3); post_rots[b, n, :2, :2] = pr
            post_trans[b, n, :2] = pt
    f32 = np.float32
    return {"rots": rots.astype(f32), "trans": trans.astype(f32),
            "intrins": intrins.astype(f32), "post_rots": post_rots.astype(f32),
            "post_trans": post_trans.astype(f32)}


def make_features(cfg: LSSConfig, seed: int = 1234, softmax: bool = True) -> Dict[str, np.ndarray]:
    """depth (B*N,D,fH,fW) softmaxed over D, feat (B*N,C,fH,fW) ~ N(0,1) and an
    upstream gradient dbev (B, C*Z, X, Y) ~ N(0,1), all float32."""
    rs = np.random.RandomState(seed + 7919)
    BN = cfg.B * cfg.N
    logits = rs.standard_normal((BN, cfg.D, cfg.fH, cfg.fW)).astype(np.float32)
    if softmax:
        e = np.exp(logits - logits.max(axis=1, keepdims=True))
        depth = (e / e.sum(axis=1, keepdims=True)).astype(np.float32)
    else:
        depth = logits
    feat = rs.standard_normal((BN, cfg.C, cfg.fH, cfg.fW)).astype(np.float32)
    return {"depth": depth, "feat": feat}


def make_dbev(cfg: LSSConfig, seed: int = 1234) -> np.ndarray:
    rs = np.random.RandomState(seed + 104729)
    nx = cfg.nx
    return rs.standard_normal((cfg.B, cfg.C * nx[2], nx[0], nx[1])).astype(np.float32)


# --------------------------------------------------------------------------
# a cheap deterministic field for the multi-GB upstream gradients of the large configurations: its value
# at any logical index can be evaluated on its own (fixtures sample it with numpy, the GPU tests fill the
# whole tensor with torch), bit-identical on both sides: 24 hashed bits scaled to [-0.5, 0.5).
# --------------------------------------------------------------------------
_HASH_MUL = 2654435761


def hash_field_np(idx, seed: int = 1234) -> np.ndarray:
    v = ((np.asarray(idx, dtype=np.int64) * _HASH_MUL + seed * 40503) & 0xFFFFFFFF) >> 8
    return (v.astype(np.float32) * np.float32(2.0 ** -24) - np.float32(0.5)).astype(np.float32)


def hash_field_torch(idx, seed: int = 1234):
    import torch
    v = ((idx.to(torch.int64) * _HASH_MUL + seed * 40503) & 0xFFFFFFFF) >> 8
    return v.to(torch.float32) * (2.0 ** -24) - 0.5


def hash_dbev_nhwc(cfg: "LSSConfig", device, seed: int = 1234):
    """The (B, X, Y, Z*C) channels-innermost storage of the logical (B, C*Z, X, Y) gradient whose element
    [b, ch, x, y] is hash_field(((b*CZ + ch)*X + x)*Y + y)."""
    import torch
    X, Y, Z = cfg.nx
    CZ = cfg.C * Z
    out = torch.empty((cfg.B, X, Y, CZ), dtype=torch.float32, device=device)
    ch = torch.arange(CZ, device=device, dtype=torch.int64).view(1, 1, CZ) * (X * Y)
    xy = (torch.arange(X, device=device, dtype=torch.int64).view(X, 1, 1) * Y +
          torch.arange(Y, device=device, dtype=torch.int64).view(1, Y, 1))
    for b in range(cfg.B):                       # one sample at a time keeps the int64 temporaries small
        out[b] = hash_field_torch(b * CZ * X * Y + ch + xy, seed)
    return out


def to_torch(d: Dict[str, np.ndarray], device="cpu"):
    import torch
    return {k: torch.from_numpy(np.ascontiguousarray(v)).to(device) for k, v in d.items()}
