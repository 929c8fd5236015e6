"""Import the UNMODIFIED reference (/root/reference/src) in the build container.

TEST INFRASTRUCTURE ONLY.  /root/reference does not exist on the GPU box, so
nothing under tests/ (-m gpu), bench.py or __graft_entry__.smoke() may import
this module at run time.  It is used by oracle/make_golden.py (to generate the
committed fixtures under tests/golden/) and by the container-only tests that
validate the numpy/C restatement against the real reference.

The reference's top-level imports pull in packages that are not installed here
(pyquaternion, matplotlib, nuscenes, efficientnet_pytorch, timm); none of them
is touched by the hot path (get_geometry / get_cam_feats / voxel_pooling /
QuickCumsum), so they are stubbed with MagicMock.
"""
import os
import sys
from unittest import mock

REFERENCE_ROOT = os.environ.get("LSS_REFERENCE_ROOT", "/root/reference")

_STUBS = [
    "pyquaternion", "matplotlib", "matplotlib.pyplot", "matplotlib.patches",
    "nuscenes", "nuscenes.utils", "nuscenes.utils.data_classes",
    "nuscenes.utils.geometry_utils", "nuscenes.map_expansion",
    "nuscenes.map_expansion.map_api", "nuscenes.nuscenes",
    "efficientnet_pytorch", "timm",
]


def available() -> bool:
    return os.path.isdir(os.path.join(REFERENCE_ROOT, "src"))


def load():
    """Return the reference's (tools, model_baseline, modules) modules."""
    if not available():
        raise RuntimeError("reference tree not present at %s" % REFERENCE_ROOT)
    sys.dont_write_bytecode = True  # the reference tree is read-only
    for m in _STUBS:
        sys.modules.setdefault(m, mock.MagicMock())
    if REFERENCE_ROOT not in sys.path:
        sys.path.insert(0, REFERENCE_ROOT)
    from src import tools, model_baseline, modules  # noqa: E402
    return tools, model_baseline, modules


class _NoEncoder:
    """Stand-in for modules.Encoder (EfficientNet.from_pretrained needs network)."""

    def __new__(cls, *a, **k):
        import torch
        return torch.nn.Identity()


def build_lss(bsize, grid_conf, data_aug_conf, outC=4, cls="LSS"):
    """Instantiate the reference's LSS / BEV_TXT with the backbone stubbed out.

    Everything on the hot path (frustum, dx/bx/nx, CamEncode's real 1x1 conv,
    get_geometry, get_cam_feats, voxel_pooling) is the reference's own code.
    """
    tools, model_baseline, modules = load()
    with mock.patch.object(model_baseline, "Encoder", _NoEncoder):
        model = getattr(model_baseline, cls)(bsize, grid_conf, data_aug_conf, outC)
    return model
