"""Import the UNMODIFIED reference: /root/reference in the build container, oracle/_ref/ (a byte-for-byte
staged copy, oracle/stage_ref.py, git-ignored) on the GPU box.

TEST INFRASTRUCTURE ONLY.  Used by oracle/make_golden.py (committed fixtures under tests/golden/), by the
container-only tests that validate the numpy/C restatement against the real reference, by the -m gpu tests
that run the reference's own model classes unpatched and patched on the B200, and by bench.py's
`reference_gpu` / `module_api` fields.  Nothing in the product path imports it.

The reference's top-level imports pull in packages that are not installed (pyquaternion, matplotlib,
nuscenes): none of them is touched by the model code, so they are stubbed with MagicMock (SURVEY.md
section 8c).  The two BACKBONE packages (efficientnet_pytorch, timm) are replaced by the small stand-ins in
oracle/shims/ that honour the output contract the reference's own Encoder / VoVNetV2 code expects
(SURVEY.md 7.3-10): the reference's `modules.Encoder`, `CamEncode`, `BevEncode`, ... run as written.
"""
import importlib
import importlib.util
import os
import sys
from unittest import mock

HERE = os.path.dirname(os.path.abspath(__file__))


def _find_root() -> str:
    env = os.environ.get("LSS_REFERENCE_ROOT")
    if env:
        return env
    for cand in ("/root/reference", os.path.join(HERE, "_ref")):
        if os.path.isdir(os.path.join(cand, "src")):
            return cand
    return "/root/reference"


REFERENCE_ROOT = _find_root()
SHIMS = os.path.join(HERE, "shims")

_STUBS = [
    "pyquaternion", "matplotlib", "matplotlib.pyplot", "matplotlib.patches",
    "nuscenes", "nuscenes.utils", "nuscenes.utils.data_classes",
    "nuscenes.utils.geometry_utils", "nuscenes.map_expansion",
    "nuscenes.map_expansion.map_api", "nuscenes.nuscenes", "wandb",
]


def available() -> bool:
    return os.path.isdir(os.path.join(REFERENCE_ROOT, "src"))


def _missing(name: str) -> bool:
    """True when the top-level package of ``name`` is neither imported nor importable."""
    top = name.split(".")[0]
    if top in sys.modules:
        return isinstance(sys.modules[top], mock.MagicMock)
    try:
        return importlib.util.find_spec(top) is None
    except (ValueError, ImportError):
        return True


def _prepare():
    if not available():
        raise RuntimeError("reference tree not present at %s (run oracle/stage_ref.py in the build container)"
                           % REFERENCE_ROOT)
    sys.dont_write_bytecode = True  # the reference tree is read-only
    for m in _STUBS:
        if m not in sys.modules and _missing(m):
            sys.modules[m] = mock.MagicMock()
    for m in ("efficientnet_pytorch", "timm"):           # the stand-in backbones, unless the real package exists
        if m not in sys.modules and _missing(m) and SHIMS not in sys.path:
            sys.path.insert(0, SHIMS)
    if REFERENCE_ROOT not in sys.path:
        sys.path.insert(0, REFERENCE_ROOT)


def load():
    """Return the reference's (tools, model_baseline, modules) modules."""
    _prepare()
    from src import tools, model_baseline, modules  # noqa: E402
    return tools, model_baseline, modules


def load_module(name: str):
    """Any module of the reference's `src` package, e.g. 'model_BEV_TXT', 'model_vovnet_transformer'."""
    _prepare()
    return importlib.import_module("src." + name)


def load_script(name: str):
    """A top-level script of the reference (e.g. 'pre_train_vovnet') imported as a module (its
    `if __name__ == '__main__'` block does not run)."""
    _prepare()
    path = os.path.join(REFERENCE_ROOT, name + ".py")
    spec = importlib.util.spec_from_file_location("_lss_ref_script_" + name, path)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


class _NoEncoder:
    """Stand-in for modules.Encoder when only the hot-path methods are exercised."""

    def __new__(cls, *a, **k):
        import torch
        return torch.nn.Identity()


def build_lss(bsize, grid_conf, data_aug_conf, outC=4, cls="LSS", module="model_baseline", backbone=False):
    """Instantiate the reference's LSS / BEV_TXT (src/model_baseline.py or src/model_BEV_TXT.py).

    backbone=False: the image encoder is replaced by Identity (hot-path methods only: frustum, dx/bx/nx,
    CamEncode's real 1x1 conv, get_geometry, get_cam_feats, voxel_pooling are the reference's own code).
    backbone=True: the reference's own Encoder runs on the stand-in EfficientNet of oracle/shims/, so
    model(imgs, rots, trans, intrins, post_rots, post_trans) works end to end."""
    _prepare()
    mod = load_module(module)
    if backbone:
        return getattr(mod, cls)(bsize, grid_conf, data_aug_conf, outC)
    with mock.patch.object(mod, "Encoder", _NoEncoder):
        return getattr(mod, cls)(bsize, grid_conf, data_aug_conf, outC)


def build_vovnet(bsize, grid_conf, data_aug_conf, outC=4, **kw):
    """The reference's VoVNetBEVTransformer (src/model_vovnet_transformer.py:354) on the stand-in timm."""
    mod = load_module("model_vovnet_transformer")
    kw.setdefault("pretrained", False)
    kw.setdefault("vovnet_type", "vovnet39")
    return mod.VoVNetBEVTransformer(bsize, grid_conf, data_aug_conf, outC=outC, **kw)
