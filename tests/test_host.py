"""CPU tests of the host-side logic: C-ABI surface, lazy handle, installer, synthetic inputs,
loud failure without CUDA.  No kernel is launched here."""
import ctypes
import os
import re
import subprocess
import sys

import numpy as np
import pytest
import torch

from lss2_multimodal_nu_b200 import _abi, functional as F, patch, synthetic as S
from lss2_multimodal_nu_b200.lazy import LiftedFrustum

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_exports_every_declared_symbol():
    """Every function include/lss_b200.h declares is exported by the built .so and bound in _abi."""
    hdr = open(os.path.join(ROOT, "include", "lss_b200.h")).read()
    declared = set(re.findall(r"\b(lss_[a-z0-9_]+)\s*\(", hdr))
    assert declared, "no declarations found"
    assert declared == set(_abi.SIGNATURES), declared ^ set(_abi.SIGNATURES)
    lib = ctypes.CDLL(_abi.LIB_PATH)
    for name in declared:
        assert hasattr(lib, name), name
    loaded = _abi.load()
    assert loaded.lss_abi_version() == _abi.ABI_VERSION == 3
    assert loaded.lss_status_string(-4) == b"workspace too small"
    out = subprocess.run(["nm", "-D", "--defined-only", _abi.LIB_PATH], capture_output=True, text=True).stdout
    exported = {l.split()[-1] for l in out.splitlines() if " T " in l and "lss_" in l}
    assert declared <= exported


def test_host_side_argument_checks():
    lib = _abi.load()
    # null pointers / bad dimensions are rejected before any CUDA call
    assert lib.lss_camera_prep(None, None, None, 1, None, None, None) == -1
    assert lib.lss_sort_ranks(None, 16, 10, None, None, None, 0, None) == -1
    assert lib.lss_sort_workspace_bytes(0, 10) == 0
    n = lib.lss_sort_workspace_bytes(346368, 320000)
    assert n >= 2 * 346368 * 4
    shape = _abi.make_shape(8, 6, 41, 8, 22, 64)
    grid = F.GridSpec.from_bounds([-50, 50, 0.5], [-50, 50, 0.5], [-10, 10, 20]).c()
    assert lib.lss_plan_workspace_bytes(shape, grid) >= (320000 + 346368) * 4   # cell histogram + slot buffer
    bad = _abi.make_shape(8, 6, 0, 8, 22, 64)
    assert lib.lss_plan_workspace_bytes(bad, grid) == 0


def test_no_cpu_fallback():
    g = F.GridSpec.from_bounds([-50, 50, 0.5], [-50, 50, 0.5], [-10, 10, 20])
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        F.quantize_rank(torch.zeros(4, 3), g, 1)
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        F.camera_prep(torch.eye(3)[None], torch.eye(3)[None], torch.eye(3)[None])


def test_product_never_imports_the_oracle():
    """The package may cite oracle files in comments, but never import, include, load or run them."""
    pkg = os.path.join(ROOT, "lss2_multimodal_nu_b200")
    bad = re.compile(r"^\s*(import|from)\s+(lss_oracle|c_oracle|ref_import)|#include.*oracle|liblss_oracle|"
                     r"sys\.path.*oracle|oracle/_build|oracle/_ref", re.M)
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                txt = open(os.path.join(dirpath, f)).read()
                assert not bad.search(txt), f


def test_gridspec_matches_reference_arithmetic():
    g = F.GridSpec.from_bounds([-50.0, 50.0, 0.5], [-50.0, 50.0, 0.5], [-10.0, 10.0, 20.0])
    assert g.dx == (0.5, 0.5, 20.0) and g.bx == (-49.75, -49.75, 0.0) and g.nx == (200, 200, 1)
    g5 = F.GridSpec.from_bounds([-51.2, 51.2, 0.2], [-51.2, 51.2, 0.2], [-10.0, 10.0, 20.0])
    assert g5.nx == (512, 512, 1) and g5.n_cells(32) == 512 * 512 * 32
    assert g5.dx[0] == float(np.float32(0.2)) and g5.bx[0] == float(np.float32(-51.2 + 0.1))


def test_lazy_handle_views_and_materialisation():
    torch.manual_seed(0)
    B, N, D, H, W, C = 2, 3, 5, 4, 6, 8
    depth = torch.rand(B * N, D, H, W).softmax(1)
    feat = torch.randn(B * N, C, H, W)
    ref = depth.unsqueeze(1) * feat.unsqueeze(2)                      # reference src/modules.py:84
    h = LiftedFrustum(depth, feat, None, None)
    assert tuple(h.shape) == (B * N, C, D, H, W) and h.dim() == 5
    # VoVNetBEVTransformer.forward: view(B,N,C,D,H,W) then permute(0,1,3,4,5,2)
    v = h.view(B, N, C, D, H, W)
    assert tuple(v.shape) == (B, N, C, D, H, W)
    p = v.permute(0, 1, 3, 4, 5, 2)
    assert p.is_pooling_layout() and tuple(p.shape) == (B, N, D, H, W, C)
    assert torch.equal(p.materialize(), ref.view(B, N, C, D, H, W).permute(0, 1, 3, 4, 5, 2))
    # PreTrainingModel.forward: permute(0,2,3,4,1) then view(B,N,D,H,W,C)
    q = h.permute(0, 2, 3, 4, 1)
    assert tuple(q.shape) == (B * N, D, H, W, C)
    q6 = q.view(B, N, D, H, W, C)
    assert q6.is_pooling_layout()
    assert torch.equal(q6.materialize(), ref.permute(0, 2, 3, 4, 1).reshape(B, N, D, H, W, C))
    # anything else falls back to the real tensor
    assert torch.equal(h.sum(dim=2), ref.sum(dim=2))
    assert h.view(B * N, C, -1).shape == (B * N, C, D * H * W)


class _StandIn(torch.nn.Module):
    """Same attribute surface as the reference's LSS class (src/model_baseline.py:11-48)."""

    def __init__(self):
        super().__init__()
        self.dx = torch.nn.Parameter(torch.tensor([0.5, 0.5, 20.0]), requires_grad=False)
        self.bx = torch.nn.Parameter(torch.tensor([-49.75, -49.75, 0.0]), requires_grad=False)
        self.nx = torch.nn.Parameter(torch.tensor([200, 200, 1]), requires_grad=False)
        self.frustum = torch.nn.Parameter(torch.zeros(4, 2, 3, 3), requires_grad=False)
        self.bsize = 1

    def get_geometry(self, *a): return "ref"
    def get_cam_feats(self, x): return "ref"
    def voxel_pooling(self, g, x): return "ref"
    def get_voxels(self, *a): return "ref"


def test_install_rebinds_methods_without_touching_state_dict():
    m = _StandIn()
    keys = list(m.state_dict().keys())
    patch.install(m)
    assert list(m.state_dict().keys()) == keys                  # strict=True checkpoints still load
    for name in ("get_geometry", "get_cam_feats", "voxel_pooling", "get_voxels"):
        assert getattr(m, name).__func__ is getattr(patch, name)
    c = patch._cache(m)
    assert c.grid.nx == (200, 200, 1) and c.grid.dx == (0.5, 0.5, 20.0)
    assert patch._cache(m) is c and list(m.state_dict().keys()) == keys
    # class-level install
    class K(_StandIn):
        pass
    patch.install(K)
    assert K.get_voxels is patch.get_voxels and _StandIn.get_voxels is not patch.get_voxels


@pytest.mark.reference
def test_install_on_the_real_reference_classes():
    import ref_import
    tools, model_baseline, modules = ref_import.load()
    cfg = S.config("tiny")
    m = ref_import.build_lss(cfg.B, cfg.grid_conf(), cfg.data_aug_conf())
    before = list(m.state_dict().keys())
    patch.install(m)
    assert list(m.state_dict().keys()) == before
    assert m.get_voxels.__func__ is patch.get_voxels and m.voxel_pooling.__func__ is patch.voxel_pooling
    assert patch._cache(m).grid.nx == tuple(int(v) for v in m.nx)
    us, vs, ds = patch._cache(m).axes
    fr = m.frustum.detach()
    assert torch.equal(us, fr[0, 0, :, 0]) and torch.equal(vs, fr[0, :, 0, 1]) and torch.equal(ds, fr[:, 0, 0, 2])
    with pytest.raises(RuntimeError, match="no CPU fallback"):          # CPU tensors are refused
        m.get_geometry(*(torch.from_numpy(v) for v in S.make_calibration(cfg).values()))


def test_synthetic_inputs_are_deterministic_and_shaped():
    cfg = S.config("config2")
    a = S.make_calibration(cfg, 5); b = S.make_calibration(cfg, 5)
    assert all(np.array_equal(a[k], b[k]) for k in a)
    assert a["rots"].shape == (8, 6, 3, 3) and a["post_trans"].shape == (8, 6, 3)
    assert (a["post_rots"][..., 2, 2] == 1).all() and (a["post_trans"][..., 2] == 0).all()
    f = S.make_features(cfg, 5)
    assert f["depth"].shape == (48, 41, 8, 22) and f["feat"].shape == (48, 64, 8, 22)
    np.testing.assert_allclose(f["depth"].sum(1), 1.0, atol=1e-5)
    alg = cfg.algorithmic_bytes()
    assert alg["fwd"] == 85468160 and alg["bwd"] == 89016320 and alg["total"] == 174484480
    assert S.config("config4").algorithmic_bytes()["total"] == 522330112


def test_geometry_provenance_records_follow_versions():
    """patch remembers which calibration a geometry tensor came from only while neither was modified."""
    import gc
    import torch
    from lss2_multimodal_nu_b200 import patch
    g = torch.zeros(2, 3)
    cal = tuple(torch.ones(2) for _ in range(5))
    patch._remember_calib(g, cal)
    assert patch._calib_of(g) == cal
    assert patch._calib_of(g.clone()) is None               # another tensor, however equal
    g += 1                                                   # in-place edit: quantise what you are given
    assert patch._calib_of(g) is None
    patch._remember_calib(g, cal)
    cal[3].mul_(2)                                           # calibration edited after get_geometry
    assert patch._calib_of(g) is None
    n = len(patch._GEOM_CALIB)
    del g
    gc.collect()
    assert len(patch._GEOM_CALIB) == n - 1                   # the record dies with the tensor


def test_install_everywhere_patches_script_defined_models_at_first_forward():
    """A model class the class-level patch cannot know (the reference's PreTrainingModel lives in the script
    run as __main__) is patched by the global forward pre-hook the first time it is called."""
    import torch
    from lss2_multimodal_nu_b200 import patch

    class ScriptModel(torch.nn.Module):
        def __init__(self):
            super().__init__()
            for n, shape in (("frustum", (4, 2, 3, 3)), ("dx", (3,)), ("bx", (3,))):
                setattr(self, n, torch.nn.Parameter(torch.ones(shape), requires_grad=False))
            self.nx = torch.nn.Parameter(torch.tensor([8, 8, 1]), requires_grad=False)

        def get_geometry(self, *a): return "stock"
        def voxel_pooling(self, g, x): return "stock"
        def forward(self, x): return x

    class Other(torch.nn.Module):
        def forward(self, x): return x

    m, o = ScriptModel(), Other()
    keys = list(m.state_dict().keys())
    assert patch.has_lss_surface(m) and not patch.has_lss_surface(o)
    patch.install_everywhere()
    try:
        o(torch.zeros(1)); m(torch.zeros(1))
    finally:
        patch.uninstall_everywhere()
    assert m.__dict__.get(patch._INSTALLED_ATTR) and not o.__dict__.get(patch._INSTALLED_ATTR)
    assert m.voxel_pooling.__func__ is patch.voxel_pooling and m.get_geometry.__func__ is patch.get_geometry
    assert list(m.state_dict().keys()) == keys               # nothing added to the checkpoint
    # calibration-shaped call arguments are recognised (model(imgs, rots, trans, intrins, post_rots, post_trans))
    B, N = 2, 6
    args = (torch.zeros(B, N, 3, 8, 8), torch.zeros(B, N, 3, 3), torch.zeros(B, N, 3), torch.zeros(B, N, 3, 3),
            torch.zeros(B, N, 3, 3), torch.zeros(B, N, 3))
    assert not patch._looks_like_calibration(args)           # CPU tensors: nothing to prefetch
    assert not patch._looks_like_calibration(args[:3])


def test_make_frustum_and_hash_field_agree_with_their_twins():
    import numpy as np
    import torch
    import lss_oracle as O
    from lss2_multimodal_nu_b200 import functional as F, synthetic as S
    for final_dim, dbound in (((128, 352), (4.0, 45.0, 1.0)), ((256, 704), (1.0, 60.0, 0.5))):
        ours = F.make_frustum(final_dim, 16, dbound).numpy()
        want = O.create_frustum(final_dim, 16, dbound)
        assert ours.shape == want.shape and (ours.view(np.uint32) == want.view(np.uint32)).all()
    idx = np.arange(0, 3_000_000_000, 7_654_321, dtype=np.int64)
    a = S.hash_field_np(idx, 77)
    b = S.hash_field_torch(torch.from_numpy(idx), 77).numpy()
    assert (a.view(np.uint32) == b.view(np.uint32)).all() and a.min() >= -0.5 and a.max() < 0.5


def test_bench_reference_arm_prints_the_contract_line():
    """`bench.py --impl reference` (the arm the driver runs beside ours): one JSON line with the same metric, unit
    and config as our arm, a CPU baseline description, zero copies, no GPU work -- it must run without a GPU."""
    import json
    import subprocess
    import sys as _sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    out = subprocess.run([_sys.executable, os.path.join(root, "bench.py"), "--impl", "reference", "--steps", "1",
                          "--warmup", "1"], capture_output=True, text=True, timeout=600, cwd=root)
    assert out.returncode == 0, out.stderr[-2000:]
    lines = [ln for ln in out.stdout.splitlines() if ln.startswith("{")]
    assert len(lines) == 1
    d = json.loads(lines[0])
    _sys.path.insert(0, root)
    import bench
    from lss2_multimodal_nu_b200 import synthetic as S
    assert d["impl"] == "reference" and d["metric"] == bench.METRIC and d["unit"] == bench.UNIT
    assert d["config"] == bench.base_config(S.config("config2"), 1, 4)          # what our arm prints for N = 1
    assert d["higher_is_better"] is True and d["vs_baseline"] is None and d["gpu_launches"] == 0
    assert d["cpu_baseline"]["kind"] == "port" and d["cpu_baseline"]["cores"] >= 1
    assert d["cpu_baseline"]["value"] == d["value"] == d["e2e"]["value"] > 0
    assert d["e2e"]["h2d_bytes_per_step"] == 0 and d["e2e"]["d2h_bytes_per_step"] == 0
    stock = d["reference_torch_cpu"]
    assert "unavailable" in stock or (stock["kind"] == "reference" and 0 < stock["value"] < d["value"] * 10)
