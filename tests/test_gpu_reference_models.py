"""The UNMODIFIED reference model classes on the B200, stock and with the drop-in installed.

oracle/_ref/ holds a byte-for-byte staged copy of the reference's Python files (oracle/stage_ref.py; in the
build container /root/reference is used directly) and oracle/shims/ the stand-in backbone packages; the
classes under test -- LSS, BEV_TXT (src/model_baseline.py, src/model_BEV_TXT.py), VoVNetBEVTransformer
(src/model_vovnet_transformer.py), PreTrainingModel (pre_train_vovnet.py) -- and every line of their
forward passes are the reference's own.  Each test runs the class as written (the reference's PyTorch
lift-splat on the GPU), installs lss2_multimodal_nu_b200.patch, runs it again on the same weights and inputs,
and compares.

What "equal" means here.  Index parity is pinned against the reference on the CPU (fixtures, SHA-256 digests
in test_gpu_parity.py).  torch's own GPU geometry is NOT bit-identical to its CPU geometry -- the batched
3x3 products go through cuBLAS with FMA -- which moves a few points per million across a voxel border
(test_gpu_geometry_of_torch_vs_ours measures it).  So values are compared on identical geometry: the stock
voxel_pooling is fed the geometry tensor our get_geometry produced (bit-exact to the CPU reference).
"""
import os
import sys

import numpy as np
import pytest
import torch

import ref_import
from lss2_multimodal_nu_b200 import functional as F
from lss2_multimodal_nu_b200 import patch
from lss2_multimodal_nu_b200 import synthetic as S

pytestmark = [pytest.mark.gpu,
              pytest.mark.skipif(not ref_import.available(), reason="reference tree not staged (oracle/stage_ref.py)")]
DEV = "cuda:0"
CAL = ("rots", "trans", "intrins", "post_rots", "post_trans")
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def calib(cfg, seed):
    c = S.make_calibration(cfg, seed)
    return [torch.from_numpy(c[k]).to(DEV) for k in CAL]


def images(cfg, seed):
    g = torch.Generator(device=DEV); g.manual_seed(seed)
    return torch.randn(cfg.B, cfg.N, 3, *cfg.final_dim, device=DEV, generator=g)


def close(a, ref, rtol, atol):
    a = a.detach().double().cpu().numpy(); ref = ref.detach().double().cpu().numpy()
    err = np.abs(a - ref); bound = atol + rtol * np.abs(ref)
    assert (err <= bound).all(), "max abs err %.3e, worst excess %.3e, %d/%d out of tolerance" % (
        err.max(), (err - bound).max(), (err > bound).sum(), err.size)


def reference_voxels_fp64(m, geom, x):
    """The reference's own voxel_pooling evaluated with float64 sums (index math stays
    float32 because geom / bx / dx are float32)."""
    torch.set_default_dtype(torch.float64)
    try:
        return m.voxel_pooling(geom, x.double())
    finally:
        torch.set_default_dtype(torch.float32)


@pytest.mark.parametrize("module,cls", [("model_baseline", "LSS"), ("model_baseline", "BEV_TXT"),
                                        ("model_BEV_TXT", "LSS"), ("model_BEV_TXT", "BEV_TXT")])
def test_lss_family_stock_vs_patched(module, cls):
    """get_geometry / get_cam_feats / voxel_pooling / get_voxels of the four EfficientNet-family classes."""
    cfg = S.config("config1", B=2)
    torch.manual_seed(1)
    m = ref_import.build_lss(cfg.B, cfg.grid_conf(), cfg.data_aug_conf(), cls=cls, module=module).to(DEV).eval()
    cal = calib(cfg, 3)
    g = torch.Generator(device=DEV); g.manual_seed(2)
    x = torch.randn(cfg.B * cfg.N, 512, cfg.fH, cfg.fW, device=DEV, generator=g)     # encoder output (model_baseline.py:137)
    keys = list(m.state_dict().keys())

    stock_vp, stock_feats = type(m).voxel_pooling, type(m).get_cam_feats     # the class keeps the reference's methods
    # ---- stock: the reference's PyTorch code on the GPU ----
    with torch.no_grad():
        geom_stock = m.get_geometry(*cal)
        lifted = m.get_cam_feats(x)                        # the materialised B x N x D x fH x fW x C tensor
    # ---- patched (instance level) ----
    patch.install(m)
    assert list(m.state_dict().keys()) == keys             # strict=True checkpoints keep loading
    geom = m.get_geometry(*cal)
    assert geom.shape == geom_stock.shape and geom.dtype == torch.float32
    # our geometry is bit-exact to torch's CPU geometry (fixtures); torch's GPU geometry is within rounding of it
    assert torch.allclose(geom, geom_stock, rtol=1e-5, atol=1e-4)
    with torch.no_grad():
        torch.set_default_dtype(torch.float64)
        try:
            want = stock_vp(m, geom, lifted.double())       # the reference's splat, float64 sums, on OUR geometry
        finally:
            torch.set_default_dtype(torch.float32)
    xg = x.clone().requires_grad_(True)
    bev = m.get_voxels(xg, *cal)                           # fused entry
    assert tuple(bev.shape) == tuple(want.shape)
    close(bev, want, 1e-5, 2e-6)
    bev2 = m.voxel_pooling(m.get_geometry(*cal), m.get_cam_feats(xg))     # the three-method path
    close(bev2, want, 1e-5, 2e-6)
    # gradients through the reference's own CamEncode vs autograd of the stock path (float32 both sides)
    w = torch.randn_like(bev)
    (bev * w).sum().backward()
    gx = xg.grad.clone(); gw = m.camencode.depthnet.weight.grad.clone()
    m.camencode.depthnet.weight.grad = None
    xs = x.clone().requires_grad_(True)
    ref = stock_vp(m, geom, stock_feats(m, xs))
    (ref * w).sum().backward()
    # (the stock float32 path carries the prefix-sum cancellation error of QuickCumsum, SURVEY.md 7.3-3, and the
    # weight gradient sums 8 448 pixels of it: the bar is relative to the largest gradient entry)
    close(gx, xs.grad, 1e-3, 2e-5 * float(xs.grad.abs().max()) + 2e-5)
    gw_ref = m.camencode.depthnet.weight.grad
    close(gw, gw_ref, 1e-3, 2e-4 * float(gw_ref.abs().max()))


def test_gpu_geometry_of_torch_vs_ours():
    """How far torch's own GPU get_geometry is from its CPU result (= ours, bit for bit): a handful of
    points in a million change voxel.  Informational bound, asserted loosely."""
    cfg = S.config("config2")
    m = ref_import.build_lss(cfg.B, cfg.grid_conf(), cfg.data_aug_conf()).to(DEV).eval()
    cal = calib(cfg, 1234)
    with torch.no_grad():
        stock = m.get_geometry(*cal)
    grid = F.GridSpec.from_tensors(m.dx, m.bx, m.nx)
    ours = F.geometry(*F.frustum_axes(m.frustum), *cal, grid, want_geom=True)
    q_stock = F.quantize_rank(stock, grid, cfg.B)
    moved = int((q_stock["cells"] != ours["cells"]).sum())
    assert moved <= max(8, cfg.P // 20000), "%d of %d points changed voxel" % (moved, cfg.P)


@pytest.mark.parametrize("module", ["model_baseline", "model_BEV_TXT"])
def test_bev_txt_full_forward_and_training_step(module):
    """BEV_TXT.forward (model_BEV_TXT.py:278-334) end to end, stock vs patched, then one optimiser step of the
    patched model (train.py:49-65: MultiLoss-shaped loss, clip_grad_norm_, Adam)."""
    cfg = S.config("config1", B=2)
    torch.manual_seed(5)
    m = ref_import.build_lss(cfg.B, cfg.grid_conf(), cfg.data_aug_conf(), cls="BEV_TXT", module=module,
                             backbone=True).to(DEV).eval()
    cal = calib(cfg, 9)
    imgs = images(cfg, 4)
    with torch.no_grad():
        stock = m(imgs, *cal)
    patch.install(m)
    with torch.no_grad():
        ours = m(imgs, *cal)
    assert patch._cache(m).prefetch_hits >= 1              # the plan was built while the backbone ran
    for a, b in zip(ours, stock):
        assert a.shape == b.shape
        close(a, b, 2e-3, 2e-3)                            # downstream conv nets amplify the 1e-5 pooling differences
    m.train()
    opt = torch.optim.Adam(m.parameters(), lr=1e-3, weight_decay=1e-7)
    bev, act, desc = m(imgs, *cal)
    loss = bev.float().pow(2).mean() + act.float().pow(2).mean() + desc.float().pow(2).mean()
    opt.zero_grad(); loss.backward()
    torch.nn.utils.clip_grad_norm_(m.parameters(), 5.0)
    opt.step()
    assert all(torch.isfinite(p).all() for p in m.parameters())
    assert m.camencode.depthnet.weight.grad.abs().sum() > 0


def test_vovnet_transformer_stock_vs_patched_and_autocast():
    """VoVNetBEVTransformer.forward (model_vovnet_transformer.py:556-639): cam_encode handle -> .view / .permute ->
    patched voxel_pooling; then the AMP path of train_vovnet_transformer.py:196."""
    cfg = S.config("config1", B=2)
    torch.manual_seed(8)
    m = ref_import.build_vovnet(cfg.B, cfg.grid_conf(), cfg.data_aug_conf()).to(DEV).eval()
    cal = calib(cfg, 6)
    imgs = images(cfg, 3)
    with torch.no_grad():
        stock = m(imgs, *cal)
    keys = list(m.state_dict().keys())
    patch.install(m)
    assert list(m.state_dict().keys()) == keys
    # the lazy handle really takes the fused path: voxel_pooling receives the two factors, in pooling layout,
    # not a materialised B x N x D x fH x fW x C tensor
    from lss2_multimodal_nu_b200.lazy import LiftedFrustum
    seen = []
    inner = m.voxel_pooling

    def spy(geom_feats, x):
        seen.append((type(x), isinstance(x, LiftedFrustum) and x.is_pooling_layout(), tuple(x.shape)))
        return inner(geom_feats, x)
    object.__setattr__(m, "voxel_pooling", spy)
    with torch.no_grad():
        ours = m(imgs, *cal)
    object.__setattr__(m, "voxel_pooling", inner)
    assert seen and seen[0][0] is LiftedFrustum and seen[0][1]
    assert seen[0][2] == (cfg.B, cfg.N, m.D, cfg.fH, cfg.fW, m.C)
    for a, b in zip(ours, stock):
        close(a, b, 5e-3, 5e-3)
    # AMP: geometry stays float32 (the reference drops to half here), features arrive as half,
    # the BEV map is float32, gradients reach the half-precision producers
    m.train()
    with torch.autocast("cuda", dtype=torch.float16):
        geom = m.get_geometry(*cal)
        assert geom.dtype == torch.float32
        bev_seg, action, desc = m(imgs, *cal)
    loss = bev_seg.float().pow(2).mean() + action.float().pow(2).mean() + desc.float().pow(2).mean()
    loss.backward()
    g = m.cam_encode.feat_proj.weight.grad
    assert g is not None and torch.isfinite(g).all() and g.abs().sum() > 0


def test_pretraining_model_through_run_launcher(tmp_path):
    """`python -m lss2_multimodal_nu_b200.run script.py`: the model class of pre_train_vovnet.py is defined in
    the script module, out of reach of the class-level patch; the launcher's forward pre-hook installs the drop-in
    at the first call.  Compared with the same script run without the launcher (stock PyTorch path)."""
    import subprocess
    script = os.path.join(ROOT, "tests", "scripts", "run_pretraining_model.py")
    out_p, out_s = str(tmp_path / "patched.pt"), str(tmp_path / "stock.pt")
    env = dict(os.environ, PYTHONPATH=ROOT + os.pathsep + os.environ.get("PYTHONPATH", ""))
    subprocess.run([sys.executable, "-m", "lss2_multimodal_nu_b200.run", script, out_p], check=True, env=env, cwd=ROOT)
    subprocess.run([sys.executable, script, out_s], check=True, env=env, cwd=ROOT)
    a, b = torch.load(out_p), torch.load(out_s)
    assert a["patched"] is True and b["patched"] is False
    assert a["state"].keys() == b["state"].keys()
    for k in a["state"]:
        assert torch.equal(a["state"][k], b["state"][k]), k
    close(a["out"], b["out"], 5e-3, 5e-3)


def test_consumer_conv_runs_channels_last_without_transposes():
    """the pooled map goes into bevencode.conv1 (reference src/modules.py:99, 7x7 stride 2)
    channels_last and its gradient comes back channels_last: no layout change on either side."""
    _, _, modules = ref_import.load()
    cfg = S.config("config1", B=2)
    torch.manual_seed(2)
    enc = modules.BevEncode(inC=cfg.C, outC=4).to(DEV)
    cal = calib(cfg, 1)
    grid = F.GridSpec.from_bounds(cfg.xbound, cfg.ybound, cfg.zbound)
    fr = F.make_frustum(cfg.final_dim, cfg.downsample, cfg.dbound).to(DEV)
    plan = F.build_plan(*F.frustum_axes(fr), *cal, grid)
    ft = S.make_features(cfg, 1)
    depth = torch.from_numpy(ft["depth"]).to(DEV).requires_grad_(True)
    feat = torch.from_numpy(ft["feat"]).to(DEV).requires_grad_(True)
    before = F.NHWC_TRANSPOSES
    names = []
    with torch.profiler.profile(activities=[torch.profiler.ProfilerActivity.CUDA]) as prof:
        bev = F.lift_splat(depth, feat, plan)
        assert bev.is_contiguous(memory_format=torch.channels_last)
        y = enc.conv1(bev)
        y.float().pow(2).mean().backward()
        torch.cuda.synchronize()
    names = [e.key for e in prof.key_averages()]
    assert F.NHWC_TRANSPOSES == before, "dBEV did not arrive channels_last"
    assert not any(("nchwToNhwc" in n) or ("nhwcToNchw" in n) for n in names), names
    assert depth.grad is not None and torch.isfinite(depth.grad).all()


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs two GPUs")
def test_module_on_a_non_current_device():
    """train.py / predict.py move the model with .to(f'cuda:{gpuid}') and never call set_device (gpuid defaults to 1):
    the tensors' device decides where the kernels run."""
    cfg = S.config("config1")
    torch.cuda.set_device(0)
    dev1 = torch.device("cuda", 1)
    m = ref_import.build_lss(cfg.B, cfg.grid_conf(), cfg.data_aug_conf()).to(dev1).eval()
    c = S.make_calibration(cfg, 2)
    cal = [torch.from_numpy(c[k]).to(dev1) for k in CAL]
    x = torch.randn(cfg.B * cfg.N, 512, cfg.fH, cfg.fW, device=dev1)
    with torch.no_grad(), torch.cuda.device(dev1):
        geom_stock = m.get_geometry(*cal)
        lifted = m.get_cam_feats(x)
        stock = reference_voxels_fp64(m, geom_stock, lifted)
    patch.install(m)
    assert torch.cuda.current_device() == 0
    with torch.no_grad():
        ours = m.voxel_pooling(geom_stock, m.get_cam_feats(x))     # the stock geometry tensor: quantised as given
        fused = m.get_voxels(x, *cal)
    torch.cuda.synchronize(dev1)
    assert ours.device == dev1 and torch.cuda.current_device() == 0
    close(ours, stock, 1e-5, 2e-6)
    assert torch.isfinite(fused).all()
