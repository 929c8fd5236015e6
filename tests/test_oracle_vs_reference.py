"""Container-only: run the UNMODIFIED reference side by side with the oracle on fresh seeds
(beyond the committed fixtures).  Skipped where /root/reference is absent (GPU box)."""
import numpy as np
import pytest
import torch

import lss_oracle as O
from lss2_multimodal_nu_b200 import synthetic as S

pytestmark = pytest.mark.reference


@pytest.mark.parametrize("cname,seed", [("tiny", 3), ("tiny", 4), ("config1", 21)])
def test_reference_side_by_side(cname, seed):
    import ref_import
    cfg = S.config(cname)
    cal = S.make_calibration(cfg, seed); ft = S.make_features(cfg, seed); dbev = S.make_dbev(cfg, seed)
    m = ref_import.build_lss(cfg.B, cfg.grid_conf(), cfg.data_aug_conf())
    t = {k: torch.from_numpy(v) for k, v in cal.items()}
    with torch.no_grad():
        geom_ref = m.get_geometry(t["rots"], t["trans"], t["intrins"], t["post_rots"], t["post_trans"]).numpy()
    geom = O.get_geometry(O.create_frustum(cfg.final_dim, cfg.downsample, cfg.dbound), **cal)
    assert (geom.view(np.uint32) == geom_ref.view(np.uint32)).all()
    depth = torch.from_numpy(ft["depth"]).requires_grad_(True)
    feat = torch.from_numpy(ft["feat"]).requires_grad_(True)
    x = (depth.unsqueeze(1) * feat.unsqueeze(2)).view(cfg.B, cfg.N, cfg.C, cfg.D, cfg.fH, cfg.fW)
    out = m.voxel_pooling(torch.from_numpy(geom_ref), x.permute(0, 1, 3, 4, 5, 2))
    out.backward(torch.from_numpy(dbev))
    dx, bx, nx = O.gen_dx_bx(cfg.xbound, cfg.ybound, cfg.zbound)
    bev32, ip = O.voxel_pooling(geom, O.lift(ft["depth"], ft["feat"]), dx, bx, nx, cfg.B, mode="reference")
    assert (bev32.view(np.uint32) == out.detach().numpy().view(np.uint32)).all()
    dd, df = O.voxel_pooling_backward(dbev, ip, ft["depth"], ft["feat"], nx, cfg.B)
    np.testing.assert_allclose(depth.grad.numpy(), dd, rtol=1e-4, atol=2e-5)   # the reference's own fp32 error
    np.testing.assert_allclose(feat.grad.numpy(), df, rtol=1e-4, atol=2e-5)


def test_use_quickcumsum_toggle_agrees():
    """The reference's only built-in cross-check (model_baseline.py:114-117): forward identical."""
    import ref_import
    cfg = S.config("tiny")
    cal = S.make_calibration(cfg, 9); ft = S.make_features(cfg, 9)
    m = ref_import.build_lss(cfg.B, cfg.grid_conf(), cfg.data_aug_conf())
    t = {k: torch.from_numpy(v) for k, v in cal.items()}
    with torch.no_grad():
        geom = m.get_geometry(t["rots"], t["trans"], t["intrins"], t["post_rots"], t["post_trans"])
        x = (torch.from_numpy(ft["depth"]).unsqueeze(1) * torch.from_numpy(ft["feat"]).unsqueeze(2))
        x = x.view(cfg.B, cfg.N, cfg.C, cfg.D, cfg.fH, cfg.fW).permute(0, 1, 3, 4, 5, 2)
        a = m.voxel_pooling(geom, x)
        m.use_quickcumsum = False
        b = m.voxel_pooling(geom, x)
    assert torch.equal(a, b)
