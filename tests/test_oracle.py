"""CPU tests: the oracle (numpy + C restatements) against the committed golden
fixtures, which are outputs of the UNMODIFIED reference (oracle/make_golden.py)."""
import hashlib
import json
import os

import numpy as np
import pytest

import c_oracle as CO
import lss_oracle as O
from lss2_multimodal_nu_b200 import synthetic as S

CAL = ("rots", "trans", "intrins", "post_rots", "post_trans")
FIXTURES = ["tiny", "edge_none_kept", "edge_one_voxel", "edge_randn_calib", "edge_nonfinite"]


def sha(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


def load(golden_dir, name):
    return dict(np.load(os.path.join(golden_dir, name + ".npz")))


def same_bits(a, b):
    a = np.ascontiguousarray(a); b = np.ascontiguousarray(b)
    return a.shape == b.shape and a.dtype == b.dtype and bool((a.view(np.uint8) == b.view(np.uint8)).all())


def same_float_bits_or_nan(a, b):
    nan = np.isnan(b)
    return bool((np.isnan(a) == nan).all() and (a.view(np.uint32)[~nan] == b.view(np.uint32)[~nan]).all())


def test_inverse3x3_known_answers(golden_dir):
    g = load(golden_dir, "inverse3x3")
    assert same_bits(O.inverse3x3(g["A"]), g["inv"])
    assert same_bits(CO.inverse3x3(g["A"]), g["inv"])


def test_linspace_arange_known_answers(golden_dir):
    g = load(golden_dir, "linspace")
    for k, v in g.items():
        kind, *args = k.split("_")
        if kind == "lin":
            got = O._linspace_f32(0, int(args[0]), int(args[1]))
        else:
            got = O._arange_f32(*map(float, args))
        assert same_bits(got, v), k


def test_grid_constants_and_frustum(golden_dir):
    g = load(golden_dir, "tiny")
    cfg = S.config("tiny")
    dx, bx, nx = O.gen_dx_bx(cfg.xbound, cfg.ybound, cfg.zbound)
    assert same_bits(dx, g["dx"]) and same_bits(bx, g["bx"]) and (nx == g["nx"]).all()
    assert same_bits(O.create_frustum(cfg.final_dim, cfg.downsample, cfg.dbound), g["frustum"])
    # 102.4 / 0.2 must give 512 cells (python-float quotient truncation)
    assert S.config("config5").nx == (512, 512, 1) and S.config("config4").D == 59 and S.config("config5").D == 118


@pytest.mark.parametrize("name", FIXTURES)
def test_geometry_bit_exact(golden_dir, name):
    g = load(golden_dir, name)
    ipr, comb = O.camera_prep(g["rots"], g["intrins"], g["post_rots"])
    assert same_bits(ipr, g["inv_post_rots"]) and same_bits(comb, g["combine"])
    geom = O.get_geometry(g["frustum"], *(g[k] for k in CAL))
    assert same_float_bits_or_nan(geom, g["geom"])
    fr = g["frustum"]
    geom_c = CO.geometry(fr[0, 0, :, 0], fr[0, :, 0, 1], fr[:, 0, 0, 2], *(g[k] for k in CAL))
    assert same_float_bits_or_nan(geom_c, g["geom"])


@pytest.mark.parametrize("name", FIXTURES)
def test_index_pipeline_bit_exact(golden_dir, name):
    g = load(golden_dir, name)
    B = g["trans"].shape[0]
    ip = O.index_pipeline(g["geom"], g["dx"], g["bx"], g["nx"], B)
    assert (ip["coords"] == g["coords"]).all()
    assert (ip["kept"] == g["kept"]).all()
    assert (ip["ranks"] == g["ranks"]).all()
    assert (ip["sorted_ranks"] == g["ranks"][g["sorts"]]).all()
    assert (ip["last_mask"] == g["last_mask"]).all()
    if len(g["ranks"]) >= 32768:      # torch's argsort is only stable on its radix path
        assert (ip["sorts"] == g["sorts"]).all()
    ic = CO.index(g["geom"], g["dx"], g["bx"], g["nx"], B)
    assert (ic["coords"] == g["coords"]).all() and (ic["kept"] == g["kept"]).all()
    assert (ic["ranks"] == g["ranks"]).all() and (ic["sorts"] == ip["sorts"]).all()


@pytest.mark.parametrize("name", FIXTURES)
def test_pooling_values(golden_dir, name):
    g = load(golden_dir, name)
    B, N = g["trans"].shape[:2]
    x32 = O.lift(g["depth"], g["feat"])
    bev32, ip = O.voxel_pooling(g["geom"], x32, g["dx"], g["bx"], g["nx"], B, mode="reference")
    assert same_bits(bev32, g["bev32"])                      # the reference's own float32 bits
    x64 = O.lift(g["depth"].astype(np.float64), g["feat"].astype(np.float64))
    bev64, _ = O.voxel_pooling(g["geom"], x64, g["dx"], g["bx"], g["nx"], B, mode="exact")
    np.testing.assert_allclose(bev64, g["bev64"], rtol=1e-12, atol=1e-12)
    dd, df = O.voxel_pooling_backward(g["dbev"], ip, g["depth"], g["feat"], g["nx"], B)
    np.testing.assert_allclose(dd, g["d_depth64"], rtol=1e-12, atol=1e-12)
    np.testing.assert_allclose(df, g["d_feat64"], rtol=1e-12, atol=1e-12)
    # C restatement: reference bits in mode 0, float64 values in mode 1
    b0, dd0, df0, K, V = CO.step(g["depth"], g["feat"], g["geom"], g["dbev"], g["dx"], g["bx"], g["nx"], B, N, mode=0)
    assert same_bits(b0, g["bev32"]) and (K, V) == (len(g["ranks"]), int(g["last_mask"].sum()))
    np.testing.assert_allclose(dd0, g["d_depth32"], rtol=1e-5, atol=2e-6)
    np.testing.assert_allclose(df0, g["d_feat32"], rtol=1e-5, atol=2e-6)
    b1, dd1, df1, _, _ = CO.step(g["depth"], g["feat"], g["geom"], g["dbev"], g["dx"], g["bx"], g["nx"], B, N, mode=1)
    np.testing.assert_allclose(b1, g["bev64"], rtol=1e-12, atol=1e-12)
    np.testing.assert_allclose(dd1, g["d_depth64"], rtol=1e-12, atol=1e-12)
    np.testing.assert_allclose(df1, g["d_feat64"], rtol=1e-12, atol=1e-12)


def test_truncation_keeps_minus_one_to_zero():
    """.long() truncates: a scaled coordinate in (-1, 0) lands in voxel 0 and is kept."""
    dx = np.array([0.5, 0.5, 20.0], np.float32); bx = np.array([-49.75, -49.75, 0.0], np.float32)
    geom = np.array([[-50.2, 0.0, -25.0], [-50.6, 0.0, 0.0], [49.99, 49.99, 9.9], [50.0, 0.0, 0.0],
                     [np.nan, 0.0, 0.0], [np.inf, 0.0, 0.0]], np.float32)
    c = O.quantize(geom, dx, bx)
    k = O.kept_mask(c, [200, 200, 1])
    assert c[0].tolist() == [0, 100, 0] and k.tolist() == [True, False, True, False, False, False]


def test_config1_fixture(golden_dir):
    g = load(golden_dir, "config1")
    cfg = S.config("config1")
    cal = S.make_calibration(cfg); ft = S.make_features(cfg); dbev = S.make_dbev(cfg)
    for k in CAL:
        assert same_bits(cal[k], g[k])
    assert sha(ft["depth"]) + sha(ft["feat"]) + sha(dbev) == str(g["inputs_sha"])
    dx, bx, nx = O.gen_dx_bx(cfg.xbound, cfg.ybound, cfg.zbound)
    geom = O.get_geometry(O.create_frustum(cfg.final_dim, cfg.downsample, cfg.dbound), **cal)
    assert sha(geom) == str(g["geom_sha"])
    ip = O.index_pipeline(geom, dx, bx, nx, cfg.B)
    assert (ip["coords"] == g["coords"]).all()
    assert (ip["kept"] == np.unpackbits(g["kept"])[:cfg.P].astype(bool)).all()
    assert (ip["ranks"] == g["ranks"]).all() and (ip["sorts"] == g["sorts"]).all()
    assert (ip["last_mask"] == np.unpackbits(g["last_mask"])[:len(g["ranks"])].astype(bool)).all()
    b0, _, _, _, _ = CO.step(ft["depth"], ft["feat"], geom, dbev, dx, bx, nx, cfg.B, cfg.N, mode=0)
    assert sha(b0) == str(g["bev32_sha"])
    b1, dd1, df1, _, _ = CO.step(ft["depth"], ft["feat"], geom, dbev, dx, bx, nx, cfg.B, cfg.N, mode=1)
    pick = tuple(g["bev_pick"].T.astype(np.int64))
    np.testing.assert_allclose(b1[pick], g["bev64_at"], rtol=1e-12, atol=1e-12)
    np.testing.assert_allclose(dd1, g["d_depth64"], rtol=1e-12, atol=1e-12)
    np.testing.assert_allclose(df1[:, ::8], g["d_feat64_sub"], rtol=1e-12, atol=1e-12)


def test_config2_digests(golden_dir):
    """Headline shape: the C restatement reproduces the reference's index tensors and its float32
    BEV output bit for bit (SHA-256 of the full tensors)."""
    with open(os.path.join(golden_dir, "config2.json")) as f:
        g = json.load(f)
    cfg = S.config("config2")
    cal = S.make_calibration(cfg); ft = S.make_features(cfg); dbev = S.make_dbev(cfg)
    assert "".join(sha(cal[k]) for k in sorted(cal)) == g["sha256"]["calibration"]
    us, vs, ds = O.frustum_axes(cfg.final_dim, cfg.downsample, cfg.dbound)
    dx, bx, nx = O.gen_dx_bx(cfg.xbound, cfg.ybound, cfg.zbound)
    geom = CO.geometry(us, vs, ds, **cal)
    assert sha(geom) == g["sha256"]["geom"]
    ic = CO.index(geom, dx, bx, nx, cfg.B)
    assert sha(ic["coords"].astype(np.int32)) == g["sha256"]["coords_i32"]
    assert sha(ic["kept"].astype(np.uint8)) == g["sha256"]["kept_u8"]
    assert sha(ic["ranks"].astype(np.int32)) == g["sha256"]["ranks_i32"]
    assert sha(ic["sorts"].astype(np.int32)) == g["sha256"]["sorts_i32"]
    b0, _, _, K, V = CO.step(ft["depth"], ft["feat"], geom, dbev, dx, bx, nx, cfg.B, cfg.N, mode=0, backward=False)
    assert (K, V) == (g["K"], g["V"]) and sha(b0) == g["sha256"]["bev32"]


def test_oracle_reproduces_reference_digests_at_config4_full_size(golden_dir):
    """The numpy restatement against the reference's own index tensors at BASELINE.json config 4, full batch
    (B=16, 4 M points): geometry, truncated coordinates, kept mask, ranks, sort order, interval mask."""
    import hashlib
    import json
    import os
    from lss2_multimodal_nu_b200 import synthetic as S

    def sha(a):
        return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()

    with open(os.path.join(golden_dir, "config4.json")) as f:
        g = json.load(f)
    cfg = S.config("config4")
    cal = S.make_calibration(cfg, 1234)
    fr = O.create_frustum(cfg.final_dim, cfg.downsample, cfg.dbound)
    dx, bx, nx = O.gen_dx_bx(cfg.xbound, cfg.ybound, cfg.zbound)
    geom = O.get_geometry(fr, **cal)
    assert sha(geom) == g["sha256"]["geom"]
    ip = O.index_pipeline(geom, dx, bx, nx, cfg.B)
    assert len(ip["ranks"]) == g["K"] and len(ip["interval_start"]) == g["V"]
    assert sha(ip["ranks"].astype(np.int32)) == g["sha256"]["ranks_i32"]
    assert sha(ip["sorts"].astype(np.int32)) == g["sha256"]["sorts_i32"]
