"""Generate tests/golden/* by running the UNMODIFIED reference in this container.

    python oracle/make_golden.py

Needs /root/reference (build container only).  The reference has no golden
vectors of its own, so these fixtures -- outputs of the
reference's own get_geometry / voxel_pooling / QuickCumsum / torch.inverse /
torch.linspace on seeded synthetic inputs -- are what pins the oracle
(oracle/lss_oracle.py, oracle/lss_oracle.c) and, through it, the CUDA path.

Fixtures written:
  tiny.npz          every intermediate, full tensors (B=2,N=3,D=16,4x6,C=8,32x32x4)
  edge_*.npz        adversarial cases: nothing kept, one voxel, randn/NaN/inf calibrations
  config1.npz       B=1 nuScenes shape: full index tensors, sampled values
  config2.json      B=8: SHA-256 digests of the index tensors and of the fp32
                    output, counts, and sampled fp64 values
  inverse3x3.npz    known answers of torch.inverse on 3x3 matrices
  linspace.npz      known answers of torch.linspace / torch.arange (frustum axes)
"""
import hashlib
import json
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, HERE)
sys.path.insert(0, ROOT)

import ref_import  # noqa: E402
from lss2_multimodal_nu_b200 import synthetic as S  # noqa: E402

OUT = os.path.join(ROOT, "tests", "golden")


def sha(a: np.ndarray) -> str:
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


def run_reference(cfg, cal, ft, dbev, want_grads=True):
    """Replay the reference's own statements and capture every intermediate."""
    tools, _, _ = ref_import.load()
    m = ref_import.build_lss(cfg.B, cfg.grid_conf(), cfg.data_aug_conf())
    m.camC = cfg.C
    t = {k: torch.from_numpy(v) for k, v in cal.items()}
    with torch.no_grad():
        geom = m.get_geometry(t["rots"], t["trans"], t["intrins"], t["post_rots"], t["post_trans"])
        inv_post_rots = torch.inverse(t["post_rots"])
        combine = t["rots"].matmul(torch.inverse(t["intrins"]))
        # index intermediates: the reference's statements (model_baseline.py:92-110)
        B = cfg.B
        Nprime = cfg.P
        gf = ((geom - (m.bx - m.dx / 2.)) / m.dx).long().view(Nprime, 3)
        batch_ix = torch.cat([torch.full([Nprime // B, 1], ix, dtype=torch.long) for ix in range(B)])
        gf = torch.cat((gf, batch_ix), 1)
        kept = (gf[:, 0] >= 0) & (gf[:, 0] < m.nx[0]) & (gf[:, 1] >= 0) & (gf[:, 1] < m.nx[1]) \
            & (gf[:, 2] >= 0) & (gf[:, 2] < m.nx[2])
        gk = gf[kept]
        ranks = gk[:, 0] * (m.nx[1] * m.nx[2] * B) + gk[:, 1] * (m.nx[2] * B) + gk[:, 2] * B + gk[:, 3]
        sorts = ranks.argsort()
        sr = ranks[sorts]
        last = torch.ones(sr.shape[0], dtype=torch.bool)
        if sr.shape[0] > 1:
            last[:-1] = sr[1:] != sr[:-1]
    res = {"frustum": m.frustum.detach().numpy(), "dx": m.dx.numpy(), "bx": m.bx.numpy(),
           "nx": m.nx.numpy(), "geom": geom.numpy(), "inv_post_rots": inv_post_rots.numpy(),
           "combine": combine.numpy(), "coords": gf[:, :3].numpy(), "kept": kept.numpy(),
           "ranks": ranks.numpy(), "sorts": sorts.numpy(), "last_mask": last.numpy()}

    def fwd_bwd(dtype):
        torch.set_default_dtype(dtype)
        try:
            depth = torch.from_numpy(ft["depth"]).to(dtype).requires_grad_(True)
            feat = torch.from_numpy(ft["feat"]).to(dtype).requires_grad_(True)
            new_x = depth.unsqueeze(1) * feat.unsqueeze(2)                       # modules.py:84
            x = new_x.view(cfg.B, cfg.N, cfg.C, cfg.D, cfg.fH, cfg.fW).permute(0, 1, 3, 4, 5, 2)
            out = m.voxel_pooling(geom, x)                                       # reference splat
            if want_grads:
                out.backward(torch.from_numpy(dbev).to(dtype))
                return out.detach().numpy(), depth.grad.numpy(), feat.grad.numpy()
            return out.detach().numpy(), None, None
        finally:
            torch.set_default_dtype(torch.float32)

    res["bev32"], res["d_depth32"], res["d_feat32"] = fwd_bwd(torch.float32)
    res["bev64"], res["d_depth64"], res["d_feat64"] = fwd_bwd(torch.float64)
    return res


def reference_indices(cfg, cal):
    """Index tensors of the reference at full size (no features): its own get_geometry and the statements of
    voxel_pooling up to the argsort (model_baseline.py:92-110) and QuickCumsum's boundary mask (tools.py:196-197)."""
    m = ref_import.build_lss(cfg.B, cfg.grid_conf(), cfg.data_aug_conf())
    t = {k: torch.from_numpy(v) for k, v in cal.items()}
    with torch.no_grad():
        geom = m.get_geometry(t["rots"], t["trans"], t["intrins"], t["post_rots"], t["post_trans"])
        B, Nprime = cfg.B, cfg.P
        gf = ((geom - (m.bx - m.dx / 2.)) / m.dx).long().view(Nprime, 3)
        batch_ix = torch.cat([torch.full([Nprime // B, 1], ix, dtype=torch.long) for ix in range(B)])
        gf = torch.cat((gf, batch_ix), 1)
        kept = (gf[:, 0] >= 0) & (gf[:, 0] < m.nx[0]) & (gf[:, 1] >= 0) & (gf[:, 1] < m.nx[1]) \
            & (gf[:, 2] >= 0) & (gf[:, 2] < m.nx[2])
        gk = gf[kept]
        ranks = gk[:, 0] * (m.nx[1] * m.nx[2] * B) + gk[:, 1] * (m.nx[2] * B) + gk[:, 2] * B + gk[:, 3]
        sorts = ranks.argsort()
        sr = ranks[sorts]
        last = torch.ones(sr.shape[0], dtype=torch.bool)
        last[:-1] = sr[1:] != sr[:-1]
    return {"geom": geom.numpy(), "coords": gf[:, :3].numpy(), "kept": kept.numpy(), "ranks": ranks.numpy(),
            "sorts": sorts.numpy(), "last_mask": last.numpy(), "gk": gk.numpy(),
            "inv_post_rots": torch.inverse(t["post_rots"]).numpy(),
            "combine": t["rots"].matmul(torch.inverse(t["intrins"])).numpy()}


def large_fixture(name):
    """config4 / config5 at their full BASELINE.json batch: SHA-256 digests of every index tensor of the
    reference, plus float64 values at sampled outputs.  The frustum tensor of these shapes (1.3 GB / 8.2 GB,
    ten copies inside the reference's voxel_pooling) does not fit this container, so the sampled values are
    evaluated from the REFERENCE's index tensors (kept, coords, sorts, intervals) with the defining sums in
    float64: bev[v] = sum over the interval of depth*feat, d_depth[p] = <dbev[cell p], feat[pixel p]>,
    d_feat[pixel] = sum_d depth*dbev[cell].  The upstream gradient is synthetic.hash_field (evaluable per element)."""
    cfg = S.config(name)
    cal, ft = S.make_calibration(cfg), S.make_features(cfg)
    r = reference_indices(cfg, cal)
    K, V = int(len(r["ranks"])), int(r["last_mask"].sum())
    X, Y, Z = cfg.nx
    B, N, D, HW, C = cfg.B, cfg.N, cfg.D, cfg.fH * cfg.fW, cfg.C
    CZ = C * Z
    depth = ft["depth"].reshape(-1).astype(np.float64)           # indexed by point id
    feat = ft["feat"].reshape(B * N, C, HW)
    kept_idx = np.nonzero(r["kept"])[0]
    sorted_pts = kept_idx[r["sorts"]]                             # point ids in the reference's sorted order
    ends = np.nonzero(r["last_mask"])[0] + 1
    starts = np.concatenate(([0], ends[:-1]))
    gk_sorted = r["gk"][r["sorts"]]                               # (K, 4): x, y, z, b per sorted point

    def pixel_of(p):
        bn = p // (D * HW)
        return bn, p % HW

    def dbev_at(b, ch, x, y):
        return S.hash_field_np(((b * CZ + ch) * X + x) * Y + y).astype(np.float64)

    rs = np.random.RandomState(23)
    # ---- BEV: sampled occupied voxels (one random channel each) + empty cells
    iv = rs.choice(V, 4096, replace=False)
    cs = rs.randint(0, C, size=4096)
    bev_pick, bev_at = [], []
    for i, c in zip(iv, cs):
        pts = sorted_pts[starts[i]:ends[i]]
        bn, hw = pixel_of(pts)
        val = float(np.sum(depth[pts] * feat[bn, c, hw].astype(np.float64)))
        x, y, z, b = (int(v) for v in gk_sorted[starts[i]])
        bev_pick.append([b, z * C + int(c), x, y]); bev_at.append(val)
    occ = np.zeros((B, Z, X, Y), dtype=bool)
    occ[gk_sorted[starts, 3], gk_sorted[starts, 2], gk_sorted[starts, 0], gk_sorted[starts, 1]] = True
    empty = np.argwhere(~occ)
    for b, z, x, y in empty[rs.choice(len(empty), 512, replace=False)]:
        bev_pick.append([int(b), int(z) * C + int(rs.randint(0, C)), int(x), int(y)]); bev_at.append(0.0)
    # ---- d_depth at sampled points, d_feat at sampled (camera, channel, pixel)
    coords = r["coords"]
    dpick = rs.choice(cfg.P, 4096, replace=False)
    dd = []
    for p in dpick:
        if not r["kept"][p]:
            dd.append(0.0); continue
        x, y, z = (int(v) for v in coords[p]); b = int(p // (cfg.P // B))
        bn, hw = pixel_of(int(p))
        g = dbev_at(b, z * C + np.arange(C), x, y)
        dd.append(float(np.sum(g * feat[bn, :, hw].astype(np.float64))))
    fpick = rs.choice(B * N * C * HW, 4096, replace=False)
    df = []
    for e in fpick:
        bn, rem = divmod(int(e), C * HW); c, hw = divmod(rem, HW)
        pts = (bn * D + np.arange(D)) * HW + hw
        k = r["kept"][pts]
        b = bn // N
        x, y, z = coords[pts, 0], coords[pts, 1], coords[pts, 2]
        g = np.where(k, dbev_at(b, np.where(k, z, 0) * C + c, np.where(k, x, 0), np.where(k, y, 0)), 0.0)
        df.append(float(np.sum(depth[pts] * g)))
    doc = {"config": name, "seed": 1234, "B": B, "P": int(cfg.P), "K": K, "V": V,
           "sha256": {"inputs": sha(ft["depth"]) + sha(ft["feat"]),
                      "calibration": "".join(sha(cal[k]) for k in sorted(cal)),
                      "inv_post_rots": sha(r["inv_post_rots"]), "combine": sha(r["combine"]),
                      "geom": sha(r["geom"]), "coords_i32": sha(coords.astype(np.int32)),
                      "kept_u8": sha(r["kept"].astype(np.uint8)), "ranks_i32": sha(r["ranks"].astype(np.int32)),
                      "sorts_i32": sha(r["sorts"].astype(np.int32)),
                      "last_mask_u8": sha(r["last_mask"].astype(np.uint8))},
           "bev_pick": bev_pick, "bev64_at": bev_at,
           "d_depth_pick": [int(v) for v in dpick], "d_depth64_at": dd,
           "d_feat_pick": [int(v) for v in fpick], "d_feat64_at": df}
    with open(os.path.join(OUT, name + ".json"), "w") as f:
        json.dump(doc, f)
    print("%s: B=%d P=%d K=%d V=%d" % (name, B, cfg.P, K, V))


def inputs(cfg, seed=1234):
    return S.make_calibration(cfg, seed), S.make_features(cfg, seed), S.make_dbev(cfg, seed)


def main():
    os.makedirs(OUT, exist_ok=True)
    if "--large" in sys.argv:            # only the full-size config4 / config5 fixtures
        for name in ("config4", "config5"):
            large_fixture(name)
        return

    # ---- tiny: everything, full
    cfg = S.config("tiny")
    cal, ft, dbev = inputs(cfg)
    r = run_reference(cfg, cal, ft, dbev)
    np.savez_compressed(os.path.join(OUT, "tiny.npz"), **cal, **ft, dbev=dbev, **r)
    print("tiny: K=%d V=%d" % (len(r["ranks"]), int(r["last_mask"].sum())))

    # ---- edge cases on the tiny shape
    def edge(name, cal_e):
        rr = run_reference(cfg, cal_e, ft, dbev)
        np.savez_compressed(os.path.join(OUT, "edge_%s.npz" % name), **cal_e, **ft, dbev=dbev, **rr)
        print("edge_%s: K=%d V=%d" % (name, len(rr["ranks"]), int(rr["last_mask"].sum())))

    far = {k: v.copy() for k, v in cal.items()}
    far["trans"] = far["trans"] + np.float32(1000.0)           # everything out of range: K = 0
    edge("none_kept", far)
    one = {k: v.copy() for k, v in cal.items()}
    one["rots"] = np.zeros_like(one["rots"])                    # all points collapse onto trans
    one["trans"] = np.zeros_like(one["trans"]) + np.float32(0.25)
    edge("one_voxel", one)
    rs = np.random.RandomState(99)
    rnd = {k: rs.standard_normal(v.shape).astype(np.float32) for k, v in cal.items()}   # test_model()-style
    edge("randn_calib", rnd)
    bad = {k: v.copy() for k, v in rnd.items()}
    bad["trans"][0, 0, 0] = np.nan
    bad["trans"][0, 1, 1] = np.inf
    bad["trans"][1, 0, 2] = -np.inf
    bad["intrins"][1, 1] = 0.0                                   # singular -> inf/nan inverse
    bad["post_trans"][1, 2, 0] = 3.0e38
    try:
        edge("nonfinite", bad)
    except Exception as e:  # torch.inverse raises on exactly-singular input
        print("edge_nonfinite with singular intrinsics rejected by reference (%s); dropping that part" % type(e).__name__)
        bad["intrins"][1, 1] = rnd["intrins"][1, 1]
        edge("nonfinite", bad)

    # ---- config1 (B=1): full index tensors, sampled values
    cfg = S.config("config1")
    cal, ft, dbev = inputs(cfg)
    r = run_reference(cfg, cal, ft, dbev)
    rs = np.random.RandomState(7)
    nz = np.argwhere(r["bev64"] != 0)
    pick = nz[rs.choice(len(nz), 20000, replace=False)]
    zero_pick = np.argwhere(r["bev64"] == 0)[rs.choice(int((r["bev64"] == 0).sum()), 2000, replace=False)]
    pick = np.concatenate([pick, zero_pick]).astype(np.int32)
    np.savez_compressed(
        os.path.join(OUT, "config1.npz"), **cal,
        inv_post_rots=r["inv_post_rots"], combine=r["combine"],
        coords=r["coords"].astype(np.int32), kept=np.packbits(r["kept"]),
        ranks=r["ranks"].astype(np.int32), sorts=r["sorts"].astype(np.int32),
        last_mask=np.packbits(r["last_mask"]),
        geom_sha=np.array(sha(r["geom"])), bev32_sha=np.array(sha(r["bev32"])),
        bev_pick=pick, bev64_at=r["bev64"][tuple(pick.T)], bev32_at=r["bev32"][tuple(pick.T)],
        d_depth64=r["d_depth64"].astype(np.float64), d_feat64_sub=r["d_feat64"][:, ::8],
        inputs_sha=np.array(sha(ft["depth"]) + sha(ft["feat"]) + sha(dbev)))
    print("config1: K=%d V=%d" % (len(r["ranks"]), int(r["last_mask"].sum())))

    # ---- config2 (B=8): digests + samples
    cfg = S.config("config2")
    cal, ft, dbev = inputs(cfg)
    r = run_reference(cfg, cal, ft, dbev)
    rs = np.random.RandomState(11)
    nz = np.argwhere(r["bev64"] != 0)
    pick = nz[rs.choice(len(nz), 4096, replace=False)]
    dpick = rs.choice(r["d_depth64"].size, 4096, replace=False)
    fpick = rs.choice(r["d_feat64"].size, 4096, replace=False)
    doc = {
        "config": "config2", "seed": 1234, "P": int(cfg.P), "K": int(len(r["ranks"])),
        "V": int(r["last_mask"].sum()),
        "sha256": {
            "inputs": sha(ft["depth"]) + sha(ft["feat"]) + sha(dbev),
            "calibration": "".join(sha(cal[k]) for k in sorted(cal)),
            "inv_post_rots": sha(r["inv_post_rots"]), "combine": sha(r["combine"]),
            "geom": sha(r["geom"]), "coords_i32": sha(r["coords"].astype(np.int32)),
            "kept_u8": sha(r["kept"].astype(np.uint8)), "ranks_i32": sha(r["ranks"].astype(np.int32)),
            "sorts_i32": sha(r["sorts"].astype(np.int32)),
            "last_mask_u8": sha(r["last_mask"].astype(np.uint8)), "bev32": sha(r["bev32"]),
        },
        "bev_nonzero": int((r["bev64"] != 0).sum()),
        "bev64_sum": float(r["bev64"].sum()), "bev64_abs_sum": float(np.abs(r["bev64"]).sum()),
        "ref_fp32_vs_fp64_maxabs": float(np.abs(r["bev32"] - r["bev64"]).max()),
        "bev_pick": pick.tolist(), "bev64_at": r["bev64"][tuple(pick.T)].tolist(),
        "d_depth_pick": dpick.tolist(), "d_depth64_at": r["d_depth64"].ravel()[dpick].tolist(),
        "d_feat_pick": fpick.tolist(), "d_feat64_at": r["d_feat64"].ravel()[fpick].tolist(),
    }
    with open(os.path.join(OUT, "config2.json"), "w") as f:
        json.dump(doc, f)
    print("config2: K=%d V=%d" % (doc["K"], doc["V"]))

    # ---- torch.inverse / linspace / arange known answers
    rs = np.random.RandomState(5)
    mats = np.concatenate([rs.standard_normal((256, 3, 3)).astype(np.float32),
                           cal["post_rots"].reshape(-1, 3, 3), cal["intrins"].reshape(-1, 3, 3)])
    np.savez_compressed(os.path.join(OUT, "inverse3x3.npz"), A=mats,
                        inv=torch.inverse(torch.from_numpy(mats)).numpy())
    lin = {}
    for end, steps in [(351, 22), (127, 8), (703, 44), (255, 16), (1599, 100), (95, 6), (63, 4)]:
        lin["lin_%d_%d" % (end, steps)] = torch.linspace(0, end, steps, dtype=torch.float).numpy()
    for lo, hi, st in [(4.0, 45.0, 1.0), (1.0, 60.0, 1.0), (1.0, 60.0, 0.5), (4.0, 20.0, 1.0)]:
        lin["ar_%g_%g_%g" % (lo, hi, st)] = torch.arange(lo, hi, st, dtype=torch.float).numpy()
    np.savez_compressed(os.path.join(OUT, "linspace.npz"), **lin)
    print("done ->", OUT)


if __name__ == "__main__":
    main()
