"""Voxel-map side of the oracle: point-to-voxel assignment and cluster sums against an independent numpy recomputation,
plane logic invariants, global-BA map and the HBA window loop."""
import numpy as np
import pytest

import oracle_api as oa
import scenes
import synth
import voxel_slam_b200 as vx


def np_key(p, vs):
    loc = (p / vs).astype(np.float32)
    loc = np.where(loc < 0, (loc - np.float32(1.0)).astype(np.float32), loc)
    return np.trunc(loc).astype(np.int64)


def np_cell_path(pw, key, vs, layer):
    """Octant path of world points down to `layer` (voxel_map.hpp:1029-1040: child = 4[x>cx]+2[y>cy]+[z>cz]; float quarter length)."""
    center = (0.5 + key) * vs
    q = np.float32(vs / 4.0)
    path = np.zeros(pw.shape[0], dtype=np.int64)
    for _ in range(layer):
        b = (pw > center).astype(np.int64)
        path = path * 8 + 4 * b[:, 0] + 2 * b[:, 1] + b[:, 2]
        center = center + ((2 * b - 1).astype(np.float32) * q).astype(np.float64)
        q = np.float32(q / np.float32(2.0))
    return path


def brute_force_clusters(sc, poses):
    W = sc["W"]
    per = {}
    for i in range(W):
        pb = sc["pts"][sc["offsets"][i]:sc["offsets"][i + 1]]
        R, t = poses[i, :9].reshape(3, 3), poses[i, 9:]
        pw = pb @ R.T + t
        key = np_key(pw, sc["mp"].voxel_size)
        per[i] = (pb, pw, key)
    return per


def test_window_factor_assignment_and_sums():
    sc = scenes.make_window(W=4, pts_per_scan=6000, L=5.0, seed=31)
    per = brute_force_clusters(sc, sc["poses_est"])
    ids, cl, s = sc["ids"], sc["clusters10"], sc["sum10"]
    assert len(ids) > 50 and set(np.unique(ids["layer"])) <= {0, 1, 2} and (ids["layer"] > 0).any()
    seen = set()
    for v in range(len(ids)):
        vid = ids[v]
        tag = (int(vid["x"]), int(vid["y"]), int(vid["z"]), int(vid["layer"]), int(vid["path"]))
        assert tag not in seen
        seen.add(tag)
        tot = np.zeros(10)
        for i in range(sc["W"]):
            pb, pw, key = per[i]
            m = np.all(key == np.array(tag[:3]), axis=1)
            if vid["layer"] > 0 and m.any():
                m[m] = np_cell_path(pw[m], np.array(tag[:3], dtype=np.float64), sc["mp"].voxel_size, int(vid["layer"])) == tag[4]
            c = oa.cluster_from_points(pb[m]) if m.any() else np.zeros(10)
            assert c[9] == cl[v, i, 9], (tag, i)                                  # bit-exact point-to-voxel assignment
            assert np.max(np.abs(c - cl[v, i]) / (np.abs(c) + 1e-9)) < 1e-12
            if m.any():
                tot += oa.cluster_from_points(pw[m])
        assert tot[9] == s[v, 9] and np.max(np.abs(tot - s[v]) / (np.abs(tot) + 1e-9)) < 1e-11
        lam = sc["eig12"][v, :3]
        assert lam[0] < sc["mp"].min_eigen_value and lam[0] / lam[2] < 0.25 and lam[0] / lam[1] <= 0.12 and s[v, 9] > 5


def test_threaded_recut_gives_the_same_factor():
    a = scenes.make_window(W=4, pts_per_scan=4000, L=5.0, seed=32, threads=1)
    b = scenes.make_window(W=4, pts_per_scan=4000, L=5.0, seed=32, threads=5)
    assert np.array_equal(a["ids"], b["ids"]) and np.array_equal(a["clusters10"], b["clusters10"]) and np.array_equal(a["eig12"], b["eig12"])


def test_max_layer_zero_pins_leaf_equals_root():
    sc = scenes.make_window(W=3, pts_per_scan=4000, L=5.0, seed=33, max_layer=0)
    assert np.all(sc["ids"]["layer"] == 0) and np.all(sc["ids"]["path"] == 0)


def make_gba(W=6, pts=3000, L=8.0, seed=41):
    tr, est = scenes.poses_true_est(W, L, seed, rot_sigma=3e-3, pos_sigma=2e-2)
    xyz, off = scenes.make_points(W, pts, L, seed, tr, dtype=np.float32)
    return tr, est, xyz, off


def test_gba_factor_rules():
    tr, est, xyz, off = make_gba()
    mp = vx.MapParams.make(voxel_size=2.0, min_eigen_value=0.1, max_layer=2)
    of = oa.build_gba_factor(mp, xyz, off, est, threads=2)
    ex = of.export()
    assert of.size() > 10
    assert np.all(ex["sum10"][:, 9] > 10)                                   # N <= 10 dead (loop_refine.hpp:360)
    assert np.all((ex["clusters10"][:, :, 9] > 0).sum(axis=1) >= 2)          # >= 2 observing frames (loop_refine.hpp:372-376)
    assert np.all(ex["fix10"] == 0) and np.all(ex["coe"] == 1.0)
    lam = ex["eig12"][:, :3]
    assert np.all(lam[:, 0] / lam[:, 1] <= 0.12) and np.all(lam[:, 0] < 0.1)
    assert np.array_equal(ex["clusters10"][:, :, 9].sum(axis=1), ex["sum10"][:, 9])
    of1 = oa.build_gba_factor(mp, xyz, off, est, threads=1)
    a, b = np.sort(ex["ids"], order=["x", "y", "z", "layer", "path"]), np.sort(of1.export()["ids"], order=["x", "y", "z", "layer", "path"])
    assert np.array_equal(a, b)


def test_hba_window_improves_poses():
    tr, est, xyz, off = make_gba(W=6, pts=4000, L=8.0, seed=42)
    coarse = vx.MapParams.make(voxel_size=2.0, min_eigen_value=0.1, max_layer=2)
    fine = vx.MapParams.make(voxel_size=1.0, min_eigen_value=0.0025, max_layer=2)
    out = oa.hba_window(coarse, fine, xyz, off, est, max_iter=4, thread_num=2)
    assert out["status"] == 0 and 1 <= out["outer_iters"] <= 4
    assert np.abs(out["poses"] - tr).max() < 0.5 * np.abs(est - tr).max()
    assert np.isfinite(out["hess"]).all() and np.abs(out["hess"]).max() > 0


def _ds_cloud(n, seed, stride=3):
    rng = np.random.default_rng(seed)
    pts = np.zeros((n, stride), dtype=np.float32)
    pts[:, :3] = (rng.uniform(-6.0, 6.0, size=(n, 3))).astype(np.float32)
    pts[: n // 10, :3] = np.round(pts[: n // 10, :3] * 2) / 2          # points exactly on cell faces, including 0 and negatives
    pts[n // 10: n // 5, :3] = pts[: n // 5 - n // 10, :3]               # exact duplicates
    if stride > 3:
        pts[:, 3:] = rng.uniform(size=(n, stride - 3)).astype(np.float32)
    return pts


def _ds_reference_python(pts, voxel_size, close):
    """tools.hpp:201-302 written out with numpy float32 scalars (every operation rounds to float, like PointType fields)."""
    f32 = np.float32
    cells = {}
    for i in range(pts.shape[0]):
        key = []
        for j in range(3):
            loc = f32(np.float64(pts[i, j]) / np.float64(voxel_size))
            if loc < 0:
                loc = f32(np.float64(loc) - 1.0)
            key.append(int(np.trunc(loc)))
        cells.setdefault(tuple(key), []).append(i)
    out = {}
    for key, ids in cells.items():
        if not close:
            x = [f32(v) for v in pts[ids[0], :3]]; cnt = f32(1)
            for i in ids[1:]:
                x = [f32(f32(f32(x[j] * cnt) + pts[i, j]) / f32(cnt + f32(1))) for j in range(3)]
                cnt = f32(cnt + f32(1))
            out[ids[0]] = (np.array(x, dtype=np.float32), float(cnt))
        else:
            c = [f32(v) for v in pts[ids[0], :3]]
            for i in ids[1:]:
                c = [f32(c[j] + pts[i, j]) for j in range(3)]
            c = [f32(c[j] / f32(len(ids))) for j in range(3)]
            nd, best = 100.0, 0
            for t, i in enumerate(ids):
                d = [np.float64(f32(c[j] - pts[i, j])) for j in range(3)]
                dis = d[0] * d[0] + d[1] * d[1] + d[2] * d[2]
                if dis < nd:
                    best, nd = t, dis
            out[ids[best]] = (pts[ids[best], :3].copy(), float(len(ids)))
    return out


@pytest.mark.parametrize("close", [False, True])
@pytest.mark.parametrize("stride", [3, 12])
def test_down_sampling_oracle_matches_float32_restatement(close, stride):
    pts = _ds_cloud(3000, 11 + stride, stride)
    for vs in (0.5, 0.125):
        o = oa.down_sampling(pts, vs, close=close, stride_floats=stride)
        ref = _ds_reference_python(pts, vs, close)
        assert len(o["index"]) == len(ref) and set(o["index"].tolist()) == set(ref.keys())
        for k, i in enumerate(o["index"].tolist()):
            assert np.array_equal(o["xyz"][k], ref[i][0]) and o["count"][k] == ref[i][1]
    assert oa.down_sampling(pts, 0.0005, close=close, stride_floats=stride) is None       # tools.hpp:203: cloud left untouched
    assert len(oa.down_sampling(pts[:0], 0.5, close=close, stride_floats=stride)["index"]) == 0


def test_submap_merge_oracle_consistency():
    """voxelslam.cpp:2428-2447: merged cloud = float(dR v + dp) in the frame of keyframe 0, then down_sampling_voxel."""
    rng = np.random.default_rng(21)
    W, per = 4, 800
    poses = np.stack([scenes.true_pose(12.0, i) for i in range(W)]) if hasattr(scenes, "true_pose") else None
    if poses is None:
        poses = np.stack([synth.true_pose(12.0, i) for i in range(W)])
    pts = rng.uniform(-8, 8, (W * per, 3)).astype(np.float32)
    off = np.arange(W + 1, dtype=np.int64) * per
    raw = oa.submap_merge(pts, off, poses, 0.0)          # < 0.001: merged cloud returned as it is
    R0, p0 = poses[0, :9].reshape(3, 3), poses[0, 9:]
    for i in range(W):
        Ri, pi = poses[i, :9].reshape(3, 3), poses[i, 9:]
        ref = (pts[off[i]:off[i + 1]].astype(np.float64) @ (R0.T @ Ri).T + R0.T @ (pi - p0))
        assert np.max(np.abs(raw["xyz"][off[i]:off[i + 1]] - ref)) < 2e-6
    o = oa.submap_merge(pts, off, poses, 0.25)
    d = oa.down_sampling(raw["xyz"], 0.25)
    assert np.array_equal(np.sort(o["index"]), np.sort(d["index"])) and int(o["count"].sum()) == W * per


def _pvec_cloud(n, seed):
    rng = np.random.default_rng(seed)
    pv = np.zeros((n, 12))
    pv[:, :3] = rng.uniform(-9.0, 9.0, (n, 3))
    pv[: n // 8, :3] = np.round(pv[: n // 8, :3] * 4) / 4            # on cell faces
    a = rng.uniform(-1e-2, 1e-2, (n, 3, 3))
    pv[:, 3:] = (a @ a.transpose(0, 2, 1)).reshape(n, 9)              # symmetric PSD covariances
    return pv


def test_down_sampling_pvec_oracle_matches_numpy():
    """voxel_map.hpp:23-64 restated with numpy float64 scalars in input order."""
    pv = _pvec_cloud(4000, 17)
    for vs in (0.5, 0.25):
        o = oa.down_sampling_pvec(pv, vs)
        keys = oa.voxel_keys(pv[:, :3], vs)[0]
        cells = {}
        for i, k in enumerate(map(tuple, keys.tolist())):
            cells.setdefault(k, []).append(i)
        assert len(o["index"]) == len(cells)
        pos = {i: t for t, i in enumerate(o["index"].tolist())}
        for ids in cells.values():
            m, cnt = pv[ids[0]].copy(), 1
            for i in ids[1:]:
                m = (m * cnt + pv[i]) / (cnt + 1); cnt += 1
            t = pos[ids[0]]
            assert np.array_equal(o["xyz"][t], m[:3].astype(np.float32)) and np.array_equal(o["var_diag"][t], m[[3, 7, 11]].astype(np.float32)) and o["count"][t] == cnt


def test_down_sampling_oracle_matches_golden_vectors():
    """tests/golden/downsample.json (independent numpy restatement, generator committed beside it) — bit-exact."""
    scenes.check_downsample_against_golden(lambda p, vs: oa.down_sampling(p, vs), lambda p, vs: oa.down_sampling(p, vs, close=True), lambda pv, vs: oa.down_sampling_pvec(pv, vs))
