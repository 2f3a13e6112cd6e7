"""CPU pins for the oracle of SURVEY §8f ranks 1 and 3 (plane_update / margi / match / odometry accumulation, voxel_map.hpp:1118-1392,
voxelslam.cpp:876-918).  There is no CUDA path for these rows yet; the oracle is pinned so that the next round can build against it."""
import numpy as np
import pytest

import oracle_api as oa
import scenes
import synth
import voxel_slam_b200 as vx

SIGMA = 0.01


@pytest.fixture(scope="module")
def world():
    W, L = 4, 6.0
    tr, est = scenes.poses_true_est(W, L, 5)
    pts, off = scenes.make_points(W, 6000, L, 5, tr)
    mp = vx.MapParams.make(voxel_size=1.0, max_layer=2)
    lm = oa.LocalMap(mp, pts, off, tr, SIGMA ** 2, mgsize=0)
    # world points at the poses the map was built with
    pw = np.concatenate([pts[off[i]:off[i + 1]] @ tr[i, :9].reshape(3, 3).T + tr[i, 9:] for i in range(W)])
    return dict(W=W, L=L, tr=tr, pts=pts, off=off, lm=lm, pw=pw, planes=lm.planes())


def _points_of(world, k):
    c, h = world["planes"]["voxel_center"][k], world["planes"]["half"][k]
    inside = np.all(np.abs(world["pw"] - c) < h * (1 - 1e-9), axis=1)
    return world["pw"][inside]


def test_plane_update_matches_monte_carlo(world):
    """plane_var (covariance of [normal | centre]) against the sample covariance of planes refitted to re-noised points."""
    P = world["planes"]
    assert len(P["N"]) > 20
    rng = np.random.default_rng(1)
    checked = 0
    for k in np.argsort(-P["N"])[:40]:
        pts = _points_of(world, k)
        if pts.shape[0] != int(P["N"][k]) or P["eig"][k, 1] < 20 * P["eig"][k, 0]:   # points on the cube faces / badly conditioned patches: skip
            continue
        n0, c0 = P["normal"][k], P["center"][k]
        assert np.allclose(c0, pts.mean(0), atol=1e-12)
        T = 3000
        noisy = pts[None] + SIGMA * rng.standard_normal((T,) + pts.shape)
        c = noisy.mean(1)
        d = noisy - c[:, None]
        cov = np.einsum("tni,tnj->tij", d, d) / pts.shape[0]
        w, U = np.linalg.eigh(cov)
        nrm = U[:, :, 0]
        nrm *= np.sign(nrm @ n0)[:, None]
        S = np.cov(np.concatenate([nrm, c], axis=1).T)
        V = P["plane_var"][k]
        # the refitted plane sees the original scatter PLUS the added noise, so the prediction is compared block-wise with 25 % slack
        for blk in (slice(0, 3), slice(3, 6)):
            assert abs(np.trace(S[blk, blk]) - np.trace(V[blk, blk])) < 0.25 * np.trace(V[blk, blk])
        assert np.linalg.norm(S - V) < 0.3 * np.linalg.norm(V)
        assert abs(P["radius"][k] - np.float32(P["eig"][k, 2])) == 0
        checked += 1
    assert checked >= 8


def test_margi_moves_scans_into_the_fixed_part(world):
    """After margi(mgsize=2) the two oldest scans live in pcr_fix: the plane table is unchanged (same points) and a second map built with
    those scans marginalised still matches the same points."""
    mp = vx.MapParams.make(voxel_size=1.0, max_layer=2)
    lm2 = oa.LocalMap(mp, world["pts"], world["off"], world["tr"], SIGMA ** 2, mgsize=2)
    a, b = world["planes"], lm2.planes()
    assert len(a["N"]) == len(b["N"])
    ia, ib = np.lexsort(a["voxel_center"].T), np.lexsort(b["voxel_center"].T)
    assert np.array_equal(a["voxel_center"][ia], b["voxel_center"][ib])
    assert np.allclose(a["center"][ia], b["center"][ib], atol=1e-12) and np.allclose(a["plane_var"][ia], b["plane_var"][ib], rtol=1e-9, atol=1e-18)


def _brute_force(world, pv, pose, rot_var, tsl_var):
    P = world["planes"]
    R, p = pose[:9].reshape(3, 3), pose[9:]
    HTH, HTz, nnt, flags = np.zeros((6, 6)), np.zeros(6), np.zeros((3, 3)), np.zeros(pv.shape[0], dtype=np.int32)
    f32 = np.float32
    for i in range(pv.shape[0]):
        x, var = pv[i, :3], pv[i, 3:].reshape(3, 3)
        phat = np.array([[0, -x[2], x[1]], [x[2], 0, -x[0]], [-x[1], x[0], 0]])
        vw = R @ var @ R.T + phat @ rot_var @ phat.T + tsl_var
        w = R @ x + p
        inside = np.where(np.all(np.abs(w - P["voxel_center"]) < P["half"][:, None], axis=1))[0]
        if inside.size != 1:
            continue
        k = inside[0]
        d = w - P["center"][k]
        dp = f32(abs(P["normal"][k] @ d)); dc = f32(d @ d)
        if f32(dc - f32(dp * dp)) > f32(9) * f32(P["radius"][k]):
            continue
        J = np.concatenate([d, -P["normal"][k]])
        sig = J @ P["plane_var"][k] @ J + P["normal"][k] @ vw @ P["normal"][k]
        if not (float(dp) < 3 * np.sqrt(sig)):
            continue
        rinv = 1.0 / (0.0005 + sig)
        res = P["normal"][k] @ d
        jac = np.concatenate([phat @ R.T @ P["normal"][k], P["normal"][k]])
        HTH += rinv * np.outer(jac, jac); HTz -= rinv * jac * res; nnt += np.outer(P["normal"][k], P["normal"][k]); flags[i] = 1
    return HTH, HTz, nnt, flags


def test_odom_accumulate_matches_brute_force(world):
    """voxelslam.cpp:876-918 against a numpy loop over the exported plane table (independent leaf search by cube containment)."""
    rng = np.random.default_rng(3)
    W, L = world["W"], world["L"]
    pose_true = synth.true_pose(L, W)                                  # the next scan of the trajectory
    body = synth.gen_scan(L, W, 1500, pose_true, seed=0x5EED0000 + 77)
    pose = synth.perturb_pose(pose_true, 99, 2e-3, 1e-2)
    var = np.tile((SIGMA ** 2 * np.eye(3)).reshape(1, 9), (body.shape[0], 1)) * rng.uniform(0.5, 2.0, (body.shape[0], 1))
    pv = np.concatenate([body, var], axis=1)
    rot_var, tsl_var = 1e-6 * np.eye(3), 1e-4 * np.eye(3)
    o = world["lm"].odom_accumulate(pv, pose, rot_var, tsl_var, passes=1)
    HTH, HTz, nnt, flags = _brute_force(world, pv, pose, rot_var, tsl_var)
    assert o["n"] == int(flags.sum()) > 300
    assert np.array_equal(o["flags"], flags)
    for a, b in ((o["HTH"], HTH), (o["HTz"], HTz), (o["nnt"], nnt)):
        assert np.max(np.abs(a - b)) < 1e-9 * np.max(np.abs(b))
    # the per-point leaf cache of the EKF loop (voxelslam.cpp:892-900) does not change the result
    o2 = world["lm"].odom_accumulate(pv, pose, rot_var, tsl_var, passes=3)
    assert o2["n"] == o["n"] and np.array_equal(o2["HTH"], o["HTH"]) and np.array_equal(o2["HTz"], o["HTz"])


# ---------------------------------------------------------------------------------------------------------------- sliding window
def _transform_cluster(c10, pose):
    """PointCluster::transform (tools.hpp:357-363) on the packed form Pxx Pxy Pxz Pyy Pyz Pzz vx vy vz N."""
    R, t = pose[:9].reshape(3, 3), pose[9:]
    P = np.array([[c10[0], c10[1], c10[2]], [c10[1], c10[3], c10[4]], [c10[2], c10[4], c10[5]]]); v = c10[6:9]; N = c10[9]
    Rv = R @ v
    Pw = R @ P @ R.T + np.outer(Rv, t) + np.outer(t, Rv) + N * np.outer(t, t)
    vw = Rv + N * t
    return np.array([Pw[0, 0], Pw[0, 1], Pw[0, 2], Pw[1, 1], Pw[1, 2], Pw[2, 2], vw[0], vw[1], vw[2], N])


@pytest.mark.parametrize("max_points", [10 ** 9, 100])
def test_sliding_window_bookkeeping(max_points):
    """voxelslam.cpp:1599-1686 (map side): after every scan, in every leaf, pcr_add == pcr_fix + sum over the window of the slot clusters
    moved to the world with the pose of their LOGICAL window position (the slot ring rotates on every slide), the slots beyond the
    window are empty, and — while nothing has been dropped at max_points — every point ever cut is in exactly one leaf."""
    W, L, n_scans, per = 5, 6.0, 13, 1500
    mp = vx.MapParams.make(voxel_size=1.0, max_layer=2)
    sim = oa.SlidingSim(mp, W, mgsize=1, max_points=max_points)
    all_world = []
    for k in range(n_scans):
        pose = synth.true_pose(L, k)
        body = synth.gen_scan(L, k, per, pose, seed=0x5EED0000 + 9)
        sim.add_scan(body, pose, var_diag=1e-4)
        all_world.append(body @ pose[:9].reshape(3, 3).T + pose[9:])
        st = sim.state()
        assert st["win_count"] == min(k + 1, W - 1) and st["win_base"] == max(0, k + 2 - W)
        assert sorted(st["ring"].tolist()) == list(range(W)) and st["ring"][0] == st["win_base"] % W
        live = st["has_sw"]
        assert live.sum() > 30
        # slots of logical positions outside the window are empty
        assert not st["slots"][:, st["win_count"]:, 9].any()
        # cluster bookkeeping (N exactly, first and second moments to rounding)
        for t in np.where(live)[0]:
            acc = st["pcr_fix"][t].copy()
            for i in range(st["win_count"]):
                if st["slots"][t, i, 9] != 0:
                    acc += _transform_cluster(st["slots"][t, i], st["poses"][i])
            assert acc[9] == st["pcr_add"][t, 9]
            assert np.max(np.abs(acc - st["pcr_add"][t])) <= 1e-9 * max(1.0, np.max(np.abs(acc)))
        if max_points >= 10 ** 9:
            # conservation: the leaves partition space and nothing was dropped
            pw = np.concatenate(all_world)
            total = 0
            for t in range(len(st["half"])):
                inside = np.all(np.abs(pw - st["voxel_center"][t]) < st["half"][t], axis=1).sum()
                assert inside == int(st["pcr_add"][t, 9]), (k, t, inside, st["pcr_add"][t, 9])
                total += inside
            assert total == pw.shape[0]
        else:
            assert np.all(st["pcr_fix"][:, 9] <= st["pcr_add"][:, 9])
