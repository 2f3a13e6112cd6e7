"""Seeded synthetic windows for the tests (3-plane corner room).  Uses the harness generator and the
ORACLE voxel map to obtain the factor (clusters, cached eig, sums) the reference would hand to the solver."""
import numpy as np

import oracle_api as oa
import synth
import voxel_slam_b200 as vx


def poses_true_est(W, L, seed, rot_sigma=2e-3, pos_sigma=1e-2):
    tr = np.stack([synth.true_pose(L, i) for i in range(W)])
    est = tr.copy()
    for i in range(1, W):
        est[i] = synth.perturb_pose(tr[i], seed * 1000 + i, rot_sigma, pos_sigma)
    return tr, est


def make_points(W, pts_per_scan, L, seed, poses_true, dtype=np.float64):
    pts = np.empty((W * pts_per_scan, 3), dtype=dtype)
    for i in range(W):
        synth.gen_scan(L, i, pts_per_scan, poses_true[i], seed=0x5EED0000 + seed, dtype=dtype, out=pts[i * pts_per_scan:(i + 1) * pts_per_scan])
    off = np.arange(W + 1, dtype=np.int64) * pts_per_scan
    return pts, off


def make_window(W=5, pts_per_scan=4000, L=6.0, seed=1, max_layer=2, voxel_size=1.0, threads=1, cut_with="est"):
    """Returns the dense factor the reference's map would build at the estimated poses (from-scratch build)."""
    tr, est = poses_true_est(W, L, seed)
    pts, off = make_points(W, pts_per_scan, L, seed, tr)
    mp = vx.MapParams.make(voxel_size=voxel_size, max_layer=max_layer)
    poses_cut = est if cut_with == "est" else tr
    of = oa.build_window_factor(mp, pts, off, poses_cut, threads=threads)
    ex = of.export()
    return dict(W=W, L=L, poses_true=tr, poses_est=est, pts=pts, offsets=off, mp=mp, oracle_factor=of, **ex)


def states_from_poses(poses12, vel=(0.5, 0.3, 0.0), g=(0.0, 0.0, -9.8)):
    W = poses12.shape[0]
    s = np.zeros((W, 24))
    s[:, :12] = poses12
    s[:, 12:15] = vel
    s[:, 21:24] = g
    return s


# ---------------------------------------------------------------- golden vectors of the down-sampling rows (tests/golden/downsample.json)
def load_downsample_golden():
    import json
    import os
    doc = json.load(open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "downsample.json")))
    pts = np.array([[float.fromhex(v) for v in p] for p in doc["points_f32"]], dtype=np.float32)
    pv = np.array([[float.fromhex(v) for v in p] for p in doc["pvec_f64"]], dtype=np.float64)
    return pts, pv, doc["cases"]


def check_downsample_against_golden(run_voxel, run_close, run_pvec):
    """run_*(array, voxel_size) -> dict(xyz, count, index[, var_diag]) from the implementation under test (oracle or CUDA); compared
    bit-exactly with the committed vectors (cells identified by the index they report: first point / picked point)."""
    pts, pv, cases = load_downsample_golden()
    for case in cases:
        vs = case["voxel_size"]
        for name, out, src in (("voxel", run_voxel(pts, vs), pts), ("close", run_close(pts, vs), pts), ("pvec", run_pvec(pv, vs), pv)):
            gold = case[name]
            assert len(out["index"]) == len(gold), (name, vs)
            for t, i in enumerate(out["index"].tolist()):
                g = gold[str(i)]
                want = np.array([float.fromhex(v) for v in g["xyz"]], dtype=np.float32)
                assert np.array_equal(out["xyz"][t].view(np.uint32), want.view(np.uint32)), (name, vs, i)
                assert out["count"][t] == g["count"], (name, vs, i)
                if name == "pvec":
                    wv = np.array([float.fromhex(v) for v in g["var_diag"]], dtype=np.float32)
                    assert np.array_equal(out["var_diag"][t].view(np.uint32), wv.view(np.uint32)), (name, vs, i)
