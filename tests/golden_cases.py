"""Input recipes shared by tests/golden/make_reference_golden.py (runs the reference's own code on them, in the build container) and
tests/test_oracle_reference_golden.py (runs the oracle on them, anywhere).  Everything is seeded and comes from the harness generators."""
import numpy as np

import scenes
import synth
import voxel_slam_b200 as vx


def states(poses12):
    return scenes.states_from_poses(poses12)


def window_case():
    W, pts, L = 4, 1500, 5.0
    tr, est = scenes.poses_true_est(W, L, 61)
    p, off = scenes.make_points(W, pts, L, 61, tr)
    return dict(W=W, mp=vx.MapParams.make(voxel_size=1.0, max_layer=2), pts=p, off=off, tr=tr, est=est)


def hba_case():
    W = 6
    tr, est = scenes.poses_true_est(W, 8.0, 62, rot_sigma=3e-3, pos_sigma=2e-2)
    xyz, off = scenes.make_points(W, 2500, 8.0, 62, tr, dtype=np.float32)
    return dict(W=W, coarse=vx.MapParams.make(voxel_size=2.0, min_eigen_value=0.1, max_layer=2), fine=vx.MapParams.make(voxel_size=1.0, min_eigen_value=0.0025, max_layer=2), xyz=xyz, off=off,
                tr=tr, est=est)


def lio_case():
    W, L = 4, 6.0
    tr, est = scenes.poses_true_est(W, L, 63)
    pts, off = scenes.make_points(W, 4000, L, 63, tr)
    scan = synth.gen_scan(L, W - 1, 2500, tr[W - 1], seed=0x5EED0000 + 63, sigma=0.05)
    pv = np.zeros((scan.shape[0], 12)); pv[:, :3] = scan; pv[:, [3, 7, 11]] = 1e-4
    st = np.zeros(24); st[:12] = synth.perturb_pose(tr[W - 1], 6300, 5e-3, 3e-2); st[12:15] = (0.3, -0.1, 0.05); st[21:24] = (0, 0, -9.8)
    cov = np.diag([1e-4] * 3 + [1e-3] * 3 + [1e-2] * 3 + [1e-6] * 6)
    return dict(mp=vx.MapParams.make(voxel_size=1.0, max_layer=2), pts=pts, off=off, tr=tr, pv=pv, state=st, cov=cov)


def pointvar_case():
    import oracle_api as oa
    rng = np.random.default_rng(64)
    n = 48
    pts = np.zeros((n, 12), dtype=np.float32)
    pts[:, :3] = rng.uniform(-40, 40, (n, 3)).astype(np.float32)
    pts[0, :3] = (3.0, -2.0, 0.0)
    A = rng.standard_normal((3, 3)) * 1e-3; B = rng.standard_normal((3, 3)) * 1e-2
    return dict(pts=pts, ext_R=oa.so3_exp(np.array([0.02, -0.01, 0.03])), ext_p=np.array([0.05, -0.02, 0.1]),
                pose=np.concatenate([oa.so3_exp(np.array([0.3, 0.1, -0.2])).ravel(), [5.0, -3.0, 1.0]]), rot_var=A @ A.T, tsl_var=B @ B.T)
