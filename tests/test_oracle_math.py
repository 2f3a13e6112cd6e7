"""CPU checks that pin the oracle (the reference ships no tests — 'parity unpinned'): numpy/scipy cross-checks, golden
vectors from an independent implementation, finite differences and metamorphic properties."""
import json
import os

import numpy as np
import pytest
from scipy.spatial.transform import Rotation

import oracle_api as oa
import scenes
import synth
import voxel_slam_b200 as vx

HERE = os.path.dirname(os.path.abspath(__file__))


def sym_from6(s):
    return np.array([[s[0], s[1], s[2]], [s[1], s[3], s[4]], [s[2], s[4], s[5]]])


def eig_cases():
    rng = np.random.default_rng(1)
    cases = []
    for _ in range(200):
        A = rng.standard_normal((3, 3)); cases.append(A + A.T)
    for _ in range(200):   # thin plane far from the origin: cov = P/N - c c^T with |c| ~ 100, lambda0 ~ 1e-4
        n = rng.standard_normal(3); n /= np.linalg.norm(n)
        Q, _ = np.linalg.qr(np.column_stack([n, rng.standard_normal((3, 2))]))
        pts = (Q[:, 1:] @ rng.uniform(-0.5, 0.5, (2, 300))).T + 0.01 * rng.standard_normal((300, 1)) * Q[:, 0] + rng.uniform(-100, 100, 3)
        c = oa.cluster_from_points(pts)
        N = c[9]; ctr = c[6:9] / N
        cases.append(sym_from6(c[:6]) / N - np.outer(ctr, ctr))
    cases += [np.diag([3.0, 1.0, 2.0]), np.eye(3), np.zeros((3, 3)), np.diag([1e-12, 1.0, 1.0]), np.array([[2.0, 1, 0], [1, 2, 0], [0, 0, 3.0]])]
    return cases


@pytest.mark.parametrize("impl", ["oracle", "hostmath"])
def test_eig3_against_numpy(impl):
    worst = 0.0
    for A in eig_cases():
        if impl == "oracle":
            w, U = oa.eig3(A)
        else:
            w, U = oa.hostmath_eig3([A[0, 0], A[0, 1], A[0, 2], A[1, 1], A[1, 2], A[2, 2]])
        wr, _ = np.linalg.eigh(A)
        scale = max(np.max(np.abs(wr)), 1e-300)
        assert np.all(np.diff(w) >= 0)
        assert np.max(np.abs(w - wr)) <= 4e-15 * scale + 1e-300
        assert np.max(np.abs(U.T @ U - np.eye(3))) < 1e-14
        assert np.max(np.abs(A @ U - U * w)) <= 1e-14 * scale + 1e-300      # columns are eigenvectors
        worst = max(worst, np.max(np.abs(w - wr)) / scale)
    assert worst < 4e-15


def test_voxel_keys_golden_bit_exact():
    cases = json.load(open(os.path.join(HERE, "golden", "voxel_keys.json")))
    assert len(cases) >= 200
    for c in cases:
        p = np.array([float.fromhex(h) for h in c["p"]])
        xyz, h = oa.voxel_keys(p[None, :], c["voxel_size"])
        assert xyz[0].tolist() == c["key"], (c, xyz)
        assert int(h[0]) == int(c["hash"])


def test_voxel_key_traps():
    """exact negative integers land one cell lower than floor(); -0.0 stays in cell 0."""
    xyz, _ = oa.voxel_keys(np.array([[-2.0, -0.0, 0.0], [-1.0, -1e-9, 0.999999]]), 1.0)
    assert xyz.tolist() == [[-3, 0, 0], [-2, -1, 0]]
    xyz, _ = oa.voxel_keys(np.array([[16777217.0, -16777217.0, 1.0]]), 1.0)   # beyond 2^24 the float cast rounds
    assert xyz.tolist() == [[16777216, -16777216, 1]]   # float(-16777217)-1 ties back to -2^24


def test_cluster_transform_equals_accumulate_after_transform():
    rng = np.random.default_rng(3)
    for hostmath in (False, True):
        for _ in range(20):
            pts = rng.uniform(-30, 30, (50, 3))
            R = Rotation.from_rotvec(rng.standard_normal(3)).as_matrix()
            t = rng.uniform(-50, 50, 3)
            pose = np.concatenate([R.ravel(), t])
            a = oa.cluster_transform(oa.cluster_from_points(pts), pose, hostmath=hostmath)
            b = oa.cluster_from_points(pts @ R.T + t)
            assert np.max(np.abs(a - b) / (np.abs(b) + 1.0)) < 1e-12


def test_so3_exp():
    rng = np.random.default_rng(4)
    for _ in range(50):
        w = rng.standard_normal(3) * rng.choice([1e-8, 1e-3, 1.0, 3.0])
        R = oa.so3_exp(w)
        assert np.max(np.abs(R - Rotation.from_rotvec(w).as_matrix())) < 1e-14
    assert np.array_equal(oa.so3_exp([1e-12, 0, 0]), np.eye(3))    # tools.hpp:55 threshold 1e-11


def test_ldlt_solve_matches_numpy():
    rng = np.random.default_rng(5)
    for n in (6, 30, 61, 150):
        B = rng.standard_normal((n, n))
        A = B @ B.T + 0.1 * np.eye(n)
        A[:6, :] = 0; A[:, :6] = 0; A[:6, :6] = np.eye(6)          # gauge-fixed shape of the LM system
        b = rng.standard_normal(n)
        x, rc = oa.ldlt_solve(A, b)
        assert rc == 0 and np.max(np.abs(x - np.linalg.solve(A, b))) < 1e-9 * np.max(np.abs(x))
    # mildly indefinite (true second derivative can be): still solved
    A = np.diag(np.r_[np.ones(5) * 3, -0.5 * np.ones(2)]) + 0.1 * np.ones((7, 7))
    b = np.arange(7.0)
    x, rc = oa.ldlt_solve(A, b)
    assert np.max(np.abs(A @ x - b)) < 1e-12


def _retract(poses, d):
    out = poses.copy()
    for i in range(poses.shape[0]):
        R = poses[i, :9].reshape(3, 3) @ Rotation.from_rotvec(d[6 * i:6 * i + 3]).as_matrix()
        out[i, :9] = R.ravel(); out[i, 9:] = poses[i, 9:] + d[6 * i + 3:6 * i + 6]
    return out


def test_gradient_and_hessian_against_finite_differences():
    """acc_evaluate2 (voxel_map.hpp:132-241) is the exact gradient / second derivative of sum lambda0 under the right
    perturbation R Exp(d_theta), t + d_t (voxel_map.hpp:407-408)."""
    sc = scenes.make_window(W=3, pts_per_scan=1500, L=4.0, seed=21)
    of, x0, W = sc["oracle_factor"], sc["poses_est"], 3
    f0 = of.residual(x0)             # refreshes the cached eig / sums at x0
    H, J, r = of.hessian(x0)
    assert abs(r - f0) < 1e-15
    n = 6 * W
    cost = lambda d: of.residual(_retract(x0, d))
    h = 1e-5
    g_fd = np.array([(cost(h * e) - cost(-h * e)) / (2 * h) for e in np.eye(n)])
    assert np.max(np.abs(g_fd - J)) < 2e-6 * np.max(np.abs(J))
    h = 2e-4
    rng = np.random.default_rng(0)
    pairs = [(a, a) for a in range(n)] + [tuple(rng.integers(0, n, 2)) for _ in range(40)]
    for a, b in pairs:
        ea, eb = np.eye(n)[a] * h, np.eye(n)[b] * h
        fd = (cost(ea + eb) - cost(ea - eb) - cost(-ea + eb) + cost(-ea - eb)) / (4 * h * h)
        sym = 0.5 * (H[a, b] + H[b, a])
        assert abs(fd - sym) < 2e-5 * np.max(np.abs(H)), (a, b, fd, sym)
    of.residual(x0)


def test_rank3_form_of_the_gpu_math_equals_the_reference_form():
    """vxs_math.cuh (what the CUDA kernels execute) compiled for the host vs acc_evaluate2 as written."""
    sc = scenes.make_window(W=6, pts_per_scan=3000, L=6.0, seed=8)
    of = sc["oracle_factor"]
    H, J, _ = of.hessian(sc["poses_est"])
    Hh, Jh, _ = of.hessian(sc["poses_est"], hostmath=True)
    assert np.max(np.abs(Hh - H)) < 1e-12 * np.max(np.abs(H)) and np.max(np.abs(Jh - J)) < 1e-12 * np.max(np.abs(J))


def test_residual_invariant_under_global_rigid_motion():
    sc = scenes.make_window(W=4, pts_per_scan=2000, L=5.0, seed=6)
    of, x = sc["oracle_factor"], sc["poses_est"]
    r0 = of.residual(x)
    Rg = Rotation.from_rotvec([0.3, -0.2, 0.5]).as_matrix(); tg = np.array([5.0, -3.0, 2.0])
    y = x.copy()
    for i in range(x.shape[0]):
        y[i, :9] = (Rg @ x[i, :9].reshape(3, 3)).ravel(); y[i, 9:] = Rg @ x[i, 9:] + tg
    # the fix clusters are zero in a from-scratch build, so the cost only depends on relative geometry
    assert np.all(sc["fix10"] == 0)
    assert abs(of.residual(y) - r0) < 1e-9 * r0
    of.residual(x)


def test_voxel_order_does_not_matter():
    sc = scenes.make_window(W=4, pts_per_scan=2000, L=5.0, seed=6)
    H, J, r = sc["oracle_factor"].hessian(sc["poses_est"])
    perm = np.random.default_rng(1).permutation(sc["eig12"].shape[0])
    of2 = oa.OracleFactor.from_dense(4, sc["clusters10"][perm], sc["fix10"][perm], None, sc["eig12"][perm], sc["sum10"][perm])
    H2, J2, r2 = of2.hessian(sc["poses_est"])
    assert np.max(np.abs(H2 - H)) < 1e-12 * np.max(np.abs(H)) and abs(r2 - r) < 1e-13 * r


def test_lidar_ba_converges_towards_truth():
    sc = scenes.make_window(W=6, pts_per_scan=5000, L=6.0, seed=12)
    out = sc["oracle_factor"].lidar_ba(sc["poses_est"], max_iter=6)
    tr = out["trace"]
    assert tr[0]["accepted"] == 1 and out["resis"][1] < 0.8 * out["resis"][0]
    assert abs(tr[0]["q1"] - (tr[0]["r1"] - tr[0]["r2"])) < 0.05 * tr[0]["q1"]          # quadratic model predicts the decrease
    before, after = np.abs(sc["poses_est"] - sc["poses_true"]).max(), np.abs(out["poses"] - sc["poses_true"]).max()
    assert after < 0.3 * before
    assert np.array_equal(out["poses"][0], sc["poses_est"][0])                           # gauge: pose 0 fixed (voxel_map.hpp:397-400)


@pytest.mark.parametrize("gravity", [False, True])
def test_li_ba_oracle_runs_and_decreases(gravity):
    W = 6
    sc = scenes.make_window(W=W, pts_per_scan=4000, L=6.0, seed=14)
    st = scenes.states_from_poses(sc["poses_est"])
    imu = synth.ImuWindow(sc["poses_true"])
    out = sc["oracle_factor"].li_ba(st, imu, with_gravity=gravity, max_iter=3)
    tr = out["trace"]
    assert len(tr) >= 1 and tr[0]["accepted"] == 1 and tr[-1]["r2"] < tr[0]["r1"]
    n = 15 * W + (3 if gravity else 0)
    assert out["hess"].shape == (n, n) and np.isfinite(out["hess"]).all()
    # the raw Hessian block the loop-closure queue reads (voxelslam.cpp:1657): block (0, DIM) diagonal is non-zero
    assert np.all(np.abs(np.diag(out["hess"][0:6, 15:21])) > 0)
