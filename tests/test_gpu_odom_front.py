"""GPU parity of the scan pre-processing between down-sampling and the EKF: var_init / calcBodyVar and pvec_update
(voxelslam.hpp:163-214) against the oracle restatement, which tests/test_ref_pin.py pins against the reference's own functions; and the
device-resident hand-over var_init -> pvec_update -> vxs_map_push_scan without a host round trip."""
import numpy as np
import pytest

import oracle_api as oa
import synth
import voxel_slam_b200 as vx

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ctx():
    c = vx.Context(0)
    yield c
    c.close()


def scan_f32(n, seed, stride=12):
    rng = np.random.default_rng(seed)
    pts = np.zeros((n, stride), dtype=np.float32)
    pts[:, :3] = rng.uniform(-60, 60, (n, 3)).astype(np.float32)
    pts[0, :3] = (3.0, -2.0, 0.0)            # z == 0: the reference moves the point to z = 1e-4 (voxelslam.hpp:165)
    pts[1, :3] = (0.01, 0.02, 35.0)
    pts[2, :3] = (25.0, -25.0, 1e-3)
    return pts


def relrow(a, b):
    return float(np.max(np.abs(a - b) / np.max(np.abs(b), axis=1, keepdims=True)))


@pytest.mark.parametrize("n,stride", [(5000, 12), (100000, 3), (1, 12)])
def test_var_init_and_pvec_update(ctx, n, stride):
    pts = scan_f32(max(n, 3), 11 + n, stride)[:n]
    ext_R = oa.so3_exp(np.array([0.02, -0.01, 0.03])); ext_p = np.array([0.05, -0.02, 0.1])
    g = ctx.var_init(pts, ext_R, ext_p, 0.02, 0.05)
    o = oa.var_init(pts, ext_R, ext_p, 0.02, 0.05)
    assert np.max(np.abs(g[:, :3] - o[:, :3])) < 1e-13 and relrow(g[:, 3:], o[:, 3:]) < 1e-11
    pose = np.concatenate([oa.so3_exp(np.array([0.3, 0.1, -0.2])).ravel(), [5.0, -3.0, 1.0]])
    rng = np.random.default_rng(9)
    A = rng.standard_normal((3, 3)) * 1e-3; B = rng.standard_normal((3, 3)) * 1e-2
    rot_var, tsl_var = A @ A.T, B @ B.T
    po, wo = oa.pvec_update(o, pose, rot_var, tsl_var)
    u = ctx.pvec_update(o, pose, rot_var, tsl_var)
    assert np.array_equal(u["pv"][:, :3], o[:, :3]) and relrow(u["pv"][:, 3:], po[:, 3:]) < 1e-12
    assert np.array_equal(u["pwld"], wo)            # world points feed the bit-exact cell assignment: same operation order as the oracle
    # resident hand-over: var_init leaves the records on the device, pvec_update works on them in place
    ctx.var_init(pts, ext_R, ext_p, 0.02, 0.05, want_out=False)
    u2 = ctx.pvec_update(None, pose, rot_var, tsl_var, n=n)
    assert relrow(u2["pv"][:, 3:], ctx.pvec_update(g, pose, rot_var, tsl_var)["pv"][:, 3:]) < 1e-15


def test_resident_scan_into_the_map(ctx):
    """var_init -> pvec_update -> vxs_map_push_scan(pv12 = NULL): the map built from the resident scan equals the one built from the host copy"""
    L, Wn, pts_n = 6.0, 4, 4000
    mp = vx.MapParams.make(voxel_size=1.0, max_layer=2)
    maps = [vx.LocalMap(ctx, mp, Wn), vx.LocalMap(ctx, mp, Wn)]
    fs = [vx.Factor(ctx, Wn), vx.Factor(ctx, Wn)]
    ident = np.eye(3); zero = np.zeros(3)
    x_buf = []
    for i in range(Wn):
        pose = synth.true_pose(L, i)
        cloud = synth.gen_scan(L, i, pts_n, pose, seed=0x5EED0000 + 41, dtype=np.float32)
        x_buf.append(pose)
        pv_host = ctx.var_init(cloud, ident, zero, 0.02, 0.05)
        upd = ctx.pvec_update(pv_host, pose, np.eye(3) * 1e-6, np.eye(3) * 1e-4)
        maps[0].push_scan(upd["pv"], np.stack(x_buf), fs[0])
        ctx.var_init(cloud, ident, zero, 0.02, 0.05, want_out=False)
        ctx.pvec_update(None, pose, np.eye(3) * 1e-6, np.eye(3) * 1e-4, n=pts_n, want_pv=False, want_pwld=False)
        maps[1].push_scan(None, np.stack(x_buf), fs[1], n=pts_n)
    try:
        a, b = maps[0].leaves(), maps[1].leaves()
        assert len(a["layer"]) == len(b["layer"]) > 50
        # nodes are created in a scheduling-dependent order: compare leaf by leaf through the cube key
        ka = np.lexsort(np.concatenate([np.round(a["voxel_center"], 9), a["layer"][:, None]], axis=1).T)
        kb = np.lexsort(np.concatenate([np.round(b["voxel_center"], 9), b["layer"][:, None]], axis=1).T)
        for k in ("voxel_center", "layer", "is_plane", "pcr_add", "slots"):
            assert np.array_equal(a[k][ka], b[k][kb]), k
        assert fs[0].counts() == fs[1].counts() and fs[0].counts()[0] > 20
        m3 = vx.LocalMap(ctx, mp, Wn)
        with pytest.raises(vx.VxsError):
            m3.push_scan(None, x_buf[0][None, :], None, n=pts_n + 1)      # no resident scan of that size
        m3.close()
    finally:
        for m in maps: m.close()
        for f in fs: f.close()
