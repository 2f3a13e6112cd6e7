"""Scratch performance probe (not a test, not bench.py): builds a window with the ORACLE map, pushes it to the GPU and
prints per-kernel CUDA-event timings.  Usage: python tests/perf_probe.py W pts_per_scan L [reps]"""
import sys
import time

import numpy as np

sys.path.insert(0, ".")
sys.path.insert(0, "tests")
import oracle_api as oa  # noqa: E402
import scenes  # noqa: E402
import synth
import voxel_slam_b200 as vx  # noqa: E402

W, pts, L = int(sys.argv[1]), int(sys.argv[2]), float(sys.argv[3])
reps = int(sys.argv[4]) if len(sys.argv) > 4 else 5
t0 = time.time()
sc = scenes.make_window(W=W, pts_per_scan=pts, L=L, seed=1, threads=8)
V = sc["eig12"].shape[0]
E = int((sc["clusters10"][:, :, 9] > 0).sum())
print(f"scene W={W} pts/scan={pts} L={L}: V={V} E={E} k_avg={E / V:.1f}  (gen+oracle map {time.time() - t0:.1f}s, cut+recut {sc['oracle_factor'].build_seconds:.2f}s)", flush=True)
ctx = vx.Context(0)
f = vx.Factor(ctx, W)
t0 = time.time()
f.push_voxels_dense(sc["clusters10"], sc["eig12"], sc["sum10"])
print(f"push {time.time() - t0:.3f}s", flush=True)
st = scenes.states_from_poses(sc["poses_est"])
imu = synth.ImuWindow(sc["poses_true"])
for name, fn in [("residual", lambda: ctx.evaluate_residual(f, sc["poses_est"])),
                 ("hessian", lambda: ctx.evaluate_hessian(f, sc["poses_est"])),
                 ("lidar_ba_1it", lambda: ctx.lidar_ba(f, sc["poses_est"], max_iter=1, want_hess=False)),
                 ("li_ba_1it", lambda: (imu.reset(), ctx.li_ba(f, st, imu, max_iter=1, want_hess=False)))]:
    fn(); fn()
    ctx.timing(True); ctx.timing_reset()
    t0 = time.time()
    for _ in range(reps):
        fn()
    wall = (time.time() - t0) / reps
    tm = ctx.timing_read(); ctx.timing(False)
    print(f"--- {name}: wall {wall * 1e3:.3f} ms/call")
    for k, (ms, calls) in sorted(tm.items(), key=lambda kv: -kv[1][0]):
        if calls:
            print(f"    {k:18s} {ms / reps:9.4f} ms/call  ({calls // reps} launches/call, {ms / calls * 1e3:8.1f} us each)")
nthr = 5
ts, _ = sc["oracle_factor"].time_hessian(sc["poses_est"], nthr, 1)
tr, _ = sc["oracle_factor"].time_residual(sc["poses_est"], nthr, 1)
print(f"CPU oracle ({nthr} threads): hessian {ts * 1e3:.1f} ms, residual {tr * 1e3:.1f} ms")
