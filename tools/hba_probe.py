import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import oracle_api as oa, scenes, voxel_slam_b200 as vx
from test_gpu_hba_batch import trajectory, oracle_window
ctx = vx.Context(0)
K, ws = 25, 10
tr, est, xyz, off = trajectory(K, 3000, 8.0, 60 + K, 3)
fine = vx.MapParams.make(voxel_size=1.0, min_eigen_value=0.0025, max_layer=2)
ref = {k0: oracle_window(fine, xyz, off, est, k0, ws)[0] for k0 in (0, 5, 10)}
for wf in ([0], [5], [10], [0, 5], [5, 0], [0, 5, 10]):
    g = ctx.hba_bottom_batch(fine, xyz, off, est, np.array(wf, dtype=np.int32), win_size=ws, want_hess=True)
    for w, k0 in enumerate(wf):
        r = ref[k0]
        print(wf, "win", k0, "pose diff", np.max(np.abs(g["poses"][w] - r["poses"])), "resis", g["resis"][w], r["resis_log"][:2], "iters", g["lm_iters"][w], "H rel",
              np.max(np.abs(g["hess"][w] - r["hess"])) / np.max(np.abs(r["hess"])))
