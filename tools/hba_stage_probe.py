"""Per-kernel time vs wall time of the three phases of the hierarchical pass (debug / profiling aid)."""
import os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import synth, voxel_slam_b200 as vx
from voxel_slam_b200 import api
K, n, per_row = int(sys.argv[1]), int(sys.argv[2]), 50
ctx = vx.Context(0)
tr = np.stack([synth.lawnmower_pose(i, per_row) for i in range(K)])
est = np.stack([tr[0]] + [synth.perturb_pose(tr[i], 100 + i, 1e-3, 1e-2) for i in range(1, K)])
xyz = api.pinned_array((K * n, 3), np.float32)
for i in range(K): synth.gen_scan_city(i, n, tr[i], out=xyz[i * n:(i + 1) * n])
off = np.arange(K + 1, dtype=np.int64) * n
fine = vx.MapParams.make(voxel_size=1.0, min_eigen_value=0.0025, max_layer=2)
wf = np.arange(0, K - 9, 5, dtype=np.int32)
def phase(name, fn, reps=2):
    fn()
    ctx.timing(True); ctx.timing_reset()
    t0 = time.perf_counter()
    for _ in range(reps): out = fn()
    wall = (time.perf_counter() - t0) * 1e3 / reps
    st = ctx.timing_read(); ctx.timing(False)
    ks = sorted(((v[0] / reps, k, v[1] / reps) for k, v in st.items() if v[1] > 0), reverse=True)
    print(f"== {name}: wall {wall:.1f} ms, kernels {sum(k[0] for k in ks):.1f} ms, launches {sum(k[2] for k in ks):.0f}")
    for t, k, c in ks[:12]: print(f"     {k:24s} {t:8.2f} ms  x{c:.0f}")
    return out
b = phase("bottom", lambda: ctx.hba_bottom_batch(fine, xyz, off, est, wf, win_size=10, thread_num=2))
m = phase("merge", lambda: ctx.submap_merge_batch(xyz, off, b["poses"], wf, 0.125))
so = m["win_offsets"]
top = phase("top", lambda: ctx.hba_window(fine, fine, m["xyz"], so, est[wf], max_iter=1, thread_num=5))
print("submap points", len(m["xyz"]), "top resis", top["resis_log"])
