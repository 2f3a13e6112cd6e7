"""kernel-level timing of one vxs_hba_pass (1 GPU)"""
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
for _ in range(2): o = ctx.hba_pass(fine, fine, xyz, off, est)
ctx.timing(True); ctx.timing_reset()
t0 = time.perf_counter(); o = ctx.hba_pass(fine, fine, xyz, off, est); wall = (time.perf_counter() - t0) * 1e3
st = ctx.timing_read(); ctx.timing(False)
ks = sorted(((v[0], k, v[1]) for k, v in st.items() if v[1] > 0), reverse=True)
print(f"wall {wall:.1f} ms (with per-kernel events), phases {o['phase_ms']}, kernels {sum(k[0] for k in ks):.1f} ms, launches {sum(k[2] for k in ks):.0f}")
for t, k, c in ks[:26]: print(f"     {k:24s} {t:8.2f} ms  x{c:.0f}")
