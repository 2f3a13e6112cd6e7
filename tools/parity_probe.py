"""GPU-vs-oracle first LM iteration at a bench-like shape, with per-frame / per-block differences (debug aid)."""
import os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import oracle_api as oa, synth, scenes, voxel_slam_b200 as vx
import bench

W, L = 50, 130.0
pts = int(sys.argv[1]) if len(sys.argv) > 1 else 60000
ctx = vx.Context(0)
tr, est, p, off = bench.scene_points(vx, W, pts, L, seed=1)
mp = vx.MapParams.make(voxel_size=1.0, min_eigen_value=0.0025, max_layer=2)
f = vx.Factor(ctx, W)
ctx.build_window_factor(mp, p, off, est, f)
ptr, fr, cl, fx, co = f.read_structure()
eig0, sum0 = f.read_back()
f.cache_save()
st0 = bench.states_from(est)
imu = synth.ImuWindow(tr)
imu.reset(); f.cache_restore()
g = ctx.li_ba(f, st0, imu, with_gravity=False, max_iter=1, want_hess=True, trace_cap=4)
of = bench.oracle_factor_from_csr(W, ptr, fr, cl, eig0, sum0)
imu2 = synth.ImuWindow(tr); imu2.reset()
o = of.li_ba(st0, imu2, with_gravity=False, max_iter=1)
print("V", ptr.shape[0] - 1, "trace gpu", g["trace"], "oracle", o["trace"])
a, b = g["states"], o["states"]
for nm, lo, hi in (("R", 0, 9), ("p", 9, 12), ("v", 12, 15), ("bg", 15, 18), ("ba", 18, 21)):
    d = np.abs(a[:, lo:hi] - b[:, lo:hi]).max(axis=1)
    print(nm, "inc", np.abs(b[:, lo:hi] - st0[:, lo:hi]).max(), "diff max", d.max(), "at frame", int(d.argmax()), "first frames", d[:4], "last", d[-3:])
H, Ho = g["hess"], o["hess"]
print("hess relinf", np.abs(H - Ho).max() / np.abs(Ho).max())
n = H.shape[0]
blk = np.abs(H - Ho).reshape(W, 15, W, 15).max(axis=(1, 3))
i, j = np.unravel_index(blk.argmax(), blk.shape)
print("worst block", i, j, blk[i, j], "block scale", np.abs(Ho.reshape(W, 15, W, 15)[i, :, j, :]).max())
# the solver on the oracle's system
gg = None
Hb, Hob = H.reshape(W, 15, W, 15), Ho.reshape(W, 15, W, 15)   # column-major n x n: index [col_frame, col_dof, row_frame, row_dof]
D = np.abs(Hb - Hob)
print("diag-block diff by dof group (max over frames): lidar 6x6", D[np.arange(W), :6, np.arange(W), :6].max(), " imu 9x9", D[np.arange(W), 6:, np.arange(W), 6:].max(),
      " cross", D[np.arange(W), :6, np.arange(W), 6:].max())
print("scale: lidar", np.abs(Hob[np.arange(W), :6, np.arange(W), :6]).max(), "imu", np.abs(Hob[np.arange(W), 6:, np.arange(W), 6:]).max())
# lidar-only Hessian of the full factor, same cached eig (restore first)
f.cache_restore()
Hl, Jl, rl = ctx.evaluate_hessian(f, est)
of2 = bench.oracle_factor_from_csr(W, ptr, fr, cl, eig0, sum0)
Hr, Jr, rr = of2.hessian(est)
print("lidar-only full factor: H relinf", np.abs(Hl - Hr).max() / np.abs(Hr).max(), "g relinf", np.abs(Jl - Jr).max() / np.abs(Jr).max(), "r", abs(rl - rr) / rr)
Db = np.abs(Hl - Hr).reshape(W, 6, W, 6).max(axis=(1, 3)); i, j = np.unravel_index(Db.argmax(), Db.shape)
print("worst lidar block", i, j, Db[i, j], "scale", np.abs(Hr.reshape(W, 6, W, 6)[i, :, j, :]).max())
print("diag blocks diff", Db[np.arange(W), np.arange(W)].max(), "offdiag", (Db - np.diag(np.diag(Db))).max())
# the IMU blocks the two runs were given
c1, b1, g1 = imu.eval(st0, False, True) if False else (None, None, None)
ia, ib = synth.ImuWindow(tr), synth.ImuWindow(tr)
ia.reset(); ib.reset()
ca, ba_, ga = ia.eval(st0, False, True); cb, bb_, gb = ib.eval(st0, False, True)
print("imu eval repeatability", abs(ca - cb), np.abs(ba_ - bb_).max(), "block scale", np.abs(ba_).max())
