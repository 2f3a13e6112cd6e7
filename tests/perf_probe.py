"""Scratch probe (not a test): per-tile-kind cost of k_syrk at the metric shape.  Usage on the GPU box: python tests/perf_probe.py"""
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CODE = r'''
import sys, os, numpy as np
sys.path.insert(0, %r); sys.path.insert(0, os.path.join(%r, "tests"))
import synth, voxel_slam_b200 as vx
W, pts, L = 50, 200000, 130.0
tr = np.stack([synth.true_pose(L, i) for i in range(W)])
p = np.empty((W * pts, 3))
for i in range(W):
    synth.gen_scan(L, i, pts, tr[i], seed=0x5EED0000 + 1, out=p[i * pts:(i + 1) * pts])
off = np.arange(W + 1, dtype=np.int64) * pts
ctx = vx.Context(0)
f = vx.Factor(ctx, W)
ctx.build_window_factor(vx.MapParams.make(), p, off, tr, f)
for _ in range(3):
    ctx.evaluate_hessian(f, tr)
ctx.timing(True); ctx.timing_reset()
for _ in range(5):
    ctx.evaluate_hessian(f, tr)
st = ctx.timing_read()
print(os.environ.get("VXS_SYRK_ONLY_TILE", "all"), os.environ.get("VXS_SYRK_STREAMK", "1"), f.counts()[:2], "k_syrk ms", st["k_syrk"][0] / 5)
''' % (ROOT, ROOT)
for tile in ["all"] + [str(t) for t in range(10)]:
    env = dict(os.environ, VXS_SYRK_STREAMK="0")
    if tile != "all":
        env["VXS_SYRK_ONLY_TILE"] = tile
    r = subprocess.run([sys.executable, "-c", CODE], env=env, capture_output=True, text=True)
    print(r.stdout.strip() or r.stderr[-500:])
r = subprocess.run([sys.executable, "-c", CODE], env=dict(os.environ, VXS_SYRK_STREAMK="1"), capture_output=True, text=True)
print(r.stdout.strip() or r.stderr[-500:])
