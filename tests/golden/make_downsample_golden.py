"""Generates tests/golden/downsample.json with an INDEPENDENT numpy restatement of the reference's voxel-grid down-sampling
(tools.hpp:201-302 down_sampling_voxel / down_sampling_close, voxel_map.hpp:23-64 down_sampling_pvec).  Every arithmetic step is rounded
to the type the reference computes in (float for PointType fields, double for pointVar).  The reference itself cannot be executed here
(C++/Eigen/PCL), so these vectors pin both the oracle and the CUDA kernels against a second implementation: points on cell faces,
negative coordinates (the `loc < 0` branch), exact duplicates, cells with one point and cells with dozens."""
import json
import os

import numpy as np

f32, f64 = np.float32, np.float64


def key(p, vs):
    k = []
    for j in range(3):
        loc = f32(f64(p[j]) / f64(vs))
        if loc < 0:
            loc = f32(f64(loc) - 1.0)
        k.append(int(np.trunc(loc)))
    return tuple(k)


def cells_of(xyz, vs):
    cells = {}
    for i in range(xyz.shape[0]):
        cells.setdefault(key(xyz[i], vs), []).append(i)
    return cells


def ds_voxel(pts, vs):
    out = {}
    for ids in cells_of(pts, vs).values():
        x = [f32(v) for v in pts[ids[0]]]; cnt = f32(1)
        for i in ids[1:]:
            x = [f32(f32(f32(x[j] * cnt) + pts[i, j]) / f32(cnt + f32(1))) for j in range(3)]
            cnt = f32(cnt + f32(1))
        out[ids[0]] = ([float(v) for v in x], float(cnt))
    return out


def ds_close(pts, vs):
    out = {}
    for ids in cells_of(pts, vs).values():
        c = [f32(v) for v in pts[ids[0]]]
        for i in ids[1:]:
            c = [f32(c[j] + pts[i, j]) for j in range(3)]
        c = [f32(c[j] / f32(len(ids))) for j in range(3)]
        nd, best = 100.0, 0
        for t, i in enumerate(ids):
            d = [f64(f32(c[j] - pts[i, j])) for j in range(3)]
            dis = d[0] * d[0] + d[1] * d[1] + d[2] * d[2]
            if dis < nd:
                best, nd = t, dis
        out[ids[best]] = ([float(v) for v in pts[ids[best]]], float(len(ids)))
    return out


def ds_pvec(pv, vs):
    out = {}
    for ids in cells_of(pv[:, :3], vs).values():
        m, cnt = pv[ids[0]].copy(), 1
        for i in ids[1:]:
            m = (m * cnt + pv[i]) / (cnt + 1); cnt += 1
        out[ids[0]] = ([float(f32(v)) for v in m[:3]], [float(f32(v)) for v in m[[3, 7, 11]]], float(cnt))
    return out


def main():
    rng = np.random.default_rng(20260923)
    n = 240
    pts = rng.uniform(-3.0, 3.0, (n, 3)).astype(np.float32)
    pts[:40] = np.round(pts[:40] * 2) / 2                     # on cell faces, incl. 0 and negatives
    pts[40:80] = pts[:40]                                    # exact duplicates
    pts[80:140] = (pts[80:140] * 0.05).astype(np.float32)   # a crowd around the origin: cells with dozens of points, all four sign octants
    pv = np.zeros((n, 12))
    pv[:, :3] = pts.astype(np.float64) + rng.uniform(-1e-4, 1e-4, (n, 3))
    a = rng.uniform(-1e-2, 1e-2, (n, 3, 3))
    pv[:, 3:] = (a @ a.transpose(0, 2, 1)).reshape(n, 9)
    doc = dict(points_f32=[[float.hex(float(v)) for v in p] for p in pts], pvec_f64=[[float.hex(float(v)) for v in p] for p in pv], cases=[])
    for vs in (0.5, 0.125):
        v, c, q = ds_voxel(pts, vs), ds_close(pts, vs), ds_pvec(pv, vs)
        doc["cases"].append(dict(voxel_size=vs,
                                 voxel={str(k): dict(xyz=[float.hex(x) for x in a_], count=b_) for k, (a_, b_) in v.items()},
                                 close={str(k): dict(xyz=[float.hex(x) for x in a_], count=b_) for k, (a_, b_) in c.items()},
                                 pvec={str(k): dict(xyz=[float.hex(x) for x in a_], var_diag=[float.hex(x) for x in b_], count=c_) for k, (a_, b_, c_) in q.items()}))
    out = os.path.join(os.path.dirname(os.path.abspath(__file__)), "downsample.json")
    json.dump(doc, open(out, "w"))
    print(n, "points,", [len(c["voxel"]) for c in doc["cases"]], "cells ->", out)


if __name__ == "__main__":
    main()
