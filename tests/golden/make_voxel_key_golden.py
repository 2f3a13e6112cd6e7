"""Generates tests/golden/voxel_keys.json with an INDEPENDENT numpy restatement of the reference's quantisation + hash
(voxel_map.hpp:1511-1518, tools.hpp:39-48).  The reference itself cannot be executed here (C++/Eigen/ROS), so these
vectors pin the oracle against a second implementation, including the adversarial cases of SURVEY.md §8c(4):
exact negative integers (land one cell lower than floor), +-0, values beyond 2^24 (float resolution), tiny negatives."""
import json
import os

import numpy as np

HASH_P, MAX_N, M64 = 116101, 10000000000, (1 << 64) - 1


def key_1d(x, vs):
    loc = np.float32(np.float64(x) / np.float64(vs))      # float loc = double / double
    if loc < 0:
        loc = np.float32(loc - np.float32(1.0))            # loc -= 1 in float
    return int(np.trunc(loc))                              # (int64_t)loc truncates toward zero


def vhash(x, y, z):
    ux, uy, uz = x & M64, y & M64, z & M64                 # hash<int64_t> is the identity on the bit pattern
    return (((((uz * HASH_P) & M64) % MAX_N + uy) & M64) * HASH_P & M64) % MAX_N + ux & M64


def main():
    cases = []
    base = [-2.0, -1.0, -0.0, 0.0, -1e-9, 0.999999, 1.0, 16777217.0, -16777217.0, 0.37, -0.37, 123.456, -123.456, 3.9999999, -3.0000001, 1e9]
    for vs in [0.25, 0.3, 1.0, 2.0, 4.0, 15.0]:
        for i, a in enumerate(base):
            p = [a * vs if abs(a) < 1e8 else a, base[(i * 7 + 3) % len(base)] * vs if abs(base[(i * 7 + 3) % len(base)]) < 1e8 else 5.5, base[(i * 5 + 1) % len(base)]]
            k = [key_1d(c, vs) for c in p]
            cases.append(dict(voxel_size=vs, p=[float.hex(float(c)) for c in p], key=k, hash=str(vhash(*k))))
    rng = np.random.default_rng(20260923)
    for vs in [0.5, 1.0, 2.0]:
        for p in rng.uniform(-300, 300, size=(40, 3)):
            k = [key_1d(c, vs) for c in p]
            cases.append(dict(voxel_size=vs, p=[float.hex(float(c)) for c in p], key=k, hash=str(vhash(*k))))
    out = os.path.join(os.path.dirname(os.path.abspath(__file__)), "voxel_keys.json")
    json.dump(cases, open(out, "w"), indent=0)
    print(len(cases), "cases ->", out)


if __name__ == "__main__":
    main()
