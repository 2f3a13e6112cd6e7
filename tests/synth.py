"""Synthetic-workload harness for the tests and bench.py (NOT product code): seeded scene generation and the stand-in for
the reference's unchanged IMU_PRE objects (preintegration.hpp — out of scope, stays on the CPU behind vxs_imu_hooks).
ctypes wrapper around tests/harness/_build/libvxs_harness.so."""
import ctypes as C
import os

import numpy as np

from voxel_slam_b200.api import ImuHooks, VxsError, _dp, _f64

HARNESS_PATH = os.path.join(os.path.dirname(os.path.abspath(__file__)), "harness", "_build", "libvxs_harness.so")
_harness = None


def harness():
    global _harness
    if _harness is None:
        if not os.path.exists(HARNESS_PATH):
            raise VxsError(-100, f"{HARNESS_PATH} is missing — run __graft_entry__.build()")
        _harness = C.CDLL(HARNESS_PATH)
        _harness.vxh_imu_create.restype = C.c_void_p
    return _harness



def true_pose(L, i):
    out = np.zeros(12)
    harness().vxh_true_pose(C.c_double(L), C.c_int(i), _dp(out))
    return out


def perturb_pose(pose12, seed, rot_sigma, pos_sigma):
    out = np.zeros(12)
    p = _f64(pose12)
    harness().vxh_perturb_pose(_dp(p), C.c_uint64(seed), C.c_double(rot_sigma), C.c_double(pos_sigma), _dp(out))
    return out


def gen_scan(L, frame, n, pose12_true, seed=0x5EED0000, off=0.37, sigma=0.01, max_range=0.0, dtype=np.float64, out=None):
    p = _f64(pose12_true)
    if dtype == np.float32:
        xyz = np.empty((n, 3), dtype=np.float32) if out is None else out
        harness().vxh_gen_scan_f32(C.c_double(L), C.c_double(off), C.c_double(sigma), C.c_double(max_range), C.c_uint64(seed), C.c_int(frame), C.c_int64(n), _dp(p),
                                   xyz.ctypes.data_as(C.POINTER(C.c_float)))
    else:
        xyz = np.empty((n, 3), dtype=np.float64) if out is None else out
        harness().vxh_gen_scan(C.c_double(L), C.c_double(off), C.c_double(sigma), C.c_double(max_range), C.c_uint64(seed), C.c_int(frame), C.c_int64(n), _dp(p), _dp(xyz))
    return xyz


def lawnmower_pose(i, per_row, step=4.0, row_gap=8.0, off=0.37):
    out = np.zeros(12)
    harness().vxh_lawnmower_pose(C.c_int(i), C.c_int(per_row), C.c_double(step), C.c_double(row_gap), C.c_double(off), _dp(out))
    return out


def gen_scan_city(frame, n, pose12_true, seed=0xC17E, G=10.0, rng_m=20.0, off=0.37, sigma=0.01, out=None):
    p = _f64(pose12_true)
    xyz = np.empty((n, 3), dtype=np.float32) if out is None else out
    harness().vxh_gen_scan_city_f32(C.c_uint64(seed), C.c_int(frame), C.c_int64(n), _dp(p), C.c_double(G), C.c_double(rng_m), C.c_double(off), C.c_double(sigma),
                                    xyz.ctypes.data_as(C.POINTER(C.c_float)))
    return xyz


class ImuWindow:
    """W-1 synthetic IMU preintegration factors (stand-in for the reference's unchanged IMU_PRE objects)."""

    def __init__(self, poses12_true, T=0.1, samples=20, gyr_noise=1e-3, acc_noise=1e-2, seed=7):
        p = _f64(poses12_true).reshape(-1, 12)
        self.W = p.shape[0]
        self._h = C.c_void_p(harness().vxh_imu_create(_dp(p), C.c_int(self.W), C.c_double(T), C.c_int(samples), C.c_double(gyr_noise), C.c_double(acc_noise), C.c_uint64(seed)))
        self.hooks = ImuHooks()
        harness().vxh_imu_hooks(self._h, C.byref(self.hooks))

    def reset(self):
        harness().vxh_imu_reset(self._h)

    def eval(self, states24, with_gravity=False, want_jac=True):
        s = _f64(states24)
        bs = 33 if with_gravity else 30
        blocks = np.zeros((self.W - 1, bs * bs))
        gvec = np.zeros((self.W - 1, bs))
        cost = C.c_double(0)
        harness().vxh_imu_eval(self._h, _dp(s), C.c_int(self.W), C.c_int(int(with_gravity)), C.c_int(int(want_jac)), _dp(blocks), _dp(gvec), C.byref(cost))
        return cost.value, blocks, gvec

    def __del__(self):
        try:
            if self._h:
                harness().vxh_imu_destroy(self._h)
                self._h = None
        except Exception:
            pass


