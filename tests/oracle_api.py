"""ctypes wrapper around oracle/_build/liboracle.so — the CPU restatement of the reference (TEST INFRASTRUCTURE ONLY)."""
import ctypes as C
import os

import numpy as np

from voxel_slam_b200.api import LM_TRACE_DTYPE, VOXEL_ID_DTYPE, LmTrace, MapParams, VoxelId, _dp, _f64

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ORACLE_PATH = os.path.join(ROOT, "oracle", "_build", "liboracle.so")
REF_PATH = os.path.join(ROOT, "oracle", "_ref", "libvxref.so")
# Two libraries export the same entry points: the hand-written restatement (oracle/vxo_*.hpp, prefix vxo_) and the reference's OWN sources
# compiled against the Eigen / PCL / ROS stand-ins (oracle/ref_capi.cpp, prefix vxr_).  This module is executed once per backend:
# `import oracle_api` binds to the restatement, `import ref_api` re-executes it with BACKEND = "ref" (tests/ref_api.py).
BACKEND = globals().get("BACKEND", "oracle")
_lib = None


class _Prefixed:
    """CDLL view that maps the vxo_ names used below onto the backend's prefix."""

    def __init__(self, cdll, prefix):
        self._cdll, self._prefix = cdll, prefix

    def __getattr__(self, name):
        return getattr(self._cdll, self._prefix + name[4:] if name.startswith("vxo_") else name)


def available():
    return os.path.exists(REF_PATH if BACKEND == "ref" else ORACLE_PATH)


def lib():
    global _lib
    if _lib is None:
        _lib = _Prefixed(C.CDLL(REF_PATH), "vxr_") if BACKEND == "ref" else _Prefixed(C.CDLL(ORACLE_PATH), "vxo_")
        _lib.vxo_factor_create.restype = C.c_void_p
        _lib.vxo_build_window_factor.restype = C.c_void_p
        _lib.vxo_build_gba_factor.restype = C.c_void_p
        _lib.vxo_factor_size.restype = C.c_int64
        for fn in ("vxo_factor_residual", "vxo_factor_hessian", "vxo_time_hessian", "vxo_time_residual"):
            getattr(_lib, fn).restype = C.c_double
    return _lib


def eig3(A):
    A = _f64(A).reshape(3, 3)
    w, U = np.zeros(3), np.zeros((3, 3))
    lib().vxo_eig3(_dp(A), _dp(w), _dp(U))
    return w, U


def hostmath_eig3(sym6):
    s = _f64(sym6)
    w, U = np.zeros(3), np.zeros((3, 3))
    lib().vxo_hostmath_eig3(_dp(s), _dp(w), _dp(U))
    return w, U


def voxel_keys(pw, voxel_size):
    p = _f64(pw).reshape(-1, 3)
    n = p.shape[0]
    xyz = np.zeros((n, 3), dtype=np.int64)
    h = np.zeros(n, dtype=np.uint64)
    lib().vxo_voxel_keys(_dp(p), C.c_int64(n), C.c_double(voxel_size), xyz.ctypes.data_as(C.POINTER(C.c_int64)), h.ctypes.data_as(C.POINTER(C.c_uint64)))
    return xyz, h


def cluster_from_points(pts):
    p = _f64(pts).reshape(-1, 3)
    out = np.zeros(10)
    lib().vxo_cluster_from_points(_dp(p), C.c_int64(p.shape[0]), _dp(out))
    return out


def cluster_transform(c10, pose12, hostmath=False):
    c, p, out = _f64(c10), _f64(pose12), np.zeros(10)
    (lib().vxo_hostmath_transform if hostmath else lib().vxo_cluster_transform)(_dp(c), _dp(p), _dp(out))
    return out


def so3_exp(w):
    w = _f64(w)
    R = np.zeros((3, 3))
    lib().vxo_so3_exp(_dp(w), _dp(R))
    return R


def ldlt_solve(A, b):
    A = np.asfortranarray(A, dtype=np.float64)
    b = _f64(b)
    x = np.zeros_like(b)
    rc = lib().vxo_ldlt_solve(_dp(A), _dp(b), C.c_int(b.shape[0]), _dp(x))
    return x, rc


class OracleFactor:
    def __init__(self, handle, W):
        self._h = C.c_void_p(handle)
        self.W = W

    @staticmethod
    def from_dense(W, clusters10, fix10, coe, eig12, sum10):
        h = lib().vxo_factor_create(C.c_int(W))
        f = OracleFactor(h, W)
        cl, e, s = _f64(clusters10), _f64(eig12), _f64(sum10)
        fx = _f64(fix10) if fix10 is not None else None
        co = _f64(coe) if coe is not None else None
        n = e.reshape(-1, 12).shape[0]
        lib().vxo_factor_push_dense(f._h, C.c_int64(n), _dp(cl), _dp(fx), _dp(co), _dp(e), _dp(s))
        return f

    def __del__(self):
        try:
            if self._h:
                lib().vxo_factor_destroy(self._h)
                self._h = None
        except Exception:
            pass

    def size(self):
        return int(lib().vxo_factor_size(self._h))

    def export(self):
        n, W = self.size(), self.W
        cl, fx, co, e, s = np.zeros((n, W, 10)), np.zeros((n, 10)), np.zeros(n), np.zeros((n, 12)), np.zeros((n, 10))
        ids = np.zeros(n, dtype=VOXEL_ID_DTYPE)
        lib().vxo_factor_export(self._h, _dp(cl), _dp(fx), _dp(co), _dp(e), _dp(s), ids.ctypes.data_as(C.POINTER(VoxelId)))
        return dict(clusters10=cl, fix10=fx, coe=co, eig12=e, sum10=s, ids=ids)

    def residual(self, poses12):
        p = _f64(poses12)
        return float(lib().vxo_factor_residual(self._h, _dp(p)))

    def hessian(self, poses12, hostmath=False):
        p = _f64(poses12)
        n = 6 * self.W
        H, J = np.zeros((n, n), order="F"), np.zeros(n)
        if hostmath:
            lib().vxo_hostmath_hessian(self._h, _dp(p), _dp(H), _dp(J))
            return H, J, None
        r = lib().vxo_factor_hessian(self._h, _dp(p), _dp(H), _dp(J))
        return H, J, float(r)

    def time_hessian(self, poses12, threads, reps=1):
        p = _f64(poses12)
        r = C.c_double(0)
        return float(lib().vxo_time_hessian(self._h, _dp(p), C.c_int(threads), C.c_int(reps), C.byref(r))), r.value

    def time_residual(self, poses12, threads, reps=1):
        p = _f64(poses12)
        r = C.c_double(0)
        return float(lib().vxo_time_residual(self._h, _dp(p), C.c_int(threads), C.c_int(reps), C.byref(r))), r.value

    def lidar_ba(self, poses12, max_iter=3, thd_num=2, trace_cap=64):
        p = _f64(poses12).copy()
        n = 6 * self.W
        H, resis = np.zeros((n, n), order="F"), np.zeros(2)
        conv, tl = C.c_int(0), C.c_int(0)
        tr = np.zeros(trace_cap, dtype=LM_TRACE_DTYPE)
        rc = lib().vxo_lidar_ba(self._h, _dp(p), C.c_int(max_iter), C.c_int(thd_num), _dp(H), _dp(resis), C.byref(conv), tr.ctypes.data_as(C.POINTER(LmTrace)),
                                C.c_int(trace_cap), C.byref(tl))
        return dict(poses=p, hess=H, resis=resis, is_converge=bool(conv.value), trace=tr[: tl.value], status=rc)

    def li_ba(self, states24, imu, with_gravity=False, max_iter=3, imu_coef=1e-4, trace_cap=64):
        s = _f64(states24).copy()
        n = 15 * self.W + (3 if with_gravity else 0)
        H, resis = np.zeros((n, n), order="F"), np.zeros(2)
        tl = C.c_int(0)
        tr = np.zeros(trace_cap, dtype=LM_TRACE_DTYPE)
        # the restatement calls back through vxs_imu_hooks (harness stand-in); the reference build drives real IMU_PRE objects (RefImuWindow)
        imu_arg = imu._h if BACKEND == "ref" else C.byref(imu.hooks)
        rc = lib().vxo_li_ba(self._h, _dp(s), C.c_int(int(with_gravity)), C.c_int(max_iter), C.c_double(imu_coef), imu_arg, _dp(H), _dp(resis),
                             tr.ctypes.data_as(C.POINTER(LmTrace)), C.c_int(trace_cap), C.byref(tl))
        return dict(states=s, hess=H, resis=resis, trace=tr[: tl.value], status=rc)


def build_window_factor(mp, pts_body, scan_offsets, poses12, fix_pts=None, threads=1, var_diag=0.0):
    pts = _f64(pts_body).reshape(-1, 3)
    off = np.ascontiguousarray(scan_offsets, dtype=np.int64)
    W = off.shape[0] - 1
    p = _f64(poses12)
    fx = _f64(fix_pts).reshape(-1, 3) if fix_pts is not None else None
    secs = C.c_double(0)
    h = lib().vxo_build_window_factor(C.byref(mp), _dp(pts), off.ctypes.data_as(C.POINTER(C.c_int64)), _dp(p), C.c_int(W), _dp(fx), C.c_int64(0 if fx is None else fx.shape[0]),
                                      C.c_int(threads), C.c_double(var_diag), C.byref(secs))
    f = OracleFactor(h, W)
    f.build_seconds = secs.value
    return f


def build_gba_factor(mp, xyz_f32, kf_offsets, poses12, threads=2, stride_floats=3):
    x = np.ascontiguousarray(xyz_f32, dtype=np.float32)
    off = np.ascontiguousarray(kf_offsets, dtype=np.int64)
    W = off.shape[0] - 1
    p = _f64(poses12)
    secs = C.c_double(0)
    h = lib().vxo_build_gba_factor(C.byref(mp), x.ctypes.data_as(C.POINTER(C.c_float)), C.c_int(stride_floats), off.ctypes.data_as(C.POINTER(C.c_int64)), _dp(p), C.c_int(W),
                                   C.c_int(threads), C.byref(secs))
    f = OracleFactor(h, W)
    f.build_seconds = secs.value
    return f


def hba_edges(hess, W, poses12):
    cap = W * (W - 1) // 2
    H = np.asfortranarray(hess, dtype=np.float64); p = _f64(poses12)
    eij = np.zeros((cap, 2), dtype=np.int32); v6 = np.zeros((cap, 6)); rot = np.zeros((cap, 9)); tra = np.zeros((cap, 3))
    lib().vxo_hba_edges.restype = C.c_int64
    m = lib().vxo_hba_edges(_dp(H), C.c_int(W), _dp(p), C.c_int64(cap), eij.ctypes.data_as(C.POINTER(C.c_int32)), _dp(v6), _dp(rot), _dp(tra))
    return dict(n=m, ij=eij[:m], v6=v6[:m], rot=rot[:m], tra=tra[:m])


def down_sampling(pts_f32, voxel_size, close=False, stride_floats=None):
    x = np.ascontiguousarray(pts_f32, dtype=np.float32)
    stride = int(stride_floats) if stride_floats else (x.shape[1] if x.ndim == 2 else 3)
    n = x.size // stride
    xyz = np.zeros((max(n, 1), 3), dtype=np.float32); cnt = np.zeros(max(n, 1), dtype=np.float32); idx = np.zeros(max(n, 1), dtype=np.int64)
    lib().vxo_down_sampling.restype = C.c_int64
    m = lib().vxo_down_sampling(C.c_int(1 if close else 0), x.ctypes.data_as(C.POINTER(C.c_float)), C.c_int(stride), C.c_int64(n), C.c_double(voxel_size),
                                xyz.ctypes.data_as(C.POINTER(C.c_float)), cnt.ctypes.data_as(C.POINTER(C.c_float)), idx.ctypes.data_as(C.POINTER(C.c_int64)), C.c_int64(n))
    if m < 0:
        return None
    return dict(xyz=xyz[:m], count=cnt[:m], index=idx[:m])


def down_sampling_pvec(pv_f64, voxel_size):
    x = np.ascontiguousarray(pv_f64, dtype=np.float64)
    n, stride = x.shape
    xyz = np.zeros((max(n, 1), 3), dtype=np.float32); nrm = np.zeros((max(n, 1), 3), dtype=np.float32); cnt = np.zeros(max(n, 1), dtype=np.float32)
    idx = np.zeros(max(n, 1), dtype=np.int64)
    lib().vxo_down_sampling_pvec.restype = C.c_int64
    m = lib().vxo_down_sampling_pvec(_dp(x), C.c_int(stride), C.c_int64(n), C.c_double(voxel_size), xyz.ctypes.data_as(C.POINTER(C.c_float)),
                                     nrm.ctypes.data_as(C.POINTER(C.c_float)), cnt.ctypes.data_as(C.POINTER(C.c_float)), idx.ctypes.data_as(C.POINTER(C.c_int64)), C.c_int64(n))
    return dict(xyz=xyz[:m], var_diag=nrm[:m], count=cnt[:m], index=idx[:m])


def submap_merge(xyz_f32, kf_offsets, poses12, voxel_size, stride_floats=None):
    x = np.ascontiguousarray(xyz_f32, dtype=np.float32)
    stride = int(stride_floats) if stride_floats else (x.shape[1] if x.ndim == 2 else 3)
    off = np.ascontiguousarray(kf_offsets, dtype=np.int64)
    W = off.shape[0] - 1
    n = int(off[-1])
    p = _f64(poses12)
    xyz = np.zeros((max(n, 1), 3), dtype=np.float32); cnt = np.zeros(max(n, 1), dtype=np.float32); idx = np.zeros(max(n, 1), dtype=np.int64)
    lib().vxo_submap_merge.restype = C.c_int64
    m = lib().vxo_submap_merge(x.ctypes.data_as(C.POINTER(C.c_float)), C.c_int(stride), off.ctypes.data_as(C.POINTER(C.c_int64)), _dp(p), C.c_int(W), C.c_double(voxel_size),
                               xyz.ctypes.data_as(C.POINTER(C.c_float)), cnt.ctypes.data_as(C.POINTER(C.c_float)), idx.ctypes.data_as(C.POINTER(C.c_int64)), C.c_int64(n))
    return dict(xyz=xyz[:m], count=cnt[:m], index=idx[:m])


def hba_window(coarse, fine, xyz_f32, kf_offsets, poses12, max_iter, thread_num=2, stride_floats=3):
    x = np.ascontiguousarray(xyz_f32, dtype=np.float32)
    off = np.ascontiguousarray(kf_offsets, dtype=np.int64)
    W = off.shape[0] - 1
    p = _f64(poses12).copy()
    H = np.zeros((6 * W, 6 * W), order="F")
    log = np.zeros(2 * max(max_iter, 1))
    it = C.c_int(0)
    rc = lib().vxo_hba_window(C.byref(coarse), C.byref(fine), x.ctypes.data_as(C.POINTER(C.c_float)), C.c_int(stride_floats), off.ctypes.data_as(C.POINTER(C.c_int64)), _dp(p),
                              C.c_int(W), C.c_int(max_iter), C.c_int(thread_num), _dp(H), _dp(log), C.byref(it))
    return dict(poses=p, hess=H, resis_log=log[: 2 * it.value], outer_iters=it.value, status=rc)


def hba_add_edge(coarse, fine, xyz_f32, kf_offsets, poses12, max_iter, thread_num=2, stride_floats=3, want_submap=True):
    """ref backend only: the reference's own HBA_add_edge (voxelslam.cpp:2319-2482, cut out at build time) -> its PGO edges and merged submap."""
    if BACKEND != "ref":
        raise RuntimeError("hba_add_edge exists for the reference build only; the oracle exposes hba_window / hba_edges / submap_merge")
    x = np.ascontiguousarray(xyz_f32, dtype=np.float32)
    off = np.ascontiguousarray(kf_offsets, dtype=np.int64)
    W = off.shape[0] - 1
    n = int(off[-1])
    p = _f64(poses12)
    cap = W * (W - 1) // 2
    eij = np.zeros((cap, 2), dtype=np.int32); v6 = np.zeros((cap, 6)); rot = np.zeros((cap, 9)); tra = np.zeros((cap, 3))
    sub = np.zeros((max(n, 1), 3), dtype=np.float32)
    nsub = C.c_int64(0)
    lib().vxo_hba_add_edge.restype = C.c_int64
    m = lib().vxo_hba_add_edge(C.byref(coarse), C.byref(fine), x.ctypes.data_as(C.POINTER(C.c_float)), C.c_int(stride_floats), off.ctypes.data_as(C.POINTER(C.c_int64)), _dp(p), C.c_int(W),
                               C.c_int(max_iter), C.c_int(thread_num), C.c_int(int(want_submap)), C.c_int64(cap), eij.ctypes.data_as(C.POINTER(C.c_int32)), _dp(v6), _dp(rot), _dp(tra),
                               C.c_int64(n), sub.ctypes.data_as(C.POINTER(C.c_float)), C.byref(nsub))
    return dict(n=m, ij=eij[:m], v6=v6[:m], rot=rot[:m], tra=tra[:m], submap=sub[: nsub.value])


def var_init(pts_f32, ext_R, ext_p, dept_err, beam_err, stride_floats=None):
    """var_init (voxelslam.hpp:187-203) -> n x 12 pointVar records (pnt | var row-major)"""
    x = np.ascontiguousarray(pts_f32, dtype=np.float32)
    stride = int(stride_floats) if stride_floats else (x.shape[1] if x.ndim == 2 else 3)
    n = x.size // stride
    out = np.zeros((max(n, 1), 12))
    lib().vxo_var_init(x.ctypes.data_as(C.POINTER(C.c_float)), C.c_int(stride), C.c_int64(n), _dp(_f64(ext_R)), _dp(_f64(ext_p)), C.c_double(dept_err), C.c_double(beam_err), _dp(out))
    return out[:n]


def pvec_update(pv12, pose12, rot_var, tsl_var):
    """pvec_update (voxelslam.hpp:205-214) -> (pv with the world variance, pwld)"""
    pv = _f64(pv12).reshape(-1, 12).copy()
    pw = np.zeros((max(pv.shape[0], 1), 3))
    lib().vxo_pvec_update(_dp(pv), C.c_int64(pv.shape[0]), _dp(_f64(pose12)), _dp(_f64(rot_var)), _dp(_f64(tsl_var)), _dp(pw))
    return pv, pw[: pv.shape[0]]


class RefImuWindow:
    """W-1 real IMU_PRE objects of the reference build (BACKEND == "ref" only), fed with the same seeded samples as synth.ImuWindow."""

    def __init__(self, poses12_true, T=0.1, samples=20, gyr_noise=1e-3, acc_noise=1e-2, seed=7):
        p = _f64(poses12_true).reshape(-1, 12)
        self.W = p.shape[0]
        lib().vxo_imu_create.restype = C.c_void_p
        self._h = C.c_void_p(lib().vxo_imu_create(_dp(p), C.c_int(self.W), C.c_double(T), C.c_int(samples), C.c_double(gyr_noise), C.c_double(acc_noise), C.c_uint64(seed)))

    def reset(self):
        lib().vxo_imu_reset(self._h)

    def eval(self, states24, with_gravity=False, want_jac=True):
        s = _f64(states24)
        bs = 33 if with_gravity else 30
        blocks, gvec, cost = np.zeros((self.W - 1, bs * bs)), np.zeros((self.W - 1, bs)), C.c_double(0)
        lib().vxo_imu_eval(self._h, _dp(s), C.c_int(self.W), C.c_int(int(with_gravity)), C.c_int(int(want_jac)), _dp(blocks), _dp(gvec), C.byref(cost))
        return cost.value, blocks, gvec

    def __del__(self):
        try:
            if self._h:
                lib().vxo_imu_destroy(self._h); self._h = None
        except Exception:
            pass


# ---------------------------------------------------------------- stateful local map after margi
class LocalMap:
    def __init__(self, mp, pts_body, scan_offsets, poses12, var_diag, mgsize=0):
        pts = _f64(pts_body).reshape(-1, 3)
        off = np.ascontiguousarray(scan_offsets, dtype=np.int64)
        self.W = off.shape[0] - 1
        lib().vxo_local_map_build.restype = C.c_void_p
        self._h = lib().vxo_local_map_build(C.byref(mp), _dp(pts), off.ctypes.data_as(C.POINTER(C.c_int64)), _dp(_f64(poses12)), C.c_int(self.W), C.c_double(var_diag), C.c_int(mgsize))
        if not self._h:
            raise RuntimeError("vxo_local_map_build failed")

    def __del__(self):
        if getattr(self, "_h", None):
            lib().vxo_local_map_free(C.c_void_p(self._h)); self._h = None

    def planes(self):
        lib().vxo_local_map_planes.restype = C.c_int64
        n = lib().vxo_local_map_planes(C.c_void_p(self._h), None, C.c_int64(0))
        rows = np.zeros((max(n, 1), 52))
        lib().vxo_local_map_planes(C.c_void_p(self._h), _dp(rows), C.c_int64(n))
        r = rows[:n]
        return dict(center=r[:, 0:3], normal=r[:, 3:6], plane_var=r[:, 6:42].reshape(-1, 6, 6), radius=r[:, 42], N=r[:, 43], voxel_center=r[:, 44:47], half=r[:, 47],
                    eig=r[:, 49:52])

    def odom_accumulate(self, pv12, pose12, rot_var, tsl_var, passes=1):
        pv = _f64(pv12).reshape(-1, 12)
        HTH, HTz, nnt = np.zeros((6, 6)), np.zeros(6), np.zeros((3, 3))
        flags = np.zeros(pv.shape[0], dtype=np.int32)
        m = lib().vxo_local_map_odom_accumulate(C.c_void_p(self._h), _dp(pv), C.c_int64(pv.shape[0]), _dp(_f64(pose12)), _dp(_f64(rot_var)), _dp(_f64(tsl_var)), C.c_int(passes),
                                                _dp(HTH), _dp(HTz), _dp(nnt), flags.ctypes.data_as(C.POINTER(C.c_int32)))
        return dict(n=m, HTH=HTH, HTz=HTz, nnt=nnt, flags=flags)


    def lio_state_estimation(self, pv12, state24, cov):
        """ref backend only: the reference's own lio_state_estimation (voxelslam.cpp:856-954, cut out at build time) on this map -> (ok, state24, cov 15x15)."""
        if BACKEND != "ref":
            raise RuntimeError("lio_state_estimation exists for the reference build only")
        pv = _f64(pv12).reshape(-1, 12)
        st, cv = _f64(state24).copy(), _f64(cov).copy()
        ok = lib().vxo_local_map_lio_state_estimation(C.c_void_p(self._h), _dp(pv), C.c_int64(pv.shape[0]), _dp(st), _dp(cv))
        return bool(ok), st, cv.reshape(15, 15)


class SlidingSim:
    """Map side of the sliding-window loop (voxelslam.cpp:1599-1686): cut + recut + tras_opt per scan, margi + ring rotation when full."""

    def __init__(self, mp, win_size, mgsize=1, max_points=100):
        self.W = win_size
        lib().vxo_sliding_sim_create.restype = C.c_void_p
        self._h = lib().vxo_sliding_sim_create(C.byref(mp), C.c_int(win_size), C.c_int(mgsize), C.c_int(max_points))

    def __del__(self):
        if getattr(self, "_h", None):
            lib().vxo_sliding_sim_free(C.c_void_p(self._h)); self._h = None

    def add_scan(self, pts_body, pose12, var_diag=1e-4):
        p = _f64(pts_body).reshape(-1, 3)
        lib().vxo_sliding_sim_add_scan(C.c_void_p(self._h), _dp(p), C.c_int64(p.shape[0]), _dp(_f64(pose12)), C.c_double(var_diag))

    def add_scan_pv(self, pv12, pose12, ba_iters=0):
        """full pointVar records (pnt | var 3x3); ba_iters > 0 runs a pose-only BA between tras_opt and margi once the window is full"""
        p = _f64(pv12).reshape(-1, 12)
        lib().vxo_sliding_sim_add_scan_pv(C.c_void_p(self._h), _dp(p), C.c_int64(p.shape[0]), _dp(_f64(pose12)), C.c_int(ba_iters))

    def factor(self):
        lib().vxo_sliding_sim_factor.restype = C.c_void_p
        return OracleFactor(lib().vxo_sliding_sim_factor(C.c_void_p(self._h)), self.W)

    def planes(self):
        lib().vxo_sliding_sim_planes.restype = C.c_int64
        n = lib().vxo_sliding_sim_planes(C.c_void_p(self._h), None, C.c_int64(0))
        rows = np.zeros((max(n, 1), 52))
        lib().vxo_sliding_sim_planes(C.c_void_p(self._h), _dp(rows), C.c_int64(n))
        r = rows[:n]
        return dict(center=r[:, 0:3], normal=r[:, 3:6], plane_var=r[:, 6:42].reshape(-1, 6, 6), radius=r[:, 42], N=r[:, 43], voxel_center=r[:, 44:47], half=r[:, 47],
                    cov_trace=r[:, 48], eig=r[:, 49:52])

    def odom_accumulate(self, pv12, pose12, rot_var, tsl_var):
        pv = _f64(pv12).reshape(-1, 12)
        HTH, HTz, nnt = np.zeros((6, 6)), np.zeros(6), np.zeros((3, 3))
        flags = np.zeros(pv.shape[0], dtype=np.int32)
        m = lib().vxo_sliding_sim_odom_accumulate(C.c_void_p(self._h), _dp(pv), C.c_int64(pv.shape[0]), _dp(_f64(pose12)), _dp(_f64(rot_var)), _dp(_f64(tsl_var)), _dp(HTH), _dp(HTz),
                                                  _dp(nnt), flags.ctypes.data_as(C.POINTER(C.c_int32)))
        return dict(n=m, HTH=HTH, HTz=HTz, nnt=nnt, flags=flags)

    def state(self):
        W = self.W
        head = np.zeros(2 + W, dtype=np.int32)
        poses = np.zeros((W, 12))
        lib().vxo_sliding_sim_state.restype = C.c_int64
        n = lib().vxo_sliding_sim_state(C.c_void_p(self._h), head.ctypes.data_as(C.POINTER(C.c_int32)), _dp(poses), None, C.c_int64(0))
        rows = np.zeros((max(n, 1), 32 + 10 * W))
        lib().vxo_sliding_sim_state(C.c_void_p(self._h), head.ctypes.data_as(C.POINTER(C.c_int32)), _dp(poses), _dp(rows), C.c_int64(n))
        r = rows[:n]
        return dict(win_count=int(head[0]), win_base=int(head[1]), ring=head[2:].copy(), poses=poses[: int(head[0])], voxel_center=r[:, 0:3], half=r[:, 3], layer=r[:, 4].astype(int),
                    is_plane=r[:, 5] > 0, isexist=r[:, 6] > 0, has_sw=r[:, 7] > 0, in_slide=r[:, 8] > 0, opt_state=r[:, 9].astype(int), last_num=r[:, 10].astype(int),
                    n_point_fix=r[:, 11].astype(int), pcr_add=r[:, 12:22], pcr_fix=r[:, 22:32], slots=r[:, 32:].reshape(n, W, 10))
