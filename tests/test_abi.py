"""The C-ABI library loads and exports every symbol include/vxs.h declares; without a GPU it fails loudly (no CPU fallback)."""
import ctypes as C
import os

import pytest
import torch

import synth
import voxel_slam_b200 as vx


def test_library_exports_every_declared_symbol():
    syms = vx.declared_symbols()
    assert len(syms) >= 25 and "vxs_li_ba" in syms and "vxs_build_window_factor" in syms
    lib = vx.lib()
    missing = [s for s in syms if not hasattr(lib, s)]
    assert not missing, missing
    assert lib.vxs_version() >= 100


def test_harness_library_loads():
    p = synth.true_pose(20.0, 3)
    assert p.shape == (12,) and abs(p[9] - 10.15) < 1e-12
    pts = synth.gen_scan(20.0, 0, 1000, synth.true_pose(20.0, 0))
    assert pts.shape == (1000, 3) and abs(pts).max() < 40
    assert (synth.gen_scan(20.0, 0, 1000, synth.true_pose(20.0, 0)) == pts).all()          # seeded, reproducible


@pytest.mark.skipif(torch.cuda.is_available(), reason="checks the no-GPU behaviour")
def test_no_cpu_fallback():
    with pytest.raises(vx.VxsError) as ei:
        vx.Context(0)
    assert ei.value.code == -1


def test_product_never_touches_the_oracle():
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    bad = []
    for d, _, files in os.walk(os.path.join(root, "voxel_slam_b200")):
        for fn in files:
            if fn.endswith((".py", ".cu", ".cuh", ".h", ".hpp", ".cpp")) or fn == "Makefile":
                txt = open(os.path.join(d, fn), errors="ignore").read()
                if "oracle/" in txt or "liboracle" in txt or "oracle_api" in txt or "vxo_" in txt:
                    bad.append(os.path.join(d, fn))
    assert not bad, bad


def test_cpp_shim_plain_layer_compiles():
    """The header-only C++ shim with the reference's class names compiles (plain layer; the Eigen-typed layer needs the reference headers)."""
    import subprocess
    import tempfile
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    with tempfile.NamedTemporaryFile("w", suffix=".cpp", delete=False) as f:
        f.write('#include "voxel_slam_b200/csrc/shim/voxel_ba_shim.hpp"\nint main() { vxs_shim::Lidar_BA_Optimizer o; return o.thd_num == 2 ? 0 : 1; }\n')
    r = subprocess.run(["g++", "-std=c++14", "-fsyntax-only", "-I", root, f.name], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr


def test_cpp_shim_typed_layer_type_checks_against_reference_signatures():
    """The Eigen/PCL-typed layer of the shim (push_voxel, damping_iter wrappers, IMU adapter, down-sampling, submap merge) type-checks against
    stand-ins that mirror the reference's names, members and call signatures (tests/shim_stubs/reference_stubs.hpp; Eigen/PCL are not
    installed here), instantiated the way voxelslam.cpp calls them."""
    import subprocess
    import tempfile
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    src = r'''
#include "tests/shim_stubs/reference_stubs.hpp"
#define VXS_SHIM_WITH_REFERENCE_TYPES
#include "voxel_slam_b200/csrc/shim/voxel_ba_shim.hpp"
void use(vxs_shim::Context& c, std::vector<IMUST>& xs, std::deque<IMU_PRE*>& imus, std::vector<Keyframe*>& smps, PVec& pvec, pcl::PointCloud<PointType>& pl) {
  vxs_shim::LidarFactor f(c, int(xs.size()));
  std::vector<PointCluster> pcrs(xs.size()); PointCluster fix, add; Eigen::Vector3d ev; Eigen::Matrix3d U;
  vxs_shim::push_voxel(f, pcrs, fix, 1.0, ev, U, add);
  Eigen::MatrixXd hess; std::vector<double> resis;
  vxs_shim::li_ba_damping_iter(xs, f, imus, &hess, 1e-4);
  vxs_shim::lidar_ba_damping_iter(xs, f, &hess, resis, 3, 2);
  vxs_shim::down_sampling_voxel(c, pl, 0.1);
  vxs_shim::down_sampling_close(c, pl, 0.1);
  vxs_shim::down_sampling_pvec(c, pvec, 0.1, pl);
  vxs_shim::submap_merge(c, xs, smps, 1.0, pl);
}
int main() { return 0; }
'''
    with tempfile.NamedTemporaryFile("w", suffix=".cpp", delete=False) as f:
        f.write(src)
    r = subprocess.run(["g++", "-std=c++14", "-fsyntax-only", "-I", root, f.name], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr


def test_bench_reference_arm_contract_on_cpu():
    """`bench.py --impl reference` needs no GPU: one JSON line with the contract keys (tiny window so that it runs in seconds)."""
    import json
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    r = subprocess.run([sys.executable, os.path.join(root, "bench.py"), "--impl", "reference", "--win", "6", "--pts-per-scan", "4000", "--L", "8", "--steps", "1", "--warmup", "1"],
                       capture_output=True, text=True, cwd=root, timeout=600)
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [l for l in r.stdout.splitlines() if l.startswith("{")]
    assert len(lines) == 1
    d = json.loads(lines[0])
    for k in ("impl", "metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "e2e", "cpu_baseline", "config"):
        assert k in d, k
    assert d["impl"] == "reference" and d["value"] > 0 and d["e2e"]["h2d_bytes_per_step"] == 0 and d["cpu_baseline"]["kind"] == "port"
    assert d["cpu_baseline"]["all_cores_variant"].get("value", 0) > 0
