"""Multi-GPU parity worker for the voxel-sharded global-BA step; launched under torchrun by tests/test_gpu_multi.py (pytest -m gpu, >= 2 GPUs):
   python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 tests/multi_gpu_worker.py
Every rank builds the same seeded window with the ORACLE map, keeps the voxels it owns (reference hash of the root cell mod n),
pushes them to its GPU and runs Lidar_BA_Optimizer::damping_iter with the NCCL all-reduce of [C | g | D | r] inside libvxs.
Rank 0 compares poses / residuals / Hessian with the single-process oracle solve."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import scenes  # noqa: E402
import voxel_slam_b200 as vx  # noqa: E402
from voxel_slam_b200 import sharding  # noqa: E402


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl")
    ctx = vx.Context(local)
    uid = [vx.Context.comm_unique_id() if rank == 0 else None]
    dist.broadcast_object_list(uid, src=0)
    ctx.comm_init(uid[0], rank, world)
    ok_all = True
    for (W, pts, L, seed) in [(6, 6000, 6.0, 5), (12, 20000, 14.0, 9)]:
        sc = scenes.make_window(W=W, pts_per_scan=pts, L=L, seed=seed)
        mine = sharding.owner_of(sc["ids"], world) == rank
        f = vx.Factor(ctx, W)
        f.push_voxels_dense(sc["clusters10"][mine], sc["eig12"][mine], sc["sum10"][mine], fix10=sc["fix10"][mine])
        g = ctx.lidar_ba(f, sc["poses_est"], max_iter=4, thd_num=1)
        r_all = ctx.evaluate_residual(f, sc["poses_true"])
        if rank == 0:
            ref = sc["oracle_factor"].lidar_ba(sc["poses_est"], max_iter=4)
            r_ref = sc["oracle_factor"].residual(sc["poses_true"])
            inc = np.max(np.abs(ref["poses"] - sc["poses_est"]))
            e_pose = np.max(np.abs(g["poses"] - ref["poses"])) / inc
            e_h = np.max(np.abs(g["hess"] - ref["hess"])) / np.max(np.abs(ref["hess"]))
            e_r = abs(g["resis"][1] - ref["resis"][1]) / ref["resis"][1]
            e_r2 = abs(r_all - r_ref) / r_ref
            ok = e_pose < 1e-5 and e_h < 1e-8 and e_r < 1e-9 and e_r2 < 1e-9 and len(g["trace"]) == len(ref["trace"])
            print(f"[multi-gpu x{world}] W={W}: shard sizes {int(mine.sum())}/{len(mine)}  pose {e_pose:.2e}  hess {e_h:.2e}  resid {e_r:.2e} {e_r2:.2e}  -> {'OK' if ok else 'FAIL'}", flush=True)
            ok_all = ok_all and ok
        f.close()
    # a rank that owns NO voxel (hash(root cell) mod n can leave a shard empty) must still enter both all-reduces (vxs_eval.cu):
    # everything on rank 0, nothing on the others
    sc = scenes.make_window(W=6, pts_per_scan=6000, L=6.0, seed=5)
    f = vx.Factor(ctx, 6)
    if rank == 0:
        f.push_voxels_dense(sc["clusters10"], sc["eig12"], sc["sum10"], fix10=sc["fix10"])
    g = ctx.lidar_ba(f, sc["poses_est"], max_iter=4, thd_num=1)
    r_all = ctx.evaluate_residual(f, sc["poses_true"])
    ref = sc["oracle_factor"].lidar_ba(sc["poses_est"], max_iter=4)
    r_ref = sc["oracle_factor"].residual(sc["poses_true"])
    inc = np.max(np.abs(ref["poses"] - sc["poses_est"]))
    ok = (np.max(np.abs(g["poses"] - ref["poses"])) / inc < 1e-5 and abs(r_all - r_ref) / r_ref < 1e-9 and len(g["trace"]) == len(ref["trace"])
          and all(a["accepted"] == b["accepted"] for a, b in zip(g["trace"], ref["trace"])))
    print(f"[multi-gpu x{world}] rank {rank} empty-shard case (V={f.counts()[0]}): {'OK' if ok else 'FAIL'}", flush=True)
    okt = torch.tensor([1 if ok else 0], device=f"cuda:{local}")
    dist.all_reduce(okt, op=dist.ReduceOp.MIN)          # every rank must have followed the same LM path
    ok_all = ok_all and int(okt.item()) == 1
    f.close()
    # hierarchical-GBA window: every rank sees all keyframe clouds, builds only the octrees it owns, all-reduces the Hessian
    import oracle_api as oa
    W = 10
    tr, est = scenes.poses_true_est(W, 8.0, 53, rot_sigma=3e-3, pos_sigma=2e-2)
    xyz, off = scenes.make_points(W, 4000, 8.0, 53, tr, dtype=np.float32)
    coarse = vx.MapParams.make(voxel_size=2.0, min_eigen_value=0.1, max_layer=2)
    fine = vx.MapParams.make(voxel_size=1.0, min_eigen_value=0.0025, max_layer=2)
    g = ctx.hba_window(coarse, fine, xyz, off, est, max_iter=4, thread_num=1)
    if rank == 0:
        r = oa.hba_window(coarse, fine, xyz, off, est, max_iter=4, thread_num=2)
        inc = np.max(np.abs(r["poses"] - est))
        e_pose = np.max(np.abs(g["poses"] - r["poses"])) / inc
        e_h = np.max(np.abs(g["hess"] - r["hess"])) / np.max(np.abs(r["hess"]))
        ok = g["outer_iters"] == r["outer_iters"] and e_pose < 1e-5 and e_h < 1e-6
        print(f"[multi-gpu x{world}] sharded HBA window: outer iters {g['outer_iters']}/{r['outer_iters']}  pose {e_pose:.2e}  hess {e_h:.2e} -> {'OK' if ok else 'FAIL'}", flush=True)
        ok_all = ok_all and ok
    # the whole hierarchical pass (vxs_hba_pass): bottom windows distributed over the ranks, merged submaps exchanged device to device over NCCL, top level voxel-sharded —
    # compared with the same pass run by ONE extra single-GPU context through the separate host-buffer calls
    K, ws, stp = 30, 10, 5
    tr, est = scenes.poses_true_est(K, 8.0, 99, rot_sigma=3e-3, pos_sigma=2e-2)
    xyz, off = scenes.make_points(K, 3000, 8.0, 99, tr, dtype=np.float32)
    p = ctx.hba_pass(fine, fine, xyz, off, est, win_size=ws, win_stride=stp, nranks=world, rank=rank)
    solo = vx.Context(local)
    wf = np.arange(0, K - ws + 1, stp, dtype=np.int32)
    b = solo.hba_bottom_batch(fine, xyz, off, est, wf, win_size=ws)
    m = solo.submap_merge_batch(xyz, off, b["poses"], wf, 0.125)
    top = solo.hba_window(fine, fine, m["xyz"], m["win_offsets"], est[wf], max_iter=1, thread_num=5)
    lo, n_mine = p["first_window"], p["window_count"]
    inc = np.max(np.abs(top["poses"] - est[wf]))
    ok = (np.max(np.abs(p["bottom_poses"][:n_mine] - b["poses"][lo:lo + n_mine])) < 1e-10 and np.array_equal(p["submap_sizes"], np.diff(m["win_offsets"]))
          and np.max(np.abs(p["top_poses"] - top["poses"])) < 1e-6 * inc and abs(p["top_resis"][1] - top["resis_log"][1]) / top["resis_log"][1] < 1e-9)
    print(f"[multi-gpu x{world}] rank {rank} hierarchical pass: windows {lo}..{lo + n_mine - 1}, top pose err {np.max(np.abs(p['top_poses'] - top['poses'])) / inc:.2e} -> {'OK' if ok else 'FAIL'}", flush=True)
    okt = torch.tensor([1 if ok else 0], device=f"cuda:{local}")
    dist.all_reduce(okt, op=dist.ReduceOp.MIN)
    ok_all = ok_all and int(okt.item()) == 1
    solo.close()
    flag = torch.tensor([1 if ok_all else 0], device=f"cuda:{local}")
    dist.broadcast(flag, src=0)
    dist.barrier()
    ctx.close()
    dist.destroy_process_group()
    sys.exit(0 if int(flag.item()) == 1 else 1)


if __name__ == "__main__":
    main()
