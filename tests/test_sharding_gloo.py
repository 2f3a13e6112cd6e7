"""N > 1 host logic on CPU (gloo, world_size 2): voxel sharding by the reference hash of the root cell + additivity of
[H, g, r] over shards — what the NCCL all-reduce of the global-BA step relies on."""
import os
import sys

import numpy as np
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, q):
    sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
    import torch
    import oracle_api as oa
    import scenes
    from voxel_slam_b200 import sharding
    os.environ["MASTER_ADDR"] = "127.0.0.1"; os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    sc = scenes.make_window(W=5, pts_per_scan=4000, L=6.0, seed=3)
    own = sharding.owner_of(sc["ids"], world)
    mine = own == rank
    # every octree (root + sub-voxels) lives on exactly one rank
    roots = {}
    for i, o in zip(sc["ids"], own):
        roots.setdefault((int(i["x"]), int(i["y"]), int(i["z"])), set()).add(int(o))
    assert all(len(v) == 1 for v in roots.values())
    of = oa.OracleFactor.from_dense(5, sc["clusters10"][mine], sc["fix10"][mine], None, sc["eig12"][mine], sc["sum10"][mine])
    H, J, r = of.hessian(sc["poses_est"])
    buf = torch.from_numpy(np.concatenate([H.ravel(order="F"), J, [r], [float(mine.sum())]]))
    dist.all_reduce(buf)                                   # the same [H | g | r] block the GPU path all-reduces with NCCL
    r2 = torch.tensor([of.residual(sc["poses_true"])], dtype=torch.float64)
    dist.all_reduce(r2)
    if rank == 0:
        Hf, Jf, rf = sc["oracle_factor"].hessian(sc["poses_est"])
        n = 30
        out = buf.numpy()
        ok = (np.max(np.abs(out[: n * n].reshape(n, n, order="F") - Hf)) < 1e-12 * np.max(np.abs(Hf)) and np.max(np.abs(out[n * n: n * n + n] - Jf)) < 1e-12 * np.max(np.abs(Jf))
              and abs(out[n * n + n] - rf) < 1e-13 * rf and int(out[-1]) == len(sc["ids"]) and 0 < mine.sum() < len(sc["ids"])
              and abs(float(r2) - sc["oracle_factor"].residual(sc["poses_true"])) < 1e-12)
        q.put(bool(ok))
    dist.barrier()
    dist.destroy_process_group()


def test_voxel_sharded_hessian_allreduce_gloo():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + (os.getpid() % 2000)
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    [p.start() for p in procs]
    [p.join(180) for p in procs]
    assert all(p.exitcode == 0 for p in procs)
    assert q.get(timeout=5) is True


def test_shard_csr_helper():
    from voxel_slam_b200 import sharding
    ptr = np.array([0, 2, 3, 6]); fr = np.array([0, 1, 2, 0, 1, 2]); cl = np.arange(60.0).reshape(6, 10)
    p2, f2, c2, (e2,) = sharding.shard_csr(ptr, fr, cl, [np.array([10, 11, 12])], [True, False, True])
    assert p2.tolist() == [0, 2, 5] and f2.tolist() == [0, 1, 0, 1, 2] and e2.tolist() == [10, 12] and c2[2, 0] == 30.0
    assert sharding.voxel_hash(-3, 0, -5) == 5254958208       # matches tests/golden/voxel_keys.json


def _hba_worker(rank, world, port, q):
    """bottom level of the hierarchical global BA over `world` ranks: every rank takes its share of the windows, solves them with the ORACLE (the windows are
    independent problems), and the per-window results gathered over gloo equal the single-process pass"""
    sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
    import torch
    import oracle_api as oa
    import scenes
    import voxel_slam_b200 as vx
    from voxel_slam_b200 import sharding
    os.environ["MASTER_ADDR"] = "127.0.0.1"; os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    K, ws, st = 14, 6, 2
    tr, est = scenes.poses_true_est(K, 8.0, 71, rot_sigma=3e-3, pos_sigma=2e-2)
    xyz, off = scenes.make_points(K, 1500, 8.0, 71, tr, dtype=np.float32)
    fine = vx.MapParams.make(voxel_size=1.0, min_eigen_value=0.0025, max_layer=2)
    wf = sharding.hba_windows(K, ws, st)
    lo, hi = sharding.window_share(len(wf), rank, world)

    def solve(k0):
        a, b = off[k0], off[k0 + ws]
        return oa.hba_window(fine, fine, xyz[a:b], off[k0:k0 + ws + 1] - a, est[k0:k0 + ws], max_iter=1, thread_num=2)["poses"]

    mine = torch.zeros((len(wf), ws, 12), dtype=torch.float64)
    for w in range(lo, hi):
        mine[w] = torch.from_numpy(solve(int(wf[w])))
    dist.all_reduce(mine)                                   # disjoint shares: the sum is the concatenation
    if rank == 0:
        shares = [sharding.window_share(len(wf), r, world) for r in range(world)]
        ok = shares[0][0] == 0 and shares[-1][1] == len(wf) and all(shares[r][1] == shares[r + 1][0] for r in range(world - 1)) and len(wf) == 5
        ok = ok and all(np.array_equal(mine[w].numpy(), solve(int(wf[w]))) for w in range(len(wf)))
        q.put(bool(ok))
    dist.barrier()
    dist.destroy_process_group()


def test_hba_bottom_windows_distributed_gloo():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 31500 + (os.getpid() % 2000)
    procs = [ctx.Process(target=_hba_worker, args=(r, 2, port, q)) for r in range(2)]
    [p.start() for p in procs]
    [p.join(240) for p in procs]
    assert all(p.exitcode == 0 for p in procs)
    assert q.get(timeout=5) is True
