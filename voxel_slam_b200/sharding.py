"""Voxel sharding across GPUs for global BA: a voxel belongs to rank hash(VOXEL_LOC of its ROOT cell) mod n.
Sharding by the root cell keeps a whole octree (root + its sub-voxels) on one rank, so cut/recut shard the same way.
The hash is the reference's (tools.hpp:39-48), evaluated in 64-bit wrap-around arithmetic."""
import numpy as np

HASH_P = 116101
MAX_N = 10000000000
_M64 = (1 << 64) - 1


def voxel_hash(x, y, z):
    x, y, z = (int(v) & _M64 for v in (x, y, z))
    return ((((z * HASH_P & _M64) % MAX_N + y & _M64) * HASH_P & _M64) % MAX_N + x) & _M64


def owner_of(ids, nranks):
    """ids: structured array with x, y, z (root cell of each factor voxel) -> owning rank per voxel."""
    return np.array([voxel_hash(i["x"], i["y"], i["z"]) % nranks for i in ids], dtype=np.int64)


def shard_csr(ptr, frame, clusters, per_voxel_arrays, keep):
    """Select the voxels flagged in `keep` from a CSR factor (host arrays)."""
    keep = np.asarray(keep, dtype=bool)
    counts = np.diff(ptr)
    ent_keep = np.repeat(keep, counts)
    new_ptr = np.concatenate([[0], np.cumsum(counts[keep])]).astype(np.int64)
    return new_ptr, frame[ent_keep], clusters[ent_keep], [a[keep] for a in per_voxel_arrays]


# ---------------------------------------------------------------------------------------------- bottom level of the hierarchical global BA
def hba_windows(K, win_size=10, win_stride=5):
    """first keyframe of every bottom-level window (thd_globalmapping: wdsize 10, mgsize 5, voxelslam.cpp:2501-2502) — the rule vxs_hba_pass uses"""
    return np.arange(0, (K - win_size) // win_stride + 1, dtype=np.int32) * win_stride if K >= win_size else np.zeros(0, dtype=np.int32)


def window_share(nwin, rank, nranks):
    """contiguous share [lo, hi) of the bottom-level windows that rank `rank` of `nranks` solves (windows are independent problems: no collective)"""
    return nwin * rank // nranks, nwin * (rank + 1) // nranks
