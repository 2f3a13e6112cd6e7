// Voxel-map construction on the GPU, build-from-scratch semantics (what motion_init voxelslam.cpp:600-628, loop_update
// :1171-1180 and every HBA pass :2374-2379 do): cut every scan into root voxels, then the top-down recut of the adaptive
// octree, then factor extraction.  Reference functions replaced (voxel_map.hpp / loop_refine.hpp):
//   cut_voxel :1504-1540, cut_voxel(fix) :1641-1671, OctoTree::allocate/push/subdivide/fix_divide :969-1116,
//   OctoTree::recut + plane_judge :1015-1019,1148-1194, tras_opt :1308-1333, LidarFactor::push_voxel :122-130,
//   OctreeGBA::cut_voxel/push/subdivide/recut LR:316-405,446-479, OctreeGBA_multi_recut LR:483-537.
//
// GPU formulation (no pointer-chasing octree, no per-voxel mutex):
//   * every point gets, in one pass, its root cell (bit-exact float quantisation) and the octant it would fall into at each
//     deeper layer — child centres depend only on the root cell and the octant path, so the whole descent is known up front;
//   * per layer: 64-bit key (node | frame) -> LSD radix sort (8-bit digits, warp match-any ranking, stable) -> contiguous
//     (node, frame) segments -> one warp per segment accumulates the body-frame cluster (pcrs_local) and the world cluster
//     (pcr_add) with shuffle reductions -> one thread per node: sums, covariance, fp64 Jacobi eigensolve, plane test ->
//     dead / plane / subdivide; only the points of subdivided nodes are re-keyed for the next layer;
//   * plane leaves passing the tras_opt filter are appended straight into the device-resident CSR factor.
// Fixed map points are a pseudo-frame W with identity pose: their clusters become pcr_fix.
#include <limits.h>
#include <string.h>
#include <algorithm>
#include <vector>
#include "vxs_internal.h"
#include "vxs_math.cuh"
#include "vxs_sortscan.cuh"

using namespace vxs;

// ------------------------------------------------------------------ scratch
struct VoxScratch : SortScratch {   // SortScratch: hist, blocksums, totals (vxs_sortscan.cuh)
  DevBuf<double> pts_d; DevBuf<float> pts_f, pts_f2; DevBuf<double> poses; DevBuf<long long> offsets, src_off; DevBuf<long long> bbox;
  DevBuf<unsigned long long> keysA, keysB; DevBuf<unsigned int> idxA, idxB; DevBuf<unsigned short> pathbits;
  DevBuf<unsigned int> flags, scanbuf;
  DevBuf<unsigned int> rec_start, node_of_rec, node_rec_start, rec_node_flag;
  DevBuf<unsigned long long> rec_key;
  DevBuf<double> rec_local, rec_world;
  // node arrays, two generations (current level / parent level)
  DevBuf<int> node_state[2]; DevBuf<long long> node_root[2]; DevBuf<int> node_path[2];
  DevBuf<double> node_eig, node_sum, node_fix; DevBuf<unsigned int> node_nent, node_sel, node_voff, node_eoff;
  DevBuf<vxs_voxel_id> ids;
  std::vector<vxs_voxel_id> ids_host;
  vxs_factor* hba_factor = nullptr;   // vxs_hba_window's factor, kept between calls (owned by the ctx's factor list)
};
static VoxScratch* scratch(vxs_ctx* c) { if (!c->vox_scratch) c->vox_scratch = new VoxScratch(); return static_cast<VoxScratch*>(c->vox_scratch); }
int vxs_comm_allreduce_max_i64(vxs_ctx* ctx, long long* buf, size_t n);   // vxs_lm.cu
void vxs_voxelize_release(vxs_ctx* c) {
  if (!c->vox_scratch) return;
  VoxScratch* s = static_cast<VoxScratch*>(c->vox_scratch);
  s->pts_d.release(); s->pts_f.release(); s->pts_f2.release(); s->poses.release(); s->offsets.release(); s->src_off.release(); s->bbox.release(); s->keysA.release(); s->keysB.release();
  s->idxA.release(); s->idxB.release(); s->pathbits.release(); s->hist.release(); s->flags.release(); s->scanbuf.release(); s->blocksums.release();
  s->totals.release(); s->rec_start.release(); s->node_of_rec.release(); s->node_rec_start.release(); s->rec_node_flag.release(); s->rec_key.release();
  s->rec_local.release(); s->rec_world.release();
  for (int g = 0; g < 2; g++) { s->node_state[g].release(); s->node_root[g].release(); s->node_path[g].release(); }
  s->node_eig.release(); s->node_sum.release(); s->node_fix.release(); s->node_nent.release(); s->node_sel.release(); s->node_voff.release(); s->node_eoff.release();
  s->ids.release();
  delete s;
  c->vox_scratch = nullptr;
}

// ------------------------------------------------------------------ bit-exact point -> cell arithmetic (no FMA contraction)
// world = R*p + t with the oracle's summation order ((r0*x + r1*y) + r2*z) + t   (voxelslam.cpp:616, loop_refine.hpp:451)
__device__ __forceinline__ double dot3_rn(double a0, double a1, double a2, double x, double y, double z, double t) {
  return __dadd_rn(__dadd_rn(__dadd_rn(__dmul_rn(a0, x), __dmul_rn(a1, y)), __dmul_rn(a2, z)), t);
}
__device__ __forceinline__ d3 world_point(const double* __restrict__ pose, d3 p) {
  return mk3(dot3_rn(pose[0], pose[1], pose[2], p.x, p.y, p.z, pose[9]), dot3_rn(pose[3], pose[4], pose[5], p.x, p.y, p.z, pose[10]),
             dot3_rn(pose[6], pose[7], pose[8], p.x, p.y, p.z, pose[11]));
}
// voxel_map.hpp:1511-1518: float loc = pw/voxel_size; if(loc<0) loc -= 1; (int64_t)loc
__device__ __forceinline__ long long quantise(double pw, double voxel_size) {
  float loc = __double2float_rn(__ddiv_rn(pw, voxel_size));
  if (loc < 0.0f) loc = __fsub_rn(loc, 1.0f);
  return __float2ll_rz(loc);
}
// tools.hpp:39-48
__device__ __forceinline__ unsigned long long voxel_hash(long long x, long long y, long long z) {
  const unsigned long long HP = 116101ull, MN = 10000000000ull;
  return (((((unsigned long long)z * HP) % MN + (unsigned long long)y) * HP) % MN) + (unsigned long long)x;
}
// octant path down to max_layer: 3 bits per layer, layer 1 in the low bits (voxel_map.hpp:1029-1040)
__device__ __forceinline__ unsigned int octant_bits(d3 pw, long long kx, long long ky, long long kz, double voxel_size, int max_layer) {
  double cx = __dmul_rn(__dadd_rn(0.5, (double)kx), voxel_size), cy = __dmul_rn(__dadd_rn(0.5, (double)ky), voxel_size), cz = __dmul_rn(__dadd_rn(0.5, (double)kz), voxel_size);
  float q = __double2float_rn(__ddiv_rn(voxel_size, 4.0));
  unsigned int bits = 0;
  for (int l = 0; l < max_layer; l++) {
    const int bx = pw.x > cx, by = pw.y > cy, bz = pw.z > cz;
    bits |= (unsigned int)(4 * bx + 2 * by + bz) << (3 * l);
    cx = __dadd_rn(cx, (double)__fmul_rn((float)(2 * bx - 1), q));
    cy = __dadd_rn(cy, (double)__fmul_rn((float)(2 * by - 1), q));
    cz = __dadd_rn(cz, (double)__fmul_rn((float)(2 * bz - 1), q));
    q = __fdiv_rn(q, 2.0f);
  }
  return bits;
}

struct PointSrc {
  const double* pd; const float* pf; int fstride;  // exactly one of pd / pf
  const long long* offsets;                        // [nframes+1] (nframes = W or W+1 with the fix pseudo-frame)
  const double* poses;                             // [nframes][12]
  int nframes;
  long long n;
  // batch mode (many independent windows in one build, vxs_hba_bottom_batch): a "frame" is a (window, slot) pair, frame = window * win_size + slot;
  // the index space is the concatenation of the windows' clouds, src_off[frame] = first point of that keyframe in the point array (keyframes
  // shared by overlapping windows are stored once).  win_size == 0: one window, the index space IS the point array.
  const long long* src_off; int win_size;
  // routed mode (top level of vxs_hba_pass on several GPUs): the points are float4 records {x, y, z, frame as int bits} that were sent to this rank
  // because it owns their root cell; there is no offsets array
  int frame_in_w;
};
__device__ __forceinline__ d3 load_point(const PointSrc& s, long long i, int fr) {
  if (s.src_off) i = s.src_off[fr] + (i - s.offsets[fr]);
  if (s.pd) return mk3(s.pd[3 * i], s.pd[3 * i + 1], s.pd[3 * i + 2]);
  const float* p = s.pf + size_t(i) * s.fstride;
  return mk3((double)p[0], (double)p[1], (double)p[2]);
}
__device__ __forceinline__ int frame_of(const PointSrc& s, long long i) {
  if (s.frame_in_w) return __float_as_int(s.pf[size_t(i) * s.fstride + 3]);
  int lo = 0, hi = s.nframes;  // offsets[lo] <= i < offsets[hi]
  while (hi - lo > 1) { const int mid = (lo + hi) >> 1; if (s.offsets[mid] <= i) lo = mid; else hi = mid; }
  return lo;
}

__global__ void k_voxel_keys(const double* __restrict__ pw, long long n, double voxel_size, long long* __restrict__ xyz, unsigned long long* __restrict__ hash) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const long long x = quantise(pw[3 * i], voxel_size), y = quantise(pw[3 * i + 1], voxel_size), z = quantise(pw[3 * i + 2], voxel_size);
  xyz[3 * i] = x; xyz[3 * i + 1] = y; xyz[3 * i + 2] = z;
  hash[i] = voxel_hash(x, y, z);
}

__global__ void k_bbox_negate_min(long long* bbox) { if (threadIdx.x < 3) bbox[threadIdx.x] = -bbox[threadIdx.x]; }   // [min | max] <-> [-min | max] around the MAX all-reduce
__global__ void __launch_bounds__(256) k_bbox(PointSrc s, double voxel_size, long long* __restrict__ bbox, long long i0, long long i1) {   // points [i0, i1)
  long long mn[3] = {LLONG_MAX, LLONG_MAX, LLONG_MAX}, mx[3] = {LLONG_MIN, LLONG_MIN, LLONG_MIN};
  for (long long i = i0 + (long long)blockIdx.x * blockDim.x + threadIdx.x; i < i1; i += (long long)gridDim.x * blockDim.x) {
    const int fr = frame_of(s, i);
    const d3 w = world_point(s.poses + 12 * fr, load_point(s, i, fr));
    const long long k[3] = {quantise(w.x, voxel_size), quantise(w.y, voxel_size), quantise(w.z, voxel_size)};
    for (int a = 0; a < 3; a++) { mn[a] = min(mn[a], k[a]); mx[a] = max(mx[a], k[a]); }
  }
  for (int a = 0; a < 3; a++) {
    for (int off = 16; off > 0; off >>= 1) { mn[a] = min(mn[a], __shfl_xor_sync(0xffffffffu, mn[a], off)); mx[a] = max(mx[a], __shfl_xor_sync(0xffffffffu, mx[a], off)); }
    if ((threadIdx.x & 31) == 0) { atomicMin(bbox + a, mn[a]); atomicMax(bbox + 3 + a, mx[a]); }
  }
}

// key0 = (linear root id << FB) | frame ; pathbits = octants of the deeper layers
__global__ void __launch_bounds__(256) k_point_keys(PointSrc s, double voxel_size, int max_layer, long long minx, long long miny, long long minz, long long ey, long long ez,
                                                    long long cells, int FB, unsigned long long* __restrict__ keys, unsigned int* __restrict__ idx, unsigned short* __restrict__ pathbits,
                                                    int shard_rank, int shard_n, unsigned int* __restrict__ owned) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= s.n) return;
  const int fr = frame_of(s, i);
  const d3 w = world_point(s.poses + 12 * fr, load_point(s, i, fr));
  const long long kx = quantise(w.x, voxel_size), ky = quantise(w.y, voxel_size), kz = quantise(w.z, voxel_size);
  unsigned long long lin = (unsigned long long)(((kx - minx) * ey + (ky - miny)) * ez + (kz - minz));
  int slot = fr;
  if (s.win_size > 0) { const int win = fr / s.win_size; slot = fr - win * s.win_size; lin += (unsigned long long)win * (unsigned long long)cells; }   // every window has its own copy of the cell space
  keys[i] = (lin << FB) | (unsigned long long)slot;
  idx[i] = (unsigned int)i;
  pathbits[i] = (unsigned short)octant_bits(w, kx, ky, kz, voxel_size, max_layer);
  // multi-GPU: a root cell (and its whole octree) belongs to rank hash(VOXEL_LOC) mod n  
  if (owned) owned[i] = (voxel_hash(kx, ky, kz) % (unsigned long long)shard_n) == (unsigned long long)shard_rank ? 1u : 0u;
}
__global__ void k_compact_owned(const unsigned long long* __restrict__ keys, const unsigned int* __restrict__ idx, const unsigned int* __restrict__ owned,
                                const unsigned int* __restrict__ pos, size_t n, unsigned long long* __restrict__ keys_out, unsigned int* __restrict__ idx_out) {
  const size_t i = size_t(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i < n && owned[i]) { keys_out[pos[i]] = keys[i]; idx_out[pos[i]] = idx[i]; }
}

// G lanes per (node, frame) record: pcrs_local[frame].push(p_body), pcr_add.push(p_world)   (voxel_map.hpp:988-989, loop_refine.hpp:318-320, 383-385).
// The point loads are gathers through the sort permutation either way, so short records (deep octree layers, sparse scans)
// are handled by narrow groups — down to one thread per record — instead of wasting a warp on two points.
template <int G>
__global__ void __launch_bounds__(256) k_rec_clusters(PointSrc s, const unsigned int* __restrict__ idx, const unsigned int* __restrict__ rec_start,
                                                      const unsigned long long* __restrict__ rec_key, unsigned int R, int FB, double* __restrict__ rec_local,
                                                      double* __restrict__ rec_world, size_t Rcap) {
  const int lane = threadIdx.x & (G - 1);
  const unsigned int group = (blockIdx.x * blockDim.x + threadIdx.x) / G, ngroups = (gridDim.x * blockDim.x) / G;
  const unsigned int iters = (R + ngroups - 1) / ngroups;
  for (unsigned int it = 0; it < iters; it++) {
    const unsigned int r = group + it * ngroups;
    const bool valid = r < R;
    unsigned int beg = 0, end = 0; int fr = 0;
    if (valid) { beg = rec_start[r]; end = rec_start[r + 1]; fr = int(rec_key[r] & ((1ull << FB) - 1ull)); }
    if (valid && s.win_size > 0 && end > beg) fr = frame_of(s, idx[beg]);        // batch mode: the key holds the slot, the (window, slot) frame follows from any point of the record
    const double* pose = s.poses + 12 * fr;
    double a[20];
#pragma unroll
    for (int k = 0; k < 20; k++) a[k] = 0.0;
    for (unsigned int j = beg + lane; j < end; j += G) {
      const d3 p = load_point(s, idx[j], fr);
      const d3 w = world_point(pose, p);
      a[0] += p.x * p.x; a[1] += p.x * p.y; a[2] += p.x * p.z; a[3] += p.y * p.y; a[4] += p.y * p.z; a[5] += p.z * p.z; a[6] += p.x; a[7] += p.y; a[8] += p.z; a[9] += 1.0;
      a[10] += w.x * w.x; a[11] += w.x * w.y; a[12] += w.x * w.z; a[13] += w.y * w.y; a[14] += w.y * w.z; a[15] += w.z * w.z; a[16] += w.x; a[17] += w.y; a[18] += w.z; a[19] += 1.0;
    }
    if (G > 1) {
#pragma unroll
      for (int k = 0; k < 20; k++)
#pragma unroll
        for (int off = G / 2; off > 0; off >>= 1) a[k] += __shfl_xor_sync(0xffffffffu, a[k], off);
    }
    if (valid && lane == 0) {
#pragma unroll
      for (int k = 0; k < 10; k++) { rec_local[size_t(k) * Rcap + r] = a[k]; rec_world[size_t(k) * Rcap + r] = a[10 + k]; }
    }
  }
}

struct DecideParams {
  double min_eigen_value; double thre; double min_point;   // for THIS layer
  int layer, max_layer, gba, W, FB;
};
enum { NODE_DEAD = 0, NODE_PLANE = 1, NODE_SUBDIVIDE = 2, NODE_LEAF = 3 };

// one thread per node: OctoTree::recut :1150-1172 / OctreeGBA::recut LR:360-399 decision + tras_opt filter :1312-1314 / LR:370-378
__global__ void __launch_bounds__(128) k_node_decide(DecideParams dp, unsigned int Nn, const unsigned int* __restrict__ node_rec_start, const unsigned long long* __restrict__ rec_key,
                                                     const double* __restrict__ rec_local, const double* __restrict__ rec_world, size_t Rcap,
                                                     const long long* __restrict__ parent_root, const int* __restrict__ parent_path, int* __restrict__ state,
                                                     long long* __restrict__ root, int* __restrict__ path, double* __restrict__ eig, double* __restrict__ sum,
                                                     double* __restrict__ fix, unsigned int* __restrict__ nent, unsigned int* __restrict__ sel, size_t Ncap) {
  const unsigned int nd = blockIdx.x * blockDim.x + threadIdx.x;
  if (nd >= Nn) return;
  const unsigned int rb = node_rec_start[nd], re = node_rec_start[nd + 1];
  cluster S; S.P.xx = S.P.xy = S.P.xz = S.P.yy = S.P.yz = S.P.zz = 0.0; S.v = mk3(0, 0, 0); S.n = 0.0;
  double fx[10];
#pragma unroll
  for (int k = 0; k < 10; k++) fx[k] = 0.0;
  unsigned int nwin = 0;
  const unsigned long long fmask = (1ull << dp.FB) - 1ull;
  for (unsigned int r = rb; r < re; r++) {
    S.P.xx += rec_world[r]; S.P.xy += rec_world[Rcap + r]; S.P.xz += rec_world[2 * Rcap + r]; S.P.yy += rec_world[3 * Rcap + r]; S.P.yz += rec_world[4 * Rcap + r];
    S.P.zz += rec_world[5 * Rcap + r]; S.v.x += rec_world[6 * Rcap + r]; S.v.y += rec_world[7 * Rcap + r]; S.v.z += rec_world[8 * Rcap + r]; S.n += rec_world[9 * Rcap + r];
    if (int(rec_key[r] & fmask) >= dp.W) { for (int k = 0; k < 10; k++) fx[k] += rec_local[size_t(k) * Rcap + r]; }
    else nwin++;
  }
  const unsigned long long nodekey = rec_key[rb] >> dp.FB;
  if (dp.layer == 0) { root[nd] = (long long)nodekey; path[nd] = 0; }
  else { const unsigned int par = (unsigned int)(nodekey >> 3); root[nd] = parent_root[par]; path[nd] = parent_path[par] * 8 + int(nodekey & 7); }
  int st = NODE_DEAD; unsigned int take = 0;
  double w[3] = {0, 0, 0}; d3 u0 = mk3(0, 0, 0), u1 = u0, u2 = u0;
  const bool alive = dp.gba ? (S.n > 10.0) : (S.n > dp.min_point && nwin > 0);
  if (alive) {
    eig3_jacobi(cov_from_sum(S), w, u0, u1, u2);
    const bool plane = (w[0] < dp.min_eigen_value) && (w[0] / w[2] < dp.thre);
    if (plane) {
      st = NODE_PLANE;
      take = (w[0] / w[1] > 0.12) ? 0u : 1u;
      if (dp.gba && nwin <= 1) take = 0u;
    } else st = (dp.layer >= dp.max_layer) ? NODE_LEAF : NODE_SUBDIVIDE;
  }
  state[nd] = st; sel[nd] = take; nent[nd] = take ? nwin : 0u;
  if (take) {
    eig[nd] = w[0]; eig[Ncap + nd] = w[1]; eig[2 * Ncap + nd] = w[2];
    eig[3 * Ncap + nd] = u0.x; eig[4 * Ncap + nd] = u1.x; eig[5 * Ncap + nd] = u2.x; eig[6 * Ncap + nd] = u0.y; eig[7 * Ncap + nd] = u1.y; eig[8 * Ncap + nd] = u2.y;
    eig[9 * Ncap + nd] = u0.z; eig[10 * Ncap + nd] = u1.z; eig[11 * Ncap + nd] = u2.z;
    sum[nd] = S.P.xx; sum[Ncap + nd] = S.P.xy; sum[2 * Ncap + nd] = S.P.xz; sum[3 * Ncap + nd] = S.P.yy; sum[4 * Ncap + nd] = S.P.yz; sum[5 * Ncap + nd] = S.P.zz;
    sum[6 * Ncap + nd] = S.v.x; sum[7 * Ncap + nd] = S.v.y; sum[8 * Ncap + nd] = S.v.z; sum[9 * Ncap + nd] = S.n;
    for (int k = 0; k < 10; k++) fix[size_t(k) * Ncap + nd] = fx[k];
  }
}

struct FactorOut {
  int32_t* ptr; int32_t* frame; int32_t* vox; double* cl; size_t Ecap; double* fix; double* coe; double* eig; double* sum; size_t Vcap;
  long long V0, E0;
  int32_t* vwin; long long cells; int win_size;      // batch mode: window of every voxel; entry frames become window * win_size + slot
};
// LidarFactor::push_voxel for every selected node, straight into the device CSR
__global__ void __launch_bounds__(128) k_emit_factor(FactorOut fo, unsigned int Nn, int W, int FB, const unsigned int* __restrict__ sel, const unsigned int* __restrict__ voff,
                                                     const unsigned int* __restrict__ eoff, const unsigned int* __restrict__ node_rec_start,
                                                     const unsigned long long* __restrict__ rec_key, const double* __restrict__ rec_local, size_t Rcap,
                                                     const double* __restrict__ eig, const double* __restrict__ sum, const double* __restrict__ fix, size_t Ncap,
                                                     const long long* __restrict__ root, const int* __restrict__ path, int layer, long long minx, long long miny, long long minz,
                                                     long long ey, long long ez, vxs_voxel_id* __restrict__ ids) {
  const unsigned int nd = blockIdx.x * blockDim.x + threadIdx.x;
  if (nd >= Nn || !sel[nd]) return;
  const long long v = fo.V0 + voff[nd];
  long long e = fo.E0 + eoff[nd];
  fo.ptr[v] = int32_t(e);
  const unsigned long long fmask = (1ull << FB) - 1ull;
  int fbase = 0;
  if (fo.win_size > 0) { const int win = int(root[nd] / fo.cells); fbase = win * fo.win_size; fo.vwin[v] = win; }
  for (unsigned int r = node_rec_start[nd]; r < node_rec_start[nd + 1]; r++) {
    const int fr = int(rec_key[r] & fmask);
    if (fr >= W) continue;
    fo.frame[e] = fbase + fr; fo.vox[e] = int32_t(v);
    for (int k = 0; k < 10; k++) fo.cl[size_t(k) * fo.Ecap + e] = rec_local[size_t(k) * Rcap + r];
    e++;
  }
  for (int k = 0; k < 10; k++) { fo.fix[size_t(k) * fo.Vcap + v] = fix[size_t(k) * Ncap + nd]; fo.sum[size_t(k) * fo.Vcap + v] = sum[size_t(k) * Ncap + nd]; }
  for (int k = 0; k < 12; k++) fo.eig[size_t(k) * fo.Vcap + v] = eig[size_t(k) * Ncap + nd];
  fo.coe[v] = 1.0;   // voxel_map.hpp:1316, loop_refine.hpp:380
  if (ids) {
    const long long lin = fo.win_size > 0 ? root[nd] % fo.cells : root[nd];
    const long long kz = lin % ez, ky = (lin / ez) % ey, kx = lin / (ez * ey);
    vxs_voxel_id id; id.x = kx + minx; id.y = ky + miny; id.z = kz + minz; id.layer = layer; id.path = path[nd];
    ids[v - fo.V0 + 0] = id;
  }
}
__global__ void k_set_last_ptr(int32_t* ptr, long long V, long long E) { ptr[V] = int32_t(E); }

// points of subdivided nodes -> keys of the next layer: ((node*8 + octant) << FB) | frame       (subdivide :1096-1116 / LR:323-356)
__global__ void k_next_flags(const unsigned int* __restrict__ recflag_ex, const unsigned int* __restrict__ recflag, const unsigned int* __restrict__ node_of_rec,
                             const int* __restrict__ state, size_t m, unsigned int* __restrict__ flag) {
  const size_t j = size_t(blockIdx.x) * blockDim.x + threadIdx.x;
  if (j >= m) return;
  const unsigned int rec = recflag_ex[j] + recflag[j] - 1u;
  flag[j] = state[node_of_rec[rec]] == NODE_SUBDIVIDE ? 1u : 0u;
}
__global__ void k_next_keys(const unsigned long long* __restrict__ keys, const unsigned int* __restrict__ idx, const unsigned int* __restrict__ recflag_ex,
                            const unsigned int* __restrict__ recflag, const unsigned int* __restrict__ node_of_rec, const unsigned int* __restrict__ flag,
                            const unsigned int* __restrict__ pos, const unsigned short* __restrict__ pathbits, size_t m, int FB, int next_layer,
                            unsigned long long* __restrict__ keys_out, unsigned int* __restrict__ idx_out) {
  const size_t j = size_t(blockIdx.x) * blockDim.x + threadIdx.x;
  if (j >= m || !flag[j]) return;
  const unsigned int rec = recflag_ex[j] + recflag[j] - 1u, nd = node_of_rec[rec], pi = idx[j];
  const unsigned int oct = (pathbits[pi] >> (3 * (next_layer - 1))) & 7u;
  keys_out[pos[j]] = ((((unsigned long long)nd << 3) | oct) << FB) | (keys[j] & ((1ull << FB) - 1ull));
  idx_out[pos[j]] = pi;
}


// ------------------------------------------------------------------ driver shared by the local map and the GBA map
// Batch mode (bs != nullptr): `nframes` = nwin * win_size virtual frames; offsets_host = prefix over the virtual frames' point counts, bs->src_off_host = first point
// of every virtual frame's keyframe in the uploaded point array (npts_total points), poses12_host = the nwin * win_size virtual-frame poses, W = win_size.
struct BatchSpec { int nwin, win_size; const int64_t* src_off_host; int64_t npts_total; };
static int build_factor(vxs_ctx* ctx, const vxs_map_params* mp, bool gba, const double* pts_d_host, const float* pts_f_host, int fstride, const int64_t* offsets_host, int nframes,
                        const double* poses12_host, int W, vxs_factor* out, vxs_voxel_id* ids_out, int64_t ids_cap, int64_t* n_out, const BatchSpec* bs = nullptr,
                        const float* pts_f_dev = nullptr /* the float points are already on the device (vxs_hba_pass): no upload */, long long own_lo = -1, long long own_hi = -1,
                        bool routed = false /* pts_f_dev holds float4 {x, y, z, frame} records of the points this rank OWNS (already exchanged by owner) */) {
  if (!ctx || !mp || !out || !offsets_host || !poses12_host || out->ctx != ctx) return VXS_ERR_ARG;
  if (mp->max_layer < 0 || mp->max_layer > 3 || !(mp->voxel_size > 0)) return vxs_fail(ctx, VXS_ERR_ARG, "max_layer must be 0..3 and voxel_size > 0");
  cudaSetDevice(ctx->device);
  VoxScratch* s = scratch(ctx);
  cudaStream_t st = ctx->stream;
  const long long N = offsets_host[nframes];
  if (N >= (1ll << 32)) return vxs_fail(ctx, VXS_ERR_ARG, "more than 2^32 points in one build");
  vxs_factor_clear(out);
  out->W = bs ? bs->nwin * bs->win_size : W;        // batch: entry frames are global (window * win_size + slot)
  out->block_W = bs ? bs->win_size : 0;
  if (n_out) *n_out = 0;
  VXS_CUDA(ctx, s->totals.reserve(16));
  if (N == 0) {
    // a rank that received no point at all (routed multi-GPU build) must still take part in the bounding-box all-reduce its peers run below
    if (ctx->nranks > 1 && !bs && own_lo == 0 && own_hi == 0) {   // the same condition as part_bbox below, at N = 0
      VXS_CUDA(ctx, s->bbox.reserve(6));
      const long long nb[6] = {-LLONG_MAX, -LLONG_MAX, -LLONG_MAX, LLONG_MIN, LLONG_MIN, LLONG_MIN};     // [-min | max] of an empty set: neutral for MAX
      VXS_CUDA(ctx, cudaMemcpyAsync(s->bbox.p, nb, sizeof nb, cudaMemcpyHostToDevice, ctx->stream));
      int rcb = vxs_comm_allreduce_max_i64(ctx, s->bbox.p, 6);
      if (rcb) return rcb;
      VXS_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    }
    return VXS_OK;
  }
  const int FB = bits_for((unsigned long long)(bs ? bs->win_size : nframes));
  // ---- upload
  PointSrc ps; ps.pd = nullptr; ps.pf = nullptr; ps.fstride = fstride; ps.nframes = nframes; ps.n = N; ps.src_off = nullptr; ps.win_size = 0; ps.frame_in_w = routed ? 1 : 0;
  const long long Nup = bs ? bs->npts_total : N;     // points to upload (batch: every keyframe once, although most belong to two windows)
  if (pts_d_host) { VXS_CUDA(ctx, s->pts_d.reserve(size_t(Nup) * 3)); VXS_CUDA(ctx, cudaMemcpyAsync(s->pts_d.p, pts_d_host, size_t(Nup) * 24, cudaMemcpyHostToDevice, st)); ps.pd = s->pts_d.p; }
  else if (pts_f_dev) ps.pf = pts_f_dev;
  else { VXS_CUDA(ctx, s->pts_f.reserve(size_t(Nup) * fstride)); VXS_CUDA(ctx, cudaMemcpyAsync(s->pts_f.p, pts_f_host, size_t(Nup) * fstride * 4, cudaMemcpyHostToDevice, st)); ps.pf = s->pts_f.p; }
  if (bs) {
    VXS_CUDA(ctx, s->src_off.reserve(size_t(nframes)));
    VXS_CUDA(ctx, cudaMemcpyAsync(s->src_off.p, bs->src_off_host, size_t(nframes) * 8, cudaMemcpyHostToDevice, st));
    ps.src_off = s->src_off.p; ps.win_size = bs->win_size;
  }
  VXS_CUDA(ctx, s->poses.reserve(size_t(nframes) * 12));
  VXS_CUDA(ctx, cudaMemcpyAsync(s->poses.p, poses12_host, size_t(bs ? nframes : W) * 96, cudaMemcpyHostToDevice, st));
  if (!bs && nframes > W) { const double ident[12] = {1, 0, 0, 0, 1, 0, 0, 0, 1, 0, 0, 0}; VXS_CUDA(ctx, cudaMemcpyAsync(s->poses.p + size_t(W) * 12, ident, 96, cudaMemcpyHostToDevice, st)); }
  VXS_CUDA(ctx, s->offsets.reserve(size_t(nframes) + 1));
  VXS_CUDA(ctx, cudaMemcpyAsync(s->offsets.p, offsets_host, (size_t(nframes) + 1) * 8, cudaMemcpyHostToDevice, st));
  ps.offsets = s->offsets.p; ps.poses = s->poses.p;
  // ---- bounding box of the root cells
  VXS_CUDA(ctx, s->bbox.reserve(6));
  const long long bb0[6] = {LLONG_MAX, LLONG_MAX, LLONG_MAX, LLONG_MIN, LLONG_MIN, LLONG_MIN};
  VXS_CUDA(ctx, cudaMemcpyAsync(s->bbox.p, bb0, sizeof bb0, cudaMemcpyHostToDevice, st));
  // multi-GPU with a known own share of the points (vxs_hba_pass: the rank's own submaps): every rank boxes its share only and the boxes are combined by one
  // 6-element MAX all-reduce, instead of every rank reading all points a first time just for the box
  const bool part_bbox = ctx->nranks > 1 && !bs && own_lo >= 0 && own_hi >= own_lo && own_hi <= N;
  const long long b0 = part_bbox ? own_lo : 0, b1 = part_bbox ? own_hi : N;
  if (b1 > b0) VXS_LAUNCH(ctx, "k_bbox", k_bbox, std::min<unsigned>(nblk(size_t(b1 - b0), 256), unsigned(ctx->sm_count) * 8), 256, 0, ps, mp->voxel_size, s->bbox.p, b0, b1);
  if (part_bbox) {
    VXS_LAUNCH(ctx, "k_bbox_negate_min", k_bbox_negate_min, 1, 32, 0, s->bbox.p);
    int rcb = vxs_comm_allreduce_max_i64(ctx, s->bbox.p, 6);
    if (rcb) return rcb;
    VXS_LAUNCH(ctx, "k_bbox_negate_min", k_bbox_negate_min, 1, 32, 0, s->bbox.p);
  }
  long long bb[6];
  VXS_CUDA(ctx, cudaMemcpyAsync(bb, s->bbox.p, sizeof bb, cudaMemcpyDeviceToHost, st));
  VXS_CUDA(ctx, cudaStreamSynchronize(st));
  const long double ex = (long double)bb[3] - bb[0] + 1, ey = (long double)bb[4] - bb[1] + 1, ez = (long double)bb[5] - bb[2] + 1;
  const long double nw = bs ? (long double)bs->nwin : 1.0L;
  if (ex * ey * ez * nw >= (long double)(1ull << 40)) return vxs_fail(ctx, VXS_ERR_RANGE, "root-cell bounding box (x windows) exceeds 2^40 cells");
  const long long eyl = (long long)ey, ezl = (long long)ez;
  const long long cells = (long long)(ex * ey * ez);
  const int root_bits = bits_for((unsigned long long)(ex * ey * ez * nw));
  // ---- layer-0 keys
  VXS_CUDA(ctx, s->keysA.reserve(size_t(N))); VXS_CUDA(ctx, s->keysB.reserve(size_t(N)));
  VXS_CUDA(ctx, s->idxA.reserve(size_t(N))); VXS_CUDA(ctx, s->idxB.reserve(size_t(N)));
  VXS_CUDA(ctx, s->pathbits.reserve(size_t(N)));
  const bool sharded = ctx->nranks > 1 && !bs && !routed;      // a batch of whole windows is distributed by window, not by voxel; routed points are all owned
  unsigned int* owned = nullptr;
  if (sharded) { VXS_CUDA(ctx, s->flags.reserve(size_t(N))); VXS_CUDA(ctx, s->scanbuf.reserve(size_t(N))); owned = s->flags.p; }
  VXS_LAUNCH(ctx, "k_point_keys", k_point_keys, nblk(size_t(N), 256), 256, 0, ps, mp->voxel_size, int(mp->max_layer), bb[0], bb[1], bb[2], eyl, ezl, cells, FB, s->keysA.p, s->idxA.p, s->pathbits.p,
             ctx->rank, ctx->nranks, owned);

  size_t m = size_t(N);
  int key_bits = root_bits + FB;
  unsigned long long* kcur = s->keysA.p; unsigned int* vcur = s->idxA.p;
  if (sharded) {  // keep only the points whose root cell this rank owns
    int rc0 = scan_u32(ctx, s, owned, s->scanbuf.p, size_t(N), s->totals.p + 5);
    if (rc0) return rc0;
    VXS_LAUNCH(ctx, "k_compact_owned", k_compact_owned, nblk(size_t(N), 256), 256, 0, s->keysA.p, s->idxA.p, owned, s->scanbuf.p, size_t(N), s->keysB.p, s->idxB.p);
    unsigned int mo = 0;
    VXS_CUDA(ctx, cudaMemcpyAsync(&mo, s->totals.p + 5, 4, cudaMemcpyDeviceToHost, st));
    VXS_CUDA(ctx, cudaStreamSynchronize(st));
    m = mo; kcur = s->keysB.p; vcur = s->idxB.p;
  }
  long long Vtot = 0, Etot = 0;
  s->ids_host.clear();
  for (int layer = 0; layer <= mp->max_layer && m > 0; layer++) {
    const int g = layer & 1;
    unsigned long long* kalt = (kcur == s->keysA.p) ? s->keysB.p : s->keysA.p;
    unsigned int* valt = (vcur == s->idxA.p) ? s->idxB.p : s->idxA.p;
    unsigned long long* ks; unsigned int* vs;
    int rc = radix_sort(ctx, s, kcur, vcur, kalt, valt, m, key_bits, &ks, &vs);
    if (rc) return rc;
    unsigned long long* kfree = (ks == s->keysA.p) ? s->keysB.p : s->keysA.p;   // the buffer not holding the sorted keys
    unsigned int* vfree = (vs == s->idxA.p) ? s->idxB.p : s->idxA.p;
    // records
    VXS_CUDA(ctx, s->flags.reserve(m)); VXS_CUDA(ctx, s->scanbuf.reserve(m));
    VXS_LAUNCH(ctx, "k_flag_heads", k_flag_heads, nblk(m, 256), 256, 0, ks, m, s->flags.p);
    rc = scan_u32(ctx, s, s->flags.p, s->scanbuf.p, m, s->totals.p + 0);
    if (rc) return rc;
    unsigned int R = 0;
    VXS_CUDA(ctx, cudaMemcpyAsync(&R, s->totals.p + 0, 4, cudaMemcpyDeviceToHost, st));
    VXS_CUDA(ctx, cudaStreamSynchronize(st));
    VXS_CUDA(ctx, s->rec_start.reserve(size_t(R) + 1)); VXS_CUDA(ctx, s->rec_key.reserve(size_t(R)));
    VXS_CUDA(ctx, s->node_of_rec.reserve(size_t(R))); VXS_CUDA(ctx, s->rec_node_flag.reserve(size_t(R) * 2));
    VXS_LAUNCH(ctx, "k_write_records", k_write_records, nblk(m, 256), 256, 0, ks, s->flags.p, s->scanbuf.p, m, s->rec_start.p, s->rec_key.p, s->totals.p + 0);
    const size_t Rcap = (size_t(R) + 31) & ~size_t(31);
    VXS_CUDA(ctx, s->rec_local.reserve(Rcap * 10)); VXS_CUDA(ctx, s->rec_world.reserve(Rcap * 10));
    {
      const double avg = double(m) / double(std::max(R, 1u));
      const int G = avg > 48.0 ? 32 : (avg > 12.0 ? 8 : (avg > 3.0 ? 4 : 1));
      const unsigned grid = std::min<unsigned>(nblk(size_t(R) * G, 256), unsigned(ctx->sm_count) * 16);
      if (G == 32) { auto kp = k_rec_clusters<32>; VXS_LAUNCH(ctx, "k_rec_clusters", kp, grid, 256, 0, ps, vs, s->rec_start.p, s->rec_key.p, R, FB, s->rec_local.p, s->rec_world.p, Rcap); }
      else if (G == 8) { auto kp = k_rec_clusters<8>; VXS_LAUNCH(ctx, "k_rec_clusters", kp, grid, 256, 0, ps, vs, s->rec_start.p, s->rec_key.p, R, FB, s->rec_local.p, s->rec_world.p, Rcap); }
      else if (G == 4) { auto kp = k_rec_clusters<4>; VXS_LAUNCH(ctx, "k_rec_clusters", kp, grid, 256, 0, ps, vs, s->rec_start.p, s->rec_key.p, R, FB, s->rec_local.p, s->rec_world.p, Rcap); }
      else { auto kp = k_rec_clusters<1>; VXS_LAUNCH(ctx, "k_rec_clusters", kp, grid, 256, 0, ps, vs, s->rec_start.p, s->rec_key.p, R, FB, s->rec_local.p, s->rec_world.p, Rcap); }
    }
    // nodes
    unsigned int* nflag = s->rec_node_flag.p; unsigned int* nex = s->rec_node_flag.p + R;
    VXS_LAUNCH(ctx, "k_flag_nodes", k_flag_nodes, nblk(R, 256), 256, 0, s->rec_key.p, size_t(R), FB, nflag);
    rc = scan_u32(ctx, s, nflag, nex, R, s->totals.p + 1);
    if (rc) return rc;
    unsigned int Nn = 0;
    VXS_CUDA(ctx, cudaMemcpyAsync(&Nn, s->totals.p + 1, 4, cudaMemcpyDeviceToHost, st));
    VXS_CUDA(ctx, cudaStreamSynchronize(st));
    VXS_CUDA(ctx, s->node_rec_start.reserve(size_t(Nn) + 1));
    VXS_LAUNCH(ctx, "k_write_nodes", k_write_nodes, nblk(R, 256), 256, 0, nflag, nex, size_t(R), s->node_of_rec.p, s->node_rec_start.p, s->totals.p + 1);
    const size_t Ncap = (size_t(Nn) + 31) & ~size_t(31);
    VXS_CUDA(ctx, s->node_state[g].reserve(Ncap)); VXS_CUDA(ctx, s->node_root[g].reserve(Ncap)); VXS_CUDA(ctx, s->node_path[g].reserve(Ncap));
    VXS_CUDA(ctx, s->node_eig.reserve(Ncap * 12)); VXS_CUDA(ctx, s->node_sum.reserve(Ncap * 10)); VXS_CUDA(ctx, s->node_fix.reserve(Ncap * 10));
    VXS_CUDA(ctx, s->node_nent.reserve(Ncap)); VXS_CUDA(ctx, s->node_sel.reserve(Ncap)); VXS_CUDA(ctx, s->node_voff.reserve(Ncap)); VXS_CUDA(ctx, s->node_eoff.reserve(Ncap));
    DecideParams dp;
    dp.min_eigen_value = mp->min_eigen_value; dp.thre = mp->plane_thre[std::min(layer, 3)]; dp.min_point = mp->min_point[std::min(layer, 3)];
    dp.layer = layer; dp.max_layer = mp->max_layer; dp.gba = gba ? 1 : 0; dp.W = W; dp.FB = FB;
    VXS_LAUNCH(ctx, "k_node_decide", k_node_decide, nblk(Nn, 128), 128, 0, dp, Nn, s->node_rec_start.p, s->rec_key.p, s->rec_local.p, s->rec_world.p, Rcap, s->node_root[g ^ 1].p,
               s->node_path[g ^ 1].p, s->node_state[g].p, s->node_root[g].p, s->node_path[g].p, s->node_eig.p, s->node_sum.p, s->node_fix.p, s->node_nent.p, s->node_sel.p, Ncap);
    rc = scan_u32(ctx, s, s->node_sel.p, s->node_voff.p, Nn, s->totals.p + 2);
    if (rc) return rc;
    rc = scan_u32(ctx, s, s->node_nent.p, s->node_eoff.p, Nn, s->totals.p + 3);
    if (rc) return rc;
    // points that go one layer down
    unsigned int nextm = 0;
    unsigned int* nxflag = nullptr; unsigned int* nxpos = nullptr;
    if (layer < mp->max_layer) {
      VXS_CUDA(ctx, s->hist.reserve(m * 2));   // reuse the histogram buffer as [flag | pos]
      nxflag = s->hist.p; nxpos = s->hist.p + m;
      VXS_LAUNCH(ctx, "k_next_flags", k_next_flags, nblk(m, 256), 256, 0, s->scanbuf.p, s->flags.p, s->node_of_rec.p, s->node_state[g].p, m, nxflag);
      rc = scan_u32(ctx, s, nxflag, nxpos, m, s->totals.p + 4);
      if (rc) return rc;
    }
    unsigned int tot[5] = {0, 0, 0, 0, 0};
    VXS_CUDA(ctx, cudaMemcpyAsync(tot, s->totals.p, sizeof tot, cudaMemcpyDeviceToHost, st));
    VXS_CUDA(ctx, cudaStreamSynchronize(st));
    const unsigned int selV = tot[2], selE = tot[3];
    nextm = layer < mp->max_layer ? tot[4] : 0;
    if (selV > 0) {
      rc = vxs_factor_reserve(out, size_t(Vtot + selV), size_t(Etot + selE));
      if (rc) return rc;
      FactorOut fo; fo.ptr = out->ptr; fo.frame = out->frame; fo.vox = out->vox; fo.cl = out->cl; fo.Ecap = out->Ecap; fo.fix = out->fix; fo.coe = out->coe; fo.eig = out->eig;
      fo.sum = out->sum; fo.Vcap = out->Vcap; fo.V0 = Vtot; fo.E0 = Etot;
      fo.vwin = nullptr; fo.cells = cells; fo.win_size = 0;
      if (bs) { VXS_CUDA(ctx, out->vwin.reserve_keep(size_t(Vtot + selV), size_t(Vtot), ctx->stream)); fo.vwin = out->vwin.p; fo.win_size = bs->win_size; }
      vxs_voxel_id* ids_dev = nullptr;
      if (ids_out) { VXS_CUDA(ctx, s->ids.reserve(selV)); ids_dev = s->ids.p; }
      VXS_LAUNCH(ctx, "k_emit_factor", k_emit_factor, nblk(Nn, 128), 128, 0, fo, Nn, W, FB, s->node_sel.p, s->node_voff.p, s->node_eoff.p, s->node_rec_start.p, s->rec_key.p,
                 s->rec_local.p, Rcap, s->node_eig.p, s->node_sum.p, s->node_fix.p, Ncap, s->node_root[g].p, s->node_path[g].p, layer, bb[0], bb[1], bb[2], eyl, ezl, ids_dev);
      if (ids_out) {
        const size_t old = s->ids_host.size();
        s->ids_host.resize(old + selV);
        VXS_CUDA(ctx, cudaMemcpyAsync(s->ids_host.data() + old, ids_dev, size_t(selV) * sizeof(vxs_voxel_id), cudaMemcpyDeviceToHost, st));
        VXS_CUDA(ctx, cudaStreamSynchronize(st));
      }
      Vtot += selV; Etot += selE;
      out->V = Vtot; out->E = Etot;   // keeps a later reserve() from dropping what was emitted
    }
    if (nextm > 0) {
      VXS_LAUNCH(ctx, "k_next_keys", k_next_keys, nblk(m, 256), 256, 0, ks, vs, s->scanbuf.p, s->flags.p, s->node_of_rec.p, nxflag, nxpos, s->pathbits.p, m, FB, layer + 1, kfree, vfree);
      kcur = kfree; vcur = vfree;
      key_bits = bits_for(((unsigned long long)Nn << 3) | 7ull) + FB;
    }
    m = nextm;
  }
  if (Vtot > 0) {
    VXS_LAUNCH(ctx, "k_set_last_ptr", k_set_last_ptr, 1, 1, 0, out->ptr, Vtot, Etot);
    // fix clusters present?  (only when fixed map points were supplied)
    out->has_fix = !bs && nframes > W;
  }
  out->V = Vtot; out->E = Etot;
  VXS_CUDA(ctx, cudaStreamSynchronize(st));
  if (n_out) *n_out = Vtot;
  if (ids_out) { const size_t ncopy = std::min<size_t>(size_t(ids_cap), s->ids_host.size()); if (ncopy) memcpy(ids_out, s->ids_host.data(), ncopy * sizeof(vxs_voxel_id)); }
  return VXS_OK;
}

// ------------------------------------------------------------------ public entry points
// ------------------------------------------------------------------ voxel-grid down-sampling (tools.hpp:201-302)
// down_sampling_voxel keeps one point per occupied cell: the running float mean of the cell's points in input order,
//   pp = (pp*cnt + p) / (cnt + 1)  per coordinate, cnt = 1, 2, ...   (tools.hpp:226-232) — float arithmetic, order dependent, so the
// cell's points are reduced sequentially by one thread in their original order (the LSD radix sort is stable);
// down_sampling_close keeps the input point nearest to the cell's float centroid (first minimum, distances in fp64, start value 100).
// The reference iterates an unordered_map, so its output ORDER is unspecified; here cells come out in ascending (x, y, z) cell order and
// every output carries the index of the first input point of its cell (the reference copies that point's other fields).
template <class T>
__global__ void __launch_bounds__(256) k_ds_bbox(const T* __restrict__ pts, int stride, long long n, double voxel_size, long long* __restrict__ bbox) {
  long long mn[3] = {LLONG_MAX, LLONG_MAX, LLONG_MAX}, mx[3] = {LLONG_MIN, LLONG_MIN, LLONG_MIN};
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    const T* p = pts + size_t(i) * stride;
    for (int a = 0; a < 3; a++) { const long long k = quantise((double)p[a], voxel_size); mn[a] = min(mn[a], k); mx[a] = max(mx[a], k); }
  }
  for (int a = 0; a < 3; a++) {
    for (int off = 16; off > 0; off >>= 1) { mn[a] = min(mn[a], __shfl_xor_sync(0xffffffffu, mn[a], off)); mx[a] = max(mx[a], __shfl_xor_sync(0xffffffffu, mx[a], off)); }
    if ((threadIdx.x & 31) == 0) { atomicMin(bbox + a, mn[a]); atomicMax(bbox + 3 + a, mx[a]); }
  }
}
template <class T>
__global__ void __launch_bounds__(256) k_ds_keys(const T* __restrict__ pts, int stride, long long n, double voxel_size, long long minx, long long miny, long long minz,
                                                 unsigned long long ey, unsigned long long ez, unsigned long long* __restrict__ keys, unsigned int* __restrict__ idx) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const T* p = pts + size_t(i) * stride;
  const unsigned long long x = (unsigned long long)(quantise((double)p[0], voxel_size) - minx), y = (unsigned long long)(quantise((double)p[1], voxel_size) - miny),
                           z = (unsigned long long)(quantise((double)p[2], voxel_size) - minz);
  keys[i] = (x * ey + y) * ez + z;
  idx[i] = (unsigned int)i;
}
// mode 0: running mean (down_sampling_voxel); mode 1: nearest to centroid (down_sampling_close)
__global__ void __launch_bounds__(128) k_ds_reduce(const float* __restrict__ pts, int stride, const unsigned int* __restrict__ idx, const unsigned int* __restrict__ seg_start,
                                                   unsigned int nseg, int mode, float* __restrict__ xyz_out, float* __restrict__ cnt_out, long long* __restrict__ pick_out) {
  const unsigned int sgi = blockIdx.x * blockDim.x + threadIdx.x;
  if (sgi >= nseg) return;
  const unsigned int b = seg_start[sgi], e = seg_start[sgi + 1];
  const float* p0 = pts + size_t(idx[b]) * stride;
  float x = p0[0], y = p0[1], z = p0[2];
  if (mode == 0) {
    float cnt = 1.0f;
    for (unsigned int j = b + 1; j < e; j++) {
      const float* p = pts + size_t(idx[j]) * stride;
      const float c1 = __fadd_rn(cnt, 1.0f);
      x = __fdiv_rn(__fadd_rn(__fmul_rn(x, cnt), p[0]), c1);
      y = __fdiv_rn(__fadd_rn(__fmul_rn(y, cnt), p[1]), c1);
      z = __fdiv_rn(__fadd_rn(__fmul_rn(z, cnt), p[2]), c1);
      cnt = c1;
    }
    xyz_out[3 * size_t(sgi)] = x; xyz_out[3 * size_t(sgi) + 1] = y; xyz_out[3 * size_t(sgi) + 2] = z;
    cnt_out[sgi] = cnt;
    pick_out[sgi] = (long long)idx[b];
  } else {
    for (unsigned int j = b + 1; j < e; j++) {
      const float* p = pts + size_t(idx[j]) * stride;
      x = __fadd_rn(x, p[0]); y = __fadd_rn(y, p[1]); z = __fadd_rn(z, p[2]);
    }
    const float fn = (float)(int)(e - b);
    x = __fdiv_rn(x, fn); y = __fdiv_rn(y, fn); z = __fdiv_rn(z, fn);
    double ndis = 100.0;
    unsigned int best = b;
    for (unsigned int j = b; j < e; j++) {
      const float* p = pts + size_t(idx[j]) * stride;
      const double xx = (double)__fsub_rn(x, p[0]), yy = (double)__fsub_rn(y, p[1]), zz = (double)__fsub_rn(z, p[2]);
      const double dis = __dadd_rn(__dadd_rn(__dmul_rn(xx, xx), __dmul_rn(yy, yy)), __dmul_rn(zz, zz));
      if (dis < ndis) { best = j; ndis = dis; }
    }
    const float* pb = pts + size_t(idx[best]) * stride;
    xyz_out[3 * size_t(sgi)] = pb[0]; xyz_out[3 * size_t(sgi) + 1] = pb[1]; xyz_out[3 * size_t(sgi) + 2] = pb[2];
    cnt_out[sgi] = fn;
    pick_out[sgi] = (long long)idx[best];
  }
}

// down_sampling_pvec (voxel_map.hpp:23-64): running fp64 mean of pnt and of var per cell, in input order; the cloud that comes out keeps
// float(pnt) and float(diag(var)) — the recurrences are elementwise, so the three diagonal entries are all that has to be carried.
// pv: pointVar records, `stride` doubles apart, pnt at [0..2], var (3x3, symmetric) at [3..11].
__global__ void __launch_bounds__(128) k_ds_reduce_pvec(const double* __restrict__ pv, int stride, const unsigned int* __restrict__ idx, const unsigned int* __restrict__ seg_start,
                                                        unsigned int nseg, float* __restrict__ xyz_out, float* __restrict__ cnt_out, long long* __restrict__ pick_out, float* __restrict__ nrm_out) {
  const unsigned int sgi = blockIdx.x * blockDim.x + threadIdx.x;
  if (sgi >= nseg) return;
  const unsigned int b = seg_start[sgi], e = seg_start[sgi + 1];
  const double* p0 = pv + size_t(idx[b]) * stride;
  double m[6] = {p0[0], p0[1], p0[2], p0[3], p0[7], p0[11]};
  int cnt = 1;
  for (unsigned int j = b + 1; j < e; j++) {
    const double* p = pv + size_t(idx[j]) * stride;
    const double q[6] = {p[0], p[1], p[2], p[3], p[7], p[11]};
    const double c = (double)cnt, c1 = (double)(cnt + 1);
#pragma unroll
    for (int k = 0; k < 6; k++) m[k] = __ddiv_rn(__dadd_rn(__dmul_rn(m[k], c), q[k]), c1);
    cnt++;
  }
  for (int k = 0; k < 3; k++) { xyz_out[3 * size_t(sgi) + k] = __double2float_rn(m[k]); nrm_out[3 * size_t(sgi) + k] = __double2float_rn(m[3 + k]); }
  cnt_out[sgi] = (float)cnt;
  pick_out[sgi] = (long long)idx[b];
}

// pts_dev: device-resident cloud (n points, stride floats)
template <class T>   // T = float (modes 0, 1) or double (mode 2, pointVar records; nrm_out receives diag(var))
static int down_sample_dev(vxs_ctx* ctx, int mode, const T* pts_dev, int stride, int64_t n, double voxel_size, float* xyz_out, float* count_out, int64_t* index_out,
                           int64_t cap, int64_t* n_out, float* nrm_out = nullptr) {
  VoxScratch* s = scratch(ctx);
  cudaStream_t st = ctx->stream;
  VXS_CUDA(ctx, s->totals.reserve(16));
  VXS_CUDA(ctx, s->bbox.reserve(6));
  const long long bb0[6] = {LLONG_MAX, LLONG_MAX, LLONG_MAX, LLONG_MIN, LLONG_MIN, LLONG_MIN};
  VXS_CUDA(ctx, cudaMemcpyAsync(s->bbox.p, bb0, sizeof bb0, cudaMemcpyHostToDevice, st));
  { auto kb = k_ds_bbox<T>; VXS_LAUNCH(ctx, "k_ds_bbox", kb, std::min<unsigned>(nblk(size_t(n), 256), unsigned(ctx->sm_count) * 8), 256, 0, pts_dev, stride, (long long)n, voxel_size, s->bbox.p); }
  long long bb[6];
  VXS_CUDA(ctx, cudaMemcpyAsync(bb, s->bbox.p, sizeof bb, cudaMemcpyDeviceToHost, st));
  VXS_CUDA(ctx, cudaStreamSynchronize(st));
  const long double ex = (long double)bb[3] - bb[0] + 1, ey = (long double)bb[4] - bb[1] + 1, ez = (long double)bb[5] - bb[2] + 1;
  if (ex * ey * ez >= (long double)(1ull << 62)) return vxs_fail(ctx, VXS_ERR_RANGE, "down-sampling grid exceeds 2^62 cells");
  const int key_bits = bits_for((unsigned long long)(ex * ey * ez));
  VXS_CUDA(ctx, s->keysA.reserve(size_t(n))); VXS_CUDA(ctx, s->keysB.reserve(size_t(n)));
  VXS_CUDA(ctx, s->idxA.reserve(size_t(n))); VXS_CUDA(ctx, s->idxB.reserve(size_t(n)));
  { auto kk = k_ds_keys<T>; VXS_LAUNCH(ctx, "k_ds_keys", kk, nblk(size_t(n), 256), 256, 0, pts_dev, stride, (long long)n, voxel_size, bb[0], bb[1], bb[2], (unsigned long long)ey, (unsigned long long)ez,
             s->keysA.p, s->idxA.p); }
  unsigned long long* ks; unsigned int* vs;
  int rc = radix_sort(ctx, s, s->keysA.p, s->idxA.p, s->keysB.p, s->idxB.p, size_t(n), key_bits, &ks, &vs);
  if (rc) return rc;
  VXS_CUDA(ctx, s->flags.reserve(size_t(n))); VXS_CUDA(ctx, s->scanbuf.reserve(size_t(n)));
  VXS_LAUNCH(ctx, "k_flag_heads", k_flag_heads, nblk(size_t(n), 256), 256, 0, ks, size_t(n), s->flags.p);
  rc = scan_u32(ctx, s, s->flags.p, s->scanbuf.p, size_t(n), s->totals.p + 0);
  if (rc) return rc;
  unsigned int R = 0;
  VXS_CUDA(ctx, cudaMemcpyAsync(&R, s->totals.p + 0, 4, cudaMemcpyDeviceToHost, st));
  VXS_CUDA(ctx, cudaStreamSynchronize(st));
  VXS_CUDA(ctx, s->rec_start.reserve(size_t(R) + 1)); VXS_CUDA(ctx, s->rec_key.reserve(size_t(R)));
  VXS_LAUNCH(ctx, "k_write_records", k_write_records, nblk(size_t(n), 256), 256, 0, ks, s->flags.p, s->scanbuf.p, size_t(n), s->rec_start.p, s->rec_key.p, s->totals.p + 0);
  // outputs: reuse the cluster scratch (4 floats + 1 int64 per cell)
  VXS_CUDA(ctx, s->rec_local.reserve(size_t(R) * 4 + 4)); VXS_CUDA(ctx, s->rec_world.reserve(size_t(R) + 1));
  float* d_xyz = reinterpret_cast<float*>(s->rec_local.p); float* d_cnt = d_xyz + 3 * size_t(R); float* d_nrm = d_cnt + size_t(R);
  long long* d_pick = reinterpret_cast<long long*>(s->rec_world.p);
  if constexpr (sizeof(T) == 8) {
    VXS_LAUNCH(ctx, "k_ds_reduce_pvec", k_ds_reduce_pvec, nblk(size_t(R), 128), 128, 0, pts_dev, stride, vs, s->rec_start.p, R, d_xyz, d_cnt, d_pick, d_nrm);
  } else {
    VXS_LAUNCH(ctx, "k_ds_reduce", k_ds_reduce, nblk(size_t(R), 128), 128, 0, pts_dev, stride, vs, s->rec_start.p, R, mode, d_xyz, d_cnt, d_pick);
  }
  *n_out = int64_t(R);
  const size_t ncopy = size_t(std::min<int64_t>(cap, int64_t(R)));
  if (xyz_out && ncopy) VXS_CUDA(ctx, cudaMemcpyAsync(xyz_out, d_xyz, ncopy * 12, cudaMemcpyDeviceToHost, st));
  if (count_out && ncopy) VXS_CUDA(ctx, cudaMemcpyAsync(count_out, d_cnt, ncopy * 4, cudaMemcpyDeviceToHost, st));
  if (index_out && ncopy) VXS_CUDA(ctx, cudaMemcpyAsync(index_out, d_pick, ncopy * 8, cudaMemcpyDeviceToHost, st));
  if (nrm_out && ncopy) VXS_CUDA(ctx, cudaMemcpyAsync(nrm_out, d_nrm, ncopy * 12, cudaMemcpyDeviceToHost, st));
  VXS_CUDA(ctx, cudaStreamSynchronize(st));
  return VXS_OK;
}
static int down_sample(vxs_ctx* ctx, int mode, const float* pts_host, int stride, int64_t n, double voxel_size, float* xyz_out, float* count_out, int64_t* index_out,
                       int64_t cap, int64_t* n_out) {
  if (!ctx || !n_out || n < 0 || stride < 3 || (n > 0 && !pts_host)) return VXS_ERR_ARG;
  *n_out = -1;
  if (voxel_size < 0.001) return VXS_OK;                 // tools.hpp:203 / 247: the cloud is left untouched
  *n_out = 0;
  if (n == 0) return VXS_OK;
  if (n >= (1ll << 32)) return vxs_fail(ctx, VXS_ERR_ARG, "more than 2^32 points in one down-sampling call");
  cudaSetDevice(ctx->device);
  VoxScratch* s = scratch(ctx);
  VXS_CUDA(ctx, s->pts_f.reserve(size_t(n) * stride));
  VXS_CUDA(ctx, cudaMemcpyAsync(s->pts_f.p, pts_host, size_t(n) * stride * 4, cudaMemcpyHostToDevice, ctx->stream));
  return down_sample_dev(ctx, mode, s->pts_f.p, stride, n, voxel_size, xyz_out, count_out, index_out, cap, n_out);
}

extern "C" int vxs_down_sampling_pvec(vxs_ctx* ctx, const double* pv, int stride_doubles, int64_t n, double voxel_size, float* xyz_out, float* var_diag_out, float* count_out,
                                      int64_t* first_index_out, int64_t cap, int64_t* n_out) {
  if (!ctx || !n_out || n < 0 || stride_doubles < 12 || (n > 0 && !pv) || !(voxel_size > 0)) return VXS_ERR_ARG;
  *n_out = 0;
  if (n == 0) return VXS_OK;
  if (n >= (1ll << 32)) return vxs_fail(ctx, VXS_ERR_ARG, "more than 2^32 points in one down-sampling call");
  cudaSetDevice(ctx->device);
  VoxScratch* s = scratch(ctx);
  VXS_CUDA(ctx, s->pts_d.reserve(size_t(n) * stride_doubles));
  VXS_CUDA(ctx, cudaMemcpyAsync(s->pts_d.p, pv, size_t(n) * stride_doubles * 8, cudaMemcpyHostToDevice, ctx->stream));
  return down_sample_dev<double>(ctx, 2, s->pts_d.p, stride_doubles, n, voxel_size, xyz_out, count_out, first_index_out, cap, n_out, var_diag_out);
}

// ------------------------------------------------------------------ submap merge of HBA_add_edge (voxelslam.cpp:2428-2447)
// every keyframe cloud is moved into the frame of keyframe 0: v' = dR v + dp in fp64, stored as float, dR = R_0^T R_i, dp = R_0^T (p_i - p_0)
__global__ void __launch_bounds__(256) k_merge_transform(const float* __restrict__ pts, int stride, const long long* __restrict__ offsets, int W, const double* __restrict__ rel12,
                                                         long long n, float* __restrict__ out) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  int lo = 0, hi = W;
  while (hi - lo > 1) { const int mid = (lo + hi) >> 1; if (offsets[mid] <= i) lo = mid; else hi = mid; }
  const double* T = rel12 + 12 * lo;
  const float* p = pts + size_t(i) * stride;
  const double x = (double)p[0], y = (double)p[1], z = (double)p[2];
  out[3 * i] = __double2float_rn(dot3_rn(T[0], T[1], T[2], x, y, z, T[9]));
  out[3 * i + 1] = __double2float_rn(dot3_rn(T[3], T[4], T[5], x, y, z, T[10]));
  out[3 * i + 2] = __double2float_rn(dot3_rn(T[6], T[7], T[8], x, y, z, T[11]));
}
extern "C" int vxs_submap_merge(vxs_ctx* ctx, const float* xyz, int stride_floats, const int64_t* kf_offsets, const double* poses12, int W, double voxel_size,
                                float* xyz_out, float* count_out, int64_t* first_index_out, int64_t cap, int64_t* n_out) {
  if (!ctx || !n_out || !kf_offsets || !poses12 || W <= 0 || stride_floats < 3) return VXS_ERR_ARG;
  const int64_t n = kf_offsets[W];
  if (n < 0 || (n > 0 && !xyz)) return VXS_ERR_ARG;
  *n_out = 0;
  if (n == 0) return VXS_OK;
  if (n >= (1ll << 32)) return vxs_fail(ctx, VXS_ERR_ARG, "more than 2^32 points in one submap merge");
  cudaSetDevice(ctx->device);
  VoxScratch* s = scratch(ctx);
  cudaStream_t st = ctx->stream;
  // relative poses on the host (W x 12, a few hundred flops): dR = R_0^T R_i, dp = R_0^T (p_i - p_0), sums in index order
  std::vector<double> rel(size_t(W) * 12);
  const double* P0 = poses12;
  for (int i = 0; i < W; i++) {
    const double* Pi = poses12 + 12 * size_t(i);
    const double d[3] = {Pi[9] - P0[9], Pi[10] - P0[10], Pi[11] - P0[11]};
    for (int r = 0; r < 3; r++) {
      for (int c = 0; c < 3; c++) rel[12 * size_t(i) + 3 * r + c] = (P0[r] * Pi[c] + P0[3 + r] * Pi[3 + c]) + P0[6 + r] * Pi[6 + c];
      rel[12 * size_t(i) + 9 + r] = (P0[r] * d[0] + P0[3 + r] * d[1]) + P0[6 + r] * d[2];
    }
  }
  VXS_CUDA(ctx, s->pts_f.reserve(size_t(n) * stride_floats));
  VXS_CUDA(ctx, cudaMemcpyAsync(s->pts_f.p, xyz, size_t(n) * stride_floats * 4, cudaMemcpyHostToDevice, st));
  VXS_CUDA(ctx, s->poses.reserve(size_t(W) * 12));
  VXS_CUDA(ctx, cudaMemcpyAsync(s->poses.p, rel.data(), rel.size() * 8, cudaMemcpyHostToDevice, st));
  VXS_CUDA(ctx, s->offsets.reserve(size_t(W) + 1));
  VXS_CUDA(ctx, cudaMemcpyAsync(s->offsets.p, kf_offsets, (size_t(W) + 1) * 8, cudaMemcpyHostToDevice, st));
  VXS_CUDA(ctx, s->pts_f2.reserve(size_t(n) * 3));
  VXS_LAUNCH(ctx, "k_merge_transform", k_merge_transform, nblk(size_t(n), 256), 256, 0, s->pts_f.p, stride_floats, s->offsets.p, W, s->poses.p, (long long)n, s->pts_f2.p);
  VXS_CUDA(ctx, cudaStreamSynchronize(st));   // rel goes out of scope
  if (voxel_size < 0.001) {                   // down_sampling_voxel would return at once: the merged cloud itself is the result
    *n_out = n;
    const size_t ncopy = size_t(std::min<int64_t>(cap, n));
    if (xyz_out && ncopy) VXS_CUDA(ctx, cudaMemcpyAsync(xyz_out, s->pts_f2.p, ncopy * 12, cudaMemcpyDeviceToHost, st));
    if (count_out) for (size_t i = 0; i < ncopy; i++) count_out[i] = 0.0f;
    if (first_index_out) for (size_t i = 0; i < ncopy; i++) first_index_out[i] = int64_t(i);
    VXS_CUDA(ctx, cudaStreamSynchronize(st));
    return VXS_OK;
  }
  return down_sample_dev(ctx, 0, s->pts_f2.p, 3, n, voxel_size, xyz_out, count_out, first_index_out, cap, n_out);
}

// ------------------------------------------------------------------ submap merge of MANY windows at once (the post-step of every bottom-level HBA_add_edge, voxelslam.cpp:2428-2447)
// Same arithmetic as vxs_submap_merge per window; the windows only share the sort: a point's key is (window, cell), the stable sort keeps the
// keyframe-by-keyframe input order inside a cell (the running float mean is order dependent), and the cells come out grouped by window.
__global__ void __launch_bounds__(256) k_merge_transform_batch(const float* __restrict__ pts, int stride, const long long* __restrict__ voff, const long long* __restrict__ soff, int nf,
                                                               const double* __restrict__ rel12, long long n, float* __restrict__ out, int* __restrict__ vf_out) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  int lo = 0, hi = nf;
  while (hi - lo > 1) { const int mid = (lo + hi) >> 1; if (voff[mid] <= i) lo = mid; else hi = mid; }
  const double* T = rel12 + 12 * lo;
  const float* p = pts + size_t(soff[lo] + (i - voff[lo])) * stride;
  const double x = (double)p[0], y = (double)p[1], z = (double)p[2];
  out[3 * i] = __double2float_rn(dot3_rn(T[0], T[1], T[2], x, y, z, T[9]));
  out[3 * i + 1] = __double2float_rn(dot3_rn(T[3], T[4], T[5], x, y, z, T[10]));
  out[3 * i + 2] = __double2float_rn(dot3_rn(T[6], T[7], T[8], x, y, z, T[11]));
  vf_out[i] = lo;
}
__global__ void __launch_bounds__(256) k_ds_keys_batch(const float* __restrict__ pts, const int* __restrict__ vf, int win_size, long long n, double voxel_size, long long minx, long long miny,
                                                       long long minz, unsigned long long ey, unsigned long long ez, unsigned long long cells, unsigned long long* __restrict__ keys,
                                                       unsigned int* __restrict__ idx) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const float* p = pts + 3 * size_t(i);
  const unsigned long long x = (unsigned long long)(quantise((double)p[0], voxel_size) - minx), y = (unsigned long long)(quantise((double)p[1], voxel_size) - miny),
                           z = (unsigned long long)(quantise((double)p[2], voxel_size) - minz);
  keys[i] = (unsigned long long)(vf[i] / win_size) * cells + (x * ey + y) * ez + z;
  idx[i] = (unsigned int)i;
}
__global__ void __launch_bounds__(256) k_merge_win_ptr(const unsigned long long* __restrict__ rec_key, unsigned int R, unsigned long long cells, int nwin, long long* __restrict__ win_ptr) {
  const int w = blockIdx.x * blockDim.x + threadIdx.x;
  if (w > nwin) return;
  unsigned int lo = 0, hi = R;
  const unsigned long long want = (unsigned long long)w * cells;
  while (lo < hi) { const unsigned int mid = (lo + hi) >> 1; if (rec_key[mid] < want) lo = mid + 1; else hi = mid; }
  win_ptr[w] = (long long)lo;
}
__global__ void __launch_bounds__(256) k_merge_local_index(long long* __restrict__ pick, const unsigned long long* __restrict__ rec_key, unsigned int R, unsigned long long cells, int win_size,
                                                           const long long* __restrict__ voff) {
  const unsigned int r = blockIdx.x * blockDim.x + threadIdx.x;
  if (r < R) pick[r] -= voff[size_t(rec_key[r] / cells) * win_size];     // index inside the window's concatenated clouds
}

// xyz_dev / dev_first_point: device copy of the input points (or NULL: upload from xyz).  dev_out: when given the merged points of all windows are appended to it on the
// device (3 floats each) instead of being copied to the host arrays.
int vxs_submap_merge_batch_impl(vxs_ctx* ctx, const float* xyz, const float* xyz_dev, int64_t dev_first_point, int stride_floats, const int64_t* kf_offsets, int K, const double* poses_win,
                                const int32_t* win_first, int nwin, int win_size, double voxel_size, int64_t max_points_per_chunk, float* xyz_out, float* count_out,
                                int64_t* first_index_out, int64_t cap, int64_t* win_offsets, int64_t* n_out, DevBuf<float>* dev_out) {
  if (!ctx || (!xyz && !xyz_dev) || !kf_offsets || !poses_win || !win_first || !win_offsets || !n_out || K <= 0 || nwin <= 0 || win_size <= 0 || stride_floats < 3 || !(voxel_size >= 0.001)) return VXS_ERR_ARG;
  for (int w = 0; w < nwin; w++) if (win_first[w] < 0 || win_first[w] + win_size > K) return vxs_fail(ctx, VXS_ERR_ARG, "vxs_submap_merge_batch: a window reaches beyond the keyframes");
  cudaSetDevice(ctx->device);
  VoxScratch* s = scratch(ctx);
  cudaStream_t st = ctx->stream;
  if (max_points_per_chunk <= 0) max_points_per_chunk = 96ll << 20;
  VXS_CUDA(ctx, s->totals.reserve(16)); VXS_CUDA(ctx, s->bbox.reserve(6));
  int64_t total = 0;
  win_offsets[0] = 0;
  for (int w0 = 0; w0 < nwin;) {
    int w1 = w0; int64_t vp = 0;
    while (w1 < nwin) {
      const int64_t np = kf_offsets[win_first[w1] + win_size] - kf_offsets[win_first[w1]];
      if (w1 > w0 && vp + np > max_points_per_chunk) break;
      vp += np; w1++;
    }
    const int nw = w1 - w0, nf = nw * win_size;
    int kf_lo = K, kf_hi = 0;
    for (int w = w0; w < w1; w++) { kf_lo = std::min(kf_lo, int(win_first[w])); kf_hi = std::max(kf_hi, int(win_first[w]) + win_size); }
    const int64_t base = kf_offsets[kf_lo], nup = kf_offsets[kf_hi] - base;
    std::vector<long long> voff(size_t(nf) + 1, 0), soff(size_t(nf), 0);
    std::vector<double> rel(size_t(nf) * 12);
    for (int w = 0; w < nw; w++) {
      const double* P0 = poses_win + size_t(w0 + w) * win_size * 12;
      for (int j = 0; j < win_size; j++) {
        const int kf = win_first[w0 + w] + j, vf = w * win_size + j;
        voff[size_t(vf) + 1] = voff[size_t(vf)] + (kf_offsets[kf + 1] - kf_offsets[kf]);
        soff[size_t(vf)] = kf_offsets[kf] - base;
        const double* Pi = P0 + 12 * size_t(j);
        const double d[3] = {Pi[9] - P0[9], Pi[10] - P0[10], Pi[11] - P0[11]};
        for (int r = 0; r < 3; r++) {
          for (int c = 0; c < 3; c++) rel[12 * size_t(vf) + 3 * r + c] = (P0[r] * Pi[c] + P0[3 + r] * Pi[3 + c]) + P0[6 + r] * Pi[6 + c];
          rel[12 * size_t(vf) + 9 + r] = (P0[r] * d[0] + P0[3 + r] * d[1]) + P0[6 + r] * d[2];
        }
      }
    }
    const long long n = voff[size_t(nf)];
    if (n >= (1ll << 32)) return vxs_fail(ctx, VXS_ERR_ARG, "more than 2^32 points in one merge chunk");
    long long R = 0;
    if (n > 0) {
      const float* in_dev = nullptr;
      if (xyz_dev) in_dev = xyz_dev + size_t(base - dev_first_point) * stride_floats;
      else {
        VXS_CUDA(ctx, s->pts_f.reserve(size_t(nup) * stride_floats));
        VXS_CUDA(ctx, cudaMemcpyAsync(s->pts_f.p, xyz + size_t(base) * stride_floats, size_t(nup) * stride_floats * 4, cudaMemcpyHostToDevice, st));
        in_dev = s->pts_f.p;
      }
      VXS_CUDA(ctx, s->poses.reserve(size_t(nf) * 12)); VXS_CUDA(ctx, s->offsets.reserve(size_t(nf) + 1)); VXS_CUDA(ctx, s->src_off.reserve(size_t(nf)));
      VXS_CUDA(ctx, cudaMemcpyAsync(s->poses.p, rel.data(), rel.size() * 8, cudaMemcpyHostToDevice, st));
      VXS_CUDA(ctx, cudaMemcpyAsync(s->offsets.p, voff.data(), voff.size() * 8, cudaMemcpyHostToDevice, st));
      VXS_CUDA(ctx, cudaMemcpyAsync(s->src_off.p, soff.data(), soff.size() * 8, cudaMemcpyHostToDevice, st));
      VXS_CUDA(ctx, s->pts_f2.reserve(size_t(n) * 3)); VXS_CUDA(ctx, s->flags.reserve(size_t(n))); VXS_CUDA(ctx, s->scanbuf.reserve(size_t(n)));
      int* vf_dev = reinterpret_cast<int*>(s->scanbuf.p);       // (window, slot) of every virtual point; consumed by the key kernel before the scan reuses the buffer
      VXS_LAUNCH(ctx, "k_merge_transform", k_merge_transform_batch, nblk(size_t(n), 256), 256, 0, in_dev, stride_floats, s->offsets.p, s->src_off.p, nf, s->poses.p, n, s->pts_f2.p, vf_dev);
      const long long bb0[6] = {LLONG_MAX, LLONG_MAX, LLONG_MAX, LLONG_MIN, LLONG_MIN, LLONG_MIN};
      VXS_CUDA(ctx, cudaMemcpyAsync(s->bbox.p, bb0, sizeof bb0, cudaMemcpyHostToDevice, st));
      { auto kb = k_ds_bbox<float>; VXS_LAUNCH(ctx, "k_ds_bbox", kb, std::min<unsigned>(nblk(size_t(n), 256), unsigned(ctx->sm_count) * 8), 256, 0, s->pts_f2.p, 3, n, voxel_size, s->bbox.p); }
      long long bb[6];
      VXS_CUDA(ctx, cudaMemcpyAsync(bb, s->bbox.p, sizeof bb, cudaMemcpyDeviceToHost, st));
      VXS_CUDA(ctx, cudaStreamSynchronize(st));      // also: the host vectors above are no longer read
      const long double ex = (long double)bb[3] - bb[0] + 1, ey = (long double)bb[4] - bb[1] + 1, ez = (long double)bb[5] - bb[2] + 1;
      if (ex * ey * ez * nw >= (long double)(1ull << 62)) return vxs_fail(ctx, VXS_ERR_RANGE, "merge grid exceeds 2^62 cells");
      const unsigned long long cells = (unsigned long long)(ex * ey * ez);
      const int key_bits = bits_for((unsigned long long)(ex * ey * ez * nw));
      VXS_CUDA(ctx, s->keysA.reserve(size_t(n))); VXS_CUDA(ctx, s->keysB.reserve(size_t(n))); VXS_CUDA(ctx, s->idxA.reserve(size_t(n))); VXS_CUDA(ctx, s->idxB.reserve(size_t(n)));
      VXS_LAUNCH(ctx, "k_ds_keys", k_ds_keys_batch, nblk(size_t(n), 256), 256, 0, s->pts_f2.p, vf_dev, win_size, n, voxel_size, bb[0], bb[1], bb[2], (unsigned long long)ey, (unsigned long long)ez,
                 cells, s->keysA.p, s->idxA.p);
      unsigned long long* ks; unsigned int* vs;
      int rc = radix_sort(ctx, s, s->keysA.p, s->idxA.p, s->keysB.p, s->idxB.p, size_t(n), key_bits, &ks, &vs);
      if (rc) return rc;
      VXS_LAUNCH(ctx, "k_flag_heads", k_flag_heads, nblk(size_t(n), 256), 256, 0, ks, size_t(n), s->flags.p);
      rc = scan_u32(ctx, s, s->flags.p, s->scanbuf.p, size_t(n), s->totals.p + 0);
      if (rc) return rc;
      unsigned int Ru = 0;
      VXS_CUDA(ctx, cudaMemcpyAsync(&Ru, s->totals.p + 0, 4, cudaMemcpyDeviceToHost, st));
      VXS_CUDA(ctx, cudaStreamSynchronize(st));
      R = Ru;
      VXS_CUDA(ctx, s->rec_start.reserve(size_t(R) + 1)); VXS_CUDA(ctx, s->rec_key.reserve(size_t(R)));
      VXS_LAUNCH(ctx, "k_write_records", k_write_records, nblk(size_t(n), 256), 256, 0, ks, s->flags.p, s->scanbuf.p, size_t(n), s->rec_start.p, s->rec_key.p, s->totals.p + 0);
      VXS_CUDA(ctx, s->rec_local.reserve(size_t(R) * 4 + 4)); VXS_CUDA(ctx, s->rec_world.reserve(size_t(R) + size_t(nw) + 2));
      float* d_xyz = reinterpret_cast<float*>(s->rec_local.p); float* d_cnt = d_xyz + 3 * size_t(R);
      long long* d_pick = reinterpret_cast<long long*>(s->rec_world.p); long long* d_wptr = d_pick + R;
      VXS_LAUNCH(ctx, "k_ds_reduce", k_ds_reduce, nblk(size_t(R), 128), 128, 0, s->pts_f2.p, 3, vs, s->rec_start.p, Ru, 0, d_xyz, d_cnt, d_pick);
      VXS_LAUNCH(ctx, "k_merge_win_ptr", k_merge_win_ptr, nblk(size_t(nw) + 1, 256), 256, 0, s->rec_key.p, Ru, cells, nw, d_wptr);
      VXS_LAUNCH(ctx, "k_merge_local_index", k_merge_local_index, nblk(size_t(R), 256), 256, 0, d_pick, s->rec_key.p, Ru, cells, win_size, s->offsets.p);
      std::vector<long long> wp(size_t(nw) + 1);
      VXS_CUDA(ctx, cudaMemcpyAsync(wp.data(), d_wptr, wp.size() * 8, cudaMemcpyDeviceToHost, st));
      if (dev_out && R > 0) {   // keep the merged cloud on the device
        VXS_CUDA(ctx, dev_out->reserve_keep(size_t(total + R) * 3, size_t(total) * 3, st));
        VXS_CUDA(ctx, cudaMemcpyAsync(dev_out->p + size_t(total) * 3, d_xyz, size_t(R) * 12, cudaMemcpyDeviceToDevice, st));
      }
      const int64_t room = dev_out ? 0 : std::max<int64_t>(0, cap - total);
      const size_t ncopy = size_t(std::min<int64_t>(room, R));
      if (xyz_out && ncopy) VXS_CUDA(ctx, cudaMemcpyAsync(xyz_out + size_t(total) * 3, d_xyz, ncopy * 12, cudaMemcpyDeviceToHost, st));
      if (count_out && ncopy) VXS_CUDA(ctx, cudaMemcpyAsync(count_out + size_t(total), d_cnt, ncopy * 4, cudaMemcpyDeviceToHost, st));
      if (first_index_out && ncopy) VXS_CUDA(ctx, cudaMemcpyAsync(first_index_out + size_t(total), d_pick, ncopy * 8, cudaMemcpyDeviceToHost, st));
      VXS_CUDA(ctx, cudaStreamSynchronize(st));
      for (int w = 0; w < nw; w++) win_offsets[w0 + w + 1] = total + wp[size_t(w) + 1];
    } else {
      for (int w = 0; w < nw; w++) win_offsets[w0 + w + 1] = total;
    }
    total += R;
    w0 = w1;
  }
  *n_out = total;
  return VXS_OK;
}
extern "C" int vxs_submap_merge_batch(vxs_ctx* ctx, const float* xyz, int stride_floats, const int64_t* kf_offsets, int K, const double* poses_win, const int32_t* win_first, int nwin,
                                      int win_size, double voxel_size, int64_t max_points_per_chunk, float* xyz_out, float* count_out, int64_t* first_index_out, int64_t cap,
                                      int64_t* win_offsets, int64_t* n_out) {
  if (!xyz) return VXS_ERR_ARG;
  return vxs_submap_merge_batch_impl(ctx, xyz, nullptr, 0, stride_floats, kf_offsets, K, poses_win, win_first, nwin, win_size, voxel_size, max_points_per_chunk, xyz_out, count_out,
                                     first_index_out, cap, win_offsets, n_out, nullptr);
}

extern "C" int vxs_down_sampling_voxel(vxs_ctx* ctx, const float* pts, int stride_floats, int64_t n, double voxel_size, float* xyz_out, float* count_out, int64_t* first_index_out,
                                       int64_t cap, int64_t* n_out) {
  return down_sample(ctx, 0, pts, stride_floats, n, voxel_size, xyz_out, count_out, first_index_out, cap, n_out);
}
extern "C" int vxs_down_sampling_close(vxs_ctx* ctx, const float* pts, int stride_floats, int64_t n, double voxel_size, float* xyz_out, float* count_out, int64_t* picked_index_out,
                                       int64_t cap, int64_t* n_out) {
  return down_sample(ctx, 1, pts, stride_floats, n, voxel_size, xyz_out, count_out, picked_index_out, cap, n_out);
}

extern "C" int vxs_voxel_keys(vxs_ctx* ctx, const double* pw, int64_t n, double voxel_size, int64_t* xyz, uint64_t* hash) {
  if (!ctx || n < 0 || (n > 0 && (!pw || !xyz || !hash)) || !(voxel_size > 0)) return VXS_ERR_ARG;
  if (n == 0) return VXS_OK;
  cudaSetDevice(ctx->device);
  VoxScratch* s = scratch(ctx);
  VXS_CUDA(ctx, s->pts_d.reserve(size_t(n) * 3)); VXS_CUDA(ctx, s->keysA.reserve(size_t(n) * 3)); VXS_CUDA(ctx, s->keysB.reserve(size_t(n)));
  VXS_CUDA(ctx, cudaMemcpyAsync(s->pts_d.p, pw, size_t(n) * 24, cudaMemcpyHostToDevice, ctx->stream));
  VXS_LAUNCH(ctx, "k_voxel_keys", k_voxel_keys, nblk(size_t(n), 256), 256, 0, s->pts_d.p, (long long)n, voxel_size, (long long*)s->keysA.p, s->keysB.p);
  VXS_CUDA(ctx, cudaMemcpyAsync(xyz, s->keysA.p, size_t(n) * 24, cudaMemcpyDeviceToHost, ctx->stream));
  VXS_CUDA(ctx, cudaMemcpyAsync(hash, s->keysB.p, size_t(n) * 8, cudaMemcpyDeviceToHost, ctx->stream));
  VXS_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  return VXS_OK;
}

extern "C" int vxs_build_window_factor(vxs_ctx* ctx, const vxs_map_params* mp, const double* pts_body, const int64_t* scan_offsets, const double* poses12, int W,
                                       const double* fix_pts, int64_t n_fix, vxs_factor* out, vxs_voxel_id* ids_out, int64_t ids_cap, int64_t* n_out) {
  if (!ctx || !scan_offsets || W <= 0 || (scan_offsets[W] > 0 && !pts_body)) return VXS_ERR_ARG;
  if (!fix_pts || n_fix <= 0) return build_factor(ctx, mp, false, pts_body, nullptr, 3, scan_offsets, W, poses12, W, out, ids_out, ids_cap, n_out);
  // fixed map points ride along as pseudo-frame W (already in world coordinates)
  const int64_t nw = scan_offsets[W];
  std::vector<double> all(size_t(nw + n_fix) * 3);
  if (nw) memcpy(all.data(), pts_body, size_t(nw) * 24);
  memcpy(all.data() + size_t(nw) * 3, fix_pts, size_t(n_fix) * 24);
  std::vector<int64_t> off(scan_offsets, scan_offsets + W + 1);
  off.push_back(nw + n_fix);
  return build_factor(ctx, mp, false, all.data(), nullptr, 3, off.data(), W + 1, poses12, W, out, ids_out, ids_cap, n_out);
}

extern "C" int vxs_build_gba_factor(vxs_ctx* ctx, const vxs_map_params* mp, const float* xyz, int stride_floats, const int64_t* kf_offsets, const double* poses12, int W,
                                    vxs_factor* out, vxs_voxel_id* ids_out, int64_t ids_cap, int64_t* n_out) {
  if (!ctx || !kf_offsets || W <= 0 || stride_floats < 3 || (kf_offsets[W] > 0 && !xyz)) return VXS_ERR_ARG;
  return build_factor(ctx, mp, true, nullptr, xyz, stride_floats, kf_offsets, W, poses12, W, out, ids_out, ids_cap, n_out);
}

// HBA_add_edge BA loop, voxelslam.cpp:2360-2399
// Map build of a chunk of independent windows (vxs_hba_bottom_batch): OctreeGBA::cut_voxel of every (window, keyframe) + OctreeGBA_multi_recut, one pass.
// win_first[w] = first keyframe of window w; keyframes [kf_lo, kf_hi) cover the chunk and are uploaded once.
int vxs_build_gba_batch(vxs_ctx* ctx, const vxs_map_params* mp, const float* xyz, int stride_floats, const int64_t* kf_offsets, const double* poses12, const int32_t* win_first, int nwin,
                        int win_size, int kf_lo, int kf_hi, vxs_factor* out, const float* xyz_dev, int64_t dev_first_point) {   // xyz_dev: device copy of the points from global index dev_first_point on (or NULL)
  const int nf = nwin * win_size;
  std::vector<int64_t> voff(size_t(nf) + 1, 0), soff(size_t(nf), 0);
  std::vector<double> vposes(size_t(nf) * 12);
  const int64_t base = kf_offsets[kf_lo];
  for (int w = 0; w < nwin; w++)
    for (int j = 0; j < win_size; j++) {
      const int kf = win_first[w] + j, vf = w * win_size + j;
      voff[size_t(vf) + 1] = voff[size_t(vf)] + (kf_offsets[kf + 1] - kf_offsets[kf]);
      soff[size_t(vf)] = kf_offsets[kf] - base;
      memcpy(vposes.data() + size_t(vf) * 12, poses12 + size_t(kf) * 12, 96);
    }
  BatchSpec bs; bs.nwin = nwin; bs.win_size = win_size; bs.src_off_host = soff.data(); bs.npts_total = kf_offsets[kf_hi] - base;
  return build_factor(ctx, mp, true, nullptr, xyz + size_t(base) * stride_floats, stride_floats, voff.data(), nf, vposes.data(), win_size, out, nullptr, 0, nullptr, &bs,
                      xyz_dev ? xyz_dev + size_t(base - dev_first_point) * stride_floats : nullptr);
}

int vxs_hba_window_impl(vxs_ctx* ctx, const vxs_map_params* coarse, const vxs_map_params* fine, const float* xyz, const float* xyz_dev, int stride_floats, const int64_t* kf_offsets,
                        double* poses12, int W, int max_iter, int thread_num, double* hess_out, double* resis_log, int* outer_iters, long long own_lo, long long own_hi) {
  if (!ctx || !coarse || !fine || !poses12 || W <= 0) return VXS_ERR_ARG;
  vxs_map_params gp = *coarse;
  const int up = 4;
  int converge_flag = 0, iters = 0, warn = 0;
  double converge_thre = 0.05;
  // the factor (and its (6W)^2 accumulators) is kept on the ctx between calls: an HBA pass per submap must not pay its cudaMalloc / cudaFree every time
  VoxScratch* vs_ = scratch(ctx);
  int rc = VXS_OK;
  if (!vs_->hba_factor) { rc = vxs_factor_create(ctx, W, &vs_->hba_factor); if (rc) return rc; }
  vxs_factor* f = vs_->hba_factor;
  for (int iterCnt = 0; iterCnt < max_iter; iterCnt++) {
    if (converge_flag == 1 || iterCnt == max_iter - 1) { const int ml = gp.max_layer; gp = *fine; gp.max_layer = ml; }   // :2362-2372 (max_layer is the shared global)
    int64_t nv = 0;
    rc = xyz_dev ? build_factor(ctx, &gp, true, nullptr, nullptr, stride_floats, kf_offsets, W, poses12, W, f, nullptr, 0, &nv, nullptr, xyz_dev, own_lo, own_hi)
                 : vxs_build_gba_factor(ctx, &gp, xyz, stride_floats, kf_offsets, poses12, W, f, nullptr, 0, &nv);
    if (rc < 0) break;
    double resis[2] = {0, 0};
    int is_converge = 0;
    rc = vxs_lidar_ba(ctx, f, poses12, up, thread_num, hess_out, resis, &is_converge, nullptr, 0, nullptr);
    if (rc < 0) break;
    if (rc > 0) warn = rc;
    if (resis_log) { resis_log[2 * iters] = resis[0]; resis_log[2 * iters + 1] = resis[1]; }
    iters++;
    if ((fabs(resis[0] - resis[1]) / resis[0] < converge_thre && is_converge) || (iterCnt == max_iter - 2 && converge_flag == 0)) {
      converge_thre = 0.01;
      if (converge_flag == 0) converge_flag = 1;
      else if (converge_flag == 1) break;
    }
  }
  if (outer_iters) *outer_iters = iters;
  return rc < 0 ? rc : warn;
}
// ---- routed top level (multi-GPU vxs_hba_pass): instead of giving every rank every submap (all-gather) and letting each rank key ALL points to find the ones it
// owns, every rank keys only its OWN submaps' points, sorts them by owner rank (stable: deterministic) and the points travel straight to their owners (all-to-all).
__global__ void __launch_bounds__(256) k_route_owner(const float* __restrict__ sub, const long long* __restrict__ woff, int nmine, int first_frame, const double* __restrict__ poses,
                                                     double voxel_size, int nranks, long long n, unsigned long long* __restrict__ keys, unsigned int* __restrict__ idx) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  int lo = 0, hi = nmine;
  while (hi - lo > 1) { const int mid = (lo + hi) >> 1; if (woff[mid] <= i) lo = mid; else hi = mid; }
  const int fr = first_frame + lo;
  const d3 w = world_point(poses + 12 * fr, mk3((double)sub[3 * i], (double)sub[3 * i + 1], (double)sub[3 * i + 2]));
  const unsigned int owner = (unsigned int)(voxel_hash(quantise(w.x, voxel_size), quantise(w.y, voxel_size), quantise(w.z, voxel_size)) % (unsigned long long)nranks);
  keys[i] = owner; idx[i] = (unsigned int)i;
}
// counts per owner from the SORTED owner keys (one thread per rank: two binary searches) — a per-point atomicAdd on R counters serialises in L2 (measured +28 ms on 87 M points)
__global__ void k_route_counts(const unsigned long long* __restrict__ skeys, long long n, int nranks, unsigned int* __restrict__ hist) {
  const int r = threadIdx.x;
  if (r >= nranks) return;
  long long lo = 0, hi = n;
  while (lo < hi) { const long long mid = (lo + hi) >> 1; if (skeys[mid] < (unsigned long long)r) lo = mid + 1; else hi = mid; }
  long long lo2 = lo, hi2 = n;
  while (lo2 < hi2) { const long long mid = (lo2 + hi2) >> 1; if (skeys[mid] <= (unsigned long long)r) lo2 = mid + 1; else hi2 = mid; }
  hist[r] = (unsigned int)(lo2 - lo);
}
__global__ void __launch_bounds__(256) k_route_pack(const float* __restrict__ sub, const long long* __restrict__ woff, int nmine, int first_frame, const unsigned int* __restrict__ sidx, long long n,
                                                    float4* __restrict__ out) {
  const long long j = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= n) return;
  const long long i = sidx[j];
  int lo = 0, hi = nmine;
  while (hi - lo > 1) { const int mid = (lo + hi) >> 1; if (woff[mid] <= i) lo = mid; else hi = mid; }
  out[j] = make_float4(sub[3 * i], sub[3 * i + 1], sub[3 * i + 2], __int_as_float(first_frame + lo));
}
int vxs_comm_alltoallv_f32(vxs_ctx* ctx, const float* send, const size_t* scount, const size_t* sdispl, float* recv, const size_t* rcount, const size_t* rdispl);   // vxs_lm.cu
int vxs_comm_allreduce(vxs_ctx* ctx, double* buf, size_t n);

int vxs_hba_top_routed(vxs_ctx* ctx, const vxs_map_params* coarse, const vxs_map_params* fine, const float* sub_mine, const int64_t* woff_host, int first_window, int nmine, int nwin,
                       double* poses12, int max_iter, int thread_num, double* resis_log, int* outer_iters, DevBuf<float>* sendbuf, DevBuf<float>* recvbuf) {
  VoxScratch* s = scratch(ctx);
  cudaStream_t st = ctx->stream;
  const int R = ctx->nranks;
  vxs_map_params gp = *coarse;
  const int up = 4;
  int converge_flag = 0, iters = 0, warn = 0, rc = VXS_OK;
  double converge_thre = 0.05;
  if (!s->hba_factor) { rc = vxs_factor_create(ctx, nwin, &s->hba_factor); if (rc) return rc; }
  vxs_factor* f = s->hba_factor;
  const long long n = woff_host[nmine];
  VXS_CUDA(ctx, s->totals.reserve(16));
  for (int iterCnt = 0; iterCnt < max_iter; iterCnt++) {
    if (converge_flag == 1 || iterCnt == max_iter - 1) { const int ml = gp.max_layer; gp = *fine; gp.max_layer = ml; }
    // ---- owner of every point of my submaps at the current poses, stable partition by owner
    VXS_CUDA(ctx, s->poses.reserve(size_t(nwin) * 12)); VXS_CUDA(ctx, s->offsets.reserve(size_t(std::max(nmine, 1)) + 1));
    VXS_CUDA(ctx, cudaMemcpyAsync(s->poses.p, poses12, size_t(nwin) * 96, cudaMemcpyHostToDevice, st));
    VXS_CUDA(ctx, cudaMemcpyAsync(s->offsets.p, woff_host, (size_t(nmine) + 1) * 8, cudaMemcpyHostToDevice, st));
    VXS_CUDA(ctx, s->flags.reserve(64));
    VXS_CUDA(ctx, cudaMemsetAsync(s->flags.p, 0, 64 * 4, st));
    std::vector<unsigned int> hist(static_cast<size_t>(R), 0u);
    unsigned int* sidx = nullptr;
    if (n > 0) {
      VXS_CUDA(ctx, s->keysA.reserve(size_t(n))); VXS_CUDA(ctx, s->keysB.reserve(size_t(n))); VXS_CUDA(ctx, s->idxA.reserve(size_t(n))); VXS_CUDA(ctx, s->idxB.reserve(size_t(n)));
      VXS_LAUNCH(ctx, "k_route_owner", k_route_owner, nblk(size_t(n), 256), 256, 0, sub_mine, s->offsets.p, nmine, first_window, s->poses.p, gp.voxel_size, R, n, s->keysA.p, s->idxA.p);
      unsigned long long* ks;
      rc = radix_sort(ctx, s, s->keysA.p, s->idxA.p, s->keysB.p, s->idxB.p, size_t(n), bits_for((unsigned long long)R), &ks, &sidx);
      if (rc) return rc;
      VXS_LAUNCH(ctx, "k_route_counts", k_route_counts, 1, 32, 0, ks, n, R, s->flags.p);
      VXS_CUDA(ctx, sendbuf->reserve((size_t(n) + size_t(n) / 8) * 4));
      VXS_LAUNCH(ctx, "k_route_pack", k_route_pack, nblk(size_t(n), 256), 256, 0, sub_mine, s->offsets.p, nmine, first_window, sidx, n, reinterpret_cast<float4*>(sendbuf->p));
      VXS_CUDA(ctx, cudaMemcpyAsync(hist.data(), s->flags.p, size_t(R) * 4, cudaMemcpyDeviceToHost, st));
    }
    // ---- everybody's send counts (R x R matrix through the double all-reduce), then the all-to-all
    VXS_CUDA(ctx, ctx->stage.reserve(size_t(R) * R));
    VXS_CUDA(ctx, cudaStreamSynchronize(st));
    std::vector<double> mat(size_t(R) * R, 0.0);
    for (int d = 0; d < R; d++) mat[size_t(ctx->rank) * R + d] = double(hist[size_t(d)]);
    VXS_CUDA(ctx, cudaMemcpyAsync(ctx->stage.p, mat.data(), mat.size() * 8, cudaMemcpyHostToDevice, st));
    rc = vxs_comm_allreduce(ctx, ctx->stage.p, mat.size());
    if (rc) return rc;
    VXS_CUDA(ctx, cudaMemcpyAsync(mat.data(), ctx->stage.p, mat.size() * 8, cudaMemcpyDeviceToHost, st));
    VXS_CUDA(ctx, cudaStreamSynchronize(st));
    const size_t Rz = static_cast<size_t>(R);
    std::vector<size_t> sc(Rz, 0), sd(Rz, 0), rcn(Rz, 0), rd(Rz, 0);
    size_t so = 0, ro = 0;
    for (int p = 0; p < R; p++) {
      sc[size_t(p)] = size_t(mat[size_t(ctx->rank) * R + p]) * 4; sd[size_t(p)] = so; so += sc[size_t(p)];
      rcn[size_t(p)] = size_t(mat[size_t(p) * R + ctx->rank]) * 4; rd[size_t(p)] = ro; ro += rcn[size_t(p)];
    }
    const long long nrecv = (long long)(ro / 4);
    VXS_CUDA(ctx, recvbuf->reserve((std::max<size_t>(ro, 4) / 4 + ro / 32) * 4));
    rc = vxs_comm_alltoallv_f32(ctx, sendbuf->p, sc.data(), sd.data(), recvbuf->p, rcn.data(), rd.data());
    if (rc) return rc;
    // ---- the map of the voxels I own, from the records I received (ascending source rank = ascending submap: the order of the all-gather path)
    std::vector<int64_t> fake_off(size_t(nwin) + 1, 0);
    fake_off[size_t(nwin)] = nrecv;
    int64_t nv = 0;
    rc = build_factor(ctx, &gp, true, nullptr, nullptr, 4, fake_off.data(), nwin, poses12, nwin, f, nullptr, 0, &nv, nullptr, recvbuf->p, 0, nrecv, true);
    if (rc < 0) break;
    double resis[2] = {0, 0};
    int is_converge = 0;
    rc = vxs_lidar_ba(ctx, f, poses12, up, thread_num, nullptr, resis, &is_converge, nullptr, 0, nullptr);
    if (rc < 0) break;
    if (rc > 0) warn = rc;
    if (resis_log) { resis_log[2 * iters] = resis[0]; resis_log[2 * iters + 1] = resis[1]; }
    iters++;
    if ((fabs(resis[0] - resis[1]) / resis[0] < converge_thre && is_converge) || (iterCnt == max_iter - 2 && converge_flag == 0)) {
      converge_thre = 0.01;
      if (converge_flag == 0) converge_flag = 1;
      else if (converge_flag == 1) break;
    }
  }
  if (outer_iters) *outer_iters = iters;
  return rc < 0 ? rc : warn;
}

extern "C" int vxs_hba_window(vxs_ctx* ctx, const vxs_map_params* coarse, const vxs_map_params* fine, const float* xyz, int stride_floats, const int64_t* kf_offsets,
                              double* poses12, int W, int max_iter, int thread_num, double* hess_out, double* resis_log, int* outer_iters) {
  return vxs_hba_window_impl(ctx, coarse, fine, xyz, nullptr, stride_floats, kf_offsets, poses12, W, max_iter, thread_num, hess_out, resis_log, outer_iters, -1, -1);
}
