// Device-wide primitives shared by the voxel-map units (vxs_voxelize.cu: build from scratch; vxs_map.cu: persistent map): exclusive scan,
// stable LSD radix sort of (uint64 key, uint32 value) pairs, and the "sorted keys -> segment records" kernels.  Header-only (static): every
// translation unit gets its own copy.
#pragma once
#include <algorithm>
#include "vxs_internal.h"

struct SortScratch { DevBuf<unsigned int> hist, blocksums, totals; };
static inline unsigned nblk(size_t n, unsigned b) { return unsigned((n + b - 1) / b); }
static inline int bits_for(unsigned long long v) { int b = 0; while ((1ull << b) <= v && b < 63) b++; return std::max(b, 1); }

// ------------------------------------------------------------------ exclusive scan (uint32), 3 kernels
#define SCAN_TILE 1024
static __device__ __forceinline__ unsigned int block_excl_scan_256x4(unsigned int v[4], unsigned int* total) {
  // 256 threads, 4 consecutive values each; returns the exclusive prefix of this thread's first value
  __shared__ unsigned int wsum[8];
  unsigned int t = v[0] + v[1] + v[2] + v[3];
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  unsigned int inc = t;
  for (int off = 1; off < 32; off <<= 1) { unsigned int o = __shfl_up_sync(0xffffffffu, inc, off); if (lane >= off) inc += o; }
  if (lane == 31) wsum[w] = inc;
  __syncthreads();
  if (w == 0) {
    unsigned int x = lane < 8 ? wsum[lane] : 0, xi = x;
    for (int off = 1; off < 8; off <<= 1) { unsigned int o = __shfl_up_sync(0xffffffffu, xi, off); if (lane >= off) xi += o; }
    if (lane < 8) wsum[lane] = xi - x;
    if (lane == 7 && total) *total = xi;
  }
  __syncthreads();
  const unsigned int r = wsum[w] + inc - t;
  __syncthreads();
  return r;
}
static __global__ void __launch_bounds__(256) k_scan_sums(const unsigned int* __restrict__ in, unsigned int* __restrict__ sums, size_t n) {
  const size_t base = size_t(blockIdx.x) * SCAN_TILE + threadIdx.x * 4;
  unsigned int v[4];
  for (int k = 0; k < 4; k++) v[k] = base + k < n ? in[base + k] : 0;
  __shared__ unsigned int tot;
  block_excl_scan_256x4(v, &tot);
  if (threadIdx.x == 0) sums[blockIdx.x] = tot;
}
static __global__ void __launch_bounds__(256) k_scan_single(unsigned int* __restrict__ data, size_t n, unsigned int* __restrict__ total_out) {
  __shared__ unsigned int carry, tot;
  if (threadIdx.x == 0) carry = 0;
  __syncthreads();
  for (size_t base0 = 0; base0 < n; base0 += SCAN_TILE) {
    const size_t base = base0 + threadIdx.x * 4;
    unsigned int v[4];
    for (int k = 0; k < 4; k++) v[k] = base + k < n ? data[base + k] : 0;
    unsigned int ex = block_excl_scan_256x4(v, &tot) + carry;
    for (int k = 0; k < 4; k++) { if (base + k < n) data[base + k] = ex; ex += v[k]; }
    __syncthreads();
    if (threadIdx.x == 0) carry += tot;
    __syncthreads();
  }
  if (threadIdx.x == 0 && total_out) *total_out = carry;
}
static __global__ void __launch_bounds__(256) k_scan_apply(const unsigned int* __restrict__ in, unsigned int* __restrict__ out, const unsigned int* __restrict__ offs, size_t n) {
  const size_t base = size_t(blockIdx.x) * SCAN_TILE + threadIdx.x * 4;
  unsigned int v[4];
  for (int k = 0; k < 4; k++) v[k] = base + k < n ? in[base + k] : 0;
  unsigned int ex = block_excl_scan_256x4(v, nullptr) + offs[blockIdx.x];
  for (int k = 0; k < 4; k++) { if (base + k < n) out[base + k] = ex; ex += v[k]; }
}
// out = exclusive scan of in (may alias), *total_dev = sum
static int scan_u32(vxs_ctx* ctx, SortScratch* s, const unsigned int* in, unsigned int* out, size_t n, unsigned int* total_dev) {
  if (n == 0) { VXS_CUDA(ctx, cudaMemsetAsync(total_dev, 0, 4, ctx->stream)); return VXS_OK; }
  const unsigned nb = nblk(n, SCAN_TILE);
  VXS_CUDA(ctx, s->blocksums.reserve(nb));
  VXS_LAUNCH(ctx, "k_scan", k_scan_sums, nb, 256, 0, in, s->blocksums.p, n);
  VXS_LAUNCH(ctx, "k_scan", k_scan_single, 1, 256, 0, s->blocksums.p, size_t(nb), total_dev);
  VXS_LAUNCH(ctx, "k_scan", k_scan_apply, nb, 256, 0, in, out, s->blocksums.p, n);
  return VXS_OK;
}

// ------------------------------------------------------------------ LSD radix sort, 8-bit digits, (uint64 key, uint32 value)
#define RS_THREADS 256
#define RS_ITEMS 16
#define RS_TILE (RS_THREADS * RS_ITEMS)
static __global__ void __launch_bounds__(RS_THREADS) k_radix_hist(const unsigned long long* __restrict__ keys, size_t n, int shift, unsigned int* __restrict__ hist, unsigned int nblocks) {
  __shared__ unsigned int cnt[256];
  cnt[threadIdx.x] = 0;
  __syncthreads();
  const size_t base = size_t(blockIdx.x) * RS_TILE;
  for (int r = 0; r < RS_ITEMS; r++) {
    const size_t i = base + size_t(r) * RS_THREADS + threadIdx.x;
    if (i < n) atomicAdd(&cnt[(keys[i] >> shift) & 255], 1u);
  }
  __syncthreads();
  hist[size_t(threadIdx.x) * nblocks + blockIdx.x] = cnt[threadIdx.x];
}
static __global__ void __launch_bounds__(RS_THREADS) k_radix_scatter(const unsigned long long* __restrict__ kin, const unsigned int* __restrict__ vin, unsigned long long* __restrict__ kout,
                                                              unsigned int* __restrict__ vout, size_t n, int shift, const unsigned int* __restrict__ base, unsigned int nblocks) {
  __shared__ unsigned int wcnt[RS_THREADS / 32][256];
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  for (int k = threadIdx.x; k < (RS_THREADS / 32) * 256; k += RS_THREADS) (&wcnt[0][0])[k] = 0;
  __syncthreads();
  const size_t wbase = size_t(blockIdx.x) * RS_TILE + size_t(w) * (32 * RS_ITEMS);
  unsigned long long key[RS_ITEMS];
  unsigned int rank[RS_ITEMS];
#pragma unroll
  for (int r = 0; r < RS_ITEMS; r++) {
    const size_t i = wbase + size_t(r) * 32 + lane;
    const bool valid = i < n;
    key[r] = valid ? kin[i] : 0ull;
    const unsigned int bin = valid ? (unsigned int)((key[r] >> shift) & 255) : 256u;
    const unsigned int peers = __match_any_sync(0xffffffffu, bin);
    const int leader = __ffs(peers) - 1;
    const unsigned int below = __popc(peers & ((1u << lane) - 1u));
    unsigned int old = 0;
    if (lane == leader && valid) { old = wcnt[w][bin]; wcnt[w][bin] = old + __popc(peers); }
    old = __shfl_sync(0xffffffffu, old, leader);
    rank[r] = old + below;
    __syncwarp();
  }
  __syncthreads();
  {  // per bin: global base of this block, then exclusive prefix over the warps of the block
    const int bin = threadIdx.x;
    unsigned int run = base[size_t(bin) * nblocks + blockIdx.x];
    for (int ww = 0; ww < RS_THREADS / 32; ww++) { const unsigned int c = wcnt[ww][bin]; wcnt[ww][bin] = run; run += c; }
  }
  __syncthreads();
#pragma unroll
  for (int r = 0; r < RS_ITEMS; r++) {
    const size_t i = wbase + size_t(r) * 32 + lane;
    if (i < n) {
      const unsigned int pos = wcnt[w][(key[r] >> shift) & 255] + rank[r];
      kout[pos] = key[r];
      vout[pos] = vin[i];
    }
  }
}
// sorts by the low `bits` bits; result pointers returned through kres / vres (ping-pong between the A and B buffers)
static int radix_sort(vxs_ctx* ctx, SortScratch* s, unsigned long long* kA, unsigned int* vA, unsigned long long* kB, unsigned int* vB, size_t n, int bits,
                      unsigned long long** kres, unsigned int** vres) {
  *kres = kA; *vres = vA;
  if (n == 0) return VXS_OK;
  const unsigned nb = nblk(n, RS_TILE);
  VXS_CUDA(ctx, s->hist.reserve(size_t(256) * nb));
  for (int shift = 0; shift < bits; shift += 8) {
    VXS_LAUNCH(ctx, "k_radix_hist", k_radix_hist, nb, RS_THREADS, 0, *kres, n, shift, s->hist.p, nb);
    int rc = scan_u32(ctx, s, s->hist.p, s->hist.p, size_t(256) * nb, s->totals.p + 15);
    if (rc) return rc;
    unsigned long long* ko = (*kres == kA) ? kB : kA;
    unsigned int* vo = (*vres == vA) ? vB : vA;
    VXS_LAUNCH(ctx, "k_radix_scatter", k_radix_scatter, nb, RS_THREADS, 0, *kres, *vres, ko, vo, n, shift, s->hist.p, nb);
    *kres = ko; *vres = vo;
  }
  return VXS_OK;
}

// ------------------------------------------------------------------ segments -> records -> nodes
static __global__ void k_flag_heads(const unsigned long long* __restrict__ keys, size_t m, unsigned int* __restrict__ flag) {
  const size_t j = size_t(blockIdx.x) * blockDim.x + threadIdx.x;
  if (j < m) flag[j] = (j == 0 || keys[j] != keys[j - 1]) ? 1u : 0u;
}
static __global__ void k_write_records(const unsigned long long* __restrict__ keys, const unsigned int* __restrict__ flag, const unsigned int* __restrict__ ex, size_t m,
                                unsigned int* __restrict__ rec_start, unsigned long long* __restrict__ rec_key, const unsigned int* __restrict__ total) {
  const size_t j = size_t(blockIdx.x) * blockDim.x + threadIdx.x;
  if (j < m && flag[j]) { rec_start[ex[j]] = (unsigned int)j; rec_key[ex[j]] = keys[j]; }
  if (j == 0) rec_start[*total] = (unsigned int)m;
}
static __global__ void k_flag_nodes(const unsigned long long* __restrict__ rec_key, size_t R, int FB, unsigned int* __restrict__ flag) {
  const size_t r = size_t(blockIdx.x) * blockDim.x + threadIdx.x;
  if (r < R) flag[r] = (r == 0 || (rec_key[r] >> FB) != (rec_key[r - 1] >> FB)) ? 1u : 0u;
}
static __global__ void k_write_nodes(const unsigned int* __restrict__ flag, const unsigned int* __restrict__ ex, size_t R, unsigned int* __restrict__ node_of_rec,
                              unsigned int* __restrict__ node_rec_start, const unsigned int* __restrict__ total) {
  const size_t r = size_t(blockIdx.x) * blockDim.x + threadIdx.x;
  if (r < R) { const unsigned int nd = ex[r] + flag[r] - 1u; node_of_rec[r] = nd; if (flag[r]) node_rec_start[nd] = (unsigned int)r; }
  if (r == 0) node_rec_start[*total] = (unsigned int)R;
}

