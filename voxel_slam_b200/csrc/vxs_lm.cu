// Levenberg–Marquardt drivers of the reference, host-orchestrated over the CUDA kernels:
//   Lidar_BA_Optimizer::damping_iter        voxel_map.hpp:367-442   -> vxs_lidar_ba
//   LI_BA_Optimizer::damping_iter           voxel_map.hpp:562-653   -> vxs_li_ba(with_gravity = 0)
//   LI_BA_OptimizerGravity::damping_iter    voxel_map.hpp:775-862   -> vxs_li_ba(with_gravity = 1)
// Per iteration the GPU does Hessian build, assembly, damped LDL^T solve and the residual-only evaluation; the host does the
// scalar accept/reject logic, the (tiny) state retraction x [+] dx and the IMU callbacks.  3n+1 doubles come back per
// iteration; the factor, the Hessian and all workspaces stay in HBM.
// Multi-GPU (voxel-sharded factor): [C | g | D | r1] is all-reduced once per Hessian build and r2 once per residual
// evaluation (vxs_eval.cu); every rank then runs the identical replicated solve.
#include <dlfcn.h>
#include <math.h>
#include <string.h>
#include <algorithm>
#include <mutex>
#include <vector>
#include "vxs_internal.h"
#include "vxs_math.cuh"

using namespace vxs;

int vxs_assemble_dev(vxs_ctx* ctx, vxs_factor* f, int S, int n, const double* blocks, const double* gvec, int bs, double imu_coef);
double* vxs_hess_r1_dev(vxs_factor* f);

// ------------------------------------------------------------------ NCCL (dlopen'ed: libvxs has no link-time dependency on it)
typedef struct { char internal[128]; } vxs_nccl_uid;
struct NcclApi {
  void* lib = nullptr;
  int (*GetUniqueId)(vxs_nccl_uid*) = nullptr;
  int (*CommInitRank)(void**, int, vxs_nccl_uid, int) = nullptr;
  int (*AllReduce)(const void*, void*, size_t, int, int, void*, cudaStream_t) = nullptr;
  int (*CommDestroy)(void*) = nullptr;
  int (*Broadcast)(const void*, void*, size_t, int, int, void*, cudaStream_t) = nullptr;
  int (*GroupStart)() = nullptr;
  int (*GroupEnd)() = nullptr;
  int (*AllGather)(const void*, void*, size_t, int, void*, cudaStream_t) = nullptr;
  int (*Send)(const void*, size_t, int, int, void*, cudaStream_t) = nullptr;
  int (*Recv)(void*, size_t, int, int, void*, cudaStream_t) = nullptr;
  const char* (*GetErrorString)(int) = nullptr;
};
static NcclApi* nccl_api() {
  static NcclApi api;
  static std::once_flag once;
  std::call_once(once, [] {
    const char* names[] = {"libnccl.so.2", "libnccl.so"};
    for (const char* nm : names) { api.lib = dlopen(nm, RTLD_NOW | RTLD_GLOBAL); if (api.lib) break; }
    if (api.lib) {
      api.GetUniqueId = (int (*)(vxs_nccl_uid*))dlsym(api.lib, "ncclGetUniqueId");
      api.CommInitRank = (int (*)(void**, int, vxs_nccl_uid, int))dlsym(api.lib, "ncclCommInitRank");
      api.AllReduce = (int (*)(const void*, void*, size_t, int, int, void*, cudaStream_t))dlsym(api.lib, "ncclAllReduce");
      api.CommDestroy = (int (*)(void*))dlsym(api.lib, "ncclCommDestroy");
      api.Broadcast = (int (*)(const void*, void*, size_t, int, int, void*, cudaStream_t))dlsym(api.lib, "ncclBroadcast");
      api.GroupStart = (int (*)())dlsym(api.lib, "ncclGroupStart");
      api.GroupEnd = (int (*)())dlsym(api.lib, "ncclGroupEnd");
      api.AllGather = (int (*)(const void*, void*, size_t, int, void*, cudaStream_t))dlsym(api.lib, "ncclAllGather");
      api.Send = (int (*)(const void*, size_t, int, int, void*, cudaStream_t))dlsym(api.lib, "ncclSend");
      api.Recv = (int (*)(void*, size_t, int, int, void*, cudaStream_t))dlsym(api.lib, "ncclRecv");
      api.GetErrorString = (const char* (*)(int))dlsym(api.lib, "ncclGetErrorString");
      if (!api.GetUniqueId || !api.CommInitRank || !api.AllReduce || !api.CommDestroy) api.lib = nullptr;
    }
  });
  return api.lib ? &api : nullptr;
}
extern "C" int vxs_comm_unique_id(unsigned char id[128]) {
  NcclApi* a = nccl_api();
  if (!a || !id) return VXS_ERR_COMM;
  vxs_nccl_uid u;
  if (a->GetUniqueId(&u) != 0) return VXS_ERR_COMM;
  memcpy(id, u.internal, 128);
  return VXS_OK;
}
extern "C" int vxs_ctx_comm_init(vxs_ctx* ctx, const unsigned char id[128], int rank, int nranks) {
  if (!ctx || !id || rank < 0 || nranks < 1 || rank >= nranks) return VXS_ERR_ARG;
  NcclApi* a = nccl_api();
  if (!a) return vxs_fail(ctx, VXS_ERR_COMM, "libnccl.so.2 not loadable");
  cudaSetDevice(ctx->device);
  vxs_nccl_uid u; memcpy(u.internal, id, 128);
  void* comm = nullptr;
  int rc = a->CommInitRank(&comm, nranks, u, rank);
  if (rc != 0) return vxs_fail(ctx, VXS_ERR_COMM, a->GetErrorString ? a->GetErrorString(rc) : "ncclCommInitRank failed");
  ctx->comm = comm; ctx->rank = rank; ctx->nranks = nranks;
  return VXS_OK;
}
extern "C" int vxs_ctx_comm_destroy(vxs_ctx* ctx) {
  if (!ctx) return VXS_ERR_ARG;
  if (ctx->comm) { NcclApi* a = nccl_api(); if (a) a->CommDestroy(ctx->comm); ctx->comm = nullptr; }
  ctx->rank = 0; ctx->nranks = 1;
  return VXS_OK;
}
int vxs_comm_allreduce(vxs_ctx* ctx, double* buf, size_t n) {
  if (ctx->nranks <= 1) return VXS_OK;
  NcclApi* a = nccl_api();
  if (!a || !ctx->comm) return vxs_fail(ctx, VXS_ERR_COMM, "communicator not initialised");
  if (ctx->timing) { vxs_stage_begin(ctx, vxs_stage_id(ctx, "nccl_allreduce")); }
  int rc = a->AllReduce(buf, buf, n, /*ncclFloat64*/ 8, /*ncclSum*/ 0, ctx->comm, ctx->stream);
  if (ctx->timing) vxs_stage_end(ctx);
  if (rc != 0) return vxs_fail(ctx, VXS_ERR_COMM, a->GetErrorString ? a->GetErrorString(rc) : "ncclAllReduce failed");
  return VXS_OK;
}

// variable-size all-gather of float data between device buffers (counts / displs in floats): ONE ncclAllGather of equal, padded slots into `pad`
// (nranks x slot floats, slot >= max count; `mine` must hold slot floats of readable memory) followed by device-to-device compaction copies — a ring / NVLS
// all-gather runs at bus bandwidth, eight grouped broadcasts of ~250 MB measured 2-3x slower.  Falls back to grouped broadcasts when pad is NULL.
int vxs_comm_allgatherv_f32(vxs_ctx* ctx, const float* mine, size_t my_count, float* all, const size_t* counts, const size_t* displs, float* pad, size_t slot) {
  if (ctx->nranks <= 1) return VXS_OK;
  NcclApi* a = nccl_api();
  if (!a || !ctx->comm || !a->Broadcast || !a->GroupStart || !a->GroupEnd) return vxs_fail(ctx, VXS_ERR_COMM, "communicator not initialised / ncclBroadcast missing");
  if (ctx->timing) vxs_stage_begin(ctx, vxs_stage_id(ctx, "nccl_allgatherv"));
  int rc = 0, rc2 = 0;
  if (pad && a->AllGather) {
    rc = a->AllGather(mine, pad, slot, /*ncclFloat32*/ 7, ctx->comm, ctx->stream);
    for (int r = 0; r < ctx->nranks && rc == 0; r++)
      if (counts[r]) { if (cudaMemcpyAsync(all + displs[r], pad + size_t(r) * slot, counts[r] * 4, cudaMemcpyDeviceToDevice, ctx->stream) != cudaSuccess) rc2 = 1; }
  } else {
    rc = a->GroupStart();
    for (int r = 0; r < ctx->nranks && rc == 0; r++) {
      if (counts[r] == 0) continue;
      const void* send = r == ctx->rank ? (const void*)mine : (const void*)(all + displs[r]);
      rc = a->Broadcast(send, all + displs[r], counts[r], /*ncclFloat32*/ 7, r, ctx->comm, ctx->stream);
    }
    rc2 = a->GroupEnd();
  }
  if (ctx->timing) vxs_stage_end(ctx);
  (void)my_count;
  if (rc != 0 || rc2 != 0) return vxs_fail(ctx, VXS_ERR_COMM, (rc && a->GetErrorString) ? a->GetErrorString(rc) : "all-gather of the submaps failed");
  return VXS_OK;
}
// all-to-all with per-peer counts / displacements (in floats): grouped ncclSend / ncclRecv, the own part by a device copy
int vxs_comm_alltoallv_f32(vxs_ctx* ctx, const float* send, const size_t* scount, const size_t* sdispl, float* recv, const size_t* rcount, const size_t* rdispl) {
  NcclApi* a = nccl_api();
  if (!a || !ctx->comm || !a->Send || !a->Recv || !a->GroupStart || !a->GroupEnd) return vxs_fail(ctx, VXS_ERR_COMM, "communicator not initialised / ncclSend missing");
  if (ctx->timing) vxs_stage_begin(ctx, vxs_stage_id(ctx, "nccl_alltoallv"));
  const int me = ctx->rank;
  int rc = 0;
  if (rcount[me]) { if (cudaMemcpyAsync(recv + rdispl[me], send + sdispl[me], rcount[me] * 4, cudaMemcpyDeviceToDevice, ctx->stream) != cudaSuccess) rc = 1; }
  int rg = a->GroupStart();
  for (int p = 0; p < ctx->nranks && rc == 0 && rg == 0; p++) {
    if (p == me) continue;
    if (scount[p]) rc = a->Send(send + sdispl[p], scount[p], /*ncclFloat32*/ 7, p, ctx->comm, ctx->stream);
    if (rc == 0 && rcount[p]) rc = a->Recv(recv + rdispl[p], rcount[p], /*ncclFloat32*/ 7, p, ctx->comm, ctx->stream);
  }
  const int rg2 = a->GroupEnd();
  if (ctx->timing) vxs_stage_end(ctx);
  if (rc || rg || rg2) return vxs_fail(ctx, VXS_ERR_COMM, "all-to-all of the routed points failed");
  return VXS_OK;
}
// element-wise MAX all-reduce of a few int64 (bounding boxes: [-min | max])
int vxs_comm_allreduce_max_i64(vxs_ctx* ctx, long long* buf, size_t n) {
  if (ctx->nranks <= 1) return VXS_OK;
  NcclApi* a = nccl_api();
  if (!a || !ctx->comm) return vxs_fail(ctx, VXS_ERR_COMM, "communicator not initialised");
  const int rc = a->AllReduce(buf, buf, n, /*ncclInt64*/ 4, /*ncclMax*/ 2, ctx->comm, ctx->stream);
  if (rc != 0) return vxs_fail(ctx, VXS_ERR_COMM, a->GetErrorString ? a->GetErrorString(rc) : "ncclAllReduce failed");
  return VXS_OK;
}

// ------------------------------------------------------------------ host-side state retraction
// x_temp = x [+] dxi   (voxel_map.hpp:405-409, 599-606, 813-822);  state stride sst (12 or 24), dof stride dst (6 or 15)
static void retract(const double* x, double* xt, const double* dxi, int W, int sst, int dst) {
  for (int j = 0; j < W; j++) {
    const double* s = x + size_t(j) * sst; double* o = xt + size_t(j) * sst; const double* d = dxi + size_t(j) * dst;
    rot3 R = load_rot(s);
    rot3 Rn = rot_mul(R, so3_exp(mk3(d[0], d[1], d[2])));
    o[0] = Rn.r00; o[1] = Rn.r01; o[2] = Rn.r02; o[3] = Rn.r10; o[4] = Rn.r11; o[5] = Rn.r12; o[6] = Rn.r20; o[7] = Rn.r21; o[8] = Rn.r22;
    for (int k = 0; k < 3; k++) o[9 + k] = s[9 + k] + d[3 + k];
    if (dst == 15) for (int k = 0; k < 9; k++) o[12 + k] = s[12 + k] + d[6 + k];  // v, bg, ba
  }
}

struct LmCommon {
  vxs_ctx* ctx; vxs_factor* f;
  int W, n, S, gauge, sst;
  std::vector<double> x, xt, dx, D, rhs;
};

// one Hessian build on the device, in two halves so that the caller can evaluate the IMU factors on the CPU in between
// (the reference does exactly that: worker threads run acc_evaluate2 while the main thread runs give_evaluate, voxel_map.hpp:487-499)
static int hessian_launch(LmCommon& c) {
  vxs_ctx* ctx = c.ctx;
  VXS_CUDA(ctx, cudaMemcpyAsync(ctx->states_a.p, c.x.data(), c.x.size() * 8, cudaMemcpyHostToDevice, ctx->stream));
  return vxs_eval_hessian_dev(ctx, c.f, ctx->states_a.p, c.sst, nullptr);
}
static int hessian_assemble(LmCommon& c, const double* blocks_h, const double* gvec_h, int bs, double imu_coef) {
  vxs_ctx* ctx = c.ctx;
  const double *bd = nullptr, *gd = nullptr;
  if (blocks_h) {
    const size_t nb = size_t(c.W - 1) * bs * bs, ng = size_t(c.W - 1) * bs;
    VXS_CUDA(ctx, ctx->himu.reserve(nb + ng));
    VXS_CUDA(ctx, cudaMemcpyAsync(ctx->himu.p, blocks_h, nb * 8, cudaMemcpyHostToDevice, ctx->stream));
    VXS_CUDA(ctx, cudaMemcpyAsync(ctx->himu.p + nb, gvec_h, ng * 8, cudaMemcpyHostToDevice, ctx->stream));
    bd = ctx->himu.p; gd = ctx->himu.p + nb;
  }
  return vxs_assemble_dev(ctx, c.f, c.S, c.n, bd, gd, bs, imu_coef);
}

// solve + bring dx, D, rhs (and one extra device scalar) back
static int solve_and_fetch(LmCommon& c, double u, const double* extra_dev, double* extra_host, int* singular) {
  vxs_ctx* ctx = c.ctx;
  const int n = c.n;
  VXS_CUDA(ctx, ctx->dx.reserve(size_t(n)));
  VXS_CUDA(ctx, ctx->dvec.reserve(size_t(n)));
  VXS_CUDA(ctx, ctx->rhs.reserve(size_t(n)));
  // results come back through pinned staging: a D2H copy into pageable memory is a synchronous staged copy per call
  const size_t need = size_t(3) * n + 8;
  if (ctx->h_pin_cap < need) {
    if (ctx->h_pin) cudaFreeHost(ctx->h_pin);
    ctx->h_pin = nullptr; ctx->h_pin_cap = 0;
    VXS_CUDA(ctx, cudaHostAlloc((void**)&ctx->h_pin, need * 8, cudaHostAllocDefault));
    ctx->h_pin_cap = need;
  }
  int* sing_pin = reinterpret_cast<int*>(ctx->h_pin + 3 * size_t(n) + 1);
  int rc = vxs_solve_damped(ctx, ctx->Hraw.p, ctx->jact.p, n, c.gauge, u, ctx->dx.p, ctx->dvec.p, ctx->rhs.p, singular ? sing_pin : nullptr);
  if (rc) return rc;
  VXS_CUDA(ctx, cudaMemcpyAsync(ctx->h_pin, ctx->dx.p, size_t(n) * 8, cudaMemcpyDeviceToHost, ctx->stream));
  VXS_CUDA(ctx, cudaMemcpyAsync(ctx->h_pin + n, ctx->dvec.p, size_t(n) * 8, cudaMemcpyDeviceToHost, ctx->stream));
  VXS_CUDA(ctx, cudaMemcpyAsync(ctx->h_pin + 2 * size_t(n), ctx->rhs.p, size_t(n) * 8, cudaMemcpyDeviceToHost, ctx->stream));
  if (extra_dev) VXS_CUDA(ctx, cudaMemcpyAsync(ctx->h_pin + 3 * size_t(n), extra_dev, 8, cudaMemcpyDeviceToHost, ctx->stream));
  VXS_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  memcpy(c.dx.data(), ctx->h_pin, size_t(n) * 8);
  memcpy(c.D.data(), ctx->h_pin + n, size_t(n) * 8);
  memcpy(c.rhs.data(), ctx->h_pin + 2 * size_t(n), size_t(n) * 8);
  if (extra_dev) *extra_host = ctx->h_pin[3 * size_t(n)];
  if (singular) *singular = *sing_pin;
  return VXS_OK;
}

static int lm_prepare(LmCommon& c) {
  vxs_ctx* ctx = c.ctx;
  cudaSetDevice(ctx->device);
  VXS_CUDA(ctx, ctx->states_a.reserve(size_t(c.W) * 24));
  VXS_CUDA(ctx, ctx->states_b.reserve(size_t(c.W) * 24));
  c.dx.assign(c.n, 0.0); c.D.assign(c.n, 0.0); c.rhs.assign(c.n, 0.0);
  return VXS_OK;
}

// ------------------------------------------------------------------ Lidar_BA_Optimizer::damping_iter
extern "C" int vxs_lidar_ba(vxs_ctx* ctx, vxs_factor* f, double* poses12, int max_iter, int thd_num, double* hess_out, double resis[2],
                            int* is_converge_out, vxs_lm_trace* trace, int trace_cap, int* trace_len) {
  if (!ctx || !f || !poses12 || f->ctx != ctx || max_iter < 0) return VXS_ERR_ARG;
  LmCommon c;
  c.ctx = ctx; c.f = f; c.W = f->W; c.S = 6; c.n = 6 * f->W; c.gauge = 6; c.sst = 12;
  int rc = lm_prepare(c);
  if (rc) return rc;
  const int W = c.W, n = c.n;
  c.x.assign(poses12, poses12 + size_t(W) * 12);
  c.xt = c.x;
  double u = 0.01, v = 2, residual1 = 0, residual2 = 0, q;
  bool is_calc_hess = true, is_converge = true;
  int ntrace = 0, warn = 0, singular = 0;
  if (trace_len) *trace_len = 0;
  for (int i = 0; i < max_iter; i++) {
    const bool built = is_calc_hess;
    if (is_calc_hess) { rc = hessian_launch(c); if (rc) return rc; rc = hessian_assemble(c, nullptr, nullptr, 0, 0.0); if (rc) return rc; }
    rc = solve_and_fetch(c, u, built ? vxs_hess_r1_dev(f) : nullptr, &residual1, &singular);
    if (rc) return rc;
    if (singular) warn = VXS_WARN_SINGULAR;
    if (i == 0 && resis) resis[0] = residual1;
    retract(c.x.data(), c.xt.data(), c.dx.data(), W, 12, 6);
    double q1 = 0;
    for (int k = 0; k < n; k++) q1 += c.dx[k] * (u * c.D[k] * c.dx[k] + c.rhs[k]);   // rhs = -JacT (gauged)
    q1 *= 0.5;
    if (ctx->nranks == 1 && f->V < thd_num) return vxs_fail(ctx, VXS_ERR_TOO_FEW_VOXELS, "Too Less Voxel (voxel_map.hpp:345-348)");
    VXS_CUDA(ctx, cudaMemcpyAsync(ctx->states_b.p, c.xt.data(), size_t(W) * 12 * 8, cudaMemcpyHostToDevice, ctx->stream));
    rc = vxs_eval_residual_dev(ctx, f, ctx->states_b.p, 12, ctx->scal.p);
    if (rc) return rc;
    VXS_CUDA(ctx, cudaMemcpyAsync(&residual2, ctx->scal.p, 8, cudaMemcpyDeviceToHost, ctx->stream));
    VXS_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    q = residual1 - residual2;
    if (trace && ntrace < trace_cap) { vxs_lm_trace t = {residual1, residual2, u, v, q1, q > 0 ? 1 : 0, built ? 1 : 0}; trace[ntrace++] = t; }
    if (q > 0) {
      c.x = c.xt;
      const double one_three = 1.0 / 3;
      q = q / q1; v = 2; q = 1 - pow(2 * q - 1, 3);
      u *= (q < one_three ? one_three : q);
      is_calc_hess = true;
    } else { u = u * v; v = 2 * v; is_calc_hess = false; is_converge = false; }
    if (fabs((residual1 - residual2) / residual1) < 1e-6) break;
  }
  if (resis) resis[1] = residual2;
  if (is_converge_out) *is_converge_out = is_converge ? 1 : 0;
  if (trace_len) *trace_len = ntrace;
  memcpy(poses12, c.x.data(), size_t(W) * 12 * 8);
  if (hess_out && max_iter > 0) {
    VXS_CUDA(ctx, cudaMemcpyAsync(hess_out, ctx->Hraw.p, size_t(n) * n * 8, cudaMemcpyDeviceToHost, ctx->stream));
    VXS_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  }
  return warn;
}

// ------------------------------------------------------------------ LI_BA_Optimizer(+Gravity)::damping_iter
extern "C" int vxs_li_ba(vxs_ctx* ctx, vxs_factor* f, double* states24, int with_gravity, int max_iter, double imu_coef, const vxs_imu_hooks* imu,
                         double* hess_out, double resis[2], vxs_lm_trace* trace, int trace_cap, int* trace_len) {
  if (!ctx || !f || !states24 || !imu || !imu->eval || !imu->update || !imu->rollback || f->ctx != ctx || max_iter < 0) return VXS_ERR_ARG;
  LmCommon c;
  c.ctx = ctx; c.f = f; c.W = f->W; c.S = 15; c.n = 15 * f->W + (with_gravity ? 3 : 0); c.gauge = with_gravity ? 6 : 15; c.sst = 24;
  int rc = lm_prepare(c);
  if (rc) return rc;
  // LI_BA_Optimizer::damping_iter hard-codes `for(int i=0; i<3; i++)` (voxel_map.hpp:581): never more than 3 iterations without gravity,
  // whatever the caller passes (fewer are allowed so that a single iteration can be stepped / timed); the gravity variant honours max_iter (:796)
  if (!with_gravity && max_iter > 3) max_iter = 3;
  const int W = c.W, n = c.n, bs = with_gravity ? 33 : 30;
  c.x.assign(states24, states24 + size_t(W) * 24);
  c.xt = c.x;
  std::vector<double> blocks(size_t(W - 1) * bs * bs), gvec(size_t(W - 1) * bs);
  double u = 0.01, v = 2, residual1 = 0, residual2 = 0, q, r_imu1 = 0;
  bool is_calc_hess = true;
  int ntrace = 0, warn = 0, singular = 0;
  if (trace_len) *trace_len = 0;
  for (int it = 0; it < max_iter; it++) {
    const bool built = is_calc_hess;
    double r1_lidar = 0;
    if (is_calc_hess) {
      // divide_thread (voxel_map.hpp:465-523): the IMU factors are evaluated by the caller's code on the CPU
      rc = hessian_launch(c);   // GPU: lidar part (asynchronous)
      if (rc) return rc;
      double cost = 0;          // CPU meanwhile: IMU part
      if (imu->eval(imu->user, c.x.data(), W, with_gravity, 1, blocks.data(), gvec.data(), &cost) != 0) return vxs_fail(ctx, VXS_ERR_CALLBACK, "imu eval");
      r_imu1 = cost * (imu_coef * 0.5);
      rc = hessian_assemble(c, blocks.data(), gvec.data(), bs, imu_coef);
      if (rc) return rc;
    }
    rc = solve_and_fetch(c, u, built ? vxs_hess_r1_dev(f) : nullptr, &r1_lidar, &singular);
    if (rc) return rc;
    if (singular) warn = VXS_WARN_SINGULAR;
    if (built) residual1 = r_imu1 + r1_lidar;
    if (it == 0 && resis) resis[0] = residual1;
    if (with_gravity) {  // x_stats_temp[0].g += dxi.tail(3) accumulates on the TEMP state (voxel_map.hpp:813), then is broadcast (:821)
      for (int k = 0; k < 3; k++) c.xt[21 + k] += c.dx[n - 3 + k];
    }
    const double g0[3] = {c.xt[21], c.xt[22], c.xt[23]};
    retract(c.x.data(), c.xt.data(), c.dx.data(), W, 24, 15);
    for (int j = 0; j < W; j++) for (int k = 0; k < 3; k++) c.xt[size_t(j) * 24 + 21 + k] = with_gravity ? g0[k] : c.x[size_t(j) * 24 + 21 + k];
    if (imu->update(imu->user, c.dx.data(), W) != 0) return vxs_fail(ctx, VXS_ERR_CALLBACK, "imu update");
    double q1 = 0;
    for (int k = 0; k < n; k++) q1 += c.dx[k] * (u * c.D[k] * c.dx[k] + c.rhs[k]);
    q1 *= 0.5;
    // only_residual (voxel_map.hpp:525-560): GPU evaluates the lidar part while the host evaluates the IMU part
    VXS_CUDA(ctx, cudaMemcpyAsync(ctx->states_b.p, c.xt.data(), size_t(W) * 24 * 8, cudaMemcpyHostToDevice, ctx->stream));
    rc = vxs_eval_residual_dev(ctx, f, ctx->states_b.p, 24, ctx->scal.p);
    if (rc) return rc;
    double r2_lidar = 0, cost2 = 0;
    VXS_CUDA(ctx, cudaMemcpyAsync(&r2_lidar, ctx->scal.p, 8, cudaMemcpyDeviceToHost, ctx->stream));
    if (imu->eval(imu->user, c.xt.data(), W, with_gravity, 0, nullptr, nullptr, &cost2) != 0) return vxs_fail(ctx, VXS_ERR_CALLBACK, "imu eval");
    VXS_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    residual2 = cost2 * (imu_coef * 0.5) + r2_lidar;
    q = residual1 - residual2;
    if (trace && ntrace < trace_cap) { vxs_lm_trace t = {residual1, residual2, u, v, q1, q > 0 ? 1 : 0, built ? 1 : 0}; trace[ntrace++] = t; }
    if (q > 0) {
      c.x = c.xt;
      const double one_three = 1.0 / 3;
      q = q / q1; v = 2; q = 1 - pow(2 * q - 1, 3);
      u *= (q < one_three ? one_three : q);
      is_calc_hess = true;
    } else {
      u = u * v; v = 2 * v; is_calc_hess = false;
      if (imu->rollback(imu->user) != 0) return vxs_fail(ctx, VXS_ERR_CALLBACK, "imu rollback");
    }
    if (fabs((residual1 - residual2) / residual1) < 1e-6) break;
  }
  if (resis) resis[1] = residual2;
  if (trace_len) *trace_len = ntrace;
  memcpy(states24, c.x.data(), size_t(W) * 24 * 8);
  if (hess_out && max_iter > 0) {
    VXS_CUDA(ctx, cudaMemcpyAsync(hess_out, ctx->Hraw.p, size_t(n) * n * 8, cudaMemcpyDeviceToHost, ctx->stream));
    VXS_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  }
  return warn;
}

// ------------------------------------------------------------------ PGO edges from the raw Hessian (voxelslam.cpp:2405-2427)
// The reference emits the edges in lexicographic (i, j) order (two nested loops); so does this: one CTA per row i counts its edges, a
// single-CTA scan turns the counts into row offsets, and the emit pass compacts every row in j order (ballot prefix), so the output — and
// the subset that survives when it does not fit `cap` — is the same on every run.
__device__ __forceinline__ bool edge_ok(const double* __restrict__ H, int n, int i, int j, double* v) {
  for (int k = 0; k < 6; k++) {
    const double hc = fabs(H[size_t(6 * j + k) * n + 6 * i + k]);
    if (hc < 1e-6) return false;
    if (v) v[k] = 1.0 / hc;
  }
  return true;
}
__global__ void __launch_bounds__(256) k_hba_edge_count(const double* __restrict__ H, int n, int W, unsigned int* __restrict__ rowcnt) {
  const int i = blockIdx.x;
  int c = 0;
  for (int j = i + 1 + threadIdx.x; j < W; j += blockDim.x) c += edge_ok(H, n, i, j, nullptr) ? 1 : 0;
  __shared__ int sh[8];
  for (int off = 16; off > 0; off >>= 1) c += __shfl_xor_sync(0xffffffffu, c, off);
  if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = c;
  __syncthreads();
  if (threadIdx.x == 0) { int t = 0; for (int w = 0; w < 8; w++) t += sh[w]; rowcnt[i] = (unsigned int)t; }
}
__global__ void __launch_bounds__(1024) k_hba_edge_scan(unsigned int* __restrict__ rowcnt, int W, unsigned int* __restrict__ total) {
  // exclusive scan of W counts in place, W small (<= a few thousand): chunks of 1024 with a carried offset
  __shared__ unsigned int buf[1024];
  __shared__ unsigned int carry;
  if (threadIdx.x == 0) carry = 0;
  __syncthreads();
  for (int base = 0; base < W; base += 1024) {
    const int idx = base + threadIdx.x;
    const unsigned int v = idx < W ? rowcnt[idx] : 0u;
    buf[threadIdx.x] = v;
    __syncthreads();
    for (int off = 1; off < 1024; off <<= 1) {
      const unsigned int t = threadIdx.x >= off ? buf[threadIdx.x - off] : 0u;
      __syncthreads();
      buf[threadIdx.x] += t;
      __syncthreads();
    }
    if (idx < W) rowcnt[idx] = carry + buf[threadIdx.x] - v;
    __syncthreads();
    if (threadIdx.x == 1023) carry += buf[1023];
    __syncthreads();
  }
  if (threadIdx.x == 0) *total = carry;
}
__global__ void __launch_bounds__(256) k_hba_edge_emit(const double* __restrict__ H, int n, int W, const double* __restrict__ poses, const unsigned int* __restrict__ rowoff, long long cap,
                                                       int* __restrict__ eij, double* __restrict__ v6, double* __restrict__ rot, double* __restrict__ tra) {
  const int i = blockIdx.x, lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  __shared__ unsigned int wcnt[8];
  __shared__ unsigned int run;
  if (threadIdx.x == 0) run = rowoff[i];
  __syncthreads();
  for (int j0 = i + 1; j0 < W; j0 += blockDim.x) {
    const int j = j0 + threadIdx.x;
    double v[6];
    const bool ok = j < W && edge_ok(H, n, i, j, v);
    const unsigned int bal = __ballot_sync(0xffffffffu, ok);
    if (lane == 0) wcnt[warp] = __popc(bal);
    __syncthreads();
    unsigned int before = run;
    for (int w = 0; w < warp; w++) before += wcnt[w];
    const long long slot = (long long)before + __popc(bal & ((1u << lane) - 1u));
    if (ok && slot < cap) {
      eij[2 * slot] = i; eij[2 * slot + 1] = j;
      for (int k = 0; k < 6; k++) v6[6 * slot + k] = v[k];
      const rot3 Ri = load_rot(poses + 12 * i), Rj = load_rot(poses + 12 * j);
      const d3 dp = mk3(poses[12 * j + 9] - poses[12 * i + 9], poses[12 * j + 10] - poses[12 * i + 10], poses[12 * j + 11] - poses[12 * i + 11]);
      const d3 t = mulT(Ri, dp);
      tra[3 * slot] = t.x; tra[3 * slot + 1] = t.y; tra[3 * slot + 2] = t.z;
      const double a[9] = {Ri.r00, Ri.r01, Ri.r02, Ri.r10, Ri.r11, Ri.r12, Ri.r20, Ri.r21, Ri.r22}, b[9] = {Rj.r00, Rj.r01, Rj.r02, Rj.r10, Rj.r11, Rj.r12, Rj.r20, Rj.r21, Rj.r22};
      for (int r = 0; r < 3; r++) for (int c = 0; c < 3; c++) rot[9 * slot + 3 * r + c] = a[r] * b[c] + a[3 + r] * b[3 + c] + a[6 + r] * b[6 + c];   // R_i^T R_j
    }
    __syncthreads();
    if (threadIdx.x == 0) { unsigned int t = 0; for (int w = 0; w < 8; w++) t += wcnt[w]; run += t; }
    __syncthreads();
  }
}

extern "C" int vxs_hba_edges(vxs_ctx* ctx, int W, const double* poses12, int64_t cap, int32_t* edge_ij, double* v6, double* rot, double* tra, int64_t* n_edges) {
  if (!ctx || W <= 0 || !poses12 || cap < 0 || !n_edges) return VXS_ERR_ARG;
  const int n = 6 * W;
  if (ctx->hraw_n != n || ctx->hraw_S != 6 || ctx->Hraw.cap < size_t(n) * n)
    return vxs_fail(ctx, VXS_ERR_ARG, "vxs_hba_edges: the Hessian resident on this ctx is not a lidar-only 6W system of this W (run vxs_lidar_ba / vxs_hba_window first)");
  cudaSetDevice(ctx->device);
  const size_t c = size_t(std::max<int64_t>(cap, 1));
  VXS_CUDA(ctx, ctx->stage.reserve(c * 18 + size_t(W) * 12));
  VXS_CUDA(ctx, ctx->stage_i64.reserve(c + size_t(W) / 2 + 4));
  double* d_v6 = ctx->stage.p; double* d_rot = d_v6 + c * 6; double* d_tra = d_rot + c * 9; double* d_pose = d_tra + c * 3;
  int* d_eij = reinterpret_cast<int*>(ctx->stage_i64.p);
  unsigned int* d_row = reinterpret_cast<unsigned int*>(ctx->stage_i64.p + c);
  unsigned int* d_cnt = reinterpret_cast<unsigned int*>(ctx->flags.p + 8);
  VXS_CUDA(ctx, cudaMemcpyAsync(d_pose, poses12, size_t(W) * 96, cudaMemcpyHostToDevice, ctx->stream));
  VXS_LAUNCH(ctx, "k_hba_edges", k_hba_edge_count, unsigned(W), 256, 0, ctx->Hraw.p, n, W, d_row);
  VXS_LAUNCH(ctx, "k_hba_edges", k_hba_edge_scan, 1, 1024, 0, d_row, W, d_cnt);
  VXS_LAUNCH(ctx, "k_hba_edges", k_hba_edge_emit, unsigned(W), 256, 0, ctx->Hraw.p, n, W, d_pose, d_row, (long long)cap, d_eij, d_v6, d_rot, d_tra);
  unsigned int cnt = 0;
  VXS_CUDA(ctx, cudaMemcpyAsync(&cnt, d_cnt, sizeof cnt, cudaMemcpyDeviceToHost, ctx->stream));
  VXS_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  *n_edges = cnt;
  const size_t m = size_t(std::min<int64_t>(cnt, cap));
  if (m) {
    VXS_CUDA(ctx, cudaMemcpyAsync(edge_ij, d_eij, m * 8, cudaMemcpyDeviceToHost, ctx->stream));
    VXS_CUDA(ctx, cudaMemcpyAsync(v6, d_v6, m * 48, cudaMemcpyDeviceToHost, ctx->stream));
    VXS_CUDA(ctx, cudaMemcpyAsync(rot, d_rot, m * 72, cudaMemcpyDeviceToHost, ctx->stream));
    VXS_CUDA(ctx, cudaMemcpyAsync(tra, d_tra, m * 24, cudaMemcpyDeviceToHost, ctx->stream));
    VXS_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  }
  return VXS_OK;
}
