// ORACLE — TEST INFRASTRUCTURE ONLY (see vxo_core.hpp).  C entry points for ctypes (tests/, smoke(), bench.py cpu_baseline).
#include <chrono>
#include "../include/vxs.h"
#include "vxo_map.hpp"
// the product's HOST-compilable math header, only for vxo_hostmath_* (a CPU check of the GPU formulas)
#include "../voxel_slam_b200/csrc/vxs_math.cuh"

using namespace vxo;

struct OracleFactor { LidarFactor f; std::vector<VoxelId> ids; explicit OracleFactor(int w) : f(w) {} };

static std::vector<State> states_from_poses12(const double* p, int W) {
  std::vector<State> xs(W);
  for (int i = 0; i < W; i++) {
    double s[24] = {0};
    std::memcpy(s, p + 12 * i, 12 * sizeof(double));
    xs[i] = state_from(s);
  }
  return xs;
}
static std::vector<State> states_from_24(const double* p, int W) { std::vector<State> xs(W); for (int i = 0; i < W; i++) xs[i] = state_from(p + 24 * i); return xs; }
static void eig12_pack(const V3& w, const M3& U, double* o) { for (int k = 0; k < 3; k++) o[k] = w[k]; for (int r = 0; r < 3; r++) for (int c = 0; c < 3; c++) o[3 + 3 * r + c] = U(r, c); }

extern "C" {

// ---------------------------------------------------------------- primitives
void vxo_eig3(const double* A9_rowmajor, double* w3, double* U9_rowmajor) {
  M3 A, U; V3 w;
  for (int r = 0; r < 3; r++) for (int c = 0; c < 3; c++) A(r, c) = A9_rowmajor[3 * r + c];
  eig3_sym(A, w, U);
  for (int k = 0; k < 3; k++) w3[k] = w[k];
  for (int r = 0; r < 3; r++) for (int c = 0; c < 3; c++) U9_rowmajor[3 * r + c] = U(r, c);
}
void vxo_voxel_keys(const double* pw, int64_t n, double voxel_size, int64_t* xyz, uint64_t* hash) {
  for (int64_t i = 0; i < n; i++) {
    VoxelLoc k = voxel_key(v3(pw[3 * i], pw[3 * i + 1], pw[3 * i + 2]), voxel_size);
    xyz[3 * i] = k.x; xyz[3 * i + 1] = k.y; xyz[3 * i + 2] = k.z;
    if (hash) hash[i] = voxel_hash(k);
  }
}
void vxo_cluster_from_points(const double* pts, int64_t n, double* c10) { PC c; for (int64_t i = 0; i < n; i++) c.push(v3(pts[3 * i], pts[3 * i + 1], pts[3 * i + 2])); pc_pack(c, c10); }
void vxo_cluster_transform(const double* c10, const double* pose12, double* out10) {
  State x = states_from_poses12(pose12, 1)[0];
  PC o; o.transform(pc_unpack(c10), x.R, x.p);
  pc_pack(o, out10);
}
void vxo_so3_exp(const double* w, double* R9) { M3 R = so3_exp(v3(w[0], w[1], w[2])); for (int r = 0; r < 3; r++) for (int c = 0; c < 3; c++) R9[3 * r + c] = R(r, c); }
int vxo_ldlt_solve(const double* A_colmajor, const double* b, int n, double* x) {
  Mat A(n, n); std::memcpy(A.a.data(), A_colmajor, size_t(n) * n * 8);
  std::vector<double> bb(b, b + n), xx;
  bool ok = ldlt_solve(A, bb, xx);
  std::memcpy(x, xx.data(), size_t(n) * 8);
  return ok ? 0 : 1;
}

// ---------------------------------------------------------------- factor handle
void* vxo_factor_create(int W) { return new OracleFactor(W); }
void vxo_factor_destroy(void* h) { delete static_cast<OracleFactor*>(h); }
int64_t vxo_factor_size(void* h) { return int64_t(static_cast<OracleFactor*>(h)->f.size()); }
int vxo_factor_win(void* h) { return static_cast<OracleFactor*>(h)->f.win_size; }
// dense [n][W][10] push (reference layout)
void vxo_factor_push_dense(void* h, int64_t n, const double* clusters10, const double* fix10, const double* coe, const double* eig12, const double* sum10) {
  LidarFactor& f = static_cast<OracleFactor*>(h)->f;
  const int W = f.win_size;
  for (int64_t v = 0; v < n; v++) {
    std::vector<PC> pcs(W);
    for (int i = 0; i < W; i++) pcs[i] = pc_unpack(clusters10 + (size_t(v) * W + i) * 10);
    PC fix = fix10 ? pc_unpack(fix10 + size_t(v) * 10) : PC();
    V3 w = v3(eig12[12 * v], eig12[12 * v + 1], eig12[12 * v + 2]);
    M3 U; for (int r = 0; r < 3; r++) for (int c = 0; c < 3; c++) U(r, c) = eig12[12 * v + 3 + 3 * r + c];
    f.push_voxel(pcs, fix, coe ? coe[v] : 1.0, w, U, pc_unpack(sum10 + size_t(v) * 10));
  }
}
void vxo_factor_export(void* h, double* clusters10, double* fix10, double* coe, double* eig12, double* sum10, vxs_voxel_id* ids) {
  OracleFactor* of = static_cast<OracleFactor*>(h);
  LidarFactor& f = of->f;
  const int W = f.win_size;
  for (size_t v = 0; v < f.size(); v++) {
    if (clusters10) for (int i = 0; i < W; i++) pc_pack(f.plvec_voxels[v][i], clusters10 + (v * W + i) * 10);
    if (fix10) pc_pack(f.sig_vecs[v], fix10 + v * 10);
    if (coe) coe[v] = f.coeffs[v];
    if (eig12) eig12_pack(f.eig_values[v], f.eig_vectors[v], eig12 + v * 12);
    if (sum10) pc_pack(f.pcr_adds[v], sum10 + v * 10);
    if (ids && v < of->ids.size()) { ids[v].x = of->ids[v].x; ids[v].y = of->ids[v].y; ids[v].z = of->ids[v].z; ids[v].layer = of->ids[v].layer; ids[v].path = of->ids[v].path; }
  }
}
double vxo_factor_residual(void* h, const double* poses12) {  // evaluate_only_residual(xs, 0, size)
  LidarFactor& f = static_cast<OracleFactor*>(h)->f;
  auto xs = states_from_poses12(poses12, f.win_size);
  double r = 0; f.evaluate_only_residual(xs, 0, int(f.size()), r);
  return r;
}
double vxo_factor_hessian(void* h, const double* poses12, double* hess, double* jact) {  // acc_evaluate2(xs, 0, size)
  LidarFactor& f = static_cast<OracleFactor*>(h)->f;
  const int n = 6 * f.win_size;
  auto xs = states_from_poses12(poses12, f.win_size);
  Mat H(n, n); std::vector<double> J(n); double r = 0;
  f.acc_evaluate2(xs, 0, int(f.size()), H, J, r);
  if (hess) std::memcpy(hess, H.a.data(), size_t(n) * n * 8);
  if (jact) std::memcpy(jact, J.data(), size_t(n) * 8);
  return r;
}
// threaded versions with the reference's fork-join structure, for CPU-baseline timing; return seconds
double vxo_time_hessian(void* h, const double* poses12, int threads, int reps, double* r_out) {
  LidarFactor& f = static_cast<OracleFactor*>(h)->f;
  const int n = 6 * f.win_size;
  auto xs = states_from_poses12(poses12, f.win_size);
  Mat H(n, n); std::vector<double> J(n);
  auto t0 = std::chrono::steady_clock::now();
  double r = 0;
  for (int i = 0; i < reps; i++) r = lidar_hessian_threads(xs, f, threads, H, J);
  double s = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
  if (r_out) *r_out = r;
  return s / reps;
}
double vxo_time_residual(void* h, const double* poses12, int threads, int reps, double* r_out) {
  LidarFactor& f = static_cast<OracleFactor*>(h)->f;
  auto xs = states_from_poses12(poses12, f.win_size);
  auto t0 = std::chrono::steady_clock::now();
  double r = 0;
  for (int i = 0; i < reps; i++) r = lidar_residual_threads(xs, f, threads);
  double s = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
  if (r_out) *r_out = r;
  return s / reps;
}

// ---------------------------------------------------------------- LM drivers
static void copy_trace(const std::vector<LmTrace>& tr, vxs_lm_trace* out, int cap, int* len) {
  int n = 0;
  for (const LmTrace& t : tr) { if (out && n < cap) { out[n].r1 = t.r1; out[n].r2 = t.r2; out[n].u = t.u; out[n].v = t.v; out[n].q1 = t.q1; out[n].accepted = t.accepted; out[n].hess_built = 0; n++; } }
  if (len) *len = n;
}
int vxo_lidar_ba(void* h, double* poses12, int max_iter, int thd_num, double* hess_out, double* resis2, int* is_converge, vxs_lm_trace* trace, int trace_cap, int* trace_len) {
  LidarFactor& f = static_cast<OracleFactor*>(h)->f;
  const int W = f.win_size, n = 6 * W;
  auto xs = states_from_poses12(poses12, W);
  Mat hess; std::vector<double> resis; std::vector<LmTrace> tr; int status = 0;
  bool conv = lidar_ba_damping_iter(xs, f, &hess, resis, max_iter, thd_num, &tr, &status);
  if (status != 0) return VXS_ERR_TOO_FEW_VOXELS;
  for (int i = 0; i < W; i++) { double s[24]; state_to(xs[i], s); std::memcpy(poses12 + 12 * i, s, 12 * 8); }
  if (hess_out) std::memcpy(hess_out, hess.a.data(), size_t(n) * n * 8);
  if (resis2) { resis2[0] = resis.size() > 0 ? resis[0] : 0; resis2[1] = resis.size() > 1 ? resis[1] : 0; }
  if (is_converge) *is_converge = conv ? 1 : 0;
  copy_trace(tr, trace, trace_cap, trace_len);
  return 0;
}
int vxo_li_ba(void* h, double* states24, int with_gravity, int max_iter, double imu_coef, const vxs_imu_hooks* imu, double* hess_out, double* resis2, vxs_lm_trace* trace,
              int trace_cap, int* trace_len) {
  LidarFactor& f = static_cast<OracleFactor*>(h)->f;
  const int W = f.win_size, n = 15 * W + (with_gravity ? 3 : 0);
  auto xs = states_from_24(states24, W);
  ImuHooks hooks;
  hooks.eval = [&](const std::vector<State>& st, bool want_jac, double* blocks, double* gvec) {
    std::vector<double> flat(size_t(W) * 24);
    for (int i = 0; i < W; i++) state_to(st[i], flat.data() + 24 * i);
    double cost = 0;
    imu->eval(imu->user, flat.data(), W, with_gravity, want_jac ? 1 : 0, blocks, gvec, &cost);
    return cost;
  };
  hooks.update = [&](const double* dxi) { imu->update(imu->user, dxi, W); };
  hooks.rollback = [&]() { imu->rollback(imu->user); };
  Mat hess; std::vector<double> resis; std::vector<LmTrace> tr;
  li_ba_damping_iter(xs, f, hooks, imu_coef, with_gravity != 0, max_iter, &hess, &resis, &tr);
  for (int i = 0; i < W; i++) state_to(xs[i], states24 + 24 * i);
  if (hess_out) std::memcpy(hess_out, hess.a.data(), size_t(n) * n * 8);
  if (resis2) {
    if (with_gravity) { resis2[0] = resis.size() > 0 ? resis[0] : 0; resis2[1] = resis.size() > 1 ? resis[1] : 0; }
    else { resis2[0] = tr.empty() ? 0 : tr.front().r1; resis2[1] = tr.empty() ? 0 : tr.back().r2; }
  }
  copy_trace(tr, trace, trace_cap, trace_len);
  return 0;
}

// ---------------------------------------------------------------- voxel map
static MapParams to_map_params(const vxs_map_params* p, int with_cov_add) {
  MapParams m;
  m.voxel_size = p->voxel_size; m.min_eigen_value = p->min_eigen_value; m.max_layer = p->max_layer; m.with_cov_add = with_cov_add != 0;
  for (int k = 0; k < 4; k++) { m.min_point[k] = p->min_point[k]; m.plane_thre[k] = p->plane_thre[k]; }
  for (int k = 4; k < 8; k++) m.plane_thre[k] = p->plane_thre[3];
  return m;
}
static GbaParams to_gba_params(const vxs_map_params* p) {
  GbaParams g;
  g.voxel_size = p->voxel_size; g.min_eigen_value = p->min_eigen_value; g.max_layer = p->max_layer;
  for (int k = 0; k < 8; k++) g.eigen_value_array[k] = p->plane_thre[k < 4 ? k : 3];
  return g;
}
// cut_voxel x W + recut + tras_opt (build from scratch).  Returns a factor handle.  var_diag>0 also runs the cov_add by-product.
void* vxo_build_window_factor(const vxs_map_params* mp, const double* pts_body, const int64_t* scan_offsets, const double* poses12, int W, const double* fix_pts,
                              int64_t n_fix, int threads, double var_diag, double* seconds_cut_recut) {
  MapParams m = to_map_params(mp, var_diag > 0);
  auto xs = states_from_poses12(poses12, W);
  std::vector<std::vector<PV>> scans(W);
  M3 var = m3_zero(); var(0, 0) = var(1, 1) = var(2, 2) = var_diag > 0 ? var_diag : 0.0;
  for (int i = 0; i < W; i++) {
    scans[i].resize(size_t(scan_offsets[i + 1] - scan_offsets[i]));
    for (size_t k = 0; k < scans[i].size(); k++) { const double* p = pts_body + 3 * (scan_offsets[i] + k); scans[i][k].pnt = v3(p[0], p[1], p[2]); scans[i][k].var = var; }
  }
  OracleFactor* of = new OracleFactor(W);
  LocalMap map;
  auto t0 = std::chrono::steady_clock::now();
  if (fix_pts && n_fix > 0) {
    std::vector<PV> fx; fx.resize(size_t(n_fix));
    for (int64_t k = 0; k < n_fix; k++) { fx[k].pnt = v3(fix_pts[3 * k], fix_pts[3 * k + 1], fix_pts[3 * k + 2]); fx[k].var = var; }
    cut_voxel_fix(map, fx, W, m);
  }
  build_window_factor(map, scans, xs, m, threads, of->f, &of->ids);
  if (seconds_cut_recut) *seconds_cut_recut = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
  local_map_free(map);
  return of;
}
void* vxo_build_gba_factor(const vxs_map_params* mp, const float* xyz, int stride_floats, const int64_t* kf_offsets, const double* poses12, int W, int threads, double* seconds) {
  GbaParams gp = to_gba_params(mp);
  auto xs = states_from_poses12(poses12, W);
  OracleFactor* of = new OracleFactor(W);
  GbaMap map;
  auto t0 = std::chrono::steady_clock::now();
  std::vector<float> packed;
  for (int i = 0; i < W; i++) {
    const size_t n = size_t(kf_offsets[i + 1] - kf_offsets[i]);
    packed.resize(n * 3);
    for (size_t k = 0; k < n; k++) for (int j = 0; j < 3; j++) packed[3 * k + j] = xyz[size_t(kf_offsets[i] + k) * stride_floats + j];
    gba_cut_voxel(map, xs[i], packed.data(), n, i, W, gp);
  }
  gba_multi_recut(map, of->f, threads, gp, &of->ids);
  if (seconds) *seconds = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
  return of;
}
// PGO edges from a raw Hessian (voxelslam.cpp:2405-2427)
int64_t vxo_hba_edges(const double* hess, int W, const double* poses12, int64_t cap, int32_t* eij, double* v6, double* rot, double* tra) {
  const int n = 6 * W;
  auto xs = states_from_poses12(poses12, W);
  int64_t m = 0;
  for (int i = 0; i < W - 1; i++) for (int j = i + 1; j < W; j++) {
    bool add = true; double v[6];
    for (int k = 0; k < 6; k++) { double hc = std::fabs(hess[size_t(6 * j + k) * n + 6 * i + k]); if (hc < 1e-6) { add = false; break; } v[k] = 1.0 / hc; }
    if (!add) continue;
    if (m < cap) {
      eij[2 * m] = i; eij[2 * m + 1] = j;
      for (int k = 0; k < 6; k++) v6[6 * m + k] = v[k];
      V3 t = tr(xs[i].R) * (xs[j].p - xs[i].p); M3 r = tr(xs[i].R) * xs[j].R;
      for (int a = 0; a < 3; a++) { tra[3 * m + a] = t[a]; for (int b = 0; b < 3; b++) rot[9 * m + 3 * a + b] = r(a, b); }
    }
    m++;
  }
  return m;
}
int vxo_hba_window(const vxs_map_params* coarse, const vxs_map_params* fine, const float* xyz, int stride_floats, const int64_t* kf_offsets, double* poses12, int W, int max_iter,
                   int thread_num, double* hess_out, double* resis_log, int* outer_iters) {
  auto xs = states_from_poses12(poses12, W);
  std::vector<std::vector<float>> clouds(W);
  Keyframes kf;
  for (int i = 0; i < W; i++) {
    const size_t n = size_t(kf_offsets[i + 1] - kf_offsets[i]);
    clouds[i].resize(n * 3);
    for (size_t k = 0; k < n; k++) for (int j = 0; j < 3; j++) clouds[i][3 * k + j] = xyz[size_t(kf_offsets[i] + k) * stride_floats + j];
    kf.xyz.push_back(clouds[i].data()); kf.npts.push_back(n);
  }
  Mat hess; std::vector<double> log;
  int it = hba_window(xs, kf, to_gba_params(coarse), to_gba_params(fine), max_iter, thread_num, hess, nullptr, &log);
  if (it < 0) return VXS_ERR_TOO_FEW_VOXELS;
  for (int i = 0; i < W; i++) { double s[24]; state_to(xs[i], s); std::memcpy(poses12 + 12 * i, s, 12 * 8); }
  if (hess_out && hess.r > 0) std::memcpy(hess_out, hess.a.data(), size_t(hess.r) * hess.c * 8);
  if (resis_log) std::memcpy(resis_log, log.data(), log.size() * 8);
  if (outer_iters) *outer_iters = it;
  return 0;
}

// ---------------------------------------------------------------- CPU execution of the GPU formulas (vxs_math.cuh), test-only
// Hessian/gradient of a dense factor through entry_jacobian + the rank-3 identity, to be compared with acc_evaluate2 above.
void vxo_hostmath_hessian(void* h, const double* poses12, double* hess, double* jact) {
  LidarFactor& f = static_cast<OracleFactor*>(h)->f;
  const int W = f.win_size, n = 6 * W;
  Mat H(n, n); std::vector<double> J(n, 0.0);
  for (size_t a = 0; a < f.size(); a++) {
    double lam[3] = {f.eig_values[a][0], f.eig_values[a][1], f.eig_values[a][2]};
    const M3& U = f.eig_vectors[a];
    vxs::voxel_consts kc = vxs::make_voxel_consts(lam, vxs::mk3(U(0, 0), U(1, 0), U(2, 0)), vxs::mk3(U(0, 1), U(1, 1), U(2, 1)), vxs::mk3(U(0, 2), U(1, 2), U(2, 2)),
                                                  vxs::mk3(f.pcr_adds[a].v[0], f.pcr_adds[a].v[1], f.pcr_adds[a].v[2]), double(f.pcr_adds[a].N), f.coeffs[a]);
    std::vector<vxs::entry_out> outs(W);
    std::vector<int> present;
    for (int i = 0; i < W; i++) if (f.plvec_voxels[a][i].N != 0) {
      double c10[10]; pc_pack(f.plvec_voxels[a][i], c10);
      vxs::cluster c; c.P.xx = c10[0]; c.P.xy = c10[1]; c.P.xz = c10[2]; c.P.yy = c10[3]; c.P.yz = c10[4]; c.P.zz = c10[5]; c.v = vxs::mk3(c10[6], c10[7], c10[8]); c.n = c10[9];
      vxs::rot3 R = vxs::load_rot(poses12 + 12 * i);
      vxs::entry_jacobian(kc, c, R, vxs::mk3(poses12[12 * i + 9], poses12[12 * i + 10], poses12[12 * i + 11]), outs[i]);
      present.push_back(i);
    }
    for (int i : present) {
      const vxs::entry_out& o = outs[i];
      for (int k = 0; k < 6; k++) J[6 * i + k] += f.coeffs[a] * o.g[k];
      for (int r = 0; r < 3; r++) for (int c = 0; c < 3; c++) {
        H(6 * i + r, 6 * i + c) += o.Drr[3 * r + c];
        H(6 * i + r, 6 * i + 3 + c) += o.Drt[3 * r + c];
        H(6 * i + 3 + c, 6 * i + r) += o.Drt[3 * r + c];
      }
      const int tt[3][3] = {{0, 1, 2}, {1, 3, 4}, {2, 4, 5}};
      for (int r = 0; r < 3; r++) for (int c = 0; c < 3; c++) H(6 * i + 3 + r, 6 * i + 3 + c) += o.Dtt[tt[r][c]];
      for (int j : present) if (j >= i)
        for (int r = 0; r < 6; r++) for (int c = 0; c < 6; c++) H(6 * i + r, 6 * j + c) -= o.x[r] * outs[j].x[c] + o.x[6 + r] * outs[j].x[6 + c] + o.x[12 + r] * outs[j].x[12 + c];
    }
  }
  for (int i = 1; i < W; i++) for (int j = 0; j < i; j++) for (int r = 0; r < 6; r++) for (int c = 0; c < 6; c++) H(6 * i + r, 6 * j + c) = H(6 * j + c, 6 * i + r);
  std::memcpy(hess, H.a.data(), size_t(n) * n * 8);
  std::memcpy(jact, J.data(), size_t(n) * 8);
}
void vxo_hostmath_eig3(const double* sym6, double* w3, double* U9_rowmajor) {
  vxs::sym3 C; C.xx = sym6[0]; C.xy = sym6[1]; C.xz = sym6[2]; C.yy = sym6[3]; C.yz = sym6[4]; C.zz = sym6[5];
  vxs::d3 u0, u1, u2;
  vxs::eig3_jacobi(C, w3, u0, u1, u2);
  U9_rowmajor[0] = u0.x; U9_rowmajor[1] = u1.x; U9_rowmajor[2] = u2.x; U9_rowmajor[3] = u0.y; U9_rowmajor[4] = u1.y; U9_rowmajor[5] = u2.y;
  U9_rowmajor[6] = u0.z; U9_rowmajor[7] = u1.z; U9_rowmajor[8] = u2.z;
}
void vxo_hostmath_transform(const double* c10, const double* pose12, double* out10) {
  vxs::cluster c, acc; c.P.xx = c10[0]; c.P.xy = c10[1]; c.P.xz = c10[2]; c.P.yy = c10[3]; c.P.yz = c10[4]; c.P.zz = c10[5]; c.v = vxs::mk3(c10[6], c10[7], c10[8]); c.n = c10[9];
  acc.P.xx = acc.P.xy = acc.P.xz = acc.P.yy = acc.P.yz = acc.P.zz = 0; acc.v = vxs::mk3(0, 0, 0); acc.n = 0;
  vxs::cluster_transform_acc(c, vxs::load_rot(pose12), vxs::mk3(pose12[9], pose12[10], pose12[11]), acc);
  out10[0] = acc.P.xx; out10[1] = acc.P.xy; out10[2] = acc.P.xz; out10[3] = acc.P.yy; out10[4] = acc.P.yz; out10[5] = acc.P.zz; out10[6] = acc.v.x; out10[7] = acc.v.y; out10[8] = acc.v.z; out10[9] = acc.n;
}


// tools.hpp:201-302; mode 0 = down_sampling_voxel, 1 = down_sampling_close.  Returns the number of cells (-1: untouched); out arrays hold min(cap, cells).
int64_t vxo_down_sampling(int mode, const float* pts, int stride, int64_t n, double voxel_size, float* xyz_out, float* cnt_out, int64_t* idx_out, int64_t cap) {
  std::vector<DsPoint> out;
  const bool done = mode == 0 ? down_sampling_voxel(pts, stride, n, voxel_size, out) : down_sampling_close(pts, stride, n, voxel_size, out);
  if (!done) return -1;
  for (int64_t i = 0; i < int64_t(out.size()) && i < cap; i++) {
    xyz_out[3 * i] = out[i].x; xyz_out[3 * i + 1] = out[i].y; xyz_out[3 * i + 2] = out[i].z; cnt_out[i] = out[i].cnt; idx_out[i] = out[i].idx;
  }
  return int64_t(out.size());
}

// voxelslam.cpp:2428-2447 submap merge; returns the number of cells (or n when voxel_size < 0.001: the merged cloud, untouched)
int64_t vxo_submap_merge(const float* pts, int stride, const int64_t* kf_offsets, const double* poses12, int W, double voxel_size, float* xyz_out, float* cnt_out, int64_t* idx_out,
                         int64_t cap) {
  std::vector<float> merged; std::vector<DsPoint> out;
  const bool done = submap_merge(pts, stride, kf_offsets, poses12, W, voxel_size, merged, out);
  if (!done) {
    const int64_t n = kf_offsets[W];
    for (int64_t i = 0; i < n && i < cap; i++) { xyz_out[3 * i] = merged[3 * i]; xyz_out[3 * i + 1] = merged[3 * i + 1]; xyz_out[3 * i + 2] = merged[3 * i + 2]; cnt_out[i] = 0; idx_out[i] = i; }
    return n;
  }
  for (int64_t i = 0; i < int64_t(out.size()) && i < cap; i++) {
    xyz_out[3 * i] = out[i].x; xyz_out[3 * i + 1] = out[i].y; xyz_out[3 * i + 2] = out[i].z; cnt_out[i] = out[i].cnt; idx_out[i] = out[i].idx;
  }
  return int64_t(out.size());
}

// voxel_map.hpp:23-64 down_sampling_pvec
int64_t vxo_down_sampling_pvec(const double* pv, int stride, int64_t n, double voxel_size, float* xyz_out, float* nrm_out, float* cnt_out, int64_t* idx_out, int64_t cap) {
  std::vector<DsPvec> out;
  down_sampling_pvec(pv, stride, n, voxel_size, out);
  for (int64_t i = 0; i < int64_t(out.size()) && i < cap; i++) {
    xyz_out[3 * i] = out[i].x; xyz_out[3 * i + 1] = out[i].y; xyz_out[3 * i + 2] = out[i].z;
    nrm_out[3 * i] = out[i].nx; nrm_out[3 * i + 1] = out[i].ny; nrm_out[3 * i + 2] = out[i].nz; cnt_out[i] = out[i].cnt; idx_out[i] = out[i].idx;
  }
  return int64_t(out.size());
}

// ---------------------------------------------------------------- stateful local map (oracle only so far)
// Build the window map from scratch with per-point variances (cov_add by-product on), recut + tras_opt (sets opt_state), then
// margi(win_count = W, mgsize) with the identity slot map and the factor as built — i.e. the map state the odometry sees after a window slide.
struct OracleLocalMap { LocalMap map; MapParams mp; int W; LidarFactor f; OracleLocalMap(int w) : W(w), f(w) {} ~OracleLocalMap() { local_map_free(map); } };
void* vxo_local_map_build(const vxs_map_params* mp, const double* pts_body, const int64_t* scan_offsets, const double* poses12, int W, double var_diag, int mgsize) {
  OracleLocalMap* h = new OracleLocalMap(W);
  h->mp = to_map_params(mp, true);
  auto xs = states_from_poses12(poses12, W);
  std::vector<std::vector<PV>> scans(W);
  M3 var = m3_zero(); var(0, 0) = var(1, 1) = var(2, 2) = var_diag;
  for (int i = 0; i < W; i++) {
    scans[i].resize(size_t(scan_offsets[i + 1] - scan_offsets[i]));
    for (size_t k = 0; k < scans[i].size(); k++) { const double* p = pts_body + 3 * (scan_offsets[i] + k); scans[i][k].pnt = v3(p[0], p[1], p[2]); scans[i][k].var = var; }
  }
  build_window_factor(h->map, scans, xs, h->mp, 1, h->f, nullptr);
  std::vector<int> slot(W);
  for (int i = 0; i < W; i++) slot[i] = i;
  for (auto& kv : h->map) if (!kv.second->margi(W, mgsize, xs, h->f, slot, h->mp)) { delete h; return nullptr; }
  return h;
}
void vxo_local_map_free(void* h) { delete static_cast<OracleLocalMap*>(h); }
static void collect_planes(OctoTree* o, std::vector<OctoTree*>& out) {
  if (o->octo_state == 0) { if (o->is_plane && o->plane.radius > 0) out.push_back(o); return; }
  for (auto c : o->leaves) if (c) collect_planes(c, out);
}
// per plane leaf (row of 52 doubles): center3 normal3 plane_var36 radius N voxel_center3 half_length cov_add-trace eig3
int64_t vxo_local_map_planes(void* hh, double* rows52, int64_t cap) {
  OracleLocalMap* h = static_cast<OracleLocalMap*>(hh);
  std::vector<OctoTree*> pl;
  for (auto& kv : h->map) collect_planes(kv.second, pl);
  for (int64_t i = 0; i < int64_t(pl.size()) && i < cap; i++) {
    double* r = rows52 + 52 * i; const OctoTree* o = pl[i];
    for (int k = 0; k < 3; k++) { r[k] = o->plane.center[k]; r[3 + k] = o->plane.normal[k]; r[44 + k] = o->voxel_center[k]; r[49 + k] = o->eig_value[k]; }
    for (int k = 0; k < 36; k++) r[6 + k] = o->plane.plane_var[k];
    r[42] = o->plane.radius; r[43] = o->pcr_add.N; r[47] = double(o->quater_length) * 2;
    double t = 0; for (int k = 0; k < 9; k++) t += o->cov_add[k * 9 + k];
    r[48] = t;
  }
  return int64_t(pl.size());
}
// the points (world) and variances that formed plane leaf `idx` are not kept by the tree; tests rebuild them from the scene instead.
int vxo_local_map_odom_accumulate(void* hh, const double* pv12, int64_t n, const double* pose12, const double* rot_var9, const double* tsl_var9, int passes, double* HTH36,
                                  double* HTz6, double* nnt9, int32_t* flags) {
  OracleLocalMap* h = static_cast<OracleLocalMap*>(hh);
  std::vector<PV> pvec; pvec.resize(size_t(n));
  for (int64_t i = 0; i < n; i++) {
    const double* p = pv12 + 12 * i;
    pvec[i].pnt = v3(p[0], p[1], p[2]);
    for (int r = 0; r < 3; r++) for (int c = 0; c < 3; c++) pvec[i].var(r, c) = p[3 + 3 * r + c];
  }
  State x = states_from_poses12(pose12, 1)[0];
  M3 rv, tv;
  for (int r = 0; r < 3; r++) for (int c = 0; c < 3; c++) { rv(r, c) = rot_var9[3 * r + c]; tv(r, c) = tsl_var9[3 * r + c]; }
  std::vector<OctoTree*> octos;
  int m = 0;
  for (int pass = 0; pass < passes; pass++) m = odom_accumulate(h->map, pvec, x, rv, tv, h->mp.voxel_size, octos, HTH36, HTz6, nnt9);   // later passes use the leaf cache
  if (flags) for (int64_t i = 0; i < n; i++) flags[i] = 0;
  if (flags) {   // recompute the per-point flags once more for the tests (same calls, cache warm)
    const M3 Rt = tr(x.R);
    for (int64_t i = 0; i < n; i++) {
      const M3 phat = hat(pvec[i].pnt);
      const M3 var_world = (x.R * pvec[i].var * Rt + phat * rv * tr(phat)) + tv;
      const V3 wld = x.R * pvec[i].pnt + x.p;
      double sd = 0; const Plane* pla = nullptr; OctoTree* oc = nullptr;
      flags[i] = match(h->map, wld, pla, var_world, sd, oc, h->mp.voxel_size);
    }
  }
  return m;
}

// ---------------------------------------------------------------- sliding-window simulator (map side of voxelslam.cpp:1599-1686)
// var_init / pvec_update (voxelslam.hpp:187-214)
void vxo_var_init(const float* pts, int stride, int64_t n, const double* ext_R9, const double* ext_p3, double dept_err, double beam_err, double* pv12) {
  M3 R; for (int r = 0; r < 3; r++) for (int c = 0; c < 3; c++) R(r, c) = ext_R9[3 * r + c];
  std::vector<PV> out;
  var_init(R, v3(ext_p3[0], ext_p3[1], ext_p3[2]), pts, stride, n, dept_err, beam_err, out);
  for (int64_t i = 0; i < n; i++) { for (int k = 0; k < 3; k++) pv12[12 * i + k] = out[size_t(i)].pnt[k]; for (int r = 0; r < 3; r++) for (int c = 0; c < 3; c++) pv12[12 * i + 3 + 3 * r + c] = out[size_t(i)].var(r, c); }
}
void vxo_pvec_update(double* pv12, int64_t n, const double* pose12, const double* rot_var9, const double* tsl_var9, double* pwld) {
  std::vector<PV> pv(static_cast<size_t>(n));
  for (int64_t i = 0; i < n; i++) { pv[size_t(i)].pnt = v3(pv12[12 * i], pv12[12 * i + 1], pv12[12 * i + 2]); for (int r = 0; r < 3; r++) for (int c = 0; c < 3; c++) pv[size_t(i)].var(r, c) = pv12[12 * i + 3 + 3 * r + c]; }
  State x = states_from_poses12(pose12, 1)[0];
  M3 rv, tv; for (int r = 0; r < 3; r++) for (int c = 0; c < 3; c++) { rv(r, c) = rot_var9[3 * r + c]; tv(r, c) = tsl_var9[3 * r + c]; }
  std::vector<V3> pw;
  pvec_update(pv, x, rv, tv, pw);
  for (int64_t i = 0; i < n; i++) { for (int r = 0; r < 3; r++) for (int c = 0; c < 3; c++) pv12[12 * i + 3 + 3 * r + c] = pv[size_t(i)].var(r, c); for (int k = 0; k < 3; k++) pwld[3 * i + k] = pw[size_t(i)][k]; }
}

void* vxo_sliding_sim_create(const vxs_map_params* mp, int win_size, int mgsize, int max_points) {
  MapParams m = to_map_params(mp, true);
  m.max_points = max_points;
  return new SlidingWindowSim(m, win_size, mgsize);
}
void vxo_sliding_sim_free(void* h) { delete static_cast<SlidingWindowSim*>(h); }
void vxo_sliding_sim_add_scan(void* h, const double* pts_body, int64_t n, const double* pose12, double var_diag) {
  SlidingWindowSim* sim = static_cast<SlidingWindowSim*>(h);
  std::vector<PV> scan; scan.resize(size_t(n));
  M3 var = m3_zero(); var(0, 0) = var(1, 1) = var(2, 2) = var_diag;
  for (int64_t k = 0; k < n; k++) { scan[k].pnt = v3(pts_body[3 * k], pts_body[3 * k + 1], pts_body[3 * k + 2]); scan[k].var = var; }
  sim->add_scan(scan, states_from_poses12(pose12, 1)[0]);
}
// full pointVar records (pnt 3 | var 3x3 row-major) and an optional pose-only BA between tras_opt and margi
void vxo_sliding_sim_add_scan_pv(void* h, const double* pv12, int64_t n, const double* pose12, int ba_iters) {
  SlidingWindowSim* sim = static_cast<SlidingWindowSim*>(h);
  std::vector<PV> scan; scan.resize(size_t(n));
  for (int64_t k = 0; k < n; k++) { const double* p = pv12 + 12 * k; scan[k].pnt = v3(p[0], p[1], p[2]); for (int r = 0; r < 3; r++) for (int c = 0; c < 3; c++) scan[k].var(r, c) = p[3 + 3 * r + c]; }
  sim->add_scan(scan, states_from_poses12(pose12, 1)[0], ba_iters);
}
// copy of the factor the last add_scan extracted (tras_opt), as a factor handle
void* vxo_sliding_sim_factor(void* h) {
  SlidingWindowSim* sim = static_cast<SlidingWindowSim*>(h);
  OracleFactor* of = new OracleFactor(sim->win_size);
  of->f = sim->voxhess;
  return of;
}
int64_t vxo_sliding_sim_planes(void* hh, double* rows52, int64_t cap) {
  SlidingWindowSim* sim = static_cast<SlidingWindowSim*>(hh);
  std::vector<OctoTree*> pl;
  for (auto& kv : sim->surf_map) collect_planes(kv.second, pl);
  for (int64_t i = 0; i < int64_t(pl.size()) && i < cap; i++) {
    double* r = rows52 + 52 * i; const OctoTree* o = pl[i];
    for (int k = 0; k < 3; k++) { r[k] = o->plane.center[k]; r[3 + k] = o->plane.normal[k]; r[44 + k] = o->voxel_center[k]; r[49 + k] = o->eig_value[k]; }
    for (int k = 0; k < 36; k++) r[6 + k] = o->plane.plane_var[k];
    r[42] = o->plane.radius; r[43] = o->pcr_add.N; r[47] = double(o->quater_length) * 2;
    double t = 0; for (int k = 0; k < 9; k++) t += o->cov_add[k * 9 + k];
    r[48] = t;
  }
  return int64_t(pl.size());
}
int vxo_sliding_sim_odom_accumulate(void* hh, const double* pv12, int64_t n, const double* pose12, const double* rot_var9, const double* tsl_var9, double* HTH36, double* HTz6, double* nnt9,
                                    int32_t* flags) {
  SlidingWindowSim* sim = static_cast<SlidingWindowSim*>(hh);
  std::vector<PV> pvec; pvec.resize(size_t(n));
  for (int64_t i = 0; i < n; i++) { const double* p = pv12 + 12 * i; pvec[i].pnt = v3(p[0], p[1], p[2]); for (int r = 0; r < 3; r++) for (int c = 0; c < 3; c++) pvec[i].var(r, c) = p[3 + 3 * r + c]; }
  State x = states_from_poses12(pose12, 1)[0];
  M3 rv, tv;
  for (int r = 0; r < 3; r++) for (int c = 0; c < 3; c++) { rv(r, c) = rot_var9[3 * r + c]; tv(r, c) = tsl_var9[3 * r + c]; }
  std::vector<OctoTree*> octos;
  const int m = odom_accumulate(sim->surf_map, pvec, x, rv, tv, sim->mp_.voxel_size, octos, HTH36, HTz6, nnt9);
  if (flags) {
    const M3 Rt = tr(x.R);
    for (int64_t i = 0; i < n; i++) {
      const M3 phat = hat(pvec[i].pnt);
      const M3 var_world = (x.R * pvec[i].var * Rt + phat * rv * tr(phat)) + tv;
      const V3 wld = x.R * pvec[i].pnt + x.p;
      double sd = 0; const Plane* pla = nullptr; OctoTree* oc = nullptr;
      flags[i] = match(sim->surf_map, wld, pla, var_world, sd, oc, sim->mp_.voxel_size);
    }
  }
  return m;
}
static void collect_leaves(OctoTree* o, std::vector<OctoTree*>& out) {
  if (o->octo_state == 0) { out.push_back(o); return; }
  for (auto c : o->leaves) if (c) collect_leaves(c, out);
}
// state: win_count, win_base, ring[W], poses of x_buf (win_count x 12).  Leaves of ALL roots of the map (slide or not): per leaf a row of
// 12 + 10 + 10 + 10*W doubles: voxel_center3 half layer is_plane isexist has_sw in_slide opt_state last_num n_point_fix | pcr_add10 | pcr_fix10 | pcrs_local[ring[i]] for i < W
int64_t vxo_sliding_sim_state(void* h, int32_t* head /* win_count, win_base, ring[W] */, double* poses12, double* rows, int64_t cap) {
  SlidingWindowSim* sim = static_cast<SlidingWindowSim*>(h);
  const int W = sim->win_size;
  head[0] = sim->win_count; head[1] = sim->win_base;
  for (int i = 0; i < W; i++) head[2 + i] = sim->ring[i];
  for (int i = 0; i < sim->win_count; i++) {
    const State& x = sim->x_buf[i];
    for (int r = 0; r < 3; r++) for (int c = 0; c < 3; c++) poses12[12 * i + 3 * r + c] = x.R(r, c);
    for (int k = 0; k < 3; k++) poses12[12 * i + 9 + k] = x.p[k];
  }
  std::vector<OctoTree*> lv;
  std::vector<char> in_slide;
  for (auto& kv : sim->surf_map) {
    const size_t b = lv.size();
    collect_leaves(kv.second, lv);
    in_slide.resize(lv.size(), sim->slide.count(kv.first) ? 1 : 0);
    (void)b;
  }
  const int rw = 32 + 10 * W;
  for (int64_t t = 0; t < int64_t(lv.size()) && t < cap; t++) {
    double* r = rows + size_t(t) * rw; const OctoTree* o = lv[t];
    for (int k = 0; k < 3; k++) r[k] = o->voxel_center[k];
    r[3] = double(o->quater_length) * 2; r[4] = o->layer; r[5] = o->is_plane; r[6] = o->isexist; r[7] = o->sw != nullptr; r[8] = in_slide[t]; r[9] = o->opt_state; r[10] = o->last_num;
    r[11] = double(o->point_fix.size());
    pc_pack(o->pcr_add, r + 12); pc_pack(o->pcr_fix, r + 22);
    for (int i = 0; i < W; i++) { if (o->sw) pc_pack(o->sw->pcrs_local[sim->ring[i]], r + 32 + 10 * i); else for (int k = 0; k < 10; k++) r[32 + 10 * i + k] = 0; }
  }
  return int64_t(lv.size());
}

}  // extern "C"
