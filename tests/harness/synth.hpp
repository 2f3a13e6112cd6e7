// Synthetic-workload harness (CPU, C++17, no dependencies).  NOT part of the hot path and NOT the oracle:
//  * seeded scene generator for the "3-plane corner room" workloads of SURVEY.md §8(d) / BASELINE.md
//  * a stand-in for the reference's UNCHANGED IMU preintegration factor (preintegration.hpp:11-331), which stays on
//    the CPU behind the vxs_imu_hooks callback of include/vxs.h.  bench.py and the tests need *some* IMU factor to
//    drive the LI-BA solvers; in a real integration the callback wraps the reference's own IMU_PRE objects
//    (INTEGRATION.md).  The same callback object feeds both the CUDA path and the oracle, so its arithmetic is not a
//    parity matter.
#pragma once
#include <cmath>
#include <cstdint>
#include <cstring>
#include <vector>

namespace vxh {

// ------------------------------------------------------------------ RNG (fixed algorithms: results do not depend on libstdc++)
struct SplitMix64 {
  uint64_t s;
  explicit SplitMix64(uint64_t seed) : s(seed) {}
  uint64_t next() { uint64_t z = (s += 0x9E3779B97F4A7C15ull); z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull; z = (z ^ (z >> 27)) * 0x94D049BB133111EBull; return z ^ (z >> 31); }
  double uni() { return double(next() >> 11) * (1.0 / 9007199254740992.0); }  // [0,1)
  double gauss() {  // Box–Muller, one value per call
    double u1 = uni(), u2 = uni();
    if (u1 < 1e-300) u1 = 1e-300;
    return std::sqrt(-2.0 * std::log(u1)) * std::cos(6.283185307179586 * u2);
  }
};

// ------------------------------------------------------------------ tiny 3x3 helpers (row-major)
inline void mat3_mul(const double* A, const double* B, double* C) { for (int i = 0; i < 3; i++) for (int j = 0; j < 3; j++) C[3 * i + j] = A[3 * i] * B[j] + A[3 * i + 1] * B[3 + j] + A[3 * i + 2] * B[6 + j]; }
inline void mat3_vec(const double* A, const double* x, double* y) { for (int i = 0; i < 3; i++) y[i] = A[3 * i] * x[0] + A[3 * i + 1] * x[1] + A[3 * i + 2] * x[2]; }
inline void mat3_tvec(const double* A, const double* x, double* y) { for (int i = 0; i < 3; i++) y[i] = A[i] * x[0] + A[3 + i] * x[1] + A[6 + i] * x[2]; }
inline void mat3_t(const double* A, double* T) { for (int i = 0; i < 3; i++) for (int j = 0; j < 3; j++) T[3 * i + j] = A[3 * j + i]; }
inline void hat3(const double* v, double* H) { H[0] = 0; H[1] = -v[2]; H[2] = v[1]; H[3] = v[2]; H[4] = 0; H[5] = -v[0]; H[6] = -v[1]; H[7] = v[0]; H[8] = 0; }
inline void exp3(const double* w, double* R) {
  double n = std::sqrt(w[0] * w[0] + w[1] * w[1] + w[2] * w[2]);
  double I[9] = {1, 0, 0, 0, 1, 0, 0, 0, 1};
  if (n < 1e-11) { std::memcpy(R, I, sizeof I); return; }
  double a[3] = {w[0] / n, w[1] / n, w[2] / n}, K[9], KK[9];
  hat3(a, K); mat3_mul(K, K, KK);
  double s = std::sin(n), c = 1.0 - std::cos(n);
  for (int i = 0; i < 9; i++) R[i] = I[i] + s * K[i] + c * KK[i];
}
inline void log3(const double* R, double* w) {
  double trR = R[0] + R[4] + R[8];
  double th = (trR > 3.0 - 1e-6) ? 0.0 : std::acos(0.5 * (trR - 1));
  double K[3] = {R[7] - R[5], R[2] - R[6], R[3] - R[1]};
  double f = (std::fabs(th) < 0.001) ? 0.5 : 0.5 * th / std::sin(th);
  for (int i = 0; i < 3; i++) w[i] = f * K[i];
}
inline void jr3(const double* v_in, double* J) {  // right Jacobian of SO(3)
  double v[3] = {v_in[0], v_in[1], v_in[2]};
  double ang = std::sqrt(v[0] * v[0] + v[1] * v[1] + v[2] * v[2]);
  double I[9] = {1, 0, 0, 0, 1, 0, 0, 0, 1};
  if (ang < 1e-9) { std::memcpy(J, I, sizeof I); return; }
  for (double& x : v) x /= ang;
  double ra = std::sin(ang) / ang, H[9];
  hat3(v, H);
  for (int i = 0; i < 3; i++) for (int j = 0; j < 3; j++) J[3 * i + j] = ra * I[3 * i + j] + (1 - ra) * v[i] * v[j] - (1 - std::cos(ang)) / ang * H[3 * i + j];
}
inline void jr3_inv(const double* R, double* J) {
  double w[3]; log3(R, w);
  double ang = std::sqrt(w[0] * w[0] + w[1] * w[1] + w[2] * w[2]);
  double I[9] = {1, 0, 0, 0, 1, 0, 0, 0, 1};
  if (ang < 1e-9) { std::memcpy(J, I, sizeof I); return; }
  double a[3] = {w[0] / ang, w[1] / ang, w[2] / ang}, H[9];
  hat3(a, H);
  double ctt = ang / 2 / std::tan(ang / 2);
  for (int i = 0; i < 3; i++) for (int j = 0; j < 3; j++) J[3 * i + j] = ctt * I[3 * i + j] + (1 - ctt) * a[i] * a[j] + ang / 2 * H[3 * i + j];
}

// generic small dense (row-major) helpers
inline bool inv_nxn(const double* A, double* Ainv, int n) {  // Gauss–Jordan, partial pivoting (n <= 16)
  double M[16 * 32];
  for (int i = 0; i < n; i++) for (int j = 0; j < n; j++) { M[size_t(i) * 2 * n + j] = A[i * n + j]; M[size_t(i) * 2 * n + n + j] = (i == j); }
  for (int c = 0; c < n; c++) {
    int p = c; double big = std::fabs(M[size_t(c) * 2 * n + c]);
    for (int r = c + 1; r < n; r++) { double v = std::fabs(M[size_t(r) * 2 * n + c]); if (v > big) { big = v; p = r; } }
    if (big == 0) return false;
    if (p != c) for (int j = 0; j < 2 * n; j++) std::swap(M[size_t(c) * 2 * n + j], M[size_t(p) * 2 * n + j]);
    double d = M[size_t(c) * 2 * n + c];
    for (int j = 0; j < 2 * n; j++) M[size_t(c) * 2 * n + j] /= d;
    for (int r = 0; r < n; r++) if (r != c) { double f = M[size_t(r) * 2 * n + c]; if (f != 0) for (int j = 0; j < 2 * n; j++) M[size_t(r) * 2 * n + j] -= f * M[size_t(c) * 2 * n + j]; }
  }
  for (int i = 0; i < n; i++) for (int j = 0; j < n; j++) Ainv[i * n + j] = M[size_t(i) * 2 * n + n + j];
  return true;
}

// ------------------------------------------------------------------ scene: 3-plane corner room, SURVEY.md §8(d)
// planes x=off, y=off, z=off over [off, L+off]^2; true pose_i = (Exp([0,0,0.01 i]), (L/2+0.05 i, L/2+0.03 i, 1.5)).
struct Scene {
  double L = 20.0, off = 0.37, sigma = 0.01, max_range = 0.0;  // max_range 0 = unlimited
  uint64_t seed = 0x5EED0000ull;
};
inline void true_pose(const Scene& sc, int i, double* pose12) {  // R row-major (9) + p (3)
  double w[3] = {0, 0, 0.01 * i};
  exp3(w, pose12);
  pose12[9] = sc.L / 2 + 0.05 * i; pose12[10] = sc.L / 2 + 0.03 * i; pose12[11] = 1.5;
}
// pose of frame i at an arbitrary place (used by the GBA lawn-mower scenes): overrides the centre
inline void perturb_pose(const double* pose12, uint64_t seed, double rot_sigma, double pos_sigma, double* out12) {
  SplitMix64 g(seed);
  double w[3] = {rot_sigma * g.gauss(), rot_sigma * g.gauss(), rot_sigma * g.gauss()}, E[9];
  exp3(w, E);
  mat3_mul(pose12, E, out12);
  for (int k = 0; k < 3; k++) out12[9 + k] = pose12[9 + k] + pos_sigma * g.gauss();
}
// n body-frame points of frame `frame` (fp64 xyz). Points are uniform over the three plane areas with N(0,sigma) along the normal,
// expressed in the body frame of pose12_true.  If sc.max_range>0 points farther than that from the sensor are re-drawn.
inline void gen_scan(const Scene& sc, int frame, int64_t n, const double* pose12_true, double* xyz_body) {
  SplitMix64 g(sc.seed + 1000003ull * uint64_t(frame + 1));
  double Rt[9]; mat3_t(pose12_true, Rt);
  const double* t = pose12_true + 9;
  for (int64_t k = 0; k < n; k++) {
    double pw[3];
    for (int tries = 0; tries < 64; tries++) {
      int plane = int(g.next() % 3);
      double a = sc.off + sc.L * g.uni(), b = sc.off + sc.L * g.uni(), c = sc.off + sc.sigma * g.gauss();
      if (sc.max_range > 0) {  // sample around the sensor instead of over the whole floor
        double ca = t[(plane + 1) % 3], cb = t[(plane + 2) % 3];
        a = ca + sc.max_range * (2 * g.uni() - 1); b = cb + sc.max_range * (2 * g.uni() - 1);
        if (a < sc.off || a > sc.off + sc.L || b < sc.off || b > sc.off + sc.L) continue;
      }
      pw[plane] = c; pw[(plane + 1) % 3] = a; pw[(plane + 2) % 3] = b;
      if (sc.max_range > 0) {
        double d2 = 0; for (int j = 0; j < 3; j++) d2 += (pw[j] - t[j]) * (pw[j] - t[j]);
        if (d2 > sc.max_range * sc.max_range) continue;
      }
      break;
    }
    double d[3] = {pw[0] - t[0], pw[1] - t[1], pw[2] - t[2]};
    mat3_vec(Rt, d, xyz_body + 3 * k);
  }
}

// ------------------------------------------------------------------ large "city grid" scene for the global-BA workloads (C4-like)
// floor z = off, walls x = off + G*i and y = off + G*j (height 0..4 m above the floor).  A keyframe at (cx, cy) sees points within
// `range` on the floor and on the two nearest wall lines of each direction, so every keyframe is constrained in all six dofs and
// co-visibility stays local (k ~ 5-15 keyframes per voxel).
inline void lawnmower_pose(int i, int per_row, double step, double row_gap, double off, double* pose12) {
  const int row = i / per_row, col = i % per_row;
  const double x = off + 3.0 + step * ((row & 1) ? (per_row - 1 - col) : col), y = off + 3.0 + row_gap * row;
  double w[3] = {0, 0, 0.3 * std::sin(0.37 * i)};
  exp3(w, pose12);
  pose12[9] = x; pose12[10] = y; pose12[11] = off + 1.5;
}
inline void gen_scan_city(uint64_t seed, int frame, int64_t n, const double* pose12_true, double G, double range, double off, double sigma, double* xyz_body) {
  SplitMix64 g(seed + 7000003ull * uint64_t(frame + 1));
  double Rt[9]; mat3_t(pose12_true, Rt);
  const double* t = pose12_true + 9;
  for (int64_t k = 0; k < n; k++) {
    double pw[3];
    const int kind = int(g.next() % 3);
    if (kind == 0) {  // floor
      pw[0] = t[0] + range * (2 * g.uni() - 1); pw[1] = t[1] + range * (2 * g.uni() - 1); pw[2] = off + sigma * g.gauss();
    } else {          // wall line x = const (kind 1) or y = const (kind 2): one of the two nearest lines
      const int a = kind - 1, b = 1 - a;
      const double cell = std::floor((t[a] - off) / G);
      const double line = off + G * (cell + double(g.next() & 1));
      pw[a] = line + sigma * g.gauss();
      pw[b] = t[b] + range * (2 * g.uni() - 1);
      pw[2] = off + 4.0 * g.uni();
    }
    double d[3] = {pw[0] - t[0], pw[1] - t[1], pw[2] - t[2]};
    mat3_vec(Rt, d, xyz_body + 3 * k);
  }
}

// ------------------------------------------------------------------ IMU factor stand-in (mirrors preintegration.hpp:11-331)
struct ImuPre {
  double R_delta[9], p_delta[3], v_delta[3], bg[3], ba[3];
  double R_bg[9], p_bg[9], p_ba[9], v_bg[9], v_ba[9];
  double dtime = 0, dbg[3], dba[3], dbg_buf[3], dba_buf[3];
  double cov[225];
  double noiseMeas[6], noiseWalk[6];  // diagonals (voxelslam.cpp:828-833)
  ImuPre() {
    double I[9] = {1, 0, 0, 0, 1, 0, 0, 0, 1};
    std::memcpy(R_delta, I, sizeof I);
    for (double* a : {p_delta, v_delta, bg, ba, dbg, dba, dbg_buf, dba_buf}) a[0] = a[1] = a[2] = 0;
    for (double* m : {R_bg, p_bg, p_ba, v_bg, v_ba}) std::memset(m, 0, 72);
    std::memset(cov, 0, sizeof cov);
    for (int i = 0; i < 3; i++) { noiseMeas[i] = 0.01; noiseMeas[3 + i] = 1.0; noiseWalk[i] = 1e-4; noiseWalk[3 + i] = 1e-4; }  // config/avia.yaml:39-42
  }
  void add_imu(const double* gyr_in, const double* acc_in, double dt) {  // preintegration.hpp:75-135 (gyr/acc already bias-corrected)
    double gyr[3] = {gyr_in[0] - bg[0], gyr_in[1] - bg[1], gyr_in[2] - bg[2]}, acc[3] = {acc_in[0] - ba[0], acc_in[1] - ba[1], acc_in[2] - ba[2]};
    dtime += dt;
    double wdt[3] = {gyr[0] * dt, gyr[1] * dt, gyr[2] * dt}, R_inc[9], R_jr[9], R_incT[9];
    exp3(wdt, R_inc); jr3(wdt, R_jr); mat3_t(R_inc, R_incT);
    double R_dt[9], R_dt2_2[9], acc_skew[9];
    for (int i = 0; i < 9; i++) { R_dt[i] = dt * R_delta[i]; R_dt2_2[i] = 0.5 * dt * dt * R_delta[i]; }
    hat3(acc, acc_skew);
    double T1[9], T2[9];
    mat3_mul(R_dt2_2, acc_skew, T1); mat3_mul(T1, R_bg, T2);
    for (int i = 0; i < 9; i++) { p_ba[i] = p_ba[i] + v_ba[i] * dt - R_dt2_2[i]; p_bg[i] = p_bg[i] + v_bg[i] * dt - T2[i]; }
    mat3_mul(R_dt, acc_skew, T1); mat3_mul(T1, R_bg, T2);
    for (int i = 0; i < 9; i++) { v_ba[i] = v_ba[i] - R_dt[i]; v_bg[i] = v_bg[i] - T2[i]; }
    mat3_mul(R_incT, R_bg, T1);
    for (int i = 0; i < 9; i++) R_bg[i] = T1[i] - R_jr[i] * dt;
    // covariance: A (9x9), B (9x6)
    double A[81], B[54];
    std::memset(A, 0, sizeof A); std::memset(B, 0, sizeof B);
    for (int i = 0; i < 9; i++) A[i * 9 + i] = 1;
    double M1[9], M2[9];
    mat3_mul(R_dt2_2, acc_skew, M1); mat3_mul(R_dt, acc_skew, M2);
    for (int r = 0; r < 3; r++) for (int c = 0; c < 3; c++) {
      A[r * 9 + c] = R_incT[3 * r + c];
      A[(3 + r) * 9 + c] = -M1[3 * r + c];
      A[(3 + r) * 9 + 6 + c] = (r == c) * dt;
      A[(6 + r) * 9 + c] = -M2[3 * r + c];
      B[r * 6 + c] = R_jr[3 * r + c] * dt;
      B[(3 + r) * 6 + 3 + c] = R_dt2_2[3 * r + c];
      B[(6 + r) * 6 + 3 + c] = R_dt[3 * r + c];
    }
    double C9[81], AC[81];
    for (int r = 0; r < 9; r++) for (int c = 0; c < 9; c++) C9[r * 9 + c] = cov[r * 15 + c];
    for (int r = 0; r < 9; r++) for (int c = 0; c < 9; c++) { double s = 0; for (int k = 0; k < 9; k++) s += A[r * 9 + k] * C9[k * 9 + c]; AC[r * 9 + c] = s; }
    for (int r = 0; r < 9; r++) for (int c = 0; c < 9; c++) {
      double s = 0;
      for (int k = 0; k < 9; k++) s += AC[r * 9 + k] * A[c * 9 + k];
      for (int k = 0; k < 6; k++) s += B[r * 6 + k] * noiseMeas[k] * B[c * 6 + k];
      cov[r * 15 + c] = s;
    }
    for (int k = 0; k < 6; k++) cov[(9 + k) * 15 + 9 + k] += noiseWalk[k] * dt;
    double Ra[3];
    mat3_vec(R_dt2_2, acc, Ra);
    for (int i = 0; i < 3; i++) p_delta[i] += v_delta[i] * dt + Ra[i];
    mat3_vec(R_dt, acc, Ra);
    for (int i = 0; i < 3; i++) v_delta[i] += Ra[i];
    mat3_mul(R_delta, R_inc, T1); std::memcpy(R_delta, T1, sizeof T1);
  }
  // preintegration.hpp:137-212 / 214-294.  st = 24 doubles (R9 p v bg ba g).  jtj (bs x bs, COLUMN-major), gg (bs); bs = 30 or 33.
  double give_evaluate(const double* st1, const double* st2, double* jtj, double* gg, bool jac_enable, bool with_g) const {
    const double *R1 = st1, *p1 = st1 + 9, *v1 = st1 + 12, *bg1 = st1 + 15, *ba1 = st1 + 18, *g1 = st1 + 21;
    const double *R2 = st2, *p2 = st2 + 9, *v2 = st2 + 12, *bg2 = st2 + 15, *ba2 = st2 + 18;
    double t3[3], E[9], R_correct[9];
    mat3_vec(R_bg, dbg, t3); exp3(t3, E); mat3_mul(R_delta, E, R_correct);
    double t_correct[3], v_correct[3], a3[3], b3[3];
    mat3_vec(p_bg, dbg, a3); mat3_vec(p_ba, dba, b3);
    for (int i = 0; i < 3; i++) t_correct[i] = p_delta[i] + a3[i] + b3[i];
    mat3_vec(v_bg, dbg, a3); mat3_vec(v_ba, dba, b3);
    for (int i = 0; i < 3; i++) v_correct[i] = v_delta[i] + a3[i] + b3[i];
    double RcT[9], R1T[9], T1[9], res_r[9];
    mat3_t(R_correct, RcT); mat3_t(R1, R1T); mat3_mul(RcT, R1T, T1); mat3_mul(T1, R2, res_r);
    double dv[3], dp[3], exp_v[3], exp_t[3], rr[15];
    for (int i = 0; i < 3; i++) { dv[i] = v2[i] - v1[i] - dtime * g1[i]; dp[i] = p2[i] - p1[i] - v1[i] * dtime - 0.5 * dtime * dtime * g1[i]; }
    mat3_vec(R1T, dv, exp_v); mat3_vec(R1T, dp, exp_t);
    log3(res_r, rr);
    for (int i = 0; i < 3; i++) { rr[3 + i] = exp_t[i] - t_correct[i]; rr[6 + i] = exp_v[i] - v_correct[i]; rr[9 + i] = bg2[i] - bg1[i]; rr[12 + i] = ba2[i] - ba1[i]; }
    double cov_inv[225];
    inv_nxn(cov, cov_inv, 15);
    if (jac_enable) {
      const int bs = with_g ? 33 : 30;
      double joc[15 * 33];  // row-major 15 x bs : [joca | jocb | jocg]
      for (int i = 0; i < 15 * bs; i++) joc[i] = 0.0;
      auto J = [&](int r, int c) -> double& { return joc[size_t(r) * bs + c]; };
      double JR_inv[9], R2T[9], M[9], M2[9], resT[9], Jr[9];
      jr3_inv(res_r, JR_inv); mat3_t(R2, R2T); mat3_t(res_r, resT);
      mat3_mul(JR_inv, R2T, M); mat3_mul(M, R1, M2);
      for (int r = 0; r < 3; r++) for (int c = 0; c < 3; c++) { J(r, c) = -M2[3 * r + c]; J(r, 15 + c) = JR_inv[3 * r + c]; }
      jr3(t3, Jr);
      mat3_mul(JR_inv, resT, M); mat3_mul(M, Jr, M2); mat3_mul(M2, R_bg, M);
      for (int r = 0; r < 3; r++) for (int c = 0; c < 3; c++) J(r, 9 + c) = -M[3 * r + c];
      double Ht[9], Hv[9];
      hat3(exp_t, Ht); hat3(exp_v, Hv);
      for (int r = 0; r < 3; r++) for (int c = 0; c < 3; c++) {
        J(3 + r, c) = Ht[3 * r + c]; J(3 + r, 3 + c) = -R1T[3 * r + c]; J(3 + r, 6 + c) = -R1T[3 * r + c] * dtime;
        J(3 + r, 9 + c) = -p_bg[3 * r + c]; J(3 + r, 12 + c) = -p_ba[3 * r + c]; J(3 + r, 15 + 3 + c) = R1T[3 * r + c];
        J(6 + r, c) = Hv[3 * r + c]; J(6 + r, 6 + c) = -R1T[3 * r + c]; J(6 + r, 9 + c) = -v_bg[3 * r + c]; J(6 + r, 12 + c) = -v_ba[3 * r + c];
        J(6 + r, 15 + 6 + c) = R1T[3 * r + c];
        J(9 + r, 9 + c) = -(r == c); J(12 + r, 12 + c) = -(r == c); J(9 + r, 15 + 9 + c) = (r == c); J(12 + r, 15 + 12 + c) = (r == c);
        if (with_g) { J(3 + r, 30 + c) = R1T[3 * r + c] * (-0.5 * dtime * dtime); J(6 + r, 30 + c) = R1T[3 * r + c] * (-dtime); }
      }
      double CJ[15 * 33];
      for (int r = 0; r < 15; r++) for (int c = 0; c < bs; c++) { double s = 0; for (int k = 0; k < 15; k++) s += cov_inv[r * 15 + k] * J(k, c); CJ[size_t(r) * bs + c] = s; }
      for (int a = 0; a < bs; a++) for (int b = 0; b < bs; b++) { double s = 0; for (int k = 0; k < 15; k++) s += J(k, a) * CJ[size_t(k) * bs + b]; jtj[size_t(b) * bs + a] = s; }
      double cr[15];
      for (int k = 0; k < 15; k++) { double t = 0; for (int m = 0; m < 15; m++) t += cov_inv[k * 15 + m] * rr[m]; cr[k] = t; }
      for (int a = 0; a < bs; a++) { double s = 0; for (int k = 0; k < 15; k++) s += J(k, a) * cr[k]; gg[a] = s; }
    }
    double cost = 0;
    for (int r = 0; r < 15; r++) { double s = 0; for (int k = 0; k < 15; k++) s += cov_inv[r * 15 + k] * rr[k]; cost += rr[r] * s; }
    return cost;
  }
  void update_state(const double* dxi15) {  // preintegration.hpp:296-303
    for (int i = 0; i < 3; i++) { dbg_buf[i] = dbg[i]; dba_buf[i] = dba[i]; dbg[i] += dxi15[9 + i]; dba[i] += dxi15[12 + i]; }
  }
  void rollback() { for (int i = 0; i < 3; i++) { dbg[i] = dbg_buf[i]; dba[i] = dba_buf[i]; } }  // voxel_map.hpp:639-643
};

// W-1 factors of a window + the three callbacks of vxs_imu_hooks
struct ImuWindow {
  std::vector<ImuPre> f;
  // constant-rate synthetic IMU between consecutive true poses: gyr = Log(Ri^T Rj)/T, acc = Ri^T(a - g) with zero world acceleration
  void build(const double* poses12_true, int W, double T, int samples, double gyr_noise, double acc_noise, uint64_t seed) {
    f.assign(W - 1, ImuPre());
    SplitMix64 g(seed);
    for (int i = 0; i + 1 < W; i++) {
      const double *Ri = poses12_true + 12 * i, *Rj = poses12_true + 12 * (i + 1);
      double RiT[9], dR[9], w[3];
      mat3_t(Ri, RiT); mat3_mul(RiT, Rj, dR); log3(dR, w);
      double dt = T / samples, Rcur[9];
      std::memcpy(Rcur, Ri, 72);
      for (int s = 0; s < samples; s++) {
        double gw[3] = {0, 0, 9.8}, acc[3], gyr[3];
        mat3_tvec(Rcur, gw, acc);
        for (int k = 0; k < 3; k++) { gyr[k] = w[k] / T + gyr_noise * g.gauss(); acc[k] += acc_noise * g.gauss(); }
        f[i].add_imu(gyr, acc, dt);
        double wd[3] = {w[0] / T * dt, w[1] / T * dt, w[2] / T * dt}, E[9], Rn[9];
        exp3(wd, E); mat3_mul(Rcur, E, Rn); std::memcpy(Rcur, Rn, 72);
      }
    }
  }
  int eval(const double* states, int W, int with_g, int want_jac, double* blocks, double* gvec, double* cost) const {
    const int bs = with_g ? 33 : 30;
    double c = 0;
    for (int i = 0; i + 1 < W; i++)
      c += f[i].give_evaluate(states + 24 * i, states + 24 * (i + 1), want_jac ? blocks + size_t(i) * bs * bs : nullptr, want_jac ? gvec + size_t(i) * bs : nullptr,
                              want_jac != 0, with_g != 0);
    *cost = c;
    return 0;
  }
  int update(const double* dxi, int W) { for (int i = 0; i + 1 < W; i++) f[i].update_state(dxi + 15 * i); return 0; }
  int rollback() { for (auto& x : f) x.rollback(); return 0; }
};

}  // namespace vxh
