// C entry points of the synthetic-workload harness (libvxs_harness.so) — scene generation and the IMU-factor stand-in
// that bench.py / tests hand to vxs_li_ba through vxs_imu_hooks.  CPU only; not the hot path, not the oracle.
#include "../../include/vxs.h"
#include "synth.hpp"

using namespace vxh;

extern "C" {

void vxh_true_pose(double L, int i, double* pose12) { Scene sc; sc.L = L; true_pose(sc, i, pose12); }
void vxh_perturb_pose(const double* pose12, uint64_t seed, double rot_sigma, double pos_sigma, double* out12) { perturb_pose(pose12, seed, rot_sigma, pos_sigma, out12); }
// n body-frame fp64 points of frame `frame`
void vxh_gen_scan(double L, double off, double sigma, double max_range, uint64_t seed, int frame, int64_t n, const double* pose12_true, double* xyz_body) {
  Scene sc; sc.L = L; sc.off = off; sc.sigma = sigma; sc.max_range = max_range; sc.seed = seed;
  gen_scan(sc, frame, n, pose12_true, xyz_body);
}
void vxh_gen_scan_f32(double L, double off, double sigma, double max_range, uint64_t seed, int frame, int64_t n, const double* pose12_true, float* xyz_body) {
  std::vector<double> tmp(size_t(n) * 3);
  vxh_gen_scan(L, off, sigma, max_range, seed, frame, n, pose12_true, tmp.data());
  for (size_t i = 0; i < tmp.size(); i++) xyz_body[i] = float(tmp[i]);
}

void vxh_lawnmower_pose(int i, int per_row, double step, double row_gap, double off, double* pose12) { lawnmower_pose(i, per_row, step, row_gap, off, pose12); }
void vxh_gen_scan_city_f32(uint64_t seed, int frame, int64_t n, const double* pose12_true, double G, double range, double off, double sigma, float* xyz_body) {
  std::vector<double> tmp(size_t(n) * 3);
  gen_scan_city(seed, frame, n, pose12_true, G, range, off, sigma, tmp.data());
  for (size_t i = 0; i < tmp.size(); i++) xyz_body[i] = float(tmp[i]);
}

struct ImuHandle { ImuWindow win, initial; };

void* vxh_imu_create(const double* poses12_true, int W, double T, int samples, double gyr_noise, double acc_noise, uint64_t seed) {
  ImuHandle* h = new ImuHandle();
  h->win.build(poses12_true, W, T, samples, gyr_noise, acc_noise, seed);
  h->initial = h->win;
  return h;
}
void vxh_imu_destroy(void* h) { delete static_cast<ImuHandle*>(h); }
void vxh_imu_reset(void* h) { ImuHandle* p = static_cast<ImuHandle*>(h); p->win = p->initial; }  // forget the bias increments of a previous solve

static int cb_eval(void* user, const double* states24, int W, int with_gravity, int want_jac, double* blocks, double* gvec, double* cost) {
  return static_cast<ImuHandle*>(user)->win.eval(states24, W, with_gravity, want_jac, blocks, gvec, cost);
}
static int cb_update(void* user, const double* dxi, int W) { return static_cast<ImuHandle*>(user)->win.update(dxi, W); }
static int cb_rollback(void* user) { return static_cast<ImuHandle*>(user)->win.rollback(); }

void vxh_imu_hooks(void* h, vxs_imu_hooks* out) { out->user = h; out->eval = cb_eval; out->update = cb_update; out->rollback = cb_rollback; }
// direct access for tests
int vxh_imu_eval(void* h, const double* states24, int W, int with_gravity, int want_jac, double* blocks, double* gvec, double* cost) {
  return cb_eval(h, states24, W, with_gravity, want_jac, blocks, gvec, cost);
}

}  // extern "C"
