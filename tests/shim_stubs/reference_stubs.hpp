// Minimal stand-ins for the reference-side types the typed layer of voxel_ba_shim.hpp touches (Eigen / PCL / tools.hpp / preintegration.hpp /
// voxel_map.hpp are not installed here).  Only names, members and call signatures are mirrored — reference file:line beside each — so that
// `g++ -fsyntax-only -DVXS_SHIM_WITH_REFERENCE_TYPES` type-checks the shim against the call surface it claims to serve.  Test infrastructure only.
#pragma once
#include <cstddef>
#include <cstring>
#include <deque>
#include <memory>
#include <vector>
#define DIM 15                                            // tools.hpp:16
namespace Eigen {
struct Vector3d { double d[3] = {0, 0, 0}; Vector3d() = default; Vector3d(double a, double b, double c) { d[0] = a; d[1] = b; d[2] = c; } double& operator[](int i) { return d[i]; } const double& operator[](int i) const { return d[i]; } void setZero() { d[0] = d[1] = d[2] = 0; } };
struct Matrix3d { double d[9] = {0}; double& operator()(int r, int c) { return d[3 * c + r]; } const double& operator()(int r, int c) const { return d[3 * c + r]; } };
struct MatrixXd {
  std::vector<double> a; int r_ = 0, c_ = 0;
  MatrixXd() = default; MatrixXd(int r, int c) : a(size_t(r) * c), r_(r), c_(c) {}
  void resize(int r, int c) { a.assign(size_t(r) * c, 0.0); r_ = r; c_ = c; }
  void setZero() { std::fill(a.begin(), a.end(), 0.0); }
  double* data() { return a.data(); } const double* data() const { return a.data(); }
};
struct VectorXd { std::vector<double> a; VectorXd() = default; explicit VectorXd(int n) : a(n) {} void setZero() { std::fill(a.begin(), a.end(), 0.0); } double* data() { return a.data(); } };
template <class T, int R, int C> struct Matrix { T d[R * C]; T& operator()(int r, int c) { return d[R * c + r]; } const T& operator()(int r, int c) const { return d[R * c + r]; } T& operator[](int i) { return d[i]; } };
template <class M> struct Map { const double* p; explicit Map(const double* q) : p(q) {} operator Matrix<double, DIM, 1>() const { Matrix<double, DIM, 1> m; std::memcpy(m.d, p, sizeof m.d); return m; } };
}  // namespace Eigen
struct IMUST { Eigen::Matrix3d R; Eigen::Vector3d p, v, bg, ba, g; Eigen::Matrix<double, DIM, DIM> cov; };   // tools.hpp:135-199
struct PointCluster { Eigen::Matrix3d P; Eigen::Vector3d v; int N = 0; };                              // tools.hpp:304-365
struct IMU_PRE {                                                                                          // preintegration.hpp:20-300
  Eigen::Vector3d dbg, dba, dbg_buf, dba_buf;                                                            // :25-26
  double give_evaluate(IMUST&, IMUST&, Eigen::MatrixXd&, Eigen::VectorXd&, bool) { return 0; }           // :137
  double give_evaluate_g(IMUST&, IMUST&, Eigen::MatrixXd&, Eigen::VectorXd&, bool) { return 0; }         // :214
  void update_state(const Eigen::Matrix<double, DIM, 1>&) {}                                             // :296
};
namespace pcl {
struct PointXYZINormal { float x, y, z, pad0, normal_x, normal_y, normal_z, pad1, intensity, curvature, pad2, pad3; };   // 48 bytes, x,y,z first
template <class T> struct PointCloud {
  using Ptr = std::shared_ptr<PointCloud<T>>;
  std::vector<T> points;
  size_t size() const { return points.size(); } void reserve(size_t n) { points.reserve(n); } void push_back(const T& t) { points.push_back(t); }
  void clear() { points.clear(); } void swap(PointCloud& o) { points.swap(o.points); }
};
}  // namespace pcl
typedef pcl::PointXYZINormal PointType;                                                                  // tools.hpp:19
struct pointVar { Eigen::Vector3d pnt; Eigen::Matrix3d var; };                                           // voxel_map.hpp:14-19
using PVec = std::vector<pointVar>;
using PVecPtr = std::shared_ptr<std::vector<pointVar>>;                                                 // voxel_map.hpp:22
struct SlideWindow { std::vector<PVec> points; std::vector<PointCluster> pcrs_local; };                   // voxel_map.hpp:896-930
namespace Eigen { template <class T> using aligned_allocator = std::allocator<T>; }
#define PLV(a) std::vector<Eigen::Vector3d, Eigen::aligned_allocator<Eigen::Vector3d>>                 // tools.hpp:13 (only PLV(3) is used)                                                                      // voxel_map.hpp:21
struct Keyframe { IMUST x0; pcl::PointCloud<PointType>::Ptr plptr; int exist, id, mp; float jour; };     // voxel_map.hpp:867-874
