// STAND-IN (test infrastructure): just enough GTSAM surface for loop_refine.hpp to compile (ScanPose::set_state, add_edge); the pose graph
// itself is out of scope and never executed by the oracle library.
#ifndef VXREF_GTSAM
#define VXREF_GTSAM
#include <Eigen/Core>
#include <memory>
#include <vector>
namespace gtsam {
typedef unsigned long long Key;
struct Point3 : Eigen::Vector3d { Point3() {} template <class O> Point3(const Eigen::DenseBase<O>& v) : Eigen::Vector3d(v) {} Point3(double x, double y, double z) : Eigen::Vector3d(x, y, z) {} };
struct Rot3 { Eigen::Matrix3d R; Rot3() { R.setIdentity(); } template <class O> Rot3(const Eigen::DenseBase<O>& m) : R(m) {} const Eigen::Matrix3d& matrix() const { return R; } };
struct Pose3 { Rot3 r; Point3 t; Pose3() {} Pose3(const Rot3& r_, const Point3& t_) : r(r_), t(t_) {} const Rot3& rotation() const { return r; } const Point3& translation() const { return t; } };
struct NonlinearFactor { typedef std::shared_ptr<NonlinearFactor> shared_ptr; virtual ~NonlinearFactor() {} };
namespace noiseModel { struct Diagonal { typedef std::shared_ptr<Diagonal> shared_ptr; Eigen::VectorXd v; static shared_ptr Variances(const Eigen::VectorXd& x) { auto p = std::make_shared<Diagonal>(); p->v = x; return p; } }; }
template <class T> struct BetweenFactor : NonlinearFactor { Key k1, k2; T z; noiseModel::Diagonal::shared_ptr n; BetweenFactor(Key a, Key b, const T& z_, noiseModel::Diagonal::shared_ptr n_) : k1(a), k2(b), z(z_), n(n_) {} };
template <class T> struct PriorFactor : NonlinearFactor { Key k; T z; noiseModel::Diagonal::shared_ptr n; PriorFactor(Key a, const T& z_, noiseModel::Diagonal::shared_ptr n_) : k(a), z(z_), n(n_) {} };
struct NonlinearFactorGraph { std::vector<NonlinearFactor::shared_ptr> f; void push_back(const NonlinearFactor::shared_ptr& x) { f.push_back(x); } void add(const NonlinearFactor::shared_ptr& x) { f.push_back(x); } size_t size() const { return f.size(); } };
struct Values {};
struct ISAM2 {};
}  // namespace gtsam
#endif
