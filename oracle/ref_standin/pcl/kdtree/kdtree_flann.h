// STAND-IN (test infrastructure): brute-force nearest neighbour with the KdTreeFLANN call shape (only icp_normal in loop_refine.hpp uses it — off the hot path).
#ifndef VXREF_PCL_KDTREE
#define VXREF_PCL_KDTREE
#include <algorithm>
#include <vector>
#include "../point_cloud.h"
#include "../point_types.h"
namespace pcl {
template <class PointT> class KdTreeFLANN {
  typename PointCloud<PointT>::ConstPtr cloud;
 public:
  void setInputCloud(const typename PointCloud<PointT>::ConstPtr& c) { cloud = c; }
  int nearestKSearch(const PointT& p, int k, std::vector<int>& idx, std::vector<float>& d2) const {
    if (!cloud || cloud->empty()) return 0;
    std::vector<std::pair<float, int>> all(cloud->size());
    for (size_t i = 0; i < cloud->size(); i++) { const PointT& q = (*cloud)[i]; const float dx = q.x - p.x, dy = q.y - p.y, dz = q.z - p.z; all[i] = {dx * dx + dy * dy + dz * dz, int(i)}; }
    k = std::min<int>(k, int(all.size()));
    std::partial_sort(all.begin(), all.begin() + k, all.end());
    idx.resize(size_t(k)); d2.resize(size_t(k));
    for (int i = 0; i < k; i++) { idx[size_t(i)] = all[size_t(i)].second; d2[size_t(i)] = all[size_t(i)].first; }
    return k;
  }
};
}  // namespace pcl
#endif
