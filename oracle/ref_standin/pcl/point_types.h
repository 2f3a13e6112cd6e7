// STAND-IN (test infrastructure, see ../Eigen/Core): the two PCL point types the reference's hot-path headers touch, with PCL's field layout
// (pcl::PointXYZINormal = 48 bytes: float data[4] {x,y,z,pad}, float data_n[4] {normal_x,y,z,pad}, intensity, curvature, 2 pad floats).
#ifndef VXREF_PCL_POINT_TYPES
#define VXREF_PCL_POINT_TYPES
namespace pcl {
struct PointXYZ {
  union { float data[4]; struct { float x, y, z; }; };
  PointXYZ() : x(0), y(0), z(0) { data[3] = 1.0f; }
  PointXYZ(float x_, float y_, float z_) : x(x_), y(y_), z(z_) { data[3] = 1.0f; }
};
struct PointXYZINormal {
  union { float data[4]; struct { float x, y, z; }; };
  union { float data_n[4]; float normal[3]; struct { float normal_x, normal_y, normal_z; }; };
  union { struct { float intensity, curvature; }; float data_c[4]; };
  PointXYZINormal() { data[0] = data[1] = data[2] = 0; data[3] = 1.0f; data_n[0] = data_n[1] = data_n[2] = data_n[3] = 0; data_c[0] = data_c[1] = data_c[2] = data_c[3] = 0; }
};
}  // namespace pcl
#endif
