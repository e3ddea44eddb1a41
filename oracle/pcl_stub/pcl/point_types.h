// TEST INFRASTRUCTURE ONLY (oracle/): minimal stand-in for <pcl/point_types.h>.
// PCL is not installed in the build image; the reference wrapper
// roswrapper/ros/src/avoid_mpc/include/kd_tree_two.h only needs the
// pcl::PointXYZ record (three floats in a 16-byte, 16-byte-aligned slot).
#pragma once
namespace pcl {
struct alignas(16) PointXYZ {
    float x, y, z, _pad;
    PointXYZ() : x(0), y(0), z(0), _pad(1.0f) {}
    PointXYZ(float _x, float _y, float _z) : x(_x), y(_y), z(_z), _pad(1.0f) {}
};
static_assert(sizeof(PointXYZ) == 16, "PointXYZ must keep PCL's 16-byte stride");
} // namespace pcl
