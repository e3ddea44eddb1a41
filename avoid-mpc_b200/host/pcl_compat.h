// pcl::PointXYZ / pcl::PointCloud for builds without PCL.  In the reference's catkin
// workspace the real headers are used (the record layout is the same: float x,y,z in a
// 16-byte, 16-byte-aligned slot; include/kd_tree_two.h:5-7,13).
#pragma once
#if __has_include(<pcl/point_types.h>) && __has_include(<pcl/point_cloud.h>)
#include <pcl/point_cloud.h>
#include <pcl/point_types.h>
#else
#include <memory>
#include <vector>
namespace pcl {
struct alignas(16) PointXYZ {
    float x, y, z, _pad;
    PointXYZ() : x(0), y(0), z(0), _pad(1.0f) {}
    PointXYZ(float _x, float _y, float _z) : x(_x), y(_y), z(_z), _pad(1.0f) {}
};
template <typename PointT> struct PointCloud {
    using Ptr = std::shared_ptr<PointCloud<PointT>>;
    std::vector<PointT> points;
    void emplace_back(float x, float y, float z) { points.emplace_back(x, y, z); }
    size_t size() const { return points.size(); }
};
} // namespace pcl
#endif
static_assert(sizeof(pcl::PointXYZ) == 16, "pcl::PointXYZ must be a 16-byte record");
