// TEST INFRASTRUCTURE ONLY (oracle/): minimal stand-in for <pcl/point_cloud.h>
// (just the members kd_tree_two.h touches: ::Ptr and ->points).
#pragma once
#include <memory>
#include <vector>
namespace pcl {
template <typename PointT> struct PointCloud {
    using Ptr = std::shared_ptr<PointCloud<PointT>>;
    std::vector<PointT> points;
};
} // namespace pcl
