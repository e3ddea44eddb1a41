// Query side of the reference's perception map
// (roswrapper/ros/src/avoid_mpc/include/FrameKDMap.h:60-67, src/FrameKDMap.cpp:254-275,
// 322-427) on libampc: QueryNearest / GetNearestDistance over the CURRENT frame's Obstacle
// and Edge clouds.  Scope notes: the reference builds the two clouds from a
// depth image inside AddVertex (FrameKDMap.cpp:34-52,90-214) and also searches up to 100
// key-frames; here the clouds are handed in ready-made (AddClouds) and only the current
// frame is searched -- what the reference's fast path does (FrameKDMap.cpp:339-345).
#ifndef FRAME_KD_MAP_H
#define FRAME_KD_MAP_H
#include "../../include/ampc.h"
#include "eigen_compat.h"
#include "pcl_compat.h"

#include <memory>
#include <vector>

class FrameKDMap {
public:
    explicit FrameKDMap(int maxPoints = 65536, int maxEdgePoints = 16384);
    using CloudPtr = pcl::PointCloud<pcl::PointXYZ>::Ptr;
    // replaces AddVertex's two InitializeNew calls + swap (FrameKDMap.cpp:44-51)
    void AddClouds(const CloudPtr &cloud, const CloudPtr &edgeCloud);
    void QueryNearest(const Eigen::Vector3d &point, int nearestPointCount,
                      std::vector<Eigen::Vector3d> &out, std::vector<double> &distances,
                      bool queryEdge = false);
    double GetNearestDistance(const Eigen::Vector3d &point);
    // batched form used by the tick loop: Q query sites at once (one kernel launch)
    void QueryNearestBatch(const std::vector<Eigen::Vector3d> &points, int nearestPointCount,
                           std::vector<std::vector<Eigen::Vector3d>> &out,
                           std::vector<std::vector<double>> &distances, bool queryEdge = false);
    int PointCount(bool edge) const { return mCount[edge ? 1 : 0]; }
    ampc_handle *Handle() { return mHandle.get(); }

private:
    std::shared_ptr<ampc_handle> mHandle;
    int mCount[2] = {0, 0};
};
#endif
