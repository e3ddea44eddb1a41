// The reference's perception map, query side and key-frame bookkeeping
// (roswrapper/ros/src/avoid_mpc/include/FrameKDMap.h:8-104, src/FrameKDMap.cpp) on libampc:
//   QueryNearest        fast path on the current frame when the point projects into its frustum
//                       and the cloud has >= k points (:322-346), otherwise every frame of the
//                       query vector is searched and the results are merged by distance (:347-376)
//   GetNearestDistance  1-NN over all frames (:400-427)
//   ProcessKeyframes    one pass of KeyframeThreadWorker (:437-488): prune old key-frames, keep
//                       only the outlier points of the last one, insert the current frame
// Every frame (current + key-frames) is one scene slot of the handle; a multi-frame query is ONE
// batched k-NN launch over the frames instead of the reference's per-query std::thread fan-out.
// Not ported: building the two clouds from the depth image
// (ProcessDepth/BuildEdgeCloud, :90-214) -- the clouds are handed in ready-made (AddClouds).
#ifndef FRAME_KD_MAP_H
#define FRAME_KD_MAP_H
#include "../../include/ampc.h"
#include "eigen_compat.h"
#include "pcl_compat.h"

#include <array>
#include <deque>
#include <memory>
#include <vector>

using Mat4 = std::array<double, 16>; // row-major homogeneous transform
Mat4 Mat4Identity();

struct MapParams { // config/mpc_parameters.yaml:59-75, already divided by resize_scale (:FrameKDMap.cpp:21-24)
    double fx = 32, fy = 32, cx = 32, cy = 24;
    int width = 64, height = 48;      // mParamWidth / mParamHeight (set in ProcessDepth, :106-107)
    double depthMax = 100, depthMin = 0.1;
    double keyframeDistanceTh = 0.1;  // keyframe_th_dist
    int keyframeCountTh = 10;         // keyframe_th_count
    int maxFrameCount = 100;          // max_frame_count
    Mat4 Tbc = {0, 0, 1, 0.05, -1, 0, 0, 0, 0, -1, 0, 0.01, 0, 0, 0, 1};
};

class FrameKDMap {
public:
    explicit FrameKDMap(int maxPoints = 65536, int maxEdgePoints = 16384, const MapParams &p = MapParams());
    using CloudPtr = pcl::PointCloud<pcl::PointXYZ>::Ptr;
    // replaces AddVertex's two InitializeNew calls + swap into mCurFrame (:44-51); Twc = Twb * Tbc
    void AddClouds(const CloudPtr &cloud, const CloudPtr &edgeCloud, const Mat4 &Twc = Mat4Identity(),
                   int rowWidthHint = 0);
    void QueryNearest(const Eigen::Vector3d &point, int nearestPointCount,
                      std::vector<Eigen::Vector3d> &out, std::vector<double> &distances,
                      bool queryEdge = false);
    double GetNearestDistance(const Eigen::Vector3d &point);
    // batched current-frame query used by the tick loop: Q query sites, one launch
    void QueryNearestBatch(const std::vector<Eigen::Vector3d> &points, int nearestPointCount,
                           std::vector<std::vector<Eigen::Vector3d>> &out,
                           std::vector<std::vector<double>> &distances, bool queryEdge = false);
    // key-frames (the reference runs ProcessKeyframes' body in a detached thread every 30 ms)
    void ProcessKeyframes();
    void InsertKeyFrame();
    void RemoveOldVertex();
    int KeyFrameCount() const { return (int)mKeyFrames.size(); }
    int KeyFramePointCount(int i, bool edge = false) const { return mKeyFrames[i].count[edge ? 1 : 0]; }
    int PointCount(bool edge) const { return mCur.count[edge ? 1 : 0]; }
    bool PtIsInFrame(const Eigen::Vector3d &ptw, const Mat4 &Twc) const;
    ampc_handle *Handle() { return mHandle.get(); }

private:
    struct Frame {
        int slot = 0;
        int count[2] = {0, 0};
        Mat4 Twc;
        std::vector<pcl::PointXYZ> pts; // host copy of the Obstacle cloud (outlier step, snapshots)
        std::vector<pcl::PointXYZ> edge;
    };
    void Upload(Frame &f);
    bool DroneBehindPts(const Mat4 &Twc, const Frame &frame);
    std::vector<const Frame *> QueryVector() const; // UpdateQueryVector (:65-75)
    void SearchFrames(const std::vector<const Frame *> &frames, const Eigen::Vector3d &p, int k, int kind,
                      std::vector<std::vector<Eigen::Vector3d>> &pts, std::vector<std::vector<double>> &d2);
    std::shared_ptr<ampc_handle> mHandle;
    MapParams mP;
    Frame mCur;
    std::deque<Frame> mKeyFrames;
    std::vector<int> mFreeSlots;
    bool mHaveCur = false;
};
#endif
