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
//   AddVertex           depth image -> Obstacle + Edge cloud + their indices, all on the device
//                       (ProcessDepth / BuildEdgeCloud incl. the OpenCV calls, :36-53,76-214)
// AddClouds is the same entry for callers that already hold the two clouds.
#ifndef FRAME_KD_MAP_H
#define FRAME_KD_MAP_H
#include "../../include/ampc.h"
#include "eigen_compat.h"
#include "pcl_compat.h"

#include <array>
#include <deque>
#include <memory>
#include <mutex>
#include <vector>

// -DAMPC_WITH_ROS in a catkin build: AddVertex also takes the reference's own argument types
#if defined(AMPC_WITH_ROS) && __has_include(<sensor_msgs/Image.h>) && __has_include(<Eigen/Dense>)
#include <sensor_msgs/Image.h>
#include <sensor_msgs/image_encodings.h>
#include <stdexcept>
#define AMPC_HAVE_ROS_TYPES 1
#endif

using Mat4 = std::array<double, 16>; // row-major homogeneous transform
Mat4 Mat4Identity();

struct DepthImage { // what cv_bridge hands ProcessDepth (:94): CV_32FC1 metres or CV_16UC1
    const void *data = nullptr;
    int rows = 0, cols = 0;
    bool isU16 = false;
    size_t step = 0; // bytes per row; 0 = tightly packed
};

struct MapParams { // config/mpc_parameters.yaml:58-75; fx..cy at full resolution, as in the yaml
    double fx = 320, fy = 320, cx = 320, cy = 240;
    double resizeScale = 10, pixel2Meter = 1;
    int width = 64, height = 48;      // mParamWidth / mParamHeight: set by AddVertex (:106-107), or by the caller
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
    // AddVertex (:36-53): a frame whose Obstacle cloud comes out empty is dropped (:41-43)
    void AddVertex(const Mat4 &Twb, const DepthImage &depth);
#ifdef AMPC_HAVE_ROS_TYPES
    // the reference's signature (include/FrameKDMap.h:61-62).  The message buffer is handed over as it
    // is (the reference copies it through cv_bridge::toCvCopy, src/FrameKDMap.cpp:94); CV_16UC1 and
    // CV_32FC1 are the two types ProcessDepth accepts (:96-103)
    void AddVertex(const Eigen::Matrix4d &mat4Twb, const sensor_msgs::ImageConstPtr &depth) {
        namespace enc = sensor_msgs::image_encodings;
        Mat4 T;
        for (int r = 0; r < 4; ++r)
            for (int c = 0; c < 4; ++c)
                T[4 * r + c] = mat4Twb(r, c);
        DepthImage d;
        d.data = depth->data.data();
        d.rows = (int)depth->height, d.cols = (int)depth->width, d.step = depth->step;
        d.isU16 = depth->encoding == enc::TYPE_16UC1 || depth->encoding == enc::MONO16;
        if (!d.isU16 && depth->encoding != enc::TYPE_32FC1)
            throw std::runtime_error("FrameKDMap::AddVertex: depth image must be 16UC1 or 32FC1");
        AddVertex(T, d);
    }
#endif
    // the two InitializeNew calls + swap into mCurFrame (:44-51) for ready-made clouds; Twc = Twb * Tbc
    void AddClouds(const CloudPtr &cloud, const CloudPtr &edgeCloud, const Mat4 &Twc = Mat4Identity(),
                   int rowWidthHint = 0);
    void QueryNearest(const Eigen::Vector3d &point, int nearestPointCount,
                      std::vector<Eigen::Vector3d> &out, std::vector<double> &distances,
                      bool queryEdge = false);
    double GetNearestDistance(const Eigen::Vector3d &point);
    // batched current-frame query: Q query sites, one launch, the current frame only
    void QueryNearestBatch(const std::vector<Eigen::Vector3d> &points, int nearestPointCount,
                           std::vector<std::vector<Eigen::Vector3d>> &out,
                           std::vector<std::vector<double>> &distances, bool queryEdge = false);
    // Q query sites with QueryNearest's semantics for EACH of them (fast path on the current frame
    // if the site projects into its frustum and the cloud has >= k points, else all frames of the
    // query vector merged by distance, :329-376): at most two launches -- the in-frame sites on the
    // current frame, the others over (frames x sites).  What ProcessWaypoints needs (:204-235).
    void QueryNearestMany(const std::vector<Eigen::Vector3d> &points, int nearestPointCount,
                          std::vector<std::vector<Eigen::Vector3d>> &out,
                          std::vector<std::vector<double>> &distances, bool queryEdge = false);
    // all points of the query vector's Obstacle clouds, current frame first (GetPtCloud, :489-503)
    CloudPtr GetPtCloud();
    // key-frames (the reference runs ProcessKeyframes' body in a detached thread every 30 ms)
    void ProcessKeyframes();
    void InsertKeyFrame();
    void RemoveOldVertex();
    int KeyFrameCount() const { return (int)mKeyFrames.size(); }
    int KeyFramePointCount(int i, bool edge = false) const { return mKeyFrames[i].cloud->count[edge ? 1 : 0]; }
    int PointCount(bool edge) const { return mCur.cloud ? mCur.cloud->count[edge ? 1 : 0] : 0; }
    // host copy of the current frame's cloud (KDTreeTwo::GetPointCloud().pts)
    std::vector<pcl::PointXYZ> CurrentPoints(bool edge = false);
    bool PtIsInFrame(const Eigen::Vector3d &ptw, const Mat4 &Twc) const;
    ampc_handle *Handle() { return mHandle.get(); }

private:
    // one scene slot of the handle = the pair of trees a reference Frame points to; shared between
    // the current frame and the key-frame made from it, like the reference's shared_ptr<KDTreeTwo>
    struct Cloud {
        int slot = 0;
        int count[2] = {0, 0};
    };
    struct Frame {
        std::shared_ptr<Cloud> cloud;
        Mat4 Twc;
    };
    std::shared_ptr<Cloud> NewCloud();
    void Upload(Cloud &c, int kind, const std::vector<pcl::PointXYZ> &pts);
    std::vector<pcl::PointXYZ> Download(const Cloud &c, int kind);
    void RefreshCounts(Cloud &c);
    bool DroneBehindPts(const Mat4 &Twc, const Frame &frame);
    std::vector<const Frame *> QueryVector() const; // UpdateQueryVector (:65-75)
    void SearchFrames(const std::vector<const Frame *> &frames, const Eigen::Vector3d &p, int k, int kind,
                      std::vector<std::vector<Eigen::Vector3d>> &pts, std::vector<std::vector<double>> &d2);
    // the same for several sites in one launch: result [site][frame]
    void SearchFramesMany(const std::vector<const Frame *> &frames, const std::vector<Eigen::Vector3d> &ps, int k,
                          int kind, std::vector<std::vector<std::vector<Eigen::Vector3d>>> &pts,
                          std::vector<std::vector<std::vector<double>>> &d2);
    // mMtxKdTree of the reference (:48,328,365,406,425,443,491): tree swap, queries and the
    // key-frame pass exclude each other; recursive because the public entries call each other
    mutable std::recursive_mutex mMtxKdTree;
    std::shared_ptr<ampc_handle> mHandle;
    MapParams mP;
    std::shared_ptr<std::vector<int>> mFreeSlots; // outlives every Cloud
    Frame mCur;
    std::deque<Frame> mKeyFrames;
    bool mHaveCur = false;
};
#endif
