#include "FrameKDMap.h"

#include <cmath>
#include <limits>
#include <stdexcept>
#include <string>

namespace {
[[noreturn]] void die(ampc_handle *h, const char *what) {
    throw std::runtime_error(std::string("FrameKDMap(GPU): ") + what + ": " + ampc_last_error(h));
}
} // namespace

FrameKDMap::FrameKDMap(int maxPoints, int maxEdgePoints) {
    ampc_config cfg{};
    cfg.N = 1, cfg.K = 1, cfg.dt = 1.0;
    cfg.max_batch = 64; // query sites per QueryNearestBatch call
    cfg.max_scenes = 1;
    cfg.max_points = maxPoints;
    cfg.max_edge_points = maxEdgePoints;
    cfg.device = 0;
    ampc_handle *h = nullptr;
    if (ampc_create(&cfg, &h) != AMPC_OK)
        die(nullptr, "ampc_create");
    mHandle.reset(h, ampc_destroy);
}

void FrameKDMap::AddClouds(const CloudPtr &cloud, const CloudPtr &edgeCloud) {
    const CloudPtr *src[2] = {&cloud, &edgeCloud};
    for (int kind = 0; kind < 2; ++kind) {
        const auto &pts = (*src[kind])->points;
        if (ampc_cloud_set(mHandle.get(), 0, kind, pts.data(), (int)pts.size(), 16) != AMPC_OK)
            die(mHandle.get(), "ampc_cloud_set");
        int32_t n = 0;
        if (ampc_cloud_count(mHandle.get(), 0, kind, &n) != AMPC_OK)
            die(mHandle.get(), "ampc_cloud_count");
        mCount[kind] = n; // after the NaN filter of KDTreeTwo::Initialize (kd_tree_two.h:99-101)
    }
}

void FrameKDMap::QueryNearestBatch(const std::vector<Eigen::Vector3d> &points, int k,
                                   std::vector<std::vector<Eigen::Vector3d>> &out,
                                   std::vector<std::vector<double>> &distances, bool queryEdge) {
    const int Q = (int)points.size();
    out.assign(Q, {});
    distances.assign(Q, {});
    const int kind = queryEdge ? AMPC_CLOUD_EDGE : AMPC_CLOUD_OBSTACLE;
    const int n = mCount[kind];
    if (Q == 0 || n == 0 || k <= 0)
        return;
    // FrameKDMap.cpp:339-345 / :293-297: ask for min(k, n) neighbours; SearchForNearest then
    // returns nothing when the cloud holds exactly that many points (kd_tree_two.h:117-124)
    const int kq = k < n ? k : n;
    std::vector<double> q(3 * Q), d2((size_t)Q * kq), pts((size_t)Q * kq * 3);
    std::vector<int32_t> cnt(Q);
    for (int i = 0; i < Q; ++i)
        q[3 * i] = points[i].x(), q[3 * i + 1] = points[i].y(), q[3 * i + 2] = points[i].z();
    // one instance (scene 0), Q queries
    if (ampc_knn_batch(mHandle.get(), kind, 1, nullptr, q.data(), Q, kq, nullptr, d2.data(),
                       pts.data(), cnt.data()) != AMPC_OK)
        die(mHandle.get(), "ampc_knn_batch");
    for (int i = 0; i < Q; ++i)
        for (int j = 0; j < cnt[i]; ++j) {
            const double *p = &pts[((size_t)i * kq + j) * 3];
            out[i].emplace_back(p[0], p[1], p[2]);
            distances[i].push_back(d2[(size_t)i * kq + j]); // squared, as in the reference
        }
}

void FrameKDMap::QueryNearest(const Eigen::Vector3d &point, int nearestPointCount,
                              std::vector<Eigen::Vector3d> &out, std::vector<double> &distances,
                              bool queryEdge) {
    std::vector<std::vector<Eigen::Vector3d>> o;
    std::vector<std::vector<double>> d;
    QueryNearestBatch({point}, nearestPointCount, o, d, queryEdge);
    out = o.empty() ? std::vector<Eigen::Vector3d>() : o[0];
    distances = d.empty() ? std::vector<double>() : d[0];
}

double FrameKDMap::GetNearestDistance(const Eigen::Vector3d &point) {
    // FrameKDMap.cpp:400-427: sqrt of the smallest squared 1-NN distance over the frames
    double nearest = std::numeric_limits<double>::max();
    if (mCount[0] == 0)
        return nearest; // empty map: returned un-rooted (:402-404)
    std::vector<Eigen::Vector3d> o;
    std::vector<double> d;
    QueryNearest(point, 1, o, d, false);
    if (!d.empty())
        nearest = d[0];
    return std::sqrt(nearest);
}
