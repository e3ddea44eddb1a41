#include "FrameKDMap.h"

#include <algorithm>
#include <cmath>
#include <limits>
#include <stdexcept>
#include <string>

namespace {
[[noreturn]] void die(ampc_handle *h, const char *what) {
    throw std::runtime_error(std::string("FrameKDMap(GPU): ") + what + ": " + ampc_last_error(h));
}
Mat4 mul(const Mat4 &a, const Mat4 &b) {
    Mat4 c{};
    for (int i = 0; i < 4; ++i)
        for (int j = 0; j < 4; ++j) {
            double s = 0;
            for (int l = 0; l < 4; ++l)
                s += a[4 * i + l] * b[4 * l + j];
            c[4 * i + j] = s;
        }
    return c;
}
Mat4 rigid_inverse(const Mat4 &T) { // [R t; 0 1]^-1 = [R' -R't; 0 1]
    Mat4 I{};
    for (int i = 0; i < 3; ++i) {
        for (int j = 0; j < 3; ++j)
            I[4 * i + j] = T[4 * j + i];
        I[4 * i + 3] = -(T[0 + i] * T[3] + T[4 + i] * T[7] + T[8 + i] * T[11]);
    }
    I[15] = 1;
    return I;
}
} // namespace

Mat4 Mat4Identity() { return {1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1}; }

FrameKDMap::FrameKDMap(int maxPoints, int maxEdgePoints, const MapParams &p) : mP(p) {
    ampc_config cfg{};
    cfg.N = 1, cfg.K = 1, cfg.dt = 1.0;
    cfg.max_batch = std::max(64, p.maxFrameCount + 4);
    cfg.max_scenes = p.maxFrameCount + 4; // current frame, the frame being built, key-frames (+1 before pruning)
    cfg.max_points = maxPoints;
    cfg.max_edge_points = maxEdgePoints;
    cfg.device = 0;
    ampc_handle *h = nullptr;
    if (ampc_create(&cfg, &h) != AMPC_OK)
        die(nullptr, "ampc_create");
    mHandle.reset(h, ampc_destroy);
    ampc_camera cam{p.fx, p.fy, p.cx, p.cy, p.resizeScale, p.pixel2Meter, p.depthMin, p.depthMax};
    if (ampc_set_camera(h, &cam) != AMPC_OK)
        die(h, "ampc_set_camera");
    mFreeSlots = std::make_shared<std::vector<int>>();
    for (int s = cfg.max_scenes - 1; s >= 0; --s)
        mFreeSlots->push_back(s);
    mCur.Twc = Mat4Identity(); // the reference leaves Frame::Twc uninitialised until the first AddVertex
}

std::shared_ptr<FrameKDMap::Cloud> FrameKDMap::NewCloud() {
    if (mFreeSlots->empty())
        throw std::runtime_error("FrameKDMap(GPU): out of scene slots");
    Cloud *c = new Cloud();
    c->slot = mFreeSlots->back();
    mFreeSlots->pop_back();
    std::shared_ptr<std::vector<int>> pool = mFreeSlots;
    return std::shared_ptr<Cloud>(c, [pool](Cloud *p) {
        pool->push_back(p->slot);
        delete p;
    });
}

void FrameKDMap::RefreshCounts(Cloud &c) {
    for (int kind = 0; kind < 2; ++kind) {
        int32_t n = 0;
        if (ampc_cloud_count(mHandle.get(), c.slot, kind, &n) != AMPC_OK)
            die(mHandle.get(), "ampc_cloud_count");
        c.count[kind] = n; // after the NaN filter of KDTreeTwo::Initialize (kd_tree_two.h:99-101)
    }
}

std::vector<pcl::PointXYZ> FrameKDMap::Download(const Cloud &c, int kind) {
    static_assert(sizeof(pcl::PointXYZ) == 16, "pcl::PointXYZ is a 16-byte record");
    std::vector<pcl::PointXYZ> pts(c.count[kind]);
    int32_t n = 0;
    if (ampc_cloud_get(mHandle.get(), c.slot, kind, pts.data(), (int32_t)pts.size(), &n) != AMPC_OK)
        die(mHandle.get(), "ampc_cloud_get");
    pts.resize(n);
    return pts;
}

std::vector<pcl::PointXYZ> FrameKDMap::CurrentPoints(bool edge) {
    std::lock_guard<std::recursive_mutex> lock(mMtxKdTree);
    return mCur.cloud ? Download(*mCur.cloud, edge ? 1 : 0) : std::vector<pcl::PointXYZ>();
}

void FrameKDMap::AddVertex(const Mat4 &Twb, const DepthImage &depth) {
    std::lock_guard<std::recursive_mutex> lock(mMtxKdTree);
    if (!depth.data || depth.rows < 1 || depth.cols < 1)
        throw std::runtime_error("FrameKDMap(GPU): empty depth image");
    const Mat4 Twc = mul(Twb, mP.Tbc);        // :117-119
    const Mat4 Tedge = mul(mCur.Twc, mP.Tbc); // mCurFrame.Twc (previous frame) * mParamTbc, :208-209
    std::shared_ptr<Cloud> c = NewCloud();
    const size_t step = depth.step ? depth.step : (size_t)depth.cols * (depth.isU16 ? 2 : 4);
    if (ampc_depth_set_batch(mHandle.get(), c->slot, 1, depth.data, depth.isU16 ? AMPC_DEPTH_U16 : AMPC_DEPTH_F32,
                             depth.rows, depth.cols, (int64_t)step, 0, Twc.data(), Tedge.data()) != AMPC_OK)
        die(mHandle.get(), "ampc_depth_set_batch");
    mP.width = (int)(depth.cols / mP.resizeScale); // :106-107
    mP.height = (int)(depth.rows / mP.resizeScale);
    RefreshCounts(*c);
    if (c->count[0] == 0)
        return; // :41-43: the previous frame stays current
    mCur.cloud = c;
    mCur.Twc = Twc;
    mHaveCur = true;
}

void FrameKDMap::Upload(Cloud &c, int kind, const std::vector<pcl::PointXYZ> &pts) {
    if (ampc_cloud_set(mHandle.get(), c.slot, kind, pts.data(), (int)pts.size(), 16) != AMPC_OK)
        die(mHandle.get(), "ampc_cloud_set");
}

void FrameKDMap::AddClouds(const CloudPtr &cloud, const CloudPtr &edgeCloud, const Mat4 &Twc, int rowWidthHint) {
    std::lock_guard<std::recursive_mutex> lock(mMtxKdTree);
    if (ampc_cloud_set_layout(mHandle.get(), AMPC_CLOUD_OBSTACLE, rowWidthHint) != AMPC_OK)
        die(mHandle.get(), "ampc_cloud_set_layout");
    std::shared_ptr<Cloud> c = NewCloud();
    Upload(*c, 0, cloud->points);
    Upload(*c, 1, edgeCloud->points);
    if (ampc_cloud_set_layout(mHandle.get(), AMPC_CLOUD_OBSTACLE, 0) != AMPC_OK)
        die(mHandle.get(), "ampc_cloud_set_layout");
    RefreshCounts(*c);
    mCur.cloud = c;
    mCur.Twc = Twc;
    mHaveCur = true;
}
bool FrameKDMap::PtIsInFrame(const Eigen::Vector3d &ptw, const Mat4 &Twc) const { // :215-231
    const Mat4 Tcw = rigid_inverse(Twc);
    const double x = Tcw[0] * ptw.x() + Tcw[1] * ptw.y() + Tcw[2] * ptw.z() + Tcw[3];
    const double y = Tcw[4] * ptw.x() + Tcw[5] * ptw.y() + Tcw[6] * ptw.z() + Tcw[7];
    const double z = Tcw[8] * ptw.x() + Tcw[9] * ptw.y() + Tcw[10] * ptw.z() + Tcw[11];
    if (z > mP.depthMax || z < 0)
        return false;
    const double sc = mP.resizeScale; // intrinsics of the resized image, :21-24
    const double u = (mP.fx / sc) * x / z + mP.cx / sc, v = (mP.fy / sc) * y / z + mP.cy / sc;
    return !(u < 0 || u >= mP.width || v < 0 || v >= mP.height);
}

std::vector<const FrameKDMap::Frame *> FrameKDMap::QueryVector() const { // :65-75
    std::vector<const Frame *> v;
    if (mHaveCur)
        v.push_back(&mCur);
    for (size_t i = 0; i + 1 < mKeyFrames.size(); ++i) // all key-frames but the last
        v.push_back(&mKeyFrames[i]);
    return v;
}

// one batched launch: instance f = frame f, the same query for every frame.  Frames follow the
// reference's per-frame rule (:293-297 + kd_tree_two.h:117-124): queryPointCount = min(k, n), and
// SearchForNearest returns nothing when n == queryPointCount, i.e. only frames with n > k answer.
void FrameKDMap::SearchFrames(const std::vector<const Frame *> &frames, const Eigen::Vector3d &p, int k,
                              int kind, std::vector<std::vector<Eigen::Vector3d>> &pts,
                              std::vector<std::vector<double>> &d2) {
    const int F = (int)frames.size();
    pts.assign(F, {});
    d2.assign(F, {});
    if (F == 0 || k <= 0)
        return;
    std::vector<int32_t> scene_of(F), cnt(F);
    std::vector<double> q(3 * F), dd((size_t)F * k), pp((size_t)F * k * 3);
    for (int f = 0; f < F; ++f) {
        scene_of[f] = frames[f]->cloud->slot;
        q[3 * f] = p.x(), q[3 * f + 1] = p.y(), q[3 * f + 2] = p.z();
    }
    if (ampc_knn_batch(mHandle.get(), kind, F, scene_of.data(), q.data(), 1, k, nullptr, dd.data(), pp.data(),
                       cnt.data()) != AMPC_OK)
        die(mHandle.get(), "ampc_knn_batch");
    for (int f = 0; f < F; ++f) {
        if (frames[f]->cloud->count[kind] <= k)
            continue; // n < k asks for n and gets none; n == k gets none
        for (int j = 0; j < cnt[f]; ++j) {
            const double *c = &pp[((size_t)f * k + j) * 3];
            pts[f].emplace_back(c[0], c[1], c[2]);
            d2[f].push_back(dd[(size_t)f * k + j]);
        }
    }
}

void FrameKDMap::SearchFramesMany(const std::vector<const Frame *> &frames, const std::vector<Eigen::Vector3d> &ps,
                                  int k, int kind, std::vector<std::vector<std::vector<Eigen::Vector3d>>> &pts,
                                  std::vector<std::vector<std::vector<double>>> &d2) {
    const int F = (int)frames.size(), Q = (int)ps.size();
    pts.assign(Q, std::vector<std::vector<Eigen::Vector3d>>(F));
    d2.assign(Q, std::vector<std::vector<double>>(F));
    if (F == 0 || Q == 0 || k <= 0)
        return;
    std::vector<int32_t> scene_of(F), cnt((size_t)F * Q);
    std::vector<double> q((size_t)3 * F * Q), dd((size_t)F * Q * k), pp((size_t)F * Q * k * 3);
    for (int f = 0; f < F; ++f) {
        scene_of[f] = frames[f]->cloud->slot;
        for (int i = 0; i < Q; ++i) {
            double *d = &q[((size_t)f * Q + i) * 3];
            d[0] = ps[i].x(), d[1] = ps[i].y(), d[2] = ps[i].z();
        }
    }
    if (ampc_knn_batch(mHandle.get(), kind, F, scene_of.data(), q.data(), Q, k, nullptr, dd.data(), pp.data(),
                       cnt.data()) != AMPC_OK)
        die(mHandle.get(), "ampc_knn_batch");
    for (int f = 0; f < F; ++f) {
        if (frames[f]->cloud->count[kind] <= k)
            continue; // n < k asks for n and gets none; n == k gets none
        for (int i = 0; i < Q; ++i)
            for (int j = 0; j < cnt[(size_t)f * Q + i]; ++j) {
                const double *c = &pp[(((size_t)f * Q + i) * k + j) * 3];
                pts[i][f].emplace_back(c[0], c[1], c[2]);
                d2[i][f].push_back(dd[((size_t)f * Q + i) * k + j]);
            }
    }
}

void FrameKDMap::QueryNearestMany(const std::vector<Eigen::Vector3d> &points, int k,
                                  std::vector<std::vector<Eigen::Vector3d>> &out,
                                  std::vector<std::vector<double>> &distances, bool queryEdge) {
    std::lock_guard<std::recursive_mutex> lock(mMtxKdTree);
    const int Q = (int)points.size();
    out.assign(Q, {});
    distances.assign(Q, {});
    const int kind = queryEdge ? AMPC_CLOUD_EDGE : AMPC_CLOUD_OBSTACLE;
    std::vector<int> fast, slow;
    const bool cur_ok = mHaveCur && mCur.cloud->count[kind] >= k;
    for (int i = 0; i < Q; ++i)
        (cur_ok && PtIsInFrame(points[i], mCur.Twc) ? fast : slow).push_back(i);
    if (!fast.empty()) { // fast path (:329-346)
        std::vector<Eigen::Vector3d> sites;
        for (int i : fast) sites.push_back(points[i]);
        std::vector<std::vector<Eigen::Vector3d>> o;
        std::vector<std::vector<double>> d;
        QueryNearestBatch(sites, k, o, d, queryEdge);
        for (size_t j = 0; j < fast.size(); ++j)
            out[fast[j]] = o[j], distances[fast[j]] = d[j];
    }
    if (!slow.empty()) { // slow path (:347-375): every frame of the query vector, merged by distance
        std::vector<Eigen::Vector3d> sites;
        for (int i : slow) sites.push_back(points[i]);
        std::vector<std::vector<std::vector<Eigen::Vector3d>>> pts;
        std::vector<std::vector<std::vector<double>>> d2;
        SearchFramesMany(QueryVector(), sites, k, kind, pts, d2);
        for (size_t j = 0; j < slow.size(); ++j) {
            std::vector<std::pair<double, Eigen::Vector3d>> all;
            for (size_t f = 0; f < pts[j].size(); ++f)
                for (size_t e = 0; e < pts[j][f].size(); ++e)
                    all.emplace_back(d2[j][f][e], pts[j][f][e]);
            std::stable_sort(all.begin(), all.end(), [](const auto &a, const auto &b) { return a.first < b.first; });
            for (int e = 0; e < k && e < (int)all.size(); ++e) {
                out[slow[j]].push_back(all[e].second);
                distances[slow[j]].push_back(all[e].first);
            }
        }
    }
}

FrameKDMap::CloudPtr FrameKDMap::GetPtCloud() { // :489-503
    std::lock_guard<std::recursive_mutex> lock(mMtxKdTree);
    auto all = std::make_shared<pcl::PointCloud<pcl::PointXYZ>>();
    for (const Frame *f : QueryVector()) {
        const std::vector<pcl::PointXYZ> pts = Download(*f->cloud, AMPC_CLOUD_OBSTACLE);
        all->points.insert(all->points.end(), pts.begin(), pts.end());
    }
    return all;
}

void FrameKDMap::QueryNearestBatch(const std::vector<Eigen::Vector3d> &points, int k,
                                   std::vector<std::vector<Eigen::Vector3d>> &out,
                                   std::vector<std::vector<double>> &distances, bool queryEdge) {
    std::lock_guard<std::recursive_mutex> lock(mMtxKdTree);
    const int Q = (int)points.size();
    out.assign(Q, {});
    distances.assign(Q, {});
    const int kind = queryEdge ? AMPC_CLOUD_EDGE : AMPC_CLOUD_OBSTACLE;
    const int n = mCur.cloud ? mCur.cloud->count[kind] : 0;
    if (Q == 0 || n == 0 || k <= 0)
        return;
    const int kq = k < n ? k : n;
    std::vector<double> q(3 * Q), d2((size_t)Q * kq), pts((size_t)Q * kq * 3);
    std::vector<int32_t> cnt(Q);
    for (int i = 0; i < Q; ++i)
        q[3 * i] = points[i].x(), q[3 * i + 1] = points[i].y(), q[3 * i + 2] = points[i].z();
    const int32_t scene0 = mCur.cloud->slot;
    if (ampc_knn_batch(mHandle.get(), kind, 1, &scene0, q.data(), Q, kq, nullptr, d2.data(), pts.data(),
                       cnt.data()) != AMPC_OK)
        die(mHandle.get(), "ampc_knn_batch");
    for (int i = 0; i < Q; ++i)
        for (int j = 0; j < cnt[i]; ++j) {
            const double *p = &pts[((size_t)i * kq + j) * 3];
            out[i].emplace_back(p[0], p[1], p[2]);
            distances[i].push_back(d2[(size_t)i * kq + j]); // squared, as in the reference
        }
}

void FrameKDMap::QueryNearest(const Eigen::Vector3d &point, int k, std::vector<Eigen::Vector3d> &out,
                              std::vector<double> &distances, bool queryEdge) {
    std::lock_guard<std::recursive_mutex> lock(mMtxKdTree);
    out.clear();
    distances.clear();
    const int kind = queryEdge ? AMPC_CLOUD_EDGE : AMPC_CLOUD_OBSTACLE;
    // fast path (:329-346): the current frame alone
    if (mHaveCur && mCur.cloud->count[kind] >= k && PtIsInFrame(point, mCur.Twc)) {
        std::vector<std::vector<Eigen::Vector3d>> o;
        std::vector<std::vector<double>> d;
        QueryNearestBatch({point}, k, o, d, queryEdge);
        out = o[0];
        distances = d[0];
        return;
    }
    // slow path (:347-375): every frame of the query vector, merge, sort by distance, keep k
    std::vector<std::vector<Eigen::Vector3d>> pts;
    std::vector<std::vector<double>> d2;
    SearchFrames(QueryVector(), point, k, kind, pts, d2);
    std::vector<std::pair<double, Eigen::Vector3d>> all;
    for (size_t f = 0; f < pts.size(); ++f)
        for (size_t j = 0; j < pts[f].size(); ++j)
            all.emplace_back(d2[f][j], pts[f][j]);
    std::stable_sort(all.begin(), all.end(), [](const auto &a, const auto &b) { return a.first < b.first; });
    for (int i = 0; i < k && i < (int)all.size(); ++i) {
        out.push_back(all[i].second);
        distances.push_back(all[i].first);
    }
}

double FrameKDMap::GetNearestDistance(const Eigen::Vector3d &point) { // :400-427
    double nearest = std::numeric_limits<double>::max();
    const std::vector<const Frame *> all = QueryVector();
    if (all.empty())
        return nearest; // returned un-rooted (:402-404)
    std::vector<const Frame *> frames;
    for (const Frame *f : all)
        if (f->cloud->count[0] > 0) // frames with an empty Obstacle cloud are skipped (:383-386)
            frames.push_back(f);
    std::vector<std::vector<Eigen::Vector3d>> pts;
    std::vector<std::vector<double>> d2;
    SearchFrames(frames, point, 1, AMPC_CLOUD_OBSTACLE, pts, d2);
    for (const auto &d : d2)
        if (!d.empty())
            nearest = std::min(nearest, d[0]);
    return std::sqrt(nearest);
}

bool FrameKDMap::DroneBehindPts(const Mat4 &Twc, const Frame &frame) { // :233-252
    const Mat4 Twb = mul(Twc, rigid_inverse(mP.Tbc));
    const Eigen::Vector3d twb(Twb[3], Twb[7], Twb[11]);
    const int ptsCount = std::min(frame.cloud->count[0], 10);
    std::vector<std::vector<Eigen::Vector3d>> pts;
    std::vector<std::vector<double>> d2;
    if (ptsCount > 0) {
    std::lock_guard<std::recursive_mutex> lock(mMtxKdTree);
        // SearchForNearest(twb, ptsCount) on that frame: nothing when the frame holds exactly ptsCount points
        std::vector<int32_t> cnt(1);
        std::vector<double> q = {twb.x(), twb.y(), twb.z()}, dd(ptsCount), pp(3 * (size_t)ptsCount);
        const int32_t slot = frame.cloud->slot;
        if (ampc_knn_batch(mHandle.get(), AMPC_CLOUD_OBSTACLE, 1, &slot, q.data(), 1, ptsCount, nullptr, dd.data(),
                           pp.data(), cnt.data()) != AMPC_OK)
            die(mHandle.get(), "ampc_knn_batch");
        for (int j = 0; j < cnt[0]; ++j) {
            // ptb = Rbw (ptw - twb); only its x component matters
            const double dx = pp[3 * j] - twb.x(), dy = pp[3 * j + 1] - twb.y(), dz = pp[3 * j + 2] - twb.z();
            const double ptb_x = Twb[0] * dx + Twb[4] * dy + Twb[8] * dz; // first row of Rwb' = first column of Rwb
            if (ptb_x <= mP.depthMin)
                return false;
        }
    }
    return true;
}

void FrameKDMap::InsertKeyFrame() { // :428-432: the key-frame shares the current frame's trees
    if (!mHaveCur)
        return;
    mKeyFrames.push_back(mCur);
}

void FrameKDMap::RemoveOldVertex() { // :59-63
    if (!mKeyFrames.empty())
        mKeyFrames.pop_front();
}

void FrameKDMap::ProcessKeyframes() { // body of KeyframeThreadWorker, :446-487
    if (!mHaveCur)
        return;
    if (mKeyFrames.empty()) {
    std::lock_guard<std::recursive_mutex> lock(mMtxKdTree);
        InsertKeyFrame();
        return;
    }
    while (!mKeyFrames.empty()) {
        if ((int)mKeyFrames.size() > mP.maxFrameCount || !DroneBehindPts(mCur.Twc, mKeyFrames.front()))
            RemoveOldVertex();
        else
            break;
    }
    if (mKeyFrames.empty())
        return;
    // points of the last key-frame farther than keyframe_th_dist from the current cloud
    Cloud &last = *mKeyFrames.back().cloud;
    const std::vector<pcl::PointXYZ> lastPts = Download(last, 0);
    const int n = (int)lastPts.size();
    std::vector<pcl::PointXYZ> outliers;
    if (n > 0 && mCur.cloud->count[0] > 0) {
        std::vector<double> q(3 * (size_t)n), d2(n);
        std::vector<int32_t> cnt(n);
        for (int i = 0; i < n; ++i)
            q[3 * i] = lastPts[i].x, q[3 * i + 1] = lastPts[i].y, q[3 * i + 2] = lastPts[i].z;
        const int32_t slot = mCur.cloud->slot;
        if (ampc_knn_batch(mHandle.get(), AMPC_CLOUD_OBSTACLE, 1, &slot, q.data(), n, 1, nullptr, d2.data(), nullptr,
                           cnt.data()) != AMPC_OK)
            die(mHandle.get(), "ampc_knn_batch");
        for (int i = 0; i < n; ++i)
            if (cnt[i] > 0 && std::sqrt(d2[i]) > mP.keyframeDistanceTh)
                outliers.push_back(lastPts[i]);
    }
    if ((int)outliers.size() < mP.keyframeCountTh)
        return;
    Upload(last, 0, outliers); // lastPtCloudPtr->InitializeNew(newCloud), :484: in place, seen by every frame sharing it
    RefreshCounts(last);
    InsertKeyFrame();
}
