// Exercises the drop-in C++ classes end to end on the GPU (run by tests/test_gpu_shim.py).
#include "AvoidanceTick.h"
#include "kd_tree_two.h"

#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstring>
#include <random>

#define CHECK(c)                                                                                   \
    do {                                                                                           \
        if (!(c)) {                                                                                \
            std::printf("CHECK failed %s:%d: %s\n", __FILE__, __LINE__, #c);                       \
            return 1;                                                                              \
        }                                                                                          \
    } while (0)

// AddVertex on two depth frames written by tests/test_gpu_shim.py together with the clouds the
// test expects for the second one (Edge points use the FIRST frame's Twc, FrameKDMap.cpp:208-209)
static int depth_case(const char *path) {
    FILE *f = std::fopen(path, "rb");
    CHECK(f != nullptr);
    int32_t hd[4];
    double par[8], Twb0[16], Twb1[16];
    CHECK(std::fread(hd, 4, 4, f) == 4 && std::fread(par, 8, 8, f) == 8);
    CHECK(std::fread(Twb0, 8, 16, f) == 16 && std::fread(Twb1, 8, 16, f) == 16);
    const int rows = hd[0], cols = hd[1], nc = hd[2], ne = hd[3];
    std::vector<float> d0((size_t)rows * cols), d1(d0.size()), cloud((size_t)nc * 4), edge((size_t)ne * 4);
    CHECK(std::fread(d0.data(), 4, d0.size(), f) == d0.size() && std::fread(d1.data(), 4, d1.size(), f) == d1.size());
    CHECK(std::fread(cloud.data(), 4, cloud.size(), f) == cloud.size());
    CHECK(std::fread(edge.data(), 4, edge.size(), f) == edge.size());
    std::fclose(f);
    MapParams mp;
    mp.fx = par[0], mp.fy = par[1], mp.cx = par[2], mp.cy = par[3], mp.resizeScale = par[4], mp.pixel2Meter = par[5];
    mp.depthMin = par[6], mp.depthMax = par[7], mp.maxFrameCount = 2;
    const int H = (int)(rows / mp.resizeScale), W = (int)(cols / mp.resizeScale);
    FrameKDMap map(H * W, H * W, mp);
    Mat4 T0, T1;
    std::copy(Twb0, Twb0 + 16, T0.begin());
    std::copy(Twb1, Twb1 + 16, T1.begin());
    DepthImage img;
    img.rows = rows, img.cols = cols;
    img.data = d0.data();
    map.AddVertex(T0, img);
    CHECK(map.PointCount(false) > 0);
    std::vector<float> sky(d0.size(), 1e4f); // nothing in range: the frame is dropped, frame 0 stays
    img.data = sky.data();
    const int before = map.PointCount(false);
    map.AddVertex(T1, img);
    CHECK(map.PointCount(false) == before);
    img.data = d1.data();
    map.AddVertex(T1, img);
    CHECK(map.PointCount(false) == nc && map.PointCount(true) == ne);
    const auto pc = map.CurrentPoints(false), pe = map.CurrentPoints(true);
    CHECK(std::memcmp(pc.data(), cloud.data(), cloud.size() * 4) == 0);
    CHECK(std::memcmp(pe.data(), edge.data(), edge.size() * 4) == 0);
    std::vector<Eigen::Vector3d> out;
    std::vector<double> dist;
    map.QueryNearest(Eigen::Vector3d(cloud[0], cloud[1], cloud[2]), 4, out, dist);
    CHECK(dist.size() == 4 && dist[0] == 0.0);
    std::printf("depth frames ok: %d obstacle points, %d edge points\n", nc, ne);
    return 0;
}

int main(int argc, char **argv) {
    if (argc > 1 && depth_case(argv[1]) != 0)
        return 1;
    std::mt19937 rng(7);
    std::uniform_real_distribution<float> ux(2.f, 30.f), uy(-6.f, 6.f), uz(0.f, 3.f);
    auto cloud = std::make_shared<pcl::PointCloud<pcl::PointXYZ>>();
    auto edge = std::make_shared<pcl::PointCloud<pcl::PointXYZ>>();
    for (int i = 0; i < 20000; ++i)
        cloud->points.emplace_back(ux(rng), uy(rng), uz(rng));
    for (int i = 0; i < 500; ++i)
        edge->points.emplace_back(ux(rng), uy(rng), uz(rng));
    cloud->points[17].x = NAN; // dropped by Initialize

    // ---- KDTreeTwo: same call pattern as FrameKDMap.cpp:262-274
    KDTreeTwo<double> tree;
    tree.InitializeNew(cloud);
    CHECK(tree.GetPointCloud().pts.size() == 19999);
    tree.SearchForNearest(5.0, 0.5, 1.5, 8);
    CHECK(tree.indices.size() == 8 && tree.squared_distances.size() == 8 && tree.closest_pts.size() == 8);
    {   // brute force in double, reference arithmetic
        const auto &pts = tree.GetPointCloud().pts;
        for (int r = 0; r < 8; ++r) {
            const auto &p = pts[tree.indices[r]];
            const double d0 = 5.0 - p.x, d1 = 0.5 - p.y, d2 = 1.5 - p.z;
            CHECK(tree.squared_distances[r] == (d0 * d0 + d1 * d1) + d2 * d2);
            if (r) CHECK(tree.squared_distances[r] >= tree.squared_distances[r - 1]);
        }
        int closer = 0;
        for (const auto &p : pts) {
            const double d0 = 5.0 - p.x, d1 = 0.5 - p.y, d2 = 1.5 - p.z;
            if ((d0 * d0 + d1 * d1) + d2 * d2 < tree.squared_distances[7]) ++closer;
        }
        CHECK(closer == 7);
    }

    // ---- ObstacleAvoidanceMPC + FrameKDMap driven by the tick loop (SetupMPC, :55-85)
    TickParams tp;
    tp.T = 1.0, tp.dt = 0.05, tp.nearestPointNum = 16;
    ObstacleAvoidanceMPC mpc0(tp.T, tp.dt, "so/mpc_obstacle_v2.so");
    ObstacleAvoidanceMPC mpc;
    mpc = mpc0; // the reference copy-assigns the solver (AvoidanceStateMachine.cpp:62-63)
    mpc.SetupWeights({50, 50, 100, 100, 1, 1, 1, 0, 0, 0, 0, 10, 50, 100, 0, 1, 1, 0, 1, 1, 0.3, 0.3, 0.5, 1.0, 1.2});
    mpc.SetupTau({6.09837416, 6.21675029, 15.79816293, 0.});
    mpc.SetupGains({0.999999, 0.999999, 0.999999, 1.});
    mpc.SetDroneAccelLimits(5., 15., 10., 10.);
    mpc.SetDroneRadius(0.5);
    FrameKDMap map(65536, 4096);
    map.AddClouds(cloud, edge);
    CHECK(map.PointCount(false) == 19999 && map.PointCount(true) == 500);
    CHECK(map.GetNearestDistance(Eigen::Vector3d(5, 0.5, 1.5)) == std::sqrt(tree.squared_distances[0]));

    AvoidanceTick tick(tp, mpc, map);
    Eigen::Vector3d pos(0, 0, 1.5), vel(3, 0, 0), acc(0, 0, 0);
    for (int t = 0; t < 5; ++t) {
        tick.SetOdom(pos, vel, acc, 0.0);
        TickResult r = tick.Step();
        CHECK(r.rounds >= 1 && r.u.size() == 4 && r.x0Array.size() == 20 && r.x0Array[0].size() == 14);
        CHECK(r.u[2] >= 5.0 - 1e-9 && r.u[2] <= 15.0 + 1e-9 && std::fabs(r.u[0]) <= 10.0 + 1e-9);
        CHECK(std::isfinite(mpc.LastCost()));
        const AccelCommand cmd = tick.MakeCommand(r);
        CHECK(cmd.mode == 1 && cmd.yaw == 0 && cmd.ax == r.u[0] && cmd.az == r.u[2]);
        TickResult unsafe = r;
        unsafe.isSafety = false;
        const AccelCommand brake = tick.MakeCommand(unsafe);
        CHECK(std::fabs(brake.ax - std::max(-10.0, std::min(10.0, -0.3 * vel.x() - 0.3 * acc.x()))) < 1e-12);
        CHECK(std::fabs(brake.az - std::max(-15.0, std::min(15.0, -0.3 * vel.z() - 0.3 * acc.z() + 9.8))) < 1e-12);
        std::printf("tick %d rounds %d status %d iters %d cost %.6f u = %.4f %.4f %.4f %.4f\n", t, r.rounds,
                    r.lastStatus, mpc.LastIterations(), mpc.LastCost(), r.u[0], r.u[1], r.u[2], r.u[3]);
        // crude plant: follow the predicted state one control period ahead
        pos = Eigen::Vector3d(r.x0Array[1][0], r.x0Array[1][1], r.x0Array[1][2]);
        vel = Eigen::Vector3d(r.x0Array[1][4], r.x0Array[1][5], r.x0Array[1][6]);
        acc = Eigen::Vector3d(r.x0Array[1][7], r.x0Array[1][8], r.x0Array[1][9]);
    }

    // ---- FrameKDMap with key-frames: slow path of QueryNearest / GetNearestDistance (FrameKDMap.cpp:347-427)
    {
        MapParams mp;
        mp.maxFrameCount = 4;
        FrameKDMap kmap(8192, 1024, mp);
        auto mat = [](double x, double y, double z) { // Twc = Twb(t = xyz, R = I) * Tbc
            Mat4 T = MapParams().Tbc;
            T[3] += x, T[7] += y, T[11] += z;
            return T;
        };
        std::vector<std::shared_ptr<pcl::PointCloud<pcl::PointXYZ>>> cl(4);
        auto none = std::make_shared<pcl::PointCloud<pcl::PointXYZ>>();
        for (int f = 0; f < 4; ++f) {
            cl[f] = std::make_shared<pcl::PointCloud<pcl::PointXYZ>>();
            for (int i = 0; i < 3000; ++i)
                cl[f]->points.emplace_back(ux(rng) + 5.f * f, uy(rng), uz(rng));
            kmap.AddClouds(cl[f], none, mat(-50, 0, 1.5));
            if (f < 3) kmap.InsertKeyFrame();
        }
        CHECK(kmap.KeyFrameCount() == 3 && kmap.PointCount(false) == 3000 && kmap.KeyFramePointCount(2) == 3000);
        // query vector = current (cl[3]) + key-frames 0 and 1; the last key-frame (cl[2]) is left out (:65-75)
        auto brute = [&](const Eigen::Vector3d &q, std::vector<int> frames) {
            std::vector<double> d;
            for (int f : frames)
                for (const auto &p : cl[f]->points) {
                    const double d0 = q.x() - p.x, d1 = q.y() - p.y, d2 = q.z() - p.z;
                    d.push_back((d0 * d0 + d1 * d1) + d2 * d2);
                }
            std::sort(d.begin(), d.end());
            return d;
        };
        std::vector<Eigen::Vector3d> out;
        std::vector<double> dist;
        const Eigen::Vector3d behind(-60, 0.3, 1.2), ahead(12, 0.2, 1.4);
        CHECK(!kmap.PtIsInFrame(behind, mat(-50, 0, 1.5)) && kmap.PtIsInFrame(ahead, mat(-50, 0, 1.5)));
        kmap.QueryNearest(behind, 8, out, dist);
        std::vector<double> want = brute(behind, {3, 0, 1});
        CHECK(out.size() == 8 && dist.size() == 8);
        for (int j = 0; j < 8; ++j) CHECK(dist[j] == want[j]);
        const Eigen::Vector3d side(12, 100, 1.4); // out of the frustum sideways
        kmap.QueryNearest(side, 16, out, dist);
        want = brute(side, {3, 0, 1});
        CHECK(dist.size() == 16);
        for (int j = 0; j < 16; ++j) CHECK(dist[j] == want[j]);
        kmap.QueryNearest(ahead, 8, out, dist); // fast path: current frame only
        want = brute(ahead, {3});
        for (int j = 0; j < 8; ++j) CHECK(dist[j] == want[j]);
        CHECK(kmap.GetNearestDistance(ahead) == std::sqrt(brute(ahead, {3, 0, 1})[0]));
        // QueryNearestMany == QueryNearest site by site, in and out of the frustum mixed (what
        // ProcessWaypoints needs: a waypoint behind the camera must see the key-frames' obstacles)
        {
            const std::vector<Eigen::Vector3d> sites = {behind, ahead, side, Eigen::Vector3d(14, -0.3, 1.6),
                                                        Eigen::Vector3d(-55, 0.1, 1.0)};
            std::vector<std::vector<Eigen::Vector3d>> mo;
            std::vector<std::vector<double>> md;
            kmap.QueryNearestMany(sites, 8, mo, md);
            int n_slow = 0;
            for (size_t i = 0; i < sites.size(); ++i) {
                kmap.QueryNearest(sites[i], 8, out, dist);
                CHECK(md[i].size() == dist.size() && dist.size() == 8);
                for (int j = 0; j < 8; ++j)
                    CHECK(md[i][j] == dist[j] && mo[i][j].x() == out[j].x() && mo[i][j].z() == out[j].z());
                n_slow += !kmap.PtIsInFrame(sites[i], mat(-50, 0, 1.5));
            }
            CHECK(n_slow == 3);
            // the nearest obstacle of `behind` lives in key-frame 0, not in the current frame
            CHECK(md[0][0] == brute(behind, {3, 0, 1})[0] && md[0][0] < brute(behind, {3})[0]);
            CHECK(kmap.GetPtCloud()->points.size() == 9000);
        }
        // ProcessKeyframes: the last key-frame keeps only its points farther than 0.1 m from the
        // current cloud, then the current frame becomes a key-frame (:462-486)
        auto cur = std::make_shared<pcl::PointCloud<pcl::PointXYZ>>();
        for (int i = 0; i < 1500; ++i) cur->points.push_back(cl[2]->points[i]);
        for (int i = 0; i < 1500; ++i) cur->points.emplace_back(ux(rng) + 40.f, uy(rng), uz(rng));
        kmap.AddClouds(cur, none, mat(-50, 0, 1.5));
        int expect = 0;
        for (const auto &p : cl[2]->points) {
            double best = 1e300;
            for (const auto &c : cur->points) {
                const double d0 = (double)p.x - c.x, d1 = (double)p.y - c.y, d2 = (double)p.z - c.z;
                best = std::min(best, (d0 * d0 + d1 * d1) + d2 * d2);
            }
            if (std::sqrt(best) > 0.1) ++expect;
        }
        CHECK(expect >= 10 && expect <= 1500);
        kmap.ProcessKeyframes();
        CHECK(kmap.KeyFrameCount() == 4 && kmap.KeyFramePointCount(2) == expect && kmap.KeyFramePointCount(3) == 3000);
        // a frame the drone has flown past is dropped (DroneBehindPts, :233-252; prune loop :451-458)
        kmap.AddClouds(cur, none, mat(9, 0, 1.5));
        kmap.ProcessKeyframes();
        CHECK(kmap.KeyFrameCount() < 4);
        std::printf("keyframes ok: outliers kept %d, frames now %d\n", expect, kmap.KeyFrameCount());
    }
    std::printf("SHIM_OK\n");
    return 0;
}
