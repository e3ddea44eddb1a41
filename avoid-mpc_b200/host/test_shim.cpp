// Exercises the drop-in C++ classes end to end on the GPU (run by tests/test_gpu_shim.py).
#include "AvoidanceTick.h"
#include "kd_tree_two.h"

#include <algorithm>
#include <cmath>
#include <cstdio>
#include <random>

#define CHECK(c)                                                                                   \
    do {                                                                                           \
        if (!(c)) {                                                                                \
            std::printf("CHECK failed %s:%d: %s\n", __FILE__, __LINE__, #c);                       \
            return 1;                                                                              \
        }                                                                                          \
    } while (0)

int main() {
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
    std::printf("SHIM_OK\n");
    return 0;
}
