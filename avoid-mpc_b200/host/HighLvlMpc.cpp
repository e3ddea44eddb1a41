#include "HighLvlMpc.h"

#include "../../include/ampc.h"

#include <stdexcept>

namespace {
[[noreturn]] void die(ampc_handle *h, const char *what) {
    throw std::runtime_error(std::string("ObstacleAvoidanceMPC: ") + what + ": " + ampc_last_error(h));
}
} // namespace

ObstacleAvoidanceMPC::ObstacleAvoidanceMPC() {}

ObstacleAvoidanceMPC::ObstacleAvoidanceMPC(double T, double dt, std::string /*soPath*/) {
    mT = T;
    mDt = dt;
    mN = T / dt; // same truncation as HighLvlMpc.cpp:9
    mNlpW0.assign(mDimX + (mDimX + mDimU) * mN, 0.0); // zero cold start, HighLvlMpc.cpp:25-27,35-42
    mWeights = {100, 100, 100, 300, 1,  1,  1,  0., 0., 0., 0.0, 10, 10,
                30,  0,   1,   1,   0., 0., 0., 1., 1., 1., 1.,  1.}; // HighLvlMpc.cpp:53-54
    mTau = {0.01, 0.01, 0.01, 0};
    mGains = {1, 1, 1, 1};
}

void ObstacleAvoidanceMPC::SetupWeights(const std::vector<double> &weights) {
    mWeights = weights;
    mParamsDirty = true;
}
void ObstacleAvoidanceMPC::SetupTau(const std::vector<double> &tau) {
    mTau = tau;
    mParamsDirty = true;
}
void ObstacleAvoidanceMPC::SetDroneRadius(const double droneRadius) {
    mDroneRadius = droneRadius;
    mParamsDirty = true;
}
void ObstacleAvoidanceMPC::SetupGains(const std::vector<double> &gains) {
    mGains = gains;
    mParamsDirty = true;
}
void ObstacleAvoidanceMPC::SetDroneAccelLimits(const double aMinZ, const double aMaxZ,
                                               const double aMaxXy, const double aMaxYawDot) {
    mLimits[0] = aMinZ, mLimits[1] = aMaxZ, mLimits[2] = aMaxXy, mLimits[3] = aMaxYawDot;
    mParamsDirty = true;
}

void ObstacleAvoidanceMPC::SetSolverOptions(double tol, int maxIter) {
    if (!(tol > 0) || maxIter < 0)
        throw std::runtime_error("ObstacleAvoidanceMPC::SetSolverOptions: need tol > 0 and maxIter >= 0");
    mTol = tol;
    mMaxIter = maxIter;
    mParamsDirty = true;
}

void ObstacleAvoidanceMPC::EnsureHandle(int K) {
    if (mHandle && K == mK)
        return;
    ampc_config cfg{};
    cfg.N = mN;
    cfg.K = K;
    cfg.dt = mDt;
    cfg.max_batch = 1;
    cfg.max_scenes = 0; // solver only: the map lives in KDTreeTwo / FrameKDMap
    cfg.device = 0;
    ampc_handle *h = nullptr;
    if (ampc_create(&cfg, &h) != AMPC_OK)
        die(nullptr, "ampc_create");
    mHandle.reset(h, ampc_destroy);
    mK = K;
    mParamsDirty = true;
}

void ObstacleAvoidanceMPC::PushParams() {
    if (!mParamsDirty)
        return;
    ampc_handle *h = mHandle.get();
    if (mWeights.size() != 25 || mTau.size() != 4 || mGains.size() != 4)
        throw std::runtime_error("ObstacleAvoidanceMPC: weights/tau/gains must have 25/4/4 entries");
    if (ampc_set_weights(h, mWeights.data()) || ampc_set_tau(h, mTau.data()) ||
        ampc_set_gains(h, mGains.data()) || ampc_set_radius(h, mDroneRadius) ||
        ampc_set_accel_limits(h, mLimits[0], mLimits[1], mLimits[2], mLimits[3]))
        die(h, "setting parameters");
    ampc_solver_opts o;
    ampc_default_solver_opts(&o);
    o.tol = mTol;
    o.max_iter = mMaxIter;
    if (ampc_set_solver_opts(h, &o))
        die(h, "ampc_set_solver_opts");
    mParamsDirty = false;
}

void ObstacleAvoidanceMPC::Solve(const std::vector<double> &vecRefStates, std::vector<double> &u,
                                 std::vector<std::vector<double>> &x0Array, bool /*faster*/) {
    const long rest = (long)vecRefStates.size() - 20 - 10L * mN;
    if (mN < 1 || rest < 0 || rest % (3L * mN) != 0)
        throw std::runtime_error("ObstacleAvoidanceMPC::Solve: vecRefStates has the wrong size");
    EnsureHandle((int)(rest / (3L * mN)));
    PushParams();
    ampc_solve_info info{};
    if (ampc_solve_batch(mHandle.get(), 1, vecRefStates.data(), mNlpW0.data(), &info) != AMPC_OK)
        die(mHandle.get(), "ampc_solve_batch");
    mLastStatus = info.status;
    mLastIters = info.iters;
    mLastCost = info.cost;
    const std::vector<double> &sol_x0 = mNlpW0; // solution in place == next warm start (:129)
    u.clear();
    u.resize(mDimU);
    for (int i = 0; i < mDimU; i++)
        u[i] = sol_x0[i + mDimX];
    x0Array.clear();
    for (size_t i = 0; i < sol_x0.size() - mDimX; i++) {
        if (i % (mDimX + mDimU) == 0)
            x0Array.push_back(std::vector<double>());
        x0Array.back().push_back(sol_x0[i]);
    }
}
