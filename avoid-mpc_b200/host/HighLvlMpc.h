// Drop-in replacement for the reference's MPC solver wrapper
// (roswrapper/ros/src/avoid_mpc/include/HighLvlMpc.h:4-34, src/HighLvlMpc.cpp): same
// class name, same public signatures, same argument meaning.  Instead of handing the NLP to
// casadi::nlpsol("ipopt", <codegen .so>) it calls libampc's C-ABI (include/ampc.h), which
// runs the sm_100a interior-point kernel.  No CasADi / IPOPT dependency.
#ifndef HIGH_LVL_MPC_H
#define HIGH_LVL_MPC_H
#include <memory>
#include <string>
#include <vector>

struct ampc_handle;

class ObstacleAvoidanceMPC {
public:
    ObstacleAvoidanceMPC();
    // soPath named the CasADi-generated library in the reference (ROS param `mpc_so`); it is
    // accepted and ignored (N is int(T/dt) as there, HighLvlMpc.cpp:9; K, which the reference
    // baked into the .so, is inferred from the first vecRefStates).
    ObstacleAvoidanceMPC(double T, double dt, std::string soPath);
    // vecRefStates = [x0 | ref N*10 | obst N*K*3 | target] (AvoidanceStateMachine.cpp:236-257).
    // u <- first control, x0Array[i] <- [X_i, U_i] for i < N, warm start kept for the next call
    // (HighLvlMpc.cpp:122-136).  `faster` selected an identically configured second solver in
    // the reference (:50-52) and is ignored.  Throws std::runtime_error if the GPU path fails.
    void Solve(const std::vector<double> &vecRefStates, std::vector<double> &u,
               std::vector<std::vector<double>> &x0Array, bool faster = false);
    void SetupWeights(const std::vector<double> &weights);
    void SetupTau(const std::vector<double> &tau);
    void SetupGains(const std::vector<double> &gains);
    void SetDroneRadius(const double droneRadius);
    void SetDroneAccelLimits(const double aMinZ, const double aMaxZ, const double aMaxXy,
                             const double aMaxYawDot);

    // additions.  The reference hands IPOPT tol = 1e-4 and max_iter = 10 (HighLvlMpc.cpp:17-23) and
    // uses whatever iterate it stops at; this class solves to convergence by default (KKT <= 1e-8,
    // at most 100 iterations).  SetSolverOptions restores a cap for hard real-time use.
    void SetSolverOptions(double tol, int maxIter);
    // (the reference ignores IPOPT's status, HighLvlMpc.cpp:116-122)
    int LastStatus() const { return mLastStatus; }   // AMPC_SOLVE_*
    int LastIterations() const { return mLastIters; }
    double LastCost() const { return mLastCost; }

private:
    void EnsureHandle(int K);
    void PushParams();
    double mT = 0, mDt = 0;
    int mN = 0, mDimX = 10, mDimU = 4, mK = -1;
    double mDroneRadius = 0;
    std::vector<double> mTau, mGains, mWeights, mNlpW0;
    double mLimits[4] = {1., 20., 10., 10.}; // aMinZ, aMaxZ, aMaxXy, aMaxYawDot (HighLvlMpc.cpp:13-16)
    bool mParamsDirty = true;
    double mTol = 1e-8;
    int mMaxIter = 100;
    int mLastStatus = -1, mLastIters = 0;
    double mLastCost = 0;
    std::shared_ptr<ampc_handle> mHandle; // shared by copies (the reference copy-assigns the
                                          // solver once, AvoidanceStateMachine.cpp:62-63)
};
#endif
