// ROS-free restatement of the TASK branch of AvoidanceStateMachine::Step
// (roswrapper/ros/src/avoid_mpc/src/AvoidanceStateMachine.cpp:322-355) and of the helpers it
// calls: GetInitPath("forward") :24-54, GetCurStateQuad :183-203, ProcessWaypoints :204-235,
// GetRefStates :236-257, PlanWapionts :259-281.  It drives the drop-in classes
// (ObstacleAvoidanceMPC, FrameKDMap) exactly as the ROS node drives the reference's, so a
// control tick can be replayed without ROS.  Publishing (:369-397) is left to the caller.
#ifndef AVOIDANCE_TICK_H
#define AVOIDANCE_TICK_H
#include "FrameKDMap.h"
#include "HighLvlMpc.h"

#include <list>
#include <string>
#include <vector>

struct TickParams { // config/mpc_parameters.yaml
    double T = 1.0, dt = 0.033;
    int maxIter = 3;          // mpc_max_iter
    int nearestPointNum = 3;  // nearest_point_num
    double speed = 10.0, height = 1.5, safetyDistance = 0.2, farestPoint = 500.0, decay = 0.015;
    bool useOdomEstimate = true;
};

// The fields of quadrotor_msgs/Command the planner fills (betaflight_ctrl/quadrotor_msgs/msg/
// Command.msg:1-17): ACCELERATION_MODE = 1, acceleration xyz, yaw.  ROS-free stand-in for the
// message published on /bfctrl/cmd (AvoidanceStateMachine.cpp:369-397).
struct AccelCommand {
    unsigned char mode = 1; // quadrotor_msgs::Command::ACCELERATION_MODE
    double ax = 0, ay = 0, az = 0;
    double yaw = 0;
};

struct TickResult {
    std::vector<double> u;                       // accel x,y,z + yaw rate (PubCmd input)
    std::vector<std::vector<double>> x0Array;    // predicted [X_i, U_i]
    bool isSafety = true;                        // false -> caller publishes the slow-down command
    int rounds = 0;
    int lastStatus = -1;
};

class AvoidanceTick {
public:
    using ObstacleList = std::vector<std::list<Eigen::Vector3d>>;
    AvoidanceTick(const TickParams &p, ObstacleAvoidanceMPC &mpc, FrameKDMap &map);
    void SetOdom(const Eigen::Vector3d &pos, const Eigen::Vector3d &vel, const Eigen::Vector3d &acc,
                 double yaw);
    TickResult Step();
    // PubCmd (:369-378) when the tick is safe, PubSlowDownCmd (:379-397: PD brake on velocity and
    // acceleration plus gravity, clamped to the acceleration limits) otherwise
    AccelCommand MakeCommand(const TickResult &r, double slowDownKp = 0.3, double slowDownKd = 0.3,
                             double aMaxXy = 10.0, double aMaxZ = 15.0) const;
    const std::vector<std::vector<double>> &RefPath() const { return mRefPath; }

private:
    void GetInitPath();
    void GetCurStateQuad(double dt);
    bool ProcessWaypoints(ObstacleList &obstacles);
    std::vector<double> GetRefStates(const ObstacleList &obstacles);
    bool PlanWapionts();
    TickParams mP;
    int mMpcN;
    ObstacleAvoidanceMPC &mMpc;
    FrameKDMap &mMap;
    std::vector<std::vector<double>> mRefPath;
    std::vector<double> mVecStateQuad;
    Eigen::Vector3d mPos, mVel, mAcc;
    double mYaw = 0;
    ObstacleList mVecObstacles;
};
#endif
