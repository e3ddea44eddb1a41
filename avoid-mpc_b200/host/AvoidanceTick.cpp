#include "AvoidanceTick.h"

#include <algorithm>
#include <cmath>

AvoidanceTick::AvoidanceTick(const TickParams &p, ObstacleAvoidanceMPC &mpc, FrameKDMap &map)
    : mP(p), mMpcN(int(p.T / p.dt)), mMpc(mpc), mMap(map), mVecStateQuad(10, 0.0) {
    // InitCircleState, AvoidanceStateMachine.cpp:14-23
    Eigen::Vector3d initPos(0, 0, mP.height), goalPos(3, 0, mP.height);
    Eigen::Vector3d dPos = (goalPos - initPos) * (1.0 / mMpcN);
    for (int i = 0; i < mMpcN; i++) {
        Eigen::Vector3d posi = initPos + dPos * double(i);
        mRefPath.push_back({posi.x(), posi.y(), posi.z(), 0, 0, 0, 0, 0, 0, 0});
    }
}

void AvoidanceTick::SetOdom(const Eigen::Vector3d &pos, const Eigen::Vector3d &vel,
                            const Eigen::Vector3d &acc, double yaw) {
    mPos = pos, mVel = vel, mAcc = acc, mYaw = yaw;
}

void AvoidanceTick::GetCurStateQuad(double dt) { // :183-203
    Eigen::Vector3d pos = mPos, vel = mVel;
    if (mP.useOdomEstimate) {
        pos = mPos + mVel * dt + mAcc * (0.5 * dt * dt);
        vel = mVel + mAcc * dt;
    }
    mVecStateQuad = {pos.x(), pos.y(), pos.z(), mYaw, vel.x(), vel.y(), vel.z(), mAcc.x(), mAcc.y(), mAcc.z()};
}

void AvoidanceTick::GetInitPath() { // :24-54, task == "forward"
    GetCurStateQuad(mP.decay);
    double goalx = std::fmin(mP.speed * mP.T + mPos.x(), mP.farestPoint);
    const double goaly = 0, goalz = mP.height;
    for (int i = 0; i < mMpcN - 1; i++) {
        const std::vector<double> &n = mRefPath[i + 1];
        mRefPath[i] = {n[0], n[1], goalz, n[3], n[4], n[5], n[6], n[7], n[8], n[9]};
    }
    mRefPath[mMpcN - 1] = {goalx, goaly, goalz, 0, mP.speed, 0, 0, 0, 0, 0};
}

bool AvoidanceTick::ProcessWaypoints(ObstacleList &obstacles) { // :204-235
    obstacles.clear();
    obstacles.resize(mMpcN);
    bool needReplan = false;
    std::vector<Eigen::Vector3d> sites;
    for (int i = 0; i < mMpcN; i++)
        sites.emplace_back(mRefPath[i][0], mRefPath[i][1], mRefPath[i][2]);
    std::vector<std::vector<Eigen::Vector3d>> pts;
    std::vector<std::vector<double>> d2;
    // mKeyFrameMap.QueryNearest per waypoint (:214): sites outside the current frustum (typically
    // waypoint 0, behind the camera) fall back to the key-frames; batched, at most two launches
    mMap.QueryNearestMany(sites, mP.nearestPointNum, pts, d2);
    for (int i = 0; i < mMpcN; i++) {
        for (int j = 0; j < mP.nearestPointNum; j++) {
            if (j < (int)pts[i].size())
                obstacles[i].emplace_back(pts[i][j].x(), pts[i][j].y(), pts[i][j].z());
            else
                obstacles[i].emplace_back(10000, 10000, 10000);
        }
        if (d2[i].empty() || std::sqrt(d2[i][0]) <= mP.safetyDistance)
            needReplan = true;
    }
    return needReplan;
}

std::vector<double> AvoidanceTick::GetRefStates(const ObstacleList &obstacles) { // :236-257
    std::vector<double> v = mVecStateQuad;
    for (int i = 0; i < mMpcN; i++)
        v.insert(v.end(), mRefPath[i].begin(), mRefPath[i].end());
    for (int i = 0; i < mMpcN; i++)
        for (const Eigen::Vector3d &o : obstacles[i]) {
            v.push_back(o.x());
            v.push_back(o.y());
            v.push_back(o.z());
        }
    std::vector<double> tgt = mRefPath.back();
    double dX = mP.speed * mP.T - std::max(0., tgt[0] - mPos.x());
    dX = std::max(0., dX);
    tgt[0] += dX;
    tgt[1] = 0.;
    v.insert(v.end(), tgt.begin(), tgt.end());
    return v;
}

bool AvoidanceTick::PlanWapionts() { // :259-281 (the loop bound is literally 1 in the reference)
    bool isSafety = true;
    for (int ptIndex = 0; ptIndex < 1; ptIndex++) {
        Eigen::Vector3d p1(mRefPath[ptIndex][0], mRefPath[ptIndex][1], mRefPath[ptIndex][2]);
        if (mMap.GetNearestDistance(p1) > mP.safetyDistance)
            continue;
        std::vector<Eigen::Vector3d> edgePts;
        std::vector<double> distances;
        mMap.QueryNearest(p1, 1, edgePts, distances, true);
        if (edgePts.empty()) {
            isSafety = false;
            continue;
        }
        mRefPath[ptIndex][0] = edgePts[0].x();
        mRefPath[ptIndex][1] = edgePts[0].y();
        mRefPath[ptIndex][2] = edgePts[0].z();
        isSafety = true;
    }
    return isSafety;
}

TickResult AvoidanceTick::Step() { // case TASK, :322-355
    TickResult r;
    GetInitPath();
    bool isSafety = true;
    const double decay = mP.decay; // the reference re-measures wall time here (:329,343)
    for (int iter = 0; iter < mP.maxIter; iter++) {
        GetCurStateQuad(decay);
        isSafety = PlanWapionts();
        const bool needReplan = ProcessWaypoints(mVecObstacles);
        if (!needReplan && iter > 0 && isSafety)
            break;
        std::vector<double> vecRefStates = GetRefStates(mVecObstacles);
        mMpc.Solve(vecRefStates, r.u, r.x0Array, iter == 0);
        for (int i = 0; i < mMpcN; i++)
            mRefPath[i].assign(r.x0Array[i].begin(), r.x0Array[i].begin() + 10);
        r.rounds++;
    }
    r.isSafety = isSafety;
    r.lastStatus = mMpc.LastStatus();
    return r;
}

AccelCommand AvoidanceTick::MakeCommand(const TickResult &r, double kp, double kd, double aMaxXy,
                                        double aMaxZ) const {
    AccelCommand cmd;
    if (r.isSafety && r.u.size() >= 3) { // PubCmd
        cmd.ax = r.u[0], cmd.ay = r.u[1], cmd.az = r.u[2];
    } else { // PubSlowDownCmd
        const Eigen::Vector3d accSlow = mVel * (-kp) - mAcc * kd + Eigen::Vector3d(0, 0, 9.8);
        cmd.ax = std::max(-aMaxXy, std::min(aMaxXy, accSlow.x()));
        cmd.ay = std::max(-aMaxXy, std::min(aMaxXy, accSlow.y()));
        cmd.az = std::max(-aMaxZ, std::min(aMaxZ, accSlow.z()));
    }
    cmd.yaw = 0;
    return cmd;
}
