"""Shipped parameter values of the reference (they parameterise the NLP).

Source: roswrapper/ros/src/avoid_mpc/config/mpc_parameters.yaml (line numbers below)
and the weight order of ParameterManager.cpp:63-68 == tools/mpc_obstacle_casadi.py:370-397.
"""
import numpy as np

# yaml:7-34  goal_{p_x,p_y,p_z,yaw,v_x,v_y,v_z,a_x,a_y,a_z}, path_{...}, u_{a_x,a_y,a_z,yaw_dot}, collide_lambda
WEIGHTS = np.array(
    [50.0, 50.0, 100.0, 100.0, 1.0, 1.0, 1.0, 0.0, 0.0, 0.0,
     0.0, 10.0, 50.0, 100.0, 0.0, 1.0, 1.0, 0.0, 1.0, 1.0,
     0.3, 0.3, 0.5, 1.0,
     1.2], dtype=np.float64)
TAU = np.array([6.09837416, 6.21675029, 15.79816293, 0.0])  # yaml:36-39
GAINS = np.array([0.999999, 0.999999, 0.999999, 1.0])       # yaml:41-44
SPEED = 10.0            # yaml:46
DRONE_RADIUS = 0.5      # yaml:47
A_MIN_Z, A_MAX_Z, A_MAX_XY, A_MAX_YAW_DOT = 5.0, 15.0, 10.0, 10.0  # yaml:49-52
HEIGHT = 1.5            # yaml:54
SAFETY_DISTANCE = 0.2   # yaml:56
T_B_C = np.array([[0.0, 0.0, 1.0, 0.05],
                  [-1.0, 0.0, 0.0, 0.0],
                  [0.0, -1.0, 0.0, 0.01],
                  [0.0, 0.0, 0.0, 1.0]])  # yaml:67-71
MPC_MAX_ITER = 3        # yaml:3 (outer k-NN <-> solve rounds per tick)

# BASELINE.json configs: T = 1.0, dt = 0.05  ->  N = int(T/dt) = 20 (HighLvlMpc.cpp:9)
BENCH_T, BENCH_DT = 1.0, 0.05


def u_bounds(a_min_z=A_MIN_Z, a_max_z=A_MAX_Z, a_max_xy=A_MAX_XY, a_max_yaw_dot=A_MAX_YAW_DOT):
    """HighLvlMpc.cpp:70-76: uMin/uMax of SetDroneAccelLimits."""
    lb = np.array([-a_max_xy, -a_max_xy, a_min_z, -a_max_yaw_dot])
    ub = np.array([a_max_xy, a_max_xy, a_max_z, a_max_yaw_dot])
    return lb, ub
