// Shared device/host helpers for libampc (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <math.h>
#include <stdint.h>

#define AMPC_FULL_MASK 0xffffffffu
#define AMPC_NX 10
#define AMPC_NU 4

namespace ampc {

__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1)
        v += __shfl_xor_sync(AMPC_FULL_MASK, v, o);
    return v;
}
__device__ __forceinline__ double warp_max(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1)
        v = fmax(v, __shfl_xor_sync(AMPC_FULL_MASK, v, o));
    return v;
}
__device__ __forceinline__ double warp_min(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1)
        v = fmin(v, __shfl_xor_sync(AMPC_FULL_MASK, v, o));
    return v;
}

// The affine dynamics (tools/mpc_obstacle_casadi.py:106-122) decouple into four chains: axis
// i < 3 with components (p_i, v_i, a_i) = state indices (i, 4+i, 7+i) driven by control i, and
// yaw (component 0 = state 3, control 3; components 1, 2 are padding that stays exactly zero).
// Chain i advances by the upper-triangular 3x3 matrix F_i and the 3-vector G_i:
struct Chain {
    double d1, c1, c2, d2, c3, c4; // F = [[d1,c1,c2],[0,d2,c3],[0,0,c4]]
    double g1, g2, g3;             // G
};

// Handle-level constants of the NLP, passed to the solve kernels by value.
struct SolveConsts {
    double Phi[100], Gam[40], gam[10]; // X+ = Phi X + Gam U + gam (RK4x4 of the affine ODE)
    double wgt[25];                    // [Q_goal(10) | Q_pen(10) | Q_u(4) | lambda]
    double radius;
    double lb[4], ub[4];
    double tol, mu_init, bound_push, bound_frac, eps_min, eps_scale, kappa_eps;
    int32_t max_iter, N, K, n_prefix;
    Chain ch[4];                       // the same Phi / Gam, chain by chain
};

} // namespace ampc
