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

// Handle-level constants of the NLP, passed to the solve kernel by value and
// staged into shared memory once per CTA.
struct SolveConsts {
    double Phi[100], Gam[40], gam[10]; // X+ = Phi X + Gam U + gam (RK4x4 of the affine ODE)
    double wgt[25];                    // [Q_goal(10) | Q_pen(10) | Q_u(4) | lambda]
    double radius;
    double lb[4], ub[4];
    double tol, mu_init, bound_push, bound_frac, eps_min, eps_scale, kappa_eps;
    int32_t max_iter, N, K, n_prefix;
};

} // namespace ampc
