// libampc.so — C-ABI (include/ampc.h) over the sm_100a kernels.
// No torch types, no CPU fallback: every compute entry point launches CUDA
// kernels and returns AMPC_ERR_CUDA (with text in ampc_last_error) if it cannot.
#include "../../include/ampc.h"

#include "common.cuh"
#include "ipm_solve.cuh"
#include "knn_tiles.cuh"
#include "depth_cloud.cuh"

#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <mutex>
#include <new>
#include <string>
#include <vector>

using namespace ampc;

static_assert(sizeof(ampc_solve_info) == sizeof(SolveOut), "ampc_solve_info layout");

namespace {

std::string g_create_error;
std::mutex g_create_mutex;

struct DevBuf {
    void *p = nullptr;
    size_t bytes = 0;
    cudaError_t reserve(size_t n) {
        if (n <= bytes)
            return cudaSuccess;
        if (p)
            cudaFree(p);
        p = nullptr;
        bytes = 0;
        cudaError_t e = cudaMalloc(&p, n);
        if (e == cudaSuccess)
            bytes = n;
        return e;
    }
    void release() {
        if (p)
            cudaFree(p);
        p = nullptr;
        bytes = 0;
    }
    template <typename T> T *as() const { return static_cast<T *>(p); }
};

} // namespace

struct ampc_handle {
    ampc_config cfg{};
    ampc_solver_opts opts{};
    double weights[25], tau[4], gains[4], radius;
    double lb[4], ub[4];
    SolveConsts consts{};
    bool consts_dirty = true;
    int n_w = 0, n_prefix = 0;
    cudaStream_t stream = nullptr;
    // clouds: [kind] slot buffers + counts
    DevBuf cloud[2], counts[2], boxes[2], gboxes[2], nan_flags[2];
    int slot_groups[2] = {0, 0};
    bool knn_two_level = true; // AMPC_KNN_TWO_LEVEL=0: the one-level search of round 1
    DevBuf sorted[2]; // Morton-bucketed copies for layout -1 (allocated on first use)
    bool sort_smem_set = false;
    int slot_points[2] = {0, 0};
    int slot_tiles[2] = {0, 0};
    int row_w[2] = {0, 0}; // layout hint per kind for the next cloud_set: 0 unorganised, > 0 row pitch, -1 Morton-bucketed
    DevBuf layout[2];      // per scene: the row pitch its tiles were built with
    DevBuf raw_stage; // staging for stride != 16 uploads
    // depth image -> clouds
    ampc_camera cam{320, 320, 320, 240, 10, 1, 0.1, 100}; // config/mpc_parameters.yaml:58-66
    DevBuf depth_stage, depth_T, depth_tab, depth_scratch, depth_flag;
    int tab_rows = 0, tab_cols = 0;
    double tab_scale = 0;
    // batch workspaces
    DevBuf queries, prefix, w, info, knn_idx, knn_d2, knn_cnt, knn_pts, scene_of, x0, ref, posx, replan;
    DevBuf ws_d, ws_i, ws_counter;
    DevBuf bo_arg, bo_cost;
    DevBuf gs_sites, gs_nb, gs_epts, gs_ecnt, gs_q0; // Edge-tree guess round (ampc_guess_round_batch)
    DevBuf tk_active, tk_q0, tk_d1, tk_c1, tk_ep, tk_ec, tk_ed, tk_safe, tk_rounds; // tick loop
    int64_t launches = 0;
    std::string err;
    // quad solve kernel: workspace of the resident warps, refill counter, tuning knobs
    DevBuf quad_ws, quad_counter, quad_order_buf;
    bool quad_order = true; // AMPC_QUAD_ORDER=0: take the instances in index order
    int n_sm = 148;
    int smem_per_sm = 228 * 1024;
    int quad_warps_per_sm = 8; // AMPC_QUAD_WARPS_PER_SM (255 registers per thread allow 8)
    int quad_per_warp = 0;     // AMPC_QUADS_PER_WARP (1, 2, 4, 8): 0 = by batch size
    int quad_cta_warps = 2;    // AMPC_QUAD_CTA_WARPS: warps per CTA of the quad kernel (they meet once per pass; measured
                               // at 32768 instances, solve-only TFLOP/s: 1: 2.47, 2: 2.59, 3: 2.53, 6: 2.36)
    int solve_kernel = 0;      // AMPC_SOLVE_KERNEL: 0 auto (by batch size), 1 warp, 2 quad
    int quad_min_batch = 8192; // AMPC_QUAD_MIN_BATCH: smallest batch the quad kernel takes in auto mode
    int solve_smem_set[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0};
    int solve_warps = 2;
    int solve_smem_pad = 0; // AMPC_SOLVE_SMEM_PAD: extra dynamic shared memory per CTA (occupancy experiments)
    int knn_smem_set = 0;
    bool index_smem_set = false;
    // optional per-kernel timing of ampc_round_batch_dev (CUDA events on the caller's stream)
    bool prof = false;
    std::vector<cudaEvent_t> prof_ev; // pairs (start, stop), tagged with a section
    std::vector<int> prof_sec;
    int prof_pending = 0;
    double prof_ms[3] = {0, 0, 0};   // index build, k-NN search, solve
    int64_t prof_n[3] = {0, 0, 0};
};

namespace {

int fail(ampc_handle *h, int code, const std::string &msg) {
    if (h)
        h->err = msg;
    return code;
}
int cuda_fail(ampc_handle *h, cudaError_t e, const char *what) {
    return fail(h, AMPC_ERR_CUDA, std::string(what) + ": " + cudaGetErrorString(e));
}
#define CK(call)                                                                                   \
    do {                                                                                           \
        cudaError_t e__ = (call);                                                                  \
        if (e__ != cudaSuccess)                                                                    \
            return cuda_fail(h, e__, #call);                                                       \
    } while (0)

// ---- discrete dynamics: 4 RK4 sub-steps of dt/4 of the affine ODE
// (tools/mpc_obstacle_casadi.py:106-122,338-357); F is affine, so Phi/Gam/gam are
// read off F applied to unit vectors (the RK4 polynomial, not expm).
void ode_rhs(const double *x, const double *u, const double *tau, double *xd) {
    xd[0] = x[4], xd[1] = x[5], xd[2] = x[6];
    xd[3] = u[3];
    xd[4] = x[7], xd[5] = x[8], xd[6] = x[9];
    xd[7] = (u[0] - x[7]) * tau[0];
    xd[8] = (u[1] - x[8]) * tau[1];
    xd[9] = (u[2] - 9.81 - x[9]) * tau[2];
}
void rk4_map(const double *x0, const double *u, const double *tau, double dt, double *xn) {
    const double h = dt / 4;
    double X[10], k1[10], k2[10], k3[10], k4[10], t[10];
    std::memcpy(X, x0, sizeof X);
    for (int m = 0; m < 4; ++m) {
        ode_rhs(X, u, tau, k1);
        for (int i = 0; i < 10; ++i) t[i] = X[i] + 0.5 * (k1[i] *= h);
        ode_rhs(t, u, tau, k2);
        for (int i = 0; i < 10; ++i) t[i] = X[i] + 0.5 * (k2[i] *= h);
        ode_rhs(t, u, tau, k3);
        for (int i = 0; i < 10; ++i) t[i] = X[i] + (k3[i] *= h);
        ode_rhs(t, u, tau, k4);
        for (int i = 0; i < 10; ++i) X[i] = X[i] + (k1[i] + 2 * k2[i] + 2 * k3[i] + (k4[i] *= h)) / 6;
    }
    std::memcpy(xn, X, sizeof X);
}
void refresh_consts(ampc_handle *h) {
    if (!h->consts_dirty)
        return;
    SolveConsts &c = h->consts;
    double z[10] = {0}, zu[4] = {0}, col[10];
    rk4_map(z, zu, h->tau, h->cfg.dt, c.gam);
    for (int j = 0; j < 10; ++j) {
        double e[10] = {0};
        e[j] = 1.0;
        rk4_map(e, zu, h->tau, h->cfg.dt, col);
        for (int i = 0; i < 10; ++i) c.Phi[i * 10 + j] = col[i] - c.gam[i];
    }
    for (int j = 0; j < 4; ++j) {
        double e[4] = {0};
        e[j] = 1.0;
        rk4_map(z, e, h->tau, h->cfg.dt, col);
        for (int i = 0; i < 10; ++i) c.Gam[i * 4 + j] = col[i] - c.gam[i];
    }
    std::memcpy(c.wgt, h->weights, sizeof c.wgt);
    c.radius = h->radius;
    std::memcpy(c.lb, h->lb, sizeof c.lb);
    std::memcpy(c.ub, h->ub, sizeof c.ub);
    c.tol = h->opts.tol;
    c.mu_init = h->opts.mu_init;
    c.bound_push = h->opts.bound_push;
    c.bound_frac = h->opts.bound_frac;
    c.eps_min = h->opts.eps_min;
    c.eps_scale = h->opts.eps_scale;
    c.kappa_eps = h->opts.kappa_eps;
    c.max_iter = h->opts.max_iter;
    c.N = h->cfg.N;
    c.K = h->cfg.K;
    c.n_prefix = h->n_prefix;
    for (int i = 0; i < 4; ++i) {
        Chain &f = c.ch[i];
        if (i < 3) {
            f.d1 = c.Phi[i * 10 + i], f.c1 = c.Phi[i * 10 + 4 + i], f.c2 = c.Phi[i * 10 + 7 + i];
            f.d2 = c.Phi[(4 + i) * 10 + 4 + i], f.c3 = c.Phi[(4 + i) * 10 + 7 + i];
            f.c4 = c.Phi[(7 + i) * 10 + 7 + i];
            f.g1 = c.Gam[i * 4 + i], f.g2 = c.Gam[(4 + i) * 4 + i], f.g3 = c.Gam[(7 + i) * 4 + i];
        } else {
            f.d1 = c.Phi[33], f.c1 = f.c2 = f.d2 = f.c3 = f.c4 = 0.0;
            f.g1 = c.Gam[15], f.g2 = f.g3 = 0.0;
        }
    }
    h->consts_dirty = false;
}

// ---- small kernels around the two hot ones --------------------------------

// stride != 16 uploads: repack raw records into float4 slots
__global__ void repack_kernel(const unsigned char *raw, int64_t scene_stride, int stride,
                              float4 *clouds, const int32_t *counts, int64_t slot_points,
                              int first_scene) {
    const int scene = first_scene + blockIdx.x; // scenes on x: no 65535 limit
    const int n = counts[scene];
    for (int i = blockIdx.y * blockDim.x + threadIdx.x; i < n; i += gridDim.y * blockDim.x) {
        const float *p = reinterpret_cast<const float *>(raw + (int64_t)blockIdx.x * scene_stride +
                                                         (int64_t)i * stride);
        clouds[(int64_t)scene * slot_points + i] = make_float4(p[0], p[1], p[2], 1.0f);
    }
}

// GetRefStates (src/AvoidanceStateMachine.cpp:236-257) minus the obstacle block
// (written by the k-NN kernel), plus the Q = N query sites of ProcessWaypoints
// (:211-215).  One thread per (instance, element).
__global__ void pack_prefix_kernel(int B, int N, int K, const double *x0, const double *ref,
                                   const double *pos_x, double speed, double T, double *prefix,
                                   int n_prefix, double *queries) {
    const int b = blockIdx.x;
    if (b >= B)
        return;
    double *p = prefix + (int64_t)b * n_prefix;
    const double *r = ref + (int64_t)b * N * 10;
    for (int i = threadIdx.x; i < 10; i += blockDim.x)
        p[i] = x0[(int64_t)b * 10 + i];
    for (int i = threadIdx.x; i < 10 * N; i += blockDim.x)
        p[10 + i] = r[i];
    for (int i = threadIdx.x; i < 3 * N; i += blockDim.x) {
        const int k = i / 3, c = i - 3 * k;
        queries[(int64_t)b * 3 * N + i] = r[10 * k + c];
    }
    if (threadIdx.x < 10) { // target rule, :250-255
        const int i = threadIdx.x;
        double v = r[10 * (N - 1) + i];
        if (i == 0) {
            const double px = pos_x ? pos_x[b] : x0[(int64_t)b * 10];
            double dX = speed * T - fmax(0.0, v - px);
            dX = fmax(0.0, dX);
            v += dX;
        }
        if (i == 1)
            v = 0.0;
        p[10 + 10 * N + 3 * K * N + i] = v;
    }
}

// needReplan of ProcessWaypoints (:228-231)
__global__ void replan_kernel(int B, int Q, int k, const double *dist2, const int32_t *count,
                              double safety, int32_t *need_replan) {
    const int b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= B)
        return;
    int flag = 0;
    for (int q = 0; q < Q; ++q) {
        const int c = count[(int64_t)b * Q + q];
        if (c == 0 || sqrt(dist2[((int64_t)b * Q + q) * k]) <= safety)
            flag = 1;
    }
    need_replan[b] = flag;
}

// ---- Edge-tree initial guesses (BASELINE config C2; PlanWapionts generalised, :259-281) ------
// Query sites of scene s: the N-1 waypoints 1..N-1 shared by all guesses, then the G guess
// positions for waypoint 0.  Guess g = g-th nearest Edge point of waypoint 0; beyond the number of
// Edge points the waypoint stays where it is.
__global__ void guess_sites_kernel(int n_scenes, int N, int G, const double *ref, const double *epts,
                                   const int32_t *ecnt, double *sites) {
    const int s = blockIdx.x;
    const int Qp = N - 1 + G;
    const double *r = ref + (int64_t)s * N * 10;
    for (int i = threadIdx.x; i < Qp; i += blockDim.x) {
        double *o = sites + ((int64_t)s * Qp + i) * 3;
        if (i < N - 1) {
            o[0] = r[10 * (i + 1)], o[1] = r[10 * (i + 1) + 1], o[2] = r[10 * (i + 1) + 2];
        } else {
            const int g = i - (N - 1);
            const double *e = g < ecnt[s] ? epts + ((int64_t)s * G + g) * 3 : r;
            o[0] = e[0], o[1] = e[1], o[2] = e[2];
        }
    }
}
// prefix of instance b = s*G + g: [x0_s | ref_s with waypoint 0 at guess g | obst | target]; the
// neighbour block of stage 0 comes from the guess's own query, stages 1.. from the shared ones
__global__ void guess_pack_kernel(int n_scenes, int N, int K, int G, const double *x0, const double *ref,
                                  const double *pos_x, double speed, double T, const double *sites,
                                  const double *nb /* [s][N-1+G][K][3] */, double *prefix, int n_prefix) {
    const int b = blockIdx.x, s = b / G, g = b - s * G;
    const int Qp = N - 1 + G;
    double *p = prefix + (int64_t)b * n_prefix;
    const double *r = ref + (int64_t)s * N * 10;
    const double *site0 = sites + ((int64_t)s * Qp + (N - 1 + g)) * 3;
    for (int i = threadIdx.x; i < 10; i += blockDim.x)
        p[i] = x0[(int64_t)s * 10 + i];
    for (int i = threadIdx.x; i < 10 * N; i += blockDim.x)
        p[10 + i] = i < 3 ? site0[i] : r[i];
    const double *nbs = nb + (int64_t)s * Qp * K * 3;
    for (int i = threadIdx.x; i < 3 * K * N; i += blockDim.x) {
        const int k = i / (3 * K), e = i - k * 3 * K;
        const int q = k == 0 ? N - 1 + g : k - 1;
        p[10 + 10 * N + i] = nbs[(int64_t)q * K * 3 + e];
    }
    if (threadIdx.x < 10) { // target rule, :250-255
        const int i = threadIdx.x;
        double v = r[10 * (N - 1) + i];
        if (N == 1 && i < 3)
            v = site0[i];
        if (i == 0) {
            const double px = pos_x ? pos_x[s] : x0[(int64_t)s * 10];
            double dX = speed * T - fmax(0.0, v - px);
            dX = fmax(0.0, dX);
            v += dX;
        }
        if (i == 1)
            v = 0.0;
        p[10 + 10 * N + 3 * K * N + i] = v;
    }
}

// ---- FP64 FMA peak of the device, measured: 8 independent dependent-FMA chains per thread,
// enough resident warps to cover the pipe latency (the denominator of the solve kernels' roofline)
__global__ void __launch_bounds__(256) fp64_fma_peak_kernel(double *out, int iters, double a, double b) {
    double x0 = threadIdx.x, x1 = x0 + 1, x2 = x0 + 2, x3 = x0 + 3, x4 = x0 + 4, x5 = x0 + 5, x6 = x0 + 6, x7 = x0 + 7;
#pragma unroll 1
    for (int i = 0; i < iters; ++i) {
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            x0 = fma(x0, a, b), x1 = fma(x1, a, b), x2 = fma(x2, a, b), x3 = fma(x3, a, b);
            x4 = fma(x4, a, b), x5 = fma(x5, a, b), x6 = fma(x6, a, b), x7 = fma(x7, a, b);
        }
    }
    out[(int64_t)blockIdx.x * blockDim.x + threadIdx.x] = ((x0 + x1) + (x2 + x3)) + ((x4 + x5) + (x6 + x7));
}

__global__ void best_of_kernel(int n_scenes, int G, const SolveOut *info, int32_t *argmin,
                               double *best) {
    const int s = blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= n_scenes)
        return;
    int a = -1;
    double c = INFINITY;
    for (int g = 0; g < G; ++g) {
        const SolveOut &o = info[(int64_t)s * G + g];
        if ((o.status == AMPC_SOLVE_CONVERGED || o.status == AMPC_SOLVE_MAX_ITER) && o.cost < c) {
            c = o.cost;
            a = g;
        }
    }
    argmin[s] = a;
    best[s] = c;
}

constexpr int PROF_RING = 192;
enum { SEC_INDEX = 0, SEC_KNN = 1, SEC_SOLVE = 2 };
int prof_drain(ampc_handle *h) {
    for (int i = 0; i < h->prof_pending; ++i) {
        float a = 0;
        CK(cudaEventSynchronize(h->prof_ev[2 * i + 1]));
        CK(cudaEventElapsedTime(&a, h->prof_ev[2 * i], h->prof_ev[2 * i + 1]));
        h->prof_ms[h->prof_sec[i]] += a;
        h->prof_n[h->prof_sec[i]]++;
    }
    h->prof_pending = 0;
    return AMPC_OK;
}
// returns the slot to close with prof_end, or -1 when profiling is off
int prof_begin(ampc_handle *h, int sec, cudaStream_t st, int *slot) {
    *slot = -1;
    if (!h->prof) return AMPC_OK;
    if (h->prof_pending == PROF_RING) {
        int rc = prof_drain(h);
        if (rc) return rc;
    }
    *slot = h->prof_pending++;
    h->prof_sec[*slot] = sec;
    CK(cudaEventRecord(h->prof_ev[2 * *slot], st));
    return AMPC_OK;
}
int prof_end(ampc_handle *h, int slot, cudaStream_t st) {
    if (slot < 0) return AMPC_OK;
    CK(cudaEventRecord(h->prof_ev[2 * slot + 1], st));
    return AMPC_OK;
}

// ---- control-tick loop on the device (AvoidanceStateMachine.cpp:328-344) --------------
// PlanWapionts (:259-281) for waypoint 0 of every still-active instance: if an obstacle is
// within safety_distance, move the waypoint to the nearest Edge point; no Edge point ->
// isSafety = false.  d1/cnt1: 1-NN on the Obstacle cloud, e*: 1-NN on the Edge cloud.
__global__ void tick_plan_kernel(int B, int N, const int32_t *active, const double *d1, const int32_t *cnt1,
                                 const double *epts, const int32_t *ecnt, bool have_edge, double safety,
                                 double *ref, int32_t *is_safety) {
    const int b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= B || active[b] == 0)
        return;
    int safe = 1;
    // GetNearestDistance returns sqrt(min dist2), or DBL_MAX when nothing was found (:400-427)
    const double nearest = cnt1[b] > 0 ? sqrt(d1[b]) : 1.7976931348623157e308;
    if (nearest <= safety) {
        if (have_edge && ecnt[b] > 0) {
            double *r0 = ref + (int64_t)b * N * 10;
            r0[0] = epts[3 * b], r0[1] = epts[3 * b + 1], r0[2] = epts[3 * b + 2];
        } else {
            safe = 0;
        }
    }
    is_safety[b] = safe;
}
// first waypoint of every instance as a 1-query site
__global__ void tick_site0_kernel(int B, int N, const double *ref, double *q0) {
    const int b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= B)
        return;
    const double *r0 = ref + (int64_t)b * N * 10;
    q0[3 * b] = r0[0], q0[3 * b + 1] = r0[1], q0[3 * b + 2] = r0[2];
}
// `if (!needReplan && iter > 0 && isSafety) break;` (:333) as a per-instance mask update
__global__ void tick_gate_kernel(int B, int iter, const int32_t *need_replan, const int32_t *is_safety,
                                 int32_t *active, int32_t *rounds) {
    const int b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= B || active[b] == 0)
        return;
    if (!need_replan[b] && iter > 0 && is_safety[b])
        active[b] = 0;
    else
        rounds[b] += 1;
}
// mRefPath[i] := x0Array[i][0:10] for i < N (:338-342) of the instances that just solved
__global__ void tick_ref_update_kernel(int B, int N, const int32_t *active, const double *w, double *ref) {
    const int b = blockIdx.x;
    if (b >= B || active[b] == 0)
        return;
    for (int e = threadIdx.x; e < 10 * N; e += blockDim.x) {
        const int i = e / 10, c = e - 10 * i;
        ref[(int64_t)b * N * 10 + e] = w[(int64_t)b * (10 + 14 * N) + 14 * i + c];
    }
}
__global__ void tick_init_kernel(int B, int32_t *active, int32_t *rounds, int32_t *is_safety) {
    const int b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b < B) {
        active[b] = 1;
        rounds[b] = 0;
        is_safety[b] = 1;
    }
}
// replan flags of the active instances only (inactive keep their last value)
__global__ void tick_replan_kernel(int B, int Q, int k, const int32_t *active, const double *dist2,
                                   const int32_t *count, double safety, int32_t *need_replan) {
    const int b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= B || active[b] == 0)
        return;
    int flag = 0;
    for (int q = 0; q < Q; ++q) {
        const int c = count[(int64_t)b * Q + q];
        if (c == 0 || sqrt(dist2[((int64_t)b * Q + q) * k]) <= safety)
            flag = 1;
    }
    need_replan[b] = flag;
}

int check_kind(ampc_handle *h, int kind) {
    if (kind != AMPC_CLOUD_OBSTACLE && kind != AMPC_CLOUD_EDGE)
        return fail(h, AMPC_ERR_INVALID, "kind must be AMPC_CLOUD_OBSTACLE or AMPC_CLOUD_EDGE");
    if (h->slot_points[kind] <= 0)
        return fail(h, AMPC_ERR_CAPACITY, "handle was created without capacity for this cloud kind");
    return AMPC_OK;
}

int launch_knn(ampc_handle *h, int kind, int B, const int32_t *scene_of_dev, const double *q_dev,
               int Q, int k, int32_t *idx, double *d2, int32_t *cnt, double *pts, int64_t pts_is,
               int64_t pts_qs, cudaStream_t st, const int32_t *active = nullptr) {
    if (k < 1 || k > KNN_KMAX)
        return fail(h, AMPC_ERR_UNSUPPORTED, "k must be in 1..32");
    if (Q < 1 || B < 1)
        return fail(h, AMPC_ERR_INVALID, "B and Q must be positive");
    // one warp per (instance, query); small batches split each cloud's tiles over `segs` warps
    int segs = 1;
    const int64_t warps = (int64_t)B * Q;
    const int target_warps = 148 * 16;
    if (warps < target_warps) {
        const int max_useful = (h->slot_points[kind] / KT_TILE + 63) / 64; // >= 64 tiles (4 groups) per segment
        segs = (int)((target_warps + warps - 1) / warps);
        if (segs > max_useful) segs = max_useful;
        if (segs > 64) segs = 64;
        if (segs < 1) segs = 1;
    }
    KnnParams P{};
    P.clouds = h->cloud[kind].as<float4>();
    P.sorted = h->sorted[kind].as<float4>();
    P.boxes = h->boxes[kind].as<float4>();
    P.gboxes = h->gboxes[kind].as<float4>();
    P.slot_groups = h->slot_groups[kind];
    P.counts = h->counts[kind].as<int32_t>();
    P.slot_points = h->slot_points[kind];
    P.slot_tiles = h->slot_tiles[kind];
    P.layout = h->layout[kind].as<int32_t>();
    P.scene_of = scene_of_dev;
    P.active = active;
    P.queries = q_dev;
    P.Q = Q;
    P.k = k;
    P.segs = segs;
    P.idx = idx;
    P.dist2 = d2;
    P.count = cnt;
    P.pts = pts;
    P.pts_inst_stride = pts_is;
    P.pts_query_stride = pts_qs;
    if (segs > 1) {
        const size_t n = (size_t)B * Q * segs * k;
        CK(h->ws_d.reserve(n * 8));
        CK(h->ws_i.reserve(n * 4));
        P.ws_d = h->ws_d.as<double>();
        P.ws_i = h->ws_i.as<uint32_t>();
    }
    const dim3 grid(B, (Q + KS_WARPS - 1) / KS_WARPS, segs);
    if (grid.y > 65535u)
        return fail(h, AMPC_ERR_UNSUPPORTED, "too many queries per instance");
    if (h->knn_two_level) {
        // group lower bounds per warp: what a slot of this handle can need, at most KS_GCHUNK
        int gch = (int)((h->slot_groups[kind] + 31) / 32 * 32);
        if (gch > KS_GCHUNK) gch = KS_GCHUNK;
        if (gch < 32) gch = 32;
        P.gchunk = gch;
        knn_search2_kernel<<<grid, KS_WARPS * 32, (size_t)KS_WARPS * gch * sizeof(float), st>>>(P);
    } else
        knn_search_kernel<<<grid, KS_WARPS * 32, 0, st>>>(P);
    h->launches++;
    CK(cudaGetLastError());
    if (segs > 1) {
        knn_merge_kernel<<<dim3(B, grid.y), KS_WARPS * 32, 0, st>>>(P);
        h->launches++;
        CK(cudaGetLastError());
    }
    return AMPC_OK;
}

// sort_mode: the scenes were given layout -1 -> after the NaN filter, bucket a copy of each
// cloud by Morton cell and build the tile boxes over the copy
// row_hint: a row pitch some of these scenes may carry in their layout word besides the handle's
// own hint (the depth kernels record the image width themselves); only sizes the group pass
int launch_index(ampc_handle *h, int kind, int first_scene, int n_scenes, cudaStream_t st, bool sort_mode = false,
                 int row_hint = 0) {
    int slot;
    int rc = prof_begin(h, SEC_INDEX, st, &slot);
    if (rc) return rc;
    // ~6 bands (6 x 32 KB) per CTA, so that the double buffer has something to overlap
    const int64_t max_bands = (h->slot_points[kind] + 8 * KI_COLS - 1) / (8 * KI_COLS) + 8;
    int parts = (int)((max_bands + 5) / 6);
    if (parts < 1) parts = 1;
    if (!h->index_smem_set) {
        CK(cudaFuncSetAttribute(cloud_index_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 2 * KI_BAND_BYTES));
        h->index_smem_set = true;
    }
    if (parts > 65535) parts = 65535;
    if (sort_mode && !h->sorted[kind].p) {
        CK(cudaStreamSynchronize(st));
        CK(h->sorted[kind].reserve((size_t)h->cfg.max_scenes * h->slot_points[kind] * 16));
    }
    cloud_index_kernel<<<dim3(n_scenes, parts), KI_THREADS, 2 * KI_BAND_BYTES, st>>>(
        h->cloud[kind].as<float4>(), h->boxes[kind].as<float4>(), h->counts[kind].as<int32_t>(),
        h->nan_flags[kind].as<int32_t>(), h->slot_points[kind], h->slot_tiles[kind], h->layout[kind].as<int32_t>(), first_scene);
    h->launches++;
    CK(cudaGetLastError());
    cloud_compact_kernel<<<n_scenes, KI_THREADS, 0, st>>>(
        h->cloud[kind].as<float4>(), h->boxes[kind].as<float4>(), h->counts[kind].as<int32_t>(),
        h->nan_flags[kind].as<int32_t>(), h->slot_points[kind], h->slot_tiles[kind], h->layout[kind].as<int32_t>(), first_scene);
    h->launches++;
    CK(cudaGetLastError());
    if (sort_mode) {
        if (!h->sort_smem_set) {
            CK(cudaFuncSetAttribute(cloud_sort_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, KM_SMEM_BYTES));
            h->sort_smem_set = true;
        }
        cloud_sort_kernel<<<n_scenes, KM_THREADS, KM_SMEM_BYTES, st>>>(
            h->cloud[kind].as<float4>(), h->sorted[kind].as<float4>(), h->boxes[kind].as<float4>(),
            h->counts[kind].as<int32_t>(), h->layout[kind].as<int32_t>(), h->slot_points[kind], h->slot_tiles[kind],
            first_scene);
        h->launches++;
        CK(cudaGetLastError());
        // tile boxes over the bucketed copy replace the ones of the original order
        cloud_index_kernel<<<dim3(n_scenes, parts), KI_THREADS, 2 * KI_BAND_BYTES, st>>>(
            h->sorted[kind].as<float4>(), h->boxes[kind].as<float4>(), h->counts[kind].as<int32_t>(),
            h->nan_flags[kind].as<int32_t>(), h->slot_points[kind], h->slot_tiles[kind], h->layout[kind].as<int32_t>(), first_scene);
        h->launches++;
        CK(cudaGetLastError());
    }
    // second level: boxes of the groups of 16 tiles.  Grid: the most groups a scene of this handle
    // can have under the layouts in play (linear, the handle's row hint, the caller's row hint)
    int g_max = GroupGeom(TileGeom(h->slot_points[kind], 0)).n_groups;
    for (int w : {h->row_w[kind], row_hint})
        if (w > 0 && tile_layout(h->slot_points[kind], w, h->slot_tiles[kind]) > 0) {
            // fewer points mean fewer rows, never more groups
            const int gw = GroupGeom(TileGeom(h->slot_points[kind], w)).n_groups;
            if (gw > g_max) g_max = gw;
        }
    if (g_max > h->slot_groups[kind]) g_max = h->slot_groups[kind];
    group_boxes_kernel<<<dim3(n_scenes, (g_max + 7) / 8), 256, 0, st>>>(
        h->boxes[kind].as<float4>(), h->gboxes[kind].as<float4>(), h->counts[kind].as<int32_t>(), h->slot_tiles[kind],
        h->slot_groups[kind], h->layout[kind].as<int32_t>(), first_scene);
    h->launches++;
    CK(cudaGetLastError());
    return prof_end(h, slot, st);
}

template <int W>
int launch_solve_w(ampc_handle *h, int B, const double *prefix_dev, double *w_dev, SolveOut *info_dev,
                   cudaStream_t st, const int32_t *active) {
    using namespace ampc::v1;
    const size_t smem = solve_smem_bytes(h->cfg.N, W) + (size_t)h->solve_smem_pad;
    if (smem > 227 * 1024)
        return fail(h, AMPC_ERR_UNSUPPORTED, "horizon too long for the per-warp shared-memory layout");
    if ((int)smem > h->solve_smem_set[W]) {
        CK(cudaFuncSetAttribute(ipm_solve_kernel<W>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        h->solve_smem_set[W] = (int)smem;
    }
    const int grid = (B + W - 1) / W;
    ipm_solve_kernel<W><<<grid, W * 32, smem, st>>>(h->consts, B, prefix_dev, w_dev, info_dev, active);
    h->launches++;
    CK(cudaGetLastError());
    return AMPC_OK;
}

// Quad kernel (ipm_quad.cuh): persistent warps (one or two per CTA) with Q = 1, 2, 4 or 8 instances in flight
// per warp.  Shared memory per warp grows with Q (and N), registers allow 8 warps per SM.
int quad_warps_per_sm(const ampc_handle *h, int Q) {
    const size_t smem = quad_smem_bytes(h->cfg.N, Q) + 1024; // + the per-CTA reservation
    int w = (int)((size_t)h->smem_per_sm / smem);
    if (w > h->quad_warps_per_sm) w = h->quad_warps_per_sm;
    return w;
}
int launch_solve_quad(ampc_handle *h, int B, const double *prefix_dev, double *w_dev, SolveOut *info_dev,
                      cudaStream_t st, const int32_t *active) {
    // 8 instances per warp amortise the sweep best (measured: B = 32768 in 29.3 ms with 8, 35.7 ms
    // with 4); fewer only when the shared-memory state of 8 does not fit (long horizons)
    int qs = 3;
    if (h->quad_per_warp > 0)
        for (qs = 0; (1 << qs) < h->quad_per_warp; ++qs) {}
    while (qs > 0 && quad_warps_per_sm(h, 1 << qs) < 1) --qs;
    const int Q = 1 << qs;
    const int per_sm = quad_warps_per_sm(h, Q);
    if (per_sm < 1)
        return fail(h, AMPC_ERR_UNSUPPORTED, "horizon too long for the shared-memory state of one instance");
    const int cap = h->n_sm * per_sm; // resident warps
    int warps = (B + Q - 1) / Q;
    if (warps > cap) warps = cap;
    // warps per CTA (AMPC_QUAD_CTA_WARPS): only when the machine is full anyway; whole CTAs per SM
    int cw = 1;
    if (warps == cap && h->quad_cta_warps > 1) {
        cw = h->quad_cta_warps < per_sm ? h->quad_cta_warps : per_sm;
        while (per_sm % cw) --cw;
    }
    const size_t smem = quad_smem_bytes(h->cfg.N, Q);
    {
        // the attribute belongs to the function, not to the handle: only ever raise it
        static std::mutex mtx;
        static int set_to[64][4] = {{0}};
        std::lock_guard<std::mutex> g(mtx);
        const int dev = h->cfg.device & 63;
        if ((int)(smem * cw) > set_to[dev][qs]) {
            const void *fn = qs == 0 ? (const void *)ipm_quad_kernel<0>
                           : qs == 1 ? (const void *)ipm_quad_kernel<1>
                           : qs == 2 ? (const void *)ipm_quad_kernel<2> : (const void *)ipm_quad_kernel<3>;
            CK(cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(smem * cw)));
            set_to[dev][qs] = (int)(smem * cw);
        }
    }
    const size_t need = (size_t)warps * quad_ws_bytes_per_warp(h->cfg.N, Q);
    if (need > h->quad_ws.bytes) {
        CK(cudaStreamSynchronize(st));
        size_t full = 0; // enough for any packing
        for (int t = 0; t <= 3; ++t) {
            const size_t v = (size_t)h->n_sm * (size_t)(quad_warps_per_sm(h, 1 << t) > 0 ? quad_warps_per_sm(h, 1 << t) : 0) *
                             quad_ws_bytes_per_warp(h->cfg.N, 1 << t);
            if (v > full) full = v;
        }
        CK(h->quad_ws.reserve(full > need ? full : need));
    }
    // counter[0]: queue head; counter[1..64]: class histogram + running offsets of the ordering
    CK(h->quad_counter.reserve(512));
    CK(cudaMemsetAsync(h->quad_counter.p, 0, 512, st));
    double *wsp = h->quad_ws.as<double>();
    int32_t *cnt = h->quad_counter.as<int32_t>();
    const int32_t *order = nullptr;
    if (h->quad_order && B > warps * Q && h->cfg.K > 0 && h->cfg.N > 1) {
        // more instances than slots: the queue is refilled, take the hard instances first
        CK(h->quad_order_buf.reserve((size_t)B * 8));
        int32_t *cls = h->quad_order_buf.as<int32_t>(), *ord = cls + B;
        solve_order_class_kernel<<<(B + 7) / 8, 256, 0, st>>>(B, h->cfg.N, h->cfg.K, h->n_prefix, h->radius, prefix_dev,
                                                           active, cls, cnt + 1);
        solve_order_scatter_kernel<<<(B + 255) / 256, 256, 0, st>>>(B, cls, cnt + 1, ord);
        h->launches += 2;
        CK(cudaGetLastError());
        order = ord;
    }
    const int grid = (warps + cw - 1) / cw;
    switch (qs) {
    case 0: ipm_quad_kernel<0><<<grid, 32 * cw, smem * cw, st>>>(h->consts, B, prefix_dev, w_dev, info_dev, active, order, wsp, cnt); break;
    case 1: ipm_quad_kernel<1><<<grid, 32 * cw, smem * cw, st>>>(h->consts, B, prefix_dev, w_dev, info_dev, active, order, wsp, cnt); break;
    case 2: ipm_quad_kernel<2><<<grid, 32 * cw, smem * cw, st>>>(h->consts, B, prefix_dev, w_dev, info_dev, active, order, wsp, cnt); break;
    default: ipm_quad_kernel<3><<<grid, 32 * cw, smem * cw, st>>>(h->consts, B, prefix_dev, w_dev, info_dev, active, order, wsp, cnt); break;
    }
    h->launches++;
    CK(cudaGetLastError());
    return AMPC_OK;
}

int launch_solve(ampc_handle *h, int B, const double *prefix_dev, double *w_dev, SolveOut *info_dev,
                 cudaStream_t st, const int32_t *active = nullptr) {
    refresh_consts(h);
    // Two kernels for one algorithm.  Small batches: one warp per instance (lowest latency, the
    // machine is not full anyway).  Large batches: four lanes per instance, eight instances per
    // persistent warp with refill from a queue (fewest instructions per solve).  Measured on B200
    // (solve stage, ms, warp / quad): B = 4096: 9.0 / 11.9, 8192: 13.8 / 13.6, 16384: 23.7 / 17.6,
    // 32768: 43.5 / 29.3.
    const bool warp_fits = ampc::v1::solve_smem_bytes(h->cfg.N, 1) <= 227 * 1024;
    const bool use_quad = h->solve_kernel == 2 || !warp_fits || (h->solve_kernel == 0 && B >= h->quad_min_batch);
    if (use_quad)
        return launch_solve_quad(h, B, prefix_dev, w_dev, info_dev, st, active);
    // warps (= instances) per CTA (AMPC_SOLVE_WARPS; measured best at C1: 2): the warps of a CTA start
    // every iteration together (shared instruction cache); horizons too long for the per-warp
    // shared memory of W warps fall back to fewer
    int W = h->solve_warps;
    while (W > 1 && ampc::v1::solve_smem_bytes(h->cfg.N, W) > 227 * 1024)
        W >>= 1;
    switch (W) {
    case 8: return launch_solve_w<8>(h, B, prefix_dev, w_dev, info_dev, st, active);
    case 4: return launch_solve_w<4>(h, B, prefix_dev, w_dev, info_dev, st, active);
    case 2: return launch_solve_w<2>(h, B, prefix_dev, w_dev, info_dev, st, active);
    default: return launch_solve_w<1>(h, B, prefix_dev, w_dev, info_dev, st, active);
    }
}

int check_batch(ampc_handle *h, int B) {
    if (!h)
        return AMPC_ERR_INVALID;
    if (B < 1)
        return fail(h, AMPC_ERR_INVALID, "B must be positive");
    if (B > h->cfg.max_batch)
        return fail(h, AMPC_ERR_CAPACITY, "B exceeds ampc_config.max_batch");
    return AMPC_OK;
}

} // namespace

extern "C" {

int ampc_api_version(void) { return AMPC_API_VERSION; }

const char *ampc_last_error(const ampc_handle *h) {
    if (h)
        return h->err.c_str();
    std::lock_guard<std::mutex> g(g_create_mutex);
    return g_create_error.c_str();
}

void ampc_default_solver_opts(ampc_solver_opts *o) {
    if (!o)
        return;
    o->tol = 1e-8;
    o->max_iter = 100;
    o->mu_init = 0.1;
    o->bound_push = 1e-2;
    o->bound_frac = 1e-2;
    o->eps_min = 1e-5;
    o->eps_scale = 1.0;
    o->kappa_eps = 100.0;
}

int ampc_create(const ampc_config *cfg, ampc_handle **out) {
    auto cfail = [](int code, const std::string &m) {
        std::lock_guard<std::mutex> g(g_create_mutex);
        g_create_error = m;
        return code;
    };
    if (!cfg || !out)
        return cfail(AMPC_ERR_INVALID, "null argument");
    *out = nullptr;
    if (cfg->N < 1 || cfg->N > 256 || cfg->K < 0 || cfg->K > KNN_KMAX || !(cfg->dt > 0) ||
        cfg->max_batch < 1 || cfg->max_scenes < 0 || cfg->max_points < 0 || cfg->max_edge_points < 0)
        return cfail(AMPC_ERR_INVALID, "bad ampc_config (need 1<=N<=256, 0<=K<=32, dt>0, max_batch>=1)");
    int ndev = 0;
    cudaError_t e = cudaGetDeviceCount(&ndev);
    if (e != cudaSuccess || ndev == 0)
        return cfail(AMPC_ERR_CUDA, std::string("no CUDA device: ") + cudaGetErrorString(e) +
                                        " (libampc has no CPU fallback)");
    if (cfg->device < 0 || cfg->device >= ndev)
        return cfail(AMPC_ERR_INVALID, "device ordinal out of range");
    e = cudaSetDevice(cfg->device);
    if (e != cudaSuccess)
        return cfail(AMPC_ERR_CUDA, std::string("cudaSetDevice: ") + cudaGetErrorString(e));
    ampc_handle *h = new (std::nothrow) ampc_handle();
    if (!h)
        return cfail(AMPC_ERR_INVALID, "out of host memory");
    h->cfg = *cfg;
    h->n_w = 10 + 14 * cfg->N;
    h->n_prefix = 20 + 10 * cfg->N + 3 * cfg->K * cfg->N;
    ampc_default_solver_opts(&h->opts);
    if (const char *e = std::getenv("AMPC_SOLVE_WARPS")) { // tuning knob: 1, 2, 4 or 8
        const int v = std::atoi(e);
        if (v == 1 || v == 2 || v == 4 || v == 8) h->solve_warps = v;
    }
    if (const char *e = std::getenv("AMPC_SOLVE_KERNEL"))
        h->solve_kernel = std::strcmp(e, "warp") == 0 ? 1 : (std::strcmp(e, "quad") == 0 ? 2 : 0);
    if (const char *e = std::getenv("AMPC_KNN_TWO_LEVEL"))
        h->knn_two_level = std::atoi(e) != 0;
    if (const char *e = std::getenv("AMPC_QUAD_ORDER"))
        h->quad_order = std::atoi(e) != 0;
    if (const char *e = std::getenv("AMPC_QUAD_MIN_BATCH")) {
        const int v = std::atoi(e);
        if (v >= 1) h->quad_min_batch = v;
    }
    if (const char *e = std::getenv("AMPC_QUADS_PER_WARP")) {
        const int v = std::atoi(e);
        if (v == 1 || v == 2 || v == 4 || v == 8) h->quad_per_warp = v;
    }
    if (const char *e = std::getenv("AMPC_QUAD_CTA_WARPS")) {
        const int v = std::atoi(e);
        if (v >= 1 && v <= 8) h->quad_cta_warps = v;
    }
    if (const char *e = std::getenv("AMPC_QUAD_WARPS_PER_SM")) {
        const int v = std::atoi(e);
        if (v >= 1 && v <= 16) h->quad_warps_per_sm = v;
    }
    {
        int v = 0;
        if (cudaDeviceGetAttribute(&v, cudaDevAttrMultiProcessorCount, cfg->device) == cudaSuccess && v > 0)
            h->n_sm = v;
        if (cudaDeviceGetAttribute(&v, cudaDevAttrMaxSharedMemoryPerMultiprocessor, cfg->device) == cudaSuccess && v > 0)
            h->smem_per_sm = v;
    }
    if (const char *e = std::getenv("AMPC_SOLVE_SMEM_PAD")) {
        const int v = std::atoi(e);
        if (v > 0 && v < 200 * 1024) h->solve_smem_pad = v & ~15;
    }
    // defaults of the reference constructor (src/HighLvlMpc.cpp:11-14,53-58)
    const double w0[25] = {100, 100, 100, 300, 1, 1, 1, 0., 0., 0., 0.0, 10, 10,
                           30,  0,   1,   1,   0., 0., 0., 1., 1., 1., 1., 1.};
    std::memcpy(h->weights, w0, sizeof w0);
    const double t0[4] = {0.01, 0.01, 0.01, 0}, g0[4] = {1, 1, 1, 1};
    std::memcpy(h->tau, t0, sizeof t0);
    std::memcpy(h->gains, g0, sizeof g0);
    h->radius = 0.0;
    const double l0[4] = {-10., -10., 1., -10.}, u0[4] = {10., 10., 20., 10.};
    std::memcpy(h->lb, l0, sizeof l0);
    std::memcpy(h->ub, u0, sizeof u0);
    auto bail = [&](cudaError_t ce, const char *what) {
        std::string m = std::string(what) + ": " + cudaGetErrorString(ce);
        ampc_destroy(h);
        return cfail(AMPC_ERR_CUDA, m);
    };
    if ((e = cudaStreamCreateWithFlags(&h->stream, cudaStreamNonBlocking)) != cudaSuccess)
        return bail(e, "cudaStreamCreate");
    h->slot_points[0] = cfg->max_points;
    h->slot_points[1] = cfg->max_edge_points;
    for (int kind = 0; kind < 2; ++kind) {
        if (cfg->max_scenes == 0 || h->slot_points[kind] == 0)
            continue;
        if ((e = h->cloud[kind].reserve((size_t)cfg->max_scenes * h->slot_points[kind] * 16)) != cudaSuccess)
            return bail(e, "cudaMalloc(clouds)");
        if ((e = h->counts[kind].reserve((size_t)cfg->max_scenes * 4)) != cudaSuccess)
            return bail(e, "cudaMalloc(counts)");
        h->slot_tiles[kind] = (int)tile_capacity(h->slot_points[kind]);
        if ((e = h->boxes[kind].reserve((size_t)cfg->max_scenes * h->slot_tiles[kind] * 32)) != cudaSuccess)
            return bail(e, "cudaMalloc(tile boxes)");
        h->slot_groups[kind] = (int)group_capacity(h->slot_tiles[kind]);
        if ((e = h->gboxes[kind].reserve((size_t)cfg->max_scenes * h->slot_groups[kind] * 32)) != cudaSuccess)
            return bail(e, "cudaMalloc(group boxes)");
        if ((e = cudaMemset(h->counts[kind].p, 0, (size_t)cfg->max_scenes * 4)) != cudaSuccess)
            return bail(e, "cudaMemset(counts)");
        if ((e = h->nan_flags[kind].reserve((size_t)cfg->max_scenes * 4)) != cudaSuccess)
            return bail(e, "cudaMalloc(nan flags)");
        if ((e = cudaMemset(h->nan_flags[kind].p, 0, (size_t)cfg->max_scenes * 4)) != cudaSuccess)
            return bail(e, "cudaMemset(nan flags)");
        if ((e = h->layout[kind].reserve((size_t)cfg->max_scenes * 4)) != cudaSuccess)
            return bail(e, "cudaMalloc(layout)");
        if ((e = cudaMemset(h->layout[kind].p, 0, (size_t)cfg->max_scenes * 4)) != cudaSuccess)
            return bail(e, "cudaMemset(layout)");
    }
    const size_t B = (size_t)cfg->max_batch, N = (size_t)cfg->N, K = (size_t)(cfg->K > 0 ? cfg->K : 1);
    struct { DevBuf *b; size_t n; } need[] = {
        {&h->queries, B * N * 3 * 8}, {&h->prefix, B * (size_t)h->n_prefix * 8},
        {&h->w, B * (size_t)h->n_w * 8}, {&h->info, B * sizeof(SolveOut)},
        {&h->knn_idx, B * N * K * 4}, {&h->knn_d2, B * N * K * 8}, {&h->knn_cnt, B * N * 4},
        {&h->scene_of, B * 4}, {&h->x0, B * 10 * 8}, {&h->ref, B * N * 10 * 8},
        {&h->posx, B * 8}, {&h->replan, B * 4}, {&h->bo_arg, B * 4}, {&h->bo_cost, B * 8}};
    for (auto &nb : need)
        if ((e = nb.b->reserve(nb.n)) != cudaSuccess)
            return bail(e, "cudaMalloc(workspace)");
    *out = h;
    return AMPC_OK;
}

void ampc_destroy(ampc_handle *h) {
    if (!h)
        return;
    cudaSetDevice(h->cfg.device);
    if (h->stream)
        cudaStreamSynchronize(h->stream);
    DevBuf *all[] = {&h->sorted[0], &h->sorted[1], &h->layout[0], &h->layout[1], &h->depth_stage, &h->depth_T, &h->depth_tab,
                     &h->depth_scratch, &h->depth_flag, &h->cloud[0], &h->cloud[1], &h->counts[0], &h->counts[1], &h->boxes[0], &h->boxes[1], &h->gboxes[0], &h->gboxes[1], &h->nan_flags[0], &h->nan_flags[1], &h->raw_stage,
                     &h->queries, &h->prefix, &h->w, &h->info, &h->knn_idx, &h->knn_d2, &h->knn_cnt,
                     &h->knn_pts, &h->scene_of, &h->x0, &h->ref, &h->posx, &h->replan, &h->ws_d,
                     &h->ws_i, &h->ws_counter, &h->quad_ws, &h->quad_counter, &h->quad_order_buf, &h->gs_sites, &h->gs_nb, &h->gs_epts, &h->gs_ecnt, &h->gs_q0, &h->bo_arg, &h->bo_cost, &h->tk_active, &h->tk_q0, &h->tk_d1,
                     &h->tk_c1, &h->tk_ep, &h->tk_ec, &h->tk_ed, &h->tk_safe, &h->tk_rounds};
    for (DevBuf *b : all)
        b->release();
    for (auto &e : h->prof_ev) cudaEventDestroy(e);
    if (h->stream)
        cudaStreamDestroy(h->stream);
    delete h;
}

int ampc_set_weights(ampc_handle *h, const double w[25]) {
    if (!h || !w) return AMPC_ERR_INVALID;
    std::memcpy(h->weights, w, sizeof h->weights);
    h->consts_dirty = true;
    return AMPC_OK;
}
int ampc_set_tau(ampc_handle *h, const double t[4]) {
    if (!h || !t) return AMPC_ERR_INVALID;
    std::memcpy(h->tau, t, sizeof h->tau);
    h->consts_dirty = true;
    return AMPC_OK;
}
int ampc_set_gains(ampc_handle *h, const double g[4]) {
    if (!h || !g) return AMPC_ERR_INVALID;
    std::memcpy(h->gains, g, sizeof h->gains);
    return AMPC_OK;
}
int ampc_set_radius(ampc_handle *h, double r) {
    if (!h) return AMPC_ERR_INVALID;
    h->radius = r;
    h->consts_dirty = true;
    return AMPC_OK;
}
int ampc_set_accel_limits(ampc_handle *h, double a_min_z, double a_max_z, double a_max_xy,
                          double a_max_yaw_dot) {
    if (!h) return AMPC_ERR_INVALID;
    if (!(a_min_z < a_max_z) || !(a_max_xy > 0) || !(a_max_yaw_dot > 0))
        return fail(h, AMPC_ERR_INVALID, "empty control box");
    const double l[4] = {-a_max_xy, -a_max_xy, a_min_z, -a_max_yaw_dot};
    const double u[4] = {a_max_xy, a_max_xy, a_max_z, a_max_yaw_dot};
    std::memcpy(h->lb, l, sizeof l);
    std::memcpy(h->ub, u, sizeof u);
    h->consts_dirty = true;
    return AMPC_OK;
}
int ampc_set_solver_opts(ampc_handle *h, const ampc_solver_opts *o) {
    if (!h || !o) return AMPC_ERR_INVALID;
    if (!(o->tol > 0) || o->max_iter < 0 || !(o->mu_init > 0) || !(o->eps_min >= 0) || !(o->kappa_eps > 0))
        return fail(h, AMPC_ERR_INVALID, "bad solver options");
    h->opts = *o;
    h->consts_dirty = true;
    return AMPC_OK;
}
int ampc_get_dynamics(ampc_handle *h, double Phi[100], double Gam[40], double gam[10]) {
    if (!h || !Phi || !Gam || !gam) return AMPC_ERR_INVALID;
    refresh_consts(h);
    std::memcpy(Phi, h->consts.Phi, sizeof h->consts.Phi);
    std::memcpy(Gam, h->consts.Gam, sizeof h->consts.Gam);
    std::memcpy(gam, h->consts.gam, sizeof h->consts.gam);
    return AMPC_OK;
}

int64_t ampc_launch_count(const ampc_handle *h) { return h ? h->launches : 0; }
void *ampc_stream(ampc_handle *h) { return h ? (void *)h->stream : nullptr; }
int ampc_synchronize(ampc_handle *h) {
    if (!h) return AMPC_ERR_INVALID;
    CK(cudaStreamSynchronize(h->stream));
    return AMPC_OK;
}

// ---- clouds -----------------------------------------------------------------
// the scenes being (re)built take the handle's current layout hint
static int set_layout(ampc_handle *h, int kind, int first_scene, int n_scenes, cudaStream_t st) {
    fill_i32_kernel<<<(n_scenes + 255) / 256, 256, 0, st>>>(h->layout[kind].as<int32_t>() + first_scene, n_scenes,
                                                            h->row_w[kind]);
    h->launches++;
    CK(cudaGetLastError());
    return AMPC_OK;
}

static int cloud_set_common(ampc_handle *h, int kind, int first_scene, int n_scenes, const void *src,
                            bool src_is_device, const int32_t *counts, int64_t scene_stride,
                            int stride, cudaStream_t st) {
    if (!h) return AMPC_ERR_INVALID;
    int rc = check_kind(h, kind);
    if (rc) return rc;
    if (n_scenes < 1 || first_scene < 0 || first_scene + n_scenes > h->cfg.max_scenes)
        return fail(h, AMPC_ERR_CAPACITY, "scene range exceeds ampc_config.max_scenes");
    if (!counts || (!src && n_scenes > 0))
        return fail(h, AMPC_ERR_INVALID, "null cloud pointer");
    if (stride < 12 || (stride & 3))
        return fail(h, AMPC_ERR_INVALID, "stride_bytes must be a multiple of 4 and >= 12");
    int maxc = 0;
    for (int s = 0; s < n_scenes; ++s) {
        if (counts[s] < 0 || counts[s] > h->slot_points[kind])
            return fail(h, AMPC_ERR_CAPACITY, "cloud has more points than the slot capacity");
        if (counts[s] > maxc) maxc = counts[s];
    }
    CK(cudaSetDevice(h->cfg.device));
    int32_t *dcounts = h->counts[kind].as<int32_t>() + first_scene;
    CK(cudaMemcpyAsync(dcounts, counts, (size_t)n_scenes * 4, cudaMemcpyHostToDevice, st));
    float4 *slots = h->cloud[kind].as<float4>();
    const int64_t slot_bytes = (int64_t)h->slot_points[kind] * 16;
    const cudaMemcpyKind dir = src_is_device ? cudaMemcpyDeviceToDevice : cudaMemcpyHostToDevice;
    if (maxc > 0) {
        if (stride == 16) {
            char *dst = reinterpret_cast<char *>(slots) + (int64_t)first_scene * slot_bytes;
            // the caller's buffer is only promised to hold counts[s] records of scene s: never read
            // past the last scene's own points
            const int last = counts[n_scenes - 1];
            if (n_scenes == 1 || (scene_stride == slot_bytes && maxc == h->slot_points[kind])) {
                const size_t bytes = (size_t)(n_scenes - 1) * slot_bytes + (size_t)last * 16;
                if (bytes) CK(cudaMemcpyAsync(dst, src, bytes, dir, st));
            } else {
                if (scene_stride < (int64_t)maxc * 16)
                    return fail(h, AMPC_ERR_INVALID, "scene_stride_bytes smaller than the largest cloud");
                CK(cudaMemcpy2DAsync(dst, (size_t)slot_bytes, src, (size_t)scene_stride, (size_t)maxc * 16,
                                     (size_t)(n_scenes - 1), dir, st));
                if (last)
                    CK(cudaMemcpyAsync(dst + (int64_t)(n_scenes - 1) * slot_bytes,
                                       static_cast<const char *>(src) + (int64_t)(n_scenes - 1) * scene_stride,
                                       (size_t)last * 16, dir, st));
            }
        } else {
            if (n_scenes > 1 && scene_stride < (int64_t)maxc * stride)
                return fail(h, AMPC_ERR_INVALID, "scene_stride_bytes smaller than the largest cloud");
            const unsigned char *raw = static_cast<const unsigned char *>(src);
            if (!src_is_device) {
                const size_t per = n_scenes == 1 ? (size_t)maxc * stride : (size_t)scene_stride;
                CK(h->raw_stage.reserve(per * n_scenes));
                const size_t bytes = per * (n_scenes - 1) + (size_t)counts[n_scenes - 1] * stride; // no over-read
                CK(cudaMemcpyAsync(h->raw_stage.p, src, bytes, cudaMemcpyHostToDevice, st));
                raw = h->raw_stage.as<unsigned char>();
                if (n_scenes == 1) scene_stride = (int64_t)per;
            }
            repack_kernel<<<dim3(n_scenes, (maxc + 255) / 256 > 64 ? 64 : (maxc + 255) / 256), 256, 0, st>>>(
                raw, scene_stride, stride, slots, h->counts[kind].as<int32_t>(), h->slot_points[kind], first_scene);
            h->launches++;
            CK(cudaGetLastError());
        }
    }
    rc = set_layout(h, kind, first_scene, n_scenes, st);
    if (rc) return rc;
    return launch_index(h, kind, first_scene, n_scenes, st, h->row_w[kind] == AMPC_LAYOUT_SORT);
}

int ampc_cloud_set(ampc_handle *h, int32_t scene, int32_t kind, const void *xyz_host, int32_t n,
                   int32_t stride_bytes) {
    if (!h) return AMPC_ERR_INVALID;
    int rc = cloud_set_common(h, kind, scene, 1, xyz_host ? xyz_host : (const void *)h, false, &n, 0,
                              stride_bytes, h->stream);
    if (rc) return rc;
    CK(cudaStreamSynchronize(h->stream));
    return AMPC_OK;
}
int ampc_cloud_set_batch(ampc_handle *h, int32_t kind, int32_t first_scene, int32_t n_scenes,
                         const void *xyz_host, const int32_t *counts, int64_t scene_stride_bytes,
                         int32_t stride_bytes) {
    if (!h) return AMPC_ERR_INVALID;
    int rc = cloud_set_common(h, kind, first_scene, n_scenes, xyz_host, false, counts,
                              scene_stride_bytes, stride_bytes, h->stream);
    if (rc) return rc;
    CK(cudaStreamSynchronize(h->stream));
    return AMPC_OK;
}
int ampc_cloud_set_batch_dev(ampc_handle *h, int32_t kind, int32_t first_scene, int32_t n_scenes,
                             const void *xyz_dev, const int32_t *counts_host,
                             int64_t scene_stride_bytes, void *stream) {
    return cloud_set_common(h, kind, first_scene, n_scenes, xyz_dev, true, counts_host,
                            scene_stride_bytes, 16, (cudaStream_t)stream);
}
int ampc_cloud_set_layout(ampc_handle *h, int32_t kind, int32_t row_width) {
    if (!h) return AMPC_ERR_INVALID;
    int rc = check_kind(h, kind);
    if (rc) return rc;
    if (row_width != AMPC_LAYOUT_UNORGANISED && row_width != AMPC_LAYOUT_SORT && row_width < 8)
        return fail(h, AMPC_ERR_INVALID, "row_width must be AMPC_LAYOUT_UNORGANISED (0), AMPC_LAYOUT_SORT (-1) or >= 8");
    h->row_w[kind] = row_width;
    return AMPC_OK;
}

int ampc_cloud_count(ampc_handle *h, int32_t scene, int32_t kind, int32_t *n_out) {
    if (!h || !n_out) return AMPC_ERR_INVALID;
    int rc = check_kind(h, kind);
    if (rc) return rc;
    if (scene < 0 || scene >= h->cfg.max_scenes)
        return fail(h, AMPC_ERR_CAPACITY, "scene out of range");
    CK(cudaSetDevice(h->cfg.device));
    CK(cudaDeviceSynchronize());
    CK(cudaMemcpy(n_out, h->counts[kind].as<int32_t>() + scene, 4, cudaMemcpyDeviceToHost));
    return AMPC_OK;
}

// ---- k-NN -------------------------------------------------------------------
int ampc_knn_batch_dev(ampc_handle *h, int32_t kind, int32_t B, const int32_t *scene_of_dev,
                       const double *queries_dev, int32_t Q, int32_t k, int32_t *idx_dev,
                       double *dist2_dev, double *pts_dev, int32_t *count_dev, void *stream) {
    if (!h) return AMPC_ERR_INVALID;
    int rc = check_kind(h, kind);
    if (rc) return rc;
    if (!queries_dev) return fail(h, AMPC_ERR_INVALID, "null queries");
    if (!scene_of_dev && B > h->cfg.max_scenes)
        return fail(h, AMPC_ERR_CAPACITY, "identity scene map needs B <= max_scenes");
    CK(cudaSetDevice(h->cfg.device));
    return launch_knn(h, kind, B, scene_of_dev, queries_dev, Q, k, idx_dev, dist2_dev, count_dev,
                      pts_dev, (int64_t)Q * k * 3, (int64_t)k * 3, (cudaStream_t)stream);
}

int ampc_knn_batch(ampc_handle *h, int32_t kind, int32_t B, const int32_t *scene_of,
                   const double *queries, int32_t Q, int32_t k, int32_t *idx_out, double *dist2_out,
                   double *pts_out, int32_t *count_out) {
    if (!h) return AMPC_ERR_INVALID;
    int rc = check_kind(h, kind);
    if (rc) return rc;
    if (B < 1 || Q < 1 || k < 1) return fail(h, AMPC_ERR_INVALID, "B, Q, k must be positive");
    if (!queries) return fail(h, AMPC_ERR_INVALID, "null queries");
    if (scene_of)
        for (int b = 0; b < B; ++b)
            if (scene_of[b] < 0 || scene_of[b] >= h->cfg.max_scenes)
                return fail(h, AMPC_ERR_CAPACITY, "scene_of entry out of range");
    CK(cudaSetDevice(h->cfg.device));
    cudaStream_t st = h->stream;
    const size_t nq = (size_t)B * Q, nr = nq * k;
    CK(h->queries.reserve(nq * 3 * 8));
    CK(h->knn_idx.reserve(nr * 4));
    CK(h->knn_d2.reserve(nr * 8));
    CK(h->knn_cnt.reserve(nq * 4));
    CK(h->scene_of.reserve((size_t)B * 4));
    if (pts_out) CK(h->knn_pts.reserve(nr * 3 * 8));
    CK(cudaMemcpyAsync(h->queries.p, queries, nq * 3 * 8, cudaMemcpyHostToDevice, st));
    if (scene_of) CK(cudaMemcpyAsync(h->scene_of.p, scene_of, (size_t)B * 4, cudaMemcpyHostToDevice, st));
    rc = ampc_knn_batch_dev(h, kind, B, scene_of ? h->scene_of.as<int32_t>() : nullptr,
                            h->queries.as<double>(), Q, k, h->knn_idx.as<int32_t>(),
                            h->knn_d2.as<double>(), pts_out ? h->knn_pts.as<double>() : nullptr,
                            h->knn_cnt.as<int32_t>(), st);
    if (rc) return rc;
    if (idx_out) CK(cudaMemcpyAsync(idx_out, h->knn_idx.p, nr * 4, cudaMemcpyDeviceToHost, st));
    if (dist2_out) CK(cudaMemcpyAsync(dist2_out, h->knn_d2.p, nr * 8, cudaMemcpyDeviceToHost, st));
    if (pts_out) CK(cudaMemcpyAsync(pts_out, h->knn_pts.p, nr * 3 * 8, cudaMemcpyDeviceToHost, st));
    if (count_out) CK(cudaMemcpyAsync(count_out, h->knn_cnt.p, nq * 4, cudaMemcpyDeviceToHost, st));
    CK(cudaStreamSynchronize(st));
    return AMPC_OK;
}

// ---- solve ------------------------------------------------------------------
int ampc_solve_batch_dev(ampc_handle *h, int32_t B, const double *p_prefix_dev, double *w_inout_dev,
                         ampc_solve_info *info_dev, void *stream) {
    int rc = check_batch(h, B);
    if (rc) return rc;
    if (!p_prefix_dev || !w_inout_dev) return fail(h, AMPC_ERR_INVALID, "null buffer");
    CK(cudaSetDevice(h->cfg.device));
    SolveOut *info = info_dev ? reinterpret_cast<SolveOut *>(info_dev) : h->info.as<SolveOut>();
    return launch_solve(h, B, p_prefix_dev, w_inout_dev, info, (cudaStream_t)stream);
}

int ampc_solve_batch(ampc_handle *h, int32_t B, const double *p_prefix, double *w_inout,
                     ampc_solve_info *info_out) {
    int rc = check_batch(h, B);
    if (rc) return rc;
    if (!p_prefix || !w_inout) return fail(h, AMPC_ERR_INVALID, "null buffer");
    CK(cudaSetDevice(h->cfg.device));
    cudaStream_t st = h->stream;
    CK(cudaMemcpyAsync(h->prefix.p, p_prefix, (size_t)B * h->n_prefix * 8, cudaMemcpyHostToDevice, st));
    CK(cudaMemcpyAsync(h->w.p, w_inout, (size_t)B * h->n_w * 8, cudaMemcpyHostToDevice, st));
    rc = launch_solve(h, B, h->prefix.as<double>(), h->w.as<double>(), h->info.as<SolveOut>(), st);
    if (rc) return rc;
    CK(cudaMemcpyAsync(w_inout, h->w.p, (size_t)B * h->n_w * 8, cudaMemcpyDeviceToHost, st));
    if (info_out)
        CK(cudaMemcpyAsync(info_out, h->info.p, (size_t)B * sizeof(SolveOut), cudaMemcpyDeviceToHost, st));
    CK(cudaStreamSynchronize(st));
    return AMPC_OK;
}

// ---- one round: k-NN at the reference waypoints + prefix packing + solve ------
int ampc_round_batch_dev(ampc_handle *h, int32_t B, const int32_t *scene_of_dev, const double *x0_dev,
                         const double *ref_dev, const double *pos_x_dev, double speed,
                         double safety_distance, double *w_inout_dev, ampc_solve_info *info_dev,
                         int32_t *need_replan_dev, void *stream) {
    int rc = check_batch(h, B);
    if (rc) return rc;
    rc = check_kind(h, AMPC_CLOUD_OBSTACLE);
    if (rc) return rc;
    if (!x0_dev || !ref_dev || !w_inout_dev) return fail(h, AMPC_ERR_INVALID, "null buffer");
    if (h->cfg.K < 1) return fail(h, AMPC_ERR_INVALID, "round needs K >= 1");
    if (!scene_of_dev && B > h->cfg.max_scenes)
        return fail(h, AMPC_ERR_CAPACITY, "identity scene map needs B <= max_scenes");
    CK(cudaSetDevice(h->cfg.device));
    cudaStream_t st = (cudaStream_t)stream;
    const int N = h->cfg.N, K = h->cfg.K;
    double *prefix = h->prefix.as<double>();
    pack_prefix_kernel<<<B, 64, 0, st>>>(B, N, K, x0_dev, ref_dev, pos_x_dev, speed, N * h->cfg.dt,
                                         prefix, h->n_prefix, h->queries.as<double>());
    h->launches++;
    CK(cudaGetLastError());
    // the k-NN kernel writes the K neighbours of waypoint q straight into the
    // obstacle block of the prefix: obst_{q,j} at 10 + 10N + 3(Kq + j)
    int slot;
    if ((rc = prof_begin(h, SEC_KNN, st, &slot))) return rc;
    rc = launch_knn(h, AMPC_CLOUD_OBSTACLE, B, scene_of_dev, h->queries.as<double>(), N, K,
                    h->knn_idx.as<int32_t>(), h->knn_d2.as<double>(), h->knn_cnt.as<int32_t>(),
                    prefix + 10 + 10 * N, h->n_prefix, 3 * K, st);
    if (rc) return rc;
    if ((rc = prof_end(h, slot, st))) return rc;
    if (need_replan_dev) {
        replan_kernel<<<(B + 127) / 128, 128, 0, st>>>(B, N, K, h->knn_d2.as<double>(),
                                                       h->knn_cnt.as<int32_t>(), safety_distance,
                                                       need_replan_dev);
        h->launches++;
        CK(cudaGetLastError());
    }
    SolveOut *info = info_dev ? reinterpret_cast<SolveOut *>(info_dev) : h->info.as<SolveOut>();
    if ((rc = prof_begin(h, SEC_SOLVE, st, &slot))) return rc;
    rc = launch_solve(h, B, prefix, w_inout_dev, info, st);
    if (rc) return rc;
    return prof_end(h, slot, st);
}

int ampc_cloud_index_dev(ampc_handle *h, int32_t kind, int32_t first_scene, int32_t n_scenes, void *stream) {
    if (!h) return AMPC_ERR_INVALID;
    int rc = check_kind(h, kind);
    if (rc) return rc;
    if (n_scenes < 1 || first_scene < 0 || first_scene + n_scenes > h->cfg.max_scenes)
        return fail(h, AMPC_ERR_CAPACITY, "scene range exceeds ampc_config.max_scenes");
    CK(cudaSetDevice(h->cfg.device));
    rc = set_layout(h, kind, first_scene, n_scenes, (cudaStream_t)stream);
    if (rc) return rc;
    return launch_index(h, kind, first_scene, n_scenes, (cudaStream_t)stream, h->row_w[kind] == AMPC_LAYOUT_SORT);
}

// ---- depth image -> Obstacle + Edge cloud (FrameKDMap::ProcessDepth, src/FrameKDMap.cpp:90-130)
int ampc_set_camera(ampc_handle *h, const ampc_camera *cam) {
    if (!h || !cam) return AMPC_ERR_INVALID;
    if (!(cam->fx > 0) || !(cam->fy > 0) || !(cam->resize_scale >= 1) || !(cam->pixel2meter > 0) ||
        !(cam->depth_max > cam->depth_min))
        return fail(h, AMPC_ERR_INVALID, "bad ampc_camera (need fx, fy, pixel2meter > 0, resize_scale >= 1, depth_max > depth_min)");
    h->cam = *cam;
    h->tab_rows = h->tab_cols = 0;
    return AMPC_OK;
}

// cv::resize's INTER_LINEAR sampling tables for (rows, cols) -> (H, W), computed as OpenCV
// does: fx = (float)((dx + 0.5) * scale - 0.5); sx = floor(fx); fx -= sx
static int depth_tables(ampc_handle *h, int rows, int cols, int H, int W, cudaStream_t st, DepthGeom *g) {
    const size_t bytes = (size_t)(3 * W + 4 * H) * 4;
    if (h->tab_rows != rows || h->tab_cols != cols || h->tab_scale != h->cam.resize_scale) {
        std::vector<int32_t> tab(3 * W + 4 * H);
        int32_t *xofs = tab.data(), *y0 = tab.data() + 3 * W, *y1 = y0 + H;
        float *xw0 = reinterpret_cast<float *>(tab.data() + W), *xw1 = xw0 + W;
        float *yw0 = reinterpret_cast<float *>(y1 + H), *yw1 = yw0 + H;
        const double sx = 1.0 / ((double)W / cols), sy = 1.0 / ((double)H / rows);
        for (int d = 0; d < W; ++d) {
            float f = (float)((d + 0.5) * sx - 0.5);
            int s = (int)std::floor(f);
            f -= (float)s;
            if (s < 0) s = 0, f = 0.f;
            xofs[d] = s >= cols - 1 ? cols - 1 : s;
            xw0[d] = 1.f - f;
            xw1[d] = s >= cols - 1 ? -1.f : f; // right border: the source sample alone
        }
        for (int d = 0; d < H; ++d) {
            float f = (float)((d + 0.5) * sy - 0.5);
            const int s = (int)std::floor(f);
            f -= (float)s;
            y0[d] = s < 0 ? 0 : (s > rows - 1 ? rows - 1 : s);
            y1[d] = s + 1 < 0 ? 0 : (s + 1 > rows - 1 ? rows - 1 : s + 1);
            yw0[d] = 1.f - f;
            yw1[d] = f;
        }
        CK(cudaStreamSynchronize(st)); // the previous tables may still be in use
        CK(h->depth_tab.reserve(bytes));
        CK(cudaMemcpy(h->depth_tab.p, tab.data(), bytes, cudaMemcpyHostToDevice));
        h->tab_rows = rows, h->tab_cols = cols, h->tab_scale = h->cam.resize_scale;
    }
    int32_t *base = h->depth_tab.as<int32_t>();
    g->xofs = base;
    g->xw0 = reinterpret_cast<float *>(base + W);
    g->xw1 = g->xw0 + W;
    g->y0 = base + 3 * W;
    g->y1 = g->y0 + H;
    g->yw0 = reinterpret_cast<float *>(base + 3 * W + 2 * H);
    g->yw1 = g->yw0 + H;
    return AMPC_OK;
}

int ampc_depth_set_batch_dev(ampc_handle *h, int32_t first_scene, int32_t n_scenes, const void *depth_dev,
                             int32_t dtype, int32_t rows, int32_t cols, int64_t row_stride_bytes,
                             int64_t image_stride_bytes, const double *T_obstacle_dev, const double *T_edge_dev,
                             void *stream) {
    if (!h) return AMPC_ERR_INVALID;
    int rc = check_kind(h, AMPC_CLOUD_OBSTACLE);
    if (rc) return rc;
    if (n_scenes < 1 || first_scene < 0 || first_scene + n_scenes > h->cfg.max_scenes)
        return fail(h, AMPC_ERR_CAPACITY, "scene range exceeds ampc_config.max_scenes");
    if (!depth_dev || !T_obstacle_dev) return fail(h, AMPC_ERR_INVALID, "null depth image or transform");
    if (dtype != AMPC_DEPTH_F32 && dtype != AMPC_DEPTH_U16)
        return fail(h, AMPC_ERR_UNSUPPORTED, "depth dtype must be AMPC_DEPTH_F32 or AMPC_DEPTH_U16"); // :96-103
    const int esz = dtype == AMPC_DEPTH_U16 ? 2 : 4;
    if (rows < 1 || cols < 1 || row_stride_bytes < (int64_t)cols * esz || (row_stride_bytes % esz) ||
        (n_scenes > 1 && image_stride_bytes < row_stride_bytes * rows) || (image_stride_bytes % esz))
        return fail(h, AMPC_ERR_INVALID, "bad depth image geometry");
    const int W = (int)(cols / h->cam.resize_scale), H = (int)(rows / h->cam.resize_scale); // :106-107
    if (W < 1 || H < 1) return fail(h, AMPC_ERR_INVALID, "resize_scale leaves no pixels");
    if ((int64_t)W * H > h->slot_points[0])
        return fail(h, AMPC_ERR_CAPACITY, "resized image has more pixels than ampc_config.max_points");
    const bool with_edge = h->slot_points[1] > 0;
    CK(cudaSetDevice(h->cfg.device));
    cudaStream_t st = (cudaStream_t)stream;
    DepthGeom g{};
    g.fx = h->cam.fx / h->cam.resize_scale, g.fy = h->cam.fy / h->cam.resize_scale;
    g.cx = h->cam.cx / h->cam.resize_scale, g.cy = h->cam.cy / h->cam.resize_scale;
    g.p2m = h->cam.pixel2meter, g.dmin = h->cam.depth_min, g.dmax = h->cam.depth_max;
    g.range = h->cam.depth_max - h->cam.depth_min;
    g.rows = rows, g.cols = cols, g.H = H, g.W = W;
    g.is_u16 = dtype == AMPC_DEPTH_U16, g.identity = (W == cols && H == rows);
    g.row_stride = row_stride_bytes, g.image_stride = image_stride_bytes;
    rc = depth_tables(h, rows, cols, H, W, st, &g);
    if (rc) return rc;
    const size_t npx = (size_t)H * W;
    if (h->depth_scratch.bytes < (size_t)n_scenes * npx * 4) {
        CK(cudaStreamSynchronize(st));
        CK(h->depth_scratch.reserve((size_t)n_scenes * npx * 4));
    }
    if (h->depth_flag.p) // an overflow reported by (or left behind by) an earlier call is not this call's
        CK(cudaMemsetAsync(h->depth_flag.p, 0, 4, st));
    if (!h->depth_flag.p) {
        CK(h->depth_flag.reserve(4));
        CK(cudaMemset(h->depth_flag.p, 0, 4));
    }
    unsigned char *infl = h->depth_scratch.as<unsigned char>();
    unsigned char *er = infl + (size_t)n_scenes * npx;
    unsigned short *mag = reinterpret_cast<unsigned short *>(er + (size_t)n_scenes * npx);
    int slot;
    rc = prof_begin(h, SEC_INDEX, st, &slot);
    if (rc) return rc;
    depth_obstacle_kernel<<<n_scenes, DC_THREADS, 0, st>>>(
        g, static_cast<const unsigned char *>(depth_dev), T_obstacle_dev, h->cloud[0].as<float4>(),
        h->counts[0].as<int32_t>(), h->layout[0].as<int32_t>(), infl, h->depth_flag.as<int32_t>(),
        h->slot_points[0], first_scene);
    h->launches++;
    CK(cudaGetLastError());
    if (with_edge) {
        edge_grad_kernel<<<dim3((W + EG_BX - 1) / EG_BX, (H + EG_BY - 1) / EG_BY, n_scenes), dim3(EG_BX, EG_BY), 0, st>>>(
            H, W, infl, er, mag);
        h->launches++;
        CK(cudaGetLastError());
        edge_cloud_kernel<<<n_scenes, DC_THREADS, 0, st>>>(
            g, T_edge_dev ? T_edge_dev : T_obstacle_dev, er, mag, h->cloud[1].as<float4>(),
            h->counts[1].as<int32_t>(), h->layout[1].as<int32_t>(), h->counts[0].as<int32_t>(),
            h->depth_flag.as<int32_t>(), h->slot_points[1], first_scene);
        h->launches++;
        CK(cudaGetLastError());
    }
    rc = prof_end(h, slot, st);
    if (rc) return rc;
    rc = launch_index(h, AMPC_CLOUD_OBSTACLE, first_scene, n_scenes, st, false, W);
    if (rc || !with_edge) return rc;
    return launch_index(h, AMPC_CLOUD_EDGE, first_scene, n_scenes, st);
}

int ampc_depth_set_batch(ampc_handle *h, int32_t first_scene, int32_t n_scenes, const void *depth_host,
                         int32_t dtype, int32_t rows, int32_t cols, int64_t row_stride_bytes,
                         int64_t image_stride_bytes, const double *T_obstacle, const double *T_edge) {
    if (!h) return AMPC_ERR_INVALID;
    if (!depth_host || !T_obstacle || n_scenes < 1 || rows < 1 || row_stride_bytes < 1)
        return fail(h, AMPC_ERR_INVALID, "null depth image or transform");
    if (n_scenes == 1) image_stride_bytes = row_stride_bytes * rows;
    if (image_stride_bytes < row_stride_bytes * rows)
        return fail(h, AMPC_ERR_INVALID, "bad depth image geometry");
    CK(cudaSetDevice(h->cfg.device));
    cudaStream_t st = h->stream;
    const size_t img_bytes = (size_t)image_stride_bytes * n_scenes, t_bytes = (size_t)n_scenes * 16 * 8;
    CK(cudaStreamSynchronize(st));
    CK(h->depth_stage.reserve(img_bytes));
    CK(h->depth_T.reserve(2 * t_bytes));
    CK(cudaMemcpyAsync(h->depth_stage.p, depth_host, img_bytes, cudaMemcpyHostToDevice, st));
    double *Td = h->depth_T.as<double>();
    CK(cudaMemcpyAsync(Td, T_obstacle, t_bytes, cudaMemcpyHostToDevice, st));
    if (T_edge) CK(cudaMemcpyAsync(Td + (size_t)n_scenes * 16, T_edge, t_bytes, cudaMemcpyHostToDevice, st));
    int rc = ampc_depth_set_batch_dev(h, first_scene, n_scenes, h->depth_stage.p, dtype, rows, cols,
                                      row_stride_bytes, image_stride_bytes, Td,
                                      T_edge ? Td + (size_t)n_scenes * 16 : nullptr, st);
    if (rc) return rc;
    int32_t flag = 0;
    CK(cudaMemcpyAsync(&flag, h->depth_flag.p, 4, cudaMemcpyDeviceToHost, st));
    CK(cudaStreamSynchronize(st));
    if (flag) {
        CK(cudaMemset(h->depth_flag.p, 0, 4));
        return fail(h, AMPC_ERR_CAPACITY, "a cloud built from a depth image exceeded its slot capacity and was truncated");
    }
    return AMPC_OK;
}

int ampc_cloud_get(ampc_handle *h, int32_t scene, int32_t kind, void *xyz16_out, int32_t max_points,
                   int32_t *n_out) {
    if (!h || !n_out) return AMPC_ERR_INVALID;
    int rc = ampc_cloud_count(h, scene, kind, n_out);
    if (rc) return rc;
    if (*n_out > max_points || (!xyz16_out && *n_out > 0))
        return fail(h, AMPC_ERR_CAPACITY, "output buffer smaller than the cloud");
    if (*n_out > 0)
        CK(cudaMemcpy(xyz16_out, h->cloud[kind].as<float4>() + (int64_t)scene * h->slot_points[kind],
                      (size_t)*n_out * 16, cudaMemcpyDeviceToHost));
    return AMPC_OK;
}

int ampc_profile_enable(ampc_handle *h, int on) {
    if (!h) return AMPC_ERR_INVALID;
    CK(cudaSetDevice(h->cfg.device));
    if (on && h->prof_ev.empty()) {
        h->prof_ev.resize(2 * PROF_RING);
        h->prof_sec.resize(PROF_RING);
        for (auto &e : h->prof_ev) CK(cudaEventCreate(&e));
    }
    int rc = prof_drain(h);
    if (rc) return rc;
    h->prof = on != 0;
    for (int i = 0; i < 3; ++i) {
        h->prof_ms[i] = 0;
        h->prof_n[i] = 0;
    }
    return AMPC_OK;
}

int ampc_profile_get(ampc_handle *h, double ms_total[3], int64_t launches[3]) {
    if (!h) return AMPC_ERR_INVALID;
    CK(cudaSetDevice(h->cfg.device));
    int rc = prof_drain(h);
    if (rc) return rc;
    for (int i = 0; i < 3; ++i) {
        if (ms_total) ms_total[i] = h->prof_ms[i];
        if (launches) launches[i] = h->prof_n[i];
    }
    return AMPC_OK;
}

int ampc_round_batch(ampc_handle *h, int32_t B, const int32_t *scene_of, const double *x0,
                     const double *ref, const double *pos_x, double speed, double safety_distance,
                     double *w_inout, ampc_solve_info *info_out, int32_t *need_replan_out) {
    int rc = check_batch(h, B);
    if (rc) return rc;
    if (!x0 || !ref || !w_inout) return fail(h, AMPC_ERR_INVALID, "null buffer");
    if (scene_of)
        for (int b = 0; b < B; ++b)
            if (scene_of[b] < 0 || scene_of[b] >= h->cfg.max_scenes)
                return fail(h, AMPC_ERR_CAPACITY, "scene_of entry out of range");
    CK(cudaSetDevice(h->cfg.device));
    cudaStream_t st = h->stream;
    const size_t N = h->cfg.N;
    CK(cudaMemcpyAsync(h->x0.p, x0, (size_t)B * 10 * 8, cudaMemcpyHostToDevice, st));
    CK(cudaMemcpyAsync(h->ref.p, ref, (size_t)B * N * 10 * 8, cudaMemcpyHostToDevice, st));
    CK(cudaMemcpyAsync(h->w.p, w_inout, (size_t)B * h->n_w * 8, cudaMemcpyHostToDevice, st));
    if (scene_of) CK(cudaMemcpyAsync(h->scene_of.p, scene_of, (size_t)B * 4, cudaMemcpyHostToDevice, st));
    if (pos_x) CK(cudaMemcpyAsync(h->posx.p, pos_x, (size_t)B * 8, cudaMemcpyHostToDevice, st));
    rc = ampc_round_batch_dev(h, B, scene_of ? h->scene_of.as<int32_t>() : nullptr, h->x0.as<double>(),
                              h->ref.as<double>(), pos_x ? h->posx.as<double>() : nullptr, speed,
                              safety_distance, h->w.as<double>(),
                              reinterpret_cast<ampc_solve_info *>(h->info.p),
                              need_replan_out ? h->replan.as<int32_t>() : nullptr, st);
    if (rc) return rc;
    CK(cudaMemcpyAsync(w_inout, h->w.p, (size_t)B * h->n_w * 8, cudaMemcpyDeviceToHost, st));
    if (info_out)
        CK(cudaMemcpyAsync(info_out, h->info.p, (size_t)B * sizeof(SolveOut), cudaMemcpyDeviceToHost, st));
    if (need_replan_out)
        CK(cudaMemcpyAsync(need_replan_out, h->replan.p, (size_t)B * 4, cudaMemcpyDeviceToHost, st));
    CK(cudaStreamSynchronize(st));
    return AMPC_OK;
}

// ---- one control tick: up to max_rounds rounds with the reference's early exit ----------
int ampc_tick_batch_dev(ampc_handle *h, int32_t B, const int32_t *scene_of_dev, const double *x0_dev,
                        double *ref_inout_dev, const double *pos_x_dev, double speed,
                        double safety_distance, int32_t max_rounds, double *w_inout_dev,
                        ampc_solve_info *info_dev, int32_t *rounds_dev, int32_t *is_safety_dev,
                        void *stream) {
    int rc = check_batch(h, B);
    if (rc) return rc;
    rc = check_kind(h, AMPC_CLOUD_OBSTACLE);
    if (rc) return rc;
    if (!x0_dev || !ref_inout_dev || !w_inout_dev) return fail(h, AMPC_ERR_INVALID, "null buffer");
    if (h->cfg.K < 1 || max_rounds < 1) return fail(h, AMPC_ERR_INVALID, "tick needs K >= 1 and max_rounds >= 1");
    if (!scene_of_dev && B > h->cfg.max_scenes)
        return fail(h, AMPC_ERR_CAPACITY, "identity scene map needs B <= max_scenes");
    CK(cudaSetDevice(h->cfg.device));
    cudaStream_t st = (cudaStream_t)stream;
    const int N = h->cfg.N, K = h->cfg.K;
    const bool have_edge = h->slot_points[AMPC_CLOUD_EDGE] > 0;
    const size_t Bs = (size_t)B;
    CK(h->tk_active.reserve(Bs * 4));
    CK(h->tk_q0.reserve(Bs * 24));
    CK(h->tk_d1.reserve(Bs * 8));
    CK(h->tk_c1.reserve(Bs * 4));
    CK(h->tk_ep.reserve(Bs * 24));
    CK(h->tk_ec.reserve(Bs * 4));
    CK(h->tk_ed.reserve(Bs * 8));
    CK(h->tk_safe.reserve(Bs * 4));
    CK(h->tk_rounds.reserve(Bs * 4));
    int32_t *active = h->tk_active.as<int32_t>();
    int32_t *safe = is_safety_dev ? is_safety_dev : h->tk_safe.as<int32_t>();
    int32_t *rounds = rounds_dev ? rounds_dev : h->tk_rounds.as<int32_t>();
    int32_t *replan = h->replan.as<int32_t>();
    SolveOut *info = info_dev ? reinterpret_cast<SolveOut *>(info_dev) : h->info.as<SolveOut>();
    double *prefix = h->prefix.as<double>();
    const int tb = 128, gb = (B + tb - 1) / tb;
    tick_init_kernel<<<gb, tb, 0, st>>>(B, active, rounds, safe);
    h->launches++;
    for (int iter = 0; iter < max_rounds; ++iter) {
        // PlanWapionts: 1-NN of waypoint 0 on the Obstacle cloud, then on the Edge cloud
        tick_site0_kernel<<<gb, tb, 0, st>>>(B, N, ref_inout_dev, h->tk_q0.as<double>());
        h->launches++;
        rc = launch_knn(h, AMPC_CLOUD_OBSTACLE, B, scene_of_dev, h->tk_q0.as<double>(), 1, 1, nullptr,
                        h->tk_d1.as<double>(), h->tk_c1.as<int32_t>(), nullptr, 0, 0, st, active);
        if (rc) return rc;
        if (have_edge) {
            rc = launch_knn(h, AMPC_CLOUD_EDGE, B, scene_of_dev, h->tk_q0.as<double>(), 1, 1, nullptr,
                            h->tk_ed.as<double>(), h->tk_ec.as<int32_t>(), h->tk_ep.as<double>(), 3, 3, st, active);
            if (rc) return rc;
        }
        tick_plan_kernel<<<gb, tb, 0, st>>>(B, N, active, h->tk_d1.as<double>(), h->tk_c1.as<int32_t>(),
                                            h->tk_ep.as<double>(), h->tk_ec.as<int32_t>(), have_edge,
                                            safety_distance, ref_inout_dev, safe);
        h->launches++;
        // ProcessWaypoints + GetRefStates
        pack_prefix_kernel<<<B, 64, 0, st>>>(B, N, K, x0_dev, ref_inout_dev, pos_x_dev, speed, N * h->cfg.dt,
                                             prefix, h->n_prefix, h->queries.as<double>());
        h->launches++;
        rc = launch_knn(h, AMPC_CLOUD_OBSTACLE, B, scene_of_dev, h->queries.as<double>(), N, K,
                        h->knn_idx.as<int32_t>(), h->knn_d2.as<double>(), h->knn_cnt.as<int32_t>(),
                        prefix + 10 + 10 * N, h->n_prefix, 3 * K, st, active);
        if (rc) return rc;
        tick_replan_kernel<<<gb, tb, 0, st>>>(B, N, K, active, h->knn_d2.as<double>(), h->knn_cnt.as<int32_t>(),
                                              safety_distance, replan);
        h->launches++;
        tick_gate_kernel<<<gb, tb, 0, st>>>(B, iter, replan, safe, active, rounds);
        h->launches++;
        // Solve + hand the solution over as the next reference path / query sites
        rc = launch_solve(h, B, prefix, w_inout_dev, info, st, active);
        if (rc) return rc;
        tick_ref_update_kernel<<<B, 64, 0, st>>>(B, N, active, w_inout_dev, ref_inout_dev);
        h->launches++;
        CK(cudaGetLastError());
    }
    return AMPC_OK;
}

int ampc_tick_batch(ampc_handle *h, int32_t B, const int32_t *scene_of, const double *x0, double *ref_inout,
                    const double *pos_x, double speed, double safety_distance, int32_t max_rounds,
                    double *w_inout, ampc_solve_info *info_out, int32_t *rounds_out, int32_t *is_safety_out) {
    int rc = check_batch(h, B);
    if (rc) return rc;
    if (!x0 || !ref_inout || !w_inout) return fail(h, AMPC_ERR_INVALID, "null buffer");
    if (scene_of)
        for (int b = 0; b < B; ++b)
            if (scene_of[b] < 0 || scene_of[b] >= h->cfg.max_scenes)
                return fail(h, AMPC_ERR_CAPACITY, "scene_of entry out of range");
    CK(cudaSetDevice(h->cfg.device));
    cudaStream_t st = h->stream;
    const size_t N = h->cfg.N;
    CK(h->tk_safe.reserve((size_t)B * 4));
    CK(h->tk_rounds.reserve((size_t)B * 4));
    CK(cudaMemcpyAsync(h->x0.p, x0, (size_t)B * 10 * 8, cudaMemcpyHostToDevice, st));
    CK(cudaMemcpyAsync(h->ref.p, ref_inout, (size_t)B * N * 10 * 8, cudaMemcpyHostToDevice, st));
    CK(cudaMemcpyAsync(h->w.p, w_inout, (size_t)B * h->n_w * 8, cudaMemcpyHostToDevice, st));
    if (scene_of) CK(cudaMemcpyAsync(h->scene_of.p, scene_of, (size_t)B * 4, cudaMemcpyHostToDevice, st));
    if (pos_x) CK(cudaMemcpyAsync(h->posx.p, pos_x, (size_t)B * 8, cudaMemcpyHostToDevice, st));
    rc = ampc_tick_batch_dev(h, B, scene_of ? h->scene_of.as<int32_t>() : nullptr, h->x0.as<double>(),
                             h->ref.as<double>(), pos_x ? h->posx.as<double>() : nullptr, speed, safety_distance,
                             max_rounds, h->w.as<double>(), reinterpret_cast<ampc_solve_info *>(h->info.p),
                             h->tk_rounds.as<int32_t>(), h->tk_safe.as<int32_t>(), st);
    if (rc) return rc;
    CK(cudaMemcpyAsync(w_inout, h->w.p, (size_t)B * h->n_w * 8, cudaMemcpyDeviceToHost, st));
    CK(cudaMemcpyAsync(ref_inout, h->ref.p, (size_t)B * N * 10 * 8, cudaMemcpyDeviceToHost, st));
    if (info_out)
        CK(cudaMemcpyAsync(info_out, h->info.p, (size_t)B * sizeof(SolveOut), cudaMemcpyDeviceToHost, st));
    if (rounds_out)
        CK(cudaMemcpyAsync(rounds_out, h->tk_rounds.p, (size_t)B * 4, cudaMemcpyDeviceToHost, st));
    if (is_safety_out)
        CK(cudaMemcpyAsync(is_safety_out, h->tk_safe.p, (size_t)B * 4, cudaMemcpyDeviceToHost, st));
    CK(cudaStreamSynchronize(st));
    return AMPC_OK;
}

int ampc_last_prefix_dev(ampc_handle *h, const double **p) {
    if (!h || !p) return AMPC_ERR_INVALID;
    *p = h->prefix.as<double>();
    return AMPC_OK;
}

// ---- best-of-G ----------------------------------------------------------------
int ampc_guess_round_batch_dev(ampc_handle *h, int32_t n_scenes, int32_t G, const double *x0_dev,
                               const double *ref_dev, const double *pos_x_dev, double speed,
                               double *w_inout_dev, ampc_solve_info *info_dev, int32_t *argmin_dev,
                               double *best_cost_dev, void *stream) {
    if (!h) return AMPC_ERR_INVALID;
    if (n_scenes < 1 || G < 1 || G > KNN_KMAX) return fail(h, AMPC_ERR_INVALID, "need n_scenes >= 1 and 1 <= G <= 32");
    int rc = check_batch(h, n_scenes * G);
    if (rc) return rc;
    if ((rc = check_kind(h, AMPC_CLOUD_OBSTACLE)) || (rc = check_kind(h, AMPC_CLOUD_EDGE))) return rc;
    if (n_scenes > h->cfg.max_scenes) return fail(h, AMPC_ERR_CAPACITY, "n_scenes exceeds ampc_config.max_scenes");
    if (!x0_dev || !ref_dev || !w_inout_dev) return fail(h, AMPC_ERR_INVALID, "null buffer");
    if (h->cfg.K < 1) return fail(h, AMPC_ERR_INVALID, "round needs K >= 1");
    CK(cudaSetDevice(h->cfg.device));
    cudaStream_t st = (cudaStream_t)stream;
    const int N = h->cfg.N, K = h->cfg.K, Qp = N - 1 + G, B = n_scenes * G;
    CK(h->gs_sites.reserve((size_t)n_scenes * Qp * 3 * 8));
    CK(h->gs_nb.reserve((size_t)n_scenes * Qp * K * 3 * 8));
    CK(h->gs_epts.reserve((size_t)n_scenes * G * 3 * 8));
    CK(h->gs_ecnt.reserve((size_t)n_scenes * 4));
    CK(h->gs_q0.reserve((size_t)n_scenes * 3 * 8));
    // (1) the G nearest Edge points of waypoint 0 of every scene
    tick_site0_kernel<<<(n_scenes + 127) / 128, 128, 0, st>>>(n_scenes, N, ref_dev, h->gs_q0.as<double>());
    h->launches++;
    CK(cudaGetLastError());
    int slot;
    if ((rc = prof_begin(h, SEC_KNN, st, &slot))) return rc;
    rc = launch_knn(h, AMPC_CLOUD_EDGE, n_scenes, nullptr, h->gs_q0.as<double>(), 1, G, h->knn_idx.as<int32_t>(),
                    h->knn_d2.as<double>(), h->gs_ecnt.as<int32_t>(), h->gs_epts.as<double>(), (int64_t)G * 3,
                    (int64_t)G * 3, st);
    if (rc) return rc;
    // (2) one Obstacle query per shared waypoint and one per guess: N-1+G per scene, not N*G
    guess_sites_kernel<<<n_scenes, 64, 0, st>>>(n_scenes, N, G, ref_dev, h->gs_epts.as<double>(),
                                                h->gs_ecnt.as<int32_t>(), h->gs_sites.as<double>());
    h->launches++;
    CK(cudaGetLastError());
    rc = launch_knn(h, AMPC_CLOUD_OBSTACLE, n_scenes, nullptr, h->gs_sites.as<double>(), Qp, K,
                    h->knn_idx.as<int32_t>(), h->knn_d2.as<double>(), h->knn_cnt.as<int32_t>(), h->gs_nb.as<double>(),
                    (int64_t)Qp * K * 3, (int64_t)K * 3, st);
    if (rc) return rc;
    if ((rc = prof_end(h, slot, st))) return rc;
    // (3) prefixes of the n_scenes x G instances, (4) solve, (5) best guess of every scene
    double *prefix = h->prefix.as<double>();
    guess_pack_kernel<<<B, 64, 0, st>>>(n_scenes, N, K, G, x0_dev, ref_dev, pos_x_dev, speed, N * h->cfg.dt,
                                        h->gs_sites.as<double>(), h->gs_nb.as<double>(), prefix, h->n_prefix);
    h->launches++;
    CK(cudaGetLastError());
    SolveOut *info = info_dev ? reinterpret_cast<SolveOut *>(info_dev) : h->info.as<SolveOut>();
    if ((rc = prof_begin(h, SEC_SOLVE, st, &slot))) return rc;
    rc = launch_solve(h, B, prefix, w_inout_dev, info, st);
    if (rc) return rc;
    if ((rc = prof_end(h, slot, st))) return rc;
    if (argmin_dev && best_cost_dev) {
        best_of_kernel<<<(n_scenes + 127) / 128, 128, 0, st>>>(n_scenes, G, info, argmin_dev, best_cost_dev);
        h->launches++;
        CK(cudaGetLastError());
    }
    return AMPC_OK;
}

int ampc_guess_round_batch(ampc_handle *h, int32_t n_scenes, int32_t G, const double *x0, const double *ref,
                           const double *pos_x, double speed, double *w_inout, ampc_solve_info *info_out,
                           int32_t *argmin_out, double *best_cost_out) {
    if (!h) return AMPC_ERR_INVALID;
    if (n_scenes < 1 || G < 1) return fail(h, AMPC_ERR_INVALID, "need n_scenes >= 1 and G >= 1");
    int rc = check_batch(h, n_scenes * G);
    if (rc) return rc;
    if (!x0 || !ref || !w_inout) return fail(h, AMPC_ERR_INVALID, "null buffer");
    CK(cudaSetDevice(h->cfg.device));
    cudaStream_t st = h->stream;
    const int N = h->cfg.N, B = n_scenes * G;
    CK(cudaMemcpyAsync(h->x0.p, x0, (size_t)n_scenes * 10 * 8, cudaMemcpyHostToDevice, st));
    CK(cudaMemcpyAsync(h->ref.p, ref, (size_t)n_scenes * N * 10 * 8, cudaMemcpyHostToDevice, st));
    if (pos_x) CK(cudaMemcpyAsync(h->posx.p, pos_x, (size_t)n_scenes * 8, cudaMemcpyHostToDevice, st));
    CK(cudaMemcpyAsync(h->w.p, w_inout, (size_t)B * h->n_w * 8, cudaMemcpyHostToDevice, st));
    rc = ampc_guess_round_batch_dev(h, n_scenes, G, h->x0.as<double>(), h->ref.as<double>(),
                                    pos_x ? h->posx.as<double>() : nullptr, speed, h->w.as<double>(),
                                    reinterpret_cast<ampc_solve_info *>(h->info.p), h->bo_arg.as<int32_t>(),
                                    h->bo_cost.as<double>(), st);
    if (rc) return rc;
    CK(cudaMemcpyAsync(w_inout, h->w.p, (size_t)B * h->n_w * 8, cudaMemcpyDeviceToHost, st));
    if (info_out) CK(cudaMemcpyAsync(info_out, h->info.p, (size_t)B * sizeof(SolveOut), cudaMemcpyDeviceToHost, st));
    if (argmin_out) CK(cudaMemcpyAsync(argmin_out, h->bo_arg.p, (size_t)n_scenes * 4, cudaMemcpyDeviceToHost, st));
    if (best_cost_out) CK(cudaMemcpyAsync(best_cost_out, h->bo_cost.p, (size_t)n_scenes * 8, cudaMemcpyDeviceToHost, st));
    CK(cudaStreamSynchronize(st));
    return AMPC_OK;
}

int ampc_measure_fp64_peak(ampc_handle *h, double *tflops_out) {
    if (!h || !tflops_out) return AMPC_ERR_INVALID;
    CK(cudaSetDevice(h->cfg.device));
    const int blocks = h->n_sm * 8, threads = 256, iters = 4096;
    DevBuf out;
    CK(out.reserve((size_t)blocks * threads * 8));
    cudaEvent_t e0, e1;
    CK(cudaEventCreate(&e0));
    CK(cudaEventCreate(&e1));
    float best = 1e30f;
    for (int rep = 0; rep < 4; ++rep) { // first repetition = warm-up
        CK(cudaEventRecord(e0, h->stream));
        fp64_fma_peak_kernel<<<blocks, threads, 0, h->stream>>>(out.as<double>(), iters, 0.999999, 1e-6);
        CK(cudaEventRecord(e1, h->stream));
        CK(cudaEventSynchronize(e1));
        float ms = 0;
        CK(cudaEventElapsedTime(&ms, e0, e1));
        if (rep > 0 && ms < best) best = ms;
    }
    h->launches += 4;
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    out.release();
    const double flop = 2.0 * 64.0 * iters * (double)blocks * threads; // 64 FMA per thread per iteration
    *tflops_out = flop / (best * 1e-3) / 1e12;
    return AMPC_OK;
}

int ampc_best_of_dev(ampc_handle *h, int32_t n_scenes, int32_t G, const ampc_solve_info *info_dev,
                     int32_t *argmin_dev, double *best_cost_dev, void *stream) {
    if (!h) return AMPC_ERR_INVALID;
    if (n_scenes < 1 || G < 1 || !info_dev || !argmin_dev || !best_cost_dev)
        return fail(h, AMPC_ERR_INVALID, "bad best_of arguments");
    CK(cudaSetDevice(h->cfg.device));
    best_of_kernel<<<(n_scenes + 127) / 128, 128, 0, (cudaStream_t)stream>>>(
        n_scenes, G, reinterpret_cast<const SolveOut *>(info_dev), argmin_dev, best_cost_dev);
    h->launches++;
    CK(cudaGetLastError());
    return AMPC_OK;
}

int ampc_best_of(ampc_handle *h, int32_t n_scenes, int32_t G, const ampc_solve_info *info,
                 int32_t *argmin_out, double *best_cost_out) {
    if (!h) return AMPC_ERR_INVALID;
    if (n_scenes < 1 || G < 1 || !info || !argmin_out || !best_cost_out)
        return fail(h, AMPC_ERR_INVALID, "bad best_of arguments");
    const int64_t B = (int64_t)n_scenes * G;
    if (B > h->cfg.max_batch) return fail(h, AMPC_ERR_CAPACITY, "n_scenes*G exceeds max_batch");
    CK(cudaSetDevice(h->cfg.device));
    cudaStream_t st = h->stream;
    CK(cudaMemcpyAsync(h->info.p, info, (size_t)B * sizeof(SolveOut), cudaMemcpyHostToDevice, st));
    int rc = ampc_best_of_dev(h, n_scenes, G, reinterpret_cast<const ampc_solve_info *>(h->info.p),
                              h->bo_arg.as<int32_t>(), h->bo_cost.as<double>(), st);
    if (rc) return rc;
    CK(cudaMemcpyAsync(argmin_out, h->bo_arg.p, (size_t)n_scenes * 4, cudaMemcpyDeviceToHost, st));
    CK(cudaMemcpyAsync(best_cost_out, h->bo_cost.p, (size_t)n_scenes * 8, cudaMemcpyDeviceToHost, st));
    CK(cudaStreamSynchronize(st));
    return AMPC_OK;
}

} // extern "C"
