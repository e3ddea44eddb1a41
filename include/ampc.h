/* ampc.h — C-ABI of the B200-native Avoid-MPC hot path (libampc.so).
 *
 * This is the drop-in boundary: plain C, opaque handle, caller-owned buffers,
 * int status returns, no exceptions and no torch/CUDA types in any signature
 * (device pointers and streams travel as void*).  Every entry point names the
 * reference interface it replaces; paths are relative to
 * roswrapper/ros/src/avoid_mpc/ of SJTU-ViSYS-team/Avoid-MPC.
 *
 * Two flavours of each batch call:
 *   ampc_xxx      host buffers in / out (copies inside the call, synchronous)
 *   ampc_xxx_dev  device buffers in / out, asynchronous on `stream`
 *
 * Sizes for a handle created with horizon N and K neighbours per stage:
 *   n_w      = 10 + 14 N                 decision vector [X_0,U_0,...,U_{N-1},X_N]
 *   n_prefix = 20 + 10 N + 3 K N         vecRefStates = [x0 | ref N*10 | obst N*K*3 | target]
 * (tools/mpc_obstacle_casadi.py:76-94,158,164,217; src/AvoidanceStateMachine.cpp:236-257)
 */
#ifndef AMPC_H
#define AMPC_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define AMPC_API_VERSION 1

/* status codes returned by every function */
enum {
    AMPC_OK = 0,
    AMPC_ERR_INVALID = 1,  /* bad argument (null, out of range, wrong size) */
    AMPC_ERR_CUDA = 2,     /* CUDA runtime error or no device: ampc_last_error() has the text */
    AMPC_ERR_CAPACITY = 3, /* batch / scene / point count exceeds what the handle was created for */
    AMPC_ERR_UNSUPPORTED = 4
};

/* per-instance solver status (ampc_solve_* `status_out`); the reference never
 * inspects IPOPT's return status (src/HighLvlMpc.cpp:116-122) */
enum {
    AMPC_SOLVE_CONVERGED = 0,
    AMPC_SOLVE_MAX_ITER = 1,
    AMPC_SOLVE_STALLED = 2, /* line search could not make progress */
    AMPC_SOLVE_NUMERIC = 3  /* NaN / inertia correction exhausted */
};

/* which per-scene cloud a call refers to (include/FrameKDMap.h:11-13:
 * Frame::pointCloud and Frame::edgeCloud) */
enum { AMPC_CLOUD_OBSTACLE = 0, AMPC_CLOUD_EDGE = 1 };

typedef struct ampc_handle ampc_handle;

typedef struct ampc_config {
    int32_t N;                /* horizon = int(T/dt)            (src/HighLvlMpc.cpp:9)            */
    int32_t K;                /* nearest_point_num              (config/mpc_parameters.yaml:5)    */
    double dt;                /* mpc_dt                         (config/mpc_parameters.yaml:2)    */
    int32_t max_batch;        /* most MPC instances per batch call                                 */
    int32_t max_scenes;       /* scene slots (one Obstacle + one Edge cloud each)                  */
    int32_t max_points;       /* capacity of one Obstacle cloud slot, points                       */
    int32_t max_edge_points;  /* capacity of one Edge cloud slot, points (0: no edge clouds)       */
    int32_t device;           /* CUDA device ordinal                                               */
} ampc_config;

/* Interior-point options (role of the ipopt.* dictionary, src/HighLvlMpc.cpp:17-23).
 * ampc_default_solver_opts() fills the values the parity tests run with. */
typedef struct ampc_solver_opts {
    double tol;        /* KKT tolerance                 (ipopt.tol)      default 1e-8 */
    int32_t max_iter;  /* iteration cap                 (ipopt.max_iter) default 100  */
    double mu_init;    /* initial barrier parameter                      default 0.1  */
    double bound_push; /* initial push of U off its bounds               default 1e-2 */
    double bound_frac; /*                                                default 1e-2 */
    double eps_min;    /* floor of the |v.n| smoothing, m/s              default 1e-5 */
    double eps_scale;  /* smoothing eps = max(eps_min, eps_scale*mu)     default 1.0  */
    double kappa_eps;  /* barrier sub-problem tolerance: mu is reduced once the barrier KKT
                          error is <= kappa_eps*mu (ipopt.barrier_tol_factor)  default 100 */
} ampc_solver_opts;

/* one record per instance written by the solve calls */
typedef struct ampc_solve_info {
    double cost;     /* NLP objective f at the returned point (un-smoothed) */
    double kkt_dual; /* | grad_U L |_inf  (x-stationarity holds exactly via the adjoint) */
    double kkt_compl;/* max slack*multiplier */
    double mu;       /* final barrier parameter */
    int32_t iters;
    int32_t status;  /* AMPC_SOLVE_* */
    int32_t n_reg;   /* iterations that needed inertia-correcting regularisation */
    int32_t n_backtrack;
} ampc_solve_info;

/* ---- lifetime ---------------------------------------------------------- */
/* replaces ObstacleAvoidanceMPC::ObstacleAvoidanceMPC(T, dt, soPath)  (src/HighLvlMpc.cpp:5-59)
 * and the FrameKDMap constructor; N and K were baked into the codegen .so there. */
int ampc_create(const ampc_config *cfg, ampc_handle **out);
void ampc_destroy(ampc_handle *h);
/* text of the last error on this handle (or of the last failed ampc_create when h == NULL) */
const char *ampc_last_error(const ampc_handle *h);
int ampc_api_version(void);

/* ---- parameters (src/HighLvlMpc.cpp:60-92) ------------------------------ */
int ampc_set_weights(ampc_handle *h, const double weights[25]);     /* SetupWeights        */
int ampc_set_tau(ampc_handle *h, const double tau[4]);              /* SetupTau            */
int ampc_set_gains(ampc_handle *h, const double gains[4]);          /* SetupGains (unused by the default dynamics, tools/mpc_obstacle_casadi.py:114-121) */
int ampc_set_radius(ampc_handle *h, double drone_radius);           /* SetDroneRadius      */
int ampc_set_accel_limits(ampc_handle *h, double a_min_z, double a_max_z, double a_max_xy,
                          double a_max_yaw_dot);                    /* SetDroneAccelLimits */
void ampc_default_solver_opts(ampc_solver_opts *o);
int ampc_set_solver_opts(ampc_handle *h, const ampc_solver_opts *o);
/* discrete dynamics X+ = Phi X + Gam U + gam the solver uses (RK4 x 4 of the
 * affine ODE, tools/mpc_obstacle_casadi.py:106-122,338-357); row-major. */
int ampc_get_dynamics(ampc_handle *h, double Phi[100], double Gam[40], double gam[10]);

/* ---- clouds: replaces KDTreeTwo::InitializeNew (include/kd_tree_two.h:75-106)
 * as called by FrameKDMap::AddVertex (src/FrameKDMap.cpp:44-47).  Points whose
 * x is NaN are dropped and the rest keep their order (kd_tree_two.h:99-101), so
 * neighbour indices refer to the filtered cloud exactly as in the reference.
 * xyz: records of `stride_bytes` (>= 12) starting with float x,y,z
 * (pcl::PointXYZ: stride 16). */
int ampc_cloud_set(ampc_handle *h, int32_t scene, int32_t kind, const void *xyz_host, int32_t n,
                   int32_t stride_bytes);
/* n_scenes clouds, scene s at xyz_host + s*scene_stride_bytes holding counts[s] points */
int ampc_cloud_set_batch(ampc_handle *h, int32_t kind, int32_t first_scene, int32_t n_scenes,
                         const void *xyz_host, const int32_t *counts, int64_t scene_stride_bytes,
                         int32_t stride_bytes);
/* same, source already in device memory (16-byte records only); async on stream */
int ampc_cloud_set_batch_dev(ampc_handle *h, int32_t kind, int32_t first_scene, int32_t n_scenes,
                             const void *xyz_dev, const int32_t *counts_host,
                             int64_t scene_stride_bytes, void *stream);
/* Optional layout hint for the clouds of `kind`.  Only pruning efficiency depends on it,
 * never the results (indices are always those of the caller's record order).  Takes effect
 * at the next cloud_set / cloud_index call.
 *   row_width >= 8           ORGANISED cloud stored row-major with `row_width` records per
 *                            image row (the resized depth image of FrameKDMap::ProcessDepth,
 *                            src/FrameKDMap.cpp:106-125): the index uses 8x8 image patches.
 *   AMPC_LAYOUT_UNORGANISED  (default) tiles are runs of 64 consecutive records; good when the
 *                            storage order is spatially coherent (scan lines, subsets of one).
 *   AMPC_LAYOUT_SORT         arbitrary storage order (what KDTreeTwo::InitializeNew accepts,
 *                            include/kd_tree_two.h:75-106): the index is built over a copy of
 *                            the cloud bucketed by Morton cell, original indices carried along
 *                            (costs three more passes over the cloud and a second slot buffer). */
#define AMPC_LAYOUT_UNORGANISED 0
#define AMPC_LAYOUT_SORT (-1)
int ampc_cloud_set_layout(ampc_handle *h, int32_t kind, int32_t row_width);
/* rebuild the index (NaN filter + tile boxes) of clouds already resident in the
 * handle's slots, e.g. after writing a new depth frame's points in place; async */
int ampc_cloud_index_dev(ampc_handle *h, int32_t kind, int32_t first_scene, int32_t n_scenes,
                         void *stream);
/* number of points held for (scene, kind) after the NaN filter (synchronises) */
int ampc_cloud_count(ampc_handle *h, int32_t scene, int32_t kind, int32_t *n_out);
/* copy the cloud of (scene, kind) back to the host as 16-byte float records (x, y, z, pad), in
 * index order: what KDTreeTwo::GetPointCloud().pts holds (include/kd_tree_two.h:134-136; read by
 * the key-frame outlier step, src/FrameKDMap.cpp:462-464).  Synchronises. */
int ampc_cloud_get(ampc_handle *h, int32_t scene, int32_t kind, void *xyz16_out, int32_t max_points,
                   int32_t *n_out);

/* ---- depth image -> both clouds: replaces FrameKDMap::ProcessDepth + BuildEdgeCloud
 * (src/FrameKDMap.cpp:76-130,176-214, incl. cv::resize / cv::erode / cv::Canny) followed by
 * the two InitializeNew calls of AddVertex (:44-47), for n_scenes depth images at once.
 * ampc_camera mirrors the perception block of config/mpc_parameters.yaml:58-66; fx..cy are the
 * full-resolution intrinsics (the library divides by resize_scale as the constructor does,
 * src/FrameKDMap.cpp:21-24).  Defaults = the shipped yaml. */
typedef struct ampc_camera {
    double fx, fy, cx, cy;
    double resize_scale; /* mParamDepthScale, >= 1 */
    double pixel2meter;
    double depth_min, depth_max; /* metres */
} ampc_camera;
int ampc_set_camera(ampc_handle *h, const ampc_camera *cam);
#define AMPC_DEPTH_F32 0 /* CV_32FC1 (the simulator's metres) */
#define AMPC_DEPTH_U16 1 /* CV_16UC1 */
/*   depth         n_scenes images, rows x cols, row_stride_bytes between rows,
 *                 image_stride_bytes between images (ignored for n_scenes == 1)
 *   T_obstacle    n_scenes x 16 doubles, row-major 4x4 camera->world transform applied to the
 *                 Obstacle points: mat4Twb * mParamTbc (:117-119)
 *   T_edge        same for the Edge points; the reference applies mCurFrame.Twc * mParamTbc,
 *                 i.e. the PREVIOUS frame's Twc times Tbc once more (:208-209) -- the caller
 *                 decides; NULL = T_obstacle
 * Fills the Obstacle slot (and the Edge slot when the handle has Edge capacity) of scenes
 * first_scene.. and builds their indices.  The Edge cloud of a scene whose Obstacle cloud came
 * out empty is empty (:125-127).  Needs rows/scale * cols/scale <= max_points; returns
 * AMPC_ERR_CAPACITY if an Edge cloud outgrew max_edge_points (it is truncated). */
int ampc_depth_set_batch(ampc_handle *h, int32_t first_scene, int32_t n_scenes, const void *depth_host,
                         int32_t dtype, int32_t rows, int32_t cols, int64_t row_stride_bytes,
                         int64_t image_stride_bytes, const double *T_obstacle, const double *T_edge);
/* same with the images and transforms already in device memory; async on stream */
int ampc_depth_set_batch_dev(ampc_handle *h, int32_t first_scene, int32_t n_scenes, const void *depth_dev,
                             int32_t dtype, int32_t rows, int32_t cols, int64_t row_stride_bytes,
                             int64_t image_stride_bytes, const double *T_obstacle_dev,
                             const double *T_edge_dev, void *stream);

/* ---- k-NN: replaces KDTreeTwo::SearchForNearest (include/kd_tree_two.h:108-133)
 * behind FrameKDMap::QueryNearest's current-frame path (src/FrameKDMap.cpp:254-275,
 * 339-345) for B instances x Q queries at once.
 *   scene_of[B]  scene slot of each instance (NULL: instance b uses scene b)
 *   queries      B*Q*3 doubles
 *   idx_out      B*Q*k int32, ascending (dist2, index); -1 beyond count
 *   dist2_out    B*Q*k doubles (SQUARED distances, as the reference); +inf beyond count
 *   pts_out      B*Q*k*3 doubles, neighbour coordinates widened from float; (1e4,1e4,1e4)
 *                beyond count (src/AvoidanceStateMachine.cpp:223-226).  May be NULL.
 *   count_out    B*Q int32: results per query = min(k, n) except 0 when n == k
 *                (the reference's quirk, kd_tree_two.h:117-124)
 * Exact: dist2 = ((dx*dx + dy*dy) + dz*dz) in double from float coordinates, each
 * operation rounded separately (include/nanoflann_two.hpp:590-599). k <= 32. */
int ampc_knn_batch(ampc_handle *h, int32_t kind, int32_t B, const int32_t *scene_of,
                   const double *queries, int32_t Q, int32_t k, int32_t *idx_out,
                   double *dist2_out, double *pts_out, int32_t *count_out);
int ampc_knn_batch_dev(ampc_handle *h, int32_t kind, int32_t B, const int32_t *scene_of_dev,
                       const double *queries_dev, int32_t Q, int32_t k, int32_t *idx_dev,
                       double *dist2_dev, double *pts_dev, int32_t *count_dev, void *stream);

/* ---- solve: replaces ObstacleAvoidanceMPC::Solve (src/HighLvlMpc.cpp:93-137)
 * for B instances.  p_prefix: B*n_prefix doubles (vecRefStates; the gains/tau/
 * weights/radius tail is appended from the handle as Solve does, :97-108).
 * w_inout: B*n_w doubles, warm start in (mNlpW0 role, :110,129; only its U part
 * is used, X is rolled out from x0), solution out.  u = w[10:14],
 * x0Array[i] = w[14i:14i+14] (:122-136).  info_out may be NULL. */
int ampc_solve_batch(ampc_handle *h, int32_t B, const double *p_prefix, double *w_inout,
                     ampc_solve_info *info_out);
int ampc_solve_batch_dev(ampc_handle *h, int32_t B, const double *p_prefix_dev, double *w_inout_dev,
                         ampc_solve_info *info_dev, void *stream);

/* ---- one round of the control tick for B instances: ProcessWaypoints +
 * GetRefStates + Solve (src/AvoidanceStateMachine.cpp:204-257,337) = Q=N k-NN
 * queries at the reference-path positions on the Obstacle cloud, prefix packing
 * with the target rule (:250-255) and the NLP solve.
 *   x0[B*10], ref[B*N*10], scene_of[B] (NULL: identity), pos_x[B] (mPos.x of the
 *   target rule; NULL: x0[0]), speed (mSpeed).
 *   need_replan_out[B] (may be NULL): nearest obstacle <= safety_distance or no
 *   neighbour at some waypoint (:228-231). */
int ampc_round_batch(ampc_handle *h, int32_t B, const int32_t *scene_of, const double *x0,
                     const double *ref, const double *pos_x, double speed, double safety_distance,
                     double *w_inout, ampc_solve_info *info_out, int32_t *need_replan_out);
int ampc_round_batch_dev(ampc_handle *h, int32_t B, const int32_t *scene_of_dev,
                         const double *x0_dev, const double *ref_dev, const double *pos_x_dev,
                         double speed, double safety_distance, double *w_inout_dev,
                         ampc_solve_info *info_dev, int32_t *need_replan_dev, void *stream);
/* ---- one control tick for B instances: the TASK loop of AvoidanceStateMachine::Step
 * (src/AvoidanceStateMachine.cpp:328-344) entirely on the device, no host round trip
 * between rounds.  Per round and per still-active instance: PlanWapionts (:259-281: if
 * waypoint 0 is within safety_distance of the Obstacle cloud it moves to the nearest Edge
 * point; no Edge point -> isSafety = 0), ProcessWaypoints + GetRefStates + Solve, then the
 * solved states become the next reference path and query sites (:338-342).  An instance
 * stops before the solve of round iter > 0 when it needs no replan and is safe (:333).
 *   ref_inout[B*N*10]  mRefPath after GetInitPath in; the last solved path out
 *   x0[B*10]           mVecStateQuad (the caller does the latency extrapolation, :183-203)
 *   rounds_out[B]      solves performed (1..max_rounds); is_safety_out[B]: 0 -> the caller
 *                      publishes the slow-down command (:347-349).  Both may be NULL.
 * u of PubCmd is w_inout[b*n_w + 10 .. 14). */
int ampc_tick_batch(ampc_handle *h, int32_t B, const int32_t *scene_of, const double *x0,
                    double *ref_inout, const double *pos_x, double speed, double safety_distance,
                    int32_t max_rounds, double *w_inout, ampc_solve_info *info_out,
                    int32_t *rounds_out, int32_t *is_safety_out);
int ampc_tick_batch_dev(ampc_handle *h, int32_t B, const int32_t *scene_of_dev, const double *x0_dev,
                        double *ref_inout_dev, const double *pos_x_dev, double speed,
                        double safety_distance, int32_t max_rounds, double *w_inout_dev,
                        ampc_solve_info *info_dev, int32_t *rounds_dev, int32_t *is_safety_dev,
                        void *stream);
/* device address of the packed prefixes of the last round (B*n_prefix doubles) */
int ampc_last_prefix_dev(ampc_handle *h, const double **p_prefix_dev);

/* ---- best-of-G reduction over the G initial guesses of each scene (Edge-tree
 * guesses, src/AvoidanceStateMachine.cpp:259-281 generalised): instances are laid
 * out scene-major (b = s*G + g).  argmin_out[s] = g of the lowest cost among
 * instances whose status is CONVERGED or MAX_ITER (ties: lowest g); -1 if none. */
int ampc_best_of(ampc_handle *h, int32_t n_scenes, int32_t G, const ampc_solve_info *info,
                 int32_t *argmin_out, double *best_cost_out);
int ampc_best_of_dev(ampc_handle *h, int32_t n_scenes, int32_t G, const ampc_solve_info *info_dev,
                     int32_t *argmin_dev, double *best_cost_dev, void *stream);

/* ---- Edge-tree initial guesses + best-of-G in one call (BASELINE config C2): PlanWapionts
 * (src/AvoidanceStateMachine.cpp:259-281) generalised from "the nearest Edge point" to the G
 * nearest, every guess solved, the cheapest kept.  Per scene s (scene slot s: Obstacle + Edge
 * cloud): guess g replaces waypoint 0 of ref_s by the g-th nearest Edge point of it (g = 0 is the
 * reference's move; a guess beyond the number of Edge points leaves the waypoint in place).
 * Instances are laid out scene-major, b = s*G + g.  Only stage 0's neighbours differ between the
 * guesses of a scene, so the Obstacle cloud is queried at N-1+G sites per scene, not N*G.
 *   x0[n_scenes*10], ref[n_scenes*N*10], pos_x[n_scenes] (NULL: x0[0]), speed: as ampc_round_batch
 *   w_inout[n_scenes*G*n_w]  warm start in / solutions out, info[n_scenes*G]
 *   argmin[n_scenes], best_cost[n_scenes]: as ampc_best_of (both NULL: skipped).  G <= 32. */
int ampc_guess_round_batch(ampc_handle *h, int32_t n_scenes, int32_t G, const double *x0, const double *ref,
                           const double *pos_x, double speed, double *w_inout, ampc_solve_info *info_out,
                           int32_t *argmin_out, double *best_cost_out);
int ampc_guess_round_batch_dev(ampc_handle *h, int32_t n_scenes, int32_t G, const double *x0_dev,
                               const double *ref_dev, const double *pos_x_dev, double speed,
                               double *w_inout_dev, ampc_solve_info *info_dev, int32_t *argmin_dev,
                               double *best_cost_dev, void *stream);

/* ---- instrumentation ---------------------------------------------------- */
/* FP64 FMA throughput of the handle's device in TFLOP/s, measured with a hand-written kernel of
 * independent dependent-FMA chains (2 flop per FMA): the denominator of the solve kernels'
 * roofline fraction.  Synchronises. */
int ampc_measure_fp64_peak(ampc_handle *h, double *tflops_out);
/* kernels launched by this handle since creation (for bench.py's gpu_launches) */
int64_t ampc_launch_count(const ampc_handle *h);
/* per-stage device time from CUDA events recorded on the caller's stream around
 * [0] the cloud index build, [1] the k-NN search, [2] the NLP solve.  enable(1)
 * resets the totals; get() waits for what was recorded and returns sums and counts. */
int ampc_profile_enable(ampc_handle *h, int on);
int ampc_profile_get(ampc_handle *h, double ms_total[3], int64_t launches[3]);
/* the handle's own stream (void* cudaStream_t) used by the host-buffer calls */
void *ampc_stream(ampc_handle *h);
int ampc_synchronize(ampc_handle *h);

#ifdef __cplusplus
}
#endif
#endif /* AMPC_H */
