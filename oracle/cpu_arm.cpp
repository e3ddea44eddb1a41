// TEST / BASELINE INFRASTRUCTURE ONLY (oracle/): the reference's CPU path for one control round,
// native and multi-threaded, as BASELINE.md section 2 asks (std::thread, one instance per thread
// at a time, -O3):  per instance
//   a1  KDTreeTwo::InitializeNew        -- the REFERENCE's own kd_tree_two.h + nanoflann_two.hpp,
//                                          compiled where they lie (FrameKDMap.cpp:44-47 rebuilds
//                                          the tree every depth frame)
//   a2/a6  N x SearchForNearest(K)      -- the reference's own code (AvoidanceStateMachine.cpp:204-235)
//   a7  GetRefStates packing            -- :236-257
//   a8  the NLP solve                   -- oracle/nlp_oracle.c (CasADi/IPOPT cannot be installed here)
// Built by oracle/Makefile into oracle/_ref/libampc_cpu_arm.so.  Used by bench.py's CPU legs only;
// the product never loads it.
#include "kd_tree_two.h" // the reference file, via -I<reference>/include

#include <atomic>
#include <chrono>
#include <cstdint>
#include <cstring>
#include <thread>
#include <vector>

extern "C" {
struct nlp_oracle_opts;
struct nlp_oracle_info_c { // layout of nlp_oracle_info (oracle/nlp_oracle.c)
    double cost;
    int32_t iters, status;
    double kkt_dual, kkt_primal, kkt_compl, mu, reg_last;
    int32_t n_reg, n_backtrack;
};
int nlp_oracle_solve(int N, int K, double dt, const double *p, double *w, const double lbu[4], const double ubu[4],
                     const nlp_oracle_opts *opt, nlp_oracle_info_c *info, double *lam_g_out);

// n instances, instance i uses scene i % n_distinct.  Returns wall seconds; per-stage seconds summed
// over instances in stage_s[3] = {tree build, k-NN, pack + solve}.
double cpu_arm_run(int n, int threads, int n_distinct, int64_t npts, const void *clouds16, const double *x0,
                   const double *ref, const double *tgt, const double *tail34, const double *W0, int N, int K,
                   double dt, const double *lb, const double *ub, const nlp_oracle_opts *opt, int32_t *status_out,
                   int32_t *iters_out, double *stage_s) {
    const int n_w = 10 + 14 * N, n_p = 20 + 10 * N + 3 * K * N + 34;
    std::atomic<int> next{0};
    std::vector<double> st((size_t)threads * 3, 0.0);
    auto worker = [&](int tid) {
        using clk = std::chrono::steady_clock;
        std::vector<double> p(n_p), w(n_w);
        for (;;) {
            const int i = next.fetch_add(1);
            if (i >= n) break;
            const int s = i % n_distinct;
            auto t0 = clk::now();
            auto cloud = std::make_shared<pcl::PointCloud<pcl::PointXYZ>>();
            cloud->points.resize((size_t)npts);
            std::memcpy(static_cast<void *>(cloud->points.data()),
                        static_cast<const char *>(clouds16) + (size_t)s * npts * 16, (size_t)npts * 16);
            KDTreeTwo<double> tree;
            tree.InitializeNew(cloud);
            auto t1 = clk::now();
            const double *r = ref + (size_t)s * N * 10;
            double *ob = p.data() + 10 + 10 * N;
            for (int q = 0; q < N; ++q) {
                tree.SearchForNearest(r[10 * q], r[10 * q + 1], r[10 * q + 2], K);
                const int m = (int)tree.closest_pts.size();
                for (int j = 0; j < K; ++j) {
                    double *o = ob + 3 * (K * q + j);
                    if (j < m)
                        o[0] = tree.closest_pts[j].x, o[1] = tree.closest_pts[j].y, o[2] = tree.closest_pts[j].z;
                    else
                        o[0] = o[1] = o[2] = 10000.0;
                }
            }
            auto t2 = clk::now();
            std::memcpy(p.data(), x0 + (size_t)s * 10, 80);
            std::memcpy(p.data() + 10, r, (size_t)N * 80);
            std::memcpy(p.data() + 10 + 10 * N + 3 * K * N, tgt + (size_t)s * 10, 80);
            std::memcpy(p.data() + 20 + 10 * N + 3 * K * N, tail34, 34 * 8);
            std::memcpy(w.data(), W0 + (size_t)s * n_w, (size_t)n_w * 8);
            nlp_oracle_info_c info;
            nlp_oracle_solve(N, K, dt, p.data(), w.data(), lb, ub, opt, &info, nullptr);
            auto t3 = clk::now();
            status_out[i] = info.status;
            iters_out[i] = info.iters;
            st[(size_t)tid * 3 + 0] += std::chrono::duration<double>(t1 - t0).count();
            st[(size_t)tid * 3 + 1] += std::chrono::duration<double>(t2 - t1).count();
            st[(size_t)tid * 3 + 2] += std::chrono::duration<double>(t3 - t2).count();
        }
    };
    const auto T0 = std::chrono::steady_clock::now();
    std::vector<std::thread> pool;
    for (int t = 0; t < threads; ++t) pool.emplace_back(worker, t);
    for (auto &t : pool) t.join();
    const double wall = std::chrono::duration<double>(std::chrono::steady_clock::now() - T0).count();
    if (stage_s) {
        stage_s[0] = stage_s[1] = stage_s[2] = 0.0;
        for (int t = 0; t < threads; ++t)
            for (int k = 0; k < 3; ++k) stage_s[k] += st[(size_t)t * 3 + k];
    }
    return wall;
}
} // extern "C"
