// Exact batched k-NN by streaming scan with tile-AABB pruning (sm_100a).
//
// Replaces KDTreeTwo::SearchForNearest / nanoflann findNeighbors+searchLevel
// (include/kd_tree_two.h:108-133, include/nanoflann_two.hpp:1563-1586,1729-1793)
// for B instances x Q queries in one launch.  For Q << Npts on a cloud that is
// rebuilt every depth frame (src/FrameKDMap.cpp:34-52) a single coalesced pass
// over the 16-byte point records IS the HBM roofline; no tree is built.
//
// One CTA per (instance, cloud segment).  Each warp streams tiles of 64 points
// (2 x 16-byte vector loads per lane, software-prefetched one tile ahead),
// reduces the tile's bounding box with shuffles, and lane q tests query q's
// exact lower bound against that query's current k-th best distance (shared by
// the CTA, tightened with atomicMin).  Only surviving (tile, query) pairs pay
// for the FP64 distances; survivors are inserted into the warp's sorted top-k
// list in shared memory.  Distances use the reference's arithmetic exactly:
// dist2 = ((dx*dx + dy*dy) + dz*dz), double, every operation rounded separately
// (nanoflann_two.hpp:590-599), so indices AND squared distances are bit-exact.
// Result order is canonical (dist2, index); see DESIGN.md for ties.
#pragma once
#include "common.cuh"

namespace ampc {

constexpr int KNN_THREADS = 256;
constexpr int KNN_WARPS = KNN_THREADS / 32;
constexpr int KNN_TILE = 64; // points per warp iteration
constexpr int KNN_KMAX = 32;

struct KnnParams {
    const float4 *clouds;    // slot s at clouds + s*slot_points
    const int32_t *counts;   // points held in each slot (after the NaN filter)
    int64_t slot_points;
    const int32_t *scene_of; // [B] or nullptr (identity)
    const double *queries;   // [B][Q][3]
    int32_t Q, k, segs;
    int32_t *idx;            // [B][Q][k] or nullptr
    double *dist2;           // [B][Q][k] or nullptr
    int32_t *count;          // [B][Q] or nullptr
    double *pts;             // neighbour coordinates, or nullptr
    int64_t pts_inst_stride; // doubles between instances (lets the caller aim at the NLP prefix)
    int64_t pts_query_stride;
    double *ws_d;            // [B][segs][Q][k] partial lists (segs > 1)
    uint32_t *ws_i;
    unsigned int *ws_counter; // [B], zero on entry, restored to zero on exit
};

__device__ __forceinline__ double knn_dist2(double qx, double qy, double qz, float px, float py,
                                            float pz) {
    const double d0 = __dsub_rn(qx, (double)px);
    double r = __dmul_rn(d0, d0);
    const double d1 = __dsub_rn(qy, (double)py);
    r = __dadd_rn(r, __dmul_rn(d1, d1));
    const double d2 = __dsub_rn(qz, (double)pz);
    r = __dadd_rn(r, __dmul_rn(d2, d2));
    return r;
}

// Lower bound of knn_dist2 over every point inside the box, computed with the
// same operations in the same order so that rounding can never make it exceed
// the distance of a contained point (rounding is monotone).
__device__ __forceinline__ double knn_box_lb(double qx, double qy, double qz, float lx, float ly,
                                             float lz, float hx, float hy, float hz) {
    const double ax = fmax(fmax(__dsub_rn((double)lx, qx), __dsub_rn(qx, (double)hx)), 0.0);
    double r = __dmul_rn(ax, ax);
    const double ay = fmax(fmax(__dsub_rn((double)ly, qy), __dsub_rn(qy, (double)hy)), 0.0);
    r = __dadd_rn(r, __dmul_rn(ay, ay));
    const double az = fmax(fmax(__dsub_rn((double)lz, qz), __dsub_rn(qz, (double)hz)), 0.0);
    r = __dadd_rn(r, __dmul_rn(az, az));
    return r;
}

// Insert (d, i) into a warp-owned list of k entries sorted by (dist2, index);
// lane j holds entry j.  No-op when the item is not among the k best.
__device__ __forceinline__ void knn_warp_insert(double *ld, uint32_t *li, int k, double d,
                                                uint32_t i, int lane) {
    const bool in = lane < k;
    const double ed = in ? ld[lane] : INFINITY;
    const uint32_t ei = in ? li[lane] : 0xffffffffu;
    const bool before = in && (ed < d || (ed == d && ei < i));
    const int pos = __popc(__ballot_sync(AMPC_FULL_MASK, before));
    __syncwarp();
    if (lane >= pos && lane < k - 1) {
        ld[lane + 1] = ed;
        li[lane + 1] = ei;
    }
    if (lane == pos && pos < k) {
        ld[pos] = d;
        li[pos] = i;
    }
    __syncwarp();
}

__device__ __forceinline__ float4 knn_ldg(const float4 *p) {
    return __ldg(p); // ld.global.nc.v4: read-only, 16-byte vector
}

__global__ void __launch_bounds__(KNN_THREADS)
knn_scan_kernel(const KnnParams P) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int Q = P.Q, k = P.k;
    double *sQ = reinterpret_cast<double *>(smem_raw);                       // Q*3
    unsigned long long *sTau = reinterpret_cast<unsigned long long *>(sQ + 3 * Q); // Q
    double *sLd = reinterpret_cast<double *>(sTau + Q);                      // [W][Q][k]
    uint32_t *sLi = reinterpret_cast<uint32_t *>(sLd + KNN_WARPS * Q * k);   // [W][Q][k]
    __shared__ int sIsLast;

    const int b = blockIdx.y, seg = blockIdx.x;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int scene = P.scene_of ? P.scene_of[b] : b;
    const int n = P.counts[scene];
    const float4 *cloud = P.clouds + (int64_t)scene * P.slot_points;

    for (int i = tid; i < 3 * Q; i += KNN_THREADS)
        sQ[i] = P.queries[(int64_t)b * 3 * Q + i];
    for (int i = tid; i < Q; i += KNN_THREADS)
        sTau[i] = 0x7ff0000000000000ull; // +inf
    for (int i = tid; i < KNN_WARPS * Q * k; i += KNN_THREADS) {
        sLd[i] = INFINITY;
        sLi[i] = 0xffffffffu;
    }
    __syncthreads();

    // ---- streaming pass -------------------------------------------------
    const int n_tiles = (n + KNN_TILE - 1) / KNN_TILE;
    const int tstride = P.segs * KNN_WARPS;
    int tile = seg * KNN_WARPS + warp;
    const float4 far = make_float4(NAN, NAN, NAN, 0.f);
    float4 nx0 = far, nx1 = far;
    if (tile < n_tiles) {
        const int i0 = tile * KNN_TILE + lane, i1 = i0 + 32;
        if (i0 < n) nx0 = knn_ldg(cloud + i0);
        if (i1 < n) nx1 = knn_ldg(cloud + i1);
    }
    for (; tile < n_tiles; tile += tstride) {
        const float4 p0 = nx0, p1 = nx1;
        const int base = tile * KNN_TILE;
        const bool v0 = base + lane < n, v1 = base + 32 + lane < n;
        { // prefetch the next tile of this warp
            const int nt = tile + tstride;
            nx0 = far;
            nx1 = far;
            if (nt < n_tiles) {
                const int i0 = nt * KNN_TILE + lane, i1 = i0 + 32;
                if (i0 < n) nx0 = knn_ldg(cloud + i0);
                if (i1 < n) nx1 = knn_ldg(cloud + i1);
            }
        }
        // tile bounding box (NaN coordinates drop out of fminf/fmaxf)
        float lx = fminf(p0.x, p1.x), ly = fminf(p0.y, p1.y), lz = fminf(p0.z, p1.z);
        float hx = fmaxf(p0.x, p1.x), hy = fmaxf(p0.y, p1.y), hz = fmaxf(p0.z, p1.z);
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            lx = fminf(lx, __shfl_xor_sync(AMPC_FULL_MASK, lx, o));
            ly = fminf(ly, __shfl_xor_sync(AMPC_FULL_MASK, ly, o));
            lz = fminf(lz, __shfl_xor_sync(AMPC_FULL_MASK, lz, o));
            hx = fmaxf(hx, __shfl_xor_sync(AMPC_FULL_MASK, hx, o));
            hy = fmaxf(hy, __shfl_xor_sync(AMPC_FULL_MASK, hy, o));
            hz = fmaxf(hz, __shfl_xor_sync(AMPC_FULL_MASK, hz, o));
        }
        for (int q0 = 0; q0 < Q; q0 += 32) {
            bool pass = false;
            if (q0 + lane < Q) {
                const int q = q0 + lane;
                const double lb = knn_box_lb(sQ[3 * q], sQ[3 * q + 1], sQ[3 * q + 2], lx, ly, lz,
                                             hx, hy, hz);
                const double t = __longlong_as_double(*(volatile unsigned long long *)&sTau[q]);
                pass = lb <= t; // false for an all-NaN box (lb is NaN)
            }
            unsigned qm = __ballot_sync(AMPC_FULL_MASK, pass);
            while (qm) {
                const int q = q0 + __ffs(qm) - 1;
                qm &= qm - 1;
                const double qx = sQ[3 * q], qy = sQ[3 * q + 1], qz = sQ[3 * q + 2];
                const double t = __longlong_as_double(*(volatile unsigned long long *)&sTau[q]);
                const double d0 = knn_dist2(qx, qy, qz, p0.x, p0.y, p0.z);
                const double d1 = knn_dist2(qx, qy, qz, p1.x, p1.y, p1.z);
                unsigned m0 = __ballot_sync(AMPC_FULL_MASK, v0 && d0 <= t);
                unsigned m1 = __ballot_sync(AMPC_FULL_MASK, v1 && d1 <= t);
                if ((m0 | m1) == 0)
                    continue;
                double *ld = sLd + (warp * Q + q) * k;
                uint32_t *li = sLi + (warp * Q + q) * k;
                double kth = ld[k - 1];
                while (m0) {
                    const int src = __ffs(m0) - 1;
                    m0 &= m0 - 1;
                    const double d = __shfl_sync(AMPC_FULL_MASK, d0, src);
                    if (d <= kth) {
                        knn_warp_insert(ld, li, k, d, (uint32_t)(base + src), lane);
                        kth = ld[k - 1];
                    }
                }
                while (m1) {
                    const int src = __ffs(m1) - 1;
                    m1 &= m1 - 1;
                    const double d = __shfl_sync(AMPC_FULL_MASK, d1, src);
                    if (d <= kth) {
                        knn_warp_insert(ld, li, k, d, (uint32_t)(base + 32 + src), lane);
                        kth = ld[k - 1];
                    }
                }
                if (lane == 0 && kth < INFINITY)
                    atomicMin(&sTau[q], (unsigned long long)__double_as_longlong(kth));
            }
        }
    }
    __syncthreads();

    // ---- merge the warps' lists: warp (q mod W) folds lists 1..W-1 into list 0
    for (int q = warp; q < Q; q += KNN_WARPS) {
        double *ld = sLd + (0 * Q + q) * k;
        uint32_t *li = sLi + (0 * Q + q) * k;
        for (int w = 1; w < KNN_WARPS; ++w) {
            const double *sd = sLd + (w * Q + q) * k;
            const uint32_t *si = sLi + (w * Q + q) * k;
            for (int j = 0; j < k; ++j) {
                const double d = sd[j];
                const uint32_t i = si[j];
                const double kd = ld[k - 1];
                const uint32_t ki = li[k - 1];
                if (!(d < kd || (d == kd && i < ki)))
                    break; // source is sorted: nothing further can enter
                knn_warp_insert(ld, li, k, d, i, lane);
            }
        }
    }
    __syncthreads();

    // ---- several segments: publish partial lists, last CTA of the instance merges
    if (P.segs > 1) {
        double *wd = P.ws_d + ((int64_t)b * P.segs + seg) * Q * k;
        uint32_t *wi = P.ws_i + ((int64_t)b * P.segs + seg) * Q * k;
        for (int i = tid; i < Q * k; i += KNN_THREADS) {
            wd[i] = sLd[i]; // list 0 of every query is contiguous: [0][q][j]
            wi[i] = sLi[i];
        }
        __threadfence();
        __syncthreads();
        if (tid == 0) {
            const unsigned t = atomicAdd(&P.ws_counter[b], 1u);
            sIsLast = (t == (unsigned)P.segs - 1);
            if (sIsLast)
                P.ws_counter[b] = 0; // ready for the next launch
        }
        __syncthreads();
        if (!sIsLast)
            return;
        __threadfence();
        for (int q = warp; q < Q; q += KNN_WARPS) {
            double *ld = sLd + q * k;
            uint32_t *li = sLi + q * k;
            for (int s = 0; s < P.segs; ++s) {
                if (s == seg)
                    continue;
                const double *sd = P.ws_d + (((int64_t)b * P.segs + s) * Q + q) * k;
                const uint32_t *si = P.ws_i + (((int64_t)b * P.segs + s) * Q + q) * k;
                for (int j = 0; j < k; ++j) {
                    const double d = __ldcg(sd + j);
                    const uint32_t i = __ldcg(si + j);
                    const double kd = ld[k - 1];
                    const uint32_t ki = li[k - 1];
                    if (!(d < kd || (d == kd && i < ki)))
                        break;
                    knn_warp_insert(ld, li, k, d, i, lane);
                }
            }
        }
        __syncthreads();
    }

    // ---- results.  SearchForNearest's count rule (kd_tree_two.h:117-124):
    // n < k -> n results, n > k -> k results, n == k -> none.
    const int cnt = (n < k) ? n : (n > k ? k : 0);
    for (int i = tid; i < Q * k; i += KNN_THREADS) {
        const int q = i / k, j = i - q * k;
        const bool ok = j < cnt;
        const uint32_t id = sLi[i];
        const int64_t o = ((int64_t)b * Q + q) * k + j;
        if (P.idx)
            P.idx[o] = ok ? (int32_t)id : -1;
        if (P.dist2)
            P.dist2[o] = ok ? sLd[i] : INFINITY;
        if (P.pts) {
            double *dst = P.pts + (int64_t)b * P.pts_inst_stride + (int64_t)q * P.pts_query_stride + 3 * j;
            if (ok) {
                const float4 p = knn_ldg(cloud + id);
                dst[0] = (double)p.x;
                dst[1] = (double)p.y;
                dst[2] = (double)p.z;
            } else { // AvoidanceStateMachine.cpp:223-226
                dst[0] = 10000.0;
                dst[1] = 10000.0;
                dst[2] = 10000.0;
            }
        }
    }
    if (P.count)
        for (int q = tid; q < Q; q += KNN_THREADS)
            P.count[(int64_t)b * Q + q] = cnt;
}

inline size_t knn_smem_bytes(int Q, int k) {
    return (size_t)Q * 3 * 8 + (size_t)Q * 8 + (size_t)KNN_WARPS * Q * k * (8 + 4);
}

// In-place, order-preserving removal of the points whose x is NaN
// (KDTreeTwo::Initialize, kd_tree_two.h:99-101).  One CTA per scene slot.
__global__ void __launch_bounds__(256)
cloud_filter_nan_kernel(float4 *clouds, int32_t *counts, int64_t slot_points, int first_scene) {
    const int scene = first_scene + blockIdx.x;
    float4 *c = clouds + (int64_t)scene * slot_points;
    const int n = counts[scene];
    __shared__ int sWarp[8];
    __shared__ int sBase;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    if (tid == 0)
        sBase = 0;
    __syncthreads();
    for (int start = 0; start < n; start += 256) {
        const int i = start + tid;
        float4 p = make_float4(0.f, 0.f, 0.f, 0.f);
        bool keep = false;
        if (i < n) {
            p = c[i];
            keep = !(p.x != p.x);
        }
        const unsigned m = __ballot_sync(AMPC_FULL_MASK, keep);
        if (lane == 0)
            sWarp[warp] = __popc(m);
        __syncthreads();
        int off = sBase;
        for (int w = 0; w < warp; ++w)
            off += sWarp[w];
        const int pos = off + __popc(m & ((1u << lane) - 1u));
        int total = 0;
        for (int w = 0; w < 8; ++w)
            total += sWarp[w];
        __syncthreads(); // every read of this chunk is done before anyone writes
        if (keep && pos != i)
            c[pos] = p;
        if (tid == 0)
            sBase += total;
        __syncthreads();
    }
    if (tid == 0)
        counts[scene] = sBase;
}

} // namespace ampc
