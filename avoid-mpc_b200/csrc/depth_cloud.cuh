// Depth image -> Obstacle cloud + Edge cloud, written straight into the scene slots in HBM
// (replaces FrameKDMap::ProcessDepth / BuildEdgeCloud and the OpenCV calls inside them,
// roswrapper/ros/src/avoid_mpc/src/FrameKDMap.cpp:76-130,176-214).
//
// Every arithmetic step keeps the C++ expression types and operation order of the reference
// (float where it uses float, double where it promotes; no FMA contraction), so the clouds are
// bit-identical to the CPU path and the k-NN indices downstream stay comparable:
//   depth_obstacle_kernel  inverse depth of the four source pixels (GetInvDepthImg), the
//                          bilinear sample cv::resize takes (a(1-f) + bf, horizontal then
//                          vertical), validity test, back-projection T * UV2Camera, and an
//                          order-preserving (row-major) compaction into the Obstacle slot;
//                          also stores the 8-bit "inflated" depth BuildEdgeCloud works on
//   edge_grad_kernel       3x3 erode, 3x3 Sobel (replicated border), L1 magnitude and the
//                          quantised gradient direction of cv::Canny's fixed-point test
//   edge_cloud_kernel      Canny's non-maximum suppression with thresholds floor(0.1) =
//                          floor(0.3) = 0 (every survivor is a strong edge), back-projection
//                          at the eroded depth, order-preserving compaction into the Edge slot
// One CTA walks one scene in row-major chunks where the output order matters; the gradient
// pass is a plain 2-D grid.  The path is HBM/L2 streaming: 4 source pixels in, 16 B out per point.
#pragma once
#include "common.cuh"

namespace ampc {

struct DepthGeom {
    double fx, fy, cx, cy;          // already divided by resize_scale (FrameKDMap.cpp:21-24)
    double p2m, dmin, dmax, range;  // range = depth_max - depth_min
    int32_t rows, cols, H, W;       // source and resized sizes
    int32_t is_u16, identity;       // identity: equal sizes, cv::resize copies
    int64_t row_stride, image_stride; // bytes
    // resize tables (device): per output column / row
    const int32_t *xofs;            // [W] left source column
    const float *xw0, *xw1;         // [W] weights 1-fx, fx; xw1 < 0: take S[xofs] alone (right border)
    const int32_t *y0, *y1;         // [H] clamped source rows
    const float *yw0, *yw1;         // [H]
};

constexpr int DC_THREADS = 1024;

// GetInvDepthImg<T> (:76-89) for one source pixel
__device__ __forceinline__ float depth_src_inv(const DepthGeom &g, const unsigned char *img, int r, int c) {
    const unsigned char *row = img + (int64_t)r * g.row_stride;
    const float pix = g.is_u16 ? (float)__ldg(reinterpret_cast<const unsigned short *>(row) + c)
                               : __ldg(reinterpret_cast<const float *>(row) + c);
    const float d = (float)__dmul_rn((double)pix, g.p2m);
    const double dd = (double)d;
    if (dd < g.dmin || dd > g.dmax)
        return 0.f;
    // float(1. / double(d)) == RN_float(1 / d): d * m for a float d (24 bits) and a float rounding
    // boundary m (25 bits) is a 49-bit integer, so 1 / d is either a boundary exactly or at least
    // 2^-49 (relative) away from one -- the 2^-53 error of the double quotient cannot move it
    // across.  One correctly rounded float reciprocal replaces the double division.
    if (fabsf(d) > 1e-18f && fabsf(d) < 1e18f)
        return __frcp_rn(d);
    return (float)(1.0 / dd);
}

__device__ __forceinline__ float depth_hsample(const DepthGeom &g, const unsigned char *img, int r, int s,
                                               float w0, float w1) {
    const float a = depth_src_inv(g, img, r, s);
    if (w1 < 0.f)
        return a;
    const float b = depth_src_inv(g, img, r, s + 1);
    return __fadd_rn(__fmul_rn(a, w0), __fmul_rn(b, w1));
}

// resized inverse depth at output pixel (row, col): the value cv::resize(INTER_LINEAR) produces
__device__ __forceinline__ float depth_inv_small(const DepthGeom &g, const unsigned char *img, int row, int col) {
    if (g.identity)
        return depth_src_inv(g, img, row, col);
    const int s = g.xofs[col];
    const float w0 = g.xw0[col], w1 = g.xw1[col];
    const float h0 = depth_hsample(g, img, g.y0[row], s, w0, w1);
    const float h1 = depth_hsample(g, img, g.y1[row], s, w0, w1);
    return __fadd_rn(__fmul_rn(h0, g.yw0[row]), __fmul_rn(h1, g.yw1[row]));
}

// T * UV2Camera(u, v, depth) stored as a pcl::PointXYZ record (:117-121,131-138)
__device__ __forceinline__ float4 depth_unproject(const DepthGeom &g, const double *T, int col, int row, double depth) {
    const double x = __dmul_rn(__dsub_rn((double)col, g.cx), depth) / g.fx;
    const double y = __dmul_rn(__dsub_rn((double)row, g.cy), depth) / g.fy;
    float o[3];
#pragma unroll
    for (int i = 0; i < 3; ++i) {
        double s = __dadd_rn(__dmul_rn(T[4 * i], x), __dmul_rn(T[4 * i + 1], y));
        s = __dadd_rn(s, __dmul_rn(T[4 * i + 2], depth));
        o[i] = (float)__dadd_rn(s, T[4 * i + 3]);
    }
    return make_float4(o[0], o[1], o[2], 0.f);
}

// ordered compaction step of one chunk: returns this thread's output position (or -1) and
// advances the running base.  sWarp: 32 ints, sBase: 1 int.
__device__ __forceinline__ int depth_chunk_slot(bool keep, int *sWarp, int *sBase) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const unsigned m = __ballot_sync(AMPC_FULL_MASK, keep);
    if (lane == 0)
        sWarp[warp] = __popc(m);
    __syncthreads();
    int v = sWarp[lane]; // DC_THREADS / 32 == 32 warps
    int incl = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const int t = __shfl_up_sync(AMPC_FULL_MASK, incl, o);
        if (lane >= o) incl += t;
    }
    const int before = __shfl_sync(AMPC_FULL_MASK, incl - v, warp);
    const int total = __shfl_sync(AMPC_FULL_MASK, incl, 31);
    const int base = *sBase;
    __syncthreads();
    if (threadIdx.x == 0)
        *sBase = base + total;
    return keep ? base + before + __popc(m & ((1u << lane) - 1u)) : -1;
}

__global__ void __launch_bounds__(DC_THREADS)
depth_obstacle_kernel(DepthGeom g, const unsigned char *__restrict__ images, const double *__restrict__ T_all,
                      float4 *__restrict__ clouds, int32_t *__restrict__ counts, int32_t *__restrict__ layout,
                      unsigned char *__restrict__ infl_all, int32_t *__restrict__ overflow, int64_t slot_points,
                      int first_scene) {
    __shared__ int sWarp[32];
    __shared__ int sBase;
    __shared__ double sT[12];
    const int b = blockIdx.x, scene = first_scene + b;
    const unsigned char *img = images + (int64_t)b * g.image_stride;
    float4 *out = clouds + (int64_t)scene * slot_points;
    unsigned char *infl = infl_all + (int64_t)b * g.H * g.W;
    if (threadIdx.x < 12)
        sT[threadIdx.x] = T_all[(int64_t)b * 16 + threadIdx.x];
    if (threadIdx.x == 0)
        sBase = 0;
    __syncthreads();
    const int npx = g.H * g.W;
    for (int start = 0; start < npx; start += DC_THREADS) {
        const int p = start + threadIdx.x;
        bool keep = false;
        int row = 0, col = 0;
        double depth = 0.0;
        if (p < npx) {
            row = p / g.W, col = p - row * g.W;
            const float v = depth_inv_small(g, img, row, col);
            const double invd = (double)v;
            // BuildEdgeCloud step 1 (:181-194): uchar(1 / invDepth / (max - min) * 200.0f) or 255
            unsigned char q = 255;
            if (invd > 1e-2)
                q = (unsigned char)((int)__dmul_rn((double)(1.f / v) / g.range, 200.0) & 0xff);
            infl[p] = q;
            // ProcessDepth loop (:110-124)
            if (!(invd < 1e-2)) {
                depth = 1.0 / invd;
                keep = depth > g.dmin && depth < g.dmax;
            }
        }
        const int slot = depth_chunk_slot(keep, sWarp, &sBase);
        if (slot >= 0 && slot < slot_points)
            out[slot] = depth_unproject(g, sT, col, row, depth);
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        int n = sBase;
        if (n > slot_points) {
            n = (int)slot_points;
            atomicExch(overflow, 1);
        }
        counts[scene] = n;
        layout[scene] = (n == npx && g.W >= 8) ? g.W : 0; // full image: 8x8 patch tiles are valid
    }
}

// eroded value and Canny gradient of every resized pixel: er (u8) and mag | dir << 12 (u16)
constexpr int EG_BX = 32, EG_BY = 8;
__global__ void __launch_bounds__(EG_BX * EG_BY)
edge_grad_kernel(int H, int W, const unsigned char *__restrict__ infl_all, unsigned char *__restrict__ er_all,
                 unsigned short *__restrict__ mag_all) {
    __shared__ unsigned char sI[EG_BY + 4][EG_BX + 4];
    __shared__ unsigned char sE[EG_BY + 2][EG_BX + 2];
    const int64_t off = (int64_t)blockIdx.z * H * W;
    const unsigned char *infl = infl_all + off;
    const int c0 = blockIdx.x * EG_BX, r0 = blockIdx.y * EG_BY;
    const int tid = threadIdx.y * EG_BX + threadIdx.x;
    for (int i = tid; i < (EG_BY + 4) * (EG_BX + 4); i += EG_BX * EG_BY) {
        const int rr = i / (EG_BX + 4), cc = i - rr * (EG_BX + 4);
        const int r = r0 - 2 + rr, c = c0 - 2 + cc;
        sI[rr][cc] = (r >= 0 && r < H && c >= 0 && c < W) ? infl[(int64_t)r * W + c] : 255; // outside: ignored by erode
    }
    __syncthreads();
    for (int i = tid; i < (EG_BY + 2) * (EG_BX + 2); i += EG_BX * EG_BY) {
        const int rr = i / (EG_BX + 2), cc = i - rr * (EG_BX + 2);
        // Sobel sees the eroded image with a replicated border: clamp the coordinate first
        const int r = min(max(r0 - 1 + rr, 0), H - 1), c = min(max(c0 - 1 + cc, 0), W - 1);
        const int lr = r - (r0 - 2), lc = c - (c0 - 2);
        int m = 255;
        if (lr >= 1 && lr < EG_BY + 3 && lc >= 1 && lc < EG_BX + 3) {
#pragma unroll
            for (int dy = -1; dy <= 1; ++dy)
#pragma unroll
                for (int dx = -1; dx <= 1; ++dx)
                    m = min(m, (int)sI[lr + dy][lc + dx]);
        }
        sE[rr][cc] = (unsigned char)m;
    }
    __syncthreads();
    const int r = r0 + threadIdx.y, c = c0 + threadIdx.x;
    if (r >= H || c >= W)
        return;
    const int y = threadIdx.y + 1, x = threadIdx.x + 1;
    const int gx = ((int)sE[y - 1][x + 1] + 2 * (int)sE[y][x + 1] + (int)sE[y + 1][x + 1]) -
                   ((int)sE[y - 1][x - 1] + 2 * (int)sE[y][x - 1] + (int)sE[y + 1][x - 1]);
    const int gy = ((int)sE[y + 1][x - 1] + 2 * (int)sE[y + 1][x] + (int)sE[y + 1][x + 1]) -
                   ((int)sE[y - 1][x - 1] + 2 * (int)sE[y - 1][x] + (int)sE[y - 1][x + 1]);
    const int ax = abs(gx), ay = abs(gy) << 15;
    const int tg22 = ax * 13573; // tan(22.5 deg) in Q15, cv::Canny's TG22
    int dir;
    if (ay < tg22)
        dir = 0; // compare left / right
    else if (ay > tg22 + (ax << 16))
        dir = 1; // compare up / down
    else
        dir = ((gx ^ gy) >= 0) ? 2 : 3; // diagonal, same / opposite signs
    er_all[off + (int64_t)r * W + c] = sE[y][x];
    mag_all[off + (int64_t)r * W + c] = (unsigned short)((ax + abs(gy)) | (dir << 12));
}

__global__ void __launch_bounds__(DC_THREADS)
edge_cloud_kernel(DepthGeom g, const double *__restrict__ T_all, const unsigned char *__restrict__ er_all,
                  const unsigned short *__restrict__ mag_all, float4 *__restrict__ clouds,
                  int32_t *__restrict__ counts, int32_t *__restrict__ layout,
                  const int32_t *__restrict__ obstacle_counts, int32_t *__restrict__ overflow,
                  int64_t slot_points, int first_scene) {
    __shared__ int sWarp[32];
    __shared__ int sBase;
    __shared__ double sT[12];
    const int b = blockIdx.x, scene = first_scene + b;
    if (obstacle_counts[scene] == 0) { // ProcessDepth returns before BuildEdgeCloud (:125-127)
        if (threadIdx.x == 0)
            counts[scene] = 0, layout[scene] = 0;
        return;
    }
    const int H = g.H, W = g.W, npx = H * W;
    const unsigned char *er = er_all + (int64_t)b * npx;
    const unsigned short *mg = mag_all + (int64_t)b * npx;
    float4 *out = clouds + (int64_t)scene * slot_points;
    if (threadIdx.x < 12)
        sT[threadIdx.x] = T_all[(int64_t)b * 16 + threadIdx.x];
    if (threadIdx.x == 0)
        sBase = 0;
    __syncthreads();
    auto mag_at = [&](int r, int c) -> int { // the magnitude buffer is zero outside the image
        return (r >= 0 && r < H && c >= 0 && c < W) ? (int)(mg[r * W + c] & 0xfff) : 0;
    };
    for (int start = 0; start < npx; start += DC_THREADS) {
        const int p = start + threadIdx.x;
        bool keep = false;
        int row = 0, col = 0;
        double depth = 0.0;
        if (p < npx) {
            row = p / W, col = p - row * W;
            const int mc = mg[p], m = mc & 0xfff, dir = mc >> 12;
            bool edge = false;
            if (m > 0) { // m > low = 0; every survivor exceeds high = 0
                if (dir == 0)
                    edge = m > mag_at(row, col - 1) && m >= mag_at(row, col + 1);
                else if (dir == 1)
                    edge = m > mag_at(row - 1, col) && m >= mag_at(row + 1, col);
                else if (dir == 2)
                    edge = m > mag_at(row - 1, col - 1) && m > mag_at(row + 1, col + 1);
                else
                    edge = m > mag_at(row - 1, col + 1) && m > mag_at(row + 1, col - 1);
            }
            if (edge) { // :199-206
                depth = __dmul_rn((double)(float)er[p], g.range) / 200.0;
                keep = !(depth > g.dmax || depth < g.dmin);
            }
        }
        const int slot = depth_chunk_slot(keep, sWarp, &sBase);
        if (slot >= 0 && slot < slot_points)
            out[slot] = depth_unproject(g, sT, col, row, depth);
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        int n = sBase;
        if (n > slot_points) {
            n = (int)slot_points;
            atomicExch(overflow, 1);
        }
        counts[scene] = n;
        layout[scene] = 0;
    }
}

__global__ void fill_i32_kernel(int32_t *p, int n, int32_t v) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n)
        p[i] = v;
}

} // namespace ampc
