// Exact batched k-NN over a flat tile-AABB index (sm_100a).
//
// Replaces KDTreeTwo::Initialize + SearchForNearest, i.e. nanoflann buildIndex /
// findNeighbors / searchLevel (include/kd_tree_two.h:88-133,
// include/nanoflann_two.hpp:1518-1541,1563-1586,1729-1793), for B instances x Q
// queries per launch.
//
// Index ("build", once per cloud = once per depth frame, src/FrameKDMap.cpp:34-52):
//   ONE streaming pass over the 16-byte point records (TMA bulk copies into double-
//   buffered shared memory) emits, per tile of 64 points, the bounding box and the number
//   of points inside (32 B per tile, 3% of the cloud).  A tile is 64 consecutive records,
//   or an 8x8 patch of the depth image when the row pitch is known (organised cloud).
//   Records whose x is NaN (kd_tree_two.h:99-101) raise a flag; a second, normally idle
//   kernel compacts such scenes in place (order preserving) and rebuilds their boxes.
//   This pass is the HBM-bound kernel of the k-NN stage; nothing is sorted.
// Search: one warp per (instance, query).  Lanes evaluate conservative FP32 lower / upper
//   bounds of the query to 32 tile boxes at a time; a few best-first picks give a tight
//   k-th-best bound, then only tiles whose lower bound does not exceed the current bound
//   are read.  The top-k list lives in registers (lane j = entry j); dense tiles are merged
//   with a warp bitonic network.
//
// Arithmetic of the distances is the reference's, operation for operation: dist2 =
// ((dx*dx + dy*dy) + dz*dz) in double from float coordinates, each operation rounded
// separately (nanoflann_two.hpp:590-599, kd_tree_two.h:34-41).  The pruning bounds are
// rounded toward the safe side, so rounding can never prune a true neighbour.  Indices
// AND squared distances are bit-exact; order is canonical (dist2, index).
#pragma once
#include "common.cuh"

namespace ampc {

constexpr int KT_TILE = 64;       // points per tile (2 x 16-byte loads per lane)
constexpr int KNN_KMAX = 32;      // top-k list = one entry per lane
constexpr int KS_WARPS = 4;       // queries (warps) per search CTA
constexpr int KS_CHUNK = 1024;    // tile lower bounds kept in shared memory per warp
constexpr int KS_DENSE = 12;      // candidates in a tile from which the bitonic merge is used
constexpr int KS_PICKS = 3;       // best-first tile picks before the storage-order sweep
constexpr int KI_THREADS = 128;   // index kernels: 4 warps per CTA

// Tile geometry.  row_w == 0: tile t = 64 consecutive records.  row_w > 0 (organised cloud,
// the row pitch of the depth image the cloud was back-projected from): tile t = an 8x8 patch
// of the image = 8 runs of 8 consecutive records (128 B each), row pitch row_w records -- a
// compact box in space.  Any set of 64 indices gives a VALID box; the hint only makes it tight.
struct TileGeom {
    int row_w, tiles_x, n_tiles;
    __host__ __device__ TileGeom(int n, int row_w_) : row_w(row_w_) {
        if (row_w > 0) {
            const int rows = (n + row_w - 1) / row_w;
            tiles_x = (row_w + 7) / 8;
            n_tiles = ((rows + 7) / 8) * tiles_x;
        } else {
            tiles_x = 0;
            n_tiles = (n + KT_TILE - 1) / KT_TILE;
        }
    }
    // record indices of slots `lane` and `lane + 32` of tile t (-1: empty slot); one division
    __device__ void points2(int t, int lane, int n, int &i0, int &i1) const {
        if (row_w > 0) {
            const int ty = t / tiles_x, tx = t - ty * tiles_x;
            const int col = 8 * tx + (lane & 7);
            const int a = (8 * ty + (lane >> 3)) * row_w + col, b = a + 4 * row_w;
            const bool ok = col < row_w;
            i0 = ok && a < n ? a : -1;
            i1 = ok && b < n ? b : -1;
        } else {
            const int a = t * KT_TILE + lane, b = a + 32;
            i0 = a < n ? a : -1;
            i1 = b < n ? b : -1;
        }
    }
    // record index of slot s (0..63) of tile t, or -1 if the slot is empty
    __host__ __device__ int point(int t, int s, int n) const {
        int i;
        if (row_w > 0) {
            const int ty = t / tiles_x, tx = t - ty * tiles_x;
            const int col = 8 * tx + (s & 7);
            if (col >= row_w)
                return -1;
            i = (8 * ty + (s >> 3)) * row_w + col;
        } else {
            i = t * KT_TILE + s;
        }
        return i < n ? i : -1;
    }
};
// tile slots per scene: enough for the linear layout and for every organised cloud that spans
// at least 8 rows; a wider row hint (n_tiles grows to ceil(w/8) for a 1-row cloud) falls back
// to the linear layout through tile_layout() below, in every kernel alike
__host__ __device__ inline int64_t tile_capacity(int64_t max_points) {
    return 2 * ((max_points + KT_TILE - 1) / KT_TILE) + 64;
}
// the layout a scene of n points is actually indexed with: its hint, unless the 8x8-patch
// tiling of that hint would need more tile slots than the scene has
__host__ __device__ inline int tile_layout(int n, int hint, int64_t slot_tiles) {
    if (hint > 0 && TileGeom(n, hint).n_tiles > slot_tiles)
        return 0;
    return hint;
}

// ---- second level of the index: groups of 16 tiles (a 4x4 block of 8x8 image patches = 32x32
// pixels of an organised cloud; 16 consecutive tiles = 1024 consecutive records otherwise), one
// 32-byte box record each.  The search tests the groups first and only looks at the tiles of the
// groups that can still hold a neighbour.
constexpr int KG_TILES = 16;
struct GroupGeom {
    int org, tiles_x, tiles_y, groups_x, n_groups, n_tiles;
    __host__ __device__ explicit GroupGeom(const TileGeom &g) {
        n_tiles = g.n_tiles;
        if (g.row_w > 0) {
            org = 1, tiles_x = g.tiles_x, tiles_y = g.tiles_x > 0 ? g.n_tiles / g.tiles_x : 0;
            groups_x = (tiles_x + 3) / 4;
            n_groups = groups_x * ((tiles_y + 3) / 4);
        } else {
            org = 0, tiles_x = tiles_y = groups_x = 0;
            n_groups = (n_tiles + KG_TILES - 1) / KG_TILES;
        }
    }
    // tile s (0..15) of group g, or -1; (ty, tx) = its patch coordinates in an organised cloud
    __device__ int tile(int g, int s, int &ty, int &tx) const {
        if (org) {
            const int gy = g / groups_x, gx = g - gy * groups_x;
            ty = 4 * gy + (s >> 2), tx = 4 * gx + (s & 3);
            return (ty < tiles_y && tx < tiles_x) ? ty * tiles_x + tx : -1;
        }
        ty = tx = 0;
        const int t = KG_TILES * g + s;
        return t < n_tiles ? t : -1;
    }
};
// group slots per scene (an organised cloud one tile row high has ceil(tiles/4) groups)
__host__ __device__ inline int64_t group_capacity(int64_t slot_tiles) { return slot_tiles / 4 + 64; }

struct KnnParams {
    const float4 *clouds;    // slot s at clouds + s*slot_points
    const float4 *sorted;    // Morton-bucketed copies (x, y, z, original index) of the scenes whose layout is -1
    const float4 *boxes;     // slot s at boxes + s*slot_tiles*2 (two float4 per tile)
    const float4 *gboxes;    // slot s at gboxes + s*slot_groups*2 (two float4 per group of tiles)
    int gchunk;              // group lower bounds kept in shared memory per warp (<= KS_GCHUNK, multiple of 32)
    int64_t slot_groups;
    const int32_t *counts;   // points held in each slot (after the NaN filter)
    int64_t slot_points;
    int64_t slot_tiles;
    const int32_t *layout;   // per scene: 0 = unorganised cloud; >0 = row pitch for 8x8 patch tiles;
                             // -1 = tiles over the Morton-bucketed copy (cloud_sort_kernel)
    const int32_t *scene_of; // [B] or nullptr (identity)
    const int32_t *active;   // [B] or nullptr: instances with 0 are skipped (outputs untouched)
    const double *queries;   // [B][Q][3]
    int32_t Q, k, segs;
    int32_t *idx;            // [B][Q][k] or nullptr
    double *dist2;           // [B][Q][k] or nullptr
    int32_t *count;          // [B][Q] or nullptr
    double *pts;             // neighbour coordinates, or nullptr
    int64_t pts_inst_stride; // doubles between instances (lets the caller aim at the NLP prefix)
    int64_t pts_query_stride;
    double *ws_d;            // [B][Q][segs][k] partial lists (segs > 1)
    uint32_t *ws_i;
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

__device__ __forceinline__ float4 knn_ldg(const float4 *p) {
    return __ldg(p); // ld.global.nc.v4
}

// ---- order-preserving float <-> int map (flip the magnitude bits of negative values) so
// that redux.sync (integer warp reduction, one instruction) yields exact float min / max
__device__ __forceinline__ int f2ord(float f) {
    const int i = __float_as_int(f);
    return i ^ ((i >> 31) & 0x7fffffff);
}
__device__ __forceinline__ float ord2f(int i) {
    return __int_as_float(i ^ ((i >> 31) & 0x7fffffff));
}
__device__ __forceinline__ float warp_fmin(float v) { // v must not be NaN
    return ord2f(__reduce_min_sync(AMPC_FULL_MASK, f2ord(v)));
}
__device__ __forceinline__ float warp_fmax(float v) {
    return ord2f(__reduce_max_sync(AMPC_FULL_MASK, f2ord(v)));
}

// Non-negative floats (and +inf) order like their bit patterns: warp minimum in ONE redux.sync
// instead of five shuffle steps.  The sign bit is masked so that a -0 cannot win or lose wrongly;
// a NaN (0x7fc00000) is larger than +inf and never the minimum unless every lane holds one.
__device__ __forceinline__ unsigned fbits_nonneg(float v) { return __float_as_uint(v) & 0x7fffffffu; }
__device__ __forceinline__ float warp_min_nonneg(float v) {
    return __uint_as_float(__reduce_min_sync(AMPC_FULL_MASK, fbits_nonneg(v)));
}
// arg-min over the lanes of a non-negative float, ties to the smallest `idx` (idx >= 0; a lane
// without a candidate passes idx = INT_MAX): two redux.sync.  Returns the minimum, idx in `bi`.
__device__ __forceinline__ float warp_argmin_nonneg(float v, int idx, int &bi) {
    const unsigned mine = fbits_nonneg(v);
    const unsigned m = __reduce_min_sync(AMPC_FULL_MASK, mine);
    bi = (int)__reduce_min_sync(AMPC_FULL_MASK, mine == m ? (unsigned)idx : 0x7fffffffu);
    return __uint_as_float(m);
}

__device__ __forceinline__ void tile_box_store(float4 *boxes, int64_t tile, float4 p0, float4 p1,
                                               bool v0, bool v1, int lane) {
    // empty slots and NaN coordinates drop out: fminf/fmaxf return the non-NaN operand
    const float big = INFINITY;
    const float x0 = v0 ? p0.x : NAN, y0 = v0 ? p0.y : NAN, z0 = v0 ? p0.z : NAN;
    const float x1 = v1 ? p1.x : NAN, y1 = v1 ? p1.y : NAN, z1 = v1 ? p1.z : NAN;
    const float lx = warp_fmin(fminf(fminf(x0, x1), big));
    const float ly = warp_fmin(fminf(fminf(y0, y1), big));
    const float lz = warp_fmin(fminf(fminf(z0, z1), big));
    const float hx = warp_fmax(fmaxf(fmaxf(x0, x1), -big));
    const float hy = warp_fmax(fmaxf(fmaxf(y0, y1), -big));
    const float hz = warp_fmax(fmaxf(fmaxf(z0, z1), -big));
    // points of the tile with no NaN coordinate: how many points certainly lie inside the box
    // (x+y+z is NaN iff the slot is empty or some coordinate is NaN -- or inf-inf, which only
    // under-counts, and an under-count is safe)
    const float s0 = x0 + y0 + z0, s1 = x1 + y1 + z1;
    const bool f0 = s0 == s0, f1 = s1 == s1;
    const int cnt = __popc(__ballot_sync(AMPC_FULL_MASK, f0)) + __popc(__ballot_sync(AMPC_FULL_MASK, f1));
    if (lane == 0) {
        boxes[2 * tile] = make_float4(lx, ly, lz, hx);
        boxes[2 * tile + 1] = make_float4(hy, hz, __int_as_float(cnt), 0.f);
    }
}

// ---- TMA bulk copy (cp.async.bulk, global -> shared, completion on an mbarrier) ----------
__device__ __forceinline__ uint32_t smem_u32(const void *p) {
    return (uint32_t)__cvta_generic_to_shared(p);
}
__device__ __forceinline__ void mbar_init(unsigned long long *bar, int count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(unsigned long long *bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes)
                 : "memory");
}
__device__ __forceinline__ void bulk_g2s(void *dst, const void *src, uint32_t bytes, unsigned long long *bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                     smem_u32(dst)),
                 "l"(src), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}
__device__ __forceinline__ void mbar_wait(unsigned long long *bar, uint32_t parity) {
    asm volatile("{\n"
                 ".reg .pred p;\n"
                 "WAIT_%=:\n"
                 "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
                 "@p bra DONE_%=;\n"
                 "bra WAIT_%=;\n"
                 "DONE_%=:\n"
                 "}" ::"r"(smem_u32(bar)),
                 "r"(parity)
                 : "memory");
}

// Index build, common path.  Grid (scenes, parts).  A CTA walks a range of "bands" of its
// scene.  A band is what 32 tiles cover: 8 image rows x up to 256 columns for an organised
// cloud (8 contiguous row segments), or 2048 consecutive records.  Each band is brought
// into shared memory by TMA bulk copies (one per row segment, double buffered: the next
// band streams in while the boxes of the current one are reduced), so HBM sees long
// linear reads whatever the tile geometry.  Warps then reduce the tiles' boxes from shared
// memory.  A record whose x is NaN raises the scene's flag (handled by cloud_compact_kernel).
constexpr int KI_COLS = 256;                   // columns (records per row segment) per band
constexpr int KI_BAND_BYTES = 8 * KI_COLS * 16; // 32 KB per buffer
__global__ void __launch_bounds__(KI_THREADS)
cloud_index_kernel(const float4 *__restrict__ clouds, float4 *__restrict__ boxes,
                   const int32_t *__restrict__ counts, int32_t *__restrict__ nan_flags,
                   int64_t slot_points, int64_t slot_tiles, const int32_t *__restrict__ layout, int first_scene) {
    extern __shared__ __align__(128) unsigned char ki_smem[];
    __shared__ __align__(8) unsigned long long bar[2];
    float4 *buf[2] = {reinterpret_cast<float4 *>(ki_smem), reinterpret_cast<float4 *>(ki_smem + KI_BAND_BYTES)};
    const int scene = first_scene + blockIdx.x; // scenes on x: no 65535 limit
    const float4 *c = clouds + (int64_t)scene * slot_points;
    float4 *bx = boxes + (int64_t)scene * slot_tiles * 2;
    const int n = counts[scene];
    const int row_w = tile_layout(n, layout[scene], slot_tiles);
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    constexpr int NW = KI_THREADS / 32;
    const TileGeom g(n, row_w);
    // band grid of this scene
    const int segs_x = row_w > 0 ? (row_w + KI_COLS - 1) / KI_COLS : 1;
    const int bands_y = row_w > 0 ? (((n + row_w - 1) / row_w) + 7) / 8 : (n + 8 * KI_COLS - 1) / (8 * KI_COLS);
    const int n_bands = bands_y * segs_x;
    const int per = (n_bands + gridDim.y - 1) / gridDim.y;
    const int b_begin = blockIdx.y * per, b_end = min(n_bands, b_begin + per);
    if (b_begin >= b_end)
        return;
    if (tid == 0) {
        mbar_init(&bar[0], 1);
        mbar_init(&bar[1], 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    // geometry of band b: rows [8*by, 8*by+8), columns [c0, c0+cols); record index of (r, col)
    auto issue = [&](int b, int which) { // one thread
        const int by = b / segs_x, bxs = b - by * segs_x;
        uint32_t total = 0;
        if (row_w > 0 && row_w <= KI_COLS) { // the 8 rows of the band are one contiguous span
            const int64_t start = (int64_t)8 * by * row_w;
            total = (uint32_t)min((int64_t)8 * row_w, (int64_t)n - start) * 16u;
            mbar_expect_tx(&bar[which], total);
            bulk_g2s(buf[which], c + start, total, &bar[which]);
        } else if (row_w > 0) {
            const int c0 = bxs * KI_COLS, cols = min(KI_COLS, row_w - c0);
            uint32_t bytes[8];
#pragma unroll
            for (int r = 0; r < 8; ++r) {
                const int64_t start = (int64_t)(8 * by + r) * row_w + c0;
                const int64_t cnt = min((int64_t)cols, (int64_t)n - start);
                bytes[r] = cnt > 0 ? (uint32_t)cnt * 16u : 0u;
                total += bytes[r];
            }
            mbar_expect_tx(&bar[which], total);
#pragma unroll
            for (int r = 0; r < 8; ++r)
                if (bytes[r])
                    bulk_g2s(buf[which] + r * KI_COLS, c + (int64_t)(8 * by + r) * row_w + c0, bytes[r], &bar[which]);
        } else {
            const int64_t start = (int64_t)b * 8 * KI_COLS;
            total = (uint32_t)min((int64_t)8 * KI_COLS, (int64_t)n - start) * 16u;
            mbar_expect_tx(&bar[which], total);
            bulk_g2s(buf[which], c + start, total, &bar[which]);
        }
    };
    bool nan_seen = false;
    const int full_rows = row_w > 0 ? n / row_w : 0, rem_cols = row_w > 0 ? n - full_rows * row_w : 0;
    if (tid == 0)
        issue(b_begin, 0);
    for (int b = b_begin; b < b_end; ++b) {
        const int it = b - b_begin, cur = it & 1;
        if (tid == 0 && b + 1 < b_end)
            issue(b + 1, cur ^ 1); // buffer cur^1 was released by the barrier ending iteration it-1
        mbar_wait(&bar[cur], (it >> 1) & 1);
        const float4 *sb = buf[cur];
        const int by = b / segs_x, bxs = b - by * segs_x;
        // Four tiles per warp pass, one per quarter-warp: lane (q, l) walks the 8 records of
        // column l of tile q (organised: the 8 image rows of the patch; unorganised: records
        // l, l + 8, ..), folding min / max / valid-count in registers; three xor-shuffles
        // finish the tile inside its 8 lanes.  Every shared-memory read is 512 contiguous
        // bytes per warp (organised) or four 128-byte runs (unorganised).
        const int qd = lane >> 3, l8 = lane & 7;
        int tiles_here, tile0, pitch = 0, col0 = 0, c0 = 0;
        if (row_w > 0) {
            c0 = bxs * KI_COLS;
            col0 = min(KI_COLS, row_w - c0);            // valid columns of this band
            tiles_here = (col0 + 7) / 8;
            tile0 = by * g.tiles_x + c0 / 8;
            pitch = row_w <= KI_COLS ? row_w : KI_COLS; // records between rows in shared memory
        } else {
            const int start = b * 8 * KI_COLS;
            col0 = min(8 * KI_COLS, n - start);         // records in this band
            tiles_here = (col0 + KT_TILE - 1) / KT_TILE;
            tile0 = start / KT_TILE;
        }
        for (int j4 = 4 * warp; j4 < tiles_here; j4 += 4 * NW) {
            const int t = j4 + qd;
            float lx = INFINITY, ly = INFINITY, lz = INFINITY, hx = -INFINITY, hy = -INFINITY, hz = -INFINITY;
            int cnt = 0;
            if (row_w > 0) {
                // rows of this column that exist: image rows < full_rows are complete, row
                // full_rows holds the first rem_cols columns
                const int col = 8 * t + l8;
                int nv = full_rows - 8 * by + (c0 + col < rem_cols ? 1 : 0);
                nv = col < col0 ? min(max(nv, 0), 8) : 0;
#pragma unroll
                for (int i = 0; i < 8; ++i) {
                    if (i < nv) {
                        const float4 p = sb[i * pitch + col];
                        nan_seen |= p.x != p.x;
                        lx = fminf(lx, p.x), ly = fminf(ly, p.y), lz = fminf(lz, p.z); // NaN drops out
                        hx = fmaxf(hx, p.x), hy = fmaxf(hy, p.y), hz = fmaxf(hz, p.z);
                        const float sum = p.x + p.y + p.z; // NaN iff some coordinate is NaN (or inf - inf:
                        cnt += sum == sum;                 // an under-count, which is safe)
                    }
                }
            } else {
#pragma unroll
                for (int i = 0; i < 8; ++i) {
                    const int sidx = KT_TILE * t + 8 * i + l8;
                    if (sidx < col0) {
                        const float4 p = sb[sidx];
                        nan_seen |= p.x != p.x;
                        lx = fminf(lx, p.x), ly = fminf(ly, p.y), lz = fminf(lz, p.z);
                        hx = fmaxf(hx, p.x), hy = fmaxf(hy, p.y), hz = fmaxf(hz, p.z);
                        const float sum = p.x + p.y + p.z;
                        cnt += sum == sum;
                    }
                }
            }
#pragma unroll
            for (int o = 1; o < 8; o <<= 1) {
                lx = fminf(lx, __shfl_xor_sync(AMPC_FULL_MASK, lx, o));
                ly = fminf(ly, __shfl_xor_sync(AMPC_FULL_MASK, ly, o));
                lz = fminf(lz, __shfl_xor_sync(AMPC_FULL_MASK, lz, o));
                hx = fmaxf(hx, __shfl_xor_sync(AMPC_FULL_MASK, hx, o));
                hy = fmaxf(hy, __shfl_xor_sync(AMPC_FULL_MASK, hy, o));
                hz = fmaxf(hz, __shfl_xor_sync(AMPC_FULL_MASK, hz, o));
                cnt += __shfl_xor_sync(AMPC_FULL_MASK, cnt, o);
            }
            if (l8 == 0 && t < tiles_here) {
                float4 *dst = bx + 2 * (int64_t)(tile0 + t);
                dst[0] = make_float4(lx, ly, lz, hx);
                dst[1] = make_float4(hy, hz, __int_as_float(cnt), 0.f);
            }
        }
        __syncthreads(); // everyone is done with buf[cur]: it may be refilled
    }
    if (__any_sync(AMPC_FULL_MASK, nan_seen) && lane == 0)
        atomicOr(&nan_flags[scene], 1);
}

// Index build, rare path (one CTA per scene, returns at once unless the scene's flag is set):
// in-place, order-preserving removal of the records whose x is NaN (kd_tree_two.h:99-101),
// then the boxes are rebuilt over the compacted cloud.
__global__ void __launch_bounds__(KI_THREADS)
cloud_compact_kernel(float4 *clouds, float4 *boxes, int32_t *counts, int32_t *nan_flags,
                     int64_t slot_points, int64_t slot_tiles, const int32_t *layout, int first_scene) {
    const int scene = first_scene + blockIdx.x;
    if (nan_flags[scene] == 0)
        return;
    const int row_w = layout[scene];
    float4 *c = clouds + (int64_t)scene * slot_points;
    float4 *bx = boxes + (int64_t)scene * slot_tiles * 2;
    const int n = counts[scene];
    __shared__ int sWarp[KI_THREADS / 32];
    __shared__ int sBase;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    constexpr int NW = KI_THREADS / 32;
    constexpr int PER_IT = NW * KT_TILE;
    if (tid == 0)
        sBase = 0;
    __syncthreads();
    for (int start = 0; start < n; start += PER_IT) {
        const int tb = start + warp * KT_TILE;
        const int j0 = tb + lane, j1 = j0 + 32;
        float4 p0 = make_float4(0, 0, 0, 0), p1 = p0;
        if (j0 < n) p0 = c[j0];
        if (j1 < n) p1 = c[j1];
        const bool k0 = j0 < n && !(p0.x != p0.x), k1 = j1 < n && !(p1.x != p1.x);
        const unsigned m0 = __ballot_sync(AMPC_FULL_MASK, k0), m1 = __ballot_sync(AMPC_FULL_MASK, k1);
        if (lane == 0)
            sWarp[warp] = __popc(m0) + __popc(m1);
        __syncthreads(); // counts visible; every read of this chunk is complete
        int off = sBase, total = 0;
#pragma unroll
        for (int w = 0; w < NW; ++w) {
            const int v = sWarp[w];
            if (w < warp) off += v;
            total += v;
        }
        const unsigned lt = (1u << lane) - 1u;
        if (k0) c[off + __popc(m0 & lt)] = p0; // lands at or below this chunk's reads
        if (k1) c[off + __popc(m0) + __popc(m1 & lt)] = p1;
        __syncthreads();
        if (tid == 0) sBase += total;
        __syncthreads();
    }
    const int m = sBase;
    if (tid == 0) {
        counts[scene] = m;
        nan_flags[scene] = 0;
    }
    const TileGeom g(m, tile_layout(m, row_w, slot_tiles));
    for (int t = warp; t < g.n_tiles; t += NW) {
        const int i0 = g.point(t, lane, m), i1 = g.point(t, lane + 32, m);
        float4 p0 = make_float4(0, 0, 0, 0), p1 = p0;
        if (i0 >= 0) p0 = c[i0];
        if (i1 >= 0) p1 = c[i1];
        tile_box_store(bx, t, p0, p1, i0 >= 0, i1 >= 0, lane);
    }
}

// ---- Morton bucketing of an unorganised cloud (layout -1) --------------------------------
// A cloud in arbitrary storage order gives 64-record tiles whose boxes span the scene, and
// the search degenerates to a full scan.  For such clouds the index is built over a COPY of
// the cloud bucketed by the 15-bit Morton code of each point (5 bits per axis of the scene's
// bounding box): a counting sort, one CTA per scene, histogram and running offsets in shared
// memory (32768 cells x 4 B = 128 KB).  The copy keeps the original index in the pad word
// (x, y, z, index), so neighbour indices remain the reference's indices; the order inside a
// cell is whatever the atomics produce, which cannot matter: the result list is ordered by
// (dist2, index) and every box is valid for whatever 64 records its tile holds.
// This is the flat (one level of 64-point leaves) form of a Morton-ordered linear BVH.
constexpr int KM_THREADS = 1024;
constexpr int KM_BITS = 5;
constexpr int KM_CELLS = 1 << (3 * KM_BITS);
constexpr int KM_SMEM_BYTES = KM_CELLS * 4;

__device__ __forceinline__ uint32_t morton_spread(uint32_t v) { // bit i of v (v < 1024) -> bit 3i
    v = (v * 0x00010001u) & 0xFF0000FFu;
    v = (v * 0x00000101u) & 0x0F00F00Fu;
    v = (v * 0x00000011u) & 0xC30C30C3u;
    v = (v * 0x00000005u) & 0x49249249u;
    return v;
}
struct MortonGrid {
    float lx, ly, lz, sx, sy, sz;
    // NaN / inf coordinates and degenerate extents fall into cell 0 of that axis
    // (fmaxf(NaN, 0) = 0); every record gets SOME cell, which is all correctness needs
    __device__ __forceinline__ uint32_t cell(float4 p) const {
        const float top = (float)((1 << KM_BITS) - 1);
        const uint32_t cx = (uint32_t)fminf(fmaxf((p.x - lx) * sx, 0.f), top);
        const uint32_t cy = (uint32_t)fminf(fmaxf((p.y - ly) * sy, 0.f), top);
        const uint32_t cz = (uint32_t)fminf(fmaxf((p.z - lz) * sz, 0.f), top);
        return morton_spread(cx) | (morton_spread(cy) << 1) | (morton_spread(cz) << 2);
    }
};

// boxes: the linear-tile boxes of the ORIGINAL cloud (cloud_index_kernel ran over it just
// before, which also applied the NaN filter); they are only reduced to the scene's box here.
__global__ void __launch_bounds__(KM_THREADS, 1)
cloud_sort_kernel(const float4 *__restrict__ clouds, float4 *__restrict__ sorted,
                  const float4 *__restrict__ boxes, const int32_t *__restrict__ counts,
                  const int32_t *__restrict__ layout, int64_t slot_points, int64_t slot_tiles,
                  int first_scene) {
    extern __shared__ __align__(16) uint32_t km_hist[]; // KM_CELLS counters, then running offsets
    __shared__ float sBox[6][KM_THREADS / 32];
    __shared__ uint32_t sWarpTotal[KM_THREADS / 32];
    const int scene = first_scene + blockIdx.x;
    if (layout[scene] != -1)
        return;
    const int n = counts[scene];
    const float4 *c = clouds + (int64_t)scene * slot_points;
    float4 *s = sorted + (int64_t)scene * slot_points;
    const float4 *bx = boxes + (int64_t)scene * slot_tiles * 2;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    constexpr int NW = KM_THREADS / 32;

    // scene box = union of the tile boxes (empty tiles hold +inf / -inf and drop out)
    float lo[3] = {INFINITY, INFINITY, INFINITY}, hi[3] = {-INFINITY, -INFINITY, -INFINITY};
    const int n_tiles = (n + KT_TILE - 1) / KT_TILE;
    for (int t = tid; t < n_tiles; t += KM_THREADS) {
        const float4 a = bx[2 * t], h = bx[2 * t + 1];
        lo[0] = fminf(lo[0], a.x), lo[1] = fminf(lo[1], a.y), lo[2] = fminf(lo[2], a.z);
        hi[0] = fmaxf(hi[0], a.w), hi[1] = fmaxf(hi[1], h.x), hi[2] = fmaxf(hi[2], h.y);
    }
#pragma unroll
    for (int a = 0; a < 3; ++a) {
        const float l = warp_fmin(lo[a]), h = warp_fmax(hi[a]);
        if (lane == 0) {
            sBox[a][warp] = l;
            sBox[3 + a][warp] = h;
        }
    }
    for (int i = tid; i < KM_CELLS; i += KM_THREADS)
        km_hist[i] = 0u;
    __syncthreads();
#pragma unroll
    for (int a = 0; a < 3; ++a) {
        lo[a] = warp_fmin(sBox[a][lane]);
        hi[a] = warp_fmax(sBox[3 + a][lane]);
    }
    MortonGrid g;
    {
        const float cells = (float)(1 << KM_BITS);
        const float ex = hi[0] - lo[0], ey = hi[1] - lo[1], ez = hi[2] - lo[2];
        g.lx = lo[0], g.ly = lo[1], g.lz = lo[2];
        g.sx = (ex > 0.f && ex < INFINITY) ? cells / ex : 0.f;
        g.sy = (ey > 0.f && ey < INFINITY) ? cells / ey : 0.f;
        g.sz = (ez > 0.f && ez < INFINITY) ? cells / ez : 0.f;
    }
    // pass 1: cell histogram
#pragma unroll 4
    for (int i = tid; i < n; i += KM_THREADS)
        atomicAdd(&km_hist[g.cell(knn_ldg(c + i))], 1u);
    __syncthreads();
    // exclusive scan of the 32768 counters: warp w owns cells [1024 w, 1024 w + 1024)
    {
        uint32_t *mine = km_hist + warp * (KM_CELLS / NW);
        uint32_t total = 0;
        for (int ch = 0; ch < KM_CELLS / NW; ch += 32)
            total += mine[ch + lane];
        total = __reduce_add_sync(AMPC_FULL_MASK, total);
        if (lane == 0)
            sWarpTotal[warp] = total;
        __syncthreads();
        uint32_t v = sWarpTotal[lane], incl = v; // NW == 32: lane l holds warp l's total
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const uint32_t u = __shfl_up_sync(AMPC_FULL_MASK, incl, o);
            if (lane >= o) incl += u;
        }
        uint32_t running = __shfl_sync(AMPC_FULL_MASK, incl - v, warp); // cells before this warp's range
        for (int ch = 0; ch < KM_CELLS / NW; ch += 32) {
            const uint32_t cnt = mine[ch + lane];
            uint32_t inc = cnt;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const uint32_t u = __shfl_up_sync(AMPC_FULL_MASK, inc, o);
                if (lane >= o) inc += u;
            }
            mine[ch + lane] = running + inc - cnt;
            running += __shfl_sync(AMPC_FULL_MASK, inc, 31);
        }
    }
    __syncthreads();
    // pass 2: scatter (x, y, z, original index) to the cell's next free slot
#pragma unroll 4
    for (int i = tid; i < n; i += KM_THREADS) {
        const float4 p = knn_ldg(c + i);
        const uint32_t slot = atomicAdd(&km_hist[g.cell(p)], 1u);
        s[slot] = make_float4(p.x, p.y, p.z, __int_as_float(i));
    }
}

// ---- register-resident sorted top-k list: lane j holds entry j -----------------
struct TopK {
    double d;
    uint32_t i;
};
__device__ __forceinline__ void topk_insert(TopK &e, int k, double d, uint32_t i, int lane) {
    const bool before = lane < k && (e.d < d || (e.d == d && e.i < i));
    const int pos = __popc(__ballot_sync(AMPC_FULL_MASK, before));
    const double ud = __shfl_up_sync(AMPC_FULL_MASK, e.d, 1);
    const uint32_t ui = __shfl_up_sync(AMPC_FULL_MASK, e.i, 1);
    if (lane == pos) {
        e.d = d;
        e.i = i;
    } else if (lane > pos) {
        e.d = ud;
        e.i = ui;
    }
}

// (dist2, index) lexicographic order
__device__ __forceinline__ bool topk_less(double ad, uint32_t ai, double bd, uint32_t bi) {
    return ad < bd || (ad == bd && ai < bi);
}
// compare-exchange with the lane `stride` away: keep the smaller (keep_min) or larger element
__device__ __forceinline__ void cmpx(double &d, uint32_t &i, int stride, bool keep_min) {
    const double od = __shfl_xor_sync(AMPC_FULL_MASK, d, stride);
    const uint32_t oi = __shfl_xor_sync(AMPC_FULL_MASK, i, stride);
    const bool other_less = topk_less(od, oi, d, i);
    if (other_less == keep_min) {
        d = od;
        i = oi;
    }
}

// Dense tile (many candidates, k <= 16): bitonic sort of the tile's 64 (dist2, index) pairs
// (2 per lane: element e = lane + 32 r), then a 32-lane bitonic merge of the 16 best with the
// list.  ~500 instructions whatever the number of candidates, against ~40 per candidate for
// the serial insertion; no long dependent chain.
__device__ __noinline__ void merge_dense_tile(TopK &e, int k, double d0, uint32_t i0, double d1,
                                              uint32_t i1, int lane) {
#pragma unroll
    for (int size = 2; size <= 32; size <<= 1) {
#pragma unroll
        for (int stride = size >> 1; stride > 0; stride >>= 1) {
            // both registers hold elements with the same (e & size) for size <= 32
            const bool asc = (lane & size) == 0;
            const bool keep_min = ((lane & stride) == 0) == asc;
            cmpx(d0, i0, stride, keep_min);
            cmpx(d1, i1, stride, keep_min);
        }
    }
    // both registers are now sorted ascending across the lanes; reversing register 1 makes the
    // 64-element sequence (register 0, then register 1) bitonic for the final merge
    {
        const double rd = __shfl_sync(AMPC_FULL_MASK, d1, 31 - lane);
        const uint32_t ri = __shfl_sync(AMPC_FULL_MASK, i1, 31 - lane);
        d1 = rd;
        i1 = ri;
    }
    if (topk_less(d1, i1, d0, i0)) { // stride 32: within the lane, smaller to register 0
        const double td = d0;
        const uint32_t ti = i0;
        d0 = d1, i0 = i1, d1 = td, i1 = ti;
    }
#pragma unroll
    for (int stride = 16; stride > 0; stride >>= 1)
        cmpx(d0, i0, stride, (lane & stride) == 0); // register 0 now holds the 32 smallest, sorted
    // merge the 16 best of the tile with the list: lanes 0..15 list (ascending, +inf padded),
    // lanes 16..31 the tile's best 16 reversed -> bitonic; 5 compare-exchange stages sort it
    double md = __shfl_sync(AMPC_FULL_MASK, d0, 31 - lane);
    uint32_t mi = __shfl_sync(AMPC_FULL_MASK, i0, 31 - lane);
    if (lane < 16) {
        md = e.d;
        mi = e.i;
    }
#pragma unroll
    for (int stride = 16; stride > 0; stride >>= 1)
        cmpx(md, mi, stride, (lane & stride) == 0);
    e.d = lane < k ? md : INFINITY;
    e.i = lane < k ? mi : 0xffffffffu;
}

// scan one tile: exact distances of its (up to) 64 points, candidates into the list.
// by_w: `cloud` is a Morton-bucketed copy whose pad word holds the point's original index.
__device__ __forceinline__ void scan_tile(const float4 *cloud, int n, const TileGeom &g, int tile,
                                          double qx, double qy, double qz, TopK &e, double &kth, int k,
                                          int lane, bool by_w) {
    int i0, i1;
    g.points2(tile, lane, n, i0, i1);
    double d0 = INFINITY, d1 = INFINITY;
    uint32_t id0 = (uint32_t)i0, id1 = (uint32_t)i1;
    if (i0 >= 0) {
        const float4 p = knn_ldg(cloud + i0);
        d0 = knn_dist2(qx, qy, qz, p.x, p.y, p.z);
        if (by_w) id0 = (uint32_t)__float_as_int(p.w);
    }
    if (i1 >= 0) {
        const float4 p = knn_ldg(cloud + i1);
        d1 = knn_dist2(qx, qy, qz, p.x, p.y, p.z);
        if (by_w) id1 = (uint32_t)__float_as_int(p.w);
    }
    const bool c0 = i0 >= 0 && d0 <= kth, c1 = i1 >= 0 && d1 <= kth;
    unsigned m0 = __ballot_sync(AMPC_FULL_MASK, c0);
    unsigned m1 = __ballot_sync(AMPC_FULL_MASK, c1);
    if (k <= 16 && __popc(m0) + __popc(m1) >= KS_DENSE) {
        merge_dense_tile(e, k, c0 ? d0 : INFINITY, c0 ? id0 : 0xffffffffu, c1 ? d1 : INFINITY,
                         c1 ? id1 : 0xffffffffu, lane);
        kth = fmin(kth, __shfl_sync(AMPC_FULL_MASK, e.d, k - 1));
        return;
    }
    while (m0) {
        const int src = __ffs(m0) - 1;
        m0 &= m0 - 1;
        const double d = __shfl_sync(AMPC_FULL_MASK, d0, src);
        if (d <= kth) {
            topk_insert(e, k, d, __shfl_sync(AMPC_FULL_MASK, id0, src), lane);
            kth = fmin(kth, __shfl_sync(AMPC_FULL_MASK, e.d, k - 1));
        }
    }
    while (m1) {
        const int src = __ffs(m1) - 1;
        m1 &= m1 - 1;
        const double d = __shfl_sync(AMPC_FULL_MASK, d1, src);
        if (d <= kth) {
            topk_insert(e, k, d, __shfl_sync(AMPC_FULL_MASK, id1, src), lane);
            kth = fmin(kth, __shfl_sync(AMPC_FULL_MASK, e.d, k - 1));
        }
    }
}

// SearchForNearest's count rule (kd_tree_two.h:117-124) and the result layout
__device__ __forceinline__ void knn_write_result(const KnnParams &P, const float4 *cloud, int n,
                                                 int b, int q, const TopK &e, int lane) {
    const int k = P.k;
    const int cnt = (n < k) ? n : (n > k ? k : 0);
    if (lane < k) {
        const bool ok = lane < cnt;
        const int64_t o = ((int64_t)b * P.Q + q) * k + lane;
        if (P.idx) P.idx[o] = ok ? (int32_t)e.i : -1;
        if (P.dist2) P.dist2[o] = ok ? e.d : INFINITY;
        if (P.pts) {
            double *dst = P.pts + (int64_t)b * P.pts_inst_stride + (int64_t)q * P.pts_query_stride + 3 * lane;
            if (ok) {
                const float4 p = knn_ldg(cloud + e.i);
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
    if (P.count && lane == 0) P.count[(int64_t)b * P.Q + q] = cnt;
}

// Conservative single-precision bounds for the pruning tests (FP32 issues at twice the
// FP64 rate): every operation is rounded toward the safe side (query widened to the float
// interval [qlo, qhi], differences/products/sums rounded down for the lower bound and up
// for the upper bound) and a final factor (1 -+ 2^-20) absorbs the round-to-nearest of
// the double-precision distance itself.  lb32 <= knn_dist2(q, p) <= ub32 for every p in
// the box; only the tightness of the pruning differs from the double version.
struct QueryF {
    float xlo, xhi, ylo, yhi, zlo, zhi;
};
__device__ __forceinline__ QueryF make_queryf(double x, double y, double z) {
    QueryF q;
    q.xlo = __double2float_rd(x), q.xhi = __double2float_ru(x);
    q.ylo = __double2float_rd(y), q.yhi = __double2float_ru(y);
    q.zlo = __double2float_rd(z), q.zhi = __double2float_ru(z);
    return q;
}
__device__ __forceinline__ float knn_box_lb32(const QueryF &q, float lx, float ly, float lz, float hx,
                                              float hy, float hz) {
    const float ax = fmaxf(fmaxf(__fsub_rd(lx, q.xhi), __fsub_rd(q.xlo, hx)), 0.f);
    const float ay = fmaxf(fmaxf(__fsub_rd(ly, q.yhi), __fsub_rd(q.ylo, hy)), 0.f);
    const float az = fmaxf(fmaxf(__fsub_rd(lz, q.zhi), __fsub_rd(q.zlo, hz)), 0.f);
    const float r = __fadd_rd(__fadd_rd(__fmul_rd(ax, ax), __fmul_rd(ay, ay)), __fmul_rd(az, az));
    return __fmul_rd(r, 0.99999904632568359375f); // 1 - 2^-20
}
__device__ __forceinline__ float knn_box_ub32(const QueryF &q, float lx, float ly, float lz, float hx,
                                              float hy, float hz) {
    const float ax = fmaxf(fabsf(__fsub_ru(q.xhi, lx)), fabsf(__fsub_rd(q.xlo, hx)));
    const float ay = fmaxf(fabsf(__fsub_ru(q.yhi, ly)), fabsf(__fsub_rd(q.ylo, hy)));
    const float az = fmaxf(fabsf(__fsub_ru(q.zhi, lz)), fabsf(__fsub_rd(q.zlo, hz)));
    const float r = __fadd_ru(__fadd_ru(__fmul_ru(ax, ax), __fmul_ru(ay, ay)), __fmul_ru(az, az));
    return __fmul_ru(r, 1.00000095367431640625f); // 1 + 2^-20
}

__global__ void __launch_bounds__(KS_WARPS * 32, 6)
knn_search_kernel(const KnnParams P) {
    __shared__ float sLB[KS_WARPS][KS_CHUNK];        // conservative lower bound per tile (NaN = done)
    __shared__ unsigned short sCand[KS_WARPS][KS_CHUNK]; // compacted list of tiles worth visiting
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int q = blockIdx.y * KS_WARPS + warp, b = blockIdx.x, seg = blockIdx.z; // b on x: no 65535 limit
    if (q >= P.Q || (P.active && P.active[b] == 0))
        return;
    const int k = P.k;
    const int scene = P.scene_of ? P.scene_of[b] : b;
    const int n = P.counts[scene];
    const float4 *cloud = P.clouds + (int64_t)scene * P.slot_points;
    const float4 *boxes = P.boxes + (int64_t)scene * P.slot_tiles * 2;
    const double *qp = P.queries + ((int64_t)b * P.Q + q) * 3;
    const double qx = qp[0], qy = qp[1], qz = qp[2];
    const int lay = tile_layout(n, P.layout[scene], P.slot_tiles);
    const bool by_w = lay < 0; // tiles over the Morton-bucketed copy; indices come from its pad word
    const float4 *tsrc = by_w ? P.sorted + (int64_t)scene * P.slot_points : cloud;
    const TileGeom g(n, lay);
    const int n_tiles = g.n_tiles;
    const int per = (n_tiles + P.segs - 1) / P.segs;
    const int t_begin = seg * per, t_end = min(n_tiles, t_begin + per);
    float *lbuf = sLB[warp];
    const QueryF qf = make_queryf(qx, qy, qz);
    unsigned short *cand = sCand[warp];
    TopK e{INFINITY, 0xffffffffu};
    // `bound`: no point farther than this can be among the k nearest.  It is the min of
    // the list's k-th entry and of the farthest-corner distance of any tile seen that holds
    // at least k points (that tile alone already has k points within that distance).
    double bound = INFINITY;

    for (int c0 = t_begin; c0 < t_end; c0 += KS_CHUNK) {
        const int cn = min(KS_CHUNK, t_end - c0);
        double ubmin = INFINITY;
        for (int j = lane; j < cn; j += 32) {
            const float4 a = knn_ldg(boxes + 2 * (int64_t)(c0 + j));
            const float4 h = knn_ldg(boxes + 2 * (int64_t)(c0 + j) + 1);
            lbuf[j] = knn_box_lb32(qf, a.x, a.y, a.z, a.w, h.x, h.y);
            if (__float_as_int(h.z) >= k) // the tile alone holds >= k points within its farthest corner
                ubmin = fmin(ubmin, (double)knn_box_ub32(qf, a.x, a.y, a.z, a.w, h.x, h.y));
        }
        bound = fmin(bound, warp_min(ubmin));
        __syncwarp();
        // compact the tiles that can still matter
        int m = 0;
        for (int j0 = 0; j0 < cn; j0 += 32) {
            const bool keep = (j0 + lane < cn) && (double)lbuf[j0 + lane] <= bound;
            const unsigned mk = __ballot_sync(AMPC_FULL_MASK, keep);
            if (keep)
                cand[m + __popc(mk & ((1u << lane) - 1u))] = (unsigned short)(j0 + lane);
            m += __popc(mk);
        }
        __syncwarp();
        // a few best-first picks tighten the bound to (nearly) the true k-th distance ...
        for (int pick = 0; pick < KS_PICKS; ++pick) {
            double best = INFINITY;
            int best_c = -1;
            for (int c = lane; c < m; c += 32) {
                const double lb = (double)lbuf[cand[c]];
                if (lb < best) { // NaN (already visited) never wins
                    best = lb;
                    best_c = c;
                }
            }
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) {
                const double ob = __shfl_xor_sync(AMPC_FULL_MASK, best, o);
                const int oc = __shfl_xor_sync(AMPC_FULL_MASK, best_c, o);
                if (ob < best || (ob == best && oc >= 0 && (best_c < 0 || oc < best_c))) {
                    best = ob;
                    best_c = oc;
                }
            }
            if (best_c < 0 || !(best <= bound))
                break;
            const int j = cand[best_c];
            double kth = fmin(bound, __shfl_sync(AMPC_FULL_MASK, e.d, k - 1));
            scan_tile(tsrc, n, g, c0 + j, qx, qy, qz, e, kth, k, lane, by_w);
            bound = fmin(bound, kth);
            __syncwarp(); // every lane has finished reading lbuf for this pick
            if (lane == 0)
                lbuf[j] = NAN;
            __syncwarp();
        }
        // ... then one sweep over the remaining candidates in storage order
        for (int cb = 0; cb < m; cb += 32) {
            const int j = (cb + lane < m) ? (int)cand[cb + lane] : 0;
            const double lb = (cb + lane < m) ? (double)lbuf[j] : NAN;
            unsigned mk = __ballot_sync(AMPC_FULL_MASK, lb <= bound);
            while (mk) {
                const int src = __ffs(mk) - 1;
                mk &= mk - 1;
                const double lbj = __shfl_sync(AMPC_FULL_MASK, lb, src);
                if (lbj <= bound) { // the bound may have tightened since the ballot
                    const int jj = __shfl_sync(AMPC_FULL_MASK, j, src);
                    double kth = fmin(bound, __shfl_sync(AMPC_FULL_MASK, e.d, k - 1));
                    scan_tile(tsrc, n, g, c0 + jj, qx, qy, qz, e, kth, k, lane, by_w);
                    bound = fmin(bound, kth);
                }
            }
        }
        __syncwarp();
    }
    if (P.segs == 1) {
        knn_write_result(P, cloud, n, b, q, e, lane);
    } else if (lane < k) {
        const int64_t o = ((((int64_t)b * P.Q + q) * P.segs) + seg) * k + lane;
        P.ws_d[o] = e.d;
        P.ws_i[o] = e.i;
    }
}

// ---- group boxes: one warp per group reduces the boxes of its (up to) 16 tiles ------------
__global__ void __launch_bounds__(256)
group_boxes_kernel(const float4 *__restrict__ boxes, float4 *__restrict__ gboxes, const int32_t *__restrict__ counts,
                   int64_t slot_tiles, int64_t slot_groups, const int32_t *__restrict__ layout, int first_scene) {
    const int scene = first_scene + blockIdx.x;
    const int lane = threadIdx.x & 31;
    const int n = counts[scene];
    const TileGeom tg(n, tile_layout(n, layout[scene], slot_tiles));
    const GroupGeom gg(tg);
    const float4 *bx = boxes + (int64_t)scene * slot_tiles * 2;
    // the grid is sized for the common layouts; a scene with more groups is covered by the stride
    for (int g = blockIdx.y * 8 + (threadIdx.x >> 5); g < gg.n_groups; g += gridDim.y * 8) {
        int ty, tx;
        const int t = lane < KG_TILES ? gg.tile(g, lane, ty, tx) : -1;
        float lx = INFINITY, ly = INFINITY, lz = INFINITY, hx = -INFINITY, hy = -INFINITY, hz = -INFINITY;
        int cnt = 0;
        if (t >= 0) {
            const float4 a = bx[2 * (int64_t)t], h = bx[2 * (int64_t)t + 1];
            lx = a.x, ly = a.y, lz = a.z, hx = a.w, hy = h.x, hz = h.y, cnt = __float_as_int(h.z);
        }
#pragma unroll
        for (int o = 8; o > 0; o >>= 1) {
            lx = fminf(lx, __shfl_xor_sync(AMPC_FULL_MASK, lx, o));
            ly = fminf(ly, __shfl_xor_sync(AMPC_FULL_MASK, ly, o));
            lz = fminf(lz, __shfl_xor_sync(AMPC_FULL_MASK, lz, o));
            hx = fmaxf(hx, __shfl_xor_sync(AMPC_FULL_MASK, hx, o));
            hy = fmaxf(hy, __shfl_xor_sync(AMPC_FULL_MASK, hy, o));
            hz = fmaxf(hz, __shfl_xor_sync(AMPC_FULL_MASK, hz, o));
            cnt += __shfl_xor_sync(AMPC_FULL_MASK, cnt, o);
        }
        if (lane == 0) {
            float4 *dst = gboxes + ((int64_t)scene * slot_groups + g) * 2;
            dst[0] = make_float4(lx, ly, lz, hx);
            dst[1] = make_float4(hy, hz, __int_as_float(cnt), 0.f);
        }
    }
}

// the (up to) 64 points of a tile, two per lane, loaded but not yet used
struct TileLoad {
    float4 p0, p1;
    int i0, i1;
};
__device__ __forceinline__ TileLoad load_tile(const float4 *cloud, int n, const TileGeom &g, int tile, int ty, int tx,
                                              int lane) {
    TileLoad L;
    if (g.row_w > 0) { // patch coordinates known: no division
        const int col = 8 * tx + (lane & 7);
        const int a = (8 * ty + (lane >> 3)) * g.row_w + col, b = a + 4 * g.row_w;
        const bool ok = col < g.row_w;
        L.i0 = ok && a < n ? a : -1;
        L.i1 = ok && b < n ? b : -1;
    } else {
        const int a = tile * KT_TILE + lane, b = a + 32;
        L.i0 = a < n ? a : -1;
        L.i1 = b < n ? b : -1;
    }
    L.p0 = L.p1 = make_float4(0.f, 0.f, 0.f, 0.f);
    if (L.i0 >= 0) L.p0 = knn_ldg(cloud + L.i0);
    if (L.i1 >= 0) L.p1 = knn_ldg(cloud + L.i1);
    return L;
}
// exact distances of a loaded tile's points, candidates into the list (as scan_tile)
__device__ __forceinline__ void consume_tile(const TileLoad &L, double qx, double qy, double qz, TopK &e, double &kth,
                                             int k, int lane, bool by_w) {
    double d0 = INFINITY, d1 = INFINITY;
    uint32_t id0 = (uint32_t)L.i0, id1 = (uint32_t)L.i1;
    if (L.i0 >= 0) {
        d0 = knn_dist2(qx, qy, qz, L.p0.x, L.p0.y, L.p0.z);
        if (by_w) id0 = (uint32_t)__float_as_int(L.p0.w);
    }
    if (L.i1 >= 0) {
        d1 = knn_dist2(qx, qy, qz, L.p1.x, L.p1.y, L.p1.z);
        if (by_w) id1 = (uint32_t)__float_as_int(L.p1.w);
    }
    const bool c0 = L.i0 >= 0 && d0 <= kth, c1 = L.i1 >= 0 && d1 <= kth;
    unsigned m0 = __ballot_sync(AMPC_FULL_MASK, c0);
    unsigned m1 = __ballot_sync(AMPC_FULL_MASK, c1);
    if (k <= 16 && __popc(m0) + __popc(m1) >= KS_DENSE) {
        merge_dense_tile(e, k, c0 ? d0 : INFINITY, c0 ? id0 : 0xffffffffu, c1 ? d1 : INFINITY,
                         c1 ? id1 : 0xffffffffu, lane);
        kth = fmin(kth, __shfl_sync(AMPC_FULL_MASK, e.d, k - 1));
        return;
    }
    while (m0) {
        const int src = __ffs(m0) - 1;
        m0 &= m0 - 1;
        const double d = __shfl_sync(AMPC_FULL_MASK, d0, src);
        if (d <= kth) {
            topk_insert(e, k, d, __shfl_sync(AMPC_FULL_MASK, id0, src), lane);
            kth = fmin(kth, __shfl_sync(AMPC_FULL_MASK, e.d, k - 1));
        }
    }
    while (m1) {
        const int src = __ffs(m1) - 1;
        m1 &= m1 - 1;
        const double d = __shfl_sync(AMPC_FULL_MASK, d1, src);
        if (d <= kth) {
            topk_insert(e, k, d, __shfl_sync(AMPC_FULL_MASK, id1, src), lane);
            kth = fmin(kth, __shfl_sync(AMPC_FULL_MASK, e.d, k - 1));
        }
    }
}

// One group: lower bounds of its tiles (lanes 0..15), the tiles that can still hold a neighbour
// visited nearest-bound first; the points of the next tile are in flight while the current one
// is merged.  `bound` as in the kernel below.
__device__ __forceinline__ void visit_group(const KnnParams &P, const float4 *boxes, const float4 *tsrc, int n,
                                            const TileGeom &g, const GroupGeom &gg, int G, const QueryF &qf, double qx,
                                            double qy, double qz, TopK &e, double &bound, int k, int lane, bool by_w) {
    int ty, tx;
    const int t = lane < KG_TILES ? gg.tile(G, lane, ty, tx) : -1;
    float lb = INFINITY, ub = INFINITY;
    if (t >= 0) {
        const float4 a = knn_ldg(boxes + 2 * (int64_t)t), h = knn_ldg(boxes + 2 * (int64_t)t + 1);
        lb = knn_box_lb32(qf, a.x, a.y, a.z, a.w, h.x, h.y);
        if (__float_as_int(h.z) >= k) // the tile alone holds >= k points within its farthest corner
            ub = knn_box_ub32(qf, a.x, a.y, a.z, a.w, h.x, h.y);
    }
    bound = fmin(bound, (double)warp_min_nonneg(ub));
    // lanes with a tile still to visit; an empty tile (lb = +inf) never is -- `bound` itself is
    // +inf while fewer than k points have been seen, so "lb <= bound" alone would not end
    bool live = t >= 0 && lb < INFINITY;
    if (!live) lb = INFINITY;
    auto argmin = [&](float &best, int &bl) { // nearest live tile, ties to the lowest lane
        best = warp_argmin_nonneg(lb, lane, bl);
    };
    float best;
    int bl;
    if (!__any_sync(AMPC_FULL_MASK, live))
        return;
    argmin(best, bl);
    if (!((double)best <= bound))
        return;
    TileLoad A = load_tile(tsrc, n, g, __shfl_sync(AMPC_FULL_MASK, t, bl), __shfl_sync(AMPC_FULL_MASK, ty, bl),
                           __shfl_sync(AMPC_FULL_MASK, tx, bl), lane);
    if (lane == bl) lb = INFINITY, live = false;
    for (;;) {
        const bool any_live = __any_sync(AMPC_FULL_MASK, live);
        argmin(best, bl);
        const bool more = any_live && (double)best <= bound;
        TileLoad B = A;
        if (more)
            B = load_tile(tsrc, n, g, __shfl_sync(AMPC_FULL_MASK, t, bl), __shfl_sync(AMPC_FULL_MASK, ty, bl),
                          __shfl_sync(AMPC_FULL_MASK, tx, bl), lane);
        double kth = fmin(bound, __shfl_sync(AMPC_FULL_MASK, e.d, k - 1));
        consume_tile(A, qx, qy, qz, e, kth, k, lane, by_w);
        bound = fmin(bound, kth);
        if (!more || !((double)best <= bound))
            break; // every remaining tile of the group has a bound at least `best`
        if (lane == bl) lb = INFINITY, live = false;
        A = B;
    }
}

// ---- exact k-NN, two-level: one warp per (instance, query).  Group lower bounds first (a 50k-
// point cloud has ~56 groups against 782 tiles), two best-first group picks to tighten the
// bound, then one sweep over the groups in storage order.  `bound`: no point farther than this
// can be among the k nearest = min of the list's k-th entry and of the farthest-corner distance of
// any group or tile seen that holds at least k points.
constexpr int KS_GCHUNK = 1024; // group lower bounds kept in shared memory per warp (16384 tiles = 1M points)
constexpr int KS_GPICKS = 2;
__global__ void __launch_bounds__(KS_WARPS * 32, 8)
knn_search2_kernel(const KnnParams P) {
    extern __shared__ float sGLB_all[]; // KS_WARPS x P.gchunk group lower bounds
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int q = blockIdx.y * KS_WARPS + warp, b = blockIdx.x, seg = blockIdx.z;
    if (q >= P.Q || (P.active && P.active[b] == 0))
        return;
    const int k = P.k;
    const int scene = P.scene_of ? P.scene_of[b] : b;
    const int n = P.counts[scene];
    const float4 *cloud = P.clouds + (int64_t)scene * P.slot_points;
    const float4 *boxes = P.boxes + (int64_t)scene * P.slot_tiles * 2;
    const float4 *gboxes = P.gboxes + (int64_t)scene * P.slot_groups * 2;
    const double *qp = P.queries + ((int64_t)b * P.Q + q) * 3;
    const double qx = qp[0], qy = qp[1], qz = qp[2];
    const int lay = tile_layout(n, P.layout[scene], P.slot_tiles);
    const bool by_w = lay < 0;
    const float4 *tsrc = by_w ? P.sorted + (int64_t)scene * P.slot_points : cloud;
    const TileGeom g(n, lay);
    const GroupGeom gg(g);
    const int per = (gg.n_groups + P.segs - 1) / P.segs;
    const int g_begin = seg * per, g_end = min(gg.n_groups, g_begin + per);
    float *glb = sGLB_all + warp * P.gchunk;
    const QueryF qf = make_queryf(qx, qy, qz);
    TopK e{INFINITY, 0xffffffffu};
    double bound = INFINITY;
    for (int c0 = g_begin; c0 < g_end; c0 += P.gchunk) {
        const int cn = min(P.gchunk, g_end - c0);
        float ubmin = INFINITY;
        for (int j = lane; j < cn; j += 32) {
            const float4 a = knn_ldg(gboxes + 2 * (int64_t)(c0 + j)), h = knn_ldg(gboxes + 2 * (int64_t)(c0 + j) + 1);
            glb[j] = knn_box_lb32(qf, a.x, a.y, a.z, a.w, h.x, h.y);
            if (__float_as_int(h.z) >= k)
                ubmin = fminf(ubmin, knn_box_ub32(qf, a.x, a.y, a.z, a.w, h.x, h.y));
        }
        bound = fmin(bound, (double)warp_min_nonneg(ubmin));
        __syncwarp();
        for (int pick = 0; pick < KS_GPICKS; ++pick) {
            float best = INFINITY;
            int bj = -1;
            for (int j = lane; j < cn; j += 32) {
                const float v = glb[j];
                if (v < best) best = v, bj = j; // NaN (visited) never wins
            }
            best = warp_argmin_nonneg(best, bj < 0 ? 0x7fffffff : bj, bj); // ties to the lowest group
            if (bj == 0x7fffffff || !((double)best <= bound))
                break;
            visit_group(P, boxes, tsrc, n, g, gg, c0 + bj, qf, qx, qy, qz, e, bound, k, lane, by_w);
            __syncwarp();
            if (lane == 0) glb[bj] = NAN;
            __syncwarp();
        }
        for (int jb = 0; jb < cn; jb += 32) {
            const float lbv = (jb + lane < cn) ? glb[jb + lane] : NAN;
            unsigned mk = __ballot_sync(AMPC_FULL_MASK, (double)lbv <= bound);
            while (mk) {
                const int src = __ffs(mk) - 1;
                mk &= mk - 1;
                const float lbj = __shfl_sync(AMPC_FULL_MASK, lbv, src);
                if ((double)lbj <= bound) // the bound may have tightened since the ballot
                    visit_group(P, boxes, tsrc, n, g, gg, c0 + jb + src, qf, qx, qy, qz, e, bound, k, lane, by_w);
            }
        }
        __syncwarp();
    }
    if (P.segs == 1) {
        knn_write_result(P, cloud, n, b, q, e, lane);
    } else if (lane < k) {
        const int64_t o = ((((int64_t)b * P.Q + q) * P.segs) + seg) * k + lane;
        P.ws_d[o] = e.d;
        P.ws_i[o] = e.i;
    }
}

// segs > 1: fold the per-segment lists of each (instance, query); one warp each
__global__ void __launch_bounds__(KS_WARPS * 32)
knn_merge_kernel(const KnnParams P) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int q = blockIdx.y * KS_WARPS + warp, b = blockIdx.x;
    if (q >= P.Q || (P.active && P.active[b] == 0))
        return;
    const int k = P.k;
    const int scene = P.scene_of ? P.scene_of[b] : b;
    const int n = P.counts[scene];
    const float4 *cloud = P.clouds + (int64_t)scene * P.slot_points;
    const int64_t o = ((int64_t)b * P.Q + q) * P.segs * k;
    TopK e{INFINITY, 0xffffffffu};
    if (lane < k) {
        e.d = P.ws_d[o + lane];
        e.i = P.ws_i[o + lane];
    }
    for (int s = 1; s < P.segs; ++s)
        for (int j = 0; j < k; ++j) {
            const double d = P.ws_d[o + (int64_t)s * k + j];
            const uint32_t i = P.ws_i[o + (int64_t)s * k + j];
            const double kd = __shfl_sync(AMPC_FULL_MASK, e.d, k - 1);
            const uint32_t ki = __shfl_sync(AMPC_FULL_MASK, e.i, k - 1);
            if (!(d < kd || (d == kd && i < ki)))
                break; // sorted source: nothing further can enter
            topk_insert(e, k, d, i, lane);
        }
    knn_write_result(P, cloud, n, b, q, e, lane);
}

} // namespace ampc
