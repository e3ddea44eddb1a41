// TEST INFRASTRUCTURE ONLY (oracle/): C-ABI shim around the REFERENCE's own
// KDTreeTwo<double> (roswrapper/ros/src/avoid_mpc/include/kd_tree_two.h:53-144)
// and its vendored nanoflann 1.5.5 (include/nanoflann_two.hpp).  The reference
// headers are included from where they lie under /root/reference (never copied
// into this repo); only the PCL point types are stubbed (oracle/pcl_stub/).
//
// Built by oracle/Makefile into oracle/_ref/libampc_ref_kdtree.so (git-ignored,
// travels to the GPU box as a prebuilt binary).  Used ONLY by tests/ (to pin
// the k-NN restatement and mint tests/golden/*) and by bench.py's CPU-baseline
// legs.  The product path never loads it.
#include "kd_tree_two.h" // the reference file, via -I<reference>/include

#include <cstdint>
#include <cstring>

extern "C" {

struct ref_tree {
    KDTreeTwo<double> tree;
};

// xyz16: n records of 16 bytes (float x,y,z,pad) == pcl::PointXYZ layout.
// Mirrors FrameKDMap::AddVertex -> KDTreeTwo::InitializeNew (FrameKDMap.cpp:44-47).
ref_tree *ref_tree_create(const void *xyz16, int64_t n) {
    auto cloud = std::make_shared<pcl::PointCloud<pcl::PointXYZ>>();
    cloud->points.resize(static_cast<size_t>(n));
    if (n > 0)
        std::memcpy(static_cast<void *>(cloud->points.data()), xyz16, static_cast<size_t>(n) * 16);
    ref_tree *t = new ref_tree();
    t->tree.InitializeNew(cloud);
    return t;
}

void ref_tree_destroy(ref_tree *t) {
    delete t;
}

int64_t ref_tree_size(ref_tree *t) {
    return static_cast<int64_t>(t->tree.GetPointCloud().pts.size());
}

// KDTreeTwo::SearchForNearest (kd_tree_two.h:108-133) incl. its quirks
// (0 results when the cloud holds exactly k points).  Returns num_results.
int ref_tree_search(ref_tree *t, double x, double y, double z, int k, int32_t *idx_out,
                    double *dist2_out, float *pts_out /* 3 floats per result, may be null */) {
    t->tree.SearchForNearest(x, y, z, k);
    const int m = static_cast<int>(t->tree.indices.size());
    for (int i = 0; i < m; ++i) {
        idx_out[i] = t->tree.indices[i];
        dist2_out[i] = t->tree.squared_distances[i];
        if (pts_out) {
            pts_out[3 * i + 0] = t->tree.closest_pts[i].x;
            pts_out[3 * i + 1] = t->tree.closest_pts[i].y;
            pts_out[3 * i + 2] = t->tree.closest_pts[i].z;
        }
    }
    return m;
}

// Batch helper for baseline timing: Q queries, results padded to k per query.
void ref_tree_search_batch(ref_tree *t, const double *q, int Q, int k, int32_t *idx_out,
                           double *dist2_out, int32_t *count_out) {
    for (int i = 0; i < Q; ++i) {
        count_out[i] = ref_tree_search(t, q[3 * i], q[3 * i + 1], q[3 * i + 2], k,
                                       idx_out + (size_t)i * k, dist2_out + (size_t)i * k, nullptr);
    }
}

} // extern "C"
