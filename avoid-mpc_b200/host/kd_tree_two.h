// Drop-in replacement for the reference's KD-tree wrapper
// (roswrapper/ros/src/avoid_mpc/include/kd_tree_two.h:53-144): same class template name,
// same public methods and the same public result vectors, backed by libampc's exact GPU
// k-NN instead of nanoflann.  Differences: results among EXACT distance ties come in
// (dist2, index) order (nanoflann's tie order is traversal dependent); k <= 32.
#ifndef KD_TREE_TWO_H
#define KD_TREE_TWO_H
#include "../../include/ampc.h"
#include "pcl_compat.h"

#include <memory>
#include <stdexcept>
#include <string>
#include <vector>

template <typename T> struct PointCloudTwo { // kd_tree_two.h:11-51 (host copy of the points)
    std::vector<pcl::PointXYZ> pts;
    inline size_t kdtree_get_point_count() const { return pts.size(); }
};

template <typename num_t> class KDTreeTwo {
public:
    std::vector<pcl::PointXYZ> closest_pts;
    std::vector<num_t> squared_distances;
    std::vector<int> indices;
    std::vector<int> colors;

    KDTreeTwo() {
        for (int i = 0; i < 3; i++)
            colors.push_back(rand() % 256);
    }
    void InitializeNew(typename pcl::PointCloud<pcl::PointXYZ>::Ptr const &xyz_cloud_new) {
        Initialize(xyz_cloud_new, true);
    }
    void AddToKDTree(typename pcl::PointCloud<pcl::PointXYZ>::Ptr const &xyz_cloud_new) {
        Initialize(xyz_cloud_new, false);
    }
    void Clear() { cloud.pts.clear(); }

    void Initialize(typename pcl::PointCloud<pcl::PointXYZ>::Ptr const &xyz_cloud_new, bool clear) {
        if (clear)
            cloud.pts.clear();
        for (const auto &p : xyz_cloud_new->points) // kd_tree_two.h:99-101: drop NaN x
            if (!(p.x != p.x))
                cloud.pts.push_back(p);
        Upload();
    }

    void SearchForNearest(num_t x, num_t y, num_t z, int n) {
        closest_pts.clear();
        squared_distances.clear();
        indices.clear();
        if (cloud.pts.size() == 0 || n <= 0)
            return;
        if (n > 32)
            throw std::runtime_error("KDTreeTwo(GPU): k > 32 is not supported");
        const double q[3] = {(double)x, (double)y, (double)z};
        std::vector<int32_t> idx(n);
        std::vector<double> d2(n);
        int32_t cnt = 0;
        if (ampc_knn_batch(h.get(), AMPC_CLOUD_OBSTACLE, 1, nullptr, q, 1, n, idx.data(), d2.data(),
                           nullptr, &cnt) != AMPC_OK)
            throw std::runtime_error(std::string("KDTreeTwo(GPU): ") + ampc_last_error(h.get()));
        for (int i = 0; i < cnt; i++) { // cnt follows kd_tree_two.h:117-124 (0 when size == n)
            closest_pts.push_back(cloud.pts[idx[i]]);
            squared_distances.push_back((num_t)d2[i]);
            indices.push_back(idx[i]);
        }
    }
    PointCloudTwo<num_t> const &GetPointCloud() { return cloud; }
    std::vector<int> const &GetColors() { return colors; }

private:
    void Upload() {
        const int n = (int)cloud.pts.size();
        if (!h || n > capacity) {
            capacity = n < 4096 ? 4096 : n + n / 2;
            ampc_config cfg{};
            cfg.N = 1, cfg.K = 1, cfg.dt = 1.0, cfg.max_batch = 1, cfg.max_scenes = 1;
            cfg.max_points = capacity, cfg.device = 0;
            ampc_handle *raw = nullptr;
            if (ampc_create(&cfg, &raw) != AMPC_OK)
                throw std::runtime_error(std::string("KDTreeTwo(GPU): ") + ampc_last_error(nullptr));
            h.reset(raw, ampc_destroy);
        }
        // InitializeNew accepts points in any order (kd_tree_two.h:88-106): larger clouds are indexed
        // over a Morton-bucketed copy; indices stay those of cloud.pts
        if (ampc_cloud_set_layout(h.get(), AMPC_CLOUD_OBSTACLE, n >= 8192 ? AMPC_LAYOUT_SORT : AMPC_LAYOUT_UNORGANISED) != AMPC_OK)
            throw std::runtime_error(std::string("KDTreeTwo(GPU): ") + ampc_last_error(h.get()));
        if (ampc_cloud_set(h.get(), 0, AMPC_CLOUD_OBSTACLE, cloud.pts.data(), n, 16) != AMPC_OK)
            throw std::runtime_error(std::string("KDTreeTwo(GPU): ") + ampc_last_error(h.get()));
    }
    PointCloudTwo<num_t> cloud;
    std::shared_ptr<ampc_handle> h;
    int capacity = 0;
};
#endif
