// libampc_multi.so -- include/ampc_multi.h: scene-sharded rounds over several devices of one
// process.  One worker thread per device drives that device's ampc_handle through the host-buffer
// C-ABI; the costs are exchanged with ncclAllGather on one communicator per device.  No kernel
// lives here: everything on the data path is libampc.so's.
#include "../../include/ampc_multi.h"

#include <cuda_runtime.h>
#include <nccl.h>

#include <cmath>
#include <cstring>
#include <limits>
#include <mutex>
#include <string>
#include <thread>
#include <vector>

struct ampc_multi {
    ampc_config cfg{};
    int n = 0;
    std::vector<int> device;
    std::vector<ampc_handle *> h;
    std::vector<ncclComm_t> comm;
    std::vector<cudaStream_t> st;
    std::vector<double *> costs_d;    // [block_cap]
    std::vector<double *> gathered_d; // [n * block_cap]
    std::vector<double *> costs_h;    // pinned staging, [block_cap]
    int block_cap = 0;
    std::string err;
    std::mutex err_mtx;
};

namespace {

std::string g_create_err;

int block_of(int batch, int n) { return (batch + n - 1) / n; }

int set_err(ampc_multi *m, int code, const std::string &msg) {
    if (m) {
        std::lock_guard<std::mutex> g(m->err_mtx);
        if (m->err.empty()) m->err = msg;
    } else {
        g_create_err = msg;
    }
    return code;
}

// run f(i) on one thread per device, return the first non-zero code
template <class F> int per_device(ampc_multi *m, F f) {
    std::vector<int> rc(m->n, 0);
    std::vector<std::thread> th;
    for (int i = 1; i < m->n; ++i) th.emplace_back([&, i] { rc[i] = f(i); });
    rc[0] = f(0);
    for (auto &t : th) t.join();
    for (int i = 0; i < m->n; ++i)
        if (rc[i]) return rc[i];
    return AMPC_OK;
}

} // namespace

extern "C" {

int ampc_multi_create(const ampc_config *cfg, const int32_t *devices, int32_t n_devices, ampc_multi **out) {
    if (!cfg || !devices || !out || n_devices < 1) return set_err(nullptr, AMPC_ERR_INVALID, "null config / device list");
    ampc_multi *m = new ampc_multi;
    m->cfg = *cfg;
    m->n = n_devices;
    m->device.assign(devices, devices + n_devices);
    m->block_cap = block_of(cfg->max_batch, n_devices);
    m->h.assign(n_devices, nullptr);
    m->st.assign(n_devices, nullptr);
    m->costs_d.assign(n_devices, nullptr);
    m->gathered_d.assign(n_devices, nullptr);
    m->costs_h.assign(n_devices, nullptr);
    auto bail = [&](int code, const std::string &msg) {
        g_create_err = msg;
        ampc_multi_destroy(m);
        return code;
    };
    for (int i = 0; i < n_devices; ++i) {
        ampc_config c = *cfg;
        c.device = devices[i];
        c.max_batch = m->block_cap;
        c.max_scenes = block_of(cfg->max_scenes > 0 ? cfg->max_scenes : cfg->max_batch, n_devices);
        int rc = ampc_create(&c, &m->h[i]);
        if (rc) return bail(rc, std::string("ampc_create on device ") + std::to_string(devices[i]) + ": " + ampc_last_error(nullptr));
        if (cudaSetDevice(devices[i]) != cudaSuccess || cudaStreamCreateWithFlags(&m->st[i], cudaStreamNonBlocking) != cudaSuccess ||
            cudaMalloc(&m->costs_d[i], sizeof(double) * m->block_cap) != cudaSuccess ||
            cudaMalloc(&m->gathered_d[i], sizeof(double) * m->block_cap * n_devices) != cudaSuccess ||
            cudaMallocHost(&m->costs_h[i], sizeof(double) * m->block_cap) != cudaSuccess)
            return bail(AMPC_ERR_CUDA, std::string("device buffers: ") + cudaGetErrorString(cudaGetLastError()));
    }
    m->comm.assign(n_devices, nullptr);
    ncclResult_t nr = ncclCommInitAll(m->comm.data(), n_devices, m->device.data());
    if (nr != ncclSuccess) {
        m->comm.clear();
        return bail(AMPC_ERR_CUDA, std::string("ncclCommInitAll: ") + ncclGetErrorString(nr));
    }
    *out = m;
    return AMPC_OK;
}

void ampc_multi_destroy(ampc_multi *m) {
    if (!m) return;
    for (size_t i = 0; i < m->comm.size(); ++i)
        if (m->comm[i]) ncclCommDestroy(m->comm[i]);
    for (int i = 0; i < m->n; ++i) {
        cudaSetDevice(m->device[i]);
        if (m->costs_d[i]) cudaFree(m->costs_d[i]);
        if (m->gathered_d[i]) cudaFree(m->gathered_d[i]);
        if (m->costs_h[i]) cudaFreeHost(m->costs_h[i]);
        if (m->st[i]) cudaStreamDestroy(m->st[i]);
        if (m->h[i]) ampc_destroy(m->h[i]);
    }
    delete m;
}

const char *ampc_multi_last_error(const ampc_multi *m) { return m ? m->err.c_str() : g_create_err.c_str(); }
int32_t ampc_multi_device_count(const ampc_multi *m) { return m ? m->n : 0; }
ampc_handle *ampc_multi_handle(ampc_multi *m, int32_t i) { return (m && i >= 0 && i < m->n) ? m->h[i] : nullptr; }
int32_t ampc_multi_block(const ampc_multi *m, int32_t batch) { return m ? block_of(batch, m->n) : 0; }

void ampc_multi_shard(const ampc_multi *m, int32_t batch, int32_t i, int32_t *first, int32_t *count) {
    const int per = block_of(batch, m->n);
    int f = i * per;
    if (f > batch) f = batch;
    int c = batch - f < per ? batch - f : per;
    if (first) *first = f;
    if (count) *count = c;
}

int ampc_multi_cloud_set_layout(ampc_multi *m, int32_t kind, int32_t row_width) {
    if (!m) return AMPC_ERR_INVALID;
    for (int i = 0; i < m->n; ++i) {
        int rc = ampc_cloud_set_layout(m->h[i], kind, row_width);
        if (rc) return set_err(m, rc, ampc_last_error(m->h[i]));
    }
    return AMPC_OK;
}

int ampc_multi_cloud_set_batch(ampc_multi *m, int32_t kind, int32_t n_scenes, const void *xyz_host,
                               const int32_t *counts, int64_t scene_stride_bytes, int32_t stride_bytes) {
    if (!m || !xyz_host || !counts || n_scenes < 1) return set_err(m, AMPC_ERR_INVALID, "null cloud batch");
    m->err.clear();
    return per_device(m, [&](int i) {
        int first, count;
        ampc_multi_shard(m, n_scenes, i, &first, &count);
        if (count < 1) return (int)AMPC_OK;
        int rc = ampc_cloud_set_batch(m->h[i], kind, 0, count, (const char *)xyz_host + (int64_t)first * scene_stride_bytes,
                                      counts + first, scene_stride_bytes, stride_bytes);
        return rc ? set_err(m, rc, ampc_last_error(m->h[i])) : (int)AMPC_OK;
    });
}

int ampc_multi_round_batch(ampc_multi *m, int32_t batch, const double *x0, const double *ref, const double *pos_x,
                           double speed, double safety_distance, double *w_inout, ampc_solve_info *info_out,
                           int32_t *need_replan_out, double *costs_all_out) {
    if (!m || !x0 || !ref || !w_inout || !info_out || batch < 1) return set_err(m, AMPC_ERR_INVALID, "null round arguments");
    if (batch > m->cfg.max_batch) return set_err(m, AMPC_ERR_CAPACITY, "batch exceeds max_batch");
    m->err.clear();
    const int N = m->cfg.N, nw = 14 * N + 10, per = block_of(batch, m->n);
    int rc = per_device(m, [&](int i) {
        int first, count;
        ampc_multi_shard(m, batch, i, &first, &count);
        int r = AMPC_OK;
        if (count > 0) {
            r = ampc_round_batch(m->h[i], count, nullptr, x0 + (size_t)first * 10, ref + (size_t)first * N * 10,
                                 pos_x ? pos_x + first : nullptr, speed, safety_distance, w_inout + (size_t)first * nw,
                                 info_out + first, need_replan_out ? need_replan_out + first : nullptr);
            if (r) set_err(m, r, ampc_last_error(m->h[i]));
        }
        // the exchange must be entered by every device even after an error on one of them
        if (cudaSetDevice(m->device[i]) != cudaSuccess) return set_err(m, AMPC_ERR_CUDA, "cudaSetDevice");
        for (int b = 0; b < per; ++b)
            m->costs_h[i][b] = (!r && b < count) ? info_out[first + b].cost : std::numeric_limits<double>::infinity();
        cudaMemcpyAsync(m->costs_d[i], m->costs_h[i], sizeof(double) * per, cudaMemcpyHostToDevice, m->st[i]);
        ncclResult_t nr = ncclAllGather(m->costs_d[i], m->gathered_d[i], per, ncclDouble, m->comm[i], m->st[i]);
        cudaError_t ce = cudaStreamSynchronize(m->st[i]);
        if (nr != ncclSuccess) return set_err(m, AMPC_ERR_CUDA, std::string("ncclAllGather: ") + ncclGetErrorString(nr));
        if (ce != cudaSuccess) return set_err(m, AMPC_ERR_CUDA, std::string("cost exchange: ") + cudaGetErrorString(ce));
        return r;
    });
    if (rc) return rc;
    if (costs_all_out) {
        std::vector<double> g((size_t)per * m->n);
        if (cudaSetDevice(m->device[0]) != cudaSuccess ||
            cudaMemcpy(g.data(), m->gathered_d[0], sizeof(double) * g.size(), cudaMemcpyDeviceToHost) != cudaSuccess)
            return set_err(m, AMPC_ERR_CUDA, "reading the gathered costs");
        for (int i = 0; i < m->n; ++i) {
            int first, count;
            ampc_multi_shard(m, batch, i, &first, &count);
            if (count > 0) std::memcpy(costs_all_out + first, g.data() + (size_t)i * per, sizeof(double) * count);
        }
    }
    return AMPC_OK;
}

const double *ampc_multi_costs_dev(ampc_multi *m, int32_t i) { return (m && i >= 0 && i < m->n) ? m->gathered_d[i] : nullptr; }

} // extern "C"
