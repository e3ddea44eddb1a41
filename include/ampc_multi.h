/* ampc_multi.h -- one process, several B200s: the scene-sharded form of ampc.h's batched round.
 *
 * The reference plans for one vehicle on one CPU; at batch scale the path shards by scene
 * (SURVEY.md 8e): every device owns a contiguous block of scenes, builds their indices, runs
 * their k-NN and solves with no data-path exchange, and the per-instance costs are all-gathered
 * (NCCL over NVLink) so that every device -- and the host -- ends the round with the costs of
 * all instances (what a best-of / arbitration step over the whole batch reads).
 *
 * libampc_multi.so = this driver (C++ host code: one worker thread per device, one NCCL
 * communicator per device, ncclAllGather of the costs) over libampc.so's C-ABI.  It replaces the
 * same reference call sites as ampc_round_batch (src/AvoidanceStateMachine.cpp:204-257,337) and
 * ampc_cloud_set_batch (src/FrameKDMap.cpp:44-47).  bench.py's multi-GPU legs use one process per
 * GPU over torch.distributed instead (the driver's launch contract); this is the in-process form
 * for a C++ host.
 */
#ifndef AMPC_MULTI_H
#define AMPC_MULTI_H

#include "ampc.h"

#ifdef __cplusplus
extern "C" {
#endif

typedef struct ampc_multi ampc_multi;

/* cfg describes the WHOLE job: max_batch / max_scenes are split evenly (rounded up) over the
 * n_devices devices listed in `devices` (cfg->device is ignored). */
int ampc_multi_create(const ampc_config *cfg, const int32_t *devices, int32_t n_devices, ampc_multi **out);
void ampc_multi_destroy(ampc_multi *m);
const char *ampc_multi_last_error(const ampc_multi *m);
int32_t ampc_multi_device_count(const ampc_multi *m);
/* the per-device handle (parameters, solver options, camera, depth ingress ... are set per handle) */
ampc_handle *ampc_multi_handle(ampc_multi *m, int32_t i);
/* block of a batch of `batch` instances / scenes owned by device i: [first, first + count) */
void ampc_multi_shard(const ampc_multi *m, int32_t batch, int32_t i, int32_t *first, int32_t *count);

/* n_scenes clouds (as ampc_cloud_set_batch): scene s goes to the device owning s in a batch of
 * n_scenes, at local slot s - first */
int ampc_multi_cloud_set_batch(ampc_multi *m, int32_t kind, int32_t n_scenes, const void *xyz_host,
                               const int32_t *counts, int64_t scene_stride_bytes, int32_t stride_bytes);
int ampc_multi_cloud_set_layout(ampc_multi *m, int32_t kind, int32_t row_width);

/* One round for `batch` instances, instance b on scene b (arguments as ampc_round_batch, host
 * buffers).  costs_all_out (may be NULL): batch doubles, the all-gathered costs as device 0 holds
 * them after the exchange. */
int ampc_multi_round_batch(ampc_multi *m, int32_t batch, const double *x0, const double *ref,
                           const double *pos_x, double speed, double safety_distance, double *w_inout,
                           ampc_solve_info *info_out, int32_t *need_replan_out, double *costs_all_out);
/* device pointer (on device i) to the gathered costs of the last round: n_devices blocks of
 * ampc_multi_block(batch) doubles, block j = the costs of device j's instances, padded with +inf */
const double *ampc_multi_costs_dev(ampc_multi *m, int32_t i);
int32_t ampc_multi_block(const ampc_multi *m, int32_t batch);

#ifdef __cplusplus
}
#endif
#endif
