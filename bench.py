#!/usr/bin/env python
"""Headline benchmark: MPC solves/sec at N=20, K=16 obstacle terms/stage, 50k-point clouds.

One "solve" = one round of the reference's control tick for one MPC instance
(src/AvoidanceStateMachine.cpp:328-344): the per-frame index build over the instance's 50k-point
Obstacle cloud (the reference rebuilds its KD-trees every depth frame, src/FrameKDMap.cpp:34-52),
Q = N k-NN queries of K neighbours, prefix packing and one NLP solve to convergence (tol 1e-8).

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference] [--mode ...]

modes (BASELINE.json configs):
  solves       configs[1]: batches of 1024 instances, 1xB200 (N GPUs: scene-sharded)      [default]
  best_of      configs[2]: 4096 scenes x 32 Edge-tree initial guesses, best-cost reduction
  scenes65536  configs[3]: 65 536 scenes sharded over the GPUs, all-gather of the costs
  knn_sweep    configs[4]: k-NN only, 10k..1M points, achieved HBM GB/s, at N GPUs
  cpu_c0       configs[0]: the reference's own CPU-runnable case, one thread
Prints ONE JSON line (rank 0).  See DESIGN.md "Measurement" for what each key means.
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "mpc_solves_per_sec"
UNIT = "solves/s"
N_H, K_NB, DT = 20, 16, 0.05
F_ITER = (N_H - 1) * K_NB * 250 + N_H * 4275  # SURVEY.md 8d work model, FP64 flop per iteration


def workload_name(B, npts):
    return f"batch={B} MPC instances/GPU, N={N_H}, K={K_NB} obstacle terms/stage, {npts}-pt cloud per instance"


def shared_config(args, world=1):
    """The part of `config` both arms print, key for key."""
    return {"workload": workload_name(args.batch, args.npts), "warm_start": args.warm, "tol": args.tol,
            "max_iter": args.max_iter, "n_gpus": world}


def hbm_peak_gbs():
    try:
        return float(json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"]), "MEASURED_PEAKS.json (measured)"
    except Exception:
        return 6650.0, "fallback 6650 GB/s"


# ------------------------------------------------------------------ CPU arm ----
def cpu_reference(n_scenes: int, npts: int, threads: int, steps: int = 1, warm="ref", K=K_NB, tol=1e-8, max_iter=100):
    """The reference's CPU path on the host cores, one instance per std::thread at a time
    (oracle/cpu_arm.cpp): KDTreeTwo::InitializeNew + N x SearchForNearest through the reference's own
    kd_tree_two.h / nanoflann (compiled into oracle/_ref) + GetRefStates packing + the oracle NLP solve.
    Returns (solves/s per step, kind, description, stage seconds of the last step)."""
    import avoid_mpc_b200 as A
    from oracle import oracle as O

    D, S = A.defaults, A.synth
    lb, ub = D.u_bounds()
    opts = O.default_opts(tol=tol, max_iter=max_iter)
    n_distinct = min(n_scenes, 48)
    clouds = np.stack([S.forest_cloud(10_000 + s, npts)[0] for s in range(n_distinct)])
    st = [S.states(10_000 + s, N_H) for s in range(n_distinct)]
    x0, ref, tgt = (np.stack([a[i] for a in st]) for i in range(3))
    W0 = np.stack([S.warm_start(warm, x0[i], ref[i], N_H) for i in range(n_distinct)])
    tail = S.full_params(np.zeros(0))
    rates, stage = [], None
    if O.cpu_arm_available():
        O.cpu_arm_run(min(n_scenes, 2 * threads), threads, clouds, x0, ref, tgt, tail, W0, N_H, K, DT, lb, ub, opts)
        for _ in range(steps):
            wall, _, _, stage = O.cpu_arm_run(n_scenes, threads, clouds, x0, ref, tgt, tail, W0, N_H, K, DT, lb, ub, opts)
            rates.append(n_scenes / wall)
        how = "native std::thread driver (oracle/cpu_arm.cpp), reference kd_tree_two.h + nanoflann for the k-NN"
    else:  # oracle/_ref was not built (no reference tree at build time): Python-driven port
        from concurrent.futures import ThreadPoolExecutor
        O.lib()

        def one(i):
            j = i % n_distinct
            tree = O.PortTree(clouds[j])
            idx, d2, cnt = tree.search(ref[j][:, :3], K)
            ob = clouds[j][idx.reshape(-1), :3].astype(np.float64).reshape(N_H, K, 3)
            p = S.full_params(S.pack_prefix(x0[j], ref[j], ob, tgt[j]))
            return O.solve(N_H, K, DT, p, W0[j], lb, ub, opts)[1].status

        with ThreadPoolExecutor(max_workers=threads) as ex:
            list(ex.map(one, range(min(n_scenes, 2 * threads))))
            for _ in range(steps):
                t0 = time.perf_counter()
                list(ex.map(one, range(n_scenes)))
                rates.append(n_scenes / (time.perf_counter() - t0))
        how = "Python thread pool over the C restatement (oracle/_ref missing)"
    desc = (f"{n_scenes} instances/step ({n_distinct} distinct {npts}-pt scenes), {threads} host threads, {how}; "
            f"per instance: tree build + {N_H}x{K}-NN + packing + oracle interior-point solve (CasADi/IPOPT not "
            f"installable), tol {tol:g}")
    return rates, "port", desc, (None if stage is None else [float(v) for v in stage])


def run_reference_arm(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    threads = os.cpu_count() or 1
    n = max(threads * 32, 256)
    rates, kind, desc, stage = cpu_reference(n, args.npts, threads, steps=args.warmup + args.steps, warm=args.warm,
                                             tol=args.tol, max_iter=args.max_iter)
    rates = rates[args.warmup:]
    v = statistics.mean(rates)
    out = {
        "impl": "reference", "metric": METRIC, "value": v, "unit": UNIT, "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * n / v, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": shared_config(args, args.gpus),
        "run": {"sample_instances_per_step": n, "stage_seconds_summed_over_threads": stage},
        "cpu_baseline": {"value": v, "unit": UNIT, "cores": threads, "kind": kind, "sample": desc},
        "e2e": {"value": v, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(out), flush=True)


# ------------------------------------------------------------------ GPU arm ----
class ClockSampler:
    """nvidia-smi polled every 50 ms by a reader thread that stamps each sample on arrival; the
    samples that arrived inside [mark_begin(), mark_end()] are reported."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
         "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, dev):
        self.dev, self.p, self.rows, self.t0, self.t1, self.thread = dev, None, [], None, None, None

    def start(self):
        try:
            self.p = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                       "-lms", "50", "-i", str(self.dev)], stdout=subprocess.PIPE, text=True)
        except OSError:
            self.p = None
            return

        def reader():
            for line in self.p.stdout:
                self.rows.append((time.perf_counter(), line))
        self.thread = threading.Thread(target=reader, daemon=True)
        self.thread.start()

    def mark_begin(self):
        self.t0 = time.perf_counter()

    def mark_end(self):
        self.t1 = time.perf_counter()

    def stop(self):
        if not self.p:
            return {"sm_mhz": None, "sm_max_mhz": None, "samples": 0, "reasons": ["nvidia-smi unavailable"]}
        self.p.terminate()
        try:
            self.p.wait(timeout=5)
        except Exception:
            self.p.kill()
        if self.thread:
            self.thread.join(timeout=2)
        sm, mx, pw, reasons = [], [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ts, line in list(self.rows):
            if self.t0 is not None and (ts < self.t0 or ts > (self.t1 or ts) + 0.05):
                continue
            f = [x.strip() for x in line.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1])), mx.append(float(f[2])), pw.append(float(f[3]))
            except ValueError:
                continue
            for nm, val in zip(names, f[5:9]):
                if val.lower().startswith("active"):
                    reasons.add(nm)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "power_w_median": statistics.median(pw) if pw else None,
                "samples": len(sm), "reasons": sorted(reasons), "window": "the timed region of `value`"}


HOST_BINDING = {}


def bind_to_gpu_numa_node(torch, local):
    """Run this process (and so first-touch its pinned host buffers) on the CPUs of the NUMA node the
    GPU hangs off, as a deployment would with numactl: with one process per GPU on a two-socket
    host, host->device copies otherwise cross the socket link.  Best effort; recorded in the line."""
    try:
        pr = torch.cuda.get_device_properties(local)
        bdf = "%04x:%02x:%02x.0" % (pr.pci_domain_id, pr.pci_bus_id, pr.pci_device_id)
        node = int(open("/sys/bus/pci/devices/%s/numa_node" % bdf).read())
        if node < 0:
            HOST_BINDING.update(numa_node=None, note="the GPU reports no NUMA node")
            return
        cpus = set()
        for part in open("/sys/devices/system/node/node%d/cpulist" % node).read().strip().split(","):
            a, _, b = part.partition("-")
            cpus.update(range(int(a), int(b or a) + 1))
        cpus &= os.sched_getaffinity(0)
        if cpus:
            os.sched_setaffinity(0, cpus)
            HOST_BINDING.update(numa_node=node, cpus=len(cpus), pci=bdf)
    except Exception as e:  # no sysfs / no permission: run unbound
        HOST_BINDING.update(numa_node=None, note="unbound (%s)" % type(e).__name__)


def dist_setup():
    import torch
    import torch.distributed as dist
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the product path has no CPU fallback")
    torch.cuda.set_device(local)
    if os.environ.get("AMPC_BENCH_NUMA_BIND", "1") != "0":
        bind_to_gpu_numa_node(torch, local)
    dev = torch.device("cuda", local)
    if world > 1:
        # NCCL prints its version banner on stdout when the communicator is created: stdout carries
        # ONE JSON line, so file descriptor 1 points at stderr until the first collective is through
        sys.stdout.flush()
        saved = os.dup(1)
        os.dup2(2, 1)
        try:
            dist.init_process_group("nccl", device_id=dev)
            dist.barrier()
            torch.cuda.synchronize()
        finally:
            sys.stdout.flush()
            os.dup2(saved, 1)
            os.close(saved)
    return torch, dist, world, rank, local, dev


def fill_scenes(A, torch, h, ids, npts, dev, chunk=1024):
    """Synthetic forest clouds of the scenes `ids` straight into the handle's slots."""
    S = A.synth
    st = torch.cuda.current_stream().cuda_stream
    for s0 in range(0, len(ids), chunk):
        c = S.forest_clouds_torch(ids[s0:s0 + chunk], npts, dev)
        h.cloud_set_batch_dev(c, first_scene=s0, stream=st)
        torch.cuda.synchronize()
        del c


def solver_stats(A, info_t):
    info = info_t.cpu().numpy().view(A.capi.INFO_DTYPE).reshape(-1)
    it = info["iters"].astype(np.float64)
    return info, {"converged_frac": float((info["status"] == 0).mean()),
                  "status_counts": np.bincount(info["status"], minlength=4).tolist(),
                  "iters_mean": float(it.mean()), "iters_p50": float(np.median(it)),
                  "iters_p90": float(np.percentile(it, 90)), "iters_max": int(it.max())}


def ncu_traffic(kernel_key):
    """dram bytes (read + write) of one launch from the committed ncu --set full summary, or None."""
    for name in ("ncu_summary_r02.json", "ncu_summary_r01.json"):
        try:
            m = json.load(open(os.path.join(ROOT, "profiles", name)))[kernel_key]["metrics"]
            scale = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
            return sum(float(m[k]["value"]) * scale[m[k]["unit"]] for k in ("dram__bytes_read.sum", "dram__bytes_write.sum")), name
        except Exception:
            continue
    return None, None


def run_ours(args):
    torch, dist, world, rank, local, dev = dist_setup()
    import avoid_mpc_b200 as A
    D, S = A.defaults, A.synth
    B, npts, F, L = args.batch, args.npts, max(1, args.in_flight), max(1, args.streams)
    F = (F + L - 1) // L * L
    BL = B * F // L  # instances per lane (one call)
    BF = B * F       # instances per step per GPU: F batches of the BASELINE size, each with its own scenes

    # ---- synthetic inputs (weak scaling: BF scenes per GPU, distinct per rank and per batch) ----
    # L lanes = independent calls in flight (own handle, scenes, outputs, CUDA stream): the tail of
    # one lane's solve kernel (its slowest instances) overlaps the next lane's kernels.
    lanes = []
    for li in range(L):
        ids = list(range(rank * BF + li * BL, rank * BF + (li + 1) * BL))
        h = A.Handle(N=N_H, K=K_NB, dt=DT, max_batch=BL, max_points=npts, device=local)
        h.set_solver_opts(tol=args.tol, max_iter=args.max_iter)
        if not args.unorganised:  # the clouds are row-major depth images: tell the index the row pitch
            h.cloud_set_layout(S.image_shape(npts)[0])
        fill_scenes(A, torch, h, ids, npts, dev)
        x0_np, ref_np, _ = S.states_batch(ids, N_H, DT)
        w0_np = np.stack([S.warm_start(args.warm, x0_np[b], ref_np[b], N_H) for b in range(BL)])
        ln = dict(h=h, st=torch.cuda.Stream(device=dev), x0_np=x0_np, ref_np=ref_np)
        ln["x0"], ln["ref"], ln["w0"] = (torch.tensor(a, device=dev) for a in (x0_np, ref_np, w0_np))
        ln["w"] = torch.empty_like(ln["w0"])
        ln["info"] = torch.zeros((BL, 48), dtype=torch.uint8, device=dev)
        ln["info_f64"] = ln["info"].view(torch.float64).view(BL, 6)
        ln["replan"] = torch.zeros(BL, dtype=torch.int32, device=dev)
        # cost exchange, double buffered: the all-gather of step s runs on its own stream while the
        # lane already computes step s+1 (it only has to be over before step s+2 reuses the buffer)
        ln["costs"] = [torch.zeros(BL, dtype=torch.float64, device=dev) for _ in range(2)]
        ln["gathered"] = [torch.zeros(BL * world, dtype=torch.float64, device=dev) for _ in range(2)] if world > 1 else None
        ln["gather_done"] = [None, None]
        ln["k"] = 0
        lanes.append(ln)
    comm = torch.cuda.Stream(device=dev) if world > 1 else None
    h = lanes[0]["h"]
    l0_ = lanes[0]
    x0, ref, w0, w, info, replan = l0_["x0"], l0_["ref"], l0_["w0"], l0_["w"], l0_["info"], l0_["replan"]
    x0_np, ref_np = l0_["x0_np"], l0_["ref_np"]
    stream = lanes[0]["st"].cuda_stream
    torch.cuda.synchronize()

    def lane_step(ln, n):
        st = ln["st"]
        with torch.cuda.stream(st):
            ln["w"][:n].copy_(ln["w0"][:n], non_blocking=True)
            # a new depth frame per solve: the index build is the one streaming pass over the clouds
            ln["h"].cloud_index_dev(0, n, stream=st.cuda_stream)
            ln["h"].round_dev(n, ln["x0"], ln["ref"], ln["w"], info_dev=ln["info"], replan_dev=ln["replan"], speed=D.SPEED,
                              safety_distance=D.SAFETY_DISTANCE, stream=st.cuda_stream)
            if world > 1 and n == BL and os.environ.get("AMPC_BENCH_GATHER") != "off":  # per-instance cost exchange: the only collective of the path
                k = ln["k"] & 1
                ln["k"] += 1
                if ln["gather_done"][k] is not None:
                    st.wait_event(ln["gather_done"][k])
                ln["costs"][k].copy_(ln["info_f64"][:, 0])
                if os.environ.get("AMPC_BENCH_GATHER", "async") == "inline":  # round 1: on the lane's own stream
                    dist.all_gather_into_tensor(ln["gathered"][k], ln["costs"][k])
                else:
                    ready = torch.cuda.Event()
                    ready.record(st)
                    comm.wait_event(ready)
                    with torch.cuda.stream(comm):
                        dist.all_gather_into_tensor(ln["gathered"][k], ln["costs"][k])
                        ln["gather_done"][k] = torch.cuda.Event()
                        ln["gather_done"][k].record(comm)

    def step(n=None):
        if n is None:
            for ln in lanes:
                lane_step(ln, BL)
        else:
            lane_step(lanes[0], n)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(steps, n=None):
        main = torch.cuda.current_stream()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        barrier()
        e0.record(main)
        for ln in lanes:
            ln["st"].wait_event(e0)
        for _ in range(steps):
            step(n)
        for ln in lanes:
            main.wait_stream(ln["st"])
        if comm is not None:
            main.wait_stream(comm)
        e1.record(main)
        barrier()
        return e0.elapsed_time(e1)

    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    for _ in range(args.warmup):
        step()
    barrier()
    if rank == 0:  # nvidia-smi needs about a second for its first sample: keep the GPU loaded meanwhile
        t_end = time.perf_counter() + 1.2
    n_load = 0
    while True:
        step()
        n_load += 1
        flag = torch.tensor([1.0 if (rank == 0 and time.perf_counter() < t_end) else 0.0], device=dev)
        if world > 1:
            dist.broadcast(flag, 0)
        if flag.item() == 0.0:
            break
    barrier()
    # ---- the timed region: exactly K steps, device time, max over ranks ----
    for ln in lanes:
        ln["h"].profile_enable(True)
    l0 = sum(ln["h"].launch_count() for ln in lanes)
    sampler.mark_begin()
    total_ms = timed(args.steps)
    sampler.mark_end()
    launches = sum(ln["h"].launch_count() for ln in lanes) - l0
    prof = {k: [0.0, 0] for k in ("index", "knn", "solve")}
    for ln in lanes:
        pr = ln["h"].profile_get()
        for k in prof:
            prof[k][0] += pr[k][0]
            prof[k][1] += pr[k][1]
        ln["h"].profile_enable(False)
    t = torch.tensor([total_ms], dtype=torch.float64, device=dev)
    info_np, sstats = solver_stats(A, torch.cat([ln["info"] for ln in lanes]))
    # scaling trace: every rank's own device time and solver work for the timed region (weak scaling:
    # each rank has different scenes, the job lasts as long as the rank with the most work)
    mine = torch.tensor([total_ms, float(info_np["iters"].astype(np.float64).sum())], dtype=torch.float64, device=dev)
    per_rank = [mine.clone() for _ in range(world)]
    if world > 1:
        dist.all_gather(per_rank, mine)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    per_rank = {"device_ms": [float(x[0].item()) for x in per_rank],
                "solver_iterations_per_step": [float(x[1].item()) for x in per_rank]}
    total_ms = float(t.item())
    clocks = sampler.stop() if rank == 0 else None
    value = world * BF * args.steps / (total_ms * 1e-3)
    sstats["need_replan_frac"] = float(torch.cat([ln["replan"] for ln in lanes]).float().mean().item())
    # per-kernel device time per call (with several lanes the kernels of different lanes time-share the GPU)
    index_ms, knn_ms, solve_ms = (prof[k][0] / max(prof[k][1], 1) for k in ("index", "knn", "solve"))

    # ---- one batch of the BASELINE size alone (latency view; the warp-per-instance kernel) ----
    for _ in range(3):
        step(B)
    torch.cuda.synchronize()
    h.profile_enable(True)
    n_single = max(5, args.steps)
    single_ms = timed(n_single, B)
    prof1 = h.profile_get()
    h.profile_enable(False)
    info1_np, s1 = solver_stats(A, info[:B])
    single = {"value": world * B * n_single / (single_ms * 1e-3), "unit": UNIT, "ms_per_batch": single_ms / n_single,
              "stage_ms": {k: prof1[k][0] / max(prof1[k][1], 1) for k in ("index", "knn", "solve")},
              "note": "one 1024-instance batch in flight: the batch waits for its slowest instance "
                      "(iterations p50 %d, max %d)" % (s1["iters_p50"], s1["iters_max"])}

    e2e = e2e_leg(A, torch, dist, dev, local, world, args)

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    # ---- roofline of the dominant kernel: the solve kernel, FP64 FMA pipe ----
    fp64_peak = h.measure_fp64_peak()
    iters_sum = float(info_np["iters"].astype(np.float64).sum())
    # achieved rate of the solve kernel: the iterations of one step over the device time the solve
    # kernels of that step occupy; with one lane that is the kernel's own duration, with several
    # lanes the kernels overlap and the step time is the honest denominator
    solve_busy_ms = solve_ms if L == 1 else (total_ms / args.steps) * solve_ms / (index_ms + knn_ms + solve_ms)
    in_step_tflops = iters_sum * F_ITER / (solve_busy_ms * 1e-3) / 1e12
    # ... and the kernel by itself: the same launches (the packed parameter vectors of the last round,
    # the same warm starts, one launch per lane in flight as in the step) with nothing else running,
    # CUDA events around every launch on its stream and around the whole loop
    def solve_only(rounds):
        main = torch.cuda.current_stream()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        per_launch = []
        torch.cuda.synchronize()
        e0.record(main)
        for ln in lanes:
            ln["st"].wait_event(e0)
        for _ in range(rounds):
            for ln in lanes:
                with torch.cuda.stream(ln["st"]):
                    ln["w"].copy_(ln["w0"], non_blocking=True)
                    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                    a.record(ln["st"])
                    ln["h"].solve_dev(BL, ln["h"].last_prefix_ptr(), ln["w"], info_dev=ln["info"], stream=ln["st"].cuda_stream)
                    b.record(ln["st"])
                    per_launch.append((a, b))
        for ln in lanes:
            main.wait_stream(ln["st"])
        e1.record(main)
        torch.cuda.synchronize()
        return e0.elapsed_time(e1), [a.elapsed_time(b) for a, b in per_launch]
    solve_only(1)
    so_rounds = 6
    so_ms, so_launch = solve_only(so_rounds)
    info_so, _ = solver_stats(A, torch.cat([ln["info"] for ln in lanes]))
    so_iters = float(info_so["iters"].astype(np.float64).sum())  # (== iters_sum: same problems, same starts)
    solve_tflops = so_rounds * so_iters * F_ITER / (so_ms * 1e-3) / 1e12
    so_launch_ms = sum(so_launch) / len(so_launch)
    step_ms = index_ms + knn_ms + solve_ms
    kern = "ipm_quad_kernel" if BL >= 8192 else "ipm_solve_kernel"
    traffic, traffic_src = ncu_traffic("ipm_quad" if BL >= 8192 else "ipm_solve")
    if traffic is not None and BL >= 8192:
        traffic *= BL / 32768.0  # the committed capture is of a 32768-instance call
    hbm_peak, hbm_src = hbm_peak_gbs()
    b_knn = 12 * npts + N_H * (24 + K_NB * 12)
    idx_traffic, _ = ncu_traffic("cloud_index")
    roofline = {
        "kernel": kern, "bound": "fp64", "bound_note": "FP64 FMA pipe (neither hbm nor tensor: ~12 KB of HBM traffic per "
        "instance, and tcgen05 has no FP64 path; the condensed DMMA variant was measured and rejected, DESIGN.md 4.3)",
        "achieved": solve_tflops, "peak": fp64_peak, "unit": "TFLOP/s", "frac": solve_tflops / fp64_peak,
        "traffic": traffic, "traffic_source": traffic_src,
        "peak_source": "hand-written dependent-FMA-chain kernel on this GPU, this run (ampc_measure_fp64_peak)",
        "work_model": "sum of interior-point iterations x %d FP64 flop (SURVEY.md 8d: (N-1) K 250 + N 4275)" % F_ITER,
        "measured": "a loop of this kernel alone after the timed region (%d launches of %d instances on the step's inputs, %d in "
                    "flight on %d streams exactly as in the step): achieved = flops of all launches / CUDA-event time of the loop"
                    % (so_rounds * L, BL, L, L),
        "avg_launch_ms": so_launch_ms, "launches_in_flight": L, "iterations_per_launch": so_iters / L,
        "per_launch": {"flop": so_iters / L * F_ITER, "tflops": so_iters / L * F_ITER / (so_launch_ms * 1e-3) / 1e12,
                       "frac": so_iters / L * F_ITER / (so_launch_ms * 1e-3) / 1e12 / fp64_peak,
                       "note": "flops of ONE launch / its own average duration: a lower bound, the launch shares the GPU "
                               "with the other lane's launch for most of that time"},
        "in_step": {"iterations_per_step": iters_sum, "avg_launch_ms": solve_ms, "tflops": in_step_tflops,
                    "frac": in_step_tflops / fp64_peak, "solve_ms_per_step_share_of_timed_region": solve_busy_ms,
                    "note": "inside the timed region the solve launches overlap the other lane's index / search kernels: "
                            "step time x this kernel's share of the summed stage times as the denominator"},
        "step_share": solve_ms / step_ms,
        # the HBM-bound part, from the single-batch pass (its kernels run alone there; with several
        # lanes in flight a kernel's event time includes time-sharing with the other lane's kernels)
        "cloud_index": {"kernel": "cloud_index_kernel (+ compaction check + group boxes)", "bound": "hbm",
                        "achieved": B * b_knn / (single["stage_ms"]["index"] * 1e-3) / 1e9,
                        "peak": hbm_peak, "unit": "GB/s",
                        "frac": B * b_knn / (single["stage_ms"]["index"] * 1e-3) / 1e9 / hbm_peak,
                        "traffic": None if idx_traffic is None else idx_traffic * B / 32768.0,  # (captured on a 32768-scene launch)
                        "peak_source": hbm_src, "algorithmic_bytes_per_launch": B * b_knn,
                        "as_laid_out_16B_GBps": B * 16 * npts / (single["stage_ms"]["index"] * 1e-3) / 1e9,
                        "avg_launch_ms": single["stage_ms"]["index"], "instances_per_launch": B,
                        "step_share": index_ms / step_ms},
        "knn_stage": {"what": "index build + two-level box-pruned search (k-NN indices bit-exact), single batch alone",
                      "index_ms": single["stage_ms"]["index"], "search_ms": single["stage_ms"]["knn"],
                      "achieved_GBps": B * b_knn / ((single["stage_ms"]["index"] + single["stage_ms"]["knn"]) * 1e-3) / 1e9,
                      "frac": B * b_knn / ((single["stage_ms"]["index"] + single["stage_ms"]["knn"]) * 1e-3) / 1e9 / hbm_peak,
                      "step_share": (index_ms + knn_ms) / step_ms}}

    out = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": total_ms / args.steps, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": shared_config(args, world),
        "run": {"batches_per_step": F, "lanes": L, "instances_per_call": BL, "instances_per_step_per_gpu": BF,
                "global_instances_per_step": world * BF,
                "step": "one control round (index build + k-NN + solve) over %d batches of %d instances, each batch with "
                        "its own %d-point scenes, submitted as %d call(s) on %d CUDA stream(s); the solve kernel refills "
                        "its warps from a queue over all instances of a call" % (F, B, npts, L, L),
                "timed_region_s": total_ms * 1e-3, "load_steps_before": n_load, "per_rank": per_rank,
                "parallelism": f"scene-sharded x{world}",
                "l2": "inputs larger than L2 (%.1f GB of clouds per step per GPU)" % (BF * npts * 16 / 1e9),
                "collective": ("all_gather of per-instance costs (NCCL), every step, %s" % (
                    "on the lane's stream" if os.environ.get("AMPC_BENCH_GATHER", "async") == "inline"
                    else "SKIPPED (scaling-trace run, not a bench line)" if os.environ.get("AMPC_BENCH_GATHER") == "off"
                    else "on its own stream one step behind the lanes")) if world > 1 else "none"},
        "e2e": e2e,
        "gpu_launches": int(launches),
        "clocks": clocks,
        "roofline": roofline,
        "stage_ms_per_step": {"index": index_ms, "knn_search": knn_ms, "solve": solve_ms},
        "single_stream": single,
        "solver": sstats,
    }
    if world == 1:
        with torch.cuda.stream(lanes[0]["st"]):
            out["cold_start"] = cold_start_leg(A, torch, h, B, x0, ref, x0_np, ref_np, w, info, replan, stream)
            out["reference_operating_point"] = truncated_leg(A, torch, h, B, x0, ref, w0, w, info, replan, stream, args)
        out["depth_path"] = depth_resident_leg(A, torch, dev, local, args)
    if world == 1 and not args.no_cpu_baseline:
        threads = os.cpu_count() or 1
        n = max(threads * 16, 128)
        rates, kind, desc, stage = cpu_reference(n, npts, threads, steps=2, warm=args.warm, tol=args.tol, max_iter=args.max_iter)
        out["cpu_baseline"] = {"value": rates[-1], "unit": UNIT, "cores": threads, "kind": kind, "sample": desc,
                               "stage_seconds_summed_over_threads": stage}
        out["cpu_c0"] = cpu_c0(args, n_inst=60)
    else:
        out["cpu_baseline"] = None
    print(json.dumps(out), flush=True)
    if world > 1:
        dist.destroy_process_group()


def e2e_leg(A, torch, dist, dev, local, world, args, rows=200, cols=250, distinct=128, lanes_n=6, batches_per_call=2):
    """`e2e`: the same round through the host-buffer C-ABI, starting from what the reference's ingress
    receives -- the depth frame (FrameKDMap::AddVertex, src/FrameKDMap.cpp:34-52): every step copies its
    depth frames (f32 metres, 4 bytes per point) and states from pinned host memory, builds both clouds
    and their indices on the device, runs k-NN + solve, and reads trajectories / costs / status back.
    `lanes_n` host threads each drive their own handle (calls of `batches_per_call` x --batch
    instances), so one call's upload overlaps the compute of the others; measured on one B200:
    3 / 6 / 8 lanes of 1024 -> 156 k / 186 k / 204 k solves/s, 6 lanes of 2048 -> 202 k at 41 GB/s
    host->device, which is what a pinned copy alone reaches on that box (42-50 GB/s): PCIe-bound.
    `depth_u16` is the same from CV_16UC1 frames (half the bytes), `cloud_upload` the same with
    ready-made 16-byte clouds uploaded instead."""
    from concurrent.futures import ThreadPoolExecutor
    D, S = A.defaults, A.synth
    B = args.batch * batches_per_call
    rank = int(os.environ.get("RANK", "0"))
    cam = dict(fx=cols / 2, fy=cols / 2, cx=cols / 2, cy=rows / 2, resize_scale=1.0)
    ids = [rank * B + b for b in range(B)]
    base = np.stack([S.forest_depth(rank * distinct + s, rows, cols, sky=False) for s in range(distinct)])
    depth_h = torch.from_numpy(base).repeat((B + distinct - 1) // distinct, 1, 1)[:B].contiguous().pin_memory()
    Twb = np.eye(4)
    Twb[2, 3] = D.HEIGHT
    T_np = np.ascontiguousarray(np.tile((Twb @ D.T_B_C).reshape(1, 16), (B, 1)))
    x0_np, ref_np, _ = S.states_batch(ids, N_H, DT)
    w0_np = np.stack([S.warm_start(args.warm, x0_np[b], ref_np[b], N_H) for b in range(B)])
    x0_h, ref_h, w0_h = (torch.tensor(a).pin_memory() for a in (x0_np, ref_np, w0_np))
    depth_np = depth_h.numpy()

    def make(edge):
        hh = A.Handle(N=N_H, K=K_NB, dt=DT, max_batch=B, max_points=rows * cols,
                      max_edge_points=rows * cols // 4 if edge else 0, device=local)
        hh.set_solver_opts(tol=args.tol, max_iter=args.max_iter)
        hh.set_camera(**cam)
        return dict(h=hh, w_h=torch.empty_like(w0_h).pin_memory(), info_h=torch.zeros((B, 48), dtype=torch.uint8).pin_memory(),
                    replan_h=torch.zeros(B, dtype=torch.int32).pin_memory())

    def run(step_fn, lanes, n):
        with ThreadPoolExecutor(max_workers=len(lanes)) as ex:
            list(ex.map(step_fn, [lanes[i % len(lanes)] for i in range(2 * len(lanes))]))
            torch.cuda.synchronize()
            if world > 1:
                dist.barrier()
            t0 = time.perf_counter()
            list(ex.map(step_fn, [lanes[i % len(lanes)] for i in range(n)]))
            torch.cuda.synchronize()
            if world > 1:
                dist.barrier()
            dt = time.perf_counter() - t0
        t = torch.tensor([dt], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    lanes = [make(True) for _ in range(lanes_n)]

    def depth_step(ln):
        ln["w_h"].copy_(w0_h)
        ln["h"].depth_set_batch(depth_np, T_np)                       # H2D depth frames; clouds + indices on the device
        ln["h"].round_host_ptrs(B, x0_h, ref_h, ln["w_h"], ln["info_h"], ln["replan_h"], speed=D.SPEED,
                                safety_distance=D.SAFETY_DISTANCE)    # H2D states, k-NN, solve, D2H results

    n = max(4 * lanes_n, min(args.steps, 24))
    dt = run(depth_step, lanes, n)
    info = lanes[0]["info_h"].numpy().view(A.capi.INFO_DTYPE).reshape(B)
    h2d = depth_h.numel() * 4 + T_np.nbytes + (x0_h.numel() + ref_h.numel() + w0_h.numel()) * 8
    d2h = w0_h.numel() * 8 + B * 48 + B * 4
    out = {"value": world * B * n / dt, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
           "steps": n, "host_threads": lanes_n, "instances_per_step": B,
           "converged_frac": float((info["status"] == 0).mean()),
           "ingress": "host depth frames %dx%d f32 (what FrameKDMap::AddVertex receives); Obstacle + Edge clouds and "
                      "their indices are built on the device" % (rows, cols),
           "h2d_GBps": h2d * n / dt / 1e9, "host_binding": dict(HOST_BINDING)}
    # the same from CV_16UC1 frames in millimetres (what a depth camera publishes; the reference's
    # GetInvDepthImg<uint16_t> branch, src/FrameKDMap.cpp:96-97, pixel2meter 0.001): half the bytes
    depth16_h = torch.from_numpy(np.clip(np.rint(depth_h.numpy() * 1000.0), 0, 65535).astype(np.uint16).view(np.int16)).pin_memory()
    depth16_np = depth16_h.numpy().view(np.uint16)
    for ln in lanes:
        ln["h"].set_camera(pixel2meter=0.001, **cam)

    def depth16_step(ln):
        ln["w_h"].copy_(w0_h)
        ln["h"].depth_set_batch(depth16_np, T_np)
        ln["h"].round_host_ptrs(B, x0_h, ref_h, ln["w_h"], ln["info_h"], ln["replan_h"], speed=D.SPEED,
                                safety_distance=D.SAFETY_DISTANCE)

    dt16 = run(depth16_step, lanes, n)
    h2d16 = h2d - depth_h.numel() * 2
    out["depth_u16"] = {"value": world * B * n / dt16, "unit": UNIT, "h2d_bytes_per_step": h2d16, "d2h_bytes_per_step": d2h,
                        "steps": n, "h2d_GBps": h2d16 * n / dt16 / 1e9,
                        "note": "CV_16UC1 millimetre frames (a depth camera's format, pixel2meter 0.001) instead of the "
                                "simulator's CV_32FC1 metres: half the host->device bytes"}
    for ln in lanes:
        ln["h"].set_camera(**cam)
    # the same with ready-made clouds uploaded (16 bytes per point): the number round 1 called e2e
    clouds_h = torch.empty((B, rows * cols, 4), dtype=torch.float32).pin_memory()
    clouds_h.copy_(S.forest_clouds_torch(ids, rows * cols, dev))
    for ln in lanes:
        ln["h"].cloud_set_layout(S.image_shape(rows * cols)[0])

    def cloud_step(ln):
        ln["w_h"].copy_(w0_h)
        ln["h"].cloud_set_batch(clouds_h)
        ln["h"].round_host_ptrs(B, x0_h, ref_h, ln["w_h"], ln["info_h"], ln["replan_h"], speed=D.SPEED,
                                safety_distance=D.SAFETY_DISTANCE)

    n2 = 2 * lanes_n
    dt2 = run(cloud_step, lanes, n2)
    h2d2 = clouds_h.numel() * 4 + (x0_h.numel() + ref_h.numel() + w0_h.numel()) * 8
    out["cloud_upload"] = {"value": world * B * n2 / dt2, "unit": UNIT, "h2d_bytes_per_step": h2d2,
                           "d2h_bytes_per_step": d2h, "steps": n2, "h2d_GBps": h2d2 * n2 / dt2 / 1e9,
                           "note": "ready-made 16-byte clouds uploaded every step instead of depth frames (PCIe-bound)"}
    # what the host side alone can move: pinned H2D copy bandwidth of this process
    buf = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    src = torch.empty(256 << 20, dtype=torch.uint8).pin_memory()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(4):
        buf.copy_(src, non_blocking=True)
    torch.cuda.synchronize()
    out["pinned_h2d_GBps_alone"] = 4 * (256 << 20) / (time.perf_counter() - t0) / 1e9
    for ln in lanes:
        ln["h"].close()
    return out


def cold_start_leg(A, torch, h, B, x0, ref, x0_np, ref_np, w, info, replan, stream, steps=5):
    """The reference's cold start (all-zero mNlpW0, src/HighLvlMpc.cpp:25-27) on one batch."""
    D, S = A.defaults, A.synth
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    for rep in range(2):
        if rep == 1:
            e0.record()  # (the caller made the lane's stream current)
        for _ in range(steps):
            w[:B].zero_()
            h.cloud_index_dev(0, B, stream=stream)
            h.round_dev(B, x0, ref, w, info_dev=info, replan_dev=replan, speed=D.SPEED, safety_distance=D.SAFETY_DISTANCE,
                        stream=stream)
    e1.record()
    torch.cuda.synchronize()
    _, st = solver_stats(A, info[:B])
    return {"warm_start": "cold", "value": B * steps / (e0.elapsed_time(e1) * 1e-3), "unit": UNIT,
            "ms_per_batch": e0.elapsed_time(e1) / steps, "solver": st, "note": "one 1024-instance batch in flight"}


def truncated_leg(A, torch, h, B, x0, ref, w0, w, info, replan, stream, args):
    """What a drop-in user sees with the reference's solver options (tol 1e-4, 10 iterations,
    src/HighLvlMpc.cpp:19-20): distance of that iterate from this solver's converged optimum."""
    D = A.defaults

    def run(tol, mi):
        h.set_solver_opts(tol=tol, max_iter=mi)
        w[:B].copy_(w0[:B])
        h.round_dev(B, x0, ref, w, info_dev=info, replan_dev=replan, speed=D.SPEED, safety_distance=D.SAFETY_DISTANCE,
                    stream=stream)
        torch.cuda.synchronize()
        return w[:B].cpu().numpy().copy(), info[:B].cpu().numpy().view(A.capi.INFO_DTYPE).reshape(B).copy()

    Wc, ic = run(args.tol, 200)
    Wt, it = run(1e-4, 10)
    h.set_solver_opts(tol=args.tol, max_iter=args.max_iter)
    ok = ic["status"] == 0
    gap = np.abs(Wt - Wc).max(axis=1)[ok]
    u0 = np.abs(Wt[:, 10:14] - Wc[:, 10:14]).max(axis=1)[ok]
    return {"options": {"tol": 1e-4, "max_iter": 10}, "instances": int(ok.sum()),
            "stopped_by_tol_frac": float((it["status"][ok] == 0).mean()),
            "linf_gap_to_converged": {"p50": float(np.median(gap)), "p90": float(np.percentile(gap, 90)), "max": float(gap.max()),
                                      "frac_below_1e-4": float((gap < 1e-4).mean()), "frac_below_1e-2": float((gap < 1e-2).mean())},
            "first_control_gap": {"p50": float(np.median(u0)), "p90": float(np.percentile(u0, 90)), "max": float(u0.max())},
            "note": "l_inf over states and controls between the tol-1e-4 / 10-iteration iterate and the KKT<=1e-8 optimum "
                    "of the same solver from the same warm start; u = w[10:14] is what the reference publishes"}


def depth_resident_leg(A, torch, dev, local, args, rows=200, cols=250, distinct=128):
    """The step upstream of the clouds: depth frames resident in HBM -> Obstacle +
    Edge clouds + tile indices (ampc_depth_set_batch_dev), the numpy restatement timed beside it."""
    D, S = A.defaults, A.synth
    B = args.batch
    hbm_peak, _ = hbm_peak_gbs()
    cam = dict(fx=cols / 2, fy=cols / 2, cx=cols / 2, cy=rows / 2, resize_scale=1.0)
    base = np.stack([S.forest_depth(s, rows, cols, sky=False) for s in range(distinct)])
    depth = torch.from_numpy(base).repeat((B + distinct - 1) // distinct, 1, 1)[:B].contiguous().to(dev)
    Twb = np.eye(4)
    Twb[2, 3] = D.HEIGHT
    T_np = np.ascontiguousarray(np.tile((Twb @ D.T_B_C).reshape(1, 16), (B, 1)))
    T = torch.from_numpy(T_np).to(dev)
    hd = A.Handle(N=N_H, K=K_NB, dt=DT, max_batch=B, max_points=rows * cols, max_edge_points=rows * cols // 4, device=local)
    hd.set_camera(**cam)
    st = torch.cuda.current_stream().cuda_stream
    for _ in range(3):
        hd.depth_set_batch_dev(depth, T, None, stream=st)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    l0, steps = hd.launch_count(), 5
    e0.record()
    for _ in range(steps):
        hd.depth_set_batch_dev(depth, T, None, stream=st)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / steps
    n_obst = float(np.mean([hd.cloud_count(s) for s in range(8)]))
    n_edge = float(np.mean([hd.cloud_count(s, A.capi.CLOUD_EDGE) for s in range(8)]))
    b_frame = rows * cols * 4 + 2 * 16 * (n_obst + n_edge)
    out = {"what": "depth frame -> Obstacle + Edge cloud + tile index, frames resident in HBM",
           "workload": f"{B} frames {rows}x{cols} f32 ({distinct} distinct scenes), resize_scale 1",
           "value": B / (ms * 1e-3), "unit": "frames/s", "ms_per_batch": ms,
           "avg_obstacle_points": n_obst, "avg_edge_points": n_edge,
           "gpu_launches_per_batch": (hd.launch_count() - l0) / steps,
           "achieved_GBps": B * b_frame / (ms * 1e-3) / 1e9,
           "frac_of_hbm_peak": B * b_frame / (ms * 1e-3) / 1e9 / hbm_peak}
    hd.close()
    if not args.no_cpu_baseline:
        from oracle import depth_oracle as DO
        ocam = DO.Camera(**cam)
        Tm = T_np[0].reshape(4, 4)
        t0 = time.perf_counter()
        for s in range(8):
            DO.process_depth(base[s], ocam, Tm, Tm)
        out["cpu_baseline"] = {"value": 8 / (time.perf_counter() - t0), "unit": "frames/s", "cores": 1, "kind": "port",
                               "sample": "8 frames, numpy restatement of ProcessDepth + BuildEdgeCloud (no k-d tree build)"}
    return out


def cpu_c0(args, n_inst=None):
    """BASELINE configs[0] (the reference's own CPU-runnable case): single instance, N=20, K=8,
    10 000-point cloud, ONE thread; p50/p90 of tree build, 20x8-NN, NLP solve and total."""
    import avoid_mpc_b200 as A
    from oracle import oracle as O
    D, S = A.defaults, A.synth
    K, npts = 8, 10000
    n_inst = n_inst or max(50, args.steps * 50)
    use_ref = O.ref_available()
    lb, ub = D.u_bounds()
    opts = O.default_opts()
    t_build, t_knn, t_solve, iters, conv = [], [], [], [], 0
    scenes = [(S.forest_cloud(20_000 + s, npts)[0], S.states(20_000 + s, N_H)) for s in range(min(n_inst, 64))]
    for i in range(n_inst):
        c, (x0, ref, tgt) = scenes[i % len(scenes)]
        t0 = time.perf_counter()
        tree = O.RefTree(c) if use_ref else O.PortTree(c)
        t1 = time.perf_counter()
        idx, d2, cnt = tree.search(ref[:, :3], K)
        t2 = time.perf_counter()
        ob = c[idx.reshape(-1), :3].astype(np.float64).reshape(N_H, K, 3)
        p = S.full_params(S.pack_prefix(x0, ref, ob, tgt))
        w0 = S.warm_start(args.warm, x0, ref, N_H)
        t3 = time.perf_counter()
        w, info = O.solve(N_H, K, DT, p, w0, lb, ub, opts)
        t4 = time.perf_counter()
        t_build.append(t1 - t0), t_knn.append(t2 - t1), t_solve.append(t4 - t3)
        iters.append(info.iters)
        conv += info.status == 0
    tot = [a + b + c_ for a, b, c_ in zip(t_build, t_knn, t_solve)]
    q = lambda v, p_: float(np.percentile(np.array(v) * 1e3, p_))  # noqa: E731
    return {"mode": "cpu_c0", "config": "single instance, N=20, K=8, 10000-pt cloud, 1 thread",
            "instances": n_inst, "knn_impl": "reference nanoflann (oracle/_ref)" if use_ref else "C restatement",
            "nlp_impl": "oracle interior-point port (tol 1e-8)", "unit": "ms",
            "tree_build": {"p50": q(t_build, 50), "p90": q(t_build, 90)},
            "knn_20x8": {"p50": q(t_knn, 50), "p90": q(t_knn, 90)},
            "nlp_solve": {"p50": q(t_solve, 50), "p90": q(t_solve, 90)},
            "total": {"p50": q(tot, 50), "p90": q(tot, 90)},
            "iters_p50": float(np.median(iters)), "converged_frac": conv / n_inst}


# ------------------------------------------------------------------ configs[2] ----
def run_best_of(args):
    """BASELINE configs[2]: 4096 scenes x 32 Edge-tree initial guesses (131 072 NLP instances), best-cost
    reduction, 1xB200.  One call: the 32 nearest Edge points of waypoint 0, N-1+G Obstacle queries per
    scene (19 of the 20 queries are shared by a scene's guesses), 131 072 solves, best-of-32."""
    torch, dist, world, rank, local, dev = dist_setup()
    import avoid_mpc_b200 as A
    D, S = A.defaults, A.synth
    n_scenes, G, npts = args.scenes or 4096, args.guesses, args.npts
    lo, hi = A.shard.scene_range(rank, world, n_scenes)
    ns = hi - lo
    ids = list(range(lo, hi))
    stream = torch.cuda.current_stream().cuda_stream
    h = A.Handle(N=N_H, K=K_NB, dt=DT, max_batch=ns * G, max_scenes=ns, max_points=npts, max_edge_points=npts // 8, device=local)
    h.set_solver_opts(tol=args.tol, max_iter=args.max_iter)
    h.cloud_set_layout(S.image_shape(npts)[0])
    fill_scenes(A, torch, h, ids, npts, dev)
    # Edge clouds: pixels with a 4-neighbour depth jump > 0.5 m; from the numpy generator of a
    # few scenes per chunk would take minutes for 4096 scenes -- the edge set is recomputed on the device
    W_img, H_img = S.image_shape(npts)
    n_edge = []
    for s0 in range(0, ns, 256):
        c = S.forest_clouds_torch(ids[s0:s0 + 256], npts, dev)
        rng = torch.linalg.norm(c[..., :3] - torch.tensor([D.T_B_C[0, 3], D.T_B_C[1, 3], D.HEIGHT + D.T_B_C[2, 3]],
                                                         device=dev, dtype=torch.float32), dim=-1).view(-1, H_img, W_img)
        jump = torch.zeros_like(rng, dtype=torch.bool)
        jump[:, :, 1:] |= (rng[:, :, 1:] - rng[:, :, :-1]).abs() > 0.5
        jump[:, :, :-1] |= (rng[:, :, 1:] - rng[:, :, :-1]).abs() > 0.5
        jump[:, 1:, :] |= (rng[:, 1:, :] - rng[:, :-1, :]).abs() > 0.5
        jump[:, :-1, :] |= (rng[:, 1:, :] - rng[:, :-1, :]).abs() > 0.5
        jump = jump.view(-1, npts)
        cap = npts // 8
        e = torch.zeros((c.shape[0], cap, 4), dtype=torch.float32, device=dev)
        cnt = np.zeros(c.shape[0], dtype=np.int32)
        for i in range(c.shape[0]):
            pts = c[i][jump[i]][:cap]
            e[i, :pts.shape[0]] = pts
            cnt[i] = pts.shape[0]
        h.cloud_set_batch_dev(e, counts=cnt, first_scene=s0, kind=A.capi.CLOUD_EDGE, stream=stream)
        torch.cuda.synchronize()
        n_edge.extend(cnt.tolist())
        del c, e
    x0_np, ref_np, _ = S.states_batch(ids, N_H, DT)
    # waypoint 0 sits next to an obstacle in the situations PlanWapionts acts on; keep the synthetic path
    x0 = torch.tensor(x0_np, device=dev)
    ref = torch.tensor(ref_np, device=dev)
    w0_np = np.stack([S.warm_start(args.warm, x0_np[s], ref_np[s], N_H) for s in range(ns)])
    w0 = torch.tensor(np.repeat(w0_np, G, axis=0), device=dev)
    w = torch.empty_like(w0)
    info = torch.zeros((ns * G, 48), dtype=torch.uint8, device=dev)
    arg = torch.zeros(ns, dtype=torch.int32, device=dev)
    best = torch.zeros(ns, dtype=torch.float64, device=dev)
    gathered = torch.zeros(n_scenes, dtype=torch.float64, device=dev) if world > 1 else None

    def step():
        w.copy_(w0, non_blocking=True)
        h.cloud_index_dev(0, ns, stream=stream)
        h.cloud_index_dev(0, ns, kind=A.capi.CLOUD_EDGE, stream=stream)
        h.guess_round_dev(ns, G, x0, ref, w, info_dev=info, argmin_dev=arg, best_dev=best, speed=D.SPEED, stream=stream)
        if world > 1 and ns * world == n_scenes:
            dist.all_gather_into_tensor(gathered, best)

    for _ in range(max(1, args.warmup // 2)):
        step()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    h.profile_enable(True)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    steps = max(2, min(args.steps, 5))
    e0.record()
    for _ in range(steps):
        step()
    e1.record()
    torch.cuda.synchronize()
    t = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms = float(t.item()) / steps
    prof = h.profile_get()
    info_np, st = solver_stats(A, info)
    warg, wbest = A.shard.best_of_scenes(info_np["cost"], info_np["status"], G)
    a_np, b_np = arg.cpu().numpy(), best.cpu().numpy()
    parity = bool((a_np == warg).all() and (b_np[warg >= 0] == wbest[warg >= 0]).all())
    # how much the extra guesses buy: cost of the best guess relative to guess 0 (the reference's move)
    c = info_np["cost"].reshape(ns, G)
    gain = float(np.median((c[:, 0] - b_np) / np.maximum(1e-12, np.abs(c[:, 0]))))
    if rank == 0:
        print(json.dumps({"mode": "best_of", "metric": METRIC, "value": n_scenes * G / (ms * 1e-3), "unit": UNIT,
                          "n_gpus": world, "scenes": n_scenes, "guesses": G, "instances": n_scenes * G,
                          "ms_per_step": ms, "scene_ticks_per_s": n_scenes / (ms * 1e-3),
                          "stage_ms": {k: prof[k][0] / max(prof[k][1], 1) for k in ("index", "knn", "solve")},
                          "obstacle_queries_per_scene": N_H - 1 + G, "edge_points_per_scene_mean": float(np.mean(n_edge)),
                          "argmin_parity_vs_host_reduction": parity,
                          "best_differs_from_guess0_frac": float((a_np != 0).mean()),
                          "median_relative_cost_gain_over_guess0": gain, "solver": st,
                          "config": {"workload": f"{n_scenes} scenes x {G} Edge-tree guesses, N={N_H}, K={K_NB}, {npts}-pt clouds",
                                     "warm_start": args.warm, "tol": args.tol, "max_iter": args.max_iter}}), flush=True)
    if world > 1:
        dist.destroy_process_group()


# ------------------------------------------------------------------ configs[3] ----
def run_scenes65536(args):
    """BASELINE configs[3]: 65 536 scenes sharded over the GPUs (8 192 per GPU on 8), one round each,
    all-gather of the per-instance costs (512 KiB).  STRONG scaling: the job is the same at every N."""
    torch, dist, world, rank, local, dev = dist_setup()
    import avoid_mpc_b200 as A
    D, S = A.defaults, A.synth
    n_scenes, npts = args.scenes or 65536, args.npts
    lo, hi = A.shard.scene_range(rank, world, n_scenes)
    ns = hi - lo
    chunk = min(ns, args.chunk)  # scenes resident at once per GPU (6.5 GB of clouds per 8192)
    stream = torch.cuda.current_stream().cuda_stream
    h = A.Handle(N=N_H, K=K_NB, dt=DT, max_batch=chunk, max_points=npts, device=local)
    h.set_solver_opts(tol=args.tol, max_iter=args.max_iter)
    h.cloud_set_layout(S.image_shape(npts)[0])
    n_chunks = (ns + chunk - 1) // chunk
    # the clouds of every chunk stay resident (the job's input); chunks run back to back
    handles = [h] + [None] * (n_chunks - 1)
    data = []
    for ci in range(n_chunks):
        ids = list(range(lo + ci * chunk, min(hi, lo + (ci + 1) * chunk)))
        if ci > 0:
            handles[ci] = A.Handle(N=N_H, K=K_NB, dt=DT, max_batch=chunk, max_points=npts, device=local)
            handles[ci].set_solver_opts(tol=args.tol, max_iter=args.max_iter)
            handles[ci].cloud_set_layout(S.image_shape(npts)[0])
        fill_scenes(A, torch, handles[ci], ids, npts, dev)
        x0_np, ref_np, _ = S.states_batch(ids, N_H, DT)
        w0_np = np.stack([S.warm_start(args.warm, x0_np[b], ref_np[b], N_H) for b in range(len(ids))])
        data.append(dict(n=len(ids), x0=torch.tensor(x0_np, device=dev), ref=torch.tensor(ref_np, device=dev),
                         w0=torch.tensor(w0_np, device=dev), w=torch.empty((len(ids), 10 + 14 * N_H), dtype=torch.float64, device=dev),
                         info=torch.zeros((len(ids), 48), dtype=torch.uint8, device=dev)))
    costs = torch.zeros(ns, dtype=torch.float64, device=dev)
    even = ns * world == n_scenes
    gathered = torch.zeros(n_scenes, dtype=torch.float64, device=dev) if world > 1 and even else None

    def job():
        o = 0
        for hh, d in zip(handles, data):
            d["w"].copy_(d["w0"], non_blocking=True)
            hh.cloud_index_dev(0, d["n"], stream=stream)
            hh.round_dev(d["n"], d["x0"], d["ref"], d["w"], info_dev=d["info"], speed=D.SPEED,
                         safety_distance=D.SAFETY_DISTANCE, stream=stream)
            costs[o:o + d["n"]].copy_(d["info"].view(torch.float64).view(d["n"], 6)[:, 0])
            o += d["n"]
        if gathered is not None:
            dist.all_gather_into_tensor(gathered, costs)

    job()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    steps = max(2, min(args.steps, 5))
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        job()
    e1.record()
    torch.cuda.synchronize()
    t = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms = float(t.item()) / steps
    _, st = solver_stats(A, torch.cat([d["info"] for d in data]))
    if rank == 0:
        print(json.dumps({"mode": "scenes65536", "metric": METRIC, "value": n_scenes / (ms * 1e-3), "unit": UNIT,
                          "n_gpus": world, "scenes": n_scenes, "scenes_per_gpu": ns, "chunk": chunk, "ms_per_job": ms,
                          "scaling": "strong", "collective": "all_gather of %d costs (%d KiB)" % (n_scenes, n_scenes * 8 // 1024)
                          if gathered is not None else "none", "solver_rank0": st,
                          "config": {"workload": f"{n_scenes} scenes, N={N_H}, K={K_NB}, {npts}-pt clouds, one round each",
                                     "warm_start": args.warm, "tol": args.tol, "max_iter": args.max_iter}}), flush=True)
    if world > 1:
        dist.destroy_process_group()


# ------------------------------------------------------------------ configs[4] ----
def run_knn_sweep(args):
    """BASELINE configs[4]: k-NN only, cloud sizes 10k..1M points, Q=20, K=16, ~2 GB of clouds per GPU;
    achieved GB/s of the k-NN stage (index build + search) against the measured HBM peak.  Under
    torchrun every rank sweeps its own scenes; times are the max over ranks, bytes the sum."""
    torch, dist, world, rank, local, dev = dist_setup()
    import avoid_mpc_b200 as A
    S = A.synth
    stream = torch.cuda.current_stream().cuda_stream
    hbm_peak, _ = hbm_peak_gbs()
    rows, shuffled = [], []

    def tmax(v):
        t = torch.tensor([v], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    for npts in (10000, 20000, 50000, 100000, 200000, 500000, 1000000):
        B = int(min(4096, max(8, 2e9 // (16 * npts))))
        ids = list(range(rank * B, (rank + 1) * B))
        clouds = S.forest_clouds_torch(ids, npts, dev)
        _, ref, _ = S.states_batch(ids, N_H, DT)
        q = torch.tensor(np.ascontiguousarray(ref[:, :, :3]), device=dev)
        h = A.Handle(N=N_H, K=K_NB, dt=DT, max_batch=B, max_points=npts, device=local)
        h.cloud_set_layout(S.image_shape(npts)[0])
        h.cloud_set_batch_dev(clouds, stream=stream)
        idx = torch.empty((B, N_H, K_NB), dtype=torch.int32, device=dev)
        d2 = torch.empty((B, N_H, K_NB), dtype=torch.float64, device=dev)
        cnt = torch.empty((B, N_H), dtype=torch.int32, device=dev)
        ev = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
        ti, ts = [], []
        for it in range(3 + args.steps):
            if world > 1:
                dist.barrier()
            ev[0].record()
            h.cloud_index_dev(0, B, stream=stream)
            ev[1].record()
            h.knn_dev(q, K_NB, idx, d2, None, cnt, stream=stream)
            ev[2].record()
            torch.cuda.synchronize()
            if it >= 3:
                ti.append(ev[0].elapsed_time(ev[1]))
                ts.append(ev[1].elapsed_time(ev[2]))
        b_knn = 12 * npts + N_H * (24 + K_NB * 12)
        t_i, t_s = tmax(statistics.median(ti)), tmax(statistics.median(ts))
        WB = world * B
        rows.append({"npts": npts, "batch_per_gpu": B, "index_ms": t_i, "search_ms": t_s,
                     "index_GBps": WB * b_knn / (t_i * 1e-3) / 1e9, "index_frac": WB * b_knn / (t_i * 1e-3) / 1e9 / (hbm_peak * world),
                     "stage_GBps": WB * b_knn / ((t_i + t_s) * 1e-3) / 1e9,
                     "stage_frac": WB * b_knn / ((t_i + t_s) * 1e-3) / 1e9 / (hbm_peak * world),
                     "as_laid_out_16B_index_GBps": WB * 16 * npts / (t_i * 1e-3) / 1e9,
                     "scene_rounds_per_s": WB / ((t_i + t_s) * 1e-3)})
        h.close()
        # the same points in arbitrary storage order (what KDTreeTwo::InitializeNew may be given):
        # 64-consecutive-record tiles (AMPC_LAYOUT_UNORGANISED) against the sorted copy
        # (AMPC_LAYOUT_SORT).  Both must return the very same indices.
        if npts in (50000, 1000000) and world == 1:
            perm = torch.randperm(npts, device=dev, generator=torch.Generator(device=dev).manual_seed(1))
            clouds = clouds[:, perm].contiguous()
            res = {}
            for name, lay, Bs in (("unorganised", A.capi.LAYOUT_UNORGANISED, min(B, 64)), ("sort", A.capi.LAYOUT_SORT, B)):
                h = A.Handle(N=N_H, K=K_NB, dt=DT, max_batch=Bs, max_points=npts, device=local)
                h.cloud_set_layout(lay)
                h.cloud_set_batch_dev(clouds[:Bs], stream=stream)
                ti, ts = [], []
                for it in range(2 + min(args.steps, 5)):
                    ev[0].record()
                    h.cloud_index_dev(0, Bs, stream=stream)
                    ev[1].record()
                    h.knn_dev(q[:Bs], K_NB, idx[:Bs], d2[:Bs], None, cnt[:Bs], stream=stream)
                    ev[2].record()
                    torch.cuda.synchronize()
                    if it >= 2:
                        ti.append(ev[0].elapsed_time(ev[1]))
                        ts.append(ev[1].elapsed_time(ev[2]))
                t_i, t_s = statistics.median(ti), statistics.median(ts)
                res[name] = (idx[:min(B, 64)].clone(), d2[:min(B, 64)].clone())
                shuffled.append({"npts": npts, "layout": name, "batch": Bs, "index_ms": t_i, "search_ms": t_s,
                                 "stage_GBps": Bs * b_knn / ((t_i + t_s) * 1e-3) / 1e9,
                                 "stage_frac": Bs * b_knn / ((t_i + t_s) * 1e-3) / 1e9 / hbm_peak,
                                 "scene_rounds_per_s": Bs / ((t_i + t_s) * 1e-3)})
                h.close()
            shuffled[-1]["identical_to_unorganised"] = bool((res["sort"][0] == res["unorganised"][0]).all().item()
                                                            and (res["sort"][1] == res["unorganised"][1]).all().item())
        del clouds
        torch.cuda.empty_cache()
    if rank == 0:
        print(json.dumps({"mode": "knn_sweep", "unit": "GB/s", "n_gpus": world, "peak_per_gpu": hbm_peak, "Q": N_H, "K": K_NB,
                          "rows": rows, "shuffled_storage_order": shuffled}), flush=True)
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--batch", type=int, default=1024)
    ap.add_argument("--npts", type=int, default=50000)
    ap.add_argument("--warm", default="ref", choices=["ref", "cold"])
    ap.add_argument("--tol", type=float, default=1e-8)
    ap.add_argument("--max-iter", type=int, default=50)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--unorganised", action="store_true", help="do not pass the image row pitch to the index")
    ap.add_argument("--mode", default="solves", choices=["solves", "best_of", "scenes65536", "knn_sweep", "cpu_c0"])
    ap.add_argument("--in-flight", type=int, default=64,
                    help="batches of --batch instances per step (each batch has its own scenes)")
    ap.add_argument("--streams", type=int, default=2, help="calls (lanes) the batches of a step are spread over")
    ap.add_argument("--scenes", type=int, default=0, help="best_of / scenes65536: number of scenes (0 = the config's)")
    ap.add_argument("--guesses", type=int, default=32)
    ap.add_argument("--chunk", type=int, default=8192, help="scenes65536: scenes per call per GPU")
    args = ap.parse_args()
    if args.warmup < 3:
        args.warmup = 3
    if args.mode == "knn_sweep":
        run_knn_sweep(args)
    elif args.mode == "cpu_c0":
        print(json.dumps(cpu_c0(args)), flush=True)
    elif args.mode == "best_of":
        run_best_of(args)
    elif args.mode == "scenes65536":
        run_scenes65536(args)
    elif args.impl == "reference":
        run_reference_arm(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
