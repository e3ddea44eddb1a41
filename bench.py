#!/usr/bin/env python
"""Headline benchmark: MPC solves/sec at N=20, K=16 obstacle terms/stage, 50k-point clouds.

One "step" = one round of the reference's control tick for a batch of independent MPC
instances (src/AvoidanceStateMachine.cpp:328-344): Q=N k-NN queries of K neighbours on each
instance's Obstacle cloud + prefix packing + one NLP solve to convergence (tol 1e-8).

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference]

Prints ONE JSON line (rank 0).  See DESIGN.md "Measurement" for what each key means.
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "mpc_solves_per_sec"
UNIT = "solves/s"
N_H, K_NB, DT = 20, 16, 0.05


def workload_name(B, npts):
    return f"batch={B} MPC instances/GPU, N={N_H}, K={K_NB} obstacle terms/stage, {npts}-pt cloud per instance"


# ------------------------------------------------------------------ CPU arm ----
def cpu_reference(n_scenes: int, npts: int, threads: int, steps: int = 1, warm="ref"):
    """The reference's CPU path on host cores: tree build + 20x16-NN through the reference's own
    KDTreeTwo/nanoflann (oracle/_ref, falls back to the C restatement) + the oracle NLP solve.
    Returns (solves_per_sec_per_step list, description)."""
    from concurrent.futures import ThreadPoolExecutor

    import avoid_mpc_b200 as A
    from oracle import oracle as O

    D, S = A.defaults, A.synth
    use_ref = O.ref_available()
    O.lib()
    if use_ref:
        O.ref_lib()
    lb, ub = D.u_bounds()
    opts = O.default_opts()
    n_distinct = min(n_scenes, 48)
    clouds = [S.forest_cloud(10_000 + s, npts)[0] for s in range(n_distinct)]
    states = [S.states(10_000 + s, N_H) for s in range(n_distinct)]

    def one(i):
        c = clouds[i % n_distinct]
        x0, ref, tgt = states[i % n_distinct]
        tree = O.RefTree(c) if use_ref else O.PortTree(c)       # a1: build (every depth frame)
        idx, d2, cnt = tree.search(ref[:, :3], K_NB)             # a2/a6: N x K-NN
        ob = c[idx.reshape(-1), :3].astype(np.float64).reshape(N_H, K_NB, 3)
        p = S.full_params(S.pack_prefix(x0, ref, ob, tgt))       # a7
        w, info = O.solve(N_H, K_NB, DT, p, S.warm_start(warm, x0, ref, N_H), lb, ub, opts)  # a8
        return info.status

    rates = []
    with ThreadPoolExecutor(max_workers=threads) as ex:
        list(ex.map(one, range(min(n_scenes, 2 * threads))))  # warm-up
        for _ in range(steps):
            t0 = time.perf_counter()
            list(ex.map(one, range(n_scenes)))
            rates.append(n_scenes / (time.perf_counter() - t0))
    kind = "port"
    desc = (f"{n_scenes} instances/step ({n_distinct} distinct {npts}-pt scenes), {threads} host threads; "
            f"k-NN = {'reference nanoflann (oracle/_ref)' if use_ref else 'C restatement'} tree build + "
            f"{N_H}x{K_NB}-NN, NLP = oracle interior-point port (CasADi/IPOPT not installable), tol 1e-8")
    return rates, kind, desc


def run_reference_arm(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    threads = os.cpu_count() or 1
    n = max(threads * 32, 256)
    rates, kind, desc = cpu_reference(n, args.npts, threads, steps=args.warmup + args.steps)
    rates = rates[args.warmup:]
    v = statistics.mean(rates)
    out = {
        "impl": "reference", "metric": METRIC, "value": v, "unit": UNIT, "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * n / v, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": workload_name(args.batch, args.npts), "sample_instances_per_step": n},
        "cpu_baseline": {"value": v, "unit": UNIT, "cores": threads, "kind": kind, "sample": desc},
        "e2e": {"value": v, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(out), flush=True)


# ------------------------------------------------------------------ GPU arm ----
class ClockSampler:
    """nvidia-smi polled every 50 ms by a reader thread that stamps each sample on arrival; the
    samples that arrived inside [mark_begin(), mark_end()] are reported.  nvidia-smi needs about
    a second to deliver its first sample, so it is started before the warm-up and the window
    covers the loaded warm-up steps plus the timed region (the GPU is busy throughout)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
         "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, dev):
        self.dev, self.p, self.rows, self.t0, self.t1, self.thread = dev, None, [], None, None, None

    def start(self):
        import threading
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
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ts, line in list(self.rows):
            if self.t0 is not None and (ts < self.t0 or ts > (self.t1 or ts) + 0.05):
                continue
            f = [x.strip() for x in line.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1])), mx.append(float(f[2]))
            except ValueError:
                continue
            for nm, val in zip(names, f[5:9]):
                if val.lower().startswith("active"):
                    reasons.add(nm)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "samples": len(sm), "reasons": sorted(reasons),
                "window": "loaded warm-up steps + single-stream pass + timed region"}


def run_ours(args):
    import torch
    import torch.distributed as dist

    import avoid_mpc_b200 as A
    D, S = A.defaults, A.synth

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the product path has no CPU fallback")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    B, npts = args.batch, args.npts

    # ---- synthetic inputs (weak scaling: B scenes per GPU, distinct per rank) ----
    ids = list(range(rank * B, (rank + 1) * B))
    clouds = S.forest_clouds_torch(ids, npts, dev)                      # (B, npts, 4) f32, resident
    x0_np, ref_np, _ = S.states_batch(ids, N_H, DT)
    w0_np = np.stack([S.warm_start(args.warm, x0_np[b], ref_np[b], N_H) for b in range(B)])
    # `streams` independent batches in flight (a batch server keeps several ticks' batches in
    # flight; each has its own depth frames, i.e. its own handle + clouds + outputs + stream).
    # streams = 1 exposes every step's full latency (the solve kernel waits for its slowest
    # instance while most SMs idle); the single-stream figure is reported alongside.
    def make_lane():
        hh = A.Handle(N=N_H, K=K_NB, dt=DT, max_batch=B, max_points=npts, device=local)
        hh.set_solver_opts(tol=args.tol, max_iter=args.max_iter)
        if not args.unorganised:  # the clouds are row-major depth images: tell the index the row pitch
            hh.cloud_set_layout(S.image_shape(npts)[0])
        st = torch.cuda.Stream(device=dev)
        hh.cloud_set_batch_dev(clouds, stream=torch.cuda.current_stream().cuda_stream)
        torch.cuda.synchronize()
        lane = dict(h=hh, st=st, w=torch.empty((B, 10 + 14 * N_H), dtype=torch.float64, device=dev),
                    info=torch.zeros((B, 48), dtype=torch.uint8, device=dev),
                    replan=torch.zeros(B, dtype=torch.int32, device=dev),
                    costs=torch.zeros(B, dtype=torch.float64, device=dev))
        lane["info_f64"] = lane["info"].view(torch.float64).view(B, 6)
        return lane

    n_streams = max(1, args.streams)
    lanes = [make_lane() for _ in range(n_streams)]
    h = lanes[0]["h"]
    x0 = torch.tensor(x0_np, device=dev)
    ref = torch.tensor(ref_np, device=dev)
    w0 = torch.tensor(w0_np, device=dev)
    gathered = torch.zeros(B * world, dtype=torch.float64, device=dev) if world > 1 else None
    torch.cuda.synchronize()

    def step(lane):
        hh, st = lane["h"], lane["st"]
        with torch.cuda.stream(st):
            lane["w"].copy_(w0, non_blocking=True)
            # a new depth frame per solve (the reference rebuilds its KD-trees every frame,
            # src/FrameKDMap.cpp:34-52): index build = the one streaming pass over the cloud
            hh.cloud_index_dev(0, B, stream=st.cuda_stream)
            hh.round_dev(B, x0, ref, lane["w"], info_dev=lane["info"], replan_dev=lane["replan"], speed=D.SPEED,
                         safety_distance=D.SAFETY_DISTANCE, stream=st.cuda_stream)
            if world > 1:  # per-instance best-cost exchange: the only collective of the path
                lane["costs"].copy_(lane["info_f64"][:, 0])
                dist.all_gather_into_tensor(gathered, lane["costs"])

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(lanes_used, steps):
        """K steps round-robin over the lanes; device time from a start event every lane waits on
        to an end event recorded after every lane has finished."""
        main = torch.cuda.current_stream()
        start, end = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        sev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(steps)]
        barrier()
        start.record(main)
        for ln in lanes_used:
            ln["st"].wait_event(start)
        for i in range(steps):
            ln = lanes_used[i % len(lanes_used)]
            sev[i][0].record(ln["st"])
            step(ln)
            sev[i][1].record(ln["st"])
        for ln in lanes_used:
            main.wait_stream(ln["st"])
        end.record(main)
        barrier()
        return start.elapsed_time(end), [a.elapsed_time(b) for a, b in sev]

    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    for _ in range(args.warmup):
        for ln in lanes:
            step(ln)
    barrier()
    # keep the GPU loaded (untimed; the same count on every rank) until the clock sampler has
    # something to report, then go straight into the measured passes
    sampler.mark_begin()
    for _ in range(args.load_rounds):
        for ln in lanes:
            step(ln)
    barrier()
    # single-stream pass: one batch in flight, every kernel runs alone -> the per-kernel times used
    # for the roofline objects (under several streams a kernel's duration includes time-sharing)
    lanes[0]["h"].profile_enable(True)
    n_single = max(3, args.steps // 2)
    single_ms, single_step_ms = timed(lanes[:1], n_single)
    prof1 = lanes[0]["h"].profile_get()
    lanes[0]["h"].profile_enable(False)
    for ln in lanes:
        ln["h"].profile_enable(True)
    l0 = sum(ln["h"].launch_count() for ln in lanes)
    total_ms, step_ms = timed(lanes, args.steps)
    launches = sum(ln["h"].launch_count() for ln in lanes) - l0
    ov = {"index": 0.0, "knn": 0.0, "solve": 0.0, "n": 0}
    for ln in lanes:
        prof = ln["h"].profile_get()
        for kk in ("index", "knn", "solve"):
            ov[kk] += prof[kk][0]
        ov["n"] += prof["solve"][1]
        ln["h"].profile_enable(False)
    index_ms, knn_ms, solve_ms = prof1["index"][0], prof1["knn"][0], prof1["solve"][0]
    rounds = prof1["solve"][1]
    t = torch.tensor([total_ms], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    total_ms = float(t.item())
    sampler.mark_end()
    clocks = sampler.stop() if rank == 0 else None
    info = lanes[0]["info"]
    replan = lanes[0]["replan"]
    w = lanes[0]["w"]
    info_np = info.cpu().numpy().view(A.capi.INFO_DTYPE).reshape(B)
    value = world * B * args.steps / (total_ms * 1e-3)
    single_value = world * B * n_single / (single_ms * 1e-3)

    # ---- end to end through the host-buffer C-ABI: clouds + states from pinned host memory,
    #      trajectories/costs/status back to host, every step ----
    clouds_h = torch.empty(clouds.shape, dtype=torch.float32, pin_memory=True)
    clouds_h.copy_(clouds)
    x0_h = torch.tensor(x0_np).pin_memory()
    ref_h = torch.tensor(ref_np).pin_memory()
    w0_h = torch.tensor(w0_np).pin_memory()
    # two host threads, each driving its own handle through the synchronous host-buffer calls:
    # the upload of one batch's clouds overlaps the k-NN + solve of the other (ctypes releases
    # the GIL).  Every step still copies all of its inputs in and all of its results out.
    from concurrent.futures import ThreadPoolExecutor
    e2e_lanes = lanes[:2]
    for ln in e2e_lanes:
        ln["w_h"] = torch.empty_like(w0_h).pin_memory()
        ln["info_h"] = torch.zeros((B, 48), dtype=torch.uint8).pin_memory()
        ln["replan_h"] = torch.zeros(B, dtype=torch.int32).pin_memory()
    w_h, info_h, replan_h = e2e_lanes[0]["w_h"], e2e_lanes[0]["info_h"], e2e_lanes[0]["replan_h"]
    torch.cuda.synchronize()

    def e2e_step(ln):
        ln["w_h"].copy_(w0_h)
        ln["h"].cloud_set_batch(clouds_h)                             # H2D of this step's clouds (+ index build)
        ln["h"].round_host_ptrs(B, x0_h, ref_h, ln["w_h"], ln["info_h"], ln["replan_h"], speed=D.SPEED,
                                safety_distance=D.SAFETY_DISTANCE)    # H2D states, k-NN, solve, D2H results

    e2e_steps = max(4, min(args.steps, 10))
    with ThreadPoolExecutor(max_workers=len(e2e_lanes)) as ex:
        list(ex.map(e2e_step, [e2e_lanes[i % len(e2e_lanes)] for i in range(2 * len(e2e_lanes))]))
        barrier()
        t0 = time.perf_counter()
        list(ex.map(e2e_step, [e2e_lanes[i % len(e2e_lanes)] for i in range(e2e_steps)]))
        barrier()
        e2e_s = time.perf_counter() - t0
    t = torch.tensor([e2e_s], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    e2e_value = world * B * e2e_steps / float(t.item())
    h2d = clouds_h.numel() * 4 + (x0_h.numel() + ref_h.numel() + w_h.numel()) * 8
    d2h = w_h.numel() * 8 + info_h.numel() + replan_h.numel() * 4

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    # ---- roofline of the HBM-bound kernel (k-NN scan): SURVEY.md §8d algorithmic bytes ----
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    hbm_peak = float(peaks.get("hbm_gbs", 6650.0))
    b_knn = 12 * npts + N_H * (24 + K_NB * 12)
    traffic = None  # dram bytes per launch of the index kernel, from the committed ncu --set full capture
    try:
        m = json.load(open(os.path.join(ROOT, "profiles", "ncu_summary_r01.json")))["cloud_index"]["metrics"]
        scale = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
        traffic = sum(float(m[k]["value"]) * scale[m[k]["unit"]] for k in ("dram__bytes_read.sum", "dram__bytes_write.sum"))
    except Exception:
        pass
    index_ms_avg = index_ms / max(rounds, 1)
    knn_ms_avg = knn_ms / max(rounds, 1)
    stage_ms = index_ms_avg + knn_ms_avg
    achieved = B * b_knn / (index_ms_avg * 1e-3) / 1e9
    denom = sum(single_step_ms)  # per-kernel times and shares come from the single-stream pass
    roofline = {"kernel": "cloud_index_kernel", "bound": "hbm", "achieved": achieved, "peak": hbm_peak,
                "unit": "GB/s", "frac": achieved / hbm_peak, "traffic": traffic,
                "traffic_note": "ncu dram__bytes_read+write of one launch at this workload (profiles/ncu_summary_r01.json); "
                                "the reference's 16-byte pcl::PointXYZ records carry 4 padding bytes per point, so traffic "
                                "= 16/12 x algorithmic + boxes",
                "peak_source": "MEASURED_PEAKS.json (measured)" if peaks else "fallback 6650 GB/s",
                "algorithmic_bytes_per_launch": B * b_knn,
                "as_laid_out_16B_GBps": B * 16 * npts / (index_ms_avg * 1e-3) / 1e9,
                "avg_launch_ms": index_ms_avg, "step_share": index_ms / denom,
                "measured": "CUDA events around the kernel on its stream, single-stream pass of this run (kernel runs alone)",
                "avg_launch_ms_with_%d_batches_in_flight" % n_streams: ov["index"] / max(ov["n"], 1),
                "knn_stage": {"what": "index build + box-pruned search (k-NN indices bit-exact)",
                              "index_ms": index_ms_avg, "search_ms": knn_ms_avg,
                              "achieved_GBps": B * b_knn / (stage_ms * 1e-3) / 1e9,
                              "frac": B * b_knn / (stage_ms * 1e-3) / 1e9 / hbm_peak,
                              "step_share": (index_ms + knn_ms) / denom}}
    # ---- NLP kernel: FP64 work vs a DGEMM-measured FP64 peak on this box ----
    a = torch.randn(4096, 4096, dtype=torch.float64, device=dev)
    for _ in range(2):
        a @ a
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(5):
        a @ a
    e1.record()
    torch.cuda.synchronize()
    fp64_peak = 5 * 2 * 4096 ** 3 / (e0.elapsed_time(e1) * 1e-3) / 1e12
    iters = info_np["iters"].astype(np.float64)
    f_iter = (N_H - 1) * K_NB * 250 + N_H * 4275
    solve_ms_avg = solve_ms / max(rounds, 1)
    nlp_tflops = float(iters.sum()) * f_iter / (solve_ms_avg * 1e-3) / 1e12
    roofline_nlp = {"kernel": "ipm_solve_kernel", "bound": "fp64 latency", "achieved": nlp_tflops, "peak": fp64_peak,
                    "unit": "TFLOP/s", "frac": nlp_tflops / fp64_peak, "peak_source": "torch f64 matmul 4096^3 on this GPU",
                    "flop_per_iter": f_iter, "avg_launch_ms": solve_ms_avg,
                    "step_share": solve_ms / denom}

    # ---- single-instance latency through the host API (cloud resident) ----
    lat = []
    for _ in range(30):
        t0 = time.perf_counter()
        h.round_host_ptrs(1, x0_h, ref_h, w_h, info_h, replan_h, speed=D.SPEED, safety_distance=D.SAFETY_DISTANCE)
        lat.append((time.perf_counter() - t0) * 1e3)
    lat = lat[5:]

    out = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": total_ms / args.steps, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": workload_name(B, npts), "global_batch": world * B, "warm_start": args.warm,
                   "tol": args.tol, "max_iter": args.max_iter, "parallelism": f"scene-sharded x{world}",
                   "streams": n_streams, "streams_note": "independent batches in flight per GPU, each with its own frames",
                   "l2": "inputs larger than L2 (%.0f MB of clouds per step per GPU)" % (B * npts * 16 / 1e6),
                   "collective": "all_gather of per-instance costs (NCCL)" if world > 1 else "none"},
        "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                "steps": e2e_steps, "host_threads": len(e2e_lanes),
                "note": "every step uploads all of its clouds and states from pinned host memory and reads all results back; "
                        "PCIe-bound (clouds are 800 KB per instance)"},
        "gpu_launches": int(launches),
        "clocks": clocks,
        "roofline": roofline,
        "roofline_nlp": roofline_nlp,
        "single_stream": {"value": single_value, "unit": UNIT, "steps": n_single, "ms_per_step": single_ms / n_single,
                          "note": "same steps, one batch in flight"},
        "kernel_ms_overlapped": {"index": ov["index"] / max(ov["n"], 1), "knn_search": ov["knn"] / max(ov["n"], 1),
                                 "solve": ov["solve"] / max(ov["n"], 1), "note": "%d batches in flight" % n_streams},
        "latency": {"batch_step_ms_p50": statistics.median(single_step_ms), "batch_step_ms_p50_overlapped": statistics.median(step_ms), "single_instance_round_ms_p50": statistics.median(lat)},
        "solver": {"converged_frac": float((info_np["status"] == 0).mean()),
                   "status_counts": np.bincount(info_np["status"], minlength=4).tolist(),
                   "iters_p50": float(np.median(iters)), "iters_p90": float(np.percentile(iters, 90)),
                   "iters_max": int(iters.max()), "need_replan_frac": float(replan.float().mean().item())},
    }
    if world == 1:
        out["depth_path"] = depth_leg(A, torch, dev, local, hbm_peak, args, cpu=not args.no_cpu_baseline)
    if world == 1 and not args.no_cpu_baseline:
        threads = os.cpu_count() or 1
        n = max(threads * 16, 128)
        rates, kind, desc = cpu_reference(n, npts, threads, steps=2, warm=args.warm)
        out["cpu_baseline"] = {"value": rates[-1], "unit": UNIT, "cores": threads, "kind": kind, "sample": desc}
    else:
        out["cpu_baseline"] = None
    print(json.dumps(out), flush=True)
    if world > 1:
        dist.destroy_process_group()


def depth_leg(A, torch, dev, local, hbm_peak, args, cpu=True, rows=200, cols=250, distinct=128):
    """The step upstream of the clouds, reported next to the headline and not
    part of it: (1) depth frames resident in HBM -> Obstacle + Edge clouds + tile indices
    (ampc_depth_set_batch_dev), with the CPU restatement of FrameKDMap::ProcessDepth timed beside
    it; (2) the control round measured end to end from HOST depth frames (what the reference's
    AddVertex receives) instead of host clouds: 4 bytes per point over PCIe instead of 16."""
    from concurrent.futures import ThreadPoolExecutor
    D, S = A.defaults, A.synth
    B = args.batch
    cam = dict(fx=cols / 2, fy=cols / 2, cx=cols / 2, cy=rows / 2, resize_scale=1.0)
    ids = [b % distinct for b in range(B)]
    base = np.stack([S.forest_depth(s, rows, cols, sky=False) for s in range(distinct)])  # 50k-point clouds
    depth_h = torch.from_numpy(base).repeat((B + distinct - 1) // distinct, 1, 1)[:B].contiguous().pin_memory()
    Twb = np.eye(4)
    Twb[2, 3] = D.HEIGHT
    T_np = np.ascontiguousarray(np.tile((Twb @ D.T_B_C).reshape(1, 16), (B, 1)))
    x0_np, ref_np, _ = S.states_batch(ids, N_H, DT)
    w0_np = np.stack([S.warm_start(args.warm, x0_np[b], ref_np[b], N_H) for b in range(B)])
    x0_h, ref_h, w0_h = (torch.tensor(a).pin_memory() for a in (x0_np, ref_np, w0_np))

    def make():
        hh = A.Handle(N=N_H, K=K_NB, dt=DT, max_batch=B, max_points=rows * cols, max_edge_points=rows * cols // 4,
                      device=local)
        hh.set_solver_opts(tol=args.tol, max_iter=args.max_iter)
        hh.set_camera(**cam)
        return dict(h=hh, w_h=torch.empty_like(w0_h).pin_memory(), info_h=torch.zeros((B, 48), dtype=torch.uint8).pin_memory(),
                    replan_h=torch.zeros(B, dtype=torch.int32).pin_memory())

    lanes = [make(), make()]
    hd = lanes[0]["h"]
    # (1) resident
    depth = depth_h.to(dev)
    T = torch.from_numpy(T_np).to(dev)
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
    # algorithmic bytes per frame: every source pixel once, the 16-byte records written, and the
    # index pass reading them once
    b_frame = rows * cols * 4 + 2 * 16 * (n_obst + n_edge)
    out = {"what": "depth frame -> Obstacle + Edge cloud + tile index, frames resident in HBM",
           "workload": f"{B} frames {rows}x{cols} f32 ({distinct} distinct scenes), resize_scale 1",
           "value": B / (ms * 1e-3), "unit": "frames/s", "ms_per_batch": ms,
           "avg_obstacle_points": n_obst, "avg_edge_points": n_edge,
           "gpu_launches_per_batch": (hd.launch_count() - l0) / steps,
           "achieved_GBps": B * b_frame / (ms * 1e-3) / 1e9,
           "frac_of_hbm_peak": B * b_frame / (ms * 1e-3) / 1e9 / hbm_peak}
    del depth

    # (2) end to end from host depth frames
    def e2e_step(ln):
        ln["w_h"].copy_(w0_h)
        ln["h"].depth_set_batch(depth_h.numpy(), T_np)                # H2D depth frames, clouds + indices on the device
        ln["h"].round_host_ptrs(B, x0_h, ref_h, ln["w_h"], ln["info_h"], ln["replan_h"], speed=D.SPEED,
                                safety_distance=D.SAFETY_DISTANCE)    # H2D states, k-NN, solve, D2H results

    n = 8
    with ThreadPoolExecutor(max_workers=2) as ex:
        list(ex.map(e2e_step, [lanes[i % 2] for i in range(4)]))
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        list(ex.map(e2e_step, [lanes[i % 2] for i in range(n)]))
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
    info = lanes[0]["info_h"].numpy().view(A.capi.INFO_DTYPE).reshape(B)
    out["e2e_from_depth"] = {"value": B * n / dt, "unit": UNIT, "steps": n, "host_threads": 2,
                             "h2d_bytes_per_step": depth_h.numel() * 4 + T_np.nbytes + (x0_h.numel() + ref_h.numel() + w0_h.numel()) * 8,
                             "d2h_bytes_per_step": w0_h.numel() * 8 + B * 48 + B * 4,
                             "converged_frac": float((info["status"] == 0).mean()),
                             "note": "same round as e2e, but the host hands over the depth frames FrameKDMap::AddVertex receives "
                                     "(f32 metres) and the clouds are built on the device"}
    if cpu:
        from oracle import depth_oracle as DO
        ocam = DO.Camera(**cam)
        Tm = T_np[0].reshape(4, 4)
        t0 = time.perf_counter()
        for s in range(8):
            DO.process_depth(base[s], ocam, Tm, Tm)
        out["cpu_baseline"] = {"value": 8 / (time.perf_counter() - t0), "unit": "frames/s", "cores": 1, "kind": "port",
                               "sample": "8 frames, numpy restatement of ProcessDepth + BuildEdgeCloud (no k-d tree build)"}
    for ln in lanes:
        ln["h"].close()
    return out


def run_cpu_c0(args):
    """BASELINE config C0 (the reference's own CPU-runnable case): single instance, N=20, K=8,
    10 000-point cloud, ONE thread; p50/p90 of tree build, 20x8-NN, NLP solve and total."""
    import avoid_mpc_b200 as A
    from oracle import oracle as O
    D, S = A.defaults, A.synth
    K, npts, n_inst = 8, 10000, max(50, args.steps * 50)
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
    q = lambda v, p_: float(np.percentile(np.array(v) * 1e3, p_))
    print(json.dumps({"mode": "cpu_c0", "config": "single instance, N=20, K=8, 10000-pt cloud, 1 thread",
                      "instances": n_inst, "knn_impl": "reference nanoflann (oracle/_ref)" if use_ref else "C restatement",
                      "nlp_impl": "oracle interior-point port (tol 1e-8)", "unit": "ms",
                      "tree_build": {"p50": q(t_build, 50), "p90": q(t_build, 90)},
                      "knn_20x8": {"p50": q(t_knn, 50), "p90": q(t_knn, 90)},
                      "nlp_solve": {"p50": q(t_solve, 50), "p90": q(t_solve, 90)},
                      "total": {"p50": q(tot, 50), "p90": q(tot, 90)},
                      "iters_p50": float(np.median(iters)), "converged_frac": conv / n_inst}), flush=True)


def run_knn_sweep(args):
    """BASELINE config C4: k-NN only, cloud sizes 10k..1M points, Q=20, K=16, ~2 GB of clouds per
    GPU; achieved GB/s of the k-NN stage (index build + search) against the measured HBM peak."""
    import torch
    import avoid_mpc_b200 as A
    S = A.synth
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    stream = torch.cuda.current_stream().cuda_stream
    try:
        hbm_peak = float(json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"])
    except Exception:
        hbm_peak = 6650.0
    rows, shuffled = [], []
    for npts in (10000, 20000, 50000, 100000, 200000, 500000, 1000000):
        B = int(min(4096, max(8, 2e9 // (16 * npts))))
        ids = list(range(B))
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
        t_i, t_s = statistics.median(ti), statistics.median(ts)
        rows.append({"npts": npts, "batch": B, "index_ms": t_i, "search_ms": t_s,
                     "index_GBps": B * b_knn / (t_i * 1e-3) / 1e9, "index_frac": B * b_knn / (t_i * 1e-3) / 1e9 / hbm_peak,
                     "stage_GBps": B * b_knn / ((t_i + t_s) * 1e-3) / 1e9,
                     "stage_frac": B * b_knn / ((t_i + t_s) * 1e-3) / 1e9 / hbm_peak,
                     "as_laid_out_16B_index_GBps": B * 16 * npts / (t_i * 1e-3) / 1e9,
                     "scene_rounds_per_s": B / ((t_i + t_s) * 1e-3)})
        h.close()
        # the same points in arbitrary storage order (what KDTreeTwo::InitializeNew may be given):
        # 64-consecutive-record tiles (AMPC_LAYOUT_UNORGANISED) against the Morton-bucketed copy
        # (AMPC_LAYOUT_SORT).  Both must return the very same indices.
        if npts in (50000, 1000000):
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
    print(json.dumps({"mode": "knn_sweep", "unit": "GB/s", "peak": hbm_peak, "Q": N_H, "K": K_NB, "rows": rows,
                      "shuffled_storage_order": shuffled}), flush=True)


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
    ap.add_argument("--mode", default="solves", choices=["solves", "knn_sweep", "cpu_c0"])
    ap.add_argument("--streams", type=int, default=8, help="independent batches in flight per GPU")
    ap.add_argument("--load-rounds", type=int, default=100,
                    help="untimed rounds over all lanes before the measured passes (clock sampling window)")
    args = ap.parse_args()
    if args.warmup < 3:
        args.warmup = 3
    if args.mode == "knn_sweep":
        run_knn_sweep(args)
    elif args.mode == "cpu_c0":
        run_cpu_c0(args)
    elif args.impl == "reference":
        run_reference_arm(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
