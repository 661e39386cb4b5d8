#!/usr/bin/env python
"""Benchmark of the scan-to-map registration hot path (see DESIGN.md "Measurement").

Workload (BASELINE.json configs 2 and 4): point-to-plane ICP of synthetic 32-beam scans (32 x 940 rays, ~27.6 k
points each) against a 1 M-point synthetic map, device-resident Gauss-Newton loop, max_iteration = 10, eps = 0
(exactly 10 iterations).  A step registers a batch of `--scans-per-gpu` scans (default 512 = config 4's 4096 scans
over 8 GPUs) on every GPU; ranks hold a replica of the map and a disjoint block of scans, with no collective on the
data path (weak scaling).  The single-scan tracking latency of config 2 is reported in the same line ("track").

  python bench.py [--gpus N] [--steps K] [--warmup W]            our arm (one process per GPU under torchrun)
  python bench.py --impl reference [...]                         CPU arm: the oracle's restatement of the
                                                                 reference loop on all host threads (bounded sample)
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "registered_points_per_sec"
UNIT = "points/s"
BYTES_PER_POINT_ITER = 96  # 16 B source float4 + 5 x 16 B gathered neighbours
MAX_ITER = 10


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--scans-per-gpu", type=int, default=512)
    ap.add_argument("--map-points", type=int, default=1_000_000)
    ap.add_argument("--cpu-sample-scans", type=int, default=8)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--workload", default="icp", choices=["icp", "ndt", "reloc"],
                    help="icp: configs 2/4 (the headline line); ndt: config 3; reloc: config 5 (extra report lines)")
    ap.add_argument("--hyp", type=int, default=65536, help="reloc: number of pose hypotheses (64x64 xy grid x 16 yaws)")
    ap.add_argument("--ndt-map-points", type=int, default=20_000_000)
    return ap.parse_args()


def make_world(args):
    from loc_lib_b200 import synth
    w = synth.World(200.0)
    return w, w.sample_map(args.map_points)


def make_scans(world, first, count, total):
    """Scans [first, first+count) of a `total`-long seeded random walk; returns (clouds (n,4), offsets, init poses, gt)."""
    from loc_lib_b200 import synth
    gt_all = world.poses(total)
    gt = gt_all[first:first + count]
    buf, counts = world.scan_batch(gt, seed=synth.SEED_SCAN + first)
    clouds = np.concatenate([buf[i, :counts[i]] for i in range(count)])
    offsets = np.concatenate([[0], np.cumsum(counts)]).astype(np.int64)
    init = np.stack([synth.perturb_pose(g, synth.SEED_POSE + 7919 * (first + i)) for i, g in enumerate(gt)])
    return clouds, offsets, init, gt


class ClockSampler(threading.Thread):
    """nvidia-smi clocks / throttle reasons of one GPU sampled every 200 ms while the timed region runs."""
    Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.rows, self.proc = index, [], None

    def run(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "200"], stdout=subprocess.PIPE, text=True)
            for line in self.proc.stdout:
                self.rows.append([c.strip() for c in line.split(",")])
        except OSError:
            pass

    def stop(self):
        if self.proc:
            self.proc.terminate()
        self.join(timeout=2)
        sm = [float(r[0]) for r in self.rows if r and r[0].replace(".", "").isdigit()]
        mx = [float(r[1]) for r in self.rows if len(r) > 1 and r[1].replace(".", "").isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = sorted({names[i] for r in self.rows if len(r) >= 6 for i in range(4) if r[2 + i].lower().startswith("active")})
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": reasons, "samples": len(sm)}


def cpu_baseline(map_cloud, clouds, offsets, init, n_scans, threads):
    """Times the oracle (std-only restatement of IcpRegistration, literal always-on ANN kd-tree search = what the
    reference runs) on `n_scans` scans with `threads` host threads.  Map build is not included (as for the GPU)."""
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import oracle_py as O
    ref = O.OracleIcp(method=O.P2PLANE, max_iteration=MAX_ITER, eps=0.0, nn_mode=O.NN_LITERAL_ANN)
    t0 = time.perf_counter()
    ref.set_target(map_cloud)
    build_s = time.perf_counter() - t0
    S = min(n_scans, len(offsets) - 1)
    t0 = time.perf_counter()
    _, _, used = ref.align_batch(clouds, offsets[:S + 1], init[:S], threads=threads)
    dt = time.perf_counter() - t0
    pts = int(offsets[S])
    return {"value": pts / dt, "unit": UNIT, "cores": int(used), "kind": "port", "scans_per_s": S / dt,
            "sample": f"{S} scans ({pts} points) x {MAX_ITER} GN iterations, oracle ANN kd-tree, {used} thread(s), "
                      f"{dt:.1f} s; kd-tree build {build_s:.1f} s not included",
            "seconds": dt}


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    world, map_cloud = make_world(args)
    threads = os.cpu_count() or 1
    S = max(4 * threads, 8)  # bounded sample of the workload: ~1-2 s of CPU work per step on all host threads
    clouds, offsets, init, _ = make_scans(world, 0, S, max(S, 4096))
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import oracle_py as O
    ref = O.OracleIcp(method=O.P2PLANE, max_iteration=MAX_ITER, eps=0.0, nn_mode=O.NN_LITERAL_ANN)
    ref.set_target(map_cloud)
    for _ in range(min(args.warmup, 1)):
        ref.align_batch(clouds, offsets[:threads + 1] if S >= threads else offsets, init[:threads], threads=threads)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        _, _, used = ref.align_batch(clouds, offsets, init, threads=threads)
    dt = time.perf_counter() - t0
    pts = int(offsets[-1]) * args.steps
    val = pts / dt
    line = {"impl": "reference", "metric": METRIC, "value": val, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": dt / args.steps * 1e3, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": workload_config(args, args.scans_per_gpu,
                                      note="reference arm: each step is a bounded sample of this workload, %d scans on %d host threads" % (S, used)),
            "cpu_baseline": {"value": val, "unit": UNIT, "cores": int(used), "kind": "port",
                             "sample": f"{S} scans x {args.steps} steps, oracle ANN kd-tree (reference semantics)"},
            "e2e": {"value": val, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "scans_per_s": S * args.steps / dt}
    print(json.dumps(line))


def workload_config(args, scans_per_step, note=""):
    return {"workload": "C2/C4: point-to-plane ICP, 32x940-ray synthetic scans (~27.6k pts) vs 1M-pt synthetic map, "
                        "10 Gauss-Newton iterations (eps=0), device-resident loop",
            "scans_per_gpu_per_step": scans_per_step, "map_points": args.map_points, "max_iteration": MAX_ITER,
            "timing": "inputs larger than L2 (scan batch %.0f MB per step); CUDA events on the launch stream" %
                      (scans_per_step * 27600 * 16 / 1e6), "note": note}


def run_ours(args):
    import torch
    import torch.distributed as dist
    import loc_lib_b200 as L
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    world_size = int(os.environ.get("WORLD_SIZE", "1"))
    if world_size > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)

    world, map_cloud = make_world(args)
    B = args.scans_per_gpu
    clouds, offsets, init, gt = make_scans(world, rank * B, B, world_size * B)
    n_pts = int(offsets[-1])

    reg = L.IcpRegistration(L.IcpOptions(method_=L.IcpMethod.P2PLANE, max_iteration_=MAX_ITER, eps_=0.0), device=local)
    stream = torch.cuda.current_stream(dev)
    reg.set_stream(stream.cuda_stream)
    t0 = time.perf_counter()
    reg.SetInputTarget(map_cloud)
    map_build_ms = (time.perf_counter() - t0) * 1e3
    map_kernel_ms = reg.last_timing()[0]

    # ---- resident inputs for `value`
    d_src = torch.from_numpy(clouds).to(dev)
    d_off = torch.from_numpy(offsets).to(dev)
    d_pin = torch.from_numpy(init).to(dev)
    d_pout = torch.zeros_like(d_pin)
    d_res = torch.zeros(B * 48, dtype=torch.uint8, device=dev)

    def step_resident():
        reg.ScanMatchBatchDevice(d_src.data_ptr(), d_off.data_ptr(), d_pin.data_ptr(), B, n_pts, d_pout.data_ptr(),
                                 d_res.data_ptr())
        return reg.last_timing()[1]

    # ---- pinned host inputs for `e2e` (the C-ABI call a MatchingInterface user makes, host buffers)
    h_src = torch.from_numpy(clouds).pin_memory()
    h_off = offsets
    h_pin = init
    h2d = h_src.numel() * 4 + h_off.nbytes + 2 * h_pin.nbytes
    d2h = B * 7 * 8 + B * 48

    def step_e2e():
        poses, res = reg.ScanMatchBatch(h_src.numpy(), h_off, h_pin)
        return poses

    def barrier():
        if world_size > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)

    for _ in range(max(args.warmup, 3)):
        step_resident()
    # sanity: the registered poses must be the right ones (median translation error vs ground truth)
    err = np.median(np.linalg.norm(d_pout.cpu().numpy()[:, 4:] - gt[:, 4:], axis=1))

    sampler = ClockSampler(local)
    sampler.start()
    time.sleep(0.25)
    barrier()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    wall0 = time.perf_counter()
    ev0.record(stream)
    launches = 0
    for _ in range(args.steps):
        launches += step_resident()
    ev1.record(stream)
    barrier()
    wall = time.perf_counter() - wall0
    dev_ms = ev0.elapsed_time(ev1)
    clocks = sampler.stop()

    # ---- end to end through host buffers
    for _ in range(2):
        step_e2e()
    barrier()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        step_e2e()
    barrier()
    e2e_s = time.perf_counter() - t0

    # ---- kernel-class breakdown for the roofline (separate instrumented pass, not part of the timed steps)
    reg.profile(True)
    for _ in range(3):
        step_resident()
    prof = reg.profile(False)

    # ---- single-scan tracking latency (config 2)
    one = clouds[:int(offsets[1])]
    h_one = torch.from_numpy(one).pin_memory().numpy()
    lat_k, lat_w = [], []
    for i in range(30):
        t0 = time.perf_counter()
        reg.ScanMatch(h_one, init[0], want_cloud=True)
        lat_w.append((time.perf_counter() - t0) * 1e3)
        lat_k.append(reg.last_timing()[0])
    track = {"scan_points": int(offsets[1]), "kernel_ms": float(np.median(lat_k[5:])), "e2e_ms": float(np.median(lat_w[5:])),
             "points_per_s_e2e": int(offsets[1]) / (np.median(lat_w[5:]) * 1e-3)}

    t_ms = torch.tensor([dev_ms, e2e_s * 1e3, float(n_pts)], dtype=torch.float64, device=dev)
    if world_size > 1:
        tmax = t_ms.clone()
        dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
        tsum = t_ms.clone()
        dist.all_reduce(tsum, op=dist.ReduceOp.SUM)
        dev_ms, e2e_ms, total_pts = float(tmax[0]), float(tmax[1]), float(tsum[2])
    else:
        e2e_ms, total_pts = e2e_s * 1e3, float(n_pts)

    if rank == 0:
        peaks = {}
        try:
            peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        except OSError:
            pass
        peak = float(peaks.get("hbm_gbs", 6650.0))
        peak_src = "measured (MEASURED_PEAKS.json hbm_gbs)" if peaks else "fallback 6650 GB/s (B200_PROFILING.md)"
        nn_ms, nn_launches = prof["search"]
        fit_ms, fit_launches = prof["fit"]
        ring_ms, _ = prof["rings"]
        solve_ms, _ = prof["solve"]
        per_launch_ms = nn_ms / max(nn_launches, 1)
        achieved = BYTES_PER_POINT_ITER * n_pts / (per_launch_ms * 1e-3) / 1e9 if per_launch_ms > 0 else None
        pipe_ms = (nn_ms + fit_ms + ring_ms + solve_ms) / max(nn_launches, 1)  # one whole Gauss-Newton iteration
        traffic = None
        try:  # dram__bytes_read + dram__bytes_write of one k_icp_nn launch of THIS command (tools/profile_remote.sh, ncu --set full)
            tj = json.load(open(os.path.join(ROOT, "profiles", "roofline_traffic.json")))
            cands = [v for k, v in tj.items() if "k_icp_nn<" in k or k.split("@")[0].endswith("k_icp_nn")]
            traffic = max(cands) if cands else None
        except (OSError, ValueError):
            pass
        value = total_pts * args.steps / (dev_ms * 1e-3)
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world_size, "steps": args.steps, "warmup": max(args.warmup, 3),
            "ms_per_step": dev_ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f64", "data": "synthetic", "config": workload_config(args, B),
            "scans_per_s": world_size * B * args.steps / (dev_ms * 1e-3),
            "point_iterations_per_s": value * MAX_ITER,
            "e2e": {"value": total_pts * args.steps / (e2e_ms * 1e-3), "unit": UNIT, "h2d_bytes_per_step": int(h2d),
                    "d2h_bytes_per_step": int(d2h), "scans_per_s": world_size * B * args.steps / (e2e_ms * 1e-3),
                    "api": "locreg_align_batch (host buffers, pinned)"},
            "gpu_launches": int(launches),
            "clocks": clocks,
            "roofline": {"bound": "hbm", "kernel": "k_icp_nn<5> (neighbour search stage 1, %.0f%% of pipeline kernel time)" %
                         (100 * nn_ms / max(nn_ms + fit_ms + ring_ms + solve_ms, 1e-9)),
                         "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak if achieved else None,
                         "traffic": traffic, "peak_source": peak_src,
                         "algorithmic_bytes_per_launch": BYTES_PER_POINT_ITER * n_pts, "launch_ms": per_launch_ms,
                         "fit_kernel_launch_ms": fit_ms / max(fit_launches, 1),
                         "rings_kernel_launch_ms": ring_ms / max(nn_launches, 1),
                         "iteration_ms": pipe_ms,
                         "iteration_achieved": BYTES_PER_POINT_ITER * n_pts / (pipe_ms * 1e-3) / 1e9 if pipe_ms > 0 else None,
                         "iteration_frac": BYTES_PER_POINT_ITER * n_pts / (pipe_ms * 1e-3) / 1e9 / peak if pipe_ms > 0 else None,
                         "note": "96 B per point-iteration (SURVEY 8d) x points per launch; the map (16 MB of points + "
                                 "neighbour lists) is mostly L2/L1 resident: the DRAM traffic (ncu, one steady-state launch) is "
                                 "the scan points plus the per-point search state (seeds, margins), and stays below this"},
            "track": track,
            "map_build": {"wall_ms": map_build_ms, "kernel_ms": map_kernel_ms, "points": int(len(map_cloud))},
            "check": {"median_translation_error_m": float(err), "wall_s_timed_region": wall},
        }
        if not args.no_cpu_baseline and world_size == 1:
            line["cpu_baseline"] = cpu_baseline(map_cloud, clouds, offsets, init, args.cpu_sample_scans, 1)
        print(json.dumps(line))
    if world_size > 1:
        dist.destroy_process_group()


def _dist_setup():
    import torch
    import torch.distributed as dist
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    world_size = int(os.environ.get("WORLD_SIZE", "1"))
    if world_size > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    torch.cuda.set_device(local)
    return rank, local, world_size, torch.device("cuda", local)


def _peak():
    try:
        return float(json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    except (OSError, KeyError, ValueError):
        return 6650.0, "fallback 6650 GB/s (B200_PROFILING.md)"


def run_ndt(args):
    """Config 3: direct NDT (NEARBY6, 1 m voxels, 10 iterations) of a 128-beam scan (~250 k points) against a 20 M-point
    map; one scan in flight, whole AlignNdt loop in one cooperative launch.  Replicas only: every rank runs the same job."""
    import torch
    import torch.distributed as dist
    import loc_lib_b200 as L
    from loc_lib_b200 import synth
    rank, local, world_size, dev = _dist_setup()
    w = synth.World(900.0)
    t0 = time.perf_counter()
    m = w.sample_map(args.ndt_map_points)
    gen_s = time.perf_counter() - t0
    gts = w.poses(4)
    scans = [w.scan(g, beams=128, azimuth=1953, seed=synth.SEED_SCAN + i) for i, g in enumerate(gts)]
    init = synth.perturb_poses(gts)
    reg = L.NdtRegistration(L.NdtOptions(max_iteration_=MAX_ITER, eps_=0.0), device=local)
    t0 = time.perf_counter()
    reg.SetInputTarget(m)
    build_wall = (time.perf_counter() - t0) * 1e3
    build_kernel = reg.last_timing()[0]
    nv = len(reg.Voxels()[3])
    pinned = [torch.from_numpy(s).pin_memory().numpy() for s in scans]
    for i in range(max(args.warmup, 3)):
        reg.ScanMatch(pinned[i % 4], init[i % 4], want_cloud=False)
    torch.cuda.synchronize(dev)
    k_ms, w_ms, pts, hits = [], [], 0, []
    t_all = time.perf_counter()
    for i in range(args.steps):
        t0 = time.perf_counter()
        _, _, pose = reg.ScanMatch(pinned[i % 4], init[i % 4], want_cloud=False)
        w_ms.append((time.perf_counter() - t0) * 1e3)
        k_ms.append(reg.last_timing()[0])
        pts += len(scans[i % 4])
        hits.append(reg.last_result["n_inlier"] / max(reg.last_result["n_effective"], 1))
    wall = time.perf_counter() - t_all
    if rank == 0:
        peak, peak_src = _peak()
        h = float(np.mean(hits))
        bytes_pt = 16 + 7 * 16 + h * 96
        dev_s = sum(k_ms) * 1e-3
        line = {"metric": METRIC, "value": pts / dev_s, "unit": UNIT, "n_gpus": world_size, "steps": args.steps,
                "warmup": max(args.warmup, 3), "ms_per_step": dev_s / args.steps * 1e3, "higher_is_better": True,
                "scaling": "replicas", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
                "config": {"workload": "C3: direct NDT (NEARBY6, voxel 1.0 m), 128x1953-ray synthetic scan vs %d-pt synthetic map, "
                                       "10 Gauss-Newton iterations (eps=0), one cooperative launch" % len(m),
                           "scan_points": int(np.mean([len(s) for s in scans])), "voxels": nv,
                           "timing": "one scan in flight; the %d-voxel table (%.0f MB) exceeds L2" % (nv, nv * 112 / 1e6)},
                "e2e": {"value": pts / wall, "unit": UNIT, "h2d_bytes_per_step": int(scans[0].nbytes), "d2h_bytes_per_step": 7 * 8 + 48,
                        "api": "locreg_align (pinned host scan in, pose out)"},
                "gpu_launches": int(reg.last_timing()[1]) * args.steps,
                "roofline": {"bound": "hbm", "kernel": "k_align_persist<NdtProblem> (10 iterations in one launch)",
                             "achieved": bytes_pt * pts * MAX_ITER / dev_s / 1e9, "peak": peak, "unit": "GB/s",
                             "frac": bytes_pt * pts * MAX_ITER / dev_s / 1e9 / peak, "traffic": None, "peak_source": peak_src,
                             "bytes_per_point_iteration": bytes_pt, "hits_per_point": h},
                "map_build": {"wall_ms": build_wall, "kernel_ms": build_kernel, "points": int(len(m)), "gen_s": gen_s}}
        print(json.dumps(line))
    if world_size > 1:
        dist.destroy_process_group()


def reloc_hypotheses(gt, n_hyp):
    """64 x 64 xy grid (0.5 m pitch, centred on gt) x 16 yaws (22.5 deg), truncated / tiled to n_hyp."""
    g = (np.arange(64) - 31.5) * 0.5
    out = []
    ax, ay, az, aw = gt[:4]
    for k in range(16):
        a = k * np.pi / 8
        bz, bw = np.sin(a / 2), np.cos(a / 2)  # yaw about the sensor's z axis: q = q_gt * (0, 0, bz, bw)
        q = np.array([ax * bw + ay * bz, ay * bw - ax * bz, az * bw + aw * bz, aw * bw - az * bz])
        for y in g:
            for x in g:
                out.append(np.concatenate([q, [gt[4] + x, gt[5] + y, gt[6]]]))
    hyp = np.array(out)
    return hyp[np.resize(np.arange(len(hyp)), n_hyp)] if n_hyp != len(hyp) else hyp


def run_reloc(args):
    """Config 5: global relocalisation, hypotheses sharded over ranks, ONE MIN all-reduce picks the winner."""
    import torch
    import torch.distributed as dist
    import loc_lib_b200 as L
    from loc_lib_b200 import dist as D
    from loc_lib_b200 import synth
    rank, local, world_size, dev = _dist_setup()
    world, map_cloud = make_world(args)
    gt = world.poses(3)[2]
    scan = world.scan(gt)
    hyp = reloc_hypotheses(gt, args.hyp)
    reg = L.IcpRegistration(L.IcpOptions(method_=L.IcpMethod.P2PLANE, max_iteration_=MAX_ITER, eps_=0.0), device=local)
    reg.SetInputTarget(map_cloud)
    warm = hyp[:: max(1, len(hyp) // 256)][:256]
    for _ in range(2):
        reg.Relocalise(scan, warm)

    def barrier():
        if world_size > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)

    barrier()
    times, kernel = [], []
    for _ in range(args.steps):
        barrier()
        t0 = time.perf_counter()
        pose, idx, score = D.relocalise_sharded(reg, scan, hyp, rank, world_size, dev)
        barrier()
        times.append(time.perf_counter() - t0)
        kernel.append(reg.last_timing()[0])
    t = torch.tensor([float(np.mean(times)), float(np.mean(kernel))], dtype=torch.float64, device=dev)
    if world_size > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    if rank == 0:
        step_s, kern_ms = float(t[0]), float(t[1])
        err = float(np.linalg.norm(pose[4:] - gt[4:]))
        line = {"metric": "relocalisation_hypotheses_per_sec", "value": len(hyp) / step_s, "unit": "hypotheses/s",
                "n_gpus": world_size, "steps": args.steps, "warmup": 2, "ms_per_step": step_s * 1e3, "higher_is_better": True,
                "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
                "config": {"workload": "C5: global relocalisation, %d pose hypotheses (64x64 xy grid 0.5 m x 16 yaws) of one 32-beam "
                                       "scan (%d pts) vs 1M-pt map, P2Plane ICP 10 iterations + score pass each, argmin by one "
                                       "NCCL MIN all-reduce" % (len(hyp), len(scan)), "hypotheses": len(hyp)},
                "registered_points_per_s": len(hyp) * len(scan) / step_s, "kernel_ms_max_rank": kern_ms,
                "best": {"index": int(idx), "score": float(score), "translation_error_m": err}}
        print(json.dumps(line))
    if world_size > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    a = parse()
    if a.impl == "reference":
        run_reference(a)
    elif a.workload == "ndt":
        run_ndt(a)
    elif a.workload == "reloc":
        run_reloc(a)
    else:
        run_ours(a)
