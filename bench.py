#!/usr/bin/env python
"""bench.py — front-end frames/s (640x480 RGB-D, ORB 1000 kp + CAPE) on N B200s.

One "step" = one pass of the hot path (ORBextractor::operator() + PlaneDetection_CAPE::
runPlaneDetection / CAPE::process for every frame) over one batch of 256 synthetic 640x480
RGB-D frames per GPU (BASELINE.json configs[2]; frames shard across GPUs as independent
batches, no collective on the data path => weak scaling, configs[3]).

  value : whole-job frames/s with the inputs already resident in HBM, timed with CUDA events
          on the handles' streams, max over ranks.
  e2e   : the same metric through the C-ABI calls a user makes, with HOST (pinned) buffers:
          H2D of that step's gray+depth and D2H of keypoints / descriptors / seg_output /
          planes are inside the timed region.
  roofline : dominant kernel of the step (per-stage CUDA events recorded during the timed
          region, read afterwards), algorithmic bytes per launch / its mean duration vs the
          measured HBM copy peak (MEASURED_PEAKS.json).
  cpu_baseline : the CPU oracle (a dependency-free port of the reference path; the reference
          itself needs OpenCV 3.4 + Eigen and cannot be built here) on a bounded sample of the
          same frames, all host threads, frame-parallel.

`--impl reference` times that CPU port on the same workload and prints the same JSON shape.
torch is used for process plumbing only (torch.distributed barrier / MAX reduce, device and
pinned host buffers); every kernel on the timed path is ours (libdrfe.so).
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time
from concurrent.futures import ThreadPoolExecutor

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(ROOT, "dr-slam_b200"))
sys.path.insert(0, ROOT)

W, H, NFEAT, BATCH = 640, 480, 1000, 256
CELL, MAX_MERGE = 20, 50.0
CYL, UNIT, SCENES = False, 1.0, (0, 1, 2)      # cylinder detection, depth unit scale (1 = metres), scene cycle
MIN_COS = float(np.float32(np.cos(np.pi / 12)))
METRIC = "front-end frames/sec (640x480 RGB-D, ORB 1000 kp + CAPE)"
WORKLOAD = "batched 256-frame synthetic TUM/ICL-shaped RGB-D sequence, ORB 1000 kp + CAPE (20-px cells, cylinders off)"


def select_workload(name):
    """c640 (default) = BASELINE.json configs[2]/[3]; c720 = configs[4], the high-res stress case."""
    global W, H, NFEAT, BATCH, CYL, UNIT, SCENES, WORKLOAD
    if name == "c720":
        W, H, NFEAT, BATCH, CYL, UNIT, SCENES = 1280, 720, 2000, 64, True, 1000.0, (2,)
        WORKLOAD = ("high-res stress: batched 64-frame synthetic 1280x720 RGB-D room-with-pillars sequence (depth in mm), "
                    "ORB 2000 kp, 8 levels + CAPE (20-px cells) with cylinder detection on")


# ---------------------------------------------------------------- byte model
def level_sizes():
    s, out = np.float32(1.0), []
    for l in range(8):
        inv = np.float32(1.0) / s
        out.append((int(np.rint(np.float32(W) * inv)), int(np.rint(np.float32(H) * inv))))
        s = np.float32(float(s) * float(np.float32(1.2)))
    return out


def algorithmic_bytes():
    """Unfused compulsory traffic per frame, split by the stage that owns it."""
    lv = level_sizes()
    P = sum(w * h for w, h in lv)
    P_src = P - lv[-1][0] * lv[-1][1]
    WH, N = W * H, NFEAT
    return {
        "pyramid": WH + P + P_src,              # gray read + pyramid write + resize reads
        "fast": P,                              # FAST reads every level once
        "quadtree": 0,
        "blur": 2 * P,                          # blur read + write
        "orient_describe": 749 * N + 512 * N + 60 * N,
        "cells": 4 * WH + 12 * WH + 12 * WH,    # depth read, cloud write, PlaneSeg read (fused in one kernel)
        "fit": (W // CELL) * (H // CELL) * (48 + 156),   # per-cell sums in, PlaneSeg + tolerance out
        "grid": 0,
        "refine": WH,                           # seg_output write (border-cell re-reads are data dependent)
        "total": WH + P + P_src + P + 2 * P + 1321 * N + 29 * WH,
    }


def measured_peak():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured"
        except Exception:
            pass
    return 6650.0, "fallback"


# ---------------------------------------------------------------- synthetic sequence
def make_sequence(drfe, first, count, threads):
    """frames first..first+count-1 of the synthetic sequence: seed = 20260000 + index, scene by block."""
    def one(i):
        return drfe.synth_frame(W, H, SCENES[(i // 64) % len(SCENES)], 20260000 + i, UNIT)
    with ThreadPoolExecutor(max_workers=threads) as ex:
        data = list(ex.map(one, range(first, first + count)))
    return np.stack([d[0] for d in data]), np.stack([d[1] for d in data]), data[0][2]


# ---------------------------------------------------------------- clocks sampler
class ClockSampler:
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.rows, self.proc = [], None
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(index), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append((time.time(), [x.strip() for x in line.split(",")]))

    def stop(self, t0, t1):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        rows = [r for (t, r) in self.rows if t0 - 0.05 <= t <= t1 + 0.15 and len(r) >= 7] or \
               [r for (_, r) in self.rows if len(r) >= 7]
        if not rows:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["no samples"]}
        sm = sorted(float(r[0]) for r in rows)
        reasons = []
        for i, name in ((3, "hw_slowdown"), (4, "hw_thermal_slowdown"), (5, "sw_thermal_slowdown"), (6, "sw_power_cap")):
            if any(r[i].lower().startswith("active") for r in rows):
                reasons.append(name)
        return {"sm_mhz": sm[len(sm) // 2], "sm_max_mhz": float(rows[0][1]), "reasons": reasons,
                "samples": len(rows), "power_w_max": max(float(r[2]) for r in rows)}


# ---------------------------------------------------------------- CPU port (oracle) timing
def cpu_port_fps(gray, depth, K, nframes, threads, min_seconds=0.0):
    """Times the CPU oracle (test infrastructure, used here ONLY as the reported baseline): passes over the first
    `nframes` frames, frame-parallel, repeated until `min_seconds` of wall time have gone by."""
    from oracle import oracle as orc
    orc.lib()
    nframes = min(nframes, len(gray))
    tl = threading.local()

    def one(i):
        if not hasattr(tl, "o"):
            tl.o = orc.OrbOracle(NFEAT, 1.2, 8, 20, 7)
            tl.c = orc.CapeOracle(H, W, CELL, CELL, CYL, MIN_COS, MAX_MERGE)
        tl.o.run(gray[i])
        cloud = tl.c.depth_to_cloud(depth[i], *K)
        tl.c.process(cloud)
        return 1

    with ThreadPoolExecutor(max_workers=threads) as ex:
        list(ex.map(one, range(min(threads, nframes))))          # warm-up (object creation)
        t0 = time.perf_counter()
        done = 0
        while True:
            list(ex.map(one, range(nframes)))
            done += nframes
            dt = time.perf_counter() - t0
            if dt >= min_seconds:
                break
    return done / dt, dt


def run_reference(args, rank):
    """--impl reference: the reference path's CPU implementation (port) on the host cores."""
    if rank != 0:
        return
    import drfe
    threads = os.cpu_count() or 1
    sample = BATCH                                   # one step = one pass over the same 256-frame batch
    gray, depth, K = make_sequence(drfe, 0, sample, threads)
    for _ in range(args.warmup):
        cpu_port_fps(gray, depth, K, min(sample, threads), threads)
    tot_t, tot_f = 0.0, 0
    for _ in range(args.steps):
        fps, dt = cpu_port_fps(gray, depth, K, sample, threads)
        tot_t += dt
        tot_f += sample
    value = tot_f / tot_t
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": "frames/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * tot_t / args.steps,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u8/f32/f64", "data": "synthetic",
        "config": {"workload": WORKLOAD, "sample_frames_per_step": sample},
        "cpu_baseline": {"value": value, "unit": "frames/s", "cores": threads, "kind": "port",
                         "sample": "%d frames per step of the same synthetic sequence, frame-parallel over %d threads; "
                                   "CPU oracle port of ORBextractor+CAPE (reference needs OpenCV 3.4 + Eigen, unbuildable here)"
                                   % (sample, threads)},
        "e2e": {"value": value, "unit": "frames/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


# ---------------------------------------------------------------- GPU arm
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--cpu-sample", type=int, default=0, help="frames in the cpu_baseline sample (0 = auto)")
    ap.add_argument("--chunk", type=int, default=0, help="frames per chunk of the pipelined e2e batch calls (0 = library default)")
    ap.add_argument("--e2e-threads", type=int, default=1, help="host threads issuing the two e2e batch calls (2: one per extractor; measured slower)")
    ap.add_argument("--workload", default="c640", choices=["c640", "c720"],
                    help="c640: the headline 640x480 / 1000 kp batch (default); c720: 1280x720 / 2000 kp / cylinders on")
    args = ap.parse_args()
    select_workload(args.workload)
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    # the contract is ONE JSON line on stdout: libraries that write to file descriptor 1 from C (NCCL prints its version
    # there under NCCL_DEBUG=VERSION) go to stderr for the rest of the run, the JSON line goes to the real stdout
    sys.stdout.flush()
    real_stdout = os.fdopen(os.dup(1), "w")
    os.dup2(2, 1)
    sys.stdout = real_stdout
    if args.impl == "reference":
        run_reference(args, rank)
        return

    import torch
    import drfe
    if not torch.cuda.is_available() or drfe.device_count() == 0:
        raise SystemExit("bench.py: no CUDA device — the front end has no CPU fallback")
    torch.cuda.set_device(local_rank)
    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))

    threads = max(1, (os.cpu_count() or 8) // max(1, world))
    import shard
    first, last = shard.weak_block(BATCH, rank)                  # this rank's own frames, no data-path collective
    gray, depth, K = make_sequence(drfe, first, last - first, threads)
    orb = drfe.ORBextractor(NFEAT, 1.2, 8, 20, 7, W, H, max_batch=BATCH, device=local_rank)
    cape = drfe.CAPE(H, W, CELL, CELL, CYL, MIN_COS, MAX_MERGE, max_batch=BATCH, device=local_rank)
    s_orb, s_cape = orb.stream(), cape.stream()

    # device-resident inputs (torch tensors are plain device memory here)
    d_gray = torch.from_numpy(gray).cuda()
    d_depth = torch.from_numpy(depth).cuda()
    torch.cuda.synchronize()

    def step_resident():
        orb.enqueue(d_gray.data_ptr(), drfe.MEM_DEVICE, BATCH, W, W * H)
        cape.enqueue_depth(d_depth.data_ptr(), *K, mem_kind=drfe.MEM_DEVICE, nframes=BATCH, row_stride=W, frame_stride=W * H)

    def barrier():
        orb.sync(); cape.sync()
        torch.cuda.synchronize()
        if dist:
            dist.barrier()

    for _ in range(args.warmup):
        step_resident()
    barrier()
    # sanity: the warm-up produced real results
    assert orb.download()[2].min() > 0 and cape.download()[2].min() > 0

    orb.set_profiling(True); cape.set_profiling(True)
    ev0, ev_orb, ev_cape = drfe.Event(), drfe.Event(), drfe.Event()
    step_evs = [drfe.Event() for _ in range(args.steps)]
    sampler = ClockSampler(local_rank)
    time.sleep(0.25)
    barrier()
    launches0 = drfe.kernel_launch_count()
    t_wall0 = time.time()
    ev0.record(s_orb)
    drfe.stream_wait_event(s_cape, ev0)
    for i in range(args.steps):
        step_resident()
        step_evs[i].record(s_orb)                                  # end of this step's ORB chain (the longer of the two)
    ev_orb.record(s_orb); ev_cape.record(s_cape)
    barrier()
    t_wall1 = time.time()
    step_ms = np.diff([0.0] + [ev0.elapsed_ms(e) for e in step_evs])
    launches = drfe.kernel_launch_count() - launches0
    ms = max(ev0.elapsed_ms(ev_orb), ev0.elapsed_ms(ev_cape))
    clocks = sampler.stop(t_wall0, t_wall1)
    stages = dict(orb.stage_times())
    stages.update(dict(cape.stage_times()))
    orb.set_profiling(False); cape.set_profiling(False)
    if dist:
        t = torch.tensor([ms], device="cuda", dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
    value = world * BATCH * args.steps / (ms * 1e-3)

    # ---- e2e: host (pinned) buffers through the public C-ABI calls, copies inside the timed region
    pin = lambda a: torch.from_numpy(a).pin_memory().numpy()  # noqa: E731
    h_gray, h_depth = pin(gray), pin(depth)
    cap = orb.cap
    PLANE_CAP = 64
    h_kps = torch.empty((BATCH, cap * 28), dtype=torch.uint8).pin_memory().numpy().view(drfe.KP_DTYPE).reshape(BATCH, cap)
    h_desc = torch.empty((BATCH, cap, 32), dtype=torch.uint8).pin_memory().numpy()
    h_cnt = torch.empty(BATCH, dtype=torch.int32).pin_memory().numpy()
    h_seg = torch.empty((BATCH, H, W), dtype=torch.uint8).pin_memory().numpy()
    h_planes = torch.empty((BATCH, PLANE_CAP * drfe.PLANE_DTYPE.itemsize), dtype=torch.uint8).pin_memory().numpy() \
        .view(drfe.PLANE_DTYPE).reshape(BATCH, PLANE_CAP)
    h_npl = torch.empty(BATCH, dtype=torch.int32).pin_memory().numpy()

    # the chunk-pipelined batch calls: per 32-frame chunk H2D | kernels | D2H on three streams per handle
    # One host thread issues both calls back to back.  Issuing them from two threads (one per extractor, --e2e-threads 2)
    # was measured slower on the B200 box (47 k -> 35 k frames/s): the two handles' H2D copies then interleave on the copy
    # engine and the longer ORB chain gets its first chunks later.
    from concurrent.futures import ThreadPoolExecutor
    pool = ThreadPoolExecutor(max_workers=1) if args.e2e_threads > 1 else None

    def step_e2e(dep, fac):
        if pool is not None:
            fut = pool.submit(orb.extract_batch, h_gray, h_kps, h_desc, h_cnt, args.chunk)
            cape.process_depth_batch(dep, *K, depth_factor=fac, seg=h_seg, planes=h_planes, nplanes=h_npl, chunk_frames=args.chunk)
            fut.result()
        else:
            orb.extract_batch(h_gray, h_kps, h_desc, h_cnt, chunk_frames=args.chunk)
            cape.process_depth_batch(dep, *K, depth_factor=fac, seg=h_seg, planes=h_planes, nplanes=h_npl, chunk_frames=args.chunk)
        orb.finish_batch()
        cape.finish_batch()

    def time_e2e(dep, fac):
        for _ in range(2):
            step_e2e(dep, fac)
        barrier()
        t0 = time.perf_counter()
        for _ in range(e2e_steps):
            step_e2e(dep, fac)
        barrier()
        dt = time.perf_counter() - t0
        if dist:
            t = torch.tensor([dt], device="cuda", dtype=torch.float64)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            dt = float(t.item())
        return world * BATCH * e2e_steps / dt

    e2e_steps = max(3, min(args.steps, 10))
    e2e_value = time_e2e(h_depth, 1.0)
    # the same with the sensor's raw 16-bit depth (TUM png, factor 1/5000) converted on the device (Frame.cc:113-115)
    fac = np.float32(np.float32(1.0 / 5000.0) * np.float32(UNIT))
    q16 = np.rint(depth / np.float32(UNIT) * 5000).astype(np.uint16)
    if UNIT != 1.0:     # the generator scales metres by UNIT after quantising; one factor reproduces it only approximately
        depth = q16.astype(np.float32) * fac
        h_depth[...] = depth
    assert np.array_equal(q16.astype(np.float32) * fac, depth)
    h_depth16 = pin(q16)
    e2e_u16 = time_e2e(h_depth16, float(fac))
    # ---- single-frame latency, the reference's call shape (BASELINE configs[1]): one 640x480 frame from host memory
    # through ORBextractor::operator() / PlaneDetection_CAPE (drfe_orb_extract, drfe_cape_process_depth), results on the host
    def median_ms(fn, n=100):
        for i in range(10):
            fn(i)
        ts = []
        for i in range(n):
            t0 = time.perf_counter()
            fn(i)
            ts.append(time.perf_counter() - t0)
        return float(np.median(ts) * 1e3)

    single = None
    if rank == 0:
        ex1 = drfe.ORBextractor(NFEAT, 1.2, 8, 20, 7, W, H, device=local_rank)
        cp1 = drfe.CAPE(H, W, CELL, CELL, CYL, MIN_COS, MAX_MERGE, device=local_rank)
        single = {"orb_extract_ms": median_ms(lambda i: ex1(gray[i % BATCH], None)),
                  "cape_process_depth_ms": median_ms(lambda i: cp1.process_depth(depth[i % BATCH], *K)),
                  "note": "median of 100 host-in/host-out single-frame calls (drfe_orb_extract, drfe_cape_process_depth)"}
        del ex1, cp1
    h2d = int(h_gray.nbytes + h_depth.nbytes)
    d2h = int(h_kps.nbytes + h_desc.nbytes + h_cnt.nbytes + h_seg.nbytes + h_planes.nbytes + h_npl.nbytes)

    if rank != 0:
        if dist:
            dist.destroy_process_group()
        return

    # ---- roofline of the dominant kernel
    alg = algorithmic_bytes()
    kernel_stages = {k: v for k, v in stages.items() if k in alg and k != "total"}
    dom = max(kernel_stages, key=kernel_stages.get)
    peak, peak_src = measured_peak()
    launches_per_stage = {"pyramid": 8}
    dom_ms = kernel_stages[dom]
    per_launch_ms = dom_ms / launches_per_stage.get(dom, 1)
    bytes_per_launch = alg[dom] * BATCH / launches_per_stage.get(dom, 1)
    achieved = bytes_per_launch / (per_launch_ms * 1e-3) / 1e9 if per_launch_ms > 0 else 0.0
    # DRAM bytes of the same kernel from the committed ncu --set full capture (profiles/), per launch of 256 frames
    traffic = None
    try:
        if args.workload == "c640":
            traffic = json.load(open(os.path.join(ROOT, "profiles", "r01_traffic.json")))["stages"][dom]["dram_bytes"]
    except Exception:
        traffic = None
    alu = None
    try:
        if args.workload == "c640" and dom == "fast":
            alu = json.load(open(os.path.join(ROOT, "profiles", "r01_traffic.json")))["stages"][dom].get("alu_pipe_pct")
    except Exception:
        alu = None
    roofline = {"bound": "hbm", "kernel": dom, "issue_bound": {"pipe": "alu (half-rate: 0.5 warp-inst/clk/SMSP)", "pipe_util_pct": alu,
                                                                "source": "ncu sm__inst_executed_pipe_alu, profiles/r01_traffic.json"}, "achieved": achieved, "peak": peak, "unit": "GB/s",
                "frac": achieved / peak, "traffic": traffic, "peak_source": peak_src,
                "note": "k_fast_strips is bound by the half-rate integer ALU pipe, not by HBM (ncu: alu pipe ~70 %, dram 4 %, "
                        "profiles/r01_*_full.txt; tools/ubench/pipes.cu); the HBM fraction is reported as the contract asks"
                        if dom == "fast" else None,
                "alg_bytes_per_launch": bytes_per_launch, "launch_ms": per_launch_ms,
                "whole_step": {"alg_bytes_per_frame": alg["total"],
                               "achieved": alg["total"] * BATCH * args.steps / (ms * 1e-3) / 1e9,
                               "frac": alg["total"] * BATCH * args.steps / (ms * 1e-3) / 1e9 / peak},
                "stage_ms": {k: round(v, 4) for k, v in stages.items()}}

    # ---- CPU baseline on a bounded sample (rank 0, N=1 only)
    cpu = None
    if world == 1:
        cores = os.cpu_count() or 1
        sample = args.cpu_sample or BATCH
        fps, dt = cpu_port_fps(gray, depth, K, sample, cores, min_seconds=10.0)
        cpu = {"value": fps, "unit": "frames/s", "cores": cores, "kind": "port",
               "sample": "passes over the first %d frames of the same batch, frame-parallel on %d threads, for %.1f s "
                         "(%d frames); CPU oracle port of ORBextractor+CAPE, -O3 x86-64-v3"
                         % (sample, cores, dt, int(round(fps * dt)))}

    line = {
        "metric": METRIC, "value": value, "unit": "frames/s", "n_gpus": world, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": ms / args.steps,
        "ms_per_step_percentiles": {"p10": float(np.percentile(step_ms, 10)), "p50": float(np.percentile(step_ms, 50)),
                                    "p90": float(np.percentile(step_ms, 90)),
                                    "note": "per step on the ORB stream (CUDA events), rank 0; the CAPE stream runs ahead"},
        "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "u8/f32/f64", "data": "synthetic",
        "config": {"workload": WORKLOAD, "frames_per_gpu_per_step": BATCH, "width": W, "height": H,
                   "nfeatures": NFEAT, "nlevels": 8, "scale_factor": 1.2, "fast": [20, 7], "cape_cell": CELL,
                   "cylinder_detection": CYL,
                   "l2": "inputs larger than L2 (%d MB of gray+depth per step per GPU, no flush needed)" % (BATCH * W * H * 5 // 1000000),
                   "sharding": "independent 256-frame batches per GPU, no collective"},
        # headline e2e: what Frame::Frame receives — the gray image and the sensor's raw 16-bit depth (imDepth is
        # converted to float INSIDE the path, Frame.cc:113-115); the float-depth variant of the same call is kept beside it
        "e2e": {"value": e2e_u16, "unit": "frames/s", "h2d_bytes_per_step": int(h_gray.nbytes + h_depth16.nbytes),
                "d2h_bytes_per_step": d2h, "steps": e2e_steps,
                "api": "drfe_orb_extract_batch + drfe_cape_process_depth_batch (gray u8 + raw u16 depth as Frame::Frame gets "
                       "them, depth scaled on the device as Frame.cc:113-115 does; pinned host buffers, 32-frame chunks "
                       "(8/16-frame chunks at both ends) pipelined H2D | kernels | D2H; the two calls issued from %d host "
                       "thread(s))" % (2 if pool is not None else 1),
                "float_depth": {"value": e2e_value, "unit": "frames/s", "h2d_bytes_per_step": h2d,
                                "note": "same call fed float depth (what PlaneDetection_CAPE::readDepthImage takes); "
                                        "H2D-bound: 393 MB per step over PCIe"}},
        "single_frame": single,
        "gpu_launches": int(launches),
        "clocks": clocks,
        "roofline": roofline,
        "cpu_baseline": cpu,
    }
    print(json.dumps(line), flush=True)
    if dist:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
