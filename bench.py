#!/usr/bin/env python
"""bench.py — front-end frames/s (640x480 RGB-D, ORB 1000 kp + CAPE) on N B200s.

One "step" = one pass of the hot path (ORBextractor::operator() + PlaneDetection_CAPE::
runPlaneDetection / CAPE::process for every frame) over one batch of 256 synthetic 640x480
RGB-D frames per GPU (BASELINE.json configs[2]; frames shard across GPUs as independent
batches, no collective on the data path => weak scaling).

  value : whole-job frames/s with the inputs already resident in HBM, timed with CUDA events
          on the handles' streams, max over ranks.
  e2e   : the same metric through the C-ABI calls a user makes, with HOST (pinned) buffers:
          H2D of that step's gray+depth and D2H of keypoints / descriptors / seg_output /
          planes are inside the timed region.  `e2e.pcie` is the host link's ceiling for exactly
          those bytes, measured in the same run (all ranks copying at once), `e2e.pcie_frac`
          = e2e / that ceiling; `e2e.pool` is the same batch through drfe_pool_extract_batch
          (ONE process, one host thread per device) — at N > 1 measured by rank 0 over all N
          devices while the other ranks wait.
  strong : BASELINE.json configs[3] as written — the SAME 256-frame sequence cut into 256/N
          frames per GPU (the headline `value` is weak scaling: 256 frames per GPU).
  c720  : BASELINE.json configs[4] (1280x720, 2000 kp, cylinders on; 64 frames per GPU),
          measured after the headline when N = 8 (or with --with-c720).
  roofline : dominant kernel of the step (per-stage CUDA events recorded during the timed
          region, read afterwards), algorithmic bytes per launch / its mean duration vs the
          measured HBM copy peak (MEASURED_PEAKS.json); `bound` says what ncu shows.
  cpu_baseline : the CPU oracle (a dependency-free port of the reference path; the reference
          itself needs OpenCV 3.4 + Eigen and cannot be built here) on a bounded sample of the
          same frames: all host threads frame-parallel (`value`), the reference's own per-frame
          arrangement of one ORB thread beside one CAPE thread (`reference_shaped`), and the
          OpenCV (cv2, IPP/AVX-512) stage times of the ORB primitives (`cv2_stage_ms`).

`--impl reference` times that CPU port on the same workload and prints the same JSON shape.
torch is used for process plumbing only (torch.distributed barrier / MAX reduce, device and
pinned host buffers); every kernel on the timed path is ours (libdrfe.so).
"""
import argparse
import importlib.util
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

METRIC = "front-end frames/sec (640x480 RGB-D, ORB 1000 kp + CAPE)"
MIN_COS = float(np.float32(np.cos(np.pi / 12)))
TRAFFIC_JSON = os.path.join("profiles", "r02_traffic.json")     # ncu --set full capture of the same step (tools/ncu_traffic.py)


class Workload:
    """c640 = BASELINE.json configs[2]/[3] (the headline); c720 = configs[4], the high-res stress case."""

    def __init__(self, name):
        self.name = name
        self.cell, self.max_merge = 20, 50.0
        if name == "c720":
            self.W, self.H, self.nfeat, self.batch, self.cyl, self.unit, self.scenes = 1280, 720, 2000, 64, True, 1000.0, (2,)
            self.text = ("high-res stress: batched 64-frame synthetic 1280x720 RGB-D room-with-pillars sequence (depth in mm), "
                         "ORB 2000 kp, 8 levels + CAPE (20-px cells) with cylinder detection on")
        else:
            self.W, self.H, self.nfeat, self.batch, self.cyl, self.unit, self.scenes = 640, 480, 1000, 256, False, 1.0, (0, 1, 2)
            self.text = "batched 256-frame synthetic TUM/ICL-shaped RGB-D sequence, ORB 1000 kp + CAPE (20-px cells, cylinders off)"

    def config(self):
        """the `config` object of the JSON line — identical in the GPU arm and in --impl reference"""
        return {"workload": self.text, "frames_per_gpu_per_step": self.batch, "width": self.W, "height": self.H,
                "nfeatures": self.nfeat, "nlevels": 8, "scale_factor": 1.2, "fast": [20, 7], "cape_cell": self.cell,
                "cylinder_detection": self.cyl,
                "l2": "inputs larger than L2 (%d MB of gray+depth per step per GPU, no flush needed)" % (self.batch * self.W * self.H * 5 // 1000000),
                "sharding": "independent %d-frame batches per GPU, no collective" % self.batch}

    # ---- byte model
    def level_sizes(self):
        s, out = np.float32(1.0), []
        for _ in range(8):
            inv = np.float32(1.0) / s
            out.append((int(np.rint(np.float32(self.W) * inv)), int(np.rint(np.float32(self.H) * inv))))
            s = np.float32(float(s) * float(np.float32(1.2)))
        return out

    def algorithmic_bytes(self):
        """Unfused compulsory traffic per frame, split by the stage that owns it."""
        lv = self.level_sizes()
        P = sum(w * h for w, h in lv)
        P_src = P - lv[-1][0] * lv[-1][1]
        WH, N = self.W * self.H, self.nfeat
        return {
            "pyramid": WH + P + P_src,              # gray read + pyramid write + resize reads
            "fast": P,                              # FAST reads every level once
            "quadtree": 0,
            "blur": 2 * P,                          # blur read + write
            "orient_describe": 749 * N + 512 * N + 60 * N,
            "cells": 4 * WH + 12 * WH + 12 * WH,    # depth read, cloud write, PlaneSeg read (the SURVEY's unfused model; the kernel keeps the cloud in registers)
            "fit": (self.W // self.cell) * (self.H // self.cell) * (48 + 156),   # per-cell sums in, PlaneSeg + tolerance out
            "grid": 0,
            "refine": WH,                           # seg_output write (border-cell re-reads are data dependent)
            "total": WH + P + P_src + P + 2 * P + 1321 * N + 29 * WH,
        }


def measured_peak():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md)"


# ---------------------------------------------------------------- synthetic sequence
_synth = None


def synth():
    """tools/synth: the input generator (its own small host library — not the product, not the oracle)"""
    global _synth
    if _synth is None:
        spec = importlib.util.spec_from_file_location("drfe_synth", os.path.join(ROOT, "tools", "synth", "synth.py"))
        _synth = importlib.util.module_from_spec(spec)
        spec.loader.exec_module(_synth)
    return _synth


def make_sequence(wl, first, count, threads):
    """frames first..first+count-1 of the synthetic sequence: seed = 20260000 + index, scene by block of 64."""
    gen = synth()

    def one(i):
        return gen.synth_frame(wl.W, wl.H, wl.scenes[(i // 64) % len(wl.scenes)], 20260000 + i, wl.unit)
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
def cpu_port_fps(wl, gray, depth, K, nframes, threads, min_seconds=0.0):
    """Times the CPU oracle (test infrastructure, used here ONLY as the reported baseline): passes over the first
    `nframes` frames, frame-parallel, repeated until `min_seconds` of wall time have gone by."""
    from oracle import oracle as orc
    orc.lib()
    nframes = min(nframes, len(gray))
    tl = threading.local()

    def one(i):
        if not hasattr(tl, "o"):
            tl.o = orc.OrbOracle(wl.nfeat, 1.2, 8, 20, 7)
            tl.c = orc.CapeOracle(wl.H, wl.W, wl.cell, wl.cell, wl.cyl, MIN_COS, wl.max_merge)
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


def cpu_reference_shaped(wl, gray, depth, K, nframes):
    """The reference's own per-frame arrangement (Frame.cc:124-134): one thread runs ORBextractor::operator(), a second one
    the CAPE wrapper, both are joined before the next frame starts -> ms per frame on 2 cores."""
    from oracle import oracle as orc
    o = orc.OrbOracle(wl.nfeat, 1.2, 8, 20, 7)
    c = orc.CapeOracle(wl.H, wl.W, wl.cell, wl.cell, wl.cyl, MIN_COS, wl.max_merge)
    n = min(nframes, len(gray))
    t_orb, t_cape = [], []

    def f_orb(i):
        t = time.perf_counter(); o.run(gray[i]); t_orb.append(time.perf_counter() - t)

    def f_cape(i):
        t = time.perf_counter(); c.process(c.depth_to_cloud(depth[i], *K)); t_cape.append(time.perf_counter() - t)
    with ThreadPoolExecutor(max_workers=2) as ex:
        for i in range(2):
            a, b = ex.submit(f_orb, i), ex.submit(f_cape, i); a.result(); b.result()
        t_orb.clear(); t_cape.clear()
        t0 = time.perf_counter()
        for i in range(n):
            a, b = ex.submit(f_orb, i), ex.submit(f_cape, i)
            a.result(); b.result()
        dt = time.perf_counter() - t0
    return {"ms_per_frame": 1e3 * dt / n, "fps": n / dt, "cores": 2, "frames": n,
            "orb_thread_ms": 1e3 * float(np.mean(t_orb)), "cape_thread_ms": 1e3 * float(np.mean(t_cape)),
            "note": "1 thread ORBextractor::operator() beside 1 thread CAPE per frame, joined per frame (Frame.cc:124-134); CPU oracle port"}


def cv2_stage_ms(wl, gray, nframes=8):
    """OpenCV's own (IPP / AVX-512) time for the ORB primitives the reference calls, one thread: the 8-level pyramid
    (cv::resize chain + copyMakeBorder), the per-cell cv::FAST calls with the 20 -> 7 fallback, the 8 GaussianBlurs.
    Cross-check of the scalar port against the library the reference really links."""
    try:
        import cv2
        from oracle import py_ref
    except Exception as e:                                       # cv2 not importable on this box
        return {"unavailable": str(e)}
    cv2.setNumThreads(1)
    ref = py_ref.OrbRef(wl.nfeat, 1.2, 8, 20, 7)
    det_ini, det_min = cv2.FastFeatureDetector_create(20, True), cv2.FastFeatureDetector_create(7, True)
    t_pyr = t_fast = t_blur = 0.0
    ncells = ncand = 0
    n = min(nframes, len(gray))
    for i in range(n):
        t = time.perf_counter()
        levels, _ = ref.pyramid(gray[i])
        t_pyr += time.perf_counter() - t
        t = time.perf_counter()
        for img in levels:
            h, w = img.shape
            bx, by = w - 32, h - 32
            nc, nr = int(np.float32(bx) / np.float32(30)), int(np.float32(by) / np.float32(30))
            wc, hc = int(np.ceil(bx / nc)), int(np.ceil(by / nr))
            for r in range(nr):
                y0 = 16 + r * hc
                if y0 >= h - 16 - 3:
                    continue
                y1 = min(y0 + hc + 6, h - 16)
                for c in range(nc):
                    x0 = 16 + c * wc
                    if x0 >= w - 16 - 6:
                        continue
                    x1 = min(x0 + wc + 6, w - 16)
                    cell = img[y0:y1, x0:x1]
                    kps = det_ini.detect(cell) or det_min.detect(cell)
                    ncells += 1
                    ncand += len(kps)
        t_fast += time.perf_counter() - t
        t = time.perf_counter()
        for img in levels:
            cv2.GaussianBlur(img, (7, 7), 2, sigmaY=2, borderType=cv2.BORDER_REFLECT_101)
        t_blur += time.perf_counter() - t
    return {"pyramid": 1e3 * t_pyr / n, "fast_cells": 1e3 * t_fast / n, "blur": 1e3 * t_blur / n, "frames": n, "threads": 1,
            "fast_calls_per_frame": ncells / n, "candidates_per_frame": ncand / n,
            "note": "cv2 %s from Python, cv2.setNumThreads(1); fast_cells includes the Python call overhead of ~%d cv::FAST calls per frame"
                    % (cv2.__version__, ncells // max(n, 1))}


def run_reference(args, wl, rank):
    """--impl reference: the reference path's CPU implementation (port) on the host cores."""
    if rank != 0:
        return
    threads = os.cpu_count() or 1
    sample = wl.batch                                # one step = one pass over the same batch
    gray, depth, K = make_sequence(wl, 0, sample, threads)
    for _ in range(args.warmup):
        cpu_port_fps(wl, gray, depth, K, min(sample, threads), threads)
    tot_t, tot_f = 0.0, 0
    for _ in range(args.steps):
        fps, dt = cpu_port_fps(wl, gray, depth, K, sample, threads)
        tot_t += dt
        tot_f += sample
    value = tot_f / tot_t
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": "frames/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * tot_t / args.steps,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u8/f32/f64", "data": "synthetic",
        "config": wl.config(),
        "cpu_baseline": {"value": value, "unit": "frames/s", "cores": threads, "kind": "port",
                         "sample": "%d frames per step of the same synthetic sequence, frame-parallel over %d threads; "
                                   "CPU oracle port of ORBextractor+CAPE (reference needs OpenCV 3.4 + Eigen, unbuildable here)"
                                   % (sample, threads)},
        "e2e": {"value": value, "unit": "frames/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


# ---------------------------------------------------------------- GPU arm
class Rig:
    """the handles, device-resident inputs and pinned host buffers of one workload on this rank's GPU"""

    def __init__(self, drfe, torch, wl, gray, depth, K, device):
        self.drfe, self.torch, self.wl, self.K = drfe, torch, wl, K
        B, W, H = wl.batch, wl.W, wl.H
        self.orb = drfe.ORBextractor(wl.nfeat, 1.2, 8, 20, 7, W, H, max_batch=B, device=device)
        self.cape = drfe.CAPE(H, W, wl.cell, wl.cell, wl.cyl, MIN_COS, wl.max_merge, max_batch=B, device=device)
        self.gray, self.depth = gray, depth
        self.d_gray = torch.from_numpy(gray).cuda()
        self.d_depth = torch.from_numpy(depth).cuda()
        torch.cuda.synchronize()

    def step_resident(self, n=None):
        wl = self.wl
        n = n or wl.batch
        self.orb.enqueue(self.d_gray.data_ptr(), self.drfe.MEM_DEVICE, n, wl.W, wl.W * wl.H)
        self.cape.enqueue_depth(self.d_depth.data_ptr(), *self.K, mem_kind=self.drfe.MEM_DEVICE, nframes=n, row_stride=wl.W,
                                frame_stride=wl.W * wl.H)

    def sync(self):
        self.orb.sync(); self.cape.sync()

    def timed_resident(self, steps, barrier, n=None):
        """device time of `steps` resident steps of n frames: CUDA events on both handle streams, the later end counts"""
        drfe = self.drfe
        s_orb, s_cape = self.orb.stream(), self.cape.stream()
        ev0, ev_orb, ev_cape = drfe.Event(), drfe.Event(), drfe.Event()
        step_evs = [(drfe.Event(), drfe.Event()) for _ in range(steps)]
        barrier()
        ev0.record(s_orb)
        drfe.stream_wait_event(s_cape, ev0)
        for i in range(steps):
            self.step_resident(n)
            step_evs[i][0].record(s_orb); step_evs[i][1].record(s_cape)
        ev_orb.record(s_orb); ev_cape.record(s_cape)
        barrier()
        ends = [max(ev0.elapsed_ms(a), ev0.elapsed_ms(b)) for a, b in step_evs]
        return max(ev0.elapsed_ms(ev_orb), ev0.elapsed_ms(ev_cape)), np.diff([0.0] + ends)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--cpu-sample", type=int, default=0, help="frames in the cpu_baseline sample (0 = auto)")
    ap.add_argument("--chunk", type=int, default=0, help="frames per chunk of the pipelined e2e batch calls (0 = library default)")
    ap.add_argument("--workload", default="c640", choices=["c640", "c720"],
                    help="c640: the headline 640x480 / 1000 kp batch (default); c720: 1280x720 / 2000 kp / cylinders on")
    ap.add_argument("--with-c720", action="store_true", help="also measure configs[4] as the `c720` key (default: only at 8 GPUs)")
    ap.add_argument("--no-extras", action="store_true", help="headline numbers only: skip strong scaling, pool, c720 and the CPU legs")
    args = ap.parse_args()
    wl = Workload(args.workload)
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
        run_reference(args, wl, rank)
        return

    import torch
    import drfe
    if not torch.cuda.is_available() or drfe.device_count() == 0:
        raise SystemExit("bench.py: no CUDA device — the front end has no CPU fallback")
    torch.cuda.set_device(local_rank)
    dist = cpu_group = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
        cpu_group = dist.new_group(backend="gloo")                # host-side barrier for the legs where only rank 0 uses the GPUs

    ncpu = os.cpu_count() or 8
    threads = max(1, ncpu // max(1, world))
    # one rank per GPU: spread the ranks' host threads over the cores instead of letting all of them start on the same ones
    try:
        cores = sorted(os.sched_getaffinity(0))
        per = max(1, len(cores) // world)
        os.sched_setaffinity(0, set(cores[local_rank * per:(local_rank + 1) * per]) or set(cores))
    except Exception:
        pass
    import shard
    B, W, H = wl.batch, wl.W, wl.H
    first, last = shard.weak_block(B, rank)                      # this rank's own frames, no data-path collective
    gray, depth, K = make_sequence(wl, first, last - first, threads)
    rig = Rig(drfe, torch, wl, gray, depth, K, local_rank)
    orb, cape = rig.orb, rig.cape

    def barrier():
        rig.sync()
        torch.cuda.synchronize()
        if dist:
            dist.barrier()

    def max_over_ranks(x):
        if not dist:
            return x
        t = torch.tensor([x], device="cuda", dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    for _ in range(args.warmup):
        rig.step_resident()
    barrier()
    # sanity: the warm-up produced real results
    assert orb.download()[2].min() > 0 and cape.download()[2].min() > 0

    orb.set_profiling(True); cape.set_profiling(True)
    sampler = ClockSampler(local_rank)
    time.sleep(0.25)
    launches0 = drfe.kernel_launch_count()
    t_wall0 = time.time()
    ms, step_ms = rig.timed_resident(args.steps, barrier)
    t_wall1 = time.time()
    launches = drfe.kernel_launch_count() - launches0
    clocks = sampler.stop(t_wall0, t_wall1)
    stages_in_step = dict(orb.stage_times())
    stages_in_step.update(dict(cape.stage_times()))
    orb.set_profiling(False); cape.set_profiling(False)
    ms = max_over_ranks(ms)
    value = world * B * args.steps / (ms * 1e-3)
    # per-kernel durations for the roofline: the same launches with ONE handle running at a time.  Inside the two-stream step a stage's
    # event interval also contains the time its CTAs wait for SM slots behind the other handle's kernels (the plane stream has the
    # lower priority: its 0.05 ms fit stage shows 0.7 ms there), so it is not a kernel duration.
    barrier()
    orb.set_profiling(True)
    for _ in range(5):
        orb.enqueue(rig.d_gray.data_ptr(), drfe.MEM_DEVICE, B, W, W * H)
    orb.sync()
    stages = dict(orb.stage_times())
    orb.set_profiling(False)
    cape.set_profiling(True)
    for _ in range(5):
        cape.enqueue_depth(rig.d_depth.data_ptr(), *K, mem_kind=drfe.MEM_DEVICE, nframes=B, row_stride=W, frame_stride=W * H)
    cape.sync()
    stages.update(dict(cape.stage_times()))
    cape.set_profiling(False)

    # ---- strong scaling (configs[3] as written): the same B-frame sequence, B / world frames per GPU
    strong = None
    if not args.no_extras:
        a, b = shard.frame_block(B, rank, world)
        n_strong = b - a                                           # (the content of the frames does not matter for the time: this rank's first n)
        if world == 1:
            strong = {"value": value, "frames_total": B, "frames_per_gpu": B, "ms_per_step": ms / args.steps, "note": "N = 1: identical to `value`"}
        else:
            for _ in range(3):
                rig.step_resident(n_strong)
            ms_s, _ = rig.timed_resident(args.steps, barrier, n_strong)
            ms_s = max_over_ranks(ms_s)
            strong = {"value": B * args.steps / (ms_s * 1e-3), "frames_total": B, "frames_per_gpu": n_strong, "ms_per_step": ms_s / args.steps,
                      "note": "BASELINE configs[3]: the same %d-frame sequence cut into contiguous blocks of %d frames per GPU "
                              "(device-resident, CUDA events, max over ranks); the headline `value` is weak scaling" % (B, n_strong)}

    # ---- e2e: host (pinned) buffers through the public C-ABI calls, copies inside the timed region
    pin = lambda a: torch.from_numpy(a).pin_memory().numpy()  # noqa: E731
    h_gray, h_depth = pin(gray), pin(depth)
    cap = orb.cap
    PLANE_CAP = 64
    h_kps = torch.empty((B, cap * 28), dtype=torch.uint8).pin_memory().numpy().view(drfe.KP_DTYPE).reshape(B, cap)
    h_desc = torch.empty((B, cap, 32), dtype=torch.uint8).pin_memory().numpy()
    h_cnt = torch.empty(B, dtype=torch.int32).pin_memory().numpy()
    h_seg = torch.empty((B, H, W), dtype=torch.uint8).pin_memory().numpy()
    h_planes = torch.empty((B, PLANE_CAP * drfe.PLANE_DTYPE.itemsize), dtype=torch.uint8).pin_memory().numpy() \
        .view(drfe.PLANE_DTYPE).reshape(B, PLANE_CAP)
    h_npl = torch.empty(B, dtype=torch.int32).pin_memory().numpy()

    # the chunk-pipelined batch calls: per 32-frame chunk H2D | kernels | D2H on three streams per handle, both calls issued by
    # one host thread back to back (two threads were measured slower: the handles' H2D copies interleave on the copy engine)
    def step_e2e(dep, fac):
        orb.extract_batch(h_gray, h_kps, h_desc, h_cnt, chunk_frames=args.chunk)
        cape.process_depth_batch(dep, *K, depth_factor=fac, seg=h_seg, planes=h_planes, nplanes=h_npl, chunk_frames=args.chunk)
        orb.finish_batch()
        cape.finish_batch()

    e2e_steps = max(3, min(2 * args.steps, 40))                    # ~5 ms each: 40 of them, so that one slow copy does not move the figure

    def time_loop(fn):
        for _ in range(2):
            fn()
        barrier()
        t0 = time.perf_counter()
        for _ in range(e2e_steps):
            fn()
        barrier()
        return max_over_ranks(time.perf_counter() - t0)

    e2e_value = world * B * e2e_steps / time_loop(lambda: step_e2e(h_depth, 1.0))
    # the same with the sensor's raw 16-bit depth (TUM png, factor 1/5000) converted on the device (Frame.cc:113-115)
    fac = np.float32(np.float32(1.0 / 5000.0) * np.float32(wl.unit))
    q16 = np.rint(depth / np.float32(wl.unit) * 5000).astype(np.uint16)
    if wl.unit != 1.0:     # the generator scales metres by the unit after quantising; one factor reproduces it only approximately
        depth = q16.astype(np.float32) * fac
        h_depth[...] = depth
    assert np.array_equal(q16.astype(np.float32) * fac, depth)
    h_depth16 = pin(q16)
    e2e_u16 = world * B * e2e_steps / time_loop(lambda: step_e2e(h_depth16, float(fac)))
    # ---- the same calls with TWO batches in flight: a second handle pair and a second set of result buffers, used alternately
    # (step s is issued before step s-1 is finished), so that the copy-in of a batch starts while the previous batch's last
    # chunks are still in their kernels and copying out.  Every step still moves all of its inputs and results.
    e2e_two = None
    if not args.no_extras:
        orb_b = drfe.ORBextractor(wl.nfeat, 1.2, 8, 20, 7, W, H, max_batch=B, device=local_rank)
        cape_b = drfe.CAPE(H, W, wl.cell, wl.cell, wl.cyl, MIN_COS, wl.max_merge, max_batch=B, device=local_rank)
        pinned_like = lambda a: torch.empty(a.nbytes, dtype=torch.uint8).pin_memory().numpy().view(a.dtype).reshape(a.shape)  # noqa: E731
        outs = [(h_kps, h_desc, h_cnt, h_seg, h_planes, h_npl), tuple(pinned_like(a) for a in (h_kps, h_desc, h_cnt, h_seg, h_planes, h_npl))]
        pairs = [(orb, cape), (orb_b, cape_b)]

        def issue(i):
            o, c = pairs[i]
            kp, de, cn, sg, pl, npl_ = outs[i]
            o.extract_batch(h_gray, kp, de, cn, chunk_frames=args.chunk)
            c.process_depth_batch(h_depth16, *K, depth_factor=float(fac), seg=sg, planes=pl, nplanes=npl_, chunk_frames=args.chunk)

        def finish(i):
            pairs[i][0].finish_batch(); pairs[i][1].finish_batch()

        def run_two(n):
            issue(0)
            for s_ in range(1, n):
                issue(s_ % 2)
                finish((s_ - 1) % 2)
            finish((n - 1) % 2)
        run_two(3)
        barrier()
        n_two = e2e_steps
        t0 = time.perf_counter()
        run_two(n_two)
        orb_b.sync(); cape_b.sync()
        barrier()
        e2e_two = {"value": world * B * n_two / max_over_ranks(time.perf_counter() - t0), "unit": "frames/s", "steps": n_two,
                   "same_results": bool(np.array_equal(outs[0][2], outs[1][2]) and np.array_equal(outs[0][3], outs[1][3])
                                        and np.array_equal(outs[0][1], outs[1][1])),
                   "note": "two handle pairs used alternately: step s is issued (drfe_orb_extract_batch + drfe_cape_process_depth_batch) before "
                           "step s-1 is finished (drfe_*_finish_batch); all of every step's H2D and D2H bytes inside the timed region"}
        del orb_b, cape_b
    h2d_u16 = int(h_gray.nbytes + h_depth16.nbytes)
    h2d_f32 = int(h_gray.nbytes + h_depth.nbytes)
    d2h = int(h_kps.nbytes + h_desc.nbytes + h_cnt.nbytes + h_seg.nbytes + h_planes.nbytes + h_npl.nbytes)

    # ---- the host link's ceiling for exactly these bytes: every rank copies one step's H2D and D2H bytes at once on two
    # streams, nothing else running.  e2e / ceiling = how close the pipelined calls come to what the box can feed.
    d_in = torch.empty(h2d_u16, dtype=torch.uint8, device="cuda")
    d_out = torch.empty(d2h, dtype=torch.uint8, device="cuda")
    p_in = torch.empty(h2d_u16, dtype=torch.uint8).pin_memory()
    p_out = torch.empty(d2h, dtype=torch.uint8).pin_memory()
    s_in, s_out = torch.cuda.Stream(), torch.cuda.Stream()

    def step_copy(up=True, down=True):
        if up:
            with torch.cuda.stream(s_in):
                d_in.copy_(p_in, non_blocking=True)
        if down:
            with torch.cuda.stream(s_out):
                p_out.copy_(d_out, non_blocking=True)
        s_in.synchronize(); s_out.synchronize()

    t_both = time_loop(step_copy) / e2e_steps
    t_up = time_loop(lambda: step_copy(True, False)) / e2e_steps
    t_down = time_loop(lambda: step_copy(False, True)) / e2e_steps
    pcie = {"ceiling_fps": world * B / t_both, "h2d_gbs_alone": world * h2d_u16 / t_up / 1e9, "d2h_gbs_alone": world * d2h / t_down / 1e9,
            "h2d_gbs_concurrent": world * h2d_u16 / t_both / 1e9, "d2h_gbs_concurrent": world * d2h / t_both / 1e9,
            "note": "all %d rank(s) copy one step's bytes (H2D %d MB + D2H %d MB per GPU) at once from / to pinned memory on two "
                    "streams, nothing else running; sums over ranks, slowest rank's time" % (world, h2d_u16 // 1000000, d2h // 1000000)}
    del d_in, d_out, p_in, p_out

    # ---- the same batch through the multi-device pool (ONE process, one host thread + handle pair per device)
    pool_res = None
    if not args.no_extras:
        if world > 1:
            barrier()
            dist.barrier(group=cpu_group)
        if rank == 0:
            try:
                ndev = world
                del rig.d_gray, rig.d_depth
                pool = drfe.Pool(list(range(ndev)), nfeatures=wl.nfeat, width=W, height=H, cell=wl.cell, cylinder_detection=wl.cyl,
                                 min_cos=MIN_COS, max_merge_dist=wl.max_merge, max_batch=B * ndev, chunk_frames=args.chunk)
                def reps(a):                                       # rank 0's frames once per device, in pinned memory of the same dtype
                    if ndev == 1:
                        return a
                    out = drfe.host_array((a.shape[0] * ndev,) + a.shape[1:], a.dtype)
                    for i in range(ndev):
                        out[i * a.shape[0]:(i + 1) * a.shape[0]] = a
                    return out
                pg, pd = reps(h_gray), reps(h_depth16)
                out = {"kps": reps(h_kps), "desc": reps(h_desc), "counts": reps(h_cnt), "seg": reps(h_seg), "planes": reps(h_planes),
                       "nplanes": reps(h_npl)}
                run = lambda: pool.extract_batch(pg, pd, *K, depth_factor=float(fac), out=out)   # noqa: E731
                for _ in range(2):
                    run()
                t0 = time.perf_counter()
                for _ in range(e2e_steps):
                    run()
                dt = time.perf_counter() - t0
                assert np.array_equal(out["counts"][:B], h_cnt) and np.array_equal(out["seg"][B * (ndev - 1):], h_seg)   # same frames, same results, on every device
                pool_res = {"value": ndev * B * e2e_steps / dt, "unit": "frames/s", "devices": ndev, "processes": 1,
                            "device_ms": [round(float(x), 3) for x in pool.device_times()],
                            "api": "drfe_pool_extract_batch: one call per step, %d frames cut into contiguous blocks of %d, one host thread "
                                   "and one ORB + CAPE handle pair per device%s" % (ndev * B, B, "" if ndev == 1 else
                                   "; measured by rank 0 alone while the other ranks wait (their GPUs are otherwise idle)")}
                pool.close()
                del pg, pd, out
            except Exception as e:                                 # the headline must not depend on this leg
                pool_res = {"error": str(e)}
        if world > 1:
            dist.barrier(group=cpu_group)

    # ---- single-frame latency, the reference's call shape (BASELINE configs[1]): one frame from host memory through
    # ORBextractor::operator() / PlaneDetection_CAPE (drfe_orb_extract, drfe_cape_process_depth), results on the host
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
    if rank == 0 and not args.no_extras:
        ex1 = drfe.ORBextractor(wl.nfeat, 1.2, 8, 20, 7, W, H, device=local_rank)
        cp1 = drfe.CAPE(H, W, wl.cell, wl.cell, wl.cyl, MIN_COS, wl.max_merge, device=local_rank)
        single = {"orb_extract_ms": median_ms(lambda i: ex1(gray[i % B], None)),
                  "cape_process_depth_ms": median_ms(lambda i: cp1.process_depth(depth[i % B], *K)),
                  "note": "median of 100 host-in/host-out single-frame calls (drfe_orb_extract, drfe_cape_process_depth)"}
        del ex1, cp1

    # ---- configs[4] (1280x720, 2000 kp, cylinders on) as an extra key: a driver-run record of the stress config
    c720 = None
    if args.workload == "c640" and (args.with_c720 or world == 8) and not args.no_extras:
        del rig, orb, cape
        torch.cuda.empty_cache()
        w7 = Workload("c720")
        f7, l7 = shard.weak_block(w7.batch, rank)
        g7, d7, K7 = make_sequence(w7, f7, l7 - f7, threads)
        rig7 = Rig(drfe, torch, w7, g7, d7, K7, local_rank)

        def barrier7():
            rig7.sync()
            torch.cuda.synchronize()
            if dist:
                dist.barrier()
        for _ in range(3):
            rig7.step_resident()
        barrier7()
        npl7 = rig7.cape.download()[2]
        ncyl7 = rig7.cape.cylinders_found()
        rig7.orb.set_profiling(True); rig7.cape.set_profiling(True)
        steps7 = max(5, args.steps // 2)
        ms7, _ = rig7.timed_resident(steps7, barrier7)
        st7 = dict(rig7.orb.stage_times()); st7.update(dict(rig7.cape.stage_times()))
        ms7 = max_over_ranks(ms7)
        alg7 = w7.algorithmic_bytes()
        peak7, _ = measured_peak()
        v7 = world * w7.batch * steps7 / (ms7 * 1e-3)
        c720 = {"metric": "front-end frames/sec (1280x720 RGB-D, ORB 2000 kp + CAPE with cylinders)", "value": v7, "unit": "frames/s",
                "n_gpus": world, "steps": steps7, "ms_per_step": ms7 / steps7, "config": w7.config(),
                "planes_per_frame": float(npl7.mean()), "cylinders_per_frame": float(np.mean(ncyl7)),
                "whole_step_hbm_frac": alg7["total"] * v7 / world / 1e9 / peak7,
                "stage_ms": {k: round(v, 4) for k, v in st7.items()}}
        del rig7

    # ---- PEAC-AHC (the plane extractor that is live in Frame::Frame) on the same depth maps: an extra key
    peac = None
    if world == 1 and args.workload == "c640" and not args.no_extras:
        q16 = np.rint(depth * np.float32(5000.0)).astype(np.uint16)
        col = np.where(q16 > 0, np.arange(W)[None, None, :], 0)      # isolated dropouts take the value to their left: INIT_STRICT
        q16 = np.take_along_axis(q16, np.maximum.accumulate(col, axis=2), axis=2)   # rejects every window with a missing pixel
        fac = float(np.float32(1.0) / np.float32(5000.0))
        pe = drfe.PEAC(W, H, max_batch=B, device=local_rank)
        d_q = torch.from_numpy(q16.view(np.int16)).cuda(local_rank)

        def peac_step():
            pe.enqueue(d_q.data_ptr(), fac, *K, nframes=B, mem_kind=drfe.MEM_DEVICE)
            pe.sync()
        peac_step()
        t0 = time.perf_counter()
        for _ in range(3):
            peac_step()
        ms_p = (time.perf_counter() - t0) / 3 * 1e3
        npl_p = pe.download()[2]
        pe1 = drfe.PEAC(W, H, device=local_rank)

        def peac_one(i):
            pe1.enqueue(q16[i % B][None], fac, *K)
            pe1.download()
        lat = median_ms(peac_one, 10)
        from oracle import oracle as _orc
        t0 = time.perf_counter()
        for i in range(4):
            _orc.peac_run(_orc.peac_cloud(q16[i * 8], fac, *K), W, H)
        cpu_ms = (time.perf_counter() - t0) / 4 * 1e3
        peac = {"metric": "PEAC-AHC plane extraction frames/sec (PlaneDetection::readDepthImage + runPlaneDetection)", "value": B / (ms_p * 1e-3),
                "unit": "frames/s", "ms_per_batch": ms_p, "frames": B, "planes_per_frame": float(npl_p.mean()),
                "single_frame_ms": lat, "cpu_port_ms_per_frame": cpu_ms, "cpu_cores": 1, "bound": "latency",
                "note": "one CTA per frame (the clustering and the region growing are sequential in the reference); depth resident on the "
                        "device, wall clock around enqueue + sync; CPU = oracle/peac_oracle.cpp on one core, 4 frames"}
        del pe, pe1, d_q

    if rank != 0:
        if dist:
            dist.destroy_process_group()
        return

    # ---- roofline of the dominant kernel
    alg = wl.algorithmic_bytes()
    kernel_stages = {k: v for k, v in stages.items() if k in alg and k != "total"}
    dom = max(kernel_stages, key=kernel_stages.get)
    peak, peak_src = measured_peak()
    launches_per_stage = {"pyramid": 8, "fit": 2}
    dom_ms = kernel_stages[dom]
    per_launch_ms = dom_ms / launches_per_stage.get(dom, 1)
    bytes_per_launch = alg[dom] * B / launches_per_stage.get(dom, 1)
    achieved = bytes_per_launch / (per_launch_ms * 1e-3) / 1e9 if per_launch_ms > 0 else 0.0
    # DRAM bytes and pipe utilisation of the same kernel from the committed ncu --set full capture, per launch of 256 frames
    traffic = alu = issue = None
    try:
        if args.workload == "c640":
            rec = json.load(open(os.path.join(ROOT, TRAFFIC_JSON)))["stages"][dom]
            traffic, alu, issue = rec.get("dram_bytes"), rec.get("alu_pipe_pct"), rec.get("issue_pct")
    except Exception:
        pass
    bound = {"fast": "alu_pipe", "pyramid": "issue", "blur": "issue", "orient_describe": "l1", "cells": "issue"}.get(dom, "hbm")
    roofline = {"bound": bound, "kernel": dom,
                "bound_evidence": {"pipe": "alu (half-rate: 0.5 warp-inst/clk/SMSP)" if bound == "alu_pipe" else bound,
                                   "alu_pipe_util_pct": alu, "issue_slot_util_pct": issue,
                                   "source": "ncu --set full, %s and profiles/r02_*_full.txt; tools/ubench/pipes.cu" % TRAFFIC_JSON},
                "achieved": achieved, "peak": peak, "unit": "GB/s",
                "frac": achieved / peak, "traffic": traffic, "traffic_source": TRAFFIC_JSON + " (committed ncu capture of this step, not measured in this run)",
                "peak_source": peak_src,
                "note": "the kernel is bound by instruction issue on the half-rate integer ALU pipe, not by HBM; `frac` is its algorithmic "
                        "bytes against the HBM copy peak, as the contract asks" if bound != "hbm" else None,
                "alg_bytes_per_launch": bytes_per_launch, "launch_ms": per_launch_ms,
                "whole_step": {"alg_bytes_per_frame": alg["total"],
                               "achieved": alg["total"] * B * args.steps / (ms * 1e-3) / 1e9,
                               "frac": alg["total"] * B * args.steps / (ms * 1e-3) / 1e9 / peak,
                               "note": "per GPU: all stages' algorithmic bytes / the step's device time, against the HBM copy peak"},
                "stage_ms": {k: round(v, 4) for k, v in stages.items()},
                "stage_ms_in_step": {k: round(v, 4) for k, v in stages_in_step.items()},
                "stage_ms_note": "stage_ms: CUDA events on the handle's stream with one handle running at a time (5 batches after the timed "
                                 "region) — kernel durations, used for `achieved`; stage_ms_in_step: the same events inside the timed two-stream "
                                 "steps, where an interval also holds the wait for SM slots behind the other handle's kernels"}

    # ---- CPU baseline on a bounded sample (rank 0, N=1 only)
    cpu = None
    if world == 1 and not args.no_extras:
        sample = args.cpu_sample or B
        fps, dt = cpu_port_fps(wl, gray, depth, K, sample, ncpu, min_seconds=10.0)
        cpu = {"value": fps, "unit": "frames/s", "cores": ncpu, "kind": "port",
               "sample": "passes over the first %d frames of the same batch, frame-parallel on %d threads, for %.1f s "
                         "(%d frames); CPU oracle port of ORBextractor+CAPE, -O3 x86-64-v3"
                         % (sample, ncpu, dt, int(round(fps * dt))),
               "reference_shaped": cpu_reference_shaped(wl, gray, depth, K, 48),
               "cv2_stage_ms": cv2_stage_ms(wl, gray, 8)}

    line = {
        "metric": METRIC, "value": value, "unit": "frames/s", "n_gpus": world, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": ms / args.steps,
        "ms_per_step_percentiles": {"p10": float(np.percentile(step_ms, 10)), "p50": float(np.percentile(step_ms, 50)),
                                    "p90": float(np.percentile(step_ms, 90)),
                                    "note": "per step = the later of the step's ORB-stream and CAPE-stream end events, rank 0.  The two handles "
                                            "run free on their own streams (like the reference's two threads), so consecutive steps overlap "
                                            "differently and the per-step figure spreads around the mean; the mean over the timed region is what `value` uses"},
        "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "u8/f32/f64", "data": "synthetic",
        "config": wl.config(),
        # headline e2e: what Frame::Frame receives — the gray image and the sensor's raw 16-bit depth (imDepth is
        # converted to float INSIDE the path, Frame.cc:113-115); the float-depth variant of the same call is kept beside it
        "e2e": {"value": e2e_u16, "unit": "frames/s", "h2d_bytes_per_step": h2d_u16,
                "d2h_bytes_per_step": d2h, "steps": e2e_steps,
                "pcie": pcie, "pcie_frac": e2e_u16 / pcie["ceiling_fps"],
                "pool": pool_res,
                "two_in_flight": e2e_two,
                "api": "drfe_orb_extract_batch + drfe_cape_process_depth_batch (gray u8 + raw u16 depth as Frame::Frame gets "
                       "them, depth scaled on the device as Frame.cc:113-115 does; pinned host buffers, 32-frame chunks "
                       "(8/16-frame chunks at both ends) pipelined H2D | kernels | D2H; one process per GPU, its host threads pinned to their own cores)",
                "float_depth": {"value": e2e_value, "unit": "frames/s", "h2d_bytes_per_step": h2d_f32,
                                "note": "same call fed float depth (what PlaneDetection_CAPE::readDepthImage takes)"}},
        "strong": strong,
        "c720": c720,
        "peac": peac,
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
