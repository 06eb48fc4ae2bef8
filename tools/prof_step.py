#!/usr/bin/env python
"""Light driver for kernel iteration on the GPU box: the bench workload without the e2e / CPU legs.

  python tools/prof_step.py [--frames 256] [--steps 5] [--only orb|cape]

ORB and CAPE are run one after the other (not concurrently, unlike bench.py), so the per-stage
CUDA-event times are not disturbed by the other stream.  Use it under ncu:
  ncu --set full --import-source on -k regex:k_ -s <skip> -c <n> -o gpurun_out/x python tools/prof_step.py --steps 2
"""
import argparse
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "dr-slam_b200"))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
import drfe  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--frames", type=int, default=256)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--only", default="")
    ap.add_argument("--width", type=int, default=640)
    ap.add_argument("--height", type=int, default=480)
    ap.add_argument("--nfeatures", type=int, default=1000)
    ap.add_argument("--cyl", type=int, default=0, help="cylinder detection on")
    ap.add_argument("--unit", type=float, default=1.0, help="depth unit scale (1000 = millimetres)")
    ap.add_argument("--scene", type=int, default=-1, help="fixed scene (default: cycle 0,1,2)")
    a = ap.parse_args()
    import torch
    B, W, H = a.frames, a.width, a.height
    uniq = min(B, 32)
    data = [drfe.synth_frame(W, H, a.scene if a.scene >= 0 else (i // 8) % 3, 20260000 + i, a.unit) for i in range(uniq)]
    gray = np.stack([data[i % uniq][0] for i in range(B)])
    depth = np.stack([data[i % uniq][1] for i in range(B)])
    K = data[0][2]
    d_gray, d_depth = torch.from_numpy(gray).cuda(), torch.from_numpy(depth).cuda()
    torch.cuda.synchronize()
    res = {}
    if a.only in ("", "orb"):
        orb = drfe.ORBextractor(a.nfeatures, 1.2, 8, 20, 7, W, H, max_batch=B)
        for _ in range(2):
            orb.enqueue(d_gray.data_ptr(), drfe.MEM_DEVICE, B, W, W * H)
        orb.sync()
        orb.set_profiling(True)
        for _ in range(a.steps):
            orb.enqueue(d_gray.data_ptr(), drfe.MEM_DEVICE, B, W, W * H)
        orb.sync()
        res.update(dict(orb.stage_times()))
        print("keypoints/frame: mean %.1f" % orb.download()[2].mean())
    if a.only in ("", "cape"):
        cape = drfe.CAPE(H, W, 20, 20, bool(a.cyl), bench.MIN_COS, 50.0, max_batch=B)
        for _ in range(2):
            cape.enqueue_depth(d_depth.data_ptr(), *K, mem_kind=drfe.MEM_DEVICE, nframes=B, row_stride=W, frame_stride=W * H)
        cape.sync()
        cape.set_profiling(True)
        for _ in range(a.steps):
            cape.enqueue_depth(d_depth.data_ptr(), *K, mem_kind=drfe.MEM_DEVICE, nframes=B, row_stride=W, frame_stride=W * H)
        cape.sync()
        res.update(dict(cape.stage_times()))
        print("planes/frame: mean %.2f" % cape.download()[2].mean())
        dbg = np.stack([cape.debug_counters(i) for i in range(min(B, 32))])
        names = ["seeds", "sweeps", "sum_ncand", "sum_nact", "cyc_argmax_list", "cyc_scan", "cyc_grow", "cyc_accum", "cyc_fit"]
        print("grid stage per frame (mean over %d frames): " % len(dbg) + ", ".join("%s %.0f" % (n, dbg[:, i].mean()) for i, n in enumerate(names)) +
              ", cyc_tail %.0f" % (dbg[:, 10] - dbg[:, 9] - dbg[:, 4:9].sum(1)).mean() +
              " (jobs %.0f, cylinders %.0f, labels %.0f, merge..end %.0f)" % ((dbg[:, 12] - dbg[:, 11]).mean(), (dbg[:, 13] - dbg[:, 12]).mean(),
                                                                           (dbg[:, 14] - dbg[:, 13]).mean(), (dbg[:, 10] - dbg[:, 14]).mean()))
    tot = sum(res.values())
    print("stage ms per %d-frame batch (serialised): " % B + ", ".join("%s %.3f" % kv for kv in res.items()) +
          " | total %.3f ms => %.0f frames/s" % (tot, B / tot * 1e3))


if __name__ == "__main__":
    main()
