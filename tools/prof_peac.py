#!/usr/bin/env python
"""PEAC-AHC on a batch (depth resident on the device): time per batch, k_peac_frame's stage counters, the CPU restatement beside it.

  python tools/prof_peac.py [--frames 256] [--steps 3]
Under ncu:  ncu --set full --import-source on -k regex:k_peac -c 3 -o gpurun_out/x python tools/prof_peac.py --frames 64 --steps 1"""
import argparse
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "dr-slam_b200"))
sys.path.insert(0, ROOT)
import drfe  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--frames", type=int, default=256)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--cpu", type=int, default=0, help="also time the CPU restatement on this many frames")
    a = ap.parse_args()
    import torch
    W, H, B = 640, 480, a.frames
    uniq = min(B, 32)
    data = [drfe.synth_frame(W, H, (i // 8) % 3, 20260000 + i) for i in range(uniq)]
    q = np.stack([np.rint(d[1] * np.float32(5000.0)).astype(np.uint16) for d in data])
    col = np.where(q > 0, np.arange(W)[None, None, :], 0)          # isolated dropouts take the value to their left
    q = np.take_along_axis(q, np.maximum.accumulate(col, axis=2), axis=2)
    qB = np.ascontiguousarray(q[np.arange(B) % uniq])
    K = data[0][2]
    fac = float(np.float32(1.0) / np.float32(5000.0))
    pe = drfe.PEAC(W, H, max_batch=B)
    d_q = torch.from_numpy(qB.view(np.int16)).cuda()
    torch.cuda.synchronize()
    ms = []
    for s in range(a.steps + 1):
        t0 = time.perf_counter()
        pe.enqueue(d_q.data_ptr(), fac, *K, nframes=B, mem_kind=drfe.MEM_DEVICE)
        pe.sync()
        ms.append((time.perf_counter() - t0) * 1e3)
    seg, planes, npl = pe.download()
    best = min(ms[1:]) if a.steps else ms[0]
    print("PEAC %d frames: %.2f ms per batch = %.3f ms per frame (%.0f frames/s); planes per frame %.2f, labelled pixels %.1f %%"
          % (B, best, best / B, B / best * 1e3, npl.mean(), 100.0 * (seg > 0).mean()))
    names = ["reset", "edges", "cluster 1", "membership + seeds", "region growing", "cluster 2", "outputs"]
    c = np.stack([pe.counters(f) for f in range(min(B, 32))])
    kc = np.diff(np.concatenate([np.zeros((len(c), 1), np.int64), c[:, 4:11].astype(np.int64)], axis=1), axis=1)
    print("k_peac_frame kilocycles per frame (mean of %d): " % len(c) + ", ".join("%s %d" % (n, v) for n, v in zip(names, kc.mean(0))) +
          "; total %d; clustering steps %.0f, region-growing queue %.0f" % (c[:, 10].mean(), c[:, 0].mean(), c[:, 1].mean()))
    if a.cpu:
        from oracle import oracle as orc
        t0 = time.perf_counter()
        for i in range(a.cpu):
            orc.peac_run(orc.peac_cloud(qB[i % uniq], fac, *K), W, H)
        print("CPU restatement (oracle/peac_oracle.cpp, one core): %.2f ms per frame" % ((time.perf_counter() - t0) / a.cpu * 1e3))


if __name__ == "__main__":
    main()
