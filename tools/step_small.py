#!/usr/bin/env python
"""Two-stream step time (ORB and CAPE handles running free, as in bench.py) for a share of n frames of the 256-frame batch — what a
GPU does per step under strong scaling (n = 256 / N).   python tools/step_small.py [n ...]"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "dr-slam_b200"))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
import drfe  # noqa: E402


def main():
    import torch
    wl = bench.Workload("c640")
    gray, depth, K = bench.make_sequence(wl, 0, wl.batch, 8)
    rig = bench.Rig(drfe, torch, wl, gray, depth, K, 0)

    def barrier():
        rig.sync()
        torch.cuda.synchronize()
    for n in [int(a) for a in sys.argv[1:]] or [256, 128, 64, 32]:
        for _ in range(3):
            rig.step_resident(n)
        ms, _ = rig.timed_resident(40, barrier, n)
        print("n = %3d frames per step: %.4f ms per step, %.0f frames/s" % (n, ms / 40, n * 40 / ms * 1e3))


if __name__ == "__main__":
    main()
