"""Generates tests/golden/peac_640x480.npz: golden vectors of the PEAC-AHC plane extractor from the C++
restatement oracle/peac_oracle.cpp — two frames' 16-bit depth maps (stored inside the fixture), seg_output, the extracted planes
(normal, centre, mse, curvature as doubles; N; rid), plane_vertices_ and the number of clustering steps.  The reference itself
cannot be built here and ships no fixtures; these vectors pin the restatement (and, on the GPU, drfe_peac_*) against drift.
   python tests/golden/make_golden_peac.py"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, os.path.join(ROOT, "dr-slam_b200"))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import drfe  # noqa: E402  (synthetic generator only, host code)
from oracle import oracle as orc  # noqa: E402
from test_peac import FAC, frame  # noqa: E402


def main():
    out = {}
    for i, (scene, seed, holes) in enumerate([(1, 20260012, ((100, 140, 300, 420),)), (2, 20260100, ((0, 60, 0, 640), (200, 260, 100, 180)))]):
        q, K = frame(drfe, scene, seed, holes)
        seg, planes, members, steps = orc.peac_run(orc.peac_cloud(q, FAC, *K), 640, 480)
        out["depth%d" % i] = q
        out["K%d" % i] = np.array(K, np.float32)
        out["seg%d" % i] = seg
        out["planes%d" % i] = planes
        out["member_offsets%d" % i] = np.cumsum([0] + [len(m) for m in members]).astype(np.int32)
        out["member_idx%d" % i] = np.concatenate(members).astype(np.int32) if members else np.zeros(0, np.int32)
        out["steps%d" % i] = np.int32(steps)
        print("frame %d: %d planes, %d steps, %d labelled pixels" % (i, len(planes), steps, int((seg > 0).sum())))
    out["depth_factor"] = np.float32(FAC)
    np.savez_compressed(os.path.join(HERE, "peac_640x480.npz"), **out)


if __name__ == "__main__":
    main()
