#!/usr/bin/env python
"""Single-frame latency of the reference's call shape (BASELINE configs[1]): ORBextractor::operator() and
PlaneDetection_CAPE::runPlaneDetection on one 640x480 frame from host memory, results back on the host.
ORB and CAPE are timed alone and, as Frame::Frame runs them, on two host threads at once."""
import os
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "dr-slam_b200"))
sys.path.insert(0, ROOT)
import drfe  # noqa: E402

MC = float(np.float32(np.cos(np.pi / 12)))


def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 300
    frames = [drfe.synth_frame(640, 480, (i // 8) % 3, 20260000 + i) for i in range(16)]
    K = frames[0][2]
    ex = drfe.ORBextractor(1000, 1.2, 8, 20, 7, 640, 480)
    cp = drfe.CAPE(480, 640, 20, 20, False, MC, 50.0)

    def orb(i):
        ex(frames[i % 16][0], None)

    def cape(i):
        cp.process_depth(frames[i % 16][1], *K)

    def both(i):
        t = threading.Thread(target=cape, args=(i,))
        t.start()
        orb(i)
        t.join()

    for name, fn in (("ORBextractor::operator()", orb), ("CAPE (depth -> planes)", cape), ("both, two host threads", both)):
        for i in range(20):
            fn(i)
        ts = []
        for i in range(n):
            t0 = time.perf_counter()
            fn(i)
            ts.append(time.perf_counter() - t0)
        ts = np.array(ts) * 1e3
        print("%-28s median %.3f ms  p10 %.3f  p90 %.3f  (%d calls)" % (name, np.median(ts), np.percentile(ts, 10), np.percentile(ts, 90), n))


if __name__ == "__main__":
    main()
