"""The C++ host adapters (dr-slam_b200/host/ORBextractor.h, CAPE.h) driven the way Frame::Frame
drives the reference extractors (Frame.cc:124-134), compared with the ctypes path and the oracle."""
import os
import re
import subprocess

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
EXE = os.path.join(ROOT, "dr-slam_b200", "host", "example_frontend")
MC = float(np.float32(np.cos(np.pi / 12)))


def fnv1a(b):
    h = 1469598103934665603
    for x in bytes(b):
        h = ((h ^ x) * 1099511628211) & 0xFFFFFFFFFFFFFFFF
    return h


@pytest.mark.parametrize("scene,seed", [(1, 20260042), (0, 20260007)])
def test_cpp_adapters_match_oracle(drfe, orc, scene, seed):
    assert os.path.exists(EXE), "run __graft_entry__.build() first"
    out = subprocess.run([EXE, str(scene), str(seed)], capture_output=True, text=True, timeout=120)
    assert out.returncode == 0, out.stderr
    head = out.stdout.splitlines()[0]
    m = re.match(r"keypoints (\d+) kp_hash (\w+) desc_hash (\w+) planes (\d+) seg_hash (\w+) plane_points (\d+) levels (\d+)", head)
    assert m, head
    gray, depth, K = drfe.synth_frame(640, 480, scene, seed)
    rk, rd = orc.OrbOracle(1000).extract(gray)
    assert int(m.group(1)) == len(rk) and int(m.group(7)) == 8
    assert int(m.group(2), 16) == fnv1a(rk.tobytes())            # cv::KeyPoint-layout keypoints, bit for bit
    ex = drfe.ORBextractor(1000, 1.2, 8, 20, 7, 640, 480)
    _, gd = ex(gray, None)
    assert int(m.group(3), 16) == fnv1a(gd.tobytes())
    assert (np.unpackbits(gd ^ rd, axis=1).sum(1) == 0).mean() >= 0.995
    o = orc.CapeOracle(480, 640, 20, 20, False, MC, 50.0)
    oseg, oplanes = o.process(o.depth_to_cloud(depth, *K))
    assert int(m.group(4)) == len(oplanes)
    assert int(m.group(5), 16) == fnv1a(oseg.tobytes())
    assert int(m.group(6)) == int((oseg > 0).sum())
    for i, line in enumerate(out.stdout.splitlines()[1:]):
        v = [float(x) for x in re.findall(r"-?\d+\.\d+", line)]
        assert np.allclose(v[:3], oplanes["normal"][i], atol=1e-5) and abs(v[3] - oplanes["d"][i]) < 1e-5


def test_handles_are_usable_from_fresh_host_threads(drfe, orc):
    """Frame::Frame starts a new std::thread per frame for ExtractORB and for the plane extractor (Frame.cc:124-134):
    handles created on one thread must work from any other, ORB and CAPE at the same time, with the same results."""
    import threading
    MC = float(np.float32(np.cos(np.pi / 12)))
    frames = [drfe.synth_frame(640, 480, i % 3, 20260300 + i) for i in range(6)]
    K = frames[0][2]
    ex = drfe.ORBextractor(1000, 1.2, 8, 20, 7, 640, 480)
    cp = drfe.CAPE(480, 640, 20, 20, False, MC, 50.0)
    want = []
    for g, d, _ in frames:                                  # sequential, this thread
        kps, desc = ex(g, None)
        npl, _, seg, planes, _ = cp.process_depth(d, *K)
        want.append((kps.copy(), desc.copy(), npl, seg.copy()))
    got = [None] * len(frames)
    for i, (g, d, _) in enumerate(frames):                  # one fresh thread per frame and extractor, as Frame does
        out = {}

        def run_orb():
            out["orb"] = ex(g, None)

        def run_cape():
            out["cape"] = cp.process_depth(d, *K)

        ts = [threading.Thread(target=run_orb), threading.Thread(target=run_cape)]
        for t in ts:
            t.start()
        for t in ts:
            t.join()
        got[i] = out
    for w, o in zip(want, got):
        kps, desc = o["orb"]
        assert all(np.array_equal(kps[n], w[0][n]) for n in kps.dtype.names) and np.array_equal(desc, w[1])
        assert o["cape"][0] == w[2] and np.array_equal(o["cape"][2], w[3])
