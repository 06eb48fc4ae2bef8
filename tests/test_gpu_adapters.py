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
    for i, line in enumerate(out.stdout.splitlines()[1:1 + len(oplanes)]):
        v = [float(x) for x in re.findall(r"-?\d+\.\d+", line)]
        assert np.allclose(v[:3], oplanes["normal"][i], atol=1e-5) and abs(v[3] - oplanes["d"][i]) < 1e-5
    # Planar_SLAM::PlaneDetection (host/PlaneExtractor.h) the way Frame::ComputePlanes reads it, against the PEAC restatement
    last = out.stdout.splitlines()[-1]
    m = re.match(r"peac planes (\d+) seg_hash (\w+) vertices (\d+) index_hash (\w+) point_hash (\w+)(.*)", last)
    assert m, last
    q = np.rint(depth * 5000).astype(np.uint16)
    col = np.where(q > 0, np.arange(640)[None, :], 0)                 # dropouts take the value to their left (as the example does)
    q = np.take_along_axis(q, np.maximum.accumulate(col, axis=1), axis=1)
    fac = float(np.float32(1.0) / np.float32(5000.0))
    cloud = orc.peac_cloud(q, fac, *K)
    pseg, pplanes, pmem, _ = orc.peac_run(cloud, 640, 480)
    assert int(m.group(1)) == len(pplanes) and int(m.group(2), 16) == fnv1a(pseg.tobytes())
    assert int(m.group(3)) == sum(len(x) for x in pmem)
    if len(pmem):
        assert int(m.group(4), 16) == fnv1a(np.concatenate(pmem).astype(np.int32).tobytes())
        assert int(m.group(5), 16) == fnv1a(np.concatenate([cloud[x] for x in pmem]).astype(np.float32).tobytes())
    parts = [part.split() for part in m.group(6).split("|")[1:]]
    coarse = [pt for pt in parts if pt and pt[0] == "coarse"]
    nrm = [pt for pt in parts if pt and pt[0] == "normals"]
    vals = [[float(x) for x in pt] for pt in parts if pt and pt[0] not in ("coarse", "normals")]
    # thirdCloudNormals(10 m): the 1/3 cloud of imDepth = float(raw) * factor and the restated PCL normals on it
    wn = orc.integral_normals(orc.third_cloud(q.astype(np.float32) * np.float32(fac), *K, 10.0))
    okn = ~np.isnan(wn[..., 0])
    assert len(nrm) == 1 and [int(x) for x in nrm[0][1:4]] == [214, 160, int(okn.sum())]
    assert int(nrm[0][4], 16) == fnv1a(wn[okn].tobytes())
    # planeCloudsVoxel(3 m, 5 cm): the member lists of the restatement, culled and voxel-filtered by the voxel-grid restatement
    want = [orc.voxel_grid(x[~(x[:, 2] > np.float32(3.0))], 0.05)[0] for x in (cloud[mm].astype(np.float32) for mm in pmem)]
    assert len(coarse) == 1 and int(coarse[0][1]) == sum(len(w) for w in want)
    assert int(coarse[0][2], 16) == fnv1a(b"".join(w.tobytes() for w in want))
    for i, v in enumerate(vals):
        assert np.array_equal(v[:3], pplanes[i, :3]) and abs(v[3] + float(pplanes[i, :3] @ pplanes[i, 3:6])) < 1e-12
    assert len(vals) == len(pplanes)


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


def write_vocabulary_file(path, voc):
    """TemplatedVocabulary::saveToTextFile format (Thirdparty/DBoW2/DBoW2/TemplatedVocabulary.h:1427-1452), what ORBvoc.txt is"""
    with open(path, "w") as f:
        f.write("%d %d  %d %d\n" % (voc["k"], voc["L"], voc["scoring"], voc["weighting"]))
        for i in range(len(voc["parent"])):
            f.write("%d %d %s %r\n" % (voc["parent"][i], voc["is_leaf"][i], " ".join(str(int(b)) for b in voc["descriptors"][i]), float(voc["weights"][i])))


def test_cpp_vocabulary_and_matchers(drfe, orc, tmp_path):
    """host/ORBVocabulary.h (loadFromTextFile + transform) and host/ORBmatcher.h (the two whole-function matchers) from C++,
    against the oracle on the same frame"""
    import struct
    scene, seed = 1, 20260042
    voc = orc.synth_vocabulary(10, 5, 61)
    vf = str(tmp_path / "voc.txt")
    write_vocabulary_file(vf, voc)
    out = subprocess.run([EXE, str(scene), str(seed), vf], capture_output=True, text=True, timeout=180)
    assert out.returncode == 0, out.stderr
    line = [ln for ln in out.stdout.splitlines() if ln.startswith("words ")][0]
    m = re.match(r"words (\d+) bow (\d+) bow_hash (\w+) fv (\d+) fv_hash (\w+) proj (-?\d+) proj_hash (\w+) bowmatch (-?\d+) bowmatch_hash (\w+)", line)
    assert m, line
    gray, depth, K = drfe.synth_frame(640, 480, scene, seed)
    keys, desc = orc.OrbOracle(1000).extract(gray)
    n = len(keys)
    V = orc.Vocabulary(**voc)
    _, _, bow, fv = V.transform(desc, 4)
    assert int(m.group(1)) == V.nwords and int(m.group(2)) == len(bow) and int(m.group(4)) == len(fv)
    hb = b"".join(struct.pack("<Id", k, v) for k, v in bow)
    hf = b"".join(struct.pack("<I", k) + np.array(l, np.uint32).tobytes() for k, l in fv)
    assert int(m.group(3), 16) == fnv1a(hb) and int(m.group(5), 16) == fnv1a(hf)
    # the two matchers: same inputs as example_frontend.cpp builds
    p = orc.frame_params(*K, [0.1, -0.05, 0.001, 0.0005, 0.0], 40.0, 640, 480)
    ku, ur, kd, gc, gi = orc.frame_post(p, keys, depth)
    f32 = np.float32
    pts = np.zeros(n, orc.LAST_POINT_DTYPE)
    z = np.where(kd > 0, kd, f32(2.0)).astype(f32)
    pts["X"] = (ku["x"] - f32(K[2])) * z / f32(K[0])
    pts["Y"] = (ku["y"] - f32(K[3])) * z / f32(K[1])
    pts["Z"], pts["angle"], pts["octave"], pts["flags"] = z, ku["angle"], ku["octave"], orc.LP_VALID | orc.LP_OBSERVED
    sf = np.array(orc.OrbOracle(1000).scale_factors(), f32)
    mk, md, holder, nm = orc.search_last_frame(p, sf, ku, ur, gc, gi, desc, np.eye(3, 4, dtype=f32).ravel(), 15.0, 0, 1, pts, desc)
    assert int(m.group(6)) == nm and nm > 500
    assert int(m.group(7), 16) == fnv1a(holder.astype(np.int32).tobytes())
    km, fm, nb = orc.search_by_bow(desc, ku["angle"], np.ones(n, np.uint8), fv, desc, keys["angle"], fv, 0.7, True)
    assert int(m.group(8)) == nb and nb > 300
    assert int(m.group(9), 16) == fnv1a(fm.astype(np.int32).tobytes())
