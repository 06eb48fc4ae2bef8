"""GPU: BASELINE-size batches (256 frames of 640x480) checked through size-independent
properties plus oracle spot checks: every frame of a batch gives exactly what the same frame
gives alone (frames are independent: no cross-frame state), results do not depend on the
position inside the batch or on how the batch is sharded, keypoints respect the reference's
geometric invariants."""
import os
import threading
from concurrent.futures import ThreadPoolExecutor

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
MC = float(np.float32(np.cos(np.pi / 12)))
B = 256


@pytest.fixture(scope="module")
def sequence(drfe):
    data = [drfe.synth_frame(640, 480, (i // 64) % 3, 20260000 + i) for i in range(B)]
    return np.stack([d[0] for d in data]), np.stack([d[1] for d in data]), data[0][2]


def test_orb_full_batch(drfe, orc, sequence):
    gray, _, _ = sequence
    ex = drfe.ORBextractor(1000, 1.2, 8, 20, 7, 640, 480, max_batch=B)
    ex.enqueue(gray)
    kps, desc, counts = ex.download()
    assert counts.min() >= 1000 and counts.max() <= ex.cap        # >= nfeatures on textured frames, <= N + overshoot
    scale = ex.GetScaleFactors()
    for f in range(B):
        k = kps[f, :counts[f]]
        lx, ly = k["x"] / scale[k["octave"]], k["y"] / scale[k["octave"]]
        lw = np.array([ex.level_size(l)[0] for l in range(8)])[k["octave"]]
        lh = np.array([ex.level_size(l)[1] for l in range(8)])[k["octave"]]
        assert np.all(lx > 18.99) and np.all(ly > 18.99) and np.all(lx < lw - 18.99) and np.all(ly < lh - 18.99)   # EDGE_THRESHOLD
        assert np.all(np.diff(k["octave"]) >= 0)                                   # levels ascending (:1076-1104)
        assert np.all((k["angle"] >= 0) & (k["angle"] < 360)) and np.all(k["response"] >= 7)
    # sharding invariance: the same frames in two half batches at different positions
    ex2 = drfe.ORBextractor(1000, 1.2, 8, 20, 7, 640, 480, max_batch=B // 2)
    for half in range(2):
        sl = slice(half * B // 2, (half + 1) * B // 2)
        ex2.enqueue(gray[sl][::-1].copy())                         # reversed order inside the shard
        k2, d2, c2 = ex2.download()
        assert np.array_equal(c2[::-1], counts[sl])
        for j in (0, 17, B // 2 - 1):
            f = sl.start + (B // 2 - 1 - j)
            assert k2[j, :c2[j]].tobytes() == kps[f, :counts[f]].tobytes() and np.array_equal(d2[j, :c2[j]], desc[f, :counts[f]])
    # parity on EVERY frame of the batch (SURVEY.md 8d run 3): the CPU oracle over all host cores, one oracle object per thread
    tl = threading.local()

    def one(f):
        if not hasattr(tl, "o"):
            tl.o = orc.OrbOracle(1000)
        rk, rd = tl.o.extract(gray[f])
        assert counts[f] == len(rk), f
        assert kps[f, :counts[f]].tobytes() == rk.tobytes(), f                     # every cv::KeyPoint field, bit for bit
        ham = np.unpackbits(desc[f, :counts[f]] ^ rd, axis=1).sum(1)
        assert (ham == 0).mean() >= 0.995 and ham.max() <= 8, f
        return int((ham == 0).sum()), len(ham)
    with ThreadPoolExecutor(max_workers=os.cpu_count() or 4) as pool:
        same = list(pool.map(one, range(B)))
    assert sum(s for s, _ in same) >= 0.995 * sum(n for _, n in same)


def test_cape_full_batch(drfe, orc, sequence):
    _, depth, K = sequence
    cp = drfe.CAPE(480, 640, 20, 20, False, MC, 50.0, max_batch=B)
    cp.enqueue_depth(depth, *K)
    seg, planes, npl = cp.download()
    assert npl.min() >= 1 and seg.max() <= npl.max()
    for f in range(B):
        assert seg[f].max() <= npl[f]
        n = planes[f, :npl[f]]["normal"]
        assert np.allclose((n ** 2).sum(1), 1, atol=1e-9) and np.all(planes[f, :npl[f]]["d"] > 0)   # d > 0 orientation rule
    cp2 = drfe.CAPE(480, 640, 20, 20, False, MC, 50.0, max_batch=64)
    cp2.enqueue_depth(depth[100:164], *K)
    s2, p2, n2 = cp2.download()
    assert np.array_equal(s2, seg[100:164]) and np.array_equal(n2, npl[100:164])
    # parity on EVERY frame of the batch: the CPU oracle over all host cores
    tl = threading.local()

    def one(f):
        if not hasattr(tl, "o"):
            tl.o = orc.CapeOracle(480, 640, 20, 20, False, MC, 50.0)
        oseg, oplanes = tl.o.process(tl.o.depth_to_cloud(depth[f], *K))
        assert np.array_equal(seg[f], oseg) and npl[f] == len(oplanes), f
        assert np.allclose(planes[f, :npl[f]]["normal"], oplanes["normal"], atol=1e-5, rtol=0), f
        assert np.allclose(planes[f, :npl[f]]["d"], oplanes["d"], atol=1e-5, rtol=1e-9), f
    with ThreadPoolExecutor(max_workers=os.cpu_count() or 4) as pool:
        list(pool.map(one, range(B)))


def test_pipelined_batch_calls_match_enqueue_download(drfe):
    """drfe_orb_extract_batch / drfe_cape_process_depth_batch (chunked H2D | kernels | D2H pipeline) deliver exactly
    what enqueue + download deliver, for chunk sizes that do and do not divide the batch, and for raw u16 depth."""
    B = 12
    frames = [drfe.synth_frame(640, 480, i % 3, 20260900 + i, 1.0) for i in range(B)]
    gray = np.stack([f[0] for f in frames])
    depth = np.stack([f[1] for f in frames])
    K = frames[0][2]
    mc = float(np.float32(np.cos(np.pi / 12)))
    orb = drfe.ORBextractor(1000, 1.2, 8, 20, 7, 640, 480, max_batch=B)
    cape = drfe.CAPE(480, 640, 20, 20, False, mc, 50.0, max_batch=B)
    orb.enqueue(gray)
    rk, rd, rc = orb.download()
    cape.enqueue_depth(depth, *K)
    rseg, rpl, rnp = cape.download()
    for chunk in (0, 5, 1):
        kps, desc, cnt = orb.extract_batch(gray, chunk_frames=chunk)
        seg, planes, npl, _, _ = cape.process_depth_batch(depth, *K, chunk_frames=chunk)
        orb.finish_batch(); cape.finish_batch()
        assert np.array_equal(cnt, rc) and np.array_equal(npl, rnp) and np.array_equal(seg, rseg)
        for f in range(B):
            assert kps[f, :cnt[f]].tobytes() == rk[f, :rc[f]].tobytes() and np.array_equal(desc[f, :cnt[f]], rd[f, :rc[f]])
            assert all(np.array_equal(planes[f, :npl[f]][n], rpl[f, :rnp[f]][n]) for n in ("normal", "d", "nr_pts", "MSE"))
    # raw sensor depth: u16 * (1/5000), converted on the device like Frame.cc:113-115
    q = np.rint(depth * 5000).astype(np.uint16)
    fac = np.float32(1.0 / 5000.0)
    assert np.array_equal(q.astype(np.float32) * fac, depth), "the synthetic depth is quantised like a TUM png"
    seg, planes, npl, _, _ = cape.process_depth_batch(q, *K, depth_factor=float(fac), chunk_frames=4)
    cape.finish_batch()
    assert np.array_equal(seg, rseg) and np.array_equal(npl, rnp)
    cape.enqueue_depth_u16(q, float(fac), *K)
    seg2, _, npl2 = cape.download()
    assert np.array_equal(seg2, rseg) and np.array_equal(npl2, rnp)
    with pytest.raises(drfe.DrfeError):
        orb.finish_batch()          # nothing in flight


def test_later_shards_of_the_sequence(drfe, orc):
    """the frames ranks 1..7 of the 8-GPU run get (sequence indices 256..2047): device capacities hold (candidate
    density grows on the small levels) and spot-checked frames match the oracle"""
    idx = list(range(256, 2048, 28))
    data = [drfe.synth_frame(640, 480, (i // 64) % 3, 20260000 + i) for i in idx]
    gray = np.stack([d[0] for d in data]); depth = np.stack([d[1] for d in data]); K = data[0][2]
    n = len(idx)
    orb = drfe.ORBextractor(1000, 1.2, 8, 20, 7, 640, 480, max_batch=n)
    cape = drfe.CAPE(480, 640, 20, 20, False, MC, 50.0, max_batch=n)
    orb.enqueue(gray); cape.enqueue_depth(depth, *K)
    kps, desc, cnt = orb.download()
    seg, planes, npl = cape.download()
    assert cnt.min() >= 1000 and npl.min() >= 1
    oo, oc = orc.OrbOracle(1000), orc.CapeOracle(480, 640, 20, 20, False, MC, 50.0)
    for f in range(0, n, 9):
        rk, rd = oo.extract(gray[f])
        assert cnt[f] == len(rk) and kps[f, :cnt[f]].tobytes() == rk.tobytes() and np.array_equal(desc[f, :cnt[f]], rd)
        oseg, opl = oc.process(oc.depth_to_cloud(depth[f], *K))
        assert np.array_equal(seg[f], oseg) and npl[f] == len(opl)


def test_default_chunk_schedule_on_a_long_batch(drfe):
    """batches of 128 frames and more use the ramped chunk schedule (8, 16, 32 ... 32, 16, 8) by default: same results as
    enqueue + download, frame by frame, for a batch whose middle part is not a multiple of 32"""
    n = 136
    base = [drfe.synth_frame(640, 480, i % 3, 20260950 + i, 1.0) for i in range(8)]
    order = [(7 * i + i // 8) % 8 for i in range(n)]
    gray = np.stack([base[j][0] for j in order])
    depth = np.stack([base[j][1] for j in order])
    K = base[0][2]
    orb = drfe.ORBextractor(1000, 1.2, 8, 20, 7, 640, 480, max_batch=n)
    cape = drfe.CAPE(480, 640, 20, 20, False, MC, 50.0, max_batch=n)
    orb.enqueue(gray)
    rk, rd, rc = orb.download()
    cape.enqueue_depth(depth, *K)
    rseg, rpl, rnp = cape.download()
    kps, desc, cnt = orb.extract_batch(gray)
    seg, planes, npl, _, _ = cape.process_depth_batch(depth, *K)
    orb.finish_batch(); cape.finish_batch()
    assert np.array_equal(cnt, rc) and np.array_equal(npl, rnp) and np.array_equal(seg, rseg)
    for f in range(n):
        assert kps[f, :cnt[f]].tobytes() == rk[f, :rc[f]].tobytes() and np.array_equal(desc[f, :cnt[f]], rd[f, :rc[f]])
        assert all(np.array_equal(planes[f, :npl[f]][nm], rpl[f, :rnp[f]][nm]) for nm in ("normal", "d", "nr_pts", "MSE"))
    first = {j: order.index(j) for j in range(8)}
    for f in range(n):                                   # repeated inputs give repeated outputs wherever they sit in a chunk
        g = first[order[f]]
        assert cnt[f] == cnt[g] and np.array_equal(desc[f, :cnt[f]], desc[g, :cnt[g]]) and np.array_equal(seg[f], seg[g])
