#!/usr/bin/env python
"""Wall time of the optional calls after a 256-frame batch (host buffers in and out, so PCIe included):
drfe_orb_frame_post, drfe_orb_search_by_projection (1000 queries per frame), drfe_cape_plane_points."""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "dr-slam_b200"))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
import drfe  # noqa: E402


def timed(fn, n=5):
    fn()
    t0 = time.perf_counter()
    for _ in range(n):
        r = fn()
    return (time.perf_counter() - t0) / n * 1e3, r


def main():
    B, W, H = 256, 640, 480
    data = [drfe.synth_frame(W, H, (i // 8) % 3, 20260000 + i) for i in range(32)]
    gray = np.stack([data[i % 32][0] for i in range(B)])
    depth = np.stack([data[i % 32][1] for i in range(B)])
    K = data[0][2]
    orb = drfe.ORBextractor(1000, 1.2, 8, 20, 7, W, H, max_batch=B)
    cape = drfe.CAPE(H, W, 20, 20, False, bench.MIN_COS, 50.0, max_batch=B)
    orb.enqueue(gray)
    kps, desc, cnt = orb.download()
    cape.enqueue_depth(depth, *K, nframes=B)
    cape.download()
    p = orb.frame_params(*K, [0.1, -0.05, 0.001, 0.0005, 0.0], 40.0)
    ms, (ku, ur, kd, gc, gi) = timed(lambda: orb.frame_post(p, depth))
    print("drfe_orb_frame_post            %7.2f ms per %d frames (depth from the host: %d MB H2D)" % (ms, B, depth.nbytes >> 20))
    rng = np.random.default_rng(1)
    Q = np.zeros((B, 1000), drfe.QUERY_DTYPE)
    src = rng.integers(0, 1000, (B, 1000))
    Q["x"] = np.take_along_axis(ku["x"], src, 1) + rng.normal(0, 3, (B, 1000)).astype(np.float32)
    Q["y"] = np.take_along_axis(ku["y"], src, 1) + rng.normal(0, 3, (B, 1000)).astype(np.float32)
    Q["r"], Q["xr"] = 7.0, Q["x"] - 20
    lv = np.take_along_axis(ku["octave"], src, 1)
    Q["min_level"], Q["max_level"] = lv - 1, lv
    QD = np.take_along_axis(desc, src[:, :, None], 1)
    ms, m = timed(lambda: orb.search_by_projection(Q, QD))
    print("drfe_orb_search_by_projection  %7.2f ms per %d frames x 1000 queries (%.1f%% matched)" % (ms, B, 100.0 * (m["best_idx"] >= 0).mean()))
    # ORBmatcher::SearchByProjection(CurrentFrame, LastFrame): 1000 last-frame points per frame, identity pose
    tp = np.zeros(B, drfe.TRACK_PARAMS_DTYPE)
    tp["Tcw"] = np.eye(3, 4, dtype=np.float32).ravel()
    tp["th"], tp["check_orientation"] = 15.0, 1
    P = np.zeros((B, 1000), drfe.LAST_POINT_DTYPE)
    z = np.where(np.take_along_axis(kd, src, 1) > 0, np.take_along_axis(kd, src, 1), 2.0)
    P["X"], P["Y"], P["Z"] = (Q["x"] - K[2]) * z / K[0], (Q["y"] - K[3]) * z / K[1], z
    P["angle"] = np.take_along_axis(ku["angle"], src, 1)
    P["octave"], P["flags"] = lv, drfe.LP_VALID | drfe.LP_OBSERVED
    ms, r = timed(lambda: orb.search_last_frame(tp, P, QD))
    print("drfe_orb_search_last_frame     %7.2f ms per %d frames x 1000 points (%.1f matches per frame, %.1f sweeps)" % (ms, B, r[3].mean(), r[4].mean()))
    ms0, _ = timed(lambda: orb.search_last_frame(tp, P, QD, np.zeros(B, np.int32)))
    print("   of which copies + launch     %7.2f ms (same call with 0 points per frame: same H2D / D2H volume)" % ms0)
    uniq = np.stack([rng.permutation(1000) for _ in range(B)])          # every point aims at its own keypoint: the tracking case
    P2, QD2 = P.copy(), np.take_along_axis(desc, uniq[:, :, None], 1)
    z2 = np.where(np.take_along_axis(kd, uniq, 1) > 0, np.take_along_axis(kd, uniq, 1), 2.0)
    P2["X"] = (np.take_along_axis(ku["x"], uniq, 1) - K[2]) * z2 / K[0]
    P2["Y"] = (np.take_along_axis(ku["y"], uniq, 1) - K[3]) * z2 / K[1]
    P2["Z"], P2["angle"], P2["octave"] = z2, np.take_along_axis(ku["angle"], uniq, 1), np.take_along_axis(ku["octave"], uniq, 1)
    ms, r = timed(lambda: orb.search_last_frame(tp, P2, QD2))
    print("drfe_orb_search_last_frame     %7.2f ms, points without collisions (%.1f matches per frame, %.1f sweeps)" % (ms, r[3].mean(), r[4].mean()))
    # Frame::ComputeBoW against a vocabulary of the ORB vocabulary's size (k = 10, L = 6: 1 111 110 nodes, 10^6 words)
    k, L = 10, 6
    sizes = [k ** l for l in range(1, L + 1)]
    first = np.cumsum([1] + sizes)                                     # node id of the first node of level l + 1
    parent = np.concatenate([np.repeat(np.arange(first[l] - (sizes[l - 1] if l else 1), first[l]), k) for l in range(L)]).astype(np.int32)
    nn = len(parent)
    leaf = np.zeros(nn, np.uint8)
    leaf[-sizes[-1]:] = 1
    vd = rng.integers(0, 256, (nn, 32), dtype=np.uint8)
    wt = np.where(leaf > 0, rng.uniform(0.5, 9.0, nn), 0.0)
    t0 = time.perf_counter()
    voc = drfe.Vocabulary(k, L, 0, 0, parent, leaf, vd, wt)
    print("drfe_vocab_create              %7.2f ms (%d nodes, %d words, %d MB on the device)" % ((time.perf_counter() - t0) * 1e3, nn, voc.words(), nn * 52 >> 20))
    L_ = orb.L
    bn, fn = np.zeros(B, np.int32), np.zeros(B, np.int32)
    bw, bv = np.zeros((B, orb.cap), np.int32), np.zeros((B, orb.cap), np.float64)
    fnode, fstart, ffeat = np.zeros((B, orb.cap), np.int32), np.zeros((B, orb.cap + 1), np.int32), np.zeros((B, orb.cap), np.int32)
    vp = lambda a: a.ctypes.data
    ms, _ = timed(lambda: L_.drfe_orb_compute_bow(orb.h, voc.h, 4, None, None, vp(bn), vp(bw), vp(bv), vp(fn), vp(fnode), vp(fstart), vp(ffeat)))
    print("drfe_orb_compute_bow           %7.2f ms per %d frames (%.0f words, %.0f nodes per frame)" % (ms, B, bn.mean(), fn.mean()))
    voc.close()
    ms, (pts, offs) = timed(lambda: cape.plane_points(B), n=3)
    print("drfe_cape_plane_points         %7.2f ms per %d frames (%d MB of points D2H, pageable destination)" % (ms, B, int(offs.max(1).sum()) * 12 >> 20))
    # the same into pinned host memory (what a caller that cares would pass)
    import ctypes as C
    import torch
    N = W * H
    pin_pts = torch.empty((B, N, 3), dtype=torch.float32).pin_memory().numpy()
    pin_off = torch.empty((B, 256), dtype=torch.int32).pin_memory().numpy()
    pin_depth = torch.from_numpy(depth).pin_memory().numpy()

    def pp():
        drfe._check(cape.L.drfe_cape_plane_points(cape.h, pin_pts.ctypes.data, N, pin_off.ctypes.data, 255))
    ms, _ = timed(pp, n=3)
    print("drfe_cape_plane_points         %7.2f ms per %d frames (pinned destination)" % (ms, B))
    ms, _ = timed(lambda: orb.frame_post(p, pin_depth))
    print("drfe_orb_frame_post            %7.2f ms per %d frames (pinned depth)" % (ms, B))
    import torch as _t
    d_depth = _t.from_numpy(depth).cuda()
    ms, _ = timed(lambda: orb.frame_post(p, d_depth.data_ptr(), mem_kind=drfe.MEM_DEVICE, row_stride=W, frame_stride=W * H))
    print("drfe_orb_frame_post            %7.2f ms per %d frames (depth already on the device)" % (ms, B))

    ms, _ = timed(lambda: orb.frame_post_shared_depth(p, cape))
    print("drfe_orb_frame_post_shared_depth %5.2f ms per %d frames (the depth the CAPE handle uploaded)" % (ms, B))
    # round 2: the 5 cm voxel filter of the per-plane lists, the 1/3-resolution cloud, the input resize
    pin_vox = _t.empty((B, N, 3), dtype=_t.float32).pin_memory().numpy()

    def vox():
        drfe._check(cape.L.drfe_cape_plane_points_voxel(cape.h, 0.05, pin_vox.ctypes.data, N, pin_off.ctypes.data, 255))
    ms, _ = timed(vox, n=3)
    nvox = sum(int(pin_off[f, 255]) for f in range(B))
    print("drfe_cape_plane_points_voxel   %7.2f ms per %d frames (pinned destination; %d centroids = %.1f MB instead of the lists)" % (ms, B, nvox, nvox * 12 / 1e6))
    ms, _ = timed(lambda: cape.third_cloud(3.0, B), n=3)
    print("drfe_cape_third_cloud          %7.2f ms per %d frames" % (ms, B))
    ms, (tc, tn) = timed(lambda: cape.third_cloud_normals(10.0, nframes=B), n=3)
    print("drfe_cape_third_cloud_normals  %7.2f ms per %d frames (PCL integral-image normals on the 214 x 160 cloud; %.0f %% of the points get a normal)"
          % (ms, B, 100.0 * (~np.isnan(tn[..., 0])).mean()))
    rs = drfe.Resizer(848, 480, 640, 480, max_batch=32)
    rgb = np.random.default_rng(2).integers(0, 256, (32, 480, 848, 3), dtype=np.uint8)
    d16 = np.random.default_rng(3).integers(0, 65536, (32, 480, 848), dtype=np.uint16)
    ms, _ = timed(lambda: rs(rgb), n=5)
    ms2, _ = timed(lambda: rs(d16), n=5)
    print("drfe_resize 848x480 -> 640x480  %7.2f ms per 32 RGB frames, %.2f ms per 32 16-bit depth maps (pageable host in / out)" % (ms, ms2))
    # PEAC-AHC (the plane extractor of Frame::Frame): one CTA per frame, so the batch is what fills the GPU
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    from test_peac import clean_depth
    from oracle import oracle as orc
    fac = float(np.float32(1.0 / 5000.0))
    q8 = np.stack([np.rint(clean_depth(data[i][1], ((100, 140, 300, 420),)) * 5000).astype(np.uint16) for i in range(0, 32, 4)])
    qB = np.ascontiguousarray(q8[np.arange(B) % 8])
    pe = drfe.PEAC(W, H, max_batch=B)

    def peac_run():
        pe.enqueue(qB, fac, *K)
        return pe.download()
    ms, (seg, planes, npl) = timed(peac_run, n=3)
    print("drfe_peac (enqueue + download)  %7.2f ms per %d frames = %.3f ms per frame (%.1f planes per frame; pageable host in / out)" % (ms, B, ms / B, npl.mean()))
    d_q = _t.from_numpy(qB.view(np.int16)).cuda()

    def peac_dev():
        pe.enqueue(d_q.data_ptr(), fac, *K, nframes=B, mem_kind=drfe.MEM_DEVICE)
        pe.sync()
    msd, _ = timed(peac_dev, n=3)
    tot = np.array([pe.counters(f)[10] for f in range(B)])
    print("drfe_peac, depth on the device, no download: %7.2f ms per %d frames = %.3f ms per frame; k_peac_frame kilocycles per frame min %d median %d max %d"
          % (msd, B, msd / B, tot.min(), np.median(tot), tot.max()))
    pin_pv = _t.empty((B, N, 3), dtype=_t.float32).pin_memory().numpy()
    pin_po = _t.empty((B, 256), dtype=_t.int32).pin_memory().numpy()

    def peac_vox():
        drfe._check(pe.L.drfe_peac_plane_points_voxel(pe.h, 3.0, 0.05, pin_pv.ctypes.data, N, pin_po.ctypes.data, 255))
    msv, _ = timed(peac_vox, n=3)
    print("drfe_peac_plane_points_voxel   %7.2f ms per %d frames (z <= 3 m, 5 cm leaf, pinned destination; %d centroids)"
          % (msv, B, sum(int(pin_po[f, 255]) for f in range(B))))
    pe1 = drfe.PEAC(W, H)

    def peac_one():
        pe1.enqueue(qB[:1], fac, *K)
        return pe1.download()
    ms1, _ = timed(peac_one, n=10)
    c = pe1.counters(0)
    names = ["reset", "edges", "cluster 1", "membership + seeds", "region growing", "cluster 2", "outputs"]
    kc = np.diff(np.concatenate([[0], c[4:11]]))
    print("   k_peac_frame stages (kilocycles): " + ", ".join("%s %d" % (n, v) for n, v in zip(names, kc)) + "; steps %d, queue %d" % (c[0], c[1]))
    t0 = time.perf_counter()
    for i in range(4):
        orc.peac_run(orc.peac_cloud(q8[i], fac, *K), W, H)
    cpu = (time.perf_counter() - t0) / 4 * 1e3
    print("drfe_peac one frame             %7.2f ms (latency);  CPU restatement, one core: %.2f ms per frame" % (ms1, cpu))
    # the per-frame sequence Tracking runs after the two extractors, on one frame (latency, host in / host out)
    orb1 = drfe.ORBextractor(1000, 1.2, 8, 20, 7, W, H)
    cape1 = drfe.CAPE(H, W, 20, 20, False, bench.MIN_COS, 50.0)
    voc = drfe.Vocabulary(k, L, 0, 0, parent, leaf, vd, wt)
    tp1, P1, QD1 = tp[:1].copy(), P2[:1].copy(), QD2[:1].copy()

    def one_frame():
        orb1(gray[0], None)
        cape1.process_depth(depth[0], *K)
        orb1.frame_post_shared_depth(p, cape1)
        orb1.compute_bow(voc)
        return orb1.search_last_frame(tp1, P1, QD1)
    ms, r = timed(one_frame, n=50)
    print("one frame: extract + planes + frame_post + ComputeBoW + SearchByProjection(last frame)  %.3f ms (%d matches)" % (ms, r[3][0]))


if __name__ == "__main__":
    main()
