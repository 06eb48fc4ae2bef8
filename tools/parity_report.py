"""Stage-by-stage parity report: CUDA path (libdrfe.so via the C ABI) vs the CPU oracle on
seeded synthetic frames.  Diagnostic tool; the pass/fail gates live in tests/."""
import argparse
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "dr-slam_b200"))
sys.path.insert(0, ROOT)
import drfe  # noqa: E402
from oracle import oracle as orc  # noqa: E402


def sort_rows(a):
    return a[np.lexsort((a[:, 2], a[:, 0], a[:, 1]))] if len(a) else a


def orb_report(w, h, nfeat, frames, batch):
    ex = drfe.ORBextractor(nfeat, 1.2, 8, 20, 7, w, h, max_batch=batch)
    o = orc.OrbOracle(nfeat)
    grays = np.stack([drfe.synth_frame(w, h, sc, seed)[0] for sc, seed in frames])
    t = time.time()
    ex.enqueue(grays)
    kps, desc, counts = ex.download()
    print("GPU batch of %d: %.1f ms (first call)" % (len(frames), (time.time() - t) * 1e3))
    ok = True
    for f in range(len(frames)):
        o.run(grays[f])
        for l in range(8):
            a, b = ex.pyramid(f, l, True), o.level(l, True)
            if not np.array_equal(a, b):
                ok = False
                print(" frame %d level %d pyramid mismatch: %d px, interior %d" % (
                    f, l, (a != b).sum(), (a[19:-19, 19:-19] != b[19:-19, 19:-19]).sum()))
            ca, cb = sort_rows(ex.candidates(f, l)), sort_rows(o.candidates(l))
            if ca.shape != cb.shape or not np.array_equal(ca, cb):
                ok = False
                sa, sb = set(map(tuple, ca)), set(map(tuple, cb))
                print(" frame %d level %d candidates: gpu %d oracle %d common %d" % (f, l, len(ca), len(cb), len(sa & sb)))
                print("    only gpu:", sorted(sa - sb)[:5], " only oracle:", sorted(sb - sa)[:5])
            ka, kb = ex.level_keypoints(f, l), o.level_keypoints(l)
            if len(ka) != len(kb) or not all(np.array_equal(ka[n], kb[n]) for n in ("x", "y", "response")):
                ok = False
                sa = set(zip(ka["x"], ka["y"], ka["response"]))
                sb = set(zip(kb["x"], kb["y"], kb["response"]))
                print(" frame %d level %d quadtree: gpu %d oracle %d common %d same-order %s tie %d" % (
                    f, l, len(ka), len(kb), len(sa & sb), False, o.level_tie(l)))
            elif not np.array_equal(ka["angle"], kb["angle"]):
                ok = False
                d = np.abs(ka["angle"] - kb["angle"])
                print(" frame %d level %d angles differ: max %.3g deg on %d kps" % (f, l, d.max(), (d > 0).sum()))
            if len(kb):
                ba, bb = ex.blurred(f, l), o.blurred(l)
                if not np.array_equal(ba, bb):
                    ok = False
                    print(" frame %d level %d blur mismatch: %d px" % (f, l, (ba != bb).sum()))
        rk, rd = o.result()
        n = counts[f]
        if n != len(rk):
            ok = False
            print(" frame %d: %d keypoints vs oracle %d" % (f, n, len(rk)))
            continue
        same = all(np.array_equal(kps[f, :n][nm], rk[nm]) for nm in KPF)
        ham = np.unpackbits(desc[f, :n] ^ rd, axis=1).sum(1)
        print(" frame %d: n=%d keypoint fields identical=%s; descriptors identical on %.2f%% (max Hamming %d)" % (
            f, n, same, (ham == 0).mean() * 100, ham.max() if n else 0))
        ok &= same
    print("ORB %dx%d: %s" % (w, h, "ALL STAGES MATCH" if ok else "MISMATCHES"))
    return ok


KPF = ("x", "y", "size", "angle", "response", "octave", "class_id")


def cape_report(w, h, frames, unit, mmd):
    mc = float(np.float32(np.cos(np.pi / 12)))
    cp = drfe.CAPE(h, w, 20, 20, False, mc, mmd, max_batch=len(frames))
    o = orc.CapeOracle(h, w, 20, 20, False, mc, mmd)
    data = [drfe.synth_frame(w, h, sc, seed, unit) for sc, seed in frames]
    depth = np.stack([d[1] for d in data])
    K = data[0][2]
    cp.enqueue_depth(depth, *K)
    seg, planes, npl = cp.download()
    ok = True
    for f in range(len(frames)):
        cloud = o.depth_to_cloud(depth[f], *K)
        oseg, oplanes = o.process(cloud)
        gc = cp.cloud(f)
        if not np.array_equal(gc, cloud):
            ok = False
            print(" frame %d cloud mismatch %d" % (f, (gc != cloud).sum()))
        cg, co = cp.cells(f), o.cells()
        bad = [n for n in cg.dtype.names if not np.array_equal(cg[n], co[n])]
        if bad:
            ok = False
            print(" frame %d cell fields differ: %s" % (f, bad))
            for n in bad[:3]:
                idx = np.flatnonzero((cg[n] != co[n]).reshape(len(cg), -1).any(1))
                print("    %s: %d cells, e.g. cell %d gpu %s oracle %s" % (n, len(idx), idx[0], cg[n][idx[0]], co[n][idx[0]]))
        pm, em = cp.grid_maps(f)
        opm, oem = o.grid_maps()
        if not np.array_equal(pm, opm) or not np.array_equal(em, oem):
            ok = False
            print(" frame %d grid maps differ: plane_map %d eroded %d" % (f, (pm != opm).sum(), (em != oem).sum()))
        if npl[f] != len(oplanes):
            ok = False
            print(" frame %d planes %d vs %d" % (f, npl[f], len(oplanes)))
        else:
            bad = [n for n in oplanes.dtype.names if not np.array_equal(planes[f, :npl[f]][n], oplanes[n])]
            if bad:
                ok = False
                print(" frame %d plane fields differ: %s" % (f, bad))
        ns = (seg[f] != oseg).sum()
        if ns:
            ok = False
        print(" frame %d: planes=%d seg mismatches=%d labelled=%.1f%%" % (f, npl[f], ns, (oseg > 0).mean() * 100))
    print("CAPE %dx%d unit %g: %s" % (w, h, unit, "ALL STAGES MATCH" if ok else "MISMATCHES"))
    return ok


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--quick", action="store_true")
    a = ap.parse_args()
    print(drfe.lib().drfe_version().decode(), "devices:", drfe.device_count())
    fr = [(0, 20260000), (1, 20260077), (0, 20260200), (2, 20260031)]
    ok = orb_report(640, 480, 1000, fr[:2] if a.quick else fr, 4)
    ok &= cape_report(640, 480, fr[:2] if a.quick else fr, 1.0, 50.0)
    if not a.quick:
        ok &= orb_report(320, 240, 500, fr[:2], 2)
        ok &= orb_report(1280, 720, 2000, fr[2:], 2)
        ok &= cape_report(640, 480, fr, 1000.0, 50.0)
        ok &= cape_report(1280, 720, fr[2:], 1000.0, 900.0)
    print("PARITY REPORT:", "OK" if ok else "FAILED")
    sys.exit(0 if ok else 1)
