#!/usr/bin/env python
"""Determinism soak: the same batch through ORB + CAPE (two streams, as bench.py runs them) and through PEAC, many times; every
download must be byte-identical to the first one.  Catches rare races the single-shot parity tests can miss (dependent launches,
stream priorities, shared-memory atomics).   python tools/soak.py [--iters 200] [--frames 256 32 1]"""
import argparse
import hashlib
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "dr-slam_b200"))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
import drfe  # noqa: E402


def digest(*arrays):
    h = hashlib.blake2b(digest_size=16)
    for a in arrays:
        h.update(np.ascontiguousarray(a).tobytes())
    return h.hexdigest()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--iters", type=int, default=200)
    ap.add_argument("--frames", type=int, nargs="*", default=[256, 32, 1])
    a = ap.parse_args()
    import torch
    wl = bench.Workload("c640")
    gray, depth, K = bench.make_sequence(wl, 0, wl.batch, 8)
    rig = bench.Rig(drfe, torch, wl, gray, depth, K, 0)
    bad = 0
    for n in a.frames:
        first = None
        for it in range(a.iters):
            rig.step_resident(n)
            kps, desc, cnt = rig.orb.download()
            seg, planes, npl = rig.cape.download()
            valid = np.arange(kps.shape[1])[None, :] < cnt[:, None]           # entries past a frame's count are not written
            pv = np.arange(planes.shape[1])[None, :] < npl[:, None]
            pl = planes[pv]                                                  # named fields only: the record's padding bytes are not written
            parts = dict(cnt=cnt, kps=kps[valid], desc=desc[valid], seg=seg, npl=npl, planes=np.concatenate([pl[nm].reshape(len(pl), -1).astype(np.float64) for nm in pl.dtype.names], axis=1))
            d = digest(*parts.values())
            dd = {k: digest(v) for k, v in parts.items()}
            if first is not None and d != first:
                print("   differing outputs:", [k for k in dd if dd[k] != first_parts[k]])
                if dd["seg"] != first_parts["seg"]:
                    fr = [f for f in range(n) if not np.array_equal(seg[f], first_arrays["seg"][f])]
                    print("   seg differs in frames", fr[:10], "pixels", int((seg != first_arrays["seg"]).sum()))
                if dd["kps"] != first_parts["kps"] or dd["cnt"] != first_parts["cnt"]:
                    print("   counts", cnt[:8], first_arrays["cnt"][:8])
            if first is None:
                first_parts, first_arrays = dd, dict(seg=seg.copy(), cnt=cnt.copy())
            if first is None:
                first = d
            elif d != first:
                bad += 1
                print("MISMATCH: ORB + CAPE, %d frames, iteration %d" % (n, it))
                break
        print("ORB + CAPE, %3d frames per step: %d identical downloads" % (n, a.iters if bad == 0 else it))
    # the chunk-pipelined batch calls (host buffers in and out, three streams per handle), against the resident path's result
    B, W, H = wl.batch, wl.W, wl.H
    rig.step_resident(B)
    kps0, desc0, cnt0 = rig.orb.download()
    seg0, planes0, npl0 = rig.cape.download()
    hk = drfe.host_array(kps0.shape, kps0.dtype); hd = drfe.host_array(desc0.shape, desc0.dtype); hc = drfe.host_array(cnt0.shape, cnt0.dtype)
    hs = drfe.host_array(seg0.shape, seg0.dtype); hp = drfe.host_array(planes0.shape, planes0.dtype); hn = drfe.host_array(npl0.shape, npl0.dtype)
    hg = drfe.host_array(gray.shape, gray.dtype); hg[...] = gray
    hz = drfe.host_array(depth.shape, depth.dtype); hz[...] = depth
    valid0 = np.arange(kps0.shape[1])[None, :] < cnt0[:, None]
    iters = max(5, a.iters // 3)
    for it in range(iters):
        rig.orb.extract_batch(hg, hk, hd, hc)
        rig.cape.process_depth_batch(hz, *K, seg=hs, planes=hp, nplanes=hn)
        rig.orb.finish_batch(); rig.cape.finish_batch()
        same = (np.array_equal(hc, cnt0) and hk[valid0].tobytes() == kps0[valid0].tobytes() and np.array_equal(hd[valid0], desc0[valid0])
                and np.array_equal(hs, seg0) and np.array_equal(hn, npl0))
        if not same:
            bad += 1
            print("MISMATCH: batch calls, iteration %d" % it)
            break
    print("batch calls (drfe_orb_extract_batch + drfe_cape_process_depth_batch), %d frames: %d results identical to the resident path's" % (B, iters))
    q = np.rint(depth[:64] * np.float32(5000.0)).astype(np.uint16)
    col = np.where(q > 0, np.arange(wl.W)[None, None, :], 0)
    q = np.take_along_axis(q, np.maximum.accumulate(col, axis=2), axis=2)
    pe = drfe.PEAC(wl.W, wl.H, max_batch=64)
    fac = float(np.float32(1.0) / np.float32(5000.0))
    first = None
    iters = max(5, a.iters // 10)
    for it in range(iters):
        pe.enqueue(q, fac, *K)
        seg, planes, npl = pe.download()
        idx, pts, offs = pe.plane_vertices()
        d = digest(seg, planes, npl, offs, idx, pts)
        if first is None:
            first = d
        elif d != first:
            bad += 1
            print("MISMATCH: PEAC, iteration %d" % it)
            break
    print("PEAC, 64 frames: %d identical downloads" % iters)
    sys.exit(1 if bad else 0)


if __name__ == "__main__":
    main()
