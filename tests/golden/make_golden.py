"""Generates the committed golden fixtures in tests/golden/*.npz.

Run in a container where cv2 is importable:   python tests/golden/make_golden.py
The expected outputs come from oracle/py_ref.py — the restatement that calls the REAL OpenCV
(cv2.resize / copyMakeBorder / FAST / GaussianBlur / fastAtan2 / erode / dilate) for every
OpenCV-owned primitive and numpy/LAPACK for the eigen-solve — NOT from the C++ oracle and not
from the CUDA path, so the fixtures pin both.  Inputs are stored inside the fixtures, so the
tests do not depend on the synthetic generator staying byte-stable.
"""
import os
import sys
import zlib

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, os.path.join(ROOT, "dr-slam_b200"))
sys.path.insert(0, ROOT)
import drfe  # noqa: E402  (only for the synthetic generator, host code)
from oracle import py_ref  # noqa: E402


def crc(a):
    return np.uint32(zlib.crc32(np.ascontiguousarray(a).tobytes()))


def orb_fixture(name, w, h, scene, seed, nfeatures):
    gray, _, _ = drfe.synth_frame(w, h, scene, seed)
    ref = py_ref.OrbRef(nfeatures, 1.2, 8, 20, 7)
    R = ref.extract(gray, keep_intermediates=True)
    out = dict(gray=gray, nfeatures=np.int32(nfeatures), kps=R["kps"], desc=R["desc"],
               per_level=np.array(ref.per_level, np.int32), umax=np.array(ref.umax, np.int32),
               scale=np.array(ref.scale, np.float32))
    for l in range(8):
        out["pyr_crc_%d" % l] = crc(R["bordered"][l])
        out["blur_crc_%d" % l] = crc(R["blurred"][l]) if R["blurred"][l] is not None else np.uint32(0)
        out["cands_%d" % l] = np.array(R["cands"][l], np.float32).reshape(-1, 3).astype(np.int16)
        lk = R["level_kps"][l]
        out["lkp_%d" % l] = np.array([(k["x"], k["y"], k["response"], k["angle"]) for k in lk], np.float32).reshape(-1, 4)
    out["pyr_level3"] = R["bordered"][3]      # one full level image, for a direct byte compare
    np.savez_compressed(os.path.join(HERE, name), **out)
    print(name, "kps", len(R["kps"]), "cands", [len(c) for c in R["cands"]])


def cape_fixture(name, w, h, scene, seed, unit, cell, max_merge, cylinder=False):
    _, depth, K = drfe.synth_frame(w, h, scene, seed, unit)
    mc = float(np.float32(np.cos(np.pi / 12)))
    cloud = py_ref.depth_to_cloud(depth, *K, cell, cell)
    res = py_ref.cape_process(cloud, h, w, cell, cell, mc, max_merge, cylinder=cylinder)
    seg, final, grid, pmap, emap = res[:5]
    q = np.rint(depth / np.float32(unit) * 5000).astype(np.uint16)
    assert np.array_equal((q.astype(np.float32) * np.float32(1.0 / 5000.0) * np.float32(unit)), depth)
    sums = np.array([[getattr(s, f) for f in py_ref.Seg.FIELDS] for s in grid], np.float64)
    out = dict(depth_q=q, unit=np.float32(unit), K=np.array(K, np.float32), cell=np.int32(cell),
               min_cos=np.float32(mc), max_merge=np.float32(max_merge), cloud_crc=crc(cloud),
               seg=seg, plane_map=pmap, eroded_map=emap,
               cell_planar=np.array([s.planar for s in grid], np.uint8),
               cell_nr_pts=np.array([s.nr_pts for s in grid], np.int32), cell_sums=sums,
               cell_normal=np.array([s.normal for s in grid]), cell_d=np.array([s.d for s in grid]),
               cell_mse=np.array([s.MSE for s in grid], np.float32),
               plane_normal=np.array([p.normal for p in final]).reshape(-1, 3),
               plane_d=np.array([p.d for p in final]), plane_nr_pts=np.array([p.nr_pts for p in final], np.int32),
               plane_mse=np.array([p.MSE for p in final], np.float32),
               plane_score=np.array([p.score for p in final], np.float32))
    if cylinder:
        c = res[5]
        out.update(cylinder=np.int32(1), nr_cylinders_final=np.int32(c["nr_cylinders_final"]), cyl_map=c["cyl_map"],
                   cyl_eroded=c["cyl_eroded"], cyl_radius=c["radius"], cyl_center=c["center"], cyl_axis=c["axis"],
                   cyl_mse=c["mse"])
    np.savez_compressed(os.path.join(HERE, name), **out)
    print(name, "planes", len(final), "planar cells", int(out["cell_planar"].sum()),
          ("cylinders %d (final %d)" % (len(res[5]["radius"]), res[5]["nr_cylinders_final"])) if cylinder else "")


if __name__ == "__main__":
    orb_fixture("orb_320x240_corridor.npz", 320, 240, 0, 20260005, 500)
    orb_fixture("orb_640x480_room.npz", 640, 480, 1, 20260077, 1000)
    cape_fixture("cape_640x480_corridor_m.npz", 640, 480, 0, 20260000, 1.0, 20, 50.0)
    cape_fixture("cape_640x480_room_mm.npz", 640, 480, 2, 20260100, 1000.0, 20, 50.0)
    cape_fixture("cape_320x240_room_mm_cell10.npz", 320, 240, 1, 20260033, 1000.0, 10, 50.0)
    # cylinder detection on (CylinderSeg.cpp; declared glibc rand() stream, seed 1 per frame)
    cape_fixture("cape_640x480_pillars_mm_cyl.npz", 640, 480, 2, 20260100, 1000.0, 20, 50.0, cylinder=True)
    cape_fixture("cape_640x480_pillars_m_cyl.npz", 640, 480, 2, 20260300, 1.0, 20, 50.0, cylinder=True)
