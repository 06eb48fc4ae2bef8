"""The oracle of the whole-function matchers exists twice — literal Python (oracle/oracle.py) and literal C++ with the
reference's containers (oracle/match_oracle.cpp), written independently from the reference source (src/ORBmatcher.cc:46-130,
1396-1535).  They must agree value for value on populated, colliding inputs; the device path is tested against the Python
one (tests/test_search_*.py), so a slip in either restatement shows up here."""
import numpy as np
import pytest

import test_search_last_frame as tlf
import test_search_projection as tsp


@pytest.mark.parametrize("mode,check,th,occupied", [(0, 1, 15.0, True), (1, 1, 7.0, False), (2, 0, 30.0, True)])
def test_last_frame_python_and_cpp_restatements_agree(drfe, orc, mode, check, th, occupied):
    gray, depth, p, ku, ur, kd, gc, gi, desc, sf = tlf.current_frame(drfe, orc, 20260470 + mode, scene=mode)
    rng = np.random.default_rng(20 + mode)
    Tcw = tlf.small_pose(rng)
    pts, pd = tlf.make_last_frame(orc, p, ku, kd, desc, Tcw, 1100, 30 + mode)
    occ = (rng.random(len(ku)) < 0.05).astype(np.uint8) if occupied else None
    a = orc.search_last_frame(p, sf, ku, ur, gc, gi, desc, Tcw.ravel(), th, mode, check, pts, pd, occ)
    b = orc.search_last_frame_cpp(p, sf, ku, ur, gc, gi, desc, Tcw.ravel(), th, mode, check, pts, pd, occ)
    assert np.array_equal(a[0], b[0]) and np.array_equal(a[1], b[1]) and np.array_equal(a[2], b[2]) and a[3] == b[3]
    assert (a[0] >= 0).sum() > 400


@pytest.mark.parametrize("nnratio,occupied", [(0.8, True), (0.6, False)])
def test_local_points_python_and_cpp_restatements_agree(drfe, orc, nnratio, occupied):
    gray, depth, p, ku, ur, gc, gi, desc, sf = tsp.frame_inputs(drfe, orc, 20260480, scene=2)
    n = len(ku)
    q, qd = tsp.make_queries(drfe, ku, ur, desc, n, 1300, 9, sf)
    tsp.plant_block(orc, p, ku, ur, gc, gi, desc, q, qd)
    fl = tsp.local_flags(orc, len(q), 12)
    occ = (np.random.default_rng(2).random(n) < 0.1).astype(np.uint8) if occupied else None
    a = orc.search_local_points(p, ku, ur, gc, gi, desc, q, qd, fl, nnratio, occ)
    b = orc.search_local_points_cpp(p, ku, ur, gc, gi, desc, q, qd, fl, nnratio, occ)
    assert a[0].tobytes() == b[0].tobytes() and np.array_equal(a[1], b[1]) and np.array_equal(a[2], b[2]) and a[3] == b[3]
    assert a[3] > 300


import test_bow  # noqa: E402


@pytest.mark.parametrize("shape", [dict(k=10, L=4, seed=81), dict(k=6, L=5, seed=82, ragged=True), dict(k=8, L=3, seed=83, scoring=1),
                                   dict(k=8, L=3, seed=84, scoring=5), dict(k=8, L=3, seed=85, weighting=2)])
def test_bow_python_and_cpp_restatements_agree(drfe, orc, shape):
    voc = orc.synth_vocabulary(shape["k"], shape["L"], shape["seed"], scoring=shape.get("scoring", 0), weighting=shape.get("weighting", 0),
                               ragged=shape.get("ragged", False))
    _, desc = test_bow.frame_descriptors(drfe, orc, 20260530 + shape["seed"])
    for levelsup in (1, 2, 4):
        a = orc.Vocabulary(**voc).transform(desc, levelsup)
        b = orc.bow_transform_cpp(voc, desc, levelsup)
        assert np.array_equal(a[0], b[0]) and np.array_equal(a[1], b[1])
        assert [k for k, _ in a[2]] == [k for k, _ in b[2]]
        assert np.array([v for _, v in a[2]]).tobytes() == np.array([v for _, v in b[2]]).tobytes()      # the doubles, bit for bit
        assert a[3] == b[3]
    assert len(a[2]) > 50


@pytest.mark.parametrize("nnratio,check", [(0.7, True), (0.9, False)])
def test_search_by_bow_python_and_cpp_restatements_agree(drfe, orc, nnratio, check):
    gray, desc = test_bow.frame_descriptors(drfe, orc, 20260540)
    keys, _ = orc.OrbOracle(1000).extract(gray)
    V = orc.Vocabulary(**orc.synth_vocabulary(10, 4, 91))
    kd, ka, valid = test_bow.make_keyframe(desc, keys["angle"], 5)
    f_fv, kf_fv = V.transform(desc, 2)[3], V.transform(kd, 2)[3]
    a = orc.search_by_bow(kd, ka, valid, kf_fv, desc, keys["angle"], f_fv, nnratio, check)
    b = orc.search_by_bow_cpp(kd, ka, valid, kf_fv, desc, keys["angle"], f_fv, nnratio, check)
    assert np.array_equal(a[0], b[0]) and np.array_equal(a[1], b[1]) and a[2] == b[2] and a[2] > 150
