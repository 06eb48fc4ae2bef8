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
