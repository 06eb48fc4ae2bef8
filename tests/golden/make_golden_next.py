"""Generates tests/golden/next_640x480_room.npz: golden vectors of the "next" rows — the per-frame steps after
extraction, Frame::ComputeBoW and the three whole-function matchers — from the literal Python restatements in
oracle/oracle.py (each follows the reference loop by loop; tests/test_search_*.py and tests/test_bow.py check them against
brute force and cv2).  Inputs are stored inside the fixture (frame, depth, vocabulary, queries), so the tests do not depend on
the synthetic generators staying byte-stable.   python tests/golden/make_golden_next.py"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, os.path.join(ROOT, "dr-slam_b200"))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import drfe  # noqa: E402  (synthetic generator only, host code)
from oracle import oracle as orc  # noqa: E402
import test_bow  # noqa: E402
import test_search_last_frame as tlf  # noqa: E402
import test_search_projection as tsp  # noqa: E402

K = (525.0, 525.0, 319.5, 239.5)
DIST = [0.1, -0.05, 0.001, 0.0005, 0.0]


def flat_fv(fv):
    return (np.array([k for k, _ in fv], np.int32), np.cumsum([0] + [len(l) for _, l in fv]).astype(np.int32),
            np.array([i for _, l in fv for i in l], np.int32))


def main():
    gray, depth, _ = drfe.synth_frame(640, 480, 1, 20260042)
    o = orc.OrbOracle(1000)
    keys, desc = o.extract(gray)
    sf = np.array(o.scale_factors(), np.float32)
    p = orc.frame_params(*K, DIST, 40.0, 640, 480)
    ku, ur, kd, gc, gi = orc.frame_post(p, keys, depth)
    out = dict(gray=gray, depth=depth, K=np.array(K, np.float64), dist=np.array(DIST, np.float64), bf=np.float64(40.0),
               keys=keys, desc=desc, keys_un=ku, u_right=ur, kp_depth=kd, grid_count=gc, grid_index=gi)
    # Frame::ComputeBoW
    voc = orc.synth_vocabulary(10, 4, 41)
    words, nodes, bow, fv = orc.Vocabulary(**voc).transform(desc, 2)
    out.update({"voc_" + k: np.asarray(v) for k, v in voc.items()})
    out.update(bow_words=words, bow_nodes=nodes, bow_key=np.array([k for k, _ in bow], np.int32), bow_value=np.array([v for _, v in bow], np.float64))
    out["fv_node"], out["fv_start"], out["fv_feat"] = flat_fv(fv)
    # SearchByProjection(CurrentFrame, LastFrame)
    rng = np.random.default_rng(7)
    Tcw = tlf.small_pose(rng)
    pts, pd = tlf.make_last_frame(orc, p, ku, kd, desc, Tcw, 900, 11)
    occ = (rng.random(len(ku)) < 0.05).astype(np.uint8)
    mk, md, holder, nm = orc.search_last_frame(p, sf, ku, ur, gc, gi, desc, Tcw.ravel(), 15.0, 0, 1, pts, pd, occ)
    out.update(lf_Tcw=Tcw, lf_points=pts, lf_desc=pd, lf_occupied=occ, lf_match_key=mk, lf_match_dist=md, lf_key_point=holder, lf_nmatches=np.int32(nm))
    # SearchByProjection(F, vpMapPoints, th) whole
    q, qd = tsp.make_queries(drfe, ku, ur, desc, len(ku), 800, 13, sf)
    tsp.plant_block(orc, p, ku, ur, gc, gi, desc, q, qd)
    fl = tsp.local_flags(orc, len(q), 5)
    rec, asg, lholder, lnm = orc.search_local_points(p, ku, ur, gc, gi, desc, q, qd, fl, 0.8, occ)
    out.update(lp_queries=q, lp_desc=qd, lp_flags=fl, lp_records=rec, lp_assigned=asg, lp_key_point=lholder, lp_nmatches=np.int32(lnm))
    # SearchByBoW
    kfd, kfa, kfv = test_bow.make_keyframe(desc, keys["angle"], 17)
    kf_fv = orc.Vocabulary(**voc).transform(kfd, 2)[3]
    km, fm, bnm = orc.search_by_bow(kfd, kfa, kfv, kf_fv, desc, keys["angle"], fv, 0.7, True)
    out.update(kf_desc=kfd, kf_angle=kfa, kf_valid=kfv, bow_kf_match=km, bow_f_match=fm, bow_nmatches=np.int32(bnm))
    out["kf_fv_node"], out["kf_fv_start"], out["kf_fv_feat"] = flat_fv(kf_fv)
    np.savez_compressed(os.path.join(HERE, "next_640x480_room.npz"), **out)
    print("matches: last frame %d, local points %d, bow %d; bow words %d" % (nm, lnm, bnm, len(bow)))


if __name__ == "__main__":
    main()
