"""Golden vectors of the "next" rows (tests/golden/next_640x480_room.npz, made by tests/golden/make_golden_next.py): the oracle
must reproduce them from the stored inputs (CPU), and the C-ABI device path must reproduce them too (GPU) — per-frame steps,
Frame::ComputeBoW, and the three whole-function matchers, every value bit for bit."""
import numpy as np
import pytest

from conftest import load_golden


def fv_list(node, start, feat):
    return [(int(node[j]), feat[start[j]:start[j + 1]].tolist()) for j in range(len(node))]


def voc_of(G):
    return dict(k=int(G["voc_k"]), L=int(G["voc_L"]), scoring=int(G["voc_scoring"]), weighting=int(G["voc_weighting"]), parent=G["voc_parent"],
                is_leaf=G["voc_is_leaf"], descriptors=G["voc_descriptors"], weights=G["voc_weights"])


def test_oracle_reproduces_next_golden(orc):
    G = load_golden("next_640x480_room.npz")
    o = orc.OrbOracle(1000)
    keys, desc = o.extract(G["gray"])
    assert keys.tobytes() == G["keys"].tobytes() and np.array_equal(desc, G["desc"])
    sf = np.array(o.scale_factors(), np.float32)
    p = orc.frame_params(*G["K"], list(G["dist"]), float(G["bf"]), 640, 480)
    ku, ur, kd, gc, gi = orc.frame_post(p, keys, G["depth"])
    assert ku.tobytes() == G["keys_un"].tobytes() and np.array_equal(ur, G["u_right"]) and np.array_equal(kd, G["kp_depth"])
    assert np.array_equal(gc, G["grid_count"]) and np.array_equal(gi, G["grid_index"])
    V = orc.Vocabulary(**voc_of(G))
    words, nodes, bow, fv = V.transform(desc, 2)
    assert np.array_equal(words, G["bow_words"]) and np.array_equal(nodes, G["bow_nodes"])
    assert [k for k, _ in bow] == G["bow_key"].tolist()
    assert np.array([v for _, v in bow], np.float64).tobytes() == G["bow_value"].tobytes()
    assert fv == fv_list(G["fv_node"], G["fv_start"], G["fv_feat"])
    mk, md, holder, nm = orc.search_last_frame(p, sf, ku, ur, gc, gi, desc, G["lf_Tcw"].ravel(), 15.0, 0, 1, G["lf_points"], G["lf_desc"], G["lf_occupied"])
    assert np.array_equal(mk, G["lf_match_key"]) and np.array_equal(md, G["lf_match_dist"]) and np.array_equal(holder, G["lf_key_point"]) and nm == G["lf_nmatches"]
    rec, asg, lh, lnm = orc.search_local_points(p, ku, ur, gc, gi, desc, G["lp_queries"], G["lp_desc"], G["lp_flags"], 0.8, G["lf_occupied"])
    assert rec.tobytes() == G["lp_records"].tobytes() and np.array_equal(asg, G["lp_assigned"]) and np.array_equal(lh, G["lp_key_point"]) and lnm == G["lp_nmatches"]
    kf_fv = fv_list(G["kf_fv_node"], G["kf_fv_start"], G["kf_fv_feat"])
    km, fm, bnm = orc.search_by_bow(G["kf_desc"], G["kf_angle"], G["kf_valid"], kf_fv, desc, keys["angle"], fv, 0.7, True)
    assert np.array_equal(km, G["bow_kf_match"]) and np.array_equal(fm, G["bow_f_match"]) and bnm == G["bow_nmatches"]


@pytest.mark.gpu
def test_gpu_reproduces_next_golden(drfe):
    G = load_golden("next_640x480_room.npz")
    ex = drfe.ORBextractor(1000, 1.2, 8, 20, 7, 640, 480)
    ex.enqueue(G["gray"][None])
    kps, desc, cnt = ex.download()
    n = int(cnt[0])
    assert kps[0, :n].tobytes() == G["keys"].tobytes() and np.array_equal(desc[0, :n], G["desc"])
    p = ex.frame_params(*G["K"], list(G["dist"]), float(G["bf"]))
    ku, ur, kd, gc, gi = ex.frame_post(p, G["depth"][None])
    assert ku[0, :n].tobytes() == G["keys_un"].tobytes() and np.array_equal(ur[0, :n], G["u_right"]) and np.array_equal(kd[0, :n], G["kp_depth"])
    assert np.array_equal(gc[0].reshape(-1), G["grid_count"]) and np.array_equal(gi[0, :len(G["grid_index"])], G["grid_index"])
    voc = drfe.Vocabulary(**voc_of(G))
    words, nodes, bow, fv = ex.compute_bow(voc, 2)[0]
    assert np.array_equal(words[:n], G["bow_words"]) and np.array_equal(nodes[:n], G["bow_nodes"])
    assert [k for k, _ in bow] == G["bow_key"].tolist()
    assert np.array([v for _, v in bow], np.float64).tobytes() == G["bow_value"].tobytes()
    assert fv == fv_list(G["fv_node"], G["fv_start"], G["fv_feat"])
    occ = np.zeros((1, ex.cap), np.uint8)
    occ[0, :n] = G["lf_occupied"]
    tp = np.zeros(1, drfe.TRACK_PARAMS_DTYPE)
    tp["Tcw"], tp["th"], tp["mode"], tp["check_orientation"] = G["lf_Tcw"].ravel(), 15.0, 0, 1
    mk, md, kp, nm, _ = ex.search_last_frame(tp, G["lf_points"][None], G["lf_desc"][None], None, occ)
    assert np.array_equal(mk[0], G["lf_match_key"]) and np.array_equal(md[0], G["lf_match_dist"]) and np.array_equal(kp[0, :n], G["lf_key_point"])
    assert nm[0] == G["lf_nmatches"]
    rec, asg, lkp, lnm = ex.search_local_points(G["lp_queries"][None], G["lp_desc"][None], G["lp_flags"][None], 0.8, None, occ)
    assert rec[0].tobytes() == G["lp_records"].tobytes() and np.array_equal(asg[0], G["lp_assigned"]) and np.array_equal(lkp[0, :n], G["lp_key_point"])
    assert lnm[0] == G["lp_nmatches"]
    nk = len(G["kf_angle"])
    kf_fv = fv_list(G["kf_fv_node"], G["kf_fv_start"], G["kf_fv_feat"])
    km, fm, bnm = ex.search_by_bow(np.array([nk], np.int32), G["kf_desc"][None], G["kf_angle"][None], G["kf_valid"][None], [kf_fv], [fv], 0.7, True)
    assert np.array_equal(km[0], G["bow_kf_match"]) and np.array_equal(fm[0, :n], G["bow_f_match"]) and bnm[0] == G["bow_nmatches"]
    voc.close()


@pytest.mark.gpu
def test_gpu_next_rows_do_not_depend_on_the_batch_position(drfe):
    """frames are independent: the golden frame placed at several positions of a 12-frame batch, between other
    frames, gives the golden results at every position — frame post, ComputeBoW and the matchers index their per-frame
    buffers by frame"""
    G = load_golden("next_640x480_room.npz")
    B, at = 12, (0, 5, 11)
    other = [drfe.synth_frame(640, 480, i % 3, 20261000 + i) for i in range(B)]
    gray = np.stack([o[0] for o in other])
    depth = np.stack([o[1] for o in other])
    for f in at:
        gray[f], depth[f] = G["gray"], G["depth"]
    ex = drfe.ORBextractor(1000, 1.2, 8, 20, 7, 640, 480, max_batch=B)
    ex.enqueue(gray)
    kps, desc, cnt = ex.download()
    n = len(G["keys"])
    p = ex.frame_params(*G["K"], list(G["dist"]), float(G["bf"]))
    ku, ur, kd, gc, gi = ex.frame_post(p, depth)
    voc = drfe.Vocabulary(**voc_of(G))
    bows = ex.compute_bow(voc, 2)
    npts, nq, nk = len(G["lf_points"]), len(G["lp_queries"]), len(G["kf_angle"])
    tp = np.zeros(B, drfe.TRACK_PARAMS_DTYPE)
    tp["Tcw"], tp["th"], tp["mode"], tp["check_orientation"] = G["lf_Tcw"].ravel(), 15.0, 0, 1
    P = np.zeros((B, npts), drfe.LAST_POINT_DTYPE)
    PD = np.zeros((B, npts, 32), np.uint8)
    Q = np.zeros((B, nq), drfe.QUERY_DTYPE)
    QD = np.zeros((B, nq, 32), np.uint8)
    FL = np.zeros((B, nq), np.uint8)
    occ = np.zeros((B, ex.cap), np.uint8)
    KD, KA, KV = np.zeros((B, nk, 32), np.uint8), np.zeros((B, nk), np.float32), np.zeros((B, nk), np.uint8)
    kn = np.zeros(B, np.int32)
    kf_fv = fv_list(G["kf_fv_node"], G["kf_fv_start"], G["kf_fv_feat"])
    kfvs = [[] for _ in range(B)]
    npv, nqv = np.zeros(B, np.int32), np.zeros(B, np.int32)
    for f in at:
        P[f], PD[f], Q[f], QD[f], FL[f] = G["lf_points"], G["lf_desc"], G["lp_queries"], G["lp_desc"], G["lp_flags"]
        occ[f, :n] = G["lf_occupied"]
        KD[f], KA[f], KV[f], kn[f], kfvs[f] = G["kf_desc"], G["kf_angle"], G["kf_valid"], nk, kf_fv
        npv[f], nqv[f] = npts, nq
    mk, md, kp, nm, _ = ex.search_last_frame(tp, P, PD, npv, occ)
    rec, asg, lkp, lnm = ex.search_local_points(Q, QD, FL, 0.8, nqv, occ)
    km, fm, bnm = ex.search_by_bow(kn, KD, KA, KV, kfvs, [b[3] for b in bows], 0.7, True)
    for f in at:
        assert cnt[f] == n and kps[f, :n].tobytes() == G["keys"].tobytes() and np.array_equal(desc[f, :n], G["desc"])
        assert ku[f, :n].tobytes() == G["keys_un"].tobytes() and np.array_equal(ur[f, :n], G["u_right"])
        assert np.array_equal(gc[f].reshape(-1), G["grid_count"])
        assert [k for k, _ in bows[f][2]] == G["bow_key"].tolist()
        assert np.array([v for _, v in bows[f][2]], np.float64).tobytes() == G["bow_value"].tobytes()
        assert bows[f][3] == fv_list(G["fv_node"], G["fv_start"], G["fv_feat"])
        assert np.array_equal(mk[f], G["lf_match_key"]) and np.array_equal(kp[f, :n], G["lf_key_point"]) and nm[f] == G["lf_nmatches"]
        assert rec[f].tobytes() == G["lp_records"].tobytes() and np.array_equal(asg[f], G["lp_assigned"]) and lnm[f] == G["lp_nmatches"]
        assert np.array_equal(km[f], G["bow_kf_match"]) and np.array_equal(fm[f, :n], G["bow_f_match"]) and bnm[f] == G["bow_nmatches"]
    for f in range(B):
        if f not in at:
            assert nm[f] == 0 and lnm[f] == 0 and bnm[f] == 0
    voc.close()
