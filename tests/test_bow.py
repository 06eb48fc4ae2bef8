"""Frame::ComputeBoW (reference src/Frame.cc:828-833) = DBoW2 TemplatedVocabulary::transform(features, BowVector&,
FeatureVector&, 4) (Thirdparty/DBoW2/DBoW2/TemplatedVocabulary.h:1126-1258, BowVector.cpp:34-84, FeatureVector.cpp:31-45).
CPU: the oracle's literal restatement against a vectorised brute-force descent and the maps' invariants.  GPU:
drfe_orb_compute_bow on the device-resident descriptors against the oracle — words, nodes, map order, and the double
values bit for bit."""
import numpy as np
import pytest


def brute_descent(voc, desc, levelsup):
    """independent vectorised descent: np.unpackbits distances, np.argmin (first minimum)"""
    n = len(voc["parent"])
    children = [[] for _ in range(n + 1)]
    for i, p in enumerate(voc["parent"]):
        children[int(p)].append(i + 1)
    D = np.vstack([np.zeros((1, 32), np.uint8), voc["descriptors"]])
    leaf_word = np.zeros(n + 1, np.int64)
    leaf_word[1:][voc["is_leaf"] > 0] = np.arange(int((voc["is_leaf"] > 0).sum()))
    words, nodes = [], []
    for d in desc:
        node, level, nid = 0, 0, None
        while children[node]:
            level += 1
            ch = children[node]
            dist = np.unpackbits(D[ch] ^ d[None, :], axis=1).sum(1)
            node = ch[int(np.argmin(dist))]
            if level == voc["L"] - levelsup:
                nid = node
        words.append(leaf_word[node])
        nodes.append(node if nid is None else nid)
    return np.array(words), np.array(nodes)


def frame_descriptors(drfe, orc, seed, scene=1):
    gray, _, _ = drfe.synth_frame(640, 480, scene, seed)
    keys, desc = orc.OrbOracle(1000).extract(gray)
    return gray, desc


@pytest.mark.parametrize("ragged", [False, True])
def test_oracle_bow_against_brute_force(drfe, orc, ragged):
    voc = orc.synth_vocabulary(6 if ragged else 10, 5 if ragged else 3, 11, ragged=ragged)
    V = orc.Vocabulary(**voc)
    assert V.nwords == int(voc["is_leaf"].sum())
    _, desc = frame_descriptors(drfe, orc, 20260501)
    desc = desc[:400]
    levelsup = 3 if ragged else 2
    words, nodes, bow, fv = V.transform(desc, levelsup)
    bw, bn = brute_descent(voc, desc, levelsup)
    kept = words >= 0
    assert kept.sum() > 300 and (~kept).sum() > 0                     # some words are stopped (weight 0)
    assert np.array_equal(words[kept], bw[kept]) and np.array_equal(nodes[kept], bn[kept])
    # the maps: keys ascending, values = count * weight up to rounding, L1 norm 1, every kept descriptor listed once
    assert [k for k, _ in bow] == sorted(set(words[kept].tolist()))
    assert abs(sum(v for _, v in bow) - 1.0) < 1e-12
    wt = {V.word_id[i]: V.weight[i] for i in range(1, len(V.weight)) if not V.children[i]}
    raw = np.array([wt[k] * (words == k).sum() for k, _ in bow])
    assert np.allclose(np.array([v for _, v in bow]), raw / raw.sum(), rtol=1e-12)
    assert [k for k, _ in fv] == sorted(set(nodes[kept].tolist()))
    flat = [i for _, l in fv for i in l]
    assert sorted(flat) == np.nonzero(kept)[0].tolist() and all(l == sorted(l) for _, l in fv)
    assert len(bow) < kept.sum()                                      # at least one word was hit twice (repeated addition path)


def test_oracle_bow_scorings(drfe, orc):
    _, desc = frame_descriptors(drfe, orc, 20260502)
    desc = desc[:200]
    base = orc.synth_vocabulary(8, 3, 5)
    l1 = orc.Vocabulary(**base).transform(desc)[2]
    l2 = orc.Vocabulary(**{**base, "scoring": 1}).transform(desc)[2]
    dot = orc.Vocabulary(**{**base, "scoring": 5}).transform(desc)[2]
    idf = orc.Vocabulary(**{**base, "weighting": 2}).transform(desc)[2]
    assert abs(sum(v * v for _, v in l2) - 1.0) < 1e-12
    assert [k for k, _ in l1] == [k for k, _ in l2] == [k for k, _ in dot] == [k for k, _ in idf]
    n = len(dot)
    r1 = np.array([v for _, v in l1]) / np.array([v for _, v in dot])
    assert np.allclose(r1, r1[0], rtol=1e-12) and abs(r1[0] * sum(v for _, v in dot) - 1) < 1e-12 and n > 50
    assert abs(sum(v for _, v in idf) - 1.0) < 1e-12


def same_maps(got, want):
    gw, gn, gbow, gfv = got
    ww, wn, wbow, wfv = want
    m = len(ww)
    assert np.array_equal(gw[:m], ww) and np.array_equal(gn[:m], wn)
    assert [k for k, _ in gbow] == [k for k, _ in wbow]
    assert np.array_equal(np.array([v for _, v in gbow], np.float64).view(np.uint64), np.array([v for _, v in wbow], np.float64).view(np.uint64))
    assert gfv == wfv


@pytest.mark.gpu
def test_gpu_compute_bow(drfe, orc):
    B = 3
    frames = [frame_descriptors(drfe, orc, 20260501 + 7 * i, scene=i % 3) for i in range(B)]
    ex = drfe.ORBextractor(1000, 1.2, 8, 20, 7, 640, 480, max_batch=B)
    ex.enqueue(np.stack([f[0] for f in frames]))
    kps, desc, cnt = ex.download()
    for f in range(B):
        assert np.array_equal(desc[f, :cnt[f]], frames[f][1])
    cases = [
        (orc.synth_vocabulary(10, 4, 21), 2),                                   # the ORB vocabulary's shape, shallower
        (orc.synth_vocabulary(6, 5, 22, ragged=True), 3),                       # leaves at several depths, 2..6 children
        (orc.synth_vocabulary(8, 3, 23, scoring=1), 4),                         # L2 norm, nid_level <= 0: every node id is the root
        (orc.synth_vocabulary(8, 3, 24, scoring=5), 1),                         # DOT_PRODUCT: no normalisation, values / size
        (orc.synth_vocabulary(8, 3, 25, weighting=2), 1),                       # IDF: addIfNotExist
    ]
    for voc, levelsup in cases:
        V = orc.Vocabulary(**voc)
        G = drfe.Vocabulary(**voc)
        assert G.words() == V.nwords
        got = ex.compute_bow(G, levelsup)
        for f in range(B):
            same_maps(got[f], V.transform(frames[f][1], levelsup))
        assert len(got[0][2]) > 100
        G.close()


@pytest.mark.gpu
def test_gpu_compute_bow_edge_cases(drfe, orc):
    """a frame with no keypoints gives empty maps; a vocabulary on another device than the extractor is refused"""
    gray, desc = frame_descriptors(drfe, orc, 20260503)
    ex = drfe.ORBextractor(1000, 1.2, 8, 20, 7, 640, 480, max_batch=2)
    ex.enqueue(np.stack([np.full((480, 640), 128, np.uint8), gray]))
    voc = orc.synth_vocabulary(10, 3, 31)
    G = drfe.Vocabulary(**voc)
    got = ex.compute_bow(G)
    assert got[0][2] == [] and got[0][3] == []
    same_maps(got[1], orc.Vocabulary(**voc).transform(desc))
    with pytest.raises(drfe.DrfeError):
        drfe.Vocabulary(10, 3, 0, 0, np.array([5], np.int32), [1], np.zeros((1, 32), np.uint8), [1.0])   # parent after the node


# ---------------------------------------------------------------- ORBmatcher::SearchByBoW (ORBmatcher.cc:160-292)
def make_keyframe(desc, angles, seed, n_extra=150):
    """a keyframe that saw the same scene: the frame's descriptors with bits flipped (some heavily), shuffled, plus
    unrelated ones; a block of identical descriptors so that several keyframe features want the same frame feature"""
    rng = np.random.default_rng(seed)
    n = len(desc)
    perm = rng.permutation(n)
    kd = desc[perm].copy()
    for i in range(n):
        for b in rng.integers(0, 256, int(rng.choice([0, 3, 8, 20, 60], p=[0.2, 0.3, 0.3, 0.15, 0.05]))):
            kd[i, b >> 3] ^= 1 << (b & 7)
    ka = np.mod(angles[perm] + rng.normal(0, 5, n) + np.where(rng.random(n) < 0.15, rng.uniform(0, 360, n), 0), 360).astype(np.float32)
    kd[20:28] = kd[20]                                                     # eight copies: in-order occupancy decides
    kd = np.vstack([kd, rng.integers(0, 256, (n_extra, 32), dtype=np.uint8)])
    ka = np.concatenate([ka, rng.uniform(0, 360, n_extra).astype(np.float32)])
    valid = (rng.random(len(kd)) < 0.85).astype(np.uint8)
    valid[20:28] = 1
    return kd, ka, valid


def test_oracle_search_by_bow_properties(drfe, orc):
    gray, desc = frame_descriptors(drfe, orc, 20260510)
    keys, _ = orc.OrbOracle(1000).extract(gray)
    V = orc.Vocabulary(**orc.synth_vocabulary(10, 4, 41))
    kd, ka, valid = make_keyframe(desc, keys["angle"], 2)
    f_fv = V.transform(desc, 2)[3]
    kf_fv = V.transform(kd, 2)[3]
    km, fm, nm = orc.search_by_bow(kd, ka, valid, kf_fv, desc, keys["angle"], f_fv, 0.7, True)
    km0, fm0, nm0 = orc.search_by_bow(kd, ka, valid, kf_fv, desc, keys["angle"], f_fv, 0.7, False)
    assert np.array_equal(km, km0) and nm0 == (km0 >= 0).sum() > 200 and nm < nm0
    node_of_f = {i: nid for nid, l in f_fv for i in l}
    node_of_k = {i: nid for nid, l in kf_fv for i in l}
    ham = lambda a, b: int(np.unpackbits(a ^ b).sum())
    m = np.nonzero(km0 >= 0)[0]
    assert len(set(km0[m])) == len(m)                                        # a frame feature is matched once
    for i in m:
        assert valid[i] and node_of_k[i] == node_of_f[km0[i]] and ham(kd[i], desc[km0[i]]) <= orc.TH_LOW and fm0[km0[i]] == i
    assert (fm0 >= 0).sum() == len(m) and (fm >= 0).sum() == nm
    assert (km0[20:28] >= 0).sum() >= 1


@pytest.mark.gpu
def test_gpu_search_by_bow(drfe, orc):
    B = 3
    frames = [frame_descriptors(drfe, orc, 20260510 + 3 * i, scene=i % 3) for i in range(B)]
    ex = drfe.ORBextractor(1000, 1.2, 8, 20, 7, 640, 480, max_batch=B)
    ex.enqueue(np.stack([f[0] for f in frames]))
    kps, desc, cnt = ex.download()
    voc = orc.synth_vocabulary(10, 4, 41)
    V, G = orc.Vocabulary(**voc), drfe.Vocabulary(**voc)
    bows = ex.compute_bow(G, 2)
    kcap = 1200
    KD, KA, KV = np.zeros((B, kcap, 32), np.uint8), np.zeros((B, kcap), np.float32), np.zeros((B, kcap), np.uint8)
    kn, kfvs = np.zeros(B, np.int32), []
    for f in range(B):
        n = int(cnt[f])
        kd, ka, valid = make_keyframe(desc[f, :n], kps[f, :n]["angle"], 50 + f, n_extra=0 if f == 2 else 150)
        if f == 2:
            kd, ka, valid = kd[:40], ka[:40], valid[:40]                   # a small keyframe: most frame nodes have no partner
        kn[f] = len(kd)
        KD[f, :kn[f]], KA[f, :kn[f]], KV[f, :kn[f]] = kd, ka, valid
        kfvs.append(V.transform(kd, 2)[3])
    for nnratio, check in ((0.7, True), (0.9, False)):
        km, fm, nm = ex.search_by_bow(kn, KD, KA, KV, kfvs, [b[3] for b in bows], nnratio, check)
        for f in range(B):
            n = int(cnt[f])
            wkm, wfm, wnm = orc.search_by_bow(KD[f, :kn[f]], KA[f, :kn[f]], KV[f, :kn[f]], kfvs[f], desc[f, :n], kps[f, :n]["angle"], bows[f][3],
                                              nnratio, check)
            assert np.array_equal(km[f, :kn[f]], wkm) and (km[f, kn[f]:] == -1).all(), f
            assert np.array_equal(fm[f, :n], wfm) and (fm[f, n:] == -1).all(), f
            assert nm[f] == wnm, f
        assert nm[0] > 200
    G.close()


@pytest.mark.gpu
def test_gpu_compute_bow_2000kp(drfe, orc):
    """1280x720 / 2000 keypoints: the per-frame sort runs on 4096 keys"""
    gray, _, _ = drfe.synth_frame(1280, 720, 2, 20260520)
    keys, desc = orc.OrbOracle(2000).extract(gray)
    ex = drfe.ORBextractor(2000, 1.2, 8, 20, 7, 1280, 720, max_batch=2)
    ex.enqueue(np.stack([gray, gray[::-1].copy()]))
    kps, gd, cnt = ex.download()
    assert np.array_equal(gd[0, :cnt[0]], desc) and cnt[0] >= 2000
    voc = orc.synth_vocabulary(10, 4, 71)
    V, G = orc.Vocabulary(**voc), drfe.Vocabulary(**voc)
    got = ex.compute_bow(G, 2)
    same_maps(got[0], V.transform(desc, 2))
    same_maps(got[1], V.transform(gd[1, :cnt[1]], 2))
    G.close()
