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
