"""Core of ORBmatcher::SearchByProjection(Frame&, vector<MapPoint*>&, th) (reference src/ORBmatcher.cc:69-116,
Frame::GetFeaturesInArea src/Frame.cc:730-779): the oracle's literal loops against brute force on the CPU, and
drfe_orb_search_by_projection (device, C ABI) against the oracle on the GPU — every field bit-exact."""
import numpy as np
import pytest

K = (525.0, 525.0, 319.5, 239.5)


def make_queries(drfe, keys_un, u_right, desc, n, nq, seed, scale_factors):
    """queries the way TrackLocalMap makes them: map points projected near existing keypoints (so that windows are
    populated), predicted level around the keypoint's octave, descriptors = the keypoint's with a few bits flipped"""
    rng = np.random.default_rng(seed)
    q = np.zeros(nq, drfe.QUERY_DTYPE)
    qd = np.zeros((nq, 32), np.uint8)
    src = rng.integers(0, n, nq)
    lvl = np.clip(keys_un["octave"][src] + rng.integers(-1, 2, nq), 0, 7)
    q["x"] = keys_un["x"][src] + rng.normal(0, 3, nq).astype(np.float32)
    q["y"] = keys_un["y"][src] + rng.normal(0, 3, nq).astype(np.float32)
    q["r"] = (np.where(rng.random(nq) < 0.5, np.float32(2.5), np.float32(4.0)) * np.float32(rng.choice([1.0, 3.0])) *
              scale_factors[lvl]).astype(np.float32)
    q["xr"] = np.where(u_right[src] > 0, u_right[src] + rng.normal(0, 4, nq), q["x"] - 20).astype(np.float32)
    q["min_level"], q["max_level"] = lvl - 1, lvl
    q["min_level"][::7], q["max_level"][::7] = -1, -1                # no level check (minLevel <= 0 and maxLevel < 0)
    q["x"][::11] += 700                                               # windows outside the grid
    flips = rng.integers(0, 256, (nq, 12))
    qd[:] = desc[src]
    for i in range(nq):
        for b in flips[i, :rng.integers(0, 12)]:
            qd[i, b >> 3] ^= 1 << (b & 7)
    return q, qd


def frame_inputs(drfe, orc, seed, scene=1):
    gray, depth, _ = drfe.synth_frame(640, 480, scene, seed)
    o = orc.OrbOracle(1000)
    keys, desc = o.extract(gray)
    p = orc.frame_params(*K, [0.1, -0.05, 0.001, 0.0005, 0.0], 40.0, 640, 480)
    ku, ur, kd, gc, gi = orc.frame_post(p, keys, depth)
    return gray, depth, p, ku, ur, gc, gi, desc, np.array(o.scale_factors(), np.float32)


def test_oracle_search_matches_brute_force(drfe, orc):
    gray, depth, p, ku, ur, gc, gi, desc, sf = frame_inputs(drfe, orc, 20260420)
    n = len(ku)
    q, qd = make_queries(drfe, ku, ur, desc, n, 300, 5, sf)
    occ = (np.random.default_rng(9).random(n) < 0.1).astype(np.uint8)
    got = orc.search_by_projection(p, ku, ur, gc, gi, desc, q, qd, occ)
    ham = np.unpackbits(qd[:, None, :] ^ desc[None, :, :], axis=2).sum(2)
    placed = np.zeros(n, bool)
    placed[gi] = True                                                  # keypoints outside the image bounds are in no cell
    some = 0
    for i in range(len(q)):
        x, y, r = q["x"][i], q["y"][i], q["r"][i]
        ok = placed & (np.abs(ku["x"] - x) < r) & (np.abs(ku["y"] - y) < r) & (occ == 0)
        if q["min_level"][i] > 0 or q["max_level"][i] >= 0:
            ok &= ku["octave"] >= q["min_level"][i]
            if q["max_level"][i] >= 0:
                ok &= ku["octave"] <= q["max_level"][i]
        ok &= ~((ur > 0) & (np.abs(q["xr"][i] - ur) > r))
        d = np.where(ok, ham[i], 999)
        if ok.any():
            some += 1
            assert got["best_dist"][i] == d.min() and d[got["best_idx"][i]] == d.min()
            assert got["best_level"][i] == ku["octave"][got["best_idx"][i]]
            if ok.sum() >= 2:
                assert got["best_dist2"][i] == np.sort(d)[1]
        else:
            assert tuple(got[i]) == (256, -1, -1, 256, -1)
    assert some > 150


@pytest.mark.gpu
def test_gpu_search_by_projection(drfe, orc):
    B = 3
    frames = [frame_inputs(drfe, orc, 20260420 + 5 * i, scene=i % 3) for i in range(B)]
    ex = drfe.ORBextractor(1000, 1.2, 8, 20, 7, 640, 480, max_batch=B)
    ex.enqueue(np.stack([f[0] for f in frames]))
    kps, desc, cnt = ex.download()
    p = ex.frame_params(*K, [0.1, -0.05, 0.001, 0.0005, 0.0], 40.0)
    ku, ur, kd, gc, gi = ex.frame_post(p, np.stack([f[1] for f in frames]))
    qcap = 600
    Q = np.zeros((B, qcap), drfe.QUERY_DTYPE)
    QD = np.zeros((B, qcap, 32), np.uint8)
    occ = (np.random.default_rng(3).random((B, ex.cap)) < 0.1).astype(np.uint8)
    nq = np.array([600, 450, 1], np.int32)
    for f in range(B):
        n = int(cnt[f])
        assert np.array_equal(desc[f, :n], frames[f][7])               # same keypoints as the oracle's
        q, qd = make_queries(drfe, ku[f, :n], ur[f, :n], desc[f, :n], n, qcap, 17 + f, frames[f][8])
        Q[f], QD[f] = q, qd
    got = ex.search_by_projection(Q, QD, nq, occ)
    for f in range(B):
        n = int(cnt[f])
        placed = int(gc[f].sum())
        want = orc.search_by_projection(frames[f][2], ku[f, :n], ur[f, :n], gc[f], gi[f, :placed], desc[f, :n], Q[f, :nq[f]], QD[f, :nq[f]], occ[f, :n])
        for name in want.dtype.names:
            assert np.array_equal(got[f, :nq[f]][name], want[name]), (f, name)
    assert (got[0]["best_idx"] >= 0).sum() > 300
    # no occupied mask
    got2 = ex.search_by_projection(Q, QD, nq, None)
    want2 = orc.search_by_projection(frames[0][2], ku[0, :int(cnt[0])], ur[0, :int(cnt[0])], gc[0], gi[0, :int(gc[0].sum())], desc[0, :int(cnt[0])], Q[0], QD[0])
    assert all(np.array_equal(got2[0][nm], want2[nm]) for nm in want2.dtype.names)


def plant_block(orc, p, ku, ur, gc, gi, desc, q, qd):
    """copies of one clearly matching query at 100..119: twenty map points that want the same keypoint"""
    core = orc.search_by_projection(p, ku, ur, gc, gi, desc, q[:100], qd[:100])
    ok = (core["best_dist"] <= 30) & ((core["best_level"] != core["best_level2"]) | (core["best_dist"] <= 0.5 * core["best_dist2"]))
    j = int(np.nonzero(ok)[0][0])
    q[100:120], qd[100:120] = q[j], qd[j]


def local_flags(orc, nq, seed):
    rng = np.random.default_rng(seed)
    return (orc.LP_VALID * (rng.random(nq) < 0.9) + orc.LP_OBSERVED * (rng.random(nq) < 0.8)).astype(np.uint8)


def test_oracle_local_points_whole_function(drfe, orc):
    """the whole SearchByProjection(F, vpMapPoints, th): with nothing assigned (no query valid for assignment purposes) the
    records equal the core's; assignments never give an observed keypoint twice; every assignment passed TH_HIGH and the
    ratio test on the record the loop left"""
    gray, depth, p, ku, ur, gc, gi, desc, sf = frame_inputs(drfe, orc, 20260421)
    n = len(ku)
    q, qd = make_queries(drfe, ku, ur, desc, n, 500, 6, sf)
    plant_block(orc, p, ku, ur, gc, gi, desc, q, qd)
    occ = (np.random.default_rng(10).random(n) < 0.1).astype(np.uint8)
    occ[orc.search_by_projection(p, ku, ur, gc, gi, desc, q[100:101], qd[100:101])["best_idx"][0]] = 0
    fl = local_flags(orc, len(q), 3)
    fl[100:120] = orc.LP_VALID | orc.LP_OBSERVED
    rec, asg, holder, nm = orc.search_local_points(p, ku, ur, gc, gi, desc, q, qd, fl, 0.8, occ)
    core = orc.search_by_projection(p, ku, ur, gc, gi, desc, q, qd, occ)
    first = np.nonzero(asg >= 0)[0][0]
    for name in core.dtype.names:                                          # up to the first assignment nothing differs from the core
        assert np.array_equal(rec[name][:first + 1][(fl[:first + 1] & orc.LP_VALID) != 0], core[name][:first + 1][(fl[:first + 1] & orc.LP_VALID) != 0])
    a = np.nonzero(asg >= 0)[0]
    assert nm == len(a) > 150
    obs = a[(fl[a] & orc.LP_OBSERVED) != 0]
    assert len(set(asg[obs])) == len(obs) and not occ[asg[a]].any()
    for i in a:
        r = rec[i]
        assert r["best_idx"] == asg[i] and r["best_dist"] <= orc.TH_HIGH
        assert not (r["best_level"] == r["best_level2"] and np.float32(r["best_dist"]) > np.float32(np.float32(0.8) * np.float32(r["best_dist2"])))
    assert (asg[100:120] >= 0).sum() >= 1
    assert all(asg[holder[idx]] == idx for idx in np.nonzero(holder >= 0)[0])


@pytest.mark.gpu
def test_gpu_search_local_points(drfe, orc):
    B = 3
    frames = [frame_inputs(drfe, orc, 20260421 + 5 * i, scene=i % 3) for i in range(B)]
    ex = drfe.ORBextractor(1000, 1.2, 8, 20, 7, 640, 480, max_batch=B)
    ex.enqueue(np.stack([f[0] for f in frames]))
    kps, desc, cnt = ex.download()
    p = ex.frame_params(*K, [0.1, -0.05, 0.001, 0.0005, 0.0], 40.0)
    ku, ur, kd, gc, gi = ex.frame_post(p, np.stack([f[1] for f in frames]))
    qcap = 1200
    Q = np.zeros((B, qcap), drfe.QUERY_DTYPE)
    QD = np.zeros((B, qcap, 32), np.uint8)
    FL = np.zeros((B, qcap), np.uint8)
    occ = (np.random.default_rng(3).random((B, ex.cap)) < 0.1).astype(np.uint8)
    nq = np.array([1200, 700, 0], np.int32)
    for f in range(B):
        n = int(cnt[f])
        q, qd = make_queries(drfe, ku[f, :n], ur[f, :n], desc[f, :n], n, qcap, 27 + f, frames[f][8])
        plant_block(orc, frames[f][2], ku[f, :n], ur[f, :n], gc[f], gi[f, :int(gc[f].sum())], desc[f, :n], q, qd)
        Q[f], QD[f], FL[f] = q, qd, local_flags(orc, qcap, 40 + f)
        FL[f, 100:120] = orc.LP_VALID | orc.LP_OBSERVED
    for nnratio, use_occ in ((0.8, True), (0.6, False)):
        rec, asg, kp, nm = ex.search_local_points(Q, QD, FL, nnratio, nq, occ if use_occ else None)
        for f in range(B):
            n, m = int(cnt[f]), int(nq[f])
            placed = int(gc[f].sum())
            wrec, wasg, wh, wnm = orc.search_local_points(frames[f][2], ku[f, :n], ur[f, :n], gc[f], gi[f, :placed], desc[f, :n], Q[f, :m], QD[f, :m],
                                                          FL[f, :m], nnratio, occ[f, :n] if use_occ else None)
            for name in wrec.dtype.names:
                assert np.array_equal(rec[f, :m][name], wrec[name]), (f, name)
            assert np.array_equal(asg[f, :m], wasg) and np.array_equal(kp[f, :n], wh) and nm[f] == wnm, f
            assert (asg[f, m:] == -1).all() and (rec[f, m:]["best_idx"] == -1).all() and (rec[f, m:]["best_dist"] == 256).all()
        assert nm[0] > 300 and nm[2] == 0
