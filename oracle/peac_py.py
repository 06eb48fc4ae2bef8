"""ORACLE — TEST INFRASTRUCTURE ONLY.  A second, independent restatement of DR-SLAM's PEAC-AHC plane extractor in plain Python,
written from the reference headers with Python's own containers (heapq, set, list) — not from oracle/peac_oracle.cpp:
  PlaneDetection::readDepthImage / runPlaneDetection         src/PlaneExtractor.cpp:28-63
  ahc::PlaneFitter::run / initGraph / ahCluster              include/peac/AHCPlaneFitter.hpp:211-259, :804-965, :976-1190
  refineDetails / findBlockMembership / floodFill            :298-382, :494-600, :434-488
  ahc::PlaneSeg, Stats                                       include/peac/AHCPlaneSeg.hpp:57-409
  ahc::ParamSet, DisjointSet                                 include/peac/AHCParamSet.hpp:46-147, DisjointSet.hpp:31-95
tests/test_peac.py requires it to agree with the C++ restatement value for value (doubles bit for bit) with solver="jacobi", and —
with solver="lapack", numpy.linalg.eigh standing where the reference has Eigen's SelfAdjointEigenSolver — to make the same
decisions (same seg_output, same plane order) with plane parameters equal to ~1e-10: the declared solver substitution (P.1 of
peac_oracle.cpp) does not change what is extracted.  The same declared orders (P.2 - P.4) apply: equal mse pops by creation number,
neighbours are tried in creation order, the final sort is stable.  Slow (pure Python): use on small frames."""
import heapq
import math

import numpy as np

DEFAULTS = dict(depthSigma=1.6e-6, stdTol_init=5.0, stdTol_merge=8.0, z_near=500.0, z_far=4000.0, angle_near=math.radians(15.0),
                angle_far=math.radians(90.0), similarityTh_merge=math.cos(math.radians(60.0)), similarityTh_refine=math.cos(math.radians(30.0)),
                depthAlpha=0.04, depthChangeTol=0.02)


def jacobi3(K):
    """cyclic Jacobi as declared in peac_oracle.cpp (P.1): pivots (0,1), (0,2), (1,2); a pivot too small to change the diagonal is zeroed
    from the fourth sweep on; eigenvalues ascending by three compare-and-swaps"""
    a = [[K[0][0], K[0][1], K[0][2]], [K[0][1], K[1][1], K[1][2]], [K[0][2], K[1][2], K[2][2]]]
    v = [[1.0, 0.0, 0.0], [0.0, 1.0, 0.0], [0.0, 0.0, 1.0]]
    for sweep in range(24):
        if a[0][1] == 0.0 and a[0][2] == 0.0 and a[1][2] == 0.0:
            break
        for p, q, r in ((0, 1, 2), (0, 2, 1), (1, 2, 0)):
            apq = a[p][q]
            if apq == 0.0:
                continue
            app, aqq = a[p][p], a[q][q]
            g = 100.0 * abs(apq)
            if sweep > 2 and abs(app) + g == abs(app) and abs(aqq) + g == abs(aqq):
                a[p][q] = a[q][p] = 0.0
                continue
            h = aqq - app
            if abs(h) + g == abs(h):
                t = apq / h
            else:
                theta = 0.5 * h / apq
                t = 1.0 / (abs(theta) + math.sqrt(1.0 + theta * theta))
                if theta < 0.0:
                    t = -t
            c = 1.0 / math.sqrt(1.0 + t * t)
            s = t * c
            a[p][p] = app - t * apq
            a[q][q] = aqq + t * apq
            a[p][q] = a[q][p] = 0.0
            arp, arq = a[r][p], a[r][q]
            a[r][p] = a[p][r] = c * arp - s * arq
            a[r][q] = a[q][r] = s * arp + c * arq
            for m in range(3):
                vp, vq = v[m][p], v[m][q]
                v[m][p] = c * vp - s * vq
                v[m][q] = s * vp + c * vq
    order = [0, 1, 2]
    d = [a[0][0], a[1][1], a[2][2]]
    if d[order[1]] < d[order[0]]:
        order[0], order[1] = order[1], order[0]
    if d[order[2]] < d[order[1]]:
        order[1], order[2] = order[2], order[1]
    if d[order[1]] < d[order[0]]:
        order[0], order[1] = order[1], order[0]
    return [d[i] for i in order], [[v[m][i] for i in order] for m in range(3)]


class Seg:
    """ahc::PlaneSeg: the nine sums, N, rid, and what Stats::compute derives from them"""
    __slots__ = ("s", "N", "rid", "mse", "center", "normal", "curvature", "nouse", "nbs", "seq")

    def __init__(self, s, N, rid, seq, solver):
        self.s, self.N, self.rid, self.seq = s, N, rid, seq
        self.nouse, self.nbs = False, set()
        self.mse = self.curvature = float("nan")
        self.center, self.normal = [0.0, 0.0, 0.0], [0.0, 0.0, 0.0]
        if N >= 4:
            self.compute(solver)

    def compute(self, solver):                                     # AHCPlaneSeg.hpp:128-162
        sx, sy, sz, sxx, syy, szz, sxy, syz, sxz = self.s
        sc = 1.0 / self.N
        self.center = [sx * sc, sy * sc, sz * sc]
        K = [[sxx - sx * sx * sc, sxy - sx * sy * sc, sxz - sx * sz * sc], [0.0, syy - sy * sy * sc, syz - sy * sz * sc], [0.0, 0.0, szz - sz * sz * sc]]
        K[1][0], K[2][0], K[2][1] = K[0][1], K[0][2], K[1][2]
        if solver == "lapack":
            w, V = np.linalg.eigh(np.array(K))
            sv, V = [float(x) for x in w], [[float(V[m][i]) for i in range(3)] for m in range(3)]
        else:
            sv, V = jacobi3(K)
        c = self.center
        if V[0][0] * c[0] + V[1][0] * c[1] + V[2][0] * c[2] <= 0:
            self.normal = [V[0][0], V[1][0], V[2][0]]
        else:
            self.normal = [-V[0][0], -V[1][0], -V[2][0]]
        self.mse = sv[0] * sc
        self.curvature = sv[0] / (sv[0] + sv[1] + sv[2])

    def similarity(self, o):
        return abs(self.normal[0] * o.normal[0] + self.normal[1] * o.normal[1] + self.normal[2] * o.normal[2])


class DisjointSet:
    def __init__(self, n):
        self.parent, self.size = list(range(n)), [1] * n

    def find(self, x):
        root = x
        while self.parent[root] != root:
            root = self.parent[root]
        while self.parent[x] != root:
            self.parent[x], x = root, self.parent[x]
        return root

    def union(self, x, y):
        xr, yr = self.find(x), self.find(y)
        if xr == yr:
            return
        if self.size[xr] < self.size[yr]:
            self.parent[xr] = yr
            self.size[yr] += self.size[xr]
        else:
            self.parent[yr] = xr
            self.size[xr] += self.size[yr]


def cloud_from_depth(depth16, factor, fx, fy, cx, cy):
    """readDepthImage: (H*W, 3) doubles; z > 5 -> (0, 0, 0).  factor, fx .. cy are floats as in the reference (float * -> double)"""
    H, W = depth16.shape
    f32 = np.float32
    z = depth16.astype(np.float64) * np.float64(f32(factor))
    jj, ii = np.meshgrid(np.arange(W, dtype=np.float64), np.arange(H, dtype=np.float64))
    x = (jj - np.float64(f32(cx))) * z / np.float64(f32(fx))
    y = (ii - np.float64(f32(cy))) * z / np.float64(f32(fy))
    far = z > 5.0
    out = np.stack([np.where(far, 0.0, x), np.where(far, 0.0, y), np.where(far, 0.0, z)], axis=-1)
    return out.reshape(-1, 3)


def run(cloud, width, height, params=None, min_support=3000, win_w=10, win_h=10, solver="jacobi"):
    """ahc::PlaneFitter::run with doRefine, ERODE_ALL_BORDER, INIT_STRICT -> (seg_output, planes [n][10] = normal, center, mse,
    curvature, N, rid; plane_vertices_ as index arrays; clustering steps of both passes)"""
    P = dict(DEFAULTS)
    if params is not None:
        P.update(params)
    X = np.asarray(cloud, np.float64).reshape(height, width, 3)
    Nh, Nw = height // win_h, width // win_w
    seq = [0]

    def new_seq():
        seq[0] += 1
        return seq[0] - 1

    def get(i, j):
        z = X[i, j, 2]
        return not (z == 0 or z != z)

    def t_mse(z, tol):
        return (P["depthSigma"] * z * z + tol) ** 2

    def t_ang_init(z):
        cz = min(max(z, P["z_near"]), P["z_far"])
        factor = (P["angle_far"] - P["angle_near"]) / (P["z_far"] - P["z_near"])
        return math.cos(factor * cz + P["angle_near"] - factor * P["z_near"])

    def discontinuous(d0, d1):
        return abs(d0 - d1) > P["depthAlpha"] * abs(d0) + P["depthChangeTol"]

    # ---- initGraph: nodes
    valid = ~((X[:, :, 2] == 0) | np.isnan(X[:, :, 2]))
    G = [None] * (Nh * Nw)
    heap = []
    ds = DisjointSet(Nh * Nw)
    for bi in range(Nh):
        for bj in range(Nw):
            r0, c0 = bi * win_h, bj * win_w
            ok = True
            for i in range(r0, min(r0 + win_h, height)):
                for j in range(c0, min(c0 + win_w, width)):
                    if not valid[i, j]:
                        ok = False
                        break
                    z = X[i, j, 2]
                    if j + 1 < width and valid[i, j + 1] and discontinuous(z, X[i, j + 1, 2]):
                        ok = False
                        break
                    if i + 1 < height and valid[i + 1, j] and discontinuous(z, X[i + 1, j, 2]):
                        ok = False
                        break
                if not ok:
                    break
            sid = new_seq()
            if not ok:
                continue
            blk = X[r0:r0 + win_h, c0:c0 + win_w].reshape(-1, 3)   # row-major, the order Stats::push sees
            x, y, z = blk[:, 0], blk[:, 1], blk[:, 2]
            s = [float(np.cumsum(t)[-1]) for t in (x, y, z, x * x, y * y, z * z, x * y, y * z, x * z)]   # cumsum adds one after the other
            node = Seg(s, len(blk), bi * Nw + bj, sid, solver)
            if node.N >= 4 and node.mse < t_mse(node.center[2], P["stdTol_init"]):
                G[bi * Nw + bj] = node
                heapq.heappush(heap, (node.mse, node.seq, node))

    def connect(a, b):
        a.nbs.add(b)
        b.nbs.add(a)

    for i in range(Nh):                                            # edges along rows (:899-931): the loop steps back and forth
        j = 1
        while j < Nw:
            c = i * Nw + j
            if G[c - 1] is None:
                j -= 1
            elif G[c] is None:
                pass
            elif j < Nw - 1 and G[c + 1] is None:
                j += 1
            else:
                th = t_ang_init(G[c].center[2])
                if (j < Nw - 1 and G[c - 1].similarity(G[c + 1]) >= th) or (j == Nw - 1 and G[c].similarity(G[c - 1]) >= th):
                    connect(G[c], G[c - 1])
                    if j < Nw - 1:
                        connect(G[c], G[c + 1])
                else:
                    j -= 1
            j += 2
    for j in range(Nw):                                            # edges along columns (:932-965)
        i = 1
        while i < Nh:
            c = i * Nw + j
            if G[c - Nw] is None:
                i -= 1
            elif G[c] is None:
                pass
            elif i < Nh - 1 and G[c + Nw] is None:
                i += 1
            else:
                th = t_ang_init(G[c].center[2])
                if (i < Nh - 1 and G[c - Nw].similarity(G[c + Nw]) >= th) or (i == Nh - 1 and G[c].similarity(G[c - Nw]) >= th):
                    connect(G[c], G[c - Nw])
                    if i < Nh - 1:
                        connect(G[c], G[c + Nw])
                else:
                    i -= 1
            i += 2

    def disconnect_all(p):
        for nb in p.nbs:
            nb.nbs.discard(p)
        p.nbs = set()

    def ah_cluster(heap):                                          # :976-1190
        extracted, steps = [], 0
        while heap:
            _, _, p = heapq.heappop(heap)
            if p.nouse:
                continue
            cand, cand_nb = None, None
            for nb in sorted(p.nbs, key=lambda n: n.seq):          # P.3
                if p.similarity(nb) < P["similarityTh_merge"]:
                    continue
                m = Seg([a + b for a, b in zip(p.s, nb.s)], p.N + nb.N, p.rid if p.N >= nb.N else nb.rid, -1, solver)
                if cand is None or cand.mse > m.mse or (cand.mse == m.mse and cand.N < m.mse):   # sic (:1048)
                    cand, cand_nb = m, nb
            if cand is not None and cand.mse < t_mse(cand.center[2], P["stdTol_merge"]):
                cand.seq = new_seq()
                heapq.heappush(heap, (cand.mse, cand.seq, cand))
                ds.union(p.rid, cand_nb.rid)                       # mergeNbsFrom (AHCPlaneSeg.hpp:378-407)
                cand.nbs = (p.nbs | cand_nb.nbs) - {p, cand_nb}
                disconnect_all(p)
                disconnect_all(cand_nb)
                for nb in cand.nbs:
                    nb.nbs.add(cand)
                p.nouse = cand_nb.nouse = True
            else:
                if p.N >= min_support:
                    extracted.append(p)
                disconnect_all(p)
            steps += 1
        extracted.sort(key=lambda n: -n.N)                         # P.4: stable
        return extracted, steps

    extracted, steps = ah_cluster(heap)

    # ---- findBlockMembership (:494-600)
    rid2plid = {p.rid: i for i, p in enumerate(extracted)}
    member = np.full(height * width, -1, np.int64)
    blk_map = [-1] * (Nh * Nw)
    is_valid = [False] * len(extracted)
    queue = []
    per_blk = win_h * win_w
    for i in range(Nh):
        for j in range(Nw):
            b = i * Nw + j
            setid = ds.find(b)
            if ds.size[setid] * per_blk >= min_support:
                nbs = ([b - 1] if j > 0 else []) + ([b + 1] if j < Nw - 1 else []) + ([b - Nw] if i > 0 else []) + ([b + Nw] if i < Nh - 1 else [])
                if all(ds.find(n) == setid for n in nbs):          # ERODE_ALL_BORDER
                    plid = rid2plid[setid]
                    blk_map[b] = plid
                    rows = np.arange(i * win_h, (i + 1) * win_h)[:, None] * width + np.arange(j * win_w, (j + 1) * win_w)[None, :]
                    member[rows.ravel()] = plid
                    is_valid[plid] = True
            if blk_map[b] < 0:
                if i > 0 and blk_map[b - Nw] >= 0:
                    sp = (i * win_h - 1) * width + j * win_w
                    queue += [(sp + k, blk_map[b - Nw]) for k in range(1, win_w)]
                if j > 0 and blk_map[b - 1] >= 0:
                    sp = (i * win_h) * width + j * win_w - 1
                    queue += [(sp + k * width, blk_map[b - 1]) for k in range(win_h - 1)]
            else:
                plid = blk_map[b]
                if i > 0 and blk_map[b - Nw] != plid:
                    sp = (i * win_h) * width + j * win_w
                    queue += [(sp + k, plid) for k in range(win_w - 1)]
                if j > 0 and blk_map[b - 1] != plid:
                    sp = (i * win_h) * width + j * win_w
                    queue += [(sp + k * width, plid) for k in range(1, win_h)]

    # ---- floodFill (:434-488)
    dist = np.full(height * width, np.finfo(np.float32).max, np.float32)
    flat = X.reshape(-1, 3)
    k = 0
    while k < len(queue):
        s_idx, plid = queue[k]
        k += 1
        pl = extracted[plid]
        sy, sx = divmod(s_idx, width)
        nbs = ([s_idx - 1] if sx > 0 else []) + ([s_idx + 1] if sx < width - 1 else []) + ([s_idx - width] if sy > 0 else []) + \
              ([s_idx + width] if sy < height - 1 else [])
        for c in nbs:
            trail = int(member[c])
            if trail <= -6 or (trail >= 0 and trail == plid):
                continue
            cy, cx = divmod(c, width)
            by, bx = cy // win_h, cx // win_w
            if by < Nh and bx < Nw and blk_map[by * Nw + bx] >= 0:
                continue
            pt = flat[c]
            ok = not (pt[2] == 0 or pt[2] != pt[2])
            if ok:
                sd = pl.normal[0] * (pt[0] - pl.center[0]) + pl.normal[1] * (pt[1] - pl.center[1]) + pl.normal[2] * (pt[2] - pl.center[2])
                cdist = np.float32(abs(sd))
                ok = float(cdist) * float(cdist) < 9 * pl.mse + 1e-5
            if ok:
                if trail >= 0:
                    n_pl = extracted[trail]
                    if pl.similarity(n_pl) >= P["similarityTh_refine"]:
                        connect(n_pl, pl)
                if cdist < dist[c]:
                    member[c] = plid
                    dist[c] = cdist
                    queue.append((c, plid))
                elif trail < 0:
                    member[c] = trail - 1
            elif trail < 0:
                member[c] = trail - 1

    # ---- one more clustering over the valid planes (:318-327), numbering (:329-343), outputs
    old = extracted
    heap2 = [(p.mse, p.seq, p) for i, p in enumerate(old) if is_valid[i]]
    heapq.heapify(heap2)
    extracted, steps2 = ah_cluster(heap2)
    plidmap = [-1] * len(old)
    for i, op in enumerate(old):
        if not is_valid[i]:
            continue
        r = ds.find(op.rid)
        for j, p in enumerate(extracted):
            if p.rid == r:
                plidmap[i] = j
                break
    lut = np.array(plidmap + [-1], np.int64)                       # member -1.. map to -1 through the last entry
    final = np.where(member >= 0, lut[np.clip(member, 0, len(old) - 1 if old else 0)] if old else -1, -1)
    seg = np.where(final >= 0, final + 1, 0).astype(np.uint8).reshape(height, width)
    planes = np.array([p.normal + p.center + [p.mse, p.curvature, float(p.N), float(p.rid)] for p in extracted], np.float64).reshape(-1, 10)
    members = [np.nonzero(final == j)[0].astype(np.int32) for j in range(len(extracted))]
    return seg, planes, members, steps + steps2
