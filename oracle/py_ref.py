"""ORACLE tier T1 — TEST INFRASTRUCTURE ONLY (golden generator / cross-check).

An independent Python restatement of the reference front end that calls the REAL OpenCV
(cv2) for every OpenCV-owned primitive the reference calls (cv::resize, copyMakeBorder,
cv::FAST, GaussianBlur, fastAtan2, erode, dilate) and restates only the reference-owned
logic literally.  It exists to pin the dependency-free C++ oracle (orb_oracle.cpp /
cape_oracle.cpp): tests/test_oracle_vs_cv2.py requires both to agree, and
tests/golden/make_golden.py uses it to write the committed fixtures.

Reference lines followed: ORBextractor.cc:77-147 (IC_Angle, computeOrbDescriptor),
:410-470 (ctor), :481-763 (DivideNode, DistributeOctTree), :765-853
(ComputeKeyPointsOctTree), :1043-1132 (operator(), ComputePyramid);
PlaneExtractor.cpp:80-152; CAPE.cpp:47-506; PlaneSeg.cpp:8-142; Histogram.cpp:8-70.
Needs cv2, so it only runs where cv2 is importable (this container); nothing at run time
on the GPU box depends on it.
"""
import ctypes
import math
import os
import re

import numpy as np

_libm = ctypes.CDLL("libm.so.6")
_libm.cosf.restype = ctypes.c_float
_libm.cosf.argtypes = [ctypes.c_float]
_libm.sinf.restype = ctypes.c_float
_libm.sinf.argtypes = [ctypes.c_float]

f32 = np.float32
EDGE = 19
HALF_PATCH = 15
PATCH = 31


def cv_round(v):
    """cvRound: round half to even."""
    return int(np.rint(v))


def load_pattern():
    here = os.path.dirname(os.path.abspath(__file__))
    txt = open(os.path.join(here, "..", "include", "drfe_orb_pattern.inc")).read()
    txt = re.sub(r"/\*.*?\*/", "", txt, flags=re.S)
    v = np.array([int(t) for t in re.findall(r"-?\d+", txt)], np.int32)
    assert v.size == 1024
    return v.reshape(512, 2)


class OrbRef:
    def __init__(self, nfeatures=1000, scale_factor=1.2, nlevels=8, ini_th=20, min_th=7):
        import cv2  # noqa: F401  (fail early where cv2 is missing)
        self.nfeatures, self.nlevels, self.ini_th, self.min_th = nfeatures, nlevels, ini_th, min_th
        sf = float(f32(scale_factor))  # ctor takes float, member is double
        self.scale = [f32(1.0)]
        for _ in range(1, nlevels):
            self.scale.append(f32(float(self.scale[-1]) * sf))
        self.inv_scale = [f32(1.0) / s for s in self.scale]
        factor = f32(1.0 / sf)
        want = f32(nfeatures) * (f32(1) - factor) / (f32(1) - f32(math.pow(float(factor), float(nlevels))))
        self.per_level = []
        tot = 0
        for _ in range(nlevels - 1):
            self.per_level.append(cv_round(want))
            tot += self.per_level[-1]
            want = f32(want * factor)
        self.per_level.append(max(nfeatures - tot, 0))
        umax = [0] * (HALF_PATCH + 1)
        r = f32(HALF_PATCH) * np.sqrt(f32(2.0)) / f32(2)
        vmax = int(math.floor(r + f32(1)))
        vmin = int(math.ceil(r))
        for v in range(vmax + 1):
            umax[v] = cv_round(math.sqrt(HALF_PATCH * HALF_PATCH - v * v))
        v0 = 0
        for v in range(HALF_PATCH, vmin - 1, -1):
            while umax[v0] == umax[v0 + 1]:
                v0 += 1
            umax[v] = v0
            v0 += 1
        self.umax = umax
        self.pattern = load_pattern()

    # ---- ComputePyramid
    def pyramid(self, image):
        import cv2
        levels, bordered = [], []
        h, w = image.shape
        for l in range(self.nlevels):
            s = self.inv_scale[l]
            sz = (cv_round(f32(w) * s), cv_round(f32(h) * s))
            if l == 0:
                cur = image.copy()
            else:
                cur = cv2.resize(levels[l - 1], sz, interpolation=cv2.INTER_LINEAR)
            levels.append(cur)
            bordered.append(cv2.copyMakeBorder(cur, EDGE, EDGE, EDGE, EDGE, cv2.BORDER_REFLECT_101))
        return levels, bordered

    # ---- DistributeOctTree, literal list semantics
    @staticmethod
    def _divide(node):
        x0, y0, x1, y1 = node["box"]
        hx = int(math.ceil(float(f32(x1 - x0) / f32(2))))
        hy = int(math.ceil(float(f32(y1 - y0) / f32(2))))
        mx, my = x0 + hx, y0 + hy
        ch = [dict(box=(x0, y0, mx, my), keys=[]), dict(box=(mx, y0, x1, my), keys=[]),
              dict(box=(x0, my, mx, y1), keys=[]), dict(box=(mx, my, x1, y1), keys=[])]
        for k in node["keys"]:
            if k[0] < mx:
                (ch[0] if k[1] < my else ch[2])["keys"].append(k)
            else:
                (ch[1] if k[1] < my else ch[3])["keys"].append(k)
        for c in ch:
            c["nomore"] = len(c["keys"]) == 1
        return ch

    def distribute(self, keys, min_x, max_x, min_y, max_y, N):
        """keys: list of (x, y, response) -> retained list in the reference's list order."""
        n_ini = int(np.round(f32(max_x - min_x) / f32(max_y - min_y)))  # C round(): no .5 cases here
        hx = f32(max_x - min_x) / f32(n_ini)
        seq = [0]

        def mk(box):
            seq[0] += 1
            return dict(box=box, keys=[], nomore=False, seq=seq[0])
        roots = [mk((int(hx * f32(i)), 0, int(hx * f32(i + 1)), max_y - min_y)) for i in range(n_ini)]
        for k in keys:
            roots[int(f32(k[0]) / hx)]["keys"].append(k)
        nodes = []
        for r in roots:
            if len(r["keys"]) == 1:
                r["nomore"] = True
            if r["keys"]:
                nodes.append(r)

        def drop(node):
            for i, n in enumerate(nodes):
                if n is node:
                    del nodes[i]
                    return

        def split(node, expandable):
            for c in self._divide(node):
                if c["keys"]:
                    seq[0] += 1
                    c["seq"] = seq[0]
                    nodes.insert(0, c)
                    if len(c["keys"]) > 1:
                        expandable.append(c)

        finish = False
        while not finish:
            prev = len(nodes)
            expandable = []
            n_expand = 0
            snapshot = list(nodes)
            for nd in snapshot:
                if nd["nomore"]:
                    continue
                before = len(expandable)
                split(nd, expandable)
                n_expand += len(expandable) - before
                drop(nd)
            if len(nodes) >= N or len(nodes) == prev:
                finish = True
            elif len(nodes) + n_expand * 3 > N:
                while not finish:
                    prev = len(nodes)
                    todo = sorted(expandable, key=lambda c: (len(c["keys"]), c["seq"]))
                    expandable = []
                    for nd in reversed(todo):
                        split(nd, expandable)
                        drop(nd)
                        if len(nodes) >= N:
                            break
                    if len(nodes) >= N or len(nodes) == prev:
                        finish = True
        out = []
        for nd in nodes:
            best = nd["keys"][0]
            for k in nd["keys"][1:]:
                if k[2] > best[2]:
                    best = k
            out.append(best)
        return out

    # ---- ComputeKeyPointsOctTree (per level, before orientation)
    def detect_level(self, img, level):
        import cv2
        h, w = img.shape
        min_bx = min_by = EDGE - 3
        max_bx, max_by = w - EDGE + 3, h - EDGE + 3
        width, height = f32(max_bx - min_bx), f32(max_by - min_by)
        n_cols, n_rows = int(width / f32(30)), int(height / f32(30))
        w_cell = int(math.ceil(float(width / f32(n_cols))))
        h_cell = int(math.ceil(float(height / f32(n_rows))))
        det_ini = cv2.FastFeatureDetector_create(self.ini_th, True)
        det_min = cv2.FastFeatureDetector_create(self.min_th, True)
        cands = []
        for i in range(n_rows):
            ini_y = min_by + i * h_cell
            max_y = ini_y + h_cell + 6
            if ini_y >= max_by - 3:
                continue
            max_y = min(max_y, max_by)
            for j in range(n_cols):
                ini_x = min_bx + j * w_cell
                max_x = ini_x + w_cell + 6
                if ini_x >= max_bx - 6:
                    continue
                max_x = min(max_x, max_bx)
                cell = np.ascontiguousarray(img[ini_y:max_y, ini_x:max_x])
                kps = det_ini.detect(cell)
                if not kps:
                    kps = det_min.detect(cell)
                for kp in kps:
                    cands.append((kp.pt[0] + j * w_cell, kp.pt[1] + i * h_cell, kp.response))
        kept = self.distribute(cands, min_bx, max_bx, min_by, max_by, self.per_level[level]) if cands else []
        size = float(int(f32(PATCH) * self.scale[level]))
        kps = [dict(x=k[0] + min_bx, y=k[1] + min_by, size=size, response=k[2], octave=level) for k in kept]
        return cands, kps

    def ic_angle(self, bordered, x, y):
        import cv2
        cy, cx = cv_round(y) + EDGE, cv_round(x) + EDGE
        m01 = m10 = 0
        for v in range(-HALF_PATCH, HALF_PATCH + 1):
            d = self.umax[abs(v)]
            row = bordered[cy + v, cx - d:cx + d + 1].astype(np.int64)
            u = np.arange(-d, d + 1)
            m10 += int((u * row).sum())
            m01 += v * int(row.sum())
        return cv2.fastAtan2(float(m01), float(m10))

    def descriptor(self, blurred, x, y, angle_deg):
        factor_pi = f32(math.pi / 180.0)
        ang = f32(angle_deg) * factor_pi
        a = f32(_libm.cosf(float(ang)))
        b = f32(_libm.sinf(float(ang)))
        cy, cx = cv_round(y), cv_round(x)
        px = self.pattern[:, 0].astype(f32)
        py = self.pattern[:, 1].astype(f32)
        rr = np.rint(px * b + py * a).astype(np.int64)   # float32 ops, no FMA in numpy
        cc = np.rint(px * a - py * b).astype(np.int64)
        vals = blurred[cy + rr, cx + cc].astype(np.int32)
        bits = (vals[0::2] < vals[1::2]).astype(np.uint8)  # 256 tests
        return np.packbits(bits.reshape(32, 8), axis=1, bitorder="little").reshape(32)

    def extract(self, image, keep_intermediates=False):
        import cv2
        levels, bordered = self.pyramid(image)
        all_c, all_k = [], []
        for l in range(self.nlevels):
            c, k = self.detect_level(levels[l], l)
            all_c.append(c)
            all_k.append(k)
        for l in range(self.nlevels):
            for kp in all_k[l]:
                kp["angle"] = self.ic_angle(bordered[l], kp["x"], kp["y"])
        out_k, out_d, blurred = [], [], []
        for l in range(self.nlevels):
            if not all_k[l]:
                blurred.append(None)
                continue
            bl = cv2.GaussianBlur(levels[l], (7, 7), 2, sigmaY=2, borderType=cv2.BORDER_REFLECT_101)
            blurred.append(bl)
            for kp in all_k[l]:
                out_d.append(self.descriptor(bl, kp["x"], kp["y"], kp["angle"]))
                s = self.scale[l] if l else f32(1)
                out_k.append((f32(kp["x"]) * s if l else f32(kp["x"]), f32(kp["y"]) * s if l else f32(kp["y"]),
                              kp["size"], kp["angle"], kp["response"], l, -1))
        res = dict(kps=np.array(out_k, dtype=[("x", "<f4"), ("y", "<f4"), ("size", "<f4"), ("angle", "<f4"),
                                               ("response", "<f4"), ("octave", "<i4"), ("class_id", "<i4")]),
                   desc=np.array(out_d, np.uint8).reshape(-1, 32))
        if keep_intermediates:
            res.update(levels=levels, bordered=bordered, cands=all_c, level_kps=all_k, blurred=blurred)
        return res


# --------------------------------------------------------------------------- CAPE
def eigen_sum_f32(v):
    """Declared Eigen MatrixXf::sum() tree for len(v) % 16 == 0 (+ tails)."""
    v = np.asarray(v, f32)
    n = v.size
    body = (n // 16) * 16
    if body == 0:
        s = f32(v[0])
        for x in v[1:]:
            s = f32(s + x)
        return s
    lanes = v[:16].copy()
    for i in range(16, body, 16):
        lanes = (lanes + v[i:i + 16]).astype(f32)
    p8 = (lanes[:8] + lanes[8:]).astype(f32)
    rest = body
    if n - body >= 8:
        p8 = (p8 + v[body:body + 8]).astype(f32)
        rest = body + 8
    p4 = (p8[:4] + p8[4:]).astype(f32)
    p2 = (p4[:2] + p4[2:]).astype(f32)
    s = f32(p2[0] + p2[1])
    for x in v[rest:]:
        s = f32(s + x)
    return s


class Seg:
    FIELDS = ("x_acc", "y_acc", "z_acc", "xx_acc", "yy_acc", "zz_acc", "xy_acc", "xz_acc", "yz_acc")

    def __init__(self):
        self.nr_pts = 0
        self.min_nr_pts = 0
        for f in self.FIELDS:
            setattr(self, f, 0.0)
        self.score = f32(0)
        self.MSE = f32(0)
        self.planar = False
        self.mean = np.zeros(3)
        self.normal = np.zeros(3)
        self.d = 0.0

    def copy(self):
        o = Seg()
        o.__dict__.update({k: (v.copy() if isinstance(v, np.ndarray) else v) for k, v in self.__dict__.items()})
        return o

    def expand(self, o):
        for f in self.FIELDS:
            setattr(self, f, getattr(self, f) + getattr(o, f))
        self.nr_pts += o.nr_pts

    def fit(self):
        n = self.nr_pts
        self.mean = np.array([self.x_acc / n, self.y_acc / n, self.z_acc / n])
        cov = np.array([[self.xx_acc - self.x_acc * self.x_acc / n, self.xy_acc - self.x_acc * self.y_acc / n,
                         self.xz_acc - self.x_acc * self.z_acc / n],
                        [0, self.yy_acc - self.y_acc * self.y_acc / n, self.yz_acc - self.y_acc * self.z_acc / n],
                        [0, 0, self.zz_acc - self.z_acc * self.z_acc / n]])
        cov = cov + np.triu(cov, 1).T
        w, v = np.linalg.eigh(cov)  # LAPACK, ascending: stands in for SelfAdjointEigenSolver
        v0 = v[:, 0]
        d = -(v0[0] * self.mean[0] + v0[1] * self.mean[1] + v0[2] * self.mean[2])
        if d > 0:
            self.normal, self.d = v0.copy(), d
        else:
            self.normal, self.d = -v0, -d
        self.MSE = f32(w[0] / n)
        self.score = f32(w[1] / w[0])


def fit_cell(X, Y, Z, cell_w):
    s = Seg()
    npts = X.size
    cell_h = npts // cell_w
    s.min_nr_pts = npts // 2
    s.planar = True
    s.nr_pts = int((Z > 0).sum())
    if s.nr_pts < s.min_nr_pts:
        s.planar = False
        return s
    for (start, stop, step, first_pair) in (
            (cell_w * (cell_h // 2), cell_w * (cell_h // 2) + cell_w, 1, 1),
            (cell_w // 2, npts - cell_w // 2, cell_w, cell_w)):
        i = start
        z_last = max(Z[i], Z[i + first_pair])
        i += step
        jumps = 0
        while i < stop:
            z = Z[i]
            if z > 0 and abs(f32(z - z_last)) < 100.0:
                z_last = z
            elif z > 0:
                jumps += 1
            i += step
        if jumps > 1:
            s.planar = False
            return s
    s.x_acc, s.y_acc, s.z_acc = float(eigen_sum_f32(X)), float(eigen_sum_f32(Y)), float(eigen_sum_f32(Z))
    s.xx_acc, s.yy_acc, s.zz_acc = float(eigen_sum_f32(X * X)), float(eigen_sum_f32(Y * Y)), float(eigen_sum_f32(Z * Z))
    s.xy_acc, s.xz_acc, s.yz_acc = float(eigen_sum_f32(X * Y)), float(eigen_sum_f32(X * Z)), float(eigen_sum_f32(Y * Z))
    s.fit()
    lim = 0.000001425 * s.mean[2] * s.mean[2] + 10
    if float(s.MSE) > lim * lim:
        s.planar = False
    return s


def depth_to_cloud(depth, fx, fy, cx, cy, cell_w, cell_h):
    """PlaneExtractor.cpp:112-152 -> cell-major (3, H*W) float32 (X block, Y block, Z block)."""
    H, W = depth.shape
    z = depth.astype(np.float64)
    jj, ii = np.meshgrid(np.arange(W, dtype=np.float64), np.arange(H, dtype=np.float64))
    x = (jj - float(f32(cx))) * z / float(f32(fx))
    y = (ii - float(f32(cy))) * z / float(f32(fy))
    ncx = W // cell_w
    r, c = np.meshgrid(np.arange(H), np.arange(W), indexing="ij")
    idx = ((r // cell_h) * ncx + (c // cell_w)) * cell_w * cell_h + (r % cell_h) * cell_w + (c % cell_w)
    out = np.zeros((3, H * W), f32)
    out[0, idx.ravel()] = x.astype(f32).ravel()
    out[1, idx.ravel()] = y.astype(f32).ravel()
    out[2, idx.ravel()] = z.astype(f32).ravel()
    return out.reshape(-1)


def glibc_rand_stream(seed=1):
    """glibc rand() (TYPE_3 additive feedback, stdlib/random_r.c) as a generator; declared stream
    of CylinderSeg's RANSAC."""
    r = [0] * 344
    w = seed if seed else 1
    r[0] = w
    for i in range(1, 31):
        hi, lo = int(w / 127773) if w >= 0 else -int(-w / 127773), 0
        lo = w - hi * 127773
        w = 16807 * lo - 2836 * hi
        if w < 0:
            w += 2147483647
        r[i] = w
    for i in range(31, 34):
        r[i] = r[i - 31]
    for i in range(34, 344):
        r[i] = (r[i - 31] + r[i - 3]) & 0xFFFFFFFF
    while True:
        v = (r[-31] + r[-3]) & 0xFFFFFFFF
        r.append(v)
        r.pop(0)
        yield v >> 1


class CylRef:
    """CylinderSeg::CylinderSeg (CylinderSeg.cpp:7-247) with numpy / LAPACK."""

    def __init__(self, grid, act, rng):
        ids = [i for i in range(len(grid)) if act[i]]
        m = len(ids)
        self.local2global = ids
        self.nr_segments = 0
        self.radii, self.centers, self.inliers, self.MSEs = [], [], [], []
        self.P1, self.P2, self.P1P2_norm, self.cylindrical = [], [], [], []
        self.axis = np.zeros(3)
        N = np.array([grid[i].normal for i in ids]).T.copy()      # 3 x m
        P = np.array([grid[i].mean for i in ids]).T.copy()
        NN = np.concatenate([N, -N], axis=1)
        cov = (NN @ NN.T) / float(NN.shape[1] - 1)
        S, V = np.linalg.eigh(cov)
        if S[2] / S[0] < 100:
            return
        vec = V[:, 0].copy()
        self.axis = vec
        Pp = P - np.outer(vec, vec @ P)
        N = N - np.outer(vec, vec @ N)
        N = N / np.sqrt((N * N).sum(0))
        K = f32(math.log(f32(1) - f32(0.8)) / math.log(1 - float(f32(0.33)) ** 3))
        m_left = m
        ids_left = list(range(m))
        left = np.ones(m, bool)
        while m_left > 5 and m_left > 0.1 * m:
            min_hyp = 0.0225 * m_left
            accepted = int(0.9 * m_left)
            max_inl = 0
            I_final = np.zeros(m, bool)
            k = 0
            while k < K:
                i1 = ids_left[next(rng) % m_left]
                i2 = ids_left[next(rng) % m_left]
                i3 = ids_left[next(rng) % m_left]
                e1 = N[:, i1] + N[:, i2] + N[:, i3]
                e2 = Pp[:, i1] + Pp[:, i2] + Pp[:, i3]
                a = 1 - (e1 @ e1) / 9
                b = (N[:, i1] * Pp[:, i1] + N[:, i2] * Pp[:, i2] + N[:, i3] * Pp[:, i3]).sum() / 3 - (e1 @ e2) / 9
                with np.errstate(all="ignore"):
                    r = np.float64(b) / np.float64(a)
                    center = (e2 - r * e1) / 3
                    D = (((Pp - r * N) - center[:, None]) ** 2).sum(0) / (r * r)
                I = D < 0.0225
                dist, inl = 0.0, 0
                for i in range(m):            # sequential MSAC sum (:137-148)
                    if left[i]:
                        if I[i]:
                            inl += 1
                            dist += D[i]
                        else:
                            dist += 0.0225
                if dist < min_hyp:
                    min_hyp, max_inl = dist, inl
                    I_final = I & left
                    if inl > accepted:
                        break
                k += 1
            if max_inl < 6:
                break
            K = f32(math.log(f32(1) - f32(0.8)) / math.log(1 - 0.5 ** 3))
            left = left & ~I_final
            ids_left = [i for i in range(m) if left[i]]
            m_left = len(ids_left)
            e1, e2, b = np.zeros(3), np.zeros(3), 0.0
            for i in np.flatnonzero(I_final):
                e1 = e1 + N[:, i]
                e2 = e2 + Pp[:, i]
                b += (N[:, i] * Pp[:, i]).sum()
            n2 = float(max_inl * max_inl)
            a = 1 - (e1 @ e1) / n2
            b = b / max_inl - (e1 @ e2) / n2
            r = b / a
            center = (e2 - r * e1) / max_inl
            r = abs(r)
            self.nr_segments += 1
            self.radii.append(f32(r))
            self.centers.append(center)
            self.inliers.append(I_final.copy())
            P2d = center + vec
            dirv = P2d - center
            n12 = np.linalg.norm(dirv)
            mse = 0.0
            for i in np.flatnonzero(I_final):
                dd = np.linalg.norm(np.cross(dirv, P[:, i] - P2d)) / n12 - r
                mse += dd * dd
            self.MSEs.append(mse / max_inl)
            self.P1.append(center.astype(f32))
            self.P2.append(P2d.astype(f32))
            self.P1P2_norm.append(f32(n12))
            self.cylindrical.append(True)


def cape_process(cloud, H, W, cw, ch, min_cos, max_merge_dist, cylinder=False):
    """CAPE::process.  Returns seg_output, final planes (list of Seg), cells, plane grid map, eroded map and —
    with cylinder=True — a dict with the cylinder results."""
    import cv2
    import sys
    sys.setrecursionlimit(20000)
    min_cos = f32(min_cos)
    max_merge_dist = f32(max_merge_dist)
    N = H * W
    CX, CY, CZ = cloud[:N], cloud[N:2 * N], cloud[2 * N:]
    ncx, ncy = W // cw, H // ch
    ncells, npc = ncx * ncy, cw * ch
    grid, tols = [], np.zeros(ncells, f32)
    sin_merge = f32(math.sqrt(1 - float(min_cos) ** 2))
    for cid in range(ncells):
        o = cid * npc
        s = fit_cell(CX[o:o + npc], CY[o:o + npc], CZ[o:o + npc], cw)
        grid.append(s)
        if s.planar:
            dx, dy, dz = CX[o + npc - 1] - CX[o], CY[o + npc - 1] - CY[o], CZ[o + npc - 1] - CZ[o]
            diam = np.sqrt(f32(f32(dx * dx + dy * dy) + dz * dz))
            t = min(max(f32(diam * sin_merge), f32(20)), max_merge_dist)
            tols[cid] = f32(t * t)
    nb = 20
    Hh = [0] * (nb * nb)
    B = [-1] * ncells
    unassigned = [False] * ncells
    remaining = 0
    for cid, s in enumerate(grid):
        if not s.planar:
            continue
        nx, ny, nz = s.normal
        pn = math.sqrt(nx * nx + ny * ny)
        polar = math.acos(-nz)
        xq = int((nb - 1) * (polar - 0.0) / (3.14 - 0.0))
        yq = 0
        if xq > 0:
            yq = int((nb - 1) * (math.atan2(nx / pn, ny / pn) - (-3.14)) / (3.14 - (-3.14)))
        B[cid] = yq * nb + xq
        Hh[B[cid]] += 1
        unassigned[cid] = True
        remaining += 1
    plane_map = np.zeros((ncy, ncx), np.int32)
    cyl_map = np.zeros((ncy, ncx), np.int32)
    segs = []
    rng = glibc_rand_stream(1)
    cyl_segments, cyl2region = [], []

    def grow(x, y, n1, d1, act):
        idx = x + ncx * y
        if not unassigned[idx] or act[idx]:
            return
        g = grid[idx]
        if (n1[0] * g.normal[0] + n1[1] * g.normal[1] + n1[2] * g.normal[2] < float(min_cos) or
                (n1[0] * g.mean[0] + n1[1] * g.mean[1] + n1[2] * g.mean[2] + d1) ** 2 > float(tols[idx])):
            return
        act[idx] = True
        if x > 0:
            grow(x - 1, y, g.normal, g.d, act)
        if x < ncx - 1:
            grow(x + 1, y, g.normal, g.d, act)
        if y > 0:
            grow(x, y - 1, g.normal, g.d, act)
        if y < ncy - 1:
            grow(x, y + 1, g.normal, g.d, act)

    while remaining > 0:
        best_bin, best_cnt = -1, 0
        for b in range(nb * nb):
            if Hh[b] > best_cnt:
                best_bin, best_cnt = b, Hh[b]
        cand = [i for i in range(ncells) if B[i] == best_bin] if best_cnt > 0 else []
        if len(cand) < 5:
            break
        seed, min_mse = cand[0], f32(2147483647)
        for i, c in enumerate(cand):
            if grid[c].MSE < min_mse:
                seed = c
                min_mse = grid[i].MSE  # sic (CAPE.cpp:130)
        acc = grid[seed].copy()
        act = [False] * ncells
        grow(seed % ncx, seed // ncx, acc.normal, acc.d, act)
        activated = 0
        for i in range(ncells):
            if act[i]:
                acc.expand(grid[i])
                activated += 1
                Hh[B[i]] -= 1
                B[i] = -1
                unassigned[i] = False
                remaining -= 1
        if activated < 4:
            continue
        acc.fit()
        if acc.score > 100:
            segs.append(acc)
            for i in range(ncells):
                if act[i]:
                    plane_map[i // ncx, i % ncx] = len(segs)
        elif cylinder and activated > 5:      # extrusion (CAPE.cpp:179-216)
            cy = CylRef(grid, act, rng)
            cyl_segments.append(cy)
            for sid in range(cy.nr_segments):
                for f in Seg.FIELDS:
                    setattr(acc, f, 0.0)
                acc.nr_pts = 0
                for c in np.flatnonzero(cy.inliers[sid]):
                    acc.expand(grid[cy.local2global[c]])
                acc.fit()
                if float(acc.MSE) < cy.MSEs[sid]:
                    segs.append(acc.copy())
                    for c in np.flatnonzero(cy.inliers[sid]):
                        g = cy.local2global[c]
                        plane_map[g // ncx, g % ncx] = len(segs)
                    cy.cylindrical[sid] = False
                else:
                    cyl2region.append((len(cyl_segments) - 1, sid))
                    for c in np.flatnonzero(cy.inliers[sid]):
                        g = cy.local2global[c]
                        cyl_map[g // ncx, g % ncx] = len(cyl2region)
    npl = len(segs)
    assoc = np.zeros((npl, npl), bool)
    for r in range(ncy - 1):
        for c in range(ncx - 1):
            v = plane_map[r, c]
            if v > 0:
                if plane_map[r, c + 1] > 0 and v != plane_map[r, c + 1]:
                    assoc[v - 1, plane_map[r, c + 1] - 1] = True
                if plane_map[r + 1, c] > 0 and v != plane_map[r + 1, c]:
                    assoc[v - 1, plane_map[r + 1, c] - 1] = True
    for r in range(npl):
        for c in range(r + 1, npl):
            assoc[r, c] = assoc[r, c] or assoc[c, r]
    merge = list(range(npl))
    for r in range(npl):
        pid = merge[r]
        expanded = False
        for c in range(r + 1, npl):
            if not assoc[r, c]:
                continue
            cosang = float(np.dot(segs[pid].normal, segs[c].normal))
            dd = (segs[r].normal[0] * segs[c].mean[0] + segs[pid].normal[1] * segs[c].mean[1] +
                  segs[pid].normal[2] * segs[c].mean[2] + segs[pid].d)
            if cosang > float(min_cos) and dd * dd < float(max_merge_dist):
                segs[pid].expand(segs[c])
                merge[c] = pid
                expanded = True
        if expanded:
            segs[pid].fit()
    cross = np.array([[0, 1, 0], [1, 1, 1], [0, 1, 0]], np.uint8)
    square = np.ones((3, 3), np.uint8)
    eroded_map = np.zeros((ncy, ncx), np.uint8)
    dist_stack = np.full(N, np.frombuffer(b"\x64\x64\x64\x64", f32)[0], f32)
    seg_stack = np.zeros(N, np.uint8)
    final = []
    for i in range(npl):
        if i != merge[i]:
            continue
        mask = np.zeros((ncy, ncx), np.uint8)
        for j in range(i, npl):
            if merge[j] == merge[i]:
                mask[plane_map == j + 1] = 1
        er = cv2.erode(mask, cross)
        if er.max() == 0:
            continue
        final.append(segs[i])
        di = cv2.dilate(mask, square)
        diff = cv2.subtract(di, er)
        nr = len(final)
        nx, ny, nz, d = (f32(segs[i].normal[0]), f32(segs[i].normal[1]), f32(segs[i].normal[2]), f32(segs[i].d))
        eroded_map[er > 0] = nr
        max_dist = f32(9) * segs[i].MSE
        for cid in np.flatnonzero(diff.ravel() > 0):
            o = cid * npc
            v = ((CX[o:o + npc] * nx + CY[o:o + npc] * ny).astype(f32) + CZ[o:o + npc] * nz).astype(f32) + d
            dist = (v.astype(f32) * v.astype(f32)).astype(f32)
            upd = (dist < max_dist) & (dist < dist_stack[o:o + npc])
            dist_stack[o:o + npc][upd] = dist[upd]
            seg_stack[o:o + npc][upd] = nr
    cyl_eroded = np.zeros((ncy, ncx), np.uint8)
    ncyl_final = 0
    for i, (reg, sid) in enumerate(cyl2region if cylinder else []):   # CAPE.cpp:323-393
        cy = cyl_segments[reg]
        mask = (cyl_map == i + 1).astype(np.uint8)
        er = cv2.erode(mask, cross)
        if er.max() == 0:
            continue
        ncyl_final += 1
        di = cv2.dilate(mask, square)
        diff = cv2.subtract(di, er)
        label = 50 + ncyl_final
        cyl_eroded[er > 0] = label
        P2 = cy.P2[sid]
        P1P2 = (P2 - cy.P1[sid]).astype(f32)
        n12, radius = float(cy.P1P2_norm[sid]), float(cy.radii[sid])
        max_dist = f32(9 * cy.MSEs[sid])
        for cid in np.flatnonzero(diff.ravel() > 0):
            o = cid * npc
            X, Y, Z = CX[o:o + npc], CY[o:o + npc], CZ[o:o + npc]
            q = np.stack([X - P2[0], Y - P2[1], Z - P2[2]]).astype(f32)
            c0 = (P1P2[1] * q[2]).astype(f32) - (P1P2[2] * q[1]).astype(f32)
            c1 = (P1P2[2] * q[0]).astype(f32) - (P1P2[0] * q[2]).astype(f32)
            c2 = (P1P2[0] * q[1]).astype(f32) - (P1P2[1] * q[0]).astype(f32)
            nrm = np.sqrt((c0 * c0 + (c1 * c1 + c2 * c2).astype(f32)).astype(f32)).astype(f32)
            dist = (nrm.astype(np.float64) / n12 - radius).astype(f32)
            dist = (dist * dist).astype(f32)
            upd = (Z > 0) & (dist < max_dist) & (dist < dist_stack[o:o + npc])
            dist_stack[o:o + npc][upd] = dist[upd]
            seg_stack[o:o + npc][upd] = label
    seg = np.zeros((H, W), np.uint8)
    for cr in range(ncy):
        for cc in range(ncx):
            cid = cr * ncx + cc
            blk = seg[cr * ch:(cr + 1) * ch, cc * cw:(cc + 1) * cw]
            if eroded_map[cr, cc] > 0:
                blk[:] = eroded_map[cr, cc]
            elif cyl_eroded[cr, cc] > 0:
                blk[:] = cyl_eroded[cr, cc]
            else:
                st = seg_stack[cid * npc:(cid + 1) * npc].reshape(ch, cw)
                blk[st > 0] = st[st > 0]
    if cylinder:
        cyl = dict(nr_cylinders_final=ncyl_final, cyl_map=cyl_map, cyl_eroded=cyl_eroded,
                   radius=np.array([cyl_segments[r].radii[s_] for r, s_ in cyl2region], f32),
                   center=np.array([cyl_segments[r].centers[s_] for r, s_ in cyl2region]).reshape(-1, 3),
                   axis=np.array([cyl_segments[r].axis for r, s_ in cyl2region]).reshape(-1, 3),
                   mse=np.array([cyl_segments[r].MSEs[s_] for r, s_ in cyl2region]))
        return seg, final, grid, plane_map, eroded_map, cyl
    return seg, final, grid, plane_map, eroded_map
