"""ORACLE — TEST INFRASTRUCTURE ONLY.

ctypes binding of oracle/libdrfe_oracle.so (the CPU restatement of the reference's ORB +
CAPE front end, see orb_oracle.cpp / cape_oracle.cpp).  Only tests/,
__graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may import
this module; the product (dr-slam_b200/) never does.
"""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "libdrfe_oracle.so")


class Keypoint(C.Structure):
    _fields_ = [("x", C.c_float), ("y", C.c_float), ("size", C.c_float), ("angle", C.c_float),
                ("response", C.c_float), ("octave", C.c_int32), ("class_id", C.c_int32)]


class Plane(C.Structure):
    _fields_ = [("nr_pts", C.c_int32), ("min_nr_pts", C.c_int32),
                ("x_acc", C.c_double), ("y_acc", C.c_double), ("z_acc", C.c_double),
                ("xx_acc", C.c_double), ("yy_acc", C.c_double), ("zz_acc", C.c_double),
                ("xy_acc", C.c_double), ("xz_acc", C.c_double), ("yz_acc", C.c_double),
                ("score", C.c_float), ("MSE", C.c_float), ("planar", C.c_int32),
                ("mean", C.c_double * 3), ("normal", C.c_double * 3), ("d", C.c_double)]


KP_DTYPE = np.dtype([("x", "<f4"), ("y", "<f4"), ("size", "<f4"), ("angle", "<f4"),
                     ("response", "<f4"), ("octave", "<i4"), ("class_id", "<i4")])
PLANE_DTYPE = np.dtype([("nr_pts", "<i4"), ("min_nr_pts", "<i4"),
                        ("x_acc", "<f8"), ("y_acc", "<f8"), ("z_acc", "<f8"),
                        ("xx_acc", "<f8"), ("yy_acc", "<f8"), ("zz_acc", "<f8"),
                        ("xy_acc", "<f8"), ("xz_acc", "<f8"), ("yz_acc", "<f8"),
                        ("score", "<f4"), ("MSE", "<f4"), ("planar", "<i4"),
                        ("mean", "<f8", (3,)), ("normal", "<f8", (3,)), ("d", "<f8")], align=True)
class FrameParams(C.Structure):
    _fields_ = [("fx", C.c_float), ("fy", C.c_float), ("cx", C.c_float), ("cy", C.c_float), ("dist", C.c_float * 5),
                ("bf", C.c_float), ("min_x", C.c_float), ("max_x", C.c_float), ("min_y", C.c_float), ("max_y", C.c_float)]


CYL_DTYPE = np.dtype([("radius", "<f4"), ("center", "<f8", (3,)), ("axis", "<f8", (3,))], align=True)
assert KP_DTYPE.itemsize == 28 and PLANE_DTYPE.itemsize == C.sizeof(Plane) and CYL_DTYPE.itemsize == 56


def build(force=False):
    if force or not os.path.exists(_LIB_PATH):
        subprocess.check_call(["make", "-C", _HERE, "-s"] + (["-B"] if force else []))
    return _LIB_PATH


_lib = None


def lib():
    global _lib
    if _lib is None:
        build()
        L = C.CDLL(_LIB_PATH)
        u8p, f32p, i32p = C.POINTER(C.c_uint8), C.POINTER(C.c_float), C.POINTER(C.c_int32)
        L.orc_orb_create.restype = C.c_void_p
        L.orc_orb_create.argtypes = [C.c_int, C.c_float, C.c_int, C.c_int, C.c_int]
        L.orc_orb_destroy.argtypes = [C.c_void_p]
        L.orc_orb_run.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int]
        L.orc_orb_features_per_level.argtypes = [C.c_void_p, C.c_int]
        L.orc_orb_scale_factor.argtypes = [C.c_void_p, C.c_int]
        L.orc_orb_scale_factor.restype = C.c_float
        L.orc_orb_umax.argtypes = [C.c_void_p, C.c_int]
        L.orc_orb_level_size.argtypes = [C.c_void_p, C.c_int, i32p, i32p]
        L.orc_orb_get_level.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_void_p]
        L.orc_orb_get_blurred.argtypes = [C.c_void_p, C.c_int, C.c_void_p]
        L.orc_orb_get_candidates.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_int]
        L.orc_orb_get_level_keypoints.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_int]
        L.orc_orb_level_tie.argtypes = [C.c_void_p, C.c_int]
        L.orc_orb_get_result.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int]
        L.orc_resize_linear_u8.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_int, C.c_int]
        L.orc_border101_u8.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_void_p]
        L.orc_fast9_nms.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_int]
        L.orc_gaussian7_u8.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_void_p]
        L.orc_fast_atan2.argtypes = [C.c_float, C.c_float]
        L.orc_fast_atan2.restype = C.c_float
        L.orc_distribute_quadtree.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int,
                                              C.c_int, C.c_void_p, C.c_int, i32p]
        L.orc_undistort_point.argtypes = [C.c_void_p, C.c_float, C.c_float, C.c_void_p, C.c_void_p]
        L.orc_frame_image_bounds.argtypes = [C.c_void_p, C.c_int, C.c_int]
        L.orc_frame_post.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p,
                                     C.c_void_p, C.c_void_p, C.c_void_p]
        L.orc_cape_create.restype = C.c_void_p
        L.orc_cape_create.argtypes = [C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_float, C.c_float]
        L.orc_cape_destroy.argtypes = [C.c_void_p]
        L.orc_cape_depth_to_cloud.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_float, C.c_float,
                                              C.c_float, C.c_float, C.c_void_p]
        L.orc_cape_process.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, i32p,
                                       C.c_void_p, C.c_int, i32p]
        L.orc_cape_get_cells.argtypes = [C.c_void_p, C.c_void_p]
        L.orc_cape_get_grid_maps.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p]
        L.orc_cape_cylinders_found.argtypes = [C.c_void_p]
        L.orc_cape_get_cyl_maps.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p]
        L.orc_glibc_rand.argtypes = [C.c_uint32, C.c_int, C.c_void_p]
        L.orc_integral_normals.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_float, C.c_float, C.c_void_p, C.c_void_p]
        L.orc_peac_fit.argtypes = [C.c_void_p, C.c_int, C.c_void_p]
        L.orc_peac_cloud.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_float, C.c_float, C.c_float, C.c_float, C.c_float, C.c_void_p]
        L.orc_peac_run.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_int, C.c_void_p,
                                   C.c_void_p, C.c_int, i32p]
        L.orc_voxel_grid.argtypes = [C.c_void_p, C.c_int, C.c_float, C.c_void_p, i32p]
        L.orc_third_cloud.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_float, C.c_float, C.c_float, C.c_float, C.c_float, C.c_void_p]
        _lib = L
    return _lib


def _p(a):
    return a.ctypes.data_as(C.c_void_p)


class OrbOracle:
    """CPU restatement of Planar_SLAM::ORBextractor (ORBextractor.cc:410-1132)."""

    def __init__(self, nfeatures=1000, scale_factor=1.2, nlevels=8, ini_th=20, min_th=7):
        self.L = lib()
        self.nlevels = nlevels
        self.nfeatures = nfeatures
        self.h = self.L.orc_orb_create(nfeatures, scale_factor, nlevels, ini_th, min_th)

    def __del__(self):
        if getattr(self, "h", None):
            self.L.orc_orb_destroy(self.h)
            self.h = None

    def features_per_level(self):
        return [self.L.orc_orb_features_per_level(self.h, l) for l in range(self.nlevels)]

    def scale_factors(self):
        return [self.L.orc_orb_scale_factor(self.h, l) for l in range(self.nlevels)]

    def umax(self):
        return [self.L.orc_orb_umax(self.h, v) for v in range(16)]

    def run(self, gray):
        gray = np.ascontiguousarray(gray, dtype=np.uint8)
        self._gray = gray
        return self.L.orc_orb_run(self.h, _p(gray), gray.shape[1], gray.shape[0], gray.strides[0])

    def level_size(self, level):
        w, h = C.c_int32(), C.c_int32()
        self.L.orc_orb_level_size(self.h, level, C.byref(w), C.byref(h))
        return w.value, h.value

    def level(self, level, bordered=False):
        w, h = self.level_size(level)
        if bordered:
            w, h = w + 38, h + 38
        out = np.empty((h, w), np.uint8)
        self.L.orc_orb_get_level(self.h, level, int(bordered), _p(out))
        return out

    def blurred(self, level):
        w, h = self.level_size(level)
        out = np.empty((h, w), np.uint8)
        n = self.L.orc_orb_get_blurred(self.h, level, _p(out))
        return out if n else None

    def candidates(self, level):
        cap = 1 << 17
        buf = np.empty((cap, 3), np.float32)
        n = self.L.orc_orb_get_candidates(self.h, level, _p(buf), cap)
        assert n <= cap
        return buf[:n].copy()

    def level_keypoints(self, level):
        cap = self.nfeatures * 4 + 64
        buf = np.empty(cap, KP_DTYPE)
        n = self.L.orc_orb_get_level_keypoints(self.h, level, _p(buf), cap)
        return buf[:n].copy()

    def level_tie(self, level):
        return self.L.orc_orb_level_tie(self.h, level)

    def result(self):
        cap = self.nfeatures * 4 + 64
        kps = np.empty(cap, KP_DTYPE)
        desc = np.empty((cap, 32), np.uint8)
        n = self.L.orc_orb_get_result(self.h, _p(kps), _p(desc), cap)
        return kps[:n].copy(), desc[:n].copy()

    def extract(self, gray):
        self.run(gray)
        return self.result()


def resize_linear(src, dw, dh):
    src = np.ascontiguousarray(src, np.uint8)
    dst = np.empty((dh, dw), np.uint8)
    lib().orc_resize_linear_u8(_p(src), src.shape[1], src.shape[0], _p(dst), dw, dh)
    return dst


def _resize_tabs(ssize, dsize, vertical=False):
    """OpenCV resize.cpp (INTER_LINEAR): fx = (float)((dx + 0.5) * scale - 0.5), sx = floor(fx), fx -= sx.  Horizontally an
    index outside the image is clamped AND its fraction set to 0 (xofs / alpha); vertically only the two ROW indices are
    clipped (clip(sy0 - ksize2 + 1 + k, 0, ssize.height)) while beta keeps the fraction — which differs when upscaling."""
    scale = 1.0 / (dsize / ssize)
    d = np.arange(dsize)
    fx = ((d + 0.5) * scale - 0.5).astype(np.float32)
    sx = np.floor(fx).astype(np.int64)
    fx = (fx - sx.astype(np.float32)).astype(np.float32)
    if vertical:
        return np.clip(sx, 0, ssize - 1), np.clip(sx + 1, 0, ssize - 1), (np.float32(1) - fx).astype(np.float32), fx
    lo = sx < 0
    fx[lo] = 0; sx[lo] = 0
    hi = sx >= ssize - 1
    fx[hi] = 0; sx[hi] = ssize - 1
    return sx, np.minimum(sx + 1, ssize - 1), (np.float32(1) - fx).astype(np.float32), fx


def resize_input(src, dw, dh):
    """cv::resize(src, dst, Size(dw, dh)) with the default INTER_LINEAR as System::TrackRGBD applies it to the colour image
    and the depth map (reference src/System.cc:325-329) — numpy restatement of OpenCV's resizeGeneric_ (HResizeLinear /
    VResizeLinear): 8U with 11-bit fixed-point coefficients, 16U / 32F with separately rounded float products, 16U stored
    through saturate_cast<ushort>(cvRound).  An exact 2 x 2 reduction is switched to INTER_AREA by cv::resize (resize.cpp:
    "if (interpolation == INTER_LINEAR && is_area_fast && iscale_x == 2 && iscale_y == 2) interpolation = INTER_AREA"): the
    rounded integer mean (a + b + c + d + 2) >> 2 for 16U; for 8U and 32F the bilinear expressions give the same values.
    src: (H, W) or (H, W, C) of uint8 / uint16 / float32."""
    src = np.asarray(src)
    sh, sw = src.shape[:2]
    if src.dtype == np.uint16 and sw == 2 * dw and sh == 2 * dh:
        q = src.astype(np.int64)
        return ((q[0::2, 0::2] + q[0::2, 1::2] + q[1::2, 0::2] + q[1::2, 1::2] + 2) >> 2).astype(np.uint16)
    x0, x1, a0, a1 = _resize_tabs(sw, dw)
    y0, y1, b0, b1 = _resize_tabs(sh, dh, vertical=True)
    s = src if src.ndim == 3 else src[:, :, None]
    if src.dtype == np.uint8:
        f32 = np.float32
        c0, c1 = np.rint(a0 * f32(2048)).astype(np.int64), np.rint(a1 * f32(2048)).astype(np.int64)
        e0, e1 = np.rint(b0 * f32(2048)).astype(np.int64), np.rint(b1 * f32(2048)).astype(np.int64)
        s = s.astype(np.int64)
        T = s[:, x0] * c0[None, :, None] + s[:, x1] * c1[None, :, None]
        D = (((e0[:, None, None] * (T[y0] >> 4)) >> 16) + ((e1[:, None, None] * (T[y1] >> 4)) >> 16) + 2) >> 2
        out = D.astype(np.uint8)
    else:
        s = s.astype(np.float32)
        T = ((s[:, x0] * a0[None, :, None]).astype(np.float32) + (s[:, x1] * a1[None, :, None]).astype(np.float32)).astype(np.float32)
        D = ((T[y0] * b0[:, None, None]).astype(np.float32) + (T[y1] * b1[:, None, None]).astype(np.float32)).astype(np.float32)
        out = np.clip(np.rint(D), 0, 65535).astype(np.uint16) if src.dtype == np.uint16 else D
    return out if src.ndim == 3 else out[:, :, 0]


def border101(src, b):
    src = np.ascontiguousarray(src, np.uint8)
    dst = np.empty((src.shape[0] + 2 * b, src.shape[1] + 2 * b), np.uint8)
    lib().orc_border101_u8(_p(src), src.shape[1], src.shape[0], b, _p(dst))
    return dst


def fast9_nms(img, threshold):
    """cv::FAST(img, threshold, nonmax=True) restatement -> (n,3) [x, y, response]."""
    assert img.dtype == np.uint8 and img.strides[1] == 1
    cap = img.shape[0] * img.shape[1] // 4 + 16
    buf = np.empty((cap, 3), np.float32)
    n = lib().orc_fast9_nms(_p(img), img.strides[0], img.shape[1], img.shape[0], threshold, _p(buf), cap)
    return buf[:n].copy()


def gaussian7(src):
    src = np.ascontiguousarray(src, np.uint8)
    dst = np.empty_like(src)
    lib().orc_gaussian7_u8(_p(src), src.shape[1], src.shape[0], _p(dst))
    return dst


def frame_params(fx, fy, cx, cy, dist, bf, width, height):
    """mK / mDistCoef / mbf + Frame::ComputeImageBounds (Frame.cc:863-891)"""
    p = FrameParams(fx, fy, cx, cy, (C.c_float * 5)(*dist), bf, 0, 0, 0, 0)
    lib().orc_frame_image_bounds(C.byref(p), width, height)
    return p


def frame_post(p, keys, depth):
    """UndistortKeyPoints + ComputeStereoFromRGBD + AssignFeaturesToGrid (Frame.cc:835-911, 224-237)"""
    keys = np.ascontiguousarray(keys, KP_DTYPE)
    depth = np.ascontiguousarray(depth, np.float32)
    n = len(keys)
    ku = np.zeros(n, KP_DTYPE)
    ur, kd = np.zeros(n, np.float32), np.zeros(n, np.float32)
    gc, gi = np.zeros(64 * 48, np.uint16), np.zeros(max(n, 1), np.uint16)
    placed = lib().orc_frame_post(C.byref(p), _p(keys), n, _p(depth), depth.shape[1], _p(ku), _p(ur), _p(kd), _p(gc), _p(gi))
    return ku, ur, kd, gc, gi[:placed].copy()


def search_by_projection(p, keys_un, u_right, grid_count, grid_index, desc, queries, qdesc, occupied=None):
    """The inner loops of ORBmatcher::SearchByProjection(Frame&, const vector<MapPoint*>&, th) (reference
    src/ORBmatcher.cc:69-116) restated literally, float32 arithmetic step by step: Frame::GetFeaturesInArea
    (src/Frame.cc:730-779) over mGrid (grid_count [64][48] + grid_index: the cells' lists concatenated x-major, as
    frame_post returns them), the occupied / right-coordinate skips, ORBmatcher::DescriptorDistance (:1712-1728) and the
    best / second-best update.  queries: records (x, y, r, xr, min_level, max_level); returns records
    (best_dist, best_idx, best_level, best_dist2, best_level2)."""
    f32 = np.float32
    gc = np.asarray(grid_count).reshape(64, 48)
    off = np.concatenate([[0], np.cumsum(gc.ravel())]).astype(np.int64)
    inv_w = f32(64) / f32(f32(p.max_x) - f32(p.min_x))
    inv_h = f32(48) / f32(f32(p.max_y) - f32(p.min_y))
    out = np.zeros(len(queries), [("best_dist", "<i4"), ("best_idx", "<i4"), ("best_level", "<i4"), ("best_dist2", "<i4"), ("best_level2", "<i4")])
    kx, ky, ko = keys_un["x"], keys_un["y"], keys_un["octave"]
    d32 = np.ascontiguousarray(desc).view(np.uint32).reshape(len(desc), 8)
    for qi, q in enumerate(queries):
        x, y, r, xr = f32(q["x"]), f32(q["y"]), f32(q["r"]), f32(q["xr"])
        lo, hi = int(q["min_level"]), int(q["max_level"])
        best, best_idx, best_lvl, best2, best_lvl2 = 256, -1, -1, 256, -1
        vind = []
        cx0 = max(0, int(np.floor(f32(f32(f32(x - f32(p.min_x)) - r) * inv_w))))
        cx1 = min(63, int(np.ceil(f32(f32(f32(x - f32(p.min_x)) + r) * inv_w))))
        cy0 = max(0, int(np.floor(f32(f32(f32(y - f32(p.min_y)) - r) * inv_h))))
        cy1 = min(47, int(np.ceil(f32(f32(f32(y - f32(p.min_y)) + r) * inv_h))))
        if cx0 < 64 and cx1 >= 0 and cy0 < 48 and cy1 >= 0:
            check = lo > 0 or hi >= 0
            for ix in range(cx0, cx1 + 1):
                for iy in range(cy0, cy1 + 1):
                    c = ix * 48 + iy
                    for j in range(off[c], off[c + 1]):
                        idx = int(grid_index[j])
                        if check:
                            if ko[idx] < lo:
                                continue
                            if hi >= 0 and ko[idx] > hi:
                                continue
                        if abs(f32(kx[idx] - x)) < r and abs(f32(ky[idx] - y)) < r:
                            vind.append(idx)
        qd = np.ascontiguousarray(qdesc[qi]).view(np.uint32)
        for idx in vind:
            if occupied is not None and occupied[idx]:
                continue
            if u_right[idx] > 0 and abs(f32(xr - u_right[idx])) > r:
                continue
            dist = 0
            for k in range(8):                       # the reference's bit tricks = a population count per word
                v = int(qd[k] ^ d32[idx, k])
                v = v - ((v >> 1) & 0x55555555)
                v = (v & 0x33333333) + ((v >> 2) & 0x33333333)
                dist += ((((v + (v >> 4)) & 0xF0F0F0F) * 0x1010101) & 0xFFFFFFFF) >> 24
            if dist < best:
                best2, best, best_lvl2, best_lvl, best_idx = best, dist, best_lvl, int(ko[idx]), idx
            elif dist < best2:
                best_lvl2, best2 = int(ko[idx]), dist
        out[qi] = (best, best_idx, best_lvl, best2, best_lvl2)
    return out


def search_local_points(p, keys_un, u_right, grid_count, grid_index, desc, queries, qdesc, qflags, nnratio, occupied=None):
    """ORBmatcher::SearchByProjection(Frame& F, const vector<MapPoint*>& vpMapPoints, th) (reference src/ORBmatcher.cc:46-130)
    whole, restated literally: the sequential loop over the map points with the skips of :55-59 (qflags & LP_VALID), the
    window search, best / second best, the ratio test (:119-122) and the in-order assignment F.mvpMapPoints[bestIdx] = pMP
    (:124), which later map points see through Observations() > 0 (:88-90).  queries as in search_by_projection (r already
    multiplied by th and the scale factor).  Returns (match records, assigned, holder, nmatches)."""
    f32 = np.float32
    gc = np.asarray(grid_count).reshape(64, 48)
    off = np.concatenate([[0], np.cumsum(gc.ravel())]).astype(np.int64)
    d32 = np.ascontiguousarray(desc).view(np.uint32).reshape(len(desc), 8)
    n = len(keys_un)
    out = np.zeros(len(queries), [("best_dist", "<i4"), ("best_idx", "<i4"), ("best_level", "<i4"), ("best_dist2", "<i4"), ("best_level2", "<i4")])
    out["best_dist"], out["best_idx"], out["best_level"], out["best_dist2"], out["best_level2"] = 256, -1, -1, 256, -1
    assigned = np.full(len(queries), -1, np.int32)
    holder = np.full(n, -1, np.int32)
    held_observed = (np.zeros(n, bool) if occupied is None else (np.asarray(occupied[:n]) != 0)).copy()
    ko = keys_un["octave"]
    nmatches = 0
    for i, q in enumerate(queries):
        if not (qflags[i] & LP_VALID):
            continue
        r, xr = f32(q["r"]), f32(q["xr"])
        vind = features_in_area(p, keys_un, off, grid_index, q["x"], q["y"], r, int(q["min_level"]), int(q["max_level"]))
        if not vind:
            continue
        qd = np.ascontiguousarray(qdesc[i]).view(np.uint32)
        best, best_idx, best_lvl, best2, best_lvl2 = 256, -1, -1, 256, -1
        for idx in vind:
            if held_observed[idx]:
                continue
            if u_right[idx] > 0 and abs(f32(xr - u_right[idx])) > r:
                continue
            dist = descriptor_distance(qd, d32[idx])
            if dist < best:
                best2, best, best_lvl2, best_lvl, best_idx = best, dist, best_lvl, int(ko[idx]), idx
            elif dist < best2:
                best_lvl2, best2 = int(ko[idx]), dist
        out[i] = (best, best_idx, best_lvl, best2, best_lvl2)
        if best <= TH_HIGH:
            if best_lvl == best_lvl2 and f32(best) > f32(f32(nnratio) * f32(best2)):
                continue
            holder[best_idx] = i
            held_observed[best_idx] = bool(qflags[i] & LP_OBSERVED)
            assigned[i] = best_idx
            nmatches += 1
    return out, assigned, holder, nmatches


def features_in_area(p, keys_un, off, grid_index, x, y, r, lo=-1, hi=-1):
    """Frame::GetFeaturesInArea(x, y, r, minLevel, maxLevel) (reference src/Frame.cc:730-779), float32 step by step;
    off = exclusive prefix sum of the [64][48] cell counts."""
    f32 = np.float32
    x, y, r = f32(x), f32(y), f32(r)
    inv_w = f32(64) / f32(f32(p.max_x) - f32(p.min_x))
    inv_h = f32(48) / f32(f32(p.max_y) - f32(p.min_y))
    kx, ky, ko = keys_un["x"], keys_un["y"], keys_un["octave"]
    vind = []
    cx0 = max(0, int(np.floor(f32(f32(f32(x - f32(p.min_x)) - r) * inv_w))))
    if cx0 >= 64:
        return vind
    cx1 = min(63, int(np.ceil(f32(f32(f32(x - f32(p.min_x)) + r) * inv_w))))
    if cx1 < 0:
        return vind
    cy0 = max(0, int(np.floor(f32(f32(f32(y - f32(p.min_y)) - r) * inv_h))))
    if cy0 >= 48:
        return vind
    cy1 = min(47, int(np.ceil(f32(f32(f32(y - f32(p.min_y)) + r) * inv_h))))
    if cy1 < 0:
        return vind
    check = lo > 0 or hi >= 0
    for ix in range(cx0, cx1 + 1):
        for iy in range(cy0, cy1 + 1):
            c = ix * 48 + iy
            for j in range(off[c], off[c + 1]):
                idx = int(grid_index[j])
                if check:
                    if ko[idx] < lo:
                        continue
                    if hi >= 0 and ko[idx] > hi:
                        continue
                if abs(f32(kx[idx] - x)) < r and abs(f32(ky[idx] - y)) < r:
                    vind.append(idx)
    return vind


def descriptor_distance(a32, b32):
    """ORBmatcher::DescriptorDistance (reference src/ORBmatcher.cc:1712-1728): the bit tricks, word by word"""
    dist = 0
    for k in range(8):
        v = int(a32[k] ^ b32[k])
        v = v - ((v >> 1) & 0x55555555)
        v = (v & 0x33333333) + ((v >> 2) & 0x33333333)
        dist += ((((v + (v >> 4)) & 0xF0F0F0F) * 0x1010101) & 0xFFFFFFFF) >> 24
    return dist


LAST_POINT_DTYPE = np.dtype([("X", "<f4"), ("Y", "<f4"), ("Z", "<f4"), ("angle", "<f4"), ("octave", "<i4"), ("flags", "<i4")])
LP_VALID, LP_OBSERVED = 1, 2
TH_HIGH, HISTO_LENGTH = 100, 30


def search_last_frame(p, scale_factors, keys_un, u_right, grid_count, grid_index, desc, Tcw, th, mode, check_orientation,
                      points, pdesc, occupied=None):
    """ORBmatcher::SearchByProjection(Frame& CurrentFrame, const Frame& LastFrame, th, bMono) (reference
    src/ORBmatcher.cc:1396-1535) restated literally: the sequential loop over the last frame's points with the in-order
    assignment CurrentFrame.mvpMapPoints[bestIdx2] = pMP, the rotation histogram, ComputeThreeMaxima (:1666-1707) and the
    removal of the other bins.  Rcw*x3Dw+tcw follows cv::gemm's small-matrix float path (products and sums in float, left
    to right, then + c; tests/test_search_last_frame.py checks that model against cv2.gemm); no FMA contraction.
    mode: 0 neither, 1 bForward, 2 bBackward (:1413-1414, computed by the caller).  mvpMapPoints of the current frame is
    modelled as `holder[idx]` = index of the last-frame point it holds (-1: untouched, -2: set to NULL by the rotation
    check); occupied[idx] != 0: holds an observed map point on entry.
    Returns (match_key, match_dist, holder, nmatches)."""
    f32, f64 = np.float32, np.float64
    T = np.asarray(Tcw, f32).reshape(3, 4)
    gc = np.asarray(grid_count).reshape(64, 48)
    off = np.concatenate([[0], np.cumsum(gc.ravel())]).astype(np.int64)
    d32 = np.ascontiguousarray(desc).view(np.uint32).reshape(len(desc), 8)
    n = len(keys_un)
    holder = np.full(n, -1, np.int32)
    held_observed = np.zeros(n, bool) if occupied is None else (np.asarray(occupied[:n]) != 0)
    held_observed = held_observed.copy()
    match_key = np.full(len(points), -1, np.int32)
    match_dist = np.full(len(points), 256, np.int32)
    rot_hist = [[] for _ in range(HISTO_LENGTH)]
    factor = f32(1.0) / f32(HISTO_LENGTH)
    nmatches = 0
    th = f32(th)
    fx, fy, cx, cy, bf = f32(p.fx), f32(p.fy), f32(p.cx), f32(p.cy), f32(p.bf)
    with np.errstate(all="ignore"):
        for i, lp in enumerate(points):
            if not (lp["flags"] & LP_VALID):
                continue
            X = (f32(lp["X"]), f32(lp["Y"]), f32(lp["Z"]))
            c3 = []
            for r in range(3):
                t0 = f32(f32(f32(T[r, 0] * X[0]) + f32(T[r, 1] * X[1])) + f32(T[r, 2] * X[2]))
                c3.append(f32(f64(t0) + f64(T[r, 3])))
            xc, yc = c3[0], c3[1]
            invzc = f32(f64(1.0) / f64(c3[2]))
            if invzc < 0:
                continue
            u = f32(f32(f32(fx * xc) * invzc) + cx)
            v = f32(f32(f32(fy * yc) * invzc) + cy)
            if u < f32(p.min_x) or u > f32(p.max_x):
                continue
            if v < f32(p.min_y) or v > f32(p.max_y):
                continue
            if np.isnan(u) or np.isnan(v):      # z == 0 and x (y) == 0: undefined in the reference ((int)floor(NaN)); declared: no match
                continue
            octave = int(lp["octave"])
            radius = f32(th * f32(scale_factors[octave]))
            if mode == 1:
                vind = features_in_area(p, keys_un, off, grid_index, u, v, radius, octave, -1)
            elif mode == 2:
                vind = features_in_area(p, keys_un, off, grid_index, u, v, radius, 0, octave)
            else:
                vind = features_in_area(p, keys_un, off, grid_index, u, v, radius, octave - 1, octave + 1)
            if not vind:
                continue
            dmp = np.ascontiguousarray(pdesc[i]).view(np.uint32)
            best, best_idx = 256, -1
            for i2 in vind:
                if held_observed[i2]:
                    continue
                if u_right[i2] > 0:
                    ur = f32(u - f32(bf * invzc))
                    er = abs(f32(ur - u_right[i2]))
                    if er > radius:
                        continue
                dist = descriptor_distance(dmp, d32[i2])
                if dist < best:
                    best, best_idx = dist, i2
            match_dist[i] = best
            if best <= TH_HIGH:
                holder[best_idx] = i
                held_observed[best_idx] = bool(lp["flags"] & LP_OBSERVED)
                match_key[i] = best_idx
                nmatches += 1
                if check_orientation:
                    rot = f32(f32(lp["angle"]) - f32(keys_un["angle"][best_idx]))
                    if rot < 0.0:
                        rot = f32(rot + f32(360.0))
                    t = f32(rot * factor)
                    b = int(np.floor(f64(t) + 0.5)) if t >= 0 else -int(np.floor(-f64(t) + 0.5))   # round(): half away from zero
                    if b == HISTO_LENGTH:
                        b = 0
                    assert 0 <= b < HISTO_LENGTH
                    rot_hist[b].append(best_idx)
    if check_orientation:
        ind1 = ind2 = ind3 = -1
        max1 = max2 = max3 = 0
        for i in range(HISTO_LENGTH):
            s = len(rot_hist[i])
            if s > max1:
                max3, max2, max1 = max2, max1, s
                ind3, ind2, ind1 = ind2, ind1, i
            elif s > max2:
                max3, max2 = max2, s
                ind3, ind2 = ind2, i
            elif s > max3:
                max3, ind3 = s, i
        if f32(max2) < f32(f32(0.1) * f32(max1)):
            ind2 = ind3 = -1
        elif f32(max3) < f32(f32(0.1) * f32(max1)):
            ind3 = -1
        for i in range(HISTO_LENGTH):
            if i != ind1 and i != ind2 and i != ind3:
                for idx in rot_hist[i]:
                    holder[idx] = -2
                    nmatches -= 1
    return match_key, match_dist, holder, nmatches


class Vocabulary:
    """DBoW2 TemplatedVocabulary<FORB::TDescriptor, FORB> as ORB-SLAM uses it (reference Thirdparty/DBoW2/DBoW2/
    TemplatedVocabulary.h), built from the arrays loadFromTextFile parses (:1338-1422): row i = node i + 1 of the file."""

    def __init__(self, k, L, scoring, weighting, parent, is_leaf, descriptors, weights):
        self.k, self.L, self.scoring, self.weighting = k, L, scoring, weighting
        n = len(parent)
        self.children = [[] for _ in range(n + 1)]
        self.desc = np.zeros((n + 1, 32), np.uint8)
        self.weight = [0.0] * (n + 1)
        self.word_id = [0] * (n + 1)
        self.nwords = 0
        for i in range(n):
            nid = i + 1
            self.children[int(parent[i])].append(nid)            # :1391
            self.desc[nid] = descriptors[i]
            self.weight[nid] = float(weights[i])
            if is_leaf[i]:                                       # :1408-1415
                self.word_id[nid] = self.nwords
                self.nwords += 1
        self.desc32 = self.desc.view(np.uint32)

    def transform_one(self, feature32, levelsup):
        """transform(feature, word_id, weight, nid, levelsup) (:1217-1258)"""
        nid_level = self.L - levelsup
        nid = 0 if nid_level <= 0 else None
        final_id, current_level = 0, 0
        while True:
            current_level += 1
            nodes = self.children[final_id]
            final_id = nodes[0]
            best_d = descriptor_distance(feature32, self.desc32[final_id])   # FORB::distance (FORB.cpp:81-101)
            for cid in nodes[1:]:
                d = descriptor_distance(feature32, self.desc32[cid])
                if d < best_d:
                    best_d, final_id = d, cid
            if current_level == nid_level:
                nid = final_id
            if not self.children[final_id]:
                break
        if nid is None:        # *nid is uninitialised in the reference when a leaf sits above level L - levelsup; declared: the leaf
            nid = final_id
        return self.word_id[final_id], self.weight[final_id], nid

    def transform(self, descriptors, levelsup=4):
        """transform(features, BowVector&, FeatureVector&, levelsup) (:1126-1194) with BowVector::addWeight /
        addIfNotExist / normalize (BowVector.cpp:34-84) and FeatureVector::addFeature (FeatureVector.cpp:31-45).
        Returns (word_id[i], node_id[i]) per descriptor (-1: stopped), the BowVector as [(word, value)] and the
        FeatureVector as [(node, [indices])], both in std::map iteration order."""
        d32 = np.ascontiguousarray(descriptors, np.uint8).reshape(-1, 32).view(np.uint32)
        v, fv = {}, {}
        words, nodes = np.full(len(d32), -1, np.int32), np.full(len(d32), -1, np.int32)
        must = self.scoring != 5                                  # ScoringObject.h:73-89
        for i in range(len(d32)):
            wid, w, nid = self.transform_one(d32[i], levelsup)
            if w > 0:
                if self.weighting <= 1:                           # TF_IDF, TF
                    if wid in v:
                        v[wid] += w
                    else:
                        v[wid] = w
                elif wid not in v:                                # IDF, BINARY
                    v[wid] = w
                fv.setdefault(nid, []).append(i)
                words[i], nodes[i] = wid, nid
        keys = sorted(v)
        if self.weighting <= 1 and v and not must:
            nd = float(len(v))
            for k in keys:
                v[k] /= nd
        if must:
            norm = 0.0
            if self.scoring != 1:                                 # L1
                for k in keys:
                    norm += abs(v[k])
            else:
                for k in keys:
                    norm += v[k] * v[k]
                norm = float(np.sqrt(np.float64(norm)))
            if norm > 0.0:
                for k in keys:
                    v[k] /= norm
        return words, nodes, [(k, v[k]) for k in keys], [(k, fv[k]) for k in sorted(fv)]


def synth_vocabulary(k, L, seed, scoring=0, weighting=0, ragged=False):
    """A random k-ary tree of depth L in the file layout of ORBvoc.txt (nodes written level by level, a node's children
    next to each other): child descriptors = the parent's with bits flipped (so descents are decided by few bits and
    ties occur), idf-like weights, a few stopped words (weight 0); ragged: some subtrees end early and some nodes have
    fewer children."""
    rng = np.random.default_rng(seed)
    parent, leaf, desc, wt = [], [], [], []
    frontier = [(0, np.zeros(32, np.uint8), 0)]
    nid = 0
    while frontier:
        nxt = []
        for pid, pdesc, depth in frontier:
            nk = k if not ragged else int(rng.integers(2, k + 1))
            for c in range(nk):
                nid += 1
                d = pdesc.copy()
                for b in rng.integers(0, 256, max(2, 96 >> depth)):
                    d[b >> 3] ^= 1 << (b & 7)
                if c and rng.random() < 0.08:
                    d = desc[-1].copy()                           # duplicate of the previous sibling: a tie, the first wins
                is_leaf = depth + 1 == L or (ragged and depth + 1 >= 3 and rng.random() < 0.1)
                parent.append(pid); leaf.append(1 if is_leaf else 0); desc.append(d)
                wt.append((0.0 if rng.random() < 0.03 else float(rng.uniform(0.5, 9.0))) if is_leaf else 0.0)
                if not is_leaf:
                    nxt.append((nid, d, depth + 1))
        frontier = nxt
    return dict(k=k, L=L, scoring=scoring, weighting=weighting, parent=np.array(parent, np.int32), is_leaf=np.array(leaf, np.uint8),
                descriptors=np.array(desc, np.uint8), weights=np.array(wt, np.float64))


TH_LOW = 50


def search_by_bow(kf_desc, kf_angle, kf_valid, kf_fv, f_desc, f_angle, f_fv, nnratio=0.7, check_orientation=True):
    """ORBmatcher::SearchByBoW(KeyFrame* pKF, Frame& F, vector<MapPoint*>& vpMapPointMatches) (reference
    src/ORBmatcher.cc:160-292) restated literally: the two FeatureVectors ([(node, [indices])] in key order) walked in
    step (:186-261; lower_bound on a mismatch), per keyframe feature with a good map point (kf_valid) the best / second
    best among the frame's features of the node that hold no match yet, TH_LOW and ratio test, rotation histogram,
    ComputeThreeMaxima and the removal of the other bins.  vpMapPointMatches is modelled as f_match[idx] = keyframe
    feature index or -1 (NULL).  Returns (kf_match before the rotation check, f_match, nmatches)."""
    f32 = np.float32
    kd32 = np.ascontiguousarray(kf_desc, np.uint8).reshape(-1, 32).view(np.uint32)
    fd32 = np.ascontiguousarray(f_desc, np.uint8).reshape(-1, 32).view(np.uint32)
    f_match = np.full(len(fd32), -1, np.int32)
    kf_match = np.full(len(kd32), -1, np.int32)
    rot_hist = [[] for _ in range(HISTO_LENGTH)]
    factor = f32(1.0) / f32(HISTO_LENGTH)
    nmatches = 0
    ki, fi = 0, 0
    while ki < len(kf_fv) and fi < len(f_fv):
        if kf_fv[ki][0] == f_fv[fi][0]:
            for real_kf in kf_fv[ki][1]:
                if not kf_valid[real_kf]:
                    continue
                best1, best_idx, best2 = 256, -1, 256
                for real_f in f_fv[fi][1]:
                    if f_match[real_f] >= 0:
                        continue
                    dist = descriptor_distance(kd32[real_kf], fd32[real_f])
                    if dist < best1:
                        best2, best1, best_idx = best1, dist, real_f
                    elif dist < best2:
                        best2 = dist
                if best1 <= TH_LOW:
                    if f32(best1) < f32(f32(nnratio) * f32(best2)):
                        f_match[best_idx] = real_kf
                        kf_match[real_kf] = best_idx
                        if check_orientation:
                            rot = f32(f32(kf_angle[real_kf]) - f32(f_angle[best_idx]))
                            if rot < 0.0:
                                rot = f32(rot + f32(360.0))
                            b = int(np.floor(np.float64(f32(rot * factor)) + 0.5))     # round(); rot >= 0 here
                            if b == HISTO_LENGTH:
                                b = 0
                            assert 0 <= b < HISTO_LENGTH
                            rot_hist[b].append(best_idx)
                        nmatches += 1
            ki += 1
            fi += 1
        elif kf_fv[ki][0] < f_fv[fi][0]:
            while ki < len(kf_fv) and kf_fv[ki][0] < f_fv[fi][0]:     # lower_bound
                ki += 1
        else:
            while fi < len(f_fv) and f_fv[fi][0] < kf_fv[ki][0]:
                fi += 1
    if check_orientation:
        ind1 = ind2 = ind3 = -1
        max1 = max2 = max3 = 0
        for i in range(HISTO_LENGTH):
            sz = len(rot_hist[i])
            if sz > max1:
                max3, max2, max1 = max2, max1, sz
                ind3, ind2, ind1 = ind2, ind1, i
            elif sz > max2:
                max3, max2 = max2, sz
                ind3, ind2 = ind2, i
            elif sz > max3:
                max3, ind3 = sz, i
        if f32(max2) < f32(f32(0.1) * f32(max1)):
            ind2 = ind3 = -1
        elif f32(max3) < f32(f32(0.1) * f32(max1)):
            ind3 = -1
        for i in range(HISTO_LENGTH):
            if i == ind1 or i == ind2 or i == ind3:
                continue
            for idx in rot_hist[i]:
                f_match[idx] = -1
                nmatches -= 1
    return kf_match, f_match, nmatches


MATCH_DTYPE = np.dtype([("best_dist", "<i4"), ("best_idx", "<i4"), ("best_level", "<i4"), ("best_dist2", "<i4"), ("best_level2", "<i4")])
QUERY_DTYPE = np.dtype([("x", "<f4"), ("y", "<f4"), ("r", "<f4"), ("xr", "<f4"), ("min_level", "<i4"), ("max_level", "<i4")])


def search_last_frame_cpp(p, scale_factors, keys_un, u_right, grid_count, grid_index, desc, Tcw, th, mode, check_orientation,
                          points, pdesc, occupied=None):
    """the C++ restatement of the same function (oracle/match_oracle.cpp, std::vector / push_back as the reference) — same
    arguments and results as search_last_frame"""
    keys_un = np.ascontiguousarray(keys_un, KP_DTYPE)
    n, m = len(keys_un), len(points)
    sf = np.ascontiguousarray(scale_factors, np.float32)
    ur = np.ascontiguousarray(u_right, np.float32)
    gc = np.ascontiguousarray(np.asarray(grid_count).ravel(), np.uint16)
    gi = np.ascontiguousarray(grid_index, np.uint16)
    d = np.ascontiguousarray(desc, np.uint8)
    T = np.ascontiguousarray(Tcw, np.float32)
    pts = np.ascontiguousarray(points, LAST_POINT_DTYPE)
    pd = np.ascontiguousarray(pdesc, np.uint8)
    occ = None if occupied is None else np.ascontiguousarray(occupied[:n], np.uint8)
    mk, md, holder = np.zeros(m, np.int32), np.zeros(m, np.int32), np.zeros(n, np.int32)
    f = lib().orc_search_last_frame
    f.restype = C.c_int
    f.argtypes = [C.c_void_p] * 4 + [C.c_int] + [C.c_void_p] * 4 + [C.c_float, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_int] + [C.c_void_p] * 4
    nm = f(C.addressof(p), _p(sf), _p(keys_un), _p(ur), n, _p(gc), _p(gi), _p(d), _p(T), float(th), int(mode), int(check_orientation), _p(pts), _p(pd), m,
           None if occ is None else _p(occ), _p(mk), _p(md), _p(holder))
    return mk, md, holder, nm


def search_local_points_cpp(p, keys_un, u_right, grid_count, grid_index, desc, queries, qdesc, qflags, nnratio, occupied=None):
    """the C++ restatement of search_local_points (oracle/match_oracle.cpp)"""
    keys_un = np.ascontiguousarray(keys_un, KP_DTYPE)
    n, m = len(keys_un), len(queries)
    ur = np.ascontiguousarray(u_right, np.float32)
    gc = np.ascontiguousarray(np.asarray(grid_count).ravel(), np.uint16)
    gi = np.ascontiguousarray(grid_index, np.uint16)
    d = np.ascontiguousarray(desc, np.uint8)
    q = np.ascontiguousarray(queries, QUERY_DTYPE)
    qd = np.ascontiguousarray(qdesc, np.uint8)
    fl = np.ascontiguousarray(qflags, np.uint8)
    occ = None if occupied is None else np.ascontiguousarray(occupied[:n], np.uint8)
    out, asg, holder = np.zeros(m, MATCH_DTYPE), np.zeros(m, np.int32), np.zeros(n, np.int32)
    f = lib().orc_search_local_points
    f.restype = C.c_int
    f.argtypes = [C.c_void_p] * 3 + [C.c_int] + [C.c_void_p] * 6 + [C.c_int, C.c_float] + [C.c_void_p] * 4
    nm = f(C.addressof(p), _p(keys_un), _p(ur), n, _p(gc), _p(gi), _p(d), _p(q), _p(qd), _p(fl), m, float(nnratio), None if occ is None else _p(occ),
           _p(out), _p(asg), _p(holder))
    return out, asg, holder, nm


def bow_transform_cpp(voc, descriptors, levelsup=4):
    """the C++ restatement of Vocabulary.transform with the reference's std::maps (oracle/match_oracle.cpp): same results"""
    d = np.ascontiguousarray(descriptors, np.uint8).reshape(-1, 32)
    n = len(d)
    parent = np.ascontiguousarray(voc["parent"], np.int32)
    leaf = np.ascontiguousarray(voc["is_leaf"], np.uint8)
    vd = np.ascontiguousarray(voc["descriptors"], np.uint8)
    wt = np.ascontiguousarray(voc["weights"], np.float64)
    words, nodes = np.zeros(n, np.int32), np.zeros(n, np.int32)
    bk, bv = np.zeros(max(n, 1), np.int32), np.zeros(max(n, 1), np.float64)
    fn = C.c_int(0)
    fnode, fstart, ffeat = np.zeros(max(n, 1), np.int32), np.zeros(n + 1, np.int32), np.zeros(max(n, 1), np.int32)
    f = lib().orc_bow_transform
    f.restype = C.c_int
    f.argtypes = [C.c_int] * 4 + [C.c_void_p] * 5 + [C.c_int, C.c_int] + [C.c_void_p] * 8
    nb = f(int(voc["L"]), int(voc["scoring"]), int(voc["weighting"]), len(parent), _p(parent), _p(leaf), _p(vd), _p(wt), _p(d), n, int(levelsup),
           _p(words), _p(nodes), _p(bk), _p(bv), C.addressof(fn), _p(fnode), _p(fstart), _p(ffeat))
    bow = [(int(bk[j]), float(bv[j])) for j in range(nb)]
    fv = [(int(fnode[j]), ffeat[fstart[j]:fstart[j + 1]].tolist()) for j in range(fn.value)]
    return words, nodes, bow, fv


def _flat_fv(fv):
    node = np.array([k for k, _ in fv], np.int32)
    start = np.cumsum([0] + [len(l) for _, l in fv]).astype(np.int32)
    feat = np.array([i for _, l in fv for i in l], np.int32)
    return np.ascontiguousarray(node), np.ascontiguousarray(start), np.ascontiguousarray(feat if len(feat) else np.zeros(1, np.int32))


def search_by_bow_cpp(kf_desc, kf_angle, kf_valid, kf_fv, f_desc, f_angle, f_fv, nnratio=0.7, check_orientation=True):
    """the C++ restatement of search_by_bow (oracle/match_oracle.cpp, the reference's map iterators and lower_bound)"""
    kd = np.ascontiguousarray(kf_desc, np.uint8).reshape(-1, 32)
    fd = np.ascontiguousarray(f_desc, np.uint8).reshape(-1, 32)
    ka, fa = np.ascontiguousarray(kf_angle, np.float32), np.ascontiguousarray(f_angle, np.float32)
    kv = np.ascontiguousarray(kf_valid, np.uint8)
    kn, ks, kf = _flat_fv(kf_fv)
    fn, fs, ff = _flat_fv(f_fv)
    km, fm = np.zeros(len(kd), np.int32), np.zeros(len(fd), np.int32)
    f = lib().orc_search_by_bow
    f.restype = C.c_int
    f.argtypes = [C.c_void_p] * 3 + [C.c_int, C.c_int] + [C.c_void_p] * 5 + [C.c_int, C.c_int] + [C.c_void_p] * 3 + [C.c_float, C.c_int, C.c_void_p, C.c_void_p]
    nm = f(_p(kd), _p(ka), _p(kv), len(kd), len(kf_fv), _p(kn) if len(kn) else None, _p(ks), _p(kf), _p(fd), _p(fa), len(fd), len(f_fv),
           _p(fn) if len(fn) else None, _p(fs), _p(ff), float(nnratio), int(check_orientation), _p(km), _p(fm))
    return km, fm, nm


def peac_params(**kw):
    """ahc::ParamSet defaults (AHCParamSet.hpp:68-78; DR-SLAM never changes them) as the 11 doubles orc_peac_run takes"""
    import math
    d = dict(depthSigma=1.6e-6, stdTol_init=5.0, stdTol_merge=8.0, z_near=500.0, z_far=4000.0, angle_near=15.0 * math.pi / 180.0,
             angle_far=90.0 * math.pi / 180.0, similarityTh_merge=math.cos(60.0 * math.pi / 180.0),
             similarityTh_refine=math.cos(30.0 * math.pi / 180.0), depthAlpha=0.04, depthChangeTol=0.02)
    d.update(kw)
    return np.array([d[k] for k in ("depthSigma", "stdTol_init", "stdTol_merge", "z_near", "z_far", "angle_near", "angle_far",
                                    "similarityTh_merge", "similarityTh_refine", "depthAlpha", "depthChangeTol")], np.float64)


def peac_cloud(depth16, depth_factor, fx, fy, cx, cy):
    """PlaneDetection::readDepthImage (PlaneExtractor.cpp:28-55): (H, W) uint16 -> (H*W, 3) float64, z > 5 culled"""
    depth16 = np.ascontiguousarray(depth16, np.uint16)
    H, W = depth16.shape
    out = np.empty((H * W, 3), np.float64)
    lib().orc_peac_cloud(_p(depth16), W, H, W, depth_factor, fx, fy, cx, cy, _p(out))
    return out


def peac_fit(sums9, n):
    """ahc::PlaneSeg::Stats::compute (AHCPlaneSeg.hpp:128-162) on {sx, sy, sz, sxx, syy, szz, sxy, syz, sxz}, N -> center, normal, mse, curvature"""
    s = np.ascontiguousarray(sums9, np.float64)
    out = np.zeros(8, np.float64)
    lib().orc_peac_fit(_p(s), int(n), _p(out))
    return out[:3].copy(), out[3:6].copy(), float(out[6]), float(out[7])


def peac_run(cloud, width, height, params=None, min_support=3000, window=10, plane_cap=256):
    """ahc::PlaneFitter::run (doRefine) -> (seg_output (H, W) uint8, planes (n, 12) float64 [normal, center, mse, curvature, N, rid],
    list of pixel-index arrays (plane_vertices_), cluster steps)"""
    cloud = np.ascontiguousarray(cloud, np.float64)
    prm = peac_params() if params is None else np.ascontiguousarray(params, np.float64)
    seg = np.zeros((height, width), np.uint8)
    planes = np.zeros((plane_cap, 12), np.float64)
    offs = np.zeros(plane_cap + 1, np.int32)
    idx = np.zeros(width * height, np.int32)
    steps = C.c_int32(0)
    ww, wh = (window, window) if np.isscalar(window) else window
    n = lib().orc_peac_run(_p(cloud), width, height, _p(prm), min_support, ww, wh, _p(seg), _p(planes), plane_cap, _p(offs), _p(idx),
                           len(idx), C.byref(steps))
    if n < 0:
        raise RuntimeError("orc_peac_run: capacity")
    return seg, planes[:n].copy(), [idx[offs[p]:offs[p + 1]].copy() for p in range(n)], steps.value


def voxel_grid(xyz, leaf=0.05):
    """pcl::VoxelGrid<PointT>::filter with setLeafSize(leaf, leaf, leaf) on an (n, 3) float32 point list (reference
    src/Frame.cc:1121-1125) -> ((m, 3) centroids in ascending leaf order, unfiltered flag)"""
    xyz = np.ascontiguousarray(xyz, np.float32).reshape(-1, 3)
    out = np.empty_like(xyz)
    flag = C.c_int32(0)
    m = lib().orc_voxel_grid(_p(xyz), len(xyz), leaf, _p(out), C.byref(flag))
    return out[:m].copy(), bool(flag.value)


def third_cloud(depth, fx, fy, cx, cy, max_point_dist):
    """the 1/3-resolution cloud of Frame::ComputePlanes_CAPE (reference src/Frame.cc:1153-1172) -> (ceil(H/3), ceil(W/3), 3)"""
    depth = np.asarray(depth, np.float32)
    assert depth.strides[1] == 4
    H, W = depth.shape
    out = np.empty(((H + 2) // 3, (W + 2) // 3, 3), np.float32)
    lib().orc_third_cloud(_p(depth), W, H, depth.strides[0] // 4, fx, fy, cx, cy, max_point_dist, _p(out))
    return out


def integral_normals(cloud, max_depth_change_factor=0.05, smoothing_size=10.0, with_distance_map=False):
    """pcl::IntegralImageNormalEstimation (AVERAGE_3D_GRADIENT) as DR-SLAM configures it (reference src/Frame.cc:1174-1187) on an
    organized (h, w, 3) float cloud -> (h, w, 3) normals, NaN where PCL leaves NaN [, the distance map]"""
    cloud = np.ascontiguousarray(cloud, np.float32)
    h, w, _ = cloud.shape
    out = np.empty((h, w, 3), np.float32)
    dm = np.empty((h, w), np.float32) if with_distance_map else None
    lib().orc_integral_normals(_p(cloud), w, h, max_depth_change_factor, smoothing_size, _p(out), _p(dm) if with_distance_map else None)
    return (out, dm) if with_distance_map else out


def glibc_rand(seed, n):
    out = np.zeros(n, np.int32)
    lib().orc_glibc_rand(int(seed), n, _p(out))
    return out


def fast_atan2(y, x):
    return lib().orc_fast_atan2(float(y), float(x))


def distribute_quadtree(xyr, min_x, max_x, min_y, max_y, n_want):
    xyr = np.ascontiguousarray(xyr, np.float32)
    out = np.empty((len(xyr) + 8, 3), np.float32)
    tie = C.c_int32(0)
    n = lib().orc_distribute_quadtree(_p(xyr), len(xyr), min_x, max_x, min_y, max_y, n_want, _p(out),
                                      len(out), C.byref(tie))
    return out[:n].copy(), tie.value


class CapeOracle:
    """CPU restatement of CAPE (CAPE.cpp) + the DR-SLAM wrapper (PlaneExtractor.cpp:111-191)."""

    def __init__(self, height=480, width=640, cell_w=20, cell_h=20, cylinder=False,
                 min_cos=float(np.float32(np.cos(np.pi / 12))), max_merge_dist=50.0):
        self.L = lib()
        self.H, self.W, self.cw, self.ch = height, width, cell_w, cell_h
        self.ncx, self.ncy = width // cell_w, height // cell_h
        self.h = self.L.orc_cape_create(height, width, cell_w, cell_h, int(cylinder), min_cos, max_merge_dist)
        if not self.h:
            raise RuntimeError("oracle: orc_cape_create failed")

    def __del__(self):
        if getattr(self, "h", None):
            self.L.orc_cape_destroy(self.h)
            self.h = None

    def depth_to_cloud(self, depth, fx, fy, cx, cy):
        depth = np.ascontiguousarray(depth, np.float32)
        cloud = np.zeros(3 * self.H * self.W, np.float32)
        self.L.orc_cape_depth_to_cloud(self.h, _p(depth), depth.strides[0] // 4, fx, fy, cx, cy, _p(cloud))
        return cloud

    def process(self, cloud, plane_cap=256):
        cloud = np.ascontiguousarray(cloud, np.float32)
        seg = np.zeros((self.H, self.W), np.uint8)
        planes = np.zeros(plane_cap, PLANE_DTYPE)
        npl, ncy = C.c_int32(0), C.c_int32(0)
        self.L.orc_cape_process(self.h, _p(cloud), _p(seg), _p(planes), plane_cap, C.byref(npl), None, 0,
                                C.byref(ncy))
        return seg, planes[:npl.value].copy()

    def plane_points(self, cloud, seg, nr_planes):
        """plane_cloud of PlaneDetection_CAPE::runPlaneDetection (reference src/PlaneExtractor.cpp:165-190): for every
        pixel in row-major order whose seg_output code is > 0, its cloud point goes to plane_cloud[code - 1].
        cloud: the cell-major organised cloud (organizePointCloudByCell, :80-99) — the un-organised cloud_array the
        reference reads is the same points in image order.  Returns a list of (n, 3) float32 arrays."""
        H, W, cw, ch = self.H, self.W, self.cw, self.ch
        r, c = np.mgrid[0:H, 0:W]
        idx = ((r // ch) * (W // cw) + (c // cw)) * (cw * ch) + (r % ch) * cw + (c % cw)     # cell_map, :135-148
        xyz = np.asarray(cloud, np.float32).reshape(3, H * W)[:, idx.ravel()].T            # image order, (H*W, 3)
        code = np.asarray(seg).ravel()
        return [xyz[code == p + 1] for p in range(nr_planes)]

    def process_full(self, cloud, plane_cap=256, cyl_cap=64):
        """CAPE::process with cylinder detection: seg_output, planes, nr_cylinders_final and the
        cylinder_segments_final list (all cylinders found, CAPE.cpp:434-445)."""
        cloud = np.ascontiguousarray(cloud, np.float32)
        seg = np.zeros((self.H, self.W), np.uint8)
        planes = np.zeros(plane_cap, PLANE_DTYPE)
        cyls = np.zeros(cyl_cap, CYL_DTYPE)
        npl, ncy = C.c_int32(0), C.c_int32(0)
        self.L.orc_cape_process(self.h, _p(cloud), _p(seg), _p(planes), plane_cap, C.byref(npl), _p(cyls), cyl_cap,
                                C.byref(ncy))
        found = self.L.orc_cape_cylinders_found(self.h)
        return seg, planes[:npl.value].copy(), ncy.value, cyls[:found].copy()

    def cyl_maps(self):
        cm = np.zeros((self.ncy, self.ncx), np.int32)
        em = np.zeros((self.ncy, self.ncx), np.uint8)
        self.L.orc_cape_get_cyl_maps(self.h, _p(cm), _p(em))
        return cm, em

    def cells(self):
        out = np.zeros(self.ncx * self.ncy, PLANE_DTYPE)
        self.L.orc_cape_get_cells(self.h, _p(out))
        return out

    def grid_maps(self):
        pm = np.zeros((self.ncy, self.ncx), np.int32)
        em = np.zeros((self.ncy, self.ncx), np.uint8)
        self.L.orc_cape_get_grid_maps(self.h, _p(pm), _p(em))
        return pm, em
