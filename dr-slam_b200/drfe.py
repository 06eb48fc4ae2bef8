"""ctypes binding of libdrfe.so (the C ABI in include/drfe.h) plus thin Python mirrors of the
reference's operator interfaces, used by the tests and bench:

    ORBextractor(nfeatures, scaleFactor, nlevels, iniThFAST, minThFAST)(image, mask)
        -> Planar_SLAM::ORBextractor::operator()  (reference include/ORBextractor.h:51-61)
    CAPE(depth_height, depth_width, cell_width, cell_height, cylinder_detection,
         min_cos_angle_4_merge, max_merge_dist).process(cloud_array)
        -> CAPE::process  (reference src/CAPE/CAPE.h:47-48)
    PlaneDetectionCAPE  -> PlaneDetection_CAPE::readDepthImage / runPlaneDetection
        (reference include/PlaneExtractor.h:84-115)

There is no CPU fallback: if libdrfe.so is missing or no CUDA device is usable these raise.
"""
import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libdrfe.so")

KP_DTYPE = np.dtype([("x", "<f4"), ("y", "<f4"), ("size", "<f4"), ("angle", "<f4"),
                     ("response", "<f4"), ("octave", "<i4"), ("class_id", "<i4")])
PLANE_DTYPE = np.dtype([("nr_pts", "<i4"), ("min_nr_pts", "<i4"),
                        ("x_acc", "<f8"), ("y_acc", "<f8"), ("z_acc", "<f8"),
                        ("xx_acc", "<f8"), ("yy_acc", "<f8"), ("zz_acc", "<f8"),
                        ("xy_acc", "<f8"), ("xz_acc", "<f8"), ("yz_acc", "<f8"),
                        ("score", "<f4"), ("MSE", "<f4"), ("planar", "<i4"),
                        ("mean", "<f8", (3,)), ("normal", "<f8", (3,)), ("d", "<f8")], align=True)
QUERY_DTYPE = np.dtype([("x", "<f4"), ("y", "<f4"), ("r", "<f4"), ("xr", "<f4"), ("min_level", "<i4"), ("max_level", "<i4")])
MATCH_DTYPE = np.dtype([("best_dist", "<i4"), ("best_idx", "<i4"), ("best_level", "<i4"), ("best_dist2", "<i4"), ("best_level2", "<i4")])
LAST_POINT_DTYPE = np.dtype([("X", "<f4"), ("Y", "<f4"), ("Z", "<f4"), ("angle", "<f4"), ("octave", "<i4"), ("flags", "<i4")])
TRACK_PARAMS_DTYPE = np.dtype([("Tcw", "<f4", (12,)), ("th", "<f4"), ("mode", "<i4"), ("check_orientation", "<i4")])
LP_VALID, LP_OBSERVED = 1, 2
CYL_DTYPE = np.dtype([("radius", "<f4"), ("center", "<f8", (3,)), ("axis", "<f8", (3,))], align=True)

MEM_HOST, MEM_DEVICE = 0, 1
OK, ERR_ARG, ERR_CUDA, ERR_CAPACITY, ERR_STATE = 0, -1, -2, -3, -4


class OrbParams(C.Structure):
    _fields_ = [("nfeatures", C.c_int32), ("scale_factor", C.c_float), ("nlevels", C.c_int32),
                ("ini_th_fast", C.c_int32), ("min_th_fast", C.c_int32)]


class CapeParams(C.Structure):
    _fields_ = [("depth_height", C.c_int32), ("depth_width", C.c_int32), ("cell_width", C.c_int32),
                ("cell_height", C.c_int32), ("cylinder_detection", C.c_int32),
                ("min_cos_angle_4_merge", C.c_float), ("max_merge_dist", C.c_float)]


class FrameParams(C.Structure):
    """mK, mDistCoef, mbf and the image bounds of Frame (Frame.cc:863-891)"""
    _fields_ = [("fx", C.c_float), ("fy", C.c_float), ("cx", C.c_float), ("cy", C.c_float), ("dist", C.c_float * 5),
                ("bf", C.c_float), ("min_x", C.c_float), ("max_x", C.c_float), ("min_y", C.c_float), ("max_y", C.c_float)]


class PeacParams(C.Structure):
    """ahc::ParamSet + the PlaneFitter members DR-SLAM leaves at their defaults"""
    _fields_ = [(n, C.c_double) for n in ("depthSigma", "stdTol_init", "stdTol_merge", "z_near", "z_far", "angle_near", "angle_far",
                                          "similarityTh_merge", "similarityTh_refine", "depthAlpha", "depthChangeTol")] + \
               [("min_support", C.c_int32), ("window_width", C.c_int32), ("window_height", C.c_int32), ("max_depth", C.c_float)]


PEAC_PLANE_DTYPE = np.dtype([("normal", "<f8", (3,)), ("center", "<f8", (3,)), ("mse", "<f8"), ("curvature", "<f8"), ("N", "<i4"), ("rid", "<i4")], align=True)


class PoolParams(C.Structure):
    _fields_ = [("orb", OrbParams), ("cape", CapeParams), ("width", C.c_int32), ("height", C.c_int32),
                ("max_batch", C.c_int32), ("chunk_frames", C.c_int32)]


class DrfeError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__("drfe error %d: %s" % (code, msg))
        self.code = code


# every symbol include/drfe.h declares (tests check the library exports all of them)
SYMBOLS = [
    "drfe_last_error", "drfe_version", "drfe_device_count", "drfe_kernel_launch_count",
    "drfe_event_create", "drfe_event_destroy", "drfe_event_record", "drfe_stream_wait_event",
    "drfe_event_elapsed_ms",
    "drfe_orb_create", "drfe_orb_destroy", "drfe_orb_get_levels", "drfe_orb_get_scale_factor",
    "drfe_orb_get_scale_factors", "drfe_orb_features_per_level", "drfe_orb_max_keypoints",
    "drfe_orb_extract", "drfe_orb_enqueue", "drfe_orb_download", "drfe_orb_sync", "drfe_orb_stream",
    "drfe_orb_extract_batch", "drfe_orb_finish_batch", "drfe_frame_image_bounds", "drfe_orb_frame_post", "drfe_orb_frame_post_shared_depth", "drfe_orb_search_by_projection", "drfe_orb_search_last_frame", "drfe_orb_search_local_points", "drfe_vocab_create", "drfe_vocab_destroy", "drfe_vocab_words", "drfe_orb_compute_bow", "drfe_orb_enqueue_color", "drfe_orb_get_gray", "drfe_orb_search_by_bow",
    "drfe_orb_level_size", "drfe_orb_get_pyramid", "drfe_orb_get_blurred", "drfe_orb_get_candidates",
    "drfe_orb_get_level_keypoints", "drfe_orb_set_profiling", "drfe_orb_stage_times",
    "drfe_cape_create", "drfe_cape_destroy", "drfe_cape_enqueue_cloud", "drfe_cape_enqueue_depth",
    "drfe_cape_enqueue_depth_u16", "drfe_cape_process_depth_batch", "drfe_cape_finish_batch",
    "drfe_cape_download", "drfe_cape_sync", "drfe_cape_stream", "drfe_cape_process",
    "drfe_cape_process_depth", "drfe_cape_num_cells", "drfe_cape_get_cloud", "drfe_cape_get_cells",
    "drfe_cape_get_grid_maps", "drfe_cape_plane_points", "drfe_cape_plane_points_voxel", "drfe_cape_third_cloud", "drfe_cape_third_cloud_normals", "drfe_cape_cylinders_found", "drfe_cape_get_cyl_maps", "drfe_cape_debug_counters", "drfe_cape_set_profiling", "drfe_cape_stage_times",
    "drfe_peac_default_params", "drfe_peac_create", "drfe_peac_destroy", "drfe_peac_stream", "drfe_peac_sync", "drfe_peac_enqueue_depth_u16",
    "drfe_peac_download", "drfe_peac_plane_vertices", "drfe_peac_plane_points_voxel", "drfe_peac_third_cloud_normals", "drfe_peac_debug_counters",
    "drfe_resizer_create", "drfe_resizer_destroy", "drfe_resizer_stream", "drfe_resizer_sync", "drfe_resize",
    "drfe_pool_create", "drfe_pool_destroy", "drfe_pool_num_devices", "drfe_pool_max_keypoints", "drfe_pool_extract_batch",
    "drfe_pool_device_times", "drfe_host_alloc", "drfe_host_free", "drfe_host_register", "drfe_host_unregister",
]

_lib = None


def lib():
    """Load libdrfe.so (raises if it has not been built: there is no fallback)."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise ImportError("libdrfe.so not found at %s — run __graft_entry__.build() "
                          "(make -C dr-slam_b200/csrc); there is no CPU fallback" % LIB_PATH)
    L = C.CDLL(LIB_PATH)
    vp, sz, i32p = C.c_void_p, C.c_size_t, C.POINTER(C.c_int32)
    L.drfe_last_error.restype = C.c_char_p
    L.drfe_version.restype = C.c_char_p
    L.drfe_device_count.argtypes = [i32p]
    L.drfe_kernel_launch_count.restype = C.c_int64
    L.drfe_event_create.argtypes = [C.POINTER(vp)]
    L.drfe_event_destroy.argtypes = [vp]
    L.drfe_event_record.argtypes = [vp, vp]
    L.drfe_stream_wait_event.argtypes = [vp, vp]
    L.drfe_event_elapsed_ms.argtypes = [vp, vp, C.POINTER(C.c_float)]
    L.drfe_orb_create.argtypes = [C.POINTER(OrbParams), C.c_int, C.c_int, C.c_int, C.c_int, C.POINTER(vp)]
    L.drfe_orb_destroy.argtypes = [vp]
    L.drfe_orb_get_levels.argtypes = [vp]
    L.drfe_orb_get_scale_factor.argtypes = [vp]
    L.drfe_orb_get_scale_factor.restype = C.c_float
    L.drfe_orb_get_scale_factors.argtypes = [vp, vp, vp, vp, vp]
    L.drfe_orb_features_per_level.argtypes = [vp, C.c_int]
    L.drfe_orb_max_keypoints.argtypes = [vp]
    L.drfe_orb_extract.argtypes = [vp, vp, C.c_int, C.c_int, sz, vp, vp, C.c_int, i32p]
    L.drfe_orb_enqueue.argtypes = [vp, C.c_int, vp, sz, sz, C.c_int]
    L.drfe_orb_download.argtypes = [vp, vp, vp, C.c_int, vp]
    L.drfe_orb_sync.argtypes = [vp]
    L.drfe_orb_extract_batch.argtypes = [vp, C.c_int, vp, sz, sz, vp, vp, C.c_int, vp, C.c_int]
    L.drfe_orb_finish_batch.argtypes = [vp]
    L.drfe_frame_image_bounds.argtypes = [vp, C.c_int, C.c_int]
    L.drfe_orb_frame_post.argtypes = [vp, vp, vp, sz, sz, C.c_int, vp, vp, vp, vp, vp, C.c_int]
    L.drfe_orb_frame_post_shared_depth.argtypes = [vp, vp, vp, vp, vp, vp, vp, vp, C.c_int]
    L.drfe_orb_search_by_projection.argtypes = [vp, vp, vp, vp, vp, C.c_int, vp]
    L.drfe_vocab_create.argtypes = [C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, vp, vp, vp, vp, C.c_int, C.POINTER(vp)]
    L.drfe_vocab_destroy.argtypes = [vp]
    L.drfe_vocab_destroy.restype = None
    L.drfe_vocab_words.argtypes = [vp]
    L.drfe_orb_compute_bow.argtypes = [vp, vp, C.c_int] + [vp] * 9
    L.drfe_orb_search_by_bow.argtypes = [vp, C.c_int] + [vp] * 12 + [C.c_float, C.c_int, vp, vp, vp]
    L.drfe_orb_enqueue_color.argtypes = [vp, C.c_int, vp, C.c_int, C.c_int, C.c_int, sz, sz, C.c_int]
    L.drfe_orb_get_gray.argtypes = [vp, C.c_int, vp]
    L.drfe_orb_search_local_points.argtypes = [vp, vp, vp, vp, vp, vp, C.c_int, C.c_float, vp, vp, vp, vp]
    L.drfe_orb_search_last_frame.argtypes = [vp, vp, vp, vp, vp, vp, C.c_int, vp, vp, vp, vp, vp]
    L.drfe_orb_stream.argtypes = [vp]
    L.drfe_orb_stream.restype = vp
    L.drfe_orb_level_size.argtypes = [vp, C.c_int, i32p, i32p]
    L.drfe_orb_get_pyramid.argtypes = [vp, C.c_int, C.c_int, C.c_int, vp]
    L.drfe_orb_get_blurred.argtypes = [vp, C.c_int, C.c_int, vp]
    L.drfe_orb_get_candidates.argtypes = [vp, C.c_int, C.c_int, vp, C.c_int, i32p]
    L.drfe_orb_get_level_keypoints.argtypes = [vp, C.c_int, C.c_int, vp, C.c_int, i32p]
    L.drfe_orb_set_profiling.argtypes = [vp, C.c_int]
    L.drfe_orb_stage_times.argtypes = [vp, vp, vp, C.c_int, i32p]
    L.drfe_cape_create.argtypes = [C.POINTER(CapeParams), C.c_int, C.c_int, C.POINTER(vp)]
    L.drfe_cape_destroy.argtypes = [vp]
    L.drfe_cape_enqueue_cloud.argtypes = [vp, C.c_int, vp, sz, C.c_int]
    L.drfe_cape_enqueue_depth.argtypes = [vp, C.c_int, vp, sz, sz, C.c_int, C.c_float, C.c_float, C.c_float, C.c_float]
    L.drfe_cape_download.argtypes = [vp, vp, vp, C.c_int, vp, vp, C.c_int, vp]
    L.drfe_cape_sync.argtypes = [vp]
    L.drfe_cape_enqueue_depth_u16.argtypes = [vp, C.c_int, vp, sz, sz, C.c_int, C.c_float, C.c_float, C.c_float, C.c_float,
                                              C.c_float]
    L.drfe_cape_process_depth_batch.argtypes = [vp, C.c_int, vp, C.c_int, C.c_float, sz, sz, C.c_float, C.c_float, C.c_float,
                                                C.c_float, vp, vp, C.c_int, vp, vp, C.c_int, vp, C.c_int]
    L.drfe_cape_finish_batch.argtypes = [vp]
    L.drfe_cape_stream.argtypes = [vp]
    L.drfe_cape_stream.restype = vp
    L.drfe_cape_process.argtypes = [vp, vp, vp, vp, C.c_int, i32p, vp, C.c_int, i32p]
    L.drfe_cape_process_depth.argtypes = [vp, vp, sz, C.c_float, C.c_float, C.c_float, C.c_float, vp, vp, C.c_int,
                                          i32p, vp, C.c_int, i32p]
    L.drfe_cape_num_cells.argtypes = [vp, i32p, i32p]
    L.drfe_cape_get_cloud.argtypes = [vp, C.c_int, vp]
    L.drfe_cape_get_cells.argtypes = [vp, C.c_int, vp]
    L.drfe_cape_get_grid_maps.argtypes = [vp, C.c_int, vp, vp]
    L.drfe_cape_cylinders_found.argtypes = [vp, vp]
    L.drfe_cape_plane_points.argtypes = [vp, vp, sz, vp, C.c_int]
    L.drfe_cape_plane_points_voxel.argtypes = [vp, C.c_float, vp, sz, vp, C.c_int]
    L.drfe_cape_third_cloud.argtypes = [vp, C.c_float, vp]
    L.drfe_cape_third_cloud_normals.argtypes = [vp, C.c_float, C.c_float, C.c_float, vp, vp]
    L.drfe_cape_get_cyl_maps.argtypes = [vp, C.c_int, vp, vp]
    L.drfe_cape_debug_counters.argtypes = [vp, C.c_int, vp]
    L.drfe_cape_set_profiling.argtypes = [vp, C.c_int]
    L.drfe_cape_stage_times.argtypes = [vp, vp, vp, C.c_int, i32p]
    L.drfe_peac_default_params.argtypes = [C.POINTER(PeacParams)]
    L.drfe_peac_create.argtypes = [C.c_int, C.c_int, C.POINTER(PeacParams), C.c_int, C.c_int, C.POINTER(vp)]
    L.drfe_peac_destroy.argtypes = [vp]
    L.drfe_peac_stream.argtypes = [vp]
    L.drfe_peac_stream.restype = vp
    L.drfe_peac_sync.argtypes = [vp]
    L.drfe_peac_enqueue_depth_u16.argtypes = [vp, C.c_int, vp, sz, sz, C.c_int, C.c_float, C.c_float, C.c_float, C.c_float, C.c_float]
    L.drfe_peac_download.argtypes = [vp, vp, vp, C.c_int, vp]
    L.drfe_peac_plane_vertices.argtypes = [vp, vp, vp, sz, vp, C.c_int]
    L.drfe_peac_plane_points_voxel.argtypes = [vp, C.c_float, C.c_float, vp, sz, vp, C.c_int]
    L.drfe_peac_third_cloud_normals.argtypes = [vp, C.c_float, C.c_float, C.c_float, vp, vp]
    L.drfe_peac_debug_counters.argtypes = [vp, C.c_int, vp]
    L.drfe_resizer_create.argtypes = [C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.POINTER(vp)]
    L.drfe_resizer_destroy.argtypes = [vp]
    L.drfe_resizer_stream.argtypes = [vp]
    L.drfe_resizer_stream.restype = vp
    L.drfe_resizer_sync.argtypes = [vp]
    L.drfe_resize.argtypes = [vp, C.c_int, vp, C.c_int, C.c_int, sz, sz, C.c_int, vp, sz, sz, C.c_int]
    L.drfe_pool_create.argtypes = [C.POINTER(PoolParams), vp, C.c_int, C.POINTER(vp)]
    L.drfe_pool_destroy.argtypes = [vp]
    L.drfe_pool_num_devices.argtypes = [vp]
    L.drfe_pool_max_keypoints.argtypes = [vp]
    L.drfe_pool_extract_batch.argtypes = [vp, C.c_int, vp, sz, sz, vp, C.c_int, C.c_float, sz, sz, C.c_float, C.c_float, C.c_float, C.c_float,
                                          vp, vp, C.c_int, vp, vp, vp, C.c_int, vp, vp, C.c_int, vp]
    L.drfe_pool_device_times.argtypes = [vp, vp, C.c_int]
    L.drfe_host_alloc.argtypes = [C.POINTER(vp), sz, C.c_int]
    L.drfe_host_free.argtypes = [vp]
    L.drfe_host_register.argtypes = [vp, sz]
    L.drfe_host_unregister.argtypes = [vp]
    _lib = L
    return L


def _check(rc):
    if rc != OK:
        raise DrfeError(rc, lib().drfe_last_error().decode())


def _ptr(a):
    if a is None:
        return None
    if isinstance(a, int):
        return C.c_void_p(a)
    return a.ctypes.data_as(C.c_void_p)


def device_count():
    n = C.c_int32(0)
    lib().drfe_device_count(C.byref(n))
    return n.value


def kernel_launch_count():
    return int(lib().drfe_kernel_launch_count())


def synth_frame(width=640, height=480, scene=0, seed=20260000, depth_unit_scale=1.0):
    """The tests' synthetic RGB-D frame.  The generator is not part of the product library: it lives in
    tools/synth (libdrfe_synth.so); this wrapper only keeps the tests' call sites short."""
    import importlib.util
    spec = importlib.util.spec_from_file_location(
        "drfe_synth", os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tools", "synth", "synth.py"))
    global _synth_mod
    if _synth_mod is None:
        _synth_mod = importlib.util.module_from_spec(spec)
        spec.loader.exec_module(_synth_mod)
    return _synth_mod.synth_frame(width, height, scene, seed, depth_unit_scale)


_synth_mod = None


class Event:
    """CUDA event on a handle stream (device-side timing)."""

    def __init__(self):
        self.e = C.c_void_p()
        _check(lib().drfe_event_create(C.byref(self.e)))

    def record(self, stream):
        _check(lib().drfe_event_record(self.e, stream))

    def elapsed_ms(self, stop):
        ms = C.c_float(0)
        _check(lib().drfe_event_elapsed_ms(self.e, stop.e, C.byref(ms)))
        return ms.value

    def __del__(self):
        if getattr(self, "e", None):
            lib().drfe_event_destroy(self.e)
            self.e = None


def stream_wait_event(stream, event):
    _check(lib().drfe_stream_wait_event(stream, event.e))


def _stage_times(fn, h):
    ms = (C.c_float * 16)()
    names = (C.c_char_p * 16)()
    n = C.c_int32(0)
    _check(fn(h, C.cast(ms, C.c_void_p), C.cast(names, C.c_void_p), 16, C.byref(n)))
    return [(names[i].decode(), float(ms[i])) for i in range(n.value)]


def host_array(shape, dtype, write_combined=False):
    """numpy array over pinned host memory from drfe_host_alloc (kept alive by the array's base object)"""
    dtype = np.dtype(dtype)
    n = int(np.prod(shape)) * dtype.itemsize
    p = C.c_void_p()
    _check(lib().drfe_host_alloc(C.byref(p), max(n, 1), int(write_combined)))

    class _Owner:
        def __init__(self, ptr):
            self.ptr = ptr

        def __del__(self):
            try:
                lib().drfe_host_free(self.ptr)
            except Exception:
                pass
    buf = (C.c_uint8 * max(n, 1)).from_address(p.value)
    buf._owner = _Owner(p)
    return np.frombuffer(buf, dtype=np.uint8, count=n).view(dtype).reshape(shape)


class PEAC:
    """Planar_SLAM::PlaneDetection on ahc::PlaneFitter (PlaneExtractor.h:61-81): readDepthImage + runPlaneDetection, batched"""

    def __init__(self, width=640, height=480, max_batch=1, device=0, **params):
        self.L = lib()
        self.h = C.c_void_p()
        self.W, self.H, self.max_batch = width, height, max_batch
        self.prm = PeacParams()
        _check(self.L.drfe_peac_default_params(C.byref(self.prm)))
        for k, v in params.items():
            setattr(self.prm, k, v)
        _check(self.L.drfe_peac_create(width, height, C.byref(self.prm), max_batch, device, C.byref(self.h)))

    def close(self):
        if getattr(self, "h", None):
            self.L.drfe_peac_destroy(self.h)
            self.h = None

    __del__ = close

    def params_array(self):
        return np.array([getattr(self.prm, n) for n, _ in PeacParams._fields_[:11]], np.float64)

    def enqueue(self, depth16, depth_factor, fx, fy, cx, cy, nframes=None, mem_kind=MEM_HOST):
        """depth16: [n, H, W] uint16 array on the host, or (with mem_kind=MEM_DEVICE and nframes) a device pointer to that layout"""
        if mem_kind == MEM_HOST:
            depth16 = np.ascontiguousarray(depth16, np.uint16)
            assert depth16.ndim == 3 and depth16.shape[1:] == (self.H, self.W)
            self._keep = depth16
            self._nframes = depth16.shape[0]
            ptr = _ptr(depth16)
        else:
            self._nframes, ptr = int(nframes), C.c_void_p(int(depth16))
        _check(self.L.drfe_peac_enqueue_depth_u16(self.h, self._nframes, ptr, self.W, self.W * self.H, mem_kind, depth_factor, fx, fy, cx, cy))

    def sync(self):
        _check(self.L.drfe_peac_sync(self.h))

    def download(self, plane_cap=255):
        nf = self._nframes
        seg = np.empty((nf, self.H, self.W), np.uint8)
        planes = np.zeros((nf, plane_cap), PEAC_PLANE_DTYPE)
        npl = np.empty(nf, np.int32)
        _check(self.L.drfe_peac_download(self.h, _ptr(seg), _ptr(planes), plane_cap, _ptr(npl)))
        return seg, planes, npl

    def plane_vertices(self, plane_cap=255, cap_per_frame=None):
        nf = self._nframes
        N = cap_per_frame or self.H * self.W
        idx = np.zeros((nf, N), np.int32)
        pts = np.zeros((nf, N, 3), np.float32)
        offs = np.zeros((nf, plane_cap + 1), np.int32)
        _check(self.L.drfe_peac_plane_vertices(self.h, _ptr(idx), _ptr(pts), N, _ptr(offs), plane_cap))
        return idx, pts, offs

    def plane_points_voxel(self, max_point_dist=float(np.finfo(np.float32).max), leaf=0.05, plane_cap=255, cap_per_frame=None):
        """Frame::ComputePlanes' coarseCloud of every plane: z <= max_point_dist, pcl::VoxelGrid(leaf) -> (points, offsets)"""
        nf = self._nframes
        N = cap_per_frame or self.H * self.W
        pts = np.zeros((nf, N, 3), np.float32)
        offs = np.zeros((nf, plane_cap + 1), np.int32)
        _check(self.L.drfe_peac_plane_points_voxel(self.h, max_point_dist, leaf, _ptr(pts), N, _ptr(offs), plane_cap))
        return pts, offs

    def third_cloud_normals(self, max_point_dist, max_depth_change_factor=0.05, smoothing_size=10.0):
        """Frame::ComputePlanes' 1/3-resolution cloud and PCL's integral-image normals on it (Frame.cc:1044-1100) -> (cloud, normals)"""
        nf = self._nframes
        cloud = np.empty((nf, (self.H + 2) // 3, (self.W + 2) // 3, 3), np.float32)
        normals = np.empty_like(cloud)
        _check(self.L.drfe_peac_third_cloud_normals(self.h, max_point_dist, max_depth_change_factor, smoothing_size, _ptr(cloud), _ptr(normals)))
        return cloud, normals

    def counters(self, frame=0):
        out = np.zeros(12, np.int32)
        _check(self.L.drfe_peac_debug_counters(self.h, frame, _ptr(out)))
        return out


class Resizer:
    """cv::resize(src, dst, Size(w, h)) (INTER_LINEAR) of System::TrackRGBD (System.cc:325-329), batched"""
    PIX = {np.dtype(np.uint8): 0, np.dtype(np.uint16): 2, np.dtype(np.float32): 5}

    def __init__(self, src_w, src_h, dst_w, dst_h, max_batch=1, device=0):
        self.L = lib()
        self.h = C.c_void_p()
        self.src, self.dst = (src_w, src_h), (dst_w, dst_h)
        _check(self.L.drfe_resizer_create(src_w, src_h, dst_w, dst_h, max_batch, device, C.byref(self.h)))

    def close(self):
        if getattr(self, "h", None):
            self.L.drfe_resizer_destroy(self.h)
            self.h = None

    __del__ = close

    def __call__(self, images):
        """images: (B, H, W) or (B, H, W, C) host array of uint8 / uint16 / float32 -> resized array of the same kind"""
        a = np.ascontiguousarray(images)
        ch = a.shape[3] if a.ndim == 4 else 1
        out = np.empty((a.shape[0], self.dst[1], self.dst[0]) + ((ch,) if a.ndim == 4 else ()), a.dtype)
        es = a.itemsize
        _check(self.L.drfe_resize(self.h, a.shape[0], _ptr(a), self.PIX[a.dtype], ch, a.shape[2] * ch * es, a.shape[1] * a.shape[2] * ch * es, MEM_HOST,
                                  _ptr(out), self.dst[0] * ch * es, self.dst[0] * self.dst[1] * ch * es, MEM_HOST))
        return out


class Pool:
    """drfe_pool: one batch of host frames sharded over several devices (contiguous blocks), results by frame index."""

    def __init__(self, devices, nfeatures=1000, scaleFactor=1.2, nlevels=8, iniThFAST=20, minThFAST=7, width=640, height=480,
                 cell=20, cylinder_detection=False, min_cos=0.97814, max_merge_dist=900.0, max_batch=256, chunk_frames=0):
        self.L = lib()
        self.h = C.c_void_p()
        self.width, self.height, self.max_batch = width, height, max_batch
        prm = PoolParams(OrbParams(nfeatures, scaleFactor, nlevels, iniThFAST, minThFAST),
                         CapeParams(height, width, cell, cell, int(cylinder_detection), min_cos, max_merge_dist), width, height, max_batch, chunk_frames)
        dev = np.ascontiguousarray(devices, np.int32)
        _check(self.L.drfe_pool_create(C.byref(prm), _ptr(dev), len(dev), C.byref(self.h)))
        self.cap = self.L.drfe_pool_max_keypoints(self.h)
        self.ndev = self.L.drfe_pool_num_devices(self.h)

    def close(self):
        if getattr(self, "h", None):
            self.L.drfe_pool_destroy(self.h)
            self.h = None

    __del__ = close

    def extract_batch(self, gray, depth, fx, fy, cx, cy, depth_factor=1.0, out=None, plane_cap=64):
        """gray (B,H,W) u8, depth (B,H,W) f32 or u16 (host) -> dict of per-frame results (out: preallocated dict to reuse)"""
        assert gray.dtype == np.uint8 and gray.ndim == 3 and gray.strides[2] == 1
        assert depth.dtype in (np.float32, np.uint16) and depth.shape == gray.shape and depth.strides[2] == depth.itemsize
        nf = gray.shape[0]
        if out is None:
            out = {"kps": np.empty((nf, self.cap), KP_DTYPE), "desc": np.empty((nf, self.cap, 32), np.uint8), "counts": np.empty(nf, np.int32),
                   "seg": np.empty((nf, self.height, self.width), np.uint8), "planes": np.empty((nf, plane_cap), PLANE_DTYPE),
                   "nplanes": np.empty(nf, np.int32)}
        es = depth.itemsize
        _check(self.L.drfe_pool_extract_batch(self.h, nf, _ptr(gray), gray.strides[1], gray.strides[0], _ptr(depth), int(depth.dtype == np.uint16),
                                              depth_factor, depth.strides[1] // es, depth.strides[0] // es, fx, fy, cx, cy,
                                              _ptr(out["kps"]), _ptr(out["desc"]), out["kps"].shape[1], _ptr(out["counts"]), _ptr(out["seg"]),
                                              _ptr(out["planes"]), out["planes"].shape[1], _ptr(out["nplanes"]), None, 0, None))
        return out

    def device_times(self):
        ms = np.zeros(self.ndev, np.float32)
        _check(self.L.drfe_pool_device_times(self.h, _ptr(ms), self.ndev))
        return ms


class Vocabulary:
    """ORBVocabulary on the device, from the arrays TemplatedVocabulary::loadFromTextFile parses"""

    def __init__(self, k, L, scoring, weighting, parent, is_leaf, descriptors, weights, device=0):
        self.L_ = lib()
        parent = np.ascontiguousarray(parent, np.int32)
        is_leaf = np.ascontiguousarray(is_leaf, np.uint8)
        descriptors = np.ascontiguousarray(descriptors, np.uint8)
        weights = np.ascontiguousarray(weights, np.float64)
        assert descriptors.shape == (len(parent), 32) and len(is_leaf) == len(parent) == len(weights)
        self.h = C.c_void_p()
        _check(self.L_.drfe_vocab_create(k, L, scoring, weighting, len(parent), _ptr(parent), _ptr(is_leaf), _ptr(descriptors),
                                         _ptr(weights), device, C.byref(self.h)))

    def words(self):
        return self.L_.drfe_vocab_words(self.h)

    def close(self):
        if self.h:
            self.L_.drfe_vocab_destroy(self.h)
            self.h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class ORBextractor:
    """Planar_SLAM::ORBextractor with its constructor arguments (ORBextractor.h:51-52).  The image
    size and the maximum frame batch are fixed per instance (device buffers are preallocated)."""

    def __init__(self, nfeatures=1000, scaleFactor=1.2, nlevels=8, iniThFAST=20, minThFAST=7,
                 width=640, height=480, max_batch=1, device=0):
        self.L = lib()
        self.h = C.c_void_p()
        self.width, self.height, self.max_batch, self.nlevels = width, height, max_batch, nlevels
        prm = OrbParams(nfeatures, scaleFactor, nlevels, iniThFAST, minThFAST)
        _check(self.L.drfe_orb_create(C.byref(prm), width, height, max_batch, device, C.byref(self.h)))
        self.cap = self.L.drfe_orb_max_keypoints(self.h)

    def close(self):
        if getattr(self, "h", None):
            self.L.drfe_orb_destroy(self.h)
            self.h = None

    __del__ = close

    # ---- getters (ORBextractor.h:63-83)
    def GetLevels(self):
        return self.L.drfe_orb_get_levels(self.h)

    def GetScaleFactor(self):
        return self.L.drfe_orb_get_scale_factor(self.h)

    def _factors(self):
        a = [np.empty(self.nlevels, np.float32) for _ in range(4)]
        _check(self.L.drfe_orb_get_scale_factors(self.h, *[_ptr(x) for x in a]))
        return a

    def GetScaleFactors(self):
        return self._factors()[0]

    def GetInverseScaleFactors(self):
        return self._factors()[1]

    def GetScaleSigmaSquares(self):
        return self._factors()[2]

    def GetInverseScaleSigmaSquares(self):
        return self._factors()[3]

    def features_per_level(self):
        return [self.L.drfe_orb_features_per_level(self.h, l) for l in range(self.nlevels)]

    # ---- operator()
    def __call__(self, image, mask=None):
        """operator()(image, mask, keypoints, descriptors): mask is ignored (ORBextractor.h:58).
        An empty image returns (None, None) — the reference returns without touching its outputs."""
        if image is None or image.size == 0:
            return None, None
        if image.dtype != np.uint8 or image.ndim != 2:
            raise AssertionError("image.type() == CV_8UC1")   # reference asserts (ORBextractor.cc:1050)
        if image.strides[1] != 1:
            image = np.ascontiguousarray(image)
        kps = np.empty(self.cap, KP_DTYPE)
        desc = np.empty((self.cap, 32), np.uint8)
        n = C.c_int32(0)
        _check(self.L.drfe_orb_extract(self.h, _ptr(image), image.shape[1], image.shape[0], image.strides[0],
                                       _ptr(kps), _ptr(desc), self.cap, C.byref(n)))
        self._nframes = 1
        return kps[:n.value].copy(), desc[:n.value].copy()

    # ---- batched
    def enqueue(self, gray, mem_kind=MEM_HOST, nframes=None, row_stride=None, frame_stride=None):
        """gray: (B,H,W) uint8 numpy array (host) or a raw device pointer (int) with explicit strides."""
        if isinstance(gray, np.ndarray):
            assert gray.dtype == np.uint8 and gray.ndim == 3 and gray.strides[2] == 1
            nframes, row_stride, frame_stride = gray.shape[0], gray.strides[1], gray.strides[0]
            self._keep = gray
        _check(self.L.drfe_orb_enqueue(self.h, nframes, _ptr(gray), row_stride, frame_stride, mem_kind))
        self._nframes = nframes

    def download(self, kps=None, desc=None, counts=None):
        nf = self._nframes
        kps = np.empty((nf, self.cap), KP_DTYPE) if kps is None else kps
        desc = np.empty((nf, self.cap, 32), np.uint8) if desc is None else desc
        counts = np.empty(nf, np.int32) if counts is None else counts
        assert kps.shape[1] == desc.shape[1]
        _check(self.L.drfe_orb_download(self.h, _ptr(kps), _ptr(desc), kps.shape[1], _ptr(counts)))
        return kps, desc, counts

    def extract_batch(self, gray, kps=None, desc=None, counts=None, chunk_frames=0):
        """Chunk-pipelined host-in / host-out batch call (asynchronous; finish_batch() waits).
        gray: (B,H,W) uint8 host array (pinned for copy/compute overlap)."""
        assert gray.dtype == np.uint8 and gray.ndim == 3 and gray.strides[2] == 1
        nf = gray.shape[0]
        kps = np.empty((nf, self.cap), KP_DTYPE) if kps is None else kps
        desc = np.empty((nf, self.cap, 32), np.uint8) if desc is None else desc
        counts = np.empty(nf, np.int32) if counts is None else counts
        self._keep = (gray, kps, desc, counts)
        _check(self.L.drfe_orb_extract_batch(self.h, nf, _ptr(gray), gray.strides[1], gray.strides[0], _ptr(kps), _ptr(desc),
                                             kps.shape[1], _ptr(counts), chunk_frames))
        self._nframes = nf
        return kps, desc, counts

    def finish_batch(self):
        _check(self.L.drfe_orb_finish_batch(self.h))

    def frame_params(self, fx, fy, cx, cy, dist, bf):
        """mK / mDistCoef / mbf + Frame::ComputeImageBounds for this handle's image size"""
        p = FrameParams(fx, fy, cx, cy, (C.c_float * 5)(*dist), bf, 0, 0, 0, 0)
        _check(self.L.drfe_frame_image_bounds(C.byref(p), self.width, self.height))
        return p

    def frame_post(self, p, depth, mem_kind=MEM_HOST, row_stride=None, frame_stride=None):
        """UndistortKeyPoints + ComputeStereoFromRGBD + AssignFeaturesToGrid on the last batch's keypoints:
        (mvKeysUn, mvuRight, mvDepth, grid counts [nf, 64, 48], grid index lists [nf, cap])"""
        nf = self._nframes
        if isinstance(depth, np.ndarray):
            assert depth.dtype == np.float32 and depth.ndim == 3 and depth.strides[2] == 4 and depth.shape[0] == nf
            row_stride, frame_stride = depth.strides[1] // 4, depth.strides[0] // 4
        ku = np.empty((nf, self.cap), KP_DTYPE)
        ur, kd = np.empty((nf, self.cap), np.float32), np.empty((nf, self.cap), np.float32)
        gc, gi = np.empty((nf, 64, 48), np.uint16), np.empty((nf, self.cap), np.uint16)
        _check(self.L.drfe_orb_frame_post(self.h, C.byref(p), _ptr(depth), row_stride, frame_stride, mem_kind, _ptr(ku), _ptr(ur),
                                          _ptr(kd), _ptr(gc), _ptr(gi), self.cap))
        return ku, ur, kd, gc, gi

    def frame_post_shared_depth(self, p, cape):
        """frame_post on the depth images the CAPE handle already holds on the device (one H2D copy of imDepth serves both
        extractors, as in Frame::Frame)"""
        nf = self._nframes
        ku = np.empty((nf, self.cap), KP_DTYPE)
        ur, kd = np.empty((nf, self.cap), np.float32), np.empty((nf, self.cap), np.float32)
        gc, gi = np.empty((nf, 64, 48), np.uint16), np.empty((nf, self.cap), np.uint16)
        _check(self.L.drfe_orb_frame_post_shared_depth(self.h, cape.h, C.byref(p), _ptr(ku), _ptr(ur), _ptr(kd), _ptr(gc), _ptr(gi), self.cap))
        return ku, ur, kd, gc, gi

    def search_by_projection(self, queries, qdesc, nqueries=None, occupied=None):
        """Core of ORBmatcher::SearchByProjection(Frame&, vector<MapPoint*>&, th) (ORBmatcher.cc:69-116) on the frames of
        the last frame_post: queries (nf, qcap) QUERY_DTYPE, qdesc (nf, qcap, 32) uint8, occupied (nf, cap) uint8 or None
        -> (nf, qcap) MATCH_DTYPE"""
        nf = self._nframes
        queries = np.ascontiguousarray(queries, QUERY_DTYPE)
        qdesc = np.ascontiguousarray(qdesc, np.uint8)
        assert queries.shape[0] == nf and qdesc.shape == queries.shape + (32,)
        qcap = queries.shape[1]
        nq = np.full(nf, qcap, np.int32) if nqueries is None else np.ascontiguousarray(nqueries, np.int32)
        if occupied is not None:
            occupied = np.ascontiguousarray(occupied, np.uint8)
            assert occupied.shape == (nf, self.cap)
        out = np.zeros((nf, qcap), MATCH_DTYPE)
        _check(self.L.drfe_orb_search_by_projection(self.h, _ptr(nq), _ptr(queries), _ptr(qdesc), _ptr(occupied), qcap, _ptr(out)))
        return out

    def search_local_points(self, queries, qdesc, qflags, nnratio=0.8, nqueries=None, occupied=None):
        """ORBmatcher::SearchByProjection(F, vpMapPoints, th) whole (ORBmatcher.cc:46-130): -> (match records (nf, qcap),
        assigned (nf, qcap), key_point (nf, cap), nmatches (nf,))"""
        nf = self._nframes
        queries = np.ascontiguousarray(queries, QUERY_DTYPE)
        qdesc = np.ascontiguousarray(qdesc, np.uint8)
        qflags = np.ascontiguousarray(qflags, np.uint8)
        assert queries.shape[0] == nf and qdesc.shape == queries.shape + (32,) and qflags.shape == queries.shape
        qcap = queries.shape[1]
        nq = np.full(nf, qcap, np.int32) if nqueries is None else np.ascontiguousarray(nqueries, np.int32)
        if occupied is not None:
            occupied = np.ascontiguousarray(occupied, np.uint8)
            assert occupied.shape == (nf, self.cap)
        out, asg = np.zeros((nf, qcap), MATCH_DTYPE), np.zeros((nf, qcap), np.int32)
        kp, nm = np.zeros((nf, self.cap), np.int32), np.zeros(nf, np.int32)
        _check(self.L.drfe_orb_search_local_points(self.h, _ptr(nq), _ptr(queries), _ptr(qdesc), _ptr(qflags), _ptr(occupied), qcap, nnratio,
                                                   _ptr(out), _ptr(asg), _ptr(kp), _ptr(nm)))
        return out, asg, kp, nm

    def search_last_frame(self, tp, points, pdesc, npoints=None, occupied=None):
        """ORBmatcher::SearchByProjection(CurrentFrame, LastFrame, th, bMono) (ORBmatcher.cc:1396-1535) with the frames of
        the last frame_post as the current frames: tp (nf,) TRACK_PARAMS_DTYPE, points (nf, pcap) LAST_POINT_DTYPE, pdesc
        (nf, pcap, 32) uint8, occupied (nf, cap) uint8 or None -> (match_key, match_dist, key_point, nmatches, sweeps)"""
        nf = self._nframes
        tp = np.ascontiguousarray(tp, TRACK_PARAMS_DTYPE)
        points = np.ascontiguousarray(points, LAST_POINT_DTYPE)
        pdesc = np.ascontiguousarray(pdesc, np.uint8)
        assert tp.shape == (nf,) and points.shape[0] == nf and pdesc.shape == points.shape + (32,)
        pcap = points.shape[1]
        npts = np.full(nf, pcap, np.int32) if npoints is None else np.ascontiguousarray(npoints, np.int32)
        if occupied is not None:
            occupied = np.ascontiguousarray(occupied, np.uint8)
            assert occupied.shape == (nf, self.cap)
        mk, md = np.zeros((nf, pcap), np.int32), np.zeros((nf, pcap), np.int32)
        kp = np.zeros((nf, self.cap), np.int32)
        nm, sw = np.zeros(nf, np.int32), np.zeros(nf, np.int32)
        _check(self.L.drfe_orb_search_last_frame(self.h, _ptr(tp), _ptr(npts), _ptr(points), _ptr(pdesc), _ptr(occupied), pcap,
                                                 _ptr(mk), _ptr(md), _ptr(kp), _ptr(nm), _ptr(sw)))
        return mk, md, kp, nm, sw

    def enqueue_color(self, pixels, rgb_order=True, coeffs=0, mem_kind=MEM_HOST):
        """cvtColor(..., CV_RGB2GRAY / BGR / RGBA / BGRA) of Tracking::GrabImageRGBD (Tracking.cc:194-207) + enqueue:
        pixels (B,H,W,3|4) uint8; coeffs 0 = Q15 (OpenCV 4.x), 1 = Q14 (OpenCV <= 3.4)"""
        assert pixels.dtype == np.uint8 and pixels.ndim == 4 and pixels.shape[3] in (3, 4) and pixels.strides[3] == 1 and pixels.strides[2] == pixels.shape[3]
        nf = pixels.shape[0]
        _check(self.L.drfe_orb_enqueue_color(self.h, nf, _ptr(pixels), pixels.shape[3], int(rgb_order), coeffs, pixels.strides[1], pixels.strides[0], mem_kind))
        self._nframes = nf

    def get_gray(self, frame=0):
        g = np.empty((self.height, self.width), np.uint8)
        _check(self.L.drfe_orb_get_gray(self.h, frame, _ptr(g)))
        return g

    def compute_bow(self, vocab, levelsup=4):
        """Frame::ComputeBoW (Frame.cc:828-833) on the last batch's descriptors -> per frame (word_id, node_id,
        [(word, value)], [(node, [indices])])"""
        nf, cap = self._nframes, self.cap
        wid, nid = np.zeros((nf, cap), np.int32), np.zeros((nf, cap), np.int32)
        bn, bw, bv = np.zeros(nf, np.int32), np.zeros((nf, cap), np.int32), np.zeros((nf, cap), np.float64)
        fn, fnode, fstart, ffeat = np.zeros(nf, np.int32), np.zeros((nf, cap), np.int32), np.zeros((nf, cap + 1), np.int32), np.zeros((nf, cap), np.int32)
        _check(self.L.drfe_orb_compute_bow(self.h, vocab.h, levelsup, _ptr(wid), _ptr(nid), _ptr(bn), _ptr(bw), _ptr(bv), _ptr(fn),
                                           _ptr(fnode), _ptr(fstart), _ptr(ffeat)))
        out = []
        for f in range(nf):
            bow = [(int(bw[f, j]), float(bv[f, j])) for j in range(bn[f])]
            fv = [(int(fnode[f, j]), ffeat[f, fstart[f, j]:fstart[f, j + 1]].tolist()) for j in range(fn[f])]
            out.append((wid[f], nid[f], bow, fv))
        return out

    @staticmethod
    def pack_feature_vectors(fvs, cap):
        """[(node, [indices])] per frame -> the flat layout of drfe_orb_compute_bow / drfe_orb_search_by_bow"""
        nf = len(fvs)
        n, node, start, feat = np.zeros(nf, np.int32), np.zeros((nf, cap), np.int32), np.zeros((nf, cap + 1), np.int32), np.zeros((nf, cap), np.int32)
        for f, fv in enumerate(fvs):
            n[f] = len(fv)
            o = 0
            for j, (nid, idx) in enumerate(fv):
                node[f, j], start[f, j] = nid, o
                feat[f, o:o + len(idx)] = idx
                o += len(idx)
            start[f, len(fv)] = o
        return n, node, start, feat

    def search_by_bow(self, kf_n, kf_desc, kf_angle, kf_valid, kf_fvs, f_fvs, nnratio=0.7, check_orientation=True):
        """ORBmatcher::SearchByBoW(pKF, F, matches) (ORBmatcher.cc:160-292), frame f of the last batch against keyframe f:
        kf_desc (nf, kcap, 32), kf_angle / kf_valid (nf, kcap), FeatureVectors as [(node, [indices])] per frame
        -> (kf_match (nf, kcap), f_match (nf, cap), nmatches (nf,))"""
        nf = self._nframes
        kf_desc = np.ascontiguousarray(kf_desc, np.uint8)
        kcap = kf_desc.shape[1]
        kf_angle = np.ascontiguousarray(kf_angle, np.float32)
        kf_valid = np.ascontiguousarray(kf_valid, np.uint8)
        kf_n = np.ascontiguousarray(kf_n, np.int32)
        assert kf_desc.shape == (nf, kcap, 32) and kf_angle.shape == (nf, kcap) == kf_valid.shape
        kn, knode, kstart, kfeat = self.pack_feature_vectors(kf_fvs, kcap)
        fn, fnode, fstart, ffeat = self.pack_feature_vectors(f_fvs, self.cap)
        km, fm, nm = np.zeros((nf, kcap), np.int32), np.zeros((nf, self.cap), np.int32), np.zeros(nf, np.int32)
        _check(self.L.drfe_orb_search_by_bow(self.h, kcap, _ptr(kf_n), _ptr(kf_desc), _ptr(kf_angle), _ptr(kf_valid), _ptr(kn), _ptr(knode),
                                             _ptr(kstart), _ptr(kfeat), _ptr(fn), _ptr(fnode), _ptr(fstart), _ptr(ffeat), nnratio,
                                             int(check_orientation), _ptr(km), _ptr(fm), _ptr(nm)))
        return km, fm, nm

    def sync(self):
        _check(self.L.drfe_orb_sync(self.h))

    def stream(self):
        return self.L.drfe_orb_stream(self.h)

    def set_profiling(self, on=True):
        _check(self.L.drfe_orb_set_profiling(self.h, int(on)))

    def stage_times(self):
        return _stage_times(self.L.drfe_orb_stage_times, self.h)

    # ---- intermediates (mvImagePyramid etc.)
    def level_size(self, level):
        w, h = C.c_int32(), C.c_int32()
        _check(self.L.drfe_orb_level_size(self.h, level, C.byref(w), C.byref(h)))
        return w.value, h.value

    def pyramid(self, frame, level, bordered=False):
        w, h = self.level_size(level)
        if bordered:
            w, h = w + 38, h + 38
        out = np.empty((h, w), np.uint8)
        _check(self.L.drfe_orb_get_pyramid(self.h, frame, level, int(bordered), _ptr(out)))
        return out

    def blurred(self, frame, level):
        w, h = self.level_size(level)
        out = np.empty((h, w), np.uint8)
        _check(self.L.drfe_orb_get_blurred(self.h, frame, level, _ptr(out)))
        return out

    def candidates(self, frame, level):
        cap = 1 << 19
        buf = np.empty((cap, 3), np.float32)
        n = C.c_int32(0)
        _check(self.L.drfe_orb_get_candidates(self.h, frame, level, _ptr(buf), cap, C.byref(n)))
        return buf[:n.value].copy()

    def level_keypoints(self, frame, level):
        buf = np.empty(self.cap, KP_DTYPE)
        n = C.c_int32(0)
        _check(self.L.drfe_orb_get_level_keypoints(self.h, frame, level, _ptr(buf), self.cap, C.byref(n)))
        return buf[:n.value].copy()


class CAPE:
    """CAPE with its constructor arguments (CAPE.h:47)."""

    def __init__(self, depth_height, depth_width, cell_width, cell_height, cylinder_detection=False,
                 min_cos_angle_4_merge=0.97814, max_merge_dist=900.0, max_batch=1, device=0):
        self.L = lib()
        self.h = C.c_void_p()
        self.H, self.W, self.cw, self.ch, self.max_batch = depth_height, depth_width, cell_width, cell_height, max_batch
        prm = CapeParams(depth_height, depth_width, cell_width, cell_height, int(cylinder_detection),
                         min_cos_angle_4_merge, max_merge_dist)
        _check(self.L.drfe_cape_create(C.byref(prm), max_batch, device, C.byref(self.h)))
        cx, cy = C.c_int32(), C.c_int32()
        self.L.drfe_cape_num_cells(self.h, C.byref(cx), C.byref(cy))
        self.ncx, self.ncy = cx.value, cy.value
        self.plane_cap = 255
        self.cylinder_detection = bool(cylinder_detection)
        self.cyl_cap = (self.ncx * self.ncy) // 6 + 1 if cylinder_detection else 0

    def close(self):
        if getattr(self, "h", None):
            self.L.drfe_cape_destroy(self.h)
            self.h = None

    __del__ = close

    def process(self, cloud_array, seg_output=None):
        """process(cloud_array, nr_planes, nr_cylinders, seg_output, plane_segments_final,
        cylinder_segments_final) -> (nr_planes, nr_cylinders, seg_output, planes, cylinders).
        cloud_array: cell-major N x 3 column-major float32 (flat 3*N: all X, all Y, all Z).
        Like the reference only labelled pixels of a passed-in seg_output are overwritten."""
        cloud = np.ascontiguousarray(cloud_array, np.float32).reshape(-1)
        assert cloud.size == 3 * self.H * self.W
        seg = np.empty((self.H, self.W), np.uint8)
        planes = np.zeros(self.plane_cap, PLANE_DTYPE)
        npl, ncyl = C.c_int32(0), C.c_int32(0)
        cyls = np.zeros(max(self.cyl_cap, 1), CYL_DTYPE)
        self._nframes = 1
        _check(self.L.drfe_cape_process(self.h, _ptr(cloud), _ptr(seg), _ptr(planes), self.plane_cap, C.byref(npl),
                                        _ptr(cyls), self.cyl_cap, C.byref(ncyl)))
        if seg_output is not None:
            seg_output[seg > 0] = seg[seg > 0]
            seg = seg_output
        return npl.value, ncyl.value, seg, planes[:npl.value].copy(), cyls[:self.cylinders_found()[0]].copy()

    def process_depth(self, depth, fx, fy, cx, cy):
        depth = np.ascontiguousarray(depth, np.float32)
        seg = np.empty((self.H, self.W), np.uint8)
        planes = np.zeros(self.plane_cap, PLANE_DTYPE)
        npl, ncyl = C.c_int32(0), C.c_int32(0)
        cyls = np.zeros(max(self.cyl_cap, 1), CYL_DTYPE)
        self._nframes = 1
        _check(self.L.drfe_cape_process_depth(self.h, _ptr(depth), depth.strides[0] // 4, fx, fy, cx, cy, _ptr(seg),
                                              _ptr(planes), self.plane_cap, C.byref(npl), _ptr(cyls), self.cyl_cap,
                                              C.byref(ncyl)))
        return npl.value, ncyl.value, seg, planes[:npl.value].copy(), cyls[:self.cylinders_found()[0]].copy()

    def plane_points(self, nframes=None, plane_cap=255, cap_per_frame=None):
        """plane_cloud of PlaneDetection_CAPE::runPlaneDetection (PlaneExtractor.cpp:165-190) for the frames of the
        last call: a list (per frame) of lists (per plane) of (n, 3) float32 arrays, pixels in row-major order."""
        nf = nframes or max(getattr(self, "_nframes", 1) or 1, 1)
        N = cap_per_frame or self.H * self.W
        pts = np.zeros((nf, N, 3), np.float32)
        offs = np.zeros((nf, plane_cap + 1), np.int32)
        _check(self.L.drfe_cape_plane_points(self.h, _ptr(pts), N, _ptr(offs), plane_cap))
        return pts, offs

    def plane_points_voxel(self, leaf=0.05, nframes=None, plane_cap=255, cap_per_frame=None):
        """pcl::VoxelGrid (leaf^3) of every plane's point list (Frame.cc:1121-1125): same layout as plane_points"""
        nf = nframes or max(getattr(self, "_nframes", 1) or 1, 1)
        N = cap_per_frame or self.H * self.W
        pts = np.zeros((nf, N, 3), np.float32)
        offs = np.zeros((nf, plane_cap + 1), np.int32)
        _check(self.L.drfe_cape_plane_points_voxel(self.h, leaf, _ptr(pts), N, _ptr(offs), plane_cap))
        return pts, offs

    def third_cloud(self, max_point_dist, nframes=None):
        """the 1/3-resolution cloud of Frame::ComputePlanes_CAPE (Frame.cc:1153-1172) of the last batch's depth"""
        nf = nframes or max(getattr(self, "_nframes", 1) or 1, 1)
        out = np.empty((nf, (self.H + 2) // 3, (self.W + 2) // 3, 3), np.float32)
        _check(self.L.drfe_cape_third_cloud(self.h, max_point_dist, _ptr(out)))
        return out

    def third_cloud_normals(self, max_point_dist, max_depth_change_factor=0.05, smoothing_size=10.0, nframes=None):
        """the 1/3-resolution cloud and PCL's integral-image normals on it (Frame.cc:1153-1216) -> (cloud, normals), NaN where PCL has NaN"""
        nf = nframes or max(getattr(self, "_nframes", 1) or 1, 1)
        cloud = np.empty((nf, (self.H + 2) // 3, (self.W + 2) // 3, 3), np.float32)
        normals = np.empty_like(cloud)
        _check(self.L.drfe_cape_third_cloud_normals(self.h, max_point_dist, max_depth_change_factor, smoothing_size, _ptr(cloud), _ptr(normals)))
        return cloud, normals

    def cylinders_found(self):
        """length of cylinder_segments_final per frame of the last call (CAPE.cpp:434-445)"""
        out = np.zeros(max(getattr(self, "_nframes", 1) or 1, 1), np.int32)
        _check(self.L.drfe_cape_cylinders_found(self.h, _ptr(out)))
        return out

    def cyl_maps(self, frame=0):
        cm = np.zeros((self.ncy, self.ncx), np.int32)
        em = np.zeros((self.ncy, self.ncx), np.uint8)
        _check(self.L.drfe_cape_get_cyl_maps(self.h, frame, _ptr(cm), _ptr(em)))
        return cm, em

    # ---- batched
    def enqueue_depth(self, depth, fx, fy, cx, cy, mem_kind=MEM_HOST, nframes=None, row_stride=None, frame_stride=None):
        if isinstance(depth, np.ndarray):
            assert depth.dtype == np.float32 and depth.ndim == 3 and depth.strides[2] == 4
            nframes, row_stride, frame_stride = depth.shape[0], depth.strides[1] // 4, depth.strides[0] // 4
            self._keep = depth
        _check(self.L.drfe_cape_enqueue_depth(self.h, nframes, _ptr(depth), row_stride, frame_stride, mem_kind,
                                              fx, fy, cx, cy))
        self._nframes = nframes

    def enqueue_depth_u16(self, depth, depth_factor, fx, fy, cx, cy, mem_kind=MEM_HOST, nframes=None, row_stride=None,
                          frame_stride=None):
        """raw 16-bit depth: z = float(d) * depth_factor on the device (Frame.cc:113-115)"""
        if isinstance(depth, np.ndarray):
            assert depth.dtype == np.uint16 and depth.ndim == 3 and depth.strides[2] == 2
            nframes, row_stride, frame_stride = depth.shape[0], depth.strides[1] // 2, depth.strides[0] // 2
            self._keep = depth
        _check(self.L.drfe_cape_enqueue_depth_u16(self.h, nframes, _ptr(depth), row_stride, frame_stride, mem_kind,
                                                  depth_factor, fx, fy, cx, cy))
        self._nframes = nframes

    def process_depth_batch(self, depth, fx, fy, cx, cy, depth_factor=1.0, seg=None, planes=None, nplanes=None,
                            chunk_frames=0):
        """Chunk-pipelined host-in / host-out batch call (asynchronous; finish_batch() waits).
        depth: (B,H,W) float32, or uint16 scaled by depth_factor."""
        assert depth.ndim == 3 and depth.dtype in (np.float32, np.uint16) and depth.strides[2] == depth.itemsize
        nf, es = depth.shape[0], depth.itemsize
        seg = np.empty((nf, self.H, self.W), np.uint8) if seg is None else seg
        planes = np.zeros((nf, self.plane_cap), PLANE_DTYPE) if planes is None else planes
        nplanes = np.empty(nf, np.int32) if nplanes is None else nplanes
        ncyl = np.zeros(nf, np.int32)
        cyls = np.zeros((nf, max(self.cyl_cap, 1)), CYL_DTYPE)
        self._keep = (depth, seg, planes, nplanes, ncyl, cyls)
        _check(self.L.drfe_cape_process_depth_batch(self.h, nf, _ptr(depth), int(depth.dtype == np.uint16), depth_factor,
                                                    depth.strides[1] // es, depth.strides[0] // es, fx, fy, cx, cy, _ptr(seg),
                                                    _ptr(planes), planes.shape[1], _ptr(nplanes), _ptr(cyls), self.cyl_cap,
                                                    _ptr(ncyl), chunk_frames))
        self._nframes = nf
        return seg, planes, nplanes, ncyl, cyls

    def finish_batch(self):
        _check(self.L.drfe_cape_finish_batch(self.h))

    def enqueue_cloud(self, cloud, mem_kind=MEM_HOST, nframes=None, frame_stride=None):
        if isinstance(cloud, np.ndarray):
            assert cloud.dtype == np.float32 and cloud.ndim == 2
            nframes, frame_stride = cloud.shape[0], cloud.strides[0] // 4
            self._keep = cloud
        _check(self.L.drfe_cape_enqueue_cloud(self.h, nframes, _ptr(cloud), frame_stride, mem_kind))
        self._nframes = nframes

    def download(self, seg=None, planes=None, nplanes=None, with_cylinders=False):
        nf = self._nframes
        seg = np.empty((nf, self.H, self.W), np.uint8) if seg is None else seg
        planes = np.zeros((nf, self.plane_cap), PLANE_DTYPE) if planes is None else planes
        nplanes = np.empty(nf, np.int32) if nplanes is None else nplanes
        ncyl = np.empty(nf, np.int32)
        if with_cylinders:
            cyls = np.zeros((nf, max(self.cyl_cap, 1)), CYL_DTYPE)
            _check(self.L.drfe_cape_download(self.h, _ptr(seg), _ptr(planes), planes.shape[1], _ptr(nplanes), _ptr(cyls),
                                             self.cyl_cap, _ptr(ncyl)))
            return seg, planes, nplanes, ncyl, cyls, self.cylinders_found()
        _check(self.L.drfe_cape_download(self.h, _ptr(seg), _ptr(planes), planes.shape[1], _ptr(nplanes), None, 0,
                                         _ptr(ncyl)))
        return seg, planes, nplanes

    def sync(self):
        _check(self.L.drfe_cape_sync(self.h))

    def stream(self):
        return self.L.drfe_cape_stream(self.h)

    def debug_counters(self, frame=0):
        out = np.zeros(16, np.int64)
        _check(self.L.drfe_cape_debug_counters(self.h, frame, _ptr(out)))
        return out

    def set_profiling(self, on=True):
        _check(self.L.drfe_cape_set_profiling(self.h, int(on)))

    def stage_times(self):
        return _stage_times(self.L.drfe_cape_stage_times, self.h)

    # ---- intermediates
    def cloud(self, frame=0):
        out = np.empty(3 * self.H * self.W, np.float32)
        _check(self.L.drfe_cape_get_cloud(self.h, frame, _ptr(out)))
        return out

    def cells(self, frame=0):
        out = np.zeros(self.ncx * self.ncy, PLANE_DTYPE)
        _check(self.L.drfe_cape_get_cells(self.h, frame, _ptr(out)))
        return out

    def grid_maps(self, frame=0):
        pm = np.zeros((self.ncy, self.ncx), np.int32)
        em = np.zeros((self.ncy, self.ncx), np.uint8)
        _check(self.L.drfe_cape_get_grid_maps(self.h, frame, _ptr(pm), _ptr(em)))
        return pm, em


class PlaneDetectionCAPE:
    """PlaneDetection_CAPE (PlaneExtractor.h:84-115): readDepthImage + runPlaneDetection with
    the public result fields nr_planes, nr_cylinders, seg_output, plane_params."""

    def __init__(self, PATCH_SIZE=20, MAX_MERGE_DIST=50.0, cylinder_detection=False, device=0):
        self.PATCH_SIZE, self.MAX_MERGE_DIST, self.cylinder_detection = PATCH_SIZE, MAX_MERGE_DIST, cylinder_detection
        self.COS_ANGLE_MAX = float(np.float32(np.cos(np.pi / 12)))   # PlaneExtractor.h:111
        self.device = device
        self._cape = None
        self.depth_img = None

    def readDepthImage(self, depthImg, K):
        if depthImg is None or depthImg.size == 0 or depthImg.dtype != np.float32:
            return False          # reference prints a warning and returns false (PlaneExtractor.cpp:104-107)
        self.depth_img, self.K_ = depthImg, np.asarray(K, np.float32)
        return True

    def runPlaneDetection(self):
        H, W = self.depth_img.shape
        if self._cape is None or (self._cape.H, self._cape.W) != (H, W):
            self._cape = CAPE(H, W, self.PATCH_SIZE, self.PATCH_SIZE, self.cylinder_detection, self.COS_ANGLE_MAX,
                              self.MAX_MERGE_DIST, device=self.device)
        K = self.K_
        (self.nr_planes, self.nr_cylinders, self.seg_output, self.plane_params,
         self.cylinder_params) = self._cape.process_depth(self.depth_img, float(K[0, 0]), float(K[1, 1]),
                                                          float(K[0, 2]), float(K[1, 2]))
