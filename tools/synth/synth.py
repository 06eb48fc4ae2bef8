"""ctypes loader of tools/synth/libdrfe_synth.so: the synthetic RGB-D frames of the tests and bench.py.
Input generator only — nothing here touches the GPU or the product library."""
import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_lib = None


def lib():
    global _lib
    if _lib is None:
        path = os.path.join(_HERE, "libdrfe_synth.so")
        if not os.path.exists(path):
            raise RuntimeError("tools/synth/libdrfe_synth.so is missing: run `python __graft_entry__.py` (build())")
        _lib = C.CDLL(path)
        vp = C.c_void_p
        _lib.drfe_synth_frame.argtypes = [C.c_int, C.c_int, C.c_int, C.c_uint32, C.c_float, vp, vp, vp, vp, vp, vp]
        _lib.drfe_synth_frame.restype = C.c_int
    return _lib


def synth_frame(width=640, height=480, scene=0, seed=20260000, depth_unit_scale=1.0):
    """Deterministic procedural RGB-D frame -> (gray u8 HxW, depth f32 HxW, (fx, fy, cx, cy))."""
    gray = np.empty((height, width), np.uint8)
    depth = np.empty((height, width), np.float32)
    K = [C.c_float() for _ in range(4)]
    rc = lib().drfe_synth_frame(width, height, scene, seed, depth_unit_scale, gray.ctypes.data_as(C.c_void_p),
                                depth.ctypes.data_as(C.c_void_p), *[C.cast(C.byref(k), C.c_void_p) for k in K])
    if rc != 0:
        raise ValueError("drfe_synth_frame: bad argument")
    return gray, depth, tuple(k.value for k in K)
