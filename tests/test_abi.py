"""CPU: the C-ABI shared library loads, exports every function include/drfe.h declares,
fails loudly without a CUDA device (no CPU fallback), and the POD structs have the layouts
the reference types have (cv::KeyPoint = 28 bytes, PlaneSeg public members)."""
import ctypes as C
import os
import re

import numpy as np
import pytest

from conftest import ROOT


def declared_functions():
    txt = open(os.path.join(ROOT, "include", "drfe.h")).read()
    txt = re.sub(r"/\*.*?\*/", "", txt, flags=re.S)
    return sorted(set(re.findall(r"\b(drfe_[a-z0-9_]+)\s*\(", txt)))


def test_library_exports_every_declared_symbol(drfe):
    L = drfe.lib()
    names = declared_functions()
    assert len(names) >= 35
    missing = [n for n in names if not hasattr(L, n)]
    assert not missing, missing
    assert sorted(drfe.SYMBOLS) == names          # the binding covers the whole header


def test_struct_layouts(drfe):
    assert drfe.KP_DTYPE.itemsize == 28                      # sizeof(cv::KeyPoint)
    assert drfe.PLANE_DTYPE.itemsize == 152                  # PlaneSeg
    assert drfe.PLANE_DTYPE.fields["mean"][1] == 96 and drfe.PLANE_DTYPE.fields["d"][1] == 144


def test_version_and_error_text(drfe):
    L = drfe.lib()
    assert b"sm_100a" in L.drfe_version()
    assert L.drfe_orb_create(None, 640, 480, 1, 0, None) == drfe.ERR_ARG
    assert b"null" in L.drfe_last_error()


def test_bad_parameters_are_rejected(drfe):
    L = drfe.lib()
    h = C.c_void_p()
    bad = drfe.OrbParams(1000, 1.0, 8, 20, 7)               # scaleFactor must be > 1
    assert L.drfe_orb_create(C.byref(bad), 640, 480, 1, 0, C.byref(h)) == drfe.ERR_ARG
    bad = drfe.OrbParams(1000, 1.2, 99, 20, 7)
    assert L.drfe_orb_create(C.byref(bad), 640, 480, 1, 0, C.byref(h)) == drfe.ERR_ARG
    cp = drfe.CapeParams(480, 640, 0, 20, 0, 0.96, 50.0)
    assert L.drfe_cape_create(C.byref(cp), 1, 0, C.byref(h)) == drfe.ERR_ARG


def test_no_cpu_fallback(drfe):
    """Without a usable CUDA device every create call must fail loudly (DRFE_ERR_CUDA)."""
    if drfe.device_count() > 0:
        pytest.skip("a CUDA device is present")
    with pytest.raises(drfe.DrfeError) as e:
        drfe.ORBextractor(1000, 1.2, 8, 20, 7)
    assert e.value.code == drfe.ERR_CUDA and "no CPU fallback" in str(e.value)
    with pytest.raises(drfe.DrfeError) as e:
        drfe.CAPE(480, 640, 20, 20)
    assert e.value.code == drfe.ERR_CUDA


def test_product_does_not_reference_the_oracle():
    """The product tree must never import, link or call anything under oracle/."""
    pkg = os.path.join(ROOT, "dr-slam_b200")
    banned = ("libdrfe_oracle", "orc_orb", "orc_cape", "orc_", "from oracle", "import oracle", "oracle/", "py_ref")
    for dp, _, files in os.walk(pkg):
        for fn in files:
            if fn.endswith((".py", ".cu", ".cpp", ".h", ".cuh", ".hpp", "Makefile")):
                txt = open(os.path.join(dp, fn), errors="ignore").read()
                hits = [b for b in banned if b in txt]
                assert not hits, (os.path.join(dp, fn), hits)


def test_synth_frame_is_deterministic(drfe):
    a = drfe.synth_frame(320, 240, 0, 20260003)
    b = drfe.synth_frame(320, 240, 0, 20260003)
    c = drfe.synth_frame(320, 240, 0, 20260004)
    assert np.array_equal(a[0], b[0]) and np.array_equal(a[1], b[1]) and a[2] == b[2]
    assert not np.array_equal(a[0], c[0])
    assert a[2] == (262.5, 262.5, 159.5, 119.5)
    g, d, _ = drfe.synth_frame(640, 480, 1, 20260001, 1000.0)
    assert g.std() > 20 and 1000 < d[d > 0].min() and d.max() < 15000 and 0.005 < (d == 0).mean() < 0.2


def test_header_is_plain_c(tmp_path):
    """include/drfe.h is the drop-in boundary: it must compile as C99 (no C++ or torch types in the signatures)"""
    import shutil
    import subprocess
    if not shutil.which("gcc"):
        pytest.skip("no gcc")
    src = tmp_path / "t.c"
    src.write_text('#include "drfe.h"\nint main(void) { drfe_keypoint k; drfe_plane p; drfe_last_point l; drfe_track_params t; drfe_proj_query q;\n'
                   '  (void)k; (void)p; (void)l; (void)t; (void)q; return sizeof(drfe_keypoint) == 28 ? 0 : 1; }\n')
    r = subprocess.run(["gcc", "-std=c99", "-Wall", "-Wextra", "-pedantic", "-Werror", "-fsyntax-only", "-I", os.path.join(ROOT, "include"), str(src)],
                       capture_output=True, text=True)
    assert r.returncode == 0, r.stderr


def test_chunk_schedule_host_logic(tmp_path):
    """ChunkPipe::schedule (the chunk boundaries of the pipelined batch calls): every batch length 1..3000 and every
    chunk wish gives between 1 and kMaxChunks non-empty chunks that tile the batch; 256 frames give the ramped
    8, 16, 32 ... 32, 16, 8 schedule.  Host-only C++ (tests/host/chunk_schedule_test.cpp), no GPU."""
    import shutil
    import subprocess
    cuda_inc = "/usr/local/cuda/include"
    if not shutil.which("g++") or not os.path.exists(os.path.join(cuda_inc, "cuda_runtime.h")):
        pytest.skip("needs g++ and the CUDA headers")
    exe = str(tmp_path / "sched")
    r = subprocess.run(["g++", "-std=c++17", "-I", os.path.join(ROOT, "dr-slam_b200", "csrc"), "-I", cuda_inc,
                        os.path.join(ROOT, "tests", "host", "chunk_schedule_test.cpp"), "-o", exe], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    out = subprocess.run([exe], capture_output=True, text=True, timeout=60)
    assert out.returncode == 0, out.stdout + out.stderr
    assert out.stdout.strip() == "11: 8 16 32 32 32 32 32 32 16 16 8"


def test_vocabulary_argument_checks_without_a_gpu(drfe):
    """drfe_vocab_create applies TemplatedVocabulary::loadFromTextFile's header checks (TemplatedVocabulary.h:1359) before it
    touches a device, and fails loudly without one"""
    L = drfe.lib()
    h = C.c_void_p()
    parent = np.zeros(3, np.int32)
    leaf = np.ones(3, np.uint8)
    desc = np.zeros((3, 32), np.uint8)
    wt = np.ones(3, np.float64)
    args = lambda k, Lv, s, w: L.drfe_vocab_create(k, Lv, s, w, 3, parent.ctypes.data, leaf.ctypes.data, desc.ctypes.data, wt.ctypes.data, 0, C.byref(h))  # noqa: E731
    for bad in ((25, 6, 0, 0), (10, 0, 0, 0), (10, 11, 0, 0), (10, 6, 6, 0), (10, 6, 0, 4), (-1, 6, 0, 0)):
        assert args(*bad) == drfe.ERR_ARG, bad
    assert L.drfe_vocab_create(10, 6, 0, 0, 3, None, leaf.ctypes.data, desc.ctypes.data, wt.ctypes.data, 0, C.byref(h)) == drfe.ERR_ARG
    assert L.drfe_vocab_words(None) == 0
    L.drfe_vocab_destroy(None)
    if drfe.device_count() == 0:
        assert args(10, 6, 0, 0) == drfe.ERR_CUDA and b"no CPU fallback" in L.drfe_last_error()
