"""The DRFE_WITH_OPENCV / DRFE_WITH_EIGEN branches of the host adapters (dr-slam_b200/host/*.h) are the code a DR-SLAM
maintainer compiles, and this container has neither library: tests/host/mock holds functional stand-ins for the few
cv:: / Eigen:: types those branches touch.  CPU: the branches compile, with the reference's argument lists.  GPU: the
program runs and reports, line for line, what the stand-in build (example_frontend) reports."""
import os
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HOST = os.path.join(ROOT, "dr-slam_b200", "host")
SRC = os.path.join(ROOT, "tests", "host", "adapters_opencv.cpp")
EXE = os.path.join(ROOT, "tests", "host", "adapters_opencv")


def build():
    cmd = ["g++", "-std=c++17", "-O1", "-Wall", "-Werror", "-ffp-contract=off", "-DDRFE_WITH_OPENCV", "-DDRFE_WITH_EIGEN",
           "-I" + os.path.join(ROOT, "tests", "host", "mock"), "-I" + HOST, "-I" + os.path.join(ROOT, "tools", "synth"),
           SRC, os.path.join(ROOT, "tools", "synth", "synth.cpp"), "-o", EXE,
           "-L" + os.path.join(ROOT, "dr-slam_b200"), "-ldrfe", "-Wl,-rpath," + os.path.join(ROOT, "dr-slam_b200"), "-lpthread"]
    return subprocess.run(cmd, capture_output=True, text=True, timeout=300)


def test_opencv_and_eigen_branches_compile():
    r = build()
    assert r.returncode == 0, r.stderr
    # the matcher / vocabulary adapters sit on top of ORBextractor.h: type-check them in the same configuration
    r = subprocess.run(["g++", "-std=c++17", "-fsyntax-only", "-Wall", "-Werror", "-DDRFE_WITH_OPENCV", "-DDRFE_WITH_EIGEN",
                        "-I" + os.path.join(ROOT, "tests", "host", "mock"), "-I" + HOST, "-x", "c++", "-"],
                       input='#include "ORBmatcher.h"\n#include "CAPE.h"\n#include "PlaneExtractor.h"\nint main() { return 0; }\n', capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stderr


@pytest.mark.gpu
@pytest.mark.parametrize("scene,seed", [(1, 20260042), (2, 20260011)])
def test_opencv_branch_reports_what_the_standin_build_reports(scene, seed):
    if not os.path.exists(EXE):
        assert build().returncode == 0
    a = subprocess.run([EXE, str(scene), str(seed)], capture_output=True, text=True, timeout=120)
    b = subprocess.run([os.path.join(HOST, "example_frontend"), str(scene), str(seed)], capture_output=True, text=True, timeout=120)
    assert a.returncode == 0 and b.returncode == 0, a.stderr + b.stderr
    la, lb = a.stdout.splitlines(), b.stdout.splitlines()
    nplanes = int(lb[0].split()[lb[0].split().index("planes") + 1])
    assert la[:1 + nplanes] == lb[:1 + nplanes]                  # keypoints, descriptors, seg_output, plane points, plane parameters
    seg_hash = lb[0].split()[lb[0].split().index("seg_hash") + 1]
    # CAPE::process on an Eigen::MatrixXf cloud = the fused depth path; plane_segments_final is appended to
    assert la[1 + nplanes] == "process planes %d cylinders 0 seg_hash %s appended %d" % (nplanes, seg_hash, nplanes)
    assert la[2 + nplanes] == "empty_image keeps 3"
    assert la[-1].startswith("peac planes") and la[-1] == lb[-1]     # Planar_SLAM::PlaneDetection (PlaneExtractor.h), cv::Mat in / out
