"""The C++14 drop-in headers (include/icp_b200/*.h: the reference's ICPOptimizer / NearestNeighborSearch
class names and signatures over the icp_gpu_* C ABI): they compile as C++14 and link against the library
(CPU), and a reference-style driver gives the same pose as the Python binding (GPU)."""
import os
import subprocess

import numpy as np
import pytest

from icp_variants_b200 import capi

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
EXE = "/tmp/icp_b200_bunny_driver"


def build_driver():
    capi.build()
    libdir = os.path.dirname(capi.LIB_PATH)
    cmd = ["g++", "-std=c++14", "-Wall", "-Wextra", "-Werror", "-I", os.path.join(ROOT, "include"),
           os.path.join(ROOT, "examples", "bunny_driver.cpp"), "-o", EXE, "-L", libdir, "-licp_gpu", f"-Wl,-rpath,{libdir}"]
    subprocess.run(cmd, check=True)
    return EXE


def test_dropin_headers_compile_as_cpp14():
    build_driver()
    # the matcher classes alone, as a reference user would instantiate them
    src = '#include "icp_b200/NearestNeighbor.h"\nint main(){ NearestNeighborSearchFlann f; NearestNeighborSearchProjective p; NearestNeighborSearchBruteForce b; f.setMatchingMaxDistance(0.1f); return (int)f.queryMatches({}).size(); }\n'
    libdir = os.path.dirname(capi.LIB_PATH)
    subprocess.run(["g++", "-std=c++14", "-x", "c++", "-", "-I", os.path.join(ROOT, "include"), "-o", "/tmp/icp_b200_nn_check",
                    "-L", libdir, "-licp_gpu", f"-Wl,-rpath,{libdir}"], input=src, text=True, check=True)


def test_driver_without_gpu_reports_and_does_not_hang():
    if capi.lib().icp_gpu_device_count() > 0:
        pytest.skip("a GPU is present")
    exe = build_driver()
    r = subprocess.run(["/tmp/icp_b200_nn_check"], stdout=subprocess.PIPE, text=True, timeout=60)
    assert "no usable CUDA device" in r.stdout and r.returncode == 0
    assert os.path.exists(exe)


def _dump(path, cloud):
    with open(path, "wb") as f:
        f.write(np.int32(len(cloud.points)).tobytes())
        f.write(np.ascontiguousarray(cloud.points, np.float32).tobytes())
        f.write(np.ascontiguousarray(cloud.normals, np.float32).tobytes())


@pytest.mark.gpu
@pytest.mark.parametrize("linear,metric", [(1, 0), (1, 1), (1, 2), (0, 1)])
def test_cpp_driver_matches_python_binding(bunny, linear, metric):
    src, tgt, _, _ = bunny
    exe = build_driver()
    _dump("/tmp/icp_b200_src.bin", src)
    _dump("/tmp/icp_b200_tgt.bin", tgt)
    r = subprocess.run([exe, "/tmp/icp_b200_src.bin", "/tmp/icp_b200_tgt.bin", str(linear), str(metric), "20"],
                       stdout=subprocess.PIPE, text=True, check=True, timeout=120)
    line = [l for l in r.stdout.splitlines() if l.startswith("POSE")][0]
    pose_cpp = capi.pose_from_c(np.array(line.split()[1:], np.float32))
    with capi.Context(0) as c:
        cfg = capi.default_config()
        cfg.metric, cfg.minimizer, cfg.n_iterations, cfg.max_distance_sq = metric, 0 if linear else 1, 20, 0.0003
        c.set_config(cfg)
        c.set_target(tgt.points, tgt.normals, tgt.colors)
        c.set_source(src.points, src.normals, src.colors)
        pose_py, _, _ = c.estimate_pose()
    assert np.allclose(pose_cpp, pose_py, rtol=0, atol=1e-7)
    # ConvergenceMeasure filled from the device (icp_gpu_convergence_errors): main.cpp:105-120's 4 correspondences
    from oracle import oracle as orc
    _, _, gs, gt = bunny
    f = [l for l in r.stdout.splitlines() if l.startswith("RMSE")][0].split()
    assert int(f[5]) == 20
    assert float(f[1]) == pytest.approx(orc.rmse(pose_py, src.points[gs], tgt.points[gt]), rel=1e-6)
    assert float(f[3]) == pytest.approx(orc.benchmark_error(pose_py, src.points[gs], tgt.points[gt]), rel=1e-6)
