"""The drop-in C++ surface (include/lidar_processing_lib/*.hpp): compiles against PCL / OpenCV
headers (here: the shims under oracle/shim), links the C-ABI library, translates status codes into
the reference's exceptions, and - on a GPU - reproduces the oracle through the node's call sequence."""
import os
import struct
import subprocess

import numpy as np
import pytest

import lidar_processing_v2_b200 as lpl
from conftest import HAVE_GPU, ROOT
from oracle.oracle import NODE_CLUSTER_CFG

EXE = os.path.join(ROOT, "tests", "native", "adaptor_main")
SRC = os.path.join(ROOT, "tests", "native", "adaptor_main.cpp")


@pytest.fixture(scope="module")
def exe():
    lpl.load_library()  # builds liblpl_b200.so if needed
    pkg = os.path.dirname(lpl.SO_PATH)
    cxx = "/usr/bin/g++" if os.path.exists("/usr/bin/g++") else "g++"
    deps = [SRC] + [os.path.join(ROOT, "include", "lidar_processing_lib", f)
                    for f in os.listdir(os.path.join(ROOT, "include", "lidar_processing_lib")) if f.endswith(".hpp")]
    if not os.path.exists(EXE) or any(os.path.getmtime(d) > os.path.getmtime(EXE) for d in deps + [lpl.SO_PATH]):
        cmd = [cxx, "-std=c++17", "-O2", "-Wall", "-Wextra", "-I", os.path.join(ROOT, "include"),
               "-I", os.path.join(ROOT, "oracle", "shim"), SRC, "-o", EXE, "-L", pkg, "-llpl_b200",
               "-Wl,-rpath," + pkg]
        res = subprocess.run(cmd, capture_output=True, text=True)
        assert res.returncode == 0, res.stderr
    return EXE


def _write_frame(path, pts, ring):
    with open(path, "wb") as f:
        f.write(struct.pack("<I", pts.shape[0]))
        f.write(np.ascontiguousarray(pts, np.float32).tobytes())
        f.write(np.ascontiguousarray(ring, np.uint16).tobytes())


@pytest.mark.skipif(HAVE_GPU, reason="checks the no-device error translation")
def test_adaptors_compile_link_and_fail_loudly_without_gpu(exe, tmp_path, golden0):
    fin, fout = str(tmp_path / "in.bin"), str(tmp_path / "out.bin")
    _write_frame(fin, golden0["pts"][:1000], golden0["ring"][:1000])
    res = subprocess.run([exe, fin, fout], capture_output=True, text=True)
    assert res.returncode == 3, (res.returncode, res.stderr)
    assert "no CUDA device" in res.stderr and not os.path.exists(fout)


def test_find_package_resolves_like_the_reference_package(tmp_path):
    """cmake/lidar_processing_libConfig.cmake: the node's own `find_package(lidar_processing_lib REQUIRED)` +
    `target_link_libraries(processor lidar_processing_lib)` (src/processor/CMakeLists.txt:17,51-56) configure and
    build against this repository with the node's warning flags (-Wall -Wextra -Werror), no edit of the node."""
    import shutil

    cmake = shutil.which("cmake")
    if cmake is None:
        pytest.skip("cmake not installed")
    lpl.load_library()
    (tmp_path / "CMakeLists.txt").write_text(
        "cmake_minimum_required(VERSION 3.16)\nproject(node_like CXX)\nset(CMAKE_CXX_STANDARD 17)\n"
        "find_package(lidar_processing_lib REQUIRED)\n"
        f"add_executable(node_like {SRC})\n"
        "target_compile_options(node_like PRIVATE -Wall -Wextra -Werror)\n"
        f"target_include_directories(node_like PRIVATE {os.path.join(ROOT, 'oracle', 'shim')})\n"
        "target_link_libraries(node_like lidar_processing_lib)\n")
    build = tmp_path / "build"
    cxx = "/usr/bin/g++" if os.path.exists("/usr/bin/g++") else "g++"
    res = subprocess.run([cmake, "-S", str(tmp_path), "-B", str(build), f"-Dlidar_processing_lib_DIR={os.path.join(ROOT, 'cmake')}",
                          f"-DCMAKE_CXX_COMPILER={cxx}", "-DCMAKE_BUILD_TYPE=Release"], capture_output=True, text=True)
    assert res.returncode == 0, res.stdout + res.stderr
    res = subprocess.run([cmake, "--build", str(build), "-j", "4"], capture_output=True, text=True)
    assert res.returncode == 0, res.stdout + res.stderr
    assert os.path.exists(build / "node_like")


@pytest.mark.gpu
def test_adaptors_reproduce_the_reference_call_sequence(exe, tmp_path, golden0, port):
    pts, ring = golden0["pts"], golden0["ring"].astype(np.uint16)
    fin, fout = str(tmp_path / "in.bin"), str(tmp_path / "out.bin")
    _write_frame(fin, pts, ring)
    res = subprocess.run([exe, fin, fout], capture_output=True, text=True)
    assert res.returncode == 0, res.stderr
    raw = open(fout, "rb").read()
    o = 0
    (n,) = struct.unpack_from("<I", raw, o); o += 4
    noise = np.frombuffer(raw, np.uint8, n, o); o += n
    labels = np.frombuffer(raw, np.uint32, n, o); o += 4 * n
    (m,) = struct.unpack_from("<I", raw, o); o += 4
    clabels = np.frombuffer(raw, np.int32, m, o); o += 4 * m
    (K,) = struct.unpack_from("<I", raw, o); o += 4
    sizes = np.frombuffer(raw, np.uint32, K, o); o += 4 * K
    hxy = np.frombuffer(raw, np.float64, 2 * int(sizes.sum()), o).reshape(-1, 2)
    o += 16 * int(sizes.sum())
    box_dt = np.dtype([("c", np.float64, 8), ("area", np.float32), ("yaw", np.float32), ("valid", np.uint32)])
    boxes = np.frombuffer(raw, box_dt, K, o)
    assert n == pts.shape[0]
    assert np.array_equal(noise, np.unpackbits(golden0["dror_exact"])[:n])
    assert np.array_equal(labels, golden0["labels"].astype(np.uint32))
    assert np.array_equal(clabels, golden0["cluster_labels"].astype(np.int32))
    off = golden0["hull_offsets"]
    assert np.array_equal(np.diff(off), sizes)
    assert np.abs(hxy - golden0["hull_xy"].astype(np.float64)).max() <= 1e-5
    # Polygonizer::boundingBoxRotatingCalipers against the reference's own boxes for this frame
    gb = np.load(os.path.join(ROOT, "tests", "golden", "kitti_polygonizer.npz"))["kitti_f000_boxes"][:, :11]
    assert np.array_equal(boxes["valid"] != 0, gb[:, 10] != 0)
    assert np.array_equal(boxes["c"].view(np.uint64), gb[:, :8].view(np.uint64))
    assert np.array_equal(boxes["area"], gb[:, 8].astype(np.float32))
    assert np.array_equal(boxes["yaw"], gb[:, 9].astype(np.float32))
    o += box_dt.itemsize * K
    # KDTree<float, 3> adaptor: k_nearest / radius_search / radius_search_k_nearest on the first 5000 points
    (nq,) = struct.unpack_from("<I", raw, o); o += 4
    assert nq == 64
    ne_dt = np.dtype([("index", np.uint32), ("distance", np.float32)])
    tp = pts[:5000, :3].astype(np.float32)
    for q in range(nq):
        t = tp[q * 7 % 5000]
        d = tp - t
        dist = (d[:, 0] * d[:, 0] + (d[:, 1] * d[:, 1] + (d[:, 2] * d[:, 2] + np.float32(0)))).astype(np.float32)
        order = np.argsort(dist, kind="stable")
        (c,) = struct.unpack_from("<I", raw, o); o += 4
        ne = np.frombuffer(raw, ne_dt, c, o); o += 8 * c
        assert c == 3 and np.array_equal(ne["index"], order[:3]) and np.array_equal(ne["distance"], dist[order[:3]])
        within = order[dist[order] <= np.float32(0.25)]
        (c,) = struct.unpack_from("<I", raw, o); o += 4
        ne = np.frombuffer(raw, ne_dt, c, o); o += 8 * c
        assert c == within.size and np.array_equal(np.sort(ne["index"]), np.sort(within))
        assert np.all(np.diff(ne["distance"]) >= 0)           # KDTree(sort = true)
        (c,) = struct.unpack_from("<I", raw, o); o += 4
        ne = np.frombuffer(raw, ne_dt, c, o); o += 8 * c
        assert c == min(2, within.size) and np.array_equal(ne["index"], within[:c])
    assert o == len(raw)
    _ = (port, NODE_CLUSTER_CFG)
