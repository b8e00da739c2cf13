"""libm_exact.cuh (the device restatements of glibc's atan2f / atanf / expf) against the running
libm: the header compiles for the host too, so the check runs on CPU. Bit-exact, no tolerance.

The ranges cover what the hot path feeds these functions: expf on -amplification * dist with
dist in [0, kernel_threshold] (segmenter.cpp:588), atanf on z / range_xy (clusterer.cpp:86,
segmenter.cpp:157), atan2f on LiDAR coordinates (dataloader.cpp:96, clusterer.cpp:74).
"""
import ctypes as C
import os
import struct
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SRC = os.path.join(ROOT, "tests", "native", "libm_check.cpp")


def _bits(x: float) -> int:
    return struct.unpack("<I", struct.pack("<f", x))[0]


@pytest.fixture(scope="module")
def lib(tmp_path_factory):
    so = str(tmp_path_factory.mktemp("libm") / "libm_check.so")
    cxx = "/usr/bin/g++" if os.path.exists("/usr/bin/g++") else "g++"
    subprocess.run([cxx, "-O2", "-ffp-contract=off", "-std=c++17", "-shared", "-fPIC", "-pthread", "-o", so, SRC],
                   check=True)
    h = C.CDLL(so)
    for fn in (h.check_expf_range, h.check_atanf_range):
        fn.restype = C.c_uint64
        fn.argtypes = [C.c_uint32, C.c_uint32, C.c_int]
    h.check_atan2f_random.restype = C.c_uint64
    h.check_atan2f_random.argtypes = [C.c_uint64, C.c_uint32]
    return h


def test_expf_every_float_of_the_jcp_range(lib):
    threads = min(8, os.cpu_count() or 1)
    # every negative float from -1e-6 down to -16 (JCP uses [-5, 0]) and a positive range
    assert lib.check_expf_range(_bits(-1e-6), _bits(-16.0), threads) == 0
    assert lib.check_expf_range(_bits(1e-6), _bits(4.0), threads) == 0


def test_atanf_every_float_of_the_elevation_range(lib):
    threads = min(8, os.cpu_count() or 1)
    # |x| from 1e-6 to 64 covers every z / range_xy of a LiDAR return (both signs are checked)
    assert lib.check_atanf_range(_bits(1e-6), _bits(64.0), threads) == 0


def test_atan2f_random_lidar_coordinates(lib):
    assert lib.check_atan2f_random(20_000_000, 1234) == 0
