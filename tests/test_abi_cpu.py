"""CPU suite: the C-ABI library loads, exports every symbol include/lpl_b200.h declares, keeps the
reference's configuration defaults, and fails loudly (no CPU fallback) without a GPU."""
import ctypes as C
import os
import re

import numpy as np
import pytest

import lidar_processing_v2_b200 as lpl
from conftest import HAVE_GPU, ROOT
from lidar_processing_v2_b200.native import EXPORTS, LPL_ERR_NO_DEVICE


def test_library_exports_every_declared_symbol():
    hdr = open(os.path.join(ROOT, "include", "lpl_b200.h")).read()
    declared = sorted(set(re.findall(r"\b(lpl_[a-z0-9_]+)\s*\(", hdr)))
    lib = lpl.load_library()
    missing = [n for n in declared if not hasattr(lib, n)]
    assert not missing, missing
    assert sorted(EXPORTS) == declared


def test_header_is_plain_c(tmp_path):
    """The boundary is a C ABI: include/lpl_b200.h compiles as C99 with pedantic warnings as errors (no C++ types,
    no CUDA or torch types in any signature), so any FFI of the reference's host language can bind it."""
    import subprocess

    src = tmp_path / "abi.c"
    src.write_text('#include "lpl_b200.h"\nint main(void) { return sizeof(lpl_segmenter_cfg) > 0 ? 0 : 1; }\n')
    inc = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "include")
    res = subprocess.run(["gcc", "-std=c99", "-Wall", "-Wextra", "-Werror", "-pedantic", "-I", inc, "-fsyntax-only", str(src)],
                         capture_output=True, text=True)
    assert res.returncode == 0, res.stderr


def test_defaults_mirror_reference_structs():
    lib = lpl.load_library()
    s = lpl.SegmenterCfg()
    lib.lpl_segmenter_default_cfg(C.byref(s))
    # segmenter.hpp:87-112
    assert (s.image_width, s.image_height, s.assume_unorganized_cloud) == (2048, 64, 0)
    assert abs(s.elevation_down_deg + 24.8) < 1e-6 and s.elevation_up_deg == 2.0
    assert (s.grid_radial_spacing_m, s.grid_slice_resolution_deg) == (2.0, 1.0)
    assert abs(s.sensor_height_m - 1.73) < 1e-6 and s.amplification_factor == 5.0
    d = lpl.DrorCfg()
    lib.lpl_dror_default_cfg(C.byref(d))
    assert abs(d.radius_multiplier_m_per_m - 0.02) < 1e-7 and abs(d.min_search_radius_m - 0.1) < 1e-7
    assert d.min_neighbours == 4  # noise_remover.hpp:41-54
    c = lpl.ClusterCfg()
    lib.lpl_cluster_default_cfg(C.byref(c))
    assert abs(c.voxel_grid_range_resolution_m - 0.4) < 1e-7 and c.min_cluster_size == 3  # clusterer.hpp:61-68
    assert abs(c.voxel_grid_elevation_resolution_deg - 1.5) < 1e-7
    assert lib.lpl_version().startswith(b"lpl_b200")


@pytest.mark.skipif(HAVE_GPU, reason="checks the no-device error path")
def test_no_cpu_fallback():
    with pytest.raises(lpl.LplError) as e:
        lpl.Context(0)
    assert e.value.code == LPL_ERR_NO_DEVICE


def test_product_package_never_imports_oracle():
    pkg = os.path.join(ROOT, "lidar_processing_v2_b200")
    for dirpath, _, files in os.walk(pkg):
        for fn in files:
            if fn.endswith((".py", ".cu", ".cuh", ".h", ".hpp", ".cpp")):
                txt = open(os.path.join(dirpath, fn), errors="replace").read()
                assert "oracle" not in txt, os.path.join(dirpath, fn)


def _write_pcd(path, pts, mode="binary", extra_field=False, padding=0):
    n = pts.shape[0]
    fields = "x y z intensity" + (" t" if extra_field else "")
    nf = 5 if extra_field else 4
    hdr = ("# .PCD v0.7 - Point Cloud Data file format\nVERSION 0.7\nFIELDS %s\nSIZE %s\nTYPE %s\nCOUNT %s\n"
           "WIDTH %d\nHEIGHT 1\nVIEWPOINT 0 0 0 1 0 0 0\nPOINTS %d\nDATA %s\n") % (
        fields, " ".join(["4"] * nf), " ".join(["F"] * nf), " ".join(["1"] * nf), n, n, mode)
    with open(path, "wb") as f:
        f.write(hdr.encode())
        if mode == "binary":
            rec = np.zeros((n, nf), np.float32)
            rec[:, :4] = pts
            f.write(rec.tobytes())
            f.write(b"\x00" * padding)  # the reference's files carry padding behind the payload
        else:
            for r in pts:
                f.write((" ".join(repr(float(v)) for v in r) + (" 0" if extra_field else "") + "\n").encode())


def test_pcd_reader(tmp_path):
    """lpl_pcd_read (replaces pcl::io::loadPCDFile<PointXYZI>, dataloader.cpp:165) on the layout of the
    reference's data/*.pcd (binary, x y z intensity, trailing padding), a wider record, and ascii."""
    import numpy as np

    from oracle.oracle import read_pcd_xyzi

    rng = np.random.default_rng(0)
    pts = np.round(rng.normal(0, 20, (5000, 4)), 3).astype(np.float32)
    for name, kw in (("a.pcd", dict(padding=3906)), ("b.pcd", dict(extra_field=True)), ("c.pcd", dict(mode="ascii"))):
        path = str(tmp_path / name)
        _write_pcd(path, pts, **kw)
        got = lpl.pcd_read(path)
        assert got.shape == pts.shape and np.array_equal(got.view(np.uint32), pts.view(np.uint32)), name
    # the oracle's own reader (tests only) agrees on the reference layout
    assert np.array_equal(read_pcd_xyzi(str(tmp_path / "a.pcd")), lpl.pcd_read(str(tmp_path / "a.pcd")))
    with pytest.raises(lpl.LplError):
        lpl.pcd_read(str(tmp_path / "missing.pcd"))
    if os.path.isdir("/root/reference/data"):
        ref_file = "/root/reference/data/0000000000.pcd"
        got = lpl.pcd_read(ref_file)
        assert got.shape == (123398, 4) and np.array_equal(got, read_pcd_xyzi(ref_file))


def test_pcd_reader_rejects_crafted_headers(tmp_path):
    """Headers whose field sizes / counts would send the x, y, z offsets outside the record (negative or huge SIZE /
    COUNT of fields that are not even read), overflow the record length or the 32-bit point count are refused with a
    status code - no read outside the staging buffer, no exception across the C ABI."""
    import lidar_processing_v2_b200 as lpl

    def pcd(fields, size, typ, count, points="4", data="binary", payload=b"\0" * 4096):
        path = tmp_path / f"bad{len(list(tmp_path.iterdir()))}.pcd"
        head = (f"# .PCD v0.7\nVERSION 0.7\nFIELDS {fields}\nSIZE {size}\nTYPE {typ}\nCOUNT {count}\n"
                f"WIDTH {points}\nHEIGHT 1\nVIEWPOINT 0 0 0 1 0 0 0\nPOINTS {points}\nDATA {data}\n")
        path.write_bytes(head.encode() + payload)
        return str(path)

    bad = [
        pcd("x a y b z", "4 1000 4 -1000 4", "F F F F F", "1 1 1 1 1"),      # offsets cancel out: y would sit at byte 1004 of a 12-byte record
        pcd("x y z pad", "4 4 4 4", "F F F F", "1 1 1 -3"),                  # negative count
        pcd("x y z pad", "4 4 4 1", "F F F U", "1 1 1 100000000"),           # record length overflow
        pcd("x y z", "4 4 4", "F F F", "1 1 1", points="99999999999"),       # more points than 32 bits hold
        pcd("x y z", "8 4 4", "F F F", "1 1 1"),                             # double x: not what the node consumes
        pcd("x y", "4 4", "F F", "1 1"),                                     # no z
        pcd("x y z", "4 4", "F F F", "1 1 1"),                               # SIZE list shorter than FIELDS
    ]
    for path in bad:
        with pytest.raises(lpl.LplError):
            lpl.pcd_read(path)
    # an ascii file whose rows are shorter than the header promises is a parse error, not an out-of-range read
    short = pcd("x y z intensity", "4 4 4 4", "F F F F", "1 1 1 1", data="ascii", payload=b"1 2 3\n4 5\n6\n\n")
    with pytest.raises(lpl.LplError):
        lpl.pcd_read(short)


def test_glibc_rand_replica_matches_libc():
    """lpl_pipeline_split_clouds colours clusters with the C library's rand() % 256 as the node does
    (processor.cpp:629-631); the replica of glibc's generator is checked against libc itself."""
    import ctypes

    import lidar_processing_v2_b200 as lpl

    libc = ctypes.CDLL("libc.so.6")
    for seed in (1, 42, 2024):
        libc.srand(seed)
        exp = np.array([libc.rand() for _ in range(2000)], np.int32)
        assert np.array_equal(lpl.glibc_rand_stream(seed, 2000), exp)
    assert np.array_equal(lpl.glibc_rand_stream(0, 10), lpl.glibc_rand_stream(1, 10))  # srand(0) == srand(1)
