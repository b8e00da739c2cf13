"""CPU suite: the C-ABI library loads, exports every symbol include/lpl_b200.h declares, keeps the
reference's configuration defaults, and fails loudly (no CPU fallback) without a GPU."""
import ctypes as C
import os
import re

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
