"""B200-native (sm_100a) implementation of LiDAR-Processing-V2's per-frame perception hot path.

The product is the C-ABI shared library ``liblpl_b200.so`` (include/lpl_b200.h, csrc/*.cu) plus the
header-only C++ adaptors in include/lidar_processing_lib/. This package holds the in-tree build
and a thin ctypes binding used by the tests and bench.py.
"""
from .build import SO_PATH, build_native  # noqa: F401
from .native import (  # noqa: F401
    JCP_AS_REFERENCE,
    BBOX_DTYPE,
    BOX_PCA,
    BOX_ROTATING_CALIPERS,
    JCP_CLEAN,
    STAGE_ALL,
    STAGE_BOXES,
    STAGE_CLUSTER,
    STAGE_DROR,
    STAGE_HULLS,
    STAGE_RING,
    STAGE_SEGMENT,
    BatchBuffers,
    PackedBuffers,
    PLANES,
    RGB_DTYPE,
    glibc_rand_stream,
    ClusterCfg,
    Context,
    DrorCfg,
    LplError,
    PinnedBuffer,
    SegmenterCfg,
    load_library,
    pcd_read,
)
