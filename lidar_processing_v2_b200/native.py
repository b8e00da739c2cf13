"""ctypes binding of include/lpl_b200.h (the C-ABI shared library liblpl_b200.so).

This is the only way Python reaches the product path; it never touches the CPU checkers and has no CPU
fallback: without the library or without a CUDA device every call raises.
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

from . import build as _build

LPL_OK = 0
LPL_ERR_INVALID_ARGUMENT = -1
LPL_ERR_CUDA = -2
LPL_ERR_CAPACITY = -3
LPL_ERR_NO_DEVICE = -4

JCP_AS_REFERENCE = 0
JCP_CLEAN = 1

STAGE_RING, STAGE_DROR, STAGE_SEGMENT, STAGE_CLUSTER, STAGE_HULLS = 1, 2, 4, 8, 16
STAGE_ALL = 31
STAGE_BOXES = 32
BOX_ROTATING_CALIPERS, BOX_PCA = 0, 1
# numpy view of lpl_bbox (80 bytes)
BBOX_DTYPE = np.dtype([("corners", np.float64, (4, 2)), ("area", np.float32), ("angle_rad", np.float32),
                       ("is_valid", np.int32), ("reserved", np.int32)])

# every symbol include/lpl_b200.h declares (tests check that the library exports them all)
EXPORTS = [
    "lpl_create", "lpl_destroy", "lpl_last_error", "lpl_version",
    "lpl_segmenter_default_cfg", "lpl_dror_default_cfg", "lpl_cluster_default_cfg",
    "lpl_segmenter_config", "lpl_dror_config", "lpl_cluster_config", "lpl_set_jcp_mode",
    "lpl_ring_partition", "lpl_dror_filter", "lpl_segment", "lpl_cluster", "lpl_convex_hull",
    "lpl_cluster_hulls", "lpl_bounding_boxes", "lpl_vehicle_match",
    "lpl_pipeline_upload", "lpl_pipeline_upload_device", "lpl_pipeline_upload_cloud2", "lpl_pipeline_upload_packed",
    "lpl_pipeline_upload_packed_xyz", "lpl_pcd_read",
    "lpl_device_bytes", "lpl_pipeline_run", "lpl_pipeline_use_graph", "lpl_pipeline_use_split", "lpl_pipeline_sync", "lpl_pipeline_status", "lpl_pipeline_download_packed",
    "lpl_pipeline_split_clouds", "lpl_glibc_rand_stream",
    "lpl_knn_build", "lpl_knn_token", "lpl_knn_k_nearest", "lpl_knn_radius_search",
    "lpl_pipeline_want_image", "lpl_pipeline_counts", "lpl_pipeline_download",
    "lpl_pipeline_download_batch", "lpl_host_alloc", "lpl_host_free",
    "lpl_profile_enable", "lpl_profile_read",
    "lpl_timer_start", "lpl_timer_stop_ms", "lpl_launch_count", "lpl_debug_segment",
    "lpl_debug_dror", "lpl_debug_cluster", "lpl_debug_hulls", "lpl_stream",
]


class SegmenterCfg(C.Structure):
    """lpl_segmenter_cfg == SegmenterConfiguration (segmenter.hpp:87-112)."""

    _fields_ = [
        ("elevation_up_deg", C.c_float),
        ("elevation_down_deg", C.c_float),
        ("image_width", C.c_int32),
        ("image_height", C.c_int32),
        ("assume_unorganized_cloud", C.c_int32),
        ("grid_radial_spacing_m", C.c_float),
        ("grid_slice_resolution_deg", C.c_float),
        ("ground_height_threshold_m", C.c_float),
        ("road_maximum_slope_m_per_m", C.c_float),
        ("min_distance_m", C.c_float),
        ("max_distance_m", C.c_float),
        ("sensor_height_m", C.c_float),
        ("kernel_threshold_distance_m", C.c_float),
        ("amplification_factor", C.c_float),
        ("z_min_m", C.c_float),
        ("z_max_m", C.c_float),
    ]


class DrorCfg(C.Structure):
    _fields_ = [
        ("radius_multiplier_m_per_m", C.c_float),
        ("min_search_radius_m", C.c_float),
        ("min_neighbours", C.c_uint32),
    ]


class ClusterCfg(C.Structure):
    _fields_ = [
        ("voxel_grid_range_resolution_m", C.c_float),
        ("voxel_grid_azimuth_resolution_deg", C.c_float),
        ("voxel_grid_elevation_resolution_deg", C.c_float),
        ("min_cluster_size", C.c_uint32),
    ]


class Frame(C.Structure):
    _fields_ = [("xyzw", C.c_void_p), ("n", C.c_uint32), ("ring", C.c_void_p)]


class Cloud2Frame(C.Structure):
    _fields_ = [("data", C.c_void_p), ("width", C.c_uint32), ("height", C.c_uint32), ("point_step", C.c_uint32),
                ("row_step", C.c_uint32), ("x_offset", C.c_int32), ("y_offset", C.c_int32), ("z_offset", C.c_int32),
                ("ring_offset", C.c_int32)]


class FrameResult(C.Structure):
    _fields_ = [
        ("noise", C.c_void_p),
        ("ring", C.c_void_p),
        ("labels", C.c_void_p),
        ("obstacle_index", C.c_void_p),
        ("cluster_labels", C.c_void_p),
        ("hull_offsets", C.c_void_p),
        ("hull_indices", C.c_void_p),
        ("hull_xy", C.c_void_p),
        ("zminmax", C.c_void_p),
        ("boxes", C.c_void_p),
        ("bgr", C.c_void_p),
        ("n", C.c_uint32),
        ("num_valid", C.c_uint32),
        ("num_obstacles", C.c_uint32),
        ("num_clusters", C.c_uint32),
        ("num_hull_vertices", C.c_uint32),
    ]


class BatchResult(C.Structure):
    _fields_ = [
        ("counts", C.c_void_p),
        ("stride", C.c_size_t),
        ("labels_u8", C.c_void_p),
        ("noise", C.c_void_p),
        ("ring", C.c_void_p),
        ("obstacle_index", C.c_void_p),
        ("cluster_labels", C.c_void_p),
        ("hull_offsets", C.c_void_p),
        ("hull_indices", C.c_void_p),
        ("hull_xy", C.c_void_p),
        ("zminmax", C.c_void_p),
        ("boxes", C.c_void_p),
    ]


# LPL_PLANE_* bits of lpl_pipeline_download_packed, in bit order: (name, dtype, floats/ints per element, count row)
PLANES = (("labels_u8", np.uint8, 1, 0), ("noise", np.uint8, 1, 0), ("ring", np.uint16, 1, 0),
          ("obstacle_index", np.uint32, 1, 2), ("cluster_labels", np.int32, 1, 2),
          ("hull_offsets", np.uint32, 1, 3), ("hull_indices", np.uint32, 1, 4),
          ("hull_xy", np.float32, 2, 4), ("zminmax", np.float32, 2, 3), ("boxes", BBOX_DTYPE, 1, 3))
PLANE_BIT = {name: 1 << i for i, (name, _, _, _) in enumerate(PLANES)}


class PackedResult(C.Structure):
    _fields_ = [
        ("counts", C.c_void_p),
        ("buffer", C.c_void_p),
        ("buffer_bytes", C.c_size_t),
        ("planes", C.c_uint32),
        ("offset", C.c_size_t * len(PLANES)),
        ("bytes_used", C.c_size_t),
    ]


class SplitResult(C.Structure):
    _fields_ = [
        ("counts", C.c_void_p),
        ("stride", C.c_size_t),
        ("ground", C.c_void_p),
        ("obstacle", C.c_void_p),
        ("unsegmented", C.c_void_p),
        ("clustered", C.c_void_p),
        ("marker_stride", C.c_size_t),
        ("marker_points", C.c_void_p),
        ("cluster_colors", C.c_void_p),
        ("colors_stride", C.c_size_t),
    ]


# numpy view of pcl::PointXYZRGB (32 bytes)
RGB_DTYPE = np.dtype([("xyz", np.float32, 3), ("w", np.float32), ("bgra", np.uint8, 4), ("pad", np.uint32, 3)])


def glibc_rand_stream(seed: int, count: int) -> np.ndarray:
    """First `count` outputs of the C library's rand() after srand(seed) (host-only utility)."""
    out = np.zeros(count, np.int32)
    load_library().lpl_glibc_rand_stream(seed, count, out.ctypes.data)
    return out


def pcd_read(path: str, lib=None) -> np.ndarray:
    """(n, 4) float32 x, y, z, intensity of a PCD file (lpl_pcd_read)."""
    lib = lib or load_library()
    n = C.c_uint32(0)
    rc = lib.lpl_pcd_read(path.encode(), None, 0, C.byref(n))
    if rc != 0:
        raise LplError(rc, f"cannot read PCD header of {path}")
    out = np.zeros((n.value, 4), np.float32)
    rc = lib.lpl_pcd_read(path.encode(), out.ctypes.data, n.value, C.byref(n))
    if rc != 0:
        raise LplError(rc, f"cannot read PCD payload of {path}")
    return out


class LplError(RuntimeError):
    def __init__(self, code: int, msg: str):
        super().__init__(f"lpl_b200 error {code}: {msg}")
        self.code = code


_lib = None


def load_library(path: str | None = None) -> C.CDLL:
    """dlopen liblpl_b200.so (building it first if the sources are newer). No GPU needed to load."""
    global _lib
    if _lib is not None and path is None:
        return _lib
    path = path or os.environ.get("LPL_B200_LIBRARY")  # a prebuilt library (experiments: tools/variants.sh)
    so = path or _build.SO_PATH
    if path is None and _build.needs_build():
        _build.build_native()
    if not os.path.exists(so):
        raise LplError(LPL_ERR_NO_DEVICE, f"{so} is missing: the CUDA extension was not built "
                       "(there is no CPU fallback)")
    L = C.CDLL(so)
    vp, u32, i32, sz = C.c_void_p, C.c_uint32, C.c_int32, C.c_size_t
    L.lpl_create.argtypes = [C.POINTER(vp), C.c_int, u32, u32, i32, i32]
    L.lpl_create.restype = C.c_int
    L.lpl_destroy.argtypes = [vp]
    L.lpl_destroy.restype = None
    L.lpl_last_error.argtypes = [vp]
    L.lpl_last_error.restype = C.c_char_p
    L.lpl_version.restype = C.c_char_p
    L.lpl_segmenter_default_cfg.argtypes = [C.POINTER(SegmenterCfg)]
    L.lpl_dror_default_cfg.argtypes = [C.POINTER(DrorCfg)]
    L.lpl_cluster_default_cfg.argtypes = [C.POINTER(ClusterCfg)]
    L.lpl_segmenter_config.argtypes = [vp, C.POINTER(SegmenterCfg)]
    L.lpl_dror_config.argtypes = [vp, C.POINTER(DrorCfg)]
    L.lpl_cluster_config.argtypes = [vp, C.POINTER(ClusterCfg)]
    L.lpl_set_jcp_mode.argtypes = [vp, C.c_int]
    L.lpl_ring_partition.argtypes = [vp, vp, sz, u32, vp]
    L.lpl_dror_filter.argtypes = [vp, vp, sz, u32, vp]
    L.lpl_segment.argtypes = [vp, vp, sz, i32, u32, vp, vp]
    L.lpl_cluster.argtypes = [vp, vp, sz, u32, vp, C.POINTER(u32)]
    L.lpl_convex_hull.argtypes = [vp, vp, sz, u32, vp, C.POINTER(u32)]
    L.lpl_cluster_hulls.argtypes = [vp, vp, sz, vp, u32, u32, vp, vp, vp, vp]
    L.lpl_bounding_boxes.argtypes = [vp, vp, sz, vp, u32, C.c_int, vp]
    L.lpl_vehicle_match.argtypes = [vp, vp, sz, vp, u32, vp, vp, vp, vp, vp]
    L.lpl_vehicle_match.restype = C.c_int
    L.lpl_pipeline_upload.argtypes = [vp, C.POINTER(Frame), u32]
    L.lpl_pipeline_upload_device.argtypes = [vp, C.POINTER(Frame), u32]
    L.lpl_pipeline_upload_cloud2.argtypes = [vp, C.POINTER(Cloud2Frame), u32]
    L.lpl_pcd_read.argtypes = [C.c_char_p, vp, u32, C.POINTER(u32)]
    L.lpl_pipeline_upload_packed.argtypes = [vp, vp, vp, u32]
    L.lpl_pipeline_upload_packed_xyz.argtypes = [vp, vp, vp, u32]
    L.lpl_pipeline_status.argtypes = [vp, u32, vp]
    L.lpl_pipeline_download_packed.argtypes = [vp, u32, C.POINTER(PackedResult)]
    L.lpl_pipeline_split_clouds.argtypes = [vp, u32, C.POINTER(SplitResult)]
    L.lpl_pipeline_split_clouds.restype = C.c_int
    L.lpl_knn_build.argtypes = [vp, vp, sz, u32]
    L.lpl_knn_build.restype = C.c_int
    L.lpl_knn_token.argtypes = [vp]
    L.lpl_knn_token.restype = C.c_ulonglong
    L.lpl_knn_k_nearest.argtypes = [vp, vp, sz, u32, u32, vp, vp, vp, vp]
    L.lpl_knn_k_nearest.restype = C.c_int
    L.lpl_knn_radius_search.argtypes = [vp, vp, sz, u32, vp, u32, vp, vp, vp]
    L.lpl_knn_radius_search.restype = C.c_int
    L.lpl_glibc_rand_stream.argtypes = [u32, u32, vp]
    L.lpl_glibc_rand_stream.restype = None
    L.lpl_pipeline_run.argtypes = [vp, u32, u32]
    L.lpl_pipeline_use_graph.argtypes = [vp, C.c_int]
    L.lpl_pipeline_use_graph.restype = C.c_int
    L.lpl_device_bytes.argtypes = [vp]
    L.lpl_device_bytes.restype = sz
    L.lpl_pipeline_use_split.argtypes = [vp, u32]
    L.lpl_pipeline_use_split.restype = C.c_int
    L.lpl_pipeline_sync.argtypes = [vp, u32]
    L.lpl_pipeline_want_image.argtypes = [vp, C.c_int]
    L.lpl_pipeline_counts.argtypes = [vp, u32, C.POINTER(FrameResult)]
    L.lpl_pipeline_download.argtypes = [vp, u32, C.POINTER(FrameResult)]
    L.lpl_pipeline_download_batch.argtypes = [vp, u32, C.POINTER(BatchResult)]
    L.lpl_host_alloc.argtypes = [C.POINTER(vp), sz]
    L.lpl_host_alloc.restype = C.c_int
    L.lpl_host_free.argtypes = [vp]
    L.lpl_host_free.restype = None
    L.lpl_profile_enable.argtypes = [vp, C.c_int]
    L.lpl_profile_read.argtypes = [vp, u32, C.POINTER(C.c_char_p), C.POINTER(C.c_float), C.POINTER(u32)]
    L.lpl_timer_start.argtypes = [vp]
    L.lpl_timer_stop_ms.argtypes = [vp, C.POINTER(C.c_float)]
    L.lpl_launch_count.argtypes = [vp, C.c_int]
    L.lpl_launch_count.restype = C.c_uint64
    L.lpl_debug_segment.argtypes = [vp, u32, vp, vp, vp, vp]
    L.lpl_debug_cluster.argtypes = [vp, u32, vp]
    L.lpl_debug_dror.argtypes = [vp, u32, C.POINTER(u32)]
    L.lpl_debug_hulls.argtypes = [vp, u32, vp]
    L.lpl_debug_dror.restype = C.c_int
    L.lpl_stream.argtypes = [vp]
    L.lpl_stream.restype = vp
    for name in ("lpl_segmenter_config", "lpl_dror_config", "lpl_cluster_config", "lpl_set_jcp_mode",
                 "lpl_ring_partition", "lpl_dror_filter", "lpl_segment", "lpl_cluster",
                 "lpl_convex_hull", "lpl_cluster_hulls", "lpl_pipeline_upload",
                 "lpl_pipeline_upload_device", "lpl_pipeline_run", "lpl_pipeline_use_graph", "lpl_pipeline_sync",
                 "lpl_pipeline_want_image", "lpl_pipeline_counts", "lpl_pipeline_download",
                 "lpl_pipeline_download_batch", "lpl_pipeline_upload_packed", "lpl_pipeline_upload_packed_xyz",
                 "lpl_pipeline_upload_cloud2", "lpl_pipeline_status", "lpl_pipeline_download_packed",
    "lpl_pipeline_split_clouds", "lpl_glibc_rand_stream",
    "lpl_knn_build", "lpl_knn_token", "lpl_knn_k_nearest", "lpl_knn_radius_search",
                 "lpl_profile_enable", "lpl_profile_read",
                 "lpl_timer_start", "lpl_timer_stop_ms", "lpl_debug_segment", "lpl_debug_cluster"):
        getattr(L, name).restype = C.c_int
    if path is None:
        _lib = L
    return L


class PinnedBuffer:
    """cudaMallocHost block exposed as a numpy array (frame staging / result planes)."""

    def __init__(self, shape, dtype):
        self.lib = load_library()
        self.shape = tuple(int(v) for v in np.atleast_1d(shape))
        self.dtype = np.dtype(dtype)
        nbytes = max(int(np.prod(self.shape)) * self.dtype.itemsize, 1)
        p = C.c_void_p()
        rc = self.lib.lpl_host_alloc(C.byref(p), nbytes)
        if rc != 0 or not p.value:
            raise LplError(rc, f"cudaMallocHost({nbytes}) failed")
        self.ptr = p.value
        buf = (C.c_char * nbytes).from_address(self.ptr)
        self.array = np.frombuffer(buf, dtype=self.dtype, count=int(np.prod(self.shape))).reshape(self.shape)

    def close(self):
        if getattr(self, "ptr", None):
            self.array = None
            self.lib.lpl_host_free(self.ptr)
            self.ptr = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class BatchBuffers:
    """Pinned host planes for lpl_pipeline_download_batch (frame-major, `stride` elements apart)."""

    PLANES = (("labels_u8", np.uint8, 1), ("noise", np.uint8, 1), ("ring", np.uint16, 1),
              ("obstacle_index", np.uint32, 1), ("cluster_labels", np.int32, 1),
              ("hull_offsets", np.uint32, 1), ("hull_indices", np.uint32, 1),
              ("hull_xy", np.float32, 2), ("zminmax", np.float32, 2), ("boxes", BBOX_DTYPE, 1))

    def __init__(self, max_frames: int, stride: int, want=("labels_u8", "obstacle_index", "cluster_labels",
                                                           "hull_offsets", "hull_xy", "zminmax")):
        self.max_frames, self.stride = max_frames, stride
        self.counts = PinnedBuffer((5, max_frames), np.uint32)
        self.planes = {}
        for name, dt, width in self.PLANES:
            if name in want:
                shape = (max_frames, stride) if width == 1 else (max_frames, stride, width)
                self.planes[name] = PinnedBuffer(shape, dt)

    def bytes_for(self, counts: np.ndarray) -> int:
        """Bytes that cross PCIe for a batch with these per-frame counts ([5][nf])."""
        nf = counts.shape[1]
        mx = counts.max(axis=1)
        width = dict(labels_u8=mx[0], noise=mx[0], ring=2 * mx[0], obstacle_index=4 * mx[2],
                     cluster_labels=4 * mx[2], hull_offsets=4 * (mx[3] + 1), hull_indices=4 * mx[4],
                     hull_xy=8 * mx[4], zminmax=8 * mx[3], boxes=80 * mx[3])
        return int(sum(int(width[k]) for k in self.planes) * nf + counts.size * 4)

    def close(self):
        self.counts.close()
        for b in self.planes.values():
            b.close()


class PackedBuffers:
    """Pinned host buffer for lpl_pipeline_download_packed: the selected planes of a batch back to back, every
    frame's occupied part only. `frame(name, f)` gives frame f of a plane as a numpy view."""

    def __init__(self, max_frames: int, nbytes: int, want=("labels_u8", "cluster_labels", "hull_offsets", "hull_xy", "zminmax")):
        self.max_frames = max_frames
        self.want = tuple(want)
        self.planes_mask = sum(PLANE_BIT[n] for n in self.want)
        self.counts = PinnedBuffer((5, max_frames), np.uint32)
        self.buf = PinnedBuffer((int(nbytes),), np.uint8)
        self.offset = {}
        self.bytes_used = 0
        self.nf = 0
        self._starts = {}

    def _set(self, nf: int, res: "PackedResult"):
        self.nf = nf
        self.bytes_used = int(res.bytes_used)
        self.offset = {name: int(res.offset[i]) for i, (name, _, _, _) in enumerate(PLANES) if name in self.want}
        cn = self.counts.array.reshape(-1)[: 5 * nf].reshape(5, nf).astype(np.int64)
        per = {0: cn[0], 2: cn[2], 3: cn[3], 4: cn[4], 5: cn[3] + 1}
        self._starts = {k: np.concatenate([[0], np.cumsum(v)]) for k, v in per.items()}
        return cn

    def frame(self, name: str, f: int) -> np.ndarray:
        i = [p[0] for p in PLANES].index(name)
        _, dt, width, row = PLANES[i]
        key = 5 if name == "hull_offsets" else row
        st = self._starts[key]
        esz = np.dtype(dt).itemsize * width
        a = self.offset[name] + int(st[f]) * esz
        n = int(st[f + 1] - st[f])
        v = self.buf.array[a:a + n * esz].view(dt)
        return v.reshape(n, width) if width > 1 else v

    def close(self):
        self.counts.close()
        self.buf.close()


def _points_arg(pts):
    """(pointer, stride_bytes, n, keepalive) for an (n, >=3) float32 array."""
    a = np.ascontiguousarray(pts, dtype=np.float32)
    if a.ndim != 2 or a.shape[1] < 3:
        raise ValueError("points must be (n, >=3) float32")
    return a.ctypes.data, a.shape[1] * 4, a.shape[0], a


class Context:
    """Owns one lpl_ctx (one CUDA stream + all device scratch)."""

    def __init__(self, device: int = 0, max_points: int = 131072, max_frames: int = 1,
                 image_height: int = 64, image_width: int = 2048):
        self.lib = load_library()
        h = C.c_void_p()
        rc = self.lib.lpl_create(C.byref(h), device, max_points, max_frames, image_height, image_width)
        if rc != 0:
            msg = {LPL_ERR_NO_DEVICE: "no CUDA device (the hot path has no CPU fallback)",
                   LPL_ERR_CAPACITY: "device allocation failed"}.get(rc, "lpl_create failed")
            raise LplError(rc, msg)
        self.h = h
        self.max_points = max_points
        self.max_frames = max_frames
        self.H, self.W = image_height, image_width

    def close(self):
        if getattr(self, "h", None):
            self.lib.lpl_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _chk(self, rc: int):
        if rc != 0:
            raise LplError(rc, self.lib.lpl_last_error(self.h).decode(errors="replace"))

    # ---- configuration
    def segmenter_default_cfg(self) -> SegmenterCfg:
        c = SegmenterCfg()
        self.lib.lpl_segmenter_default_cfg(C.byref(c))
        return c

    def segmenter_config(self, cfg: SegmenterCfg):
        self._chk(self.lib.lpl_segmenter_config(self.h, C.byref(cfg)))

    def dror_config(self, mult=0.02, min_radius=0.1, min_neighbours=4):
        self._chk(self.lib.lpl_dror_config(self.h, C.byref(DrorCfg(mult, min_radius, min_neighbours))))

    def cluster_config(self, range_m=0.4, az_deg=1.0, el_deg=1.5, min_size=3):
        self._chk(self.lib.lpl_cluster_config(self.h, C.byref(ClusterCfg(range_m, az_deg, el_deg, min_size))))

    def set_jcp_mode(self, mode: int):
        self._chk(self.lib.lpl_set_jcp_mode(self.h, mode))

    # ---- single-frame entry points
    def ring_partition(self, pts) -> np.ndarray:
        p, stride, n, keep = _points_arg(pts)
        out = np.zeros(n, np.uint16)
        self._chk(self.lib.lpl_ring_partition(self.h, p, stride, n, out.ctypes.data))
        return out

    def dror_filter(self, pts) -> np.ndarray:
        p, stride, n, keep = _points_arg(pts)
        out = np.zeros(n, np.uint8)
        self._chk(self.lib.lpl_dror_filter(self.h, p, stride, n, out.ctypes.data))
        return out

    def segment(self, pts, ring=None, want_image=False):
        """pts (n, >=3) float32; ring (n,) uint16 or None (ring-less point type)."""
        a = np.ascontiguousarray(pts, dtype=np.float32)
        n = a.shape[0]
        if ring is not None:
            # build PointXYZIR-like records: x y z pad | intensity ring(u16) pad  (32 bytes)
            rec = np.zeros((n, 8), np.float32)
            rec[:, :3] = a[:, :3]
            rec.view(np.uint16).reshape(n, 16)[:, 10] = np.asarray(ring, np.uint16)
            p, stride, roff, keep = rec.ctypes.data, 32, 20, rec
        else:
            p, stride, roff, keep = a.ctypes.data, a.shape[1] * 4, -1, a
        labels = np.zeros(n, np.uint32)
        img = np.zeros((self.H, self.W, 3), np.uint8) if want_image else None
        self._chk(self.lib.lpl_segment(self.h, p, stride, roff, n, labels.ctypes.data,
                                       img.ctypes.data if want_image else None))
        return (labels, img) if want_image else labels

    def cluster(self, pts):
        p, stride, n, keep = _points_arg(pts)
        out = np.full(n, -1, np.int32)
        k = C.c_uint32(0)
        self._chk(self.lib.lpl_cluster(self.h, p, stride, n, out.ctypes.data, C.byref(k)))
        return out, k.value

    def convex_hull(self, xy) -> np.ndarray:
        a = np.ascontiguousarray(xy, dtype=np.float64)
        n = a.shape[0]
        idx = np.zeros(max(n, 1), np.int32)
        cnt = C.c_uint32(0)
        self._chk(self.lib.lpl_convex_hull(self.h, a.ctypes.data, a.shape[1] * 8, n, idx.ctypes.data,
                                           C.byref(cnt)))
        return idx[: cnt.value].copy()

    def bounding_boxes(self, hull_xy, offsets, method: int = BOX_ROTATING_CALIPERS) -> np.ndarray:
        """Oriented boxes of len(offsets) - 1 convex hulls given back to back in hull_xy ((m, 2) doubles)."""
        a = np.ascontiguousarray(hull_xy, dtype=np.float64).reshape(-1, 2)
        off = np.ascontiguousarray(offsets, dtype=np.uint32)
        K = max(len(off) - 1, 0)
        out = np.zeros(max(K, 1), BBOX_DTYPE)
        self._chk(self.lib.lpl_bounding_boxes(self.h, a.ctypes.data if a.size else None, 16, off.ctypes.data, K,
                                              method, out.ctypes.data))
        return out[:K].copy()

    def vehicle_match(self, hull_xy, offsets, z_min_max, cluster_sizes, boxes):
        """Vehicle class per cluster (-1 none) and the hull's polygon area (lpl_vehicle_match)."""
        a = np.ascontiguousarray(hull_xy, dtype=np.float64).reshape(-1, 2)
        off = np.ascontiguousarray(offsets, dtype=np.uint32)
        K = max(len(off) - 1, 0)
        z = np.ascontiguousarray(z_min_max, np.float64).reshape(-1, 2)
        sz_ = np.ascontiguousarray(cluster_sizes, np.uint32)
        bx = np.ascontiguousarray(boxes, BBOX_DTYPE)
        cls = np.full(max(K, 1), -1, np.int32)
        area = np.zeros(max(K, 1), np.float64)
        self._chk(self.lib.lpl_vehicle_match(self.h, a.ctypes.data if a.size else None, 16, off.ctypes.data, K, z.ctypes.data,
                                             sz_.ctypes.data, bx.ctypes.data, cls.ctypes.data, area.ctypes.data))
        return cls[:K].copy(), area[:K].copy()

    def cluster_hulls(self, pts, labels, num_clusters=None):
        p, stride, n, keep = _points_arg(pts)
        lab = np.ascontiguousarray(labels, np.int32)
        K = int(lab.max()) + 1 if (num_clusters is None and n) else int(num_clusters or 0)
        K = max(K, 0)
        off = np.zeros(K + 1, np.uint32)
        hidx = np.zeros(max(n, 1), np.int32)
        hxy = np.zeros((max(n, 1), 2), np.float32)
        zmm = np.zeros((max(K, 1), 2), np.float32)
        self._chk(self.lib.lpl_cluster_hulls(self.h, p, stride, lab.ctypes.data, n, K, off.ctypes.data,
                                             hidx.ctypes.data, hxy.ctypes.data, zmm.ctypes.data))
        tot = int(off[K]) if K else 0
        return off, hxy[:tot].copy(), hidx[:tot].copy(), zmm[:K].copy()

    # ---- batched pipeline
    def upload(self, frames, rings=None, device=False):
        """frames: list of (n, 4) float32 arrays (host) or list of (device_ptr, n) when device."""
        nf = len(frames)
        arr = (Frame * nf)()
        keep = []
        for i, fr in enumerate(frames):
            if device:
                ptr, n = fr
                arr[i].xyzw, arr[i].n = ptr, n
                arr[i].ring = rings[i] if rings is not None else None
            else:
                a = np.ascontiguousarray(fr, dtype=np.float32)
                assert a.ndim == 2 and a.shape[1] == 4, "pipeline frames are (n, 4) float32"
                keep.append(a)
                arr[i].xyzw, arr[i].n = a.ctypes.data, a.shape[0]
                if rings is not None and rings[i] is not None:
                    r = np.ascontiguousarray(rings[i], np.uint16)
                    keep.append(r)
                    arr[i].ring = r.ctypes.data
        fn = self.lib.lpl_pipeline_upload_device if device else self.lib.lpl_pipeline_upload
        self._chk(fn(self.h, arr, nf))
        self._keep = keep
        return nf

    def upload_packed(self, xyzw, counts) -> int:
        """xyzw: (sum(counts), 4) float32, the frames back to back (ideally pinned); counts: points per frame."""
        a = np.ascontiguousarray(xyzw, dtype=np.float32)
        cn = np.ascontiguousarray(counts, dtype=np.uint32)
        assert a.ndim == 2 and a.shape[1] == 4 and int(cn.sum()) == a.shape[0]
        self._chk(self.lib.lpl_pipeline_upload_packed(self.h, a.ctypes.data if a.size else None, cn.ctypes.data, len(cn)))
        self._keep = [a, cn]
        return len(cn)

    def upload_packed_xyz(self, xyz, counts) -> int:
        """xyz: (sum(counts), 3) float32, 12 bytes per point (the std::array<float, 3> cloud of NoiseRemover::filter)."""
        a = np.ascontiguousarray(xyz, dtype=np.float32)
        cn = np.ascontiguousarray(counts, dtype=np.uint32)
        assert a.ndim == 2 and a.shape[1] == 3 and int(cn.sum()) == a.shape[0]
        self._chk(self.lib.lpl_pipeline_upload_packed_xyz(self.h, a.ctypes.data if a.size else None, cn.ctypes.data, len(cn)))
        self._keep = [a, cn]
        return len(cn)

    def upload_cloud2(self, messages) -> int:
        """messages: dicts with data (uint8 array), width, height, point_step, row_step, x/y/z_offset and
        ring_offset (-1 = none) - the PointCloud2 fields Processor::convert reads."""
        nf = len(messages)
        arr = (Cloud2Frame * nf)()
        keep = []
        for f, m in enumerate(messages):
            data = np.ascontiguousarray(m["data"], np.uint8)
            keep.append(data)
            arr[f] = Cloud2Frame(data.ctypes.data if data.size else None, m["width"], m["height"], m["point_step"],
                                 m["row_step"], m["x_offset"], m["y_offset"], m["z_offset"], m.get("ring_offset", -1))
        self._chk(self.lib.lpl_pipeline_upload_cloud2(self.h, arr, nf))
        self._keep = keep  # the copies are asynchronous: the host arrays must outlive them
        return nf

    def run(self, nf: int, stages: int = STAGE_ALL):
        self._chk(self.lib.lpl_pipeline_run(self.h, nf, stages))

    def use_graph(self, enable: bool):
        self._chk(self.lib.lpl_pipeline_use_graph(self.h, 1 if enable else 0))

    def device_bytes(self) -> int:
        """Device memory held by the context (lpl_device_bytes)."""
        return int(self.lib.lpl_device_bytes(self.h))

    def use_split(self, parts: int):
        """Sub-batches of one run on concurrent streams (lpl_pipeline_use_split)."""
        self._chk(self.lib.lpl_pipeline_use_split(self.h, parts))

    def sync(self, nf: int):
        self._chk(self.lib.lpl_pipeline_sync(self.h, nf))

    def want_image(self, enable: bool):
        self._chk(self.lib.lpl_pipeline_want_image(self.h, 1 if enable else 0))

    def counts(self, f: int) -> FrameResult:
        r = FrameResult()
        self._chk(self.lib.lpl_pipeline_counts(self.h, f, C.byref(r)))
        return r

    def download(self, f: int, want_image=False, want_boxes=False) -> dict:
        r = self.counts(f)
        out = dict(
            noise=np.zeros(r.n, np.uint8), ring=np.zeros(r.n, np.uint16), labels=np.zeros(r.n, np.uint32),
            obstacle_index=np.zeros(r.num_obstacles, np.uint32),
            cluster_labels=np.zeros(r.num_obstacles, np.int32),
            hull_offsets=np.zeros(r.num_clusters + 1, np.uint32),
            hull_indices=np.zeros(r.num_hull_vertices, np.uint32),
            hull_xy=np.zeros((r.num_hull_vertices, 2), np.float32),
            zminmax=np.zeros((r.num_clusters, 2), np.float32),
        )
        if want_image:
            out["bgr"] = np.zeros((self.H, self.W, 3), np.uint8)
        if want_boxes:
            out["boxes"] = np.zeros(r.num_clusters, BBOX_DTYPE)
        for k, v in out.items():
            setattr(r, k, v.ctypes.data if v.size else None)
        self._chk(self.lib.lpl_pipeline_download(self.h, f, C.byref(r)))
        out.update(n=r.n, num_valid=r.num_valid, num_obstacles=r.num_obstacles,
                   num_clusters=r.num_clusters, num_hull_vertices=r.num_hull_vertices)
        return out

    def download_batch(self, nf: int, bufs: BatchBuffers) -> np.ndarray:
        """One strided D2H copy per plane for the whole batch; returns counts[5][nf] (a view)."""
        r = BatchResult()
        r.counts = bufs.counts.ptr
        r.stride = bufs.stride
        for name, b in bufs.planes.items():
            setattr(r, name, b.ptr)
        # counts are laid out [5][nf] for this call's nf
        self._chk(self.lib.lpl_pipeline_download_batch(self.h, nf, C.byref(r)))
        return bufs.counts.array.reshape(-1)[: 5 * nf].reshape(5, nf)

    def download_packed(self, nf: int, bufs: PackedBuffers) -> np.ndarray:
        """ONE D2H transfer of the occupied bytes of the selected planes; returns counts[5][nf]."""
        r = PackedResult()
        r.counts = bufs.counts.ptr
        r.buffer = bufs.buf.ptr
        r.buffer_bytes = bufs.buf.array.nbytes
        r.planes = bufs.planes_mask
        self._chk(self.lib.lpl_pipeline_download_packed(self.h, nf, C.byref(r)))
        return bufs._set(nf, r)

    def split_clouds(self, nf: int, stride: int, markers: bool = True, colors=None) -> dict:
        """Label split + clustered cloud (+ marker lines) of the last batch (lpl_pipeline_split_clouds). Returns per
        frame lists of RGB_DTYPE record arrays and (n, 3) float64 marker vertices. colors: (nf, K, 3) uint8 or None
        (the node's rand() stream)."""
        counts = np.zeros((5, nf), np.uint32)
        planes = {k: np.zeros((nf, stride), RGB_DTYPE) for k in ("ground", "obstacle", "unsegmented", "clustered")}
        mk_stride = max(stride // 2, 1)
        mk = np.zeros((nf, mk_stride, 3), np.float64) if markers else None
        r = SplitResult()
        r.counts = counts.ctypes.data
        r.stride = stride
        for k, v in planes.items():
            setattr(r, k, v.ctypes.data)
        r.marker_stride = mk_stride
        r.marker_points = mk.ctypes.data if markers else None
        keep = None
        if colors is not None:
            keep = np.ascontiguousarray(colors, np.uint8)
            r.cluster_colors = keep.ctypes.data
            r.colors_stride = keep.shape[1]
        self._chk(self.lib.lpl_pipeline_split_clouds(self.h, nf, C.byref(r)))
        out = {k: [planes[k][f, : counts[i, f]] for f in range(nf)] for i, k in enumerate(("ground", "obstacle", "unsegmented", "clustered"))}
        out["markers"] = [mk[f, : counts[4, f]] for f in range(nf)] if markers else None
        out["counts"] = counts
        return out

    # ---- general neighbour queries (KDTree<float, 3> of the reference)
    def knn_build(self, pts):
        p, stride, n, keep = _points_arg(pts)
        self._chk(self.lib.lpl_knn_build(self.h, p, stride, n))

    def knn_token(self) -> int:
        return int(self.lib.lpl_knn_token(self.h))

    def k_nearest(self, queries, k: int, radius_sqr=None):
        """-> (idx [m][k] uint32, dist [m][k] float32 squared distances, count [m])."""
        p, stride, m, keep = _points_arg(queries)
        idx = np.zeros((m, k), np.uint32)
        dist = np.zeros((m, k), np.float32)
        cnt = np.zeros(m, np.uint32)
        r = None if radius_sqr is None else np.ascontiguousarray(radius_sqr, np.float32)
        self._chk(self.lib.lpl_knn_k_nearest(self.h, p, stride, m, k, None if r is None else r.ctypes.data, idx.ctypes.data,
                                             dist.ctypes.data, cnt.ctypes.data))
        return idx, dist, cnt

    def radius_search(self, queries, radius_sqr, max_per_query: int):
        p, stride, m, keep = _points_arg(queries)
        idx = np.zeros((m, max_per_query), np.uint32)
        dist = np.zeros((m, max_per_query), np.float32)
        cnt = np.zeros(m, np.uint32)
        r = np.ascontiguousarray(radius_sqr, np.float32)
        self._chk(self.lib.lpl_knn_radius_search(self.h, p, stride, m, r.ctypes.data, max_per_query, idx.ctypes.data,
                                                 dist.ctypes.data, cnt.ctypes.data))
        return idx, dist, cnt

    def status(self, nf: int) -> np.ndarray:
        """Per-frame capacity flags of the last run (0 = good)."""
        st = np.zeros(nf, np.uint32)
        self._chk(self.lib.lpl_pipeline_status(self.h, nf, st.ctypes.data))
        return st

    # ---- measurement / debugging
    def profile(self, enable: bool):
        self._chk(self.lib.lpl_profile_enable(self.h, 1 if enable else 0))

    def profile_read(self) -> list:
        """[(kernel name, ms)] of the last run, in launch order."""
        names = (C.c_char_p * 96)()
        ms = (C.c_float * 96)()
        cnt = C.c_uint32(0)
        self._chk(self.lib.lpl_profile_read(self.h, 96, names, ms, C.byref(cnt)))
        return [(names[i].decode(), float(ms[i])) for i in range(cnt.value)]

    def timer_start(self):
        self._chk(self.lib.lpl_timer_start(self.h))

    def timer_stop_ms(self) -> float:
        ms = C.c_float(0)
        self._chk(self.lib.lpl_timer_stop_ms(self.h, C.byref(ms)))
        return ms.value

    def launch_count(self, reset=False) -> int:
        return int(self.lib.lpl_launch_count(self.h, 1 if reset else 0))

    def debug_segment(self, f: int = 0) -> dict:
        cnt = np.zeros(8, np.uint32)
        self._chk(self.lib.lpl_debug_segment(self.h, f, None, None, None, cnt.ctypes.data))
        slices, rings = int(cnt[5]), int(cnt[6])
        elev = np.zeros(slices * rings, np.float32)
        plane = np.zeros(4, np.float32)
        best = C.c_uint32(0)
        self._chk(self.lib.lpl_debug_segment(self.h, f, elev.ctypes.data, plane.ctypes.data, C.byref(best),
                                             cnt.ctypes.data))
        return dict(elevation=elev.reshape(slices, rings), plane=plane, best_inliers=best.value,
                    n_binned=int(cnt[0]), n_candidates=int(cnt[1]), n_queued=int(cnt[2]),
                    rounds=int(cnt[3]), border_rows=int(cnt[4]), status=int(cnt[7]))

    def debug_counters(self, f: int = 0) -> dict:
        """Cheap per-frame counters of the last run (no plane downloads)."""
        cnt = np.zeros(8, np.uint32)
        self._chk(self.lib.lpl_debug_segment(self.h, f, None, None, None, cnt.ctypes.data))
        nu = C.c_uint32(0)
        self._chk(self.lib.lpl_debug_dror(self.h, f, C.byref(nu)))
        return dict(n_binned=int(cnt[0]), n_candidates=int(cnt[1]), n_queued=int(cnt[2]), rounds=int(cnt[3]),
                    border_rows=int(cnt[4]), cells=int(cnt[5]) * int(cnt[6]), status=int(cnt[7]),
                    n_unresolved=int(nu.value))

    def debug_hulls(self, f: int = 0) -> dict:
        v = np.zeros(2, np.uint32)
        self._chk(self.lib.lpl_debug_hulls(self.h, f, v.ctypes.data))
        return dict(n_hull_sort=int(v[0]), n_voxels=int(v[1]))

    def debug_cluster(self, f: int = 0) -> np.ndarray:
        dims = np.zeros(3, np.int32)
        self._chk(self.lib.lpl_debug_cluster(self.h, f, dims.ctypes.data))
        return dims
