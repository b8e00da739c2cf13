"""TEST INFRASTRUCTURE ONLY — ctypes loaders for the two CPU checkers.

* ``RefOracle``  : oracle/_ref/libref_oracle.so = the reference's own segmenter.cpp /
  clusterer.cpp / noise_remover.cpp compiled unmodified (oracle/Makefile, oracle/ref_capi.cpp).
* ``PortOracle`` : oracle/liboracle_port.so = oracle/port.cpp, our CPU restatement.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
import this module. The product path (lidar_processing_v2_b200) never does.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
REF_SO = os.path.join(HERE, "_ref", "libref_oracle.so")
PORT_SO = os.path.join(HERE, "liboracle_port.so")

UNKNOWN, GROUND, OBSTACLE = 0, 1, 2


def build(verbose: bool = False) -> None:
    """Compile the checkers (port always; _ref only where /root/reference exists)."""
    res = subprocess.run(["make", "-C", HERE, "all"], capture_output=True, text=True)
    if verbose or res.returncode != 0:
        print(res.stdout, res.stderr)
    res.check_returncode()


class SegCfg(C.Structure):
    """POD mirror of SegmenterConfiguration (segmenter.hpp:87-112)."""

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


def default_seg_cfg(**over) -> SegCfg:
    c = SegCfg(2.0, -24.8, 2048, 64, 0, 2.0, 1.0, 0.2, 0.2, 2.0, 100.0, 1.73, 1.0, 5.0, -3.0, 4.0)
    for k, v in over.items():
        setattr(c, k, v)
    return c


# clustering configuration of the processor node (processor.param.yaml:31-35)
NODE_CLUSTER_CFG = dict(range_m=0.4, az_deg=1.0, el_deg=3.0, min_size=3)


def _f32(a):
    a = np.ascontiguousarray(a, dtype=np.float32)
    assert a.ndim == 2 and a.shape[1] >= 3
    return a


def _p(a, t):
    return a.ctypes.data_as(C.POINTER(t))


class RefOracle:
    """The unmodified reference library behind oracle/ref_capi.cpp."""

    def __init__(self, path: str = REF_SO):
        if not os.path.exists(path):
            raise FileNotFoundError(path)
        L = self.lib = C.CDLL(path)
        L.ref_last_error.restype = C.c_char_p
        L.ref_segmenter_create.restype = C.c_void_p
        L.ref_segmenter_destroy.argtypes = [C.c_void_p]
        L.ref_segmenter_config.argtypes = [C.c_void_p, C.POINTER(SegCfg)]
        L.ref_segment.argtypes = [C.c_void_p, C.c_void_p, C.c_int32, C.c_void_p, C.c_uint32,
                                  C.c_void_p, C.c_void_p]
        L.ref_segment.restype = C.c_int
        L.ref_segment_timed.argtypes = [C.c_void_p, C.c_void_p, C.c_int32, C.c_void_p, C.c_uint32,
                                        C.c_int32]
        L.ref_segment_timed.restype = C.c_double
        L.ref_segment_grid_dims.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p]
        L.ref_segment_intermediates.argtypes = [C.c_void_p] + [C.c_void_p] * 4
        L.ref_clusterer_create.restype = C.c_void_p
        L.ref_clusterer_destroy.argtypes = [C.c_void_p]
        L.ref_clusterer_config.argtypes = [C.c_void_p, C.c_float, C.c_float, C.c_float, C.c_uint32]
        L.ref_clusterer_reserve.argtypes = [C.c_void_p, C.c_uint32, C.c_uint32]
        L.ref_cluster.argtypes = [C.c_void_p, C.c_void_p, C.c_int32, C.c_uint32, C.c_void_p,
                                  C.c_void_p]
        L.ref_cluster.restype = C.c_int
        L.ref_cluster_timed.argtypes = [C.c_void_p, C.c_void_p, C.c_int32, C.c_uint32, C.c_int32]
        L.ref_cluster_timed.restype = C.c_double
        L.ref_dror.argtypes = [C.c_void_p, C.c_int32, C.c_uint32, C.c_float, C.c_float, C.c_uint32,
                               C.c_int32, C.c_void_p]
        L.ref_dror.restype = C.c_int
        L.ref_dror_timed.argtypes = [C.c_void_p, C.c_int32, C.c_uint32, C.c_int32]
        L.ref_dror_timed.restype = C.c_double
        L.ref_shim_dilate5x5.argtypes = [C.c_void_p, C.c_int32, C.c_int32, C.c_void_p]
        if hasattr(L, "ref_kdtree_query"):
            L.ref_kdtree_query.argtypes = [C.c_void_p, C.c_uint32, C.c_void_p, C.c_uint32, C.c_uint32, C.c_void_p, C.c_int32,
                                           C.c_void_p, C.c_void_p, C.c_void_p]
            L.ref_kdtree_query.restype = C.c_int
        self.has_polygonizer = hasattr(L, "ref_convex_hull")
        if self.has_polygonizer:
            L.ref_convex_hull.argtypes = [C.c_void_p, C.c_int32, C.c_void_p]
            L.ref_convex_hull.restype = C.c_int32
            L.ref_antipodal_pairs.argtypes = [C.c_void_p, C.c_int32, C.c_void_p]
            L.ref_antipodal_pairs.restype = C.c_int32
            L.ref_bounding_box.argtypes = [C.c_void_p, C.c_int32, C.c_int32, C.c_void_p]
        self._seg = C.c_void_p(L.ref_segmenter_create())
        self._clu = C.c_void_p(L.ref_clusterer_create())
        self.seg_cfg = default_seg_cfg()
        L.ref_segmenter_config(self._seg, C.byref(self.seg_cfg))
        self.cluster_config(**NODE_CLUSTER_CFG)

    def __del__(self):
        try:
            self.lib.ref_segmenter_destroy(self._seg)
            self.lib.ref_clusterer_destroy(self._clu)
        except Exception:
            pass


    # -- KDTree<float, 3> (kdtree.hpp:216-337) ----------------------------------------
    def kdtree_query(self, pts, queries, k, radius_sqr=None):
        """The reference's own radius_search (sorted by distance), one query at a time; its k_nearest template does not
        compile (kdtree.hpp:246), so radius_sqr is required. Returns (idx [m][k], dist [m][k], count [m])."""
        p = np.ascontiguousarray(np.asarray(pts, np.float32)[:, :3])
        q = np.ascontiguousarray(np.asarray(queries, np.float32)[:, :3])
        m = q.shape[0]
        idx = np.zeros((m, k), np.uint32)
        dist = np.zeros((m, k), np.float32)
        cnt = np.zeros(m, np.uint32)
        r = None if radius_sqr is None else np.ascontiguousarray(radius_sqr, np.float32)
        rc = self.lib.ref_kdtree_query(p.ctypes.data, p.shape[0], q.ctypes.data, m, k, None if r is None else r.ctypes.data,
                                       0 if r is None else 1, idx.ctypes.data, dist.ctypes.data, cnt.ctypes.data)
        if rc != 0:
            raise RuntimeError(self.lib.ref_last_error().decode())
        return idx, dist, cnt

    # -- polygonizer (convex hull, antipodal pairs, oriented boxes) ---------------------
    def convex_hull(self, xy):
        xy = np.ascontiguousarray(xy, np.float64)
        n = xy.shape[0]
        idx = np.zeros(max(n, 1), np.int32)
        k = self.lib.ref_convex_hull(xy.ctypes.data, n, idx.ctypes.data)
        return idx[:k].copy()

    def antipodal_pairs(self, hull_xy):
        hull_xy = np.ascontiguousarray(hull_xy, np.float64)
        n = hull_xy.shape[0]
        out = np.zeros((3 * n + 4, 2), np.int32)
        k = self.lib.ref_antipodal_pairs(hull_xy.ctypes.data, n, out.ctypes.data)
        return out[:k].copy()

    def bounding_box(self, hull_xy, method=0):
        """[11] = 4 corners (x, y), area, angle_rad, is_valid; method 0 rotating calipers, 1 PCA."""
        hull_xy = np.ascontiguousarray(hull_xy, np.float64)
        out = np.zeros(11, np.float64)
        self.lib.ref_bounding_box(hull_xy.ctypes.data, hull_xy.shape[0], method, out.ctypes.data)
        return out

    # -- segmentation -----------------------------------------------------------------
    def segment_config(self, cfg: SegCfg):
        self.seg_cfg = cfg
        self.lib.ref_segmenter_config(self._seg, C.byref(cfg))

    def segment(self, pts, ring=None, want_image=False):
        pts = _f32(pts)
        n = pts.shape[0]
        labels = np.zeros(n, np.uint32)
        H, W = self.seg_cfg.image_height, self.seg_cfg.image_width
        img = np.zeros((H, W, 3), np.uint8) if want_image else None
        rp = None
        if ring is not None:
            ring = np.ascontiguousarray(ring, np.uint16)
            rp = ring.ctypes.data
        rc = self.lib.ref_segment(self._seg, pts.ctypes.data, pts.shape[1], rp, n,
                                  labels.ctypes.data, img.ctypes.data if want_image else None)
        if rc != 0:
            raise RuntimeError(self.lib.ref_last_error().decode())
        return (labels, img) if want_image else labels

    def segment_timed(self, pts, ring, reps=1) -> float:
        pts = _f32(pts)
        ring = np.ascontiguousarray(ring, np.uint16)
        return self.lib.ref_segment_timed(self._seg, pts.ctypes.data, pts.shape[1],
                                          ring.ctypes.data, pts.shape[0], reps)

    def segment_intermediates(self):
        s, r = C.c_int32(), C.c_int32()
        self.lib.ref_segment_grid_dims(self._seg, C.byref(s), C.byref(r))
        H, W = self.seg_cfg.image_height, self.seg_cfg.image_width
        elev = np.zeros(s.value * r.value, np.float32)
        cmap = np.zeros(H * W, np.int32)
        depth = np.zeros(H * W, np.float32)
        ncand = C.c_uint32()
        self.lib.ref_segment_intermediates(self._seg, elev.ctypes.data, cmap.ctypes.data,
                                           depth.ctypes.data, C.byref(ncand))
        return dict(slices=s.value, rings=r.value, elevation=elev.reshape(s.value, r.value),
                    cloud_map=cmap.reshape(H, W), depth=depth.reshape(H, W),
                    ransac_candidates=ncand.value)

    # -- clustering -------------------------------------------------------------------
    def cluster_config(self, range_m=0.4, az_deg=1.0, el_deg=1.5, min_size=3):
        self.lib.ref_clusterer_config(self._clu, range_m, az_deg, el_deg, min_size)

    def cluster_reserve(self, buckets, elements):
        self.lib.ref_clusterer_reserve(self._clu, buckets, elements)

    def cluster(self, pts, want_dims=False):
        pts = _f32(pts)
        n = pts.shape[0]
        labels = np.full(n, -1, np.int32)
        dims = np.zeros(3, np.int32)
        rc = self.lib.ref_cluster(self._clu, pts.ctypes.data, pts.shape[1], n, labels.ctypes.data,
                                  dims.ctypes.data)
        if rc != 0:
            raise OverflowError(self.lib.ref_last_error().decode())
        return (labels, dims) if want_dims else labels

    def cluster_timed(self, pts, reps=1) -> float:
        pts = _f32(pts)
        return self.lib.ref_cluster_timed(self._clu, pts.ctypes.data, pts.shape[1], pts.shape[0],
                                          reps)

    # -- DROR -------------------------------------------------------------------------
    def dror(self, pts, mode="exact", mult=0.02, min_radius=0.1, min_neighbours=4):
        pts = _f32(pts)
        n = pts.shape[0]
        labels = np.zeros(n, np.uint8)
        rc = self.lib.ref_dror(pts.ctypes.data, pts.shape[1], n, mult, min_radius, min_neighbours,
                               0 if mode == "as_is" else 1, labels.ctypes.data)
        if rc != 0:
            raise RuntimeError(self.lib.ref_last_error().decode())
        return labels

    def dror_timed(self, pts, reps=1) -> float:
        pts = _f32(pts)
        return self.lib.ref_dror_timed(pts.ctypes.data, pts.shape[1], pts.shape[0], reps)

    def shim_dilate(self, img):
        img = np.ascontiguousarray(img, np.uint8)
        out = np.zeros_like(img)
        self.lib.ref_shim_dilate5x5(img.ctypes.data, img.shape[0], img.shape[1], out.ctypes.data)
        return out


def have_ref() -> bool:
    return os.path.exists(REF_SO)


class PortSegDebug(C.Structure):
    _fields_ = [
        ("elevation", C.c_void_p),
        ("cloud_map", C.c_void_p),
        ("pre_jcp_code", C.c_void_p),
        ("plane", C.c_float * 4),
        ("best_inliers", C.c_uint32),
        ("n_candidates", C.c_uint32),
        ("n_binned", C.c_uint32),
        ("n_queued", C.c_uint32),
        ("n_undecided", C.c_uint32),
        ("rounds", C.c_uint32),
        ("max_cell", C.c_uint32),
        ("n_nonempty_cells", C.c_uint32),
        ("slices", C.c_int32),
        ("rings", C.c_int32),
    ]


JCP_AS_IS, JCP_CLEAN, JCP_AS_IS_DATAFLOW, JCP_CLEAN_DATAFLOW = 0, 1, 2, 3


class PortOracle:
    """oracle/port.cpp — our CPU restatement of the reference algorithms."""

    def __init__(self, path: str = PORT_SO):
        if not os.path.exists(path):
            build()
        L = self.lib = C.CDLL(path)
        L.port_ring_partition.argtypes = [C.c_void_p, C.c_int32, C.c_uint32, C.c_void_p]
        L.port_dror.argtypes = [C.c_void_p, C.c_int32, C.c_uint32, C.c_float, C.c_float,
                                C.c_uint32, C.c_void_p]
        L.port_segment.argtypes = [C.POINTER(SegCfg), C.c_void_p, C.c_int32, C.c_void_p,
                                   C.c_uint32, C.c_int32, C.c_void_p, C.c_void_p,
                                   C.POINTER(PortSegDebug)]
        L.port_segment.restype = C.c_int
        L.port_cluster.argtypes = [C.c_void_p, C.c_int32, C.c_uint32, C.c_float, C.c_float,
                                   C.c_float, C.c_uint32, C.c_void_p, C.c_void_p, C.c_void_p]
        L.port_cluster.restype = C.c_int
        L.port_convex_hull.argtypes = [C.c_void_p, C.c_int32, C.c_void_p]
        L.port_convex_hull.restype = C.c_int32
        L.port_antipodal_pairs.argtypes = [C.c_void_p, C.c_int32, C.c_void_p]
        L.port_antipodal_pairs.restype = C.c_int32
        L.port_bounding_box.argtypes = [C.c_void_p, C.c_int32, C.c_int32, C.c_void_p]
        L.port_cluster_hulls.argtypes = [C.c_void_p, C.c_int32, C.c_uint32, C.c_void_p, C.c_int32,
                                         C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
        L.port_cluster_hulls.restype = C.c_int32
        L.port_rng_draws.argtypes = [C.c_uint32, C.c_uint32, C.c_void_p]
        L.port_std_rng_draws.argtypes = [C.c_uint32, C.c_uint32, C.c_void_p]
        self.seg_cfg = default_seg_cfg()
        self.cluster_cfg = dict(NODE_CLUSTER_CFG)

    def ring_partition(self, pts):
        pts = _f32(pts)
        ring = np.zeros(pts.shape[0], np.uint16)
        self.lib.port_ring_partition(pts.ctypes.data, pts.shape[1], pts.shape[0], ring.ctypes.data)
        return ring

    def dror(self, pts, mult=0.02, min_radius=0.1, min_neighbours=4):
        pts = _f32(pts)
        labels = np.zeros(pts.shape[0], np.uint8)
        self.lib.port_dror(pts.ctypes.data, pts.shape[1], pts.shape[0], mult, min_radius,
                           min_neighbours, labels.ctypes.data)
        return labels

    def segment_config(self, cfg: SegCfg):
        self.seg_cfg = cfg

    def segment(self, pts, ring=None, jcp_mode=JCP_AS_IS, want_image=False, want_debug=False):
        pts = _f32(pts)
        n = pts.shape[0]
        cfg = self.seg_cfg
        H, W = cfg.image_height, cfg.image_width
        labels = np.zeros(n, np.uint32)
        img = np.zeros((H, W, 3), np.uint8) if want_image else None
        rp = None
        if ring is not None:
            ring = np.ascontiguousarray(ring, np.uint16)
            rp = ring.ctypes.data
        dbg = PortSegDebug()
        keep = {}
        if want_debug:
            slices = int(np.float32(2.0 * np.pi) / (np.float32(cfg.grid_slice_resolution_deg) *
                                                     np.float32(np.pi / 180.0)))
            rings = int(np.float32(cfg.max_distance_m) / np.float32(cfg.grid_radial_spacing_m))
            keep["elevation"] = np.zeros((slices + 1) * (rings + 1), np.float32)
            keep["cloud_map"] = np.zeros(H * W, np.int32)
            keep["pre_jcp_code"] = np.zeros(H * W, np.uint8)
            dbg.elevation = keep["elevation"].ctypes.data
            dbg.cloud_map = keep["cloud_map"].ctypes.data
            dbg.pre_jcp_code = keep["pre_jcp_code"].ctypes.data
        self.lib.port_segment(C.byref(cfg), pts.ctypes.data, pts.shape[1], rp, n, jcp_mode,
                              labels.ctypes.data, img.ctypes.data if want_image else None,
                              C.byref(dbg))
        out = [labels]
        if want_image:
            out.append(img)
        if want_debug:
            d = {k: getattr(dbg, k) for k in ("best_inliers", "n_candidates", "n_binned",
                                              "n_queued", "n_undecided", "rounds", "max_cell",
                                              "n_nonempty_cells", "slices", "rings")}
            d["plane"] = np.array(list(dbg.plane), np.float32)
            d["elevation"] = keep["elevation"][: d["slices"] * d["rings"]].reshape(d["slices"],
                                                                                  d["rings"])
            d["cloud_map"] = keep["cloud_map"].reshape(H, W)
            d["pre_jcp_code"] = keep["pre_jcp_code"].reshape(H, W)
            out.append(d)
        return out[0] if len(out) == 1 else tuple(out)

    def cluster(self, pts, want_dims=False, **over):
        pts = _f32(pts)
        cfg = dict(self.cluster_cfg)
        cfg.update(over)
        n = pts.shape[0]
        labels = np.full(n, -1, np.int32)
        dims = np.zeros(3, np.int32)
        nvox = C.c_uint32(0)
        k = self.lib.port_cluster(pts.ctypes.data, pts.shape[1], n, cfg["range_m"], cfg["az_deg"],
                                  cfg["el_deg"], cfg["min_size"], labels.ctypes.data,
                                  dims.ctypes.data, C.byref(nvox))
        self.last_num_clusters = k
        self.last_num_voxels = nvox.value
        return (labels, dims) if want_dims else labels

    # -- polygonizer (convex hull, antipodal pairs, oriented boxes) ---------------------
    def convex_hull(self, xy):
        xy = np.ascontiguousarray(xy, np.float64)
        n = xy.shape[0]
        idx = np.zeros(max(n, 1), np.int32)
        k = self.lib.port_convex_hull(xy.ctypes.data, n, idx.ctypes.data)
        return idx[:k].copy()

    def antipodal_pairs(self, hull_xy):
        hull_xy = np.ascontiguousarray(hull_xy, np.float64)
        n = hull_xy.shape[0]
        out = np.zeros((3 * n + 4, 2), np.int32)
        k = self.lib.port_antipodal_pairs(hull_xy.ctypes.data, n, out.ctypes.data)
        return out[:k].copy()

    def bounding_box(self, hull_xy, method=0):
        """[11] = 4 corners (x, y), area, angle_rad, is_valid; method 0 rotating calipers, 1 PCA."""
        hull_xy = np.ascontiguousarray(hull_xy, np.float64)
        out = np.zeros(11, np.float64)
        self.lib.port_bounding_box(hull_xy.ctypes.data, hull_xy.shape[0], method, out.ctypes.data)
        return out

    def cluster_hulls(self, pts, labels):
        """Per-cluster gather + hull (processor.cpp:627-658, polygonizer.cpp:33-91)."""
        pts = _f32(pts)
        labels = np.ascontiguousarray(labels, np.int32)
        n = pts.shape[0]
        K = int(labels.max()) + 1 if n else 0
        K = max(K, 0)
        off = np.zeros(K + 1, np.uint32)
        xy = np.zeros((max(n, 1), 2), np.float64)
        hidx = np.zeros(max(n, 1), np.int32)
        zmm = np.zeros((max(K, 1), 2), np.float64)
        tot = self.lib.port_cluster_hulls(pts.ctypes.data, pts.shape[1], n, labels.ctypes.data, K,
                                          off.ctypes.data, xy.ctypes.data, hidx.ctypes.data,
                                          zmm.ctypes.data)
        return off, xy[:tot].copy(), hidx[:tot].copy(), zmm[:K].copy()

    def rng_draws(self, n, count, std=False):
        out = np.zeros(count, np.uint32)
        (self.lib.port_std_rng_draws if std else self.lib.port_rng_draws)(n, count, out.ctypes.data)
        return out


def read_pcd_xyzi(path: str) -> np.ndarray:
    """PCD v0.7 binary, FIELDS x y z intensity (4 x float32). Reads exactly POINTS*16 bytes
    (the KITTI files under /root/reference/data carry trailing padding)."""
    with open(path, "rb") as f:
        npts = None
        while True:
            line = f.readline()
            if not line:
                raise ValueError("bad PCD header")
            tok = line.split()
            if tok and tok[0] == b"POINTS":
                npts = int(tok[1])
            if tok and tok[0] == b"DATA":
                assert tok[1] == b"binary"
                break
        raw = f.read(npts * 16)
    return np.frombuffer(raw, np.float32).reshape(npts, 4).copy()


def label_hash(labels) -> str:
    """FNV-1a over the u32 labels."""
    h = 0xCBF29CE484222325
    for v in np.asarray(labels, np.uint32).tolist():
        h = ((h ^ v) * 0x100000001B3) & 0xFFFFFFFFFFFFFFFF
    return f"{h:016x}"
