"""CPU suite: pins the oracle (port restatement + unmodified reference build) against the committed
golden fixtures, which were produced by the reference's own sources (tools/make_golden.py).
The reference's test-suite holds no vectors for this path."""
import json
import os

import numpy as np
import pytest

from oracle.oracle import JCP_AS_IS, JCP_AS_IS_DATAFLOW, JCP_CLEAN, JCP_CLEAN_DATAFLOW, NODE_CLUSTER_CFG, label_hash
from tools import frames as F


def _check_frame(o, g, is_port):
    pts = g["pts"]
    ring = o.ring_partition(pts) if is_port else g["ring"].astype(np.uint16)
    if is_port:
        assert np.array_equal(ring, g["ring"].astype(np.uint16))
        labels, img, dbg = o.segment(pts, ring, want_image=True, want_debug=True)
        elev = dbg["elevation"]
    else:
        labels, img = o.segment(pts, ring, want_image=True)
        elev = o.segment_intermediates()["elevation"]
    assert np.array_equal(labels, g["labels"].astype(np.uint32))
    assert np.array_equal(np.packbits(img.reshape(-1) > 0), g["image"])
    assert np.array_equal(elev.view(np.uint32), g["elevation"].view(np.uint32))
    assert np.array_equal(o.segment(pts, None), g["labels_noring"].astype(np.uint32))
    obs = np.ascontiguousarray(pts[labels == 2])
    if is_port:
        cl, dims = o.cluster(obs, want_dims=True, **NODE_CLUSTER_CFG)
    else:
        o.cluster_config(**NODE_CLUSTER_CFG)
        cl, dims = o.cluster(obs, want_dims=True)
    assert np.array_equal(cl, g["cluster_labels"].astype(np.int32))
    assert np.array_equal(dims, g["voxel_dims"])
    return obs, cl


def test_port_matches_golden(port, golden0, golden100):
    for g in (golden0, golden100):
        obs, cl = _check_frame(port, g, True)
        off, xy, idx, zmm = port.cluster_hulls(obs, cl)
        assert np.array_equal(off, g["hull_offsets"])
        assert np.array_equal(xy.astype(np.float32), g["hull_xy"])
        assert np.array_equal(zmm.astype(np.float32), g["zminmax"])
        assert np.array_equal(np.packbits(port.dror(g["pts"])), g["dror_exact"])


def test_reference_build_matches_golden(ref, golden0):
    _check_frame(ref, golden0, False)
    n = golden0["pts"].shape[0]
    assert np.array_equal(np.packbits(ref.dror(golden0["pts"], mode="exact")), golden0["dror_exact"])
    # the as-is result (stale KD-tree stack, hazard H1) is one-directional: it only ever turns
    # NOISE into VALID
    as_is = np.unpackbits(golden0["dror_as_is"])[:n]
    exact = np.unpackbits(golden0["dror_exact"])[:n]
    assert int(((as_is == 1) & (exact == 0)).sum()) == 0
    assert int(as_is.sum()) == 69 and int(exact.sum()) == 1050


def test_known_answers_survey(golden0):
    """SURVEY.md 8c provisional known-answers, re-derived from the committed fixture."""
    lab = golden0["labels"]
    assert golden0["pts"].shape[0] == 123398
    assert [int((lab == k).sum()) for k in (1, 2, 0)] == [67718, 46500, 9180]
    assert int(golden0["cluster_labels"].max()) + 1 == 262
    assert int((golden0["cluster_labels"] < 0).sum()) == 151
    assert golden0["hull_xy"].shape[0] == 1803
    assert int(golden0["ring"].min()) == 0 and int(golden0["ring"].max()) == 63


def test_rng_stream_matches_libstdcxx(port):
    """std::mt19937{42} + uniform_int_distribution (segmenter.cpp:369-371): the restated Lemire
    mapping must equal libstdc++'s own distribution (hazard H8)."""
    import random

    first = port.rng_draws(0xFFFFFFFF, 4)  # n = 2^32 - 1 is (almost) the raw stream
    assert first.shape == (4,)
    for n in (2, 3, 7, 1000, 38328, 40552, 123457, 2 ** 31 + 11):
        a = port.rng_draws(n, 300)
        b = port.rng_draws(n, 300, std=True)
        assert np.array_equal(a, b), n
        assert int(a.max()) < n
    _ = random


def test_port_equals_reference_on_synthetic(port, ref):
    for seed in (4001, 4002):
        pts, ring = F.synth_scan(seed)
        a = port.segment(pts, ring)
        b = ref.segment(pts, ring)
        assert np.array_equal(a, b)
        assert np.array_equal(port.segment(pts, None), ref.segment(pts, None))
        obs = np.ascontiguousarray(pts[a == 2])
        for cfg in (NODE_CLUSTER_CFG, dict(range_m=0.4, az_deg=1.0, el_deg=1.5, min_size=3),
                    dict(range_m=1.0, az_deg=2.0, el_deg=2.0, min_size=10)):
            ref.cluster_config(**cfg)
            assert np.array_equal(port.cluster(obs, **cfg), ref.cluster(obs))
        assert np.array_equal(port.dror(pts), ref.dror(pts, mode="exact"))


def test_port_equals_reference_on_whole_row_queues(port, ref):
    """Concentric walls: the JCP queue holds whole image rows (the longest in-row dependency chains).
    The GPU parity test for this scene checks against the port, so the port is pinned to the
    unmodified reference sources here, labels and image."""
    for seed, kw in ((7, {}), (8, dict(radii=(5.0, 5.6, 7.5, 11.0), height=1.2))):
        pts, ring = F.synth_ring_walls(seed, **kw)
        a, img_a, dbg = port.segment(pts, ring, want_image=True, want_debug=True)
        b, img_b = ref.segment(pts, ring, want_image=True)
        assert dbg["n_queued"] > 8000
        assert np.array_equal(a, b) and np.array_equal(img_a, img_b)
        assert np.array_equal(port.segment(pts, None), ref.segment(pts, None))


def test_port_equals_reference_128_beams(port, ref):
    from oracle.oracle import default_seg_cfg

    pts, ring = F.synth_scan(3000, beams=128, n_boxes=120, n_poles=80, dropout=0.01)
    cfg = default_seg_cfg(image_height=128)
    port.segment_config(cfg)
    ref.segment_config(cfg)
    try:
        assert np.array_equal(port.segment(pts, ring), ref.segment(pts, ring))
    finally:
        port.segment_config(default_seg_cfg())
        ref.segment_config(default_seg_cfg())


def test_jcp_dataflow_equals_raster(port, golden0, golden100):
    """The raster-order Gauss-Seidel sweep of the reference only couples a pixel to EARLIER queued
    pixels, so any schedule that respects those dependencies gives the same result (the GPU sweeps
    row by row with the in-row recurrence resolved by a scan). The port carries a data-flow schedule
    as a cross-check of that order independence."""
    for g in (golden0, golden100):
        ring = g["ring"].astype(np.uint16)
        assert np.array_equal(port.segment(g["pts"], ring, jcp_mode=JCP_AS_IS),
                              port.segment(g["pts"], ring, jcp_mode=JCP_AS_IS_DATAFLOW))
        assert np.array_equal(port.segment(g["pts"], ring, jcp_mode=JCP_CLEAN),
                              port.segment(g["pts"], ring, jcp_mode=JCP_CLEAN_DATAFLOW))


def test_edge_cases(port, ref):
    empty = np.zeros((0, 4), np.float32)
    assert port.segment(empty, np.zeros(0, np.uint16)).shape == (0,)
    assert ref.segment(empty, np.zeros(0, np.uint16)).shape == (0,)
    assert port.cluster(empty).shape == (0,)
    assert port.dror(empty).shape == (0,)
    assert port.ring_partition(empty).shape == (0,)
    # isolated points are all noise; 4 coincident points are all valid (self counts)
    far = np.array([[10, 0, 0, 0], [20, 0, 0, 0], [30, 5, 0, 0]], np.float32)
    assert port.dror(far).tolist() == [1, 1, 1]
    assert ref.dror(far, mode="exact").tolist() == [1, 1, 1]
    same = np.tile(np.array([[10, 1, 0, 0]], np.float32), (4, 1))
    assert port.dror(same).tolist() == [0, 0, 0, 0]
    # a cluster below min size is dropped; hull of < 3 points is the identity
    two = np.array([[10, 0, 0, 0], [10.05, 0, 0, 0]], np.float32)
    assert port.cluster(two, **NODE_CLUSTER_CFG).tolist() == [-1, -1]
    assert port.convex_hull(np.array([[0.0, 0.0], [1.0, 1.0]])).tolist() == [0, 1]
    # collinear points collapse to the two extremes; square keeps 4 corners CCW from the lexicographic minimum
    line = np.stack([np.arange(6.0), np.arange(6.0)], -1)
    assert port.convex_hull(line).tolist() == [0, 5]
    sq = np.array([[1, 1], [0, 0], [1, 0], [0, 1], [0.5, 0.5]], np.float64)
    assert port.convex_hull(sq).tolist() == [1, 2, 0, 3]


def test_dilate_shim_matches_opencv(ref):
    cv2 = pytest.importorskip("cv2")
    rng = np.random.default_rng(0)
    img = (rng.random((64, 2048)) < 0.02).astype(np.uint8) * 255
    exp = cv2.dilate(img, cv2.getStructuringElement(cv2.MORPH_RECT, (5, 5)))
    assert np.array_equal(ref.shim_dilate(img), exp)


@pytest.mark.skipif(not F.have_pack(), reason="data/kitti154.npz not built")
def test_port_matches_reference_summary_all_154(port):
    """Every KITTI frame: the port's outputs hash to what the unmodified reference produced."""
    summ = json.load(open(os.path.join(F.GOLDEN_DIR, "kitti154_summary.json")))["frames"]
    frames = F.load_pack()
    assert len(frames) == len(summ) == 154
    for s, pts in list(zip(summ, frames))[::3]:
        ring = port.ring_partition(pts)
        assert label_hash(ring) == s["ring_hash"]
        labels = port.segment(pts, ring)
        assert label_hash(labels) == s["label_hash"], s["frame"]
        obs = np.ascontiguousarray(pts[labels == 2])
        cl = port.cluster(obs, **NODE_CLUSTER_CFG)
        assert int(cl.max()) + 1 == s["clusters"]
        assert label_hash(cl.astype(np.int64).astype(np.uint32)) == s["cluster_hash"]
        off, xy, idx, zmm = port.cluster_hulls(obs, cl)
        assert xy.shape[0] == s["hull_vertices"]
        assert label_hash(xy.astype(np.float32).view(np.uint32).reshape(-1)) == s["hull_xy_hash"]
        assert int(port.dror(pts).sum()) == s["dror_noise_exact"]


def _random_hull_inputs(rng, trials):
    for trial in range(trials):
        n = int(rng.integers(1, 60))
        kind = trial % 4
        if kind == 0:
            xy = np.round(rng.normal(0, 3, (n, 2)), 3)
        elif kind == 1:
            xy = np.round(rng.uniform(-5, 5, (n, 2)), 1)  # many ties and collinear runs
        elif kind == 2:
            t = rng.uniform(0, 2 * np.pi, n)
            xy = np.round(np.c_[np.cos(t) * 4, np.sin(t) * 2], 3)
        else:
            xy = np.round(np.c_[rng.uniform(0, 10, n), rng.integers(0, 2, n).astype(float)], 2)  # two parallel lines
        yield xy.astype(np.float32).astype(np.float64)


def test_polygonizer_port_equals_reference(port, ref):
    """convexHull, findAntipodalPairsOfConvexHull, boundingBoxRotatingCalipers and the PCA box of the
    restatement against the reference's own polygonizer.cpp (oracle/_ref): bit-exact. (The PCA box
    goes through the Eigen stand-in on both sides: that one is a consistency check, not a pin.)"""
    if not ref.has_polygonizer:
        pytest.skip("oracle/_ref predates the polygonizer wrappers")
    rng = np.random.default_rng(7)
    for xy in _random_hull_inputs(rng, 1500):
        a, b = port.convex_hull(xy), ref.convex_hull(xy)
        assert a.shape == b.shape and np.array_equal(xy[a], xy[b])
        h = xy[b]
        assert np.array_equal(port.antipodal_pairs(h), ref.antipodal_pairs(h))
        for m in (0, 1):
            assert np.array_equal(port.bounding_box(h, m).view(np.uint64), ref.bounding_box(h, m).view(np.uint64))


def test_polygonizer_golden(port, golden0, golden100):
    """The hulls of the golden fixtures (made with the restated hull) equal the reference's own
    convexHull output, and the restated boxes equal the reference's (tests/golden/kitti_polygonizer.npz,
    tools/make_golden_boxes.py)."""
    gp = np.load(os.path.join(F.GOLDEN_DIR, "kitti_polygonizer.npz"))
    for name, g in (("kitti_f000", golden0), ("kitti_f100", golden100)):
        off, hxy = gp[name + "_hull_offsets"], gp[name + "_hull_xy"]
        assert np.array_equal(off, g["hull_offsets"])
        assert np.array_equal(hxy, g["hull_xy"])
        boxes = gp[name + "_boxes"]
        for k in range(len(off) - 1):
            h = hxy[off[k]:off[k + 1]].astype(np.float64)
            assert len(port.antipodal_pairs(h)) == gp[name + "_pairs"][k]
            assert np.array_equal(port.bounding_box(h, 0).view(np.uint64), boxes[k, :11].view(np.uint64))
            assert np.array_equal(port.bounding_box(h, 1).view(np.uint64), boxes[k, 11:].view(np.uint64))
