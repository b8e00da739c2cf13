"""GPU suite: the CUDA path (always through the C ABI, lidar_processing_v2_b200.native) against
the CPU oracle on the same inputs, the committed golden fixtures, and size-independent properties
at full batch size. Integer / label / index outputs are compared bit-exactly; hull vertex
coordinates are floats copied from the input, also compared bit-exactly (tolerance 0 <= 1e-5)."""
import json
import os

import numpy as np
import pytest

import lidar_processing_v2_b200 as lpl
import parity
from oracle.oracle import JCP_AS_IS, JCP_CLEAN, NODE_CLUSTER_CFG, PortOracle, default_seg_cfg, label_hash
from tools import frames as F

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ctx():
    c = lpl.Context(0, max_points=131072, max_frames=8)
    yield c
    c.close()


def _assert_parity(rep):
    bad = {k: v for k, v in rep.items() if k.endswith("_diff") and v != 0}
    assert not bad and rep.get("seg_status", 0) == 0, rep


def test_device_footprint(ctx):
    # one slab per context; stage-local scratch planes share storage (capi.cu: carve): < 0.7 KB per point of capacity
    # including the per-frame range-image / polar-grid / DROR-grid planes (0.75 KB in round 1)
    per_point = ctx.device_bytes() / (8 * 131072)
    assert 300 < per_point < 700, per_point


def test_native_library_is_the_path(ctx):
    # the extension in-tree is what runs: kernels were launched by this context
    ctx.launch_count(reset=True)
    ctx.dror_filter(np.zeros((10, 4), np.float32))
    assert ctx.launch_count() > 0
    assert os.path.exists(lpl.SO_PATH)


def test_stagewise_vs_unmodified_reference(ctx, ref, golden0, golden100):
    for g in (golden0, golden100):
        _assert_parity(parity.stage_report(ctx, ref, g["pts"]))


def test_stagewise_vs_golden_fixture(ctx, golden0):
    g = golden0
    pts = g["pts"]
    assert np.array_equal(ctx.ring_partition(pts), g["ring"].astype(np.uint16))
    assert np.array_equal(np.packbits(ctx.dror_filter(pts)), g["dror_exact"])
    labels, img = ctx.segment(pts, g["ring"].astype(np.uint16), want_image=True)
    assert np.array_equal(labels, g["labels"].astype(np.uint32))
    assert np.array_equal(np.packbits(img.reshape(-1) > 0), g["image"])
    assert np.array_equal(ctx.debug_segment(0)["elevation"].view(np.uint32), g["elevation"].view(np.uint32))
    assert np.array_equal(ctx.segment(pts, None), g["labels_noring"].astype(np.uint32))
    obs = np.ascontiguousarray(pts[labels == 2])
    ctx.cluster_config(**NODE_CLUSTER_CFG)
    cl, k = ctx.cluster(obs)
    assert np.array_equal(cl, g["cluster_labels"].astype(np.int32)) and k == 262
    assert np.array_equal(ctx.debug_cluster(0), g["voxel_dims"])
    off, xy, idx, zmm = ctx.cluster_hulls(obs, cl)
    assert np.array_equal(off, g["hull_offsets"])
    assert np.abs(xy - g["hull_xy"]).max() <= 1e-5 and np.array_equal(xy, g["hull_xy"])
    assert np.array_equal(zmm, g["zminmax"])


def test_clean_jcp_mode(ctx, port, golden0):
    _assert_parity(parity.stage_report(ctx, port, golden0["pts"], jcp_mode=JCP_CLEAN, check_ringless=False))
    ctx.set_jcp_mode(lpl.JCP_AS_REFERENCE)


def test_synthetic_frames(ctx, ref):
    for seed in (4001, 4007):
        pts, ring = F.synth_scan(seed)
        _assert_parity(parity.stage_report(ctx, ref, pts, ring_given=ring))


def test_cluster_configs_and_azimuth_seam(ctx, port):
    """H3: the reference's literal azimuth wrap (num_azimuth = ceil(max/res) + 1) must be kept."""
    rng = np.random.default_rng(5)
    # a ring of points around the sensor crossing the 0 / 2 pi seam, plus blobs
    a = rng.uniform(0, 2 * np.pi, 4000)
    r = 12.0 + rng.normal(0, 0.05, a.size)
    ringpts = np.stack([r * np.cos(a), r * np.sin(a), rng.uniform(-1, 0.5, a.size)], -1)
    blobs = rng.normal(0, 0.3, (3000, 3)) + rng.uniform(-40, 40, (30, 1, 3)).repeat(100, 1).reshape(-1, 3) * [1, 1, 0.02]
    pts = np.zeros((7000, 4), np.float32)
    pts[:, :3] = np.round(np.concatenate([ringpts, blobs]) * 1000) / 1000
    for cfg in (NODE_CLUSTER_CFG, dict(range_m=0.4, az_deg=1.0, el_deg=1.5, min_size=3),
                dict(range_m=1.0, az_deg=2.0, el_deg=2.0, min_size=10),
                dict(range_m=0.2, az_deg=0.5, el_deg=1.0, min_size=1)):
        ctx.cluster_config(**cfg)
        got, k = ctx.cluster(pts)
        exp, dims = port.cluster(pts, want_dims=True, **cfg)
        assert np.array_equal(got, exp), cfg
        assert np.array_equal(ctx.debug_cluster(0), dims)
        assert k == int(exp.max()) + 1
    ctx.cluster_config(**NODE_CLUSTER_CFG)


def test_convex_hull_entry_point(ctx, port):
    rng = np.random.default_rng(1)
    cases = [
        np.round(rng.normal(0, 5, (500, 2)), 3),
        np.round(rng.uniform(-1, 1, (5000, 2)), 2),            # many exact duplicates and collinear runs
        np.stack([np.arange(50.0), 2 * np.arange(50.0)], -1),   # collinear -> 2 vertices
        np.array([[0, 0], [1, 0], [1, 1], [0, 1], [0.5, 0.5]], float),
        np.array([[3.0, 4.0]]), np.array([[0.0, 0.0], [1.0, 1.0]]),
        np.repeat(np.array([[2.0, 2.0]]), 7, 0),                 # all identical
        rng.normal(0, 8, (30000, 2)),                             # 30 chunks of the shared-memory hull pass + the join
        np.c_[100 * np.cos(np.linspace(0, 2 * np.pi, 3000, endpoint=False)),
              100 * np.sin(np.linspace(0, 2 * np.pi, 3000, endpoint=False))],   # ~3000 points in convex position: more hull
                                                                  # vertices than a warp's buffer holds -> the global-memory path
        np.c_[np.linspace(-5, 5, 1500), np.linspace(-5, 5, 1500) ** 2],        # a parabola: every point a vertex, 2 chunks
    ]
    for xy in cases:
        xy = xy.astype(np.float32).astype(np.float64)
        got = ctx.convex_hull(xy)
        exp = port.convex_hull(xy)
        assert got.shape == exp.shape
        # indices may differ between equal points (unstable sort), coordinates may not
        assert np.abs(xy[got] - xy[exp]).max() <= 1e-5 if got.size else True
        assert np.array_equal(xy[got], xy[exp])
    # coordinates that are not float-representable take the fp64 path (the reference accepts any doubles)
    for xy in (np.array([[0.1, 0.2], [1.0, 1.0], [0.0, 3.0]]), rng.normal(0, 3, (4000, 2)), rng.uniform(-1, 1, (2, 2)),
               np.c_[np.linspace(0, 1, 300) ** 3, np.linspace(0, 1, 300)]):
        got = ctx.convex_hull(xy)
        exp = port.convex_hull(xy)
        assert got.shape == exp.shape and np.array_equal(xy[got], xy[exp])


def _boxes_equal(got, exp, exact=True):
    """got: BBOX_DTYPE records, exp: [K][11] oracle rows (4 corners, area, angle, valid)."""
    assert np.array_equal(got["is_valid"] != 0, exp[:, 10] != 0)
    v = exp[:, 10] != 0
    if exact:
        assert np.array_equal(got["corners"].reshape(-1, 8)[v].view(np.uint64), exp[v, :8].view(np.uint64))
        assert np.array_equal(got["area"][v].view(np.uint32), exp[v, 8].astype(np.float32).view(np.uint32))
        assert np.array_equal(got["angle_rad"][v].view(np.uint32), exp[v, 9].astype(np.float32).view(np.uint32))
    else:
        assert np.abs(got["corners"].reshape(-1, 8)[v] - exp[v, :8]).max(initial=0.0) <= 1e-9
        assert np.abs(got["area"][v] - exp[v, 8]).max(initial=0.0) <= 1e-5 * max(1.0, np.abs(exp[v, 8]).max(initial=0.0))
        assert np.abs(got["angle_rad"][v] - exp[v, 9]).max(initial=0.0) <= 1e-6
    assert not got["corners"][~v].any()


def test_bounding_boxes_entry_point(ctx, port):
    """lpl_bounding_boxes against the restated / reference polygonizer (src/polygonizer.cpp:93-362):
    rotating calipers bit-exact; PCA within 1e-9 m (device atan2 vs glibc for the yaw, 1e-6 rad)."""
    rng = np.random.default_rng(3)
    hulls = []
    for n in list(rng.integers(3, 40, 300)) + [3, 4, 200, 700, 1, 2, 0]:
        t = np.sort(rng.uniform(0, 2 * np.pi, int(n)))
        r = rng.uniform(0.5, 6.0)
        xy = np.c_[np.cos(t) * r * rng.uniform(0.3, 1.0), np.sin(t) * r] + rng.uniform(-40, 40, 2)
        xy = np.round(xy, 3).astype(np.float32).astype(np.float64)
        hulls.append(xy[port.convex_hull(xy)] if len(xy) else xy.reshape(0, 2))
    hulls.append(np.array([[0, 0], [2, 0], [2, 1], [0, 1]], float))          # rectangle: parallel edges everywhere
    hulls.append(np.array([[0, 0], [1, 0], [0.5, 1e-7]], float))               # sliver
    off = np.concatenate([[0], np.cumsum([len(h) for h in hulls])]).astype(np.uint32)
    allxy = np.concatenate(hulls)
    for method, exact in ((lpl.BOX_ROTATING_CALIPERS, True), (lpl.BOX_PCA, False)):
        got = ctx.bounding_boxes(allxy, off, method)
        exp = np.stack([port.bounding_box(h, method) for h in hulls])
        _boxes_equal(got, exp, exact)
    assert ctx.bounding_boxes(np.zeros((0, 2)), np.zeros(1, np.uint32)).shape == (0,)


def test_pipeline_boxes_vs_reference_golden(ctx, golden0):
    """LPL_STAGE_BOXES on the chained pipeline: one rotating-calipers box per cluster hull, against the
    boxes the reference's own polygonizer.cpp produced for this frame (tests/golden/kitti_polygonizer.npz)."""
    gp = np.load(os.path.join(F.GOLDEN_DIR, "kitti_polygonizer.npz"))
    ctx.cluster_config(**NODE_CLUSTER_CFG)
    nf = ctx.upload([golden0["pts"]])
    ctx.run(nf, (lpl.STAGE_ALL & ~lpl.STAGE_DROR) | lpl.STAGE_BOXES)
    ctx.sync(nf)
    out = ctx.download(0, want_boxes=True)
    assert np.array_equal(out["hull_offsets"], gp["kitti_f000_hull_offsets"])
    assert np.array_equal(out["hull_xy"], gp["kitti_f000_hull_xy"])
    _boxes_equal(out["boxes"], gp["kitti_f000_boxes"][:, :11], exact=True)


def test_pointcloud2_ingest(ctx, golden0, golden100):
    """lpl_pipeline_upload_cloud2 (replaces Processor::convert<PointT>, processor.cpp:42-179): raw
    PointXYZIR / PointXYZ records unpacked on the device give the very same results as the packed upload."""
    ctx.cluster_config(**NODE_CLUSTER_CFG)
    frames = [golden0["pts"], golden100["pts"][:70001]]
    rings = [golden0["ring"].astype(np.uint16), golden100["ring"].astype(np.uint16)[:70001]]
    nf = ctx.upload(frames, rings=rings)
    stages = lpl.STAGE_ALL & ~lpl.STAGE_RING
    ctx.run(nf, stages)
    ctx.sync(nf)
    want = [ctx.download(f) for f in range(nf)]
    msgs = []
    for pts, ring in zip(frames, rings):
        n = pts.shape[0]
        rec = np.zeros((n, 32), np.uint8)           # pcl::PointXYZIR: x y z pad | intensity ring pad
        rec[:, 0:12] = pts[:, :3].copy().view(np.uint8).reshape(n, 12)
        rec[:, 16:20] = pts[:, 3:4].copy().view(np.uint8).reshape(n, 4)
        rec[:, 20:22] = ring.view(np.uint8).reshape(n, 2)
        msgs.append(dict(data=rec.reshape(-1), width=n, height=1, point_step=32, row_step=32 * n,
                         x_offset=0, y_offset=4, z_offset=8, ring_offset=20))
    nf = ctx.upload_cloud2(msgs)
    ctx.run(nf, stages)
    ctx.sync(nf)
    for f in range(nf):
        got = ctx.download(f)
        for k in ("ring", "noise", "labels", "cluster_labels", "hull_offsets", "hull_xy", "zminmax"):
            assert np.array_equal(got[k], want[f][k]), (f, k)
    # organized 2-row layout with a row pitch and unaligned (packed 14-byte) records, no ring
    pts = golden0["pts"][:4000]
    rec = np.zeros((2, 2000 * 14 + 6), np.uint8)
    body = np.zeros((4000, 14), np.uint8)
    body[:, 1:13] = pts[:, :3].copy().view(np.uint8).reshape(4000, 12)
    rec[:, :2000 * 14] = body.reshape(2, 2000 * 14)
    ctx.upload_cloud2([dict(data=rec.reshape(-1), width=2000, height=2, point_step=14, row_step=2000 * 14 + 6,
                            x_offset=1, y_offset=5, z_offset=9, ring_offset=-1)])
    ctx.run(1, lpl.STAGE_RING | lpl.STAGE_DROR)
    ctx.sync(1)
    a = ctx.download(0)
    ctx.upload([pts])
    ctx.run(1, lpl.STAGE_RING | lpl.STAGE_DROR)
    ctx.sync(1)
    b = ctx.download(0)
    assert a["n"] == 4000 and np.array_equal(a["ring"], b["ring"]) and np.array_equal(a["noise"], b["noise"])


def test_packed_upload_equals_per_frame_upload(ctx, golden0, golden100):
    frames = [golden0["pts"][:50000], golden100["pts"], golden0["pts"][:0], golden0["pts"][:7]]
    ctx.cluster_config(**NODE_CLUSTER_CFG)
    nf = ctx.upload(frames)
    ctx.run(nf, lpl.STAGE_ALL)
    ctx.sync(nf)
    want = [ctx.download(f) for f in range(nf)]
    nf = ctx.upload_packed(np.concatenate(frames), [f.shape[0] for f in frames])
    ctx.run(nf, lpl.STAGE_ALL)
    ctx.sync(nf)
    for f in range(nf):
        got = ctx.download(f)
        for k in ("ring", "noise", "labels", "cluster_labels", "hull_offsets", "hull_xy", "zminmax"):
            assert np.array_equal(got[k], want[f][k]), (f, k)


def test_packed_xyz_upload_and_packed_download(ctx, golden0, golden100):
    """The 12-byte std::array<float, 3> batch upload (noise_remover.hpp:68) and the one-transfer packed
    download give what the per-frame calls give, for every plane, on a ragged batch with an empty frame."""
    frames = [golden0["pts"][:50000], golden100["pts"], golden0["pts"][:0], golden0["pts"][:7], golden0["pts"][:2049]]
    ctx.cluster_config(**NODE_CLUSTER_CFG)
    stages = lpl.STAGE_ALL | lpl.STAGE_BOXES
    nf = ctx.upload(frames)
    ctx.run(nf, stages)
    ctx.sync(nf)
    want = [ctx.download(f, want_boxes=True) for f in range(nf)]
    xyz = np.ascontiguousarray(np.concatenate(frames)[:, :3])
    nf = ctx.upload_packed_xyz(xyz, [f.shape[0] for f in frames])
    ctx.run(nf, stages)
    names = [p[0] for p in lpl.PLANES]
    bufs = lpl.PackedBuffers(8, 8 * 131072 * 40, want=names)
    counts = ctx.download_packed(nf, bufs)
    assert not ctx.status(nf).any()
    total = 0
    for f in range(nf):
        w = want[f]
        assert counts[:, f].tolist() == [w["n"], w["num_valid"], w["num_obstacles"], w["num_clusters"], w["num_hull_vertices"]]
        assert np.array_equal(bufs.frame("labels_u8", f), w["labels"].astype(np.uint8))
        for k in ("noise", "ring", "obstacle_index", "cluster_labels", "hull_offsets", "hull_indices", "hull_xy", "zminmax"):
            assert np.array_equal(bufs.frame(k, f), w[k]), (f, k)
        assert bufs.frame("boxes", f).tobytes() == w["boxes"].tobytes()
        total += w["n"] * 4 + w["num_obstacles"] * 8 + (w["num_clusters"] + 1) * 4 + w["num_hull_vertices"] * 12 + w["num_clusters"] * 88
    assert total <= bufs.bytes_used <= total + 16 * len(names)   # only the occupied bytes cross PCIe (+ plane alignment)
    # a subset of the planes, and a host buffer that is too small
    sub = lpl.PackedBuffers(8, 1 << 20, want=("labels_u8", "hull_xy"))
    ctx.download_packed(nf, sub)
    assert np.array_equal(sub.frame("hull_xy", 1), want[1]["hull_xy"])
    tiny = lpl.PackedBuffers(8, 1024, want=("labels_u8",))
    with pytest.raises(lpl.LplError) as e:
        ctx.download_packed(nf, tiny)
    assert e.value.code == lpl.native.LPL_ERR_CAPACITY
    for b in (bufs, sub, tiny):
        b.close()


def test_run_without_hulls_reports_no_stale_vertices(ctx, golden0):
    ctx.cluster_config(**NODE_CLUSTER_CFG)
    nf = ctx.upload([golden0["pts"]])
    no_dror = lpl.STAGE_ALL & ~lpl.STAGE_DROR   # the golden counts are those of the node's chain (no DROR)
    ctx.run(nf, no_dror)
    ctx.sync(nf)
    assert ctx.download(0)["num_hull_vertices"] == 1803
    ctx.run(nf, no_dror & ~lpl.STAGE_HULLS)
    ctx.sync(nf)
    out = ctx.download(0)
    assert out["num_clusters"] == 262 and out["num_hull_vertices"] == 0 and not out["hull_offsets"].any()


def test_processor_glue_split_clouds_and_markers(port, golden0, golden100):
    """lpl_pipeline_split_clouds against the node's host loops (processor.cpp:562-579 label split, :627-647 clustered
    cloud with std::rand() colours, :254-343 marker line lists), restated here in numpy on the oracle's chain."""
    frames = [golden0["pts"], golden100["pts"][:70001].copy(), golden0["pts"][:0], golden0["pts"][:300].copy()]
    c = lpl.Context(0, max_points=131072, max_frames=4)
    try:
        c.cluster_config(**NODE_CLUSTER_CFG)
        nf = c.upload(frames)
        c.run(nf, lpl.STAGE_ALL & ~lpl.STAGE_DROR)     # the node's chain
        c.sync(nf)
        res = [c.download(f) for f in range(nf)]
        got = c.split_clouds(nf, 131072)
        rnd = lpl.glibc_rand_stream(1, 3 * sum(r["num_clusters"] for r in res)) % 256   # a fresh context = a fresh process
        ro = 0
        for f, pts in enumerate(frames):
            exp = parity.oracle_chain(port, pts, dror=False)
            lab = exp["labels"]
            for name, sel, rgb in (("ground", lab == 1, (124, 252, 0)), ("obstacle", lab == 2, (200, 0, 0)),
                                   ("unsegmented", (lab != 1) & (lab != 2), (255, 255, 0))):
                g = got[name][f]
                assert g.shape[0] == int(sel.sum()), (f, name)
                assert np.array_equal(g["xyz"], pts[sel][:, :3]) and np.all(g["w"] == 1.0)
                assert np.all(g["bgra"] == np.array([rgb[2], rgb[1], rgb[0], 255], np.uint8)) and not g["pad"].any()
            obs = pts[lab == 2]
            cl = exp["cluster_labels"]
            K = exp["num_clusters"]
            order = np.concatenate([np.flatnonzero(cl == k) for k in range(K)]) if K else np.zeros(0, np.int64)
            g = got["clustered"][f]
            assert g.shape[0] == order.shape[0]
            assert np.array_equal(g["xyz"], obs[order][:, :3])
            cols = rnd[ro:ro + 3 * K].reshape(K, 3).astype(np.uint8)   # r, g, b per cluster, in label order
            ro += 3 * K
            exp_bgra = np.stack([cols[cl[order], 2], cols[cl[order], 1], cols[cl[order], 0], np.full(order.shape[0], 255, np.uint8)], -1) \
                if K else np.zeros((0, 4), np.uint8)
            assert np.array_equal(g["bgra"], exp_bgra)
            # markers
            off, xy, zmm = exp["hull_offsets"], exp["hull_xy"].astype(np.float64), exp["zminmax"].astype(np.float64)
            lines = []
            for k in range(K):
                p = xy[off[k]:off[k + 1]]
                n = p.shape[0]
                if n < 3:
                    continue
                for z in (zmm[k, 0], zmm[k, 1]):
                    for i in list(range(1, n)) + [0]:
                        a, b = p[i - 1], p[i]
                        lines += [[a[0], a[1], z], [b[0], b[1], z]]
                for q in p:
                    lines += [[q[0], q[1], zmm[k, 0]], [q[0], q[1], zmm[k, 1]]]
            lines = np.array(lines, np.float64).reshape(-1, 3)
            assert got["markers"][f].shape == lines.shape and np.array_equal(got["markers"][f], lines), f
        # caller-supplied colours
        kmax = max(r["num_clusters"] for r in res)
        mine = np.random.default_rng(0).integers(0, 256, (nf, kmax, 3)).astype(np.uint8)
        got2 = c.split_clouds(nf, 131072, markers=False, colors=mine)
        g = got2["clustered"][0]
        cl0 = res[0]["cluster_labels"]
        order = np.concatenate([np.flatnonzero(cl0 == k) for k in range(res[0]["num_clusters"])])
        assert np.array_equal(g["bgra"][:, 2], mine[0, cl0[order], 0]) and np.array_equal(g["bgra"][:, 0], mine[0, cl0[order], 2])
        # the hull results of the batch are still intact after the split reused the sort buffers
        assert np.array_equal(c.download(0)["hull_xy"], res[0]["hull_xy"])
    finally:
        c.close()


def test_general_neighbour_queries_vs_reference_kdtree(ctx, ref, golden0):
    """lpl_knn_radius_search against the reference's own KDTree<float, 3>::radius_search (kdtree.hpp:283-337):
    neighbourhoods as sets, squared distances bit for bit. lpl_knn_k_nearest against a brute-force restatement
    (parity unpinned: the reference's k_nearest template does not compile - kdtree.hpp:246 calls a
    PriorityQueue::push_back that does not exist - and is never instantiated): the k smallest of the reference's
    dist_sqr expression (kdtree.hpp:131-143), ties by point index."""
    pts = golden0["pts"][::6]                      # ~20k points
    rng = np.random.default_rng(2)
    q = np.concatenate([pts[rng.choice(pts.shape[0], 300, replace=False)],
                        np.c_[rng.uniform(-60, 60, (100, 2)), rng.uniform(-2, 1, 100), np.zeros(100)].astype(np.float32)])
    ctx.knn_build(pts)
    assert ctx.knn_token() != 0
    for k in (1, 5, 24, 100):
        gi, gd, gc = ctx.k_nearest(q, k)
        P, Q = pts[:, :3].astype(np.float32), q[:, :3].astype(np.float32)
        d0, d1, d2 = (Q[:, None, 0] - P[None, :, 0]), (Q[:, None, 1] - P[None, :, 1]), (Q[:, None, 2] - P[None, :, 2])
        dall = (d0 * d0 + (d1 * d1 + (d2 * d2 + np.float32(0)))).astype(np.float32)
        ri = np.argsort(dall, axis=1, kind="stable")[:, :k].astype(np.uint32)    # stable: ties by index
        rd = np.take_along_axis(dall, ri.astype(np.int64), 1)
        assert np.all(gc == k)
        assert np.array_equal(gd.view(np.uint32), rd.view(np.uint32))
        uniq = np.ones_like(gd, bool)
        uniq[:, 1:] &= gd[:, 1:] != gd[:, :-1]
        uniq[:, :-1] &= gd[:, :-1] != gd[:, 1:]
        assert np.array_equal(gi[uniq], ri[uniq])
        # every returned index really is at the returned distance
        d = pts[gi.reshape(-1), :3].astype(np.float32) - np.repeat(q[:, :3], k, 0)
        assert np.array_equal((d[:, 0] * d[:, 0] + (d[:, 1] * d[:, 1] + (d[:, 2] * d[:, 2]))).astype(np.float32) == gd.reshape(-1),
                              np.ones(gd.size, bool))
    r2 = (np.maximum(0.02 * np.hypot(q[:, 0], q[:, 1]), 0.3) ** 2).astype(np.float32)
    gi, gd, gc = ctx.radius_search(q, r2, 256)
    ri, rd, rc = ref.kdtree_query(pts, q, 256, radius_sqr=r2)
    assert np.array_equal(gc, rc) and gc.max() <= 256 and gc.max() > 20
    for j in range(q.shape[0]):
        assert np.array_equal(np.sort(gi[j, :gc[j]]), np.sort(ri[j, :rc[j]]))
        assert np.array_equal(np.sort(gd[j, :gc[j]]).view(np.uint32), rd[j, :rc[j]].view(np.uint32))
    # k nearest within a radius (radius_search_k_nearest's candidates), and truncation by max_per_query
    gi2, gd2, gc2 = ctx.k_nearest(q, 4, radius_sqr=r2)
    assert np.array_equal(gc2, np.minimum(gc, 4))
    gi3, gd3, gc3 = ctx.radius_search(q, r2, 8)
    assert np.array_equal(gc3, gc) and np.array_equal(gi3[:, :8][gc[:, None] > np.arange(8)], gi[:, :8][gc[:, None] > np.arange(8)])
    # another call of the context replaces the resident point set: the token says so
    ctx.dror_filter(pts[:100])
    assert ctx.knn_token() == 0
    with pytest.raises(lpl.LplError):
        ctx.k_nearest(q, 3)


def _vehicle_match_node(hull, zmm, size, box):
    """processor.cpp:680-757 + processor.hpp:60-192, restated in float64 numpy (the node is ROS code, not in the oracle)."""
    base = np.array([[4.3, 4.6, 1.6, 1.9, 1.4, 1.5], [4.6, 5.0, 1.6, 1.9, 1.4, 1.5], [4.6, 5.2, 1.7, 2.1, 1.7, 1.8],
                     [5.2, 5.8, 1.9, 2.2, 1.8, 2.0], [4.9, 5.2, 1.7, 2.1, 1.7, 1.8]])
    tol = np.array([0.8, 0.8, 0.5, 0.5, 0.5, 0.5]) * np.array([-1, 1, -1, 1, -1, 1])
    dim = base + tol
    vol = np.stack([dim[:, 0] * dim[:, 2] * dim[:, 4], dim[:, 1] * dim[:, 3] * dim[:, 5]], -1)
    ar = np.stack([dim[:, 0] * dim[:, 2], dim[:, 1] * dim[:, 3]], -1)
    n = hull.shape[0]
    area = 0.0
    if n > 2:
        for i in range(n - 1):
            area += hull[i, 0] * hull[i + 1, 1] - hull[i + 1, 0] * hull[i, 1]
        area += hull[n - 1, 0] * hull[0, 1] - hull[0, 0] * hull[n - 1, 1]
    area = abs(area) * 0.5
    h = zmm[1] - zmm[0]
    cls = -1
    if n >= 3 and size > 150 and dim[:, 4].min() < h < dim[:, 5].max():
        v = area * h
        if vol[:, 0].min() < v < vol[:, 1].max() and box["is_valid"]:
            if area / float(box["area"]) > 0.4:
                c = box["corners"]
                e1 = np.sqrt((c[0, 0] - c[1, 0]) ** 2 + (c[0, 1] - c[1, 1]) ** 2)
                e2 = np.sqrt((c[1, 0] - c[2, 0]) ** 2 + (c[1, 1] - c[2, 1]) ** 2)
                ln, wd = max(e1, e2), min(e1, e2)
                for k in range(5):
                    if (dim[k, 0] < ln < dim[k, 1] and dim[k, 2] < wd < dim[k, 3] and dim[k, 4] < h < dim[k, 5]
                            and vol[k, 0] < v < vol[k, 1] and ar[k, 0] < area < ar[k, 1]):
                        cls = k
                        break
    return cls, area


def test_vehicle_shape_matching(ctx, port, golden0):
    """lpl_vehicle_match against the node's disabled polygon simplification (processor.cpp:680-757): class index
    and polygon area per cluster, on synthetic car-sized rectangles of every class and on the clusters of a KITTI frame."""
    rng = np.random.default_rng(4)
    hulls, zmm, sizes = [], [], []
    for ln, wd, ht in [(4.4, 1.7, 1.45), (4.8, 1.8, 1.45), (5.0, 1.9, 1.75), (5.5, 2.0, 1.9), (5.0, 2.0, 1.75), (9.0, 2.5, 3.0),
                       (1.0, 0.5, 1.7), (4.4, 1.7, 0.3), (4.5, 1.8, 1.5), (3.2, 1.0, 1.2), (6.7, 2.8, 2.6)]:
        for _ in range(4):
            t = rng.uniform(0, np.pi)
            R = np.array([[np.cos(t), -np.sin(t)], [np.sin(t), np.cos(t)]])
            rect = np.array([[-ln / 2, -wd / 2], [ln / 2, -wd / 2], [ln / 2, wd / 2], [-ln / 2, wd / 2]]) @ R.T + rng.uniform(-30, 30, 2)
            rect = np.round(rect, 3).astype(np.float32).astype(np.float64)
            hulls.append(rect[port.convex_hull(rect)])
            z0 = np.float32(rng.uniform(-1.8, -1.5))
            zmm.append([float(z0), float(np.float32(z0 + ht))])
            sizes.append(int(rng.integers(100, 400)))
    # the clusters of a KITTI frame
    pts = golden0["pts"]
    lab = golden0["labels"]
    obs = np.ascontiguousarray(pts[lab == 2])
    cl = golden0["cluster_labels"].astype(np.int32)
    off0, xy0 = golden0["hull_offsets"], golden0["hull_xy"].astype(np.float64)
    for k in range(len(off0) - 1):
        hulls.append(xy0[off0[k]:off0[k + 1]])
        zmm.append([float(v) for v in golden0["zminmax"][k]])
        sizes.append(int((cl == k).sum()))
    off = np.concatenate([[0], np.cumsum([len(h) for h in hulls])]).astype(np.uint32)
    allxy = np.concatenate(hulls)
    boxes = ctx.bounding_boxes(allxy, off, lpl.BOX_ROTATING_CALIPERS)
    cls, area = ctx.vehicle_match(allxy, off, np.array(zmm), np.array(sizes, np.uint32), boxes)
    exp = [_vehicle_match_node(h, z, s, b) for h, z, s, b in zip(hulls, zmm, sizes, boxes)]
    assert np.array_equal(cls, np.array([e[0] for e in exp], np.int32))
    assert np.array_equal(area.view(np.uint64), np.array([e[1] for e in exp], np.float64).view(np.uint64))
    assert set(cls[:44].tolist()) >= {-1, 0} and (cls[:20] >= 0).sum() >= 4   # car-sized rectangles do match


def test_edge_cases(ctx, port):
    ctx.cluster_config(**NODE_CLUSTER_CFG)
    empty = np.zeros((0, 4), np.float32)
    assert ctx.ring_partition(empty).shape == (0,)
    assert ctx.dror_filter(empty).shape == (0,)
    assert ctx.segment(empty, np.zeros(0, np.uint16)).shape == (0,)
    cl, k = ctx.cluster(empty)
    assert cl.shape == (0,) and k == 0
    for n in (1, 2, 3, 31, 257, 2049):
        rng = np.random.default_rng(n)
        pts = np.zeros((n, 4), np.float32)
        pts[:, :3] = np.round(rng.uniform(-30, 30, (n, 3)) * [1, 1, 0.05] * 1000) / 1000
        ring = rng.integers(0, 64, n).astype(np.uint16)
        assert np.array_equal(ctx.ring_partition(pts), port.ring_partition(pts))
        assert np.array_equal(ctx.dror_filter(pts), port.dror(pts))
        assert np.array_equal(ctx.segment(pts, ring), port.segment(pts, ring))
        got, k = ctx.cluster(pts)
        assert np.array_equal(got, port.cluster(pts, **NODE_CLUSTER_CFG))
    # the point at the origin (KITTI frame 0 has one): range_xy = 0
    z = np.zeros((5, 4), np.float32)
    z[1:, 0] = [5, 5.01, 5.02, 5.03]
    assert np.array_equal(ctx.dror_filter(z), port.dror(z))
    assert np.array_equal(ctx.segment(z, np.zeros(5, np.uint16)), port.segment(z, np.zeros(5, np.uint16)))
    # more points than the context was created for -> capacity error, not a crash
    with pytest.raises(lpl.LplError) as e:
        ctx.dror_filter(np.zeros((140000, 4), np.float32))
    assert e.value.code == lpl.native.LPL_ERR_CAPACITY


def test_dror_configs_and_permutation_invariance(ctx, port, golden0):
    pts = golden0["pts"][:60000]
    for cfg in (dict(mult=0.02, min_radius=0.1, min_neighbours=4), dict(mult=0.05, min_radius=0.2, min_neighbours=8),
                dict(mult=0.01, min_radius=0.05, min_neighbours=2)):
        ctx.dror_config(**cfg)
        assert np.array_equal(ctx.dror_filter(pts), port.dror(pts, **cfg)), cfg
    ctx.dror_config()
    base = ctx.dror_filter(pts)
    perm = np.random.default_rng(0).permutation(pts.shape[0])
    assert np.array_equal(ctx.dror_filter(np.ascontiguousarray(pts[perm])), base[perm])


def test_chained_batch_ragged(ctx, port, golden0, golden100):
    frames = [golden0["pts"], F.synth_scan(4100)[0], golden100["pts"][:70001].copy(), np.zeros((0, 4), np.float32),
              golden100["pts"], golden0["pts"][:5].copy()]
    ctx.cluster_config(**NODE_CLUSTER_CFG)
    ctx.set_jcp_mode(lpl.JCP_AS_REFERENCE)
    for dror in (True, False):
        stages = lpl.STAGE_ALL if dror else (lpl.STAGE_ALL & ~lpl.STAGE_DROR)
        nf = ctx.upload(frames)
        ctx.run(nf, stages)
        ctx.sync(nf)
        bufs = lpl.BatchBuffers(8, 131072, want=[p[0] for p in lpl.BatchBuffers.PLANES])
        counts = ctx.download_batch(nf, bufs).copy()
        for f in range(nf):
            got = ctx.download(f)
            exp = parity.oracle_chain(port, frames[f], dror)
            rep = parity.chain_report(got, exp)
            assert all(v == 0 for v in rep.values()), (f, dror, rep)
            # the batched download returns the same planes as the per-frame one
            n, m, k, hv = got["n"], got["num_obstacles"], got["num_clusters"], got["num_hull_vertices"]
            assert counts[:, f].tolist() == [n, got["num_valid"], m, k, hv]
            P = {name: b.array for name, b in bufs.planes.items()}
            assert np.array_equal(P["labels_u8"][f, :n], got["labels"].astype(np.uint8))
            assert np.array_equal(P["noise"][f, :n], got["noise"])
            assert np.array_equal(P["ring"][f, :n], got["ring"])
            assert np.array_equal(P["obstacle_index"][f, :m], got["obstacle_index"])
            assert np.array_equal(P["cluster_labels"][f, :m], got["cluster_labels"])
            assert np.array_equal(P["hull_offsets"][f, :k + 1], got["hull_offsets"])
            assert np.array_equal(P["hull_xy"][f, :hv], got["hull_xy"])
            assert np.array_equal(P["zminmax"][f, :k], got["zminmax"])
        bufs.close()


def test_128_beam_spec_size_full_chain(port):
    """BASELINE.json configs[2] at the specified size: 128 beams x 2048 columns, 400 boxes + 300 poles, 1 % dropout
    (~253k points) with its ring field through the WHOLE chain - DROR, segmentation, clustering, hulls - against the
    port, frame by frame in one batch of two scenes."""
    scenes = [F.synth_scan(3000 + i, beams=128, n_boxes=400, n_poles=300, n_walls=0, dropout=0.01) for i in range(2)]
    frames = [s[0] for s in scenes]
    rings = [s[1] for s in scenes]
    assert min(f.shape[0] for f in frames) > 240_000
    c = lpl.Context(0, max_points=max(f.shape[0] for f in frames), max_frames=2, image_height=128, image_width=2048)
    cfg = c.segmenter_default_cfg()
    cfg.image_height = 128
    c.segmenter_config(cfg)
    c.cluster_config(**NODE_CLUSTER_CFG)
    port.segment_config(default_seg_cfg(image_height=128))
    try:
        nf = c.upload(frames, rings=rings)
        c.run(nf, lpl.STAGE_ALL & ~lpl.STAGE_RING)
        c.sync(nf)
        for f in range(nf):
            got = c.download(f)
            exp = parity.oracle_chain(port, frames[f], dror=True, ring=rings[f])
            rep = parity.chain_report(got, exp)
            assert all(v == 0 for v in rep.values()), (f, rep)
            assert got["num_clusters"] > 100 and got["num_hull_vertices"] > 1000
    finally:
        port.segment_config(default_seg_cfg())
        c.close()


def test_128_beam_config(port):
    """BASELINE.json configs[2]: 128 x 2048 organised frames."""
    pts, ring = F.synth_scan(3000, beams=128, n_boxes=120, n_poles=80, dropout=0.01)
    c = lpl.Context(0, max_points=pts.shape[0], max_frames=1, image_height=128, image_width=2048)
    cfg = c.segmenter_default_cfg()
    cfg.image_height = 128
    c.segmenter_config(cfg)
    port.segment_config(default_seg_cfg(image_height=128))
    try:
        exp = port.segment(pts, ring)
        got = c.segment(pts, ring)
        assert np.array_equal(got, exp)
        obs = np.ascontiguousarray(pts[exp == 2])
        c.cluster_config(**NODE_CLUSTER_CFG)
        cl, k = c.cluster(obs)
        assert np.array_equal(cl, port.cluster(obs, **NODE_CLUSTER_CFG))
    finally:
        port.segment_config(default_seg_cfg())
        c.close()


def test_image_width_not_a_multiple_of_16(port, golden0):
    """64 x 1000 range image: the dilation tile cannot be described by a TMA tensor map (row stride
    must be a multiple of 16 bytes) and is staged with plain loads instead; same results as the oracle."""
    pts, ring = golden0["pts"], golden0["ring"].astype(np.uint16)
    c = lpl.Context(0, max_points=pts.shape[0], max_frames=1, image_height=64, image_width=1000)
    cfg = c.segmenter_default_cfg()
    cfg.image_width = 1000
    c.segmenter_config(cfg)
    port.segment_config(default_seg_cfg(image_width=1000))
    try:
        exp, img_o = port.segment(pts, ring, want_image=True)
        got, img_g = c.segment(pts, ring, want_image=True)
        assert np.array_equal(got, exp)
        assert np.array_equal(img_g, img_o)
    finally:
        port.segment_config(default_seg_cfg())
        c.close()


def test_dense_polar_cells(ctx, port):
    """Thousands of points in single polar cells near the sensor: exercises the oversized-cell path of
    the RECM statistics (> 256 heights per cell) and the RANSAC rank selection when one cell holds more
    candidates than its shared-memory staging (> 1024), against the oracle."""
    rng = np.random.default_rng(11)
    parts = []
    for az_deg, r0, n in ((10.3, 3.0, 3000), (10.6, 5.0, 1500), (200.2, 2.5, 2200), (95.0, 7.0, 400)):
        a = np.deg2rad(az_deg + rng.uniform(0, 0.35, n))
        r = r0 + rng.uniform(0, 0.9, n)
        z = -1.73 + rng.normal(0, 0.03, n)
        z[rng.random(n) < 0.05] += rng.uniform(0.6, 1.5)     # a few obstacle returns inside the cells
        parts.append(np.stack([r * np.cos(a), r * np.sin(a), z], -1))
    scan, ring_scan = F.synth_scan(4003)
    xyz = np.concatenate(parts + [scan[::7, :3]])
    pts = np.zeros((xyz.shape[0], 4), np.float32)
    pts[:, :3] = np.round(xyz * 1000) / 1000
    order = rng.permutation(pts.shape[0])
    pts = np.ascontiguousarray(pts[order])
    ring = rng.integers(0, 64, pts.shape[0]).astype(np.uint16)
    exp, img_o, dbg_o = port.segment(pts, ring, want_image=True, want_debug=True)
    got, img_g = ctx.segment(pts, ring, want_image=True)
    dbg_g = ctx.debug_segment(0)
    assert dbg_g["n_candidates"] == dbg_o["n_candidates"] and dbg_o["n_candidates"] > 5000
    assert np.array_equal(dbg_o["plane"].view(np.uint32), dbg_g["plane"].view(np.uint32))
    assert np.array_equal(dbg_o["elevation"].view(np.uint32), dbg_g["elevation"].view(np.uint32))
    assert np.array_equal(got, exp) and np.array_equal(img_g, img_o)
    assert np.array_equal(ctx.segment(pts, None), port.segment(pts, None))


def test_jcp_whole_rows_queued(ctx, port):
    """Concentric walls: boundaries run along entire image rows, so the JCP queue holds whole rows of
    consecutive pixels (2048-long dependency chains, several chunks of the row scan with the carried
    state), with and without the ring field; labels and image against the oracle."""
    pts, ring = F.synth_ring_walls(7)
    exp, img_o, dbg_o = port.segment(pts, ring, want_image=True, want_debug=True)
    assert dbg_o["n_queued"] > 30000
    got, img_g = ctx.segment(pts, ring, want_image=True)
    assert ctx.debug_segment(0)["n_queued"] == dbg_o["n_queued"]
    assert np.array_equal(got, exp) and np.array_equal(img_g, img_o)
    assert np.array_equal(ctx.segment(pts, None), port.segment(pts, None))
    pts2, ring2 = F.synth_ring_walls(8, radii=(5.0, 5.6, 7.5, 11.0), height=1.2)
    assert np.array_equal(ctx.segment(pts2, ring2), port.segment(pts2, ring2))


def test_unorganized_2m_cloud_properties(port):
    """BASELINE.json configs[4] at reduced size against the oracle, and at full size (2 M points)
    through size-independent properties: idempotence and DROR monotonicity in the radius."""
    small = F.synth_unorganized(5000, n=200_000, n_blobs=500)
    c = lpl.Context(0, max_points=2_000_000, max_frames=1)
    try:
        assert np.array_equal(c.dror_filter(small), port.dror(small))
        cl, k = c.cluster(small)
        exp = port.cluster(small, range_m=0.4, az_deg=1.0, el_deg=1.5, min_size=3)
        assert np.array_equal(cl, exp)
        off, xy, idx, zmm = c.cluster_hulls(small, cl)
        eo, exy, eidx, ezmm = port.cluster_hulls(small, exp)
        assert np.array_equal(off, eo) and np.array_equal(xy, exy.astype(np.float32))
        big = F.synth_unorganized(5001)
        a = c.dror_filter(big)
        assert np.array_equal(a, c.dror_filter(big))                 # idempotent / deterministic
        c.dror_config(mult=0.04, min_radius=0.2, min_neighbours=4)
        b = c.dror_filter(big)
        assert int(((b == 1) & (a == 0)).sum()) == 0                 # a larger radius never adds noise
        cl1, k1 = c.cluster(big)
        cl2, k2 = c.cluster(big)
        assert k1 == k2 and np.array_equal(cl1, cl2)
        # canonical labelling: cluster ids are ranked by their minimum point index
        first = np.full(k1, big.shape[0], np.int64)
        np.minimum.at(first, cl1[cl1 >= 0], np.flatnonzero(cl1 >= 0))
        assert np.all(np.diff(first) > 0)
        sizes = np.bincount(cl1[cl1 >= 0], minlength=k1)
        assert sizes.min() >= 3
    finally:
        c.close()


def test_unorganized_2m_cloud_full_size_parity(port):
    """BASELINE.json configs[4] at FULL size against the port (which has no voxel cap): the chained
    ring-less pipeline on the 2,000,000-point cloud (DROR mask, labels, cluster partition, hulls), and - the
    part of the config that stresses thousands of small clusters - fine-voxel clustering + hulls + oriented
    boxes of the 735k DROR-valid non-ground points (5,877 clusters, 333,739 voxels: above the reference's
    200k-voxel table and above the shared-memory union-find, so the global path runs)."""
    big = F.synth_unorganized(5000)
    c = lpl.Context(0, max_points=big.shape[0], max_frames=1)
    try:
        c.cluster_config(**NODE_CLUSTER_CFG)
        nf = c.upload([big])
        c.run(nf, lpl.STAGE_ALL & ~lpl.STAGE_RING)
        c.sync(nf)
        got = c.download(0)
        exp = parity.oracle_chain(port, big, dror=True, ring=None)
        rep = parity.chain_report(got, exp, skip=("ring",))
        assert all(v == 0 for v in rep.values()), rep
        assert int(exp["noise"].sum()) > 10_000 and exp["num_clusters"] > 300
        sub = np.ascontiguousarray(big[(exp["noise"] == 0) & (big[:, 2] > -1.55)])
        fine = dict(range_m=0.2, az_deg=0.25, el_deg=0.5, min_size=3)
        c.cluster_config(**fine)
        cl, k = c.cluster(sub)
        ecl = port.cluster(sub, **fine)
        assert port.last_num_voxels > 200_000 and k > 5000
        assert np.array_equal(cl, ecl)
        off, xy, idx, zmm = c.cluster_hulls(sub, cl)
        eo, exy, eidx, ezmm = port.cluster_hulls(sub, ecl)
        assert np.array_equal(off, eo) and np.array_equal(xy, exy.astype(np.float32))
        assert np.array_equal(zmm, ezmm.astype(np.float32))
        boxes = c.bounding_boxes(exy, eo, lpl.BOX_ROTATING_CALIPERS)
        sel = np.random.default_rng(0).choice(k, 400, replace=False)
        ebox = np.stack([port.bounding_box(exy[eo[j]:eo[j + 1]], lpl.BOX_ROTATING_CALIPERS) for j in sel])
        _boxes_equal(boxes[sel], ebox, exact=True)
    finally:
        c.close()


@pytest.mark.skipif(not F.have_pack(), reason="data/kitti154.npz not built")
def test_full_sequence_chained_with_dror_vs_reference():
    """BASELINE.json configs[1] at full size, the whole chain INCLUDING DROR as one 154-frame batch: every
    frame's DROR mask, labels, cluster labels, hull offsets / vertices and z extents hash to what the
    unmodified reference (exact DROR semantics) produced (tests/golden/kitti154_chain_dror.json,
    tools/make_golden_chain.py)."""
    import hashlib

    def sha(a, dt):
        return hashlib.sha1(np.ascontiguousarray(a, dtype=dt).tobytes()).hexdigest()

    gold = json.load(open(os.path.join(F.GOLDEN_DIR, "kitti154_chain_dror.json")))["frames"]
    frames = F.load_pack()
    assert len(gold) == len(frames) == 154
    c = lpl.Context(0, max_points=max(f.shape[0] for f in frames), max_frames=len(frames))
    try:
        c.cluster_config(**NODE_CLUSTER_CFG)
        xyz = np.ascontiguousarray(np.concatenate(frames)[:, :3])
        nf = c.upload_packed_xyz(xyz, [f.shape[0] for f in frames])   # the bench's e2e upload path
        bufs = lpl.PackedBuffers(nf, nf * 131072 * 8, want=("labels_u8", "noise", "cluster_labels", "hull_offsets",
                                                             "hull_xy", "zminmax"))
        # one stream, then the batch as 3 and as 2 concurrent sub-batches (lpl_pipeline_use_split); every variant runs
        # plainly (first run), as a freshly captured graph and as a replayed graph
        for parts in (1, 3, 2):
            c.use_split(parts)
            for rep in range(3):
                c.run(nf, lpl.STAGE_ALL)
                counts = c.download_packed(nf, bufs)
                bad = []
                for f, g in enumerate(gold):
                    ok = (counts[0, f] == g["n"] and counts[3, f] == g["clusters"] and counts[4, f] == g["hull_vertices"]
                          and sha(bufs.frame("noise", f), np.uint8) == g["noise_sha1"]
                          and sha(bufs.frame("labels_u8", f), np.uint8) == g["labels_sha1"]
                          and sha(bufs.frame("cluster_labels", f), np.int32) == g["cluster_sha1"]
                          and sha(bufs.frame("hull_offsets", f), np.uint32) == g["hull_offsets_sha1"]
                          and sha(bufs.frame("hull_xy", f), np.float32) == g["hull_xy_sha1"]
                          and sha(bufs.frame("zminmax", f), np.float32) == g["zminmax_sha1"])
                    if not ok:
                        bad.append(f)
                assert not bad, (parts, rep, bad)
        bufs.close()
    finally:
        c.close()


@pytest.mark.skipif(not F.have_pack(), reason="data/kitti154.npz not built")
def test_full_sequence_batch_vs_reference_summary():
    """BASELINE.json configs[1] at full size: all 154 frames as ONE batch; labels / clusters / hulls
    hash to what the unmodified reference produced frame by frame (tests/golden/kitti154_summary.json)."""
    summ = json.load(open(os.path.join(F.GOLDEN_DIR, "kitti154_summary.json")))["frames"]
    frames = F.load_pack()
    c = lpl.Context(0, max_points=max(f.shape[0] for f in frames), max_frames=len(frames))
    try:
        c.cluster_config(**NODE_CLUSTER_CFG)
        nf = c.upload(frames)
        c.run(nf, lpl.STAGE_ALL & ~lpl.STAGE_DROR)   # the summary chain has no DROR (the node never wires it)
        c.sync(nf)
        for f, s in enumerate(summ):
            got = c.download(f)
            assert label_hash(got["ring"]) == s["ring_hash"], f
            assert label_hash(got["labels"]) == s["label_hash"], f
            assert got["num_clusters"] == s["clusters"], f
            assert label_hash(got["cluster_labels"].astype(np.int64).astype(np.uint32)) == s["cluster_hash"], f
            assert label_hash(got["hull_offsets"]) == s["hull_offsets_hash"], f
            assert label_hash(got["hull_xy"].view(np.uint32).reshape(-1)) == s["hull_xy_hash"], f
        # DROR stage-wise on the raw clouds
        c.run(nf, lpl.STAGE_DROR)
        c.sync(nf)
        for f, s in enumerate(summ):
            got = c.download(f)
            assert int(got["noise"].sum()) == s["dror_noise_exact"], f
            assert label_hash(got["noise"]) == s["dror_hash"], f
    finally:
        c.close()


def test_frame_pipeline_stream(port, golden0, golden100):
    """The double-buffered two-stream pipeline returns the same per-frame results as a single context."""
    from lidar_processing_v2_b200.stream import run_stream

    frames = [golden0["pts"], golden100["pts"], golden0["pts"][:60000].copy(), golden100["pts"][:90000].copy(),
              golden0["pts"]]
    stats = run_stream(frames, device=0, batch=2)
    exp = [parity.oracle_chain(port, f, dror=True) for f in frames]
    assert stats["frames"] == len(frames) and stats["world"] == 1
    assert stats["points"] == sum(f.shape[0] for f in frames)
    assert stats["obstacles"] == sum(int(e["obstacle_index"].shape[0]) for e in exp)
    assert stats["clusters"] == sum(e["num_clusters"] for e in exp)
    assert stats["hull_vertices"] == sum(int(e["hull_xy"].shape[0]) for e in exp)
