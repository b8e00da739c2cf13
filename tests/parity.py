"""Shared comparison helpers: CUDA path (through the C ABI) vs the CPU oracle, stage by stage.

Used by tests/test_gpu_parity.py, tools/gpu_diag.py and __graft_entry__.smoke(). Everything that
touches oracle/ lives here or in tests/ - never in the product package.
"""
from __future__ import annotations

import numpy as np

from oracle.oracle import JCP_AS_IS, JCP_CLEAN, NODE_CLUSTER_CFG, PortOracle, default_seg_cfg

import lidar_processing_v2_b200 as lpl


def seg_cfg_to_lpl(cfg) -> lpl.SegmenterCfg:
    out = lpl.SegmenterCfg()
    for name, _ in lpl.SegmenterCfg._fields_:
        setattr(out, name, getattr(cfg, name))
    return out


def stage_report(ctx: lpl.Context, oracle, pts: np.ndarray, ring_given=None, node_cluster_cfg=True,
                 jcp_mode=JCP_AS_IS, check_ringless=True) -> dict:
    """Run every stage on the GPU with the oracle's upstream output as input and count mismatches.

    `oracle` is a PortOracle or RefOracle (same method names). Returns a flat dict of counters;
    all "*_diff" entries must be zero for parity.
    """
    rep = {}
    port = oracle if isinstance(oracle, PortOracle) else PortOracle()
    n = pts.shape[0]
    # ---- stage 0: ring partition
    ring_o = port.ring_partition(pts)
    ring_g = ctx.ring_partition(pts)
    rep["ring_diff"] = int((ring_o != ring_g).sum())
    ring = ring_o if ring_given is None else ring_given
    # ---- stage 1: DROR (exact semantics)
    d_o = oracle.dror(pts, mode="exact") if not isinstance(oracle, PortOracle) else oracle.dror(pts)
    d_g = ctx.dror_filter(pts)
    rep["dror_diff"] = int((d_o != d_g).sum())
    rep["dror_noise"] = int(d_o.sum())
    # ---- stage 2: segmentation (with ring)
    ctx.set_jcp_mode(lpl.JCP_AS_REFERENCE if jcp_mode == JCP_AS_IS else lpl.JCP_CLEAN)
    if isinstance(oracle, PortOracle):
        l_o, img_o, dbg_o = oracle.segment(pts, ring, jcp_mode=jcp_mode, want_image=True, want_debug=True)
    else:
        assert jcp_mode == JCP_AS_IS, "the unmodified reference only has the as-is JCP behaviour"
        l_o, img_o = oracle.segment(pts, ring, want_image=True)
        dbg_o = oracle.segment_intermediates()
    l_g, img_g = ctx.segment(pts, ring, want_image=True)
    dbg_g = ctx.debug_segment(0)
    rep["seg_diff"] = int((l_o != l_g).sum())
    rep["seg_image_diff"] = int((img_o != img_g).any(axis=2).sum())
    rep["seg_elev_diff"] = int((dbg_o["elevation"] != dbg_g["elevation"]).sum())
    rep["seg_counts"] = [int((l_o == k).sum()) for k in range(3)]
    if "plane" in dbg_o:
        rep["seg_plane_diff"] = int((dbg_o["plane"].view(np.uint32) != dbg_g["plane"].view(np.uint32)).sum())
        rep["seg_cand_diff"] = int(dbg_o["n_candidates"] != dbg_g["n_candidates"])
        rep["seg_queue_diff"] = int(dbg_o["n_queued"] != dbg_g["n_queued"])
    rep["seg_rounds"] = dbg_g["rounds"]
    rep["seg_status"] = dbg_g["status"]
    if check_ringless:
        if isinstance(oracle, PortOracle):
            l2_o = oracle.segment(pts, None, jcp_mode=jcp_mode)
        else:
            l2_o = oracle.segment(pts, None)
        l2_g = ctx.segment(pts, None)
        rep["seg_noring_diff"] = int((l2_o != l2_g).sum())
    # ---- stage 3: clustering of the oracle's obstacle cloud
    obs = np.ascontiguousarray(pts[l_o == 2])
    ccfg = NODE_CLUSTER_CFG if node_cluster_cfg else dict(range_m=0.4, az_deg=1.0, el_deg=1.5, min_size=3)
    ctx.cluster_config(**ccfg)
    if isinstance(oracle, PortOracle):
        c_o, dims_o = oracle.cluster(obs, want_dims=True, **ccfg)
    else:
        oracle.cluster_config(**ccfg)
        c_o, dims_o = oracle.cluster(obs, want_dims=True)
    c_g, k_g = ctx.cluster(obs)
    rep["clu_diff"] = int((c_o != c_g).sum())
    rep["clu_dims_diff"] = int((dims_o != ctx.debug_cluster(0)).sum()) if obs.shape[0] else 0
    rep["clu_k"] = [int(c_o.max() + 1) if c_o.size else 0, int(k_g)]
    # ---- stage 4: hulls of the oracle's clusters
    off_o, xy_o, idx_o, zmm_o = port.cluster_hulls(obs, c_o)
    off_g, xy_g, idx_g, zmm_g = ctx.cluster_hulls(obs, c_o)
    rep["hull_off_diff"] = int((off_o != off_g).sum()) if off_o.shape == off_g.shape else -1
    if xy_o.shape == xy_g.shape:
        rep["hull_xy_diff"] = int((xy_o.astype(np.float32) != xy_g).any(axis=1).sum())
        # reported indices must address the same coordinates (duplicates may pick another index)
        rep["hull_idx_coord_diff"] = int((obs[idx_g, :2] != xy_g).any(axis=1).sum()) if idx_g.size else 0
    else:
        rep["hull_xy_diff"] = -1
        rep["hull_idx_coord_diff"] = -1
    rep["hull_z_diff"] = int((zmm_o.astype(np.float32) != zmm_g).sum()) if zmm_o.shape == zmm_g.shape else -1
    rep["hull_vertices"] = int(xy_o.shape[0])
    return rep


def parity_ok(rep: dict) -> bool:
    return all(v == 0 for k, v in rep.items() if k.endswith("_diff")) and rep.get("seg_status", 0) == 0


def oracle_chain(port: PortOracle, pts, dror: bool, jcp_mode=JCP_AS_IS, cluster_cfg=None, ring="partition"):
    """Chained CPU pipeline: ring -> [DROR -> compaction] -> segment -> obstacle
    compaction -> cluster -> hulls. Returns a dict in input-index space like Context.download.
    ring: "partition" (stage 0 from the point order), an array (the cloud's own ring field) or None
    (ring-less point type: height index from the elevation angle)."""
    ccfg = cluster_cfg or NODE_CLUSTER_CFG
    n = pts.shape[0]
    if isinstance(ring, str):
        ring = port.ring_partition(pts)
    noise = port.dror(pts) if dror else np.zeros(n, np.uint8)
    keep = np.flatnonzero(noise == 0)
    lv = port.segment(np.ascontiguousarray(pts[keep]), None if ring is None else np.ascontiguousarray(ring[keep]),
                      jcp_mode=jcp_mode)
    if ring is None:
        ring = np.zeros(n, np.uint16)
    labels = np.zeros(n, np.uint32)
    labels[keep] = lv
    obs_idx = keep[lv == 2]
    obs = np.ascontiguousarray(pts[obs_idx])
    cl = port.cluster(obs, **ccfg)
    off, xy, hidx, zmm = port.cluster_hulls(obs, cl)
    return dict(ring=ring, noise=noise, labels=labels, obstacle_index=obs_idx.astype(np.uint32),
                cluster_labels=cl, hull_offsets=off, hull_xy=xy.astype(np.float32), zminmax=zmm.astype(np.float32),
                num_clusters=int(cl.max() + 1) if cl.size else 0)


def chain_report(got: dict, exp: dict, skip=()) -> dict:
    rep = {}
    for k in ("ring", "noise", "labels", "obstacle_index", "cluster_labels", "hull_offsets"):
        if k in skip:
            continue
        a, b = np.asarray(got[k]), np.asarray(exp[k])
        rep[k + "_diff"] = int((a != b).sum()) if a.shape == b.shape else -1
    a, b = np.asarray(got["hull_xy"]), np.asarray(exp["hull_xy"])
    rep["hull_xy_diff"] = int((a != b).sum()) if a.shape == b.shape else -1
    # z extents bit for bit: the reference keeps the first of equal extremes, observable as the sign of a zero
    a, b = np.asarray(got["zminmax"], np.float32), np.asarray(exp["zminmax"], np.float32)
    rep["zminmax_diff"] = int((a.view(np.uint32) != b.view(np.uint32)).sum()) if a.shape == b.shape else -1
    return rep
