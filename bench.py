#!/usr/bin/env python
"""Benchmark of the per-frame LiDAR perception hot path on B200 (driver contract: see README/DESIGN).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
    python -m torch.distributed.run --nproc-per-node N ... bench.py --gpus N ...

A *step* is one pass of the whole hot path (ring partition -> DROR -> JCP/RECM segmentation ->
curved-voxel clustering -> per-cluster hulls) over one batch of frames on every rank:

  workload "kitti154"  BASELINE.json configs[1]: the reference's 154 KITTI HDL-64E frames
                       (18.7 M points, 300 MB of float4 input > the 126 MB L2, so no L2 flush is needed)
                       as one batch; every rank runs its own copy (weak scaling, no collective on
                       the data path - NCCL only reduces the timing).
  workload "synth64"   BASELINE.json configs[3]: seeded synthetic HDL-64E sweeps (also the fallback when
                       data/kitti154.npz is absent); "synth128" = configs[2] (128 beams x 2048 columns with a
                       ring field); "cloud2m" = configs[4] (unorganised 2 M-point clouds, ring-less chain).
                       Select with --workload; the default and the headline is kitti154.

`value`  = frames/s with the inputs already resident in HBM (CUDA events on the context stream).
`e2e`    = frames/s through the C ABI with HOST buffers: every step uploads the batch from pinned
           host memory (one packed transfer), runs the pipeline and reads labels / clusters / hulls
           back; --e2e-ctx contexts (streams, default 4) rotate so copies overlap compute.
`roofline` describes the dominant kernel (largest share of the step) from per-kernel CUDA events
recorded inside the timed region; `cpu_baseline` / `--impl reference` time the reference's own CPU
code (oracle/_ref, compiled from the unmodified sources) on the host cores.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "frames/s (HDL-64E ~120k pts/frame: ring partition + DROR + JCP segmentation + curved-voxel clustering + hulls)"
UNIT = "frames/s"


# --------------------------------------------------------------------------------------------
# workload
# --------------------------------------------------------------------------------------------
WORKLOADS = ("kitti154", "synth64", "synth128", "cloud2m")


def load_frames(limit=None, workload=None):
    """-> (frames, workload name, description, options). `workload` None = kitti154 when the pack is there.
    options: image_height (Context + segmenter configuration), stages (None = the node's whole chain),
    cpu (False: the reference CPU arm cannot run it)."""
    from tools import frames as F

    opts = {"image_height": 64, "stages": None, "cpu": True, "rings": None}
    if workload in (None, "kitti154") and F.have_pack():
        fr = F.load_pack(limit=limit)
        return fr, "kitti154", "KITTI HDL-64E frames of the reference repository (data/*.pcd, repacked)", opts
    if workload == "synth128":
        # BASELINE.json configs[2]: 128 beams x 2048 columns, 400 boxes + 300 poles, 1 % dropout, seed 3000 + frame
        n = limit or 64
        base = [F.synth_scan(3000 + i, beams=128, n_boxes=400, n_poles=300, n_walls=0, dropout=0.01)
                for i in range(min(n, 16))]
        opts["image_height"] = 128
        # the reference's ring partition (dataloader.cpp:68-137) is written for 64 rings: a 128-beam cloud
        # arrives with its ring field (pcl::PointXYZIR input of Segmenter::segment) and skips that stage
        opts["stages"] = "ring_field"
        opts["rings"] = [base[i % len(base)][1] for i in range(n)]
        return ([base[i % len(base)][0] for i in range(n)], "synth128",
                f"synthetic 128-beam x 2048-column sweeps with a ring field (seeded ray-cast scenes, {len(base)} "
                f"distinct scenes cycled); chain: DROR + segmentation + clustering + hulls", opts)
    if workload == "cloud2m":
        # BASELINE.json configs[4]: 2 M-point unorganised clouds (no ring structure: ring-less segmentation)
        n = limit or 5
        opts["stages"] = "ringless"
        opts["cpu"] = False  # the reference Clusterer's 200k-voxel table overflows on this cloud
        return ([F.synth_unorganized(5000 + i) for i in range(n)], "cloud2m",
                "synthetic unorganised 2,000,000-point clouds (ground disc + 5,000 blobs + 5 % background), "
                "ring-less chain: DROR + segmentation + clustering + hulls", opts)
    # BASELINE.json configs[3]: synthetic HDL-64E sweeps, 6 % dropout, 60 boxes + 40 poles + 2 walls, seed 4000 + frame
    n = limit or 154
    base = [F.synth_scan(4000 + i)[0] for i in range(min(n, 32))]
    why = "" if workload == "synth64" else "; data/kitti154.npz absent"
    return ([base[i % len(base)] for i in range(n)], "synth64",
            f"synthetic HDL-64E sweeps (seeded ray-cast scenes, {len(base)} distinct scenes cycled{why})", opts)


# --------------------------------------------------------------------------------------------
# clocks
# --------------------------------------------------------------------------------------------
class ClockSampler:
    """SM clock and throttle reasons sampled DURING the timed region: an NVML polling thread (every
    2 ms: the timed region of the default run is ~50 ms, too short for `nvidia-smi -lms`), with the
    nvidia-smi loop of B200_PROFILING.md as the fallback when NVML cannot be loaded."""

    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")
    REASON_BITS = {0x8: "hw_slowdown", 0x40: "hw_thermal_slowdown", 0x20: "sw_thermal_slowdown", 0x4: "sw_power_cap",
                   0x80: "hw_power_brake_slowdown"}

    def __init__(self, index: int):
        self.index = index
        self.rows = []
        self.proc = None
        self.nvml = None
        self.handle = None
        self.samples = []  # (sm_mhz, reasons bitmask)
        self.stop_flag = threading.Event()
        self.th = None
        try:
            import pynvml

            pynvml.nvmlInit()
            handle = None
            try:
                import torch

                uuid = str(torch.cuda.get_device_properties(index).uuid)
                handle = pynvml.nvmlDeviceGetHandleByUUID(("GPU-" + uuid) if not uuid.startswith("GPU-") else uuid)
            except Exception:
                handle = pynvml.nvmlDeviceGetHandleByIndex(index)
            pynvml.nvmlDeviceGetClockInfo(handle, pynvml.NVML_CLOCK_SM)
            self.nvml, self.handle = pynvml, handle
        except Exception:
            self.nvml = None

    def _poll(self):
        nv, h = self.nvml, self.handle
        get_reasons = getattr(nv, "nvmlDeviceGetCurrentClocksEventReasons", None) or nv.nvmlDeviceGetCurrentClocksThrottleReasons
        while not self.stop_flag.is_set():
            try:
                self.samples.append((float(nv.nvmlDeviceGetClockInfo(h, nv.NVML_CLOCK_SM)), int(get_reasons(h))))
            except Exception:
                pass
            time.sleep(0.002)

    def _sample(self):
        nv, h = self.nvml, self.handle
        get_reasons = getattr(nv, "nvmlDeviceGetCurrentClocksEventReasons", None) or nv.nvmlDeviceGetCurrentClocksThrottleReasons
        try:
            self.samples.append((float(nv.nvmlDeviceGetClockInfo(h, nv.NVML_CLOCK_SM)), int(get_reasons(h))))
        except Exception:
            pass

    def start(self):
        if self.nvml is not None:
            self._sample()  # first calls are slow: take them before the timed region
            self.samples.clear()
            self.th = threading.Thread(target=self._poll, daemon=True)
            self.th.start()
            return
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "100",
                 "-i", str(self.index)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.th = threading.Thread(target=self._read, daemon=True)
            self.th.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append(line.strip())

    def stop(self) -> dict:
        if self.nvml is not None:
            self.stop_flag.set()
            self.th.join(timeout=2)
            sm = [x[0] for x in self.samples]
            bits = 0
            for x in self.samples:
                bits |= x[1]
            try:
                mx = float(self.nvml.nvmlDeviceGetMaxClockInfo(self.handle, self.nvml.NVML_CLOCK_SM))
            except Exception:
                mx = None
            return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": mx,
                    "reasons": sorted(n for b, n in self.REASON_BITS.items() if bits & b),
                    "samples": len(sm), "source": "nvml, 2 ms polling inside the timed region"}
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            t = [x.strip() for x in r.split(",")]
            if len(t) < 7:
                continue
            try:
                sm.append(float(t[0]))
                mx.append(float(t[1]))
            except ValueError:
                continue
            for nm, v in zip(names, t[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(nm)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm), "source": "nvidia-smi -lms 100"}


# --------------------------------------------------------------------------------------------
# roofline bookkeeping: algorithmic bytes per kernel (DESIGN.md "Kernels and their rooflines")
# --------------------------------------------------------------------------------------------
def algorithmic_bytes(name: str, s: dict) -> float:
    """Compulsory HBM bytes of one launch of kernel `name` over the batch described by `s`
    (N input points, V DROR-valid points, NB binned points, M obstacle points, K clusters,
    HV hull vertices, PX pixels, CELLS polar cells, Q queued pixels, C RANSAC candidates,
    U DROR-unresolved points, F frames)."""
    N, V, NB, M, K, HV = s["N"], s["V"], s["NB"], s["M"], s["K"], s["HV"]
    NH = s.get("NH", M)  # points that survive the octagon filter and enter the hull sort
    PX, CELLS, Q, C, U, F = s["PX"], s["CELLS"], s["Q"], s["C"], s["U"], s["F"]
    t = {
        "ring_count": 16 * N, "ring_write": 16 * N + 2 * N,
        "dror_near": 16 * N + 1 * N + 4 * U,
        "dror_mark": 20 * U + 16384 * F, "dror_grid_count": 16 * N, "dror_grid_scan": 8 * 131072 * F,
        "dror_grid_scatter": 16 * N + 16 * U * 8,
        "dror_query": 20 * U + U,
        "seg_bin": (16 + 2 + 1) * N + 12 * N, "seg_cell_scan": 8 * CELLS, "seg_scatter": (16 + 8) * N + 8 * NB,
        "seg_cell": 8 * NB + 8 * CELLS, "seg_elev": 12 * CELLS,
        "seg_label": (16 + 4) * N + 4 * NB + 1 * N + 16 * C,
        "ransac_draw": 1024 * F + 8 * CELLS, "ransac_plane": 120 * 64 * F, "ransac_count": 16 * C,
        "seg_image": (16 + 4 + 4 + 1) * N + 1 * NB + 8 * NB, "seg_px": 8 * PX + 17 * PX,  # + 17 B for every pixel that has a winner (not counted)
        "seg_dilate": 2 * PX, "jcp_queue": 2 * PX + 4 * Q,
        "jcp_pre": Q * (25 * 17 + 24 * 4 + 8), "jcp_resolve": PX * 1 + Q * (8 + 4 + 96) + Q * 1 + 8 * Q // 3,
        "jcp_rows": PX * 1 + Q * (8 + 4 + 96) + Q * 1,
        "take_obstacles": 1 * N + 16 * M + 20 * M,
        "clu_sph": 16 * M + 16 * M, "clu_insert": 16 * M + 4 * M, "clu_edges": 8 * M + 52 * M // 3,
        "clu_union_sm": 52 * M // 3 + 8 * M // 3, "clu_union": 8 * M, "clu_flatten": 8 * M,
        "clu_rank": 8 * M, "clu_labels": 8 * M + 16 * M,
        # hulls: the stage as a whole must read every obstacle point's (x, y) and label once and write
        # the vertices (8M + 4M + 4 Hv + 12 K); the chain kernels are charged with that
        "hull_octagon": 64 * K, "hull_keep": (16 + 4) * M + 16 * NH, "hull_seg_scan": 8 * K,
        "hull_tilesort": 32 * NH, "hull_merge": 32 * NH,
        "hull_thin_big": 16 * NH + 4 * HV + 12 * K, "hull_thin": 16 * NH + 4 * HV + 12 * K,
        "hull_final": 16 * HV + 12 * K, "hull_off_scan": 8 * K, "hull_gather": 4 * HV + 16 * HV + 12 * HV,
        "obb_frames": 8 * HV + 80 * K,
        "label_count": 4 * M,
    }
    return float(t.get(name, 0.0))


def batch_stats(ctx, nf, seg_dbg=True) -> dict:
    s = dict(N=0, V=0, NB=0, M=0, K=0, HV=0, Q=0, C=0, U=0, NH=0, F=nf, PX=nf * ctx.H * ctx.W, CELLS=0)
    for f in range(nf):
        r = ctx.counts(f)
        s["N"] += r.n
        s["V"] += r.num_valid
        s["M"] += r.num_obstacles
        s["K"] += r.num_clusters
        s["HV"] += r.num_hull_vertices
        if seg_dbg:
            d = ctx.debug_counters(f)
            s["NB"] += d["n_binned"]
            s["C"] += d["n_candidates"]
            s["Q"] += d["n_queued"]
            s["U"] += d["n_unresolved"]
            s["CELLS"] += d["cells"]
            s["NH"] += ctx.debug_hulls(f)["n_hull_sort"]
    return s


# --------------------------------------------------------------------------------------------
# reference CPU path (oracle/_ref = the reference's own sources; port only for the two stages the
# reference keeps outside the library: ring partition and the per-cluster gather + hull call)
# --------------------------------------------------------------------------------------------
_W = {}


def _cpu_worker_init(limit, workload=None):
    from oracle.oracle import PortOracle, RefOracle, default_seg_cfg, have_ref

    _W["frames"], _, _, opts = load_frames(limit, workload)
    _W["rings"] = opts["rings"]
    _W["ref"] = RefOracle() if have_ref() else None
    _W["port"] = PortOracle()
    if opts["image_height"] != 64:
        for o in (_W["ref"], _W["port"]):
            if o is not None:
                o.segment_config(default_seg_cfg(image_height=opts["image_height"]))


def _cpu_worker_frame(i):
    """Whole hot path on frame i, chained as in DESIGN.md (ring -> DROR -> segment VALID ->
    cluster OBSTACLE -> hulls), through the reference's own code where it is a library function."""
    ref, port, pts = _W["ref"], _W["port"], _W["frames"][i]
    ring = port.ring_partition(pts) if _W.get("rings") is None else _W["rings"][i]
    noise = ref.dror(pts, mode="as_is") if ref is not None else port.dror(pts)
    keep = noise == 0
    pv = np.ascontiguousarray(pts[keep])
    rv = np.ascontiguousarray(ring[keep])
    labels = ref.segment(pv, rv) if ref is not None else port.segment(pv, rv)
    obs = np.ascontiguousarray(pv[labels == 2])
    cl = ref.cluster(obs) if ref is not None else port.cluster(obs)
    off, xy, idx, zmm = port.cluster_hulls(obs, cl)
    return int(off[-1]) if off.size else 0


class CpuReference:
    """The reference's CPU path on `procs` worker processes (the library is single-threaded and
    its objects are not thread-safe, so one process per core with its own instances)."""

    def __init__(self, procs: int, limit: int, workload=None):
        import multiprocessing as mp

        from oracle.oracle import build, have_ref

        build()
        self.kind = "reference" if have_ref() else "port"
        self.procs = procs
        self.pool = mp.get_context("spawn").Pool(procs, initializer=_cpu_worker_init, initargs=(limit, workload))
        self.pool.map(_cpu_worker_frame, list(range(min(limit, procs))))  # warm-up: imports, page faults

    def run(self, idx) -> float:
        t0 = time.perf_counter()
        self.pool.map(_cpu_worker_frame, list(idx), chunksize=1)
        return time.perf_counter() - t0

    def close(self):
        self.pool.close()
        self.pool.join()


def host_cores() -> int:
    try:
        return max(1, len(os.sched_getaffinity(0)))
    except Exception:
        return max(1, os.cpu_count() or 1)


def run_reference_arm(args, rank, world):
    if rank != 0:
        return 0
    frames, workload, data_desc, opts = load_frames(args.frames, args.workload)
    if not opts["cpu"]:
        print(json.dumps({"impl": "reference", "unavailable":
                          f"workload {workload}: the reference Clusterer's fixed 200k-voxel table overflows"}), flush=True)
        return 0
    cores = host_cores()
    per_step = min(len(frames), 2 * cores)
    cpu = CpuReference(cores, per_step, args.workload)
    for _ in range(min(args.warmup, 1)):
        cpu.run(range(per_step))
    t = 0.0
    for _ in range(args.steps):
        t += cpu.run(range(per_step))
    cpu.close()
    fps = per_step * args.steps / t
    desc = (f"{per_step} frames of {workload} per step through the reference's own CPU code "
            f"(oracle/_ref: unmodified segmenter/clusterer/noise_remover sources; ring partition and hull "
            f"gather restated) on {cores} worker processes")
    line = {
        "impl": "reference", "metric": METRIC, "value": fps, "unit": UNIT, "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * t / args.steps,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
        "data": data_desc,
        "config": {"workload": workload, "frames_per_step": per_step, "stages": "ring+dror+segment+cluster+hulls",
                   "host_threads": cores},
        "cpu_baseline": {"value": fps, "unit": UNIT, "cores": cores, "kind": cpu.kind, "sample": desc},
        "e2e": {"value": fps, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)
    return 0


# --------------------------------------------------------------------------------------------
# our arm
# --------------------------------------------------------------------------------------------
def run_ours(args, rank, local_rank, world):
    import torch

    import lidar_processing_v2_b200 as lpl

    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device - the hot path has no CPU fallback")
    torch.cuda.set_device(local_rank)
    dist = None
    if world > 1:
        import torch.distributed as dist_

        dist = dist_
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))

    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    frames, workload, data_desc, opts = load_frames(args.frames, args.workload)
    # every rank runs the same number of frames; rotate the sequence so ranks do not share inputs
    rot = (rank * 19) % len(frames)
    frames = frames[rot:] + frames[:rot]
    rings = opts["rings"]
    if rings is not None:
        rings = rings[rot:] + rings[:rot]
    nf = len(frames)
    max_pts = max(f.shape[0] for f in frames)
    stages = lpl.STAGE_ALL
    if opts["stages"] in ("ringless", "ring_field"):
        stages = lpl.STAGE_ALL & ~lpl.STAGE_RING
    img_h = opts["image_height"]

    def make_ctx(max_frames):
        c = lpl.Context(local_rank, max_points=max_pts, max_frames=max_frames, image_height=img_h)
        if img_h != 64:
            cfg = c.segmenter_default_cfg()
            cfg.image_height = img_h
            c.segmenter_config(cfg)
        c.cluster_config(range_m=0.4, az_deg=1.0, el_deg=3.0, min_size=3)  # processor.param.yaml:31-35
        return c

    ctx = make_ctx(nf)
    total_pts = sum(f.shape[0] for f in frames)

    # ---- device-resident throughput ("value")
    ctx.upload(frames, rings=rings)
    ctx.sync(nf)
    for _ in range(args.warmup):
        ctx.run(nf, stages)
    ctx.sync(nf)
    sampler = ClockSampler(local_rank)
    ctx.profile(True)
    prof = {}
    barrier()
    sampler.start()
    ctx.launch_count(reset=True)
    t_wall = time.perf_counter()
    ctx.timer_start()
    for _ in range(args.steps):
        ctx.run(nf, stages)
        # per-kernel events of this step; reading them waits for the step (the stream itself stays
        # busy: the next run is enqueued right after)
        for name, ms in ctx.profile_read():
            a = prof.setdefault(name, [0.0, 0])
            a[0] += ms
            a[1] += 1
    ms_total = ctx.timer_stop_ms()
    barrier()
    t_wall = time.perf_counter() - t_wall
    clocks = sampler.stop()
    launches = ctx.launch_count()
    ctx.profile(False)
    ctx.sync(nf)
    stats = batch_stats(ctx, nf)

    # ---- end to end through the C ABI with host buffers ("e2e")
    e2e = run_e2e(lpl, ctx, frames, local_rank, args, barrier, stages, img_h, rings)

    # ---- p50 latency of single-frame batches (H2D -> all results on the host)
    lat = run_latency(lpl, frames, local_rank, stages, max_pts, make_ctx, rings) if rank == 0 else None

    t_ms = torch.tensor([ms_total, e2e["seconds"] * 1e3], dtype=torch.float64, device="cuda")
    if dist is not None:
        dist.all_reduce(t_ms, op=dist.ReduceOp.MAX)
    ms_max, e2e_ms_max = float(t_ms[0]), float(t_ms[1])
    if rank != 0:
        if dist is not None:
            dist.destroy_process_group()
        return 0

    value = world * nf * args.steps / (ms_max * 1e-3)
    e2e_value = world * nf * args.steps / (e2e_ms_max * 1e-3)

    # ---- roofline of the dominant kernel
    import json as _json

    peaks_path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(peaks_path):
        peak = float(_json.load(open(peaks_path))["hbm_gbs"])
        peak_src = "MEASURED_PEAKS.json hbm_gbs (of measured)"
    else:
        peak, peak_src = 6650.0, "B200_PROFILING.md fallback (of fallback)"
    kernels = []
    for name, (ms, cnt) in prof.items():
        per_launch_ms = ms / cnt
        by = algorithmic_bytes(name, stats)
        kernels.append({"kernel": name, "launches": cnt, "ms_per_launch": per_launch_ms,
                        "share": ms / (ms_total if ms_total > 0 else 1.0),
                        "alg_bytes_per_launch": by, "gbs": by / (per_launch_ms * 1e6) if per_launch_ms > 0 else 0.0,
                        "frac_of_hbm_peak": (by / (per_launch_ms * 1e6) / peak) if per_launch_ms > 0 else 0.0})
    # kernels launched twice per step under one name (two-pass compactions) are merged above
    kernels.sort(key=lambda k: -k["share"])
    top = kernels[0]
    roofline = {"bound": "hbm", "kernel": top["kernel"], "achieved": top["gbs"], "peak": peak, "unit": "GB/s",
                "frac": top["gbs"] / peak, "traffic": None, "peak_source": peak_src,
                "share_of_step": top["share"], "ms_per_launch": top["ms_per_launch"],
                "note": "sequential sub-steps (JCP relaxation, hull chains, RECM scans) are latency-bound; see "
                        "DESIGN.md and profiles/ for stall counters"}
    traffic_path = os.path.join(ROOT, "profiles", "traffic.json")
    if os.path.exists(traffic_path):
        try:
            roofline["traffic"] = _json.load(open(traffic_path)).get(top["kernel"])
        except Exception:
            pass

    # ---- CPU baseline on this box's host cores (bounded sample)
    cpu_baseline = None
    if world == 1 and not args.no_cpu and opts["cpu"]:
        cores = host_cores()
        ns = min(nf, 4 * cores)
        cpu = CpuReference(cores, ns, args.workload)
        secs = cpu.run(range(ns))
        cpu.close()
        cpu_baseline = {"value": ns / secs, "unit": UNIT, "cores": cores, "kind": cpu.kind,
                        "sample": f"first {ns} frames of {workload} (unrotated), whole chained pipeline, "
                                  f"{cores} worker processes, {secs:.1f} s"}

    if os.environ.get("LPL_BENCH_KERNELS"):
        with open(os.environ["LPL_BENCH_KERNELS"], "w") as fh:
            _json.dump(kernels, fh, indent=1)
    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": ms_max / args.steps, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": data_desc,
        "config": {"workload": workload, "frames_per_step_per_gpu": nf, "points_per_step_per_gpu": total_pts,
                   "stages": ("ring+" if stages & lpl.STAGE_RING else "") + "dror+segment+cluster+hulls", "l2": f"inputs ({16 * total_pts / 1e6:.0f} MB/step) larger than the 126 MB L2, no flush",
                   "parallelism": f"frame-sharded x{world}, no data-path collective"},
        "points_per_s": world * total_pts * args.steps / (ms_max * 1e-3),
        "clocks": clocks,
        "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": e2e["h2d"], "d2h_bytes_per_step": e2e["d2h"],
                "pipelining": f"{args.e2e_parts} part-batches per step over {args.e2e_ctx} contexts (streams) in rotation, pinned host memory"},
        "gpu_launches": int(launches),
        "roofline": roofline,
        "kernels": [{k: (round(v, 6) if isinstance(v, float) else v) for k, v in kk.items()} for kk in kernels[:20]],
        "latency_ms": lat,
        "wall_s_timed_region": t_wall,
    }
    if cpu_baseline is not None:
        line["cpu_baseline"] = cpu_baseline
    print(json.dumps(line), flush=True)
    if dist is not None:
        dist.destroy_process_group()
    return 0


def run_e2e(lpl, ctx0, frames, device, args, barrier, stages, img_h=64, rings=None):
    """Upload (pinned host -> device) + run + batch download through the package's FramePipeline
    (args.e2e_ctx contexts / CUDA streams rotating over args.e2e_parts part-batches of the step)."""
    from lidar_processing_v2_b200.stream import FramePipeline

    nf = len(frames)
    max_pts = max(f.shape[0] for f in frames)
    n_parts, n_ctx = args.e2e_parts, args.e2e_ctx
    per = (nf + n_parts - 1) // n_parts
    pipe = FramePipeline(device, max_pts, per, stages=stages, n_ctx=n_ctx, image_height=img_h)
    parts = [frames[a:a + per] for a in range(0, nf, per)]
    pinned, views, packed, ring_views = [], [], [], []
    for a in range(0, nf, per) if rings is not None else ():
        rbuf = lpl.PinnedBuffer((sum(r.shape[0] for r in rings[a:a + per]),), np.uint16)
        rv, o = [], 0
        for r in rings[a:a + per]:
            rbuf.array[o:o + r.shape[0]] = r
            rv.append(rbuf.array[o:o + r.shape[0]])
            o += r.shape[0]
        pinned.append(rbuf)
        ring_views.append(rv)
    for part in parts:
        buf = lpl.PinnedBuffer((sum(f.shape[0] for f in part), 4), np.float32)
        v, o = [], 0
        for f in part:
            buf.array[o:o + f.shape[0]] = f
            v.append(buf.array[o:o + f.shape[0]])
            o += f.shape[0]
        pinned.append(buf)
        views.append(v)
        # the frames of a part-batch lie back to back in pinned memory: one H2D transfer per batch
        packed.append((buf.array, np.array([f.shape[0] for f in part], np.uint32)))

    def run_steps(k):
        for _ in range(k):
            for i in range(len(views)):
                if rings is not None:
                    pipe.submit(views[i], rings=ring_views[i])  # per-frame copies: points + ring field
                    continue
                pipe.submit(views[i], packed=packed[i])  # returns (and thereby downloads) the batch this slot held before
        pipe.drain()

    run_steps(max(args.warmup, 1))
    barrier()
    pipe.h2d_bytes = pipe.d2h_bytes = 0
    t0 = time.perf_counter()
    run_steps(args.steps)
    barrier()
    secs = time.perf_counter() - t0
    h2d, d2h = pipe.h2d_bytes // args.steps, pipe.d2h_bytes // args.steps
    pipe.close()
    return {"seconds": secs, "h2d": h2d, "d2h": d2h}


def run_latency(lpl, frames, device, stages, max_pts, make_ctx, rings=None):
    ctx = make_ctx(1)
    stride = ((max_pts + 2047) // 2048) * 2048
    out = lpl.BatchBuffers(1, stride)
    pin = lpl.PinnedBuffer((max_pts, 4), np.float32)
    ts = []
    sel = list(range(min(64, len(frames)))) + list(range(min(8, len(frames))))
    for k in sel:
        f = frames[k]
        v = pin.array[: f.shape[0]]
        v[:] = f
        t0 = time.perf_counter()
        ctx.upload([v], rings=None if rings is None else [rings[k]])
        ctx.run(1, stages)
        ctx.download_batch(1, out)
        ts.append((time.perf_counter() - t0) * 1e3)
    ts = np.array(ts[min(8, len(ts) // 2):])
    ctx.close()
    return {"p50": float(np.percentile(ts, 50)), "p95": float(np.percentile(ts, 95)), "frames": int(ts.size),
            "what": "batch of 1: pinned H2D + all stages + labels/clusters/hulls D2H"}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default=None, choices=WORKLOADS,
                    help="default: kitti154 (BASELINE.json configs[1]); the synthetic shapes are configs[2..4]")
    ap.add_argument("--frames", type=int, default=None, help="frames per batch (default: the whole sequence)")
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    ap.add_argument("--e2e-parts", type=int, default=1, help="part-batches one step is split into on the e2e path")
    ap.add_argument("--e2e-ctx", type=int, default=4, help="contexts (CUDA streams) the e2e path rotates over")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if args.warmup < 3 and args.impl == "ours":
        args.warmup = 3
    if args.impl == "reference":
        return run_reference_arm(args, rank, world)
    return run_ours(args, rank, local_rank, world)


if __name__ == "__main__":
    sys.exit(main())
