"""Verbose stage-by-stage parity + timing report on a GPU box (debugging aid, not a test).

usage: python tools/gpu_diag.py [--frames N] [--batch B] [--quick]
Writes gpurun_out/diag.json.
"""
from __future__ import annotations

import argparse
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import lidar_processing_v2_b200 as lpl  # noqa: E402
from oracle.oracle import JCP_AS_IS, JCP_CLEAN, PortOracle, RefOracle, have_ref  # noqa: E402
from tools import frames as F  # noqa: E402

import parity  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--frames", type=int, default=4)
    ap.add_argument("--batch", type=int, default=8)
    ap.add_argument("--quick", action="store_true")
    args = ap.parse_args()
    out = {"reports": []}
    port = PortOracle()
    ref = RefOracle() if have_ref() else None
    print("ref oracle:", "yes" if ref else "no", flush=True)
    ctx = lpl.Context(0, max_points=131072, max_frames=max(args.batch, 1))
    g0 = F.load_golden("kitti_f000")
    pts = g0["pts"]
    t = time.time()
    rep = parity.stage_report(ctx, ref or port, pts)
    rep["name"] = "kitti_f000 vs " + ("ref" if ref else "port")
    rep["secs"] = round(time.time() - t, 2)
    print(json.dumps(rep), flush=True)
    out["reports"].append(rep)
    if not args.quick:
        rep = parity.stage_report(ctx, port, pts, jcp_mode=JCP_CLEAN)
        rep["name"] = "kitti_f000 vs port (clean JCP)"
        print(json.dumps(rep), flush=True)
        out["reports"].append(rep)
        sp, sr = F.synth_scan(4001)
        rep = parity.stage_report(ctx, ref or port, sp, ring_given=sr)
        rep["name"] = "synth_4001"
        print(json.dumps(rep), flush=True)
        out["reports"].append(rep)
        if F.have_pack():
            fr = F.load_pack(limit=args.frames)
            for i, p in enumerate(fr[1:], 1):
                rep = parity.stage_report(ctx, ref or port, p, check_ringless=False)
                rep["name"] = f"pack[{i}]"
                print(json.dumps(rep), flush=True)
                out["reports"].append(rep)
    # ---- chained batch
    frames = [pts] + [F.synth_scan(4100 + i)[0] for i in range(args.batch - 1)]
    ctx.cluster_config(range_m=0.4, az_deg=1.0, el_deg=3.0, min_size=3)
    ctx.set_jcp_mode(lpl.JCP_AS_REFERENCE)
    for dror in (False, True):
        stages = lpl.STAGE_ALL if dror else (lpl.STAGE_ALL & ~lpl.STAGE_DROR)
        nf = ctx.upload(frames)
        ctx.run(nf, stages)
        ctx.sync(nf)
        for f in range(nf):
            got = ctx.download(f)
            exp = parity.oracle_chain(port, frames[f], dror)
            rep = parity.chain_report(got, exp)
            rep["name"] = f"chain dror={dror} frame={f}"
            rep["counts"] = [got["n"], got["num_valid"], got["num_obstacles"], got["num_clusters"],
                             got["num_hull_vertices"]]
            print(json.dumps(rep), flush=True)
            out["reports"].append(rep)
    # ---- timing of the batch (device-resident after upload)
    for it in range(3):
        nf = ctx.upload(frames)
        ctx.sync(nf)
        ctx.launch_count(reset=True)
        ctx.timer_start()
        ctx.run(nf, lpl.STAGE_ALL)
        ms = ctx.timer_stop_ms()
        print(f"batch of {nf}: {ms:.3f} ms -> {nf / ms * 1e3:.0f} frames/s, launches {ctx.launch_count()}", flush=True)
        out["batch_ms"] = ms
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    with open(os.path.join(ROOT, "gpurun_out", "diag.json"), "w") as f:
        json.dump(out, f, indent=1)
    bad = [r["name"] for r in out["reports"] if not all(v == 0 for k, v in r.items() if k.endswith("_diff"))]
    print("MISMATCHING:", bad)
    return 0


if __name__ == "__main__":
    sys.exit(main())
