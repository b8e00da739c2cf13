"""How much of a step do concurrent sub-batches hide? The 154-frame batch is split over k contexts
(one CUDA stream each); every step enqueues all of them, the wall clock over `steps` steps is taken
after a sync of every stream. k = 1 is the single-stream step bench.py times as `value`.

    python tools/overlap_probe.py [--steps 20] [--splits 1,2,3,4]
"""
from __future__ import annotations

import argparse
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

import lidar_processing_v2_b200 as lpl  # noqa: E402
from bench import load_frames  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--splits", default="1,2,3,4")
    ap.add_argument("--workload", default=None)
    ap.add_argument("--prio", default="", help="comma list of stream priorities per context (0 .. -5), e.g. -5,0,0")
    args = ap.parse_args()
    frames, workload, _, opts = load_frames(None, args.workload)
    stages = lpl.STAGE_ALL if opts["stages"] is None else (lpl.STAGE_ALL & ~lpl.STAGE_RING)
    max_pts = max(f.shape[0] for f in frames)
    out = {}
    for k in [int(v) for v in args.splits.split(",")]:
        per = (len(frames) + k - 1) // k
        parts = [frames[a:a + per] for a in range(0, len(frames), per)]
        ctxs = []
        prios = [int(v) for v in args.prio.split(",")] if args.prio else []
        for j, p in enumerate(parts):
            os.environ["LPL_STREAM_PRIORITY"] = str(prios[j] if j < len(prios) else 0)
            c = lpl.Context(0, max_points=max_pts, max_frames=len(p), image_height=opts["image_height"])
            c.cluster_config(range_m=0.4, az_deg=1.0, el_deg=3.0, min_size=3)
            c.upload(p, rings=None)
            c.sync(len(p))
            ctxs.append(c)
        for _ in range(3):
            for c, p in zip(ctxs, parts):
                c.run(len(p), stages)
        for c, p in zip(ctxs, parts):
            c.sync(len(p))
        t0 = time.perf_counter()
        for _ in range(args.steps):
            for c, p in zip(ctxs, parts):
                c.run(len(p), stages)
        for c, p in zip(ctxs, parts):
            c.sync(len(p))
        dt = time.perf_counter() - t0
        out[k] = {"ms_per_step": 1e3 * dt / args.steps, "frames_per_s": len(frames) * args.steps / dt}
        for c in ctxs:
            c.close()
    print(json.dumps({"workload": workload, "frames": len(frames), "splits": out}))


if __name__ == "__main__":
    main()
