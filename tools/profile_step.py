"""Minimal driver for ncu: upload one batch and run the whole pipeline a few times.

usage: python tools/profile_step.py [--frames N] [--steps K] [--stages MASK]
(no timing, no oracle; numbers printed under a profiler are never bench values)
"""
from __future__ import annotations

import argparse
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import lidar_processing_v2_b200 as lpl  # noqa: E402
from tools import frames as F  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--frames", type=int, default=64)
    ap.add_argument("--steps", type=int, default=2)
    ap.add_argument("--stages", type=int, default=lpl.STAGE_ALL)
    ap.add_argument("--workload", default=None, help="synth128 | cloud2m | synth64 (bench.py shapes); default: the KITTI pack")
    a = ap.parse_args()
    if a.workload:
        import bench

        fr, _, _, opts = bench.load_frames(None, a.workload)
        stages = lpl.STAGE_ALL & ~lpl.STAGE_RING if opts["stages"] in ("ringless", "ring_field") else lpl.STAGE_ALL
        ctx = bench.make_ctx_factory(lpl, 0, max(f.shape[0] for f in fr), opts["image_height"])(len(fr))
        nf = ctx.upload(fr, rings=opts["rings"])
    else:
        fr = F.load_pack(limit=a.frames) if F.have_pack() else [F.synth_scan(4000 + i)[0] for i in range(a.frames)]
        stages = a.stages
        ctx = lpl.Context(0, max_points=max(f.shape[0] for f in fr), max_frames=len(fr))
        ctx.cluster_config(range_m=0.4, az_deg=1.0, el_deg=3.0, min_size=3)
        nf = ctx.upload(fr)
    for _ in range(a.steps):
        ctx.run(nf, stages)
    ctx.sync(nf)
    print("launches", ctx.launch_count())


if __name__ == "__main__":
    main()
