"""Workload for compute-sanitizer: smoke() (stage-wise calls + a ragged chained batch) and a 16-frame batch run as one
stream and as two concurrent sub-batches (plain, then as a captured and replayed graph), results compared.
usage (on a GPU box): compute-sanitizer --tool memcheck|synccheck|initcheck|racecheck python tools/sanitize_run.py"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import __graft_entry__ as G  # noqa: E402
import lidar_processing_v2_b200 as lpl  # noqa: E402
from tools import frames as F  # noqa: E402


def main():
    G.smoke()
    g0, g1 = F.load_golden("kitti_f000")["pts"], F.load_golden("kitti_f100")["pts"]
    frames = [(g0 if i % 2 == 0 else g1)[: 60000 + 4000 * i].copy() for i in range(16)]
    ctx = lpl.Context(0, max_points=131072, max_frames=16)
    ctx.cluster_config(range_m=0.4, az_deg=1.0, el_deg=3.0, min_size=3)
    nf = ctx.upload(frames)
    outs = {}
    for parts in (1, 2):
        ctx.use_split(parts)
        for _ in range(3):
            ctx.run(nf, lpl.STAGE_ALL)
        ctx.sync(nf)
        outs[parts] = [ctx.download(f) for f in range(nf)]
    for f in range(nf):
        for k in ("labels", "noise", "cluster_labels", "hull_offsets", "hull_xy", "zminmax"):
            assert np.array_equal(outs[1][f][k], outs[2][f][k]), (f, k)
    print("sanitize_run ok: 16 frames, sub-batches 1 and 2 agree")


if __name__ == "__main__":
    main()
