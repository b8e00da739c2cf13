"""Per-frame hull-stage statistics on the KITTI batch: obstacle points, points entering the hull sort
(after the octagon filter), occupied voxels, clusters, and the cluster-size tail.
usage (on a GPU box): python tools/hull_stats.py"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
import lidar_processing_v2_b200 as lpl  # noqa: E402


def main():
    frames, workload, _, _ = bench.load_frames(None)
    nf = len(frames)
    ctx = lpl.Context(0, max_points=max(f.shape[0] for f in frames), max_frames=nf)
    ctx.cluster_config(range_m=0.4, az_deg=1.0, el_deg=3.0, min_size=3)
    ctx.upload(frames)
    ctx.run(nf, lpl.STAGE_ALL)
    ctx.sync(nf)
    rows = []
    big = []
    for f in range(nf):
        r = ctx.counts(f)
        h = ctx.debug_hulls(f)
        rows.append((r.num_obstacles, h["n_hull_sort"], h["n_voxels"], r.num_clusters, r.num_hull_vertices))
        if f % 16 == 0:
            out = ctx.download(f)
            sizes = np.bincount(out["cluster_labels"][out["cluster_labels"] >= 0])
            big.append(np.sort(sizes)[-5:][::-1])
    a = np.array(rows, float)
    print(workload, "means: obstacles %.0f, into hull sort %.0f (%.1f %%), voxels %.0f, clusters %.1f, hull vertices %.0f"
          % (a[:, 0].mean(), a[:, 1].mean(), 100 * a[:, 1].sum() / a[:, 0].sum(), a[:, 2].mean(), a[:, 3].mean(), a[:, 4].mean()))
    print("max into hull sort", int(a[:, 1].max()), "max voxels", int(a[:, 2].max()))
    print("five largest clusters of every 16th frame:", [list(map(int, b)) for b in big])


if __name__ == "__main__":
    main()
