"""Experiment: end-to-end frames/s of the FramePipeline rotation for several (contexts, graph, sub-batches) choices.
usage (on a GPU box): python tools/e2e_sweep.py [ctx:graph:split ...]"""
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
import lidar_processing_v2_b200 as lpl  # noqa: E402
from lidar_processing_v2_b200.stream import FramePipeline  # noqa: E402


def run(frames, n_ctx, graph, split, seconds=1.5):
    nf = len(frames)
    counts = np.array([f.shape[0] for f in frames], np.uint32)
    buf = lpl.PinnedBuffer((int(counts.sum()), 3), np.float32)
    o = 0
    for f in frames:
        buf.array[o:o + f.shape[0]] = f[:, :3]
        o += f.shape[0]
    pipe = FramePipeline(0, int(counts.max()), nf, n_ctx=n_ctx, graph=graph, split=split)
    for _ in range(2 * n_ctx):
        pipe.submit(None, packed=(buf.array, counts))
    pipe.drain()
    t0 = time.perf_counter()
    k = 0
    while time.perf_counter() - t0 < seconds:
        for _ in range(n_ctx):
            pipe.submit(None, packed=(buf.array, counts))
        k += n_ctx
    pipe.drain()
    dt = time.perf_counter() - t0
    pipe.close()
    buf.close()
    return nf * k / dt


def main():
    combos = [tuple(int(v) for v in a.split(":")) for a in sys.argv[1:]] or [(4, 0, 1), (4, 0, 2), (4, 1, 1), (4, 1, 2), (3, 0, 1), (3, 0, 2),
                                                                           (2, 0, 2), (2, 1, 2), (6, 0, 1)]
    frames, workload, _, _ = bench.load_frames(None)
    for n_ctx, graph, split in combos:
        print(f"{workload} ctx={n_ctx} graph={graph} split={split}: {run(frames, n_ctx, graph, split):.0f} frames/s", flush=True)


if __name__ == "__main__":
    main()
