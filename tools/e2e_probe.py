"""Experiment: where does the end-to-end step time go? Variants of the FramePipeline loop with the
result download reduced / removed and with the host-side call times accumulated.
usage (on a GPU box): python tools/e2e_probe.py"""
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
import lidar_processing_v2_b200 as lpl  # noqa: E402
from lidar_processing_v2_b200.stream import FramePipeline  # noqa: E402


def run(frames, n_ctx, want, steps=8, warm=2):
    nf = len(frames)
    max_pts = max(f.shape[0] for f in frames)
    pipe = FramePipeline(0, max_pts, nf, n_ctx=n_ctx, want=want)
    buf = lpl.PinnedBuffer((sum(f.shape[0] for f in frames), 4), np.float32)
    views, o = [], 0
    for f in frames:
        buf.array[o:o + f.shape[0]] = f
        views.append(buf.array[o:o + f.shape[0]])
        o += f.shape[0]
    counts_arr = np.array([f.shape[0] for f in frames], np.uint32)
    t_collect = t_upload = t_run = 0.0

    def step():
        nonlocal t_collect, t_upload, t_run
        i = pipe.turn
        a = time.perf_counter()
        pipe.collect(i)
        b = time.perf_counter()
        n = pipe.ctx[i].upload_packed(buf.array, counts_arr)
        c = time.perf_counter()
        pipe.ctx[i].run(n, pipe.stages)
        d = time.perf_counter()
        pipe.inflight[i] = n
        pipe.turn = (pipe.turn + 1) % pipe.n_ctx
        t_collect += b - a
        t_upload += c - b
        t_run += d - c

    for _ in range(warm):
        step()
    pipe.drain()
    t_collect = t_upload = t_run = 0.0
    t0 = time.perf_counter()
    for _ in range(steps):
        step()
    pipe.drain()
    dt = time.perf_counter() - t0
    pipe.close()
    buf.close()
    return dt / steps * 1e3, t_collect / steps * 1e3, t_upload / steps * 1e3, t_run / steps * 1e3


def main():
    frames, workload, _, _ = bench.load_frames(None)
    full = ("labels_u8", "obstacle_index", "cluster_labels", "hull_offsets", "hull_xy", "zminmax")
    for n_ctx, want, tag in ((4, full, "all results"), (4, ("hull_offsets",), "counts + hull offsets only"),
                             (2, full, "all results"), (1, full, "all results, one context")):
        ms, tc, tu, tr = run(frames, n_ctx, want)
        print(f"{workload} ctx={n_ctx} {tag}: {ms:.2f} ms/step  host: collect {tc:.2f} upload {tu:.2f} run {tr:.2f} ms",
              flush=True)


if __name__ == "__main__":
    main()
