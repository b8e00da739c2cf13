"""Host side of the frame-parallel deployment: frame sharding across ranks, the double-buffered
per-GPU stream pipeline, and the end-of-run statistics reduction.

Frames are independent (the reference resets every buffer per call, segmenter.cpp:73-85,
clusterer.cpp:104-106, and re-seeds its RNG per call, segmenter.cpp:369), so the data path needs no
collective: each rank owns a contiguous block of frames and one GPU; torch.distributed (NCCL on
GPUs, gloo in the CPU tests) only sums counters and takes the slowest rank's time at the end.
"""
from __future__ import annotations

import time

import numpy as np

from . import native as _n


def shard_range(n_frames: int, rank: int, world: int) -> tuple[int, int]:
    """Contiguous block [start, stop) of frames for `rank`; the first n_frames % world ranks get one more."""
    if world <= 0 or not (0 <= rank < world):
        raise ValueError("rank / world out of range")
    base, extra = divmod(n_frames, world)
    start = rank * base + min(rank, extra)
    return start, start + base + (1 if rank < extra else 0)


def reduce_stats(counters: dict, elapsed_s: float, dist=None, device=None) -> dict:
    """Sum integer counters and take the maximum elapsed time over all ranks (a few hundred bytes).
    `dist` is an initialised torch.distributed module or None for a single process."""
    keys = sorted(counters)
    if dist is None or not dist.is_initialized() or dist.get_world_size() == 1:
        return dict(counters, elapsed_s=float(elapsed_s), world=1)
    import torch

    t = torch.tensor([float(counters[k]) for k in keys], dtype=torch.float64, device=device)
    e = torch.tensor([float(elapsed_s)], dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.SUM)
    dist.all_reduce(e, op=dist.ReduceOp.MAX)
    out = {k: (int(round(v)) if isinstance(counters[k], (int, np.integer)) else float(v))
           for k, v in zip(keys, t.tolist())}
    out["elapsed_s"] = float(e.item())
    out["world"] = dist.get_world_size()
    return out


class FramePipeline:
    """`n_ctx` contexts (one CUDA stream each) working on batches in rotation: while one batch's
    results cross PCIe and the next batch is uploaded from pinned host memory, the other batches'
    kernels keep the SMs busy. `submit` returns the results of the batch submitted `n_ctx` calls
    earlier."""

    def __init__(self, device: int, max_points: int, batch: int, stages: int = _n.STAGE_ALL,
                 cluster_cfg: dict | None = None, want=("labels_u8", "cluster_labels", "hull_offsets", "hull_xy",
                                                         "zminmax"), n_ctx: int = 2,
                 image_height: int = 64, packed_results: bool = True, result_bytes_per_point: int = 8,
                 graph: bool | None = None, split: int | None = None):
        """`want`: result planes that come back (obstacle_index is derivable on the host: the ascending positions
        of label 2). packed_results: one D2H transfer per batch of exactly the occupied bytes
        (lpl_pipeline_download_packed) instead of one strided copy per plane. graph / split: override the per-context
        CUDA-graph replay and sub-batch count (defaults: graph replay on, no sub-batches in a rotation)."""
        if n_ctx < 1:
            raise ValueError("n_ctx must be >= 1")
        self.stages = stages
        self.batch = batch
        self.n_ctx = n_ctx
        self.ctx = [_n.Context(device, max_points=max_points, max_frames=batch, image_height=image_height)
                    for _ in range(n_ctx)]
        for c in self.ctx:
            # whole-chain graph per batch: +2 % for one stream working alone; in a rotation of four it measured -5 %
            # early in round 2 and +1 % with the final kernels (tools/e2e_sweep.py: 36.25k vs 35.89k frames/s) - on
            c.use_graph(True if graph is None else bool(graph))
            if split is not None:
                c.use_split(int(split))
            elif n_ctx > 1:
                c.use_split(1)  # the rotation already overlaps whole batches (measured: 36.9k vs 36.7k frames/s with 2 sub-batches)
            if image_height != 64:
                cfg = c.segmenter_default_cfg()
                cfg.image_height = image_height
                c.segmenter_config(cfg)
            c.cluster_config(**(cluster_cfg or dict(range_m=0.4, az_deg=1.0, el_deg=3.0, min_size=3)))
        stride = ((max_points + 2047) // 2048) * 2048
        self.packed_results = packed_results
        if packed_results:
            self.out = [_n.PackedBuffers(batch, batch * stride * result_bytes_per_point, want=want) for _ in range(n_ctx)]
        else:
            self.out = [_n.BatchBuffers(batch, stride, want=want) for _ in range(n_ctx)]
        self.inflight = [0] * n_ctx
        self.turn = 0
        self.h2d_bytes = 0
        self.d2h_bytes = 0

    def submit(self, frames, packed=None, rings=None):
        """Enqueue a batch (list of (n, 4) float32 arrays, ideally views of pinned memory). When the
        frames lie back to back in one buffer, pass it as packed=(array, counts): the batch then crosses
        PCIe as one transfer; an (N, 3) array is the 12-byte std::array<float, 3> layout, an (N, 4) array the
        16-byte PCL layout. Returns (counts[5][nf], result buffers) of the batch that previously used this
        slot, or None."""
        i = self.turn
        done = self.collect(i)
        if packed is not None:
            if packed[0].shape[1] == 3:
                nf = self.ctx[i].upload_packed_xyz(packed[0], packed[1])
            else:
                nf = self.ctx[i].upload_packed(packed[0], packed[1])
            self.h2d_bytes += int(np.sum(packed[1])) * 4 * int(packed[0].shape[1])
        else:
            nf = self.ctx[i].upload(frames, rings=rings)
            self.h2d_bytes += sum(int(f.shape[0]) for f in frames) * 16
            if rings is not None:
                self.h2d_bytes += sum(int(r.shape[0]) for r in rings if r is not None) * 2
        self.ctx[i].run(nf, self.stages)
        self.inflight[i] = nf
        self.turn = (self.turn + 1) % self.n_ctx
        return done

    def collect(self, i: int):
        nf = self.inflight[i]
        if nf == 0:
            return None
        if self.packed_results:
            counts = self.ctx[i].download_packed(nf, self.out[i])
            self.d2h_bytes += self.out[i].bytes_used + counts.size * 4
        else:
            counts = self.ctx[i].download_batch(nf, self.out[i])
            self.d2h_bytes += self.out[i].bytes_for(counts)
        self.inflight[i] = 0
        return counts, self.out[i]

    def drain(self):
        """Results of everything still in flight, oldest first."""
        res = []
        for k in range(self.n_ctx):
            r = self.collect((self.turn + k) % self.n_ctx)
            if r is not None:
                res.append(r)
        return res

    def close(self):
        for c in self.ctx:
            c.close()
        for o in self.out:
            o.close()


def run_stream(frames, device: int, batch: int, rank: int = 0, world: int = 1, dist=None) -> dict:
    """Process this rank's shard of `frames` through a FramePipeline; returns the reduced statistics."""
    a, b = shard_range(len(frames), rank, world)
    mine = frames[a:b]
    stats = dict(frames=0, points=0, obstacles=0, clusters=0, hull_vertices=0)
    if not mine:
        return reduce_stats(stats, 0.0, dist)
    pipe = FramePipeline(device, max(f.shape[0] for f in mine), batch)

    def account(res):
        if res is None:
            return
        counts, _ = res
        stats["frames"] += int(counts.shape[1])
        stats["points"] += int(counts[0].sum())
        stats["obstacles"] += int(counts[2].sum())
        stats["clusters"] += int(counts[3].sum())
        stats["hull_vertices"] += int(counts[4].sum())

    t0 = time.perf_counter()
    for s in range(0, len(mine), batch):
        account(pipe.submit(mine[s:s + batch]))
    for r in pipe.drain():
        account(r)
    elapsed = time.perf_counter() - t0
    pipe.close()
    return reduce_stats(stats, elapsed, dist)
