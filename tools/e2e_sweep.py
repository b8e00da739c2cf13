"""Experiment: end-to-end frames/s of bench.run_e2e for several (part-batches, contexts) choices.
usage (on a GPU box): python tools/e2e_sweep.py [parts:ctx ...]"""
import argparse
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
import lidar_processing_v2_b200 as lpl  # noqa: E402


def main():
    combos = [tuple(int(v) for v in a.split(":")) for a in sys.argv[1:]] or [(2, 2), (3, 3), (4, 2), (4, 3), (4, 4), (6, 3)]
    frames, workload, _, _ = bench.load_frames(None)
    for parts, nctx in combos:
        args = argparse.Namespace(steps=6, warmup=2, e2e_parts=parts, e2e_ctx=nctx)
        r = bench.run_e2e(lpl, None, frames, 0, args, lambda: None, lpl.STAGE_ALL)
        print(f"{workload} parts={parts} ctx={nctx}: {len(frames) * args.steps / r['seconds']:.0f} frames/s "
              f"({r['seconds'] / args.steps * 1e3:.2f} ms/step)", flush=True)


if __name__ == "__main__":
    main()
