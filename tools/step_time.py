"""Device-resident step time of a workload without per-kernel events (what bench.py's `value` is made of).
usage (on a GPU box): [LPL_WORKLOAD=kitti154|synth64|synth128|cloud2m] [LPL_FRAMES=n] python tools/step_time.py"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
import lidar_processing_v2_b200 as lpl  # noqa: E402


def main():
    for wl in (os.environ.get("LPL_WORKLOAD") or "kitti154,synth128,cloud2m,kitti1").split(","):
        one = wl == "kitti1"
        frames, workload, _, opts = bench.load_frames(None, "kitti154" if one else wl)
        rings = opts["rings"]
        if one:
            frames = frames[:1]
        if os.environ.get("LPL_FRAMES"):
            frames = frames[: int(os.environ["LPL_FRAMES"])]
            rings = rings[: len(frames)] if rings is not None else None
        nf = len(frames)
        stages = lpl.STAGE_ALL & ~lpl.STAGE_RING if opts["stages"] in ("ringless", "ring_field") else lpl.STAGE_ALL
        ctx = bench.make_ctx_factory(lpl, 0, max(f.shape[0] for f in frames), opts["image_height"])(nf)
        ctx.upload(frames, rings=rings)
        if os.environ.get("LPL_PARTS"):
            ctx.use_split(int(os.environ["LPL_PARTS"]))
        for _ in range(5):
            ctx.run(nf, stages)
        ctx.sync(nf)
        steps = 20
        ctx.timer_start()
        for _ in range(steps):
            ctx.run(nf, stages)
        ms = ctx.timer_stop_ms() / steps
        print(f"{wl} x{nf}: {ms:.3f} ms/step  {nf / ms * 1e3:.0f} frames/s", flush=True)
        del ctx


if __name__ == "__main__":
    main()
