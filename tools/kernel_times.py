"""Per-kernel CUDA-event times of one pipeline step on the KITTI batch (or synthetic frames).
usage (on a GPU box): python tools/kernel_times.py [kernel-name-substring ...]"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
import lidar_processing_v2_b200 as lpl  # noqa: E402


def main():
    want = sys.argv[1:]
    frames, workload, _, opts = bench.load_frames(None, os.environ.get("LPL_WORKLOAD"))
    rings = opts["rings"]
    if os.environ.get("LPL_FRAMES"):
        frames = frames[: int(os.environ["LPL_FRAMES"])]
        rings = rings[: len(frames)] if rings is not None else None
    nf = len(frames)
    cap = int(max(f.shape[0] for f in frames) * float(os.environ.get("LPL_CAP_SCALE", "1")))  # >1: how much do capacity-sized grids cost?
    stages = lpl.STAGE_ALL & ~lpl.STAGE_RING if opts["stages"] in ("ringless", "ring_field") else lpl.STAGE_ALL
    ctx = bench.make_ctx_factory(lpl, 0, cap, opts["image_height"])(nf)
    ctx.upload(frames, rings=rings)
    for _ in range(3):
        ctx.run(nf, stages)
    ctx.sync(nf)
    ctx.profile(True)
    acc = {}
    steps = 5
    for _ in range(steps):
        ctx.run(nf, stages)
        ctx.sync(nf)
        for name, ms in ctx.profile_read():
            acc[name] = acc.get(name, 0.0) + ms
    tot = sum(acc.values()) / steps
    print(f"{workload} x{nf}: {tot:.3f} ms/step", flush=True)
    for name, ms in sorted(acc.items(), key=lambda kv: -kv[1]):
        if not want or any(w in name for w in want):
            print(f"  {name:20s} {ms / steps:.3f} ms")


if __name__ == "__main__":
    main()
