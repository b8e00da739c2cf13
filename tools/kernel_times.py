"""Per-kernel CUDA-event times of one pipeline step on the KITTI batch (or synthetic frames).
usage (on a GPU box): python tools/kernel_times.py [kernel-name-substring ...]"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
import lidar_processing_v2_b200 as lpl  # noqa: E402


def main():
    want = sys.argv[1:]
    frames, workload, _, _ = bench.load_frames(None, os.environ.get("LPL_WORKLOAD"))
    if os.environ.get("LPL_FRAMES"):
        frames = frames[: int(os.environ["LPL_FRAMES"])]
    nf = len(frames)
    cap = int(max(f.shape[0] for f in frames) * float(os.environ.get("LPL_CAP_SCALE", "1")))  # >1: how much do capacity-sized grids cost?
    ctx = lpl.Context(0, max_points=cap, max_frames=nf)
    ctx.cluster_config(range_m=0.4, az_deg=1.0, el_deg=3.0, min_size=3)
    ctx.upload(frames)
    for _ in range(3):
        ctx.run(nf, lpl.STAGE_ALL)
    ctx.sync(nf)
    ctx.profile(True)
    acc = {}
    steps = 5
    for _ in range(steps):
        ctx.run(nf, lpl.STAGE_ALL)
        ctx.sync(nf)
        for name, ms in ctx.profile_read():
            acc[name] = acc.get(name, 0.0) + ms
    tot = sum(acc.values()) / steps
    print(f"{workload}: {tot:.3f} ms/step", flush=True)
    for name, ms in sorted(acc.items(), key=lambda kv: -kv[1]):
        if not want or any(w in name for w in want):
            print(f"  {name:20s} {ms / steps:.3f} ms")


if __name__ == "__main__":
    main()
