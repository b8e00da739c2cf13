"""Host ceiling probe: aggregate pinned-memory H2D / D2H bandwidth when N ranks copy at the same time.

    python -m torch.distributed.run --nproc-per-node N --master-addr 127.0.0.1 --master-port 29511 tools/h2d_probe.py

Every rank copies a 300 MB pinned buffer to its GPU (and 60 MB back, concurrently on a second stream) in a loop
for ~1 s after a barrier; rank 0 prints the per-rank and the summed GB/s. If the sum stops growing with N the
limit is the host (memory / PCIe root), not the GPUs: that is the ceiling of the e2e metric at N GPUs.
"""
import json
import os
import time

import torch
import torch.distributed as dist


def main():
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    up = torch.empty(300_000_000, dtype=torch.uint8).pin_memory()
    up.random_(0, 255)
    dn = torch.empty(60_000_000, dtype=torch.uint8).pin_memory()
    d_up = torch.empty_like(up, device="cuda")
    d_dn = torch.empty_like(dn, device="cuda")
    s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
    res = {}
    for mode in ("h2d", "h2d+d2h"):
        for _ in range(2):
            d_up.copy_(up, non_blocking=True)
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        n = 0
        while time.perf_counter() - t0 < 1.0:
            with torch.cuda.stream(s1):
                d_up.copy_(up, non_blocking=True)
            if mode != "h2d":
                with torch.cuda.stream(s2):
                    dn.copy_(d_dn, non_blocking=True)
            s1.synchronize()
            s2.synchronize()
            n += 1
        dt = time.perf_counter() - t0
        res[mode] = [n * 0.3 / dt, (n * 0.06 / dt) if mode != "h2d" else 0.0]
    t = torch.tensor([res["h2d"][0], res["h2d+d2h"][0], res["h2d+d2h"][1]], dtype=torch.float64, device="cuda")
    allv = [torch.zeros_like(t) for _ in range(world)]
    if world > 1:
        dist.all_gather(allv, t)
    else:
        allv = [t]
    if rank == 0:
        m = torch.stack(allv).cpu()
        print(json.dumps({"n_gpus": world, "cpus": len(os.sched_getaffinity(0)),
                          "h2d_alone_gbs_per_rank": [round(v, 1) for v in m[:, 0].tolist()], "h2d_alone_gbs_sum": round(float(m[:, 0].sum()), 1),
                          "h2d_with_d2h_gbs_sum": round(float(m[:, 1].sum()), 1), "d2h_gbs_sum": round(float(m[:, 2].sum()), 1)}), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
