import torch, time, os
print("cpus", os.cpu_count(), "affinity", len(os.sched_getaffinity(0)))
x = torch.empty(300_000_000, dtype=torch.uint8).pin_memory()
d = torch.empty_like(x, device="cuda")
for rep in range(3):
    torch.cuda.synchronize(); t=time.perf_counter(); d.copy_(x, non_blocking=True); torch.cuda.synchronize(); dt=time.perf_counter()-t
    print("H2D 300MB one copy: %.2f ms %.1f GB/s" % (dt*1e3, 0.3/dt))
h = torch.empty(100_000_000, dtype=torch.uint8).pin_memory()
for rep in range(2):
    torch.cuda.synchronize(); t=time.perf_counter(); h.copy_(d[:100_000_000], non_blocking=True); torch.cuda.synchronize(); dt=time.perf_counter()-t
    print("D2H 100MB: %.2f ms %.1f GB/s" % (dt*1e3, 0.1/dt))
# chunks of 2 MB
s = torch.cuda.Stream()
with torch.cuda.stream(s):
    torch.cuda.synchronize(); t=time.perf_counter()
    for i in range(150):
        d[i*2_000_000:(i+1)*2_000_000].copy_(x[i*2_000_000:(i+1)*2_000_000], non_blocking=True)
    torch.cuda.synchronize(); dt=time.perf_counter()-t
print("H2D 150 x 2MB: %.2f ms %.1f GB/s" % (dt*1e3, 0.3/dt))
