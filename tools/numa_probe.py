import os, time, torch
cpu = int(open("/proc/self/stat").read().split()[38])
x = torch.empty(1 << 28, dtype=torch.uint8).pin_memory()   # 256 MB
d = torch.empty(1 << 28, dtype=torch.uint8, device="cuda")
for _ in range(2):
    d.copy_(x, non_blocking=True); torch.cuda.synchronize()
t0 = time.perf_counter(); d.copy_(x, non_blocking=True); torch.cuda.synchronize(); h2d = (1 << 28) / (time.perf_counter() - t0) / 1e9
t0 = time.perf_counter(); x.copy_(d, non_blocking=True); torch.cuda.synchronize(); d2h = (1 << 28) / (time.perf_counter() - t0) / 1e9
print("cpu %d affinity %d cpus: H2D %.1f GB/s D2H %.1f GB/s" % (cpu, len(os.sched_getaffinity(0)), h2d, d2h), flush=True)
