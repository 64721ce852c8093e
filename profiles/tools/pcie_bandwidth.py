import torch, time
x = torch.empty(256 << 20, dtype=torch.uint8).pin_memory()
d = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
for sz in (1 << 20, 8 << 20, 64 << 20, 256 << 20):
    for name, a, b in (("h2d", d, x), ("d2h", x, d)):
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        reps = max(1, (256 << 20) // sz)
        for r in range(reps):
            a[:sz].copy_(b[:sz], non_blocking=True)
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
        print(name, sz >> 20, "MiB", "%.1f GB/s" % (reps * sz / dt / 1e9))
# both directions at once
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
y = torch.empty(256 << 20, dtype=torch.uint8).pin_memory()
d2 = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
torch.cuda.synchronize(); t0 = time.perf_counter()
with torch.cuda.stream(s1): d.copy_(x, non_blocking=True)
with torch.cuda.stream(s2): y.copy_(d2, non_blocking=True)
torch.cuda.synchronize(); dt = time.perf_counter() - t0
print("duplex %.1f GB/s each way" % ((256 << 20) / dt / 1e9))
