"""H2D rate of one 760 MB pinned tensor copied as 1, 2 or 4 concurrent pieces (one CUDA stream each)."""
import time, torch
n = 65536 * 100 * 29
src = torch.empty(n, dtype=torch.float32).pin_memory()
src.uniform_()
dst = torch.empty(n, dtype=torch.float32, device="cuda:0")
for parts in (1, 2, 4, 1, 2, 4):
    streams = [torch.cuda.Stream() for _ in range(parts)]
    step = (n + parts - 1) // parts
    def go():
        for i, s in enumerate(streams):
            with torch.cuda.stream(s):
                dst[i * step:(i + 1) * step].copy_(src[i * step:(i + 1) * step], non_blocking=True)
    go(); torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(5):
        go()
    torch.cuda.synchronize()
    dt = (time.perf_counter() - t0) / 5
    print(f"{parts} stream(s): {dt * 1e3:.2f} ms per 760 MB = {n * 4 / dt / 1e9:.1f} GB/s")
