import torch, time
d = torch.empty(64<<20, dtype=torch.uint8, device='cuda')
hp = torch.empty(64<<20, dtype=torch.uint8).pin_memory()
hq = torch.empty(64<<20, dtype=torch.uint8)
for name, h in (("pinned", hp), ("pageable", hq)):
    for sz in (1<<20, 11<<20, 64<<20):
        for direction in ("d2h", "h2d"):
            for _ in range(2):
                torch.cuda.synchronize(); t0 = time.perf_counter()
                if direction == "d2h": h[:sz].copy_(d[:sz], non_blocking=True)
                else: d[:sz].copy_(h[:sz], non_blocking=True)
                torch.cuda.synchronize(); dt = time.perf_counter() - t0
            print(name, direction, sz >> 20, "MB", "%.2f ms" % (dt * 1e3), "%.1f GB/s" % (sz / dt / 1e9))
