import torch, time
dev = torch.device("cuda:0")
for mb in (33, 264, 1024):
    h = torch.empty(mb * 1024 * 1024, dtype=torch.uint8).pin_memory()
    d = torch.empty_like(h, device=dev)
    for _ in range(3): d.copy_(h, non_blocking=True)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    n = max(4, 4096 // mb)
    e0.record()
    for _ in range(n): d.copy_(h, non_blocking=True)
    e1.record(); torch.cuda.synchronize()
    print(f"{mb} MB x {n}: {mb * 1.048576e6 * n / (e0.elapsed_time(e1) * 1e-3) / 1e9:.2f} GB/s")
