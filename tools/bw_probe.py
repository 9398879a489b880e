import torch, time
a = torch.empty(167_000_000, dtype=torch.uint8).pin_memory(); d = torch.empty_like(a, device='cuda')
for n in (167_000_000, 21_000_000):
    for _ in range(3): d[:n].copy_(a[:n], non_blocking=True)
    torch.cuda.synchronize(); t0=time.perf_counter()
    for _ in range(10): d[:n].copy_(a[:n], non_blocking=True)
    torch.cuda.synchronize(); dt=(time.perf_counter()-t0)/10
    print('H2D', n/1e6, 'MB', dt*1e3, 'ms', n/dt/1e9, 'GB/s')
    for _ in range(3): a[:n].copy_(d[:n], non_blocking=True)
    torch.cuda.synchronize(); t0=time.perf_counter()
    for _ in range(10): a[:n].copy_(d[:n], non_blocking=True)
    torch.cuda.synchronize(); dt=(time.perf_counter()-t0)/10
    print('D2H', n/1e6, 'MB', dt*1e3, 'ms', n/dt/1e9, 'GB/s')
