import torch, time
a = torch.empty(167_000_000, dtype=torch.uint8).pin_memory(); d = torch.empty_like(a, device='cuda')
x = torch.randn(8192, 8192, device='cuda', dtype=torch.float32)
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
def copy():
    with torch.cuda.stream(s1): d.copy_(a, non_blocking=True)
def comp():
    with torch.cuda.stream(s2):
        for _ in range(2): y = x @ x
for f, name in ((copy, 'copy'), (comp, 'compute'), (lambda: (copy(), comp()), 'both')):
    for _ in range(2): f()
    torch.cuda.synchronize(); t0 = time.perf_counter()
    for _ in range(5): f()
    torch.cuda.synchronize(); print(name, (time.perf_counter() - t0) / 5 * 1e3, 'ms')
