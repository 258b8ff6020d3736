import torch, time
d = torch.empty(8294400//4, dtype=torch.float32, device='cuda'); h = torch.empty_like(d, device='cpu').pin_memory()
s = torch.cuda.Stream()
for n in (1, 20):
    torch.cuda.synchronize(); t0 = time.perf_counter()
    with torch.cuda.stream(s):
        for _ in range(n*10): h.copy_(d, non_blocking=True)
    torch.cuda.synchronize(); dt = (time.perf_counter()-t0)/(n*10)
print("D2H 8.3MB pinned: %.3f ms -> %.1f GB/s" % (dt*1e3, 8.2944e6/dt/1e9))
h2 = torch.empty(1719360//4, dtype=torch.float32).pin_memory(); d2 = torch.empty_like(h2, device='cuda')
torch.cuda.synchronize(); t0 = time.perf_counter()
for _ in range(200): d2.copy_(h2, non_blocking=True)
torch.cuda.synchronize(); dt = (time.perf_counter()-t0)/200
print("H2D 1.7MB pinned: %.3f ms -> %.1f GB/s" % (dt*1e3, 1.71936e6/dt/1e9))
