import torch, time
n = 1<<28
h = torch.empty(n, dtype=torch.float64, pin_memory=True); h.fill_(1.0)
d = torch.empty(n, dtype=torch.float64, device='cuda')
for _ in range(2):
    torch.cuda.synchronize(); t=time.perf_counter(); d.copy_(h, non_blocking=True); torch.cuda.synchronize(); dt=time.perf_counter()-t
    print("H2D pinned GB/s", n*8/dt/1e9)
for _ in range(2):
    torch.cuda.synchronize(); t=time.perf_counter(); h.copy_(d, non_blocking=True); torch.cuda.synchronize(); dt=time.perf_counter()-t
    print("D2H pinned GB/s", n*8/dt/1e9)
t=time.perf_counter(); xs=[torch.empty(n, dtype=torch.float64, device='cuda') for _ in range(28)]; torch.cuda.synchronize(); print("alloc 28x2GiB s", time.perf_counter()-t)
