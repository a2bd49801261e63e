import torch, time, sys
sys.path.insert(0, "/root/repo")
from pybmc_b200 import _device as D
dev = torch.device("cuda", 0)
for mb in (7.2, 23.6, 100):
    n = int(mb * 1e6 / 4)
    t = torch.randn(n, device=dev)
    buf = torch.empty(n, pin_memory=True)
    for _ in range(3):
        buf.copy_(t, non_blocking=True); torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(10):
        buf.copy_(t, non_blocking=True); torch.cuda.synchronize()
    dt = (time.perf_counter() - t0) / 10
    h = torch.randn(n).pin_memory()
    for _ in range(3):
        t.copy_(h, non_blocking=True); torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(10):
        t.copy_(h, non_blocking=True); torch.cuda.synchronize()
    dt2 = (time.perf_counter() - t0) / 10
    for _ in range(3):
        a = D.to_host(t)
    t0 = time.perf_counter()
    for _ in range(10):
        a = D.to_host(t)
    dt3 = (time.perf_counter() - t0) / 10
    print(f"{mb} MB: D2H pinned {dt*1e3:.3f} ms = {mb/1e3/dt:.1f} GB/s | H2D pinned {dt2*1e3:.3f} ms = {mb/1e3/dt2:.1f} GB/s | D.to_host {dt3*1e3:.3f} ms = {mb/1e3/dt3:.1f} GB/s")
