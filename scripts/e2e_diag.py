import os, sys, time, ctypes as C
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
n, T, P = 256, 333, 3000
host = torch.empty((n, T, P), dtype=torch.float32).pin_memory()
dev = torch.empty_like(host, device='cuda')
torch.cuda.synchronize()
for _ in range(3):
    t0 = time.perf_counter(); dev.copy_(host, non_blocking=True); torch.cuda.synchronize(); dt = time.perf_counter() - t0
print('torch pinned H2D 1.02GB: %.1f ms  %.1f GB/s' % (dt * 1e3, host.numel() * 4 / dt / 1e9))
pageable = torch.empty((n, T, P), dtype=torch.float32)
t0 = time.perf_counter(); dev.copy_(pageable); torch.cuda.synchronize(); dt = time.perf_counter() - t0
print('torch pageable H2D: %.1f ms  %.1f GB/s' % (dt * 1e3, host.numel() * 4 / dt / 1e9))
