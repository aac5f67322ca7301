"""Probe: cost of cudaHostRegister/Unregister on a 512 MiB malloc'd buffer vs pageable/pinned copy rates."""
import time

import numpy as np
import torch

rt = torch.cuda.cudart()
torch.cuda.init()
n = 512 << 20
a = np.empty(n, dtype=np.uint8)
a[:] = 1
d = torch.empty(n, dtype=torch.uint8, device="cuda")
t = torch.from_numpy(a)
for _ in range(2):
    t0 = time.perf_counter(); d.copy_(t); torch.cuda.synchronize(); t1 = time.perf_counter()
    print(f"pageable H2D 512 MiB: {(t1-t0)*1e3:.1f} ms  {n/(t1-t0)/1e9:.1f} GB/s")
for _ in range(2):
    t0 = time.perf_counter(); t.copy_(d); torch.cuda.synchronize(); t1 = time.perf_counter()
    print(f"pageable D2H 512 MiB: {(t1-t0)*1e3:.1f} ms  {n/(t1-t0)/1e9:.1f} GB/s")
for flags in (0,):
    t0 = time.perf_counter(); r = rt.cudaHostRegister(a.ctypes.data, n, flags); t1 = time.perf_counter()
    print(f"cudaHostRegister(flags={flags}) rc={r}: {(t1-t0)*1e3:.1f} ms")
    t0 = time.perf_counter(); d.copy_(t, non_blocking=True); torch.cuda.synchronize(); t1b = time.perf_counter()
    print(f"registered H2D 512 MiB: {(t1b-t0)*1e3:.1f} ms  {n/(t1b-t0)/1e9:.1f} GB/s")
    t0 = time.perf_counter(); r = rt.cudaHostUnregister(a.ctypes.data); t1 = time.perf_counter()
    print(f"cudaHostUnregister rc={r}: {(t1-t0)*1e3:.1f} ms")
p = torch.empty(n, dtype=torch.uint8, pin_memory=True)
t0 = time.perf_counter(); d.copy_(p, non_blocking=True); torch.cuda.synchronize(); t1 = time.perf_counter()
print(f"pinned H2D 512 MiB: {(t1-t0)*1e3:.1f} ms  {n/(t1-t0)/1e9:.1f} GB/s")
t0 = time.perf_counter(); p.copy_(d, non_blocking=True); torch.cuda.synchronize(); t1 = time.perf_counter()
print(f"pinned D2H 512 MiB: {(t1-t0)*1e3:.1f} ms  {n/(t1-t0)/1e9:.1f} GB/s")
# host memcpy rate (one thread) into pinned memory
pn = p.numpy()
t0 = time.perf_counter(); pn[:] = a; t1 = time.perf_counter()
print(f"host memcpy 512 MiB (1 thread): {(t1-t0)*1e3:.1f} ms  {n/(t1-t0)/1e9:.1f} GB/s")
