"""Per-operator wall time of the host-array entry point (pb_host_apply: pinned host array -> slabs in ->
sweeps -> slabs out) at n^3: python tools/prof_e2e.py [n] [ops...]

Environment: PB_HOST_SLABS (slabs per field), PB_HOST_NO_ZSTREAM=1 (the Gaussian filter's z sweep after all
slabs have arrived, as before round 2), PB_BOUNDED=1."""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

from pyranda_b200 import ParcopPlan

n = int(sys.argv[1]) if len(sys.argv) > 1 else 512
ops = sys.argv[2:] or ["ddx", "ddy", "ddz", "sfilter", "gfilter"]
periodic = os.environ.get("PB_BOUNDED", "0") != "1"
L = 2 * np.pi * (n - 1) / n if periodic else 1.0
p = ParcopPlan(n, n, n, 0, L, 0, L, 0, L, periodic=(periodic,) * 3, device=0)
p.set_mesh()
hin = torch.rand((n, n, n), dtype=torch.float64).pin_memory()
hout = torch.empty((n, n, n), dtype=torch.float64).pin_memory()
a_in, a_out = hin.numpy().T, hout.numpy().T
gb = n ** 3 * 8 / 1e9
total = 0.0
for name in ops:
    p.apply_host_into(name, a_in, a_out)
    ts = []
    for _ in range(5):
        t0 = time.perf_counter()
        p.apply_host_into(name, a_in, a_out)
        ts.append(time.perf_counter() - t0)
    ms = sorted(ts)[len(ts) // 2] * 1e3
    total += ms
    print("%-9s %8.2f ms   %.2f Gpoints/s   %.1f GB/s each way if fully overlapped" % (name, ms, n ** 3 / ms / 1e6, gb / ms * 1e3))
print("step      %8.2f ms   %.2f Gpoints/s" % (total, len(ops) * n ** 3 / total / 1e6))
# one slab-sized copy each way, alone: what the link gives
d = torch.empty((n, n, n), dtype=torch.float64, device="cuda")
for label, fn in (("H2D", lambda: d.copy_(hin, non_blocking=True)), ("D2H", lambda: hout.copy_(d, non_blocking=True))):
    fn(); torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    print("%s alone  %8.2f ms   %.1f GB/s" % (label, (time.perf_counter() - t0) / 3 * 1e3, gb / ((time.perf_counter() - t0) / 3)))
# both directions at once on two streams: the floor of an operator whose slabs stream in and out
d2 = torch.empty((n, n, n), dtype=torch.float64, device="cuda")
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
torch.cuda.synchronize()
t0 = time.perf_counter()
for _ in range(3):
    with torch.cuda.stream(s1):
        d.copy_(hin, non_blocking=True)
    with torch.cuda.stream(s2):
        hout.copy_(d2, non_blocking=True)
torch.cuda.synchronize()
dt = (time.perf_counter() - t0) / 3
print("H2D + D2H together %8.2f ms   %.1f GB/s each way" % (dt * 1e3, gb / dt))
