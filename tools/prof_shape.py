"""Time operators on an arbitrary (nx, ny, nz) periodic field: python tools/prof_shape.py nx ny nz [ops...]"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

from pyranda_b200 import ParcopPlan

nx, ny, nz = (int(v) for v in sys.argv[1:4])
ops = sys.argv[4:] or ["ddx", "ddy", "ddz"]
Ls = [2 * np.pi * (k - 1) / k for k in (nx, ny, nz)]
p = ParcopPlan(nx, ny, nz, 0, Ls[0], 0, Ls[1], 0, Ls[2], periodic=(True,) * 3, device=0)
p.set_mesh()
f = p.empty_device()
f.copy_(torch.rand((nz, ny, nx), dtype=torch.float64, device="cuda").permute(2, 1, 0))
out = p.empty_device()
st = torch.cuda.current_stream().cuda_stream
for name in ops:
    for _ in range(3):
        p.apply_ptr(name, f.data_ptr(), out.data_ptr(), st)
torch.cuda.synchronize()
npts = nx * ny * nz
batches = {name: [] for name in ops}
for _ in range(5):
    for name in ops:
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(5):
            p.apply_ptr(name, f.data_ptr(), out.data_ptr(), st)
        e1.record(); torch.cuda.synchronize()
        batches[name].append(e0.elapsed_time(e1) / 5)
for name in ops:
    ms = sorted(batches[name])[2]
    sweeps = 3 if name in ("sfilter", "gfilter", "laplacian", "ring") else 1
    print("%-9s %8.3f ms  %7.1f Gpts/s  %6.1f GB/s algorithmic (%.1f%% of 6553.6)" % (
        name, ms, npts / ms / 1e6, sweeps * 16 * npts / ms / 1e6, sweeps * 16 * npts / ms / 1e6 / 65.536))
