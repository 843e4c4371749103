"""Run each benchmark operator a few times at n^3 (for ncu captures): python tools/prof_ops.py [n] [ops...]"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

from pyranda_b200 import ParcopPlan, _lib

if os.environ.get("PB_LIB"):  # A/B experiments: another build of the library
    _lib.LIB_PATH = os.environ["PB_LIB"]

n = int(sys.argv[1]) if len(sys.argv) > 1 else 512
ops = sys.argv[2:] or ["ddx", "ddy", "ddz", "sfilter", "gfilter"]
periodic = os.environ.get("PB_BOUNDED", "0") != "1"
if os.environ.get("PB_TUNE"):
    a, b, c = (int(v) for v in os.environ["PB_TUNE"].split(","))
    _lib.load().pb_set_tuning(a, b, c)
L = 2 * np.pi * (n - 1) / n if periodic else 1.0
p = ParcopPlan(n, n, n, 0, L, 0, L, 0, L, periodic=(periodic,) * 3, device=0)
p.set_mesh()
f = p.empty_device()
f.copy_(torch.rand((n, n, n), dtype=torch.float64, device="cuda"))
out = p.empty_device()
reps = int(os.environ.get("PB_REPS", "3"))
for name in ops:
    for _ in range(reps):
        p.apply_ptr(name, f.data_ptr(), out.data_ptr(), torch.cuda.current_stream().cuda_stream)
torch.cuda.synchronize()
if os.environ.get("PB_TIME"):
    # several interleaved batches per operator; the median batch is reported (boxes and clocks drift)
    batches = {name: [] for name in ops}
    for _ in range(5):
        for name in ops:
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(10):
                p.apply_ptr(name, f.data_ptr(), out.data_ptr(), torch.cuda.current_stream().cuda_stream)
            e1.record(); torch.cuda.synchronize()
            batches[name].append(e0.elapsed_time(e1) / 10)
    for name in ops:
        ms = sorted(batches[name])[len(batches[name]) // 2]
        sweeps = 3 if name in ("sfilter", "gfilter", "laplacian", "ring") else 1
        print("%-9s %8.3f ms  %7.1f Gpts/s  %6.1f GB/s algorithmic (%.1f%% of 6553.6)" % (
            name, ms, n ** 3 / ms / 1e6, sweeps * 16 * n ** 3 / ms / 1e6, sweeps * 16 * n ** 3 / ms / 1e6 / 65.536))
