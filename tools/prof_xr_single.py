"""The cross-rank sweep kernel of ONE rank of a z-slab on one GPU, for ncu (profilers take one process):
the plan of rank 1 of `world` ranks, record buffers in local memory, PB_XR_NOPOLL=3 (no records sent
or polled -- the numbers are wrong, the instruction stream and memory traffic are the kernel's own).

  PB_XR_NOPOLL=3 python tools/prof_xr_single.py 512 4 sfilterz ddz        (n, world, operators)
"""
import ctypes
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

from pyranda_b200 import ParcopPlan
from pyranda_b200._lib import OP, XRingC, check

assert os.environ.get("PB_XR_NOPOLL") == "3", "this tool only makes sense without the exchange"
n = int(sys.argv[1]) if len(sys.argv) > 1 else 512
world = int(sys.argv[2]) if len(sys.argv) > 2 else 4
ops = sys.argv[3:] or ["sfilterz", "ddz"]
L = 2 * np.pi
p = ParcopPlan(n, n, n * world, 0, L, 0, L, 0, L, periodic=(True,) * 3, pz=world, coords=(0, 0, 1), device=0)
p.set_mesh()
lib = p.L
f = p.empty_device(); f.copy_(torch.rand((n, n, n), dtype=torch.float64, device="cuda"))
out = p.empty_device()
halo = [torch.zeros(4 * n * n, dtype=torch.float64, device="cuda") for _ in range(2)]
rec = [torch.zeros(8 * n * n * 4, dtype=torch.int64, device="cuda") for _ in range(2)]
reps = int(os.environ.get("PB_REPS", "3"))
epoch = 0
for name in ops:
    code = OP[name]
    ts = []
    for it in range(reps + 1):
        epoch += 1
        x = XRingC()
        x.epoch = epoch
        x.en_in, x.st_in = rec[0].data_ptr(), rec[1].data_ptr()
        for k in range(3):
            x.en_out[k] = rec[0].data_ptr(); x.st_out[k] = rec[1].data_ptr()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        check(lib, lib.pb_z_ring(p._h, code, f.data_ptr(), halo[0].data_ptr(), halo[1].data_ptr(), out.data_ptr(), ctypes.byref(x), 0, 0.0,
                                 torch.cuda.current_stream().cuda_stream))
        e1.record(); torch.cuda.synchronize()
        if it:
            ts.append(e0.elapsed_time(e1))
    print("%-9s %.4f ms (rank 1 of %d, %d^3 slab, mode %d, no exchange)" % (name, sorted(ts)[len(ts) // 2], world, n, lib.pb_z_ring_mode(p._h, code)), flush=True)
