"""Time operators on an (nx, ny, nz) field under the ring-kernel policies:
python tools/prof_ring.py nx ny nz periodic(0/1) mode:lines[,mode:lines...] ops..."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

from pyranda_b200 import ParcopPlan, _lib

nx, ny, nz = (int(v) for v in sys.argv[1:4])
periodic = sys.argv[4] == "1"
policies = [tuple(int(v) for v in p.split(":")) for p in sys.argv[5].split(",")]
ops = sys.argv[6:] or ["ddx", "ddy", "ddz", "sfilter", "gfilter"]
L = _lib.load()
Ls = [2 * np.pi * (k - 1) / k if periodic else 1.0 for k in (nx, ny, nz)]
p = ParcopPlan(nx, ny, nz, 0, Ls[0], 0, Ls[1], 0, Ls[2], periodic=(periodic,) * 3, device=0)
p.set_mesh()
f = p.empty_device()
f.copy_(torch.rand((nz, ny, nx), dtype=torch.float64, device="cuda").permute(2, 1, 0))
out = p.empty_device()
st = torch.cuda.current_stream().cuda_stream
npts = nx * ny * nz
print("grid", nx, ny, nz, "periodic" if periodic else "bounded", flush=True)
for mode, lines in policies:
    L.pb_set_ring(mode, lines)
    res = {}
    for name in ops:
        try:
            for _ in range(3):
                p.apply_ptr(name, f.data_ptr(), out.data_ptr(), st)
            torch.cuda.synchronize()
        except Exception as exc:  # noqa: BLE001
            print("  ring %d:%d %-9s FAILED %s" % (mode, lines, name, exc), flush=True)
            continue
        ts = []
        for _ in range(5):
            r0 = L.pb_ring_launch_count()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(5):
                p.apply_ptr(name, f.data_ptr(), out.data_ptr(), st)
            e1.record(); torch.cuda.synchronize()
            ts.append(e0.elapsed_time(e1) / 5)
            ring = (L.pb_ring_launch_count() - r0) / 5
        ms = sorted(ts)[2]
        sweeps = 3 if name in ("sfilter", "gfilter", "laplacian", "ring") else 1
        print("  ring %d:%-2d %-9s %8.3f ms  %6.1f GB/s (%.1f%% of 6553.6)  ring launches/op %.0f" % (
            mode, lines, name, ms, sweeps * 16 * npts / ms / 1e6, sweeps * 16 * npts / ms / 1e6 / 65.536, ring), flush=True)
