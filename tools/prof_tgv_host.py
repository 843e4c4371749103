"""How much of a Taylor-Green RK4 step is the host enqueueing it: wall time per step at 64^3 (the GPU work is
negligible, the figure is the host floor of the interpreter + ctypes + launches), 128^3 and 256^3, and at each size
the time the host needs to enqueue a step (no synchronisation inside) beside the synchronised step time."""
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch

from decks import TGV_EOM, TGV_IC, tgv_mesh
from pyranda_b200.sim import pyrandaSim

for n in (64, 128, 256):
    ss = pyrandaSim("TGvortex", tgv_mesh(n))
    ss.EOM(TGV_EOM)
    ss.setIC(TGV_IC)
    t, dt = 0.0, float(ss.variables["dt"]) * 0.5
    for _ in range(2):
        t = ss.rk4(t, dt)
    torch.cuda.synchronize()
    K = 5
    t0 = time.perf_counter()
    for _ in range(K):
        t = ss.rk4(t, dt)  # dt is a host float here: no device read inside the step
    t_enq = (time.perf_counter() - t0) / K
    torch.cuda.synchronize()
    t_all = (time.perf_counter() - t0) / K
    print("n = %3d   host enqueue %.2f ms per step   step (synchronised after %d steps) %.2f ms" % (n, t_enq * 1e3, K, t_all * 1e3), flush=True)
    del ss
    torch.cuda.empty_cache()
