"""One Taylor-Green RK4 step at n^3 inside a cudaProfilerStart/Stop window (for the ncu launch list):
ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file out.csv python tools/prof_tgv.py [n]"""
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch

from decks import TGV_EOM, TGV_IC, tgv_mesh
from pyranda_b200.sim import pyrandaSim

n = int(sys.argv[1]) if len(sys.argv) > 1 else 256
ss = pyrandaSim("TGvortex", tgv_mesh(n))
ss.EOM(TGV_EOM)
ss.setIC(TGV_IC)
t, dt = 0.0, ss.variables["dt"] * 0.5
for _ in range(2):
    t = ss.rk4(t, dt); dt = ss.variables["dt"] * 0.5
torch.cuda.synchronize()
t0 = time.perf_counter()
for _ in range(3):
    t = ss.rk4(t, dt); dt = ss.variables["dt"] * 0.5
torch.cuda.synchronize()
print("ms per RK4 step (3 steps, wall):", (time.perf_counter() - t0) / 3 * 1e3, flush=True)
torch.cuda.profiler.start()
t = ss.rk4(t, dt)
torch.cuda.synchronize()
torch.cuda.profiler.stop()
