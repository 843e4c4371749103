"""Where one RK4 step of a deck goes: torch.profiler kernel table (library kernels, generated pointwise
kernels `fz`, torch's own).  python tools/prof_deck.py cylinder [nx ny nz] | rt3d [npts] | tgv [n]"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p_ in (ROOT, os.path.join(ROOT, "tests")):
    if p_ not in sys.path:
        sys.path.insert(0, p_)
import numpy as np
import torch
from torch.profiler import ProfilerActivity, profile

which = sys.argv[1]
args = [int(v) for v in sys.argv[2:]]
from pyranda_b200.sim import pyrandaSim
cfl = 1.0
if which == "cylinder":
    from decks import CYLINDER_CURV_EOM, CYLINDER_CURV_IC, zoom_mesh_1d
    nx, ny, nz = args or (1024, 512, 64)
    Lx = float(np.pi) * 2.0 * (nx - 1.0) / nx
    Ly = float(np.pi) * 2.0 * (ny - 1.0) / ny
    xS = zoom_mesh_1d(nx, -2. * Lx, 2. * Lx, -2., 2., 1.0, 4 * Lx / float(nx) * .3)
    yS = zoom_mesh_1d(ny, -2. * Ly, 2. * Ly, -2., 2., 1.0, 4 * Ly / float(ny) * .3)
    Lz = float(np.pi) * 2.0 * (nz - 1.0) / nz
    zU = np.linspace(0.0, Lz, nz)
    opt = {"coordsys": 3, "function": lambda i, j, k: (xS[i], yS[j], zU[k]), "periodic": [False, False, True], "periodicGrid": False,
           "x1": [-2 * Lx, -2 * Ly, 0.0], "xn": [2 * Lx, 2 * Ly, Lz], "nn": [nx, ny, nz]}
    ss = pyrandaSim("cylinder_curvilinear", opt)
    ss.EOM(CYLINDER_CURV_EOM)
    ss.setIC(CYLINDER_CURV_IC)
elif which == "rt3d":
    from decks import RT_EOM, RT_IC, RT_PARMS, rt_mesh, rt_xbar
    npts = args[0] if args else 256
    ss = pyrandaSim("RT_3D", rt_mesh(npts, two_d=False))
    ss.addUserDefinedFunction("xbar", rt_xbar)
    parm = RT_PARMS(npts)
    ss.EOM(RT_EOM, parm)
    np.random.seed(1234)
    ss.setIC(RT_IC, parm)
    cfl = 0.1
else:
    from decks import TGV_EOM, TGV_IC, tgv_mesh
    ss = pyrandaSim("tgv", tgv_mesh(args[0] if args else 256))
    ss.EOM(TGV_EOM)
    ss.setIC(TGV_IC)
    cfl = 0.5
t = 0.0
for _ in range(2):
    t = ss.rk4(t, float(ss.variables["dt"]) * cfl)
torch.cuda.synchronize()
import time
t0 = time.perf_counter()
t = ss.rk4(t, float(ss.variables["dt"]) * cfl)
torch.cuda.synchronize()
print("one step, wall: %.2f ms" % ((time.perf_counter() - t0) * 1e3))
with profile(activities=[ProfilerActivity.CPU, ProfilerActivity.CUDA]) as prof:
    t = ss.rk4(t, float(ss.variables["dt"]) * cfl)
    torch.cuda.synchronize()
ev = [e for e in prof.key_averages() if e.device_type == torch.autograd.DeviceType.CUDA or e.self_device_time_total > 0]
rows = sorted(((e.self_device_time_total, e.count, e.key) for e in prof.key_averages() if e.self_device_time_total > 0), reverse=True)
tot = sum(r[0] for r in rows)
print("device time in the step: %.2f ms over %d kernels/copies" % (tot / 1e3, sum(r[1] for r in rows)))
for us, cnt, key in rows[:40]:
    print("%9.2f ms %6d x  %5.1f%%  %s" % (us / 1e3, cnt, 100.0 * us / tot, key[:110]))
