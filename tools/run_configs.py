"""BASELINE configs[3] and configs[4] at size (also imported by bench.py):

  rt3d      examples/RT3D.py, Rayleigh-Taylor, 1.5 Npts x Npts x Npts, non-periodic x with one-sided closures,
            gfilter artificial diffusivities, BC + dt packages, the deck's own user function; z-slab over the
            ranks of the process group (torchrun) or one GPU
  cylinder  examples/cylinder_curv.py deck on its zoomed curvilinear mesh extruded in z: nx x ny x nz, metric-term
            derivatives (curvilinear div / grad / cell-volume weighted filter / ring), IBM + BC packages; one GPU

  python tools/run_configs.py cylinder [nx ny nz steps]
  torchrun --nproc-per-node 8 tools/run_configs.py rt3d [npts steps]
"""
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p_ in (ROOT, os.path.join(ROOT, "tests")):
    if p_ not in sys.path:
        sys.path.insert(0, p_)
import numpy as np


def _time_steps(ss, steps, cfl, sync):
    t, dt = 0.0, float(ss.variables["dt"]) * cfl
    t = ss.rk4(t, dt)  # warm-up (kernel compilation, plans)
    dt = float(ss.variables["dt"]) * cfl
    sync()
    t0 = time.perf_counter()
    for _ in range(steps):
        t = ss.rk4(t, dt)
        dt = float(ss.variables["dt"]) * cfl
    sync()
    return (time.perf_counter() - t0) / steps, t


def run_rt3d(npts=256, steps=3):
    """Returns a dict (rank 0) with ms per RK4 step of the RT 3-D deck on the current process group."""
    import torch
    import torch.distributed as dist
    from decks import RT_EOM, RT_IC, RT_PARMS, rt_mesh, rt_xbar
    world = dist.get_world_size() if dist.is_initialized() else 1
    mesh = rt_mesh(npts, two_d=False)
    if world > 1:
        from pyranda_b200.distributed import distributed_sim
        ss = distributed_sim("RT_3D", mesh, device=torch.cuda.current_device())
    else:
        from pyranda_b200.sim import pyrandaSim
        ss = pyrandaSim("RT_3D", mesh, device=torch.cuda.current_device())
    ss.addUserDefinedFunction("xbar", rt_xbar)
    parm = RT_PARMS(npts)
    ss.EOM(RT_EOM, parm)
    np.random.seed(1234)
    ss.setIC(RT_IC, parm)

    def sync():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
    sec, t = _time_steps(ss, steps, 0.1, sync)
    tt = torch.tensor([sec], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
    nn = (ss.nx, ss.ny, ss.nz)
    rho = ss.variables["rho"]
    ok = bool(torch.isfinite(rho).all().item())
    return {"config": "examples/RT3D.py deck, %d x %d x %d, x non-periodic (one-sided closures), z-slab x%d (BASELINE configs[3])" % (nn + (world,)),
            "grid": list(nn), "n_gpus": world, "steps": steps, "ms_per_rk4_step": tt.item() * 1e3,
            "gpoints_per_s": nn[0] * nn[1] * nn[2] / tt.item() / 1e9, "finite": ok, "time": t,
            "note": "the deck's user function `xbar` is the reference's own Python loop over x (one host read per plane)"}


def run_cylinder(nx=1024, ny=512, nz=64, steps=3):
    """The curvilinear cylinder deck on an extruded mesh + per-operator timings with metric fields."""
    import torch
    from decks import CYLINDER_CURV_EOM, CYLINDER_CURV_IC, zoom_mesh_1d
    from pyranda_b200.sim import pyrandaSim
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
    sec, t = _time_steps(ss, steps, 1.0, torch.cuda.synchronize)
    ok = bool(torch.isfinite(ss.variables["rho"]).all().item())
    # operators with metric fields, device resident
    plan = ss.B.plan
    f, g = ss.variables["rho"], ss.variables["p"]
    npts = nx * ny * nz
    # algorithmic bytes per point: contraction with the nine inverse-metric fields + determinant + three sweeps
    cases = {"div (curvilinear)": (lambda: plan.divergence(f, g, f), 112),
             "grad (curvilinear)": (lambda: plan.grads(f), 3 * 16 + 9 * 8 + 48),
             "filter (cell-volume weighted)": (lambda: plan.sfilter(f), 48 + 24 + 24),
             "ring (per-point length scales)": (lambda: plan.pring(f), 3 * 24)}
    ops = {}
    for name, (fn, bpp) in cases.items():
        for _ in range(2):
            fn()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(5):
            fn()
        e1.record(); torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 5
        ops[name] = {"ms": ms, "algorithmic_bytes_per_point": bpp, "algorithmic_GBps": bpp * npts / (ms * 1e-3) / 1e9}
    return {"config": "examples/cylinder_curv.py deck on its zoomed curvilinear mesh extruded in z, %d x %d x %d (BASELINE configs[4])" % (nx, ny, nz),
            "grid": [nx, ny, nz], "steps": steps, "ms_per_rk4_step": sec * 1e3, "gpoints_per_s": npts / sec / 1e9, "finite": ok,
            "per_op": ops}


if __name__ == "__main__":
    import json
    import torch
    which = sys.argv[1]
    if which == "rt3d":
        import torch.distributed as dist
        local = int(os.environ.get("LOCAL_RANK", "0"))
        torch.cuda.set_device(local)
        if int(os.environ.get("WORLD_SIZE", "1")) > 1:
            os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")
            dist.init_process_group("nccl", device_id=torch.device("cuda", local))
        res = run_rt3d(*(int(v) for v in sys.argv[2:4]))
        if int(os.environ.get("RANK", "0")) == 0:
            print(json.dumps(res), flush=True)
        if dist.is_initialized():
            dist.destroy_process_group()
    else:
        print(json.dumps(run_cylinder(*(int(v) for v in sys.argv[2:6]))), flush=True)
