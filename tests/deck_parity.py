"""Shared by the emulated (CPU) and the GPU test: the reference's 2-D example decks through the device
backend of the interpreter (`pyranda_b200.sim.CudaBackend` on a ParcopPlan) against the same deck on
the oracle-backed driver, a few RK4 steps from the same initial condition.

Decks: examples/RT3D.py in 2-D (BASELINE config 4: bounded x, BC package, sponge), examples/
cylinder_curv.py (BASELINE config 5: curvilinear zoom mesh, IBM + BC packages) and examples/
cylinder_curv2.py (O-grid, periodic direction with a non-periodic grid, bc.slip); and a box with
symmetry planes (SYMM operators through div / grad / lap / filters / ring / dd4)."""
import numpy as np

from decks import (CYLINDER_CURV_EOM, CYLINDER_CURV_IC, OMESH_EOM, OMESH_IC, RT_EOM, RT_IC, RT_PARMS, cylinder_curv_mesh,
                   cylinder_omesh, rt_mesh, rt_xbar, SYMM_EOM, SYMM_IC, symm_box_mesh)
from oracle_backend import make_sim

CASES = {
    "symm_box": lambda n: (symm_box_mesh(), SYMM_EOM, SYMM_IC, None, ("phi", "gx", "gy", "gz", "r")),
    "RT_2D": lambda n: (rt_mesh(n), RT_EOM, RT_IC, RT_PARMS(n), ("rho", "Yh", "Et", "p", "mybar")),
    "RT_3D": lambda n: (rt_mesh(n, two_d=False), RT_EOM, RT_IC, RT_PARMS(n), ("rho", "Yh", "Et", "p")),
    "cylinder_curv": lambda n: (cylinder_curv_mesh(n), CYLINDER_CURV_EOM, CYLINDER_CURV_IC, None, ("rho", "u", "v", "p")),
    "cylinder_omesh": lambda n: (cylinder_omesh(n), OMESH_EOM, OMESH_IC, None, ("rho", "u", "v", "p")),
}


def device_sim(name, mesh, **plan_kw):
    from pyranda_b200 import ParcopPlan
    from pyranda_b200.sim import CudaBackend, curvilinear_coordinates, parse_mesh, pyrandaSim
    opt = parse_mesh(mesh) if isinstance(mesh, str) else mesh
    cs = int(opt.get("coordsys", 0))
    plan = ParcopPlan(*opt["nn"], opt["x1"][0], opt["xn"][0], opt["x1"][1], opt["xn"][1], opt["x1"][2], opt["xn"][2],
                      periodic=tuple(opt["periodic"]), coordsys=cs,
                      symmetric=tuple(tuple(bool(b) for b in pair) for pair in opt.get("symmetric", ((False, False),) * 3)), **plan_kw)
    if cs == 3:
        plan.set_mesh(*curvilinear_coordinates(opt), periodic_grid=bool(opt.get("periodicGrid", True)))
    else:
        plan.set_mesh()
    return pyrandaSim(name, opt, backend=CudaBackend(plan))


def worst_difference(case, npts, oracle_mod, nsteps=5, **plan_kw):
    """max over the listed variables of |device - oracle| / max|oracle| after `nsteps` RK4 steps."""
    mesh, eom, ic, parm, names = CASES[case](npts)
    sims = [make_sim(oracle_mod, case, mesh), device_sim(case, mesh, **plan_kw)]
    for ss in sims:
        ss.addUserDefinedFunction("xbar", rt_xbar)
        ss.EOM(eom, parm)
        np.random.seed(1234)
        ss.setIC(ic, parm)
    t = [0.0, 0.0]
    with np.errstate(all="ignore"):
        for _ in range(nsteps):
            dt = float(sims[0].variables["dt"]) * 0.1
            for k, ss in enumerate(sims):
                t[k] = ss.rk4(t[k], dt)
    worst = 0.0
    for nm in names:
        a = sims[0].variables[nm]
        b = sims[1].variables[nm].cpu().numpy()
        worst = max(worst, float(np.abs(a - b).max() / np.abs(a).max()))
    return worst
