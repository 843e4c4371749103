"""GPU: whole decks at the sizes BASELINE.json names, device path vs the same deck driver on the CPU
oracle's operators (same dt on both sides, fields compared point by point).

  configs[0]  examples/3Dadvect.py at 64^3 periodic
  configs[1]  examples/TaylorGreen.py: 3 RK4 steps at 256^3 (the pipelined kernels, hoisted flux
              arguments, the stage kernel with the flux arithmetic inside) and 100 steps at 64^3
              (north_star: within 1e-10 after 100 RK4 steps)."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _pair(oracle_mod, mesh, eom, ic):
    from oracle_backend import make_sim
    from pyranda_b200.sim import pyrandaSim
    sims = [make_sim(oracle_mod, "ref", mesh), pyrandaSim("gpu", mesh)]
    for ss in sims:
        ss.EOM(eom)
        ss.setIC(ic)
    return sims


def _errors(ref, gpu, names, floor_from=None):
    out = {}
    for nm in names:
        a = gpu.variables[nm].cpu().numpy()
        b = ref.variables[nm]
        scale = np.abs(b).max()
        if floor_from is not None:
            scale = max(scale, np.abs(ref.variables[floor_from]).max())
        out[nm] = float(np.abs(a - b).max() / scale)
    return out


def test_advect3d_deck_64(oracle_mod):
    from decks import ADVECT3D_EOM, ADVECT3D_IC, advect3d_mesh
    ref, gpu = _pair(oracle_mod, advect3d_mesh(64), ADVECT3D_EOM, ADVECT3D_IC)
    dt = float(ref.variables["dt"]) * 0.5
    assert abs(float(gpu.variables["dt"]) * 0.5 - dt) < 1e-13 * dt
    t = [0.0, 0.0]
    for _ in range(20):
        t[0] = ref.rk4(t[0], dt)
        t[1] = gpu.rk4(t[1], dt)
    errs = _errors(ref, gpu, ("rho", "rhou", "rhov", "rhow", "Et", "p"), floor_from="rhow")
    print("3Dadvect 64^3, 20 steps:", errs)
    assert max(errs.values()) < 1e-11, errs


def test_taylor_green_256_three_steps(oracle_mod):
    from decks import TGV_EOM, TGV_IC, tgv_mesh
    ref, gpu = _pair(oracle_mod, tgv_mesh(256), TGV_EOM, TGV_IC)
    assert gpu.fuser is not None and gpu.fuser.enabled
    dt = float(ref.variables["dt"]) * 0.5
    t = [0.0, 0.0]
    for _ in range(3):
        t[0] = ref.rk4(t[0], dt)
        t[1] = gpu.rk4(t[1], dt)
    plan = gpu._flux_plan()
    assert plan["gid"] is not None and all(it["stage"] is not None for it in plan["items"])   # hoisted + fused stage ran
    errs = _errors(ref, gpu, ("rho", "rhou", "rhov", "rhow", "Et", "p", "enst"), floor_from="rhou")
    print("TGV 256^3, 3 steps:", errs)
    assert max(errs.values()) < 1e-12, errs
    assert abs(float(gpu.variables["dt"]) - float(ref.variables["dt"])) < 1e-12 * float(ref.variables["dt"])


def test_taylor_green_64_hundred_steps(oracle_mod):
    from decks import TGV_EOM, TGV_IC, tgv_mesh
    ref, gpu = _pair(oracle_mod, tgv_mesh(64), TGV_EOM, TGV_IC)
    dt = float(ref.variables["dt"]) * 0.5
    t = [0.0, 0.0]
    for _ in range(100):
        t[0] = ref.rk4(t[0], dt)
        t[1] = gpu.rk4(t[1], dt)
    errs = _errors(ref, gpu, ("rho", "rhou", "rhov", "rhow", "Et", "p"), floor_from="rhou")
    print("TGV 64^3, 100 steps:", errs)
    assert max(errs.values()) < 1e-10, errs
    # The artificial viscosities are gbar(ring(.)) of S and of div u.  The flow is nearly solenoidal:
    # div u is itself a small difference of O(1) derivatives, so the 8th-derivative detector of it works
    # on a field whose leading digits are cancellation noise and amplifies the 1e-15 relative differences
    # of the two arithmetic paths.  mu (detector of the O(1) strain rate) agrees to 1e-9 of its own
    # maximum; beta is judged by what it multiplies -- beta * div u against the pressure it is added to --
    # and only loosely (1e-4) against its own maximum.
    visc = _errors(ref, gpu, ("mu", "beta"))
    print("TGV 64^3, 100 steps, artificial viscosities:", visc)
    assert visc["mu"] < 1e-9 and visc["beta"] < 1e-4, visc
    p = np.abs(ref.variables["p"]).max()
    for nm, grad in (("mu", "S"), ("beta", "div")):
        d = np.abs(gpu.variables[nm].cpu().numpy() - ref.variables[nm]).max() * np.abs(ref.variables[grad]).max()
        assert d < 1e-12 * p, (nm, d / p)


def test_restart_and_viz_dump_on_device_fields(tmp_path):
    """§8 f4 on the device backend (pyranda.py:431-470,475-588): writeRestart / readRestart stage the
    device state through the host once and the restarted run continues bit for bit; ss.write dumps the
    device fields as legacy VTK with the reference's file names."""
    import os
    from decks import TGV_EOM, TGV_IC, tgv_mesh
    from pyranda_b200.sim import pyrandaSim
    a = pyrandaSim("tgv", tgv_mesh(64))
    a.EOM(TGV_EOM)
    a.setIC(TGV_IC)
    t = 0.0
    for _ in range(2):
        t = a.rk4(t, float(a.variables["dt"]) * 0.5)
    a.writeRestart(tmp_path / "state")
    for _ in range(2):
        t = a.rk4(t, float(a.variables["dt"]) * 0.5)
    b = pyrandaSim("tgv", tgv_mesh(64))
    tb = b.readRestart(tmp_path / "state")
    assert b.cycle == 2 and len(b.equations) == len(a.equations)
    for _ in range(2):
        tb = b.rk4(tb, float(b.variables["dt"]) * 0.5)
    assert tb == t
    for nm in ("rho", "rhou", "Et", "p", "mu"):
        assert np.array_equal(a.variables[nm].cpu().numpy(), b.variables[nm].cpu().numpy()), nm
    path = a.write(["rho", "p"], root=str(tmp_path))
    assert path.endswith(os.path.join("vis%07d" % a.cycle, "proc-000000.%07d.vtk" % a.cycle))
    raw = open(path, "rb").read()
    n = 64 ** 3
    assert b"DIMENSIONS 64 64 64" in raw
    tag = b"SCALARS p float\nLOOKUP_TABLE default\n"
    at = raw.index(tag) + len(tag)
    p = np.frombuffer(raw[at:at + 4 * n], dtype=">f4")
    want = a.variables["p"].permute(2, 1, 0).contiguous().cpu().numpy().ravel()   # x fastest
    assert np.allclose(p, want, rtol=1e-6, atol=1e-7)
