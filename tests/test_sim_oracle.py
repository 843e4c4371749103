"""The deck interpreter + RK4 driver (pyranda_b200.sim) on the numpy / oracle backend, pinned to the
reference's golden scalars for whole simulations (tolerance 1e-4, tests/run_tests.py:84)."""
import os

import numpy as np
import pytest

from decks import run_tgv, tgv_mesh
from oracle_backend import make_sim


def baseline(key):
    """A golden curve of the reference's regression suite (tests/golden/make_baseline_fixtures.py)."""
    with np.load(os.path.join(os.path.dirname(__file__), "golden", "reference_baselines.npz")) as z:
        return z[key]


def test_translate():
    from pyranda_b200.sim import translate
    s = translate(":mu: = gbar( abs(ring(:S:)) ) * :rho: * 1.0e-4")
    assert s == 'self.variables["mu"]=self.gfilter(xp.abs(self.ring(self.variables["S"])))*self.variables["rho"]*1.0e-4'
    assert translate("dt.courant(:u:,:v:,:w:,:cs:)*1.0").startswith("self.dt_courant(")
    assert translate("numpy.minimum(:dt:,0.2*dt.diff(:beta:,:rho:))") == \
        'xp.minimum(self.variables["dt"],0.2*self.dt_diff(self.variables["beta"],self.variables["rho"]))'
    assert translate("sin(meshx)*pi") == 'xp.sin(self.variables["meshx"])*xp.pi'
    assert translate("gbarx(:a:)+expo") == 'self.gfilterx(self.variables["a"])+expo'


def test_taylor_green_golden(oracle_mod):
    """tests/cases/testTaylorGreen.py:3: enstrophy ratio 1.00106718415 at t = 0.1 on 32^3."""
    ss = make_sim(oracle_mod, "TGvortex", tgv_mesh(32))
    enst, time, n = run_tgv(ss, tstop=0.1)
    # the reference accepts 1e-4 (tests/run_tests.py:84); the baseline is printed with 12 digits and is met to 5e-12
    assert abs(enst - 1.00106718415) / 1.00106718415 < 1e-10, (enst, time, n)


def test_advection_1d_golden(oracle_mod):
    """tests/cases/test1DAdvection.py:3-4 (examples/advection.py): sum((phi - phi0)^2) after one
    period of periodic 1-D advection, N = 50, 100 and 200 (the N = 300 / 400 baselines, 3e-14 and
    5e-15, are sums of squared rounding errors and pin nothing)."""
    from pyranda_b200.sim import pyrandaSim
    from oracle_backend import NumpyOracleBackend
    for npts, golden in ((50, 9.65612670412e-09), (100, 7.6111541815e-11), (200, 5.93285138337e-13)):
        L = np.pi * 2.0
        mesh = {"x1": [0.0, 0.0, 0.0], "xn": [L * (npts - 1) / npts, 1.0, 1.0], "nn": [npts, 1, 1], "periodic": [True, False, False]}
        o = oracle_mod.Oracle(npts, 1, 1, 0.0, mesh["xn"][0], 0, 1, 0, 1, periodic=(True, False, False))
        ss = pyrandaSim("advection", mesh, backend=NumpyOracleBackend(o))
        ss.EOM(" ddt(:phi:)  =  -:c: * ddx(:phi:) ")
        ss.setIC("r   = sqrt( (meshx-pi)**2  )\n:phi: = 1.0 + 0.1 * exp( -(r/(pi/4.0))**2 )\n:phi2: = 1.0*:phi:\n:c:   = 1.0")
        v = 1.0
        dt_max = v / npts * L * .90
        tt = L / v * 1.0
        dt, time = dt_max, 0.0
        while tt > time:
            time = ss.rk4(time, dt)
            dt = min(dt_max, (tt - time))
        err = float(np.sum((ss.variables["phi"] - ss.variables["phi2"]) ** 2))
        # 4e-11 / 8e-10 / 9e-9 relative: the baselines themselves are sums of squares of 1e-5 ... 1e-7 errors
        assert abs(err - golden) / golden < 1e-6, (npts, err, golden)


def test_rayleigh_taylor_2d_golden_curve(oracle_mod):
    """tests/cases/testRT.py (examples/RT3D.py 32 1 1, BASELINE config 4 in 2-D): the reference's own
    baseline file tests/baselines/RT_2D.dat -- mixing width every 20 cycles up to t = 100, 828 RK4
    steps on 44 x 32 -- through the interpreter on the oracle.  Non-periodic x with one-sided closures,
    ddx/ddy, grad, fbar, gbar, ring, the BC package, the dt package, a user-defined function and the
    seeded random3D().  The reference's tolerance is 1e-4; the curve is reproduced to ~1e-12."""
    import os
    from decks import RT_EOM, RT_IC, RT_PARMS, rt_mesh, rt_xbar
    gold = baseline("RT_2D")
    npts = 32
    ss = make_sim(oracle_mod, "RT_2D", rt_mesh(npts))
    ss.addUserDefinedFunction("xbar", rt_xbar)  # examples/RT3D.py:73-81, verbatim
    parm = RT_PARMS(npts)
    ss.EOM(RT_EOM, parm)
    np.random.seed(1234)  # examples/RT3D.py:225
    ss.setIC(RT_IC, parm)
    trapz = getattr(np, "trapezoid", None) or np.trapz
    time, cfl = 0.0, 1.0
    dt = float(ss.variables["dt"]) * cfl
    dtmax = dt * .1
    mix_w, time_w = [], []
    while time < 100.0:  # examples/RT3D.py:248-262
        time = ss.rk4(time, dt)
        dt = min(float(ss.variables["dt"]) * cfl, dtmax * 1.1)
        dtmax = dt * 1.0
        if ss.cycle % 20 == 0:
            mix_w.append(trapz(ss.variables["mix"].mean(axis=(1, 2)), ss.variables["meshx"][:, 0, 0]))
            time_w.append(time)
    assert len(time_w) == gold.shape[1] == 41
    assert np.abs(np.array(time_w) / gold[0] - 1).max() < 1e-9
    assert np.abs(np.array(mix_w) / gold[1] - 1).max() < 1e-9


def test_kelvin_helmholtz_2d_golden_curve(oracle_mod):
    """tests/cases/test2deuler.py KH-2d-64 (examples/KH.py 64 0 . 1): density along j = 3 ny / 4 at
    t = 1.5 against the reference's baseline file (tolerance there 1e-4; reproduced to ~1e-12).
    Periodic 2-D: ddx/ddy, grad, fbar, gbar, ring, the dt package, deck dictionaries, where()."""
    import os
    from decks import KH_EOM, KH_EOM_PARMS, KH_IC, KH_IC_PARMS, kh_mesh
    gold = baseline("KH-2d-64")
    npts = 64
    ss = make_sim(oracle_mod, "KH", kh_mesh(npts))
    ss.EOM(KH_EOM, KH_EOM_PARMS)
    ss.setIC(KH_IC, KH_IC_PARMS)
    t_final, dt_max, time = 1.5, 1.0, 0.0
    dt = float(ss.variables["dt"])
    while t_final > time:  # examples/KH.py:141-146
        time = ss.rk4(time, dt)
        dt = min(float(ss.variables["dt"]), 1.1 * dt)
        dt = min(dt_max, dt)
        dt = min(dt, (t_final - time))
    j = int(3 * npts / 4)
    assert np.abs(ss.variables["meshx"][:, j, 0] - gold[0]).max() < 1e-14
    assert np.abs(ss.variables["rho"][:, j, 0] - gold[1]).max() < 1e-9


@pytest.mark.parametrize("npts", [64, 128])
def test_euler_2d_sod_golden_curve(npts, oracle_mod):
    """tests/cases/test2deuler.py euler-2d-64 / euler-2d-128 (examples/euler.py, cylindrical Sod
    problem in a box): density along j = ny / 2 at t = pi / 4 against the reference's baseline files.
    Bounded x and y: one-sided closures of d1, both filters and the ring detector, bc.extrap,
    bc.const.  The curves are reproduced to the last printed digit."""
    import os
    from decks import EULER2D_EOM, EULER2D_IC, euler2d_mesh
    gold = baseline("euler-2d-%d" % npts)
    ss = make_sim(oracle_mod, "sod", euler2d_mesh(npts))
    ss.EOM(EULER2D_EOM)
    ss.setIC(EULER2D_IC)
    dt_max = 1.0 / npts * 0.75  # examples/euler.py:137-155
    tt = np.pi * 2.0 * .125
    time, dt = 0.0, dt_max
    while tt > time:
        time = ss.rk4(time, dt)
        dt = min(dt_max, (tt - time))
    j = int(npts / 2)
    assert np.abs(ss.variables["meshx"][:, j, 0] - gold[0]).max() < 1e-14
    assert np.abs(ss.variables["rho"][:, j, 0] - gold[1]).max() < 1e-12


@pytest.mark.parametrize("npts", [32, 64])
def test_cylinder_ibm_golden_curve(npts, oracle_mod):
    """tests/cases/testCylinder.py cylinder-2d-32 / -64 (examples/cylinder.py): Mach-2 flow over an
    immersed cylinder, pressure along j = ny / 2 at t = 1.5 against the reference's baseline files.
    Bounded box: one-sided closures, grad, fbar, gbar, ring, the IBM package (ibmV with a frame
    velocity, ibmS), the BC package and the dt package.  Reproduced to the last printed digit."""
    import os
    from decks import CYLINDER_EOM, CYLINDER_IC, cylinder_mesh
    gold = baseline("cylinder-2d-%d" % npts)
    ss = make_sim(oracle_mod, "cylinder_test", cylinder_mesh(npts))
    ss.EOM(CYLINDER_EOM)
    ss.setIC(CYLINDER_IC)
    tt, cfl, time = 1.5, 1.0, 0.0
    dt = float(ss.variables["dt"]) * cfl * .1  # examples/cylinder.py:147-153
    with np.errstate(all="ignore"):  # the IBM package divides by the level set on the surface, as the reference does
        while tt > time:
            time = ss.rk4(time, dt)
            dt = min(float(ss.variables["dt"]) * cfl, 1.1 * dt)
            dt = min(dt, (tt - time))
    j = int(npts / 2)
    assert np.abs(ss.variables["meshx"][:, j, 0] - gold[0]).max() < 1e-14
    assert np.abs(ss.variables["p"][:, j, 0] - gold[1]).max() < 1e-11


def test_curvilinear_cylinder_golden_curve(oracle_mod):
    """tests/cases/testCylinder.py cylinder_curved-2d-64 (examples/cylinder_curv.py, the deck of
    BASELINE config 5 at 64 x 64): velocity magnitude along j = ny / 2 at t = 3 against the
    reference's baseline file.  coordsys = 3 on the tanh-stretched zoomMesh: metrics from the compact
    derivatives of the coordinates, div through contravariant fluxes, grad, the cell-volume weighted
    filter, gbar, ring with per-point lengths, the IBM, BC and dt packages (curvilinear Courant
    branch).  Reference tolerance 1e-4; reproduced to ~1e-13."""
    import os
    from decks import CYLINDER_CURV_EOM, CYLINDER_CURV_IC, cylinder_curv_mesh
    gold = baseline("cylinder_curved-2d-64")
    npts = 64
    ss = make_sim(oracle_mod, "cylinder_curvilinear", cylinder_curv_mesh(npts))
    ss.EOM(CYLINDER_CURV_EOM)
    ss.setIC(CYLINDER_CURV_IC)
    tt, cfl, time = 3.0, 1.0, 0.0
    dt = float(ss.variables["dt"]) * cfl * .01  # examples/cylinder_curv.py:165,187-191
    with np.errstate(all="ignore"):
        while tt > time:
            time = ss.rk4(time, dt)
            dt = min(float(ss.variables["dt"]) * cfl, dt * 1.1)
            dt = min(dt, (tt - time))
    j = int(npts / 2)
    assert np.abs(ss.variables["meshx"][:, j, 0] - gold[0]).max() < 1e-13
    assert np.abs(ss.variables["umag"][:, j, 0] - gold[1]).max() < 1e-10


def test_omesh_cylinder_golden_curve(oracle_mod):
    """tests/cases/testCylinder.py cylinder_omesh-2d-64 (examples/cylinder_curv2.py): Mach-1.5 flow
    around a cylinder on an O-grid, |u| along j = ny / 2 at t = 3 against the reference's baseline
    file.  coordsys = 3 with a periodic direction whose coordinates are not periodic (periodicGrid =
    False: metrics through the one-sided first derivative, mesh.f90:251-304) and the free-slip wall
    `bc.slip` of this repository's BC package inside the step.  Reproduced to ~2e-13."""
    import os
    from decks import OMESH_EOM, OMESH_IC, cylinder_omesh
    gold = baseline("cylinder_omesh-2d-64")
    npts = 64
    ss = make_sim(oracle_mod, "cylinder_omesh", cylinder_omesh(npts))
    ss.EOM(OMESH_EOM)
    ss.setIC(OMESH_IC)
    tt, cfl, time = 3.0, 0.8, 0.0
    dt = float(ss.variables["dt"]) * cfl * .1
    with np.errstate(all="ignore"):
        while tt > time:  # examples/cylinder_curv2.py:150-167
            time = ss.rk4(time, dt)
            dt = float(ss.variables["dt"]) * cfl
            dt = min(0.2 * float(ss.variables["dtB"]), dt)
            dt = min(dt, (tt - time))
    j = int(npts / 2)
    assert np.abs(ss.variables["meshx"][:, j, 0] - gold[0]).max() < 1e-13
    assert np.abs(ss.variables["umag"][:, j, 0] - gold[1]).max() < 1e-10


@pytest.mark.parametrize("npts", [16, 64])
def test_heat_1d_steady_profile(npts, oracle_mod):
    """tests/cases/testHeat1D.py heat1D-analytic-N -- 0.0 (examples/heat1D.py): the linear steady
    profile of ddt(phi) = c lap(phi) between two bc.const ends stays put for 500 steps; the summed
    error against the analytic line is the baseline's 0.0.  Pins the bounded second derivative."""
    L = np.pi * 2.0
    Lp = L * (npts - 1.0) / npts
    mesh = {"x1": [0.0, 0.0, 0.0], "xn": [Lp, Lp, Lp], "nn": [npts, 1, 1], "periodic": [False, True, True]}
    ss = make_sim(oracle_mod, "heat_equation", mesh)
    ss.EOM("ddt(:phi:)  =  :c: * lap(:phi:)\nbc.const(['phi'],['x1'],2.0)\nbc.const(['phi'],['xn'],1.0)")
    ss.setIC("xnn = meshx[-1,0,0]\n:phi: = 1.0 + 1.0*(xnn - meshx)/xnn\n:c:   = 1.0")
    dt_max = L / npts * .005
    tt = dt_max * 500
    time, dt = 0.0, dt_max
    while tt > time:
        time = ss.rk4(time, dt)
        dt = min(dt_max, (tt - time))
    x = ss.variables["meshx"]
    anl = 1.0 + 1.0 * (x[-1, 0, 0] - x) / x[-1, 0, 0]
    assert ss.cycle >= 500 and np.sum(np.abs(anl - ss.variables["phi"])[:, 0]) < 1e-11


def test_deck_functions_reach_every_operator(oracle_mod):
    """divT, ringV, dd4x..z, lap, gbarx..z in deck lines (pyranda.py:817-858) evaluate to the operators'
    own results."""
    mesh = "xdom = (0.0, 1.0, 16)\nydom = (0.0, 1.0, 18)\nzdom = (0.0, 1.0, 20)"
    ss = make_sim(oracle_mod, "ops", mesh)
    ss.EOM("ddt(:a:) = -ddx(:a:)\n[:tx:,:ty:,:tz:] = divT(:a:,:b:,:a:,:b:,:a:,:b:,:a:,:b:,:a:)\n:r: = ringV(:a:,:b:,:a:)\n"
           ":d4: = dd4x(:a:) + dd4y(:a:) + dd4z(:a:)\n:g: = gbarx(:a:) + gbary(:a:) + gbarz(:a:) + lap(:b:)")
    ss.setIC(":a: = sin(3.0*meshx)*cos(2.0*meshy)*cos(meshz)\n:b: = meshx*meshy + cos(4.0*meshz)")
    o, a, b = ss.B.o, ss.variables["a"], ss.variables["b"]
    for got, ref in zip((ss.variables[k] for k in ("tx", "ty", "tz")), o.divergencetensor(a, b, a, b, a, b, a, b, a)):
        assert np.array_equal(got, ref)
    assert np.array_equal(ss.variables["r"], o.pringv(a, b, a))
    assert np.array_equal(ss.variables["d4"], o.dd4x(a) + o.dd4y(a) + o.dd4z(a))
    assert np.array_equal(ss.variables["g"], o.gfilterdir(a, 1) + o.gfilterdir(a, 2) + o.gfilterdir(a, 3) + o.plaplacian(b))
    ss2 = make_sim(oracle_mod, "names", mesh)
    ss2.EOM("ddt(:a:) = -ddx(:a:)\n:k: = meshi + 100.0*meshj + 10000.0*meshk\n:v: = meshVar('CellVol')")
    ss2.setIC(":a: = meshx")
    i, j, k = np.meshgrid(np.arange(16), np.arange(18), np.arange(20), indexing="ij")
    assert np.array_equal(ss2.variables["k"], i + 100.0 * j + 10000.0 * k)
    assert np.array_equal(ss2.variables["v"], o.getvar("CellVol"))


def test_viz_dump_vtk(oracle_mod, tmp_path):
    """ss.write (pyranda.py:431-470): a binary legacy-VTK structured grid per dump with the
    reference's directory / file names and the .visit index; read back here."""
    ss = make_sim(oracle_mod, "viz", "xdom = (0.0, 1.0, 16)\nydom = (0.0, 2.0, 18)\nzdom = (0.0, 3.0, 20)")
    ss.EOM("ddt(:a:) = -ddx(:a:)\n:b: = 2.0*:a:")
    ss.setIC(":a: = sin(3.0*meshx)*cos(2.0*meshy)*cos(meshz)")
    ss.rk4(0.0, 1.0e-3)
    path = ss.write(["a", "b"], root=str(tmp_path))
    assert path.endswith(os.path.join("vis0000001", "proc-000000.0000001.vtk"))
    raw = open(path, "rb").read()
    assert raw.startswith(b"# vtk DataFile Version 3.0") and b"DIMENSIONS 16 18 20" in raw
    n = 16 * 18 * 20
    at = raw.index(b"POINTS %d float\n" % n) + len(b"POINTS %d float\n" % n)
    pts = np.frombuffer(raw[at:at + 12 * n], dtype=">f4").reshape(n, 3)
    assert np.allclose(pts[:, 1], ss.variables["meshy"].ravel(order="F"), rtol=1e-6)
    at = raw.index(b"SCALARS b float\nLOOKUP_TABLE default\n") + len(b"SCALARS b float\nLOOKUP_TABLE default\n")
    b = np.frombuffer(raw[at:at + 4 * n], dtype=">f4")
    assert np.allclose(b, ss.variables["b"].ravel(order="F"), rtol=1e-6, atol=1e-7)
    visit = open(os.path.join(str(tmp_path), "pyranda.visit")).read()
    assert "!NBLOCKS 1" in visit and "proc-000000.0000001.vtk" in visit


def test_restart_roundtrip(oracle_mod, tmp_path):
    """writeRestart / readRestart (pyranda.py:475-588): a restarted run continues bit for bit."""
    from decks import TGV_EOM, TGV_IC, tgv_mesh
    from oracle_backend import make_sim
    a = make_sim(oracle_mod, "tgv", tgv_mesh(16))
    a.EOM(TGV_EOM)
    a.setIC(TGV_IC)
    t = 0.0
    for _ in range(2):
        t = a.rk4(t, float(a.variables["dt"]) * 0.5)
    a.writeRestart(tmp_path / "state")
    for _ in range(2):
        t = a.rk4(t, float(a.variables["dt"]) * 0.5)
    b = make_sim(oracle_mod, "tgv", tgv_mesh(16))
    tb = b.readRestart(tmp_path / "state")
    assert b.cycle == 2 and len(b.equations) == len(a.equations)
    for _ in range(2):
        tb = b.rk4(tb, float(b.variables["dt"]) * 0.5)
    assert tb == t
    for nm in ("rho", "rhou", "Et", "p", "mu"):
        assert np.array_equal(a.variables[nm], b.variables[nm]), nm
    # the reference's call shape: no arguments -> restart_<cycle>/proc-NNNNN.npz (pyranda.py:475-490)
    import os
    cwd = os.getcwd()
    os.chdir(tmp_path)
    try:
        written = a.writeRestart()
        assert written == os.path.join("restart_%s" % str(a.cycle).zfill(7), "proc-00000.npz") and os.path.exists(written)
        c = make_sim(oracle_mod, "tgv", tgv_mesh(16))
        c.cycle = a.cycle
        assert c.readRestart() == t and np.array_equal(c.variables["rho"], a.variables["rho"])
    finally:
        os.chdir(cwd)
