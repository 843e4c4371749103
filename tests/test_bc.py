"""The BC package of the interpreter (pyranda_b200/bc.py; pyrandaBC.py:40-186): plane semantics on
numpy and on Fortran-strided torch tensors, and a bounded deck through the oracle-backed driver."""
import numpy as np
import pytest

from decks import BC_EOM, BC_IC, bc_mesh
from pyranda_b200.bc import BoundaryConditions


def _fields(kind):
    rng = np.random.default_rng(3)
    a = np.asfortranarray(rng.uniform(-1, 1, size=(6, 5, 4)))
    if kind == "numpy":
        return a.copy(order="F"), a
    import torch
    t = torch.from_numpy(a.transpose(2, 1, 0).copy(order="C")).permute(2, 1, 0)  # own memory, strides (1, nx, nx*ny)
    return t, a


@pytest.mark.parametrize("kind", ["numpy", "torch"])
def test_planes(kind):
    f, a = _fields(kind)
    g, _ = _fields(kind)
    bc = BoundaryConditions({"f": f, "g": g})
    bc.extrap("f", ["x1", "yn"])
    bc.extrap(["f"], "zn", order=1)
    bc.const(["f", "g"], ["y1"], 2.5)
    bc.field("g", ["xn", "z1"], f)
    e = a.copy()
    e[0, :, :] = 2 * e[1, :, :] - e[2, :, :]
    e[:, -1, :] = 2 * e[:, -2, :] - e[:, -3, :]
    e[:, :, -1] = e[:, :, -2]
    e[:, 0, :] = 2.5
    eg = a.copy()
    eg[:, 0, :] = 2.5
    eg[-1, :, :] = e[-1, :, :]
    eg[:, :, 0] = e[:, :, 0]
    assert np.array_equal(np.asarray(f), e) and np.array_equal(np.asarray(g), eg)
    bc2 = BoundaryConditions({"f": f}, owns={"x1": False})
    before = np.asarray(f).copy()
    bc2.const("f", ["x1", "xn"], 9.0)  # not this rank's boundaries: untouched
    assert np.array_equal(np.asarray(f), before)
    with pytest.raises(ValueError):
        bc.const("f", "w1", 0.0)


@pytest.mark.parametrize("kind", ["numpy", "torch"])
def test_symmetric_planes(kind):
    """bc.symm (pyrandaBC.py:748-786): ghost planes mirror their images, optionally with the sign flipped."""
    f, a = _fields(kind)
    g, _ = _fields(kind)
    bc = BoundaryConditions({"f": f, "g": g})
    bc.symm("f", ["x1", "yn"], npts=2)
    bc.symm(["g"], "z1", anti=True, npts=2)
    e = a.copy()
    e[0:2, :, :] = e[2:4, :, :][::-1, :, :]
    e[:, -2:, :] = e[:, -4:-2, :][:, ::-1, :]
    eg = a.copy()
    eg[:, :, 0:2] = -eg[:, :, 2:4][:, :, ::-1]
    assert np.array_equal(np.asarray(f), e) and np.array_equal(np.asarray(g), eg)


def test_bounded_deck_with_bc_lines(oracle_mod):
    from oracle_backend import make_sim
    ss = make_sim(oracle_mod, "bc", bc_mesh(32))
    ss.EOM(BC_EOM)
    ss.setIC(BC_IC)
    t, dt = 0.0, 2.0e-3
    for _ in range(5):
        t = ss.rk4(t, dt)
    phi, g2 = ss.variables["phi"], ss.variables["grad2"]
    assert np.all(phi[0, :, :] == 0.0)
    assert np.array_equal(phi[-1, 1:, :], (2 * phi[-2, :, :] - phi[-3, :, :])[1:, :])
    assert np.array_equal(phi[:, 0, :], phi[:, 1, :])
    assert np.array_equal(g2[0, :, :], phi[0, :, :])
    assert np.isfinite(phi).all() and 0.5 < phi.max() < 1.1


def gen_farfield():
    import importlib.util
    import os
    spec = importlib.util.spec_from_file_location("make_bc_golden", os.path.join(os.path.dirname(__file__), "golden", "make_bc_golden.py"))
    gen = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(gen)
    return gen.FARFIELD


def _golden_case(kind, oracle_mod):
    """Fields and metrics of tests/golden/make_bc_golden.py (the reference's own BC package ran there)."""
    import importlib.util
    import os
    spec = importlib.util.spec_from_file_location("make_bc_golden", os.path.join(os.path.dirname(__file__), "golden", "make_bc_golden.py"))
    gen = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(gen)
    X, Y, Z, Xd, Yd, Zd = gen.mesh()
    o = oracle_mod.Oracle(*gen.N, 0, 1, 0, 1, 0, 1, coordsys=3, mesh_xyz=(Xd, Yd, Zd))
    f = gen.fields(X, Y, Z)
    gold = np.load(os.path.join(os.path.dirname(__file__), "golden", "bc_20x18.npz"))
    if kind == "numpy":
        return {k: a.copy(order="F") for k, a in f.items()}, o.getvar, gold
    import torch
    conv = lambda a: torch.from_numpy(np.ascontiguousarray(a.transpose(2, 1, 0))).permute(2, 1, 0)
    return {k: conv(a) for k, a in f.items()}, (lambda name: conv(o.getvar(name))), gold


@pytest.mark.parametrize("kind", ["numpy", "torch"])
def test_exit_and_slip_match_the_reference_package(kind, oracle_mod):
    """bc.exit (pyrandaBC.py:468-522) bc.slip (:186-466) and bc.farfield (:524-746) against golden planes produced by the
    reference's own package on a curvilinear 2-D grid (tests/golden/make_bc_golden.py)."""
    v, getvar, gold = _golden_case(kind, oracle_mod)
    bc = BoundaryConditions(v, getvar=getvar)
    bc.exit(["rho", "w"], ["x1", "xn", "y1", "yn"])
    bc.exit("u", ["x1", "yn"], norm=True)
    for k in ("rho", "w", "u"):
        assert np.abs(np.asarray(v[k]) - gold["exit_" + k]).max() < 1e-14, k
    v, getvar, gold = _golden_case(kind, oracle_mod)
    bc = BoundaryConditions(v, getvar=getvar)
    bc.slip([["u", "v"]], ["x1", "yn"])
    bc.slip([["u", "v", "w"]], ["xn", "y1"])
    for k in ("u", "v", "w"):
        assert np.abs(np.asarray(v[k]) - gold["slip_" + k]).max() < 1e-13, k
    # the wall-normal velocity is gone and the speed did not grow
    n1, n2, _ = bc._normals("y1")
    un = np.asarray(v["u"])[:, 0, :] * np.asarray(n1) + np.asarray(v["v"])[:, 0, :] * np.asarray(n2)
    assert np.abs(un).max() < 1e-13
    v, getvar, gold = _golden_case(kind, oracle_mod)
    v["u"][1, :, :] *= 3.0
    bc = BoundaryConditions(v, getvar=getvar)
    for d, ref in gen_farfield().items():
        bc.BCdata["farfield-properties-%s" % d] = dict(ref, rho="rho", u="u", v="v", w="w", p="p")
    bc.farfield(["yn", "x1"])
    for k in ("rho", "u", "v", "w", "p"):
        assert np.abs(np.asarray(v[k]) - gold["far_" + k]).max() < 1e-13, k
