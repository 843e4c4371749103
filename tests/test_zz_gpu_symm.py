"""SYMM boundaries on the CUDA path (compact.f90:77-91, stencils.f90:2390-2453):
parity against the oracle, and the size-independent property that a symmetry plane is the interior
stencil applied to the mirrored field.  (The file name sorts last on purpose: newest GPU tests run
after the established ones.)
"""
import numpy as np
import pytest

from conftest import domain, rel_linf, synthetic_field

pytestmark = pytest.mark.gpu

TOL = 1e-12
UNARY = ["ddx", "ddy", "ddz", "dd8x", "dd8y", "dd8z", "d2x", "d2y", "d2z", "plaplacian", "pring", "sfilter", "gfilter",
         "dd4x", "dd4y", "dd4z"]
CASES = [((True, True), (True, False), (False, True)), ((True, False), (False, False), (True, True))]


@pytest.mark.parametrize("symm", CASES)
@pytest.mark.parametrize("n", [(64, 48, 32), (256, 256, 32)])  # the second size runs the TMA kernels
def test_symmetry_planes_match_oracle(n, symm, oracle_mod):
    from pyranda_b200 import ParcopPlan
    (x1, xn), (y1, yn), (z1, zn) = domain(n, False)
    o = oracle_mod.Oracle(*n, x1, xn, y1, yn, z1, zn, periodic=(False,) * 3, symmetric=symm)
    p = ParcopPlan(*n, x1, xn, y1, yn, z1, zn, periodic=(False,) * 3, symmetric=symm)
    p.set_mesh()
    f = synthetic_field(o.getvar("x"), o.getvar("y"), o.getvar("z"))
    for name in UNARY:
        assert rel_linf(getattr(p, name)(f), getattr(o, name)(f)) < TOL, (name, n)
    for d, name in enumerate(("ddx_odd", "ddy_odd", "ddz_odd")):
        assert rel_linf(getattr(p, name)(f), o.dir_op("d1", d, f, bc=-1)) < TOL, (name, n)
    for d, name in enumerate(("dd8x_odd", "dd8y_odd", "dd8z_odd")):
        assert rel_linf(getattr(p, name)(f), o.dir_op("d8", d, f, bc=-1)) < TOL, (name, n)
    g = np.asfortranarray(np.cos(2 * f) + 0.3 * f)
    h = np.asfortranarray(f * f - 0.5)
    assert rel_linf(p.divergence(f, g, h), o.divergence(f, g, h)) < TOL
    ins = (f, g, h, 2 * g, f + h, -f, h * g, 0.5 * f, g - h)
    for a, b in zip(p.divergencetensor(*ins), o.divergencetensor(*ins)):
        assert rel_linf(a, b) < TOL
    assert rel_linf(p.pringv(f, g, h), o.pringv(f, g, h)) < TOL


@pytest.mark.parametrize("N", [64, 512])
def test_symmetry_equals_mirrored_periodic_line(N):
    """cos(kx) / sin(kx) on N cells of [0, pi] with SYMM ends against the periodic operators on 2N
    cells of [0, 2 pi] -- no oracle involved, any size."""
    import torch
    from pyranda_b200 import ParcopPlan
    ny = nz = 32
    dx = np.pi / N
    per = ParcopPlan(2 * N, ny, nz, dx / 2, 2 * np.pi - dx / 2, 0, 1, 0, 1, periodic=(True, False, False))
    sym = ParcopPlan(N, ny, nz, dx / 2, np.pi - dx / 2, 0, 1, 0, 1, periodic=(False,) * 3,
                     symmetric=((True, True), (False, False), (False, False)))
    per.set_mesh(); sym.set_mesh()
    X = per.getvar("x")
    amp = np.random.default_rng(7).uniform(-1, 1, size=6)
    even = np.asfortranarray(sum(a * np.cos(k * X) for k, a in enumerate(amp)))
    odd = np.asfortranarray(sum(a * np.sin((k + 1) * X) for k, a in enumerate(amp)))
    # the 8th derivative of smooth modes cancels ~ (N / 6)^8 / 4e3 digits: it is compared on the filters' scale instead
    for name, tol in (("ddx", 1e-11), ("d2x", 1e-10), ("sfilter", 1e-12), ("gfilter", 1e-12)):
        assert rel_linf(getattr(sym, name)(np.asfortranarray(even[:N])), getattr(per, name)(even)[:N]) < tol, (name, N)
    assert rel_linf(sym.ddx_odd(np.asfortranarray(odd[:N])), per.ddx(odd)[:N]) < 1e-11


@pytest.mark.parametrize("periodic", [True, False])
@pytest.mark.parametrize("n", [(64, 48, 32), (256, 256, 32), (1040, 32, 32)])
def test_fourth_derivative(n, periodic, oracle_mod):
    """dd4x/dd4y/dd4z (parcop.f90:255-277, stencils.f90:430-513): explicit, 7-point, no metric scale;
    host arrays and device tensors; on a periodic axis also against the stencil's transfer function."""
    import torch
    from pyranda_b200 import ParcopPlan
    (x1, xn), (y1, yn), (z1, zn) = domain(n, periodic)
    o = oracle_mod.Oracle(*n, x1, xn, y1, yn, z1, zn, periodic=(periodic,) * 3)
    p = ParcopPlan(*n, x1, xn, y1, yn, z1, zn, periodic=(periodic,) * 3)
    p.set_mesh()
    f = synthetic_field(o.getvar("x"), o.getvar("y"), o.getvar("z"))
    t = p.empty_device()
    t.copy_(torch.from_numpy(f))
    for name in ("dd4x", "dd4y", "dd4z"):
        ref = getattr(o, name)(f)
        assert rel_linf(getattr(p, name)(f), ref) < TOL, (name, n, periodic)
        assert rel_linf(getattr(p, name)(t).cpu().numpy(), ref) < TOL, (name, n, periodic, "device")
    if periodic:
        a, b, c, d = 28.0 / 3.0, -6.5, 2.0, -1.0 / 6.0
        X, k = o.getvar("x"), 5
        th = k * p.dx
        T = a + 2 * b * np.cos(th) + 2 * c * np.cos(2 * th) + 2 * d * np.cos(3 * th)
        g = np.asfortranarray(np.sin(k * X))
        assert np.abs(p.dd4x(g) - T * g).max() < 1e-12


def test_curvilinear_tensor_divergence_and_vector_ring(oracle_mod):
    """coordsys = 3: divT is the curvilinear divergence of each row (operators.f90:176-179), ringV
    scales by the per-point d1 / d2 / d3 (:680-683)."""
    import torch
    from pyranda_b200 import ParcopPlan
    n = (64, 48, 32)
    xs = [np.linspace(0, 1, k) for k in n]
    X, Y, Z = np.meshgrid(*xs, indexing="ij")
    Xd = X + 0.05 * np.sin(2 * np.pi * Y) * np.sin(np.pi * X)
    Yd = Y + 0.04 * np.sin(2 * np.pi * X) * Z
    Zd = Z * (1 + 0.1 * X)
    o = oracle_mod.Oracle(*n, 0, 1, 0, 1, 0, 1, coordsys=3, mesh_xyz=(Xd, Yd, Zd))
    p = ParcopPlan(*n, 0, 1, 0, 1, 0, 1, coordsys=3)
    p.set_mesh(Xd, Yd, Zd)
    f = synthetic_field(X, Y, Z)
    g = np.asfortranarray(np.cos(2 * f) + 0.3 * f)
    h = np.asfortranarray(f * f - 0.5)
    ins = (f, g, h, 2 * g, f + h, -f, h * g, 0.5 * f, g - h)
    ref = o.divergencetensor(*ins)
    for a, b in zip(p.divergencetensor(*ins), ref):
        assert rel_linf(a, b) < 1e-11
    dev = []
    for a in ins:
        t = p.empty_device(); t.copy_(torch.from_numpy(a)); dev.append(t)
    for a, b in zip(p.divergencetensor(*dev), ref):
        assert rel_linf(a.cpu().numpy(), b) < 1e-11
    assert rel_linf(p.pringv(f, g, h), o.pringv(f, g, h)) < 1e-11
    assert rel_linf(p.pringv(*dev[:3]).cpu().numpy(), o.pringv(f, g, h)) < 1e-11


def test_exit_slip_farfield_on_device_fields(oracle_mod):
    """bc.exit / bc.slip / bc.farfield (pyrandaBC.py:186-746) on CUDA tensors with Fortran strides against the
    golden planes of the reference's own package (tests/golden/make_bc_golden.py)."""
    import torch
    from pyranda_b200.bc import BoundaryConditions
    from test_bc import _golden_case
    from test_bc import gen_farfield
    for case in ("exit", "slip", "far"):
        v, getvar, gold = _golden_case("torch", oracle_mod)
        if case == "far":
            v["u"][1, :, :] *= 3.0
        dev = {k: a.permute(2, 1, 0).contiguous().cuda().permute(2, 1, 0) for k, a in v.items()}
        bc = BoundaryConditions(dev, getvar=lambda name: getvar(name).permute(2, 1, 0).contiguous().cuda().permute(2, 1, 0))
        if case == "exit":
            bc.exit(["rho", "w"], ["x1", "xn", "y1", "yn"])
            bc.exit("u", ["x1", "yn"], norm=True)
            names = ("rho", "w", "u")
        elif case == "far":
            for d, ref in gen_farfield().items():
                bc.BCdata["farfield-properties-%s" % d] = dict(ref, rho="rho", u="u", v="v", w="w", p="p")
            bc.farfield(["yn", "x1"])
            names = ("rho", "u", "v", "w", "p")
        else:
            bc.slip([["u", "v"]], ["x1", "yn"])
            bc.slip([["u", "v", "w"]], ["xn", "y1"])
            names = ("u", "v", "w")
        for k in names:
            assert dev[k].is_cuda and np.abs(dev[k].cpu().numpy() - gold[case + "_" + k]).max() < 1e-13, (case, k)


@pytest.mark.parametrize("case", ["RT_2D", "RT_3D", "cylinder_curv", "cylinder_omesh", "symm_box"])
def test_example_decks_on_the_gpu(case, oracle_mod):
    """BASELINE configs 4 and 5 as 2-D decks (examples/RT3D.py, examples/cylinder_curv.py) and the
    O-grid deck with bc.slip: five RK4 steps on device-resident fields (fused pointwise kernels, BC /
    IBM / dt packages on the device) against the oracle-backed driver, whose full runs reproduce the
    reference's golden curves (tests/test_sim_oracle.py).  north_star tolerance for 100 steps: 1e-10."""
    from deck_parity import worst_difference
    assert worst_difference(case, 32 if case == "RT_3D" else 64, oracle_mod) < 1e-10
