"""The real CUDA kernel sources, compiled for the host against tests/emul/cuda_emul.h (one
std::thread per CUDA thread, std::barrier for __syncthreads), checked against the oracle.

This is how indexing / barrier placement of kernels.cu and the dispatch logic of api.cu are
debugged in the GPU-less development container.  The emulated library is test infrastructure: the
product never loads it (pyranda_b200._lib only knows libparcop_b200.so).
"""
import os
import subprocess

import numpy as np
import pytest

from conftest import domain, rel_linf, synthetic_field

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
EMUL = os.path.join(ROOT, "tests", "emul")


@pytest.fixture(scope="module")
def emul_lib():
    subprocess.check_call(["make", "-C", EMUL, "-s"])
    from pyranda_b200 import _lib
    L = _lib.load(os.path.join(EMUL, "libparcop_emul.so"))
    L.pb_set_tuning(16, 16, 16)  # short chunks so that small grids still exercise P > 1
    return L


def _pair(n, periodic, oracle_mod, lib):
    from pyranda_b200 import ParcopPlan
    (x1, xn), (y1, yn), (z1, zn) = domain(n, periodic)
    o = oracle_mod.Oracle(*n, x1, xn, y1, yn, z1, zn, periodic=(periodic,) * 3)
    p = ParcopPlan(*n, x1, xn, y1, yn, z1, zn, periodic=(periodic,) * 3, lib=lib)
    p.set_mesh()
    return o, p, synthetic_field(o.getvar("x"), o.getvar("y"), o.getvar("z"))


@pytest.mark.parametrize("periodic", [True, False])
def test_emulated_operators_match_oracle(periodic, oracle_mod, emul_lib):
    o, p, f = _pair((32, 48, 32), periodic, oracle_mod, emul_lib)
    for name in ("ddx", "ddy", "ddz", "dd8x", "dd8z", "d2y", "sfilter", "gfilter", "plaplacian", "pring", "dd4x", "dd4y", "dd4z"):
        assert rel_linf(getattr(p, name)(f), getattr(o, name)(f)) < 1e-13, (name, periodic)
    assert rel_linf(p.divergence(f, 2 * f, f * f), o.divergence(f, 2 * f, f * f)) < 1e-13
    for a, b in zip(p.grads(f), o.grads(f)):
        assert rel_linf(a, b) < 1e-13


@pytest.mark.parametrize("periodic", [True, False])
def test_emulated_host_pipeline_streams_the_gaussian_filter(periodic, oracle_mod, emul_lib):
    """pb_host_apply moves host arrays in slabs; the explicit filter's z sweep runs slab by slab behind the
    x / y sweeps (neighbouring slabs' planes as halo planes, wrap planes / closure rows at the ends)."""
    o, p, f = _pair((16, 32, 64), periodic, oracle_mod, emul_lib)
    for name in ("gfilter", "sfilter", "ddx", "ddy", "ddz"):
        assert rel_linf(getattr(p, name)(f), getattr(o, name)(f)) < 1e-13, (name, periodic)
    for d in (1, 2, 3):
        assert rel_linf(p.gfilterdir(f, d), o.dir_op("gf", d - 1, f)) < 1e-13, (d, periodic)


@pytest.mark.parametrize("periodic", [True, False])
def test_emulated_tensor_divergence_and_vector_ring(periodic, oracle_mod, emul_lib):
    """divergenceTensor and pRingV (parcop.f90:213-223,324-333), Cartesian."""
    o, p, f = _pair((32, 24, 20), periodic, oracle_mod, emul_lib)
    g = np.asarray(np.cos(2 * f) + 0.3 * f, order="F")
    h = np.asarray(f * f - 0.5, order="F")
    ins = (f, g, h, 2 * g, f + h, -f, h * g, 0.5 * f, g - h)
    for a, b in zip(p.divergencetensor(*ins), o.divergencetensor(*ins)):
        assert rel_linf(a, b) < 1e-13
    assert rel_linf(p.pringv(f, g, h), o.pringv(f, g, h)) < 1e-13


SYMM_CASES = [((True, True), (True, False), (False, True)), ((True, False), (False, False), (True, True))]


@pytest.mark.parametrize("symm", SYMM_CASES)
def test_emulated_symmetry_planes(symm, oracle_mod, emul_lib):
    """SYMM ends (compact.f90:77-91, stencils.f90:2390-2453): every one-field operator with the even
    closure rows, the divergences with the odd first derivative on the normal flux."""
    from pyranda_b200 import ParcopPlan
    n = (32, 48, 32)
    (x1, xn), (y1, yn), (z1, zn) = domain(n, False)
    o = oracle_mod.Oracle(*n, x1, xn, y1, yn, z1, zn, periodic=(False,) * 3, symmetric=symm)
    p = ParcopPlan(*n, x1, xn, y1, yn, z1, zn, periodic=(False,) * 3, symmetric=symm, lib=emul_lib)
    p.set_mesh()
    f = synthetic_field(o.getvar("x"), o.getvar("y"), o.getvar("z"))
    for name in ("ddx", "ddy", "ddz", "dd8x", "dd8y", "dd8z", "d2x", "d2y", "d2z", "sfilter", "gfilter", "plaplacian", "pring",
                 "dd4x", "dd4y", "dd4z"):
        assert rel_linf(getattr(p, name)(f), getattr(o, name)(f)) < 1e-13, name
    for d, name in enumerate(("ddx_odd", "ddy_odd", "ddz_odd")):
        assert rel_linf(getattr(p, name)(f), o.dir_op("d1", d, f, bc=-1)) < 1e-13, name
    for d, name in enumerate(("dd8x_odd", "dd8y_odd", "dd8z_odd")):
        assert rel_linf(getattr(p, name)(f), o.dir_op("d8", d, f, bc=-1)) < 1e-13, name
    g = np.asarray(np.cos(2 * f) + 0.3 * f, order="F")
    h = np.asarray(f * f - 0.5, order="F")
    assert rel_linf(p.divergence(f, g, h), o.divergence(f, g, h)) < 1e-13
    ins = (f, g, h, 2 * g, f + h, -f, h * g, 0.5 * f, g - h)
    for a, b in zip(p.divergencetensor(*ins), o.divergencetensor(*ins)):
        assert rel_linf(a, b) < 1e-13
    assert rel_linf(p.pringv(f, g, h), o.pringv(f, g, h)) < 1e-13  # the normal component through the odd d8


def test_emulated_symmetry_equals_mirrored_periodic_line(oracle_mod, emul_lib):
    """A symmetry plane is the interior stencil on the mirrored field: cos(kx) / sin(kx) on N cells
    of [0, pi] with SYMM ends give what the periodic operators give on 2N cells of [0, 2 pi]."""
    from pyranda_b200 import ParcopPlan
    N, ny, nz = 32, 16, 16
    dx = np.pi / N
    per = ParcopPlan(2 * N, ny, nz, dx / 2, 2 * np.pi - dx / 2, 0, 1, 0, 1, periodic=(True, False, False), lib=emul_lib)
    sym = ParcopPlan(N, ny, nz, dx / 2, np.pi - dx / 2, 0, 1, 0, 1, periodic=(False,) * 3,
                     symmetric=((True, True), (False, False), (False, False)), lib=emul_lib)
    per.set_mesh(); sym.set_mesh()
    X = per.getvar("x")
    rng = np.random.default_rng(7)
    amp = rng.uniform(-1, 1, size=6)
    even = np.asfortranarray(sum(a * np.cos(k * X) for k, a in enumerate(amp)))
    odd = np.asfortranarray(sum(a * np.sin((k + 1) * X) for k, a in enumerate(amp)))
    # (the 8th derivative of these smooth modes cancels seven digits: weights ~4e3 on values ~1 give ~2e-3)
    for name, tol in (("ddx", 1e-12), ("d2x", 1e-11), ("dd8x", 1e-8), ("sfilter", 1e-12), ("gfilter", 1e-12)):
        assert rel_linf(getattr(sym, name)(np.asfortranarray(even[:N])), getattr(per, name)(even)[:N]) < tol, name
    assert rel_linf(sym.ddx_odd(np.asfortranarray(odd[:N])), per.ddx(odd)[:N]) < 1e-12


def test_emulated_partial_tiles_and_single_chunk(oracle_mod, emul_lib):
    """nx not a multiple of the tile width, line count not a multiple of the x tile, P == 1."""
    emul_lib.pb_set_tuning(16, 16, 64)
    try:
        for periodic in (True, False):
            o, p, f = _pair((24, 20, 18), periodic, oracle_mod, emul_lib)
            for name in ("ddx", "ddy", "ddz", "sfilter"):
                assert rel_linf(getattr(p, name)(f), getattr(o, name)(f)) < 1e-13, (name, periodic)
    finally:
        emul_lib.pb_set_tuning(16, 16, 16)


def test_emulated_curvilinear(oracle_mod, emul_lib):
    """coordsys = 3: metrics, div, grad and the cell-volume weighted filter on a distorted grid."""
    from pyranda_b200 import ParcopPlan
    n = (32, 24, 16)
    xs = [np.linspace(0, 1, k) for k in n]
    X, Y, Z = np.meshgrid(*xs, indexing="ij")
    Xd = X + 0.05 * np.sin(2 * np.pi * Y) * np.sin(np.pi * X)
    Yd = Y + 0.04 * np.sin(2 * np.pi * X) * Z
    Zd = Z * (1 + 0.1 * X)
    o = oracle_mod.Oracle(*n, 0, 1, 0, 1, 0, 1, coordsys=3, mesh_xyz=(Xd, Yd, Zd))
    p = ParcopPlan(*n, 0, 1, 0, 1, 0, 1, coordsys=3, lib=emul_lib)
    p.set_mesh(Xd, Yd, Zd)
    for name in ("dtJ", "dAx", "dBy", "dCz", "dAy", "d1", "d2", "d3", "CellVol", "GridLen"):
        assert rel_linf(p.getvar(name), o.getvar(name)) < 1e-12, name
    f = synthetic_field(X, Y, Z)
    assert rel_linf(p.divergence(f, 2 * f, -f), o.divergence(f, 2 * f, -f)) < 1e-12
    for a, b in zip(p.grads(f), o.grads(f)):
        assert rel_linf(a, b) < 1e-12
    assert rel_linf(p.sfilter(f), o.sfilter(f)) < 1e-12
    assert rel_linf(p.pring(f), o.pring(f)) < 1e-12
    # divT: divergence of each row (operators.f90:176-179); ringV with the per-point d1, d2, d3 (:680-683)
    g = np.asarray(np.cos(2 * f) + 0.3 * f, order="F")
    h = np.asarray(f * f - 0.5, order="F")
    ins = (f, g, h, 2 * g, f + h, -f, h * g, 0.5 * f, g - h)
    for a, b in zip(p.divergencetensor(*ins), o.divergencetensor(*ins)):
        assert rel_linf(a, b) < 1e-12
    assert rel_linf(p.pringv(f, g, h), o.pringv(f, g, h)) < 1e-12


_SLOW = pytest.mark.skipif(not os.environ.get("PB_SLOW_TESTS"), reason="minutes under emulation; the same deck runs in the gloo and the GPU tests (PB_SLOW_TESTS=1 to include)")


@pytest.mark.parametrize("case", ["RT_2D", pytest.param("RT_3D", marks=_SLOW), "cylinder_curv", "cylinder_omesh", "symm_box"])
def test_emulated_example_decks_on_the_device_backend(case, oracle_mod, emul_lib):
    """BASELINE configs 4 and 5 as 2-D decks (and the O-grid deck with bc.slip) through the device
    backend of the interpreter, compute in the emulated build of the CUDA sources, against the
    oracle-backed driver whose full runs reproduce the reference's golden curves
    (tests/test_sim_oracle.py)."""
    from deck_parity import worst_difference
    n, steps = (16, 2) if case == "RT_3D" else (32, 5)  # the 3-D deck is slow under one-thread-per-CUDA-thread emulation
    assert worst_difference(case, n, oracle_mod, nsteps=steps, lib=emul_lib, tensor_device="cpu") < 1e-11


def test_copied_variable_keeps_its_value_through_the_inplace_stage(oracle_mod, emul_lib):
    """`:phi0: = :phi:` in the initial conditions (examples/simple.py, grid_conv.py, swirl.py): the
    reference rebinds the conserved array every stage, so phi0 keeps the initial field.  The device
    stage update is in place and must not drag the copy (or the mesh, `:phi: = meshx`) along; a
    C-contiguous conserved field must not be walked with the wrong strides either."""
    from deck_parity import device_sim
    from oracle_backend import make_sim
    mesh = "xdom = (0.0, 1.0, 32, periodic=True)\nydom = (0.0, 1.0, 16, periodic=True)\nzdom = (0.0, 1.0, 1)"
    eom = "ddt(:phi:) = - :c: * ddx(:phi:) - 0.5 * ddy(:phi:)\nddt(:psi:) = - ddx(:psi:)"
    ic = ":phi: = exp(-(meshx-0.5)**2/0.01)*cos(6.283185307179586*meshy)\n:phi0: = :phi:\n:c: = 1.0\n:psi: = meshx"
    sims = [make_sim(oracle_mod, "copy", mesh), device_sim("copy", mesh, lib=emul_lib, tensor_device="cpu")]
    for ss in sims:
        ss.EOM(eom)
        ss.setIC(ic)
    # a conserved field handed over C-contiguous (as a user function would return it)
    sims[1].variables["phi"] = sims[1].variables["phi"].contiguous()
    t = [0.0, 0.0]
    for _ in range(4):
        for k, ss in enumerate(sims):
            t[k] = ss.rk4(t[k], 2e-3)
    a, b = sims
    for nm in ("phi", "phi0", "psi", "meshx"):
        assert rel_linf(b.variables[nm].numpy(), a.variables[nm]) < 1e-13, nm
    assert np.abs(a.variables["phi"] - a.variables["phi0"]).max() > 1e-3          # the field moved ...
    assert np.abs(b.variables["phi0"].numpy() - a.variables["phi0"]).max() == 0.0  # ... the copy did not
