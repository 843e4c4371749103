"""The drop-in boundary as a pyranda maintainer would bind it: the f2py-shaped module
`pyranda_b200.parcop.parcop` and the `pyrandaMPI` class, driven with the reference's own call
sequence (pyranda/pyrandaMPI.py:151-174 setup, :246-247 setPatch, :657-661 getVar, :664-740 the
der / gfil / sfil dispatch objects; pyrandaMesh.py:93-135 setup_mesh) and checked against the oracle.

CPU: the host-emulated build of the same sources (tests/emul).  GPU: libparcop_b200.so."""
import os
import subprocess

import numpy as np
import pytest

from conftest import domain, rel_linf, synthetic_field

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
EMUL = os.path.join(ROOT, "tests", "emul")
TOL = 1e-12


def _emul():
    subprocess.check_call(["make", "-C", EMUL, "-s"])
    from pyranda_b200 import _lib
    L = _lib.load(os.path.join(EMUL, "libparcop_emul.so"))
    L.pb_set_tuning(16, 16, 16)
    return {"lib": L, "tensor_device": "cpu"}


def _f2py_sequence(oracle_mod, kw, n=(32, 24, 20), periodic=False):
    from pyranda_b200.parcop import parcop
    (x1, xn), (y1, yn), (z1, zn) = domain(n, periodic)
    b = "PERI" if periodic else "NONE"
    patch, level = 3, 1
    # pyrandaMPI.py:151-155
    parcop.setup(patch, level, 0, n[0], n[1], n[2], 1, 1, 1, 0, x1, xn, y1, yn, z1, zn, b, b, b, b, b, b, **kw)
    parcop.set_patch(patch, level)                       # :246-247
    parcop.setup_mesh(patch, level)                      # pyrandaMesh.py:133
    o = oracle_mod.Oracle(*n, x1, xn, y1, yn, z1, zn, periodic=(periodic,) * 3)
    for nm in ("x", "y", "z", "d1", "d2", "d3", "CellVol", "GridLen"):
        got = parcop.getvar(nm, *n)                      # :657-661: name and the three sizes
        assert got.shape == tuple(n) and got.flags.f_contiguous
        assert rel_linf(got, o.getvar(nm)) < 1e-14, nm
    assert rel_linf(parcop.xgrid(*n), o.getvar("x")) < 1e-14
    f = synthetic_field(o.getvar("x"), o.getvar("y"), o.getvar("z"))
    g = np.asfortranarray(np.cos(2 * f) + 0.3 * f)
    h = np.asfortranarray(f * f - 0.5)
    one = ("ddx", "ddy", "ddz", "dd4x", "dd4y", "dd4z", "dd8x", "dd8y", "dd8z", "plaplacian", "pring", "sfilter", "gfilter")
    for nm in one:                                       # parcop_der / parcop_sfil / parcop_gfil bodies
        got = getattr(parcop, nm)(f)
        assert isinstance(got, np.ndarray) and got.shape == tuple(n)
        assert rel_linf(got, getattr(o, nm)(f)) < TOL, nm
    for d in (1, 2, 3):
        assert rel_linf(parcop.gfilterdir(f, d), o.gfilterdir(f, d)) < TOL
    assert rel_linf(parcop.divergence(f, g, h), o.divergence(f, g, h)) < TOL
    gr = parcop.grads(f)
    assert isinstance(gr, tuple) and len(gr) == 3       # f2py returns the three intent(out) arrays as a tuple
    for a, bb in zip(gr, o.grads(f)):
        assert rel_linf(a, bb) < TOL
    ins = (f, g, h, 2 * g, f + h, -f, h * g, 0.5 * f, g - h)
    dv = parcop.divergencetensor(*ins)
    assert isinstance(dv, tuple) and len(dv) == 3
    for a, bb in zip(dv, o.divergencetensor(*ins)):
        assert rel_linf(a, bb) < TOL
    assert rel_linf(parcop.pringv(f, g, h), o.pringv(f, g, h)) < TOL
    # a second patch, then back: the module-global current patch (parcop.f90:196-200)
    parcop.setup(4, 1, 0, 16, 16, 16, 1, 1, 1, 0, 0.0, 1.0, 0.0, 1.0, 0.0, 1.0, b, b, b, b, b, b, **kw)
    parcop.setup_mesh(4, 1)
    assert parcop.getvar("x").shape == (16, 16, 16)
    parcop.set_patch(patch, level)
    assert rel_linf(parcop.ddx(f), o.ddx(f)) < TOL
    with pytest.raises(Exception):
        parcop.set_patch(9, 9)


def _mpi_class_sequence(oracle_mod, kw, n=(32, 24, 20), periodic=True, device_fields=False):
    from pyranda_b200.pyrandaMPI import pyrandaMPI
    (x1, xn), (y1, yn), (z1, zn) = domain(n, periodic)
    mesh = {"nn": list(n), "x1": [x1, y1, z1], "xn": [xn, yn, zn], "periodic": [periodic] * 3, "coordsys": 0}
    pm = pyrandaMPI(mesh, **kw)
    pm.setPatch()
    o = oracle_mod.Oracle(*n, x1, xn, y1, yn, z1, zn, periodic=(periodic,) * 3)
    assert (pm.ax, pm.ay, pm.az) == tuple(n) and tuple(pm.chunk_3d_size) == tuple(n)
    assert pm.x1proc and pm.znproc and pm.master
    assert rel_linf(pm.getVar("GridLen"), o.getvar("GridLen")) < 1e-14
    f = synthetic_field(o.getvar("x"), o.getvar("y"), o.getvar("z"))
    g = np.asfortranarray(np.cos(2 * f) + 0.3 * f)
    h = np.asfortranarray(f * f - 0.5)
    if device_fields:
        import torch

        def dev(a):
            t = pm.emptyScalar()
            t.copy_(torch.as_tensor(np.ascontiguousarray(a), device=t.device))
            return t
        back = lambda t: t.cpu().numpy()
    else:
        dev = back = lambda a: a
    F, G, H = dev(f), dev(g), dev(h)
    for nm, ref in (("ddx", "ddx"), ("ddy", "ddy"), ("ddz", "ddz"), ("dd4x", "dd4x"), ("dd8y", "dd8y"), ("dd8z", "dd8z"),
                    ("laplacian", "plaplacian"), ("ring", "pring")):
        assert rel_linf(back(getattr(pm.der, nm)(F)), getattr(o, ref)(f)) < TOL, nm
    assert rel_linf(back(pm.der.div(F, G, H)), o.divergence(f, g, h)) < TOL
    for a, bb in zip(pm.der.grad(F), o.grads(f)):
        assert rel_linf(back(a), bb) < TOL
    assert rel_linf(back(pm.der.ringV(F, G, H)), o.pringv(f, g, h)) < TOL
    assert rel_linf(back(pm.fil.filter(F)), o.sfilter(f)) < TOL
    assert rel_linf(back(pm.gfil.filter(F)), o.gfilter(f)) < TOL
    for d in (1, 2, 3):
        assert rel_linf(back(pm.gfil.filterDir(F, d)), o.gfilterdir(f, d)) < TOL
    if device_fields:
        assert abs(pm.sum3D(F) - f.sum()) <= 1e-12 * np.abs(f).sum()
        assert pm.max3D(F) == f.max() and pm.min3D(F) == f.min()
        assert float(pm.emptyScalar().abs().max()) == 0.0


@pytest.mark.parametrize("periodic", [True, False])
def test_f2py_module_call_sequence_emulated(periodic, oracle_mod):
    _f2py_sequence(oracle_mod, _emul(), periodic=periodic)


def test_pyrandaMPI_class_emulated(oracle_mod):
    _mpi_class_sequence(oracle_mod, _emul())


@pytest.mark.gpu
@pytest.mark.parametrize("periodic", [True, False])
def test_f2py_module_call_sequence_gpu(periodic, oracle_mod):
    _f2py_sequence(oracle_mod, {}, n=(64, 48, 32), periodic=periodic)


@pytest.mark.gpu
@pytest.mark.parametrize("device_fields", [False, True])
def test_pyrandaMPI_class_gpu(device_fields, oracle_mod):
    _mpi_class_sequence(oracle_mod, {}, n=(64, 48, 32), device_fields=device_fields)
