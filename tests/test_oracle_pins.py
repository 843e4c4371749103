"""Pin the CPU oracle (oracle/parcop_oracle.c) before anything is compared with it.

1. the reference's own golden scalar for its operator unit test (tests/cases/testUnit.py:3,
   examples/unit_test.py) -- periodic and bounded ddx/ddy/ddz, fbar, gbar at 32^3;
2. analytic transfer functions of the stencils on periodic grids;
3. emulated MPI ranks (the SPIKE reduced system) == one rank;
4. the committed golden vectors under tests/golden/.
The Fortran/MPI reference cannot be built in this image, so (1) at the reference's tolerance
(1e-4; we get ~1e-10) is the tightest pin against the reference itself.
"""
import os

import numpy as np
import pytest

from conftest import domain, rel_linf, synthetic_field

GOLDEN_OPS = ("ddx", "ddy", "ddz", "sfilter", "gfilter", "pring", "plaplacian")
GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def unit_test_error_sum(make_ops):
    """Restatement of /root/reference/examples/unit_test.py with `make_ops(periodic)` supplying the
    operator object; returns the ERROR scalar the reference prints."""
    Nx = 32
    L = float(str(np.pi * 2.0 * (Nx - 1) / Nx))  # the deck round-trips L through str()
    per, xyz_p = make_ops(Nx, L, True)
    bnd, xyz_b = make_ops(Nx, 1.0, False)
    npts = Nx ** 3
    err = 0.0
    dd = lambda o: (o.ddx, o.ddy, o.ddz)
    for d in range(3):
        err += np.sum(np.abs(dd(per)[d](np.sin(xyz_p[d])) - np.cos(xyz_p[d])))
    for d in range(3):
        err += np.sum(np.abs(dd(bnd)[d](xyz_b[d] * 0.0))) / npts
    for d in range(3):
        err += np.sum(np.abs(dd(bnd)[d](xyz_b[d] * 0.0 + 1.0))) / npts
    for power in (1.0, 2.0, 3.0, 4.0):
        for d in range(3):
            err += np.sum(np.abs(dd(bnd)[d](xyz_b[d] ** power) - power * xyz_b[d] ** (power - 1))) / npts
    mx, px = xyz_b[0], xyz_p[0]
    for fb, fp in ((bnd.sfilter, per.sfilter), (bnd.gfilter, per.gfilter)):
        err += np.sum(np.abs(fb(mx * 0.0))) / npts
        err += np.sum(np.abs(fb(mx * 0.0 + 1.0) - (mx * 0.0 + 1.0))) / npts
        err += np.sum(np.abs(fb(mx) - mx)) / npts
        err += np.sum(np.abs(fp(np.sin(px)) - np.sin(px))) / npts
        err += np.sum(np.abs(fp(px * 0.0 + 1.0) - 1.0)) / npts
        err += np.sum(np.abs(fp(px * 0.0))) / npts
    return err


REFERENCE_UNIT_TEST = 0.017490746954566275  # /root/reference/tests/cases/testUnit.py:3


def test_reference_unit_test_golden(oracle_mod):
    def make(Nx, L, periodic):
        o = oracle_mod.Oracle(Nx, Nx, Nx, 0.0, L, 0.0, L, 0.0, L, periodic=(periodic,) * 3)
        return o, [o.getvar(c) for c in "xyz"]
    err = unit_test_error_sum(make)
    # the reference accepts 1e-4 (tests/run_tests.py:84); the restatement lands within 1e-9
    assert abs(err - REFERENCE_UNIT_TEST) / REFERENCE_UNIT_TEST < 1e-9, err


@pytest.mark.parametrize("n", [32, 64, 100])
def test_transfer_functions_periodic(n, oracle_mod):
    (x1, xn), _, _ = domain((n, n, n), True)
    o = oracle_mod.Oracle(n, 16, 16, x1, xn, 0, 1, 0, 1, periodic=(True, False, False))
    x = o.getvar("x")
    h = o.dx
    w1, w2, w8, wf, wg = (oracle_mod.get_weight(k) for k in ("d1", "d2", "d8", "sf", "gf"))
    for k in (1.0, 3.0, 7.0):
        th = k * h
        s, c = np.sin(k * x), np.cos(k * x)
        lhs = lambda w: w["ali"][2] + 2 * w["ali"][3] * np.cos(th) + 2 * w["ali"][4] * np.cos(2 * th)
        kp = 2 * sum(w1["ari"][3 + j] * np.sin(j * th) for j in (1, 2, 3)) / lhs(w1) / h
        assert np.abs(o.ddx(s) - kp * c).max() < 2e-13 * max(1, k)
        k2 = (w2["ari"][3] + 2 * sum(w2["ari"][3 + j] * np.cos(j * th) for j in (1, 2, 3))) / lhs(w2) / h ** 2
        assert np.abs(o.d2x(s) - k2 * s).max() < 1e-12 * abs(k2)
        k8 = (w8["ari"][4] + 2 * sum(w8["ari"][4 + j] * np.cos(j * th) for j in (1, 2, 3, 4))) / lhs(w8)
        # the d8 weights are O(4e3) and cancel to O(k^8 h^8): the error floor is ~4e3 * eps
        assert np.abs(o.dd8x(s) - k8 * s).max() < 2e-11 * max(abs(k8), 1)
        # filters are stored in difference form: T = 1 + (B - A)/A
        tf = 1 + (wf["ari"][4] + 2 * sum(wf["ari"][4 + j] * np.cos(j * th) for j in (1, 2, 3, 4))) / lhs(wf)
        assert np.abs(o.dir_op("sf", 0, s, bc=1) - tf * s).max() < 1e-13
        tg = 1 + (wg["ari"][4] + 2 * sum(wg["ari"][4 + j] * np.cos(j * th) for j in (1, 2, 3, 4)))
        assert np.abs(o.gfilterdir(s, 1) - tg * s).max() < 1e-14


def test_bounded_exactness(oracle_mod):
    n = 40
    o = oracle_mod.Oracle(n, n, n, 0, 1, 0, 1, 0, 1)
    x, y, z = (o.getvar(c) for c in "xyz")
    assert np.abs(o.ddx(x ** 3) - 3 * x ** 2).max() < 1e-11
    assert np.abs(o.ddy(y ** 2) - 2 * y).max() < 1e-12
    assert np.abs(o.d2z(z ** 3) - 6 * z).max() < 1e-9
    assert np.abs(o.plaplacian(x ** 2 + y ** 2 + z ** 2) - 6.0).max() < 1e-9
    assert np.abs(o.sfilter(1 + x + 2 * y) - (1 + x + 2 * y)).max() < 1e-13
    assert np.all(o.ddx(np.ones_like(x)) == 0.0)


@pytest.mark.parametrize("periodic", [True, False])
@pytest.mark.parametrize("nranks", [2, 4])
def test_emulated_ranks_equal_one_rank(periodic, nranks, oracle_mod):
    n = (64, 64, 64)
    (x1, xn), (y1, yn), (z1, zn) = domain(n, periodic)
    one = oracle_mod.Oracle(*n, x1, xn, y1, yn, z1, zn, periodic=(periodic,) * 3)
    many = oracle_mod.Oracle(*n, x1, xn, y1, yn, z1, zn, periodic=(periodic,) * 3, px=nranks, py=nranks, pz=nranks)
    f = synthetic_field(one.getvar("x"), one.getvar("y"), one.getvar("z"))
    for kind in ("d1", "d2", "d8", "sf", "gf"):
        for d in range(3):
            bc = 1 if kind in ("sf", "gf") else 0
            assert rel_linf(many.dir_op(kind, d, f, bc), one.dir_op(kind, d, f, bc)) < 1e-13, (kind, d)


def test_symmetric_weights_fold_the_interior_stencil(oracle_mod):
    """stencils.f90:2390-2453: bc=+1/-1 rows are the interior stencil reflected about the face."""
    for kind, sl, sr in (("d1", -1, +1), ("d2", +1, +1), ("sf", +1, +1)):
        w = oracle_mod.get_weight(kind)
        nor = w["nor"]
        ari = w["ari"].copy()
        if kind == "sf":  # undo the difference form for the comparison
            ari[2:7] += w["ali"]
        row0 = w["arb1"][2][0].copy()  # bc = +1, first row
        if kind == "sf":
            row0[2:7] += w["alb1"][2][0]
        expect = ari.copy()
        for j in range(nor):  # ghost point -1-j mirrors onto point j
            expect[nor + 1 + j] += sr * expect[nor - 1 - j] if j + 1 <= nor else 0
            expect[nor - 1 - j] = 0.0
        # row 0 sits on point 0: stencil slot nor-1-j (point -1-j) folds onto slot nor+j (point j)
        fold = ari.copy()
        for j in range(nor):
            fold[nor + j] += sr * ari[nor - 1 - j]
            fold[nor - 1 - j] = 0.0
        assert np.allclose(row0, fold, atol=0, rtol=0), kind


def test_curvilinear_metrics_reduce_to_cartesian(oracle_mod):
    """coordsys=3 on an undistorted grid must reproduce the Cartesian operators."""
    n = (32, 24, 20)
    xs = [np.linspace(0, 1, k) for k in n]
    X, Y, Z = np.meshgrid(*xs, indexing="ij")
    cart = oracle_mod.Oracle(*n, 0, 1, 0, 1, 0, 1)
    curv = oracle_mod.Oracle(*n, 0, 1, 0, 1, 0, 1, coordsys=3, mesh_xyz=(X, Y, Z))
    f = synthetic_field(X, Y, Z)
    assert np.abs(curv.getvar("dtJ") - 1.0).max() < 1e-11
    assert rel_linf(curv.divergence(f, 2 * f, -f), cart.divergence(f, 2 * f, -f)) < 1e-10
    for a, b in zip(curv.grads(f), cart.grads(f)):
        assert rel_linf(a, b) < 1e-10
    assert rel_linf(curv.sfilter(f), cart.sfilter(f)) < 1e-12


def test_golden_vectors(oracle_mod):
    """Committed fixtures (tests/golden/make_golden.py): the oracle must keep producing them bit-for-bit
    up to the libm / compiler differences between machines (1e-14)."""
    data = np.load(os.path.join(GOLDEN, "ops_16x16x16.npz"))
    for periodic in (True, False):
        tag = "per" if periodic else "bnd"
        n = tuple(int(v) for v in data["n"])
        (x1, xn), (y1, yn), (z1, zn) = domain(n, periodic)
        o = oracle_mod.Oracle(*n, x1, xn, y1, yn, z1, zn, periodic=(periodic,) * 3)
        f = np.asfortranarray(data["f_" + tag])
        for name in GOLDEN_OPS:
            assert rel_linf(getattr(o, name)(f), data["%s_%s" % (name, tag)]) < 1e-14, (name, tag)


def test_fourth_derivative_transfer_function(oracle_mod):
    """e4d4 (stencils.f90:430-513) on a periodic axis: the oracle's dd4 of sin(kx) is the 7-point
    stencil's transfer function times sin(kx), which tends to (k dx)^4 (no metric scale,
    compact_operators.f90:229)."""
    N = 48
    o = oracle_mod.Oracle(N, 8, 8, 0.0, 2 * np.pi * (N - 1) / N, 0, 1, 0, 1, periodic=(True, False, False))
    X = o.getvar("x")
    dx = 2 * np.pi / N
    a, b, c, d = 28.0 / 3.0, -6.5, 2.0, -1.0 / 6.0
    for k in (1, 2, 5, 11):
        T = a + 2 * b * np.cos(k * dx) + 2 * c * np.cos(2 * k * dx) + 2 * d * np.cos(3 * k * dx)
        f = np.asfortranarray(np.sin(k * X))
        assert np.abs(o.dd4x(f) - T * f).max() < 1e-13
        if k <= 2:
            assert abs(T - (k * dx) ** 4) < 0.3 * (k * dx) ** 8  # 4th-order accurate: error O(th^8)
