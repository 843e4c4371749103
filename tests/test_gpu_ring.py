"""GPU: the ring kernel (thread-block clusters, 64-line tiles) and direct oracle parity at the
configurations the benchmark times (512^3 periodic and bounded, 1024-point lines in every direction).
Tolerance 1e-12 relative L-infinity per operator application (BASELINE.json north_star)."""
import numpy as np
import pytest

from conftest import domain, rel_linf, synthetic_field

pytestmark = pytest.mark.gpu
TOL = 1e-12


def _pair(n, periodic, oracle_mod):
    from pyranda_b200 import ParcopPlan
    (x1, xn), (y1, yn), (z1, zn) = domain(n, periodic)
    o = oracle_mod.Oracle(*n, x1, xn, y1, yn, z1, zn, periodic=(periodic,) * 3)
    p = ParcopPlan(*n, x1, xn, y1, yn, z1, zn, periodic=(periodic,) * 3)
    p.set_mesh()
    return o, p, synthetic_field(o.getvar("x"), o.getvar("y"), o.getvar("z"))


def _lib():
    from pyranda_b200 import _lib
    return _lib.load()


@pytest.mark.parametrize("periodic", [True, False])
@pytest.mark.parametrize("n,lines,names", [
    ((64, 1024, 16), 32, ("ddy", "d2y", "dd8y", "sfilter", "plaplacian", "pring")),   # 4-CTA clusters
    ((72, 16, 1024), 16, ("ddz", "d2z", "dd8z", "sfilter")),                          # 2-CTA clusters, partial tile
    ((1024, 64, 64), 0, ("ddx", "sfilter", "gfilter")),                               # 1024-point x lines
    ((160, 128, 16), 0, ("ddy", "sfilter", "plaplacian", "pring")),                   # 64-line tiles (128-point lines)
    ((96, 16, 128), 0, ("ddz", "sfilter")),
])
def test_ring_kernel_lines(n, lines, names, periodic, oracle_mod):
    L = _lib()
    L.pb_set_ring(1, lines)
    try:
        o, p, f = _pair(n, periodic, oracle_mod)
        for nm in names:
            r0 = L.pb_ring_launch_count()
            got = getattr(p, nm)(f)
            if n[0] != 1024:
                assert L.pb_ring_launch_count() > r0, "the ring kernel did not run"
            assert rel_linf(got, getattr(o, nm)(f)) < TOL, (nm, n, periodic)
    finally:
        L.pb_set_ring(1, 0)


@pytest.mark.parametrize("periodic", [True, False])
@pytest.mark.parametrize("lines", [16, 32])
def test_ring_kernel_wherever_it_fits(lines, periodic, oracle_mod):
    L = _lib()
    L.pb_set_ring(2, lines)
    try:
        for n, names in (((64, 512, 16), ("ddy", "sfilter", "pring")), ((64, 16, 256), ("ddz", "sfilter"))):
            o, p, f = _pair(n, periodic, oracle_mod)
            for nm in names:
                assert rel_linf(getattr(p, nm)(f), getattr(o, nm)(f)) < TOL, (nm, n, periodic, lines)
    finally:
        L.pb_set_ring(1, 0)


@pytest.mark.parametrize("periodic", [True, False])
def test_benchmark_operators_512_against_the_oracle(periodic, oracle_mod):
    """BASELINE configs[1] size, the five benchmark operators + ring + laplacian, device-resident
    path (the persistent kernels walk > 50 tiles per CTA here), compared point by point."""
    import torch
    n = (512, 512, 512)
    o, p, f = _pair(n, periodic, oracle_mod)
    t = p.empty_device()
    t.copy_(torch.from_numpy(f).cuda())
    for nm in ("ddx", "ddy", "ddz", "sfilter", "gfilter", "pring", "plaplacian"):
        got = getattr(p, nm)(t).cpu().numpy()
        ref = getattr(o, nm)(f)
        assert rel_linf(got, ref) < TOL, (nm, periodic, rel_linf(got, ref))
        del got, ref
