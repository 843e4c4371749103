"""numpy + CPU-oracle backend for pyranda_b200.sim.pyrandaSim (TEST INFRASTRUCTURE).

Lets the very same deck driver that runs on the GPU run on the host with the oracle's operators,
so multi-step results can be compared (CUDA vs oracle) and the reference's golden scalars for whole
simulations (Taylor-Green, 1-D advection) can pin the oracle."""
import numpy as np


class NumpyOracleBackend:
    def __init__(self, oracle_obj):
        self.o = oracle_obj
        self.xp = np
        self.shape = oracle_obj.shape

    def zeros(self): return np.zeros(self.shape, order="F")
    def asfield(self, a): return np.asfortranarray(a, dtype=np.float64)
    def tohost(self, a): return a
    def isfield(self, a): return isinstance(a, np.ndarray)
    def _f(self, a): return a if isinstance(a, np.ndarray) else np.full(self.shape, float(a), order="F")
    def ddx(self, v): return self.o.ddx(self._f(v))
    def ddy(self, v): return self.o.ddy(self._f(v))
    def ddz(self, v): return self.o.ddz(self._f(v))
    def dd4x(self, v): return self.o.dd4x(self._f(v))
    def dd4y(self, v): return self.o.dd4y(self._f(v))
    def dd4z(self, v): return self.o.dd4z(self._f(v))
    def dd8x(self, v): return self.o.dd8x(self._f(v))
    def dd8y(self, v): return self.o.dd8y(self._f(v))
    def dd8z(self, v): return self.o.dd8z(self._f(v))
    def filter(self, v): return self.o.sfilter(self._f(v))
    def gfilter(self, v): return self.o.gfilter(self._f(v))
    def gfilterdir(self, v, d): return self.o.gfilterdir(self._f(v), d)
    def ring(self, v): return self.o.pring(self._f(v))
    def laplacian(self, v): return self.o.plaplacian(self._f(v))
    def div(self, a, b, c): return self.o.divergence(self._f(a), self._f(b), self._f(c))
    def grad(self, v): return self.o.grads(self._f(v))
    def divT(self, *f9): return self.o.divergencetensor(*[self._f(a) for a in f9])
    def ringV(self, a, b, c): return self.o.pringv(self._f(a), self._f(b), self._f(c))
    def getvar(self, name): return self.o.getvar(name)
    def sum3D(self, a): return float(np.sum(a))
    def max3D(self, a): return float(np.max(a))
    def min3D(self, a): return float(np.min(a))

    def rk4_stage(self, dt, A, B, F, PHI, U):  # pyranda.py:800-804
        tmp1 = A * PHI
        PHI[...] = dt * F + tmp1
        tmp2 = B * PHI
        return U + tmp2


def make_sim(oracle_mod, name, mesh):
    from pyranda_b200.sim import curvilinear_coordinates, parse_mesh, pyrandaSim
    opt = parse_mesh(mesh) if isinstance(mesh, str) else mesh
    kw = {}
    if int(opt.get("coordsys", 0)) == 3:
        kw = {"coordsys": 3, "mesh_xyz": curvilinear_coordinates(opt), "periodic_grid": bool(opt.get("periodicGrid", True))}
    if "symmetric" in opt:
        kw["symmetric"] = tuple(tuple(bool(b) for b in pair) for pair in opt["symmetric"])
    o = oracle_mod.Oracle(*opt["nn"], opt["x1"][0], opt["xn"][0], opt["x1"][1], opt["xn"][1], opt["x1"][2], opt["xn"][2],
                          periodic=tuple(opt["periodic"]), **kw)
    return pyrandaSim(name, opt, backend=NumpyOracleBackend(o))
