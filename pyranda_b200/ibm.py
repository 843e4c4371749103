"""Immersed-boundary package of the EOM interpreter: ibmS / ibmV / ibmWall.

Reference: pyranda/pyrandaIBM.py:14-16 (constants), :34-73 (deck entry points), :75-89 (the
extension march `smooth_terrain`), :92-194 (`slip_velocity`).  A level set `phi` (negative inside the
body) and its gradient are given; fields are extended into the body by a few pseudo-time steps of
advection along the gradient, each followed by a Gaussian filter, and velocities get their normal (slip
wall) or whole (no-slip wall) component removed and rebuilt as a linear profile through the zero
level.  Everything is `grad`, `gfilter` and masked pointwise algebra on the interpreter's array
namespace, so with the CUDA backend it runs on the device-resident fields; each pointwise expression is
handed to the interpreter's fuser as a whole (one generated kernel instead of one pass per operation).
"""


class ImmersedBoundary:
    ITER, CFL, EPS = 2, 0.5, 0.1  # pyrandaIBM.py:14-16

    def __init__(self, sim):
        self.sim = sim
        self._code = {}

    # ---- deck entry points (pyrandaIBM.py:34-73) ----
    def scalar(self, value, phi, gphi):
        return self._extend(phi, gphi, self._field(value), 0.0)

    def velocity_slip(self, vel, phi, gphi, phivar=None):
        return self._wall(phi, gphi, vel, phivar, slip=True)

    def velocity_wall(self, vel, phi, gphi, phivar=None):
        return self._wall(phi, gphi, vel, None, slip=False)  # the reference drops phivar here (:60-61)

    # ---- helpers ----
    def _field(self, a):
        return a if self.sim.B.isfield(a) else self.sim.emptyScalar(float(a))

    def _ev(self, src, **local):
        """Pointwise algebra as ONE expression: with the device backend the interpreter's fuser turns it into one
        generated kernel (same IEEE operations in the same order as the array evaluation, fuse.py); otherwise it
        is evaluated as written on the array namespace."""
        code = self._code.get(src)
        if code is None:
            fz = getattr(self.sim, "fuser", None)
            code = self._code[src] = compile(fz.transform(src) if fz is not None else src, "<ibm>", "eval")
        ns = getattr(self.sim, "_ns", None)
        return eval(code, ns if ns is not None else {"xp": self.sim.xp}, local)

    def _extend(self, sdf, g, val, epsi):
        """pyrandaIBM.py:75-89: march `val` along grad(phi) where phi <= epsi, filtering each step."""
        sim = self.sim
        val = val * 1.0
        for _ in range(self.ITER):
            tx, ty, tz = sim.grad(val)
            val = self._ev("xp.where(sdf <= epsi, val + cfl * gl * (tx * g0 + ty * g1 + tz * g2), val)", sdf=sdf, epsi=epsi, val=val,
                           cfl=self.CFL, gl=sim.GridLen, tx=tx, ty=ty, tz=tz, g0=g[0], g1=g[1], g2=g[2])
            val = self._ev("xp.where(sdf <= epsi, filt, val)", sdf=sdf, epsi=epsi, filt=sim.gfilter(val), val=val)
        return val

    def _wall(self, sdf, g, vel, frame, slip):
        """pyrandaIBM.py:92-194."""
        sim = self.sim
        lens = sim.GridLen * self.EPS
        v = [self._field(c) * 1.0 for c in vel]
        if frame:  # interface velocity: work in its frame (:100-108)
            v = [c - f for c, f in zip(v, frame)]
        v = [self._extend(sdf, g, c, 0.0) for c in v]
        kw = dict(sdf=sdf, lens=lens, v0=v[0], v1=v[1], v2=v[2])
        if slip:
            normal = self._ev("v0 * g0 + v1 * g1 + v2 * g2", g0=g[0], g1=g[1], g2=g[2], **kw)
            axis = g
        else:
            normal = self._ev("xp.sqrt(v0 * v0 + v1 * v1 + v2 * v2)", **kw)
            axis = [self._ev("c * (1.0 / (normal + 1.0e-16))", c=c, normal=normal) for c in v]
        ramp = self._ev("xp.where(sdf < lens, 0.0, normal / sdf)", normal=normal, sdf=sdf, lens=lens)
        v = [self._ev("c - xp.where(sdf < lens, normal, 0.0) * a", c=c, a=a, normal=normal, sdf=sdf, lens=lens) for c, a in zip(v, axis)]
        ramp = self._extend(sdf, g, ramp, lens)  # linear profile through the zero level (:131-133,178-180)
        v = [self._ev("c + xp.where(sdf < lens, ramp * sdf, 0.0) * a", c=c, a=a, ramp=ramp, sdf=sdf, lens=lens) for c, a in zip(v, axis)]
        if frame:
            v = [c + f for c, f in zip(v, frame)]
        return v
