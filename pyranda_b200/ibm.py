"""Immersed-boundary package of the EOM interpreter: ibmS / ibmV / ibmWall.

Reference: pyranda/pyrandaIBM.py:14-16 (constants), :34-73 (deck entry points), :75-89 (the
extension march `smooth_terrain`), :92-194 (`slip_velocity`).  A level set `phi` (negative inside the
body) and its gradient are given; fields are extended into the body by a few pseudo-time steps of
advection along the gradient, each followed by a Gaussian filter, and velocities get their normal (slip
wall) or whole (no-slip wall) component removed and rebuilt as a linear profile through the zero
level.  Everything is `grad`, `gfilter` and masked pointwise algebra on the interpreter's array
namespace, so with the CUDA backend it runs on the device-resident fields.
"""


class ImmersedBoundary:
    ITER, CFL, EPS = 2, 0.5, 0.1  # pyrandaIBM.py:14-16

    def __init__(self, sim):
        self.sim = sim

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

    def _extend(self, sdf, g, val, epsi):
        """pyrandaIBM.py:75-89: march `val` along grad(phi) where phi <= epsi, filtering each step."""
        sim, xp = self.sim, self.sim.xp
        val = val * 1.0
        inside = sdf <= epsi
        for _ in range(self.ITER):
            tx, ty, tz = sim.grad(val)
            term = tx * g[0] + ty * g[1] + tz * g[2]
            val = xp.where(inside, val + self.CFL * sim.GridLen * term, val)
            val = xp.where(inside, sim.gfilter(val), val)
        return val

    def _wall(self, sdf, g, vel, frame, slip):
        """pyrandaIBM.py:92-194."""
        sim, xp = self.sim, self.sim.xp
        lens = sim.GridLen * self.EPS
        v = [self._field(c) * 1.0 for c in vel]
        if frame:  # interface velocity: work in its frame (:100-108)
            v = [c - f for c, f in zip(v, frame)]
        v = [self._extend(sdf, g, c, 0.0) for c in v]
        near = sdf < lens
        if slip:
            normal = v[0] * g[0] + v[1] * g[1] + v[2] * g[2]
            axis = g
        else:
            normal = xp.sqrt(v[0] * v[0] + v[1] * v[1] + v[2] * v[2])
            inv = 1.0 / (normal + 1.0e-16)
            axis = [c * inv for c in v]
        vn = xp.where(near, normal, 0.0)
        ramp = xp.where(near, 0.0, normal / sdf)
        v = [c - vn * a for c, a in zip(v, axis)]
        ramp = self._extend(sdf, g, ramp, lens)  # linear profile through the zero level (:131-133,178-180)
        vn = xp.where(near, ramp * sdf, 0.0)
        v = [c + vn * a for c, a in zip(v, axis)]
        if frame:
            v = [c + f for c, f in zip(v, frame)]
        return v
