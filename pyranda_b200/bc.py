"""Boundary-plane packages of the EOM interpreter: bc.extrap / bc.const / bc.field / bc.symm /
bc.exit / bc.slip / bc.farfield.

Reference: pyranda/pyrandaBC.py:40-186,748-786 (the `BC` package: `bc.extrap(vars, dirs, order)`,
`bc.const(vars, dirs, val)`, `bc.field(var, dirs, field)`, `bc.symm(vars, dirs, anti, npts)` lines
inside an EOM string), :468-522 (`bc.exit(vars, dirs, norm)`, the bounded essentially-non-oscillatory
outflow extrapolation) and :186-466 (`bc.slip([[u, v(, w)]], dirs)`, free slip on a curvilinear
wall: extrapolate, then remove the wall-normal velocity) and :524-746 (`bc.farfield(dirs)`,
characteristic far-field values from the Riemann invariants; its reference state is stored by the
deck in `BCdata['farfield-properties-<dir>']`, as in examples/cylinder_O_grid.py:84).  They sit in
`updateVars`, i.e. they run after every RK4 stage, so with device-resident fields they must not
leave the GPU: everything here is in-place slicing on the field object (a CUDA tensor with Fortran
strides, or a numpy array in the oracle-backed test driver) -- one tiny strided kernel per plane.

A boundary is named by axis and side ('x1', 'xn', 'y1', 'yn', 'z1', 'zn'); `owns[name]` tells whether
this rank holds that physical boundary (pyrandaMPI x1proc ... znproc: always true on one GPU, the
first / last rank of a z-slab for 'z1' / 'zn').
"""

_AXIS = {"x": 0, "y": 1, "z": 2}


def _as_list(a):
    return list(a) if isinstance(a, (list, tuple)) else [a]


def _xp(f):
    """The array module of a field: torch for device tensors, numpy for the oracle-backed driver."""
    if type(f).__module__.startswith("torch"):
        import torch
        return torch
    import numpy
    return numpy


class BoundaryConditions:
    def __init__(self, variables, owns=None, getvar=None):
        self.variables = variables
        self.owns = owns if owns is not None else {n: True for n in ("x1", "xn", "y1", "yn", "z1", "zn")}
        self.getvar = getvar   # mesh metrics by name (pyranda.getVar), needed by bc.slip
        self.BCdata = {}       # wall normals, cached per boundary (pyrandaBC.py:24,281-291)

    @staticmethod
    def _plane(direction, depth):
        """Index tuple of the plane `depth` points inside the boundary `direction`."""
        axis = _AXIS[direction[0]]
        at = depth if direction[1] == "1" else -1 - depth
        idx = [slice(None)] * 3
        idx[axis] = at
        return tuple(idx)

    def _each(self, var, direction):
        for d in _as_list(direction):
            if d[0] not in _AXIS or d[1:] not in ("1", "n"):
                raise ValueError("unknown boundary '%s'" % d)
            if not self.owns.get(d, False):
                continue
            for v in _as_list(var):
                yield self.variables[v], d

    # pyrandaBC.py:58-116
    def extrap(self, var, direction, order=2):
        for f, d in self._each(var, direction):
            if order == 2:
                f[self._plane(d, 0)] = 2 * f[self._plane(d, 1)] - f[self._plane(d, 2)]
            else:
                f[self._plane(d, 0)] = f[self._plane(d, 1)]

    # pyrandaBC.py:118-160
    def const(self, var, direction, val):
        for f, d in self._each(var, direction):
            f[self._plane(d, 0)] = val

    # pyrandaBC.py:748-786: the first `npts` planes mirror the next `npts` (sign flipped for `anti`)
    def symm(self, var, direction, anti=False, npts=4):
        sign = -1.0 if anti else 1.0
        for f, d in self._each(var, direction):
            axis = _AXIS[d[0]]
            ghost, image = [slice(None)] * 3, [slice(None)] * 3
            if d[1] == "1":
                ghost[axis], image[axis] = slice(0, npts), slice(npts, 2 * npts)
            else:
                ghost[axis], image[axis] = slice(-npts, None), slice(-2 * npts, -npts)
            src = f[tuple(image)]
            src = src.flip(axis) if hasattr(src, "flip") else src[tuple(slice(None, None, -1) if k == axis else slice(None) for k in range(3))]
            f[tuple(ghost)] = src * sign

    # pyrandaBC.py:162-186
    def field(self, var, direction, field):
        for f, d in self._each(var, direction):
            f[self._plane(d, 0)] = field[self._plane(d, 0)]

    # pyrandaBC.py:468-522: x and y boundaries only (the reference has no z branch)
    def exit(self, var, direction, norm=False):
        for f, d in self._each(var, direction):
            if d[0] == "z":
                continue
            u1, u2, u3 = f[self._plane(d, 0)], f[self._plane(d, 1)], f[self._plane(d, 2)]
            f[self._plane(d, 0)] = self._beno(u1, u2, u3, norm)

    @staticmethod
    def _beno(u1, u2, u3, norm):
        """pyrandaBC.py:510-521: the smaller of the first- and second-order extrapolation increments,
        none at an extremum; `norm` keeps a normal velocity between 0 and its inner neighbour."""
        xp = _xp(u1)
        d1 = u2 - u1
        d2 = (2.0 * u2 - u3) - u1
        beno = xp.where(xp.abs(d1) <= xp.abs(d2), u1 + d1, u1 + d2)
        beno = xp.where(d1 * d2 <= 0.0, u1, beno)
        if norm:
            zero = u2 * 0.0
            beno = xp.maximum(xp.minimum(beno, xp.maximum(zero, u2)), xp.minimum(zero, u2))
        return beno

    # pyrandaBC.py:202-265: unit normal of the constant-A / B / C surfaces from the inverse metrics
    def _normals(self, d):
        key = "slipbc-" + d
        if key not in self.BCdata:
            if self.getvar is None:
                raise ValueError("bc.slip needs the mesh metrics (curvilinear mesh)")
            g = {k: self.getvar(k) for k in ("dAx", "dAy", "dAz", "dBx", "dBy", "dBz", "dCx", "dCy", "dCz", "dtJ")}
            J = g["dtJ"]
            m = {"xA": (-g["dBz"] * g["dCy"] + g["dBy"] * g["dCz"]) * J, "xB": (g["dAz"] * g["dCy"] - g["dAy"] * g["dCz"]) * J,
                 "xC": (-g["dAz"] * g["dBy"] + g["dAy"] * g["dBz"]) * J,
                 "yA": (g["dBz"] * g["dCx"] - g["dBx"] * g["dCz"]) * J, "yB": (-g["dAz"] * g["dCx"] + g["dAx"] * g["dCz"]) * J,
                 "yC": (g["dAz"] * g["dBx"] - g["dAx"] * g["dBz"]) * J,
                 "zA": (-g["dBy"] * g["dCx"] + g["dBx"] * g["dCy"]) * J, "zB": (g["dAy"] * g["dCx"] - g["dAx"] * g["dCy"]) * J,
                 "zC": (-g["dAy"] * g["dBx"] + g["dAx"] * g["dBy"]) * J}
            t1, t2 = {"x": ("B", "C"), "y": ("A", "C"), "z": ("A", "B")}[d[0]]
            a1, a2, a3 = m["x" + t1], m["y" + t1], m["z" + t1]
            b1, b2, b3 = m["x" + t2], m["y" + t2], m["z" + t2]
            n1, n2, n3 = (a2 * b3 - a3 * b2), -(a1 * b3 - a3 * b1), (a1 * b2 - a2 * b1)
            mag = _xp(n1).sqrt(n1 * n1 + n2 * n2 + n3 * n3)
            pl = self._plane(d, 0)
            self.BCdata[key] = [(n1 / mag)[pl], (n2 / mag)[pl], (n3 / mag)[pl]]
        return self.BCdata[key]

    # pyrandaBC.py:186-200,267-466: `bc.slip([['u','v']], ['y1'])`; x and y walls, as in the reference
    def slip(self, var, direction):
        for d in _as_list(direction):
            if d[0] not in ("x", "y") or d[1:] not in ("1", "n"):
                continue
            for velocity in _as_list(var):
                velocity = _as_list(velocity)
                norms = self._normals(d)
                self.extrap(velocity, d, order=2)  # "for free slip, always extrapolate first"
                if not self.owns.get(d, False):
                    continue
                pl = self._plane(d, 0)
                U = [self.variables[v] for v in velocity]
                xp = _xp(U[0])
                udotn = 0.0
                for u, n in zip(U, norms):
                    udotn = udotn + u[pl] * n
                mag0 = 0.0
                for u in U:
                    mag0 = mag0 + u[pl] ** 2
                mag0 = xp.sqrt(mag0)
                for u, n in zip(U, norms):
                    u[pl] = u[pl] - udotn * n
                magF = 0.0
                for u in U:
                    magF = magF + u[pl] ** 2
                magF = xp.sqrt(magF)
                for u in U:  # the projection must not lengthen the vector
                    u[pl] = xp.where(magF > mag0, u[pl] * mag0 / magF, u[pl])

    # pyrandaBC.py:612-746 with Reimann (:566-608): boundary plane from the plane next to it and the
    # free-stream state, through the incoming / outgoing Riemann invariants along the face normal
    def farfield(self, direction):
        for d in _as_list(direction):
            if d[0] not in _AXIS or d[1:] not in ("1", "n"):
                raise ValueError("unknown boundary '%s'" % d)
            if not self.owns.get(d, False):
                continue
            ref = self.BCdata["farfield-properties-%s" % d]
            gamma = ref["gamma"]
            nx, ny, nz = self._normals(d)
            inner, edge = self._plane(d, 1), self._plane(d, 0)
            V = {k: self.variables[ref[k]] for k in ("rho", "u", "v", "w", "p")}
            rhoi, Ui, Vi, Wi, Pi = (V[k][inner] for k in ("rho", "u", "v", "w", "p"))
            xp = _xp(Ui)
            one = Ui * 0.0 + 1.0
            rhoo, Uo, Vo, Wo, Po = (one * ref[k] for k in ("rho0", "u0", "v0", "w0", "p0"))
            Vni = Ui * nx + Vi * ny + Wi * nz
            Vno = Uo * nx + Vo * ny + Wo * nz
            SSo = xp.sqrt(gamma * Po / rhoo)
            SSi = xp.sqrt(gamma * Pi / rhoi)
            mach = xp.sqrt(Ui ** 2 + Vi ** 2 + Wi ** 2) / SSi
            Rplus = xp.where(mach >= 1.0, Vno + 2.0 * SSo / (gamma - 1.0), Vni + 2.0 * SSi / (gamma - 1.0))
            Rminus = xp.where(mach >= 1.0, Vni - 2.0 * SSi / (gamma - 1.0), Vno - 2.0 * SSo / (gamma - 1.0))
            Vnormal = (Rminus + Rplus) * 0.5
            SSb = (Rplus - Rminus) * (gamma - 1.0) * 0.25
            out = Vnormal >= 0.0
            Ub = xp.where(out, Ui + (Vnormal - Vni) * nx, Uo + (Vnormal - Vno) * nx)
            Vb = xp.where(out, Vi + (Vnormal - Vni) * ny, Vo + (Vnormal - Vno) * ny)
            Wb = xp.where(out, Wi + (Vnormal - Vni) * nz, Wo + (Vnormal - Vno) * nz)
            aux = 1.0 / (gamma - 1.0)
            rhob = xp.where(out, ((rhoi ** gamma * SSb ** 2) / (gamma * Pi)) ** aux, ((rhoo ** gamma * SSb ** 2) / (gamma * Po)) ** aux)
            Pb = rhob * SSb ** 2 / gamma
            V["rho"][edge], V["u"][edge], V["v"][edge], V["w"][edge], V["p"][edge] = rhob, Ub, Vb, Wb, Pb
