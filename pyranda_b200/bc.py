"""Boundary-plane packages of the EOM interpreter: bc.extrap / bc.const / bc.field / bc.symm.

Reference: pyranda/pyrandaBC.py:40-186,748-786 (the `BC` package: `bc.extrap(vars, dirs, order)`,
`bc.const(vars, dirs, val)`, `bc.field(var, dirs, field)`, `bc.symm(vars, dirs, anti, npts)` lines
inside an EOM string).  They sit in
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


class BoundaryConditions:
    def __init__(self, variables, owns=None):
        self.variables = variables
        self.owns = owns if owns is not None else {n: True for n in ("x1", "xn", "y1", "yn", "z1", "zn")}

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
