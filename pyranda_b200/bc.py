"""Boundary-plane packages of the EOM interpreter: bc.extrap / bc.const / bc.field.

Reference: pyranda/pyrandaBC.py:40-186 (the `BC` package: `bc.extrap(vars, dirs, order)`,
`bc.const(vars, dirs, val)`, `bc.field(var, dirs, field)` lines inside an EOM string).  They sit in
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

    # pyrandaBC.py:162-186
    def field(self, var, direction, field):
        for f, d in self._each(var, direction):
            f[self._plane(d, 0)] = field[self._plane(d, 0)]
