"""ctypes front-end of the CPU oracle (oracle/parcop_oracle.c).

TEST INFRASTRUCTURE ONLY -- see the header of parcop_oracle.c.  Imported by tests/, by
__graft_entry__.smoke() and by bench.py's cpu_baseline / ``--impl reference`` legs; never by
pyranda_b200/.

The class mirrors the Python-visible surface of the reference's f2py module
(pyranda/parcop/parcop.f90:23-379 as seen from pyranda/pyrandaMPI.py:151-155,664-740): Fortran
``intent(out)`` arrays come back as return values, inputs are Fortran-ordered float64.
"""
import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None

KINDS = {"d1": 0, "d2": 1, "d8": 2, "sf": 3, "gf": 4, "d4": 5}
BC = {"NONE": 0, "PERI": 1, "SYMM": 2}

_dp = ctypes.POINTER(ctypes.c_double)


def build(force=False):
    so = os.path.join(_HERE, "libparcop_oracle.so")
    src = os.path.join(_HERE, "parcop_oracle.c")
    if force or not os.path.exists(so) or os.path.getmtime(so) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", _HERE, "-s"])
    return so


def lib():
    global _LIB
    if _LIB is None:
        so = os.path.join(_HERE, "libparcop_oracle.so")
        if not os.path.exists(so):
            build()
        L = ctypes.CDLL(so)
        L.po_setup.restype = ctypes.c_void_p
        L.po_setup.argtypes = [ctypes.c_int] * 7 + [ctypes.c_double] * 6 + [ctypes.c_int] * 6
        L.po_free.argtypes = [ctypes.c_void_p]
        L.po_setup_mesh.argtypes = [ctypes.c_void_p]
        L.po_setup_mesh_x3.argtypes = [ctypes.c_void_p, _dp, _dp, _dp, ctypes.c_int]
        for name in ("po_ddx", "po_ddy", "po_ddz", "po_lap", "po_ring"):
            getattr(L, name).argtypes = [ctypes.c_void_p, _dp, _dp]
        L.po_dd8.argtypes = [ctypes.c_void_p, ctypes.c_int, _dp, _dp]
        L.po_d2.argtypes = [ctypes.c_void_p, ctypes.c_int, _dp, _dp]
        L.po_dd4.argtypes = [ctypes.c_void_p, ctypes.c_int, _dp, _dp]
        L.po_div.argtypes = [ctypes.c_void_p, _dp, _dp, _dp, _dp]
        L.po_divT.argtypes = [ctypes.c_void_p] + [_dp] * 12
        L.po_divT.restype = ctypes.c_int
        L.po_ringV.argtypes = [ctypes.c_void_p, _dp, _dp, _dp, _dp]
        L.po_grad.argtypes = [ctypes.c_void_p, _dp, _dp, _dp, _dp]
        L.po_filter.argtypes = [ctypes.c_void_p, ctypes.c_int, _dp, _dp]
        L.po_gfilter_dir.argtypes = [ctypes.c_void_p, ctypes.c_int, _dp, _dp]
        L.po_dir_op.argtypes = [ctypes.c_void_p, ctypes.c_int, ctypes.c_int, ctypes.c_int, _dp, _dp]
        L.po_eval_raw.argtypes = [ctypes.c_void_p, ctypes.c_int, ctypes.c_int, ctypes.c_int, _dp, _dp]
        L.po_getvar.argtypes = [ctypes.c_void_p, ctypes.c_char_p, _dp]
        L.po_spacing.restype = ctypes.c_double
        L.po_spacing.argtypes = [ctypes.c_void_p, ctypes.c_int]
        L.po_get_weight.argtypes = [ctypes.c_int, ctypes.POINTER(ctypes.c_int)] + [_dp] * 6
        L.po_get_tables.argtypes = [ctypes.c_void_p, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_int, _dp, _dp, _dp]
        L.po_num_threads.restype = ctypes.c_int
        _LIB = L
    return _LIB


def _f(a):
    return np.asfortranarray(a, dtype=np.float64)


def _p(a):
    return a.ctypes.data_as(_dp)


def get_weight(kind):
    """Coefficient tables of one operator (stencils.f90), for table-level tests."""
    ints = (ctypes.c_int * 6)()
    ali = np.zeros(5); ari = np.zeros(9)
    alb1 = np.zeros((4, 4, 5)); alb2 = np.zeros((4, 4, 5))
    arb1 = np.zeros((4, 4, 9)); arb2 = np.zeros((4, 4, 9))
    lib().po_get_weight(KINDS[kind], ints, _p(ali), _p(ari), _p(alb1), _p(alb2), _p(arb1), _p(arb2))
    nol, nor, ncl, ncr, null_option, implicit = list(ints)
    return dict(nol=nol, nor=nor, ncl=ncl, ncr=ncr, null_option=null_option, implicit=bool(implicit),
                ali=ali[:ncl], ari=ari[:ncr], alb1=alb1[:, :, :ncl], alb2=alb2[:, :, :ncl],
                arb1=arb1[:, :, :ncr], arb2=arb2[:, :, :ncr])


class Oracle:
    """One (patch, level) of the reference: parcop.setup + setup_mesh + the operator calls.

    px/py/pz emulate an MPI decomposition in-process; fields passed in are GLOBAL arrays.
    """

    def __init__(self, nx, ny, nz, x1=0.0, xn=1.0, y1=0.0, yn=1.0, z1=0.0, zn=1.0,
                 periodic=(False, False, False), px=1, py=1, pz=1, coordsys=0,
                 symmetric=((False, False), (False, False), (False, False)), mesh_xyz=None,
                 periodic_grid=True):
        L = lib()
        bcs = []
        for d in range(3):  # pyrandaMPI.py:101-131
            b1 = bn = "NONE"
            if periodic[d]:
                b1 = bn = "PERI"
            if symmetric[d][0]:
                b1 = "SYMM"
            if symmetric[d][1]:
                bn = "SYMM"
            bcs += [BC[b1], BC[bn]]
        self.shape = (nx, ny, nz)
        self._h = L.po_setup(nx, ny, nz, px, py, pz, coordsys, x1, xn, y1, yn, z1, zn, *bcs)
        if not self._h:
            raise ValueError("po_setup failed (sizes not divisible by the processor grid?)")
        if coordsys == 3:
            x, y, z = (_f(a) for a in mesh_xyz)
            rc = L.po_setup_mesh_x3(self._h, _p(x), _p(y), _p(z), int(bool(periodic_grid)))
        else:
            rc = L.po_setup_mesh(self._h)
        if rc != 0:
            raise ValueError("unsupported coordsys %d" % coordsys)
        self.dx, self.dy, self.dz = (L.po_spacing(self._h, d) for d in range(3))

    def __del__(self):
        try:
            if self._h:
                lib().po_free(self._h)
                self._h = None
        except Exception:
            pass

    def _new(self):
        return np.zeros(self.shape, dtype=np.float64, order="F")

    def _unary(self, fn, val, *pre):
        v = _f(val)
        assert v.shape == self.shape, (v.shape, self.shape)
        out = self._new()
        fn(self._h, *pre, _p(v), _p(out))
        return out

    # --- parcop.f90:225-301 -------------------------------------------------------------
    def ddx(self, val): return self._unary(lib().po_ddx, val)
    def ddy(self, val): return self._unary(lib().po_ddy, val)
    def ddz(self, val): return self._unary(lib().po_ddz, val)
    def dd8x(self, val): return self._unary(lib().po_dd8, val, 0)
    def dd8y(self, val): return self._unary(lib().po_dd8, val, 1)
    def dd8z(self, val): return self._unary(lib().po_dd8, val, 2)
    def dd4x(self, val): return self._unary(lib().po_dd4, val, 0)
    def dd4y(self, val): return self._unary(lib().po_dd4, val, 1)
    def dd4z(self, val): return self._unary(lib().po_dd4, val, 2)
    def d2x(self, val): return self._unary(lib().po_d2, val, 0)
    def d2y(self, val): return self._unary(lib().po_d2, val, 1)
    def d2z(self, val): return self._unary(lib().po_d2, val, 2)
    # --- parcop.f90:303-368 -------------------------------------------------------------
    def plaplacian(self, val): return self._unary(lib().po_lap, val)
    def pring(self, val): return self._unary(lib().po_ring, val)
    def sfilter(self, val): return self._unary(lib().po_filter, val, 0)
    def gfilter(self, val): return self._unary(lib().po_filter, val, 1)
    def gfilterdir(self, val, direction): return self._unary(lib().po_gfilter_dir, val, int(direction))

    def dir_op(self, kind, direction, val, bc=0):
        """d1x..filterz of compact_operators.f90 with the symmetry selector `bc`."""
        return self._unary(lib().po_dir_op, val, KINDS[kind], int(direction), int(bc))

    def eval_raw(self, kind, direction, val, iop=1):
        return self._unary(lib().po_eval_raw, val, KINDS[kind], int(direction), int(iop))

    # --- parcop.f90:202-211, 371-379 ----------------------------------------------------
    def divergence(self, fx, fy, fz):
        fx, fy, fz = _f(fx), _f(fy), _f(fz)
        out = self._new()
        lib().po_div(self._h, _p(fx), _p(fy), _p(fz), _p(out))
        return out

    def divergencetensor(self, fxx, fxy, fxz, fyx, fyy, fyz, fzx, fzy, fzz):
        """parcop.f90:213-223 (Cartesian)."""
        ins = [_f(a) for a in (fxx, fxy, fxz, fyx, fyy, fyz, fzx, fzy, fzz)]
        outs = [self._new(), self._new(), self._new()]
        if lib().po_divT(self._h, *[_p(a) for a in ins], *[_p(a) for a in outs]) != 0:
            raise ValueError("divT: only the Cartesian branch is restated")
        return tuple(outs)

    def pringv(self, vx, vy, vz):
        """parcop.f90:324-333."""
        a, b, c = _f(vx), _f(vy), _f(vz)
        out = self._new()
        lib().po_ringV(self._h, _p(a), _p(b), _p(c), _p(out))
        return out

    def grads(self, val):
        v = _f(val)
        a, b, c = self._new(), self._new(), self._new()
        lib().po_grad(self._h, _p(v), _p(a), _p(b), _p(c))
        return a, b, c

    def getvar(self, name):
        out = self._new()
        if lib().po_getvar(self._h, name.encode(), _p(out)) != 0:
            raise KeyError(name)
        return out

    def tables(self, kind, direction, rank=0, iop=1):
        n = self.shape[direction]
        al = np.zeros((9, n)); rc = np.zeros((4, n)); aa = np.zeros(16 * 4 * 64)
        nal = lib().po_get_tables(self._h, KINDS[kind], direction, iop, rank, _p(al), _p(rc), _p(aa))
        return nal, al, rc, aa


def num_threads():
    return lib().po_num_threads()
