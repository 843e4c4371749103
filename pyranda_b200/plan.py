"""ParcopPlan: one (patch, level) of the reference's operator state, living on one B200.

Host-side mirror of what ``parcop.setup`` + ``setup_mesh`` create in the reference
(pyranda/parcop/parcop.f90:23-81, pyranda/pyrandaMPI.py:151-155) and of the per-operator f2py
entry points (parcop.f90:202-379).  Every method dispatches through the C ABI in
include/parcop_b200.h; nothing is computed in Python.

Array conventions (same as f2py): float64, Fortran order, shape (ax, ay, az).
  * numpy arrays  -> host path: H2D copy, device operator, D2H copy (``pb_host_*``).
  * torch CUDA tensors -> device-resident path (``pb_apply`` on ``data_ptr()``), result is a new
    CUDA tensor with Fortran strides; used by the RK4 loop so fields never leave HBM.
"""
import ctypes

import numpy as np

from . import _lib
from ._lib import OP, REDUCE, ParcopError, check

_vp = ctypes.c_void_p


def _is_torch(a):
    return type(a).__module__.split(".")[0] == "torch"


class ParcopPlan:
    def __init__(self, nx, ny, nz, x1=0.0, xn=1.0, y1=0.0, yn=1.0, z1=0.0, zn=1.0,
                 periodic=(False, False, False), px=1, py=1, pz=1, coords=(0, 0, 0), coordsys=0,
                 symmetric=((False, False), (False, False), (False, False)), device=-1, lib=None,
                 tensor_device="cuda"):
        self.L = lib if lib is not None else _lib.load()
        # where device-resident fields live; "cpu" only makes sense with the emulated library (tests)
        self.tensor_device = tensor_device
        bcs = []
        for d in range(3):  # pyrandaMPI.py:101-131
            b1 = bn = b"NONE"
            if periodic[d]:
                b1 = bn = b"PERI"
            if symmetric[d][0]:
                b1 = b"SYMM"
            if symmetric[d][1]:
                bn = b"SYMM"
            bcs += [b1, bn]
        h = _vp()
        check(self.L, self.L.pb_plan_create(ctypes.byref(h), nx, ny, nz, px, py, pz, coords[0], coords[1],
                                            coords[2], coordsys, float(x1), float(xn), float(y1), float(yn),
                                            float(z1), float(zn), *bcs, int(device)))
        self._h = h
        self.global_shape = (nx, ny, nz)
        self.procs = (px, py, pz)
        self.coords = tuple(coords)
        self.periodic = tuple(bool(p) for p in periodic)
        self.coordsys = coordsys
        ax, ay, az = ctypes.c_int(), ctypes.c_int(), ctypes.c_int()
        check(self.L, self.L.pb_plan_extents(h, ctypes.byref(ax), ctypes.byref(ay), ctypes.byref(az)))
        self.shape = (ax.value, ay.value, az.value)
        self.npts = ax.value * ay.value * az.value
        dx, dy, dz = ctypes.c_double(), ctypes.c_double(), ctypes.c_double()
        check(self.L, self.L.pb_plan_spacing(h, ctypes.byref(dx), ctypes.byref(dy), ctypes.byref(dz)))
        self.dx, self.dy, self.dz = dx.value, dy.value, dz.value
        self._keep = []

    def close(self):
        if getattr(self, "_h", None):
            self.L.pb_plan_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # ------------------------------------------------------------------ mesh
    def set_mesh(self, x=None, y=None, z=None, periodic_grid=True):
        """parcop.setup_mesh / setup_mesh_x3 (parcop.f90:64-81)."""
        if x is None:
            check(self.L, self.L.pb_plan_set_mesh(self._h, None, None, None, int(periodic_grid)))
        else:
            xs = [np.asfortranarray(a, dtype=np.float64) for a in (x, y, z)]
            for a in xs:
                if a.shape != self.shape:
                    raise ParcopError("mesh array shape %s != local extents %s" % (a.shape, self.shape))
            check(self.L, self.L.pb_plan_set_mesh(self._h, xs[0].ctypes.data, xs[1].ctypes.data,
                                                  xs[2].ctypes.data, int(periodic_grid)))

    def getvar(self, name):
        """parcop.getvar / xgrid / dxgrid / mesh_getcellvol / mesh_getgridlen -> numpy array."""
        out = np.zeros(self.shape, dtype=np.float64, order="F")
        check(self.L, self.L.pb_getvar(self._h, name.encode(), out.ctypes.data))
        return out

    def getvar_ptr(self, name):
        p = _vp()
        check(self.L, self.L.pb_getvar_device(self._h, name.encode(), ctypes.byref(p)))
        return p.value

    # ------------------------------------------------------------------ array plumbing
    def _host_in(self, val):
        a = np.asfortranarray(val, dtype=np.float64)
        if a.shape != self.shape:  # compact_d1.f90:65-68 stops the program; we raise
            raise ParcopError("operand shape %s != local extents %s" % (a.shape, self.shape))
        return a

    def _host_out(self):
        return np.empty(self.shape, dtype=np.float64, order="F")

    def empty_device(self, like=None):
        """A CUDA field with Fortran strides (pyrandaMPI.emptyScalar, device resident)."""
        import torch
        ax, ay, az = self.shape
        if like is not None:
            dev = like.device
        elif self.tensor_device == "cuda":
            dev = torch.device("cuda", torch.cuda.current_device())
        else:
            dev = torch.device(self.tensor_device)
        return torch.empty((az, ay, ax), dtype=torch.float64, device=dev).permute(2, 1, 0)

    def _dev_in(self, t):
        import torch
        ax, ay, az = self.shape
        if tuple(t.shape) != self.shape or t.dtype != torch.float64 or t.device.type != torch.device(self.tensor_device).type:
            raise ParcopError("operand must be a float64 %s tensor of shape %s" % (self.tensor_device, self.shape))
        if t.stride() != (1, ax, ax * ay):
            f = self.empty_device(t)
            f.copy_(t)
            t = f
        return t

    def _stream(self):
        import torch
        return torch.cuda.current_stream().cuda_stream if self.tensor_device == "cuda" else 0

    # ------------------------------------------------------------------ operators
    def apply_ptr(self, op, in_ptr, out_ptr, stream=0):
        """Raw device-pointer call (pb_apply)."""
        check(self.L, self.L.pb_apply(self._h, OP[op] if isinstance(op, str) else int(op), in_ptr, out_ptr, stream))

    def apply(self, op, val):
        code = OP[op]
        if _is_torch(val):
            t = self._dev_in(val)
            out = self.empty_device(t)
            check(self.L, self.L.pb_apply(self._h, code, t.data_ptr(), out.data_ptr(), self._stream()))
            return out
        a = self._host_in(val)
        out = self._host_out()
        check(self.L, self.L.pb_host_apply(self._h, code, a.ctypes.data, out.ctypes.data))
        return out

    def apply_host_into(self, op, a_in, a_out):
        """Host path into a caller-owned (e.g. pinned) Fortran-ordered output array."""
        if not (a_in.flags.f_contiguous and a_out.flags.f_contiguous and a_in.shape == self.shape == a_out.shape):
            raise ParcopError("apply_host_into needs Fortran-ordered float64 arrays of the local extents")
        check(self.L, self.L.pb_host_apply(self._h, OP[op], a_in.ctypes.data, a_out.ctypes.data))

    # parcop.f90:225-368, one method per f2py subroutine (lower-case names as f2py exports them)
    def ddx(self, val): return self.apply("ddx", val)
    def ddy(self, val): return self.apply("ddy", val)
    def ddz(self, val): return self.apply("ddz", val)
    # d1x/d1y/d1z(v, dv, bc=-1) (compact_operators.f90:13-128): fields odd across the symmetry planes
    def ddx_odd(self, val): return self.apply("ddx_odd", val)
    def ddy_odd(self, val): return self.apply("ddy_odd", val)
    def ddz_odd(self, val): return self.apply("ddz_odd", val)
    def dd8x_odd(self, val): return self.apply("dd8x_odd", val)  # d8x(v, dv, bc=-1), compact_operators.f90:280-313
    def dd8y_odd(self, val): return self.apply("dd8y_odd", val)
    def dd8z_odd(self, val): return self.apply("dd8z_odd", val)
    def dd4x(self, val): return self.apply("dd4x", val)
    def dd4y(self, val): return self.apply("dd4y", val)
    def dd4z(self, val): return self.apply("dd4z", val)
    def dd8x(self, val): return self.apply("dd8x", val)
    def dd8y(self, val): return self.apply("dd8y", val)
    def dd8z(self, val): return self.apply("dd8z", val)
    def d2x(self, val): return self.apply("d2x", val)
    def d2y(self, val): return self.apply("d2y", val)
    def d2z(self, val): return self.apply("d2z", val)
    def plaplacian(self, val): return self.apply("laplacian", val)
    def pring(self, val): return self.apply("ring", val)
    def sfilter(self, val): return self.apply("sfilter", val)
    def gfilter(self, val): return self.apply("gfilter", val)

    def gfilterdir(self, val, direction):
        return self.apply(("gfilterx", "gfiltery", "gfilterz")[int(direction) - 1], val)

    def divergencetensor(self, fxx, fxy, fxz, fyx, fyy, fyz, fzx, fzy, fzz):
        """parcop.f90:213-223: (dfx, dfy, dfz), host arrays or device tensors."""
        ins = (fxx, fxy, fxz, fyx, fyy, fyz, fzx, fzy, fzz)
        if _is_torch(fxx):
            t = [self._dev_in(a) for a in ins]
            outs = [self.empty_device(t[0]) for _ in range(3)]
            check(self.L, self.L.pb_divergence_tensor(self._h, *[a.data_ptr() for a in t], *[o.data_ptr() for o in outs], self._stream()))
            return tuple(outs)
        h = [self._host_in(a) for a in ins]
        outs = [self._host_out() for _ in range(3)]
        vp = ctypes.c_void_p
        check(self.L, self.L.pb_host_divergence_tensor(self._h, (vp * 9)(*[a.ctypes.data for a in h]), (vp * 3)(*[o.ctypes.data for o in outs])))
        return tuple(outs)

    def pringv(self, vx, vy, vz):
        """parcop.f90:324-333."""
        if _is_torch(vx):
            a, b, c = self._dev_in(vx), self._dev_in(vy), self._dev_in(vz)
            out = self.empty_device(a)
            check(self.L, self.L.pb_ring_vector(self._h, a.data_ptr(), b.data_ptr(), c.data_ptr(), out.data_ptr(), self._stream()))
            return out
        a, b, c = self._host_in(vx), self._host_in(vy), self._host_in(vz)
        out = self._host_out()
        check(self.L, self.L.pb_host_ring_vector(self._h, a.ctypes.data, b.ctypes.data, c.ctypes.data, out.ctypes.data))
        return out

    def sfilterdir(self, val, direction):
        return self.apply(("sfilterx", "sfiltery", "sfilterz")[int(direction) - 1], val)

    def divergence(self, fx, fy, fz):
        if _is_torch(fx):
            a, b, c = self._dev_in(fx), self._dev_in(fy), self._dev_in(fz)
            out = self.empty_device(a)
            check(self.L, self.L.pb_divergence(self._h, a.data_ptr(), b.data_ptr(), c.data_ptr(), out.data_ptr(), self._stream()))
            return out
        a, b, c = self._host_in(fx), self._host_in(fy), self._host_in(fz)
        out = self._host_out()
        check(self.L, self.L.pb_host_divergence(self._h, a.ctypes.data, b.ctypes.data, c.ctypes.data, out.ctypes.data))
        return out

    def grads(self, val):
        if _is_torch(val):
            t = self._dev_in(val)
            gx, gy, gz = self.empty_device(t), self.empty_device(t), self.empty_device(t)
            check(self.L, self.L.pb_grads(self._h, t.data_ptr(), gx.data_ptr(), gy.data_ptr(), gz.data_ptr(), self._stream()))
            return gx, gy, gz
        a = self._host_in(val)
        gx, gy, gz = self._host_out(), self._host_out(), self._host_out()
        check(self.L, self.L.pb_host_grads(self._h, a.ctypes.data, gx.ctypes.data, gy.ctypes.data, gz.ctypes.data))
        return gx, gy, gz

    # ------------------------------------------------------------------ RK4 / reductions
    def rk4_stage(self, dt, A, B, F, PHI, U):
        """PHI = dt*F + A*PHI ; U += B*PHI, in place on device tensors (pyranda.py:800-804)."""
        check(self.L, self.L.pb_rk4_stage(self._h, self.npts, float(dt), float(A), float(B),
                                          F.data_ptr(), PHI.data_ptr(), U.data_ptr(), self._stream()))

    def reduce(self, kind, t):
        """Local sum / max / min of a device field (pyrandaMPI.py:307-326, before the allreduce)."""
        out = ctypes.c_double()
        check(self.L, self.L.pb_reduce(self._h, REDUCE[kind], t.numel(), t.data_ptr(), ctypes.byref(out), self._stream()))
        return out.value

    def reduce_device(self, kind, t):
        """The same reduction without the host round trip: returns a 0-dim CUDA tensor."""
        import torch
        out = torch.empty((), dtype=torch.float64, device=t.device)
        check(self.L, self.L.pb_reduce_device(self._h, REDUCE[kind], t.numel(), t.data_ptr(), out.data_ptr(), self._stream()))
        return out

    def launch_count(self):
        return int(self.L.pb_launch_count())
