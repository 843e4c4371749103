"""The f2py module surface of the reference (pyranda/parcop/parcop.f90, built with
``f2py --lower``), re-pointed at libparcop_b200.so.

Every function has the reference's Python-visible call shape: ``intent(out)`` arrays are return
values, sizes are inferred from the operand, names are lower case.  State is module-global and
selected with ``set_patch`` exactly like the Fortran module's current-patch pointers
(parcop.f90:196-200, objects.f90:77-89); each (patch, level) slot owns one ``ParcopPlan``.

Operands may be numpy arrays (host path: H2D, kernels, D2H -- what an unmodified pyranda sees) or
float64 CUDA tensors with Fortran strides (device-resident path used by pyranda_b200.sim).
"""
import numpy as np

from ..plan import ParcopPlan
from .._lib import ParcopError

_plans = {}
_current = None


def _cur():
    if _current is None:
        raise ParcopError("parcop.setup / set_patch has not been called")
    return _plans[_current]


def setup(patch, level, comm, nx, ny, nz, px, py, pz, coordsys, x1, xn, y1, yn, z1, zn,
          bx1, bxn, by1, byn, bz1, bzn, coords=(0, 0, 0), device=-1, lib=None, tensor_device="cuda"):
    """parcop.f90:23-61.  `comm` is accepted for call compatibility (the reference passes an MPI
    Fortran handle); rank coordinates come from `coords` (z-slab: (0, 0, rank)).  `lib` / `tensor_device`
    select another build of the library (the host-emulated one of tests/emul in CPU tests)."""
    global _current
    periodic = tuple(str(b).strip().upper() == "PERI" for b in (bx1, by1, bz1))
    symmetric = tuple((str(a).strip().upper() == "SYMM", str(b).strip().upper() == "SYMM")
                      for a, b in ((bx1, bxn), (by1, byn), (bz1, bzn)))
    key = (int(patch), int(level))
    if key in _plans:
        _plans[key].close()
    _plans[key] = ParcopPlan(nx, ny, nz, x1, xn, y1, yn, z1, zn, periodic=periodic, px=px, py=py, pz=pz,
                             coords=coords, coordsys=coordsys, symmetric=symmetric, device=device, lib=lib,
                             tensor_device=tensor_device)
    _current = key


def set_patch(patch, level):
    """parcop.f90:196-200."""
    global _current
    key = (int(patch), int(level))
    if key not in _plans:
        raise ParcopError("patch %s level %s was never set up" % key)
    _current = key


def plan(patch=None, level=None):
    """The ParcopPlan behind a slot (an extension: the reference has no handle to return)."""
    return _cur() if patch is None else _plans[(int(patch), int(level))]


def setup_mesh(patch, level):
    """parcop.f90:64-70."""
    _plans[(int(patch), int(level))].set_mesh()


def setup_mesh_x3(patch, level, x1, x2, x3, meshper):
    """parcop.f90:73-81."""
    _plans[(int(patch), int(level))].set_mesh(x1, x2, x3, periodic_grid=bool(meshper))


# ---- mesh getters (parcop.f90:84-191); the size arguments are accepted and ignored ----------------
def getvar(vname, nx=None, ny=None, nz=None):
    return _cur().getvar(str(vname).strip())


def xgrid(nx=None, ny=None, nz=None): return _cur().getvar("x")
def ygrid(nx=None, ny=None, nz=None): return _cur().getvar("y")
def zgrid(nx=None, ny=None, nz=None): return _cur().getvar("z")
def dxgrid(nx=None, ny=None, nz=None): return _cur().getvar("d1")
def dygrid(nx=None, ny=None, nz=None): return _cur().getvar("d2")
def dzgrid(nx=None, ny=None, nz=None): return _cur().getvar("d3")
def mesh_getcellvol(nx=None, ny=None, nz=None): return _cur().getvar("CellVol")
def mesh_getgridlen(nx=None, ny=None, nz=None): return _cur().getvar("GridLen")


# ---- operators (parcop.f90:202-379) ---------------------------------------------------------------
def ddx(val): return _cur().ddx(val)
def ddy(val): return _cur().ddy(val)
def ddz(val): return _cur().ddz(val)
def dd4x(val): return _cur().dd4x(val)  # parcop.f90:255-277
def dd4y(val): return _cur().dd4y(val)
def dd4z(val): return _cur().dd4z(val)
def dd8x(val): return _cur().dd8x(val)
def dd8y(val): return _cur().dd8y(val)
def dd8z(val): return _cur().dd8z(val)
def plaplacian(val): return _cur().plaplacian(val)
def pring(val): return _cur().pring(val)
def sfilter(val): return _cur().sfilter(val)
def gfilter(val): return _cur().gfilter(val)
def gfilterdir(val, direction): return _cur().gfilterdir(val, direction)
def divergence(fx, fy, fz): return _cur().divergence(fx, fy, fz)
def divergencetensor(fxx, fxy, fxz, fyx, fyy, fyz, fzx, fzy, fzz):  # parcop.f90:213-223
    return _cur().divergencetensor(fxx, fxy, fxz, fyx, fyy, fyz, fzx, fzy, fzz)
def pringv(vx, vy, vz): return _cur().pringv(vx, vy, vz)  # parcop.f90:324-333
def grads(val): return _cur().grads(val)


# communicator getters (parcop.f90:406-440): there is no MPI here; ranks are torch.distributed ranks
def commx(): return None
def commy(): return None
def commz(): return None
def commxy(): return None
def commxz(): return None
def commyz(): return None
