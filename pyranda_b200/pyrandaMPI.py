"""Mirror of the parts of pyranda/pyrandaMPI.py that sit on the hot path.

`pyrandaMPI(mesh)` keeps the reference's attribute names (nx, ax, chunk_3d_*, x1proc, ...), its
`der / fil / gfil` dispatch objects (pyrandaMPI.py:207-209, 664-740), `emptyScalar`
(:249-250) and the `sum3D / max3D / min3D` reductions (:307-326), with fields living on the GPU.
MPI is replaced by torch.distributed: one process per GPU, z-slab decomposition (1, 1, G).
Host utilities of the reference class (ghost exchange for IO, gathers for probes) are out of scope.
"""
import numpy as np

from .plan import ParcopPlan


class parcop_der:
    """pyrandaMPI.py:664-712."""

    def __init__(self, eng): self._e = eng
    def ddx(self, val): return self._e.op("ddx", val)
    def ddy(self, val): return self._e.op("ddy", val)
    def ddz(self, val): return self._e.op("ddz", val)
    def dd4x(self, val): return self._e.op("dd4x", val)  # pyrandaMPI.py:678-686
    def dd4y(self, val): return self._e.op("dd4y", val)
    def dd4z(self, val): return self._e.op("dd4z", val)
    def dd8x(self, val): return self._e.op("dd8x", val)
    def dd8y(self, val): return self._e.op("dd8y", val)
    def dd8z(self, val): return self._e.op("dd8z", val)
    def laplacian(self, val): return self._e.op("laplacian", val)
    def ring(self, val): return self._e.op("ring", val)
    def div(self, fx, fy, fz): return self._e.div(fx, fy, fz)
    def divT(self, *f9): return self._e.divT(*f9)                  # pyrandaMPI.py:699-700
    def ringV(self, vx, vy, vz): return self._e.ringV(vx, vy, vz)  # pyrandaMPI.py:711-712
    def grad(self, val): return self._e.grad(val)


class parcop_sfil:
    """pyrandaMPI.py:734-740."""

    def __init__(self, eng): self._e = eng
    def filter(self, val): return self._e.op("sfilter", val)


class parcop_gfil:
    """pyrandaMPI.py:721-730."""

    def __init__(self, eng): self._e = eng
    def filter(self, val): return self._e.op("gfilter", val)
    def filterDir(self, val, direction): return self._e.op(("gfilterx", "gfiltery", "gfilterz")[int(direction) - 1], val)


class pyrandaMPI:
    def __init__(self, mesh, comm=None, device=None, lib=None, tensor_device=None):
        opt = mesh.options if hasattr(mesh, "options") else mesh
        if isinstance(opt, str):
            from .sim import parse_mesh
            opt = parse_mesh(opt)
        self.nx, self.ny, self.nz = (int(v) for v in opt["nn"])
        x1, xn = opt["x1"], opt["xn"]
        self.dx = (xn[0] - x1[0]) / max(self.nx - 1, 1)  # pyrandaMPI.py:43-45
        self.dy = (xn[1] - x1[1]) / max(self.ny - 1, 1)
        self.dz = (xn[2] - x1[2]) / max(self.nz - 1, 1)
        self.periodic = tuple(bool(p) for p in opt.get("periodic", (False,) * 3))
        self.coordsys = int(opt.get("coordsys", 0))
        # meshOptions['symmetric'] -> "SYMM" boundary strings (pyrandaMPI.py:51,120-131)
        self.symmetric = tuple((bool(a), bool(b)) for a, b in opt.get("symmetric", ((False, False),) * 3))
        self.order = (10, 10, 10)
        self.filter_type = ("compact", "compact", "compact")
        world, rank = 1, 0
        try:
            import torch.distributed as dist
            if dist.is_available() and dist.is_initialized():
                world, rank = dist.get_world_size(comm), dist.get_rank(comm)
        except ImportError:
            pass
        self.comm, self.world, self.rank = comm, world, rank
        self.px, self.py, self.pz = 1, 1, world
        dev = -1 if device is None else int(device)
        args = (self.nx, self.ny, self.nz, x1[0], xn[0], x1[1], xn[1], x1[2], xn[2])
        if world > 1:
            from .distributed import DistributedParcop
            self._dist = DistributedParcop(*args, periodic=self.periodic, coordsys=self.coordsys, device=dev, group=comm,
                                           symmetric=self.symmetric, lib=lib, tensor_device=tensor_device)
            self.plan = self._dist.plan
        else:
            self._dist = None
            kw = {} if tensor_device is None else {"tensor_device": tensor_device}
            self.plan = ParcopPlan(*args, periodic=self.periodic, coordsys=self.coordsys, device=dev, symmetric=self.symmetric,
                                   lib=lib, **kw)
        self.ax, self.ay, self.az = self.plan.shape
        # makeMesh always ends in setup_mesh / setup_mesh_x3 (pyrandaMPI.py:151-174, pyrandaMesh.py:93-135):
        # ring, getVar and the curvilinear operators need the mesh arrays
        if self.coordsys == 3:
            from .sim import curvilinear_coordinates
            lo = (0, 0, rank * self.az)
            self.plan.set_mesh(*curvilinear_coordinates(opt, lo, self.plan.shape), periodic_grid=bool(opt.get("periodicGrid", True)))
        else:
            self.plan.set_mesh()
        self.chunk_3d_size = np.array(self.plan.shape, dtype=np.int32)
        self.chunk_3d_lo = np.array([0, 0, rank * self.az], dtype=np.int32)
        self.chunk_3d_hi = self.chunk_3d_lo + self.chunk_3d_size - 1
        self.master = rank == 0
        self.x1proc = self.xnproc = self.y1proc = self.ynproc = True
        self.z1proc, self.znproc = rank == 0, rank == world - 1
        self.der, self.fil, self.gfil = parcop_der(self), parcop_sfil(self), parcop_gfil(self)

    # ---- dispatch ----
    def op(self, name, val):
        if self._dist is not None:
            return self._dist.apply(name, val)
        return self.plan.apply(name, val)

    def div(self, fx, fy, fz):
        return self._dist.divergence(fx, fy, fz) if self._dist is not None else self.plan.divergence(fx, fy, fz)

    def grad(self, val):
        return self._dist.grads(val) if self._dist is not None else self.plan.grads(val)

    def divT(self, *f9):
        return self._dist.divergencetensor(*f9) if self._dist is not None else self.plan.divergencetensor(*f9)

    def ringV(self, vx, vy, vz):
        return self._dist.pringv(vx, vy, vz) if self._dist is not None else self.plan.pringv(vx, vy, vz)

    def setPatch(self):  # pyrandaMPI.py:246-247: one plan per instance, nothing to select
        pass

    def emptyScalar(self):  # pyrandaMPI.py:249-250
        t = self.plan.empty_device()
        t.zero_()
        return t

    def getVar(self, vname):  # pyrandaMPI.py:657-661
        return self.plan.getvar(vname)

    # ---- pyrandaMPI.py:307-326 ----
    def sum3D(self, data):
        return self._dist.sum3D(data) if self._dist is not None else self.plan.reduce("sum", data)

    def max3D(self, data):
        return self._dist.max3D(data) if self._dist is not None else self.plan.reduce("max", data)

    def min3D(self, data):
        return self._dist.min3D(data) if self._dist is not None else self.plan.reduce("min", data)

    def iprint(self, sprnt):
        if self.master:
            print(sprnt, flush=True)
