"""Device-resident mirror of pyrandaSim's time loop (reference: pyranda/pyranda.py).

What is mirrored, and only that: the mesh mini-language (pyrandaMesh.py:33-56), the EOM / IC string
interpreter (pyranda.py:231-416,812-869, pyrandaUtils.py:22-66, pyrandaEq.py:16-45), the operator
forwards (pyranda.py:607-736), the low-storage 5-stage RK4 (pyranda.py:758-810) and the time-step
package (pyrandaTimestep.py:42-77).  User decks (the strings) are unchanged; what changes is where
the arrays live: every variable is a float64 CUDA tensor with Fortran strides, operators go through
the C ABI (ParcopPlan), pointwise expressions run as device tensor ops, and the stage update is the
fused pb_rk4_stage kernel -- nothing round-trips to the host inside a step except the scalar
results of max / min / sum reductions (as in the reference, which allreduces them).

The class is generic over a small `backend` object so the SAME driver can be run on numpy arrays
with another operator provider; tests use that to compare the CUDA path with the CPU oracle after
many steps.  The product ships only CudaBackend.
"""
import math
import os
import re

import numpy as np


# --------------------------------------------------------------------------------------------------
class CudaBackend:
    """Fields are CUDA tensors; operators are libparcop_b200 kernels behind a ParcopPlan."""

    def __init__(self, plan):
        import torch
        self.torch = torch
        self.plan = plan
        self.xp = _TorchNS(torch)
        self.device = (torch.device("cuda", torch.cuda.current_device()) if plan.tensor_device == "cuda"
                       else torch.device(plan.tensor_device))

    def zeros(self):
        t = self.plan.empty_device()
        t.zero_()
        return t

    def asfield(self, a):
        t = self.plan.empty_device()
        t.copy_(self.torch.as_tensor(np.ascontiguousarray(a), device=self.device))
        return t

    def tohost(self, t):
        return t.cpu().numpy() if hasattr(t, "cpu") else t

    def isfield(self, a):
        return isinstance(a, self.torch.Tensor)

    def _f(self, a):
        """Operators need a full Fortran-strided field; broadcast scalars / fix strides."""
        if not self.isfield(a):
            t = self.plan.empty_device()
            t.fill_(float(a))
            return t
        return a

    # operators (pyranda.py:607-736 -> pyrandaMPI.py:664-740)
    def ddx(self, v): return self.plan.ddx(self._f(v))
    def ddy(self, v): return self.plan.ddy(self._f(v))
    def ddz(self, v): return self.plan.ddz(self._f(v))
    def divT(self, *f9): return self.plan.divergencetensor(*[self._f(a) for a in f9])
    def ringV(self, a, b, c): return self.plan.pringv(self._f(a), self._f(b), self._f(c))
    def dd4x(self, v): return self.plan.dd4x(self._f(v))
    def dd4y(self, v): return self.plan.dd4y(self._f(v))
    def dd4z(self, v): return self.plan.dd4z(self._f(v))
    def dd8x(self, v): return self.plan.dd8x(self._f(v))
    def dd8y(self, v): return self.plan.dd8y(self._f(v))
    def dd8z(self, v): return self.plan.dd8z(self._f(v))
    def filter(self, v): return self.plan.sfilter(self._f(v))
    def gfilter(self, v): return self.plan.gfilter(self._f(v))
    def gfilterdir(self, v, d): return self.plan.gfilterdir(self._f(v), d)
    def ring(self, v): return self.plan.pring(self._f(v))
    def laplacian(self, v): return self.plan.plaplacian(self._f(v))
    def div(self, a, b, c): return self.plan.divergence(self._f(a), self._f(b), self._f(c))
    def grad(self, v): return self.plan.grads(self._f(v))
    def getvar(self, name): return self.asfield(self.plan.getvar(name))

    # reductions (pyrandaMPI.py:307-326)
    def sum3D(self, a): return self.plan.reduce("sum", self._c(a)) if self.isfield(a) else float(a)
    def max3D(self, a): return self.plan.reduce("max", self._c(a)) if self.isfield(a) else float(a)
    def min3D(self, a): return self.plan.reduce("min", self._c(a)) if self.isfield(a) else float(a)
    # the same, result left on the device (no stream sync): the time-step controller uses these
    def max3D_dev(self, a): return self.plan.reduce_device("max", self._c(a))
    def min3D_dev(self, a): return self.plan.reduce_device("min", self._c(a))

    def _c(self, a):
        """`a` as a field with Fortran strides (1, ax, ax*ay): the kernels walk flat memory."""
        ax, ay, az = self.plan.shape
        return a if a.stride() == (1, ax, ax * ay) else self.plan.empty_device(a).copy_(a)

    def storage_key(self, a):
        return a.data_ptr()

    def clone(self, a):
        return self.plan.empty_device(a).copy_(a)

    def rk4_stage(self, dt, A, B, F, PHI, U):
        """pyranda.py:800-804 fused; returns the new U (updated in place)."""
        F = self._c(self._f(F))
        U = self._c(U)
        ax, ay, az = self.plan.shape
        for t in (F, PHI, U):
            assert t.stride() == (1, ax, ax * ay), "rk4_stage operands must share the Fortran layout"
        self.plan.rk4_stage(dt, A, B, F, PHI, U)
        return U


class _TorchNS:
    """The handful of `numpy.` functions decks and sMap use (pyranda.py:842-858), on tensors or floats."""

    def __init__(self, torch):
        self.t = torch
        self.pi = math.pi

    def _u(self, name, a):
        return getattr(self.t, name)(a) if isinstance(a, self.t.Tensor) else getattr(math, {"abs": "fabs"}.get(name, name))(a)

    def sqrt(self, a): return self._u("sqrt", a)
    def abs(self, a): return self._u("abs", a)
    def sin(self, a): return self._u("sin", a)
    def cos(self, a): return self._u("cos", a)
    def tanh(self, a): return self._u("tanh", a)
    def exp(self, a): return self._u("exp", a)
    def sign(self, a): return self.t.sign(a) if isinstance(a, self.t.Tensor) else float(np.sign(a))

    def _b(self, name, a, b):
        T = self.t.Tensor
        if isinstance(a, T) or isinstance(b, T):
            ref = a if isinstance(a, T) else b
            a = a if isinstance(a, T) else self.t.full_like(ref, float(a))
            b = b if isinstance(b, T) else self.t.full_like(ref, float(b))
            return getattr(self.t, name)(a, b)
        return getattr(np, name)(a, b)

    def minimum(self, a, b): return self._b("minimum", a, b)
    def maximum(self, a, b): return self._b("maximum", a, b)

    def where(self, c, a, b):
        T = self.t.Tensor
        a = a if isinstance(a, T) else self.t.full_like(c, float(a), dtype=self.t.float64)
        b = b if isinstance(b, T) else self.t.full_like(c, float(b), dtype=self.t.float64)
        return self.t.where(c, a, b)


# --------------------------------------------------------------------------------------------------
_FUNCS = {  # deck function -> method of the simulation object (pyranda.py:817-858)
    "ddx": "self.ddx", "ddy": "self.ddy", "ddz": "self.ddz", "div": "self.div", "divT": "self.divT", "ringV": "self.ringV", "grad": "self.grad",
    "fbar": "self.filter", "gbar": "self.gfilter", "gbarx": "self.gfilterx", "gbary": "self.gfiltery",
    "gbarz": "self.gfilterz", "lap": "self.laplacian", "ring": "self.ring", "dd8x": "self.dd8x",
    "dd4x": "self.dd4x", "dd4y": "self.dd4y", "dd4z": "self.dd4z",  # pyranda.py:833-835
    "dd8y": "self.dd8y", "dd8z": "self.dd8z", "sum": "self.B.sum3D", "max": "self.B.max3D", "min": "self.B.min3D",
    "mean": "self.mean", "sign": "xp.sign", "abs": "xp.abs", "sqrt": "xp.sqrt", "sin": "xp.sin", "cos": "xp.cos",
    "tanh": "xp.tanh", "exp": "xp.exp", "where": "xp.where", "3d": "self.emptyScalar", "random3D": "self.random3D", "meshVar": "self.B.getvar",
    "dt.courant": "self.dt_courant", "dt.diff": "self.dt_diff", "dt.diffDir": "self.dt_diff_dir",
    "bc.extrap": "self.bc.extrap", "bc.const": "self.bc.const", "bc.field": "self.bc.field", "bc.symm": "self.bc.symm",
    "bc.exit": "self.bc.exit", "bc.slip": "self.bc.slip", "bc.farfield": "self.bc.farfield",  # pyrandaBC.py:28-38
    "ibmV": "self.ibm.velocity_slip", "ibmWall": "self.ibm.velocity_wall", "ibmS": "self.ibm.scalar",  # pyrandaIBM.py:27-31
    "numpy.minimum": "xp.minimum",
    "numpy.maximum": "xp.maximum", "numpy.sqrt": "xp.sqrt", "numpy.abs": "xp.abs", "numpy.where": "xp.where",
}
_NAMES = {"simtime": "self.time", "deltat": "self.deltat", "pi": "xp.pi", "meshx": 'self.variables["meshx"]',
          "meshy": 'self.variables["meshy"]', "meshz": 'self.variables["meshz"]', "gridLen": "self.GridLen",
          "meshi": "self.mesh.indices[0].data", "meshj": "self.mesh.indices[1].data", "meshk": "self.mesh.indices[2].data"}  # pyranda.py:864-866
_VAR = re.compile(r":([A-Za-z_]\w*):")
_CALL = re.compile(r"(?<![\w.\"])((?:dt\.|numpy\.|bc\.)?[A-Za-z_3]\w*)\(")
_WORD = re.compile(r"(?<![\w.\"])([A-Za-z_]\w*)(?![\w(\"])")


def translate(expr, user=()):
    """Deck expression -> Python source evaluated with `self` (the sim) and `xp` in scope.
    `user`: names added with addUserDefinedFunction, called as f(self, ...) (pyranda.py:156-161)."""
    expr = expr.replace(" ", "").strip()
    expr = _VAR.sub(lambda m: 'self.variables["%s"]' % m.group(1), expr)

    def call(m):
        if m.group(1) in user:
            return "self.userDefined['%s'](self," % m.group(1)
        return (_FUNCS[m.group(1)] + "(") if m.group(1) in _FUNCS else m.group(0)
    expr = _CALL.sub(call, expr)
    expr = _WORD.sub(lambda m: _NAMES.get(m.group(1), m.group(1)), expr)
    return expr


def _lines(text):
    out = []
    for ln in text.split("\n"):
        ln = ln.strip()
        if ln and not ln.startswith("#"):
            out.append(ln)
    return out


class _Equation:
    def __init__(self, text, fuser=None, user=()):
        self.text = text
        # an assignment has only variables (or ddt(variable)) left of its "="; a package call such as
        # `bc.extrap([...], [...], order=1)` is evaluated for its side effect
        head = text.split("=", 1)[0]
        assign = "=" in text and re.fullmatch(r"\s*(ddt\(\s*)?\[?\s*:\w+:\s*(,\s*:\w+:\s*)*\]?\s*\)?\s*", head) is not None
        lhs, rhs = text.split("=", 1) if assign else (None, text)
        self.lhs = _VAR.findall(lhs) if lhs is not None else None
        self.kind = "PDE" if "ddt(" in text else "ALG"  # pyrandaEq.py:42-43
        self.src = _drop_abs_of_ring(translate(rhs, user))
        self.raw = self.src   # before fusion: the flux plan re-fuses the PDE lines after hoisting their operator arguments
        self.opaque = bool(user) and any(("userDefined['%s']" % u) in self.src for u in user) or any(
            k in self.src for k in ("self.bc.", "self.ibm.", "self.random3D", "self.emptyScalar"))
        self.reads = set(_VAR.findall(rhs))
        self.pure = None  # AST when the right-hand side is arithmetic over variables only (groupable)
        if fuser is not None:  # arithmetic between operator calls -> one generated kernel each (fuse.py)
            if self.kind == "ALG" and self.lhs is not None and len(self.lhs) == 1:
                self.pure = fuser.pure_tree(self.src)
            self.src = fuser.transform(self.src)
        self.code = compile(self.src, "<eom>", "eval")


def _drop_abs_of_ring(src):
    """abs(ring(f)) == ring(f) bit for bit: the detector is a maximum of |d8 f| * d^2 (operators.f90:615-643),
    so the deck's abs() around it needs no pass over the field."""
    import ast
    try:
        tree = ast.parse(src, mode="eval")
    except SyntaxError:
        return src
    changed = [False]

    class T(ast.NodeTransformer):
        def visit_Call(self, node):
            self.generic_visit(node)
            f = node.func
            if (isinstance(f, ast.Attribute) and isinstance(f.value, ast.Name) and f.value.id == "xp" and f.attr == "abs"
                    and len(node.args) == 1 and not node.keywords):
                a = node.args[0]
                if (isinstance(a, ast.Call) and isinstance(a.func, ast.Attribute) and isinstance(a.func.value, ast.Name)
                        and a.func.value.id == "self" and a.func.attr in ("ring", "ringV")):
                    changed[0] = True
                    return a
            return node
    tree = T().visit(tree)
    return ast.unparse(ast.fix_missing_locations(tree)) if changed[0] else src


def parse_mesh(text):
    """`xdom = (x1, xn, n, periodic=True)` lines (pyrandaMesh.py:33-56)."""
    opt = {"x1": [0.0, 0.0, 0.0], "xn": [1.0, 1.0, 1.0], "nn": [1, 1, 1], "periodic": [False, False, False],
           "symmetric": [[False, False], [False, False], [False, False]]}  # pyrandaMesh.py:222

    def setter(ind):
        def f(x1, xn, nn, periodic=False):
            opt["x1"][ind], opt["xn"][ind], opt["nn"][ind], opt["periodic"][ind] = float(x1), float(xn), int(nn), bool(periodic)
        return f
    ns = {"xdom": None, "ydom": None, "zdom": None}
    for ln in _lines(text):
        ln = ln.replace(" ", "")
        for k, name in enumerate(("xdom", "ydom", "zdom")):
            if ln.startswith(name + "=("):
                eval("f" + ln[len(name) + 1:], {"f": setter(k)})
    return opt


class _Data:
    """`.data` holder, the one attribute of pyrandaVar that user-defined functions of the example
    decks read (`pysim.mesh.indices[0].data`, `pysim.mesh.coords[0].data`)."""

    def __init__(self, data):
        self.data = data


class _MeshView:
    """`pysim.mesh` as user code sees it (pyrandaMesh.py:58-208): coords, global indices, nn, GridLen."""

    def __init__(self, sim, lo):
        B = sim.B
        self.nn = [sim.nx, sim.ny, sim.nz]
        self.coordsys = sim.coordsys
        self.coords = [_Data(sim.variables[k]) for k in ("meshx", "meshy", "meshz")]
        self.GridLen = sim.GridLen
        self.d1, self.d2, self.d3 = sim.d1, sim.d2, sim.d3
        self.shape = list(tuple(sim.zero.shape))
        self._B, self._lo, self._indices = B, lo, None

    @property
    def indices(self):
        """Global index fields iloc / jloc / kloc (pyrandaMesh.py:63-81), built on first use: three
        more fields that most decks never read."""
        if self._indices is None:
            idx = np.meshgrid(*[np.arange(self.shape[d]) + self._lo[d] for d in range(3)], indexing="ij")
            self._indices = [_Data(self._B.asfield(np.asfortranarray(a, dtype=np.float64))) for a in idx]
        return self._indices


class _PyMPIView:
    """`pysim.PyMPI` as user code sees it: sizes, ownership flags and the directional sums of
    pyrandaMPI.py:307-357 on fields of the backend (device tensors or numpy arrays).  On a z-slab
    the sums over z are completed across the ranks by the backend's `allsum` / `allgather`."""

    def __init__(self, sim, lo):
        self._sim = sim
        B = sim.B
        self.nx, self.ny, self.nz = sim.nx, sim.ny, sim.nz
        self.ax, self.ay, self.az = tuple(sim.zero.shape)
        self.chunk_3d_lo = np.array(lo, dtype=np.int32)
        self.chunk_3d_hi = self.chunk_3d_lo + np.array([self.ax, self.ay, self.az], dtype=np.int32) - 1
        owns = getattr(B, "owns", None) or {}
        for nm in ("x1", "xn", "y1", "yn", "z1", "zn"):
            setattr(self, nm + "proc", bool(owns.get(nm, True)))
        self.master = lo[2] == 0
        self.sum3D, self.max3D, self.min3D = B.sum3D, B.max3D, B.min3D

    def _sum(self, data, axes):
        return data.sum(axis=axes) if isinstance(data, np.ndarray) else data.sum(dim=axes)

    def _z_complete(self, part, z_summed):
        """Local partial sums -> global result: summed over the ranks when z was reduced, gathered
        along z otherwise (one rank: nothing to do)."""
        B = self._sim.B
        if z_summed:
            return B.allsum(part) if hasattr(B, "allsum") else part
        return B.allgather_z(part) if hasattr(B, "allgather_z") else part

    # reduction along two directions -> global 1-D profile (pyrandaMPI.py:349-357)
    def yzsum(self, data): return self._z_complete(self._sum(data, (1, 2)), True)
    def xzsum(self, data): return self._z_complete(self._sum(data, (0, 2)), True)
    def xysum(self, data): return self._z_complete(self._sum(data, (0, 1)), False)
    # reduction along one direction -> global 2-D array (pyrandaMPI.py:330-347)
    def xsum(self, data): return self._z_complete(self._sum(data, (0,)), False)
    def ysum(self, data): return self._z_complete(self._sum(data, (1,)), False)
    def zsum(self, data): return self._z_complete(self._sum(data, (2,)), True)
    xbar, ybar, zbar = xsum, ysum, zsum

    def iprint(self, sprnt):
        if self.master:
            print(sprnt, flush=True)


def curvilinear_coordinates(opt, lo=(0, 0, 0), shape=None):
    """Coordinates of a `coordsys = 3` mesh as pyrandaMesh.makeMesh builds them
    (pyrandaMesh.py:93-129): the uniform grid between x1 and xn, overwritten point by point with
    `function(i, j, k)` of the GLOBAL indices when the options carry one.  `lo` / `shape`: the local
    block of a decomposed mesh."""
    nn = opt["nn"]
    shape = tuple(nn) if shape is None else tuple(shape)
    ax = [np.linspace(opt["x1"][d], opt["xn"][d], num=nn[d])[lo[d]:lo[d] + shape[d]] for d in range(3)]
    x, y, z = (np.asfortranarray(a) for a in np.meshgrid(*ax, indexing="ij"))
    fn = opt.get("function")
    if fn:
        # a function of index ARRAYS (numpy fancy indexing, as the example decks' lambdas allow) fills the
        # whole block at once; anything else is evaluated point by point like pyrandaMesh.makeMesh does
        try:
            I, J, K = np.meshgrid(*[np.arange(shape[d]) + lo[d] for d in range(3)], indexing="ij")
            got = fn(I, J, K)
            arrs = [np.asfortranarray(np.broadcast_to(np.asarray(g, dtype=np.float64), shape)) for g in got]
            if len(arrs) == 3:
                probe = [(0, 0, 0), (shape[0] - 1, shape[1] - 1, shape[2] - 1), (shape[0] // 2, shape[1] // 3, shape[2] // 2)]
                if all(np.allclose([a[q] for a in arrs], fn(q[0] + lo[0], q[1] + lo[1], q[2] + lo[2]), rtol=0, atol=0) for q in probe):
                    return tuple(arrs)
        except Exception:
            pass
        for i in range(shape[0]):
            for j in range(shape[1]):
                for k in range(shape[2]):
                    x[i, j, k], y[i, j, k], z[i, j, k] = fn(i + lo[0], j + lo[1], k + lo[2])
    return x, y, z


class pyrandaSim:
    """`ss = pyrandaSim(name, mesh); ss.EOM(eom); ss.setIC(ic); ss.rk4(time, dt)` on the device."""

    def __init__(self, name, mesh, backend=None, device=0):
        self.name = name
        opt = parse_mesh(mesh) if isinstance(mesh, str) else mesh
        self.meshOptions = opt
        self.nx, self.ny, self.nz = opt["nn"]
        self.npts = self.nx * self.ny * self.nz
        self.coordsys = int(opt.get("coordsys", 0))
        if backend is None:
            from .plan import ParcopPlan
            plan = ParcopPlan(self.nx, self.ny, self.nz, opt["x1"][0], opt["xn"][0], opt["x1"][1], opt["xn"][1],
                              opt["x1"][2], opt["xn"][2], periodic=tuple(opt["periodic"]), device=device,
                              coordsys=self.coordsys,
                              symmetric=tuple(tuple(s) for s in opt.get("symmetric", ((False, False),) * 3)))
            if self.coordsys == 3:  # pyrandaMesh.py:93-135: the grid comes from Python
                plan.set_mesh(*curvilinear_coordinates(opt), periodic_grid=bool(opt.get("periodicGrid", True)))
            else:
                plan.set_mesh()
            backend = CudaBackend(plan)
        self.B = backend
        self.xp = backend.xp
        self.variables = {}
        self.userDefined = {}
        self.vizDumpHistory = []
        self._cmetric = None
        self.equations = []
        self.conserved = []
        self.time, self.deltat, self.cycle = 0.0, 0.0, 0
        for k, nm in enumerate(("meshx", "meshy", "meshz")):
            self.variables[nm] = backend.getvar("xyz"[k])
        self.d1, self.d2, self.d3 = (backend.getvar(n) for n in ("d1", "d2", "d3"))
        self.GridLen = backend.getvar("GridLen")
        self.zero = backend.zeros()
        self._ns = {"xp": self.xp, "numpy": self.xp, "self": self}
        lo = tuple(getattr(backend, "chunk_lo", (0, 0, 0)))  # first global index of this rank's block
        self.mesh = _MeshView(self, lo)
        self.PyMPI = _PyMPIView(self, lo)
        from .bc import BoundaryConditions
        # the `BC` package (pyrandaBC.py), on the fields in place; a z-slab backend says which faces are its own
        self.bc = BoundaryConditions(self.variables, getattr(backend, "owns", None), getvar=backend.getvar)
        from .ibm import ImmersedBoundary
        self.ibm = ImmersedBoundary(self)             # the `IBM` package (pyrandaIBM.py)
        self.fuser = None
        self._plan = {}
        self._fplan = None
        self._hoisted = {}
        if isinstance(backend, CudaBackend) and os.environ.get("PB_NO_FUSE", "0") != "1":
            from .fuse import Fuser
            self.fuser = Fuser(self.xp)
            self._ns["__fz"] = self.fuser.call
            # the time-step controller as fused expressions (same operations, same order as below)
            self._courant = compile(self.fuser.transform(
                "xp.abs(u)/self.d1 + xp.abs(v)/self.d2 + xp.abs(w)/self.d3 + xp.abs(c)/self.GridLen"), "<dt>", "eval")
            self._diffrate = compile(self.fuser.transform(
                "density*self.GridLen*self.GridLen/xp.maximum(1.0e-12, bulk)"), "<dt>", "eval")

    # ---- operator forwards (pyranda.py:607-736) ----
    def ddx(self, v): return 0.0 if self.nx <= 1 else self.B.ddx(v)
    def ddy(self, v): return 0.0 if self.ny <= 1 else self.B.ddy(v)
    def ddz(self, v): return 0.0 if self.nz <= 1 else self.B.ddz(v)
    def divT(self, *f9): return self.B.divT(*f9)          # pyranda.py:673-674
    def ringV(self, vx, vy, vz): return self.B.ringV(vx, vy, vz)  # pyranda.py:686-687
    def dd4x(self, v): return self.B.dd4x(v)  # pyranda.py:622-629
    def dd4y(self, v): return self.B.dd4y(v)
    def dd4z(self, v): return self.B.dd4z(v)
    def dd8x(self, v): return self.B.dd8x(v)
    def dd8y(self, v): return self.B.dd8y(v)
    def dd8z(self, v): return self.B.dd8z(v)
    def filter(self, v): return self.B.filter(v)
    def gfilter(self, v): return self.B.gfilter(v)
    def gfilterx(self, v): return self.B.gfilterdir(v, 1)
    def gfiltery(self, v): return self.B.gfilterdir(v, 2)
    def gfilterz(self, v): return self.B.gfilterdir(v, 3)
    def ring(self, v): return self.B.ring(v)
    def laplacian(self, v): return self.B.laplacian(v)
    def grad(self, v): return self.B.grad(v)

    def div(self, f1, f2=None, f3=None):  # pyranda.py:640-671
        z = self.zero
        if f2 is None and f3 is None:
            if self.nx > 1: return self.B.div(f1, z, z)
            if self.ny > 1: return self.B.div(z, f1, z)
            if self.nz > 1: return self.B.div(z, z, f1)
            return 0.0
        if f3 is None:
            if self.nx > 1 and self.ny > 1: return self.B.div(f1, f2, z)
            if self.nz > 1 and self.ny > 1: return self.B.div(z, f1, f2)
            if self.nz > 1 and self.nx > 1: return self.B.div(f1, z, f2)
            if self.nx > 1: return self.B.div(f1, z, z)
            if self.ny > 1: return self.B.div(z, f1, z)
            if self.nz > 1: return self.B.div(z, z, f1)
            return 0.0
        return self.B.div(f1, f2, f3)

    def emptyScalar(self, val=0.0):
        return self.B.zeros() + val

    def mean(self, a):
        return 1.0 / float(self.npts) * self.B.sum3D(a)

    # ---- pyrandaTimestep.py:42-77 ----
    def dt_courant(self, u, v, w, c):
        if self.coordsys == 3:  # :45-55: velocities along the (2-D) grid lines over the local spacings
            xp = self.xp
            if self._cmetric is None:
                g = {k: self.B.getvar(k) for k in ("dAx", "dAy", "dBx", "dBy")}
                self._cmetric = (g, xp.sqrt(g["dAx"] * g["dAx"] + g["dAy"] * g["dAy"]),
                                 xp.sqrt(g["dBx"] * g["dBx"] + g["dBy"] * g["dBy"]))
            g, magA, magB = self._cmetric
            uA = (u * g["dAx"] + v * g["dAy"]) / magA
            uB = (u * g["dBx"] + v * g["dBy"]) / magB
            vrate = xp.abs(uA) / self.d1 + xp.abs(uB) / self.d2
            return 1.0 / self.B.max3D(vrate + xp.abs(c) / self.GridLen)
        if self.fuser is not None:  # one kernel + a reduction whose result stays on the device
            return 1.0 / self.B.max3D_dev(eval(self._courant, self._ns, {"u": u, "v": v, "w": w, "c": c}))
        xp = self.xp
        vrate = xp.abs(u) / self.d1 + xp.abs(v) / self.d2 + xp.abs(w) / self.d3
        crate = xp.abs(c) / self.GridLen
        return 1.0 / self.B.max3D(vrate + crate)

    def dt_diff(self, bulk, density):
        if self.fuser is not None:
            return self.B.min3D_dev(eval(self._diffrate, self._ns, {"bulk": bulk, "density": density}))
        delta = self.GridLen
        drate = density * delta * delta / self.xp.maximum(1.0e-12, bulk)
        return self.B.min3D(drate)

    def dt_diff_dir(self, bulk, density, delta):  # pyrandaTimestep.py:79-85: the same limit with a caller-given length
        return self.B.min3D(density * delta * delta / self.xp.maximum(1.0e-12, bulk))

    # ---- interpreter (pyranda.py:231-416) ----
    @staticmethod
    def _apply_dict(text, d):
        """pyranda.py:219-228: every key of the dictionary is replaced by str(value), textually."""
        for key in (d or {}):
            text = text.replace(key, str(d[key]))
        return text

    def addUserDefinedFunction(self, name, function):
        """pyranda.py:156-161: `name(args)` in a deck line calls function(sim, args)."""
        if name in _FUNCS:
            raise ValueError("cannot add user-defined function '%s': the name is taken" % name)
        self.userDefined[name] = function

    def random3D(self):
        """`random3D()` of a deck (pyranda.py:857): numpy's global stream, so numpy.random.seed applies."""
        return self.B.asfield(np.asfortranarray(np.random.random(tuple(self.zero.shape))))

    def EOM(self, eom, eomDict=None):
        eom = self._apply_dict(eom, eomDict)
        for ln in _lines(eom):
            eq = _Equation(ln, self.fuser, tuple(self.userDefined))
            self.equations.append(eq)
            for nm in _VAR.findall(ln):
                self.variables.setdefault(nm, self.B.zeros())
            if eq.kind == "PDE":
                self.conserved.append(eq.lhs[0])

    def setIC(self, ics, icDict=None):
        ics = self._apply_dict(ics, icDict)
        local = {}
        for ln in _lines(ics):
            for nm in _VAR.findall(ln):
                self.variables.setdefault(nm, self.B.zeros())
            exec(translate(ln, tuple(self.userDefined)), self._ns, local)
        self.updateVars()

    def _flux_plan(self):
        """Once per deck: the operator arguments of all PDE lines hoisted into one kernel (fuse.hoist),
        the lines re-fused, and for lines of the form `arithmetic(operator results)` the split that lets
        the RK4 stage kernel evaluate that arithmetic itself."""
        pdes = [eq for eq in self.equations if eq.kind == "PDE"]
        if self._fplan is not None and self._fplan[0] == len(self.equations):
            return self._fplan[1]
        plan = {"gid": None, "items": []}
        if self.fuser is not None and os.environ.get("PB_NO_HOIST", "0") != "1" and not any(eq.opaque for eq in pdes):
            gid, srcs = self.fuser.hoist([eq.raw for eq in pdes])
            plan["gid"] = gid
            for eq, src in zip(pdes, srcs):
                fsrc = self.fuser.transform(src)
                split = self.fuser.split_stage(fsrc)
                item = {"name": eq.lhs[0], "code": compile(fsrc, "<eom>", "eval"), "stage": None}
                if split is not None:
                    item["stage"] = (split[0], compile(split[1], "<eom>", "eval"))
                plan["items"].append(item)
        else:
            plan["items"] = [{"name": eq.lhs[0], "code": eq.code, "stage": None} for eq in pdes]
        self._fplan = (len(self.equations), plan)
        return plan

    def updateFlux(self):  # pyranda.py:376-394
        plan = self._flux_plan()
        if plan["gid"] is not None:
            self.fuser.run_group(plan["gid"], self.variables, out=self._hoisted)
        flux = {it["name"]: eval(it["code"], self._ns) for it in plan["items"]}
        self._hoisted.clear()
        return flux

    def _host_only(self):
        """Variables nothing on the device depends on during a step: assigned by algebraic lines and read
        only by lines that assign such variables (the time-step controller's `:dt:` chain, `:cs:` feeding
        it, diagnostics like `:enst:`).  Between the stages of one rk4() call nobody can look at them, so
        the lines that only assign them run after the last stage only.  A deck with package calls or user
        functions (which may read any variable) keeps everything."""
        if any(eq.opaque for eq in self.equations) or os.environ.get("PB_NO_LAZY", "0") == "1":
            return set()
        algs = [eq for eq in self.equations if eq.kind == "ALG" and eq.lhs]
        dead = set(nm for eq in algs for nm in eq.lhs) - set(self.conserved)
        changed = True
        while changed:
            changed = False
            for eq in self.equations:
                keeps = eq.kind == "PDE" or not eq.lhs or any(nm not in dead for nm in eq.lhs)
                if keeps and (eq.reads & dead):  # a line that runs every stage reads it
                    dead -= eq.reads
                    changed = True
        return dead

    def _alg_plan(self, final=True):
        """Consecutive algebraic equations that are pure arithmetic become one multi-output kernel."""
        plan, run = [], []
        dead = set() if final else self._host_only()

        def flush():
            if len(run) >= 2:
                plan.append(("group", self.fuser.make_group([(e.lhs[0], e.pure) for e in run])))
            else:
                plan.extend(("eq", e) for e in run)
            run.clear()
        for eq in self.equations:
            if eq.kind != "ALG":
                continue
            if eq.lhs and all(nm in dead for nm in eq.lhs):
                continue  # host-only result, not the last stage
            if self.fuser is not None and eq.pure is not None:
                run.append(eq)
            else:
                flush()
                plan.append(("eq", eq))
        flush()
        return plan

    def updateVars(self, final=True):  # pyranda.py:397-416
        if self.fuser is not None:
            cur = self._plan.get(final)
            if cur is None or cur[0] != len(self.equations):
                cur = self._plan[final] = (len(self.equations), self._alg_plan(final))
            for kind, item in cur[1]:
                if kind == "group":
                    self.fuser.run_group(item, self.variables)
                    continue
                rhs = eval(item.code, self._ns)
                if not item.lhs:
                    continue
                if len(item.lhs) == 1:
                    self.variables[item.lhs[0]] = rhs
                else:
                    for nm, r in zip(item.lhs, rhs):
                        self.variables[nm] = r
            return
        for eq in self.equations:
            if eq.kind != "ALG":
                continue
            rhs = eval(eq.code, self._ns)
            if not eq.lhs:
                continue
            if len(eq.lhs) == 1:
                self.variables[eq.lhs[0]] = rhs
            else:
                for nm, r in zip(eq.lhs, rhs):
                    self.variables[nm] = r

    def var(self, name):
        return self.variables[name]

    # ---- viz dumps (pyranda.py:431-470, pyrandaIO.py:53-231) ----
    def write(self, wVars=(), root=None):
        """`ss.write(['rho', 'u'])`: one legacy-VTK STRUCTURED_GRID file per rank and dump under
        `<root>/vis<cycle>/proc-<rank>.<cycle>.vtk` plus the `pyranda.visit` index, with the
        reference's names (pyranda.py:440-470).  Binary, big-endian float32 as VTK requires; CYCLE and
        TIME field data as pyrandaIO.py:79-84.  Off the step loop: each variable is staged device ->
        host once.  On a z-slab every block carries one plane of its neighbours (the reference's
        pyrandaMPI.ghost, pyrandaIO.py:62-67,107-108), so the pieces of a dump share a plane."""
        names = list(wVars) if wVars else list(self.conserved)
        root = self.name if root is None else root
        rank = int(self.PyMPI.chunk_3d_lo[2] // max(self.PyMPI.az, 1))
        nblocks = max(self.nz // max(self.PyMPI.az, 1), 1)
        dump = "vis" + str(self.cycle).zfill(7)
        os.makedirs(os.path.join(root, dump), exist_ok=True)
        path = os.path.join(root, dump, "proc-%s.%s.vtk" % (str(rank).zfill(6), str(self.cycle).zfill(7)))
        ghost = getattr(self.B, "ghost_host", None)  # z-slab backends: neighbour planes appended
        host = lambda a: np.asarray(ghost(a) if ghost is not None else self.B.tohost(a), dtype=np.float64)
        xyz = [host(self.variables[k]) for k in ("meshx", "meshy", "meshz")]
        ax, ay, az = xyz[0].shape
        with open(path, "wb") as fid:
            fid.write(b"# vtk DataFile Version 3.0\nvtk output\nBINARY\nDATASET STRUCTURED_GRID\n")
            fid.write(b"FIELD FieldData 2\nCYCLE 1 1 int\n")
            fid.write(np.array([self.cycle], dtype=">i4").tobytes())
            fid.write(b"\nTIME 1 1 double\n")
            fid.write(np.array([self.time], dtype=">f8").tobytes())
            fid.write(("\nDIMENSIONS %d %d %d\nPOINTS %d float\n" % (ax, ay, az, ax * ay * az)).encode())
            pts = np.stack([a.ravel(order="F") for a in xyz], axis=1)  # x fastest, as the reference's loops
            fid.write(pts.astype(">f4").tobytes())
            fid.write(("\nPOINT_DATA %d\n" % (ax * ay * az)).encode())
            for nm in names:
                fid.write(("SCALARS %s float\nLOOKUP_TABLE default\n" % nm).encode())
                fid.write(host(self.variables[nm]).ravel(order="F").astype(">f4").tobytes())
                fid.write(b"\n")
        self.vizDumpHistory.append([self.cycle, self.time])
        if rank == 0:
            with open(os.path.join(root, "pyranda.visit"), "w") as vid:
                vid.write("!NBLOCKS %s \n" % nblocks)
                for cyc, tm in self.vizDumpHistory:
                    vid.write("!TIME %s \n" % tm)
                    for p_ in range(nblocks):
                        vid.write("%s\n" % os.path.join("vis" + str(cyc).zfill(7), "proc-%s.%s.vtk" % (str(p_).zfill(6), str(cyc).zfill(7))))
        return path

    # ---- restart (pyranda.py:475-588: the whole state, for a later run) ----
    def _restart_file(self, path, suffix):
        """`path` as given on one rank; on a z-slab every rank has its own file (`<path>.proc-NNNNN`).  Without a
        path: `restart_<cycle or suffix>/proc-NNNNN.npz`, the directory naming of pyranda.py:484-490."""
        lo = getattr(self.PyMPI, "chunk_3d_lo", (0, 0, 0))
        az = max(int(getattr(self.PyMPI, "az", self.nz)), 1)
        rank, split = int(lo[2] // az), az < self.nz
        if path is None:
            d = "restart" + (suffix if suffix else "_" + str(self.cycle).zfill(7))
            os.makedirs(d, exist_ok=True)
            return os.path.join(d, "proc-%s.npz" % str(rank).zfill(5))
        path = str(path)
        if split:
            path = (path[:-4] if path.endswith(".npz") else path) + ".proc-%s.npz" % str(rank).zfill(5)
        return path if path.endswith(".npz") else path + ".npz"

    def writeRestart(self, path=None, suffix=None):
        """One .npz per rank with every variable (fields staged device -> host, numbers as they are), the
        equation strings, time, cycle, the last time step and the rank's extents.  A bare `ss.writeRestart()`
        works as in the reference (pyranda.py:475-490).  Off the step loop: the only D2H of a run."""
        path = self._restart_file(path, suffix)
        fields, scalars = {}, {}
        for nm, val in self.variables.items():
            if self.B.isfield(val) and getattr(val, "ndim", 3) == 3:
                fields["f_" + nm] = np.asarray(self.B.tohost(val))
            else:
                scalars[nm] = float(val)
        meta = {"time": self.time, "cycle": self.cycle, "deltat": float(self.deltat), "scalars": scalars,
                "equations": [eq.text for eq in self.equations], "nn": [self.nx, self.ny, self.nz]}
        meta["lo"] = [int(v) for v in getattr(self.PyMPI, "chunk_3d_lo", (0, 0, 0))]
        np.savez(path, __meta__=np.array(repr(meta)), **fields)
        return path

    def readRestart(self, path=None, suffix=None):
        """Restores a state written by writeRestart on a simulation built with the same mesh and partition."""
        import ast as _ast
        with np.load(self._restart_file(path, suffix)) as z:
            meta = _ast.literal_eval(str(z["__meta__"]))
            if list(meta["nn"]) != [self.nx, self.ny, self.nz]:
                raise ValueError("restart was written on a %s grid" % (meta["nn"],))
            lo = [int(v) for v in getattr(self.PyMPI, "chunk_3d_lo", (0, 0, 0))]
            if list(meta.get("lo", lo)) != lo:
                raise ValueError("restart was written by the rank at %s, this rank starts at %s" % (meta["lo"], lo))
            if not self.equations:
                self.EOM("\n".join(meta["equations"]))
            for key in z.files:
                if key.startswith("f_"):
                    self.variables[key[2:]] = self.B.asfield(np.asfortranarray(z[key]))
        self.variables.update(meta["scalars"])
        self.time, self.cycle, self.deltat = meta["time"], meta["cycle"], meta["deltat"]
        return self.time

    # ---- pyranda.py:758-810 ----
    ARK = (0.0, -6234157559845. / 12983515589748., -6194124222391. / 4410992767914.,
           -31623096876824. / 15682348800105., -12251185447671. / 11596622555746.)
    BRK = (494393426753. / 4806282396855., 4047970641027. / 5463924506627., 9795748752853. / 13190207949281.,
           4009051133189. / 8539092990294., 1348533437543. / 7166442652324.)
    ETA = (494393426753. / 4806282396855., 4702696611523. / 9636871101405., 3614488396635. / 5249666457482.,
           9766892798963. / 10823461281321., 1.0)

    def _unshare_conserved(self):
        """The stage update is in place on the device.  The reference rebinds `variables[U].data` to a
        new array every stage (pyranda.py:800-804), so a deck line such as `:phi0: = :phi:` or
        `:phi: = meshx` leaves the copy untouched; here a conserved field that shares its storage
        with another variable or a mesh array gets its own before every stage update."""
        B = self.B
        if not hasattr(B, "storage_key"):
            return
        count = {}
        for v in list(self.variables.values()) + [self.GridLen, self.d1, self.d2, self.d3, self.zero]:
            if B.isfield(v):
                k = B.storage_key(v)
                count[k] = count.get(k, 0) + 1
        for U in self.conserved:
            u = self.variables[U]
            if B.isfield(u) and count.get(B.storage_key(u), 0) > 1:
                count[B.storage_key(u)] -= 1
                self.variables[U] = B.clone(u)

    def _fused_stage(self, dt, A, B, PHI):
        """One stage on the device: every operator of every PDE line first (as updateFlux does, all
        right-hand sides before any update), then per conserved variable ONE kernel that evaluates the
        line's remaining arithmetic and the stage update (no flux field is written and read back)."""
        plan = self._flux_plan()
        if plan["gid"] is not None:
            self.fuser.run_group(plan["gid"], self.variables, out=self._hoisted)
        pending = []
        for it in plan["items"]:
            if it["stage"] is not None:
                pending.append((it, eval(it["stage"][1], self._ns)))
            else:
                pending.append((it, eval(it["code"], self._ns)))
        self._hoisted.clear()
        self._unshare_conserved()
        for it, val in pending:
            U = it["name"]
            u = self.B._c(self.variables[U])
            if it["stage"] is not None:
                if self.fuser.stage(it["stage"][0], val, dt, A, B, PHI[U], u):
                    self.variables[U] = u
                    continue
                val = self.fuser.call(it["stage"][0], *val)
            self.variables[U] = self.B.rk4_stage(dt, A, B, val, PHI[U], u)

    def rk4(self, time, dt):
        PHI = {U: self.B.zeros() for U in self.conserved}
        time_i = time
        dt = float(dt)  # a device scalar from the time-step controller: the one host read of the step
        self.deltat = dt
        for ii in range(5):
            if self.fuser is not None:
                self._fused_stage(dt, self.ARK[ii], self.BRK[ii], PHI)
            else:
                FLUX = self.updateFlux()
                self._unshare_conserved()
                for U in self.conserved:
                    self.variables[U] = self.B.rk4_stage(dt, self.ARK[ii], self.BRK[ii], FLUX[U], PHI[U], self.variables[U])
            time = time_i + self.ETA[ii] * dt
            self.time = time
            self.updateVars(final=(ii == 4))
        self.cycle += 1
        return time
