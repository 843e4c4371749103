"""Generates tests/golden/bc_20x18.npz by running the REFERENCE's own BC package
(/root/reference/pyranda/pyrandaBC.py, loaded with a stub for its package base class, because
`import pyranda` needs mpi4py and the compiled parcop module) on a small curvilinear grid, the mesh
metrics served by the CPU oracle.  Covers `bc.exit` (exitbc / BENO, with and without `norm`) and
`bc.slip` (slipbc) on the x1 / xn / y1 / yn boundaries and `bc.farfield` (farfieldbc / Reimann) with a
subsonic and a supersonic free stream.  Run in the development container only:

    python tests/golden/make_bc_golden.py
"""
import os
import sys
import types

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import oracle  # noqa: E402

REF = "/root/reference/pyranda/pyrandaBC.py"
N = (20, 18, 1)


def load_reference_bc():
    pkg = types.ModuleType("refpkg")
    pkg.__path__ = []
    base = types.ModuleType("refpkg.pyrandaPackage")

    class pyrandaPackage:
        def __init__(self, name, pysim):
            self.name, self.pyranda = name, pysim
    base.pyrandaPackage = pyrandaPackage
    sys.modules["refpkg"], sys.modules["refpkg.pyrandaPackage"] = pkg, base
    mod = types.ModuleType("refpkg.pyrandaBC")
    mod.__package__ = "refpkg"
    exec(compile(open(REF).read(), REF, "exec"), mod.__dict__)
    return mod.pyrandaBC


def mesh():
    xs = [np.linspace(0, 1, k) for k in N]
    X, Y, Z = np.meshgrid(*xs, indexing="ij")
    Xd = X + 0.06 * np.sin(2 * np.pi * Y) * np.sin(np.pi * X)
    Yd = Y + 0.05 * np.sin(2 * np.pi * X) * (1 + 0.3 * Z)
    Zd = Z * (1 + 0.1 * X)
    return X, Y, Z, Xd, Yd, Zd


def fields(X, Y, Z):
    rng = np.random.default_rng(11)
    mk = lambda a: np.asfortranarray(a + 0.2 * rng.uniform(-1, 1, size=X.shape))
    return {"u": mk(np.sin(3 * X) * np.cos(2 * Y)), "v": mk(np.cos(4 * X + Y)), "w": mk(0.3 * np.sin(5 * Z + X)),
            "rho": mk(1.0 + 0.3 * np.cos(5 * X) * np.sin(3 * Y)), "p": mk(1.0 + 0.3 * np.sin(4 * X - Y))}


FARFIELD = {"yn": {"rho0": 1.0, "p0": 1.0, "u0": 0.4, "v0": 0.1, "w0": 0.0, "gamma": 1.4},
            "x1": {"rho0": 0.9, "p0": 0.2, "u0": 1.5, "v0": -0.2, "w0": 0.1, "gamma": 1.4}}


def main():
    X, Y, Z, Xd, Yd, Zd = mesh()
    o = oracle.Oracle(*N, 0, 1, 0, 1, 0, 1, coordsys=3, mesh_xyz=(Xd, Yd, Zd))

    class Var:
        def __init__(self, a): self.data = a

    class MPI:
        ax, ay, az = N
        x1proc = xnproc = y1proc = ynproc = z1proc = znproc = True

    class Sim:
        PyMPI = MPI()
        def __init__(self, f): self.variables = {k: Var(a.copy(order="F")) for k, a in f.items()}
        def getVar(self, name): return o.getvar(name)

    out = {}
    f0 = fields(X, Y, Z)
    sim = Sim(f0)
    bc = load_reference_bc()(sim)
    bc.exitbc(["rho", "w"], ["x1", "xn", "y1", "yn"])
    bc.exitbc("u", ["x1", "yn"], norm=True)
    for k in ("rho", "w", "u"):
        out["exit_" + k] = sim.variables[k].data
    sim = Sim(f0)
    bc = load_reference_bc()(sim)
    bc.slipbc([["u", "v"]], ["x1", "yn"])
    bc.slipbc([["u", "v", "w"]], ["xn", "y1"])
    for k in ("u", "v", "w"):
        out["slip_" + k] = sim.variables[k].data
    sim = Sim(f0)
    sim.variables["u"].data[1, :, :] *= 3.0  # part of the x1 face supersonic
    bc = load_reference_bc()(sim)
    for d, ref in FARFIELD.items():
        bc.BCdata["farfield-properties-%s" % d] = dict(ref, rho="rho", u="u", v="v", w="w", p="p")
    bc.farfieldbc(["yn", "x1"])
    for k in ("rho", "u", "v", "w", "p"):
        out["far_" + k] = sim.variables[k].data
    np.savez_compressed(os.path.join(ROOT, "tests", "golden", "bc_20x18.npz"), **out)
    print({k: float(np.abs(a - f0[k.split("_")[1]]).max()) for k, a in out.items()})


if __name__ == "__main__":
    main()
