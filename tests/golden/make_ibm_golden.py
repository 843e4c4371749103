"""Generates tests/golden/ibm_24x24.npz by running the REFERENCE's own immersed-boundary package
(/root/reference/pyranda/pyrandaIBM.py, loaded with a stub for its package base class, because
`import pyranda` needs mpi4py and the compiled parcop module) on a small 2-D grid, with `grad` and
`gfilter` served by the CPU oracle.  Run in the development container only:

    python tests/golden/make_ibm_golden.py
"""
import os
import sys
import types

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import oracle  # noqa: E402

REF = "/root/reference/pyranda/pyrandaIBM.py"


def load_reference_ibm():
    pkg = types.ModuleType("refpkg")
    pkg.__path__ = []
    base = types.ModuleType("refpkg.pyrandaPackage")

    class pyrandaPackage:  # the two attributes the IBM package uses
        def __init__(self, name, pysim):
            self.name, self.pyranda = name, pysim
    base.pyrandaPackage = pyrandaPackage
    sys.modules["refpkg"], sys.modules["refpkg.pyrandaPackage"] = pkg, base
    mod = types.ModuleType("refpkg.pyrandaIBM")
    mod.__package__ = "refpkg"
    exec(compile(open(REF).read(), REF, "exec"), mod.__dict__)
    return mod.pyrandaIBM


def main():
    n = (24, 24, 1)
    o = oracle.Oracle(*n, 0.0, 1.0, 0.0, 1.0, 0.0, 1.0, periodic=(False, False, False))
    x, y = o.getvar("x"), o.getvar("y")

    class Mesh:
        GridLen = o.getvar("GridLen")

    class Sim:
        mesh = Mesh()
        def grad(self, v): return list(o.grads(np.asfortranarray(v)))
        def gfilter(self, v): return o.gfilter(np.asfortranarray(v))
    ibm = load_reference_ibm()(Sim())
    phi = np.asfortranarray(np.sqrt((x - 0.5) ** 2 + (y - 0.45) ** 2) - 0.22)  # a cylinder
    gphi = [np.asfortranarray(g) for g in o.grads(phi)]
    u = np.asfortranarray(1.0 + 0.3 * np.sin(3 * x) * np.cos(2 * y))
    v = np.asfortranarray(0.2 * np.cos(4 * x + y))
    w = np.asfortranarray(0.0 * x)
    rho = np.asfortranarray(1.0 + 0.1 * np.cos(5 * x) * np.sin(3 * y))
    frame = [np.asfortranarray(0.05 + 0.0 * x), np.asfortranarray(-0.02 + 0.0 * x), w]
    with np.errstate(divide="ignore", invalid="ignore"):
        out = {"ibmS": ibm.ibmS(rho, phi, gphi)}
        for k, c in enumerate(ibm.ibmVel([u, v, w], phi, gphi)):
            out["ibmV%d" % k] = c
        for k, c in enumerate(ibm.ibmVel([u, v, w], phi, gphi, phivar=frame)):
            out["ibmVf%d" % k] = c
        for k, c in enumerate(ibm.ibmWall([u, v, w], phi, gphi)):
            out["ibmW%d" % k] = c
    np.savez_compressed(os.path.join(ROOT, "tests", "golden", "ibm_24x24.npz"), phi=phi, u=u, v=v, rho=rho,
                        fu=frame[0], fv=frame[1], **out)
    print({k: float(np.abs(a).max()) for k, a in out.items()})


if __name__ == "__main__":
    main()
