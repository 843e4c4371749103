"""Packs the golden curves of the REFERENCE's own regression suite (tests/baselines/*.dat of
LLNL/pyranda, compared there at a relative tolerance of 1e-4 by tests/run_tests.py) into
tests/golden/reference_baselines.npz, so the oracle pins of tests/test_sim_oracle.py do not need the
reference checkout at test time.  Data only (two rows of numbers per curve), float64, unchanged.

key                      reference file / producing command (tests/cases/*.py)
RT_2D                    RT_2D.dat                      examples/RT3D.py 32 1 1
cylinder-2d-32 / -64     cylinder-2d-32.dat / -64.dat   examples/cylinder.py N 1 name
cylinder_curved-2d-64    cylinder_curved-2d-64.dat      examples/cylinder_curv.py 64 1 name
cylinder_omesh-2d-64     cylinder_omesh-2d-64.dat       examples/cylinder_curv2.py 64 1 name
euler-2d-64 / -128       euler-2d-64.dat / -128.dat     examples/euler.py N 1 name
KH-2d-64                 KelvinHelmholtzKH-2d-64.dat    examples/KH.py 64 0 KH-2d-64 1

Run in the development container only:   python tests/golden/make_baseline_fixtures.py
"""
import os

import numpy as np

REF = "/root/reference/tests/baselines"
FILES = {"RT_2D": "RT_2D.dat", "cylinder-2d-32": "cylinder-2d-32.dat", "cylinder-2d-64": "cylinder-2d-64.dat",
         "cylinder_curved-2d-64": "cylinder_curved-2d-64.dat", "cylinder_omesh-2d-64": "cylinder_omesh-2d-64.dat",
         "euler-2d-64": "euler-2d-64.dat", "euler-2d-128": "euler-2d-128.dat", "KH-2d-64": "KelvinHelmholtzKH-2d-64.dat"}

if __name__ == "__main__":
    out = {k: np.loadtxt(os.path.join(REF, f)) for k, f in FILES.items()}
    np.savez_compressed(os.path.join(os.path.dirname(os.path.abspath(__file__)), "reference_baselines.npz"), **out)
    print({k: v.shape for k, v in out.items()})
