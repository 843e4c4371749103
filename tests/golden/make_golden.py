"""Regenerate tests/golden/ops_16x16x16.npz from the CPU oracle.

The reference (Fortran + MPI + f2py) cannot be imported or built in the development image, so these
vectors come from the oracle, which is itself pinned to the reference's golden scalars
(tests/test_oracle_pins.py).  They travel to the GPU box, where the `-m gpu` tests also compare the
CUDA path against them without needing anything from /root/reference.
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
sys.path.insert(0, os.path.dirname(HERE))
from conftest import domain, synthetic_field  # noqa: E402
from oracle import oracle  # noqa: E402

n = (16, 16, 16)
out = {"n": np.array(n)}
for periodic in (True, False):
    tag = "per" if periodic else "bnd"
    (x1, xn), (y1, yn), (z1, zn) = domain(n, periodic)
    o = oracle.Oracle(*n, x1, xn, y1, yn, z1, zn, periodic=(periodic,) * 3)
    f = synthetic_field(o.getvar("x"), o.getvar("y"), o.getvar("z"))
    out["f_" + tag] = f
    for name in ("ddx", "ddy", "ddz", "sfilter", "gfilter", "pring", "plaplacian"):
        out["%s_%s" % (name, tag)] = getattr(o, name)(f)
np.savez_compressed(os.path.join(HERE, "ops_16x16x16.npz"), **out)
print("wrote", os.path.join(HERE, "ops_16x16x16.npz"))
