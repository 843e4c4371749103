"""pyranda_b200 -- B200-native replacement for the hot path of LLNL/pyranda's `parcop` module.

Only what the compact-operator path needs lives here: the CUDA sources and C ABI (csrc/,
include/parcop_b200.h), and the host-side mirror of the reference's Python interface to them
(parcop.py == the f2py module surface, pyrandaMPI.py == der / fil / gfil dispatch, sim.py == the
RK4 stage loop on device-resident fields).
"""
from ._lib import ParcopError, load  # noqa: F401
from .plan import ParcopPlan  # noqa: F401

__all__ = ["ParcopPlan", "ParcopError", "load"]
