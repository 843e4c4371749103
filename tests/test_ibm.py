"""The immersed-boundary package of the interpreter (pyranda_b200/ibm.py) against golden vectors
produced by the reference's own pyrandaIBM.py (tests/golden/make_ibm_golden.py), on numpy arrays and
on torch tensors (the code path the CUDA backend takes), and through deck lines."""
import os

import numpy as np
import pytest

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "ibm_24x24.npz")


def _sim(oracle_mod):
    from oracle_backend import make_sim
    mesh = "\n".join(["xdom = (0.0, 1.0, 24)", "ydom = (0.0, 1.0, 24)", "zdom = (0.0, 1.0, 1)"])
    return make_sim(oracle_mod, "ibm", mesh)


def _check(got, want, tol=1e-13):
    assert np.abs(np.asarray(got) - want).max() <= tol * max(1.0, np.abs(want).max())


def test_matches_reference_package(oracle_mod):
    g = np.load(GOLD)
    ss = _sim(oracle_mod)
    phi = np.asfortranarray(g["phi"])
    gphi = list(ss.grad(phi))
    u, v, w = np.asfortranarray(g["u"]), np.asfortranarray(g["v"]), 0.0  # the third component as a number
    with np.errstate(divide="ignore", invalid="ignore"):
        _check(ss.ibm.scalar(np.asfortranarray(g["rho"]), phi, gphi), g["ibmS"])
        for k, c in enumerate(ss.ibm.velocity_slip([u, v, w], phi, gphi)):
            _check(c, g["ibmV%d" % k])
        frame = [np.asfortranarray(g["fu"]), np.asfortranarray(g["fv"]), np.zeros_like(phi)]
        for k, c in enumerate(ss.ibm.velocity_slip([u, v, w], phi, gphi, phivar=frame)):
            _check(c, g["ibmVf%d" % k])
        for k, c in enumerate(ss.ibm.velocity_wall([u, v, w], phi, gphi)):
            _check(c, g["ibmW%d" % k])


def test_torch_namespace_takes_the_same_path(oracle_mod):
    """The CUDA backend evaluates the package on tensors through _TorchNS: same numbers on CPU tensors."""
    import torch
    from pyranda_b200.ibm import ImmersedBoundary
    from pyranda_b200.sim import _TorchNS
    g = np.load(GOLD)
    ref = _sim(oracle_mod)

    def T(a):
        return torch.from_numpy(np.array(a, order="C", copy=True))

    class Backend:
        def isfield(self, a): return isinstance(a, torch.Tensor)

    class Sim:
        xp = _TorchNS(torch)
        GridLen = T(ref.GridLen)
        B = Backend()
        def grad(self, v): return [T(c) for c in ref.grad(np.asfortranarray(v.numpy()))]
        def gfilter(self, v): return T(ref.gfilter(np.asfortranarray(v.numpy())))
        def emptyScalar(self, val): return torch.full_like(self.GridLen, float(val))
    ibm = ImmersedBoundary(Sim())
    phi = T(g["phi"])
    gphi = [T(c) for c in ref.grad(np.asfortranarray(g["phi"]))]
    _check(ibm.scalar(T(g["rho"]), phi, gphi).numpy(), g["ibmS"])
    for k, c in enumerate(ibm.velocity_slip([T(g["u"]), T(g["v"]), 0.0], phi, gphi)):
        _check(c.numpy(), g["ibmV%d" % k])
    for k, c in enumerate(ibm.velocity_wall([T(g["u"]), T(g["v"]), 0.0], phi, gphi)):
        _check(c.numpy(), g["ibmW%d" % k])


def test_deck_lines(oracle_mod):
    """examples/cylinder.py style: gradient of the level set once, ibmV / ibmS inside updateVars."""
    g = np.load(GOLD)
    ss = _sim(oracle_mod)
    ss.EOM("""
[:gx:,:gy:,:gz:] = grad(:phi:)
[:u:,:v:,:w:]    = ibmV( [:u:,:v:,0.0], :phi:, [:gx:,:gy:,:gz:] )
:rho:            = ibmS( :rho:, :phi:, [:gx:,:gy:,:gz:] )
""")
    for nm in ("phi", "u", "v", "rho"):
        ss.variables[nm] = np.asfortranarray(g[nm])
    with np.errstate(divide="ignore", invalid="ignore"):
        ss.updateVars()
    _check(ss.variables["u"], g["ibmV0"])
    _check(ss.variables["v"], g["ibmV1"])
    _check(ss.variables["rho"], g["ibmS"])
