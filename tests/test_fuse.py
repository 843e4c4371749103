"""Expression fusion of the device EOM interpreter (pyranda_b200/fuse.py): the source
transformation is checked on the CPU (fallback evaluation == plain evaluation, NVRTC accepts every
generated kernel); the kernels themselves run in the GPU tests."""
import numpy as np
import pytest

from decks import TGV_EOM
from pyranda_b200.fuse import Fuser, kernel_source
from pyranda_b200.sim import _lines, translate


class _NS:  # the numpy namespace the interpreter hands to equations
    sqrt, abs, sin, cos, tanh, exp = np.sqrt, np.abs, np.sin, np.cos, np.tanh, np.exp
    minimum, maximum, where, pi = np.minimum, np.maximum, np.where, np.pi


class _Sim:
    def __init__(self, rng):
        self.variables = {}
        self.rng = rng
        self.calls = 0

    def _op(self, *a):
        self.calls += 1
        return sum(np.asarray(x, dtype=float) for x in a) * 0.5 + 1.0

    ddx = ddy = ddz = ring = gfilter = filter = _op

    def dt_courant(self, *a):
        return 0.1

    def dt_diff(self, *a):
        return 0.2


def test_fused_source_matches_plain_evaluation():
    rng = np.random.default_rng(0)
    sim = _Sim(rng)
    fz = Fuser(_NS, enabled=False)  # fallback path: the transformed source must mean the same
    names = set()
    import re
    for ln in _lines(TGV_EOM):
        names.update(re.findall(r":([A-Za-z_]\w*):", ln))
    for nm in names:
        sim.variables[nm] = rng.uniform(0.5, 1.5, size=(4, 3, 2))
    sim.variables["gamma"] = 1.4
    ns = {"xp": _NS, "numpy": _NS, "self": sim, "__fz": fz.call}
    nfused = 0
    for ln in _lines(TGV_EOM):
        rhs = ln.split("=", 1)[1]
        src = translate(rhs)
        fsrc = fz.transform(src)
        nfused += fsrc.count("__fz(")
        a, b = eval(src, ns), eval(fsrc, ns)
        assert np.array_equal(np.asarray(a), np.asarray(b)), ln
    assert nfused >= 30


def test_generated_kernels_compile():
    nvrtc = pytest.importorskip("cuda.bindings.nvrtc")
    from pyranda_b200.fuse import _Nvrtc
    fz = Fuser(_NS)
    for ln in _lines(TGV_EOM):
        fz.transform(translate(ln.split("=", 1)[1]))
    rt = _Nvrtc()
    assert fz.specs
    for spec in fz.specs:
        for mask in ([True] * spec.nleaves, [True] + [False] * (spec.nleaves - 1)):
            cubin = rt.compile_to_cubin(kernel_source(spec.cexpr, mask))
            assert len(cubin) > 100
