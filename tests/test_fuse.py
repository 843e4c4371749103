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


def test_equation_groups_match_sequential_evaluation():
    """Consecutive pure-arithmetic equations become one multi-output kernel: the group (evaluated
    through its Python fallback here) must equal the sequential interpreter, and NVRTC must accept
    the generated multi-output kernels."""
    import re
    from pyranda_b200.fuse import group_kernel_source
    from pyranda_b200.sim import _Equation
    rng = np.random.default_rng(1)
    fz = Fuser(_NS, enabled=False)
    eqs = [_Equation(ln, fz) for ln in _lines(TGV_EOM)]
    names = set()
    for ln in _lines(TGV_EOM):
        names.update(re.findall(r":([A-Za-z_]\w*):", ln))
    base = {nm: rng.uniform(0.5, 1.5, size=(3, 4, 2)) for nm in names}
    base["gamma"] = 1.4
    pure = [e for e in eqs if e.kind == "ALG" and e.pure is not None]
    assert [e.lhs[0] for e in pure][:4] == ["u", "v", "w", "p"]
    # sequential reference
    sim = _Sim(rng)
    sim.variables = dict(base)
    ns = {"xp": _NS, "numpy": _NS, "self": sim, "__fz": fz.call}
    for e in pure:
        sim.variables[e.lhs[0]] = eval(e.code, ns)
    # one group over the same equations
    gid = fz.make_group([(e.lhs[0], e.pure) for e in pure])
    got = dict(base)
    fz.run_group(gid, got)
    for e in pure:
        assert np.array_equal(np.asarray(got[e.lhs[0]]), np.asarray(sim.variables[e.lhs[0]]), equal_nan=True), e.text
    pytest.importorskip("cuda.bindings.nvrtc")
    from pyranda_b200.fuse import _Nvrtc
    g = fz.groups[gid]
    mask = [nm != "gamma" for nm in g["inputs"]]
    assert len(_Nvrtc().compile_to_cubin(group_kernel_source(g["stmts"], mask))) > 100


def test_hoisted_flux_arguments_and_stage_split():
    """fuse.hoist: the operator arguments of all PDE lines in one multi-output group, the lines
    rewritten to read them; evaluated through the fallback path the fluxes equal the plain ones."""
    rng = np.random.default_rng(1)
    sim = _Sim(rng)
    sim._hoisted = {}
    fz = Fuser(_NS, enabled=False)
    import re
    names = set()
    for ln in _lines(TGV_EOM):
        names.update(re.findall(r":([A-Za-z_]\w*):", ln))
    for nm in names:
        sim.variables[nm] = rng.uniform(0.5, 1.5, size=(4, 3, 2))
    ns = {"xp": _NS, "numpy": _NS, "self": sim, "__fz": fz.call}
    pde = [translate(ln.split("=", 1)[1]) for ln in _lines(TGV_EOM) if "ddt(" in ln]
    gid, srcs = fz.hoist(pde)
    assert gid is not None and len(fz.groups[gid]["outs"]) == 15   # 3 directions x 5 equations, all distinct
    assert len(fz.groups[gid]["inputs"]) == 14                      # rho u v w rhou rhov rhow Et + 6 stresses, each read once
    fz.run_group(gid, sim.variables, out=sim._hoisted)
    for src, new in zip(pde, srcs):
        assert "self._hoisted" in new
        fsrc = fz.transform(new)
        split = fz.split_stage(fsrc)
        assert split is not None                                    # -ddx(.) - ddy(.) - ddz(.): arithmetic of three operator results
        leaves = eval(split[1], ns)
        assert len(leaves) == 3
        assert np.array_equal(eval(src, ns), fz.call(split[0], *leaves))
        assert np.array_equal(eval(src, ns), eval(fsrc, ns))


def test_host_only_lines_are_found():
    """The `:dt:` chain, `:cs:` and the diagnostics of the Taylor-Green deck are needed by the host only."""
    from pyranda_b200.sim import pyrandaSim, _Equation
    eqs = [_Equation(ln) for ln in _lines(TGV_EOM)]

    class S:
        equations = eqs
        conserved = [e.lhs[0] for e in eqs if e.kind == "PDE"]
    dead = pyrandaSim._host_only(S)
    assert dead == {"dt", "cs", "enst", "tke"}, dead


def test_where_with_a_comparison_is_one_fused_select():
    """xp.where(a < b, x, y) (the immersed-boundary package's masked algebra, pyranda_b200/ibm.py) becomes one
    generated kernel with a select; the fallback evaluation equals the plain one and NVRTC takes the source."""
    rng = np.random.default_rng(1)
    sdf, val, tx, g0, lens = (rng.standard_normal((6, 5, 4)) for _ in range(5))
    fz = Fuser(_NS, enabled=False)
    for src in ("xp.where(sdf <= epsi, val + cfl * gl * (tx * g0), val)", "c - xp.where(sdf < lens, normal, 0.0) * a",
                "xp.where(sdf < lens, 0.0, normal / sdf)", "xp.where(sdf >= 0.5, xp.sqrt(xp.abs(val)), 0.0)"):
        loc = dict(sdf=sdf, epsi=0.0, val=val, cfl=0.5, gl=np.abs(g0), tx=tx, g0=g0, lens=lens, c=val, normal=tx, a=g0)
        out = fz.transform(src)
        assert out.startswith("__fz(") and "where" not in out, out   # the whole expression is one fused call
        got = eval(out, {"xp": _NS, "__fz": fz.call}, loc)
        assert np.array_equal(got, eval(src, {"xp": _NS}, loc))
    spec = fz.specs[0]
    assert "?" in spec.cexpr
    try:
        from pyranda_b200.fuse import _Nvrtc
        rt = _Nvrtc()
    except Exception:
        pytest.skip("NVRTC bindings not importable here")
    for spec in fz.specs:
        rt.compile_to_cubin(kernel_source(spec.cexpr, [True] * spec.nleaves))
