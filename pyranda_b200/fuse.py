"""Fused pointwise expressions for the device-resident EOM interpreter.

pyranda evaluates every equation string with numpy, one temporary per arithmetic operation
(pyranda.py:397-416, pyrandaEq.py).  On the GPU that is one kernel and one 8 B/point round trip per
operation, and once the compact operators run near the HBM roofline it dominates a Taylor-Green step.
Here the arithmetic between operator calls is compiled: the Python AST of a translated equation is
cut at everything that is not arithmetic (variables, operator calls such as self.ddx(...), reductions),
every maximal arithmetic subtree becomes ONE elementwise CUDA kernel generated as source, compiled
with NVRTC for sm_100a (--fmad=false: the same IEEE operations in the same order as the unfused
expression, so results are bit-identical to the torch evaluation) and launched on the current
stream.  Leaves that turn out to be Python scalars are passed by value; anything unexpected (mixed
shapes or strides, non-fp64, 0-dim tensors) falls back to the unfused evaluation.

This is plumbing of the host interpreter, not part of the operator library: without
NVRTC / cuda-python the interpreter simply runs unfused.
"""
import ast
import ctypes
import math

_BIN = {ast.Add: "+", ast.Sub: "-", ast.Mult: "*", ast.Div: "/"}
_CMP = {ast.Lt: "<", ast.LtE: "<=", ast.Gt: ">", ast.GtE: ">="}
_CFUN = {"sqrt": "sqrt", "abs": "fabs", "sin": "sin", "cos": "cos", "tanh": "tanh", "exp": "exp",
         "minimum": "fmin", "maximum": "fmax"}


def _is_num(node):
    return isinstance(node, ast.Constant) and isinstance(node.value, (int, float)) and not isinstance(node.value, bool)


def _is_arith(node):
    if isinstance(node, ast.BinOp):
        return type(node.op) in _BIN or isinstance(node.op, ast.Pow)
    if isinstance(node, ast.UnaryOp):
        return isinstance(node.op, (ast.USub, ast.UAdd))
    if isinstance(node, ast.Call) and isinstance(node.func, ast.Attribute) and isinstance(node.func.value, ast.Name):
        if node.func.value.id == "xp" and node.func.attr == "where" and not node.keywords:
            # xp.where(a < b, x, y) with one ordering comparison: a select in the generated kernel
            c = node.args[0] if len(node.args) == 3 else None
            return isinstance(c, ast.Compare) and len(c.ops) == 1 and type(c.ops[0]) in _CMP
        return node.func.value.id == "xp" and node.func.attr in _CFUN and not node.keywords
    return False


class _Spec:
    """One fused subtree: C expression over v0..vk, the same expression as a Python function."""

    def __init__(self, cexpr, pysrc, nleaves):
        self.cexpr, self.nleaves = cexpr, nleaves
        self.pysrc = pysrc
        self.fallback = None
        self.kernels = {}


class Fuser:
    def __init__(self, xp, enabled=True):
        self.xp = xp
        self.enabled = enabled
        self.specs = []
        self.groups = []
        self._rt = None
        self.launches = 0

    # ------------------------------------------------------------------ source transformation
    def transform(self, src):
        """Python expression source -> source in which every maximal arithmetic subtree is a call
        __fz(id, leaf0, leaf1, ...)."""
        tree = ast.parse(src, mode="eval")
        tree.body = self._rewrite(tree.body)
        ast.fix_missing_locations(tree)
        return ast.unparse(tree)

    def _rewrite(self, node):
        if _is_arith(node):
            leaves, keys = [], {}
            cexpr, pyexpr = self._emit(node, leaves, keys)
            if not leaves:  # constants only
                return node
            spec = _Spec(cexpr, "lambda xp, %s: %s" % (", ".join("L%d" % i for i in range(len(leaves))), pyexpr), len(leaves))
            self.specs.append(spec)
            return ast.Call(func=ast.Name(id="__fz", ctx=ast.Load()),
                            args=[ast.Constant(len(self.specs) - 1)] + leaves, keywords=[])
        for field, old in ast.iter_fields(node):
            if isinstance(old, list):
                setattr(node, field, [self._rewrite(o) if isinstance(o, ast.AST) else o for o in old])
            elif isinstance(old, ast.AST):
                setattr(node, field, self._rewrite(old))
        return node

    def _emit(self, node, leaves, keys):
        """(C expression, Python expression) of an arithmetic subtree; leaves are collected in order."""
        if _is_num(node):
            v = float(node.value)
            return "(%s)" % repr(v), repr(node.value)
        if not _is_arith(node):
            new = self._rewrite(node)  # operator calls inside keep their own fused arguments
            key = ast.dump(new)
            if key not in keys:
                keys[key] = len(leaves)
                leaves.append(new)
            return "v%d" % keys[key], "L%d" % keys[key]
        if isinstance(node, ast.UnaryOp):
            c, p = self._emit(node.operand, leaves, keys)
            s = "-" if isinstance(node.op, ast.USub) else "+"
            return "(%s%s)" % (s, c), "(%s%s)" % (s, p)
        if isinstance(node, ast.Call) and node.func.attr == "where":
            cmp_ = node.args[0]
            (lc, lp), (rc, rp) = self._emit(cmp_.left, leaves, keys), self._emit(cmp_.comparators[0], leaves, keys)
            (ac, ap), (bc, bp) = self._emit(node.args[1], leaves, keys), self._emit(node.args[2], leaves, keys)
            op = _CMP[type(cmp_.ops[0])]
            return ("((%s %s %s) ? %s : %s)" % (lc, op, rc, ac, bc), "xp.where((%s %s %s), %s, %s)" % (lp, op, rp, ap, bp))
        if isinstance(node, ast.Call):
            parts = [self._emit(a, leaves, keys) for a in node.args]
            return ("%s(%s)" % (_CFUN[node.func.attr], ", ".join(c for c, _ in parts)),
                    "xp.%s(%s)" % (node.func.attr, ", ".join(p for _, p in parts)))
        lc, lp = self._emit(node.left, leaves, keys)
        rc, rp = self._emit(node.right, leaves, keys)
        if isinstance(node.op, ast.Pow):
            if _is_num(node.right) and float(node.right.value) == 2.0:
                return "(%s * %s)" % (lc, lc), "(%s ** 2)" % lp  # what numpy and torch do for x**2
            if _is_num(node.right) and float(node.right.value) == 0.5:
                return "sqrt(%s)" % lc, "(%s ** 0.5)" % lp
            return "pow(%s, %s)" % (lc, rc), "(%s ** %s)" % (lp, rp)
        op = _BIN[type(node.op)]
        return "(%s %s %s)" % (lc, op, rc), "(%s %s %s)" % (lp, op, rp)

    # ------------------------------------------------------------------ groups of equations
    @staticmethod
    def _var_name(node):
        """'x' for the node self.variables["x"], else None."""
        if (isinstance(node, ast.Subscript) and isinstance(node.value, ast.Attribute) and node.value.attr == "variables"
                and isinstance(node.value.value, ast.Name) and node.value.value.id == "self"):
            sl = node.slice
            if isinstance(sl, ast.Constant) and isinstance(sl.value, str):
                return sl.value
        return None

    def pure_tree(self, src):
        """The AST of `src` when it is arithmetic over variables and numbers only, else None."""
        tree = ast.parse(src, mode="eval").body

        def ok(n):
            if _is_num(n) or self._var_name(n) is not None:
                return True
            if not _is_arith(n):
                return False
            kids = n.args if isinstance(n, ast.Call) else ([n.operand] if isinstance(n, ast.UnaryOp) else [n.left, n.right])
            return all(ok(k) for k in kids)
        return tree if (_is_arith(tree) and ok(tree)) else None

    def make_group(self, assignments):
        """assignments: [(lhs name, pure AST)] evaluated in order -> one kernel with several outputs.
        Inputs are the variables read before the group assigns them; later equations see the values
        earlier ones produced (kept in registers), exactly like the sequential interpreter."""
        inputs, produced, stmts = [], {}, []

        def emit(n):
            if _is_num(n):
                return "(%s)" % repr(float(n.value)), repr(n.value)
            nm = self._var_name(n)
            if nm is not None:
                if nm in produced:
                    return "o%d" % produced[nm], "O[%d]" % produced[nm]
                if nm not in inputs:
                    inputs.append(nm)
                return "v%d" % inputs.index(nm), "V[%d]" % inputs.index(nm)
            if isinstance(n, ast.UnaryOp):
                c, p = emit(n.operand)
                s = "-" if isinstance(n.op, ast.USub) else "+"
                return "(%s%s)" % (s, c), "(%s%s)" % (s, p)
            if isinstance(n, ast.Call):
                parts = [emit(a) for a in n.args]
                return ("%s(%s)" % (_CFUN[n.func.attr], ", ".join(c for c, _ in parts)),
                        "xp.%s(%s)" % (n.func.attr, ", ".join(p for _, p in parts)))
            lc, lp = emit(n.left)
            rc, rp = emit(n.right)
            if isinstance(n.op, ast.Pow):
                if _is_num(n.right) and float(n.right.value) == 2.0:
                    return "(%s * %s)" % (lc, lc), "(%s ** 2)" % lp
                if _is_num(n.right) and float(n.right.value) == 0.5:
                    return "sqrt(%s)" % lc, "(%s ** 0.5)" % lp
                return "pow(%s, %s)" % (lc, rc), "(%s ** %s)" % (lp, rp)
            op = _BIN[type(n.op)]
            return "(%s %s %s)" % (lc, op, rc), "(%s %s %s)" % (lp, op, rp)

        outs = []
        for name, tree in assignments:
            c, p = emit(tree)
            stmts.append((c, p))
            produced[name] = len(outs)  # a variable assigned twice: later readers see the latest value
            outs.append(name)
        self.groups.append({"inputs": inputs, "outs": outs, "stmts": stmts, "kernels": {}, "py": None})
        return len(self.groups) - 1

    def run_group(self, gid, variables, out=None):
        """Evaluates the group into `out` (default: `variables`), fused when every input is a field or a number."""
        g = self.groups[gid]
        src_vars = variables
        if out is not None:
            variables = out
        vals = [src_vars[nm] for nm in g["inputs"]]
        kern = self._group_kernel(g, vals) if self.enabled else None
        if kern is None:
            if g["py"] is None:
                g["py"] = [eval("lambda xp, V, O: " + p) for _, p in g["stmts"]]
            O = []
            for name, fn in zip(g["outs"], g["py"]):
                O.append(fn(self.xp, vals, O))
                variables[name] = O[-1]
            return
        import torch
        fn, mask, ref = kern
        n = ref.numel()
        outs = [torch.empty_strided(ref.shape, ref.stride(), dtype=torch.float64, device=ref.device) for _ in g["outs"]]
        args = [n] + [v.data_ptr() if m else float(v) for v, m in zip(vals, mask)] + [o.data_ptr() for o in outs]
        types = ([ctypes.c_long] + [ctypes.c_void_p if m else ctypes.c_double for m in mask] + [ctypes.c_void_p] * len(outs))
        self._runtime().launch(fn, n, args, types, torch.cuda.current_stream().cuda_stream)
        self.launches += 1
        for name, o in zip(g["outs"], outs):
            variables[name] = o

    def _group_kernel(self, g, vals):
        import torch
        mask, ref = [], None
        for v in vals:
            if isinstance(v, torch.Tensor):
                if v.dim() == 0 or v.dtype != torch.float64 or not v.is_cuda:
                    return None
                if ref is None:
                    ref = v
                elif v.shape != ref.shape or v.stride() != ref.stride():
                    return None
                mask.append(True)
            elif isinstance(v, (int, float)) and not isinstance(v, bool):
                mask.append(False)
            else:
                return None
        if ref is None or not _dense(ref):
            return None
        # every output must depend on at least one field, or the interpreter would have produced a number
        dep = []
        for c, _ in g["stmts"]:
            uses_field = any(("v%d" % i) in _tokens(c) for i, m in enumerate(mask) if m) or any(
                ("o%d" % j) in _tokens(c) and dep[j] for j in range(len(dep)))
            dep.append(uses_field)
        if not all(dep):
            return None
        key = tuple(mask)
        try:
            if key not in g["kernels"]:
                g["kernels"][key] = self._runtime().compile(group_kernel_source(g["stmts"], mask))
        except Exception:
            self.enabled = False
            return None
        return g["kernels"][key], mask, ref

    # ------------------------------------------------------------------ hoisted operator arguments
    OPERATORS = ("ddx", "ddy", "ddz", "div", "divT", "grad", "filter", "gfilter", "gfilterx", "gfiltery", "gfilterz", "ring",
                 "ringV", "laplacian", "dd4x", "dd4y", "dd4z", "dd8x", "dd8y", "dd8z")

    def _is_operator_call(self, node):
        return (isinstance(node, ast.Call) and isinstance(node.func, ast.Attribute) and isinstance(node.func.value, ast.Name)
                and node.func.value.id == "self" and node.func.attr in self.OPERATORS and not node.keywords)

    def _pure(self, n):
        if _is_num(n) or self._var_name(n) is not None:
            return True
        if not _is_arith(n):
            return False
        kids = n.args if isinstance(n, ast.Call) else ([n.operand] if isinstance(n, ast.UnaryOp) else [n.left, n.right])
        return all(self._pure(k) for k in kids)

    def hoist(self, sources, holder="self._hoisted"):
        """The flux equations of a deck call operators on arithmetic of the variables:
        -ddx(:rhou:*:u: - :tauxx:) - ddy(...) ...  Every such argument (over ALL the given sources) moves
        into ONE multi-output kernel that reads each variable once; the sources come back with the
        arguments replaced by `holder["hK"]`.  Returns (group id or None, new sources).  Valid because
        nothing assigns a variable while the right-hand sides of the PDE lines are evaluated."""
        trees = [ast.parse(src, mode="eval") for src in sources]
        found, order = {}, []

        def visit(node):
            for field, old in ast.iter_fields(node):
                kids = old if isinstance(old, list) else [old]
                new = []
                for k in kids:
                    if isinstance(k, ast.AST):
                        visit(k)
                        if self._is_operator_call(node) and field == "args" and _is_arith(k) and self._pure(k):
                            key = ast.dump(k)
                            if key not in found:
                                found[key] = "h%d" % len(order)
                                order.append((found[key], k))
                            k = ast.Subscript(value=ast.parse(holder, mode="eval").body, slice=ast.Constant(found[key]), ctx=ast.Load())
                    new.append(k)
                setattr(node, field, new if isinstance(old, list) else new[0])

        for t in trees:
            visit(t)
        if len(order) < 2:
            return None, list(sources)
        for t in trees:
            ast.fix_missing_locations(t)
        return self.make_group(order), [ast.unparse(t) for t in trees]

    # ------------------------------------------------------------------ RK4 stage with the flux expression inside
    @staticmethod
    def split_stage(src):
        """`__fz(k, leaf0, leaf1, ...)` at the top of a transformed right-hand side -> (k, source of the
        tuple of leaves), else None: the stage kernel can then evaluate the flux expression itself."""
        body = ast.parse(src, mode="eval").body
        if (isinstance(body, ast.Call) and isinstance(body.func, ast.Name) and body.func.id == "__fz" and body.args
                and isinstance(body.args[0], ast.Constant) and not body.keywords):
            tup = ast.Tuple(elts=list(body.args[1:]), ctx=ast.Load())
            return int(body.args[0].value), ast.unparse(ast.fix_missing_locations(ast.Expression(body=tup)))
        return None

    def stage(self, idx, vals, dt, A, B, PHI, U):
        """PHI = dt*F + A*PHI ; U = U + B*PHI (pyranda.py:800-804) with F = spec idx of `vals` evaluated in
        the same kernel (no flux field is written or read back).  Returns False when the operands do
        not qualify; the caller then evaluates F and runs the library's stage kernel."""
        if not self.enabled:
            return False
        import torch
        spec = self.specs[idx]
        mask = []
        for v in vals:
            if isinstance(v, torch.Tensor):
                if v.dim() == 0 or v.dtype != torch.float64 or not v.is_cuda or v.shape != U.shape or v.stride() != U.stride():
                    return False
                mask.append(True)
            elif isinstance(v, (int, float)) and not isinstance(v, bool):
                mask.append(False)
            else:
                return False
        if not any(mask) or PHI.stride() != U.stride() or PHI.shape != U.shape or not _dense(U):
            return False
        try:
            rt = self._runtime()
            key = ("stage",) + tuple(mask)
            kern = spec.kernels.get(key)
            if kern is None:
                kern = spec.kernels[key] = rt.compile(stage_kernel_source(spec.cexpr, mask))
        except Exception:
            self.enabled = False
            return False
        n = U.numel()
        args = [n] + [v.data_ptr() if m else float(v) for v, m in zip(vals, mask)] + [float(dt), float(A), float(B), PHI.data_ptr(), U.data_ptr()]
        types = ([ctypes.c_long] + [ctypes.c_void_p if m else ctypes.c_double for m in mask]
                 + [ctypes.c_double] * 3 + [ctypes.c_void_p] * 2)
        rt.launch(kern, n, args, types, torch.cuda.current_stream().cuda_stream)
        self.launches += 1
        return True

    # ------------------------------------------------------------------ run time
    def call(self, idx, *vals):
        spec = self.specs[idx]
        if spec.fallback is None:
            spec.fallback = eval(spec.pysrc)
        if not self.enabled:
            return spec.fallback(self.xp, *vals)
        import torch
        mask, ref = [], None
        for v in vals:
            if isinstance(v, torch.Tensor):
                if v.dim() == 0 or v.dtype != torch.float64 or not v.is_cuda:
                    return spec.fallback(self.xp, *vals)
                if ref is None:
                    ref = v
                elif v.shape != ref.shape or v.stride() != ref.stride():
                    return spec.fallback(self.xp, *vals)
                mask.append(True)
            elif isinstance(v, (int, float)) and not isinstance(v, bool):
                mask.append(False)
            else:
                return spec.fallback(self.xp, *vals)
        if ref is None:
            return spec.fallback(self.xp, *vals)
        n = ref.numel()
        if not _dense(ref):
            return spec.fallback(self.xp, *vals)
        try:
            rt = self._runtime()
            kern = spec.kernels.get(tuple(mask))
            if kern is None:
                kern = spec.kernels[tuple(mask)] = rt.compile(kernel_source(spec.cexpr, mask))
        except Exception:
            self.enabled = False  # no NVRTC / driver bindings: run unfused from now on
            return spec.fallback(self.xp, *vals)
        out = torch.empty_strided(ref.shape, ref.stride(), dtype=torch.float64, device=ref.device)
        args = [n] + [v.data_ptr() if m else float(v) for v, m in zip(vals, mask)] + [out.data_ptr()]
        types = [ctypes.c_long] + [ctypes.c_void_p if m else ctypes.c_double for m in mask] + [ctypes.c_void_p]
        rt.launch(kern, n, args, types, torch.cuda.current_stream().cuda_stream)
        self.launches += 1
        return out

    def _runtime(self):
        if self._rt is None:
            self._rt = _Nvrtc()
        return self._rt


def _dense(ref):
    """Dense storage in any axis order: the kernels walk the flat storage."""
    run = 1
    for s_, e_ in sorted(zip(ref.stride(), ref.shape)):
        if e_ != 1 and s_ != run:
            return False
        run *= e_
    return True


def _tokens(cexpr):
    import re
    return set(re.findall(r"[vo]\d+", cexpr))


def group_kernel_source(stmts, mask):
    params = ["long n"]
    loads = []
    for i, m in enumerate(mask):
        if m:
            params.append("const double *__restrict__ a%d" % i)
            loads.append("    const double v%d = a%d[t];" % (i, i))
        else:
            params.append("double v%d" % i)
    body = []
    for j, (c, _) in enumerate(stmts):
        params.append("double *__restrict__ out%d" % j)
        body.append("    const double o%d = %s;\n    out%d[t] = o%d;" % (j, c, j, j))
    return ("extern \"C\" __global__ void __launch_bounds__(256) fz(%s) {\n"
            "  for (long t = blockIdx.x * (long)blockDim.x + threadIdx.x; t < n; t += (long)gridDim.x * blockDim.x) {\n"
            "%s\n%s\n  }\n}\n" % (", ".join(params), "\n".join(loads), "\n".join(body)))


def stage_kernel_source(cexpr, mask):
    """The RK4 stage update with the flux expression evaluated in place (same operations and order as
    the flux kernel followed by rk4_stage_kernel: --fmad=false keeps dt*F + tmp1 two roundings)."""
    params = ["long n"]
    loads = []
    for i, m in enumerate(mask):
        if m:
            params.append("const double *__restrict__ a%d" % i)
            loads.append("    const double v%d = a%d[t];" % (i, i))
        else:
            params.append("double v%d" % i)
    params += ["double dt", "double A", "double B", "double *__restrict__ PHI", "double *__restrict__ U"]
    return ("extern \"C\" __global__ void __launch_bounds__(256) fz(%s) {\n"
            "  for (long t = blockIdx.x * (long)blockDim.x + threadIdx.x; t < n; t += (long)gridDim.x * blockDim.x) {\n"
            "%s\n    const double F = %s;\n    const double tmp1 = A * PHI[t];\n    const double phi = dt * F + tmp1;\n"
            "    PHI[t] = phi;\n    const double tmp2 = B * phi;\n    U[t] = U[t] + tmp2;\n  }\n}\n"
            % (", ".join(params), "\n".join(loads), cexpr))


def kernel_source(cexpr, mask):
    params = ["long n"]
    loads = []
    for i, m in enumerate(mask):
        if m:
            params.append("const double *__restrict__ a%d" % i)
            loads.append("    const double v%d = a%d[t];" % (i, i))
        else:
            params.append("double v%d" % i)
    params.append("double *__restrict__ out")
    return ("extern \"C\" __global__ void __launch_bounds__(256) fz(%s) {\n"
            "  for (long t = blockIdx.x * (long)blockDim.x + threadIdx.x; t < n; t += (long)gridDim.x * blockDim.x) {\n"
            "%s\n    out[t] = %s;\n  }\n}\n" % (", ".join(params), "\n".join(loads), cexpr))


class _Nvrtc:
    """NVRTC + driver API through cuda-python: source -> cubin for sm_100a -> CUfunction."""

    def __init__(self):
        from cuda.bindings import driver, nvrtc
        self.driver, self.nvrtc = driver, nvrtc
        self.cache = {}

    def compile_to_cubin(self, src, arch="sm_100a"):
        nv = self.nvrtc
        err, prog = nv.nvrtcCreateProgram(src.encode(), b"fz.cu", 0, [], [])
        assert err == nv.nvrtcResult.NVRTC_SUCCESS, err
        opts = [b"--gpu-architecture=" + arch.encode(), b"--fmad=false"]
        (err,) = nv.nvrtcCompileProgram(prog, len(opts), opts)
        if err != nv.nvrtcResult.NVRTC_SUCCESS:
            _, size = nv.nvrtcGetProgramLogSize(prog)
            log = b" " * size
            nv.nvrtcGetProgramLog(prog, log)
            raise RuntimeError("NVRTC: " + log.decode(errors="replace"))
        _, size = nv.nvrtcGetCUBINSize(prog)
        cubin = b" " * size
        (err,) = nv.nvrtcGetCUBIN(prog, cubin)
        assert err == nv.nvrtcResult.NVRTC_SUCCESS, err
        nv.nvrtcDestroyProgram(prog)
        return cubin

    def compile(self, src):
        if src in self.cache:
            return self.cache[src]
        d = self.driver
        cubin = self.compile_to_cubin(src)
        err, mod = d.cuModuleLoadData(cubin)
        assert err == d.CUresult.CUDA_SUCCESS, err
        err, fn = d.cuModuleGetFunction(mod, b"fz")
        assert err == d.CUresult.CUDA_SUCCESS, err
        self.cache[src] = (fn, mod)
        return self.cache[src]

    def launch(self, kern, n, args, types, stream):
        d = self.driver
        blocks = max(1, min((n + 255) // 256, 148 * 16))
        (err,) = d.cuLaunchKernel(kern[0], blocks, 1, 1, 256, 1, 1, 0, d.CUstream(int(stream)),
                                  (tuple(args), tuple(types)), 0)
        assert err == d.CUresult.CUDA_SUCCESS, err
