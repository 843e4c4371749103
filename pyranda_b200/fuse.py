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
        # dense storage in any axis order: the kernel walks the flat storage
        st = sorted(zip(ref.stride(), ref.shape))
        run = 1
        for s_, e_ in st:
            if e_ != 1 and s_ != run:
                return spec.fallback(self.xp, *vals)
            run *= e_
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
