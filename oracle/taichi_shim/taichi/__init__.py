"""Minimal pure-Python emulation of the Taichi API surface used by uc-vision/taichi_image.

TEST INFRASTRUCTURE ONLY.  It exists so that the UNMODIFIED reference sources under /root/reference
can be imported and executed in this container (Taichi itself is not installed, not vendored, not
pinned by the reference, and there is no network).  Kernels run as plain Python loops, so only small
inputs are practical; oracle/gen_golden.py uses it to produce the golden vectors in tests/golden/.

Taichi semantics restated here (Taichi documentation defaults; they cannot be verified against a
real Taichi build in this image):
  * default_fp = f32, default_ip = i32: Python float literals and float arithmetic are f32;
  * ti.cast(float -> int) truncates toward zero; stores to f16 round to nearest even;
  * ti.round rounds half away from zero;
  * ti.atomic_min/max/add on kernel locals and struct fields update them in place (implemented by an
    AST rewrite `ti.atomic_min(x, v)` -> `x = atomic_min(x, v)`); loops run sequentially, so float
    reductions are accumulated in index order;
  * loading an f16 element and combining it with f32 values / literals computes in f32;
  * `@` on mat3/vec3 is a row-by-row dot product accumulated left to right, without FMA contraction;
  * min / max (ti.atomic_*, tm.min/max, Vector.max) follow IEEE minNum / maxNum (LLVM llvm.minnum /
    llvm.maxnum, CUDA fminf / fmaxf): a NaN operand is ignored.
"""
from __future__ import annotations

import ast
import copy
import inspect
import itertools
import sys
import textwrap

import numpy as np

try:
    import torch
except Exception:  # pragma: no cover
    torch = None

_F32 = np.float32


# ------------------------------------------------------------------ dtypes
class DType:
    def __init__(self, name, np_type):
        self.name, self.np = name, np_type

    def __call__(self, x):
        return cast(x, self)

    def __repr__(self):
        return f"ti.{self.name}"

    @property
    def is_float(self):
        return np.issubdtype(self.np, np.floating)


u8 = uint8 = DType("u8", np.uint8)
u16 = uint16 = DType("u16", np.uint16)
u32 = uint32 = DType("u32", np.uint32)
i8 = int8 = DType("i8", np.int8)
i16 = int16 = DType("i16", np.int16)
i32 = int32 = DType("i32", np.int32)
i64 = int64 = DType("i64", np.int64)
f16 = float16 = DType("f16", np.float16)
f32 = float32 = DType("f32", np.float32)
f64 = float64 = DType("f64", np.float64)
_BY_NP = {np.dtype(d.np): d for d in (u8, u16, u32, i8, i16, i32, i64, f16, f32, f64)}

cuda, cpu, gpu = "cuda", "cpu", "gpu"
DEBUG, INFO, WARN, ERROR, TRACE = "debug", "info", "warn", "error", "trace"


def init(*args, **kwargs):
    return None


def loop_config(**kwargs):
    return None


def static(x, *rest):
    return x if not rest else (x,) + rest


def template():
    return _Template()


class _Template:
    pass


# ------------------------------------------------------------------ vectors / matrices
def _is_float_dtype(dt):
    return np.issubdtype(dt, np.floating)


def _raw(x):
    if isinstance(x, Vec):
        return x.a
    if isinstance(x, (tuple, list)):
        return np.asarray([_raw(v) for v in x])
    return x


def _is_floaty(x):
    if isinstance(x, Vec):
        return _is_float_dtype(x.a.dtype)
    if isinstance(x, (float, np.floating)):
        return True
    if isinstance(x, (tuple, list)):
        return any(_is_floaty(v) for v in x)
    if isinstance(x, np.ndarray):
        return _is_float_dtype(x.dtype)
    return False


class Vec:
    """Fixed-size vector with Taichi promotion: anything float -> f32, all-int -> i32 (bit ops keep
    unsigned small types)."""
    __slots__ = ("a",)
    __array_ufunc__ = None        # numpy scalars defer to Vec.__r*__ instead of broadcasting over it
    _SW = {"x": 0, "y": 1, "z": 2, "w": 3, "r": 0, "g": 1, "b": 2}

    def __init__(self, values, dtype=None):
        a = np.array(_raw(values))
        if a.ndim == 0:
            a = a.reshape(1)
        if dtype is not None:
            a = _cast_array(a, dtype)
        elif _is_float_dtype(a.dtype):
            a = a.astype(_F32)
        elif a.dtype == np.int64:
            a = a.astype(np.int32)
        self.a = a

    # -- access
    def __len__(self):
        return len(self.a)

    def __iter__(self):
        return iter(self.a)

    def __getitem__(self, i):
        if isinstance(i, slice):
            return Vec(self.a[i].copy())
        return self.a[int(i)]

    def __setitem__(self, i, v):
        self.a[int(i)] = v

    def __getattr__(self, name):
        sw = Vec._SW
        if name and all(c in sw for c in name):
            if len(name) == 1:
                return self.a[sw[name]]
            return Vec(self.a[[sw[c] for c in name]].copy())
        raise AttributeError(name)

    def __repr__(self):
        return f"Vec({self.a.tolist()}, {self.a.dtype})"

    def copy(self):
        return Vec(self.a.copy())

    def __deepcopy__(self, memo):
        return self.copy()

    # -- arithmetic
    def _bin(self, other, op, reverse=False, bitwise=False):
        o = _raw(other)
        if bitwise:
            l, r = (o, self.a) if reverse else (self.a, o)
            return Vec(op(l, r))
        if _is_floaty(self) or _is_floaty(other):
            l = self.a.astype(_F32)
            r = np.asarray(o, dtype=_F32)
        else:
            l = self.a.astype(np.int32)
            r = np.asarray(o, dtype=np.int32)
        if reverse:
            l, r = r, l
        with np.errstate(all="ignore"):
            return Vec(op(l, r))

    def __add__(self, o): return self._bin(o, np.add)
    def __radd__(self, o): return self._bin(o, np.add, True)
    def __sub__(self, o): return self._bin(o, np.subtract)
    def __rsub__(self, o): return self._bin(o, np.subtract, True)
    def __mul__(self, o): return self._bin(o, np.multiply)
    def __rmul__(self, o): return self._bin(o, np.multiply, True)

    def __truediv__(self, o):
        l = self.a.astype(_F32)
        with np.errstate(all="ignore"):
            return Vec(l / np.asarray(_raw(o), dtype=_F32))

    def __rtruediv__(self, o):
        with np.errstate(all="ignore"):
            return Vec(np.asarray(_raw(o), dtype=_F32) / self.a.astype(_F32))

    def __floordiv__(self, o): return self._bin(o, np.floor_divide)
    def __rfloordiv__(self, o): return self._bin(o, np.floor_divide, True)
    def __rpow__(self, o): return self._bin(o, np.power, True)
    def __mod__(self, o): return self._bin(o, np.mod)
    def __pow__(self, o): return self._bin(o, np.power)
    def __neg__(self): return Vec(-self.a)
    def __and__(self, o): return self._bin(o, np.bitwise_and, bitwise=True)
    def __or__(self, o): return self._bin(o, np.bitwise_or, bitwise=True)
    def __lshift__(self, o): return self._bin(o, np.left_shift, bitwise=True)
    def __rshift__(self, o): return self._bin(o, np.right_shift, bitwise=True)
    def __lt__(self, o): return Vec(self.a < _raw(o))
    def __le__(self, o): return Vec(self.a <= _raw(o))
    def __gt__(self, o): return Vec(self.a > _raw(o))
    def __ge__(self, o): return Vec(self.a >= _raw(o))

    def max(self): return np.fmax.reduce(self.a)
    def min(self): return np.fmin.reduce(self.a)
    def sum(self):
        s = self.a[0]
        for v in self.a[1:]:
            s = s + v
        return s


class Mat:
    def __init__(self, values, n):
        self.m = np.asarray(_raw(values), dtype=_F32).reshape(n, n)
        self.n = n

    def __matmul__(self, v):
        x = v.a.astype(_F32) if isinstance(v, Vec) else np.asarray(v, _F32)
        out = []
        for r in range(self.n):
            acc = self.m[r, 0] * x[0]
            for k in range(1, self.n):
                acc = acc + self.m[r, k] * x[k]
            out.append(acc)
        return Vec(np.array(out, dtype=_F32))

    def inverse(self):
        return Mat(np.linalg.inv(self.m.astype(np.float64)).astype(_F32), self.n)


class VecType:
    def __init__(self, n, dtype):
        self.n, self.dtype = n, dtype

    def __call__(self, *args):
        flat = []
        for a in args:
            if isinstance(a, Vec):
                flat.extend(a.a.tolist() if False else list(a.a))
            elif isinstance(a, (tuple, list, np.ndarray)):
                flat.extend(list(np.asarray(_raw(a)).reshape(-1)))
            else:
                flat.append(a)
        if len(flat) == 1:
            flat = flat * self.n
        assert len(flat) == self.n, f"vector({self.n}) built from {len(flat)} values"
        return Vec(np.array(flat), self.dtype)


class MatType:
    def __init__(self, n):
        self.n = n

    def __call__(self, *args):
        if len(args) == 1:
            args = args[0]
        return Mat(list(np.asarray(_raw(args), dtype=np.float64).reshape(-1)), self.n)


def Vector(values, dt=None):
    return Vec(values, dt)


# ------------------------------------------------------------------ casting
def _cast_array(a, dtype):
    a = np.asarray(a)
    with np.errstate(all="ignore"):
        if not dtype.is_float and _is_float_dtype(a.dtype):
            return np.trunc(a).astype(np.int64).astype(dtype.np)
        return a.astype(dtype.np)


def cast(x, dtype):
    if isinstance(x, Vec):
        return Vec(_cast_array(x.a, dtype))
    return _cast_array(np.asarray(x), dtype)[()]


def round(x, dtype=None):  # noqa: A001  (ti.round: half away from zero)
    v = _raw(x)
    r = np.sign(v) * np.floor(np.abs(v) + _F32(0.5))
    if isinstance(x, Vec):
        return Vec(r) if dtype is None else cast(Vec(r), dtype)
    return _F32(r) if dtype is None else cast(r, dtype)


def select(cond, a, b):
    c = _raw(cond)
    if isinstance(a, Vec) or isinstance(b, Vec) or isinstance(cond, Vec):
        return Vec(np.where(c, _raw(a), _raw(b)))
    return a if c else b


def _f32(x):
    return x if isinstance(x, (np.floating, np.integer)) and not isinstance(x, np.float64) else _F32(x)


def atomic_min(a, b):
    return np.fmin(_F32(a), _F32(b)) if _is_floaty(a) or _is_floaty(b) else min(a, b)


def atomic_max(a, b):
    return np.fmax(_F32(a), _F32(b)) if _is_floaty(a) or _is_floaty(b) else max(a, b)


def atomic_add(a, b):
    return a + b


# ------------------------------------------------------------------ ndarray arguments
class NdAnn:
    def __init__(self, dtype=None, ndim=None):
        self.dtype, self.ndim = dtype, ndim


class NdArray:
    """View of a numpy array as a Taichi ndarray (optionally of vector elements)."""

    def __init__(self, arr, ann: NdAnn):
        self.arr = arr
        self.vec = isinstance(ann.dtype, VecType)
        if self.vec:
            self.ndim = arr.ndim - 1
            assert arr.shape[-1] == ann.dtype.n, f"vector({ann.dtype.n}) ndarray got shape {arr.shape}"
            assert ann.ndim is None or ann.ndim == self.ndim, f"ndim {ann.ndim} vs array {arr.shape}"
        else:
            self.ndim = arr.ndim
        self.dtype = _BY_NP[arr.dtype]

    @property
    def shape(self):
        return tuple(self.arr.shape[: self.ndim])

    @staticmethod
    def _index(i):
        if i is None:
            return ()
        if isinstance(i, Vec):
            return tuple(int(v) for v in i.a)
        if isinstance(i, tuple):
            out = []
            for v in i:
                out.extend(NdArray._index(v) if isinstance(v, (Vec, tuple)) else (int(v),))
            return tuple(out)
        return (int(i),)

    @staticmethod
    def _up(v):
        if v.dtype == np.float16:      # f16 loads take part in f32 arithmetic
            return v.astype(_F32)
        return v

    def __getitem__(self, i):
        idx = self._index(i)
        if any(k < 0 or k >= s for k, s in zip(idx, self.arr.shape)):
            raise IndexError(f"ndarray index {idx} out of bounds for {self.arr.shape}")
        v = self.arr[idx]
        if self.vec:
            return Vec(self._up(np.array(v)))
        return self._up(np.asarray(v))[()]

    def __setitem__(self, i, val):
        idx = self._index(i)
        if any(k < 0 or k >= s for k, s in zip(idx, self.arr.shape)):
            raise IndexError(f"ndarray index {idx} out of bounds for {self.arr.shape}")
        self.arr[idx] = _cast_array(np.asarray(_raw(val)), self.dtype)


class _NdRange:
    def __init__(self, dims):
        self.dims = [range(*d) if isinstance(d, (tuple, list)) else range(int(d)) for d in dims]

    def __iter__(self):
        if len(self.dims) == 1:
            return iter(self.dims[0])
        return itertools.product(*self.dims)


def ndrange(*dims):
    return _NdRange(dims)


def grouped(x):
    if isinstance(x, NdArray):
        x = _NdRange(x.shape)
    for idx in itertools.product(*x.dims):
        yield Vec(np.array(idx, dtype=np.int32))


class _Types:
    @staticmethod
    def vector(n, dtype):
        return VecType(n, dtype)

    @staticmethod
    def matrix(n, m, dtype):
        assert n == m
        return MatType(n)

    @staticmethod
    def ndarray(dtype=None, ndim=None, **kw):
        return NdAnn(dtype, ndim)


types = _Types()


# ------------------------------------------------------------------ decorators
class _AtomicRewrite(ast.NodeTransformer):
    _FN = {"atomic_min": "__ti_atomic_min", "atomic_max": "__ti_atomic_max", "atomic_add": "__ti_atomic_add"}

    def visit_Expr(self, node):
        c = node.value
        if (isinstance(c, ast.Call) and isinstance(c.func, ast.Attribute) and c.func.attr in self._FN
                and isinstance(c.func.value, ast.Name) and c.func.value.id == "ti" and len(c.args) == 2):
            target = copy.deepcopy(c.args[0])
            for n in ast.walk(target):
                if hasattr(n, "ctx"):
                    n.ctx = ast.Load()
            target.ctx = ast.Store()
            call = ast.Call(func=ast.Name(id=self._FN[c.func.attr], ctx=ast.Load()), args=c.args, keywords=[])
            return ast.copy_location(ast.Assign(targets=[target], value=call), node)
        return node


def _rebuild(fn, frame=None):
    """Recompile `fn` with in-place atomics rewritten; closure variables (and the locals of the scope
    that applies the decorator, which annotations may reference) become globals of the copy."""
    src = textwrap.dedent(inspect.getsource(fn))
    tree = ast.parse(src)
    fdef = tree.body[0]
    fdef.decorator_list = []
    tree = ast.fix_missing_locations(_AtomicRewrite().visit(tree))
    ns = dict(fn.__globals__)
    if frame is not None and frame.f_locals is not frame.f_globals:
        ns.update(frame.f_locals)
    cv = inspect.getclosurevars(fn)
    ns.update(cv.nonlocals)
    ns.update({"__ti_atomic_min": atomic_min, "__ti_atomic_max": atomic_max, "__ti_atomic_add": atomic_add})
    # strip annotations that reference closure-only names evaluated at def time
    code = compile(tree, inspect.getsourcefile(fn) or "<taichi_shim>", "exec", dont_inherit=True)
    exec(code, ns)
    new = ns[fdef.name]
    new.__ti_original__ = fn
    return new


def func(fn):
    return _rebuild(fn, sys._getframe(1))


def _to_numpy(x):
    if torch is not None and isinstance(x, torch.Tensor):
        assert x.device.type == "cpu", "taichi_shim runs on CPU tensors only"
        if x.dtype == torch.uint16:
            return x.view(torch.int16).numpy().view(np.uint16)
        return x.numpy()
    return x


def kernel(fn):
    impl = _rebuild(fn, sys._getframe(1))
    sig = inspect.signature(impl)
    params = list(sig.parameters.values())

    def wrapper(*args, **kwargs):
        bound = sig.bind(*args, **kwargs)
        call = {}
        for p in params:
            v = bound.arguments[p.name]
            ann = p.annotation
            if isinstance(ann, NdAnn):
                arr = _to_numpy(v)
                assert isinstance(arr, np.ndarray), f"{fn.__name__}: {p.name} must be an array"
                call[p.name] = NdArray(arr, ann)
            elif isinstance(ann, DType):
                call[p.name] = cast(v, ann)
            else:
                call[p.name] = v
        return impl(**call)

    wrapper.__name__ = fn.__name__
    wrapper.__ti_kernel__ = True
    return wrapper


def dataclass(cls):
    fields = dict(getattr(cls, "__annotations__", {}))

    def __init__(self, *args, **kwargs):
        names = list(fields)
        vals = dict(zip(names, args))
        vals.update(kwargs)
        for n in names:
            v = vals.get(n, 0)
            t = fields[n]
            if isinstance(t, DType):
                v = cast(v, t)
            elif isinstance(t, VecType):
                v = t(v) if not isinstance(v, Vec) else Vec(v.a.copy(), t.dtype)
            else:
                v = copy.deepcopy(v)
            object.__setattr__(self, n, v)

    cls.__init__ = __init__
    return cls


def data_oriented(cls):
    return cls


def field(*args, **kwargs):
    raise NotImplementedError("ti.field is not emulated")


from . import math  # noqa: E402,F401
