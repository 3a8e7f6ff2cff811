"""taichi.math subset used by the reference (TEST INFRASTRUCTURE ONLY, see taichi/__init__.py)."""
import numpy as np

from . import Vec, VecType, MatType, f32, i32, _raw, _F32

vec2, vec3, vec4 = VecType(2, f32), VecType(3, f32), VecType(4, f32)
ivec2, ivec3, ivec4 = VecType(2, i32), VecType(3, i32), VecType(4, i32)
mat3 = MatType(3)
inf = float("inf")


def _wrap(x, r):
    return Vec(r) if isinstance(x, Vec) else (r if isinstance(r, np.generic) else _F32(r))


def _f(x):
    r = _raw(x)
    if isinstance(r, np.ndarray):
        return r.astype(_F32) if np.issubdtype(r.dtype, np.floating) else r
    if isinstance(r, float):
        return _F32(r)
    return r


def clamp(x, xmin, xmax):
    a = _f(x)
    with np.errstate(all="ignore"):
        r = np.fmin(np.fmax(a, _f(xmin)), _f(xmax))
    if isinstance(x, Vec) or isinstance(xmin, Vec) or isinstance(xmax, Vec):
        return Vec(r)
    return r


def mix(x, y, a):
    return x * (1.0 - a) + y * a


def dot(a, b):
    a = a if isinstance(a, Vec) else Vec(a)
    b = b if isinstance(b, Vec) else Vec(b)
    p = a * b
    return p.sum()


def pow(x, y):  # noqa: A001
    with np.errstate(all="ignore"):
        if isinstance(x, Vec) or isinstance(y, Vec):
            return Vec(np.power(np.asarray(_raw(x), _F32), np.asarray(_raw(y), _F32)))
        return np.power(_F32(x), _F32(y))


def exp(x):
    with np.errstate(all="ignore"):
        return _wrap(x, np.exp(np.asarray(_raw(x), _F32)))[()] if not isinstance(x, Vec) else Vec(np.exp(x.a.astype(_F32)))


def log(x):
    with np.errstate(all="ignore"):
        return Vec(np.log(x.a.astype(_F32))) if isinstance(x, Vec) else np.log(_F32(x))


def max(a, b):  # noqa: A001
    with np.errstate(all="ignore"):
        r = np.fmax(_f(a), _f(b))
    return Vec(r) if isinstance(a, Vec) or isinstance(b, Vec) else r


def min(a, b):  # noqa: A001
    with np.errstate(all="ignore"):
        r = np.fmin(_f(a), _f(b))
    return Vec(r) if isinstance(a, Vec) or isinstance(b, Vec) else r
