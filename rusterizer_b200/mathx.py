"""Host-side f32 linear algebra mirroring the reference's `math` module.

Only builds the matrices / vectors that cross the C-ABI boundary as plain data
(host-only for project / rotate* / translate / normalized).
Every operation is one IEEE binary32 rounding in the reference's source order:
  * dot: sum starts at 0.0, sequential            (math/vector.rs:17-23)
  * Mat x Mat: R[i][j] = dot(row_i(A), col_j(B))  (math/matrix.rs:56-79)
  * Mat x Vec: r[i] = dot(row_i(M), v)            (math/vector.rs:219-240)
Transcendentals: Rust's f32::sin / cos / tan lower to the platform libm (sinf / cosf / tanf), which glibc does NOT
round correctly for every argument (sinf(0.29452431) is one ulp off the correctly rounded value).  To hand the raster
path the very inputs the Rust host would, sin / cos / tan call the same libm through ctypes (sinf_array etc. for the
mesh generators); tests/test_host.py compares everything built here, bit for bit, with the C++ mirror.  sqrt is
correctly rounded in every libm.
"""
from __future__ import annotations

import math

import numpy as np

F = np.float32


def _f(x) -> np.float32:
    return np.float32(x)


def _libm():
    import ctypes
    import ctypes.util

    lib = ctypes.CDLL(ctypes.util.find_library("m") or "libm.so.6")
    for name in ("sinf", "cosf", "tanf"):
        fn = getattr(lib, name)
        fn.restype, fn.argtypes = ctypes.c_float, [ctypes.c_float]
    return lib


_M = _libm()


def sin(x) -> np.float32:
    return F(_M.sinf(float(F(x))))


def cos(x) -> np.float32:
    return F(_M.cosf(float(F(x))))


def tan(x) -> np.float32:
    return F(_M.tanf(float(F(x))))


def sinf_array(a: np.ndarray) -> np.ndarray:
    """libm sinf over an f32 array (what a Rust loop over f32::sin computes)."""
    return np.array([_M.sinf(float(v)) for v in np.asarray(a, np.float32).reshape(-1)], np.float32).reshape(np.shape(a))


def cosf_array(a: np.ndarray) -> np.ndarray:
    return np.array([_M.cosf(float(v)) for v in np.asarray(a, np.float32).reshape(-1)], np.float32).reshape(np.shape(a))


def sqrt(x) -> np.float32:
    return F(math.sqrt(float(F(x))))


def vec(v) -> np.ndarray:
    return np.asarray(v, dtype=np.float32)


def dot(a, b) -> np.float32:
    s = F(0.0)
    for x, y in zip(vec(a), vec(b)):
        s = F(s + F(x * y))
    return s


def vlen(v) -> np.float32:
    """Vector::len, math/vector.rs:38-40"""
    acc = F(0.0)
    for e in vec(v):
        acc = F(acc + F(e * e))
    return sqrt(acc)


def normalized(v) -> np.ndarray:
    """Vector::normalized, math/vector.rs:42-44"""
    v = vec(v)
    return (v / vlen(v)).astype(np.float32)


def cross3(a, b) -> np.ndarray:
    """Vec3::cross, math/vector.rs:196-203"""
    a, b = vec(a), vec(b)
    return vec([F(F(a[1] * b[2]) - F(a[2] * b[1])),
                F(F(a[2] * b[0]) - F(a[0] * b[2])),
                F(F(a[0] * b[1]) - F(a[1] * b[0]))])


def mat4(*x) -> np.ndarray:
    return np.asarray(x, dtype=np.float32).reshape(4, 4)


def identity() -> np.ndarray:
    return np.eye(4, dtype=np.float32)


def matmul(a, b) -> np.ndarray:
    a, b = vec(a).reshape(4, 4), vec(b).reshape(4, 4)
    r = np.empty((4, 4), np.float32)
    for i in range(4):
        for j in range(4):
            r[i, j] = dot(a[i], b[:, j])
    return r


def matvec(m, v) -> np.ndarray:
    m, v = vec(m).reshape(4, 4), vec(v)
    return vec([dot(m[i], v) for i in range(4)])


def translate(x, y, z) -> np.ndarray:
    """math/transform.rs:10-17"""
    return mat4(1, 0, 0, x, 0, 1, 0, y, 0, 0, 1, z, 0, 0, 0, 1)


def rotate_x(rad) -> np.ndarray:
    """math/transform.rs:19-41"""
    c, s = cos(rad), sin(rad)
    return mat4(1, 0, 0, 0, 0, c, -s, 0, 0, s, c, 0, 0, 0, 0, 1)


def rotate_y(rad) -> np.ndarray:
    """math/transform.rs:43-65"""
    c, s = cos(rad), sin(rad)
    return mat4(c, 0, s, 0, 0, 1, 0, 0, -s, 0, c, 0, 0, 0, 0, 1)


def rotate_z(rad) -> np.ndarray:
    """math/transform.rs:67-89"""
    c, s = cos(rad), sin(rad)
    return mat4(c, -s, 0, 0, s, c, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1)


def rotate(x, y, z) -> np.ndarray:
    """math/transform.rs:91-96 : rotate_z(z) * rotate_y(y) * rotate_x(x), left-assoc"""
    return matmul(matmul(rotate_z(z), rotate_y(y)), rotate_x(x))


def project(near, far, aspect_ratio, vert_fov) -> np.ndarray:
    """math/mod.rs:92-122"""
    near, far, aspect_ratio, vert_fov = F(near), F(far), F(aspect_ratio), F(vert_fov)
    assert near > 0.0
    half_width = F(tan(F(vert_fov / F(2.0))) * near)
    half_height = F(aspect_ratio * half_width)
    return mat4(
        F(near / half_width), 0, 0, 0,
        0, F(near / half_height), 0, 0,
        0, 0, F(F(-F(far + near)) / F(far - near)), F(F(F(F(-2.0) * far) * near) / F(far - near)),
        0, 0, -1.0, 0,
    )
