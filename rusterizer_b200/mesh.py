"""Host-side mesh container + generators mirroring the reference's `mesh.rs`.

A Mesh is the FFI input of `Renderer.render`: positions f32[Nv][3], attributes
f32[Nv][6] = (r,g,b,a,u,v) (graphics_primitives.rs:10-13), indices u32[3*Nt]
(the reference stores `usize`; the boundary narrows to u32, see include/rz.h).
"""
from __future__ import annotations

from dataclasses import dataclass

import numpy as np

from . import mathx

F = np.float32
_RED, _GREEN, _BLUE, _WHITE = (1, 0, 0, 1), (0, 1, 0, 1), (0, 0, 1, 1), (1, 1, 1, 1)


@dataclass
class Mesh:
    """mesh.rs:5-12"""
    vertices: np.ndarray    # (Nv,3) f32
    indices: np.ndarray     # (3*Nt,) u32
    attributes: np.ndarray  # (Nv,6) f32

    def __post_init__(self):
        self.vertices = np.ascontiguousarray(self.vertices, dtype=np.float32).reshape(-1, 3)
        self.indices = np.ascontiguousarray(self.indices, dtype=np.uint32).reshape(-1)
        self.attributes = np.ascontiguousarray(self.attributes, dtype=np.float32).reshape(-1, 6)
        assert self.vertices.shape[0] == self.attributes.shape[0]

    @property
    def n_vertices(self) -> int:
        return self.vertices.shape[0]

    @property
    def n_triangles(self) -> int:
        return self.indices.size // 3


def _attrs(colors, uvs):
    return np.concatenate([np.asarray(colors, np.float32), np.asarray(uvs, np.float32)], axis=1)


def centered_quad(width: float) -> Mesh:
    """mesh.rs:15-41"""
    h = F(F(width) / F(2.0))
    v = [[-h, h, 2.0], [h, h, 2.0], [h, -h, 2.0], [-h, -h, 2.0]]
    a = _attrs([_RED, _BLUE, _GREEN, _WHITE], [[0, 0], [1, 0], [1, 1], [0, 1]])
    return Mesh(v, [0, 1, 2, 0, 2, 3], a)


def triangle() -> Mesh:
    """mesh.rs:44-66"""
    v = [[-1.0, -1.0, 2.0], [0.0, 1.0, 2.0], [1.0, -1.0, 2.0]]
    a = _attrs([_RED, _BLUE, _GREEN], [[0, 1], [1, 0], [1, 1]])
    return Mesh(v, [0, 1, 2], a)


def cube(width: float) -> Mesh:
    """mesh.rs:69-149"""
    base = np.array([
        [-0.5, 0.5, -0.5], [0.5, 0.5, -0.5], [0.5, -0.5, -0.5], [-0.5, -0.5, -0.5],   # front
        [0.5, 0.5, 0.5], [-0.5, 0.5, 0.5], [-0.5, -0.5, 0.5], [0.5, -0.5, 0.5],       # back
        [-0.5, 0.5, 0.5], [-0.5, 0.5, -0.5], [-0.5, -0.5, -0.5], [-0.5, -0.5, 0.5],   # left
        [0.5, 0.5, -0.5], [0.5, 0.5, 0.5], [0.5, -0.5, 0.5], [0.5, -0.5, -0.5],       # right
        [-0.5, 0.5, -0.5], [-0.5, 0.5, 0.5], [0.5, 0.5, 0.5], [0.5, 0.5, -0.5],       # top
        [-0.5, -0.5, 0.5], [-0.5, -0.5, -0.5], [0.5, -0.5, -0.5], [0.5, -0.5, 0.5],   # bottom
    ], dtype=np.float32)
    v = (base * F(width)).astype(np.float32)
    idx = []
    for i in range(6):
        o = i * 4
        idx += [o + 0, o + 1, o + 2, o + 0, o + 2, o + 3]
    colors = [(_RED, _BLUE, _GREEN)[i % 3] for i in range(24)]
    uvs = [[0, 0], [1, 0], [1, 1], [0, 1]] * 6
    return Mesh(v, idx, _attrs(colors, uvs))


def sphere(radius: float, n_phi_samples: int = 17, n_theta_samples: int = 9) -> Mesh:
    """mesh.rs:152-207, with the two sample counts (hard-coded 17 x 9 there) as parameters so
    the same generator yields the 1M-triangle benchmark sphere (1001 x 501)."""
    r = F(radius)
    i = np.arange(n_theta_samples, dtype=np.float32)
    j = np.arange(n_phi_samples, dtype=np.float32)
    theta_ratio = (i / F(n_theta_samples - 1)).astype(np.float32)
    phi_ratio = (j / F(n_phi_samples - 1)).astype(np.float32)
    phi = ((F(np.pi) * F(2.0)) * phi_ratio).astype(np.float32)
    theta = (F(np.pi) * theta_ratio).astype(np.float32)

    # f32::sin / f32::cos of mesh.rs:176-178 = libm sinf / cosf (not correctly rounded for every argument): same calls
    st, ct = mathx.sinf_array(theta), mathx.cosf_array(theta)
    sp, cp = mathx.sinf_array(phi), mathx.cosf_array(phi)
    rs = (r * st).astype(np.float32)[:, None]
    x = (rs * cp[None, :]).astype(np.float32)
    y = np.broadcast_to((r * ct).astype(np.float32)[:, None], x.shape)
    z = (rs * sp[None, :]).astype(np.float32)
    verts = np.stack([x, y, z], axis=-1).reshape(-1, 3)

    ii, jj = np.meshgrid(np.arange(n_theta_samples - 1), np.arange(n_phi_samples - 1), indexing="ij")
    a = n_phi_samples * ii + jj
    b = n_phi_samples * ii + (jj + 1)
    c = n_phi_samples * (ii + 1) + (jj + 1)
    d = n_phi_samples * (ii + 1) + jj
    idx = np.stack([a, b, c, a, c, d], axis=-1).reshape(-1).astype(np.uint32)

    attrs = np.empty((verts.shape[0], 6), np.float32)
    attrs[:, 0:3] = np.abs(verts)
    attrs[:, 3] = 1.0
    attrs[:, 4] = np.broadcast_to(phi_ratio[None, :], x.shape).reshape(-1)
    attrs[:, 5] = np.broadcast_to(theta_ratio[:, None], x.shape).reshape(-1)
    return Mesh(verts, idx, attrs)
