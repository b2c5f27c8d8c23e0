"""Host-side camera mirroring the reference's `camera.rs` (produces the `view` matrix that
crosses the C-ABI boundary as 16 floats)."""
from __future__ import annotations

import math

import numpy as np

from . import mathx

F = np.float32


class Camera:
    """camera.rs:3-7.  `Camera()` is `Camera::default()` (camera.rs:46-54)."""

    def __init__(self, pos=(0.0, 0.0, -5.0), up=(0.0, 1.0, 0.0), dir=(0.0, 0.0, 1.0)):
        self.pos = mathx.vec(pos)
        self.up = mathx.normalized(up)
        self.dir = mathx.normalized(dir)

    @classmethod
    def orbit(cls, angle: float, radius: float = 5.0, height: float = 0.0) -> "Camera":
        """Orbit constructor used by the camera-sweep config: position on a
        circle of `radius` in the xz-plane, looking at the origin, up +y.  Not in the reference
        (it only has `Default`); it feeds the same get_view_matrix()."""
        a = F(angle)
        pos = mathx.vec([F(F(radius) * mathx.sin(a)), F(height), F(F(-F(radius)) * mathx.cos(a))])
        d = mathx.vec([F(-pos[0]), F(-pos[1]), F(-pos[2])])
        return cls(pos=pos, up=(0.0, 1.0, 0.0), dir=d)

    def get_view_matrix(self) -> np.ndarray:
        """camera.rs:10-43"""
        cam_z = (self.dir * F(-1.0)).astype(np.float32)  # Neg = *= -1.0 (math/vector.rs:86-96)
        cam_x = mathx.normalized(mathx.cross3(cam_z, self.up))
        cam_y = mathx.normalized(mathx.cross3(cam_x, cam_z))
        rotation_inv = mathx.mat4(
            cam_x[0], cam_y[0], cam_z[0], 0.0,
            cam_x[1], cam_y[1], cam_z[1], 0.0,
            cam_x[2], cam_y[2], cam_z[2], 0.0,
            0.0, 0.0, 0.0, 1.0,
        ).T.copy()
        vec_to_pos = (self.pos - mathx.vec([0.0, 0.0, 0.0])).astype(np.float32)
        neg = (vec_to_pos * F(-1.0)).astype(np.float32)
        translation_inv = mathx.translate(neg[0], neg[1], neg[2])
        return mathx.matmul(rotation_inv, translation_inv)
