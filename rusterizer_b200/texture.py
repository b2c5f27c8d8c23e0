"""Host-side texture container mirroring the reference's `texture.rs` layout
(row-major u8[h][w][texel_width], texel_width in {3,4}, origin top-left; texture.rs:8-13,51)."""
from __future__ import annotations

import hashlib
from dataclasses import dataclass

import numpy as np

# sha256 of the decoded RGBA bytes of the reference's images/checkerboard.png
CHECKERBOARD_SHA256 = "5828e42c6a0fd3863013375a564a04668fc1aa55274dde86e647d44d3af90539"


@dataclass
class Texture:
    texels: np.ndarray  # (h, w, texel_width) u8

    def __post_init__(self):
        self.texels = np.ascontiguousarray(self.texels, dtype=np.uint8)
        assert self.texels.ndim == 3 and self.texels.shape[2] in (3, 4)

    @property
    def width(self):
        return self.texels.shape[1]

    @property
    def height(self):
        return self.texels.shape[0]

    @property
    def texel_width(self):
        return self.texels.shape[2]

    @classmethod
    def from_png_file(cls, path) -> "Texture":
        """texture.rs:26-45.  The reference hard-codes texel_width = 4 whatever the file holds (texture.rs:41),
        which is only right for RGBA files such as images/checkerboard.png; here the width follows the decoded
        layout (RGB8 -> 3, RGBA8 -> 4), the two layouts read_texel handles (texture.rs:47-63)."""
        from .image import read_png  # host-side I/O only

        return cls(read_png(path))

    @classmethod
    def checkerboard(cls) -> "Texture":
        """The decoded content of the reference fixture images/checkerboard.png regenerated
        procedurally: 400x400 RGBA, 4x4 squares of 100 px, pure black/white, top-left black,
        alpha 255.  Verified against the fixture's sha256 (tests/test_host.py)."""
        yy, xx = np.mgrid[0:400, 0:400]
        white = ((xx // 100 + yy // 100) % 2).astype(np.uint8) * 255
        t = np.empty((400, 400, 4), np.uint8)
        t[..., 0] = t[..., 1] = t[..., 2] = white
        t[..., 3] = 255
        return cls(t)

    def sha256(self) -> str:
        return hashlib.sha256(self.texels.tobytes()).hexdigest()
