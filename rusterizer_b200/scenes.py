"""Scene builders for the BASELINE.json configs.  Pure host-side data:
the same Scene feeds the CUDA path (through the C ABI) and, in tests/bench, the CPU oracle.

  C1  default `cargo run --release` scene (main.rs:91-155) at a fixed `elapsed`
  C2  1M-triangle textured UV-sphere, 1920x1080
  C3  250K triangles that all straddle the near plane, 3840x2160
  C4  8192x8192 framebuffer (tile-row ranges across GPUs)
  C5  orbiting-camera sweep of the C2 mesh
"""
from __future__ import annotations

import math
from dataclasses import dataclass, field

import numpy as np

from . import mathx
from .camera import Camera
from .mesh import Mesh, centered_quad, cube, sphere, triangle
from .texture import Texture

F = np.float32
FS_TEXTURE, FS_COLOR, FS_DEBUG = 0, 1, 2
FS_TEXTURE_BLEND = 3  # registry extension (include/rz.h): (texture.sample(u, v) + attr.color) / 2.0


def fs_with_texture(fs: int, index: int) -> int:
    """RZ_FS_WITH_TEXTURE: the sampling shaders read Uniforms::get_texture(index) (uniform.rs:35-37)."""
    return fs | (index << 8)


@dataclass
class Draw:
    mesh: Mesh
    world: np.ndarray
    fs: int = FS_TEXTURE  # shader id, optionally combined with a texture index: fs_with_texture(fs, index)


@dataclass
class Scene:
    name: str
    width: int
    height: int
    view: np.ndarray
    projection: np.ndarray
    draws: list = field(default_factory=list)
    texture: Texture | None = None          # texture 0
    extra_textures: list = field(default_factory=list)  # textures 1.. (bind order, uniform.rs:29-33)
    scissor: tuple | None = None            # (x0, y0, x1, y1), the extension sketched at rasterizer/mod.rs:349-350
    msaa: int = 4                           # samples per pixel: 4 = the reference (mod.rs:23); 1, 2, 8 = runtime extension
    guard_band: float = 1.0                 # guard band factor (mod.rs:417-419); 1 = the reference's clip planes

    @property
    def n_triangles(self) -> int:
        return sum(d.mesh.n_triangles for d in self.draws)

    @property
    def n_vertices(self) -> int:
        return sum(d.mesh.n_vertices for d in self.draws)

    def algorithmic_bytes(self) -> int:
        """Compulsory HBM traffic of one frame: 36 B/vertex + 12 B/triangle +
        3 matrices per draw + the texture once + the resolved u32 image once."""
        tex = self.texture.texels.nbytes if self.texture is not None else 0
        return 36 * self.n_vertices + 12 * self.n_triangles + 192 * len(self.draws) + tex + 4 * self.width * self.height


def default_projection(width: int, height: int) -> np.ndarray:
    """main.rs:137-142"""
    return mathx.project(1.0, 200.0, F(F(height) / F(width)), F(math.pi / 2))


def default_scene(elapsed: float = 1.0, fs: int = FS_TEXTURE, width: int = 1280, height: int = 720) -> Scene:
    """C1: Mode::Demo, main.rs:93-105 -- cube(1.0) then sphere(0.5), checkerboard texture."""
    t = F(elapsed)
    w0 = mathx.rotate(t, t, 0.0)
    w1 = mathx.matmul(mathx.rotate(t, 0.0, F(math.pi / 4)), mathx.translate(0.0, 3.0, 0.0))
    return Scene("default", width, height, Camera().get_view_matrix(), default_projection(width, height),
                 [Draw(cube(1.0), w0, fs), Draw(sphere(0.5), w1, fs)], Texture.checkerboard())


def clip_test_scene(elapsed: float = 1.0, fs: int = FS_TEXTURE, width: int = 1280, height: int = 720) -> Scene:
    """C1 variant: Mode::ClipTest, main.rs:106-125 -- one triangle following the window border."""
    t = F(elapsed)
    w = mathx.matmul(mathx.matmul(mathx.rotate_z(t), mathx.translate(7.3, 0.0, 0.0)), mathx.rotate_z(F(-t)))
    return Scene("clip_test", width, height, Camera().get_view_matrix(), default_projection(width, height),
                 [Draw(triangle(), w, fs)], Texture.checkerboard())


def sphere_scene(n_phi: int = 1001, n_theta: int = 501, radius: float = 2.0, width: int = 1920, height: int = 1080,
                 fs: int = FS_TEXTURE, camera: Camera | None = None, mesh: Mesh | None = None) -> Scene:
    """C2 (and, scaled down, its parity-test versions): UV-sphere generalising mesh.rs:152-207;
    1001 x 501 samples = 501 501 vertices, exactly 1 000 000 triangles."""
    mesh = mesh if mesh is not None else sphere(radius, n_phi, n_theta)
    cam = camera or Camera()
    return Scene(f"sphere_{mesh.n_triangles}", width, height, cam.get_view_matrix(), default_projection(width, height),
                 [Draw(mesh, mathx.rotate(0.3, 0.3, 0.0), fs)], Texture.checkerboard())


def grid_mesh(nx: int, ny: int, x0: float, x1: float, y0: float, y1: float, z: float, jitter: float = 0.0,
              seed: int = 1234) -> Mesh:
    """(nx x ny) quads in the plane z (+ per-vertex z jitter), wound like mesh.rs centered_quad."""
    rng = np.random.RandomState(seed)
    xs = np.linspace(x0, x1, nx + 1, dtype=np.float32)
    ys = np.linspace(y1, y0, ny + 1, dtype=np.float32)  # top row first (y up)
    X, Y = np.meshgrid(xs, ys)
    Z = (np.float32(z) + (rng.rand(ny + 1, nx + 1).astype(np.float32) - F(0.5)) * F(2 * jitter)).astype(np.float32)
    verts = np.stack([X, Y, Z], -1).reshape(-1, 3)
    u = np.broadcast_to(np.linspace(0, 1, nx + 1, dtype=np.float32)[None, :], X.shape)
    v = np.broadcast_to(np.linspace(0, 1, ny + 1, dtype=np.float32)[:, None], X.shape)
    attrs = np.empty((verts.shape[0], 6), np.float32)
    attrs[:, 0] = u.reshape(-1)
    attrs[:, 1] = v.reshape(-1)
    attrs[:, 2] = 0.5
    attrs[:, 3] = 1.0
    attrs[:, 4] = u.reshape(-1)
    attrs[:, 5] = v.reshape(-1)
    jj, ii = np.meshgrid(np.arange(ny), np.arange(nx), indexing="ij")
    tl = jj * (nx + 1) + ii
    tr, bl, br = tl + 1, tl + (nx + 1), tl + (nx + 2)
    idx = np.stack([tl, tr, br, tl, br, bl], -1).reshape(-1).astype(np.uint32)
    return Mesh(verts, idx, attrs)


def overdraw_scene(nx: int = 500, ny: int = 250, width: int = 1920, height: int = 1080, back_to_front: bool = True,
                   fs: int = FS_TEXTURE) -> Scene:
    """C2 secondary mesh: 4 stacked viewport-filling grids at camera depths 4..7 with z jitter,
    drawn back-to-front (worst case for the depth test) or front-to-back.  4*2*nx*ny triangles."""
    aspect = height / width
    depths = [7.0, 6.0, 5.0, 4.0] if back_to_front else [4.0, 5.0, 6.0, 7.0]
    draws = []
    for k, d in enumerate(depths):
        m = grid_mesh(nx, ny, -0.98 * d, 0.98 * d, -0.98 * d * aspect, 0.98 * d * aspect, d - 5.0, jitter=0.01,
                      seed=1234 + k)
        draws.append(Draw(m, mathx.identity(), fs))
    return Scene("overdraw", width, height, Camera().get_view_matrix(), default_projection(width, height), draws,
                 Texture.checkerboard())


def near_clip_scene(nx: int = 500, ny: int = 250, width: int = 3840, height: int = 2160, fs: int = FS_TEXTURE,
                    seed: int = 42) -> Scene:
    """C3: a field of 2*nx*ny triangles (250K by default) that ALL straddle the near plane (near = 1):
    every quad is a shingle whose top edge sits at camera depth 1.1 and whose bottom edge at
    0.9 * (1 +- 10% jitter), slightly wider than the viewport so side planes clip too."""
    rng = np.random.RandomState(seed)
    aspect = height / width
    xs = np.linspace(-1.1, 1.1, nx + 1, dtype=np.float32)
    ys = np.linspace(0.62 * aspect / 0.5625, -0.62 * aspect / 0.5625, ny + 1, dtype=np.float32)
    d_far = F(1.1)
    ii, jj = np.meshgrid(np.arange(nx), np.arange(ny), indexing="xy")
    ii, jj = ii.reshape(-1), jj.reshape(-1)
    n = ii.size
    d_near = (F(0.9) * (F(1.0) + (rng.rand(n).astype(np.float32) - F(0.5)) * F(0.2))).astype(np.float32)
    verts = np.empty((n, 4, 3), np.float32)
    # camera depth d <-> world z = d - 5 (default camera at z = -5 looking down +z)
    verts[:, 0] = np.stack([xs[ii], ys[jj], np.full(n, d_far - F(5.0), np.float32)], -1)          # TL
    verts[:, 1] = np.stack([xs[ii + 1], ys[jj], np.full(n, d_far - F(5.0), np.float32)], -1)      # TR
    verts[:, 2] = np.stack([xs[ii + 1], ys[jj + 1], d_near - F(5.0)], -1)                          # BR
    verts[:, 3] = np.stack([xs[ii], ys[jj + 1], d_near - F(5.0)], -1)                              # BL
    attrs = np.empty((n, 4, 6), np.float32)
    attrs[..., 0:3] = np.array([[1, 0, 0], [0, 0, 1], [0, 1, 0], [1, 1, 1]], np.float32)[None]
    attrs[..., 3] = 1.0
    base_u = (ii / F(nx)).astype(np.float32)[:, None]
    base_v = (jj / F(ny)).astype(np.float32)[:, None]
    attrs[..., 4] = base_u + np.array([0, 1, 1, 0], np.float32)[None] / F(nx)
    attrs[..., 5] = base_v + np.array([0, 0, 1, 1], np.float32)[None] / F(ny)
    np.clip(attrs[..., 4:6], 0.0, 1.0, out=attrs[..., 4:6])
    q = (np.arange(n, dtype=np.uint32) * 4)[:, None]
    idx = (q + np.array([0, 1, 2, 0, 2, 3], np.uint32)[None]).reshape(-1)
    mesh = Mesh(verts.reshape(-1, 3), idx, attrs.reshape(-1, 6))
    return Scene("near_clip", width, height, Camera().get_view_matrix(), default_projection(width, height),
                 [Draw(mesh, mathx.identity(), fs)], Texture.checkerboard())


def fullscreen_quad_scene(width: int = 8192, height: int = 8192, fs: int = FS_TEXTURE) -> Scene:
    """C4 (ii): a 2-triangle quad (mesh.rs centered_quad, z = 2 => camera depth 7) that covers the
    whole square viewport and is clipped by all four side planes."""
    aspect = height / width
    side = 2.0 * 7.0 * max(1.0, aspect) * 1.05
    return Scene("fullscreen_quad", width, height, Camera().get_view_matrix(), default_projection(width, height),
                 [Draw(centered_quad(side), mathx.identity(), fs)], Texture.checkerboard())


def orbit_cameras(n_frames: int = 1024, radius: float = 5.0):
    """C5: frame k looks at the origin from angle 2*pi*k/n on a circle in the xz-plane."""
    return [Camera.orbit(2.0 * math.pi * k / n_frames, radius) for k in range(n_frames)]


# ---------------------------------------------------------------------------------------------
def render_scene(renderer, scene: Scene, device_meshes=None):
    """Issue a Scene through the reference-shaped API (main.rs:135-145,170-173).  Works with the CUDA
    Renderer (rusterizer_b200.render) and with any object of the same surface."""
    u = renderer.uniforms()
    blk = u.write_block()
    blk.view = scene.view
    blk.projection = scene.projection
    for k, d in enumerate(scene.draws):
        blk.world = d.world
        renderer.render(device_meshes[k] if device_meshes else d.mesh, 0, d.fs)
