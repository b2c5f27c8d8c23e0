"""rusterizer_b200 -- B200-native (sm_100a CUDA) implementation of rusterizer's per-frame raster path.

Host-side mirror of the reference crate's render/shader API surface:

    Renderer, Uniforms           <- render.rs / uniform.rs   (calls the C ABI in include/rz.h)
    Mesh + generators            <- mesh.rs
    Camera                       <- camera.rs
    Texture                      <- texture.rs
    mathx                        <- math/

The compute path is the CUDA library `rusterizer_b200/librz_b200.so`; there is no CPU fallback --
constructing a Renderer without that library (or without a GPU) raises.
"""
from . import mathx  # noqa: F401
from .camera import Camera  # noqa: F401
from .mesh import Mesh, centered_quad, cube, sphere, triangle  # noqa: F401
from .texture import Texture  # noqa: F401

FS_TEXTURE, FS_COLOR, FS_DEBUG = 0, 1, 2  # `enum FS`, main.rs:23-27
FS_TEXTURE_BLEND = 3  # registry extension, see include/rz.h (RZ_FS_WITH_TEXTURE adds a texture index)
VS_MVP = 0  # the crate's only vertex shader, main.rs:147-152


def __getattr__(name):  # lazy: importing the package must not need the GPU library
    if name in ("Renderer", "Uniforms", "load_library", "RzError"):
        from . import render

        return getattr(render, name)
    raise AttributeError(name)
