"""Host-side mirror of the reference's `render.rs` / `uniform.rs` over the C ABI (include/rz.h).

    Renderer.new(width, height)            render.rs:48
    Renderer.uniforms() -> Uniforms        render.rs:71
    Uniforms.write_block()/read_block()    uniform.rs:39-45
    Uniforms.bind_texture(index, tex)      uniform.rs:29-33
    Renderer.render(mesh, vs, fs)          render.rs:98-103   (vs / fs are shader IDs here)
    Renderer.framebuffer()                 rasterizer/mod.rs:520 via Renderer::display render.rs:121

The compute path is the CUDA library only; `load_library()` raises if it is missing and
`Renderer()` raises if no GPU is present.  Nothing here imports the CPU oracle.
"""
from __future__ import annotations

import ctypes as C
import os
from pathlib import Path

import numpy as np

from .mesh import Mesh
from .texture import Texture

_HERE = Path(__file__).resolve().parent
LIB_NAME = "librz_b200.so"

COUNTER_FIELDS = (
    "n_tris_in", "n_degenerate", "n_outside", "n_inside", "n_clipped_in",
    "n_tris_setup", "n_bbox_px", "n_covered_px", "n_shaded_px", "n_samples_written",
    "n_tex_oob", "n_clip_overflow",
)

# every symbol include/rz.h declares (tests check the library exports all of them)
ABI_SYMBOLS = (
    "rz_create", "rz_destroy", "rz_set_stream", "rz_bind_texture", "rz_write_block", "rz_read_block",
    "rz_mesh_create", "rz_mesh_destroy", "rz_render", "rz_render_host", "rz_framebuffer",
    "rz_framebuffer_async", "rz_framebuffer_host_async", "rz_sync", "rz_discard_frame", "rz_shared_alloc", "rz_shared_open",
    "rz_shared_close", "rz_shared_free", "rz_signal", "rz_wait_flags", "rz_set_row_range", "rz_set_row_interleave", "rz_set_scissor", "rz_set_msaa", "rz_set_guard_band", "rz_tile_width", "rz_tile_height",
    "rz_counters", "rz_reset_counters", "rz_timings", "rz_launch_count", "rz_debug_capture",
    "rz_debug_read", "rz_debug_tile_times", "rz_debug_vertex_stage", "rz_last_error", "rz_version",
)

ERROR_NAMES = {0: "RZ_OK", -1: "RZ_E_INVALID", -2: "RZ_E_CUDA", -3: "RZ_E_NO_DEVICE", -4: "RZ_E_TEXTURE",
               -5: "RZ_E_INDEX", -6: "RZ_E_CAPACITY", -7: "RZ_E_NOMEM", -8: "RZ_E_PEER"}


class RzError(RuntimeError):
    def __init__(self, code: int, msg: str):
        super().__init__(f"{ERROR_NAMES.get(code, code)}: {msg}")
        self.code = code


class CountersT(C.Structure):
    _fields_ = [(n, C.c_uint64) for n in COUNTER_FIELDS]


class TimingsT(C.Structure):
    _fields_ = [("geometry_ms", C.c_float), ("bin_ms", C.c_float), ("tile_ms", C.c_float), ("total_ms", C.c_float)]


_lib = None


def library_path() -> Path:
    return Path(os.environ.get("RZ_B200_LIB", _HERE / LIB_NAME))


def load_library() -> C.CDLL:
    """Load librz_b200.so and declare the C ABI.  Raises (loudly) when the CUDA library has not
    been built -- there is no fallback path."""
    global _lib
    if _lib is not None:
        return _lib
    path = library_path()
    if not path.exists():
        raise FileNotFoundError(
            f"{path} not found: build the CUDA library first (python -c 'import __graft_entry__ as g; g.build()'). "
            "rusterizer_b200 has no CPU fallback.")
    L = C.CDLL(str(path))
    vp, fp, u32p, u8p = C.c_void_p, C.POINTER(C.c_float), C.POINTER(C.c_uint32), C.POINTER(C.c_uint8)
    L.rz_create.argtypes = [C.c_int, C.c_uint32, C.c_uint32, C.POINTER(vp)]
    L.rz_destroy.argtypes = [vp]
    L.rz_destroy.restype = None
    L.rz_set_stream.argtypes = [vp, vp]
    L.rz_bind_texture.argtypes = [vp, C.c_uint32, u8p, C.c_uint32, C.c_uint32, C.c_uint32]
    L.rz_write_block.argtypes = [vp, fp, fp, fp]
    L.rz_read_block.argtypes = [vp, fp, fp, fp]
    L.rz_mesh_create.argtypes = [vp, vp, vp, C.c_uint32, vp, C.c_uint64, C.POINTER(vp)]
    L.rz_mesh_destroy.argtypes = [vp]
    L.rz_mesh_destroy.restype = None
    L.rz_render.argtypes = [vp, vp, C.c_uint32, C.c_uint32]
    L.rz_render_host.argtypes = [vp, vp, vp, C.c_uint32, vp, C.c_uint64, C.c_uint32, C.c_uint32]
    L.rz_framebuffer.argtypes = [vp, vp, C.POINTER(vp)]
    L.rz_framebuffer_async.argtypes = [vp, vp, C.POINTER(vp)]
    L.rz_framebuffer_host_async.argtypes = [vp, vp]
    L.rz_sync.argtypes = [vp]
    L.rz_discard_frame.argtypes = [vp]
    L.rz_shared_alloc.argtypes = [vp, C.c_uint64, C.POINTER(vp), u8p]
    L.rz_shared_open.argtypes = [vp, u8p, C.POINTER(vp)]
    L.rz_shared_close.argtypes = [vp, vp]
    L.rz_shared_free.argtypes = [vp, vp]
    L.rz_signal.argtypes = [vp, C.POINTER(vp), C.c_uint32, C.c_uint32]
    L.rz_wait_flags.argtypes = [vp, vp, C.c_uint32, C.c_uint32, C.c_uint32, C.c_uint32]
    L.rz_set_row_range.argtypes = [vp, C.c_uint32, C.c_uint32]
    L.rz_set_row_interleave.argtypes = [vp, C.c_uint32, C.c_uint32, C.c_uint32]
    L.rz_set_scissor.argtypes = [vp, C.c_uint32, C.c_uint32, C.c_uint32, C.c_uint32]
    L.rz_set_msaa.argtypes = [vp, C.c_uint32]
    L.rz_set_guard_band.argtypes = [vp, C.c_float]
    L.rz_tile_width.restype = C.c_uint32
    L.rz_tile_height.restype = C.c_uint32
    L.rz_counters.argtypes = [vp, C.POINTER(CountersT)]
    L.rz_reset_counters.argtypes = [vp]
    L.rz_timings.argtypes = [vp, C.POINTER(TimingsT)]
    L.rz_launch_count.argtypes = [vp]
    L.rz_launch_count.restype = C.c_uint64
    L.rz_debug_capture.argtypes = [vp, C.c_int]
    L.rz_debug_read.argtypes = [vp, vp, vp, vp]
    L.rz_debug_tile_times.argtypes = [vp, vp, C.c_uint32, C.POINTER(C.c_uint32)]
    L.rz_debug_vertex_stage.argtypes = [vp, vp, fp]
    L.rz_last_error.argtypes = [vp]
    L.rz_last_error.restype = C.c_char_p
    L.rz_version.restype = C.c_char_p
    _lib = L
    return L


def _f32(a):
    return np.ascontiguousarray(a, dtype=np.float32)


def _fp(a):
    return a.ctypes.data_as(C.POINTER(C.c_float))


class DeviceMesh:
    """A Mesh uploaded once to HBM (rz_mesh_create)."""

    def __init__(self, renderer: "Renderer", mesh: Mesh):
        self._r = renderer
        self.n_vertices, self.n_triangles = mesh.n_vertices, mesh.n_triangles
        self.handle = C.c_void_p()
        L = renderer._L
        rc = L.rz_mesh_create(renderer._ctx, mesh.vertices.ctypes.data, mesh.attributes.ctypes.data, mesh.n_vertices,
                              mesh.indices.ctypes.data, mesh.indices.size, C.byref(self.handle))
        renderer._check(rc)

    def close(self):
        if self.handle:
            self._r._L.rz_mesh_destroy(self.handle)
            self.handle = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class UniformBlock:
    """uniform.rs:4-9: assigning a field writes through to the ctx."""

    def __init__(self, renderer: "Renderer"):
        object.__setattr__(self, "_r", renderer)

    def _read(self):
        w, v, p = (np.empty(16, np.float32) for _ in range(3))
        self._r._check(self._r._L.rz_read_block(self._r._ctx, _fp(w), _fp(v), _fp(p)))
        return {"world": w.reshape(4, 4), "view": v.reshape(4, 4), "projection": p.reshape(4, 4)}

    def __getattr__(self, name):
        if name in ("world", "view", "projection"):
            return self._read()[name]
        raise AttributeError(name)

    def __setattr__(self, name, value):
        if name not in ("world", "view", "projection"):
            raise AttributeError(name)
        m = _f32(value).reshape(16)
        args = {"world": None, "view": None, "projection": None}
        args[name] = _fp(m)
        self._r._check(self._r._L.rz_write_block(self._r._ctx, args["world"], args["view"], args["projection"]))


class Uniforms:
    """uniform.rs:11-46"""

    def __init__(self, renderer: "Renderer"):
        self._r = renderer
        self._block = UniformBlock(renderer)
        self._textures: list[Texture] = []

    def write_block(self) -> UniformBlock:
        return self._block

    def read_block(self) -> UniformBlock:
        return self._block

    def bind_texture(self, index: int, tex: Texture):
        t = tex.texels
        rc = self._r._L.rz_bind_texture(self._r._ctx, index, t.ctypes.data_as(C.POINTER(C.c_uint8)), tex.width,
                                        tex.height, tex.texel_width)
        self._r._check(rc)
        self._textures.append(tex)

    def get_texture(self, index: int) -> Texture:
        return self._textures[index]


class Renderer:
    """render.rs:38-127 without the minifb window: `framebuffer()` returns the u32 image that
    `display()` would hand to `update_with_buffer` (render.rs:121-124)."""

    def __init__(self, width: int, height: int, device: int = 0):
        self._L = load_library()
        self.width, self.height = int(width), int(height)
        self._ctx = C.c_void_p()
        rc = self._L.rz_create(device, self.width, self.height, C.byref(self._ctx))
        if rc != 0:
            raise RzError(rc, (self._L.rz_last_error(None) or b"").decode())
        self._uniforms = Uniforms(self)
        self._keepalive: list = []
        self.msaa = 4

    @classmethod
    def new(cls, width: int, height: int, device: int = 0) -> "Renderer":
        return cls(width, height, device)

    def _check(self, rc: int):
        if rc != 0:
            raise RzError(rc, (self._L.rz_last_error(self._ctx) or b"").decode())

    def close(self):
        if getattr(self, "_ctx", None):
            self._L.rz_destroy(self._ctx)
            self._ctx = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # ---- reference surface ----
    def uniforms(self) -> Uniforms:
        return self._uniforms

    def upload(self, mesh: Mesh) -> DeviceMesh:
        return DeviceMesh(self, mesh)

    def render(self, mesh, vertex_shader: int = 0, fragment_shader: int = 0):
        """Renderer::render (render.rs:98-114).  `mesh` is a host Mesh (copied to the device during
        the call, like the reference borrowing &mesh) or a DeviceMesh uploaded earlier."""
        if isinstance(mesh, DeviceMesh):
            self._check(self._L.rz_render(self._ctx, mesh.handle, vertex_shader, fragment_shader))
            self._keepalive.append(mesh)
        else:
            self._check(self._L.rz_render_host(self._ctx, mesh.vertices.ctypes.data, mesh.attributes.ctypes.data,
                                               mesh.n_vertices, mesh.indices.ctypes.data, mesh.indices.size,
                                               vertex_shader, fragment_shader))

    def render_arrays(self, pos_ptr: int, attr_ptr: int, nv: int, idx_ptr: int, n_idx: int, vs: int = 0, fs: int = 0):
        """rz_render_host on raw host pointers (e.g. pinned torch tensors)."""
        self._check(self._L.rz_render_host(self._ctx, pos_ptr, attr_ptr, nv, idx_ptr, n_idx, vs, fs))

    def framebuffer(self, out: np.ndarray | None = None) -> np.ndarray:
        """Rasterizer::framebuffer (rasterizer/mod.rs:520-522): resolve + clear; returns u32[H][W]."""
        if out is None:
            out = np.empty((self.height, self.width), np.uint32)
        assert out.dtype == np.uint32 and out.size == self.width * self.height and out.flags.c_contiguous
        self._check(self._L.rz_framebuffer(self._ctx, out.ctypes.data, None))
        self._keepalive.clear()
        return out

    def framebuffer_into(self, host_ptr: int):
        self._check(self._L.rz_framebuffer(self._ctx, host_ptr, None))
        self._keepalive.clear()

    def framebuffer_device(self) -> int:
        """Synchronous frame without the D2H copy; returns the device pointer of the u32 image."""
        p = C.c_void_p()
        self._check(self._L.rz_framebuffer(self._ctx, None, C.byref(p)))
        self._keepalive.clear()
        return p.value

    def framebuffer_async(self, device_dst: int | None = None) -> int:
        p = C.c_void_p()
        self._check(self._L.rz_framebuffer_async(self._ctx, device_dst, C.byref(p)))
        return p.value

    def framebuffer_host_async(self, host_ptr: int):
        """Streaming display: frame + D2H of the image on a copy stream, no host sync (valid after sync())."""
        self._check(self._L.rz_framebuffer_host_async(self._ctx, host_ptr))

    def sync(self):
        self._check(self._L.rz_sync(self._ctx))
        self._keepalive.clear()

    def discard_frame(self):
        """Drop the draws recorded for the current frame without executing them."""
        self._check(self._L.rz_discard_frame(self._ctx))
        self._keepalive.clear()

    # ---- peer memory (screen-space sharding over NVLink, see include/rz.h) ----
    def shared_alloc(self, nbytes: int) -> tuple[int, bytes]:
        p, h = C.c_void_p(), (C.c_uint8 * 64)()
        self._check(self._L.rz_shared_alloc(self._ctx, nbytes, C.byref(p), h))
        return p.value, bytes(h)

    def shared_open(self, handle: bytes) -> int:
        p, h = C.c_void_p(), (C.c_uint8 * 64).from_buffer_copy(handle)
        self._check(self._L.rz_shared_open(self._ctx, h, C.byref(p)))
        return p.value

    def shared_close(self, ptr: int):
        self._check(self._L.rz_shared_close(self._ctx, ptr))

    def shared_free(self, ptr: int):
        self._check(self._L.rz_shared_free(self._ctx, ptr))

    def signal(self, flag_ptrs, value: int):
        """Raise up to 16 (local or peer-mapped) flags to `value` once all earlier work of the stream is done."""
        ptrs = [flag_ptrs] if isinstance(flag_ptrs, int) else list(flag_ptrs)
        arr = (C.c_void_p * len(ptrs))(*ptrs)
        self._check(self._L.rz_signal(self._ctx, arr, len(ptrs), value & 0xFFFFFFFF))

    def wait_flags(self, flags_ptr: int, n: int, stride_bytes: int, value: int, timeout_ms: int = 0):
        self._check(self._L.rz_wait_flags(self._ctx, flags_ptr, n, stride_bytes, value & 0xFFFFFFFF, timeout_ms))

    # ---- extras ----
    def set_stream(self, cuda_stream: int | None):
        self._check(self._L.rz_set_stream(self._ctx, cuda_stream))

    def set_row_range(self, row_begin: int, row_end: int):
        self._check(self._L.rz_set_row_range(self._ctx, row_begin, row_end))

    def set_scissor(self, x0: int, y0: int, x1: int, y1: int):
        """Scissor rect [x0,x1) x [y0,y1) (the extension sketched at rasterizer/mod.rs:349-350)."""
        self._check(self._L.rz_set_scissor(self._ctx, x0, y0, x1, y1))

    def set_msaa(self, samples: int):
        """Samples per pixel: 1, 2, 4 (the reference, rasterizer/mod.rs:23) or 8."""
        self._check(self._L.rz_set_msaa(self._ctx, samples))
        self.msaa = int(samples)

    def set_guard_band(self, factor: float):
        """Guard-band clipping (rasterizer/mod.rs:417-419): side clip planes at |x|, |y| <= factor * w."""
        self._check(self._L.rz_set_guard_band(self._ctx, factor))

    def set_row_interleave(self, band_tile_rows: int, rank: int, world: int):
        self._check(self._L.rz_set_row_interleave(self._ctx, band_tile_rows, rank, world))

    def counters(self) -> dict:
        c = CountersT()
        self._check(self._L.rz_counters(self._ctx, C.byref(c)))
        return {n: int(getattr(c, n)) for n in COUNTER_FIELDS}

    def reset_counters(self):
        self._check(self._L.rz_reset_counters(self._ctx))

    def timings(self) -> dict:
        t = TimingsT()
        self._check(self._L.rz_timings(self._ctx, C.byref(t)))
        return {k: float(getattr(t, k)) for k, _ in TimingsT._fields_}

    def launch_count(self) -> int:
        return int(self._L.rz_launch_count(self._ctx))

    def debug_capture(self, enable: bool = True):
        self._check(self._L.rz_debug_capture(self._ctx, 1 if enable else 0))

    def debug_read(self):
        shape = (self.height, self.width, self.msaa)
        d, c, o = np.empty(shape, np.float32), np.empty(shape, np.uint32), np.empty(shape, np.uint32)
        self._check(self._L.rz_debug_read(self._ctx, d.ctypes.data, c.ctypes.data, o.ctypes.data))
        return d, c, o

    def tile_times(self) -> np.ndarray:
        """(n_tiles, 8) u64 rows, see rz_debug_tile_times in include/rz.h (zeros for unused rows)."""
        n = ((self.width + 15) // 16) * ((self.height + 15) // 16)
        out = np.zeros((n, 8), np.uint64)
        got = C.c_uint32()
        self._check(self._L.rz_debug_tile_times(self._ctx, out.ctypes.data, n, C.byref(got)))
        return out[: got.value]

    def vertex_stage(self, mesh: DeviceMesh) -> np.ndarray:
        out = np.empty((mesh.n_vertices, 4), np.float32)
        self._check(self._L.rz_debug_vertex_stage(self._ctx, mesh.handle, _fp(out)))
        return out


def tile_size() -> tuple[int, int]:
    L = load_library()
    return int(L.rz_tile_width()), int(L.rz_tile_height())
