"""ctypes wrapper around the CPU oracle (oracle/rz_oracle.c).

TEST INFRASTRUCTURE ONLY: imported by tests/, __graft_entry__.smoke() and the
cpu_baseline / --impl reference legs of bench.py.  The product package
(rusterizer_b200/) never imports this module.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess
from pathlib import Path

import numpy as np

_HERE = Path(__file__).resolve().parent

COUNTER_FIELDS = (
    "n_tris_in", "n_degenerate", "n_outside", "n_inside", "n_clipped_in",
    "n_tris_setup", "n_bbox_px", "n_covered_px", "n_shaded_px", "n_samples_written",
    "n_tex_oob", "n_clip_overflow",
)


class Counters(C.Structure):
    _fields_ = [(n, C.c_uint64) for n in COUNTER_FIELDS]

    def as_dict(self):
        return {n: int(getattr(self, n)) for n in COUNTER_FIELDS}


def build(fast: bool = False, out_dir: os.PathLike | None = None) -> Path:
    """Compile the oracle with gcc.  fast=True adds -O3 -march=native (timing build; must be
    compiled on the machine that runs it)."""
    out_dir = Path(out_dir) if out_dir else _HERE / "_build"
    out_dir.mkdir(parents=True, exist_ok=True)
    out = out_dir / ("liboracle_fast.so" if fast else "liboracle.so")
    src = _HERE / "rz_oracle.c"
    if out.exists() and out.stat().st_mtime >= src.stat().st_mtime:
        return out
    opt = ["-O3", "-march=native"] if fast else ["-O2"]
    cmd = ["gcc", *opt, "-fPIC", "-shared", "-std=c11", "-ffp-contract=off", "-fno-fast-math",
           "-o", str(out), str(src), "-lm"]
    subprocess.run(cmd, check=True)
    return out


def _f32p(a):
    return a.ctypes.data_as(C.POINTER(C.c_float))


def _u32p(a):
    return a.ctypes.data_as(C.POINTER(C.c_uint32))


def _u8p(a):
    return a.ctypes.data_as(C.POINTER(C.c_uint8))


def f32(a):
    return np.ascontiguousarray(a, dtype=np.float32)


class OracleLib:
    def __init__(self, path: os.PathLike | None = None, fast: bool = False):
        self.path = Path(path) if path else build(fast=fast)
        L = self.lib = C.CDLL(str(self.path))
        fp, u32p, u8p = C.POINTER(C.c_float), C.POINTER(C.c_uint32), C.POINTER(C.c_uint8)
        u64p = C.POINTER(C.c_uint64)
        L.orc_create.restype = C.c_void_p
        L.orc_create.argtypes = [C.c_uint32, C.c_uint32]
        L.orc_create_rows.restype = C.c_void_p
        L.orc_create_rows.argtypes = [C.c_uint32, C.c_uint32, C.c_uint32, C.c_uint32]
        L.orc_destroy.argtypes = [C.c_void_p]
        L.orc_bind_texture.argtypes = [C.c_void_p, C.c_uint32, u8p, C.c_uint32, C.c_uint32, C.c_uint32]
        L.orc_write_block.argtypes = [C.c_void_p, fp, fp, fp]
        L.orc_set_scissor.argtypes = [C.c_void_p, C.c_uint32, C.c_uint32, C.c_uint32, C.c_uint32]
        L.orc_set_msaa.argtypes = [C.c_void_p, C.c_uint32]
        L.orc_debug_asserts.argtypes = [u64p, C.c_int]
        L.orc_set_guard_band.argtypes = [C.c_void_p, C.c_float]
        L.orc_render.argtypes = [C.c_void_p, fp, fp, C.c_uint32, u32p, C.c_uint64, C.c_uint32, C.c_uint32]
        L.orc_rasterize.argtypes = [C.c_void_p, fp, fp, C.c_uint64, C.c_uint32]
        L.orc_vertex_stage.argtypes = [C.c_void_p, fp, C.c_uint32, fp]
        L.orc_framebuffer.restype = u32p
        L.orc_framebuffer.argtypes = [C.c_void_p]
        L.orc_depth_samples.restype = fp
        L.orc_depth_samples.argtypes = [C.c_void_p]
        L.orc_color_samples.restype = u32p
        L.orc_color_samples.argtypes = [C.c_void_p]
        L.orc_owner_samples.restype = u32p
        L.orc_owner_samples.argtypes = [C.c_void_p]
        L.orc_counters.argtypes = [C.c_void_p, C.POINTER(Counters)]
        L.orc_reset_counters.argtypes = [C.c_void_p]
        L.orc_mat4_mul.argtypes = [fp, fp, fp]
        L.orc_mat4_vec.argtypes = [fp, fp, fp]
        L.orc_triangle_2x_area.restype = C.c_float
        L.orc_triangle_2x_area.argtypes = [fp]
        L.orc_to_argb.restype = C.c_uint32
        L.orc_to_argb.argtypes = [fp]
        L.orc_box_filter_color.restype = C.c_uint32
        L.orc_box_filter_color.argtypes = [u32p]
        L.orc_try_clip.restype = C.c_int
        L.orc_try_clip.argtypes = [fp, fp, fp, fp, C.c_int]
        L.orc_perspective_divide.argtypes = [fp, fp]
        L.orc_pixel_bbox.argtypes = [fp, u64p]
        L.orc_viewport_setup.argtypes = [C.c_uint32, C.c_uint32, fp, fp, fp, fp, fp, fp]
        L.orc_eval_pixel.argtypes = [fp, fp, fp, C.c_uint64, C.c_uint64, C.c_uint32, u32p, fp, fp, fp]
        L.orc_eval_single.restype = C.c_int
        L.orc_eval_single.argtypes = [fp, C.c_float, C.c_float, fp, fp]
        L.orc_tex_sample.restype = C.c_uint64
        L.orc_tex_sample.argtypes = [u8p, C.c_uint32, C.c_uint32, C.c_uint32, C.c_float, C.c_float, fp]
        L.orc_tile_grid.argtypes = [C.c_uint32, C.c_uint32, u32p, u32p]
        L.orc_tile_idx.restype = C.c_uint32
        L.orc_tile_idx.argtypes = [C.c_uint32, C.c_uint32, C.c_uint32]
        L.orc_tiles_marked.restype = C.c_uint32
        L.orc_tiles_marked.argtypes = [C.c_void_p, C.c_int]

    # ---- unit-level helpers (KATs) ----
    def mat4_mul(self, a, b):
        a, b = f32(a).reshape(16), f32(b).reshape(16)
        r = np.empty(16, np.float32)
        self.lib.orc_mat4_mul(_f32p(a), _f32p(b), _f32p(r))
        return r.reshape(4, 4)

    def mat4_vec(self, m, v):
        m, v = f32(m).reshape(16), f32(v).reshape(4)
        r = np.empty(4, np.float32)
        self.lib.orc_mat4_vec(_f32p(m), _f32p(v), _f32p(r))
        return r

    def triangle_2x_area(self, xy):
        xy = f32(xy).reshape(6)
        return np.float32(self.lib.orc_triangle_2x_area(_f32p(xy)))

    def to_argb(self, rgba):
        rgba = f32(rgba).reshape(4)
        return int(self.lib.orc_to_argb(_f32p(rgba)))

    def box_filter_color(self, colors):
        c = np.ascontiguousarray(colors, dtype=np.uint32).reshape(4)
        return int(self.lib.orc_box_filter_color(_u32p(c)))

    def try_clip(self, pos, attrs=None, cap=16):
        """pos (3,4) clip-space; returns (kind, tris_pos (n,3,4), tris_attrs (n,3,6));
        kind: 'outside' | 'inside' | 'clipped'."""
        pos = f32(pos).reshape(12)
        attrs = f32(np.zeros((3, 6)) if attrs is None else attrs).reshape(18)
        op = np.zeros(cap * 12, np.float32)
        oa = np.zeros(cap * 18, np.float32)
        n = self.lib.orc_try_clip(_f32p(pos), _f32p(attrs), _f32p(op), _f32p(oa), cap)
        if n < 0:
            return "outside", op[:0].reshape(0, 3, 4), oa[:0].reshape(0, 3, 6)
        if n == 0:
            return "inside", op[:0].reshape(0, 3, 4), oa[:0].reshape(0, 3, 6)
        return "clipped", op[: n * 12].reshape(n, 3, 4), oa[: n * 18].reshape(n, 3, 6)

    def perspective_divide(self, clip):
        clip = f32(clip).reshape(12)
        out = np.empty(12, np.float32)
        self.lib.orc_perspective_divide(_f32p(clip), _f32p(out))
        return out.reshape(3, 4)

    def pixel_bbox(self, xy):
        xy = f32(xy).reshape(6)
        out = np.zeros(4, np.uint64)
        self.lib.orc_pixel_bbox(_f32p(xy), out.ctypes.data_as(C.POINTER(C.c_uint64)))
        return tuple(int(v) for v in out)  # min_x, max_x, min_y, max_y

    def viewport_setup(self, width, height, ndc):
        ndc = f32(ndc).reshape(12)
        pts, nrm = np.empty(6, np.float32), np.empty(6, np.float32)
        z, w, inv = np.empty(3, np.float32), np.empty(3, np.float32), np.empty(1, np.float32)
        self.lib.orc_viewport_setup(width, height, _f32p(ndc), _f32p(pts), _f32p(nrm), _f32p(z), _f32p(w), _f32p(inv))
        return dict(points=pts.reshape(3, 2), normals=nrm.reshape(3, 2), depths=z, depths_camera_space=w,
                    inv_2x_area=inv[0])

    def eval_pixel(self, screen, w, attrs, x, y, interp_mask=0xFF):
        screen, w, attrs = f32(screen).reshape(9), f32(w).reshape(3), f32(attrs).reshape(18)
        mask = np.zeros(1, np.uint32)
        ev, d, a = np.empty(12, np.float32), np.empty(4, np.float32), np.empty(6, np.float32)
        self.lib.orc_eval_pixel(_f32p(screen), _f32p(w), _f32p(attrs), x, y, interp_mask, _u32p(mask), _f32p(ev),
                                _f32p(d), _f32p(a))
        return dict(mask=int(mask[0]), evals=ev.reshape(4, 3), depths=d, attr=a)

    def eval_single(self, screen, x, y):
        screen = f32(screen).reshape(9)
        e, n = np.empty(3, np.float32), np.empty(6, np.float32)
        ins = self.lib.orc_eval_single(_f32p(screen), C.c_float(x), C.c_float(y), _f32p(e), _f32p(n))
        return bool(ins), e, n.reshape(3, 2)

    def tex_sample(self, texels, u, v):
        t = np.ascontiguousarray(texels, dtype=np.uint8)
        h, w, tw = t.shape
        out = np.empty(4, np.float32)
        oob = self.lib.orc_tex_sample(_u8p(t), w, h, tw, C.c_float(u), C.c_float(v), _f32p(out))
        return out, int(oob)

    def tile_grid(self, width, height):
        a, b = np.zeros(1, np.uint32), np.zeros(1, np.uint32)
        self.lib.orc_tile_grid(width, height, _u32p(a), _u32p(b))
        return int(a[0]), int(b[0])

    def tile_idx(self, width, row, col):
        return int(self.lib.orc_tile_idx(width, row, col))


DEBUG_ASSERTS = ("clamp_bary_range (mod.rs:104)", "ndc_range (mod.rs:319-321)", "z_range (mod.rs:329)",
                 "texture_uv_range (texture.rs:66-67)", "texel_xy_range (texture.rs:49-50)")


def debug_asserts(lib: "OracleLib | None" = None, reset: bool = True) -> dict:
    """How often each range-guarding debug_assert! of the reference would have fired in a debug build since the last
    reset (the oracle computes the --release behaviour either way).  Per process and per library instance."""
    lib = lib or get_lib()
    out = (C.c_uint64 * 5)()
    lib.lib.orc_debug_asserts(out, 1 if reset else 0)
    return dict(zip(DEBUG_ASSERTS, (int(v) for v in out)))


_LIBS: dict = {}


def get_lib(fast: bool = False) -> OracleLib:
    if fast not in _LIBS:
        _LIBS[fast] = OracleLib(fast=fast)
    return _LIBS[fast]


class OracleRenderer:
    """The reference's Renderer/Rasterizer surface (render.rs:47-127) on the CPU oracle."""

    def __init__(self, width: int, height: int, lib: OracleLib | None = None, rows: tuple | None = None):
        """rows=(r0, r1): hold (and return) only the pixel rows [r0, r1) of the frame -- a large frame is checked band
        by band, one process per band (see oracle.banded_render)."""
        self.L = lib or get_lib()
        self.width, self.height = width, height
        self.row0, self.row1 = (0, height) if rows is None else (int(rows[0]), min(int(rows[1]), height))
        self.ctx = self.L.lib.orc_create_rows(width, height, self.row0, self.row1)
        self.ns = 4
        if not self.ctx:
            raise MemoryError("orc_create failed")

    def close(self):
        if self.ctx:
            self.L.lib.orc_destroy(self.ctx)
            self.ctx = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def bind_texture(self, index: int, texels: np.ndarray):
        t = np.ascontiguousarray(texels, dtype=np.uint8)
        h, w, tw = t.shape
        rc = self.L.lib.orc_bind_texture(self.ctx, index, _u8p(t), w, h, tw)
        if rc != 0:
            raise ValueError(f"orc_bind_texture failed: {rc}")

    def set_scissor(self, x0, y0, x1, y1):
        self.L.lib.orc_set_scissor(self.ctx, x0, y0, x1, y1)

    def set_msaa(self, n: int):
        """Samples per pixel (1, 2, 4, 8); the reference's N_MSAA_SAMPLES = 4 is the default."""
        if self.L.lib.orc_set_msaa(self.ctx, n) != 0:
            raise ValueError(f"orc_set_msaa({n})")
        self.ns = n

    def set_guard_band(self, g: float):
        if self.L.lib.orc_set_guard_band(self.ctx, g) != 0:
            raise ValueError(f"orc_set_guard_band({g})")

    def write_block(self, world=None, view=None, projection=None):
        arrs = [None if m is None else f32(m).reshape(16) for m in (world, view, projection)]
        ptrs = [None if a is None else _f32p(a) for a in arrs]
        self.L.lib.orc_write_block(self.ctx, *ptrs)

    def render(self, pos, attrs, idx, vs_id=0, fs_id=0):
        pos, attrs = f32(pos).reshape(-1, 3), f32(attrs).reshape(-1, 6)
        idx = np.ascontiguousarray(idx, dtype=np.uint32).reshape(-1)
        assert pos.shape[0] == attrs.shape[0]
        rc = self.L.lib.orc_render(self.ctx, _f32p(pos), _f32p(attrs), pos.shape[0], _u32p(idx), idx.size, vs_id, fs_id)
        if rc != 0:
            raise RuntimeError(f"orc_render failed: {rc}")

    def rasterize(self, clip_pos, attrs, fs_id=0):
        clip_pos, attrs = f32(clip_pos).reshape(-1, 3, 4), f32(attrs).reshape(-1, 3, 6)
        rc = self.L.lib.orc_rasterize(self.ctx, _f32p(clip_pos), _f32p(attrs), clip_pos.shape[0], fs_id)
        if rc != 0:
            raise RuntimeError(f"orc_rasterize failed: {rc}")

    def vertex_stage(self, pos):
        pos = f32(pos).reshape(-1, 3)
        out = np.empty((pos.shape[0], 4), np.float32)
        self.L.lib.orc_vertex_stage(self.ctx, _f32p(pos), pos.shape[0], _f32p(out))
        return out

    def _view(self, ptr, dtype, per_px):
        rows = self.row1 - self.row0
        n = self.width * rows * per_px
        return np.ctypeslib.as_array(ptr, shape=(n,)).view(dtype).reshape(rows, self.width, per_px).copy()

    def depth_samples(self):
        return self._view(self.L.lib.orc_depth_samples(self.ctx), np.float32, self.ns)

    def color_samples(self):
        return self._view(self.L.lib.orc_color_samples(self.ctx), np.uint32, self.ns)

    def owner_samples(self):
        return self._view(self.L.lib.orc_owner_samples(self.ctx), np.uint32, self.ns)

    def framebuffer(self):
        p = self.L.lib.orc_framebuffer(self.ctx)
        rows = self.row1 - self.row0
        return np.ctypeslib.as_array(p, shape=(rows * self.width,)).reshape(rows, self.width).copy()

    def counters(self):
        c = Counters()
        self.L.lib.orc_counters(self.ctx, C.byref(c))
        return c.as_dict()

    def reset_counters(self):
        self.L.lib.orc_reset_counters(self.ctx)

    def tiles_marked(self, prev=False):
        return int(self.L.lib.orc_tiles_marked(self.ctx, 1 if prev else 0))


# ---------------------------------------------------------------------------------------------------------------
# Banded oracle: a frame too large for one core / one address space (3840x2160 with 1.6e9 bbox pixels, 8192x8192 with
# 3.2 GB of sample state) is rendered as horizontal bands by a pool of processes.  Pixels are independent given the
# ordered triangle stream (rasterizer/mod.rs:443-473), so the bands together are exactly the frame the reference
# would produce; per-pixel counters add up, per-triangle counters are the same in every band.
PER_PIXEL_COUNTERS = ("n_bbox_px", "n_covered_px", "n_shaded_px", "n_samples_written", "n_tex_oob")


def _band_worker(job):
    (make_scene, make_args, r0, r1, fast, gpu_paths) = job
    scene = make_scene(*make_args)
    r = OracleRenderer(scene.width, scene.height, get_lib(fast=fast), rows=(r0, r1))
    if scene.texture is not None:
        r.bind_texture(0, scene.texture.texels)
        for k, t in enumerate(getattr(scene, "extra_textures", [])):
            r.bind_texture(k + 1, t.texels)
    r.write_block(view=scene.view, projection=scene.projection)
    if getattr(scene, "scissor", None):
        r.set_scissor(*scene.scissor)
    if getattr(scene, "msaa", 4) != 4:
        r.set_msaa(scene.msaa)
    if getattr(scene, "guard_band", 1.0) != 1.0:
        r.set_guard_band(scene.guard_band)
    for d in scene.draws:
        r.write_block(world=d.world)
        r.render(d.mesh.vertices, d.mesh.attributes, d.mesh.indices, 0, d.fs)
    out = {"rows": (r0, r1), "counters": r.counters(), "mismatch": {}}
    if gpu_paths:  # compare this band with the GPU's arrays (memory-mapped .npy files) right here: only counts travel back
        H, W = scene.height, scene.width
        for name, arr in (("owner", r.owner_samples()), ("depth", r.depth_samples().view(np.uint32)), ("color", r.color_samples())):
            if name in gpu_paths:
                g = np.load(gpu_paths[name], mmap_mode="r")[r0:r1]
                g = g.view(np.uint32) if g.dtype != np.uint32 else g
                bad = arr != g
                n = int(bad.sum())
                out["mismatch"][name] = (n, (np.argwhere(bad)[:3] + [r0, 0, 0]).tolist() if n else [])
    fb = r.framebuffer()
    if gpu_paths and "fb" in gpu_paths:
        g = np.load(gpu_paths["fb"], mmap_mode="r")[r0:r1]
        bad = fb != g
        n = int(bad.sum())
        out["mismatch"]["fb"] = (n, (np.argwhere(bad)[:3] + [r0, 0]).tolist() if n else [])
    else:
        out["fb"] = fb
    r.close()
    return out


def banded_render(make_scene, make_args=(), height=None, n_bands=None, workers=None, fast=True, gpu_paths=None):
    """Render make_scene(*make_args) with the oracle as `n_bands` row bands in a process pool.  make_scene must be a
    picklable module-level callable (every worker rebuilds the scene: the builders are deterministic).  Returns
    dict(fb=[H][W] or None, counters=..., mismatch={array: (count, first positions)}); with gpu_paths = {"fb" | "owner" |
    "depth" | "color": path of a .npy file holding the GPU's array} every band is compared inside its worker."""
    import multiprocessing as mp

    workers = workers or max(1, (os.cpu_count() or 2))
    n_bands = n_bands or workers * 4  # more bands than workers: the bands of a scene are not equally expensive
    H = height
    edges = [H * k // n_bands for k in range(n_bands + 1)]
    jobs = [(make_scene, tuple(make_args), edges[k], edges[k + 1], fast, gpu_paths) for k in range(n_bands) if edges[k + 1] > edges[k]]
    get_lib(fast=fast)  # compile once, before the pool forks
    with mp.get_context("fork").Pool(workers) as pool:
        parts = pool.map(_band_worker, jobs, chunksize=1)
    counters = dict(parts[0]["counters"])
    for k in PER_PIXEL_COUNTERS:
        counters[k] = sum(p["counters"][k] for p in parts)
    mismatch = {}
    for p in parts:
        for name, (n, first) in p["mismatch"].items():
            tot, firsts = mismatch.get(name, (0, []))
            mismatch[name] = (tot + n, (firsts + first)[:5])
    fb = None
    if parts and "fb" in parts[0]:
        fb = np.concatenate([p["fb"] for p in sorted(parts, key=lambda p: p["rows"][0])], axis=0)
    return {"fb": fb, "counters": counters, "mismatch": mismatch}
