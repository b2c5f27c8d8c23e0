"""CPU-side tests (no GPU): host helpers, scene definitions, the C-ABI surface, multi-rank sharding."""
import ctypes
import os
import re
import subprocess
import sys
from pathlib import Path

import numpy as np
import pytest

ROOT = Path(__file__).resolve().parent.parent


def test_checkerboard_matches_reference_fixture_hash():
    """decoded images/checkerboard.png = 400x400 RGBA, 4x4 squares, top-left black."""
    from rusterizer_b200.texture import CHECKERBOARD_SHA256, Texture

    t = Texture.checkerboard()
    assert t.texels.shape == (400, 400, 4) and t.texel_width == 4
    assert t.sha256() == CHECKERBOARD_SHA256
    assert tuple(t.texels[0, 0]) == (0, 0, 0, 255) and tuple(t.texels[0, 100]) == (255, 255, 255, 255)


def test_default_camera_and_projection():
    """camera.rs:46-54 + main.rs:137-142 (values quoted in SURVEY.md section 8d)."""
    from rusterizer_b200 import mathx
    from rusterizer_b200.camera import Camera
    from rusterizer_b200.scenes import default_projection

    v = Camera().get_view_matrix()
    assert np.array_equal(np.abs(v), np.array([[1, 0, 0, 0], [0, 1, 0, 0], [0, 0, 1, 5], [0, 0, 0, 1]], np.float32))
    assert v[2, 2] == -1 and v[2, 3] == -5
    p = default_projection(1280, 720)
    assert p[0, 0] == np.float32(1.0) and abs(p[1, 1] - 1.7777778) < 1e-6
    assert abs(p[2, 2] + 1.0100503) < 1e-6 and abs(p[2, 3] + 2.0100503) < 1e-6 and p[3, 2] == -1
    # orbit camera at angle 0 is the default camera
    assert np.array_equal(Camera.orbit(0.0).get_view_matrix(), v)
    assert mathx.vlen(Camera.orbit(1.0).pos) == pytest.approx(5.0, abs=1e-5)


def test_mesh_generators():
    """mesh.rs:15-207"""
    from rusterizer_b200 import mesh

    c = mesh.cube(1.0)
    assert c.n_vertices == 24 and c.n_triangles == 12
    assert list(c.indices[:6]) == [0, 1, 2, 0, 2, 3]
    assert np.array_equal(c.attributes[:4, 4:], [[0, 0], [1, 0], [1, 1], [0, 1]])
    s = mesh.sphere(0.5)
    assert s.n_vertices == 17 * 9 and s.n_triangles == 256
    assert list(s.indices[:6]) == [0, 1, 18, 0, 18, 17]
    assert np.allclose(np.linalg.norm(s.vertices, axis=1), 0.5, atol=1e-6)
    assert np.array_equal(s.attributes[:, 0:3], np.abs(s.vertices))
    assert mesh.triangle().n_triangles == 1 and mesh.centered_quad(2.0).n_triangles == 2


def test_benchmark_scene_sizes():
    """BASELINE.md section 4 / SURVEY.md section 8d: C2 has exactly 1 000 000 triangles, 501 501 vertices and
    38 988 628 algorithmic bytes; C3 has 250 000 triangles."""
    from rusterizer_b200 import scenes

    c2 = scenes.sphere_scene()
    assert (c2.n_triangles, c2.n_vertices, c2.width, c2.height) == (1_000_000, 501_501, 1920, 1080)
    assert c2.algorithmic_bytes() == 38_988_628
    c3 = scenes.near_clip_scene()
    assert c3.n_triangles == 250_000 and (c3.width, c3.height) == (3840, 2160)
    assert scenes.overdraw_scene().n_triangles == 1_000_000
    assert len(scenes.orbit_cameras(1024)) == 1024


def _header_symbols():
    text = (ROOT / "include" / "rz.h").read_text()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(rz_[a-z_0-9]+)\s*\(", text)))


def test_c_abi_library_exports_every_declared_symbol():
    """The drop-in boundary: librz_b200.so must export exactly what include/rz.h declares."""
    import __graft_entry__ as g
    from rusterizer_b200 import render

    if not render.library_path().exists():
        g.build()
    lib = ctypes.CDLL(str(render.library_path()))
    declared = _header_symbols()
    assert declared == sorted(render.ABI_SYMBOLS)
    for name in declared:
        assert hasattr(lib, name), f"{name} declared in include/rz.h but not exported"
    lib.rz_version.restype = ctypes.c_char_p
    assert b"sm_100a" in lib.rz_version()
    lib.rz_tile_width.restype = ctypes.c_uint32
    assert lib.rz_tile_width() == 16 and lib.rz_tile_height() == 16


def test_no_cpu_fallback_without_gpu():
    """The product path fails loudly when there is no CUDA device (rz_create -> RZ_E_NO_DEVICE)."""
    import torch

    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    from rusterizer_b200.render import Renderer, RzError

    with pytest.raises(RzError) as e:
        Renderer(64, 64)
    assert e.value.code == -3 and "no CPU fallback" in str(e.value)


def test_cpp_host_mirror_builds_and_fails_loudly_without_gpu():
    """rusterizer.hpp + demo_main.cpp (the crate's Mode::Demo frame) compile and link against the C ABI;
    without a GPU the demo must stop at rz_create with RZ_E_NO_DEVICE."""
    import torch

    import __graft_entry__ as g

    g.build()
    demo = ROOT / "rusterizer_b200" / "host" / "rz_demo"
    assert demo.exists()
    if torch.cuda.is_available():
        pytest.skip("a GPU is present (the demo is run by the gpu tests)")
    p = subprocess.run([str(demo)], capture_output=True, text=True)
    assert p.returncode == 1 and "no CUDA device" in p.stderr


def test_cpp_and_python_host_mirrors_build_the_same_data(tmp_path):
    """Both host mirrors restate math/mod.rs:92-122, math/transform.rs, camera.rs:10-43 and mesh.rs:15-207; what they
    build crosses the C ABI as plain arrays, so it has to be the same data.  host/host_dump.cpp prints the C++ side.
    Bit-exact for everything: projection, rotations, the default and orbit cameras, quad, triangle, cube, the crate's
    17 x 9 sphere and other sphere sizes.  Both mirrors call libm's sinf / cosf / tanf like Rust's f32::sin / cos / tan do
    (glibc's sinf is not correctly rounded for every argument -- sinf(0.29452431), a theta and phi of the 65 x 33 sphere,
    is one ulp off -- so a "rounded double" sin would hand the raster path slightly different meshes than the Rust host)."""
    from rusterizer_b200 import mathx, mesh as M
    from rusterizer_b200.camera import Camera

    exe = tmp_path / "host_dump"
    host = ROOT / "rusterizer_b200" / "host"
    subprocess.run(["g++", "-std=c++17", "-O2", "-ffp-contract=off", "-Wall", "-I" + str(ROOT / "include"), str(host / "host_dump.cpp"),
                    "-o", str(exe), "-L" + str(ROOT / "rusterizer_b200"), "-lrz_b200", "-Wl,-rpath," + str(ROOT / "rusterizer_b200"),
                    "-Wl,-rpath,/usr/local/cuda/lib64"], check=True)
    rec = {}
    for line in subprocess.run([str(exe)], capture_output=True, text=True, check=True).stdout.strip().split("\n"):
        name, kind, *vals = line.split()
        rec[name] = (np.array([int(v, 16) for v in vals], np.uint32).view(np.float32) if kind == "f"
                     else np.array([int(v) for v in vals], np.uint32))
    F = np.float32
    pi = F(3.14159274101257324)
    py = {
        "project": mathx.project(1.0, 200.0, F(720.0) / F(1280.0), pi / F(2.0)),
        "project_1080": mathx.project(1.0, 200.0, F(1080.0) / F(1920.0), pi / F(2.0)),
        "rotate_110": mathx.rotate(1.0, 1.0, 0.0), "rotate_0303": mathx.rotate(0.3, 0.3, 0.0), "rotate_xyz": mathx.rotate(2.5, -0.7, 4.1),
        "demo_sphere_world": mathx.matmul(mathx.rotate(1.0, 0.0, pi / F(4.0)), mathx.translate(0.0, 3.0, 0.0)),
        "view_default": Camera().get_view_matrix(),
    }
    for k in (1, 77, 256, 511, 1000):
        py[f"view_orbit_{k}"] = Camera.orbit(F(2.0) * pi * F(k) / F(1024.0)).get_view_matrix()
    for name, m in (("centered_quad", M.centered_quad(9.0)), ("triangle", M.triangle()), ("cube", M.cube(1.0)),
                    ("sphere_default", M.sphere(0.5)), ("sphere_65_33", M.sphere(2.0, 65, 33))):
        py[name + ".vertices"], py[name + ".attributes"], py[name + ".indices"] = m.vertices, m.attributes, m.indices
    assert set(py) == set(rec)
    for k, v in py.items():
        a, b = np.asarray(v).reshape(-1), rec[k]
        assert a.shape == b.shape, k
        if a.dtype != np.float32:
            assert np.array_equal(a.astype(np.uint32), b), k
        else:
            assert np.array_equal(a.view(np.uint32), b.view(np.uint32)), k


def test_product_package_never_imports_the_oracle():
    """oracle/ is test infrastructure: nothing under rusterizer_b200/ may import, load or call it."""
    for path in (ROOT / "rusterizer_b200").rglob("*"):
        if path.suffix in (".py", ".cu", ".cuh", ".h", ".hpp") and path.is_file():
            text = path.read_text()
            assert "liboracle" not in text and "rz_oracle" not in text, path
            assert not re.search(r"^\s*(from|import)\s+oracle", text, flags=re.M), path


def test_row_ranges_partition_the_framebuffer():
    from rusterizer_b200 import sharding

    for H in (1080, 8192, 720, 17, 16):
        for world in (1, 2, 3, 4, 8):
            rows = [sharding.row_range(r, world, H, 16) for r in range(world)]
            assert rows[0][0] == 0 and rows[-1][1] == H
            for (a0, a1), (b0, b1) in zip(rows, rows[1:]):
                assert a1 == b0 and a0 <= a1
            for a0, a1 in rows:
                assert a0 == a1 or (a0 % 16 == 0 and (a1 % 16 == 0 or a1 == H))  # empty strips allowed at the tail
    assert sharding.frame_ids(1, 4, 10) == [1, 5, 9]
    assert sorted(sum((sharding.frame_ids(r, 8, 1024) for r in range(8)), [])) == list(range(1024))


_GLOO_WORKER = r"""
import os, sys
sys.path.insert(0, {root!r}); sys.path.insert(0, {root!r} + "/tests")
import numpy as np, torch, torch.distributed as dist
from rusterizer_b200 import scenes, sharding
from helpers import oracle_render
dist.init_process_group("gloo", rank=int(os.environ["RANK"]), world_size=int(os.environ["WORLD_SIZE"]))
rank, world = dist.get_rank(), dist.get_world_size()
# tile-row sharding: every rank holds its strip of the frame, the gather must reassemble the frame
sc = scenes.default_scene(1.0, width=200, height=120)
full = oracle_render(sc)["fb"]
r0, r1 = sharding.row_range(rank, world, sc.height, 16)
per = sharding.strip_rows(sc.height, world, 16)
strip = np.zeros((per, sc.width), np.uint32)
strip[: r1 - r0] = full[r0:r1]
out = sharding.gather_strips(torch.from_numpy(strip.view(np.int32)), sc.height).numpy().view(np.uint32)
assert out.shape == full.shape and np.array_equal(out, full), "gathered frame differs"
# frame sharding: the union of the ranks' frame ids covers the sweep exactly once
ids = sharding.frame_ids(rank, world, 10)
gathered = [None] * world
dist.all_gather_object(gathered, ids)
assert sorted(sum(gathered, [])) == list(range(10))
dist.barrier(); dist.destroy_process_group()
print("rank", rank, "ok")
"""


def test_sharding_world_size_2_gloo(tmp_path):
    """N>1 host logic on CPU: 2 ranks over gloo -- strip gather reassembles the oracle's frame."""
    script = tmp_path / "worker.py"
    script.write_text(_GLOO_WORKER.format(root=str(ROOT)))
    env = dict(os.environ, MASTER_ADDR="127.0.0.1", MASTER_PORT="29531", WORLD_SIZE="2")
    procs = [subprocess.Popen([sys.executable, str(script)], env=dict(env, RANK=str(r)), stdout=subprocess.PIPE,
                              stderr=subprocess.STDOUT, text=True) for r in range(2)]
    outs = [p.communicate(timeout=240)[0] for p in procs]
    for p, o in zip(procs, outs):
        assert p.returncode == 0, o


_PEER_WORKER = r"""
import os, sys, time
sys.path.insert(0, {root!r}); sys.path.insert(0, {root!r} + "/tests")
import numpy as np, torch.distributed as dist
from multiprocessing import shared_memory
from rusterizer_b200 import sharding
dist.init_process_group("gloo", rank=int(os.environ["RANK"]), world_size=int(os.environ["WORLD_SIZE"]))
rank, world = dist.get_rank(), dist.get_world_size()
W, H, TH = 64, 112, 16


class FakeRenderer:
    # The peer-memory surface of render.Renderer (include/rz.h) on host shared memory: 'device pointers' are
    # (block << 40 | offset), handles are shm names, frames are copied synchronously from `self.next_frame`.
    def __init__(self):
        self.width, self.height, self.blocks, self.rows, self.il = W, H, {{}}, (0, H), (0, 0, 1)
        self.next_frame = None
    def _view(self, ptr, n):
        return np.ndarray((n,), np.uint32, self.blocks[ptr >> 40].buf, offset=ptr & ((1 << 40) - 1))
    def shared_alloc(self, nbytes):
        shm = shared_memory.SharedMemory(create=True, size=nbytes)
        np.ndarray((nbytes // 4,), np.uint32, shm.buf)[:] = 0
        bid = len(self.blocks) + 1 + 100 * rank
        self.blocks[bid] = shm
        return bid << 40, shm.name.encode().ljust(64, b"\0")
    def shared_open(self, handle):
        shm = shared_memory.SharedMemory(name=handle.rstrip(b"\0").decode())
        bid = len(self.blocks) + 1 + 100 * rank
        self.blocks[bid] = shm
        return bid << 40
    def shared_close(self, ptr):
        self.blocks.pop(ptr >> 40).close()
    def shared_free(self, ptr):
        shm = self.blocks.pop(ptr >> 40); shm.close(); shm.unlink()
    def set_row_range(self, a, b): self.rows = (a, b)
    def set_row_interleave(self, band, r, w): self.il = (band, r, w)
    def owns(self, y):
        band, r, w = self.il
        return self.rows[0] <= y < self.rows[1] and (band == 0 or (y // TH // band) % w == r)
    def framebuffer_async(self, dst):
        img = self._view(dst - self.rows[0] * W * 4, W * H).reshape(H, W)
        for y in range(H):
            if self.owns(y): img[y] = self.next_frame[y]
    def signal(self, ptrs, value):
        for p in ([ptrs] if isinstance(ptrs, int) else ptrs): self._view(p, 1)[0] = value
    def wait_flags(self, ptr, n, stride, value, timeout_ms=0):
        t0 = time.time()
        for i in range(n):
            while int(self._view(ptr + i * stride, 1)[0]) < value:
                assert time.time() - t0 < 20, "peer flag never arrived"
                time.sleep(0.0005)
    def sync(self): pass


for band in (0, 1, 2):
    r = FakeRenderer()
    pf = sharding.PeerFrame(r, root=0, n_buffers=2, interleave_band=band)
    rng = np.random.default_rng(99)
    frames = [rng.integers(0, 2 ** 32, (H, W), dtype=np.uint64).astype(np.uint32) for _ in range(7)]  # same on every rank
    for f, frame in enumerate(frames):
        r.next_frame = frame
        img = pf.finish_frame()
        if rank == 0:
            got = r._view(img, W * H).reshape(H, W)
            assert np.array_equal(got, frame), (band, f)   # complete and not stale: every rank's rows have arrived
            pf.release()
        else:
            assert img is None
    pf.close()
dist.barrier(); dist.destroy_process_group()
print("rank", rank, "ok")
"""


def test_peer_frame_protocol_world_size_3_gloo(tmp_path):
    """The N>1 host logic of sharding.PeerFrame on CPU: 3 ranks over gloo with a fake renderer whose 'peer memory'
    is host shared memory -- handle exchange, contiguous and interleaved row ownership, completion flags, the two
    alternating images and the acknowledgement back-pressure deliver every frame complete to the presenter."""
    script = tmp_path / "peer_worker.py"
    script.write_text(_PEER_WORKER.format(root=str(ROOT)))
    env = dict(os.environ, MASTER_ADDR="127.0.0.1", MASTER_PORT="29533", WORLD_SIZE="3")
    procs = [subprocess.Popen([sys.executable, str(script)], env=dict(env, RANK=str(r)), stdout=subprocess.PIPE,
                              stderr=subprocess.STDOUT, text=True) for r in range(3)]
    outs = [p.communicate(timeout=240)[0] for p in procs]
    for p, o in zip(procs, outs):
        assert p.returncode == 0, o
