"""GPU parity tests proper (run on the B200 with `-m gpu`): the CUDA path, called through the
C ABI, against the CPU oracle on the same seeded inputs.

Bar (BASELINE.json north_star): coverage masks and depth-test outcomes bit-exact -> we compare the
per-sample owner key, the per-sample depth BITS and the per-sample packed colour; resolved 8-bit
colour within +-1 LSB -> we assert exact equality of the u32 image (tolerance 0, stricter than
required).  Work counters must match the oracle's too.
"""
import numpy as np
import pytest

from helpers import compare, gpu_render, oracle_render
from rusterizer_b200 import mathx, scenes
from rusterizer_b200.mesh import Mesh

pytestmark = pytest.mark.gpu


def check(scene, **kw):
    o = oracle_render(scene)
    g = gpu_render(scene, debug=True, **kw)
    msgs = compare(o, g)
    assert not msgs, f"{scene.name}: " + "; ".join(msgs)
    assert o["counters"]["n_tex_oob"] == 0 and o["counters"]["n_clip_overflow"] == 0
    return o, g


@pytest.mark.parametrize("elapsed", [0.0, 1.0, 2.5])
@pytest.mark.parametrize("fs", [0, 1, 2])
def test_default_scene(elapsed, fs):
    """C1: the crate's default scene (cube + sphere, 2 draws) at fixed `elapsed`, all three FS."""
    check(scenes.default_scene(elapsed, fs=fs))


@pytest.mark.parametrize("elapsed", [0.0, 0.4, 3.0, 3.5, 5.9])
def test_clip_test_scene(elapsed):
    """C1 --clip-test: one triangle orbiting the window border (main.rs:106-125)."""
    check(scenes.clip_test_scene(elapsed))


def test_sphere_small_triangles():
    """C2 scaled down: 80K ~pixel-sized triangles."""
    check(scenes.sphere_scene(201, 201, width=640, height=360))


def test_sphere_device_resident_mesh():
    check(scenes.sphere_scene(101, 51, width=480, height=270), device_resident=True)


def test_sphere_odd_resolution():
    """Width not a multiple of 4 or of the tile size: scalar write-back path, ragged tiles."""
    check(scenes.sphere_scene(101, 51, width=333, height=187))


@pytest.mark.parametrize("b2f", [True, False])
def test_overdraw_layers(b2f):
    """4 stacked jittered grids, back-to-front / front-to-back: the in-order depth test and the
    post-depth shading-position rule (rasterizer/mod.rs:70-83) under heavy overdraw."""
    check(scenes.overdraw_scene(60, 30, width=480, height=270, back_to_front=b2f))


def test_near_clip_field():
    """C3 scaled down: every triangle straddles the near plane (Sutherland-Hodgman + fan)."""
    o, _ = check(scenes.near_clip_scene(80, 40, width=640, height=360))
    assert o["counters"]["n_clipped_in"] > 0.9 * o["counters"]["n_tris_in"]


def test_fullscreen_quad_clipped_by_four_planes():
    """C4(ii) scaled down: 2 huge triangles -> large-triangle binning + pixel-parallel tile path."""
    check(scenes.fullscreen_quad_scene(512, 512))


def test_mixed_large_and_small_in_one_frame():
    """Large and small triangles interleaved in submission order inside the same tiles."""
    s = scenes.sphere_scene(61, 31, width=400, height=240)
    quad = scenes.fullscreen_quad_scene(400, 240).draws[0]
    s.draws = [s.draws[0], quad, scenes.Draw(s.draws[0].mesh, mathx.rotate(1.0, 0.2, 0.0), 1), quad]
    check(s)


def test_coincident_triangles_keep_first():
    """Equal depths: strict `<` keeps the EARLIER triangle (rasterizer/mod.rs:374)."""
    s = scenes.sphere_scene(41, 21, width=320, height=200)
    d = s.draws[0]
    s.draws = [d, scenes.Draw(d.mesh, d.world, 1), scenes.Draw(d.mesh, d.world, 2)]
    o, g = check(s)
    nt = d.mesh.n_triangles
    owned = o["owner"][o["owner"] != 0xFFFFFFFF] // 8
    assert owned.max() < nt  # nothing from the 2nd / 3rd draw ever wins


def test_random_soup():
    """Seeded random triangle soup (arbitrary sizes, both windings, slivers, off-screen parts)."""
    rng = np.random.RandomState(7)
    nt = 3000
    ctr = rng.uniform(-3, 3, (nt, 1, 3)).astype(np.float32)
    ctr[..., 2] = rng.uniform(-3.5, 6, (nt, 1)).astype(np.float32)
    size = (10 ** rng.uniform(-2.5, 0.3, (nt, 1, 1))).astype(np.float32)
    verts = (ctr + rng.uniform(-1, 1, (nt, 3, 3)).astype(np.float32) * size).reshape(-1, 3)
    attrs = rng.uniform(0, 1, (nt * 3, 6)).astype(np.float32)
    mesh = Mesh(verts, np.arange(nt * 3, dtype=np.uint32), attrs)
    for fs in (0, 1):
        s = scenes.sphere_scene(width=512, height=288, mesh=mesh, fs=fs)
        check(s)


def test_near_degenerate_fuzz_backface_margin():
    """near-collinear triangles whose computed area is 0 or slightly NEGATIVE can
    still cover a sample through the tie-break.  The geometry stage culls back-facing triangles
    only beyond a rigorous rounding margin; this fuzz (slivers through sample points, both
    windings, at pixel scale) must stay bit-exact, including the survey's own counterexample."""
    rng = np.random.RandomState(11)
    W, H = 1536, 1152
    nt = 60000
    # slivers: two far-apart points and a third one very close to the segment between them,
    # nudged to pass (almost) exactly through an RGSS sample position
    a = rng.uniform(8, 1100, (nt, 2))
    d = rng.uniform(-60, 60, (nt, 2))
    b = a + d
    t = rng.uniform(0.2, 0.8, (nt, 1))
    samp = np.array([[0.625, 0.125], [0.875, 0.625], [0.375, 0.875], [0.125, 0.375]])
    c = np.floor(a + t * d) + samp[rng.randint(0, 4, nt)]
    a = c - t * d
    b = c + (1 - t) * d
    c = c + rng.normal(0, 1, (nt, 2)) * (10.0 ** rng.uniform(-7, -2, (nt, 1)))
    pts = np.stack([a, b, c], 1).astype(np.float32)  # screen-space xy
    # the survey's counterexample (area2 = -3.05e-05, yet pixel (1059,1009) sample 1 is covered)
    pts[0] = [[1045.8563232421875, 1019.8419799804688], [1098.50146484375, 981.4735717773438],
              [1058.9920654296875, 1010.2684936523438]]
    flip = rng.rand(nt) < 0.5
    pts[flip] = pts[flip][:, ::-1]
    # screen -> clip with w = 1 under identity matrices: x_ndc = 2x/W - 1, y_ndc = 1 - 2y/H
    verts = np.zeros((nt, 3, 3), np.float32)
    verts[..., 0] = pts[..., 0] * np.float32(2.0 / W) - np.float32(1.0)
    verts[..., 1] = np.float32(1.0) - pts[..., 1] * np.float32(2.0 / H)
    verts[..., 2] = rng.uniform(-0.5, 0.5, (nt, 1)).astype(np.float32)
    attrs = rng.uniform(0, 1, (nt * 3, 6)).astype(np.float32)
    mesh = Mesh(verts.reshape(-1, 3), np.arange(nt * 3, dtype=np.uint32), attrs)
    s = scenes.sphere_scene(width=W, height=H, mesh=mesh, fs=1)
    s.view, s.projection = mathx.identity(), mathx.identity()
    s.draws[0].world = mathx.identity()
    o, g = check(s)
    assert o["counters"]["n_covered_px"] > 100


def test_exotic_float_inputs():
    """inf / NaN / 1e30 vertex coordinates and attributes: the reference's arithmetic is mirrored
    literally (NaN-propagating clamp, NaN-ignoring bbox, tie-break on NaN edge values), so even the
    garbage must agree.  Triangles whose screen coordinates are not finite take the literal
    per-pixel path of the tile kernel."""
    rng = np.random.RandomState(3)
    nt = 400
    verts = rng.uniform(-2, 2, (nt, 3, 3)).astype(np.float32)
    verts[..., 2] = rng.uniform(-2, 4, (nt, 3)).astype(np.float32)
    special = np.array([np.inf, -np.inf, np.nan, 1e30, -1e30, 1e19, 3e38], np.float32)
    for t in range(0, nt, 3):  # every third triangle gets one or two exotic coordinates
        for _ in range(rng.randint(1, 3)):
            verts[t, rng.randint(3), rng.randint(3)] = special[rng.randint(len(special))]
    attrs = rng.uniform(0, 1, (nt * 3, 6)).astype(np.float32)
    attrs[rng.randint(0, nt * 3, 40), rng.randint(0, 6, 40)] = np.nan
    mesh = Mesh(verts.reshape(-1, 3), np.arange(nt * 3, dtype=np.uint32), attrs)
    for fs in (1, 0):
        s = scenes.sphere_scene(width=320, height=200, mesh=mesh, fs=fs)
        o = oracle_render(s)
        g = gpu_render(s, debug=True)
        o["counters"].pop("n_tex_oob"), g["counters"].pop("n_tex_oob")  # NaN uv -> index 0, but count may differ
        msgs = compare(o, g)
        assert not msgs, "; ".join(msgs)


def test_wild_flag_is_per_frame():
    """One context, alternating frames with and without non-finite triangles.  The geometry stage raises
    FrameState::has_wild when it emits a triangle for the literal per-pixel walk and the tile stage only looks for
    such items when the flag is set: it must neither leak into the next (tame) frame nor be missed in a frame that
    follows a tame one.  Also mixes list lengths: the tame frame has pixel lists for the register replay of phase B."""
    from rusterizer_b200.render import Renderer

    rng = np.random.RandomState(11)
    nt = 300
    verts = rng.uniform(-2, 2, (nt, 3, 3)).astype(np.float32)
    verts[..., 2] = rng.uniform(-2, 4, (nt, 3)).astype(np.float32)
    special = np.array([np.inf, -np.inf, np.nan, 1e30, -1e30, 3e38], np.float32)
    for t in range(0, nt, 4):
        verts[t, rng.randint(3), rng.randint(3)] = special[rng.randint(len(special))]
    wild_mesh = Mesh(verts.reshape(-1, 3), np.arange(nt * 3, dtype=np.uint32), rng.uniform(0, 1, (nt * 3, 6)).astype(np.float32))
    wild = scenes.sphere_scene(width=320, height=200, mesh=wild_mesh, fs=1)
    tame = scenes.sphere_scene(65, 33, width=320, height=200, fs=0)
    r = Renderer(320, 200)
    r.uniforms().bind_texture(0, tame.texture)
    ow, ot = oracle_render(wild), oracle_render(tame)
    for sc, o in ((tame, ot), (wild, ow), (tame, ot), (wild, ow), (wild, ow), (tame, ot)):
        g = gpu_render(sc, debug=True, renderer=r)
        msgs = compare(o, g)
        assert not msgs, sc.name + ": " + "; ".join(msgs)
    r.close()


def test_empty_and_degenerate_inputs():
    """Empty mesh, zero-area triangles, a frame with no draws."""
    from rusterizer_b200.render import Renderer

    r = Renderer(200, 120)
    fb = r.framebuffer()
    assert (fb == 0xFF191919).all()
    empty = Mesh(np.zeros((0, 3)), np.zeros(0, np.uint32), np.zeros((0, 6)))
    r.render(empty, 0, 1)
    degenerate = Mesh([[0, 0, 0], [0, 1, 0], [0, 0, 0]], [0, 1, 2], np.zeros((3, 6)))
    r.render(degenerate, 0, 1)
    fb = r.framebuffer()
    assert (fb == 0xFF191919).all()
    assert r.counters()["n_degenerate"] == 1
    r.close()


def test_frames_are_independent_and_repeatable():
    """framebuffer() clears: rendering the same scene twice gives the same image; an empty frame
    after it gives the clear colour (resolve_and_clear, rasterizer/mod.rs:478-518)."""
    from rusterizer_b200.render import Renderer

    s = scenes.default_scene(1.0, width=320, height=180)
    r = Renderer(s.width, s.height)
    r.uniforms().bind_texture(0, s.texture)
    scenes.render_scene(r, s)
    a = r.framebuffer()
    scenes.render_scene(r, s)
    b = r.framebuffer()
    c = r.framebuffer()
    assert np.array_equal(a, b) and (c == 0xFF191919).all()
    o = oracle_render(s)
    assert np.array_equal(o["fb"], a)
    r.close()


def test_streaming_host_frames_match_synchronous_frames():
    """rz_render_host + rz_framebuffer_host_async (upload / kernels / download of neighbouring frames
    overlap on three streams, two staging sets, two output buffers): every streamed frame of a camera
    sweep is bit-identical to the oracle's frame, whatever is in flight around it."""
    import torch

    from rusterizer_b200.render import Renderer

    cams = scenes.orbit_cameras(16)
    base = scenes.sphere_scene(129, 65, width=320, height=192)
    mesh = base.draws[0].mesh
    want = [oracle_render(scenes.sphere_scene(129, 65, width=320, height=192, camera=c))["fb"] for c in cams[:6]]
    r = Renderer(base.width, base.height)
    r.uniforms().bind_texture(0, base.texture)
    blk = r.uniforms().write_block()
    blk.projection = base.projection
    blk.world = base.draws[0].world
    pos = torch.from_numpy(mesh.vertices).pin_memory()
    att = torch.from_numpy(mesh.attributes).pin_memory()
    idx = torch.from_numpy(mesh.indices.view(np.int32)).pin_memory()
    outs = [torch.zeros((base.height, base.width), dtype=torch.int32).pin_memory() for _ in range(6)]
    for rep in range(2):  # the second pass runs with warm (already sized) buffers and both parities used
        for k in range(6):
            blk.view = cams[k].get_view_matrix()
            r.render_arrays(pos.data_ptr(), att.data_ptr(), mesh.n_vertices, idx.data_ptr(), mesh.indices.size, 0, 0)
            r.framebuffer_host_async(outs[k].data_ptr())
            if k % 2 == 1:
                r.sync()  # at most two frames in flight per the API contract
        r.sync()
        for k in range(6):
            assert np.array_equal(outs[k].numpy().view(np.uint32), want[k]), f"pass {rep} frame {k}"
            outs[k].zero_()
    # the synchronous call still works after streaming, on the same ctx
    blk.view = cams[2].get_view_matrix()
    r.render(mesh, 0, 0)
    assert np.array_equal(r.framebuffer(), want[2])
    r.close()


def test_vertex_stage_bitwise():
    """Stage 1 alone: clip-space positions bitwise equal to the oracle's vertex stage."""
    from oracle.oracle import OracleRenderer
    from rusterizer_b200.render import Renderer

    s = scenes.sphere_scene(101, 51, width=320, height=180)
    d = s.draws[0]
    r = Renderer(s.width, s.height)
    blk = r.uniforms().write_block()
    blk.view, blk.projection, blk.world = s.view, s.projection, d.world
    got = r.vertex_stage(r.upload(d.mesh))
    o = OracleRenderer(s.width, s.height)
    o.write_block(world=d.world, view=s.view, projection=s.projection)
    want = o.vertex_stage(d.mesh.vertices)
    assert np.array_equal(got.view(np.uint32), want.view(np.uint32))
    r.close()


def test_error_behaviour():
    """Same failure points as the reference's panics, as error codes."""
    from rusterizer_b200.render import Renderer, RzError

    r = Renderer(64, 64)
    tex = scenes.Texture.checkerboard()
    with pytest.raises(RzError) as e:  # uniform.rs:31 assert!(len == index)
        r.uniforms().bind_texture(1, tex)
    assert e.value.code == -4
    tri = scenes.triangle()
    with pytest.raises(RzError) as e:  # uniform.rs:36 get_texture(0) on an empty Vec
        r.render(tri, 0, 0)
    assert e.value.code == -4
    with pytest.raises(RzError):
        r.render(tri, 0, 7)
    bad = Mesh(tri.vertices, [0, 1, 5], tri.attributes)  # render.rs:83-87 slice index panic
    r.render(bad, 0, 1)
    with pytest.raises(RzError) as e:
        r.framebuffer()
    assert e.value.code == -5
    r.close()


def test_capacity_growth_dense_tile():
    """A whole 20K-triangle sphere inside a handful of tiles: tile bins overflow their initial
    capacity; rz_framebuffer grows them and replays the frame; lists longer than the in-smem sort
    capacity take the global-memory sort path."""
    s = scenes.sphere_scene(101, 101, radius=0.15, width=256, height=256)
    check(s)
    s = scenes.overdraw_scene(8, 8, width=64, height=64)
    layers = [scenes.Draw(d.mesh, d.world, d.fs) for d in s.draws] * 200  # 800 draws into 16 tiles
    s.draws = layers
    check(s)
    # ~5 px cells, 40 layers: chunks of 255 small items whose fragments overflow the shared-memory
    # fragment pool -> the chunk is halved until it fits
    s = scenes.overdraw_scene(12, 12, width=64, height=64)
    s.draws = [scenes.Draw(d.mesh, d.world, d.fs) for d in s.draws] * 10
    check(s)


def test_async_overflow_is_replayed_by_sync():
    """An async frame issued FIRST on a fresh context (device buffers still at their initial size) overflows its tile
    bins; it cannot know.  rz_sync grows the buffers from the device's own counts and replays the frame into the same
    destination, so the image is complete -- here equal to the oracle's.  With several unsynchronised async frames
    only the last one can be re-created: RZ_E_CAPACITY says so, and the next frame just works."""
    import torch

    from rusterizer_b200.render import Renderer, RzError

    s = scenes.overdraw_scene(8, 8, width=64, height=64)
    s.draws = [scenes.Draw(d.mesh, d.world, d.fs) for d in s.draws] * 100  # 400 draws, 51 200 triangles into 16 tiles
    o = oracle_render(s)
    r = Renderer(s.width, s.height)
    r.uniforms().bind_texture(0, s.texture)
    scenes.render_scene(r, s)
    ptr = r.framebuffer_async()
    r.sync()  # grows + replays, returns RZ_OK (one async frame pending)

    class _Raw:
        __cuda_array_interface__ = {"shape": (s.height, s.width), "typestr": "<i4", "data": (ptr, False), "version": 3}

    got = torch.as_tensor(_Raw(), device="cuda").cpu().numpy().view(np.uint32)
    assert np.array_equal(got, o["fb"])
    r.close()
    # host-streaming form, two frames in flight on a fresh context: the last image is complete, the error names the rest
    r = Renderer(s.width, s.height)
    r.uniforms().bind_texture(0, s.texture)
    outs = [torch.empty((s.height, s.width), dtype=torch.int32).pin_memory() for _ in range(2)]
    for k in range(2):
        scenes.render_scene(r, s)
        r.framebuffer_host_async(outs[k].data_ptr())
    with pytest.raises(RzError) as e:
        r.sync()
    assert e.value.code == -6
    assert np.array_equal(outs[1].numpy().view(np.uint32), o["fb"])
    scenes.render_scene(r, s)  # buffers are large enough now
    r.framebuffer_host_async(outs[0].data_ptr())
    r.sync()
    assert np.array_equal(outs[0].numpy().view(np.uint32), o["fb"])
    # a mesh destroyed between the async frame and rz_sync: the frame cannot be replayed, rz_sync says so (no crash)
    r2 = Renderer(s.width, s.height)
    r2.uniforms().bind_texture(0, s.texture)
    dms = [r2.upload(d.mesh) for d in s.draws]
    scenes.render_scene(r2, s, dms)
    r2.framebuffer_async()
    for m in dms:
        m.close()
    with pytest.raises(RzError) as e:
        r2.sync()
    assert e.value.code == -6
    r2.close()
    # discard: recorded draws are dropped, the next frame is just the clear colour
    scenes.render_scene(r, s)
    r.discard_frame()
    assert (r.framebuffer() == 0xFF191919).all()
    r.close()


def test_cpp_host_demo_runs(tmp_path):
    """The C++ mirror of the crate API renders the Mode::Demo frame on the GPU (main.rs:93-105): texture loaded
    from a PNG file (Texture::from_png_file, texture.rs:26-45), image written as PNG instead of a window."""
    import re
    import subprocess
    from pathlib import Path

    from rusterizer_b200 import image
    from rusterizer_b200.texture import Texture

    demo = Path(__file__).resolve().parent.parent / "rusterizer_b200" / "host" / "rz_demo"
    tex_png, out_png = tmp_path / "checkerboard.png", tmp_path / "frame.png"
    tex_png.write_bytes(image.encode_png(Texture.checkerboard().texels))
    p = subprocess.run([str(demo), "1.0", str(out_png), str(tex_png)], capture_output=True, text=True)
    assert p.returncode == 0, p.stderr
    frame = image.read_png(out_png)
    m = re.search(r"touched_px=(\d+)", p.stdout)
    assert frame.shape == (720, 1280, 3) and int((frame != 0x19).any(axis=2).sum()) == int(m.group(1))
    m = re.search(r"tris_in=(\d+) samples_written=(\d+) touched_px=(\d+)", p.stdout)
    assert m and int(m.group(1)) == 268
    o = oracle_render(scenes.default_scene(1.0))
    # libm vs Python trig may differ in the last bit of a matrix entry, so compare loosely
    assert abs(int(m.group(2)) - o["counters"]["n_samples_written"]) < 0.01 * o["counters"]["n_samples_written"]


def test_peer_store_primitives_two_contexts_one_gpu():
    """The screen-space sharding protocol of sharding.PeerFrame on ONE GPU: two contexts (two 'ranks', two
    streams) store their tile rows into one shared image; completion and back-pressure travel through the
    flag kernels (rz_signal / rz_wait_flags).  The assembled frames equal the oracle's frames."""
    import torch

    from rusterizer_b200.render import Renderer

    cams = scenes.orbit_cameras(8)
    s = scenes.sphere_scene(97, 49, width=320, height=192)
    mesh = s.draws[0].mesh
    want = [oracle_render(scenes.sphere_scene(97, 49, width=320, height=192, camera=c))["fb"] for c in cams[:4]]
    ctx = []
    for q in range(2):
        r = Renderer(s.width, s.height)
        r.uniforms().bind_texture(0, s.texture)
        b = r.uniforms().write_block()
        b.projection, b.world = s.projection, s.draws[0].world
        r.set_row_range(*( (0, 96) if q == 0 else (96, 192) ))
        ctx.append((r, b))
    root = ctx[0][0]
    img, handle = root.shared_alloc(s.width * s.height * 4)
    assert len(handle) == 64
    flags, _ = root.shared_alloc(256)
    ack, _ = ctx[1][0].shared_alloc(128)

    class _Raw:
        __cuda_array_interface__ = {"shape": (s.height, s.width), "typestr": "<i4", "data": (img, False), "version": 3}

    view = torch.as_tensor(_Raw(), device="cuda")
    outs = []
    for f in range(4):
        seq = f + 1
        for q, (r, b) in enumerate(ctx):
            b.view = cams[f].get_view_matrix()
            r.render(mesh, 0, 0)
        # "rank 1": wait until the presenter has consumed the previous frame, store rows, raise its flag
        r1 = ctx[1][0]
        if f >= 1:
            r1.wait_flags(ack, 1, 128, f)
        r1.framebuffer_async(img + 96 * s.width * 4)
        r1.signal(flags + 128, seq)
        # presenter: own rows, wait for rank 1, consume (copy out on its stream), acknowledge
        root.framebuffer_async(img)
        root.wait_flags(flags + 128, 1, 128, seq)
        root.sync()
        outs.append(view.cpu().numpy().view(np.uint32).copy())
        root.signal([ack], seq)
    for r, _ in ctx:
        r.sync()
    for f in range(4):
        assert np.array_equal(outs[f], want[f]), f"frame {f}"
    # a flag that never arrives: the wait kernel gives up and rz_sync reports RZ_E_PEER
    from rusterizer_b200.render import RzError

    root.wait_flags(flags + 128, 1, 128, 1000, timeout_ms=50)
    with pytest.raises(RzError) as ei:
        root.sync()
    assert ei.value.code == -8
    root.sync()  # the condition is cleared once reported
    root.shared_free(img); root.shared_free(flags); ctx[1][0].shared_free(ack)
    for r, _ in ctx:
        r.close()


def test_rgb_texture_non_square_and_out_of_range_uvs():
    """Texture::read_texel with texel_width 3 (texture.rs:51-61), a non-square non-power-of-two texture, and UVs
    outside [0,1]: the reference neither wraps nor clamps (texture.rs:65-83) -- coordinates past the buffer are a
    panic there, a clamped read counted in n_tex_oob here and in the oracle; everything else must match bit for bit."""
    rng = np.random.default_rng(11)
    tex = scenes.Texture(rng.integers(0, 256, (37, 61, 3), dtype=np.uint8))
    base = scenes.default_scene(1.0, width=320, height=180)
    s = scenes.Scene("rgb_tex", base.width, base.height, base.view, base.projection, base.draws, tex)
    check(s)
    # UVs scaled to [-0.25, 1.5]: part of the samples fall outside the texture
    draws = []
    for d in base.draws:
        a = d.mesh.attributes.copy()
        a[:, 4:6] = a[:, 4:6] * 1.75 - 0.25
        draws.append(scenes.Draw(Mesh(d.mesh.vertices, d.mesh.indices, a), d.world, d.fs))
    s2 = scenes.Scene("rgb_tex_oob", base.width, base.height, base.view, base.projection, draws, tex)
    o = oracle_render(s2)
    g = gpu_render(s2, debug=True)
    assert o["counters"]["n_tex_oob"] > 0
    msgs = compare(o, g)
    assert not msgs, "; ".join(msgs)


def test_many_draws_in_one_frame():
    """150 draws accumulate into one frame in submission order (main.rs:170-173): more draws than the per-draw
    table's first allocation, alternating fragment shaders, overlapping in depth."""
    rng = np.random.default_rng(5)
    base = scenes.default_scene(0.5, width=256, height=144)
    quad, tri = scenes.centered_quad(1.0), scenes.triangle()
    draws = []
    for k in range(150):
        world = mathx.matmul(mathx.translate(float(rng.uniform(-3, 3)), float(rng.uniform(-2, 2)), float(rng.uniform(-1, 6))),
                             mathx.rotate(float(rng.uniform(0, 6)), float(rng.uniform(0, 6)), float(rng.uniform(0, 6))))
        draws.append(scenes.Draw(quad if k % 2 else tri, world, k % 3))
    s = scenes.Scene("many_draws", base.width, base.height, base.view, base.projection, draws, base.texture)
    check(s)
    check(s, device_resident=True)


def test_shader_registry_texture_index_and_blend():
    """Registry extension (include/rz.h): sampling shaders take a texture index
    (Uniforms::get_texture(index), uniform.rs:35-37) and FS TextureBlend = (sample + attr.color) / 2.0 with the
    Color operators of color.rs:88-111 -- three textures (RGBA and RGB), five draws, against the oracle."""
    from rusterizer_b200.render import Renderer, RzError

    rng = np.random.default_rng(23)
    base = scenes.default_scene(1.0, width=320, height=180)
    cube, sph = base.draws[0], base.draws[1]
    t1 = scenes.Texture(rng.integers(0, 256, (64, 48, 4), dtype=np.uint8))
    t2 = scenes.Texture(rng.integers(0, 256, (33, 77, 3), dtype=np.uint8))
    W = scenes.fs_with_texture
    draws = [scenes.Draw(cube.mesh, cube.world, W(scenes.FS_TEXTURE, 1)),
             scenes.Draw(sph.mesh, sph.world, W(scenes.FS_TEXTURE_BLEND, 2)),
             scenes.Draw(scenes.centered_quad(3.0), mathx.translate(-2.0, 0.5, 1.0), W(scenes.FS_TEXTURE_BLEND, 0)),
             scenes.Draw(scenes.triangle(), mathx.translate(2.0, -1.0, -1.0), scenes.FS_COLOR),
             scenes.Draw(scenes.centered_quad(1.5), mathx.matmul(mathx.translate(2.5, 1.5, 0.0), mathx.rotate(0.4, 0.2, 0.1)),
                         scenes.FS_TEXTURE)]
    s = scenes.Scene("registry", base.width, base.height, base.view, base.projection, draws, base.texture, [t1, t2])
    check(s)
    check(s, device_resident=True)
    # a near-clipped triangle through the blend shader (interpolated attributes live in an AttrRec)
    c = scenes.clip_test_scene(0.4, width=320, height=180)
    c.draws[0].fs = W(scenes.FS_TEXTURE_BLEND, 1)
    c.extra_textures = [t1]
    check(c)
    # failure points: unbound texture index (uniform.rs:36 index panic), index on a non-sampling shader
    r = Renderer(64, 64)
    r.uniforms().bind_texture(0, base.texture)
    with pytest.raises(RzError) as e:
        r.render(scenes.triangle(), 0, W(scenes.FS_TEXTURE, 1))
    assert e.value.code == -4
    with pytest.raises(RzError) as e:
        r.render(scenes.triangle(), 0, W(scenes.FS_COLOR, 1))
    assert e.value.code == -1
    r.close()


@pytest.mark.parametrize("rect", [(37, 21, 251, 163), (0, 0, 320, 180), (100, 50, 101, 51), (64, 48, 64, 120), (300, 170, 9999, 9999)])
def test_scissor_rect(rect):
    """rz_set_scissor (the extension the reference sketches at rasterizer/mod.rs:349-350): the triangles' pixel bounding
    boxes are intersected with the rect instead of the viewport -- small and large triangles, a near-clipped one,
    rects that cut through tiles, a one-pixel rect, an empty rect and one that is clamped to the viewport."""
    base = scenes.default_scene(1.0, width=320, height=180)
    clip = scenes.clip_test_scene(0.4, width=320, height=180)
    sph = scenes.sphere_scene(65, 33, width=320, height=180)
    draws = base.draws + [scenes.Draw(clip.draws[0].mesh, clip.draws[0].world, scenes.FS_COLOR),
                          scenes.Draw(sph.draws[0].mesh, mathx.translate(1.5, -1.0, 2.0), scenes.FS_TEXTURE)]
    s = scenes.Scene("scissor", 320, 180, base.view, base.projection, draws, base.texture, scissor=rect)
    o, g = check(s)
    x0, y0, x1, y1 = (min(rect[0], 320), min(rect[1], 180), min(rect[2], 320), min(rect[3], 180))
    outside = np.ones((180, 320), bool)
    outside[y0:y1, x0:x1] = False
    assert (g["fb"][outside] == 0xFF191919).all()
    full = oracle_render(scenes.Scene("noscissor", 320, 180, base.view, base.projection, draws, base.texture))
    assert np.array_equal(g["fb"][~outside], full["fb"][~outside])  # inside the rect nothing changes


@pytest.mark.parametrize("msaa", [1, 2, 8])
def test_runtime_msaa_sample_counts(msaa):
    """1, 2 and 8 samples per pixel (the reference fixes 4, rasterizer/mod.rs:23).  Per-sample
    owner, depth bits and colour, the box-filtered image and all counters equal the oracle's with the same count."""
    for mk in (lambda: scenes.default_scene(1.0, width=320, height=180), lambda: scenes.clip_test_scene(0.7, width=320, height=180),
               lambda: scenes.sphere_scene(129, 65, width=320, height=200), lambda: scenes.overdraw_scene(40, 20, width=320, height=192),
               lambda: scenes.near_clip_scene(60, 30, 320, 180), lambda: scenes.default_scene(2.5, fs=2, width=200, height=120)):
        s = mk()
        s.msaa = msaa
        o = oracle_render(s)
        g = gpu_render(s, debug=True)
        assert g["depth"].shape[-1] == msaa
        msgs = compare(o, g)
        assert not msgs, f"{s.name} x{msaa}: " + "; ".join(msgs)


def test_msaa_count_switches_between_frames():
    """One context renders 4 -> 8 -> 1 -> 4 samples per pixel; every frame matches the oracle of its count, and changing
    the count while draws are recorded is refused."""
    from rusterizer_b200.render import Renderer, RzError

    s = scenes.default_scene(1.0, width=256, height=144)
    r = Renderer(s.width, s.height)
    r.uniforms().bind_texture(0, s.texture)
    for n in (4, 8, 1, 4):
        s.msaa = n
        r.set_msaa(n)
        scenes.render_scene(r, s)
        assert np.array_equal(r.framebuffer(), oracle_render(s)["fb"]), n
    scenes.render_scene(r, s)
    with pytest.raises(RzError):
        r.set_msaa(2)
    with pytest.raises(RzError):
        r.set_msaa(3)
    r.close()


@pytest.mark.parametrize("guard", [1.5, 8.0])
def test_guard_band_clipping(guard):
    """guard-band clipping (rasterizer/mod.rs:417-419), per sample against the oracle; fewer
    triangles go through Sutherland-Hodgman than with the reference's planes."""
    for mk in (lambda: scenes.clip_test_scene(0.7, width=320, height=180), lambda: scenes.near_clip_scene(60, 30, 320, 180),
               lambda: scenes.fullscreen_quad_scene(256, 256), lambda: scenes.default_scene(1.0, width=320, height=180)):
        s = mk()
        base = oracle_render(s)["counters"]["n_clipped_in"]
        s.guard_band = guard
        o = oracle_render(s)
        msgs = compare(o, gpu_render(s, debug=True))
        assert not msgs, f"{s.name} g={guard}: " + "; ".join(msgs)
        assert o["counters"]["n_clipped_in"] <= base


def test_fuzz_campaign_slice():
    """Sixty frames of the fuzz campaign (tests/fuzz_parity.py: random resolutions, soups, grids, huge triangles,
    shaders, sample counts, guard bands, scissor rects); profiles/r02_fuzz.txt records a 300 s run of the same."""
    import fuzz_parity

    for seed in range(9000, 9060):
        s = fuzz_parity.make_case(seed)
        msgs = compare(oracle_render(s), gpu_render(s, debug=True, device_resident=bool(seed & 1)))
        assert not msgs, f"seed {seed}: " + "; ".join(msgs)
