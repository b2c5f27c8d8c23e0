"""Cross-check of the C oracle (oracle/rz_oracle.c) against a second restatement of the reference written
independently as a numpy array program (oracle/py_restatement.py): every depth sample, every packed colour sample
and every resolved pixel of whole frames must agree bit for bit.  The reference (Rust) cannot be built in this
image, so this is what stands behind the oracle's whole-frame behaviour besides the reference's own unit-test
vectors (tests/test_oracle_kats.py).  CPU only."""
import numpy as np
import pytest

from helpers import oracle_render
from oracle.py_restatement import PyRasterizer
from rusterizer_b200 import mathx, scenes
from rusterizer_b200.mesh import Mesh
from rusterizer_b200.scenes import Draw, Scene
from rusterizer_b200.texture import Texture


def py_render(scene):
    r = PyRasterizer(scene.width, scene.height, msaa=scene.msaa, guard_band=scene.guard_band)
    if scene.texture is not None:
        r.bind_texture(0, scene.texture.texels)
    r.view, r.projection = scene.view, scene.projection
    for d in scene.draws:
        r.world = d.world
        r.render(d.mesh.vertices, d.mesh.attributes, d.mesh.indices, d.fs)
    depth, color = r.depth.copy(), r.color.copy()
    return dict(depth=depth, color=color, fb=r.framebuffer())


def crosscheck(scene, min_drawn=50):
    o, p = oracle_render(scene), py_render(scene)
    od, pd = o["depth"].view(np.uint32), p["depth"].view(np.uint32)
    assert np.array_equal(od, pd), f"{scene.name}: depth bits differ at {np.argwhere(od != pd)[:5].tolist()}"
    assert np.array_equal(o["color"], p["color"]), f"{scene.name}: colour samples differ at {np.argwhere(o['color'] != p['color'])[:5].tolist()}"
    assert np.array_equal(o["fb"], p["fb"]), f"{scene.name}: resolved image differs"
    assert (o["fb"] != 0xFF191919).sum() > min_drawn, f"{scene.name}: nothing was drawn"
    return o


@pytest.mark.parametrize("fs", [0, 1, 2])
def test_default_scene_all_shaders(fs):
    """The crate's Mode::Demo frame (cube + sphere, two draws, main.rs:94-104) with FS Texture / Color / Debug."""
    crosscheck(scenes.default_scene(1.0, fs=fs, width=192, height=108))


@pytest.mark.parametrize("elapsed", [0.0, 0.7, 2.5])
def test_clip_scene(elapsed):
    """The crate's --clip-test triangle (main.rs:106-125): Sutherland-Hodgman, fan, attribute interpolation."""
    crosscheck(scenes.clip_test_scene(elapsed, width=160, height=90))


def test_small_triangles_sphere():
    crosscheck(scenes.sphere_scene(49, 25, width=200, height=120))


@pytest.mark.parametrize("b2f", [True, False])
def test_overdraw_order(b2f):
    """Stacked jittered grids drawn back-to-front / front-to-back: submission order and the strict < depth test."""
    crosscheck(scenes.overdraw_scene(nx=12, ny=7, width=160, height=96, back_to_front=b2f))


def test_near_plane_field():
    """Every triangle straddles the near plane (BASELINE configs[2] scaled down)."""
    o = crosscheck(scenes.near_clip_scene(nx=10, ny=6, width=192, height=108))
    assert o["counters"]["n_clipped_in"] > 50


def test_random_soup_and_coincident():
    """Random triangles of all sizes and windings, partly off screen, plus exact duplicates (the second copy must
    lose the strict < test) and an RGB (3-byte) non-square texture."""
    rng = np.random.RandomState(5)
    nt = 160
    ctr = rng.uniform(-3, 3, (nt, 1, 3)).astype(np.float32)
    ctr[..., 2] = rng.uniform(-3.5, 6, (nt, 1)).astype(np.float32)
    size = (10 ** rng.uniform(-1.5, 0.4, (nt, 1, 1))).astype(np.float32)
    verts = (ctr + rng.uniform(-1, 1, (nt, 3, 3)).astype(np.float32) * size).reshape(nt, 3, 3)
    verts = np.concatenate([verts, verts[:40]], 0)  # coincident copies, submitted later
    nt = verts.shape[0]
    attrs = rng.uniform(0, 1, (nt * 3, 6)).astype(np.float32)
    mesh = Mesh(verts.reshape(-1, 3), np.arange(nt * 3, dtype=np.uint32), attrs)
    for fs in (1, 0):
        s = scenes.sphere_scene(width=176, height=100, mesh=mesh, fs=fs)
        if fs == 0:
            tex = np.random.RandomState(9).randint(0, 256, (23, 37, 3)).astype(np.uint8)
            s = Scene(s.name + "_rgb", s.width, s.height, s.view, s.projection, s.draws, Texture(tex))
        crosscheck(s)


def test_wide_range_colours_expose_low_bits():
    """FS Color with attribute values up to 6e4: Color::to_argb neither clamps nor masks (color.rs:15-20), so the
    integer part of c * 255 (up to 2^24, one ulp = 1) lands in the packed sample and a last-bit difference in the
    perspective-correct interpolation (five divisions, mod.rs:85-99) changes the colour sample."""
    rng = np.random.RandomState(21)
    nt = 120
    ctr = rng.uniform(-2.5, 2.5, (nt, 1, 3)).astype(np.float32)
    ctr[..., 2] = rng.uniform(-3.0, 5, (nt, 1)).astype(np.float32)
    verts = (ctr + rng.uniform(-1, 1, (nt, 3, 3)).astype(np.float32) * np.float32(0.8)).reshape(-1, 3)
    attrs = rng.uniform(0, 60000, (nt * 3, 6)).astype(np.float32)
    mesh = Mesh(verts, np.arange(nt * 3, dtype=np.uint32), attrs)
    crosscheck(scenes.sphere_scene(width=160, height=96, mesh=mesh, fs=1))


@pytest.mark.parametrize("seed", [3, 4])
def test_non_finite_inputs(seed):
    """inf / NaN / 1e30 coordinates and NaN colours: NaN-ignoring min/max of the bounding box (bounding_box.rs:16-25),
    saturating `as usize` / `as u32` casts, NaN-propagating clamp, the tie-break on NaN edge values (mod.rs:148-170)
    and comparisons that are false for NaN in try_clip (clipping.rs:86-104) must come out the same in both."""
    rng = np.random.RandomState(seed)
    nt = 120
    verts = rng.uniform(-2, 2, (nt, 3, 3)).astype(np.float32)
    verts[..., 2] = rng.uniform(-2, 4, (nt, 3)).astype(np.float32)
    special = np.array([np.inf, -np.inf, np.nan, 1e30, -1e30, 1e19, 3e38], np.float32)
    for t in range(0, nt, 3):
        for _ in range(rng.randint(1, 3)):
            verts[t, rng.randint(3), rng.randint(3)] = special[rng.randint(len(special))]
    attrs = rng.uniform(0, 1, (nt * 3, 6)).astype(np.float32)
    attrs[rng.randint(0, nt * 3, 15), rng.randint(0, 4, 15)] = np.nan
    mesh = Mesh(verts.reshape(-1, 3), np.arange(nt * 3, dtype=np.uint32), attrs)
    with np.errstate(all="ignore"):
        crosscheck(scenes.sphere_scene(width=96, height=60, mesh=mesh, fs=1))


def sample_grid_scene(seed, n_tris=400, size=64):
    """Triangles whose vertices sit on the 1/8-pixel lattice of a size x size target (identity matrices, w = 1), so
    that many edges pass EXACTLY through RGSS sample positions and shared edges abound: exercises the tie-break of
    EdgeFunctions::inside (mod.rs:148-170) and equal depths under the strict < test."""
    rng = np.random.RandomState(seed)
    base = rng.randint(0, size * 8 - 40, (n_tris, 1, 2))
    sxy = (base + rng.randint(0, 41, (n_tris, 3, 2))).astype(np.float32) / np.float32(8.0)  # screen coordinates
    pos = np.empty((n_tris, 3, 3), np.float32)
    pos[..., 0] = sxy[..., 0] * np.float32(2.0 / size) - np.float32(1.0)
    pos[..., 1] = np.float32(1.0) - sxy[..., 1] * np.float32(2.0 / size)
    pos[..., 2] = (rng.randint(-4, 5, (n_tris, 1)) / np.float32(8.0)).astype(np.float32)  # few distinct depths: ties
    # every triangle twice, the copy with reversed winding, so both orientations of every edge occur
    pos = np.concatenate([pos, pos[:, ::-1]], 0)
    attrs = rng.uniform(0, 1, (pos.shape[0] * 3, 6)).astype(np.float32)
    mesh = Mesh(pos.reshape(-1, 3), np.arange(pos.shape[0] * 3, dtype=np.uint32), attrs)
    eye = mathx.identity()
    return Scene(f"sample_grid_{seed}", size, size, eye, eye, [Draw(mesh, eye, 1)], Texture.checkerboard())


@pytest.mark.parametrize("seed", [1, 2])
def test_edges_through_sample_points(seed):
    sc = sample_grid_scene(seed)
    # the lattice construction must survive the viewport transform exactly, or the test is not testing ties
    v = sc.draws[0].mesh.vertices
    sx = np.float32(sc.width) * (v[:, 0] + np.float32(1.0)) / np.float32(2.0)
    assert np.array_equal(sx * 8, np.round(sx * 8))
    crosscheck(sc)


@pytest.mark.parametrize("shape", [(17, 29, 4), (31, 8, 3), (2, 2, 4)])
def test_texture_sample_function(shape):
    """Texture::sample (texture.rs:65-83) as exact f32: 8-bit frame colours hide last-bit differences of the
    bilinear blend, so the two restatements are also compared on the four output floats, bit for bit, for random
    (u, v) and for the lattice points where floor == ceil."""
    from oracle.oracle import get_lib
    from oracle.py_restatement import Texture as PyTexture

    L = get_lib()
    rng = np.random.RandomState(shape[0])
    tex = rng.randint(0, 256, shape).astype(np.uint8)
    h, w, _ = shape
    u = rng.uniform(0, 1, 3000).astype(np.float32)
    v = rng.uniform(0, 1, 3000).astype(np.float32)
    lat_u = (np.arange(w, dtype=np.float32) / np.float32(w - 1)).astype(np.float32)
    lat_v = (np.arange(h, dtype=np.float32) / np.float32(h - 1)).astype(np.float32)
    u = np.concatenate([u, np.repeat(lat_u, h), np.float32([0, 1, 0, 1])])
    v = np.concatenate([v, np.tile(lat_v, w), np.float32([0, 0, 1, 1])])
    got = np.stack(PyTexture(tex).sample(u, v), -1).astype(np.float32)
    want = np.stack([L.tex_sample(tex, float(a), float(b))[0] for a, b in zip(u, v)])
    assert np.array_equal(got.view(np.uint32), want.view(np.uint32))


@pytest.mark.parametrize("seed", range(12))
def test_seeded_fuzz(seed):
    """Seeded random scenes of four kinds (lattice triangles, wide-range colour soup, the demo scene and the clip-test
    triangle at random times with random shaders); 200 further seeds were run once without a mismatch."""
    rng = np.random.RandomState(1000 + seed)
    kind = seed % 4
    if kind == 0:
        sc = sample_grid_scene(100 + seed, n_tris=150, size=32)
    elif kind == 1:
        nt = 60
        ctr = rng.uniform(-3, 3, (nt, 1, 3)).astype(np.float32)
        ctr[..., 2] = rng.uniform(-4.5, 6, (nt, 1)).astype(np.float32)
        size = (10 ** rng.uniform(-1.5, 0.6, (nt, 1, 1))).astype(np.float32)
        verts = (ctr + rng.uniform(-1, 1, (nt, 3, 3)).astype(np.float32) * size).reshape(-1, 3)
        attrs = rng.uniform(0, 50000, (nt * 3, 6)).astype(np.float32)
        sc = scenes.sphere_scene(width=96, height=64, mesh=Mesh(verts, np.arange(nt * 3, dtype=np.uint32), attrs), fs=1)
    elif kind == 2:
        sc = scenes.default_scene(float(rng.uniform(0, 6)), fs=int(rng.randint(0, 3)), width=96, height=54)
    else:
        sc = scenes.clip_test_scene(float(rng.uniform(0, 6.3)), fs=int(rng.randint(0, 3)), width=96, height=54)
    o, p = oracle_render(sc), py_render(sc)
    assert np.array_equal(o["depth"].view(np.uint32), p["depth"].view(np.uint32))
    assert np.array_equal(o["color"], p["color"])
    assert np.array_equal(o["fb"], p["fb"])


def test_two_frames_clear_between():
    """resolve_and_clear leaves clear depth and colour behind: the second frame of a context equals a fresh one."""
    sc = scenes.sphere_scene(33, 17, width=128, height=80, fs=1)
    r = PyRasterizer(sc.width, sc.height)
    r.view, r.projection = sc.view, sc.projection
    out = []
    for _ in range(2):
        for d in sc.draws:
            r.world = d.world
            r.render(d.mesh.vertices, d.mesh.attributes, d.mesh.indices, d.fs)
        out.append(r.framebuffer())
    assert np.array_equal(out[0], out[1])
    assert np.array_equal(out[0], oracle_render(sc)["fb"])


@pytest.mark.parametrize("msaa", [1, 2, 8])
def test_runtime_sample_counts(msaa):
    """N_MSAA_SAMPLES (mod.rs:23) as a runtime value -- sample pattern, CoverageMask::all(), the
    shading-position rule and the box filter all follow the count.  Both restatements agree per sample."""
    for mk in (lambda: scenes.default_scene(1.0, width=160, height=90), lambda: scenes.clip_test_scene(0.7, width=128, height=72),
               lambda: scenes.overdraw_scene(nx=10, ny=6, width=128, height=80), lambda: scenes.sphere_scene(33, 17, width=128, height=80, fs=2)):
        sc = mk()
        sc.msaa = msaa
        o = crosscheck(sc, min_drawn=20)
        assert o["depth"].shape[-1] == msaa


@pytest.mark.parametrize("guard", [1.5, 4.0])
def test_guard_band_clipping(guard):
    """guard-band clipping (mod.rs:417-419).  The side planes move out to |x|, |y| <= g * w, so
    triangles that leave the viewport inside the band are rasterised unclipped (only their pixel box is bounded); the
    image differs from g = 1 by at most the rounding of the clipped attributes."""
    for mk in (lambda: scenes.clip_test_scene(0.7, width=128, height=72), lambda: scenes.near_clip_scene(20, 10, 160, 90),
               lambda: scenes.fullscreen_quad_scene(96, 96)):
        sc = mk()
        ref = oracle_render(sc)
        sc.guard_band = guard
        o = crosscheck(sc, min_drawn=20)
        assert o["counters"]["n_clipped_in"] <= ref["counters"]["n_clipped_in"]
        a, b = ref["fb"].view(np.uint8).astype(np.int16), o["fb"].view(np.uint8).astype(np.int16)
        # same picture: only pixels on a checkerboard edge (where the last bit of u, v picks the texel) may change visibly
        assert (np.abs(a - b).reshape(sc.height, sc.width, 4).max(-1) > 2).mean() < 0.01
        assert ((ref["fb"] != 0xFF191919) == (o["fb"] != 0xFF191919)).mean() > 0.999


def test_oracle_debug_assert_mode():
    """the reference's range debug_assert!s are compiled out of --release; the oracle reports how often
    each would have fired in a debug build.  The crate's own scenes are clean; a guard band deliberately leaves the
    [-1, 1] NDC range (mod.rs:319-321), and attributes of a (1, 1, x) pattern push u past 1.0."""
    from oracle.oracle import debug_asserts

    debug_asserts()
    oracle_render(scenes.default_scene(1.0, width=160, height=90))
    oracle_render(scenes.sphere_scene(33, 17, width=160, height=90))
    clean = debug_asserts()
    assert clean["ndc_range (mod.rs:319-321)"] == 0 and clean["z_range (mod.rs:329)"] == 0 and clean["texel_xy_range (texture.rs:49-50)"] == 0
    sc = scenes.clip_test_scene(0.7, width=128, height=72)
    sc.guard_band = 4.0
    oracle_render(sc)
    assert debug_asserts()["ndc_range (mod.rs:319-321)"] > 0


def test_fuzz_frames_agree():
    """Frames of the fuzz campaign (tests/fuzz_parity.py: random soups with slivers and duplicates, grids, huge triangles,
    depths across the near and far planes, 1/2/4/8 samples, guard bands) through both restatements.  The campaign itself
    compares the CUDA path with the C oracle on a GPU box; this is the CPU-side check that the oracle it leans on agrees
    with the independently written one on the same kind of input."""
    import fuzz_parity

    done, seed = 0, 7000
    while done < 16:
        s = fuzz_parity.make_case(seed)
        seed += 1
        n_tris = sum(d.mesh.indices.size // 3 for d in s.draws)
        if n_tris > 2500 or s.width * s.height > 80000 or getattr(s, "scissor", None):
            continue  # keep the CPU suite quick; the second restatement has no scissor rect
        o, p = oracle_render(s), py_render(s)
        assert np.array_equal(o["depth"].view(np.uint32), p["depth"].view(np.uint32)), f"seed {seed - 1}: depth bits differ"
        assert np.array_equal(o["color"], p["color"]), f"seed {seed - 1}: colour samples differ"
        assert np.array_equal(o["fb"], p["fb"]), f"seed {seed - 1}: resolved image differs"
        done += 1
