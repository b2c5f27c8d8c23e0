"""Known-answer tests that PIN the CPU oracle to the reference's own unit tests.

Every test below is a port of a `#[test]` in the reference crate (file:line in each
docstring, paths relative to /root/reference/src).  All comparisons are exact f32 /
integer equality, like the `assert_eq!`s they come from.
"""
import math

import numpy as np
import pytest

F = np.float32
RED = [1.0, 0.0, 0.0, 1.0, 0.0, 0.0]  # Color::red() + uvs [0,0]
ATTRS_RED = [RED, RED, RED]


def eq(a, b):
    a, b = np.asarray(a, np.float32), np.asarray(b, np.float32)
    assert a.shape == b.shape, (a.shape, b.shape)
    assert np.array_equal(a, b), f"\n{a!r}\n!=\n{b!r}"


# ---------------------------------------------------------------- rasterizer/mod.rs

def test_perspective_divide(oracle_lib):
    """rasterizer/mod.rs:530-560"""
    clip = [[-0.5, 0.9, 0.0, 10.0], [0.08, 0.3, 0.0, 2.0], [0.5, -0.3, 0.0, 1.0]]
    expected = [[-0.05, 0.089999996, 0.0, 10.0], [0.04, 0.15, 0.0, 2.0], [0.5, -0.3, 0.0, 1.0]]
    eq(oracle_lib.perspective_divide(clip), expected)


def test_viewport_transform_1(oracle_lib):
    """rasterizer/mod.rs:563-608"""
    ndc = [[-1.0, 0.5, -0.5, 5.0], [1.0, 0.5, 0.0, 6.0], [0.0, -0.5, 0.5, 7.0]]
    r = oracle_lib.viewport_setup(400, 500, ndc)
    eq(r["depths_camera_space"], [5.0, 6.0, 7.0])
    eq(r["depths"], [0.25, 0.5, 0.75])
    assert r["inv_2x_area"] == F(0.00001)
    eq(r["points"], [[0.0, 125.0], [400.0, 125.0], [200.0, 375.0]])


def test_viewport_transform_2(oracle_lib):
    """rasterizer/mod.rs:611-655"""
    ndc = [[-0.25, 1.0, -1.0, 5.0], [0.5, 0.0, 0.0, 6.0], [0.25, -1.0, 1.0, 7.0]]
    r = oracle_lib.viewport_setup(400, 500, ndc)
    eq(r["depths_camera_space"], [5.0, 6.0, 7.0])
    eq(r["depths"], [0.0, 0.5, 1.0])
    assert r["inv_2x_area"] == F(0.00002)
    eq(r["points"], [[150.0, 0.0], [300.0, 250.0], [250.0, 500.0]])


def test_coverage_mask_bit_ops():
    """rasterizer/mod.rs:658-702 -- CoverageMask is a u8 with bit i = sample i; the oracle
    uses the same `(mask & !(1<<i)) | (v<<i)` update (rz_oracle.c eval_cov)."""
    m = 0

    def set_(m, i, v):
        return (m & ~(1 << i) & 0xFF) | (int(v) << i)

    def get(m, i):
        return ((1 << i) & m) != 0

    assert m == 0
    m = set_(m, 0, True)
    assert m != 0 and get(m, 0) and not get(m, 1) and not get(m, 2) and not get(m, 3)
    m = set_(m, 2, True)
    assert get(m, 0) and not get(m, 1) and get(m, 2) and not get(m, 3)
    m = set_(m, 3, False)
    assert get(m, 0) and not get(m, 1) and get(m, 2) and not get(m, 3)
    m = set_(m, 0, False)
    assert not get(m, 0) and get(m, 2) and m != 0
    m = set_(m, 2, False)
    assert m == 0


SCREEN_TRI = [[100.0, 300.0, 0.5], [200.0, 150.0, 0.5], [300.0, 300.0, 0.5]]  # mod.rs:704-720
W567 = [5.0, 6.0, 7.0]


def test_edge_functions_basic(oracle_lib):
    """rasterizer/mod.rs:723-762"""
    for (x, y), mask in [((200, 200), 0b1111), ((99, 299), 0), ((101, 299), 0b1111), ((200, 149), 0),
                         ((200, 151), 0b1111), ((301, 300), 0)]:
        assert oracle_lib.eval_pixel(SCREEN_TRI, W567, ATTRS_RED, x, y)["mask"] == mask, (x, y)


def test_edge_functions_partial(oracle_lib):
    """rasterizer/mod.rs:765-779"""
    for (x, y), mask in [((299, 299), 0b1100), ((150, 224), 0b0111), ((250, 225), 0b1100)]:
        assert oracle_lib.eval_pixel(SCREEN_TRI, W567, ATTRS_RED, x, y)["mask"] == mask, (x, y)


def test_edge_functions_tie_breaker(oracle_lib):
    """rasterizer/mod.rs:782-801"""
    ins, e, n = oracle_lib.eval_single(SCREEN_TRI, 150.0, 225.0)
    assert ins and e[0] == 0.0 and n[0][0] > 0.0
    ins, e, n = oracle_lib.eval_single(SCREEN_TRI, 250.0, 225.0)
    assert (not ins) and e[1] == 0.0 and n[1][0] < 0.0
    ins, e, n = oracle_lib.eval_single(SCREEN_TRI, 250.0, 300.0)
    assert ins and e[2] == 0.0 and n[2][0] == 0.0 and n[2][1] < 0.0


def test_fragment_creation_same_depth(oracle_lib):
    """rasterizer/mod.rs:804-811"""
    eq(oracle_lib.eval_pixel(SCREEN_TRI, W567, ATTRS_RED, 200, 200)["depths"], [0.5] * 4)


def test_fragment_creation_same_depth_partial_coverage(oracle_lib):
    """rasterizer/mod.rs:814-820"""
    eq(oracle_lib.eval_pixel(SCREEN_TRI, W567, ATTRS_RED, 299, 299)["depths"], [0.0, 0.0, 0.5, 0.5])


def test_fragment_creation_interp_depth(oracle_lib):
    """rasterizer/mod.rs:823-872 (16 exact f32 depths)"""
    tri = [[100.0, 300.0, 0.5], [200.0, 150.0, 0.3], [300.0, 300.0, 0.8]]
    cases = [
        ((101, 299), [0.50039583, 0.5019375, 0.50177085, 0.5002292]),
        ((200, 151), [0.30356252, 0.30510417, 0.3049375, 0.30339584]),
        ((298, 299), [0.7958958, 0.79743755, 0.79727083, 0.79572916]),
        ((200, 258), [0.55322915, 0.5547708, 0.5546042, 0.55306244]),
    ]
    for (x, y), d in cases:
        eq(oracle_lib.eval_pixel(tri, W567, ATTRS_RED, x, y)["depths"], d)


def test_fragment_creation_interp_attr_same_depth(oracle_lib):
    """rasterizer/mod.rs:882-908 (8 exact perspective-correct UVs)"""
    attrs = [[1, 0, 0, 1, 0.0, 0.0], [1, 0, 0, 1, 0.0, 1.0], [1, 0, 0, 1, 1.0, 1.0]]
    cases = [
        ((100, 299), [0.00020831265, 0.006041646]),
        ((200, 150), [0.004791677, 0.99895835]),
        ((299, 299), [0.99645835, 0.9972917]),
        ((200, 258), [0.3641667, 0.6408334]),
    ]
    for (x, y), uv in cases:
        r = oracle_lib.eval_pixel(SCREEN_TRI, [5.0, 5.0, 5.0], attrs, x, y)
        eq(r["attr"][4:6], uv)


# ---------------------------------------------------------------- rasterizer/clipping.rs

def test_clip_fully_inside(oracle_lib):
    """rasterizer/clipping.rs:225-238"""
    kind, _, _ = oracle_lib.try_clip([[-0.5, 0.0, 0.0, 1.0], [0.0, 1.0, 0.0, 1.0], [0.5, 0.0, 0.0, 1.0]], ATTRS_RED)
    assert kind == "inside"


def test_clip_cull_degenerate(oracle_lib):
    """rasterizer/clipping.rs:241-254"""
    kind, _, _ = oracle_lib.try_clip([[0.0, 0.0, 0.0, 1.0], [0.0, 1.0, 0.0, 1.0], [0.0, 0.0, 0.0, 1.0]], ATTRS_RED)
    assert kind == "outside"


def test_clip_outside(oracle_lib):
    """rasterizer/clipping.rs:257-269"""
    kind, _, _ = oracle_lib.try_clip([[-0.6, 1.0, -1.0, 0.5], [0.6, 1.2, -2.0, 0.5], [0.6, 1.0, -1.5, 0.5]], ATTRS_RED)
    assert kind == "outside"


def test_clip_partial_right_side_overlap(oracle_lib):
    """rasterizer/clipping.rs:272-304"""
    kind, tris, _ = oracle_lib.try_clip([[1.5, 0.0, 0.0, 2.0], [2.5, 1.0, 0.0, 2.0], [0.6, 1.0, 0.0, 2.0]], ATTRS_RED)
    assert kind == "clipped" and len(tris) == 2
    eq(tris[0], [[1.5, 0.0, 0.0, 2.0], [2.0, 0.5, 0.0, 2.0], [2.0, 1.0, 0.0, 2.0]])
    eq(tris[1], [[1.5, 0.0, 0.0, 2.0], [2.0, 1.0, 0.0, 2.0], [0.6, 1.0, 0.0, 2.0]])


def test_clip_two_side_overlap(oracle_lib):
    """rasterizer/clipping.rs:307-346"""
    kind, tris, _ = oracle_lib.try_clip(
        [[-7.0629444, 5.0629444, 5.060302, 7.0], [-6.0629444, 7.0629444, 5.060302, 7.0],
         [-5.0629444, 5.0629444, 5.060302, 7.0]], ATTRS_RED)
    assert kind == "clipped" and len(tris) == 3
    eq(tris[0], [[-7.0, 5.0629444, 5.060302, 7.0], [-7.0, 5.188833, 5.060302, 7.0], [-6.0944166, 7.0, 5.060302, 7.0]])
    eq(tris[1], [[-7.0, 5.0629444, 5.060302, 7.0], [-6.0944166, 7.0, 5.060302, 7.0], [-6.031472, 7.0, 5.060302, 7.0]])
    eq(tris[2], [[-7.0, 5.0629444, 5.060302, 7.0], [-6.031472, 7.0, 5.060302, 7.0],
                 [-5.0629444, 5.0629444, 5.060302, 7.0]])


def test_clip_partial_right_side_overlap_single(oracle_lib):
    """rasterizer/clipping.rs:349-374"""
    kind, tris, _ = oracle_lib.try_clip([[2.4, 0.0, 0.0, 2.0], [2.5, 1.0, 0.0, 2.0], [0.6, 1.0, 0.0, 2.0]], ATTRS_RED)
    assert kind == "clipped" and len(tris) == 1
    eq(tris[0], [[2.0, 0.22222227, 0.0, 2.0], [2.0, 1.0, 0.0, 2.0], [0.6, 1.0, 0.0, 2.0]])


def test_clip_on_right_edge(oracle_lib):
    """rasterizer/clipping.rs:377-389"""
    kind, _, _ = oracle_lib.try_clip(
        [[5.1227903, -1.0, 5.060302, 7.0], [7.0, -0.7544193, 5.060302, 7.0], [7.0, -1.0, 5.060302, 7.0]], ATTRS_RED)
    assert kind == "inside"


def test_clip_late_outside(oracle_lib):
    """rasterizer/clipping.rs:392-406"""
    kind, _, _ = oracle_lib.try_clip(
        [[-4.7000513, -4.7000513, 1.3230639, 3.2999487], [-3.7000513, -2.7000513, 1.3230639, 3.2999487],
         [-2.7000513, -4.7000513, 1.3230639, 3.2999487]], ATTRS_RED)
    assert kind == "outside"


def test_clip_complete_coverage(oracle_lib):
    """rasterizer/clipping.rs:409-426"""
    kind, tris, _ = oracle_lib.try_clip(
        [[-10.700051, 10.000513, 1.3, 3.2999487], [15.700051, 0.0, 1.3230639, 1.3],
         [-10.700051, -10.700051, 1.3, 3.2999487]], ATTRS_RED)
    assert kind == "clipped" and len(tris) == 2


CLIP_SMALL = [[1.6960512, -2.4062815, 0.42041498, 2.4062815], [1.6891757, -2.4017043, 0.41579157, 2.401704],
              [1.686294, -2.3693168, 0.4157151, 2.4016285]]


def test_clip_small_triangle(oracle_lib):
    """rasterizer/clipping.rs:429-447"""
    kind, tris, _ = oracle_lib.try_clip(CLIP_SMALL, ATTRS_RED)
    assert kind == "clipped" and len(tris) == 2


def test_clipped_tris_are_inside(oracle_lib):
    """rasterizer/clipping.rs:450-470"""
    kind, tris, attrs = oracle_lib.try_clip(
        [[2.0700936, -3.0516267, 0.98517126, 2.965418], [2.0748749, -2.6156826, 1.0383241, 3.018042],
         [2.0742767, -2.8663764, 1.1517053, 3.130295]], ATTRS_RED)
    assert kind == "clipped"
    for t, a in zip(tris, attrs):
        assert oracle_lib.try_clip(t, a)[0] == "inside"


# ---------------------------------------------------------------- rasterizer/buffers.rs

RED_C, BLUE_C, GREEN_C = 0xFFFF0000, 0xFF0000FF, 0xFF00FF00


def test_average_same_color(oracle_lib):
    """rasterizer/buffers.rs:173-177"""
    for c in (BLUE_C, GREEN_C, RED_C):
        assert oracle_lib.box_filter_color([c] * 4) == c


def test_average_two_colors(oracle_lib):
    """rasterizer/buffers.rs:180-196"""
    assert oracle_lib.box_filter_color([RED_C, BLUE_C, RED_C, BLUE_C]) == 0xFF7F007F
    assert oracle_lib.box_filter_color([RED_C, RED_C, BLUE_C, BLUE_C]) == 0xFF7F007F
    assert oracle_lib.box_filter_color([RED_C, GREEN_C, RED_C, GREEN_C]) == 0xFF7F7F00


def test_average_three_colors(oracle_lib):
    """rasterizer/buffers.rs:199-204"""
    assert oracle_lib.box_filter_color([RED_C, GREEN_C, RED_C, BLUE_C]) == 0xFF7F3F3F


def test_average_colors(oracle_lib):
    """rasterizer/buffers.rs:207-213"""
    assert oracle_lib.box_filter_color([0xFF35B565, 0xFFF3FA12, 0xFF3E5469, 0xFF435623]) == 0xFF6A9640


def test_buffer_tiles_pow_2_square(oracle_lib):
    """rasterizer/buffers.rs:216-254 (grid geometry + tile_idx)"""
    T = 64
    assert oracle_lib.tile_grid(128, 128) == (2, 2)
    assert oracle_lib.tile_idx(128, 0, 0) == 0
    assert oracle_lib.tile_idx(128, T - 1, T - 1) == 0
    assert oracle_lib.tile_idx(128, 127, 127) == 3
    assert oracle_lib.tile_idx(128, 64, 64) == 64 // T * (128 // T) + 64 // T


def test_buffer_tiles_uneven_rect(oracle_lib):
    """rasterizer/buffers.rs:257-318"""
    T = 64
    assert oracle_lib.tile_grid(442, 711) == (7, 12)
    assert oracle_lib.tile_idx(442, 0, 0) == 0
    assert oracle_lib.tile_idx(442, T - 1, T - 1) == 0
    assert oracle_lib.tile_idx(442, 710, 441) == 7 * 12 - 1
    assert oracle_lib.tile_idx(442, 64, 64) == 64 // T * (442 // T + 1) + 64 // T


def test_buffer_tiles_double_buffer():
    """rasterizer/buffers.rs:299-317: marked()/prev_marked()/next() through whole frames"""
    from oracle.oracle import OracleRenderer

    r = OracleRenderer(442, 711)
    # one clip-space triangle covering a few pixels near the top-left, Color FS
    tri = [[[-0.99, 0.99, 0.0, 1.0], [-0.9, 0.99, 0.0, 1.0], [-0.99, 0.9, 0.0, 1.0]]]
    r.rasterize(tri, [[RED, RED, RED]], fs_id=1)
    assert r.tiles_marked() == 1 and r.tiles_marked(prev=True) == 0
    r.framebuffer()
    assert r.tiles_marked() == 0 and r.tiles_marked(prev=True) == 1
    r.framebuffer()
    assert r.tiles_marked() == 0 and r.tiles_marked(prev=True) == 0


# ---------------------------------------------------------------- rasterizer/bounding_box.rs

def test_bounding_box(oracle_lib):
    """rasterizer/bounding_box.rs:48-74"""
    assert oracle_lib.pixel_bbox([[100.0, 200.0], [230.0, 200.0], [230.0, 300.0]]) == (100, 230, 200, 300)
    assert oracle_lib.pixel_bbox([[50.9, 200.0], [230.0, 100.0], [500.0, 200.9]]) == (50, 500, 100, 201)


# ---------------------------------------------------------------- color.rs

def test_argb(oracle_lib):
    """color.rs:130-155"""
    assert oracle_lib.to_argb([1, 1, 1, 1]) == 0xFFFFFFFF
    assert oracle_lib.to_argb([1, 0, 0, 1]) == 0xFFFF0000
    assert oracle_lib.to_argb([0, 1, 0, 1]) == 0xFF00FF00
    assert oracle_lib.to_argb([0, 0, 1, 1]) == 0xFF0000FF
    assert oracle_lib.to_argb([0, 0, 0, 1]) == 0xFF000000


# ---------------------------------------------------------------- math/matrix.rs

MAT = np.arange(1, 17, dtype=np.float32).reshape(4, 4)


def test_matrix_mul_identity(oracle_lib):
    """math/matrix.rs:217-224"""
    eq(oracle_lib.mat4_mul(np.eye(4), np.eye(4)), np.eye(4))


def test_matrix_rows_cols_transpose():
    """math/matrix.rs:227-262: row-major [[f32;4];4]; row(i), col(j), transpose -- the layout
    the oracle and the C ABI use (numpy row-major 4x4)."""
    eq(MAT[0], [1, 2, 3, 4])
    eq(MAT[3], [13, 14, 15, 16])
    eq(MAT[:, 0], [1, 5, 9, 13])
    eq(MAT[:, 3], [4, 8, 12, 16])
    eq(MAT.T, [[1, 5, 9, 13], [2, 6, 10, 14], [3, 7, 11, 15], [4, 8, 12, 16]])


def test_matrix_mul(oracle_lib):
    """math/matrix.rs:265-286"""
    eq(oracle_lib.mat4_mul(MAT, MAT),
       [[90, 100, 110, 120], [202, 228, 254, 280], [314, 356, 398, 440], [426, 484, 542, 600]])
    eq(oracle_lib.mat4_mul(MAT, MAT.T),
       [[30, 70, 110, 150], [70, 174, 278, 382], [110, 278, 446, 614], [150, 382, 614, 846]])


# ---------------------------------------------------------------- math/vector.rs (Mat x Vec)

VECS = [[3.0, 10.34, 1.0, 0.0], [13.0, 10.90, -15.0, 0.0], [-10345.124, 0.9123, -15.0, 0.0],
        [3.0, 10.34, 1.0, 1.0], [13.0, 10.90, -15.0, 1.0], [-10345.124, 0.9123, -15.0, 1.0]]


def test_mat4_mul_identity_vec(oracle_lib):
    """math/vector.rs:383-406"""
    for v in VECS:
        eq(oracle_lib.mat4_vec(np.eye(4), v), v)


def test_mat4_mul_translate(oracle_lib):
    """math/vector.rs:409-436"""
    from rusterizer_b200 import mathx

    t = mathx.translate(4.0, -2.0, 4.5)
    expected = VECS[:3] + [[7.0, 8.34, 5.5, 1.0], [17.0, 8.90, -10.5, 1.0], [-10341.124, -1.0877, -10.5, 1.0]]
    for v, e in zip(VECS, expected):
        eq(oracle_lib.mat4_vec(t, v), e)


def test_mat4_mul_rotate_lh(oracle_lib):
    """math/vector.rs:439-466"""
    from rusterizer_b200 import mathx

    v = [[1.0, 0.0, 0.0], [0.0, 1.0, 0.0], [0.0, 0.0, 1.0]]
    c = -0.00000004371139
    expected = [[1.0, 0.0, 0.0], [0.0, c, 1.0], [0.0, -1.0, c],
                [c, 0.0, -1.0], [0.0, 1.0, 0.0], [1.0, 0.0, c],
                [c, 1.0, 0.0], [-1.0, c, 0.0], [0.0, 0.0, 1.0]]
    h = F(math.pi / 2)
    mats = [mathx.rotate_x(h), mathx.rotate_y(h), mathx.rotate_z(h)]
    for i, m in enumerate(mats):
        for j, vec in enumerate(v):
            for w in (0.0, 1.0):
                eq(oracle_lib.mat4_vec(m, vec + [w]), expected[i * 3 + j] + [w])


def test_vector_len_normalized_arith():
    """math/vector.rs:247-380: host-side f32 vector helpers (len, normalized, neg, mul, div, add)."""
    from rusterizer_b200 import mathx

    assert mathx.vlen([0.0, 0.0, 0.0]) == 0.0
    for v in ([1, 0, 0], [0, 1, 0], [0, 0, 1], [-1, 0, 0], [0, -1, 0], [0, 0, -1]):
        assert mathx.vlen(v) == 1.0
    assert mathx.vlen([3.0, 10.0, 1.0]) == F(10.488089)
    for v in ([3.0, 10.760, 1.0], [8.0, 10.0, 1.0], [3.0, 143.5, 1.0], [63.0, 2234.5, -1.0],
              [23.0, -1546.1, 1324.0], [99.0, 14.0, -123.0]):
        assert mathx.vlen(mathx.normalized(v)) == 1.0
    eq(-mathx.vec([23.0, -1546.1, 1324.0]), [-23.0, 1546.1, -1324.0])
    eq(mathx.vec([3.0, 10.90, 1.0]) * F(10.0), [30.0, 109.0, 10.0])
    eq(mathx.vec([13.0, 10.90, -15.0]) * F(2.0), [26.0, 21.8, -30.0])
    eq(mathx.vec([13.0, 10.90, -15.0]) * F(-3.0), [-39.0, -32.699997, 45.0])
    eq(mathx.vec([3.0, 10.90, 1.0]) / F(10.0), [0.3, 1.0899999, 0.1])
    eq(mathx.vec([13.0, 10.90, -15.0]) / F(2.0), [6.5, 5.45, -7.5])
    eq(mathx.vec([13.0, 10.90, -15.0]) / F(-3.0), [-4.3333335, -3.6333332, 5.0])
    v = [mathx.vec([3.0, 10.34, 1.0]), mathx.vec([13.0, 10.90, -15.0]), mathx.vec([-10345.124, 0.9123, -15.0])]
    expected = [[6.0, 20.68, 2.0], [16.0, 21.24, -14.0], [-10342.124, 11.2523, -14.0],
                [16.0, 21.24, -14.0], [26.0, 21.80, -30.0], [-10332.124, 11.8123, -30.0],
                [-10342.124, 11.2523, -14.0], [-10332.124, 11.8123, -30.0], [-20690.248, 1.8246, -30.0]]
    for i in range(3):
        for j in range(3):
            eq(v[i] + v[j], expected[i * 3 + j])


def test_registry_blend_shader_arithmetic():
    """FS TextureBlend (registry extension, include/rz.h) = (texture.sample(u, v) + attr.color) / 2.0 with one
    IEEE rounding per Color operator (color.rs:88-111).  A full-screen quad with constant colour and constant uv:
    every pixel is pack((texel/255 + colour) / 2)."""
    from oracle.oracle import OracleRenderer

    tex = np.zeros((2, 2, 4), np.uint8)
    tex[...] = (200, 100, 50, 255)
    r = OracleRenderer(16, 16)
    r.bind_texture(0, np.zeros((4, 4, 4), np.uint8))
    r.bind_texture(1, tex)
    pos = np.array([[-1, -1, 0.5], [1, -1, 0.5], [1, 1, 0.5], [-1, 1, 0.5]], np.float32) * np.float32([0.9, 0.9, 1])
    col = np.float32([0.25, 0.5, 0.75, 1.0])
    attrs = np.tile(np.concatenate([col, np.float32([0.5, 0.5])]), (4, 1)).astype(np.float32)
    idx = np.array([0, 2, 1, 0, 3, 2], np.uint32)
    r.render(pos, attrs, idx, 0, 3 | (1 << 8))
    fb = r.framebuffer()
    f = np.float32
    want = [int(f(f(f(f(t) / f(255.0)) + c) / f(2.0)) * f(255.0)) for t, c in zip((200, 100, 50), col[:3])]
    assert fb[8, 8] == (0xFF000000 | (want[0] << 16) | (want[1] << 8) | want[2])
    with pytest.raises(Exception):
        r.render(pos, attrs, idx, 0, 3 | (2 << 8))  # texture 2 is not bound
    r.close()


def test_scissor_bounds_the_bbox_walk():
    """Scissor extension (rasterizer/mod.rs:349-350): the bbox walk is bounded by the rect instead of the viewport,
    so n_bbox_px is exactly the rect's share and pixels outside keep the clear colour."""
    from oracle.oracle import OracleRenderer

    r = OracleRenderer(32, 32)
    pos = np.array([[-1, -1, 0.5], [1, -1, 0.5], [1, 1, 0.5], [-1, 1, 0.5]], np.float32)
    attrs = np.tile(np.float32([1, 0, 0, 1, 0, 0]), (4, 1))
    idx = np.array([0, 2, 1, 0, 3, 2], np.uint32)
    r.set_scissor(5, 7, 20, 9)
    r.render(pos, attrs, idx, 0, 1)
    c = r.counters()
    fb = r.framebuffer()
    assert c["n_bbox_px"] == 2 * (20 - 5) * (9 - 7)      # both triangles' boxes cover the viewport
    inside = np.zeros((32, 32), bool)
    inside[7:9, 5:20] = True
    assert (fb[~inside] == 0xFF191919).all() and (fb[inside] == 0xFFFF0000).all()
    r.close()
