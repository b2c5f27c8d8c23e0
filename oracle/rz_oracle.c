/*
 * rz_oracle.c -- CPU ORACLE (TEST INFRASTRUCTURE, NOT PRODUCT CODE)
 *
 * A plain-C restatement of the reference CPU rasteriser's per-frame path
 * (NiklasJonsson/rusterizer).  It exists only so that tests/, smoke() and the
 * cpu_baseline / --impl reference legs of bench.py can check and time the
 * reference algorithm.  Nothing under rusterizer_b200/ may link, load or call
 * it: the product path is CUDA only.
 *
 * Parity status: PINNED against the reference's own unit-test golden vectors
 * (tests/test_oracle_kats.py ports src/rasterizer/mod.rs:525-909,
 * clipping.rs:197-471, buffers.rs:159-319, bounding_box.rs:44-75,
 * color.rs:125-156, math/matrix.rs:196-287, math/vector.rs:242-467).
 * The reference itself cannot be compiled here (no Rust toolchain), so there
 * is no oracle/_ref; whole-frame behaviour (ordering, post-depth shading
 * position, texture sampling, resolve) is pinned by source only.
 *
 * Arithmetic contract: every + - * / below is ONE IEEE-754 binary32 operation
 * in the reference's source order.  Build with -ffp-contract=off and without
 * -ffast-math (oracle/Makefile does).  Each function cites the reference
 * file:line it follows (paths relative to the reference's src/).
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#define N_MSAA 4                       /* rasterizer/mod.rs:23 : the reference's (only) sample count */
#define MAX_MSAA 8                     /* runtime sample counts 1 / 2 / 4 / 8 (extension) */
#define CLEAR_COLOR 0xFF191919u        /* rasterizer/buffers.rs:5 */
#define CLEAR_DEPTH 3.40282347e+38f    /* f32::MAX, rasterizer/buffers.rs:6 */
#define TILE_SIZE 64u                  /* rasterizer/buffers.rs:8 */
#define CULL_EPS 0.000001f             /* rasterizer/clipping.rs:12 */
#define ORC_MAX_POLY 40
#define ORC_NO_OWNER 0xFFFFFFFFu

/* rasterizer/mod.rs:109-114 */
static const float RGSS[N_MSAA][2] = {
    {5.0f / 8.0f, 1.0f / 8.0f},
    {7.0f / 8.0f, 5.0f / 8.0f},
    {3.0f / 8.0f, 7.0f / 8.0f},
    {1.0f / 8.0f, 3.0f / 8.0f},
};

/* Sample patterns for the runtime sample counts the reference does not define (N_MSAA_SAMPLES is a constant there,
 * mod.rs:23,109-114): 1 = the pixel centre, 2 and 8 = the standard patterns of the D3D11 specification (in sixteenths
 * of a pixel, exactly representable).  4 stays the reference's rotated grid. */
static const float PAT1[1][2] = {{0.5f, 0.5f}};
static const float PAT2[2][2] = {{0.75f, 0.75f}, {0.25f, 0.25f}};
static const float PAT8[8][2] = {{0.5625f, 0.3125f}, {0.4375f, 0.6875f}, {0.8125f, 0.5625f}, {0.3125f, 0.1875f},
                                 {0.1875f, 0.8125f}, {0.0625f, 0.4375f}, {0.6875f, 0.9375f}, {0.9375f, 0.0625f}};
static const float (*pattern_of(int ns))[2] { return ns == 1 ? PAT1 : ns == 2 ? PAT2 : ns == 8 ? PAT8 : RGSS; }

typedef struct {
    float v[6]; /* r g b a u v : graphics_primitives.rs:10-13, color.rs:7-12 */
} attr_t;

typedef struct {
    float p[3][4]; /* clip/NDC x y z w */
    attr_t a[3];
} tri_t; /* graphics_primitives.rs:64-71 */

typedef struct {
    uint64_t n_tris_in, n_degenerate, n_outside, n_inside, n_clipped_in;
    uint64_t n_tris_setup, n_bbox_px, n_covered_px, n_shaded_px, n_samples_written;
    uint64_t n_tex_oob, n_clip_overflow;
} orc_counters_t;

typedef struct {
    uint8_t *buf;
    uint32_t w, h, tw;
    size_t len;
} tex_t;

typedef struct orc_ctx {
    uint32_t width, height;
    int ns;          /* samples per pixel: N_MSAA_SAMPLES (mod.rs:23) made a runtime value; default 4 */
    float guard;     /* guard band factor g >= 1 (mod.rs:417-419): x / y clip planes at |x| <= g*w; default 1 = the reference */
    uint32_t row0, row1; /* oracle-only row window [row0,row1): the sample buffers hold just these rows, so a large frame can
                            be checked band by band (one process per band); default = the whole frame */
    uint32_t sc_x0, sc_y0, sc_x1, sc_y1; /* scissor rect (extension suggested at mod.rs:349-350); default = viewport */
    uint32_t *color;   /* [w*h][ns]  buffers.rs:83-105 */
    float *depth;      /* [w*h][ns]  buffers.rs:129-147 */
    uint32_t *owner;   /* [w*h][ns]  oracle-only: order key of the last writer */
    uint32_t *resolve; /* [w*h] */
    /* BufferTiles, buffers.rs:11-79 */
    uint32_t n_horizontal, n_vertical;
    uint8_t *tile_mask[2];
    uint32_t mask_idx;
    float world[16], view[16], proj[16];
    tex_t tex[16];
    uint32_t n_tex;
    uint32_t tri_base; /* running triangle number inside the frame (order key) */
    orc_counters_t cnt;
    float *vs_out; /* scratch */
    size_t vs_cap;
} orc_ctx;

/* "debug_assert mode" (SURVEY.md section 5): the reference guards value ranges with debug_assert!s that are compiled out of
 * --release.  The oracle always computes the release behaviour and only COUNTS how often each of them would have fired
 * in a debug build.  Per process (the static helpers have no context); read and reset with orc_debug_asserts(). */
enum { DA_CLAMP_BARY, DA_NDC_RANGE, DA_Z_RANGE, DA_TEX_UV, DA_TEXEL_XY, DA_COUNT };
static uint64_t g_dassert[DA_COUNT];
void orc_debug_asserts(uint64_t *out5, int reset) {
    for (int i = 0; i < DA_COUNT; i++) {
        if (out5) out5[i] = g_dassert[i];
        if (reset) g_dassert[i] = 0;
    }
}

/* ---------------- math subset ---------------- */

/* math/vector.rs:17-23 : sum starts at 0.0, sequential */
static float dot4(const float *a, const float *b) {
    float s = 0.0f;
    for (int k = 0; k < 4; k++) s = s + a[k] * b[k];
    return s;
}
static float dot2(float ax, float ay, float bx, float by) {
    float s = 0.0f;
    s = s + ax * bx;
    s = s + ay * by;
    return s;
}
/* math/matrix.rs:56-79 : R[i][j] = dot(row_i(A), col_j(B)) */
void orc_mat4_mul(const float *A, const float *B, float *R) {
    float out[16];
    for (int i = 0; i < 4; i++)
        for (int j = 0; j < 4; j++) {
            float col[4] = {B[0 * 4 + j], B[1 * 4 + j], B[2 * 4 + j], B[3 * 4 + j]};
            out[i * 4 + j] = dot4(&A[i * 4], col);
        }
    memcpy(R, out, sizeof out);
}
/* math/vector.rs:219-240 */
void orc_mat4_vec(const float *M, const float *v, float *r) {
    float out[4];
    for (int i = 0; i < 4; i++) out[i] = dot4(&M[i * 4], v);
    memcpy(r, out, sizeof out);
}
/* math/vector.rs:183-185 */
static float cross2(float ax, float ay, float bx, float by) { return ax * by - bx * ay; }

/* rasterizer/mod.rs:15-21 (xy of the first three points) */
static float area2(float x0, float y0, float x1, float y1, float x2, float y2) {
    float v10x = x1 - x0, v10y = y1 - y0;
    float v20x = x2 - x0, v20y = y2 - y0;
    return cross2(v10x, v10y, v20x, v20y);
}
float orc_triangle_2x_area(const float *xy6) {
    return area2(xy6[0], xy6[1], xy6[2], xy6[3], xy6[4], xy6[5]);
}

/* Rust f32::clamp(0.0, 1.0): NaN stays NaN.  rasterizer/mod.rs:102-106 */
static float clamp_bary(float x) {
    if (!(x >= 0.0f - 0.0001f && x <= 1.0f + 0.0001f)) g_dassert[DA_CLAMP_BARY]++; /* debug_assert! mod.rs:104 */
    if (x < 0.0f) x = 0.0f;
    if (x > 1.0f) x = 1.0f;
    return x;
}

/* Rust `f32 as usize`: saturating, NaN -> 0 */
static uint64_t f32_as_usize(float f) {
    if (!(f > 0.0f)) return 0;
    if (f >= 18446744073709551616.0f) return UINT64_MAX;
    return (uint64_t)f;
}
/* Rust `f32 as u32`: saturating, NaN -> 0 */
static uint32_t f32_as_u32(float f) {
    if (!(f > 0.0f)) return 0;
    if (f >= 4294967296.0f) return 0xFFFFFFFFu;
    return (uint32_t)f;
}

/* attribute arithmetic, graphics_primitives.rs:21-61 + color.rs:75-123 */
static attr_t attr_sub(attr_t a, attr_t b) { attr_t r; for (int k = 0; k < 6; k++) r.v[k] = a.v[k] - b.v[k]; return r; }
static attr_t attr_add(attr_t a, attr_t b) { attr_t r; for (int k = 0; k < 6; k++) r.v[k] = a.v[k] + b.v[k]; return r; }
static attr_t attr_mul(attr_t a, float s) { attr_t r; for (int k = 0; k < 6; k++) r.v[k] = a.v[k] * s; return r; }

/* ---------------- color ---------------- */

/* color.rs:15-20 : truncating, saturating, fields OR-ed without masking */
uint32_t orc_to_argb(const float *rgba) {
    return (f32_as_u32(rgba[3] * 255.0f) << 24) | (f32_as_u32(rgba[0] * 255.0f) << 16) |
           (f32_as_u32(rgba[1] * 255.0f) << 8) | f32_as_u32(rgba[2] * 255.0f);
}

/* rasterizer/buffers.rs:111-125 with N_MSAA_SAMPLES = ns */
static uint32_t box_filter_n(const uint32_t *c, int ns) {
    uint32_t r = 0, g = 0, b = 0;
    for (int i = 0; i < ns; i++) {
        r += (c[i] & 0x00FF0000u) >> 16;
        g += (c[i] & 0x0000FF00u) >> 8;
        b += c[i] & 0x000000FFu;
    }
    return (0xFFu << 24) | ((r / (uint32_t)ns) << 16) | ((g / (uint32_t)ns) << 8) | (b / (uint32_t)ns);
}
uint32_t orc_box_filter_color(const uint32_t *c) { return box_filter_n(c, N_MSAA); }

/* ---------------- clipping ---------------- */

/* rasterizer/clipping.rs:29-38.  Guard band (extension sketched at mod.rs:417-419): the four side planes move out to
 * |x|, |y| <= g*w; g*w is one f32 product, and 1.0f * w == w, so g = 1 is the reference bit for bit. */
static float distance_measure(int plane, const float *p, float guard) {
    switch (plane) {
    case 0: return guard * p[3] + p[0]; /* LEFT   */
    case 1: return guard * p[3] - p[0]; /* RIGHT  */
    case 2: return guard * p[3] + p[1]; /* BOTTOM */
    case 3: return guard * p[3] - p[1]; /* TOP    */
    case 4: return p[3] + p[2]; /* NEAR   */
    default: return p[3] - p[2]; /* FAR   */
    }
}

/* rasterizer/clipping.rs:43-51 ; point ops math/point.rs:108-117,130-139 */
static float compute_intersection(const float *p0, float d0, const float *p1, float d1, float *out) {
    float alpha = d0 / (d0 - d1);
    float one_minus = 1.0f - alpha;
    for (int k = 0; k < 4; k++) out[k] = p0[k] * one_minus + p1[k] * alpha;
    return alpha;
}

/*
 * rasterizer/clipping.rs:62-195.
 * Returns: -1 = Outside, 0 = Inside, n>0 = Clipped into n triangles written to out[].
 * *overflow is set when the polygon outgrew ORC_MAX_POLY (never expected).
 */
static int try_clip(const tri_t *t, tri_t *out, int out_cap, int *overflow, float guard) {
    if (fabsf(area2(t->p[0][0], t->p[0][1], t->p[1][0], t->p[1][1], t->p[2][0], t->p[2][1])) < CULL_EPS)
        return -1;

    int inside[3][2] = {{1, 1}, {1, 1}, {1, 1}};
    int outside[3][2] = {{1, 1}, {1, 1}, {1, 1}};
    for (int i = 0; i < 3; i++) {
        const float *v = t->p[i];
        float w = v[3], nw = -v[3];
        for (int ax = 0; ax < 3; ax++) {
            /* guard band: "inside" (no clipping needed) is judged against the widened side planes, "outside" (nothing
             * can be visible) still against the view frustum itself; z is never widened */
            float gw = ax < 2 ? guard * w : w, ngw = -gw;
            inside[ax][0] &= v[ax] >= ngw;
            inside[ax][1] &= v[ax] <= gw;
            outside[ax][0] &= v[ax] < nw;
            outside[ax][1] &= v[ax] > w;
        }
    }
    int any_out = 0, all_in = 1;
    for (int ax = 0; ax < 3; ax++)
        for (int s = 0; s < 2; s++) {
            any_out |= outside[ax][s];
            all_in &= inside[ax][s];
        }
    if (any_out) return -1;
    if (all_in) return 0;

    float ov[ORC_MAX_POLY][4], iv[ORC_MAX_POLY][4];
    attr_t oa[ORC_MAX_POLY], ia[ORC_MAX_POLY];
    int n_out = 3;
    for (int i = 0; i < 3; i++) {
        memcpy(ov[i], t->p[i], sizeof ov[i]);
        oa[i] = t->a[i];
    }
    for (int plane = 0; plane < 6; plane++) {
        int n_in = n_out;
        memcpy(iv, ov, sizeof(float) * 4 * n_in);
        memcpy(ia, oa, sizeof(attr_t) * n_in);
        n_out = 0;
        for (int i = 0; i < n_in; i++) {
            int prev_i = (i + n_in - 1) % n_in;
            const float *pv = iv[prev_i], *cv = iv[i];
            float pd = distance_measure(plane, pv, guard);
            float cd = distance_measure(plane, cv, guard);
            int pin = pd >= 0.0f, cin = cd >= 0.0f;
            if (n_out + 2 > ORC_MAX_POLY) {
                *overflow = 1;
                break;
            }
            if (pin && cin) {
                memcpy(ov[n_out], cv, sizeof ov[0]);
                oa[n_out++] = ia[i];
            } else if (pin && !cin) {
                float alpha = compute_intersection(pv, pd, cv, cd, ov[n_out]);
                oa[n_out++] = attr_add(attr_mul(attr_sub(ia[i], ia[prev_i]), alpha), ia[prev_i]);
            } else if (!pin && cin) {
                float alpha = compute_intersection(pv, pd, cv, cd, ov[n_out]);
                oa[n_out++] = attr_add(attr_mul(attr_sub(ia[i], ia[prev_i]), alpha), ia[prev_i]);
                memcpy(ov[n_out], cv, sizeof ov[0]);
                oa[n_out++] = ia[i];
            }
        }
    }
    if (n_out == 0) return -1;
    if (n_out < 3) return -1; /* reference would underflow usize (App. B-11); S-H never yields 1-2 */
    int n_tris = n_out - 2;
    if (n_tris > out_cap) {
        *overflow = 1;
        n_tris = out_cap;
    }
    for (int i = 0; i < n_tris; i++) {
        memcpy(out[i].p[0], ov[0], sizeof ov[0]);
        memcpy(out[i].p[1], ov[i + 1], sizeof ov[0]);
        memcpy(out[i].p[2], ov[i + 2], sizeof ov[0]);
        out[i].a[0] = oa[0];
        out[i].a[1] = oa[i + 1];
        out[i].a[2] = oa[i + 2];
    }
    return n_tris;
}

/* KAT entry: pos[12] (3 x xyzw), attrs[18]; out_pos[cap*12], out_attrs[cap*18] */
int orc_try_clip(const float *pos, const float *attrs, float *out_pos, float *out_attrs, int cap) {
    tri_t t, out[ORC_MAX_POLY];
    memcpy(t.p, pos, sizeof t.p);
    memcpy(t.a, attrs, sizeof t.a);
    int ovf = 0;
    int n = try_clip(&t, out, cap < ORC_MAX_POLY ? cap : ORC_MAX_POLY, &ovf, 1.0f);
    for (int i = 0; i < n; i++) {
        memcpy(out_pos + i * 12, out[i].p, sizeof out[i].p);
        memcpy(out_attrs + i * 18, out[i].a, sizeof out[i].a);
    }
    return n;
}

/* ---------------- raster triangle ---------------- */

typedef struct {
    float px[3], py[3];      /* EdgeFunctions.points   mod.rs:118 */
    float nx[3], ny[3];      /* EdgeFunctions.normals  mod.rs:119 */
    float cov_eval[MAX_MSAA][3];
    uint8_t cov_mask;
    int ns;                  /* N_MSAA_SAMPLES */
    const float (*pat)[2];   /* sample pattern (RGSS_SAMPLE_PATTERN for ns = 4) */
    float w[3];              /* depths_camera_space */
    float z[3];              /* depths */
    attr_t a[3];
    float inv_2x_area;
} rtri_t; /* mod.rs:178-184 */

/* rasterizer/mod.rs:284-313 */
void orc_perspective_divide(const float *clip12, float *ndc12) {
    for (int i = 0; i < 3; i++) {
        const float *v = clip12 + 4 * i;
        ndc12[4 * i + 0] = v[0] / v[3];
        ndc12[4 * i + 1] = v[1] / v[3];
        ndc12[4 * i + 2] = v[2] / v[3];
        ndc12[4 * i + 3] = v[3];
    }
}

/* rasterizer/mod.rs:315-345 + RasterizerTriangle::new mod.rs:187-222 */
static void viewport_setup(uint32_t width, uint32_t height, const float *ndc12, const attr_t *attrs, rtri_t *r) {
    const float zmin = 0.0f, zmax = 1.0f;
    float sx[3], sy[3], sz[3];
    for (int i = 0; i < 3; i++) {
        const float *v = ndc12 + 4 * i;
        sx[i] = (float)width * (v[0] + 1.0f) / 2.0f;
        sy[i] = (float)height * (1.0f - (v[1] + 1.0f) / 2.0f);
        sz[i] = (v[2] + 1.0f) * 0.5f * (zmax - zmin) + zmin;
        for (int a = 0; a < 3; a++)
            if (!(v[a] <= 1.0f && v[a] >= -1.0f)) g_dassert[DA_NDC_RANGE]++; /* debug_assert! mod.rs:319-321 */
        if (!(sz[i] >= zmin && sz[i] <= zmax)) g_dassert[DA_Z_RANGE]++;      /* debug_assert! mod.rs:329 */
        r->w[i] = v[3];
    }
    /* v0 = p1-p0, v1 = p2-p1, v2 = p0-p2 ; n_k = (-v_k.y, v_k.x) */
    float v0x = sx[1] - sx[0], v0y = sy[1] - sy[0];
    float v1x = sx[2] - sx[1], v1y = sy[2] - sy[1];
    float v2x = sx[0] - sx[2], v2y = sy[0] - sy[2];
    r->nx[0] = -v0y; r->ny[0] = v0x;
    r->nx[1] = -v1y; r->ny[1] = v1x;
    r->nx[2] = -v2y; r->ny[2] = v2x;
    r->inv_2x_area = 1.0f / area2(sx[0], sy[0], sx[1], sy[1], sx[2], sy[2]);
    for (int i = 0; i < 3; i++) {
        r->px[i] = sx[i];
        r->py[i] = sy[i];
        r->z[i] = sz[i];
        r->a[i] = attrs[i];
    }
    memset(r->cov_eval, 0, sizeof r->cov_eval);
    r->cov_mask = 0;
    r->ns = N_MSAA;
    r->pat = RGSS;
}

/* rasterizer/mod.rs:125-132 */
static void eval_single(const rtri_t *r, float x, float y, float *e) {
    for (int k = 0; k < 3; k++) e[k] = dot2(r->nx[k], r->ny[k], x - r->px[k], y - r->py[k]);
}

/* rasterizer/mod.rs:148-170 */
static int inside(const rtri_t *r, const float *e) {
    for (int k = 0; k < 3; k++) {
        if (e[k] > 0.0f) continue;
        if (e[k] < 0.0f) return 0;
        if (r->nx[k] > 0.0f) continue;
        if (r->nx[k] < 0.0f) return 0;
        if (r->ny[k] < 0.0f) continue;
        return 0;
    }
    return 1;
}

/* rasterizer/mod.rs:134-146 */
static void eval_cov(rtri_t *r, uint64_t x, uint64_t y) {
    for (int i = 0; i < r->ns; i++) {
        float xs = (float)x + r->pat[i][0];
        float ys = (float)y + r->pat[i][1];
        eval_single(r, xs, ys, r->cov_eval[i]);
        int v = inside(r, r->cov_eval[i]);
        r->cov_mask = (uint8_t)((r->cov_mask & ~(1u << i)) | ((unsigned)v << i));
    }
}

/* rasterizer/mod.rs:225-253 */
static void fragment_depths(const rtri_t *r, float *sampled) {
    for (int i = 0; i < r->ns; i++) {
        sampled[i] = 0.0f;
        if ((r->cov_mask >> i) & 1) {
            const float *e = r->cov_eval[i];
            float b0 = clamp_bary(e[1] * r->inv_2x_area);
            float b1 = clamp_bary(e[2] * r->inv_2x_area);
            float b2 = clamp_bary(1.0f - b0 - b1);
            sampled[i] = b0 * r->z[0] + b1 * r->z[1] + b2 * r->z[2];
        }
    }
}

/* rasterizer/mod.rs:69-100 */
static attr_t interpolate(const rtri_t *r, uint64_t x, uint64_t y, uint8_t cov) {
    float xs = (float)x + 0.5f, ys = (float)y + 0.5f;
    if (cov != (uint8_t)((1u << r->ns) - 1u)) { /* CoverageMask::all() (mod.rs:42-44) */
        for (int i = 0; i < r->ns; i++)
            if ((cov >> i) & 1) {
                xs = (float)x + r->pat[i][0];
                ys = (float)y + r->pat[i][1];
                break;
            }
    }
    float e[3];
    eval_single(r, xs, ys, e);
    float f_u = e[1] / r->w[0];
    float f_v = e[2] / r->w[1];
    float f_w = e[0] / r->w[2];
    float sum = f_u + f_v + f_w;
    float u = clamp_bary(f_u / sum);
    float v = clamp_bary(f_v / sum);
    float w = clamp_bary(1.0f - u - v);
    return attr_add(attr_add(attr_mul(r->a[0], u), attr_mul(r->a[1], v)), attr_mul(r->a[2], w));
}

/* rasterizer/bounding_box.rs:13-42 (fold with NaN-ignoring min/max, floor/ceil, saturating cast) */
void orc_pixel_bbox(const float *xy6, uint64_t *out4) {
    float mnx = 3.40282347e+38f, mxx = -3.40282347e+38f, mny = 3.40282347e+38f, mxy = -3.40282347e+38f;
    for (int i = 0; i < 3; i++) {
        mnx = fminf(mnx, xy6[2 * i]);
        mxx = fmaxf(mxx, xy6[2 * i]);
        mny = fminf(mny, xy6[2 * i + 1]);
        mxy = fmaxf(mxy, xy6[2 * i + 1]);
    }
    out4[0] = f32_as_usize(floorf(mnx));
    out4[1] = f32_as_usize(ceilf(mxx));
    out4[2] = f32_as_usize(floorf(mny));
    out4[3] = f32_as_usize(ceilf(mxy));
}

/* ---------------- texture ---------------- */

/* texture.rs:47-63 + color.rs:22-29.  Out-of-buffer reads (a panic in the
 * reference) are clamped to the last byte and counted. */
static void read_texel(orc_ctx *c, const tex_t *t, uint64_t x, uint64_t y, float *rgba) {
    if (!(x < t->w) || !(y < t->h)) g_dassert[DA_TEXEL_XY]++; /* debug_assert! texture.rs:49-50 */
    uint64_t start = x * t->tw + y * t->tw * t->w;
    uint8_t b[4] = {0, 0, 0, 255};
    uint32_t n = t->tw == 4 ? 4 : 3;
    for (uint32_t k = 0; k < n; k++) {
        uint64_t o = start + k;
        if (o >= t->len) {
            c->cnt.n_tex_oob++;
            o = t->len - 1;
        }
        b[k] = t->buf[o];
    }
    for (int k = 0; k < 4; k++) rgba[k] = (float)b[k] / 255.0f;
}

/* texture.rs:65-83 */
static void tex_sample(orc_ctx *c, const tex_t *t, float u, float v, float *out) {
    if (!(u >= 0.0f && u <= 1.0f) || !(v >= 0.0f && v <= 1.0f)) g_dassert[DA_TEX_UV]++; /* debug_assert! texture.rs:66-67 */
    float x = u * (float)(t->w - 1);
    float y = v * (float)(t->h - 1);
    uint64_t x0 = f32_as_usize(floorf(x)), x1 = f32_as_usize(ceilf(x));
    uint64_t y0 = f32_as_usize(floorf(y)), y1 = f32_as_usize(ceilf(y));
    float tl[4], tr[4], bl[4], br[4];
    read_texel(c, t, x0, y0, tl);
    read_texel(c, t, x1, y0, tr);
    read_texel(c, t, x0, y1, bl);
    read_texel(c, t, x1, y1, br);
    float xf = x - truncf(x); /* f32::fract */
    float yf = y - truncf(y);
    for (int k = 0; k < 4; k++) {
        float r0 = tl[k] * (1.0f - xf) + tr[k] * xf;
        float r1 = bl[k] * (1.0f - xf) + br[k] * xf;
        out[k] = r0 * (1.0f - yf) + r1 * yf;
    }
}

/* ---------------- context ---------------- */

orc_ctx *orc_create_rows(uint32_t width, uint32_t height, uint32_t row0, uint32_t row1);
orc_ctx *orc_create(uint32_t width, uint32_t height) { return orc_create_rows(width, height, 0, height); }

/* Oracle-only: a context whose per-sample buffers (and returned framebuffer) hold only the pixel rows [row0,row1) of a
 * width x height frame.  Triangles are transformed with the full viewport and their pixel boxes are additionally bounded
 * by the window, exactly like the scissor extension bounds them; pixels are independent of each other (mod.rs:443-473),
 * so the union of the bands of a frame IS the frame.  Per-pixel counters add up over the bands, per-triangle counters
 * are identical in every band. */
orc_ctx *orc_create_rows(uint32_t width, uint32_t height, uint32_t row0, uint32_t row1) {
    orc_ctx *c = (orc_ctx *)calloc(1, sizeof *c);
    if (row1 > height) row1 = height;
    if (row0 > row1) row0 = row1;
    size_t n = (size_t)width * (row1 - row0);
    if (n == 0) n = 1;
    c->width = width;
    c->height = height;
    c->row0 = row0;
    c->row1 = row1;
    c->ns = N_MSAA;
    c->guard = 1.0f;
    c->sc_x0 = 0; c->sc_y0 = 0; c->sc_x1 = width; c->sc_y1 = height;
    c->color = (uint32_t *)malloc(n * N_MSAA * sizeof(uint32_t));
    c->depth = (float *)malloc(n * N_MSAA * sizeof(float));
    c->owner = (uint32_t *)malloc(n * N_MSAA * sizeof(uint32_t));
    c->resolve = (uint32_t *)malloc(n * sizeof(uint32_t));
    for (size_t i = 0; i < n * N_MSAA; i++) {
        c->color[i] = CLEAR_COLOR;
        c->depth[i] = CLEAR_DEPTH;
        c->owner[i] = ORC_NO_OWNER;
    }
    for (size_t i = 0; i < n; i++) c->resolve[i] = CLEAR_COLOR;
    /* buffers.rs:19-47 */
    c->n_horizontal = width / TILE_SIZE + (width % TILE_SIZE == 0 ? 0 : 1);
    c->n_vertical = height / TILE_SIZE + (height % TILE_SIZE == 0 ? 0 : 1);
    size_t nt = (size_t)c->n_horizontal * c->n_vertical;
    c->tile_mask[0] = (uint8_t *)calloc(nt ? nt : 1, 1);
    c->tile_mask[1] = (uint8_t *)calloc(nt ? nt : 1, 1);
    static const float ident[16] = {1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1};
    memcpy(c->world, ident, sizeof ident); /* uniform.rs:18-27 */
    memcpy(c->view, ident, sizeof ident);
    memcpy(c->proj, ident, sizeof ident);
    return c;
}

void orc_destroy(orc_ctx *c) {
    if (!c) return;
    free(c->color); free(c->depth); free(c->owner); free(c->resolve);
    free(c->tile_mask[0]); free(c->tile_mask[1]); free(c->vs_out);
    for (uint32_t i = 0; i < c->n_tex; i++) free(c->tex[i].buf);
    free(c);
}

/* Runtime sample count (extension: N_MSAA_SAMPLES is a constant in the reference, mod.rs:23): 1, 2, 4 or 8 samples per
 * pixel with the patterns above.  Call between frames; the sample buffers are re-created in their cleared state. */
int orc_set_msaa(orc_ctx *c, uint32_t ns) {
    if (ns != 1 && ns != 2 && ns != 4 && ns != 8) return -1;
    size_t n = (size_t)c->width * (c->row1 - c->row0);
    if (n == 0) n = 1;
    free(c->color); free(c->depth); free(c->owner);
    c->ns = (int)ns;
    c->color = (uint32_t *)malloc(n * ns * sizeof(uint32_t));
    c->depth = (float *)malloc(n * ns * sizeof(float));
    c->owner = (uint32_t *)malloc(n * ns * sizeof(uint32_t));
    for (size_t i = 0; i < n * ns; i++) {
        c->color[i] = CLEAR_COLOR;
        c->depth[i] = CLEAR_DEPTH;
        c->owner[i] = ORC_NO_OWNER;
    }
    return 0;
}

/* Guard band (extension sketched at mod.rs:417-419): g >= 1 widens the four side clip planes to |x|, |y| <= g*w, so
 * triangles that leave the viewport but stay inside the band are rasterised unclipped (their pixel boxes are bounded
 * by the viewport, mod.rs:347-361).  g = 1 is the reference. */
int orc_set_guard_band(orc_ctx *c, float g) {
    if (!(g >= 1.0f) || g > 1048576.0f) return -1;
    c->guard = g;
    return 0;
}

/* Scissor rect [x0,x1) x [y0,y1), clamped to the viewport; x0 >= x1 or y0 >= y1 draws nothing. */
int orc_set_scissor(orc_ctx *c, uint32_t x0, uint32_t y0, uint32_t x1, uint32_t y1) {
    c->sc_x0 = x0 < c->width ? x0 : c->width; c->sc_x1 = x1 < c->width ? x1 : c->width;
    c->sc_y0 = y0 < c->height ? y0 : c->height; c->sc_y1 = y1 < c->height ? y1 : c->height;
    return 0;
}

/* uniform.rs:29-33 : index must equal the number of bound textures */
int orc_bind_texture(orc_ctx *c, uint32_t index, const uint8_t *texels, uint32_t w, uint32_t h, uint32_t tw) {
    if (index != c->n_tex || c->n_tex >= 16) return -1;
    if (tw != 3 && tw != 4) return -2;
    tex_t *t = &c->tex[c->n_tex++];
    t->w = w; t->h = h; t->tw = tw;
    t->len = (size_t)w * h * tw;
    t->buf = (uint8_t *)malloc(t->len ? t->len : 1);
    memcpy(t->buf, texels, t->len);
    return 0;
}

void orc_write_block(orc_ctx *c, const float *world, const float *view, const float *proj) {
    if (world) memcpy(c->world, world, 64);
    if (view) memcpy(c->view, view, 64);
    if (proj) memcpy(c->proj, proj, 64);
}

/* main.rs:67-77 : FS ids follow `enum FS` (main.rs:23-27): 0 Texture, 1 Color, 2 Debug.
 * Registry extension (include/rz.h): bits 8..15 of fs_id = texture index for the sampling shaders
 * (Uniforms::get_texture(index), uniform.rs:35-37); 3 = TextureBlend = (sample + attr.color) / 2.0 with the
 * component-wise Color Add and Div<f32> of color.rs:88-111. */
static int fragment_shader(orc_ctx *c, uint32_t fs_id, const float *fc_depths, const attr_t *a, float *rgba) {
    const uint32_t shader = fs_id & 0xFFu, ti = (fs_id >> 8) & 0xFFu;
    switch (shader) {
    case 0:
        if (ti >= c->n_tex) return -1;
        tex_sample(c, &c->tex[ti], a->v[4], a->v[5], rgba);
        return 0;
    case 1:
        rgba[0] = a->v[0]; rgba[1] = a->v[1]; rgba[2] = a->v[2]; rgba[3] = a->v[3];
        return 0;
    case 2: /* Color::grayscale(depths[0]) color.rs:66-73 */
        rgba[0] = rgba[1] = rgba[2] = fc_depths[0];
        rgba[3] = 1.0f;
        return 0;
    case 3: {
        if (ti >= c->n_tex) return -1;
        float t[4];
        tex_sample(c, &c->tex[ti], a->v[4], a->v[5], t);
        for (int k = 0; k < 4; k++) {
            volatile float sum = t[k] + a->v[k]; /* Color + Color, color.rs:99-109 */
            rgba[k] = sum / 2.0f;                /* Color / f32,   color.rs:88-97  */
        }
        return 0;
    }
    default:
        return -1;
    }
}

/* rasterizer/mod.rs:399-476 for one input triangle (already assembled) */
static int rasterize_one(orc_ctx *c, const tri_t *raw, uint32_t fs_id, uint32_t order_base) {
    tri_t clipped[ORC_MAX_POLY];
    const tri_t *list = raw;
    int n = 1, ovf = 0;
    int res = try_clip(raw, clipped, ORC_MAX_POLY, &ovf, c->guard);
    if (ovf) c->cnt.n_clip_overflow++;
    if (res < 0) {
        /* distinguish the degenerate cull for the counters only */
        if (fabsf(area2(raw->p[0][0], raw->p[0][1], raw->p[1][0], raw->p[1][1], raw->p[2][0], raw->p[2][1])) < CULL_EPS)
            c->cnt.n_degenerate++;
        else
            c->cnt.n_outside++;
        return 0;
    }
    if (res == 0) {
        c->cnt.n_inside++;
    } else {
        c->cnt.n_clipped_in++;
        list = clipped;
        n = res;
    }
    const uint32_t W = c->width, H = c->height;
    for (int ti = 0; ti < n; ti++) {
        float ndc[12];
        rtri_t r;
        orc_perspective_divide(&list[ti].p[0][0], ndc);
        viewport_setup(W, H, ndc, list[ti].a, &r);
        r.ns = c->ns;
        r.pat = pattern_of(c->ns);
        const int NS = c->ns;
        c->cnt.n_tris_setup++;
        uint32_t key = order_base * 8u + (uint32_t)(ti < 8 ? ti : 7);
        /* mod.rs:347-361 */
        float xy6[6] = {r.px[0], r.py[0], r.px[1], r.py[1], r.px[2], r.py[2]};
        uint64_t bb[4];
        orc_pixel_bbox(xy6, bb);
        /* "the user would supply a scissoring rect that could be used to bound the triangles" (mod.rs:349-350):
         * the viewport bounds 0..width / 0..height become the scissor rect (default: the viewport itself) */
        uint64_t min_x = bb[0] > c->sc_x0 ? bb[0] : c->sc_x0, max_x = bb[1] < c->sc_x1 ? bb[1] : c->sc_x1;
        uint64_t min_y = bb[2] > c->sc_y0 ? bb[2] : c->sc_y0, max_y = bb[3] < c->sc_y1 ? bb[3] : c->sc_y1;
        if (min_y < c->row0) min_y = c->row0; /* oracle-only row window (orc_create_rows) */
        if (max_y > c->row1) max_y = c->row1;
        for (uint64_t i = min_y; i < max_y; i++) {
            for (uint64_t j = min_x; j < max_x; j++) {
                c->cnt.n_bbox_px++;
                eval_cov(&r, j, i);
                if (!r.cov_mask) continue;
                c->cnt.n_covered_px++;
                float sd[MAX_MSAA];
                fragment_depths(&r, sd);
                size_t idx = (size_t)(i - c->row0) * W + j;
                /* depth_coverage mod.rs:363-378 */
                uint8_t dcov = 0;
                for (int s = 0; s < NS; s++)
                    if (((r.cov_mask >> s) & 1) && sd[s] < c->depth[idx * NS + s]) dcov |= (uint8_t)(1u << s);
                if (!dcov) continue;
                c->cnt.n_shaded_px++;
                attr_t a = interpolate(&r, j, i, dcov);
                float rgba[4];
                if (fragment_shader(c, fs_id, sd, &a, rgba) != 0) return -1;
                /* write_pixel mod.rs:380-397 */
                uint32_t argb = orc_to_argb(rgba);
                c->tile_mask[c->mask_idx][(i / TILE_SIZE) * c->n_horizontal + (j / TILE_SIZE)] = 1;
                for (int s = 0; s < NS; s++)
                    if ((dcov >> s) & 1) {
                        c->color[idx * NS + s] = argb;
                        c->depth[idx * NS + s] = sd[s];
                        c->owner[idx * NS + s] = key;
                        c->cnt.n_samples_written++;
                    }
            }
        }
    }
    return 0;
}

/*
 * Renderer::render render.rs:98-114 : vertex stage over every mesh vertex with the
 * MVP vertex shader (main.rs:147-152, the matrix product recomputed per vertex as
 * the closure does), primitive assembly (render.rs:75-96), rasterize.
 * pos f32[nv][3], attrs f32[nv][6], idx u32[n_idx] (the reference uses usize).
 */
int orc_render(orc_ctx *c, const float *pos, const float *attrs, uint32_t nv, const uint32_t *idx, uint64_t n_idx,
               uint32_t vs_id, uint32_t fs_id) {
    if (vs_id != 0 || (fs_id & 0xFFu) > 3 || (fs_id >> 16) != 0) return -3;
    if (((fs_id & 0xFFu) == 0 || (fs_id & 0xFFu) == 3) && ((fs_id >> 8) & 0xFFu) >= c->n_tex) return -4;
    if (((fs_id & 0xFFu) == 1 || (fs_id & 0xFFu) == 2) && (fs_id >> 8) != 0) return -3;
    if (c->vs_cap < nv) {
        free(c->vs_out);
        c->vs_out = (float *)malloc((size_t)(nv ? nv : 1) * 16);
        c->vs_cap = nv;
    }
    for (uint32_t i = 0; i < nv; i++) {
        float pv[16], pvw[16];
        orc_mat4_mul(c->proj, c->view, pv);
        orc_mat4_mul(pv, c->world, pvw);
        float v[4] = {pos[3 * i], pos[3 * i + 1], pos[3 * i + 2], 1.0f};
        orc_mat4_vec(pvw, v, c->vs_out + 4 * (size_t)i);
    }
    uint64_t n_tris = n_idx / 3;
    for (uint64_t t = 0; t < n_tris; t++) {
        tri_t tri;
        for (int k = 0; k < 3; k++) {
            uint32_t vi = idx[3 * t + k];
            if (vi >= nv) return -5; /* reference panics (slice index) */
            memcpy(tri.p[k], c->vs_out + 4 * (size_t)vi, 16);
            memcpy(tri.a[k].v, attrs + 6 * (size_t)vi, 24);
        }
        c->cnt.n_tris_in++;
        if (rasterize_one(c, &tri, fs_id, c->tri_base + (uint32_t)t) != 0) return -6;
    }
    c->tri_base += (uint32_t)n_tris;
    return 0;
}

/* Rasterizer::rasterize entry for already-assembled clip-space triangles (mod.rs:399-404):
 * clip_pos f32[nt][3][4], attrs f32[nt][3][6] */
int orc_rasterize(orc_ctx *c, const float *clip_pos, const float *attrs, uint64_t nt, uint32_t fs_id) {
    for (uint64_t t = 0; t < nt; t++) {
        tri_t tri;
        memcpy(tri.p, clip_pos + 12 * t, sizeof tri.p);
        memcpy(tri.a, attrs + 18 * t, sizeof tri.a);
        c->cnt.n_tris_in++;
        if (rasterize_one(c, &tri, fs_id, c->tri_base + (uint32_t)t) != 0) return -6;
    }
    c->tri_base += (uint32_t)nt;
    return 0;
}

/* Vertex stage only (render.rs:104-108): out f32[nv][4] */
void orc_vertex_stage(orc_ctx *c, const float *pos, uint32_t nv, float *out) {
    for (uint32_t i = 0; i < nv; i++) {
        float pv[16], pvw[16];
        orc_mat4_mul(c->proj, c->view, pv);
        orc_mat4_mul(pv, c->world, pvw);
        float v[4] = {pos[3 * i], pos[3 * i + 1], pos[3 * i + 2], 1.0f};
        orc_mat4_vec(pvw, v, out + 4 * (size_t)i);
    }
}

/* rasterizer/mod.rs:478-522 (resolve_and_clear through the dirty-tile double buffer) */
const uint32_t *orc_framebuffer(orc_ctx *c) {
    const uint32_t W = c->width, H = c->height;
    size_t nt = (size_t)c->n_horizontal * c->n_vertical;
    uint8_t *prev = c->tile_mask[(c->mask_idx + 1) % 2];
    uint8_t *cur = c->tile_mask[c->mask_idx];
    for (size_t t = 0; t < nt; t++) {
        if (!prev[t]) continue;
        uint32_t tx = (uint32_t)(t % c->n_horizontal), ty = (uint32_t)(t / c->n_horizontal);
        uint32_t x1 = (tx + 1) * TILE_SIZE < W ? (tx + 1) * TILE_SIZE : W;
        uint32_t y1 = (ty + 1) * TILE_SIZE < H ? (ty + 1) * TILE_SIZE : H;
        uint32_t y0 = ty * TILE_SIZE;
        if (y0 < c->row0) y0 = c->row0;
        if (y1 > c->row1) y1 = c->row1;
        for (uint32_t y = y0; y < y1; y++)
            for (uint32_t x = tx * TILE_SIZE; x < x1; x++) c->resolve[(size_t)(y - c->row0) * W + x] = CLEAR_COLOR;
    }
    for (size_t t = 0; t < nt; t++) {
        if (!cur[t]) continue;
        uint32_t tx = (uint32_t)(t % c->n_horizontal), ty = (uint32_t)(t / c->n_horizontal);
        uint32_t x1 = (tx + 1) * TILE_SIZE < W ? (tx + 1) * TILE_SIZE : W;
        uint32_t y1 = (ty + 1) * TILE_SIZE < H ? (ty + 1) * TILE_SIZE : H;
        uint32_t y0 = ty * TILE_SIZE;
        if (y0 < c->row0) y0 = c->row0;
        if (y1 > c->row1) y1 = c->row1;
        for (uint32_t y = y0; y < y1; y++)
            for (uint32_t x = tx * TILE_SIZE; x < x1; x++) {
                size_t idx = (size_t)(y - c->row0) * W + x;
                c->resolve[idx] = box_filter_n(&c->color[idx * c->ns], c->ns);
                for (int s = 0; s < c->ns; s++) {
                    c->color[idx * c->ns + s] = CLEAR_COLOR;
                    c->depth[idx * c->ns + s] = CLEAR_DEPTH;
                    c->owner[idx * c->ns + s] = ORC_NO_OWNER;
                }
            }
    }
    /* BufferTiles::next buffers.rs:58-63 */
    c->mask_idx = (c->mask_idx + 1) % 2;
    memset(c->tile_mask[c->mask_idx], 0, nt);
    c->tri_base = 0;
    return c->resolve;
}

/* Oracle-only views of the per-sample state BEFORE orc_framebuffer() clears it. */
const float *orc_depth_samples(orc_ctx *c) { return c->depth; }
const uint32_t *orc_color_samples(orc_ctx *c) { return c->color; }
const uint32_t *orc_owner_samples(orc_ctx *c) { return c->owner; }

void orc_counters(orc_ctx *c, orc_counters_t *out) { *out = c->cnt; }
void orc_reset_counters(orc_ctx *c) { memset(&c->cnt, 0, sizeof c->cnt); }

/* ---------------- unit-level KAT entry points ---------------- */

/* viewport_transform + RasterizerTriangle::new.  ndc12 in; out: pts[6], normals[6], z[3], w[3], inv */
void orc_viewport_setup(uint32_t width, uint32_t height, const float *ndc12, float *pts6, float *normals6, float *z3,
                        float *w3, float *inv) {
    attr_t a[3];
    memset(a, 0, sizeof a);
    rtri_t r;
    viewport_setup(width, height, ndc12, a, &r);
    for (int i = 0; i < 3; i++) {
        pts6[2 * i] = r.px[i]; pts6[2 * i + 1] = r.py[i];
        normals6[2 * i] = r.nx[i]; normals6[2 * i + 1] = r.ny[i];
        z3[i] = r.z[i]; w3[i] = r.w[i];
    }
    *inv = r.inv_2x_area;
}

/* RasterizerTriangle::new from screen-space vertices (used by the reference's raster tests):
 * screen9 = 3 x (x,y,z); w3; attrs18 */
static void rtri_from_screen(const float *screen9, const float *w3, const float *attrs18, rtri_t *r) {
    float sx[3], sy[3];
    for (int i = 0; i < 3; i++) {
        sx[i] = screen9[3 * i]; sy[i] = screen9[3 * i + 1];
        r->z[i] = screen9[3 * i + 2];
        r->w[i] = w3[i];
        memcpy(r->a[i].v, attrs18 + 6 * i, 24);
        r->px[i] = sx[i]; r->py[i] = sy[i];
    }
    float v0x = sx[1] - sx[0], v0y = sy[1] - sy[0];
    float v1x = sx[2] - sx[1], v1y = sy[2] - sy[1];
    float v2x = sx[0] - sx[2], v2y = sy[0] - sy[2];
    r->nx[0] = -v0y; r->ny[0] = v0x;
    r->nx[1] = -v1y; r->ny[1] = v1x;
    r->nx[2] = -v2y; r->ny[2] = v2x;
    r->inv_2x_area = 1.0f / area2(sx[0], sy[0], sx[1], sy[1], sx[2], sy[2]);
    memset(r->cov_eval, 0, sizeof r->cov_eval);
    r->cov_mask = 0;
    r->ns = N_MSAA;
    r->pat = RGSS;
}

/* eval(x,y) + fragment() + interpolate(x,y,mask_for_interp or own coverage if 0xFF):
 * outputs: mask, edge evals [4][3], sampled depths [4], interpolated attrs [6] */
void orc_eval_pixel(const float *screen9, const float *w3, const float *attrs18, uint64_t x, uint64_t y,
                    uint32_t interp_mask, uint32_t *mask, float *evals12, float *depths4, float *attr6) {
    rtri_t r;
    rtri_from_screen(screen9, w3, attrs18, &r);
    eval_cov(&r, x, y);
    *mask = r.cov_mask;
    memcpy(evals12, r.cov_eval, sizeof(float) * N_MSAA * 3); /* the KAT entry is the reference's 4-sample form */
    fragment_depths(&r, depths4);
    uint8_t m = interp_mask == 0xFF ? r.cov_mask : (uint8_t)interp_mask;
    attr_t a = interpolate(&r, x, y, m);
    memcpy(attr6, a.v, 24);
}

/* eval_single + inside at an arbitrary point: returns inside flag, writes e[3] and normals[6] */
int orc_eval_single(const float *screen9, float x, float y, float *e3, float *normals6) {
    float w3[3] = {1, 1, 1}, attrs[18] = {0};
    rtri_t r;
    rtri_from_screen(screen9, w3, attrs, &r);
    eval_single(&r, x, y, e3);
    for (int i = 0; i < 3; i++) { normals6[2 * i] = r.nx[i]; normals6[2 * i + 1] = r.ny[i]; }
    return inside(&r, e3);
}

/* Texture::sample on an ad-hoc texture; returns number of out-of-buffer byte reads */
uint64_t orc_tex_sample(const uint8_t *texels, uint32_t w, uint32_t h, uint32_t tw, float u, float v, float *rgba) {
    orc_ctx c;
    memset(&c, 0, sizeof c);
    tex_t t;
    t.buf = (uint8_t *)texels; t.w = w; t.h = h; t.tw = tw; t.len = (size_t)w * h * tw;
    tex_sample(&c, &t, u, v, rgba);
    return c.cnt.n_tex_oob;
}

/* BufferTiles geometry (buffers.rs:19-56) for the tile KATs */
void orc_tile_grid(uint32_t width, uint32_t height, uint32_t *n_horizontal, uint32_t *n_vertical) {
    *n_horizontal = width / TILE_SIZE + (width % TILE_SIZE == 0 ? 0 : 1);
    *n_vertical = height / TILE_SIZE + (height % TILE_SIZE == 0 ? 0 : 1);
}
uint32_t orc_tile_idx(uint32_t width, uint32_t row, uint32_t col) {
    uint32_t nh = width / TILE_SIZE + (width % TILE_SIZE == 0 ? 0 : 1);
    return (row / TILE_SIZE) * nh + (col / TILE_SIZE);
}
/* number of dirty tiles in the current / previous mask (BufferTiles::marked / prev_marked) */
uint32_t orc_tiles_marked(orc_ctx *c, int prev) {
    size_t nt = (size_t)c->n_horizontal * c->n_vertical;
    const uint8_t *m = c->tile_mask[prev ? (c->mask_idx + 1) % 2 : c->mask_idx];
    uint32_t n = 0;
    for (size_t t = 0; t < nt; t++) n += m[t];
    return n;
}
