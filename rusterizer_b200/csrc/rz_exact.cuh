// rz_exact.cuh -- bit-exact f32 building blocks shared by every kernel.
//
// The parity contract (BASELINE.json north_star) is "one IEEE-754 binary32
// rounding per source-level operation of the reference, in source order, never fused".  All
// parity-critical arithmetic therefore goes through the explicit round-to-nearest intrinsics
// below (__fmul_rn/__fadd_rn/... are never contracted into FFMA by nvcc), and the translation
// unit is additionally compiled with -fmad=false -prec-div=true -ftz=false.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace rz {

// Programmatic dependent launch (PDL): every kernel of the frame is launched with
// programmaticStreamSerializationAllowed, so its CTAs may become resident while the previous kernel
// is still running; pdl_wait() blocks until that kernel has completed and its writes are visible.
// Each kernel calls pdl_launch() + pdl_wait() before touching global memory, which keeps the stream's
// sequential semantics and only hides the launch latency between the dependent kernels.
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdl_launch() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }

// 256-bit global accesses (sm_100: LDG.E.256 / STG.E.256) of one aligned 32-byte sector, as two float4 halves.
// Plain (coherent) accesses: usable on buffers an earlier kernel of the PDL chain wrote.
__device__ __forceinline__ void ld_sector(const float4 *p, float4 &a, float4 &b) {
    asm volatile("ld.global.v8.f32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                 : "=f"(a.x), "=f"(a.y), "=f"(a.z), "=f"(a.w), "=f"(b.x), "=f"(b.y), "=f"(b.z), "=f"(b.w)
                 : "l"(p));
}
__device__ __forceinline__ void st_sector(float4 *p, const float4 a, const float4 b) {
    asm volatile("st.global.v8.f32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};" ::"l"(p), "f"(a.x), "f"(a.y), "f"(a.z), "f"(a.w), "f"(b.x),
                 "f"(b.y), "f"(b.z), "f"(b.w)
                 : "memory");
}

__device__ __forceinline__ float fadd(float a, float b) { return __fadd_rn(a, b); }
__device__ __forceinline__ float fsub(float a, float b) { return __fsub_rn(a, b); }
__device__ __forceinline__ float fmul(float a, float b) { return __fmul_rn(a, b); }
__device__ __forceinline__ float fdiv(float a, float b) { return __fdiv_rn(a, b); }

// Vector::dot for N=2 (math/vector.rs:17-23): sum starts at 0.0, sequential.
__device__ __forceinline__ float dot2z(float ax, float ay, float bx, float by) {
    return fadd(fadd(0.0f, fmul(ax, bx)), fmul(ay, by));
}
// Vector::dot for N=4
__device__ __forceinline__ float dot4z(float a0, float a1, float a2, float a3, float b0, float b1, float b2, float b3) {
    float s = fadd(0.0f, fmul(a0, b0));
    s = fadd(s, fmul(a1, b1));
    s = fadd(s, fmul(a2, b2));
    s = fadd(s, fmul(a3, b3));
    return s;
}
// Vector<_,2>::cross (math/vector.rs:183-185)
__device__ __forceinline__ float cross2(float ax, float ay, float bx, float by) {
    return fsub(fmul(ax, by), fmul(bx, ay));
}
// Rust f32::clamp(0.0, 1.0) as used by clamp_bary (rasterizer/mod.rs:102-106): NaN stays NaN.
__device__ __forceinline__ float clamp01(float x) {
    if (x < 0.0f) x = 0.0f;
    if (x > 1.0f) x = 1.0f;
    return x;
}
// Rust `f32 as u32` / `as usize` (saturating, NaN -> 0).  cvt.rzi.u32.f32 has these semantics.
__device__ __forceinline__ uint32_t sat_u32(float f) { return __float2uint_rz(f); }

// RGSS sample offsets (rasterizer/mod.rs:109-114); all exactly representable.
__device__ __forceinline__ float rgss_x(int i) { return i == 0 ? 0.625f : (i == 1 ? 0.875f : (i == 2 ? 0.375f : 0.125f)); }
__device__ __forceinline__ float rgss_y(int i) { return i == 0 ? 0.125f : (i == 1 ? 0.625f : (i == 2 ? 0.875f : 0.375f)); }

// Screen-space triangle: RasterizerTriangle (rasterizer/mod.rs:178-222) without the attributes.
struct Setup {
    float px[3], py[3]; // EdgeFunctions.points
    float nx[3], ny[3]; // EdgeFunctions.normals
    float z[3];         // depths
    float w[3];         // depths_camera_space
    float inv;          // inv_2x_area
};

// RasterizerTriangle::new (rasterizer/mod.rs:187-222) from screen points.
__device__ __forceinline__ void setup_edges(Setup &s) {
    float v0x = fsub(s.px[1], s.px[0]), v0y = fsub(s.py[1], s.py[0]);
    float v1x = fsub(s.px[2], s.px[1]), v1y = fsub(s.py[2], s.py[1]);
    float v2x = fsub(s.px[0], s.px[2]), v2y = fsub(s.py[0], s.py[2]);
    s.nx[0] = -v0y; s.ny[0] = v0x;
    s.nx[1] = -v1y; s.ny[1] = v1x;
    s.nx[2] = -v2y; s.ny[2] = v2x;
    // triangle_2x_area (rasterizer/mod.rs:15-21): cross(p1-p0, p2-p0)
    float v20x = fsub(s.px[2], s.px[0]), v20y = fsub(s.py[2], s.py[0]);
    s.inv = fdiv(1.0f, cross2(v0x, v0y, v20x, v20y));
}

// Edge normals only (for shading, which needs neither inv_2x_area nor the depths).
__device__ __forceinline__ void setup_normals(Setup &s) {
    float v0x = fsub(s.px[1], s.px[0]), v0y = fsub(s.py[1], s.py[0]);
    float v1x = fsub(s.px[2], s.px[1]), v1y = fsub(s.py[2], s.py[1]);
    float v2x = fsub(s.px[0], s.px[2]), v2y = fsub(s.py[0], s.py[2]);
    s.nx[0] = -v0y; s.ny[0] = v0x;
    s.nx[1] = -v1y; s.ny[1] = v1x;
    s.nx[2] = -v2y; s.ny[2] = v2x;
}

// EdgeFunctions::eval_single for one edge (rasterizer/mod.rs:125-132)
__device__ __forceinline__ float edge_eval(const Setup &s, int k, float x, float y) {
    return dot2z(s.nx[k], s.ny[k], fsub(x, s.px[k]), fsub(y, s.py[k]));
}

// One edge of EdgeFunctions::inside (rasterizer/mod.rs:148-170), tie-break included.
__device__ __forceinline__ bool edge_pass(float e, float nx, float ny) {
    if (e > 0.0f) return true;
    if (e < 0.0f) return false;
    if (nx > 0.0f) return true;
    if (nx < 0.0f) return false;
    return ny < 0.0f;
}

// EdgeFunctions::eval (rasterizer/mod.rs:134-146): 4-sample coverage mask of pixel (X,Y).
__device__ __forceinline__ uint32_t coverage_mask(const Setup &s, int X, int Y) {
    uint32_t mask = 0;
    float fx = (float)X, fy = (float)Y;
#pragma unroll
    for (int i = 0; i < 4; i++) {
        float xs = fadd(fx, rgss_x(i)), ys = fadd(fy, rgss_y(i));
        bool in = edge_pass(edge_eval(s, 0, xs, ys), s.nx[0], s.ny[0]) &&
                  edge_pass(edge_eval(s, 1, xs, ys), s.nx[1], s.ny[1]) &&
                  edge_pass(edge_eval(s, 2, xs, ys), s.nx[2], s.ny[2]);
        mask |= (in ? 1u : 0u) << i;
    }
    return mask;
}

// Fast exact coverage for FINITE screen coordinates (the caller guarantees |coords| < 2^20, so every
// edge value is finite).  Two observations make one compare per edge enough:
//   * EdgeFunctions::inside (mod.rs:148-170) reduces to  e > 0 || (e == 0 && tie_k)  with the
//     per-edge constant tie_k = n.x > 0 || (n.x == 0 && n.y < 0); that is  e >= thr_k  with
//     thr_k = 0.0 when tie_k holds and the smallest positive subnormal otherwise (-0.0 >= 0.0 holds);
//   * the leading `0.0 +` of Vector::dot only turns a -0.0 into +0.0, which `>=` cannot see, so it is
//     dropped HERE (the depth / attribute paths keep it).
// thr[k] comes from edge_thresholds().
__device__ __forceinline__ void edge_thresholds(const Setup &s, float *thr) {
#pragma unroll
    for (int k = 0; k < 3; k++)
        thr[k] = (s.nx[k] > 0.0f || (!(s.nx[k] < 0.0f) && s.ny[k] < 0.0f)) ? 0.0f : 1.401298464e-45f;
}
__device__ __forceinline__ uint32_t coverage_mask_fast(const Setup &s, const float *thr, int X, int Y) {
    uint32_t mask = 0;
    const float fx = (float)X, fy = (float)Y;
#pragma unroll
    for (int i = 0; i < 4; i++) {
        const float xs = fadd(fx, rgss_x(i)), ys = fadd(fy, rgss_y(i));
        bool in = true;
#pragma unroll
        for (int k = 0; k < 3; k++) {
            const float e = fadd(fmul(s.nx[k], fsub(xs, s.px[k])), fmul(s.ny[k], fsub(ys, s.py[k])));
            in = in & (e >= thr[k]);
        }
        mask |= (in ? 1u : 0u) << i;
    }
    return mask;
}
__device__ __forceinline__ bool setup_is_tame(const Setup &s) {
    bool ok = true;
#pragma unroll
    for (int k = 0; k < 3; k++) ok = ok & (fabsf(s.px[k]) < 1048576.0f) & (fabsf(s.py[k]) < 1048576.0f);
    return ok; // false for NaN / inf / absurd coordinates: those take the literal path
}

// interpolate_depth closure of RasterizerTriangle::fragment (rasterizer/mod.rs:226-236)
__device__ __forceinline__ float sample_depth(const Setup &s, int X, int Y, int i) {
    float xs = fadd((float)X, rgss_x(i)), ys = fadd((float)Y, rgss_y(i));
    float e1 = edge_eval(s, 1, xs, ys), e2 = edge_eval(s, 2, xs, ys);
    float b0 = clamp01(fmul(e1, s.inv));
    float b1 = clamp01(fmul(e2, s.inv));
    float b2 = clamp01(fsub(fsub(1.0f, b0), b1));
    return fadd(fadd(fmul(b0, s.z[0]), fmul(b1, s.z[1])), fmul(b2, s.z[2]));
}

// Coverage and depth test of a pixel's four samples against one triangle, for the pixel-parallel walks (finite
// coordinates).  d holds the four depths of the pixel (strict <, rasterizer/mod.rs:374); returns the coverage mask, mp
// receives the post-depth-test mask.  `full`: the binner proved every sample of the item's box inside (ENTRY_FULL), the
// coverage test is skipped.  Two separate code paths on purpose: in the general one the compiler shares the edge values
// between coverage_mask_fast and sample_depth, and a merged form loses that.
__device__ __forceinline__ uint32_t cover_depth4(const Setup &s, const float *thr, bool full, int X, int Y, float4 &d, uint32_t &mp) {
    mp = 0u;
    if (full) {
        const float z0 = sample_depth(s, X, Y, 0), z1 = sample_depth(s, X, Y, 1), z2 = sample_depth(s, X, Y, 2), z3 = sample_depth(s, X, Y, 3);
        if (z0 < d.x) { d.x = z0; mp |= 1u; }
        if (z1 < d.y) { d.y = z1; mp |= 2u; }
        if (z2 < d.z) { d.z = z2; mp |= 4u; }
        if (z3 < d.w) { d.w = z3; mp |= 8u; }
        return 0xFu;
    }
    const uint32_t m = coverage_mask_fast(s, thr, X, Y);
    if (m & 1u) { const float z = sample_depth(s, X, Y, 0); if (z < d.x) { d.x = z; mp |= 1u; } }
    if (m & 2u) { const float z = sample_depth(s, X, Y, 1); if (z < d.y) { d.y = z; mp |= 2u; } }
    if (m & 4u) { const float z = sample_depth(s, X, Y, 2); if (z < d.z) { d.z = z; mp |= 4u; } }
    if (m & 8u) { const float z = sample_depth(s, X, Y, 3); if (z < d.w) { d.w = z; mp |= 8u; } }
    return m;
}

// PixelBoundingBox::from + Rasterizer::bounding_box (bounding_box.rs:13-42, mod.rs:347-361),
// clamped to the viewport.  Half-open pixel ranges; empty when min >= max.
struct BBox {
    uint32_t x0, x1, y0, y1;
};
// `sc` = {x0, y0, x1, y1}: the viewport, or the scissor rect the reference suggests at mod.rs:349-350.
__device__ __forceinline__ BBox pixel_bbox(const Setup &s, const uint4 sc) {
    float mnx = fminf(fminf(fminf(3.40282347e+38f, s.px[0]), s.px[1]), s.px[2]);
    float mxx = fmaxf(fmaxf(fmaxf(-3.40282347e+38f, s.px[0]), s.px[1]), s.px[2]);
    float mny = fminf(fminf(fminf(3.40282347e+38f, s.py[0]), s.py[1]), s.py[2]);
    float mxy = fmaxf(fmaxf(fmaxf(-3.40282347e+38f, s.py[0]), s.py[1]), s.py[2]);
    BBox b;
    b.x0 = max(sat_u32(floorf(mnx)), sc.x);
    b.x1 = min(sat_u32(ceilf(mxx)), sc.z);
    b.y0 = max(sat_u32(floorf(mny)), sc.y);
    b.y1 = min(sat_u32(ceilf(mxy)), sc.w);
    return b;
}

// Color::to_argb (color.rs:15-20): truncating saturating casts, fields OR-ed without masking.
__device__ __forceinline__ uint32_t to_argb(float r, float g, float b, float a) {
    return (sat_u32(fmul(a, 255.0f)) << 24) | (sat_u32(fmul(r, 255.0f)) << 16) | (sat_u32(fmul(g, 255.0f)) << 8) |
           sat_u32(fmul(b, 255.0f));
}

// ColorBuffer::box_filter_color (rasterizer/buffers.rs:111-125)
__device__ __forceinline__ uint32_t box_filter(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3) {
    uint32_t r = ((c0 >> 16) & 0xFF) + ((c1 >> 16) & 0xFF) + ((c2 >> 16) & 0xFF) + ((c3 >> 16) & 0xFF);
    uint32_t g = ((c0 >> 8) & 0xFF) + ((c1 >> 8) & 0xFF) + ((c2 >> 8) & 0xFF) + ((c3 >> 8) & 0xFF);
    uint32_t b = (c0 & 0xFF) + (c1 & 0xFF) + (c2 & 0xFF) + (c3 & 0xFF);
    return 0xFF000000u | ((r >> 2) << 16) | ((g >> 2) << 8) | (b >> 2);
}

} // namespace rz
