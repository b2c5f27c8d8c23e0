// rz_tile.cuh -- stage 3+4: per-tile raster, MSAA depth test, shading, resolve.
//
// One CTA of 256 threads owns one 16x16 pixel tile.  The tile's 4-sample depth and colour state
// (the reference's DepthBuffer / ColorBuffer, rasterizer/buffers.rs:83-157) lives in shared
// memory from clear to resolve and only the box-filtered u32 image goes back to HBM.
//
// Submission order is honoured exactly, but without serialising on it.
// Within a chunk of <= 255 items (and <= 8192 bbox pixels) every covered pixel of every item
// becomes a FRAGMENT record {item, pixel, coverage, 4 sample depths} on a per-pixel list, and
// each fragment F decides by itself, from the other fragments G of its pixel, what the ordered
// replay of the reference would have done:
//     post-depth mask  m(F)[s] = cov_F[s] & z_F[s] < z_old[s] & no G earlier than F with z_G[s] <= z_F[s]
//     still visible    v(F)[s] = m(F)[s]  & no G later   than F with z_G[s] <  z_F[s]
// ("earlier/later" by order key; strict-< depth test, rasterizer/mod.rs:374).  m(F) selects the
// shading position (rasterizer/mod.rs:70-83) and feeds the counters; only fragments with v(F) != 0
// are shaded, and they write colour + depth of exactly the samples in v(F).  Nothing in a chunk
// depends on the order in which threads run, so a tile whose list fits one chunk is never sorted.
//   phase A0 thread = item          bin entry (order key, record index, in-tile pixel box) -> scan of the box areas
//   phase A1 thread = (item, pixel pair) exact coverage of two horizontally adjacent box pixels (shared y terms);
//                                   covered pixels become fragment records on per-pixel lists
//   phase A2 thread = fragment      the covered samples' depths (RasterizerTriangle::fragment, mod.rs:225-253)
//   phase B  thread = pixel         m(F), v(F): the pixel's few fragments replayed in key order (lists of up to
//                                   four fragments are ordered and replayed in registers)
//   phase C  thread = fragment      interpolate + fragment shader + pack; write the visible samples
// Tiles that need several chunks sort their list by order key first.  Items with non-finite or absurd
// coordinates take a literal pixel-parallel walk (a thread owns a pixel for the whole run, so it
// applies the triangles in order by construction) that evaluates EdgeFunctions::inside verbatim; the geometry
// stage flags frames that contain such items (FrameState::has_wild), all others skip the search for them.
// A single-chunk tile passes six CTA barriers: one at the top of the trip, two in A0, one after A1, B, C.
// Setup (inv_2x_area, pixel box, tie-break bits) is done once per triangle by the geometry stage; the units of phase A1
// and the fragments of phase C read the triangle's record straight from global memory (L1: ~7 units share a record).
#pragma once
#include "rz_exact.cuh"
#include "rz_geom.cuh"
#include "rz_types.cuh"

namespace rz {

// Texture::read_texel + Color::from_rgba (texture.rs:47-63, color.rs:22-29).  Reads past the
// buffer (a panic in the reference) are clamped to the last byte and counted.
// `lut[b]` holds (b as f32) / 255.0 computed with the same IEEE division, once per CTA.
__device__ __forceinline__ void read_texel(const TexInfo &t, const float *lut, uint32_t x, uint32_t y, float *rgba,
                                           uint32_t &oob) {
    const unsigned long long start =
        (unsigned long long)x * t.tw + (unsigned long long)y * t.tw * (unsigned long long)t.w;
    uint32_t b0, b1, b2, b3 = 255u;
    if (t.tw == 4 && start + 3 < t.len) {
        const uint32_t v = __ldg(reinterpret_cast<const uint32_t *>(t.data + start));
        b0 = v & 0xFF; b1 = (v >> 8) & 0xFF; b2 = (v >> 16) & 0xFF; b3 = v >> 24;
    } else {
        uint32_t bb[4] = {0, 0, 0, 255};
        const uint32_t nch = t.tw == 4 ? 4u : 3u;
        for (uint32_t k = 0; k < nch; k++) {
            unsigned long long o = start + k;
            if (o >= t.len) {
                oob++;
                o = t.len - 1;
            }
            bb[k] = __ldg(t.data + o);
        }
        b0 = bb[0]; b1 = bb[1]; b2 = bb[2]; b3 = bb[3];
    }
    rgba[0] = lut[b0];
    rgba[1] = lut[b1];
    rgba[2] = lut[b2];
    rgba[3] = lut[b3];
}

// Texture::sample (texture.rs:65-83) + Color::to_argb.  ALPHA = false skips the alpha channel: it only
// lands in bits 24..31 of the packed colour (channels bleed upwards, never downwards, color.rs:15-20)
// and the resolve forces those bits to 0xFF (buffers.rs:121-124), so the image is identical; the
// parity instrumentation keeps it to compare the per-sample colours bit for bit.
template <bool ALPHA>
__device__ __forceinline__ void sample_texture(const TexInfo &t, const float *lut, float u, float v, uint32_t &oob, float *o) {
    const float x = fmul(u, (float)(t.w - 1)), y = fmul(v, (float)(t.h - 1));
    const uint32_t x0 = sat_u32(floorf(x)), x1 = sat_u32(ceilf(x));
    const uint32_t y0 = sat_u32(floorf(y)), y1 = sat_u32(ceilf(y));
    const float xf = fsub(x, truncf(x)), yf = fsub(y, truncf(y)); // f32::fract
    const float omx = fsub(1.0f, xf), omy = fsub(1.0f, yf);
    float tl[4], tr[4], bl[4], br[4];
    uint32_t v00 = 0, v10 = 0, v01 = 0, v11 = 0;
    bool packed = false;
    if ((t.bound & 2u) && x1 < t.w && y1 < t.h) {
        // RGBA8 texture below 4 GB with all four texels inside it (floor <= ceil, so x1 and y1 decide; always the
        // case for u, v in [0, 1]): the byte offsets x * 4 + y * 4 * width of read_texel fit 32 bits, word index
        const uint32_t *tex = reinterpret_cast<const uint32_t *>(t.data);
        const uint32_t r0 = y0 * t.w, r1 = y1 * t.w;
        v00 = __ldg(tex + r0 + x0); v10 = __ldg(tex + r0 + x1);
        v01 = __ldg(tex + r1 + x0); v11 = __ldg(tex + r1 + x1);
        packed = true;
    } else {
        const unsigned long long rs = (unsigned long long)t.tw * t.w; // bytes per texel row
        const unsigned long long r0 = (unsigned long long)y0 * rs, r1 = (unsigned long long)y1 * rs;
        const unsigned long long c0 = (unsigned long long)x0 * t.tw, c1 = (unsigned long long)x1 * t.tw;
        if (t.tw == 4 && r1 + c1 + 3 < t.len && r1 + c0 + 3 < t.len && r0 + c1 + 3 < t.len) {
            // RGBA8, all four texels inside the buffer -> four 32-bit loads
            v00 = __ldg(reinterpret_cast<const uint32_t *>(t.data + r0 + c0));
            v10 = __ldg(reinterpret_cast<const uint32_t *>(t.data + r0 + c1));
            v01 = __ldg(reinterpret_cast<const uint32_t *>(t.data + r1 + c0));
            v11 = __ldg(reinterpret_cast<const uint32_t *>(t.data + r1 + c1));
            packed = true;
        }
    }
    if (packed) {
#pragma unroll
        for (int k = 0; k < (ALPHA ? 4 : 3); k++) {
            tl[k] = lut[(v00 >> (8 * k)) & 0xFFu]; tr[k] = lut[(v10 >> (8 * k)) & 0xFFu];
            bl[k] = lut[(v01 >> (8 * k)) & 0xFFu]; br[k] = lut[(v11 >> (8 * k)) & 0xFFu];
        }
    } else {
        read_texel(t, lut, x0, y0, tl, oob);
        read_texel(t, lut, x1, y0, tr, oob);
        read_texel(t, lut, x0, y1, bl, oob);
        read_texel(t, lut, x1, y1, br, oob);
    }
#pragma unroll
    for (int k = 0; k < (ALPHA ? 4 : 3); k++) {
        const float q0 = fadd(fmul(tl[k], omx), fmul(tr[k], xf));
        const float q1 = fadd(fmul(bl[k], omx), fmul(br[k], xf));
        o[k] = fadd(fmul(q0, omy), fmul(q1, yf));
    }
    if (!ALPHA) o[3] = 0.0f;
}
template <bool ALPHA>
__device__ __forceinline__ uint32_t pack_argb(const float *o) {
    if (!ALPHA) return 0xFF000000u | to_argb(o[0], o[1], o[2], 0.0f);
    return to_argb(o[0], o[1], o[2], o[3]);
}

// Fragment::interpolate (rasterizer/mod.rs:69-100) + the built-in fragment shaders
// (main.rs:67-77) + Color::to_argb.  mpost is the POST-depth-test mask, depth0 the pre-test
// sampled depth of sample 0 (0.0 when uncovered), as FragCoords.depths[0] (mod.rs:458-463).
// `s` needs the screen points and edge normals only; depths_camera_space, the shader id and the
// attribute locations come from the triangle's ShadeRec.
// EXT = false compiles the reference's three shaders on texture 0 only (the registry extension costs registers in
// the hottest loop of the frame; frames that do not use it run the lean instantiation).
// shade_at: the same at an explicit sample position (xs, ys) -- the caller applies the position rule of
// Fragment::interpolate (mod.rs:70-83) for its sample pattern.
// `depth0` is a callable: only the depth-visualising shader reads FragCoords.depths[0], and the pixel-parallel walks
// would have to evaluate it (one more sample depth per fragment) just to throw it away.
template <bool ALPHA, bool EXT, typename D0>
__device__ __forceinline__ uint32_t shade_at(const FrameParams &P, const Setup &s, uint32_t rec, const float *lut, float xs,
                                             float ys, D0 depth0, uint32_t &oob) {
    // (records are written by the geometry kernels of this frame: plain loads, not the read-only path)
    float4 shf, wq; // the whole 32-byte record in one 256-bit load; wq = depths_camera_space
    ld_sector(reinterpret_cast<const float4 *>(&P.shade[rec]), shf, wq);
    const uint4 sh = make_uint4(__float_as_uint(shf.x), __float_as_uint(shf.y), __float_as_uint(shf.z), __float_as_uint(shf.w));
    const uint32_t info = sh.x, fs = info & 3u, texidx = EXT ? (info >> 3) & 31u : 0u;
    if (fs == 2u) { // Color::grayscale(depths[0])
        const float g = depth0();
        return to_argb(g, g, g, 1.0f);
    }
    const bool clipped = (info & 4u) != 0u;
    const float *a0, *a1, *a2;
    if (clipped) { // interpolated attributes live in an AttrRec (written by clip_kernel in this frame)
        a0 = P.attrs[sh.y].a;
        a1 = a0 + 6;
        a2 = a0 + 12;
    } else {       // unclipped: straight from the mesh
        const float *attr = P.draws[info >> 8].attr;
        a0 = attr + 6 * (size_t)sh.y;
        a1 = attr + 6 * (size_t)sh.z;
        a2 = attr + 6 * (size_t)sh.w;
    }
    const float e0 = edge_eval(s, 0, xs, ys), e1 = edge_eval(s, 1, xs, ys), e2 = edge_eval(s, 2, xs, ys);
    const float fu = fdiv(e1, wq.x), fv = fdiv(e2, wq.y), fw = fdiv(e0, wq.z);
    const float sum = fadd(fadd(fu, fv), fw);
    const float u = clamp01(fdiv(fu, sum));
    const float v = clamp01(fdiv(fv, sum));
    const float w = clamp01(fsub(fsub(1.0f, u), v));
#define RZ_LDA(p) (*(p)) // plain loads: the pointers may address AttrRecs written by clip_kernel in this frame
#define RZ_INTERP(c) fadd(fadd(fmul(RZ_LDA(a0 + (c)), u), fmul(RZ_LDA(a1 + (c)), v)), fmul(RZ_LDA(a2 + (c)), w))
    if (fs == 1u) return to_argb(RZ_INTERP(0), RZ_INTERP(1), RZ_INTERP(2), RZ_INTERP(3));
    // (u, v) of a vertex sit at byte offset 24 * i + 16 of an array that starts on a 256-byte boundary (cudaMalloc;
    // AttrRec is 16-byte aligned with the three attributes at 0, 24, 48): one 64-bit load per vertex
    const float2 *u0p = reinterpret_cast<const float2 *>(a0 + 4), *u1p = reinterpret_cast<const float2 *>(a1 + 4),
                 *u2p = reinterpret_cast<const float2 *>(a2 + 4);
    const float2 t0 = RZ_LDA(u0p), t1 = RZ_LDA(u1p), t2 = RZ_LDA(u2p);
    const float tu = fadd(fadd(fmul(t0.x, u), fmul(t1.x, v)), fmul(t2.x, w));
    const float tv = fadd(fadd(fmul(t0.y, u), fmul(t1.y, v)), fmul(t2.y, w));
    float o[4];
    if (!EXT || texidx == 0u) {
        sample_texture<ALPHA>(P.tex0, lut, tu, tv, oob, o);
    } else {
        const TexInfo t = P.tex_table[texidx];
        sample_texture<ALPHA>(t, lut, tu, tv, oob, o);
    }
    if (EXT && fs == 3u) { // FS TextureBlend: (texture.sample(u, v) + attr.color) / 2.0   (Color Add + Div<f32>, color.rs:88-111)
#pragma unroll
        for (int k = 0; k < (ALPHA ? 4 : 3); k++) o[k] = fdiv(fadd(o[k], RZ_INTERP(k)), 2.0f);
    }
#undef RZ_INTERP
#undef RZ_LDA
    return pack_argb<ALPHA>(o);
}

// 4-sample form: the position rule of Fragment::interpolate (mod.rs:70-83) for the rotated-grid pattern -- the pixel
// centre when all four samples passed the depth test, else the first passing sample.
template <bool ALPHA, bool EXT, typename D0>
__device__ __forceinline__ uint32_t shade(const FrameParams &P, const Setup &s, uint32_t rec, const float *lut, int X,
                                          int Y, uint32_t mpost, D0 depth0, uint32_t &oob) {
    float xs, ys;
    if (mpost == 0xFu) {
        xs = fadd((float)X, 0.5f);
        ys = fadd((float)Y, 0.5f);
    } else {
        const int i = __ffs(mpost) - 1;
        xs = fadd((float)X, rgss_x(i));
        ys = fadd((float)Y, rgss_y(i));
    }
    return shade_at<ALPHA, EXT>(P, s, rec, lut, xs, ys, depth0, oob);
}

// Bitonic network in its "flip then halve" form: every compare-exchange puts the smaller element at
// the lower index, so n need not be a power of two (the missing tail behaves like +inf padding).
template <typename T, typename Less>
__device__ __forceinline__ void block_sort(T *a, int n, Less less) {
    // Thread t owns the elements t, t + NT, ...: an aligned block of 32 elements belongs to one warp, so an
    // exchange distance below 32 keeps both partners inside the warp and __syncwarp orders the stages; only
    // the wide stages (and the first narrow one after them) need the CTA barrier -- 14 instead of 45 for 512 keys.
    bool prev_wide = false;
    for (int k = 2; (k >> 1) < n; k <<= 1) {
        {
            const bool wide = k > 32;
            if (wide || prev_wide) __syncthreads(); else __syncwarp();
            prev_wide = wide;
        }
        for (int i = threadIdx.x; i < n; i += NT) {
            const int p = i ^ (k - 1);
            if (p > i && p < n) {
                T x = a[i], y = a[p];
                if (less(y, x)) { a[i] = y; a[p] = x; }
            }
        }
        for (int j = k >> 2; j > 0; j >>= 1) {
            const bool wide = j >= 32;
            if (wide || prev_wide) __syncthreads(); else __syncwarp();
            prev_wide = wide;
            for (int i = threadIdx.x; i < n; i += NT) {
                const int p = i ^ j;
                if (p > i && p < n) {
                    T x = a[i], y = a[p];
                    if (less(y, x)) { a[i] = y; a[p] = x; }
                }
            }
        }
    }
    __syncthreads();
}

// A large item staged in shared memory for the pixel-parallel walks (20 words).
struct __align__(16) BigSetup {
    float px[3], py[3], nx[3], ny[3], z[3];
    float inv;
    uint32_t key, rec;
    uint32_t box; // lx0 | ly0 << 8 | bw << 16 | bh << 24   (tile-local)
    uint32_t tie; // tie-break bits of the three edges | block mask << 8 (bin entry, rz_types.cuh)
};
static_assert(sizeof(BigSetup) == 80, "BigSetup must be 20 words");

constexpr uint32_t FR_NONE = 0xFFFFu;

// Pixel of a thread in the pixel-parallel phases: warp w owns the 8x4 pixel block w of the tile (block row w / 2, block
// column w % 2), lane l the pixel (l % 8, l / 8) inside it.  Eight consecutive lanes are eight consecutive pixels of a
// row: float4 sample records stay conflict-free and the epilogue's 128-bit row stores still collect four neighbours.
__device__ __forceinline__ void pixel_of_thread(int &lx, int &ly) {
    const int w = threadIdx.x >> 5, l = threadIdx.x & 31;
    lx = (w & 1) * BLOCK_W + (l & 7);
    ly = (w >> 1) * BLOCK_H + (l >> 3);
}

// Asynchronous global -> shared copies (LDGSTS): no register staging, the issuing thread does not wait.
__device__ __forceinline__ void cp_async16(void *smem_dst, const void *gsrc) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"((uint32_t)__cvta_generic_to_shared(smem_dst)), "l"(gsrc) : "memory");
}
__device__ __forceinline__ void cp_async8(void *smem_dst, const void *gsrc) {
    asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"((uint32_t)__cvta_generic_to_shared(smem_dst)), "l"(gsrc) : "memory");
}
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_all;" ::: "memory"); }

#ifndef RZ_B_REG_MAXFRAG
#define RZ_B_REG_MAXFRAG 640 // chunks with more fragments (> 2.5 per pixel: overdraw) replay their pixel lists by re-walking
#endif
#ifndef RZ_BUDGET_TRIG
#define RZ_BUDGET_TRIG 3u // a chunk is cut beforehand when its predicted fragments exceed 7/8 * (1 + TRIG/8) of the pool
#endif
__device__ __forceinline__ uint32_t fit_units(uint32_t units, uint32_t nfrag) {
    return nfrag > 32u ? max(256u, units * (uint32_t)(POOL * 7 / 8) / nfrag) : (uint32_t)UNIT_CAP;
}

struct FragPool {
    float4 z[POOL];        // the 4 sample depths (0.0 where uncovered, like Fragment.sampled_depths)
    uint32_t meta[POOL];   // item | pixel << 8 | coverage << 16
    uint16_t next[POOL];   // per-pixel list link
    uint8_t fin[POOL];     // post-depth mask | still-visible mask << 4
};

template <bool DBG>
struct TileSmemT {
    float depth[TILE_PX * 4];
    uint32_t color[TILE_PX * 4];
    uint32_t okey[DBG ? TILE_PX * 4 : 4]; // owner keys (parity instrumentation only)
    float lut[256];                       // (b as f32) / 255.0   (Color::from_rgba, color.rs:22-29)
    uint32_t it_key[CHUNK + 1];           // items of the current chunk: order key (the rank in the list once it is sorted)
    uint32_t it_rec[CHUNK + 1];                // record index | tie-break bits << 29
    uint32_t it_okey[DBG ? CHUNK + 1 : 1];       // the triangle's order key as the oracle reports it (parity instrumentation only)
    uint32_t it_box[CHUNK + 1];                // lx0 | ly0 << 4 | bw << 8 | ceil(65536 / bw) << 13   (j / bw == (j * rcp) >> 16 for j < 256)
    float4 it_q0[CHUNK + 1], it_q1[CHUNK + 1];          // the items' records, quarters 0 and 1 (screen points, z0, z1) and the first half of
    float2 it_q2[CHUNK + 1];                   //   quarter 2 (z2, inv_2x_area): copied asynchronously (cp.async) while phase A0 scans
    uint32_t pre[NT + 1];                 // exclusive prefix of the in-tile box areas (work units)
    uint8_t unit_item[UNIT_CAP];          // work unit -> item
    union {
        FragPool fr;
        unsigned long long sorted[SORT_CAP];
        BigSetup big[CHUNK];
    } u;
    uint32_t head[TILE_PX];               // per-pixel fragment list heads
    uint32_t scan[NT / 32];
    uint32_t first_big, first_small, nfrag, ovf;
    uint4 ent;                            // the tile of this trip: {tile id, list length, first bin entry, work index}; x = ~0: none left
    uint32_t bucket_end[ORDER_BUCKETS];   // prefix of the list-length class sizes (busy-list work order)
    uint32_t clr_cursor[NT / 32];         // per-warp cursor of the empty-tile clears
    uint32_t unit_budget;                 // adaptive work-unit budget of a chunk (fragment pool occupancy predictor)
};

// Sort the tile's list by order key.  Lists of up to SORT_CAP entries: (key, position) pairs are sorted in shared
// memory, then the entries are gathered in key order and written back over the head of the bin in COMPACT form --
// uint2 {record | tie bits, box}; the order key of entry i is now simply i (returns true).  Longer lists are sorted in
// place in HBM as full entries (returns false).
__device__ __forceinline__ bool sort_tile_list_ptr(unsigned long long *sorted, uint4 *bin, int n) {
    __syncthreads();
    if (n <= SORT_CAP) {
        for (int i = threadIdx.x; i < n; i += NT)
            sorted[i] = ((unsigned long long)__ldcg(reinterpret_cast<const uint32_t *>(bin + i)) << 32) | (uint32_t)i;
        __syncthreads();
        block_sort(sorted, n, [](unsigned long long a, unsigned long long b) { return a < b; });
        uint2 e[SORT_CAP / NT];
#pragma unroll
        for (int k = 0; k < SORT_CAP / NT; k++) {
            const int i = threadIdx.x + k * NT;
            if (i < n) {
                const uint4 v = __ldcg(bin + (uint32_t)sorted[i]);
                e[k] = make_uint2(v.y, v.z);
            }
        }
        __syncthreads(); // every entry has been read: the compact list may overwrite the head of the bin
#pragma unroll
        for (int k = 0; k < SORT_CAP / NT; k++) {
            const int i = threadIdx.x + k * NT;
            if (i < n) __stcg(reinterpret_cast<uint2 *>(bin) + i, e[k]);
        }
        __syncthreads();
        return true;
    }
    block_sort(bin, n, [](const uint4 &a, const uint4 &b) { return a.x < b.x; });
    __syncthreads();
    return false;
}
template <typename SM>
__device__ __forceinline__ bool sort_tile_list(SM &S, uint4 *bin, int n) {
    return sort_tile_list_ptr(S.u.sorted, bin, n);
}

// Stage the records of `cnt` chunk items in shared memory for a pixel-parallel walk (thread = item).  The staging
// area aliases the fragment pool, which these walks do not use.  FROM_TABLE: the item table of phase A0 already holds
// the record (cp.async); otherwise it is read from global memory.
template <bool FROM_TABLE, typename SM>
__device__ __forceinline__ void stage_big(const FrameParams &P, SM &S, int cnt, uint32_t key, uint32_t rec_tie, int bx0, int by0,
                                          int bw, int bh, uint32_t blocks) {
    if ((int)threadIdx.x < cnt) {
        BigSetup &b = S.u.big[threadIdx.x];
        const uint32_t rec = rec_tie & ENTRY_REC_MASK;
        float4 r0, r1;
        float2 r2;
        if (FROM_TABLE) {
            r0 = S.it_q0[threadIdx.x]; r1 = S.it_q1[threadIdx.x]; r2 = S.it_q2[threadIdx.x];
        } else {
            const float4 *rr = reinterpret_cast<const float4 *>(&P.recs[rec]);
            r0 = rr[0]; r1 = rr[1];
            const float4 q2 = rr[2];
            r2 = make_float2(q2.x, q2.y);
        }
        Setup s;
        s.px[0] = r0.x; s.py[0] = r0.y; s.px[1] = r0.z; s.py[1] = r0.w; s.px[2] = r1.x; s.py[2] = r1.y;
        setup_normals(s);
#pragma unroll
        for (int k = 0; k < 3; k++) {
            b.px[k] = s.px[k]; b.py[k] = s.py[k]; b.nx[k] = s.nx[k]; b.ny[k] = s.ny[k];
        }
        b.z[0] = r1.z; b.z[1] = r1.w; b.z[2] = r2.x;
        b.inv = r2.y; b.key = key; b.rec = rec;
        b.box = (uint32_t)bx0 | ((uint32_t)by0 << 8) | ((uint32_t)bw << 16) | ((uint32_t)bh << 24);
        b.tie = (rec_tie >> 29) | (blocks << 8); // tie-break bits of the three edges | 8x4 block mask << 8
    }
}
__device__ __forceinline__ void big_to_setup(const BigSetup &b, Setup &q) {
#pragma unroll
    for (int k = 0; k < 3; k++) {
        q.px[k] = b.px[k]; q.py[k] = b.py[k]; q.nx[k] = b.nx[k]; q.ny[k] = b.ny[k];
        q.z[k] = b.z[k];
    }
    q.inv = b.inv;
}

// Chunk dominated by large items: pixel-parallel walk with deferred shading (see the call site).  Only compiled
// into the DIRECT instantiations of the tile kernel: its register needs perturb the allocation of the fragment
// path that small-triangle workloads live in, so frames without large triangles run an instantiation that does not
// contain it.  The chunk's items are staged in S.u.big (stage_big) and S.it_key holds their order keys.
template <bool DBG, bool EXT>
__device__ __forceinline__ uint4 direct_chunk(const FrameParams &P, TileSmemT<DBG> &S, int cnt, bool sorted, int X, int Y) {
    const int tid = threadIdx.x;
    int lx, ly;
    pixel_of_thread(lx, ly);
    const int pix = ly * TW + lx; // this thread's slot in S.depth / S.color / S.okey
    uint32_t c_cov = 0, c_shaded = 0, c_samples = 0, c_oob = 0;
    const BigSetup *B = S.u.big;
    // the walk needs submission order: an unsorted chunk (it is the whole tile list then) is ranked by
    // order key right in the item table instead of being sorted and reloaded
    if (tid < cnt) {
        uint32_t rank = (uint32_t)tid;
        if (!sorted) {
            rank = 0;
            const uint32_t mykey = S.it_key[tid];
            for (int j = 0; j < cnt; j++) rank += S.it_key[j] < mykey ? 1u : 0u;
        }
        S.unit_item[rank] = (uint8_t)tid; // the unit table is free in this path
    }
    __syncthreads();
    float4 d = *reinterpret_cast<const float4 *>(&S.depth[pix * 4]);
    uint32_t own = 0xFFFFFFFFu; // 4 x u8: chunk item that owns sample k (0xFF: untouched in this chunk)
    uint32_t omp = 0u;          // 4 x 5 bits: that fragment's post-depth mask | (sample 0 covered) << 4
    for (int r = 0; r < cnt; r++) {
        const int it = (int)S.unit_item[r];
        if (!((B[it].tie >> (8 + (tid >> 5))) & 1u)) continue; // the binner proved this warp's 8x4 block uncovered (warp-uniform)
        const uint32_t box = B[it].box;
        const uint32_t rx = (uint32_t)lx - (box & 0xFFu), ry = (uint32_t)ly - ((box >> 8) & 0xFFu);
        if (rx >= ((box >> 16) & 0xFFu) || ry >= (box >> 24)) continue; // outside the in-tile box
        Setup q;
        big_to_setup(B[it], q);
        const uint32_t tie = B[it].tie & 7u;
        float thr[3];
#pragma unroll
        for (int k = 0; k < 3; k++) thr[k] = ((tie >> k) & 1u) ? 0.0f : 1.401298464e-45f;
        // (ENTRY_FULL is not used here: a second code path in this loop costs the general one 5 % -- measured on the C3 frame)
        const uint32_t m = coverage_mask_fast(q, thr, X, Y);
        if (!m) continue;
        c_cov++;
        uint32_t mp = 0;
        if (m & 1u) { const float z = sample_depth(q, X, Y, 0); if (z < d.x) { d.x = z; mp |= 1u; } } // strict < (mod.rs:374)
        if (m & 2u) { const float z = sample_depth(q, X, Y, 1); if (z < d.y) { d.y = z; mp |= 2u; } }
        if (m & 4u) { const float z = sample_depth(q, X, Y, 2); if (z < d.z) { d.z = z; mp |= 4u; } }
        if (m & 8u) { const float z = sample_depth(q, X, Y, 3); if (z < d.w) { d.w = z; mp |= 8u; } }
        if (!mp) continue;
        c_shaded++;
        c_samples += __popc(mp);
        const uint32_t tag = mp | ((m & 1u) << 4);
#pragma unroll
        for (int k = 0; k < 4; k++)
            if ((mp >> k) & 1u) {
                own = (own & ~(0xFFu << (8 * k))) | ((uint32_t)it << (8 * k));
                omp = (omp & ~(0x1Fu << (5 * k))) | (tag << (5 * k));
            }
    }
    *reinterpret_cast<float4 *>(&S.depth[pix * 4]) = d;
#pragma unroll
    for (int k = 0; k < 4; k++) {
        const uint32_t it = (own >> (8 * k)) & 0xFFu;
        if (it == 0xFFu) continue;
        bool first = true; // shade every surviving owner once (MSAA: one shader call per fragment)
#pragma unroll
        for (int j = 0; j < k; j++) first = first && ((own >> (8 * j)) & 0xFFu) != it;
        if (!first) continue;
        const uint32_t tag = (omp >> (5 * k)) & 0x1Fu;
        Setup q;
        big_to_setup(B[it], q);
        const auto depth0 = [&]() { return (tag & 16u) ? sample_depth(q, X, Y, 0) : 0.0f; }; // FragCoords.depths[0] (mod.rs:458-463)
        const uint32_t argb = shade<DBG, EXT>(P, q, B[it].rec, S.lut, X, Y, tag & 0xFu, depth0, c_oob);
#pragma unroll
        for (int j = k; j < 4; j++)
            if (((own >> (8 * j)) & 0xFFu) == it) {
                S.color[pix * 4 + j] = argb;
                if (DBG) S.okey[pix * 4 + j] = DBG ? S.it_okey[it] : 0u;
            }
    }
    __syncthreads();
    return make_uint4(c_cov, c_shaded, c_samples, c_oob);
}

// Tiles nothing was binned into: the box filter of four clear samples is the clear colour
// (buffers.rs:5,111-125).  A warp handles a GROUP of 8 horizontally adjacent tiles (4 lanes per tile, one
// 128-bit store per lane and row), so a row of empty tiles leaves the SM as 512 contiguous bytes -- full
// lines for HBM and, when the image lives in a peer GPU's memory, full-size NVLink write packets.  Every warp
// walks its own strided slice of the shard's groups (cursor g), at most max_visits groups per call.
constexpr int CLEAR_GROUP = 8;
template <bool DBG>
__device__ __forceinline__ void clear_empty_tiles(const FrameParams &P, int lane, uint32_t &g, uint32_t stride, uint32_t max_visits) {
    const uint32_t groups_x = (P.tiles_x + CLEAR_GROUP - 1) / CLEAR_GROUP;
    const uint32_t shard_groups = groups_x * (P.ty_end - P.ty_begin);
    for (uint32_t n = 0; g < shard_groups && n < max_visits; g += stride, n++) {
        const uint32_t ty = P.ty_begin + g / groups_x, tx = (g % groups_x) * CLEAR_GROUP + (uint32_t)lane / 4u;
        if (!owns_tile_row(P, ty)) continue;
        const bool empty = tx < P.tiles_x && P.tile_count[ty * P.tiles_x + tx] == 0u;
        const unsigned emask = DBG ? __ballot_sync(0xffffffffu, empty) : 0u; // bit 4t: tile t of the group is empty
        const int y0 = (int)ty * TH, xq = (int)tx * TW + (lane & 3) * 4;
        if (empty) {
            if ((P.W & 3u) == 0u) {
                if (xq < (int)P.W)
#pragma unroll 4
                    for (int row = 0; row < TH; row++)
                        if (y0 + row < (int)P.H)
                            *reinterpret_cast<uint4 *>(&P.out[(size_t)(y0 + row) * P.W + xq]) =
                                make_uint4(CLEAR_COLOR, CLEAR_COLOR, CLEAR_COLOR, CLEAR_COLOR);
            } else {
                for (int row = 0; row < TH; row++)
                    for (int k = 0; k < 4; k++)
                        if (y0 + row < (int)P.H && xq + k < (int)P.W) P.out[(size_t)(y0 + row) * P.W + xq + k] = CLEAR_COLOR;
            }
        }
        if (DBG && emask) { // parity instrumentation: the per-sample state of the empty tiles, coalesced over the group row
            const int gx0 = (int)((g % groups_x) * CLEAR_GROUP) * TW;
            for (int row = 0; row < TH && y0 + row < (int)P.H; row++)
                for (int j = lane; j < CLEAR_GROUP * TW * 4; j += 32) { // sample j of the group row
                    const int px = gx0 + j / 4;
                    if (!((emask >> (4 * (j / (TW * 4)))) & 1u) || px >= (int)P.W) continue;
                    const size_t o = ((size_t)(y0 + row) * P.W + px) * 4 + (j & 3);
                    if (P.dbg_depth) P.dbg_depth[o] = CLEAR_DEPTH;
                    if (P.dbg_color) P.dbg_color[o] = CLEAR_COLOR;
                    if (P.dbg_owner) P.dbg_owner[o] = NO_OWNER;
                }
        }
    }
}

// Tiles whose whole list has at most FAST_N items (full-screen quads, cube faces, the 8192^2 frames whose triangles span
// several tiles each): no item table, no scan, no fragment pool and a single barrier.  Every thread reads the few bin
// entries itself (uniform addresses: one transaction per warp), ranks them by order key, keeps its pixel's four
// samples in registers and applies the items in submission order -- exact coverage, sample depths, strict-< depth test
// (the literal sequence of rasterizer/mod.rs:443-473) -- with deferred shading: each surviving owner is shaded once.
// Returns false (nothing touched) when an item needs the literal walk.  The final samples are left in the thread's own
// words of S.depth / S.color / S.okey for the common epilogue.
template <bool DBG, bool EXT>
__device__ __forceinline__ bool fast_tile(const FrameParams &P, TileSmemT<DBG> &S, const uint4 *bin, int n, int X, int Y, uint32_t &c_cov,
                                          uint32_t &c_shaded, uint32_t &c_samples, uint32_t &c_oob) {
    const int tid = threadIdx.x;
    int lx, ly;
    pixel_of_thread(lx, ly);
    const int pix = ly * TW + lx; // this thread's slot in S.depth / S.color / S.okey
    static_assert(FAST_N <= 4, "the entries are ordered by a 4-element network");
    uint4 e[4];
    uint32_t anyw = 0;
#pragma unroll
    for (int i = 0; i < 4; i++) {
        e[i] = i < n ? __ldcg(bin + i) : make_uint4(0xFFFFFFFFu, 0u, 0u, 0u);
        if (i < n) anyw |= e[i].z;
    }
    if (anyw & ENTRY_WILD) return false;
    // submission order: bubble the (at most four) entries by key; uniform across the CTA, all in registers
#define RZ_CXE(a, b) { if (e[b].x < e[a].x) { const uint4 t_ = e[a]; e[a] = e[b]; e[b] = t_; } }
    RZ_CXE(0, 1) RZ_CXE(2, 3) RZ_CXE(0, 2) RZ_CXE(1, 3) RZ_CXE(1, 2)
#undef RZ_CXE
    float4 d = make_float4(CLEAR_DEPTH, CLEAR_DEPTH, CLEAR_DEPTH, CLEAR_DEPTH);
    uint32_t own = 0xFFFFFFFFu; // 4 x u8: item that owns sample k (0xFF: untouched)
    uint32_t omp = 0u;          // 4 x 5 bits: that fragment's post-depth mask | (sample 0 covered) << 4
#pragma unroll
    for (int it = 0; it < 4; it++) {
        if (it >= n) break;
        const uint32_t boxw = e[it].z;
        if (!((boxw >> (ENTRY_BLOCKS_SHIFT + (tid >> 5))) & 1u)) continue; // this warp's 8x4 block cannot be covered
        const uint32_t rx = (uint32_t)lx - (boxw & 15u), ry = (uint32_t)ly - ((boxw >> 4) & 15u);
        if (rx > ((boxw >> 8) & 15u) || ry > ((boxw >> 12) & 15u)) continue; // outside the in-tile box
        const float4 *rr = reinterpret_cast<const float4 *>(&P.recs[e[it].y & ENTRY_REC_MASK]);
        const float4 r0 = rr[0], r1 = rr[1];
        Setup q;
        q.px[0] = r0.x; q.py[0] = r0.y; q.px[1] = r0.z; q.py[1] = r0.w; q.px[2] = r1.x; q.py[2] = r1.y;
        setup_normals(q);
        float thr[3];
#pragma unroll
        for (int k = 0; k < 3; k++) thr[k] = ((e[it].y >> (29 + k)) & 1u) ? 0.0f : 1.401298464e-45f;
        const float4 r2 = rr[2];
        q.z[0] = r1.z; q.z[1] = r1.w; q.z[2] = r2.x;
        q.inv = r2.y;
        uint32_t mp;
        const uint32_t m = cover_depth4(q, thr, (boxw & ENTRY_FULL) != 0u, X, Y, d, mp); // (uniform across the CTA)
        if (!m) continue;
        c_cov++;
        if (!mp) continue;
        c_shaded++;
        c_samples += __popc(mp);
        const uint32_t tag = mp | ((m & 1u) << 4);
#pragma unroll
        for (int k = 0; k < 4; k++)
            if ((mp >> k) & 1u) {
                own = (own & ~(0xFFu << (8 * k))) | ((uint32_t)it << (8 * k));
                omp = (omp & ~(0x1Fu << (5 * k))) | (tag << (5 * k));
            }
    }
    *reinterpret_cast<float4 *>(&S.depth[pix * 4]) = d;
#pragma unroll
    for (int k = 0; k < 4; k++) {
        const uint32_t it = (own >> (8 * k)) & 0xFFu;
        if (it == 0xFFu) continue;
        bool first = true; // shade every surviving owner once (MSAA: one shader call per fragment)
#pragma unroll
        for (int j = 0; j < k; j++) first = first && ((own >> (8 * j)) & 0xFFu) != it;
        if (!first) continue;
        const uint32_t tag = (omp >> (5 * k)) & 0x1Fu;
        const uint4 ei = it == 0u ? e[0] : (it == 1u ? e[1] : (it == 2u ? e[2] : e[3]));
        const uint32_t rec = ei.y & ENTRY_REC_MASK;
        const float4 *rr = reinterpret_cast<const float4 *>(&P.recs[rec]);
        const float4 r0 = rr[0], r1 = rr[1];
        Setup q;
        q.px[0] = r0.x; q.py[0] = r0.y; q.px[1] = r0.z; q.py[1] = r0.w; q.px[2] = r1.x; q.py[2] = r1.y;
        setup_normals(q);
        const auto depth0 = [&]() { // FragCoords.depths[0] (mod.rs:458-463)
            if (!(tag & 16u)) return 0.0f;
            const float4 r2 = rr[2];
            Setup qd = q;
            qd.z[0] = r1.z; qd.z[1] = r1.w; qd.z[2] = r2.x;
            qd.inv = r2.y;
            return sample_depth(qd, X, Y, 0);
        };
        const uint32_t argb = shade<DBG, EXT>(P, q, rec, S.lut, X, Y, tag & 0xFu, depth0, c_oob);
#pragma unroll
        for (int j = k; j < 4; j++)
            if (((own >> (8 * j)) & 0xFFu) == it) {
                S.color[pix * 4 + j] = argb;
                if (DBG) S.okey[pix * 4 + j] = ei.x;
            }
    }
    return true;
}

template <bool DBG, bool EXT, bool DIRECT>
__global__ void __launch_bounds__(NT, DBG ? 3 : RZ_TILE_CTAS) tile_kernel(FrameParams P) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    typedef TileSmemT<DBG> SM;
    SM &S = *reinterpret_cast<SM *>(smem_raw);

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    int lx, ly;
    pixel_of_thread(lx, ly);
    const int pix = ly * TW + lx; // this thread's pixel in the pixel-parallel walks and in the resolve
    uint32_t c_cov = 0, c_shaded = 0, c_samples = 0, c_oob = 0;
    pdl_launch();
    pdl_wait();

    // ---- phase 1: persistent loop over the tiles that received triangles (dynamic work stealing) ----
    S.lut[tid] = fdiv((float)tid, 255.0f);
    const bool wild = P.fs->has_wild != 0u; // the frame holds items for the literal per-pixel walk
    // prefix of the class sizes: work item w belongs to the first class with w < end
    // (kept in shared memory: eight registers held over the whole tile loop were spilled instead)
    if (tid == 0) {
        uint32_t acc = 0;
#pragma unroll
        for (int b = 0; b < ORDER_BUCKETS; b++) {
            acc += P.fs->bucket_n[b];
            S.bucket_end[b] = acc;
        }
    }
    // empty-tile clears: with a caller-owned destination they are spread over the loop (a few tile groups per
    // rasterised tile, cursor kept in shared memory), the rest follows after the loop
    if (lane == 0) S.clr_cursor[warp] = blockIdx.x * (NT / 32) + warp;
    if (tid == 0) S.unit_budget = UNIT_CAP;
    // Work stealing.  After the last phase barrier of a tile, thread 0 steals the next work index and loads that tile's
    // busy entry {tile id, list length, first bin entry}; both round trips overlap the write-back of the current tile,
    // the entry is published in shared memory at the very end of the trip and read by everybody after the barrier at the
    // top of the next one, so a trip starts with its bin address in hand.  (Stealing further ahead -- reserving a tile
    // while another is still being processed -- was measured: the reserved tiles of slow CTAs are held hostage at the
    // end of the frame, C2 tile stage 83 -> 131 us.)
    auto load_entry = [&](uint32_t w) -> uint4 {
        const uint32_t n_busy_ = S.bucket_end[ORDER_BUCKETS - 1];
        if (w >= n_busy_) return make_uint4(0xFFFFFFFFu, 0u, 0u, 0u);
        uint32_t b = 0, start = 0;
#pragma unroll
        for (int k = 0; k < ORDER_BUCKETS - 1; k++)
            if (w >= S.bucket_end[k]) {
                b = k + 1;
                start = S.bucket_end[k];
            }
        uint4 e = __ldcg(P.busy + (size_t)b * P.tiles_x * P.tiles_y + (w - start));
        e.w = w;
        return e;
    };
    if (tid == 0) S.ent = load_entry(atomicAdd(&P.fs->tile_cursor, 1u)); // (S.bucket_end was written by this same thread)
    for (;;) {
    __syncthreads(); // previous tile fully retired (also covers S.lut and S.bucket_end on the first trip)
    const uint4 ent = S.ent;
    if (ent.x == 0xFFFFFFFFu) break;
    const uint32_t tile = ent.x, bin_off = ent.z, work = ent.w;
    int n = (int)ent.y;
    const uint32_t tx = tile % P.tiles_x, ty = tile / P.tiles_x;
    const int tileX0 = tx * TW, tileY0 = ty * TH;
    const int X = tileX0 + lx, Y = tileY0 + ly;
    unsigned long long t_start = 0, t_ph[5] = {0, 0, 0, 0, 0};
#define RZ_STAMP(k)                                                                                  \
    if (DBG && P.dbg_tile_time && tid == 0 && t_ph[k] == 0) asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t_ph[k]));
    if (DBG && P.dbg_tile_time && tid == 0) asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t_start));

    // clear (the state resolve_and_clear leaves behind, rasterizer/mod.rs:497-506)
#pragma unroll
    for (int k = 0; k < 4; k++) {
        S.depth[pix * 4 + k] = CLEAR_DEPTH; // (its own pixel's slot: the barrier-free walk writes it without a barrier in between)
        S.color[pix * 4 + k] = CLEAR_COLOR;
        if (DBG) S.okey[pix * 4 + k] = NO_OWNER;
    }

    bool fast_done = false;
    if (DIRECT && n > 0 && n <= FAST_N && !wild) {
        fast_done = fast_tile<DBG, EXT>(P, S, P.bins + bin_off, n, X, Y, c_cov, c_shaded, c_samples, c_oob);
        __syncthreads(); // every thread has read S.ent (thread 0 overwrites it at the end of the trip)
    }
    if (n > 0 && !fast_done) {
        S.head[tid] = FR_NONE;
        uint4 *bin = P.bins + bin_off;
        bool sorted = false;  // the list is in submission order
        bool compact = false; // ... and in its compact form (sort_tile_list)
        if (n > CHUNK) {
            compact = sort_tile_list(S, bin, n);
            sorted = true;
        }
        // (no barrier needed here: sort_tile_list ends with one, and the clears above are separated from their
        // first readers by the barriers of phase A0)

        for (int pos = 0; pos < n;) {
            // ---- load this window's entries (thread = item) ----
            const int item = pos + tid;
            const bool valid = tid < CHUNK && item < n;
            uint32_t rec_tie = 0, key = 0, blocks = 0;
            int bx0 = 0, by0 = 0, bw = 0, bh = 0; // in-tile box, tile-local origin
            bool big = false;
            if (valid) {
                uint32_t boxw;
                if (compact) {
                    const uint2 e = __ldcg(reinterpret_cast<const uint2 *>(bin) + item);
                    key = (uint32_t)item; // sorted: the rank orders the items
                    rec_tie = e.x; boxw = e.y;
                } else {
                    const uint4 e = __ldcg(bin + item);
                    key = e.x; rec_tie = e.y; boxw = e.z;
                }
                bx0 = (int)(boxw & 15u); by0 = (int)((boxw >> 4) & 15u);
                bw = (int)((boxw >> 8) & 15u) + 1; bh = (int)((boxw >> 12) & 15u) + 1;
                blocks = (boxw >> ENTRY_BLOCKS_SHIFT) & 0xFFu;
                big = wild && (boxw & ENTRY_WILD) != 0u; // NaN / inf / absurd coordinates: literal per-pixel path
                if (DBG) S.it_okey[tid] = compact ? P.recs[rec_tie & ENTRY_REC_MASK].key : key;
            }
            const int nvalid = min(CHUNK, n - pos);
            // items for the literal walk split the window into runs; a frame without any has none to look for
            int first_big = CHUNK + 1, first_small = 0;
            if (wild) {
                if (tid == 0) {
                    S.first_big = CHUNK + 1;
                    S.first_small = CHUNK + 1;
                }
                __syncthreads();
                if (valid && big) atomicMin(&S.first_big, (uint32_t)tid);
                if (valid && !big) atomicMin(&S.first_small, (uint32_t)tid);
                __syncthreads();
                first_big = (int)S.first_big;
                first_small = (int)S.first_small;
            }

            if (!sorted && first_big <= CHUNK) { // large items need the ordered walk
                compact = sort_tile_list(S, bin, n);
                sorted = true;
                continue;
            }

            if (first_big == 0) {
                // ================= run of items for the literal walk: pixel-parallel =================
                const int run = min(first_small, nvalid);
                const BigSetup *B = S.u.big;
                stage_big<false>(P, S, run, key, rec_tie, bx0, by0, bw, bh, blocks);
                __syncthreads();
                float d[4];
                uint32_t col[4], ok[4];
#pragma unroll
                for (int k = 0; k < 4; k++) {
                    d[k] = S.depth[pix * 4 + k];
                    col[k] = S.color[pix * 4 + k];
                    if (DBG) ok[k] = S.okey[pix * 4 + k];
                }
                for (int it = 0; it < run; it++) {
                    const uint32_t box = B[it].box;
                    const uint32_t rx = (uint32_t)lx - (box & 0xFF), ry = (uint32_t)ly - ((box >> 8) & 0xFF);
                    if (rx >= ((box >> 16) & 0xFF) || ry >= (box >> 24)) continue;
                    Setup q;
                    big_to_setup(B[it], q);
                    const uint32_t m = coverage_mask(q, X, Y);
                    if (!m) continue;
                    c_cov++;
                    float zs[4];
                    uint32_t mp = 0;
#pragma unroll
                    for (int k = 0; k < 4; k++) {
                        zs[k] = ((m >> k) & 1u) ? sample_depth(q, X, Y, k) : 0.0f;
                        if (((m >> k) & 1u) && zs[k] < d[k]) mp |= 1u << k; // strict < (mod.rs:374)
                    }
                    if (!mp) continue;
                    c_shaded++;
                    c_samples += __popc(mp);
                    const uint32_t argb = shade<true, EXT>(P, q, B[it].rec, S.lut, X, Y, mp, [&]() { return zs[0]; }, c_oob);
#pragma unroll
                    for (int k = 0; k < 4; k++)
                        if ((mp >> k) & 1u) {
                            d[k] = zs[k];
                            col[k] = argb;
                            if (DBG) ok[k] = DBG ? S.it_okey[it] : 0u;
                        }
                }
#pragma unroll
                for (int k = 0; k < 4; k++) {
                    S.depth[pix * 4 + k] = d[k];
                    S.color[pix * 4 + k] = col[k];
                    if (DBG) S.okey[pix * 4 + k] = ok[k];
                }
                __syncthreads();
                pos += run;
                continue;
            }

            // ================= chunk of small items =================
            int cnt = sorted ? min(first_big, nvalid) : nvalid;
            // ---- phase A0: item table + exclusive scan of the in-tile box areas ----
            // work units of phase A1 = horizontal pixel PAIRS of the in-tile box (the two pixels share the y terms of
            // the twelve edge evaluations and the unit's decoding); a box of odd width ends its rows with a half-used pair
            const int pw = (bw + 1) >> 1;
            const uint32_t area = (tid < cnt) ? (uint32_t)(pw * bh) : 0u;
            if (tid < cnt) {
                // the record travels to shared memory on its own while the scan below runs (waited for at the barrier
                // that ends phase A0)
                const char *rp = reinterpret_cast<const char *>(&P.recs[rec_tie & ENTRY_REC_MASK]);
                cp_async16(&S.it_q0[tid], rp);
                cp_async16(&S.it_q1[tid], rp + 16);
                cp_async8(&S.it_q2[tid], rp + 32);
                S.it_key[tid] = key; S.it_rec[tid] = rec_tie;
                S.it_box[tid] = (uint32_t)bx0 | ((uint32_t)by0 << 4) | ((uint32_t)bw << 8) | (((65536u + (uint32_t)pw - 1u) / (uint32_t)pw) << 13);
            }
            uint32_t incl = area;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const uint32_t v = __shfl_up_sync(0xffffffffu, incl, o);
                if (lane >= o) incl += v;
            }
            if (lane == 31) S.scan[warp] = incl;
            if (tid == 0) {
                S.nfrag = 0;
                S.ovf = 0;
            }
            __syncthreads();
            uint32_t wbase = 0, total_units = 0;
#pragma unroll
            for (int k = 0; k < NT / 32; k++) {
                const uint32_t v = S.scan[k];
                if (k < warp) wbase += v;
                total_units += v;
            }
            // ================= chunk dominated by large items: pixel-parallel, deferred shading =================
            // When the items of the chunk cover the tile broadly (average in-tile box >= DIRECT_MIN_AREA pixels) the
            // fragment machinery below only adds barriers: every thread keeps its pixel, walks the chunk's items in
            // submission order (exact coverage, sample depths, strict-< depth test: the literal sequence of
            // rasterizer/mod.rs:443-473), remembers per sample which item wrote it last together with that fragment's
            // post-depth mask, and shades each surviving owner once at the end.  No fragment pool, no unit table.
            const bool go_direct = DIRECT && cnt > 0 && 2u * total_units >= (uint32_t)DIRECT_MIN_AREA * (uint32_t)cnt; // (units are pixel pairs)
            {
                const uint32_t first = wbase + incl - area; // exclusive prefix: first work unit of item <tid>
                S.pre[tid] = first;
                if (tid == NT - 1) S.pre[NT] = wbase + incl;
                if (!go_direct)
                    for (uint32_t k = 0; k < area && first + k < (uint32_t)UNIT_CAP; k++) S.unit_item[first + k] = (uint8_t)tid;
            }
            cp_async_wait_all();
            __syncthreads();
            if (go_direct) {
                stage_big<true>(P, S, cnt, key, rec_tie, bx0, by0, bw, bh, blocks);
                const uint4 dc = direct_chunk<DBG, EXT>(P, S, cnt, sorted, X, Y); // (starts with a barrier)
                c_cov += dc.x; c_shaded += dc.y; c_samples += dc.z; c_oob += dc.w;
                pos += cnt;
                continue;
            }

            // unit budget of this chunk: what the unit table can index, and what the fragment pool is expected to
            // hold given the fragments-per-unit ratio the last chunks of this CTA saw (a chunk whose fragments
            // overflow the pool has to redo its whole coverage pass, so it is cheaper to cut it beforehand)
            // S.unit_budget aims at 7/8 of the pool; a chunk is only cut when the prediction exceeds the pool by
            // a clear margin (~1.2x), borderline chunks are simply tried
            const uint32_t budget = min((uint32_t)UNIT_CAP, S.unit_budget);
            if (S.pre[cnt] > min((uint32_t)UNIT_CAP, budget + budget * RZ_BUDGET_TRIG / 8u)) {
                // keep the longest prefix of items that fits
                // (chunks must follow submission order, so an unsorted list is sorted first)
                if (!sorted) {
                    compact = sort_tile_list(S, bin, n);
                    sorted = true;
                    continue;
                }
                int lo = 1, hi = cnt; // largest c with pre[c] <= budget (pre[1] <= 256 always fits)
                while (hi - lo > 0) {
                    const int mid = (lo + hi + 1) >> 1;
                    if (S.pre[mid] <= budget) lo = mid; else hi = mid - 1;
                }
                cnt = lo;
            }

            RZ_STAMP(0) // A0 done (entries loaded, item table + scan)
            bool need_sort = false;
            for (;;) {
                // ---- phase A1: thread = (item, box pixel) work unit, consecutive lanes = consecutive units ----
                const int units = (int)S.pre[cnt];
                uint32_t cov_try = 0;
                for (int u0 = 0; u0 < units; u0 += NT) {
                    const int u = u0 + tid;
                    uint32_t m0 = 0, m1 = 0, p = 0, it = 0; // coverage of the pair's two pixels p, p + 1
                    if (u < units) {
                        it = S.unit_item[u];
                        const uint32_t box = S.it_box[it];
                        const int ibw = (int)((box >> 8) & 0x1Fu), ipw = (ibw + 1) >> 1;
                        const int j = u - (int)S.pre[it];
                        const int ry = (int)(((uint32_t)j * (box >> 13)) >> 16), rx = 2 * (j - ry * ipw); // j / ipw, j < 256, ipw <= 8
                        const int lpx = (int)(box & 0xFu) + rx, lpy = (int)((box >> 4) & 0xFu) + ry;
                        const bool second = rx + 1 < ibw; // (false: the row's last, half-used pair)
                        p = (uint32_t)(lpy * TW + lpx);
                        const uint32_t rec_t = S.it_rec[it];
                        const float4 r0 = S.it_q0[it], r1 = S.it_q1[it];
                        // EdgeFunctions normals (mod.rs:200-205) and the single-compare form of inside() (rz_exact.cuh)
                        const float n0x = -fsub(r0.w, r0.y), n0y = fsub(r0.z, r0.x);
                        const float n1x = -fsub(r1.y, r0.w), n1y = fsub(r1.x, r0.z);
                        const float n2x = -fsub(r0.y, r1.y), n2y = fsub(r0.x, r1.x);
                        const float t0 = (rec_t & (1u << 29)) ? 0.0f : 1.401298464e-45f;
                        const float t1 = (rec_t & (1u << 30)) ? 0.0f : 1.401298464e-45f;
                        const float t2 = (rec_t & (1u << 31)) ? 0.0f : 1.401298464e-45f;
                        const float fx = (float)(tileX0 + lpx), fx1 = (float)(tileX0 + lpx + 1), fy = (float)(tileY0 + lpy);
#pragma unroll
                        for (int i = 0; i < 4; i++) {
                            const float xs = fadd(fx, rgss_x(i)), xs1 = fadd(fx1, rgss_x(i)), ys = fadd(fy, rgss_y(i));
                            const float y0 = fmul(n0y, fsub(ys, r0.y)), y1 = fmul(n1y, fsub(ys, r0.w)), y2 = fmul(n2y, fsub(ys, r1.y));
                            {
                                const float e0 = fadd(fmul(n0x, fsub(xs, r0.x)), y0);
                                const float e1 = fadd(fmul(n1x, fsub(xs, r0.z)), y1);
                                const float e2 = fadd(fmul(n2x, fsub(xs, r1.x)), y2);
                                m0 |= ((e0 >= t0) & (e1 >= t1) & (e2 >= t2)) ? (1u << i) : 0u;
                            }
                            {
                                const float e0 = fadd(fmul(n0x, fsub(xs1, r0.x)), y0);
                                const float e1 = fadd(fmul(n1x, fsub(xs1, r0.z)), y1);
                                const float e2 = fadd(fmul(n2x, fsub(xs1, r1.x)), y2);
                                m1 |= ((e0 >= t0) & (e1 >= t1) & (e2 >= t2)) ? (1u << i) : 0u;
                            }
                        }
                        if (!second) m1 = 0u;
                    }
                    const uint32_t bal0 = __ballot_sync(0xffffffffu, m0 != 0u), bal1 = __ballot_sync(0xffffffffu, m1 != 0u);
                    if (bal0 | bal1) { // warp-aggregated fragment allocation (ballots + popc prefix): first pixels, then second pixels
                        uint32_t slot = 0;
                        if (lane == 0) slot = atomicAdd(&S.nfrag, (uint32_t)(__popc(bal0) + __popc(bal1)));
                        slot = __shfl_sync(0xffffffffu, slot, 0);
                        if (m0) {
                            const uint32_t s0 = slot + __popc(bal0 & lanemask_lt());
                            cov_try++;
                            if (s0 < POOL) {
                                S.u.fr.meta[s0] = it | (p << 8) | (m0 << 16);
                                S.u.fr.next[s0] = (uint16_t)atomicExch(&S.head[p], s0);
                            } else {
                                S.ovf = 1u;
                            }
                        }
                        if (m1) {
                            const uint32_t s1 = slot + __popc(bal0) + __popc(bal1 & lanemask_lt());
                            cov_try++;
                            if (s1 < POOL) {
                                S.u.fr.meta[s1] = it | ((p + 1u) << 8) | (m1 << 16);
                                S.u.fr.next[s1] = (uint16_t)atomicExch(&S.head[p + 1u], s1);
                            } else {
                                S.ovf = 1u;
                            }
                        }
                    }
                }
                __syncthreads();
                // S.nfrag keeps counting past the pool: the chunk's true fragment count.  fit_units() = the units that
                // would have filled 7/8 of the pool at this chunk's fragments-per-unit ratio (units <= 6144: 32-bit)
                if (!S.ovf) {
                    c_cov += cov_try;
                    if (tid == 0) S.unit_budget = min((uint32_t)UNIT_CAP, (3u * S.unit_budget + fit_units((uint32_t)units, S.nfrag)) / 4u + 64u);
                    break;
                }
                const uint32_t fit = fit_units((uint32_t)units, S.nfrag);
                // The chunk's fragments do not fit the pool.  Nothing has touched the tile state yet: drop them
                // and retry with the prefix of items the measured ratio says will fit.  Chunks must follow
                // submission order, so an unsorted list is sorted first.
                __syncthreads();
                S.head[tid] = FR_NONE;
                if (tid == 0) {
                    S.nfrag = 0;
                    S.ovf = 0;
                    S.unit_budget = fit;
                }
                if (!sorted) {
                    need_sort = true;
                    break;
                }
                {
                    int lo = 1, hi = max(1, cnt - 1); // largest c < cnt with pre[c] <= fit (fit < units, so it shrinks)
                    while (hi - lo > 0) {
                        const int mid = (lo + hi + 1) >> 1;
                        if (S.pre[mid] <= fit) lo = mid; else hi = mid - 1;
                    }
                    cnt = lo;
                }
                __syncthreads();
            }
            if (need_sort) {
                compact = sort_tile_list(S, bin, n);
                sorted = true;
                continue;
            }
            RZ_STAMP(1) // A1 done
            const int nfrag = (int)S.nfrag;
            // ---- phase A2: thread = fragment; the covered samples' depths (RasterizerTriangle::fragment, mod.rs:225-253),
            // every lane busy (in phase A1 only ~40 % of the units are covered)
            for (int f = tid; f < nfrag; f += NT) {
                const uint32_t meta = S.u.fr.meta[f];
                const uint32_t it = meta & 0xFFu, p = (meta >> 8) & 0xFFu, m = (meta >> 16) & 0xFu;
                const float4 r0 = S.it_q0[it], r1 = S.it_q1[it];
                const float2 r2 = S.it_q2[it];
                const float n1x = -fsub(r1.y, r0.w), n1y = fsub(r1.x, r0.z);
                const float n2x = -fsub(r0.y, r1.y), n2y = fsub(r0.x, r1.x);
                const float fx = (float)(tileX0 + (int)(p % TW)), fy = (float)(tileY0 + (int)(p / TW));
                float zz[4];
#pragma unroll
                for (int i = 0; i < 4; i++) {
                    const float xs = fadd(fx, rgss_x(i)), ys = fadd(fy, rgss_y(i));
                    const float e1 = dot2z(n1x, n1y, fsub(xs, r0.z), fsub(ys, r0.w)); // eval_single (mod.rs:125-132)
                    const float e2 = dot2z(n2x, n2y, fsub(xs, r1.x), fsub(ys, r1.y));
                    const float b0 = clamp01(fmul(e1, r2.y));
                    const float b1 = clamp01(fmul(e2, r2.y));
                    const float b2 = clamp01(fsub(fsub(1.0f, b0), b1));
                    const float z = fadd(fadd(fmul(b0, r1.z), fmul(b1, r1.w)), fmul(b2, r2.x));
                    zz[i] = ((m >> i) & 1u) ? z : 0.0f;
                }
                S.u.fr.z[f] = make_float4(zz[0], zz[1], zz[2], zz[3]);
            }
            __syncthreads();
            RZ_STAMP(2) // sample depths done
            // ---- phase B: thread = pixel; replay this pixel's fragments in submission order ----
            // (depth test exactly as Rasterizer::depth_coverage + write_pixel, mod.rs:363-397); the last writer
            // of every sample is the fragment that stays visible.  A pixel holds 2-3 fragments on average.
            {
                const uint32_t h = S.head[tid];
                const bool reg_lists = nfrag <= RZ_B_REG_MAXFRAG; // dense chunks: most lists exceed the register path
                if (h != FR_NONE) {
                    float4 d = *reinterpret_cast<const float4 *>(&S.depth[tid * 4]);
                    uint32_t own0 = FR_NONE, own1 = FR_NONE, own2 = FR_NONE, own3 = FR_NONE;
                    // one fragment of the replay: Rasterizer::depth_coverage + the depth half of write_pixel
                    auto replay = [&](uint32_t g) {
                        const uint32_t m = (S.u.fr.meta[g] >> 16) & 0xFu;
                        const float4 z = S.u.fr.z[g];
                        uint32_t mp = 0;
                        if ((m & 1u) && z.x < d.x) { mp |= 1u; d.x = z.x; own0 = g; } // strict < (mod.rs:374)
                        if ((m & 2u) && z.y < d.y) { mp |= 2u; d.y = z.y; own1 = g; }
                        if ((m & 4u) && z.z < d.z) { mp |= 4u; d.z = z.z; own2 = g; }
                        if ((m & 8u) && z.w < d.w) { mp |= 8u; d.w = z.w; own3 = g; }
                        S.u.fr.fin[g] = (uint8_t)mp; // post-depth mask; the visible bits are added below
                        if (mp) {
                            c_shaded++;
                            c_samples += __popc(mp);
                        }
                    };
                    // Lists of up to four fragments (nearly all of them) are walked ONCE: every fragment goes into
                    // a register as key << 32 | coverage << 16 | fragment, a 5-exchange network orders them by key,
                    // absent slots carry the key 0xFFFFFFFF (no triangle has it: a frame holds < 2^29 triangles,
                    // key = 8 * number + fan index) and sort to the end.  Keys within a pixel list are distinct (one
                    // fragment per item).  The masks stay in registers: one byte store per fragment at the end.
                    const unsigned long long ABSENT = 0xFFFFFFFF00000000ull | FR_NONE;
                    unsigned long long e0 = ABSENT, e1 = ABSENT, e2 = ABSENT, e3 = ABSENT;
                    uint32_t g = h;
#define RZ_ENTRY(e)                                                                                          \
    {                                                                                                        \
        const uint32_t meta_ = S.u.fr.meta[g];                                                               \
        e = ((unsigned long long)S.it_key[meta_ & 0xFFu] << 32) | (meta_ & 0xF0000u) | g;                    \
        g = S.u.fr.next[g];                                                                                  \
    }
                    if (reg_lists) RZ_ENTRY(e0)
                    if (reg_lists && g != FR_NONE) {
                        RZ_ENTRY(e1)
                        if (g != FR_NONE) {
                            RZ_ENTRY(e2)
                            if (g != FR_NONE) RZ_ENTRY(e3)
                        }
                    }
#undef RZ_ENTRY
                    bool replayed = false;
                    if (reg_lists && g == FR_NONE) {
                        replayed = true;
#define RZ_CX(a, b) { const unsigned long long lo_ = min(a, b), hi_ = max(a, b); a = lo_; b = hi_; }
                        RZ_CX(e0, e1) RZ_CX(e2, e3) RZ_CX(e0, e2) RZ_CX(e1, e3) RZ_CX(e1, e2)
#undef RZ_CX
                        uint32_t mpk = 0u;      // post-depth mask of slot i in bits 4i..4i+3
                        uint32_t ownk = 0xFFFFu; // slot that wrote sample s last in bits 4s..4s+3 (0xF: none)
#define RZ_SLOT(i, e)                                                                                        \
    if ((i) == 0 || e != ABSENT) {                                                                           \
        const uint32_t m = ((uint32_t)e >> 16) & 0xFu;                                                       \
        const float4 z = S.u.fr.z[(uint32_t)e & 0xFFFFu];                                                    \
        uint32_t mp = 0;                                                                                     \
        if ((m & 1u) && z.x < d.x) { mp |= 1u; d.x = z.x; ownk = (ownk & ~0x000Fu) | (uint32_t)(i); }        \
        if ((m & 2u) && z.y < d.y) { mp |= 2u; d.y = z.y; ownk = (ownk & ~0x00F0u) | ((uint32_t)(i) << 4); } \
        if ((m & 4u) && z.z < d.z) { mp |= 4u; d.z = z.z; ownk = (ownk & ~0x0F00u) | ((uint32_t)(i) << 8); } \
        if ((m & 8u) && z.w < d.w) { mp |= 8u; d.w = z.w; ownk = (ownk & ~0xF000u) | ((uint32_t)(i) << 12); }\
        mpk |= mp << (4 * (i));                                                                              \
        if (mp) {                                                                                            \
            c_shaded++;                                                                                      \
            c_samples += __popc(mp);                                                                         \
        }                                                                                                    \
    }
                        RZ_SLOT(0, e0) RZ_SLOT(1, e1) RZ_SLOT(2, e2) RZ_SLOT(3, e3)
#undef RZ_SLOT
                        // fin = post-depth mask | still-visible mask << 4 (the samples this slot wrote last)
#define RZ_FIN(i, e)                                                                                         \
    if ((i) == 0 || e != ABSENT) {                                                                           \
        const uint32_t vis = ((ownk & 0xFu) == (uint32_t)(i) ? 1u : 0u) | (((ownk >> 4) & 0xFu) == (uint32_t)(i) ? 2u : 0u) | \
                             (((ownk >> 8) & 0xFu) == (uint32_t)(i) ? 4u : 0u) | (((ownk >> 12) & 0xFu) == (uint32_t)(i) ? 8u : 0u); \
        S.u.fr.fin[(uint32_t)e & 0xFFFFu] = (uint8_t)(((mpk >> (4 * (i))) & 0xFu) | (vis << 4));             \
    }
                        RZ_FIN(0, e0) RZ_FIN(1, e1) RZ_FIN(2, e2) RZ_FIN(3, e3)
#undef RZ_FIN
                    }
                    if (!replayed)
                    {
                        // longer lists: the next fragment in key order is found by re-walking the list
                        uint32_t last_key = 0;
                        bool first = true;
                        for (;;) {
                            uint32_t best = FR_NONE, best_key = 0xFFFFFFFFu;
                            for (uint32_t g2 = h; g2 != FR_NONE; g2 = S.u.fr.next[g2]) {
                                const uint32_t k = S.it_key[S.u.fr.meta[g2] & 0xFFu];
                                if ((first || k > last_key) && k < best_key) {
                                    best = g2;
                                    best_key = k;
                                }
                            }
                            if (best == FR_NONE) break;
                            replay(best);
                            last_key = best_key;
                            first = false;
                        }
                    }
                    if (own0 != FR_NONE) S.u.fr.fin[own0] |= 0x10u;
                    if (own1 != FR_NONE) S.u.fr.fin[own1] |= 0x20u;
                    if (own2 != FR_NONE) S.u.fr.fin[own2] |= 0x40u;
                    if (own3 != FR_NONE) S.u.fr.fin[own3] |= 0x80u;
                    *reinterpret_cast<float4 *>(&S.depth[tid * 4]) = d;
                    S.head[tid] = FR_NONE;
                }
            }
            __syncthreads();
            RZ_STAMP(3) // B done
            // ---- phase C: thread = fragment; shade what is still visible and write its samples ----
            for (int f = tid; f < nfrag; f += NT) {
                const uint32_t fin = S.u.fr.fin[f];
                const uint32_t meta = S.u.fr.meta[f];
                const uint32_t p = (meta >> 8) & 0xFFu;
                const uint32_t vis = fin >> 4;
                if (!vis) continue;
                const uint32_t it = meta & 0xFFu;
                const uint32_t rec = S.it_rec[it] & ENTRY_REC_MASK;
                Setup q;
                {
                    const float4 r0 = S.it_q0[it], r1 = S.it_q1[it];
                    q.px[0] = r0.x; q.py[0] = r0.y; q.px[1] = r0.z; q.py[1] = r0.w; q.px[2] = r1.x; q.py[2] = r1.y;
                    setup_normals(q);
                }
                const uint32_t argb = shade<DBG, EXT>(P, q, rec, S.lut, tileX0 + (int)(p % TW), tileY0 + (int)(p / TW),
                                            fin & 0xFu, [&]() { return S.u.fr.z[f].x; }, c_oob);
#pragma unroll
                for (int k = 0; k < 4; k++)
                    if ((vis >> k) & 1u) {
                        S.color[p * 4 + k] = argb;
                        if (DBG) S.okey[p * 4 + k] = DBG ? S.it_okey[it] : 0u;
                    }
            }
            __syncthreads();
            RZ_STAMP(4) // C done
            pos += cnt;
        }
    }
    else if (!fast_done) __syncthreads(); // (a busy tile never has an empty list; keeps the hand-over below safe if it ever did)
    // every path through the chunk loop ends with a barrier: all threads have read S.ent long ago
    uint4 ent_next = make_uint4(0u, 0u, 0u, 0u);
    if (tid == 0) ent_next = load_entry(atomicAdd(&P.fs->tile_cursor, 1u)); // issued here, consumed at the end of the trip
    // ---- resolve (ColorBuffer::box_filter_color, buffers.rs:111-125) and write back ----
    const uint32_t res = box_filter(S.color[pix * 4], S.color[pix * 4 + 1], S.color[pix * 4 + 2], S.color[pix * 4 + 3]);
    if (DBG && X < (int)P.W && Y < (int)P.H) {
        const size_t o = ((size_t)Y * P.W + X) * 4;
#pragma unroll
        for (int k = 0; k < 4; k++) {
            if (P.dbg_depth) P.dbg_depth[o + k] = S.depth[pix * 4 + k];
            if (P.dbg_color) P.dbg_color[o + k] = S.color[pix * 4 + k];
            if (P.dbg_owner) P.dbg_owner[o + k] = S.okey[pix * 4 + k];
        }
    }
    if ((P.W & 3u) == 0u) {
        // 128-bit row stores without staging: lane L (L % 4 == 0) collects the pixels of lanes L..L+3, which are
        // four consecutive pixels of one row of the warp's 8x4 block (pixel_of_thread)
        const uint32_t r1 = __shfl_down_sync(0xffffffffu, res, 1), r2 = __shfl_down_sync(0xffffffffu, res, 2),
                       r3 = __shfl_down_sync(0xffffffffu, res, 3);
        if ((lane & 3) == 0 && X < (int)P.W && Y < (int)P.H)
            *reinterpret_cast<uint4 *>(&P.out[(size_t)Y * P.W + X]) = make_uint4(res, r1, r2, r3);
    } else if (X < (int)P.W && Y < (int)P.H) {
        P.out[(size_t)Y * P.W + X] = res;
    }

    if (DBG && P.dbg_tile_time && tid == 0) {
        unsigned long long t_end;
        uint32_t smid;
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t_end));
        asm volatile("mov.u32 %0, %%smid;" : "=r"(smid));
        unsigned long long *o = P.dbg_tile_time + 8 * (size_t)work;
        o[0] = (unsigned long long)tile | ((unsigned long long)n << 32);
        o[1] = t_start; o[2] = t_end; o[3] = smid;
        unsigned long long pk = 0; // first-chunk phase stamps: 16-bit deltas in units of 16 ns
        for (int k = 0; k < 4; k++) {
            unsigned long long d = t_ph[k] > t_start ? (t_ph[k] - t_start) >> 4 : 0;
            pk |= (d > 0xFFFFull ? 0xFFFFull : d) << (16 * k);
        }
        o[4] = pk;
        o[5] = t_ph[4] > t_start ? t_ph[4] - t_start : 0;
        o[6] = 0; o[7] = 0;
    }
    if (P.spread_clears) {
        const uint32_t n_busy = S.bucket_end[ORDER_BUCKETS - 1];
        const uint32_t stride = gridDim.x * (NT / 32), iters = max(1u, (n_busy + gridDim.x - 1) / gridDim.x);
        const uint32_t groups = ((P.tiles_x + CLEAR_GROUP - 1) / CLEAR_GROUP) * (P.ty_end - P.ty_begin);
        uint32_t g = S.clr_cursor[warp];
        clear_empty_tiles<DBG>(P, lane, g, stride, ((groups + stride - 1) / stride + iters) / iters);
        __syncwarp();
        if (lane == 0) S.clr_cursor[warp] = g;
    }
    if (tid == 0) S.ent = ent_next; // publish the next trip's tile (everybody read S.ent right after the barrier at the top)
    } // persistent tile loop

    // ---- the rest of the tiles nothing was binned into ----
    {
        __syncwarp();
        uint32_t g = S.clr_cursor[warp];
        clear_empty_tiles<DBG>(P, lane, g, gridDim.x * (NT / 32), 0xFFFFFFFFu);
    }

    // ---- counters: warp reduce -> per-warp partials -> one striped global RED per counter ----
    {
        c_cov = __reduce_add_sync(0xffffffffu, c_cov);
        c_shaded = __reduce_add_sync(0xffffffffu, c_shaded);
        c_samples = __reduce_add_sync(0xffffffffu, c_samples);
        c_oob = __reduce_add_sync(0xffffffffu, c_oob);
        __syncthreads(); // S.pre is free now
        if (lane == 0) {
            S.pre[warp * 4 + 0] = c_cov; S.pre[warp * 4 + 1] = c_shaded;
            S.pre[warp * 4 + 2] = c_samples; S.pre[warp * 4 + 3] = c_oob;
        }
        __syncthreads();
        if (tid < 4) {
            unsigned long long sum = 0ull;
#pragma unroll
            for (int w = 0; w < NT / 32; w++) sum += S.pre[w * 4 + tid];
            const int slot = tid == 0 ? C_COVERED_PX : (tid == 1 ? C_SHADED_PX : (tid == 2 ? C_SAMPLES : C_TEX_OOB));
            if (sum) atomicAdd(&P.fs->counters[blockIdx.x % CNT_STRIPES][slot], sum);
        }
    }
}

static_assert(RZ_TILE_CTAS * (sizeof(TileSmemT<false>) + 1024) <= 227 * 1024, "tile kernel must fit RZ_TILE_CTAS CTAs per SM");
static_assert(CHUNK <= 255 && POOL <= 65535 && UNIT_CAP <= 65535, "item ids are bytes, fragment and unit ids 16 bits");

} // namespace rz
