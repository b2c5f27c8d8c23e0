// rz_tile.cuh -- stage 3+4: per-tile raster, MSAA depth test, shading, resolve.
//
// One CTA of 256 threads owns one 16x16 pixel tile.  The tile's 4-sample depth and colour state
// (the reference's DepthBuffer / ColorBuffer, rasterizer/buffers.rs:83-157) lives in shared
// memory from clear to resolve and only the box-filtered u32 image goes back to HBM.
//
// Submission order is honoured exactly:
//   * the tile's list is sorted by order key (8 * triangle number + fan index);
//   * runs of LARGE items (in-tile bbox > 32 px) are walked pixel-parallel: a thread owns a pixel
//     for the whole run, so it applies the triangles in order by construction;
//   * runs of SMALL items are processed triangle-parallel in chunks of <= 255 items:
//       phase 1  thread = item : exact coverage over its bbox, sample depths -> smem records,
//                                sets bit <item> in the bitmask of every pixel it covers;
//       phase 2  thread = pixel: walks the set bits in ascending (= submission) order and replays
//                                the reference's strict-< depth test, recording each fragment's
//                                post-depth mask and the final per-sample owner;
//       phase 3  thread = item : shades its fragments at the position the post-depth mask selects
//                                (rasterizer/mod.rs:70-83) -- only those that still own a sample at
//                                the end of the chunk can be seen, the others are skipped.
#pragma once
#include "rz_exact.cuh"
#include "rz_geom.cuh"
#include "rz_types.cuh"

namespace rz {

// Texture::read_texel + Color::from_rgba (texture.rs:47-63, color.rs:22-29).  Reads past the
// buffer (a panic in the reference) are clamped to the last byte and counted.
__device__ __forceinline__ void read_texel(const TexInfo &t, uint32_t x, uint32_t y, float *rgba, uint32_t &oob) {
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
    rgba[0] = fdiv((float)b0, 255.0f);
    rgba[1] = fdiv((float)b1, 255.0f);
    rgba[2] = fdiv((float)b2, 255.0f);
    rgba[3] = fdiv((float)b3, 255.0f);
}

// Texture::sample (texture.rs:65-83) + Color::to_argb
__device__ __forceinline__ uint32_t sample_texture_argb(const TexInfo &t, float u, float v, uint32_t &oob) {
    const float x = fmul(u, (float)(t.w - 1)), y = fmul(v, (float)(t.h - 1));
    const uint32_t x0 = sat_u32(floorf(x)), x1 = sat_u32(ceilf(x));
    const uint32_t y0 = sat_u32(floorf(y)), y1 = sat_u32(ceilf(y));
    float tl[4], tr[4], bl[4], br[4];
    read_texel(t, x0, y0, tl, oob);
    read_texel(t, x1, y0, tr, oob);
    read_texel(t, x0, y1, bl, oob);
    read_texel(t, x1, y1, br, oob);
    const float xf = fsub(x, truncf(x)), yf = fsub(y, truncf(y)); // f32::fract
    const float omx = fsub(1.0f, xf), omy = fsub(1.0f, yf);
    float o[4];
#pragma unroll
    for (int k = 0; k < 4; k++) {
        const float r0 = fadd(fmul(tl[k], omx), fmul(tr[k], xf));
        const float r1 = fadd(fmul(bl[k], omx), fmul(br[k], xf));
        o[k] = fadd(fmul(r0, omy), fmul(r1, yf));
    }
    return to_argb(o[0], o[1], o[2], o[3]);
}

// Fragment::interpolate (rasterizer/mod.rs:69-100) + the built-in fragment shaders
// (main.rs:67-77) + Color::to_argb.  mpost is the POST-depth-test mask, depth0 the pre-test
// sampled depth of sample 0 (0.0 when uncovered), as FragCoords.depths[0] (mod.rs:458-463).
__device__ __forceinline__ uint32_t shade(const Setup &s, const AttrRec *ar, uint32_t fs, const TexInfo &tex, int X,
                                          int Y, uint32_t mpost, float depth0, uint32_t &oob) {
    if (fs == 2u) return to_argb(depth0, depth0, depth0, 1.0f); // Color::grayscale(depths[0])
    float xs, ys;
    if (mpost == 0xFu) {
        xs = fadd((float)X, 0.5f);
        ys = fadd((float)Y, 0.5f);
    } else {
        const int i = __ffs(mpost) - 1;
        xs = fadd((float)X, rgss_x(i));
        ys = fadd((float)Y, rgss_y(i));
    }
    const float e0 = edge_eval(s, 0, xs, ys), e1 = edge_eval(s, 1, xs, ys), e2 = edge_eval(s, 2, xs, ys);
    const float fu = fdiv(e1, s.w[0]), fv = fdiv(e2, s.w[1]), fw = fdiv(e0, s.w[2]);
    const float sum = fadd(fadd(fu, fv), fw);
    const float u = clamp01(fdiv(fu, sum));
    const float v = clamp01(fdiv(fv, sum));
    const float w = clamp01(fsub(fsub(1.0f, u), v));
    const float *a = ar->a; // vertex k, component c at a[6k + c]
#define RZ_INTERP(c) fadd(fadd(fmul(__ldg(a + (c)), u), fmul(__ldg(a + 6 + (c)), v)), fmul(__ldg(a + 12 + (c)), w))
    if (fs == 1u) return to_argb(RZ_INTERP(0), RZ_INTERP(1), RZ_INTERP(2), RZ_INTERP(3));
    const float tu = RZ_INTERP(4), tv = RZ_INTERP(5);
#undef RZ_INTERP
    return sample_texture_argb(tex, tu, tv, oob);
}

// Bitonic network in its "flip then halve" form: every compare-exchange puts the smaller key at
// the lower index, so n need not be a power of two (the missing tail behaves like +inf padding).
template <typename T>
__device__ __forceinline__ void block_sort(T *a, int n) {
    for (int k = 2; (k >> 1) < n; k <<= 1) {
        for (int i = threadIdx.x; i < n; i += NT) {
            const int p = i ^ (k - 1);
            if (p > i && p < n) {
                T x = a[i], y = a[p];
                if (x > y) { a[i] = y; a[p] = x; }
            }
        }
        __syncthreads();
        for (int j = k >> 2; j > 0; j >>= 1) {
            for (int i = threadIdx.x; i < n; i += NT) {
                const int p = i ^ j;
                if (p > i && p < n) {
                    T x = a[i], y = a[p];
                    if (x > y) { a[i] = y; a[p] = x; }
                }
            }
            __syncthreads();
        }
    }
}

// A large item staged in shared memory for the pixel-parallel walk (24 words).
struct __align__(16) BigSetup {
    float px[3], py[3], nx[3], ny[3], z[3], w[3];
    float inv;
    uint32_t key, fs, rec;
    uint32_t box; // lx0 | ly0 << 8 | bw << 16 | bh << 24   (tile-local)
    uint32_t pad;
};
static_assert(sizeof(BigSetup) == 96, "BigSetup must be 24 words");

struct TileSmem {
    float depth[TILE_PX * 4];
    uint32_t color[TILE_PX * 4];
    uint32_t okey[TILE_PX * 4];              // owner keys (parity instrumentation only)
    unsigned long long sorted[SORT_CAP];
    uint32_t pixmask[(CHUNK + 1) / 32][TILE_PX];
    float4 pool[POOL];                       // fragment records: 4 sample depths; aliased by BigSetup[]
    uint8_t mpost[POOL];
    uint32_t owner[TILE_PX];                 // 4 x u8 item id per pixel; reused as resolve staging
    uint32_t pxm[NT];                        // per item: which bbox pixels are covered
    uint32_t meta[NT];                       // per item: base | lx0 << 12 | ly0 << 16 | bw << 20
    uint32_t scan[NT / 32];
    uint32_t first_big, first_small, cut;
    unsigned long long cnt[4];
};
static_assert(sizeof(BigSetup) * CHUNK <= sizeof(float4) * POOL, "BigSetup run must fit in the pool");

template <bool DBG>
__global__ void __launch_bounds__(NT) tile_kernel(FrameParams P) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    TileSmem &S = *reinterpret_cast<TileSmem *>(smem_raw);

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const uint32_t tx = blockIdx.x % P.tiles_x, ty = P.ty_begin + blockIdx.x / P.tiles_x;
    const uint32_t tile = ty * P.tiles_x + tx;
    const int tileX0 = tx * TW, tileY0 = ty * TH;
    const int lx = tid % TW, ly = tid / TW;
    const int X = tileX0 + lx, Y = tileY0 + ly;

    // clear (the state resolve_and_clear leaves behind, rasterizer/mod.rs:497-506)
#pragma unroll
    for (int k = 0; k < 4; k++) {
        S.depth[tid * 4 + k] = CLEAR_DEPTH;
        S.color[tid * 4 + k] = CLEAR_COLOR;
        if (DBG) S.okey[tid * 4 + k] = NO_OWNER;
    }
#pragma unroll
    for (int k = 0; k < (CHUNK + 1) / 32; k++) S.pixmask[k][tid] = 0u;
    if (tid < 4) S.cnt[tid] = 0ull;

    uint32_t c_cov = 0, c_shaded = 0, c_samples = 0, c_oob = 0;

    const uint32_t n_total = P.tile_count[tile];
    const int n = (int)min(n_total, P.bin_cap);
    if (n > 0) {
        unsigned long long *bin = P.bins + (size_t)tile * P.bin_cap;
        const unsigned long long *list;
        if (n <= SORT_CAP) {
            for (int i = tid; i < n; i += NT) S.sorted[i] = bin[i];
            __syncthreads();
            block_sort(S.sorted, n);
            list = S.sorted;
        } else {
            __syncthreads();
            block_sort(bin, n); // rare: in place in global memory
            list = bin;
        }
        __syncthreads();

        for (int pos = 0; pos < n;) {
            // ---- load this window's items (thread = item) ----
            const int item = pos + tid;
            const bool valid = tid < CHUNK && item < n;
            Setup s;
            uint32_t rec = 0, key = 0, fs = 0;
            int bx0 = 0, by0 = 0, bw = 0, bh = 0; // in-tile bbox, tile-local origin
            bool big = false;
            if (valid) {
                rec = (uint32_t)list[item];
                load_setup(P.recs, rec, s, key, fs);
                BBox b = pixel_bbox(s, P.W, P.H);
                const int x0 = max((int)b.x0, tileX0), x1 = min((int)b.x1, tileX0 + TW);
                const int y0 = max((int)b.y0, tileY0), y1 = min((int)b.y1, tileY0 + TH);
                if (x0 < x1 && y0 < y1) {
                    bx0 = x0 - tileX0; by0 = y0 - tileY0; bw = x1 - x0; bh = y1 - y0;
                }
                big = bw * bh > SMALL_PX;
            }
            if (tid == 0) {
                S.first_big = CHUNK + 1;
                S.first_small = CHUNK + 1;
                S.cut = CHUNK + 1;
            }
            __syncthreads();
            if (valid && big) atomicMin(&S.first_big, (uint32_t)tid);
            if (valid && !big) atomicMin(&S.first_small, (uint32_t)tid);
            __syncthreads();
            const int nvalid = min(CHUNK, n - pos);

            if (S.first_big == 0) {
                // ================= run of large items: pixel-parallel =================
                const int run = min((int)S.first_small, nvalid);
                BigSetup *B = reinterpret_cast<BigSetup *>(S.pool);
                if (tid < run) {
                    BigSetup &b = B[tid];
#pragma unroll
                    for (int k = 0; k < 3; k++) {
                        b.px[k] = s.px[k]; b.py[k] = s.py[k]; b.nx[k] = s.nx[k]; b.ny[k] = s.ny[k];
                        b.z[k] = s.z[k]; b.w[k] = s.w[k];
                    }
                    b.inv = s.inv; b.key = key; b.fs = fs; b.rec = rec;
                    b.box = (uint32_t)bx0 | ((uint32_t)by0 << 8) | ((uint32_t)bw << 16) | ((uint32_t)bh << 24);
                }
                __syncthreads();
                float d[4];
                uint32_t col[4], ok[4];
#pragma unroll
                for (int k = 0; k < 4; k++) {
                    d[k] = S.depth[tid * 4 + k];
                    col[k] = S.color[tid * 4 + k];
                    if (DBG) ok[k] = S.okey[tid * 4 + k];
                }
                for (int it = 0; it < run; it++) {
                    const uint32_t box = B[it].box;
                    const uint32_t rx = (uint32_t)lx - (box & 0xFF), ry = (uint32_t)ly - ((box >> 8) & 0xFF);
                    if (rx >= ((box >> 16) & 0xFF) || ry >= (box >> 24)) continue;
                    Setup q;
#pragma unroll
                    for (int k = 0; k < 3; k++) {
                        q.px[k] = B[it].px[k]; q.py[k] = B[it].py[k]; q.nx[k] = B[it].nx[k]; q.ny[k] = B[it].ny[k];
                        q.z[k] = B[it].z[k]; q.w[k] = B[it].w[k];
                    }
                    q.inv = B[it].inv;
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
                    const uint32_t argb = shade(q, &P.attrs[B[it].rec], B[it].fs, P.tex0, X, Y, mp, zs[0], c_oob);
#pragma unroll
                    for (int k = 0; k < 4; k++)
                        if ((mp >> k) & 1u) {
                            d[k] = zs[k];
                            col[k] = argb;
                            if (DBG) ok[k] = B[it].key;
                        }
                }
#pragma unroll
                for (int k = 0; k < 4; k++) {
                    S.depth[tid * 4 + k] = d[k];
                    S.color[tid * 4 + k] = col[k];
                    if (DBG) S.okey[tid * 4 + k] = ok[k];
                }
                __syncthreads();
                pos += run;
                continue;
            }

            // ================= chunk of small items: triangle-parallel =================
            int cnt = min((int)S.first_big, nvalid);
            const bool mine = tid < cnt;
            // ---- phase 1a: exact coverage of every bbox pixel (<= 32) ----
            unsigned long long cov_lo = 0ull, cov_hi = 0ull;
            uint32_t pxm = 0, ncov = 0;
            if (mine) {
                int j = 0;
                for (int ry = 0; ry < bh; ry++)
                    for (int rx = 0; rx < bw; rx++, j++) {
                        const uint32_t m = coverage_mask(s, tileX0 + bx0 + rx, tileY0 + by0 + ry);
                        if (m) {
                            if (j < 16) cov_lo |= (unsigned long long)m << (4 * j);
                            else cov_hi |= (unsigned long long)m << (4 * (j - 16));
                            pxm |= 1u << j;
                            ncov++;
                        }
                    }
            }
            // block exclusive scan of ncov -> record base
            uint32_t incl = ncov;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const uint32_t v = __shfl_up_sync(0xffffffffu, incl, o);
                if (lane >= o) incl += v;
            }
            if (lane == 31) S.scan[warp] = incl;
            __syncthreads();
            uint32_t wbase = 0;
#pragma unroll
            for (int k = 0; k < NT / 32; k++)
                if (k < warp) wbase += S.scan[k];
            const uint32_t base = wbase + incl - ncov;
            if (mine && base + ncov > POOL) atomicMin(&S.cut, (uint32_t)tid); // defer the rest to the next chunk
            __syncthreads();
            cnt = min(cnt, (int)S.cut);
            const bool act = tid < cnt;
            // ---- phase 1b: sample depths -> records; publish per-pixel item bits ----
            if (act) {
                S.pxm[tid] = pxm;
                S.meta[tid] = base | ((uint32_t)bx0 << 12) | ((uint32_t)by0 << 16) | ((uint32_t)bw << 20);
                c_cov += ncov;
                uint32_t bits = pxm, r = base;
                while (bits) {
                    const int j = __ffs(bits) - 1;
                    bits &= bits - 1;
                    const int ry = j / bw, rx = j - ry * bw;
                    const uint32_t m = (uint32_t)((j < 16 ? cov_lo >> (4 * j) : cov_hi >> (4 * (j - 16))) & 0xFull);
                    const int PX = tileX0 + bx0 + rx, PY = tileY0 + by0 + ry;
                    float4 z;
                    z.x = (m & 1u) ? sample_depth(s, PX, PY, 0) : 0.0f;
                    z.y = (m & 2u) ? sample_depth(s, PX, PY, 1) : 0.0f;
                    z.z = (m & 4u) ? sample_depth(s, PX, PY, 2) : 0.0f;
                    z.w = (m & 8u) ? sample_depth(s, PX, PY, 3) : 0.0f;
                    S.pool[r] = z;
                    S.mpost[r] = (uint8_t)m; // pre-depth mask for phase 2, replaced by the post-depth mask
                    const int p = (by0 + ry) * TW + bx0 + rx;
                    atomicOr(&S.pixmask[tid >> 5][p], 1u << (tid & 31));
                    r++;
                }
            }
            __syncthreads();
            // ---- phase 2: thread = pixel, ordered depth resolve ----
            {
                float d0 = S.depth[tid * 4], d1 = S.depth[tid * 4 + 1], d2 = S.depth[tid * 4 + 2],
                      d3 = S.depth[tid * 4 + 3];
                uint32_t own = 0xFFFFFFFFu;
                const int nwords = (cnt + 31) >> 5;
                for (int wd = 0; wd < nwords; wd++) {
                    uint32_t bits = S.pixmask[wd][tid];
                    if (!bits) continue;
                    S.pixmask[wd][tid] = 0u;
                    while (bits) {
                        const int it = (wd << 5) + __ffs(bits) - 1;
                        bits &= bits - 1;
                        const uint32_t meta = S.meta[it];
                        const int ibw = (int)(meta >> 20);
                        const int j = (ly - (int)((meta >> 16) & 0xF)) * ibw + (lx - (int)((meta >> 12) & 0xF));
                        const uint32_t r = (meta & 0xFFFu) + __popc(S.pxm[it] & ((1u << j) - 1u));
                        const uint32_t m = S.mpost[r];
                        const float4 z = S.pool[r];
                        uint32_t mp = 0;
                        if ((m & 1u) && z.x < d0) { mp |= 1u; d0 = z.x; own = (own & 0xFFFFFF00u) | (uint32_t)it; }
                        if ((m & 2u) && z.y < d1) { mp |= 2u; d1 = z.y; own = (own & 0xFFFF00FFu) | ((uint32_t)it << 8); }
                        if ((m & 4u) && z.z < d2) { mp |= 4u; d2 = z.z; own = (own & 0xFF00FFFFu) | ((uint32_t)it << 16); }
                        if ((m & 8u) && z.w < d3) { mp |= 8u; d3 = z.w; own = (own & 0x00FFFFFFu) | ((uint32_t)it << 24); }
                        S.mpost[r] = (uint8_t)mp;
                    }
                }
                S.depth[tid * 4] = d0; S.depth[tid * 4 + 1] = d1; S.depth[tid * 4 + 2] = d2; S.depth[tid * 4 + 3] = d3;
                S.owner[tid] = own;
            }
            __syncthreads();
            // ---- phase 3: thread = item, shade the fragments that are still visible ----
            if (act) {
                uint32_t bits = pxm, r = base;
                while (bits) {
                    const int j = __ffs(bits) - 1;
                    bits &= bits - 1;
                    const uint32_t mp = S.mpost[r];
                    const float depth0 = S.pool[r].x;
                    r++;
                    if (!mp) continue;
                    c_shaded++;
                    c_samples += __popc(mp);
                    const int ry = j / bw, rx = j - ry * bw;
                    const int p = (by0 + ry) * TW + bx0 + rx;
                    const uint32_t own = S.owner[p];
                    uint32_t f = 0;
#pragma unroll
                    for (int k = 0; k < 4; k++)
                        if (((mp >> k) & 1u) && ((own >> (8 * k)) & 0xFFu) == (uint32_t)tid) f |= 1u << k;
                    if (!f) continue; // overwritten later in this chunk: its colour can never be seen
                    const uint32_t argb =
                        shade(s, &P.attrs[rec], fs, P.tex0, tileX0 + bx0 + rx, tileY0 + by0 + ry, mp, depth0, c_oob);
#pragma unroll
                    for (int k = 0; k < 4; k++)
                        if ((f >> k) & 1u) {
                            S.color[p * 4 + k] = argb;
                            if (DBG) S.okey[p * 4 + k] = key;
                        }
                }
            }
            __syncthreads();
            pos += cnt;
        }
    }
    __syncthreads();

    // ---- resolve (ColorBuffer::box_filter_color, buffers.rs:111-125) and write back ----
    const uint32_t res = box_filter(S.color[tid * 4], S.color[tid * 4 + 1], S.color[tid * 4 + 2], S.color[tid * 4 + 3]);
    if (DBG && X < (int)P.W && Y < (int)P.H) {
        const size_t o = ((size_t)Y * P.W + X) * 4;
#pragma unroll
        for (int k = 0; k < 4; k++) {
            if (P.dbg_depth) P.dbg_depth[o + k] = S.depth[tid * 4 + k];
            if (P.dbg_color) P.dbg_color[o + k] = S.color[tid * 4 + k];
            if (P.dbg_owner) P.dbg_owner[o + k] = S.okey[tid * 4 + k];
        }
    }
    if ((P.W & 3u) == 0u) {
        S.owner[tid] = res; // stage, then 64 threads issue 128-bit row stores
        __syncthreads();
        if (tid < TH * (TW / 4)) {
            const int row = tid / (TW / 4), q = tid % (TW / 4);
            const int Yr = tileY0 + row, Xq = tileX0 + q * 4;
            if (Yr < (int)P.H && Xq < (int)P.W)
                *reinterpret_cast<uint4 *>(&P.out[(size_t)Yr * P.W + Xq]) =
                    *reinterpret_cast<const uint4 *>(&S.owner[row * TW + q * 4]);
        }
    } else if (X < (int)P.W && Y < (int)P.H) {
        P.out[(size_t)Y * P.W + X] = res;
    }

    // ---- counters ----
    c_cov = __reduce_add_sync(0xffffffffu, c_cov);
    c_shaded = __reduce_add_sync(0xffffffffu, c_shaded);
    c_samples = __reduce_add_sync(0xffffffffu, c_samples);
    c_oob = __reduce_add_sync(0xffffffffu, c_oob);
    if (lane == 0) {
        if (c_cov) atomicAdd(&S.cnt[0], (unsigned long long)c_cov);
        if (c_shaded) atomicAdd(&S.cnt[1], (unsigned long long)c_shaded);
        if (c_samples) atomicAdd(&S.cnt[2], (unsigned long long)c_samples);
        if (c_oob) atomicAdd(&S.cnt[3], (unsigned long long)c_oob);
    }
    __syncthreads();
    if (tid == 0) {
        if (S.cnt[0]) atomicAdd(&P.fs->counters[C_COVERED_PX], S.cnt[0]);
        if (S.cnt[1]) atomicAdd(&P.fs->counters[C_SHADED_PX], S.cnt[1]);
        if (S.cnt[2]) atomicAdd(&P.fs->counters[C_SAMPLES], S.cnt[2]);
        if (S.cnt[3]) atomicAdd(&P.fs->counters[C_TEX_OOB], S.cnt[3]);
    }
}

} // namespace rz
