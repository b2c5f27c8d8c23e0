// rz_msaa.cuh -- tile stage for the runtime sample counts 1, 2 and 8.
//
// The reference fixes N_MSAA_SAMPLES = 4 (rasterizer/mod.rs:23) with the rotated-grid pattern (mod.rs:109-114); the
// 4-sample frame is what rz_tile.cuh is built around (float4 depths, 4-bit masks, the fragment machinery).  The other
// counts are an extension of the crate's surface, served by this compact kernel: one CTA per 16x16 tile, the tile's
// list sorted by order key, every thread owns one pixel and applies the items in submission order -- the literal
// sequence of Rasterizer::rasterize (mod.rs:443-473) with N_MSAA_SAMPLES = NS: EdgeFunctions::eval + inside per sample,
// RasterizerTriangle::fragment, depth_coverage (strict <), Fragment::interpolate at the pixel centre when all NS samples
// passed else at the first passing sample, the fragment shader, write_pixel; then box_filter_color over NS samples.
// Geometry, clipping, binning and the shaders are shared with the 4-sample path.
#pragma once
#include "rz_exact.cuh"
#include "rz_geom.cuh"
#include "rz_tile.cuh"
#include "rz_types.cuh"

namespace rz {

// Sample patterns: 1 = pixel centre; 2, 8 = the D3D11 standard patterns (sixteenths of a pixel, exact in f32);
// 4 = the reference's RGSS (only here for completeness -- 4-sample frames run rz_tile.cuh).
__device__ __constant__ float MSAA_PAT[8 + 4 + 2 + 1][2] = {
    {0.5625f, 0.3125f}, {0.4375f, 0.6875f}, {0.8125f, 0.5625f}, {0.3125f, 0.1875f},
    {0.1875f, 0.8125f}, {0.0625f, 0.4375f}, {0.6875f, 0.9375f}, {0.9375f, 0.0625f}, // 8 samples at [0..8)
    {0.625f, 0.125f},   {0.875f, 0.625f},   {0.375f, 0.875f},   {0.125f, 0.375f},   // 4 samples at [8..12)
    {0.75f, 0.75f},     {0.25f, 0.25f},                                             // 2 samples at [12..14)
    {0.5f, 0.5f}};                                                                  // 1 sample  at [14]
template <int NS>
__device__ __forceinline__ int pat_base() { return NS == 8 ? 0 : NS == 4 ? 8 : NS == 2 ? 12 : 14; }

constexpr int MSAA_CHUNK = 128; // items staged in shared memory at a time

template <bool DBG>
struct MsaaSmemT {
    unsigned long long sorted[SORT_CAP]; // sort_tile_list_ptr works in here
    BigSetup big[MSAA_CHUNK];
    uint32_t okey[DBG ? MSAA_CHUNK : 1];
    float lut[256];
    uint32_t bucket_end[ORDER_BUCKETS];
    uint32_t clr_cursor[NT / 32];
    uint32_t cur_tile;
};
template <int NS, bool DBG>
__global__ void __launch_bounds__(NT, 2) msaa_tile_kernel(FrameParams P) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    typedef MsaaSmemT<DBG> SM;
    SM &S = *reinterpret_cast<SM *>(smem_raw);
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int lx = tid % TW, ly = tid / TW;
    const int pb = pat_base<NS>();
    const uint32_t FULL = (1u << NS) - 1u; // CoverageMask::all() (mod.rs:42-44)
    uint32_t c_cov = 0, c_shaded = 0, c_samples = 0, c_oob = 0;
    pdl_launch();
    pdl_wait();
    S.lut[tid] = fdiv((float)tid, 255.0f);
    if (tid == 0) {
        uint32_t acc = 0;
#pragma unroll
        for (int b = 0; b < ORDER_BUCKETS; b++) {
            acc += P.fs->bucket_n[b];
            S.bucket_end[b] = acc;
        }
        S.cur_tile = atomicAdd(&P.fs->tile_cursor, 1u);
    }
    if (lane == 0) S.clr_cursor[warp] = blockIdx.x * (NT / 32) + warp;
    for (;;) {
        __syncthreads();
        const uint32_t n_busy = S.bucket_end[ORDER_BUCKETS - 1];
        const uint32_t work = S.cur_tile;
        if (work >= n_busy) break;
        uint32_t tile, bin_off;
        int n;
        {
            uint32_t b = 0, start = 0;
#pragma unroll
            for (int k = 0; k < ORDER_BUCKETS - 1; k++)
                if (work >= S.bucket_end[k]) {
                    b = k + 1;
                    start = S.bucket_end[k];
                }
            const uint4 e = __ldcg(P.busy + (size_t)b * P.tiles_x * P.tiles_y + (work - start));
            tile = e.x;
            n = (int)e.y;
            bin_off = e.z;
        }
        const int tileX0 = (int)(tile % P.tiles_x) * TW, tileY0 = (int)(tile / P.tiles_x) * TH;
        const int X = tileX0 + lx, Y = tileY0 + ly;
        uint4 *bin = P.bins + bin_off;
        // submission order: sort the list by order key (always; a one-item list is sorted)
        bool compact = false;
        if (n > 1) compact = sort_tile_list_ptr(S.sorted, bin, n);
        float d[NS];
        uint32_t col[NS], ok[NS];
#pragma unroll
        for (int k = 0; k < NS; k++) {
            d[k] = CLEAR_DEPTH;
            col[k] = CLEAR_COLOR;
            ok[k] = NO_OWNER;
        }
        for (int pos = 0; pos < n; pos += MSAA_CHUNK) {
            const int cnt = min(MSAA_CHUNK, n - pos);
            __syncthreads(); // the previous chunk has been walked
            if (tid < cnt) {
                uint32_t rec_tie, boxw, key;
                if (compact) {
                    const uint2 e = __ldcg(reinterpret_cast<const uint2 *>(bin) + pos + tid);
                    rec_tie = e.x; boxw = e.y;
                    key = 0;
                } else {
                    const uint4 e = __ldcg(bin + pos + tid);
                    key = e.x; rec_tie = e.y; boxw = e.z;
                }
                const uint32_t rec = rec_tie & ENTRY_REC_MASK;
                const float4 *rr = reinterpret_cast<const float4 *>(&P.recs[rec]);
                const float4 r0 = rr[0], r1 = rr[1], r2 = rr[2];
                BigSetup &b = S.big[tid];
                Setup s;
                s.px[0] = r0.x; s.py[0] = r0.y; s.px[1] = r0.z; s.py[1] = r0.w; s.px[2] = r1.x; s.py[2] = r1.y;
                setup_normals(s);
#pragma unroll
                for (int k = 0; k < 3; k++) {
                    b.px[k] = s.px[k]; b.py[k] = s.py[k]; b.nx[k] = s.nx[k]; b.ny[k] = s.ny[k];
                }
                b.z[0] = r1.z; b.z[1] = r1.w; b.z[2] = r2.x;
                b.inv = r2.y; b.key = key; b.rec = rec;
                b.box = (boxw & 15u) | (((boxw >> 4) & 15u) << 8) | ((((boxw >> 8) & 15u) + 1u) << 16) | ((((boxw >> 12) & 15u) + 1u) << 24);
                b.tie = 0;
                if (DBG) S.okey[tid] = __float_as_uint(r2.z);
            }
            __syncthreads();
            for (int it = 0; it < cnt; it++) {
                const BigSetup &B = S.big[it];
                const uint32_t box = B.box;
                const uint32_t rx = (uint32_t)lx - (box & 0xFF), ry = (uint32_t)ly - ((box >> 8) & 0xFF);
                if (rx >= ((box >> 16) & 0xFF) || ry >= (box >> 24)) continue;
                Setup q;
                big_to_setup(B, q);
                // EdgeFunctions::eval (mod.rs:134-146) over the NS sample positions; the edge values feed the depths
                uint32_t m = 0;
                float zs[NS];
#pragma unroll
                for (int i = 0; i < NS; i++) {
                    const float xs = fadd((float)X, MSAA_PAT[pb + i][0]), ys = fadd((float)Y, MSAA_PAT[pb + i][1]);
                    const float e0 = edge_eval(q, 0, xs, ys), e1 = edge_eval(q, 1, xs, ys), e2 = edge_eval(q, 2, xs, ys);
                    const bool in = edge_pass(e0, q.nx[0], q.ny[0]) && edge_pass(e1, q.nx[1], q.ny[1]) && edge_pass(e2, q.nx[2], q.ny[2]);
                    m |= (in ? 1u : 0u) << i;
                    // RasterizerTriangle::fragment (mod.rs:225-253); 0.0 where uncovered
                    const float b0 = clamp01(fmul(e1, q.inv));
                    const float b1 = clamp01(fmul(e2, q.inv));
                    const float b2 = clamp01(fsub(fsub(1.0f, b0), b1));
                    const float z = fadd(fadd(fmul(b0, q.z[0]), fmul(b1, q.z[1])), fmul(b2, q.z[2]));
                    zs[i] = in ? z : 0.0f;
                }
                if (!m) continue;
                c_cov++;
                uint32_t mp = 0; // depth_coverage (mod.rs:363-378): strict <
#pragma unroll
                for (int i = 0; i < NS; i++)
                    if (((m >> i) & 1u) && zs[i] < d[i]) mp |= 1u << i;
                if (!mp) continue;
                c_shaded++;
                c_samples += __popc(mp);
                // Fragment::interpolate (mod.rs:70-83): the centre if every sample passed, else the first passing sample
                float xs = fadd((float)X, 0.5f), ys = fadd((float)Y, 0.5f);
                if (mp != FULL) {
                    const int i = __ffs(mp) - 1;
                    xs = fadd((float)X, MSAA_PAT[pb + i][0]);
                    ys = fadd((float)Y, MSAA_PAT[pb + i][1]);
                }
                const uint32_t argb = shade_at<true, true>(P, q, B.rec, S.lut, xs, ys, [&]() { return zs[0]; }, c_oob);
#pragma unroll
                for (int i = 0; i < NS; i++)
                    if ((mp >> i) & 1u) { // write_pixel (mod.rs:380-397)
                        d[i] = zs[i];
                        col[i] = argb;
                        if (DBG) ok[i] = DBG ? S.okey[it] : 0u;
                    }
            }
        }
        __syncthreads();
        if (tid == 0) S.cur_tile = atomicAdd(&P.fs->tile_cursor, 1u);
        // ColorBuffer::box_filter_color (buffers.rs:111-125) over NS samples
        uint32_t r = 0, g = 0, b = 0;
#pragma unroll
        for (int i = 0; i < NS; i++) {
            r += (col[i] >> 16) & 0xFF;
            g += (col[i] >> 8) & 0xFF;
            b += col[i] & 0xFF;
        }
        if (X < (int)P.W && Y < (int)P.H) {
            P.out[(size_t)Y * P.W + X] = 0xFF000000u | ((r / NS) << 16) | ((g / NS) << 8) | (b / NS);
            if (DBG) {
                const size_t o = ((size_t)Y * P.W + X) * NS;
#pragma unroll
                for (int i = 0; i < NS; i++) {
                    if (P.dbg_depth) P.dbg_depth[o + i] = d[i];
                    if (P.dbg_color) P.dbg_color[o + i] = col[i];
                    if (P.dbg_owner) P.dbg_owner[o + i] = ok[i];
                }
            }
        }
    }
    // tiles nothing was binned into: the clear colour (and, for the parity instrumentation, the cleared samples)
    {
        const uint32_t shard_tiles = P.tiles_x * (P.ty_end - P.ty_begin);
        for (uint32_t t = blockIdx.x; t < shard_tiles; t += gridDim.x) {
            const uint32_t tile = P.ty_begin * P.tiles_x + t;
            if (!owns_tile_row(P, tile / P.tiles_x) || P.tile_count[tile] != 0u) continue;
            const int X = (int)(tile % P.tiles_x) * TW + lx, Y = (int)(tile / P.tiles_x) * TH + ly;
            if (X < (int)P.W && Y < (int)P.H) {
                P.out[(size_t)Y * P.W + X] = CLEAR_COLOR;
                if (DBG) {
                    const size_t o = ((size_t)Y * P.W + X) * NS;
                    for (int i = 0; i < NS; i++) {
                        if (P.dbg_depth) P.dbg_depth[o + i] = CLEAR_DEPTH;
                        if (P.dbg_color) P.dbg_color[o + i] = CLEAR_COLOR;
                        if (P.dbg_owner) P.dbg_owner[o + i] = NO_OWNER;
                    }
                }
            }
        }
    }
    // counters
    c_cov = __reduce_add_sync(0xffffffffu, c_cov);
    c_shaded = __reduce_add_sync(0xffffffffu, c_shaded);
    c_samples = __reduce_add_sync(0xffffffffu, c_samples);
    c_oob = __reduce_add_sync(0xffffffffu, c_oob);
    if (lane == 0) {
        const int st = blockIdx.x % CNT_STRIPES;
        if (c_cov) atomicAdd(&P.fs->counters[st][C_COVERED_PX], (unsigned long long)c_cov);
        if (c_shaded) atomicAdd(&P.fs->counters[st][C_SHADED_PX], (unsigned long long)c_shaded);
        if (c_samples) atomicAdd(&P.fs->counters[st][C_SAMPLES], (unsigned long long)c_samples);
        if (c_oob) atomicAdd(&P.fs->counters[st][C_TEX_OOB], (unsigned long long)c_oob);
    }
}

} // namespace rz
