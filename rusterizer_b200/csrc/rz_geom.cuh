// rz_geom.cuh -- stage 1+2: vertex transform, clip, triangle reconstruction, setup, binning.
//
// One thread per input triangle (Renderer::render + the front half of Rasterizer::rasterize,
// render.rs:98-114, rasterizer/mod.rs:425-441).  Surviving triangles are compacted into the
// record arrays with a warp-aggregated atomic (ballot + popc prefix), tagged with their
// submission-order key, and binned into 16x16 screen tiles.  Small triangles are rasterised
// exactly right here so that triangles covering no sample never reach a tile list; large ones
// are queued for the cooperative binning kernel below.
#pragma once
#include "rz_exact.cuh"
#include "rz_types.cuh"

namespace rz {

__device__ __forceinline__ unsigned lanemask_lt() {
    unsigned m;
    asm("mov.u32 %0, %%lanemask_lt;" : "=r"(m));
    return m;
}

// Warp-aggregated "allocate one slot" (stream compaction by ballot/prefix): the converged lanes
// elect a leader that bumps the global cursor once; every lane gets base + its rank.
__device__ __forceinline__ uint32_t alloc_slot(uint32_t *cursor) {
    unsigned m = __activemask();
    int leader = __ffs(m) - 1;
    uint32_t base = 0;
    if ((int)(threadIdx.x & 31) == leader) base = atomicAdd(cursor, (uint32_t)__popc(m));
    base = __shfl_sync(m, base, leader);
    return base + __popc(m & lanemask_lt());
}

// Append an entry to a tile's bin.  Neighbouring triangles land in the same tile, so the lanes that are converged
// here are first grouped by tile (match.any): one atomic per group reserves the slots, instead of up to 32
// same-address atomics serialising in L2.  `rec_tie` = record index | tie-break bits << 29, `box` = in-tile pixel box
// (see the entry layout in rz_types.cuh).
__device__ __forceinline__ void push_bin(const FrameParams &P, uint32_t tile, uint32_t key, uint32_t rec_tie, uint32_t box) {
    const uint2 tb = __ldg(reinterpret_cast<const uint2 *>(&P.tile_bin[tile])); // planned on the host before the frame;
                                                                                // in flight together with the atomic below
    const unsigned peers = __match_any_sync(__activemask(), tile);
    const int leader = __ffs(peers) - 1;
    uint32_t base = 0;
    if ((int)(threadIdx.x & 31) == leader) {
        base = atomicAdd(&P.tile_count[tile], (uint32_t)__popc(peers));
    }
    base = __shfl_sync(peers, base, leader);
    const uint32_t slot = base + __popc(peers & lanemask_lt());
    if (slot < tb.y)
        P.bins[(size_t)tb.x + slot] = make_uint4(key, rec_tie, box, 0u);
    else
        atomicOr(&P.fs->err, ERR_BIN_OVF);
}
// in-tile pixel box of the pixel rectangle [x0,x1) x [y0,y1) (non-empty inside tile (tx, ty))
__device__ __forceinline__ uint32_t tile_box(uint32_t x0, uint32_t x1, uint32_t y0, uint32_t y1, uint32_t tx, uint32_t ty) {
    const uint32_t X0 = max(x0, tx * TW), X1 = min(x1, tx * TW + TW), Y0 = max(y0, ty * TH), Y1 = min(y1, ty * TH + TH);
    return (X0 - tx * TW) | ((Y0 - ty * TH) << 4) | ((X1 - X0 - 1u) << 8) | ((Y1 - Y0 - 1u) << 12);
}
// tie-break bits of EdgeFunctions::inside (mod.rs:160-168): bit k <=> E_k == 0 counts as inside
__device__ __forceinline__ uint32_t tie_bits(const Setup &s) {
    uint32_t t = 0;
#pragma unroll
    for (int k = 0; k < 3; k++) t |= ((s.nx[k] > 0.0f || (!(s.nx[k] < 0.0f) && s.ny[k] < 0.0f)) ? 1u : 0u) << k;
    return t;
}

struct GeomLocal {
    uint32_t c[C_COUNT];
    unsigned long long bbox;
};

// Rasterizer::perspective_divide + viewport_transform for one vertex (rasterizer/mod.rs:284-331):
// clip (x,y,z,w) -> screen (x, y, depth) ; w is kept as depths_camera_space.
__device__ __forceinline__ float4 project_vertex(const float *c, float Wf, float Hf) {
    const float w = c[3];
    const float nx = fdiv(c[0], w), ny = fdiv(c[1], w), nz = fdiv(c[2], w);
    float4 r;
    // `/ 2.0` is computed as `* 0.5`: scaling by a power of two rounds identically
    r.x = fmul(fmul(Wf, fadd(nx, 1.0f)), 0.5f);
    r.y = fmul(Hf, fsub(1.0f, fmul(fadd(ny, 1.0f), 0.5f)));
    r.z = fadd(fmul(fmul(fadd(nz, 1.0f), 0.5f), 1.0f), 0.0f); // (z+1)*0.5*(zmax-zmin)+zmin
    r.w = w;
    return r;
}

// RasterizerTriangle::new + bounding_box (rasterizer/mod.rs:187-222,347-361) for one screen-space
// triangle (s.px/py/z/w filled), then cull / record / bin.
// Unclipped triangles pass their vertex indices (ca == nullptr); clipped ones pass the three
// interpolated VertexAttributes ca[0..2] (6 floats each, local memory).
__device__ __forceinline__ void emit_setup(const FrameParams &P, const DrawParams &D, Setup &s, uint32_t i0, uint32_t i1,
                                           uint32_t i2, const float (*ca)[6], uint32_t key, GeomLocal &lc) {
    setup_normals(s);
    lc.c[C_TRIS_SETUP]++;

    BBox b = pixel_bbox(s, P.scissor);
    b.y0 = max(b.y0, P.row_begin); // screen-space shard owned by this ctx
    b.y1 = min(b.y1, P.row_end);
    if (b.x0 >= b.x1 || b.y0 >= b.y1) return;
    const uint32_t bw = b.x1 - b.x0, bh = b.y1 - b.y0;
    {
        const uint32_t rows = owned_rows(P, b.y0, b.y1);
        if (rows == 0u) return; // touches none of this ctx's interleaved bands
        lc.bbox += (unsigned long long)bw * rows;
    }

    // ---- exact culls (results identical to walking the bbox) ----
    // (1) Back-facing with a rigorous margin.  The three edge functions of any sample sum to the
    // signed 2x area A; a covered sample has all three computed values >= 0.  With u = 2^-24 every
    // computed edge value is within 4u*(|dy||sx-px| + |dx||sy-py|) of its exact value and the computed
    // area within 8u*Bw*Bh of A, where Bw,Bh are the bbox extents and |s-p| <= B+1 inside the pixel
    // bbox.  Hence coverage implies  area2 >= -u*(32*Bw*Bh + 12*(Bw+Bh)); we cull only below twice
    // that bound (DESIGN.md "Exact culls").  NaN/inf compare false and fall through to the exact path.
    const float area2 = cross2(fsub(s.px[1], s.px[0]), fsub(s.py[1], s.py[0]), fsub(s.px[2], s.px[0]),
                               fsub(s.py[2], s.py[0]));
    {
        const float mnx = fminf(fminf(s.px[0], s.px[1]), s.px[2]), mxx = fmaxf(fmaxf(s.px[0], s.px[1]), s.px[2]);
        const float mny = fminf(fminf(s.py[0], s.py[1]), s.py[2]), mxy = fmaxf(fmaxf(s.py[0], s.py[1]), s.py[2]);
        const float Bw = mxx - mnx, Bh = mxy - mny;
        const float tau = 1.1920929e-07f * (34.0f * Bw * Bh + 14.0f * (Bw + Bh)) + 1e-30f;
        if (area2 < -tau) return;
    }
    const bool small = bw <= GEOM_SMALL_DIM && bh <= GEOM_SMALL_DIM;
    const uint32_t tx0 = b.x0 / TW, ty0 = b.y0 / TH;
    uint32_t tmask = 0;
    if (small) {
        const uint32_t tx1 = (b.x1 - 1) / TW, ty1 = (b.y1 - 1) / TH;
        if (P.msaa == 4u && area2 < GEOM_THIN_AREA2 && bw * bh <= GEOM_THIN_PX && setup_is_tame(s)) {
            // (2) thin / tiny triangles usually touch no sample at all: rasterise them exactly right
            // here so they never reach a tile list (pole slivers of a UV-sphere, distant meshes).  Finite
            // coordinates only (anything else is simply binned): the single-compare form of EdgeFunctions::inside
            // that the tile stage uses (coverage_mask_fast) is exact under that precondition.
            float thr[3];
            edge_thresholds(s, thr);
            for (uint32_t Y = b.y0; Y < b.y1; Y++)
                for (uint32_t X = b.x0; X < b.x1 && owns_tile_row(P, Y / TH); X++)
                    if (coverage_mask_fast(s, thr, (int)X, (int)Y)) tmask |= 1u << (((Y / TH - ty0) << 1) | (X / TW - tx0));
            if (!tmask) return; // contributes nothing (its bbox pixels are already counted)
        } else {
            tmask = 1u | (tx1 > tx0 ? 2u : 0u) | (ty1 > ty0 ? 4u : 0u) | ((tx1 > tx0 && ty1 > ty0) ? 8u : 0u);
            if (P.il_band) {
                if (!owns_tile_row(P, ty0)) tmask &= ~3u;
                if (!owns_tile_row(P, ty0 + 1)) tmask &= ~12u;
                if (!tmask) return;
            }
        }
    }

    // non-finite / absurd coordinates take the literal per-pixel walk of the tile stage; frames without any
    // (all of them, in practice) skip the detection of such items in every chunk
    const bool tame = setup_is_tame(s);
    if (!tame) atomicOr(&P.fs->has_wild, 1u);
    const uint32_t stripe = blockIdx.x % REC_STRIPES, stripe_cap = P.rec_cap / REC_STRIPES;
    const uint32_t local = alloc_slot(&P.fs->rec_cursor[stripe]);
    if (local >= stripe_cap) {
        atomicOr(&P.fs->err, ERR_REC_OVF);
        return;
    }
    const uint32_t rec = stripe * stripe_cap + local;
    uint32_t clip_attr = 0;
    if (!ca) { // the fragment shader will gather these from the mesh: pull them into L2 now (fire and forget)
        asm volatile("prefetch.global.L2 [%0];" ::"l"(D.attr + 6 * (size_t)i0));
        asm volatile("prefetch.global.L2 [%0];" ::"l"(D.attr + 6 * (size_t)i1));
        asm volatile("prefetch.global.L2 [%0];" ::"l"(D.attr + 6 * (size_t)i2));
    }
    if (ca) {
        clip_attr = alloc_slot(&P.fs->n_clip_attr);
        if (clip_attr >= P.attr_cap) {
            atomicOr(&P.fs->err, ERR_ATTR_OVF);
            return;
        }
        float4 *ar = reinterpret_cast<float4 *>(&P.attrs[clip_attr]);
        ar[0] = make_float4(ca[0][0], ca[0][1], ca[0][2], ca[0][3]);
        ar[1] = make_float4(ca[0][4], ca[0][5], ca[1][0], ca[1][1]);
        ar[2] = make_float4(ca[1][2], ca[1][3], ca[1][4], ca[1][5]);
        ar[3] = make_float4(ca[2][0], ca[2][1], ca[2][2], ca[2][3]);
        ar[4] = make_float4(ca[2][4], ca[2][5], 0.0f, 0.0f);
    }
    // RasterizerTriangle::new (mod.rs:187-222): inv_2x_area = 1.0 / triangle_2x_area(points), once per triangle
    // (area2 above is the same expression, mod.rs:15-21)
    const float inv = fdiv(1.0f, area2);
    float4 *rr = reinterpret_cast<float4 *>(&P.recs[rec]);
    rr[0] = make_float4(s.px[0], s.py[0], s.px[1], s.py[1]);
    rr[1] = make_float4(s.px[2], s.py[2], s.z[0], s.z[1]);
    rr[2] = make_float4(s.z[2], inv, __uint_as_float(key), 0.0f);
    // (the 32-byte shade record leaves in one 256-bit store)
    st_sector(reinterpret_cast<float4 *>(&P.shade[rec]),
              make_float4(__uint_as_float((D.fs & 3u) | (ca ? 4u : 0u) | (((D.fs >> 8) & 31u) << 3) | (D.draw << 8)),
                          __uint_as_float(ca ? clip_attr : i0), __uint_as_float(i1), __uint_as_float(i2)),
              make_float4(s.w[0], s.w[1], s.w[2], 0.0f));
    const uint32_t rec_tie = rec | (tie_bits(s) << 29);
    const uint32_t wild_bit = (tame ? 0u : ENTRY_WILD) | (0xFFu << ENTRY_BLOCKS_SHIFT); // small triangles: every block may be covered

    if (small) {
#pragma unroll
        for (uint32_t q = 0; q < 4; q++)
            if ((tmask >> q) & 1u) {
                const uint32_t tx = tx0 + (q & 1u), ty = ty0 + (q >> 1);
                push_bin(P, ty * P.tiles_x + tx, key, rec_tie, tile_box(b.x0, b.x1, b.y0, b.y1, tx, ty) | wild_bit);
            }
    } else {
        // one queue item per slab of LARGE_SLAB_ROWS tile rows x LARGE_CHUNK_COLS tile columns; the slots of all of them are
        // reserved with ONE atomic (a full-screen triangle on an 8192-row target has 64 slabs: a round trip per slab
        // kept its thread busy for 50 us)
        const uint32_t tyB = (b.y1 - 1) / TH + 1, txB = (b.x1 - 1) / TW + 1;
        const uint32_t n_slabs = (tyB - ty0 + LARGE_SLAB_ROWS - 1) / LARGE_SLAB_ROWS;
        const uint32_t n_chunks = (txB - tx0 + LARGE_CHUNK_COLS - 1) / LARGE_CHUNK_COLS;
        const uint32_t n_items = n_slabs * n_chunks;
        // (the lanes of the warp that arrive here with the same item count share the atomic)
        uint32_t li0 = 0;
        {
            const unsigned peers = __match_any_sync(__activemask(), n_items);
            const int leader = __ffs(peers) - 1;
            if ((int)(threadIdx.x & 31) == leader) li0 = atomicAdd(&P.fs->n_large, (uint32_t)__popc(peers) * n_items);
            li0 = __shfl_sync(peers, li0, leader) + (uint32_t)__popc(peers & lanemask_lt()) * n_items;
        }
        if (li0 + n_items > P.large_cap) atomicOr(&P.fs->err, ERR_LARGE_OVF); // (the frame is replayed with a larger queue)
        uint4 *q = reinterpret_cast<uint4 *>(P.large + li0); // {rec, key, rows, cols}
        uint32_t room = li0 < P.large_cap ? P.large_cap - li0 : 0u;
        for (uint32_t sy = ty0; sy < tyB; sy += LARGE_SLAB_ROWS) {
            const uint32_t rows = sy | (min(sy + (uint32_t)LARGE_SLAB_ROWS, tyB) << 16);
            for (uint32_t sx = tx0; sx < txB && room; sx += LARGE_CHUNK_COLS, room--)
                *q++ = make_uint4(rec, key, rows, sx | (min(sx + (uint32_t)LARGE_CHUNK_COLS, txB) << 16));
        }
    }
}

// clipping::distance_measure (rasterizer/clipping.rs:29-38); planes in CLIP_PLANES order (53-60)
// Guard band (the extension sketched at rasterizer/mod.rs:417-419): the four side planes sit at |x|, |y| <= g * w; g * w is
// one f32 product and 1.0f * w == w, so g = 1 is the reference bit for bit.
__device__ __forceinline__ float clip_distance(int plane, const float *p, float guard) {
    float c = p[plane >> 1];
    const float w = plane < 4 ? fmul(guard, p[3]) : p[3];
    return (plane & 1) ? fsub(w, c) : fadd(w, c);
}

// outcode bits of one clip-space vertex (clipping.rs:86-104): for axis a in x,y,z
//   bit a     : v[a] >= -w     bit 3+a : v[a] <= w     bit 6+a : v[a] < -w     bit 9+a : v[a] > w
// With a guard band the "inside" bits (no clipping needed) are judged against the widened side planes, the "outside"
// bits (nothing can be visible) still against the view frustum; z is never widened.
__device__ __forceinline__ uint32_t clip_code(const float *c, float guard) {
    const float w = c[3], nw = -w;
    uint32_t code = 0;
#pragma unroll
    for (int a = 0; a < 3; a++) {
        const float gw = a < 2 ? fmul(guard, w) : w, ngw = -gw;
        code |= (c[a] >= ngw ? 1u : 0u) << a;
        code |= (c[a] <= gw ? 1u : 0u) << (3 + a);
        code |= (c[a] < nw ? 1u : 0u) << (6 + a);
        code |= (c[a] > w ? 1u : 0u) << (9 + a);
    }
    return code;
}

// block-level counter reduction: warp reduce -> per-warp partials -> one striped global RED per counter
__device__ __forceinline__ void flush_geom_counters(const FrameParams &P, const GeomLocal &lc, uint32_t (*s_part)[C_COUNT],
                                                    unsigned long long *s_bbox) {
    {
        const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
#pragma unroll
        for (int k = 0; k <= C_TRIS_SETUP; k++) {
            const uint32_t v = __reduce_add_sync(0xffffffffu, lc.c[k]);
            if (lane == 0) s_part[warp][k] = v;
        }
        {
            const uint32_t v = __reduce_add_sync(0xffffffffu, lc.c[C_CLIP_OVF]);
            if (lane == 0) s_part[warp][C_CLIP_OVF] = v;
        }
        unsigned long long v = lc.bbox;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
        if (lane == 0) s_bbox[warp] = v;
    }
    __syncthreads();
    if (threadIdx.x < C_COUNT) {
        const int k = threadIdx.x;
        unsigned long long sum = 0ull;
        if (k <= C_TRIS_SETUP || k == C_CLIP_OVF) {
#pragma unroll
            for (int w = 0; w < NT / 32; w++) sum += s_part[w][k];
        } else if (k == C_BBOX_PX) {
#pragma unroll
            for (int w = 0; w < NT / 32; w++) sum += s_bbox[w];
        }
        if (sum) atomicAdd(&P.fs->counters[blockIdx.x % CNT_STRIPES][k], sum);
    }
}

// Stage 1a -- vertex stage (render.rs:104-108): one thread per mesh vertex.  The vertex shader
// (main.rs:147-152) is ((P*V)*W) * (x,y,z,1); the matrix product is hoisted to the host in the same
// operation order.  Besides the clip-space position it stores what every
// triangle using the vertex would recompute: the perspective divide + viewport transform and the
// 12 trivial-accept/reject comparisons.
// The first vertex kernel of a frame also zeroes the per-frame part of the frame state and the tile
// counters (zero16/n16 != 0): the previous frame's tile kernel has completed by pdl_wait(), and this
// frame's geometry kernels start after this kernel.
__global__ void __launch_bounds__(NT) vertex_kernel(FrameParams P, DrawParams D, uint4 *zero16, uint32_t n16) {
    pdl_launch();
    pdl_wait();
    for (uint32_t i = blockIdx.x * NT + threadIdx.x; i < n16; i += gridDim.x * NT) zero16[i] = make_uint4(0u, 0u, 0u, 0u);
    if (blockIdx.x == 0 && threadIdx.x == 0 && D.draw != 0xFFFFFFFFu) {
        // per-draw table entry for the clip and tile kernels, written on the device: a host-side copy of the
        // table would cost a host-stream synchronisation per frame (pageable cudaMemcpyAsync)
        DrawInfo *di = const_cast<DrawInfo *>(P.draws) + D.draw;
        di->attr = D.attr; di->pos = D.pos; di->idx = D.idx;
        di->nv = D.nv; di->tri_base = D.tri_base; di->fs = D.fs; di->pad = 0u;
#pragma unroll
        for (int k = 0; k < 16; k++) di->M[k] = D.M[k];
    }
    // VERTEX_PER_THREAD vertices per thread, a CTA-strided chunk each: all position loads are issued before
    // the first use, so one wave of CTAs keeps enough bytes in flight to stream the mesh from HBM
    float x[VERTEX_PER_THREAD], y[VERTEX_PER_THREAD], z[VERTEX_PER_THREAD];
    const uint32_t v0 = blockIdx.x * (NT * VERTEX_PER_THREAD) + threadIdx.x;
#pragma unroll
    for (int k = 0; k < VERTEX_PER_THREAD; k++) {
        const uint32_t v = v0 + k * NT;
        if (v < D.nv) {
            const float *p = D.pos + 3 * (size_t)v;
            x[k] = __ldg(p); y[k] = __ldg(p + 1); z[k] = __ldg(p + 2);
        }
    }
#pragma unroll
    for (int k = 0; k < VERTEX_PER_THREAD; k++) {
        const uint32_t v = v0 + k * NT;
        if (v >= D.nv) break;
        float c[4];
#pragma unroll
        for (int r = 0; r < 4; r++)
            c[r] = dot4z(D.M[4 * r], D.M[4 * r + 1], D.M[4 * r + 2], D.M[4 * r + 3], x[k], y[k], z[k], 1.0f);
        st_sector(D.vtx + 2 * (size_t)v, project_vertex(c, (float)P.W, (float)P.H),
                  make_float4(c[0], c[1], c[2], __uint_as_float(clip_code(c, P.guard))));
    }
}

// Stage 1b/2 -- primitive assembly, clip, setup, binning: one thread per input triangle
// (render.rs:75-96, rasterizer/mod.rs:425-441).
#ifndef RZ_GEOM_MIN_CTAS
#define RZ_GEOM_MIN_CTAS 4 // 64 registers, nothing spilled: 32 resident warps per SM.  (5 CTAs = 48 registers spill 128 bytes per
                           // thread and cost 3-4 us on the C2 frame, 16 us on the overdraw frame; 3 CTAs are no faster than 4.)
#endif
__global__ void __launch_bounds__(NT, RZ_GEOM_MIN_CTAS) geom_kernel(FrameParams P, DrawParams D) {
    __shared__ uint32_t s_part[NT / 32][C_COUNT];
    __shared__ unsigned long long s_bbox[NT / 32];
    pdl_launch();
    pdl_wait();

    GeomLocal lc;
#pragma unroll
    for (int k = 0; k < C_COUNT; k++) lc.c[k] = 0;
    lc.bbox = 0ull;

    // Persistent CTAs: CTA b takes the 256-triangle chunks b, b + grid, b + 2*grid, ...  The index triple of
    // the next chunk is loaded before the current one is processed (register double buffer), so the DRAM
    // latency of the index stream never sits on the critical path, the warps of a CTA run free of
    // each other (no per-chunk barrier), and the counters are flushed once per CTA.
    const uint32_t n_chunks = (D.nt + NT - 1) / NT;
    uint32_t n0 = 0, n1 = 0, n2 = 0;
    {
        const uint32_t t = blockIdx.x * NT + threadIdx.x;
        if (t < D.nt) {
            n0 = __ldg(&D.idx[3 * (size_t)t]); n1 = __ldg(&D.idx[3 * (size_t)t + 1]); n2 = __ldg(&D.idx[3 * (size_t)t + 2]);
        }
    }
    for (uint32_t chunk = blockIdx.x; chunk < n_chunks; chunk += gridDim.x) {
    const uint32_t t = chunk * NT + threadIdx.x;
    const uint32_t i0 = n0, i1 = n1, i2 = n2;
    {
        const size_t tn = (size_t)t + (size_t)gridDim.x * NT;
        if (tn < D.nt) {
            n0 = __ldg(&D.idx[3 * tn]); n1 = __ldg(&D.idx[3 * tn + 1]); n2 = __ldg(&D.idx[3 * tn + 2]);
        }
    }
    if (t < D.nt) {
        if (i0 >= D.nv || i1 >= D.nv || i2 >= D.nv) {
            atomicOr(&P.fs->err, ERR_INDEX); // the reference panics here (render.rs:83-87)
        } else {
            lc.c[C_TRIS_IN]++;
            const uint32_t key0 = (D.tri_base + t) * 8u;
            // Everything a triangle may need from its three vertices is requested up front, so the
            // independent gathers overlap instead of paying one L2 round trip per decision.
            // (written by this draw's vertex kernel, the preceding launch: plain loads, not the read-only path)
            // one 256-bit load per vertex (sm_100 LDG.256: the record is one aligned 32-byte sector)
            float4 sv0, cq0, sv1, cq1, sv2, cq2;
            ld_sector(D.vtx + 2 * (size_t)i0, sv0, cq0);
            ld_sector(D.vtx + 2 * (size_t)i1, sv1, cq1);
            ld_sector(D.vtx + 2 * (size_t)i2, sv2, cq2);
            const float2 q0 = make_float2(cq0.x, cq0.y), q1 = make_float2(cq1.x, cq1.y), q2 = make_float2(cq2.x, cq2.y);
            const uint32_t code = __float_as_uint(cq0.w) & __float_as_uint(cq1.w) & __float_as_uint(cq2.w);
            // clipping::try_clip (rasterizer/clipping.rs:62-195): degenerate test on clip-space xy first
            const float a2x = cross2(fsub(q1.x, q0.x), fsub(q1.y, q0.y), fsub(q2.x, q0.x), fsub(q2.y, q0.y));
            if (fabsf(a2x) < 0.000001f) {
                lc.c[C_DEGENERATE]++;
            } else {
                if (code & 0xFC0u) { // all three vertices outside one plane
                    lc.c[C_OUTSIDE]++;
                } else if ((code & 0x3Fu) == 0x3Fu) { // all inside all planes
                    lc.c[C_INSIDE]++;
                    Setup s;
                    s.px[0] = sv0.x; s.py[0] = sv0.y; s.z[0] = sv0.z; s.w[0] = sv0.w;
                    s.px[1] = sv1.x; s.py[1] = sv1.y; s.z[1] = sv1.z; s.w[1] = sv1.w;
                    s.px[2] = sv2.x; s.py[2] = sv2.y; s.z[2] = sv2.z; s.w[2] = sv2.w;
                    emit_setup(P, D, s, i0, i1, i2, nullptr, key0, lc);
                } else {
                    // straddles a clip plane: queued for clip_kernel (keeps this kernel's register count,
                    // and with it the number of resident warps, independent of the Sutherland-Hodgman path)
                    const uint32_t q = alloc_slot(&P.fs->n_clipq);
                    if (q < P.rec_cap) P.clipq[q] = ((unsigned long long)D.draw << 32) | t;
                    else atomicOr(&P.fs->err, ERR_REC_OVF);
                }
            }
        }
    }

    } // persistent chunk loop
    flush_geom_counters(P, lc, s_part, s_bbox);
}

// Stage 1c -- triangles that straddle a clip plane (ClipResult::Clipped candidates), all draws of the
// frame: one thread per queued triangle, persistent grid.  The clip-space vertices are recomputed from
// the mesh (the same operation sequence as vertex_kernel, hence the same bits), clipped by
// Sutherland-Hodgman and re-triangulated as a fan (clipping.rs:118-190).
__global__ void __launch_bounds__(NT) clip_kernel(FrameParams P) {
    __shared__ uint32_t s_part[NT / 32][C_COUNT];
    __shared__ unsigned long long s_bbox[NT / 32];
    pdl_launch();
    pdl_wait();
    const uint32_t nq = min(P.fs->n_clipq, P.rec_cap);
    if (nq == 0u) return; // nothing straddles a clip plane (the common case for a mesh inside the frustum)
    GeomLocal lc;
#pragma unroll
    for (int k = 0; k < C_COUNT; k++) lc.c[k] = 0;
    lc.bbox = 0ull;
    for (uint32_t qi = blockIdx.x * NT + threadIdx.x; qi < nq; qi += gridDim.x * NT) {
        const unsigned long long e = P.clipq[qi];
        const uint32_t t = (uint32_t)e;
        const DrawInfo &di = P.draws[(uint32_t)(e >> 32)];
        DrawParams D; // what emit_setup needs
        D.attr = di.attr; D.fs = di.fs; D.draw = (uint32_t)(e >> 32);
        const uint32_t key0 = (di.tri_base + t) * 8u;
        float pv[2][MAX_POLY][4];
        float pa[2][MAX_POLY][6];
        for (int v = 0; v < 3; v++) {
            const uint32_t vi = __ldg(&di.idx[3 * (size_t)t + v]);
            const float *p = di.pos + 3 * (size_t)vi;
            const float x = __ldg(p), y = __ldg(p + 1), z = __ldg(p + 2);
            for (int r = 0; r < 4; r++)
                pv[0][v][r] = dot4z(di.M[4 * r], di.M[4 * r + 1], di.M[4 * r + 2], di.M[4 * r + 3], x, y, z, 1.0f);
            const float *a = di.attr + 6 * (size_t)vi;
            for (int k = 0; k < 6; k++) pa[0][v][k] = __ldg(a + k);
        }
        int n_out = 3, cur = 0;
        bool ovf = false;
        // Sutherland-Hodgman against LEFT,RIGHT,BOTTOM,TOP,NEAR,FAR (clipping.rs:118-171)
        for (int plane = 0; plane < 6; plane++) {
            {   // a plane every vertex is inside of leaves the polygon as it is (the loop below would copy each vertex,
                // in order, through local memory): most clipped triangles cross one plane of the six
                bool all_in = true;
                for (int i = 0; i < n_out; i++) all_in = all_in && clip_distance(plane, pv[cur][i], P.guard) >= 0.0f;
                if (all_in) continue;
            }
            const int n_in = n_out, in = cur, out = cur ^ 1;
            n_out = 0;
            for (int i = 0; i < n_in; i++) {
                const int prev = (i + n_in - 1) % n_in;
                const float *pvv = pv[in][prev], *cvv = pv[in][i];
                const float pd = clip_distance(plane, pvv, P.guard), cd = clip_distance(plane, cvv, P.guard);
                const bool pin = pd >= 0.0f, cin = cd >= 0.0f;
                if (!pin && !cin) continue;
                if (n_out + 2 > MAX_POLY) {
                    ovf = true;
                    break;
                }
                if (pin != cin) { // crossing: intersection first (clipping.rs:43-51)
                    const float alpha = fdiv(pd, fsub(pd, cd));
                    const float om = fsub(1.0f, alpha);
                    for (int k = 0; k < 4; k++) pv[out][n_out][k] = fadd(fmul(pvv[k], om), fmul(cvv[k], alpha));
                    for (int k = 0; k < 6; k++) // (cur - prev) * alpha + prev (clipping.rs:149,161)
                        pa[out][n_out][k] = fadd(fmul(fsub(pa[in][i][k], pa[in][prev][k]), alpha), pa[in][prev][k]);
                    n_out++;
                }
                if (cin) {
                    for (int k = 0; k < 4; k++) pv[out][n_out][k] = cvv[k];
                    for (int k = 0; k < 6; k++) pa[out][n_out][k] = pa[in][i][k];
                    n_out++;
                }
            }
            cur = out;
        }
        if (ovf) lc.c[C_CLIP_OVF]++;
        if (n_out < 3) {
            lc.c[C_OUTSIDE]++; // late outside (clipping.rs:175-177)
        } else {
            lc.c[C_CLIPPED_IN]++;
            const float Wf = (float)P.W, Hf = (float)P.H;
            const float4 s0 = project_vertex(pv[cur][0], Wf, Hf);
            float4 sb = project_vertex(pv[cur][1], Wf, Hf);
            for (int i = 0; i + 2 < n_out; i++) { // fan (0, i+1, i+2) (clipping.rs:185-190)
                const float4 sc = project_vertex(pv[cur][i + 2], Wf, Hf);
                Setup s;
                s.px[0] = s0.x; s.py[0] = s0.y; s.z[0] = s0.z; s.w[0] = s0.w;
                s.px[1] = sb.x; s.py[1] = sb.y; s.z[1] = sb.z; s.w[1] = sb.w;
                s.px[2] = sc.x; s.py[2] = sc.y; s.z[2] = sc.z; s.w[2] = sc.w;
                float ca[3][6];
                for (int k = 0; k < 6; k++) {
                    ca[0][k] = pa[cur][0][k]; ca[1][k] = pa[cur][i + 1][k]; ca[2][k] = pa[cur][i + 2][k];
                }
                emit_setup(P, D, s, 0u, 0u, 0u, ca, key0 + (uint32_t)min(i, 7), lc);
                sb = sc;
            }
        }
    }
    flush_geom_counters(P, lc, s_part, s_bbox);
}

// Screen points of a record (quarters q0, q1) + edge normals.  Records are written by the geometry kernels of the
// same frame (the preceding launches of the PDL chain): plain loads after griddepcontrol.wait, never the
// non-coherent read-only path.
__device__ __forceinline__ void load_points(const RasterRec *recs, uint32_t rec, Setup &s) {
    const float4 *rr = reinterpret_cast<const float4 *>(&recs[rec]);
    const float4 r0 = rr[0], r1 = rr[1];
    s.px[0] = r0.x; s.py[0] = r0.y; s.px[1] = r0.z; s.py[1] = r0.w;
    s.px[2] = r1.x; s.py[2] = r1.y; s.z[0] = r1.z; s.z[1] = r1.w;
    setup_normals(s);
}

// Large-triangle binning: persistent warps steal (triangle, slab of tile rows) items; the 32
// lanes test the slab's tiles in parallel.  A tile is skipped only when, for some edge, the
// most favourable sample position of tile-cap-bbox already fails that edge -- exact because the
// f32 edge function is monotone in x and in y.
__global__ void __launch_bounds__(NT) large_bin_kernel(FrameParams P) {
    pdl_launch();
    pdl_wait();
    const uint32_t n = min(P.fs->n_large, P.large_cap);
    if (n == 0u) return; // nothing queued (meshes of small triangles): skip the round trip of the cursor atomic
    const int lane = threadIdx.x & 31;
    for (;;) { // one warp per (triangle, slab) item, stolen from a global cursor
        uint32_t item = 0;
        if (lane == 0) item = atomicAdd(&P.fs->large_next, 1u);
        item = __shfl_sync(0xffffffffu, item, 0);
        if (item >= n) break;
        const LargeItem li = P.large[item];
        Setup s;
        load_points(P.recs, li.rec, s);
        BBox b = pixel_bbox(s, P.scissor);
        b.y0 = max(b.y0, P.row_begin);
        b.y1 = min(b.y1, P.row_end);
        const uint32_t rec_tie = li.rec | (tie_bits(s) << 29);
        const uint32_t wild_bit = setup_is_tame(s) ? 0u : ENTRY_WILD;
        // Block masks pay for slivers (a thin triangle crossing a tile touches about half of its 8x4 blocks); a fat
        // triangle covers nearly all blocks of the tiles it touches, so its entries simply carry 0xFF.  Heuristic only:
        // any mask that includes every block the triangle can cover is correct.
        bool sliver;
        {
            const float a2 = fabsf(cross2(fsub(s.px[1], s.px[0]), fsub(s.py[1], s.py[0]), fsub(s.px[2], s.px[0]), fsub(s.py[2], s.py[0])));
            const float bw_f = fmaxf(fmaxf(s.px[0], s.px[1]), s.px[2]) - fminf(fminf(s.px[0], s.px[1]), s.px[2]);
            const float bh_f = fmaxf(fmaxf(s.py[0], s.py[1]), s.py[2]) - fminf(fminf(s.py[0], s.py[1]), s.py[2]);
            sliver = a2 < RZ_SLIVER_FRAC * bw_f * bh_f; // area2 / 2 against a fraction of the bounding box area
        }
        const uint32_t tx0 = li.cols & 0xFFFFu, ntx = (li.cols >> 16) - tx0, li_ty0 = li.rows & 0xFFFFu;
        const uint32_t total = ntx * ((li.rows >> 16) - li_ty0);
        // bounds of the sample positions: the 4-sample rotated grid spans [1/8, 7/8] of a pixel, the other patterns
        // are simply bounded by the pixel itself
        const float o_lo = P.msaa == 4u ? 0.125f : 0.0f, o_hi = P.msaa == 4u ? 0.875f : 1.0f;
        for (uint32_t t0 = 0; t0 < total; t0 += 32) { // (warp-uniform trip count: the block tests below are warp collectives)
            const uint32_t t = t0 + lane;
            const uint32_t tx = tx0 + t % ntx, ty = li_ty0 + t / ntx;
            const uint32_t X0 = max(b.x0, tx * TW), X1 = min(b.x1, tx * TW + TW);
            const uint32_t Y0 = max(b.y0, ty * TH), Y1 = min(b.y1, ty * TH + TH);
            bool keep = t < total && X0 < X1 && Y0 < Y1 && owns_tile_row(P, ty);
            bool full = true;
            if (keep) {
                const float sx_lo = fadd((float)X0, o_lo), sx_hi = fadd((float)(X1 - 1), o_hi);
                const float sy_lo = fadd((float)Y0, o_lo), sy_hi = fadd((float)(Y1 - 1), o_hi);
#pragma unroll
                for (int k = 0; k < 3; k++) {
                    const float cx = s.nx[k] >= 0.0f ? sx_hi : sx_lo;
                    const float cy = s.ny[k] >= 0.0f ? sy_hi : sy_lo;
                    keep = keep && edge_pass(edge_eval(s, k, cx, cy), s.nx[k], s.ny[k]);
                }
                // a tile every sample of which is inside (all three edges pass at their LEAST favourable corner; monotone
                // again) needs no block tests: the interior tiles of a large triangle
#pragma unroll
                for (int k = 0; k < 3 && keep; k++) {
                    const float cx = s.nx[k] >= 0.0f ? sx_lo : sx_hi;
                    const float cy = s.ny[k] >= 0.0f ? sy_lo : sy_hi;
                    full = full && edge_pass(edge_eval(s, k, cx, cy), s.nx[k], s.ny[k]);
                }
            }
            // the same exact reject one level down: the eight 8x4 pixel blocks of the tile (one per warp of the tile
            // kernel's pixel-parallel walks).  A tile none of whose blocks survives is not binned at all.  The tiles of
            // this trip that need the test are handled four at a time, lane = (tile slot, block): a sliver's item has a
            // dozen tiles, so a lane per tile would leave most of the warp idle during the 24 edge evaluations.
            // (non-finite coordinates: every block, the monotonicity argument needs finite values)
            // (fat triangles: every block of a touched tile may be covered, no block tests)
            const bool need = keep && !full && sliver && !wild_bit;
            uint32_t blocks = keep ? 0xFFu : 0u;
            unsigned need_mask = __ballot_sync(0xffffffffu, need);
            while (need_mask) {
                const int slot = lane >> 3;
                const uint32_t blk = (uint32_t)lane & 7u;
                unsigned m = need_mask; // the slot-th set bit of need_mask: the lane whose tile this lane helps with
#pragma unroll
                for (int i = 0; i < 3; i++)
                    if (i < slot) m &= m - 1u;
                const bool act = m != 0u;
                const int src = act ? __ffs(m) - 1 : 0;
                const uint32_t stx = __shfl_sync(0xffffffffu, tx, src), sty = __shfl_sync(0xffffffffu, ty, src);
                bool kb = false;
                if (act) {
                    const uint32_t SX0 = max(b.x0, stx * TW), SX1 = min(b.x1, stx * TW + TW);
                    const uint32_t SY0 = max(b.y0, sty * TH), SY1 = min(b.y1, sty * TH + TH);
                    const uint32_t bx0 = max(SX0, stx * TW + (blk & 1u) * BLOCK_W), bx1 = min(SX1, stx * TW + (blk & 1u) * BLOCK_W + BLOCK_W);
                    const uint32_t by0 = max(SY0, sty * TH + (blk >> 1) * BLOCK_H), by1 = min(SY1, sty * TH + (blk >> 1) * BLOCK_H + BLOCK_H);
                    if (bx0 < bx1 && by0 < by1) {
                        const float bx_lo = fadd((float)bx0, o_lo), bx_hi = fadd((float)(bx1 - 1), o_hi);
                        const float by_lo = fadd((float)by0, o_lo), by_hi = fadd((float)(by1 - 1), o_hi);
                        kb = true;
#pragma unroll
                        for (int k = 0; k < 3; k++) {
                            const float cx = s.nx[k] >= 0.0f ? bx_hi : bx_lo;
                            const float cy = s.ny[k] >= 0.0f ? by_hi : by_lo;
                            kb = kb && edge_pass(edge_eval(s, k, cx, cy), s.nx[k], s.ny[k]);
                        }
                    }
                }
                const unsigned res = __ballot_sync(0xffffffffu, kb);
                const int rank = __popc(need_mask & lanemask_lt());
                if (((need_mask >> lane) & 1u) && rank < 4) blocks = (res >> (8 * rank)) & 0xFFu;
#pragma unroll
                for (int i = 0; i < 4; i++) need_mask &= need_mask - 1u; // (x & (x - 1) of 0 is 0)
            }
            if (blocks) push_bin(P, ty * P.tiles_x + tx, li.key, rec_tie, tile_box(b.x0, b.x1, b.y0, b.y1, tx, ty) | wild_bit | (blocks << ENTRY_BLOCKS_SHIFT) |
                                      ((full && !wild_bit) ? ENTRY_FULL : 0u));
        }
    }
}

// Work order of the tile stage: tiles are handed out longest-list-first (LPT scheduling), so the few
// tiles with hundreds of triangles start early instead of forming the tail of the frame.  One thread
// per tile of the shard files its tile under a list-length class.
__device__ __forceinline__ uint32_t order_class(uint32_t n) {
    return n >= 512u ? 0u : n >= 384u ? 1u : n >= 256u ? 2u : n >= 192u ? 3u : n >= 128u ? 4u : n >= 96u ? 5u : n >= 48u ? 6u : 7u;
}
__global__ void __launch_bounds__(NT) order_kernel(FrameParams P) {
    pdl_launch();
    pdl_wait();
    const uint32_t shard_tiles = P.tiles_x * (P.ty_end - P.ty_begin);
    const uint32_t i = blockIdx.x * NT + threadIdx.x;
    if (i >= shard_tiles) return;
    const uint32_t tile = P.ty_begin * P.tiles_x + i;
    const uint32_t n = P.tile_count[tile];
    if (n == 0u) return;
    {   // how many tiles hold only a few items (the host picks the tile-kernel instantiation with the short-list walk by it)
        const unsigned few = __ballot_sync(__activemask(), n <= (uint32_t)FAST_N);
        if (few && (threadIdx.x & 31) == (unsigned)(__ffs(few) - 1)) atomicAdd(&P.fs->n_few_tiles, (uint32_t)__popc(few));
    }
    const uint32_t b = order_class(n);
    const unsigned peers = __match_any_sync(__activemask(), b);
    const int leader = __ffs(peers) - 1;
    uint32_t base = 0;
    if ((int)(threadIdx.x & 31) == leader) base = atomicAdd(&P.fs->bucket_n[b], (uint32_t)__popc(peers));
    base = __shfl_sync(peers, base, leader);
    // the entry carries the list length (bounded by the bin's capacity) and the tile's first bin entry too: the tile stage
    // learns all of it with one load
    const uint2 tb = __ldg(reinterpret_cast<const uint2 *>(&P.tile_bin[tile]));
    P.busy[(size_t)b * P.tiles_x * P.tiles_y + base + __popc(peers & lanemask_lt())] = make_uint4(tile, min(n, tb.y), tb.x, 0u);
}

// ---- completion flags over peer memory (screen-space sharding, one process per GPU) ----
// A rank whose tile kernel stored its rows straight into another GPU's image (NVLink peer stores) raises
// flag[rank] there; the owner of the image spins until every flag has reached the frame's sequence number.
// Launched WITHOUT programmatic serialization: the stream order guarantees the tile kernel has completed.
struct FlagList {
    uint32_t *p[16];
};
__global__ void signal_kernel(FlagList flags, uint32_t n, uint32_t value) {
    if (threadIdx.x >= n) return;
    __threadfence_system();
    asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(flags.p[threadIdx.x]), "r"(value) : "memory");
}
// *timed_out is set (and the kernel returns) if a flag does not arrive within timeout_ns: a dead peer must
// not hang the GPU.
__global__ void wait_flags_kernel(const uint32_t *flags, uint32_t n, uint32_t stride_words, uint32_t value, unsigned long long timeout_ns,
                                  uint32_t *timed_out) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    unsigned long long t0, t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t0));
    for (;;) {
        uint32_t v;
        asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(flags + (size_t)i * stride_words) : "memory");
        if ((int32_t)(v - value) >= 0) break;
        __nanosleep(200);
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
        if (t - t0 > timeout_ns) {
            atomicOr(timed_out, 1u);
            break;
        }
    }
}

// First kernel of a frame: zero the per-frame part of the frame state and the tile counters.
__global__ void __launch_bounds__(NT) frame_begin_kernel(uint4 *p, uint32_t n16) {
    pdl_launch();
    pdl_wait();
    const uint32_t i = blockIdx.x * NT + threadIdx.x;
    if (i < n16) p[i] = make_uint4(0u, 0u, 0u, 0u);
}

} // namespace rz
