// rz_types.cuh -- device-side data layout of one frame (see DESIGN.md "Data layout in HBM").
#pragma once
#include <stdint.h>

namespace rz {

constexpr int TW = 16;              // screen tile width  (pixels)
constexpr int TH = 16;              // screen tile height (pixels)
constexpr int TILE_PX = TW * TH;    // 256 = threads per tile CTA
constexpr int NT = 256;             // threads per CTA in every kernel
#ifndef RZ_CHUNK
#define RZ_CHUNK 255
#endif
constexpr int CHUNK = RZ_CHUNK;     // items per triangle-parallel chunk (item id fits u8, 0xFF = none; <= 255)
#ifndef RZ_UNIT_CAP
#define RZ_UNIT_CAP 6144
#endif
constexpr int UNIT_CAP = RZ_UNIT_CAP; // (item, pixel) work units per chunk (one byte each in shared memory)
#ifndef RZ_POOL
#define RZ_POOL 1024
#endif
constexpr int POOL = RZ_POOL;       // per-chunk fragment records held in shared memory
#ifndef RZ_DIRECT_MIN_AREA
#define RZ_DIRECT_MIN_AREA 96
#endif
constexpr int DIRECT_MIN_AREA = RZ_DIRECT_MIN_AREA; // average in-tile bbox (pixels) from which a chunk is walked pixel-parallel
#ifndef RZ_FAST_N
#define RZ_FAST_N 2
#endif
constexpr int FAST_N = RZ_FAST_N;           // tiles with at most this many items are walked pixel-parallel straight from their bin
constexpr int SORT_CAP = 2048;      // tile lists up to this length are sorted in shared memory
constexpr int GEOM_SMALL_DIM = 16;  // bbox extent up to which a triangle is binned directly (spans at most 2x2 tiles)
#ifndef RZ_GEOM_THIN_PX
#define RZ_GEOM_THIN_PX 64
#endif
#ifndef RZ_GEOM_THIN_AREA2
#define RZ_GEOM_THIN_AREA2 1.0f
#endif
constexpr int GEOM_THIN_PX = RZ_GEOM_THIN_PX;    // bbox area up to which a thin triangle is pre-rasterised exactly
constexpr float GEOM_THIN_AREA2 = RZ_GEOM_THIN_AREA2; // 2x screen area below which a small triangle is pre-rasterised
#ifndef RZ_VERTEX_PER_THREAD
#define RZ_VERTEX_PER_THREAD 1 // 1, 2 and 4 measured on one box: 41.0 / 42.0 / 42.8 us for the C2 geometry stage
#endif
constexpr int VERTEX_PER_THREAD = RZ_VERTEX_PER_THREAD; // vertices per thread of the vertex stage
#ifndef RZ_LARGE_SLAB_ROWS
#define RZ_LARGE_SLAB_ROWS 8
#endif
constexpr int LARGE_SLAB_ROWS = RZ_LARGE_SLAB_ROWS; // tile rows per large-triangle binning work item
#ifndef RZ_LARGE_CHUNK_COLS
#define RZ_LARGE_CHUNK_COLS 64
#endif
constexpr int LARGE_CHUNK_COLS = RZ_LARGE_CHUNK_COLS; // ... and tile columns: a screen-sized triangle becomes many items, one warp each
constexpr int MAX_POLY = 10;        // clipped polygon vertex budget (=> <= 8 fan triangles, 3 key bits)

constexpr uint32_t CLEAR_COLOR = 0xFF191919u;    // rasterizer/buffers.rs:5
constexpr float CLEAR_DEPTH = 3.40282347e+38f;   // f32::MAX, rasterizer/buffers.rs:6
constexpr uint32_t NO_OWNER = 0xFFFFFFFFu;

// error / overflow flags accumulated in FrameState::err
enum : uint32_t {
    ERR_REC_OVF = 1u,    // record array too small
    ERR_BIN_OVF = 2u,    // a tile list outgrew its bin
    ERR_LARGE_OVF = 4u,  // large-triangle queue too small
    ERR_INDEX = 8u,      // mesh index >= nv
    ERR_ATTR_OVF = 16u,  // clipped-attribute array too small
};

// counter slots in FrameState::counters (same order as rz_counters_t)
enum {
    C_TRIS_IN, C_DEGENERATE, C_OUTSIDE, C_INSIDE, C_CLIPPED_IN, C_TRIS_SETUP, C_BBOX_PX, C_COVERED_PX,
    C_SHADED_PX, C_SAMPLES, C_TEX_OOB, C_CLIP_OVF, C_COUNT
};

// Device-side frame bookkeeping.  `counters` and `err` persist across frames (read by rz_counters /
// rz_sync); everything from `n_records` on, and tile_count[] which follows in the same allocation,
// is zeroed by one memset at the start of every frame.
#ifndef RZ_TILE_CTAS
#define RZ_TILE_CTAS 4 // resident tile CTAs per SM (64 registers, 55 KB shared memory each)
#endif
constexpr int ORDER_BUCKETS = 8; // busy tiles are handed out longest-list-first in 8 classes
constexpr int REC_STRIPES = 64; // record slots are handed out from per-stripe cursors (CTA id % stripes):
                                // one hot cursor would serialise ~10^4 same-address atomics in L2
constexpr int CNT_STRIPES = 32; // counters are striped over CTAs to spread the global atomics
struct FrameState {
    unsigned long long counters[CNT_STRIPES][16];
    uint32_t err;         // sticky ERR_* flags
    uint32_t peer_timeout; // set by wait_flags_kernel when a peer's flag did not arrive
    uint32_t pad0[2];
    uint32_t n_clipq;     // triangles queued for the clip kernel           <- per-frame part starts here
    uint32_t n_large;     // large-triangle binning work items
    uint32_t n_clip_attr; // AttrRec slots handed out to clipped triangles
    uint32_t large_next;  // work-stealing cursor of the large binning kernel
    uint32_t has_wild;    // a triangle with NaN / inf / absurd screen coordinates was emitted (tile stage: literal walk)
    uint32_t tile_cursor; // work-stealing cursor of the tile kernel
    uint32_t n_few_tiles; // busy tiles whose list has at most FAST_N items (order_kernel): they take the barrier-free walk
    uint32_t pad1;
    uint32_t rec_cursor[REC_STRIPES]; // emitted (post-clip, post-cull) triangles per stripe
    uint32_t bucket_n[ORDER_BUCKETS]; // non-empty tiles per list-length class (class 0 = longest lists)
};

// Raster record: RasterizerTriangle (rasterizer/mod.rs:178-222) as coverage and depth need it -- the three screen
// points, the depths, inv_2x_area and the submission-order key.  Computed ONCE per triangle by the geometry stage (the
// tile stage used to redo the setup per (triangle, tile)).  48 B = 3 x float4:
//   q0 = p0x p0y p1x p1y      q1 = p2x p2y z0 z1      q2 = z2 inv key -
struct __align__(16) RasterRec {
    float p0x, p0y, p1x, p1y;
    float p2x, p2y, z0, z1;
    float z2, inv;
    uint32_t key;   // submission order: 8 * (triangle number in frame) + fan index
    uint32_t pad;
};
static_assert(sizeof(RasterRec) == 48, "RasterRec is three 16-byte quarters");

// Shade record: what only visible fragments need -- the fragment shader id, where the three VertexAttributes live and
// depths_camera_space.  Unclipped triangles point at the mesh's own attribute array through their vertex indices
// (nothing is copied); clipped ones at an AttrRec.  32 B = one sector.
struct __align__(16) ShadeRec {
    uint32_t info;        // fs (2 bits) | clipped << 2 | texture index << 3 (5 bits) | draw << 8
    uint32_t i0, i1, i2;  // vertex indices into the draw's attribute array (unclipped); i0 = AttrRec index (clipped)
    float w0, w1, w2;
    uint32_t pad;
};

// Tile bin entry, 16 B, written by the geometry stage for every (triangle, tile) pair:
//   x = order key
//   y = record index (29 bits) | tie-break bits of the three edges << 29  (EdgeFunctions::inside, mod.rs:160-168:
//       bit k set <=> a sample exactly on edge k counts as inside, i.e. n.x > 0 || (n.x == 0 && n.y < 0))
//   z = in-tile pixel box: lx0 | ly0 << 4 | (bw - 1) << 8 | (bh - 1) << 12, bit 16 = non-finite / absurd coordinates
//       (literal per-pixel walk), bits 17..24 = block mask: bit b set <=> the triangle may cover a sample of the 8x4
//       pixel block b of the tile (b = block row * 2 + block column; exact corner reject by the large-triangle binner,
//       0xFF from the small-triangle path).  The pixel-parallel walks give one block to each warp and skip the rest.
//   w = unused
constexpr uint32_t ENTRY_REC_MASK = 0x1FFFFFFFu;
constexpr uint32_t ENTRY_WILD = 1u << 16;
constexpr int ENTRY_BLOCKS_SHIFT = 17; // 8-bit block mask
constexpr uint32_t ENTRY_FULL = 1u << 25; // every sample of the in-tile box is inside the triangle (proved by the binner at the
                                          // least favourable corner of each edge): the short-list walk (fast_tile) skips the coverage test
#ifndef RZ_SLIVER_FRAC
#define RZ_SLIVER_FRAC 0.25f // a large triangle whose 2x area is below this fraction of its bounding box area gets exact block masks
#endif
constexpr int BLOCK_W = 8, BLOCK_H = 4;  // pixel block of one warp in the pixel-parallel phases (2 x 4 blocks per tile)
// Per-tile bin: `off` = first entry in FrameParams::bins, `cap` = entries the tile may hold (planned on the host from the
// counts of the last frame that overflowed: memory is O(total entries), a single hot tile no longer sizes every bin)
struct TileBin {
    uint32_t off, cap;
};

// Interpolated attributes of a clipped triangle: VertexAttribute x3 (graphics_primitives.rs:10-13),
// 18 floats padded to 80 B.
struct __align__(16) AttrRec {
    float a[20];
};

// Per-draw data the shading step and the clip kernel dereference (one entry per rz_render call)
struct DrawInfo {
    const float *attr;   // [nv][6]
    const float *pos;    // [nv][3]
    const uint32_t *idx; // [3*nt]
    uint32_t nv, tri_base, fs, pad;
    float M[16];         // (projection * view) * world
};

// Large-triangle binning work item
struct __align__(16) LargeItem {
    uint32_t rec, key;
    uint32_t rows; // tile rows [ty0, ty1):    ty0 | ty1 << 16   (a framebuffer has at most 4096 tiles per axis)
    uint32_t cols; // tile columns [tx0, tx1): tx0 | tx1 << 16
};

struct TexInfo {
    const uint8_t *data;
    unsigned long long len;
    uint32_t w, h, tw, bound;
};

// Everything a kernel needs about the frame; passed by value.
struct FrameParams {
    uint32_t W, H;               // framebuffer size
    uint32_t tiles_x, tiles_y;   // tile grid of the whole framebuffer
    uint32_t ty_begin, ty_end;   // tile rows owned by this ctx (screen-space shard)
    uint32_t row_begin, row_end; // same in pixel rows
    uint4 scissor;               // {x0, y0, x1, y1}: bounds every triangle's pixel bbox (default = the viewport)
    uint32_t msaa;               // samples per pixel: 4 = the reference (mod.rs:23); 1, 2, 8 run the generic tile kernel
    float guard;                 // guard band factor g >= 1: side clip planes at |x|, |y| <= g * w (mod.rs:417-419); 1 = reference
    uint32_t il_band, il_rank, il_world; // interleaved ownership of tile-row bands (il_band == 0: off), see owns_tile_row()
    uint32_t rec_cap, large_cap;
    FrameState *fs;
    uint32_t *tile_count;        // [tiles_x * tiles_y]
    uint4 *busy;                 // [ORDER_BUCKETS][tiles_x * tiles_y] non-empty tiles per class: {tile id, list length, first bin entry, -}
    uint4 *bins;                 // bin entries of all tiles (see TileBin); tile t owns [tile_bin[t].off, +cap)
    const TileBin *tile_bin;     // [tiles_x * tiles_y]
    RasterRec *recs;
    ShadeRec *shade;
    unsigned long long *clipq;   // [rec_cap] draw << 32 | triangle: triangles that straddle a clip plane
    AttrRec *attrs;              // clipped triangles only
    const DrawInfo *draws;
    uint32_t attr_cap;
    LargeItem *large;
    uint32_t *out;               // resolved framebuffer u32[H][W]
    uint32_t spread_clears;      // caller-owned destination (possibly a peer GPU's memory): empty-tile clears are
                                 // issued a few per rasterised tile instead of as one burst at the end
    float *dbg_depth;            // optional [H][W][4]
    uint32_t *dbg_color;
    uint32_t *dbg_owner;
    unsigned long long *dbg_tile_time; // optional [tiles][4]: tile id | n << 32, start ns, end ns, SM id
    TexInfo tex0;                // texture 0 by value (the built-in FS Texture, main.rs:69-71)
    const TexInfo *tex_table;    // every bound texture (fragment shaders with a texture index != 0)
};

// Screen-space sharding: besides the contiguous row range a ctx may own every il_world-th band of il_band
// tile rows (band k belongs to rank k % il_world), which balances a centred object across the GPUs.
__host__ __device__ inline bool owns_tile_row(const FrameParams &P, uint32_t ty) {
    return P.il_band == 0u || (ty / P.il_band) % P.il_world == P.il_rank;
}
// number of owned pixel rows in [y0, y1) (rows inside the ctx's row range)
__host__ __device__ inline uint32_t owned_rows(const FrameParams &P, uint32_t y0, uint32_t y1) {
    if (P.il_band == 0u) return y1 - y0;
    const uint32_t band_px = P.il_band * TH, period = band_px * P.il_world, lo = P.il_rank * band_px;
    // f(y) = owned rows in [0, y)
    const uint32_t r1 = y1 % period, r0 = y0 % period;
    const uint32_t f1 = (y1 / period) * band_px + (r1 > lo ? (r1 - lo < band_px ? r1 - lo : band_px) : 0u);
    const uint32_t f0 = (y0 / period) * band_px + (r0 > lo ? (r0 - lo < band_px ? r0 - lo : band_px) : 0u);
    return f1 - f0;
}

// One draw call (Renderer::render, render.rs:98-114)
struct DrawParams {
    const float *pos;      // [nv][3]
    const float *attr;     // [nv][6]
    const uint32_t *idx;   // [3*nt]
    float4 *vtx;           // [nv][2] vertex-stage output, one 32 B sector per vertex:
                           //   [0] = screen x, y, depth, clip w   (perspective divide + viewport)
                           //   [1] = clip x, y, z, bits(the 12 trivial accept/reject comparisons)
    uint32_t nv, nt;
    uint32_t tri_base;     // triangle number of this draw's first triangle inside the frame
    uint32_t fs;           // shader id | texture index << 8
    uint32_t draw;         // index into FrameParams::draws
    float M[16];           // (projection * view) * world, row-major (main.rs:147-152)
};

} // namespace rz
