// rz_api.cu -- C ABI (include/rz.h) and host-side frame orchestration for the sm_100a kernels.
//
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -fmad=false -prec-div=true -prec-sqrt=true
//        -ftz=false -Xcompiler -ffp-contract=off ... (see __graft_entry__.build()).
#include <cuda_runtime.h>
#include <nvtx3/nvToolsExt.h> // header-only; ranges are no-ops unless a profiler is attached

#include <algorithm>
#include <mutex>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <new>
#include <string>
#include <vector>

#include "../../include/rz.h"
#include "rz_exact.cuh"
#include "rz_geom.cuh"
#include "rz_tile.cuh"
#include "rz_msaa.cuh"
#include "rz_types.cuh"

using namespace rz;

#ifndef RZ_LARGE_BIN_CTAS
#define RZ_LARGE_BIN_CTAS 2 // CTAs per SM of the large-triangle binning kernel
#endif

struct rz_mesh {
    rz_ctx *ctx;
    int device = 0; // kept here too: a mesh may be destroyed after its ctx
    float *d_pos = nullptr;
    float *d_attr = nullptr;
    uint32_t *d_idx = nullptr;
    uint32_t nv = 0;
    uint64_t n_idx = 0;
    size_t cap_pos = 0, cap_attr = 0, cap_idx = 0; // bytes (staging meshes are reused and grown)
};

struct DrawCmd {
    const rz_mesh *mesh;
    uint32_t fs;
    float M[16];
};

struct Texture {
    uint8_t *d_data = nullptr;
    uint32_t w = 0, h = 0, tw = 0;
    size_t len = 0;
};

struct rz_ctx {
    int device = 0;
    uint32_t W = 0, H = 0, tiles_x = 0, tiles_y = 0;
    uint32_t row_begin = 0, row_end = 0;
    uint32_t il_band = 0, il_rank = 0, il_world = 1;
    uint32_t sc_x0 = 0, sc_y0 = 0, sc_x1 = 0, sc_y1 = 0; // scissor rect (rz_set_scissor), default = viewport
    uint32_t msaa = 4;     // samples per pixel (rz_set_msaa); 4 = the reference
    uint32_t dbg_msaa = 0; // sample count the debug capture buffers were sized for
    float guard = 1.0f;    // guard band factor (rz_set_guard_band); 1 = the reference's clip planes
    cudaStream_t own_stream = nullptr, stream = nullptr;
    float world[16], view[16], proj[16];
    std::vector<Texture> textures;
    std::vector<DrawCmd> draws;
    // Host-mesh draws (rz_render_host) of the current frame use staging[parity][k].  The uploads run on
    // their own stream and the two staging sets alternate per frame, so the H2D copy of frame i+1
    // overlaps the kernels of frame i (and the D2H copy of frame i-1 on `down_stream`).
    std::vector<rz_mesh *> staging[2];
    size_t staging_used = 0;
    int parity = 0, out_parity = 0;
    cudaStream_t up_stream = nullptr, down_stream = nullptr;
    cudaEvent_t ev_uploaded = nullptr;
    cudaEvent_t ev_consumed[2] = {nullptr, nullptr};  // the frame that read staging[p] has finished
    cudaEvent_t ev_frame_done[2] = {nullptr, nullptr};
    cudaEvent_t ev_d2h_done[2] = {nullptr, nullptr};  // the image in d_out_ring[p] has reached the host
    uint32_t *d_out_ring[2] = {nullptr, nullptr};     // [0] aliases d_out

    // device buffers
    unsigned char *d_state = nullptr; // FrameState + tile_count[]
    uint4 *d_bins = nullptr;          // bin entries of all tiles
    TileBin *d_tile_bin = nullptr;    // [tiles] {first entry, capacity}
    uint64_t bins_cap = 0;            // entries allocated in d_bins
    bool bins_planned = false;
    std::vector<TileBin> bin_plan;    // host copy of the plan: capacities only ever grow (alternating scenes converge)
    uint64_t bin_floor = 0;
    RasterRec *d_recs = nullptr;
    unsigned long long *d_clipq = nullptr;
    ShadeRec *d_shade = nullptr;
    AttrRec *d_attrs = nullptr;
    DrawInfo *d_draws = nullptr;
    uint32_t draw_cap = 0, attr_cap = 0;
    LargeItem *d_large = nullptr;
    uint32_t *d_out = nullptr;
    TexInfo *d_textab = nullptr; // [RZ_MAX_TEXTURES]
    unsigned long long *d_cnt_backup = nullptr;
    float *d_dbg_depth = nullptr;
    uint32_t *d_dbg_color = nullptr, *d_dbg_owner = nullptr;
    unsigned long long *d_dbg_time = nullptr;
    float4 *d_vtx = nullptr; // vertex-stage scratch [vert_cap][2], sized for the largest mesh
    uint32_t vert_cap = 0;
    uint32_t rec_cap = 0, large_cap = 0;
    bool debug = false;
    bool use_direct = true; // tile-kernel instantiation with the large-item path (see enqueue_frame)
    uint32_t last_n_large = 0; // large-triangle binning work items of the last frame the host has seen (sizes that kernel's grid)
    // The last frame issued through an async entry point, kept so that rz_sync can grow the device buffers and replay it
    // when it turns out to have overflowed them (an async call cannot know: the cursors live on the device).
    struct AsyncFrame {
        bool valid = false;
        std::vector<DrawCmd> draws;
        uint32_t *out_base = nullptr;
        uint32_t *out_host = nullptr; // rz_framebuffer_host_async: destination of the D2H copy
        int ring = 0;                 //   ... and the output buffer it ran in
        size_t staging_used = 0;
        int parity = 0;
    } last_async;
    uint32_t async_pending = 0;       // async frames issued since the last rz_sync / rz_framebuffer

    FrameState *h_state = nullptr; // pinned
    cudaEvent_t ev[4] = {nullptr, nullptr, nullptr, nullptr};
    rz_timings_t timings = {0, 0, 0, 0};
    uint64_t launches = 0;
    int num_sms = 148;
    int sticky = RZ_OK;
    std::string err;
};

static thread_local std::string g_err;

// Live contexts: rz_mesh_destroy has to drop a ctx's remembered async frame when it references the mesh (the replay of
// rz_sync must never touch a destroyed mesh), and a mesh may outlive its ctx.
static std::mutex g_ctx_mutex;
static std::vector<rz_ctx *> g_ctxs;

// Launch with programmatic stream serialization (PDL): see pdl_wait() in rz_exact.cuh.
template <typename... KArgs, typename... Args>
static cudaError_t launch_pdl(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st, Args... args) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = grid;
    cfg.blockDim = block;
    cfg.dynamicSmemBytes = smem;
    cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    return cudaLaunchKernelEx(&cfg, kernel, KArgs(args)...);
}

static int fail(rz_ctx *ctx, int code, const char *fmt, ...) {
    char buf[512];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof buf, fmt, ap);
    va_end(ap);
    if (ctx) ctx->err = buf;
    g_err = buf;
    return code;
}

#define CU(ctx, call)                                                                                   \
    do {                                                                                                \
        cudaError_t e_ = (call);                                                                        \
        if (e_ != cudaSuccess) {                                                                        \
            if (ctx) (ctx)->sticky = RZ_E_CUDA;                                                         \
            return fail(ctx, e_ == cudaErrorMemoryAllocation ? RZ_E_NOMEM : RZ_E_CUDA, "%s: %s", #call, \
                        cudaGetErrorString(e_));                                                        \
        }                                                                                               \
    } while (0)

// Mat4 * Mat4 exactly as math/matrix.rs:56-79 (dot = sequential sum from 0.0, vector.rs:17-23).
// This file is compiled with -Xcompiler -ffp-contract=off so the host never fuses either.
static void mat4_mul(const float *A, const float *B, float *R) {
    float out[16];
    for (int i = 0; i < 4; i++)
        for (int j = 0; j < 4; j++) {
            volatile float s = 0.0f;
            for (int k = 0; k < 4; k++) {
                volatile float p = A[i * 4 + k] * B[k * 4 + j];
                s = s + p;
            }
            out[i * 4 + j] = s;
        }
    memcpy(R, out, sizeof out);
}

// FrameState, then tile_count[tiles] (zeroed every frame), then busy[ORDER_BUCKETS][tiles] (not zeroed; 64-bit
// entries, so the region starts on the next 16-byte boundary)
static size_t zeroed_state_bytes(const rz_ctx *c) { return sizeof(FrameState) + sizeof(uint32_t) * (size_t)c->tiles_x * c->tiles_y; }
static size_t busy_offset(const rz_ctx *c) { return (zeroed_state_bytes(c) + 15) / 16 * 16; }
static size_t state_bytes(const rz_ctx *c) { return busy_offset(c) + sizeof(uint4) * ORDER_BUCKETS * (size_t)c->tiles_x * c->tiles_y; }

static int free_frame_buffers(rz_ctx *c) {
    cudaFree(c->d_bins); c->d_bins = nullptr;
    cudaFree(c->d_tile_bin); c->d_tile_bin = nullptr;
    cudaFree(c->d_recs); c->d_recs = nullptr;
    cudaFree(c->d_shade); c->d_shade = nullptr;
    cudaFree(c->d_clipq); c->d_clipq = nullptr;
    cudaFree(c->d_attrs); c->d_attrs = nullptr;
    cudaFree(c->d_draws); c->d_draws = nullptr;
    cudaFree(c->d_large); c->d_large = nullptr;
    cudaFree(c->d_vtx); c->d_vtx = nullptr;
    return RZ_OK;
}

static int ensure_vertex_scratch(rz_ctx *c, uint32_t nv) {
    if (nv <= c->vert_cap) return RZ_OK;
    cudaFree(c->d_vtx);
    c->d_vtx = nullptr; c->vert_cap = 0;
    CU(c, cudaMalloc(&c->d_vtx, (size_t)nv * 2 * sizeof(float4)));
    c->vert_cap = nv;
    return RZ_OK;
}

// Tile bins.  Every tile owns a slice {off, cap} of one entry array, planned on the host: at first a uniform small
// capacity, after an overflow from the tile counts of the frame that overflowed (the counters keep counting past the
// capacity).  Every tile gets at least `floor` entries -- the longest list of the frame + 25 %, as long as that
// uniform layout stays inside BIN_BUDGET bytes, so a moving object does not overflow the tiles it moves into -- and
// 1.5 x its own count beyond that: memory is O(total entries), one hot tile no longer sizes every bin.
static const uint64_t BIN_BUDGET = 512ull << 20;
static int plan_bins(rz_ctx *c, const uint32_t *counts) {
    const size_t tiles = (size_t)c->tiles_x * c->tiles_y;
    std::vector<TileBin> plan(tiles);
    const uint64_t budget_per_tile = std::max<uint64_t>(32, BIN_BUDGET / sizeof(uint4) / tiles);
    uint64_t floor_cap = std::max<uint64_t>(c->bin_floor, std::min<uint64_t>(256, budget_per_tile));
    if (counts) {
        uint32_t mx = 0;
        for (size_t t = 0; t < tiles; t++) mx = std::max(mx, counts[t]);
        floor_cap = std::max<uint64_t>(floor_cap, std::min<uint64_t>((uint64_t)mx * 5 / 4 + 64, budget_per_tile));
    }
    c->bin_floor = floor_cap;
    const bool have_old = c->bin_plan.size() == tiles;
    uint64_t off = 0;
    for (size_t t = 0; t < tiles; t++) {
        uint64_t cap = floor_cap;
        if (counts) cap = std::max<uint64_t>(cap, (uint64_t)counts[t] * 3 / 2 + 64);
        if (have_old) cap = std::max<uint64_t>(cap, c->bin_plan[t].cap); // never shrink: frames that alternate converge
        if (off + cap > 0xFFFFFFFFull) return fail(c, RZ_E_NOMEM, "the tile bins of this frame need more than 2^32 entries");
        plan[t].off = (uint32_t)off;
        plan[t].cap = (uint32_t)cap;
        off += cap;
    }
    if (off > c->bins_cap) {
        cudaFree(c->d_bins); c->d_bins = nullptr;
        c->bins_cap = 0;
        CU(c, cudaMalloc(&c->d_bins, off * sizeof(uint4)));
        c->bins_cap = off;
    }
    if (!c->d_tile_bin) CU(c, cudaMalloc(&c->d_tile_bin, tiles * sizeof(TileBin)));
    CU(c, cudaMemcpyAsync(c->d_tile_bin, plan.data(), tiles * sizeof(TileBin), cudaMemcpyHostToDevice, c->stream));
    CU(c, cudaStreamSynchronize(c->stream)); // pageable source: complete before it is touched again
    c->bin_plan.swap(plan);
    c->bins_planned = true;
    return RZ_OK;
}

static int ensure_capacity(rz_ctx *c, uint32_t rec_cap, uint32_t large_cap, uint32_t attr_cap) {
    if (rec_cap > c->rec_cap) {
        if (rec_cap > ENTRY_REC_MASK) return fail(c, RZ_E_NOMEM, "a frame may hold at most 2^29 triangle records");
        cudaFree(c->d_recs); c->d_recs = nullptr;
        cudaFree(c->d_shade); c->d_shade = nullptr;
        cudaFree(c->d_clipq); c->d_clipq = nullptr;
        c->rec_cap = 0;
        CU(c, cudaMalloc(&c->d_recs, (size_t)rec_cap * sizeof(RasterRec)));
        CU(c, cudaMalloc(&c->d_shade, (size_t)rec_cap * sizeof(ShadeRec)));
        CU(c, cudaMalloc(&c->d_clipq, (size_t)rec_cap * sizeof(unsigned long long)));
        c->rec_cap = rec_cap;
    }
    if (attr_cap > c->attr_cap) {
        cudaFree(c->d_attrs); c->d_attrs = nullptr;
        c->attr_cap = 0;
        CU(c, cudaMalloc(&c->d_attrs, (size_t)attr_cap * sizeof(AttrRec)));
        c->attr_cap = attr_cap;
    }
    if (large_cap > c->large_cap) {
        cudaFree(c->d_large); c->d_large = nullptr;
        CU(c, cudaMalloc(&c->d_large, (size_t)large_cap * sizeof(LargeItem)));
        c->large_cap = large_cap;
    }
    return RZ_OK;
}

extern "C" {

const char *rz_version(void) { return "rusterizer_b200 0.1 (sm_100a)"; }
uint32_t rz_tile_width(void) { return TW; }
uint32_t rz_tile_height(void) { return TH; }

const char *rz_last_error(rz_ctx *ctx) { return ctx ? ctx->err.c_str() : g_err.c_str(); }

int rz_create(int device, uint32_t width, uint32_t height, rz_ctx **out) {
    if (!out || width == 0 || height == 0 || width > 65535u || height > 65535u)
        return fail(nullptr, RZ_E_INVALID, "rz_create: bad arguments (width/height must be 1..65535)");
    *out = nullptr;
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0)
        return fail(nullptr, RZ_E_NO_DEVICE, "rz_create: no CUDA device (this library has no CPU fallback)");
    if (device < 0 || device >= ndev) return fail(nullptr, RZ_E_NO_DEVICE, "rz_create: device %d out of range", device);
    rz_ctx *c = new (std::nothrow) rz_ctx();
    if (!c) return fail(nullptr, RZ_E_NOMEM, "rz_create: out of host memory");
    c->device = device;
    c->W = width; c->H = height;
    c->tiles_x = (width + TW - 1) / TW;
    c->tiles_y = (height + TH - 1) / TH;
    c->row_begin = 0; c->row_end = height;
    c->sc_x1 = width; c->sc_y1 = height;
    static const float ident[16] = {1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1}; // uniform.rs:18-27
    memcpy(c->world, ident, 64); memcpy(c->view, ident, 64); memcpy(c->proj, ident, 64);
#define CU_NEW(call)                                                                       \
    do {                                                                                   \
        cudaError_t e_ = (call);                                                           \
        if (e_ != cudaSuccess) {                                                           \
            int rc_ = fail(nullptr, RZ_E_CUDA, "%s: %s", #call, cudaGetErrorString(e_));  \
            rz_destroy(c);                                                                 \
            return rc_;                                                                    \
        }                                                                                  \
    } while (0)
    CU_NEW(cudaSetDevice(device));
    cudaDeviceProp prop;
    CU_NEW(cudaGetDeviceProperties(&prop, device));
    c->num_sms = prop.multiProcessorCount;
    CU_NEW(cudaStreamCreateWithFlags(&c->own_stream, cudaStreamNonBlocking));
    c->stream = c->own_stream;
    CU_NEW(cudaMalloc(&c->d_state, state_bytes(c)));
    CU_NEW(cudaMemset(c->d_state, 0, state_bytes(c)));
    CU_NEW(cudaMalloc(&c->d_out, (size_t)width * height * sizeof(uint32_t)));
    CU_NEW(cudaMemset(c->d_out, 0, (size_t)width * height * sizeof(uint32_t))); // rows a screen-space shard does not own are never written
    CU_NEW(cudaMalloc(&c->d_cnt_backup, sizeof(unsigned long long) * 16 * CNT_STRIPES));
    CU_NEW(cudaHostAlloc(&c->h_state, sizeof(FrameState), cudaHostAllocDefault));
    for (int i = 0; i < 4; i++) CU_NEW(cudaEventCreate(&c->ev[i]));
    CU_NEW(cudaStreamCreateWithFlags(&c->up_stream, cudaStreamNonBlocking));
    CU_NEW(cudaStreamCreateWithFlags(&c->down_stream, cudaStreamNonBlocking));
    CU_NEW(cudaEventCreateWithFlags(&c->ev_uploaded, cudaEventDisableTiming));
    for (int i = 0; i < 2; i++) {
        CU_NEW(cudaEventCreateWithFlags(&c->ev_consumed[i], cudaEventDisableTiming));
        CU_NEW(cudaEventCreateWithFlags(&c->ev_frame_done[i], cudaEventDisableTiming));
        CU_NEW(cudaEventCreateWithFlags(&c->ev_d2h_done[i], cudaEventDisableTiming));
    }
    c->d_out_ring[0] = c->d_out;
    {
        struct { const void *f; size_t smem; } k[] = {
            {(const void *)tile_kernel<false, false, false>, sizeof(TileSmemT<false>)}, {(const void *)tile_kernel<false, false, true>, sizeof(TileSmemT<false>)},
            {(const void *)tile_kernel<false, true, false>, sizeof(TileSmemT<false>)},  {(const void *)tile_kernel<false, true, true>, sizeof(TileSmemT<false>)},
            {(const void *)tile_kernel<true, false, false>, sizeof(TileSmemT<true>)},   {(const void *)tile_kernel<true, false, true>, sizeof(TileSmemT<true>)},
            {(const void *)tile_kernel<true, true, false>, sizeof(TileSmemT<true>)},    {(const void *)tile_kernel<true, true, true>, sizeof(TileSmemT<true>)}};
        for (auto &e : k) CU_NEW(cudaFuncSetAttribute(e.f, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)e.smem));
        struct { const void *f; size_t smem; } km[] = {
            {(const void *)msaa_tile_kernel<1, false>, sizeof(MsaaSmemT<false>)}, {(const void *)msaa_tile_kernel<1, true>, sizeof(MsaaSmemT<true>)},
            {(const void *)msaa_tile_kernel<2, false>, sizeof(MsaaSmemT<false>)}, {(const void *)msaa_tile_kernel<2, true>, sizeof(MsaaSmemT<true>)},
            {(const void *)msaa_tile_kernel<8, false>, sizeof(MsaaSmemT<false>)}, {(const void *)msaa_tile_kernel<8, true>, sizeof(MsaaSmemT<true>)}};
        for (auto &e : km) CU_NEW(cudaFuncSetAttribute(e.f, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)e.smem));
    }
#undef CU_NEW
    {
        std::lock_guard<std::mutex> lock(g_ctx_mutex);
        g_ctxs.push_back(c);
    }
    *out = c;
    return RZ_OK;
}

void rz_destroy(rz_ctx *c) {
    if (!c) return;
    {
        std::lock_guard<std::mutex> lock(g_ctx_mutex);
        g_ctxs.erase(std::remove(g_ctxs.begin(), g_ctxs.end(), c), g_ctxs.end());
    }
    cudaSetDevice(c->device);
    if (c->stream) cudaStreamSynchronize(c->stream);
    if (c->up_stream) cudaStreamSynchronize(c->up_stream);
    if (c->down_stream) cudaStreamSynchronize(c->down_stream);
    free_frame_buffers(c);
    cudaFree(c->d_state); cudaFree(c->d_out); cudaFree(c->d_out_ring[1]); cudaFree(c->d_cnt_backup);
    cudaFree(c->d_dbg_depth); cudaFree(c->d_dbg_color); cudaFree(c->d_dbg_owner); cudaFree(c->d_dbg_time);
    for (auto &t : c->textures) cudaFree(t.d_data);
    cudaFree(c->d_textab);
    for (int p = 0; p < 2; p++)
        for (auto *m : c->staging[p]) rz_mesh_destroy(m);
    if (c->h_state) cudaFreeHost(c->h_state);
    for (int i = 0; i < 4; i++)
        if (c->ev[i]) cudaEventDestroy(c->ev[i]);
    if (c->ev_uploaded) cudaEventDestroy(c->ev_uploaded);
    for (int i = 0; i < 2; i++) {
        if (c->ev_consumed[i]) cudaEventDestroy(c->ev_consumed[i]);
        if (c->ev_frame_done[i]) cudaEventDestroy(c->ev_frame_done[i]);
        if (c->ev_d2h_done[i]) cudaEventDestroy(c->ev_d2h_done[i]);
    }
    if (c->up_stream) cudaStreamDestroy(c->up_stream);
    if (c->down_stream) cudaStreamDestroy(c->down_stream);
    if (c->own_stream) cudaStreamDestroy(c->own_stream);
    delete c;
}

int rz_set_stream(rz_ctx *c, void *cuda_stream) {
    if (!c) return RZ_E_INVALID;
    CU(c, cudaSetDevice(c->device));
    CU(c, cudaStreamSynchronize(c->stream));
    c->stream = cuda_stream ? (cudaStream_t)cuda_stream : c->own_stream;
    return RZ_OK;
}

int rz_bind_texture(rz_ctx *c, uint32_t index, const uint8_t *texels, uint32_t width, uint32_t height,
                    uint32_t texel_width) {
    if (!c || !texels || width == 0 || height == 0) return fail(c, RZ_E_INVALID, "rz_bind_texture: bad arguments");
    if (texel_width != 3 && texel_width != 4)
        return fail(c, RZ_E_INVALID, "rz_bind_texture: texel_width must be 3 or 4 (texture.rs:48)");
    if (index != c->textures.size())
        return fail(c, RZ_E_TEXTURE, "rz_bind_texture: index %u != number of bound textures %zu (uniform.rs:31)", index,
                    c->textures.size());
    if (c->textures.size() >= RZ_MAX_TEXTURES)
        return fail(c, RZ_E_TEXTURE, "rz_bind_texture: at most %d textures can be bound", RZ_MAX_TEXTURES);
    CU(c, cudaSetDevice(c->device));
    if (!c->d_textab) CU(c, cudaMalloc(&c->d_textab, sizeof(TexInfo) * RZ_MAX_TEXTURES));
    Texture t;
    t.w = width; t.h = height; t.tw = texel_width;
    t.len = (size_t)width * height * texel_width;
    CU(c, cudaMalloc(&t.d_data, t.len));
    TexInfo ti;
    ti.data = t.d_data; ti.len = t.len; ti.w = t.w; ti.h = t.h; ti.tw = t.tw;
    ti.bound = 1u | ((t.tw == 4 && t.len < (1ull << 32)) ? 2u : 0u); // bit 1: RGBA8 with 32-bit byte offsets (rz_tile.cuh)
    cudaError_t e = cudaMemcpyAsync(t.d_data, texels, t.len, cudaMemcpyHostToDevice, c->stream);
    if (e == cudaSuccess) e = cudaMemcpyAsync(c->d_textab + c->textures.size(), &ti, sizeof ti, cudaMemcpyHostToDevice, c->stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(c->stream);
    if (e != cudaSuccess) { // nothing was bound: give the texels back
        cudaFree(t.d_data);
        c->sticky = RZ_E_CUDA;
        return fail(c, RZ_E_CUDA, "rz_bind_texture: %s", cudaGetErrorString(e));
    }
    c->textures.push_back(t);
    return RZ_OK;
}

int rz_write_block(rz_ctx *c, const float *world, const float *view, const float *projection) {
    if (!c) return RZ_E_INVALID;
    if (world) memcpy(c->world, world, 64);
    if (view) memcpy(c->view, view, 64);
    if (projection) memcpy(c->proj, projection, 64);
    return RZ_OK;
}

int rz_read_block(rz_ctx *c, float *world, float *view, float *projection) {
    if (!c) return RZ_E_INVALID;
    if (world) memcpy(world, c->world, 64);
    if (view) memcpy(view, c->view, 64);
    if (projection) memcpy(projection, c->proj, 64);
    return RZ_OK;
}

static int mesh_upload(rz_ctx *c, rz_mesh *m, const float *pos, const float *attr, uint32_t nv, const uint32_t *idx,
                       uint64_t n_idx, cudaStream_t st) {
    const size_t bp = (size_t)nv * 12, ba = (size_t)nv * 24, bi = (size_t)n_idx * 4;
    if (bp > m->cap_pos) { cudaFree(m->d_pos); m->d_pos = nullptr; m->cap_pos = 0; CU(c, cudaMalloc(&m->d_pos, bp)); m->cap_pos = bp; }
    if (ba > m->cap_attr) { cudaFree(m->d_attr); m->d_attr = nullptr; m->cap_attr = 0; CU(c, cudaMalloc(&m->d_attr, ba)); m->cap_attr = ba; }
    if (bi > m->cap_idx) { cudaFree(m->d_idx); m->d_idx = nullptr; m->cap_idx = 0; CU(c, cudaMalloc(&m->d_idx, bi)); m->cap_idx = bi; }
    if (bp) CU(c, cudaMemcpyAsync(m->d_pos, pos, bp, cudaMemcpyHostToDevice, st));
    if (ba) CU(c, cudaMemcpyAsync(m->d_attr, attr, ba, cudaMemcpyHostToDevice, st));
    if (bi) CU(c, cudaMemcpyAsync(m->d_idx, idx, bi, cudaMemcpyHostToDevice, st));
    m->nv = nv;
    m->n_idx = n_idx;
    return RZ_OK;
}

int rz_mesh_create(rz_ctx *c, const float *positions, const float *attributes, uint32_t nv, const uint32_t *indices,
                   uint64_t n_idx, rz_mesh **out) {
    if (!c || !out) return RZ_E_INVALID;
    *out = nullptr;
    if ((nv && (!positions || !attributes)) || (n_idx && !indices) || n_idx % 3 != 0 || n_idx / 3 > 0x1FFFFFFFull)
        return fail(c, RZ_E_INVALID, "rz_mesh_create: bad arguments (n_idx must be a multiple of 3, < 2^29 triangles)");
    CU(c, cudaSetDevice(c->device));
    rz_mesh *m = new (std::nothrow) rz_mesh();
    if (!m) return fail(c, RZ_E_NOMEM, "rz_mesh_create: out of host memory");
    m->ctx = c;
    m->device = c->device;
    int rc = mesh_upload(c, m, positions, attributes, nv, indices, n_idx, c->stream);
    if (rc == RZ_OK && cudaStreamSynchronize(c->stream) != cudaSuccess) rc = fail(c, RZ_E_CUDA, "rz_mesh_create: sync failed");
    if (rc != RZ_OK) {
        rz_mesh_destroy(m);
        return rc;
    }
    *out = m;
    return RZ_OK;
}

void rz_mesh_destroy(rz_mesh *m) {
    if (!m) return;
    {   // a remembered async frame that draws this mesh can no longer be replayed
        std::lock_guard<std::mutex> lock(g_ctx_mutex);
        for (rz_ctx *c : g_ctxs)
            for (const DrawCmd &d : c->last_async.draws)
                if (d.mesh == m) {
                    c->last_async.valid = false;
                    c->last_async.draws.clear();
                    break;
                }
    }
    cudaSetDevice(m->device);
    cudaFree(m->d_pos); cudaFree(m->d_attr); cudaFree(m->d_idx);
    delete m;
}

static int record_draw(rz_ctx *c, const rz_mesh *mesh, uint32_t vs_id, uint32_t fs_id) {
    if (vs_id != RZ_VS_MVP) return fail(c, RZ_E_INVALID, "rz_render: unknown vertex shader id %u", vs_id);
    const uint32_t shader = fs_id & 0xFFu, tex = (fs_id >> 8) & 0xFFu;
    if (shader > RZ_FS_TEXTURE_BLEND || (fs_id >> 16) != 0)
        return fail(c, RZ_E_INVALID, "rz_render: unknown fragment shader id 0x%x", fs_id);
    const bool textured = shader == RZ_FS_TEXTURE || shader == RZ_FS_TEXTURE_BLEND;
    if (tex != 0 && !textured) return fail(c, RZ_E_INVALID, "rz_render: fragment shader %u takes no texture index", shader);
    if (textured && tex >= c->textures.size())
        return fail(c, RZ_E_TEXTURE, "rz_render: the fragment shader reads texture %u but %zu are bound (uniform.rs:36)", tex,
                    c->textures.size());
    if (c->draws.size() >= (1u << 24)) return fail(c, RZ_E_INVALID, "rz_render: more than 2^24 draws in one frame");
    DrawCmd d;
    d.mesh = mesh;
    d.fs = fs_id;
    float pv[16];
    mat4_mul(c->proj, c->view, pv);   // projection * view
    mat4_mul(pv, c->world, d.M);      // (projection * view) * world   (main.rs:147-152, left-assoc)
    c->draws.push_back(d);
    return RZ_OK;
}

int rz_render(rz_ctx *c, const rz_mesh *mesh, uint32_t vs_id, uint32_t fs_id) {
    if (!c || !mesh) return RZ_E_INVALID;
    if (mesh->ctx != c) return fail(c, RZ_E_INVALID, "rz_render: mesh belongs to another ctx");
    return record_draw(c, mesh, vs_id, fs_id);
}

int rz_render_host(rz_ctx *c, const float *positions, const float *attributes, uint32_t nv, const uint32_t *indices,
                   uint64_t n_idx, uint32_t vs_id, uint32_t fs_id) {
    if (!c) return RZ_E_INVALID;
    if ((nv && (!positions || !attributes)) || (n_idx && !indices) || n_idx % 3 != 0 || n_idx / 3 > 0x1FFFFFFFull)
        return fail(c, RZ_E_INVALID, "rz_render_host: bad arguments");
    CU(c, cudaSetDevice(c->device));
    auto &staging = c->staging[c->parity];
    if (c->staging_used == staging.size()) {
        rz_mesh *m = new (std::nothrow) rz_mesh();
        if (!m) return fail(c, RZ_E_NOMEM, "rz_render_host: out of host memory");
        m->ctx = c;
        m->device = c->device;
        staging.push_back(m);
    }
    // this staging set was last read two frames ago: wait (on the upload stream only) for that frame
    if (c->staging_used == 0) CU(c, cudaStreamWaitEvent(c->up_stream, c->ev_consumed[c->parity], 0));
    rz_mesh *m = staging[c->staging_used];
    int rc = mesh_upload(c, m, positions, attributes, nv, indices, n_idx, c->up_stream);
    if (rc != RZ_OK) return rc;
    rc = record_draw(c, m, vs_id, fs_id);
    if (rc == RZ_OK) c->staging_used++;
    return rc;
}

static FrameParams make_params(rz_ctx *c, uint32_t *out_base) {
    FrameParams P;
    memset(&P, 0, sizeof P);
    P.W = c->W; P.H = c->H;
    P.tiles_x = c->tiles_x; P.tiles_y = c->tiles_y;
    P.row_begin = c->row_begin; P.row_end = c->row_end;
    P.il_band = c->il_band; P.il_rank = c->il_rank; P.il_world = c->il_world;
    P.scissor = make_uint4(c->sc_x0, c->sc_y0, c->sc_x1, c->sc_y1);
    P.msaa = c->msaa; P.guard = c->guard;
    P.ty_begin = c->row_begin / TH;
    P.ty_end = (c->row_end + TH - 1) / TH;
    P.rec_cap = c->rec_cap; P.large_cap = c->large_cap;
    P.fs = reinterpret_cast<FrameState *>(c->d_state);
    P.tile_count = reinterpret_cast<uint32_t *>(c->d_state + sizeof(FrameState));
    P.busy = reinterpret_cast<uint4 *>(c->d_state + busy_offset(c));
    P.bins = c->d_bins; P.tile_bin = c->d_tile_bin; P.recs = c->d_recs; P.shade = c->d_shade; P.clipq = c->d_clipq; P.attrs = c->d_attrs; P.large = c->d_large;
    P.draws = c->d_draws; P.attr_cap = c->attr_cap;
    P.out = out_base;
    P.spread_clears = (out_base != c->d_out && out_base != c->d_out_ring[1]) ? 1u : 0u;
    if (c->debug) {
        P.dbg_depth = c->d_dbg_depth; P.dbg_color = c->d_dbg_color; P.dbg_owner = c->d_dbg_owner;
        P.dbg_tile_time = c->d_dbg_time;
    }
    if (!c->textures.empty()) {
        const Texture &t = c->textures[0];
        P.tex_table = c->d_textab;
        P.tex0.data = t.d_data; P.tex0.len = t.len; P.tex0.w = t.w; P.tex0.h = t.h; P.tex0.tw = t.tw;
        P.tex0.bound = 1u | ((t.tw == 4 && t.len < (1ull << 32)) ? 2u : 0u);
    }
    return P;
}

// NVTX ranges around the launches of each stage (host side: they bracket the enqueue, tools correlate the kernels).
struct NvtxRange {
    explicit NvtxRange(const char *name) { nvtxRangePushA(name); }
    ~NvtxRange() { nvtxRangePop(); }
};

// Enqueue one whole frame on the ctx stream.  timed: record stage events.
static int enqueue_frame(rz_ctx *c, uint32_t *out_base, bool timed) {
    uint64_t total_tris = 0;
    uint32_t max_nv = 0;
    for (auto &d : c->draws) {
        total_tris += d.mesh->n_idx / 3;
        max_nv = std::max(max_nv, d.mesh->nv);
    }
    {
        int rc = ensure_vertex_scratch(c, max_nv);
        if (rc != RZ_OK) return rc;
    }
    if (total_tris > 0x1FFFFFFFull) return fail(c, RZ_E_INVALID, "frame has more than 2^29 triangles");
    {
        // records are allocated from REC_STRIPES arenas filled round-robin by CTA: leave 25 % headroom
        uint32_t want_rec = std::max<uint64_t>(c->rec_cap, (total_tris * 5 / 4 + 256 * REC_STRIPES + REC_STRIPES - 1) / REC_STRIPES * REC_STRIPES);
        uint32_t want_large = std::max<uint32_t>(c->large_cap, 1u << 16);
        uint32_t want_attr = std::max<uint32_t>(c->attr_cap, 1u << 14);
        int rc = ensure_capacity(c, want_rec, want_large, want_attr);
        if (rc != RZ_OK) return rc;
        if (!c->bins_planned && (rc = plan_bins(c, nullptr)) != RZ_OK) return rc;
        // per-draw table for the shading step
        if (c->draws.size() > c->draw_cap) {
            cudaFree(c->d_draws); c->d_draws = nullptr;
            c->draw_cap = 0;
            const uint32_t cap = (uint32_t)std::max<size_t>(c->draws.size() * 2, 64);
            CU(c, cudaMalloc(&c->d_draws, (size_t)cap * sizeof(DrawInfo)));
            c->draw_cap = cap;
        }
    }
    FrameParams P = make_params(c, out_base);
    cudaStream_t st = c->stream;
    (void)cudaGetLastError(); // do not blame this frame for a stale, non-sticky error of an earlier call
    if (c->staging_used) { // host-mesh draws: the frame starts when their uploads have landed
        CU(c, cudaEventRecord(c->ev_uploaded, c->up_stream));
        CU(c, cudaStreamWaitEvent(st, c->ev_uploaded, 0));
    }
    const size_t off = offsetof(FrameState, n_clipq); // 16-byte aligned by construction
    const uint32_t n16 = (uint32_t)((zeroed_state_bytes(c) - off + 15) / 16);
    uint4 *zero16 = reinterpret_cast<uint4 *>(c->d_state + off);
    bool zeroed = false; // the first vertex kernel of the frame zeroes the frame state; frames without one do it here
    // ... and so do frames whose first vertex kernel is too small a grid for the job (a quad on an 8192 x 8192 target:
    // one CTA would zero 262144 tile counters by itself)
    uint32_t first_vertex_threads = 0;
    for (auto &d : c->draws)
        if (d.mesh->n_idx / 3 > 0) {
            first_vertex_threads = std::max(1u, (d.mesh->nv + NT * VERTEX_PER_THREAD - 1) / (NT * VERTEX_PER_THREAD)) * NT;
            break;
        }
    if (total_tris == 0 || n16 > 16u * first_vertex_threads) {
        CU(c, launch_pdl(frame_begin_kernel, dim3((n16 + NT - 1) / NT), dim3(NT), 0, st, zero16, n16));
        c->launches++;
        zeroed = true;
    }
    if (timed) CU(c, cudaEventRecord(c->ev[0], st));
    uint32_t tri_base = 0, draw_index = 0;
    NvtxRange frame_range("rz frame");
    nvtxRangePushA("rz geometry (vertex, geom, clip)");
    for (auto &d : c->draws) {
        const uint32_t nt = (uint32_t)(d.mesh->n_idx / 3);
        const uint32_t this_draw = draw_index++;
        if (nt == 0) continue;
        DrawParams D;
        D.pos = d.mesh->d_pos; D.attr = d.mesh->d_attr; D.idx = d.mesh->d_idx;
        D.nv = d.mesh->nv; D.nt = nt; D.tri_base = tri_base; D.fs = d.fs; D.draw = this_draw;
        memcpy(D.M, d.M, 64);
        D.vtx = c->d_vtx;
        CU(c, launch_pdl(vertex_kernel, dim3(std::max(1u, (D.nv + NT * VERTEX_PER_THREAD - 1) / (NT * VERTEX_PER_THREAD))), dim3(NT), 0, st, P, D, zeroed ? (uint4 *)nullptr : zero16,
                         zeroed ? 0u : n16));
        zeroed = true;
        static const int tune_geom = getenv("RZ_TUNE_GEOM_CTAS_PER_SM") ? atoi(getenv("RZ_TUNE_GEOM_CTAS_PER_SM")) : RZ_GEOM_MIN_CTAS;
        const uint32_t geom_ctas = tune_geom > 0 ? std::min<uint32_t>((nt + NT - 1) / NT, (uint32_t)c->num_sms * tune_geom) : (nt + NT - 1) / NT;
        CU(c, launch_pdl(geom_kernel, dim3(geom_ctas), dim3(NT), 0, st, P, D));
        c->launches += 2;
        tri_base += nt;
    }
    if (tri_base) { // triangles that straddle a clip plane, all draws of the frame (persistent grid)
        const uint32_t ctas = (uint32_t)std::min<uint64_t>((uint64_t)c->num_sms * 4, (tri_base + NT - 1) / NT);
        CU(c, launch_pdl(clip_kernel, dim3(ctas), dim3(NT), 0, st, P));
        c->launches++;
    }
    nvtxRangePop();
    if (timed) CU(c, cudaEventRecord(c->ev[1], st));
    nvtxRangePushA("rz binning (large_bin, order)");
    // two CTAs per SM keep the launch cheap for frames without large triangles; a frame like the last one with
    // thousands of (triangle, slab) items gets the full five (48 registers)
    CU(c, launch_pdl(large_bin_kernel, dim3(c->num_sms * (c->last_n_large > 4096u ? 5 : RZ_LARGE_BIN_CTAS)), dim3(NT), 0, st, P));
    c->launches++;
    const uint32_t n_tiles = P.tiles_x * (P.ty_end - P.ty_begin);
    if (n_tiles) {
        CU(c, launch_pdl(order_kernel, dim3((n_tiles + NT - 1) / NT), dim3(NT), 0, st, P));
        c->launches++;
    }
    nvtxRangePop();
    if (timed) CU(c, cudaEventRecord(c->ev[2], st));
    NvtxRange tile_range("rz tile raster + resolve");
    static const float tune_tile = getenv("RZ_TUNE_TILE_CTAS_PER_SM") ? (float)atof(getenv("RZ_TUNE_TILE_CTAS_PER_SM")) : (float)RZ_TILE_CTAS;
    const dim3 tile_grid(std::min<uint32_t>(n_tiles, c->debug ? (uint32_t)c->num_sms * 3u : (uint32_t)((float)c->num_sms * tune_tile))); // persistent CTAs, 4 per SM
    if (n_tiles && c->msaa != 4u) {
        // runtime sample counts other than the reference's 4: the generic pixel-parallel tile kernel (rz_msaa.cuh)
        const dim3 g(std::min<uint32_t>(n_tiles, (uint32_t)c->num_sms * 2u));
#define RZ_MSAA_LAUNCH(NS, D) CU(c, launch_pdl(msaa_tile_kernel<NS, D>, g, dim3(NT), sizeof(MsaaSmemT<D>), st, P))
        if (c->msaa == 1u) { if (c->debug) RZ_MSAA_LAUNCH(1, true); else RZ_MSAA_LAUNCH(1, false); }
        else if (c->msaa == 2u) { if (c->debug) RZ_MSAA_LAUNCH(2, true); else RZ_MSAA_LAUNCH(2, false); }
        else { if (c->debug) RZ_MSAA_LAUNCH(8, true); else RZ_MSAA_LAUNCH(8, false); }
#undef RZ_MSAA_LAUNCH
        c->launches++;
    } else if (n_tiles) {
        bool ext = false; // does any draw use the shader-registry extension (texture index != 0, TextureBlend)?
        for (auto &d : c->draws) ext = ext || (d.fs >> 8) != 0 || (d.fs & 0xFFu) == RZ_FS_TEXTURE_BLEND;
        // DIRECT: the pixel-parallel path for chunks of large items is only compiled into the instantiations used
        // when the last frame whose state the host has seen queued large triangles (a stale guess costs speed only)
        const bool dir = c->use_direct;
#define RZ_TILE_LAUNCH(D, E, R) CU(c, launch_pdl(tile_kernel<D, E, R>, tile_grid, dim3(NT), sizeof(TileSmemT<D>), st, P))
        if (c->debug) {
            if (ext) { if (dir) RZ_TILE_LAUNCH(true, true, true); else RZ_TILE_LAUNCH(true, true, false); }
            else     { if (dir) RZ_TILE_LAUNCH(true, false, true); else RZ_TILE_LAUNCH(true, false, false); }
        } else {
            if (ext) { if (dir) RZ_TILE_LAUNCH(false, true, true); else RZ_TILE_LAUNCH(false, true, false); }
            else     { if (dir) RZ_TILE_LAUNCH(false, false, true); else RZ_TILE_LAUNCH(false, false, false); }
        }
#undef RZ_TILE_LAUNCH
        c->launches++;
    }
    if (timed) CU(c, cudaEventRecord(c->ev[3], st));
    if (c->staging_used) CU(c, cudaEventRecord(c->ev_consumed[c->parity], st));
    CU(c, cudaGetLastError());
    return RZ_OK;
}

// Which tile-kernel instantiation the NEXT frame runs: the one with the large-item / short-list paths when the last frame
// whose state the host has seen queued large triangles or had mostly tiles with a handful of items (a stale guess only
// costs speed -- every instantiation is correct for every input).
static bool wants_direct(const FrameState *fs) {
    uint32_t busy = 0;
    for (int b = 0; b < ORDER_BUCKETS; b++) busy += fs->bucket_n[b];
    return fs->n_large > 0 || (busy > 0 && fs->n_few_tiles * 2u >= busy);
}

static void end_frame(rz_ctx *c) {
    c->draws.clear();
    if (c->staging_used) c->parity ^= 1;
    c->staging_used = 0;
}

static int map_err_flags(rz_ctx *c, uint32_t flags) {
    if (flags & ERR_INDEX) return fail(c, RZ_E_INDEX, "a mesh index is >= nv (the reference panics at render.rs:83-87)");
    if (flags & (ERR_REC_OVF | ERR_BIN_OVF | ERR_LARGE_OVF | ERR_ATTR_OVF))
        return fail(c, RZ_E_CAPACITY, "an async frame outgrew its device buffers (flags 0x%x); re-issue via rz_framebuffer()", flags);
    return RZ_OK;
}

// Grow the device buffers an overflowed frame asked for.  c->h_state holds the frame's final state: the cursors keep
// counting past the capacities, so one look says how much is needed.
static int grow_after_overflow(rz_ctx *c, uint32_t flags) {
    cudaStream_t st = c->stream;
    uint32_t want_rec = c->rec_cap, want_large = c->large_cap, want_attr = c->attr_cap;
    if (flags & ERR_ATTR_OVF) want_attr = std::max<uint64_t>((uint64_t)c->h_state->n_clip_attr * 5 / 4 + 1024, (uint64_t)c->attr_cap * 2);
    if (flags & ERR_REC_OVF) {
        uint32_t mx = 0; // the fullest stripe decides
        for (int i = 0; i < REC_STRIPES; i++) mx = std::max(mx, c->h_state->rec_cursor[i]);
        want_rec = std::max<uint64_t>(((uint64_t)mx * 5 / 4 + 64) * REC_STRIPES, (uint64_t)c->rec_cap * 3 / 2);
        want_rec = std::max<uint64_t>(want_rec, (uint64_t)c->h_state->n_clipq); // the clip queue shares the capacity
    }
    if (flags & ERR_LARGE_OVF) want_large = std::max<uint64_t>((uint64_t)c->h_state->n_large * 5 / 4 + 1024, (uint64_t)c->large_cap * 2);
    if (flags & ERR_BIN_OVF) {
        std::vector<uint32_t> counts((size_t)c->tiles_x * c->tiles_y);
        CU(c, cudaMemcpyAsync(counts.data(), c->d_state + sizeof(FrameState), counts.size() * 4, cudaMemcpyDeviceToHost, st));
        CU(c, cudaStreamSynchronize(st));
        int rc = plan_bins(c, counts.data());
        if (rc != RZ_OK) return rc;
    }
    return ensure_capacity(c, want_rec, want_large, want_attr);
}

int rz_framebuffer(rz_ctx *c, uint32_t *out_host, const uint32_t **out_device) {
    if (!c) return RZ_E_INVALID;
    if (c->sticky != RZ_OK) return c->sticky;
    CU(c, cudaSetDevice(c->device));
    cudaStream_t st = c->stream;
    FrameState *dfs = reinterpret_cast<FrameState *>(c->d_state);
    // surface errors of earlier async frames first
    CU(c, cudaMemcpyAsync(c->h_state, dfs, sizeof(FrameState), cudaMemcpyDeviceToHost, st));
    CU(c, cudaStreamSynchronize(st));
    if (c->h_state->err) {
        // an error of an EARLIER async frame (nobody called rz_sync): report it, keep the draws recorded for the current
        // frame -- the caller may simply call rz_framebuffer() again
        uint32_t flags = c->h_state->err;
        CU(c, cudaMemsetAsync(&dfs->err, 0, sizeof(uint32_t), st));
        c->async_pending = 0;
        return map_err_flags(c, flags);
    }
    c->async_pending = 0;
    CU(c, cudaMemcpyAsync(c->d_cnt_backup, dfs->counters, sizeof(unsigned long long) * 16 * CNT_STRIPES, cudaMemcpyDeviceToDevice, st));
    int rc = RZ_OK;
    for (int attempt = 0; attempt < 8; attempt++) {
        rc = enqueue_frame(c, c->d_out, true);
        if (rc != RZ_OK) break;
        CU(c, cudaMemcpyAsync(c->h_state, dfs, sizeof(FrameState), cudaMemcpyDeviceToHost, st));
        CU(c, cudaStreamSynchronize(st));
        const uint32_t flags = c->h_state->err;
        c->use_direct = wants_direct(c->h_state);
        c->last_n_large = c->h_state->n_large;
        if (!flags) break;
        CU(c, cudaMemsetAsync(&dfs->err, 0, sizeof(uint32_t), st));
        if (flags & ERR_INDEX) {
            rc = map_err_flags(c, flags);
            break;
        }
        // grow what overflowed (the cursors kept counting past the capacity) and replay the frame
        CU(c, cudaMemcpyAsync(dfs->counters, c->d_cnt_backup, sizeof(unsigned long long) * 16 * CNT_STRIPES, cudaMemcpyDeviceToDevice, st));
        rc = grow_after_overflow(c, flags);
        if (rc != RZ_OK) break;
        if (attempt == 7) rc = fail(c, RZ_E_CAPACITY, "frame still overflows after 8 growth attempts");
    }
    end_frame(c);
    if (rc != RZ_OK) return rc;
    float ms;
    if (cudaEventElapsedTime(&ms, c->ev[0], c->ev[1]) == cudaSuccess) c->timings.geometry_ms = ms;
    if (cudaEventElapsedTime(&ms, c->ev[1], c->ev[2]) == cudaSuccess) c->timings.bin_ms = ms;
    if (cudaEventElapsedTime(&ms, c->ev[2], c->ev[3]) == cudaSuccess) c->timings.tile_ms = ms;
    if (cudaEventElapsedTime(&ms, c->ev[0], c->ev[3]) == cudaSuccess) c->timings.total_ms = ms;
    if (out_host) {
        CU(c, cudaMemcpyAsync(out_host, c->d_out, (size_t)c->W * c->H * 4, cudaMemcpyDeviceToHost, st));
        CU(c, cudaStreamSynchronize(st));
    }
    if (out_device) *out_device = c->d_out;
    return RZ_OK;
}

// the frame just enqueued by an async entry point, for rz_sync's overflow replay
static void remember_async(rz_ctx *c, uint32_t *out_base, uint32_t *out_host, int ring) {
    auto &a = c->last_async;
    a.valid = true;
    a.draws = c->draws;
    a.out_base = out_base; a.out_host = out_host; a.ring = ring;
    a.staging_used = c->staging_used; a.parity = c->parity;
    c->async_pending++;
}

int rz_framebuffer_async(rz_ctx *c, uint32_t *device_dst, const uint32_t **out_device) {
    if (!c) return RZ_E_INVALID;
    if (c->sticky != RZ_OK) return c->sticky;
    CU(c, cudaSetDevice(c->device));
    // an external destination holds only this ctx's rows [row_begin,row_end), starting at device_dst
    // (with interleaved bands the row range is the whole frame, so device_dst is the full image)
    uint32_t *base = device_dst ? device_dst - (size_t)c->row_begin * c->W : c->d_out;
    int rc = enqueue_frame(c, base, false);
    if (rc == RZ_OK) remember_async(c, base, nullptr, 0);
    end_frame(c);
    if (rc != RZ_OK) return rc;
    if (out_device) *out_device = device_dst ? device_dst : c->d_out;
    return RZ_OK;
}

// D2H copy of the image rendered into output buffer `p`, on the download stream
static int enqueue_download(rz_ctx *c, int p, uint32_t *out_host) {
    CU(c, cudaEventRecord(c->ev_frame_done[p], c->stream));
    CU(c, cudaStreamWaitEvent(c->down_stream, c->ev_frame_done[p], 0));
    const size_t off = (size_t)c->row_begin * c->W, cnt = (size_t)(c->row_end - c->row_begin) * c->W;
    CU(c, cudaMemcpyAsync(out_host + off, c->d_out_ring[p] + off, cnt * sizeof(uint32_t), cudaMemcpyDeviceToHost,
                          c->down_stream));
    CU(c, cudaEventRecord(c->ev_d2h_done[p], c->down_stream));
    return RZ_OK;
}

int rz_framebuffer_host_async(rz_ctx *c, uint32_t *out_host) {
    if (!c || !out_host) return RZ_E_INVALID;
    if (c->sticky != RZ_OK) return c->sticky;
    CU(c, cudaSetDevice(c->device));
    const int p = c->out_parity;
    if (!c->d_out_ring[p]) {
        CU(c, cudaMalloc(&c->d_out_ring[p], (size_t)c->W * c->H * sizeof(uint32_t)));
        CU(c, cudaMemsetAsync(c->d_out_ring[p], 0, (size_t)c->W * c->H * sizeof(uint32_t), c->stream));
    }
    // the image rendered into this buffer two frames ago must have left for the host
    CU(c, cudaStreamWaitEvent(c->stream, c->ev_d2h_done[p], 0));
    int rc = enqueue_frame(c, c->d_out_ring[p], false);
    if (rc == RZ_OK) remember_async(c, c->d_out_ring[p], out_host, p);
    end_frame(c);
    if (rc != RZ_OK) return rc;
    rc = enqueue_download(c, p, out_host);
    if (rc != RZ_OK) return rc;
    c->out_parity ^= 1;
    return RZ_OK;
}

int rz_discard_frame(rz_ctx *c) {
    if (!c) return RZ_E_INVALID;
    end_frame(c);
    return RZ_OK;
}

int rz_sync(rz_ctx *c) {
    if (!c) return RZ_E_INVALID;
    if (c->sticky != RZ_OK) return c->sticky;
    CU(c, cudaSetDevice(c->device));
    FrameState *dfs = reinterpret_cast<FrameState *>(c->d_state);
    CU(c, cudaMemcpyAsync(c->h_state, dfs, sizeof(FrameState), cudaMemcpyDeviceToHost, c->stream));
    CU(c, cudaStreamSynchronize(c->stream));
    CU(c, cudaStreamSynchronize(c->down_stream));
    c->use_direct = wants_direct(c->h_state);
    c->last_n_large = c->h_state->n_large;
    const uint32_t pending = c->async_pending;
    c->async_pending = 0;
    if (c->h_state->peer_timeout) {
        CU(c, cudaMemsetAsync(&dfs->peer_timeout, 0, sizeof(uint32_t), c->stream));
        return fail(c, RZ_E_PEER, "rz_wait_flags: a peer did not signal within the timeout");
    }
    if (c->h_state->err) {
        uint32_t flags = c->h_state->err;
        CU(c, cudaMemsetAsync(&dfs->err, 0, sizeof(uint32_t), c->stream));
        if ((flags & ERR_INDEX) || !c->last_async.valid) return map_err_flags(c, flags);
        // An async frame outgrew the device buffers.  Grow them from the counts the device kept and replay the LAST
        // async frame (its draw list and destination were kept), so the caller's latest image is complete; frames
        // issued before it since the previous rz_sync cannot be re-created and are reported.  The work counters include
        // the overflowed attempt.
        std::vector<DrawCmd> recorded;
        recorded.swap(c->draws); // draws recorded after the frame stay recorded
        const size_t staging_used = c->staging_used;
        const int parity = c->parity;
        int rc = RZ_OK;
        for (int attempt = 0; attempt < 8 && flags; attempt++) {
            rc = grow_after_overflow(c, flags);
            if (rc != RZ_OK) break;
            auto &a = c->last_async;
            c->draws = a.draws;
            c->staging_used = a.staging_used; c->parity = a.parity;
            rc = enqueue_frame(c, a.out_base, false);
            c->draws.clear();
            if (rc == RZ_OK && a.out_host) rc = enqueue_download(c, a.ring, a.out_host);
            if (rc != RZ_OK) break;
            CU(c, cudaMemcpyAsync(c->h_state, dfs, sizeof(FrameState), cudaMemcpyDeviceToHost, c->stream));
            CU(c, cudaStreamSynchronize(c->stream));
            CU(c, cudaStreamSynchronize(c->down_stream));
            flags = c->h_state->err;
            if (flags) CU(c, cudaMemsetAsync(&dfs->err, 0, sizeof(uint32_t), c->stream));
        }
        c->draws.swap(recorded);
        c->staging_used = staging_used; c->parity = parity;
        if (rc != RZ_OK) return rc;
        if (flags) return map_err_flags(c, flags);
        if (pending > 1)
            return fail(c, RZ_E_CAPACITY, "an async frame outgrew the device buffers: the buffers were grown and the last frame was "
                                          "replayed (its image is complete); the %u frame(s) issued before it may be incomplete", pending - 1);
    }
    return RZ_OK;
}

// ---- peer memory (one process per GPU; CUDA IPC over NVLink) ----
int rz_shared_alloc(rz_ctx *c, uint64_t bytes, void **dev_ptr, uint8_t *handle64) {
    if (!c || !dev_ptr || !handle64 || bytes == 0) return RZ_E_INVALID;
    static_assert(sizeof(cudaIpcMemHandle_t) == 64, "IPC handle is 64 bytes");
    CU(c, cudaSetDevice(c->device));
    void *p = nullptr;
    CU(c, cudaMalloc(&p, bytes));
    cudaError_t e = cudaMemset(p, 0, bytes);
    cudaIpcMemHandle_t h;
    if (e == cudaSuccess) e = cudaIpcGetMemHandle(&h, p);
    if (e != cudaSuccess) {
        cudaFree(p);
        return fail(c, RZ_E_CUDA, "rz_shared_alloc: %s", cudaGetErrorString(e));
    }
    memcpy(handle64, &h, 64);
    *dev_ptr = p;
    return RZ_OK;
}

int rz_shared_open(rz_ctx *c, const uint8_t *handle64, void **dev_ptr) {
    if (!c || !handle64 || !dev_ptr) return RZ_E_INVALID;
    CU(c, cudaSetDevice(c->device));
    cudaIpcMemHandle_t h;
    memcpy(&h, handle64, 64);
    cudaError_t e = cudaIpcOpenMemHandle(dev_ptr, h, cudaIpcMemLazyEnablePeerAccess);
    if (e != cudaSuccess) {
        (void)cudaGetLastError();
        return fail(c, RZ_E_CUDA, "rz_shared_open: %s (peer access between the two GPUs is required)", cudaGetErrorString(e));
    }
    return RZ_OK;
}

int rz_shared_close(rz_ctx *c, void *dev_ptr) {
    if (!c || !dev_ptr) return RZ_E_INVALID;
    CU(c, cudaSetDevice(c->device));
    CU(c, cudaStreamSynchronize(c->stream));
    CU(c, cudaIpcCloseMemHandle(dev_ptr));
    return RZ_OK;
}

int rz_shared_free(rz_ctx *c, void *dev_ptr) {
    if (!c || !dev_ptr) return RZ_E_INVALID;
    CU(c, cudaSetDevice(c->device));
    CU(c, cudaStreamSynchronize(c->stream));
    CU(c, cudaFree(dev_ptr));
    return RZ_OK;
}

int rz_signal(rz_ctx *c, uint32_t *const *flags, uint32_t n, uint32_t value) {
    if (!c || !flags || n == 0 || n > 16) return RZ_E_INVALID;
    if (c->sticky != RZ_OK) return c->sticky;
    CU(c, cudaSetDevice(c->device));
    FlagList fl;
    memset(&fl, 0, sizeof fl);
    for (uint32_t i = 0; i < n; i++) {
        if (!flags[i]) return RZ_E_INVALID;
        fl.p[i] = flags[i];
    }
    signal_kernel<<<1, 32, 0, c->stream>>>(fl, n, value);
    c->launches++;
    CU(c, cudaGetLastError());
    return RZ_OK;
}

int rz_wait_flags(rz_ctx *c, const uint32_t *flags, uint32_t n, uint32_t stride_bytes, uint32_t value, uint32_t timeout_ms) {
    if (!c || !flags || n == 0 || stride_bytes % 4 != 0) return RZ_E_INVALID;
    if (c->sticky != RZ_OK) return c->sticky;
    CU(c, cudaSetDevice(c->device));
    FrameState *dfs = reinterpret_cast<FrameState *>(c->d_state);
    wait_flags_kernel<<<(n + 31) / 32, 32, 0, c->stream>>>(flags, n, stride_bytes / 4, value,
                                                           (unsigned long long)(timeout_ms ? timeout_ms : 2000u) * 1000000ull, &dfs->peer_timeout);
    c->launches++;
    CU(c, cudaGetLastError());
    return RZ_OK;
}

int rz_set_row_range(rz_ctx *c, uint32_t row_begin, uint32_t row_end) {
    if (!c) return RZ_E_INVALID;
    if (row_begin >= row_end || row_end > c->H || row_begin % TH != 0 || (row_end % TH != 0 && row_end != c->H))
        return fail(c, RZ_E_INVALID, "rz_set_row_range: rows must be tile-aligned (%d) and inside the framebuffer", TH);
    c->row_begin = row_begin;
    c->row_end = row_end;
    return RZ_OK;
}

int rz_set_scissor(rz_ctx *c, uint32_t x0, uint32_t y0, uint32_t x1, uint32_t y1) {
    if (!c) return RZ_E_INVALID;
    c->sc_x0 = std::min(x0, c->W); c->sc_x1 = std::min(x1, c->W);
    c->sc_y0 = std::min(y0, c->H); c->sc_y1 = std::min(y1, c->H);
    return RZ_OK;
}

static int alloc_debug_buffers(rz_ctx *c) {
    if (c->d_dbg_depth && c->dbg_msaa == c->msaa) return RZ_OK;
    cudaFree(c->d_dbg_depth); cudaFree(c->d_dbg_color); cudaFree(c->d_dbg_owner);
    c->d_dbg_depth = nullptr; c->d_dbg_color = nullptr; c->d_dbg_owner = nullptr;
    const size_t n = (size_t)c->W * c->H * c->msaa;
    CU(c, cudaMalloc(&c->d_dbg_depth, n * 4));
    CU(c, cudaMalloc(&c->d_dbg_color, n * 4));
    CU(c, cudaMalloc(&c->d_dbg_owner, n * 4));
    c->dbg_msaa = c->msaa;
    return RZ_OK;
}

int rz_set_msaa(rz_ctx *c, uint32_t samples) {
    if (!c) return RZ_E_INVALID;
    if (samples != 1 && samples != 2 && samples != 4 && samples != 8)
        return fail(c, RZ_E_INVALID, "rz_set_msaa: %u samples per pixel (supported: 1, 2, 4, 8; the reference has 4, rasterizer/mod.rs:23)", samples);
    if (!c->draws.empty()) return fail(c, RZ_E_INVALID, "rz_set_msaa: draws are recorded for the current frame (one sample count per frame)");
    CU(c, cudaSetDevice(c->device));
    CU(c, cudaStreamSynchronize(c->stream));
    c->msaa = samples;
    if (c->d_dbg_depth) return alloc_debug_buffers(c);
    return RZ_OK;
}

int rz_set_guard_band(rz_ctx *c, float factor) {
    if (!c) return RZ_E_INVALID;
    if (!(factor >= 1.0f) || factor > 1048576.0f) return fail(c, RZ_E_INVALID, "rz_set_guard_band: the factor must be in [1, 2^20]");
    c->guard = factor;
    return RZ_OK;
}

int rz_set_row_interleave(rz_ctx *c, uint32_t band_tile_rows, uint32_t rank, uint32_t world) {
    if (!c) return RZ_E_INVALID;
    if (band_tile_rows == 0 || world <= 1) {
        c->il_band = 0; c->il_rank = 0; c->il_world = 1;
        return RZ_OK;
    }
    if (rank >= world) return fail(c, RZ_E_INVALID, "rz_set_row_interleave: rank %u >= world %u", rank, world);
    c->il_band = band_tile_rows; c->il_rank = rank; c->il_world = world;
    return RZ_OK;
}

int rz_counters(rz_ctx *c, rz_counters_t *out) {
    if (!c || !out) return RZ_E_INVALID;
    CU(c, cudaSetDevice(c->device));
    CU(c, cudaMemcpyAsync(c->h_state, c->d_state, sizeof(FrameState), cudaMemcpyDeviceToHost, c->stream));
    CU(c, cudaStreamSynchronize(c->stream));
    unsigned long long k[16] = {0};
    for (int st = 0; st < CNT_STRIPES; st++)
        for (int i = 0; i < 16; i++) k[i] += c->h_state->counters[st][i];
    out->n_tris_in = k[C_TRIS_IN]; out->n_degenerate = k[C_DEGENERATE]; out->n_outside = k[C_OUTSIDE];
    out->n_inside = k[C_INSIDE]; out->n_clipped_in = k[C_CLIPPED_IN]; out->n_tris_setup = k[C_TRIS_SETUP];
    out->n_bbox_px = k[C_BBOX_PX]; out->n_covered_px = k[C_COVERED_PX]; out->n_shaded_px = k[C_SHADED_PX];
    out->n_samples_written = k[C_SAMPLES]; out->n_tex_oob = k[C_TEX_OOB]; out->n_clip_overflow = k[C_CLIP_OVF];
    return RZ_OK;
}

int rz_reset_counters(rz_ctx *c) {
    if (!c) return RZ_E_INVALID;
    CU(c, cudaSetDevice(c->device));
    CU(c, cudaMemsetAsync(c->d_state, 0, sizeof(unsigned long long) * 16 * CNT_STRIPES, c->stream));
    return RZ_OK;
}

int rz_timings(rz_ctx *c, rz_timings_t *out) {
    if (!c || !out) return RZ_E_INVALID;
    *out = c->timings;
    return RZ_OK;
}

uint64_t rz_launch_count(rz_ctx *c) { return c ? c->launches : 0; }

int rz_debug_capture(rz_ctx *c, int enable) {
    if (!c) return RZ_E_INVALID;
    CU(c, cudaSetDevice(c->device));
    if (enable) {
        int rc = alloc_debug_buffers(c);
        if (rc != RZ_OK) return rc;
        if (!c->d_dbg_time) {
            CU(c, cudaMalloc(&c->d_dbg_time, (size_t)c->tiles_x * c->tiles_y * 64));
            CU(c, cudaMemset(c->d_dbg_time, 0, (size_t)c->tiles_x * c->tiles_y * 64));
        }
    }
    c->debug = enable != 0;
    return RZ_OK;
}

int rz_debug_read(rz_ctx *c, float *depth, uint32_t *color, uint32_t *owner) {
    if (!c) return RZ_E_INVALID;
    if (!c->d_dbg_depth) return fail(c, RZ_E_INVALID, "rz_debug_read: capture was never enabled");
    CU(c, cudaSetDevice(c->device));
    const size_t bytes = (size_t)c->W * c->H * 4 * c->dbg_msaa;
    CU(c, cudaStreamSynchronize(c->stream));
    if (depth) CU(c, cudaMemcpy(depth, c->d_dbg_depth, bytes, cudaMemcpyDeviceToHost));
    if (color) CU(c, cudaMemcpy(color, c->d_dbg_color, bytes, cudaMemcpyDeviceToHost));
    if (owner) CU(c, cudaMemcpy(owner, c->d_dbg_owner, bytes, cudaMemcpyDeviceToHost));
    return RZ_OK;
}

int rz_debug_tile_times(rz_ctx *c, uint64_t *out, uint32_t max_tiles, uint32_t *n_written) {
    if (!c || !out || !n_written) return RZ_E_INVALID;
    if (!c->d_dbg_time) return fail(c, RZ_E_INVALID, "rz_debug_tile_times: capture was never enabled");
    CU(c, cudaSetDevice(c->device));
    CU(c, cudaStreamSynchronize(c->stream));
    const uint32_t n = std::min<uint32_t>(max_tiles, c->tiles_x * c->tiles_y);
    CU(c, cudaMemcpy(out, c->d_dbg_time, (size_t)n * 64, cudaMemcpyDeviceToHost));
    *n_written = n;
    return RZ_OK;
}

int rz_debug_vertex_stage(rz_ctx *c, const rz_mesh *mesh, float *out_clip) {
    if (!c || !mesh || !out_clip) return RZ_E_INVALID;
    CU(c, cudaSetDevice(c->device));
    if (mesh->nv == 0) return RZ_OK;
    CU(c, cudaStreamSynchronize(c->stream));
    int rc = ensure_vertex_scratch(c, mesh->nv);
    if (rc != RZ_OK) return rc;
    FrameParams P = make_params(c, c->d_out);
    DrawParams D;
    memset(&D, 0, sizeof D);
    D.draw = 0xFFFFFFFFu; // no per-draw table entry
    float pv[16];
    mat4_mul(c->proj, c->view, pv);
    mat4_mul(pv, c->world, D.M);
    D.pos = mesh->d_pos; D.nv = mesh->nv;
    D.vtx = c->d_vtx;
    CU(c, launch_pdl(vertex_kernel, dim3((mesh->nv + NT * VERTEX_PER_THREAD - 1) / (NT * VERTEX_PER_THREAD)), dim3(NT), 0, c->stream, P, D,
                     (uint4 *)nullptr, 0u));
    c->launches++;
    std::vector<float> tmp((size_t)mesh->nv * 8);
    CU(c, cudaMemcpyAsync(tmp.data(), c->d_vtx, tmp.size() * 4, cudaMemcpyDeviceToHost, c->stream));
    CU(c, cudaStreamSynchronize(c->stream));
    for (size_t v = 0; v < mesh->nv; v++) { // (clip x, y, z) from word 4..6, clip w from word 3
        out_clip[4 * v + 0] = tmp[8 * v + 4]; out_clip[4 * v + 1] = tmp[8 * v + 5];
        out_clip[4 * v + 2] = tmp[8 * v + 6]; out_clip[4 * v + 3] = tmp[8 * v + 3];
    }
    return RZ_OK;
}

} // extern "C"
