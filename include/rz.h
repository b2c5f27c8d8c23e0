/*
 * rz.h -- C ABI of the B200-native raster path (librz_b200.so).
 *
 * Drop-in boundary for rusterizer's per-frame raster pipeline.  Every entry point names the
 * reference interface it replaces (paths relative to the reference's src/).  Plain pointers and
 * sizes only: a Rust `extern "C"` block, a cgo stub or Python ctypes can bind it unchanged
 * (see INTEGRATION.md).  There is no CPU fallback behind this ABI: rz_create fails when no
 * CUDA device is present.
 *
 * Conventions
 *   - return value: RZ_OK (0) or a negative RZ_E_* code; text via rz_last_error(ctx)
 *   - matrices: 16 floats, row-major [[f32;4];4]            (math/matrix.rs:13, uniform.rs:5-9)
 *   - positions: f32[nv][3]  (Point3D<WorldSpace>)          (math/point.rs:22, mesh.rs:9)
 *   - attributes: f32[nv][6] = r,g,b,a,u,v (VertexAttribute) (graphics_primitives.rs:10-13)
 *   - indices: u32 (the reference stores usize, mesh.rs:10; nv < 2^32 so the narrowing is lossless)
 *   - textures: u8[h][w][texel_width], texel_width 3|4, origin top-left (texture.rs:8-13,51)
 *   - framebuffer: u32[height][width] 0xAARRGGBB, alpha 0xFF, origin top-left
 *                                                           (rasterizer/buffers.rs:121-124)
 *   - shaders are selected by identifier because Rust fn pointers (render.rs:33-36) cannot run
 *     on the GPU: VS ids follow main.rs:147-152, FS ids follow `enum FS` main.rs:23-27.
 *   - a ctx is NOT thread-safe (the reference takes &mut self everywhere); one ctx per GPU/stream.
 */
#ifndef RZ_H
#define RZ_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define RZ_OK 0
#define RZ_E_INVALID -1   /* bad argument (null pointer, unknown shader id, zero size ...)        */
#define RZ_E_CUDA -2      /* CUDA runtime error (sticky on the ctx)                               */
#define RZ_E_NO_DEVICE -3 /* no CUDA device / device index out of range                           */
#define RZ_E_TEXTURE -4   /* bind index != number of bound textures (assert! at uniform.rs:31),
                             or FS Texture used with no texture bound (index panic uniform.rs:36) */
#define RZ_E_INDEX -5     /* a mesh index >= nv (slice-index panic at render.rs:83-87)            */
#define RZ_E_CAPACITY -6  /* async frames outgrew the device buffers; rz_sync has grown them and replayed
                             the last one, earlier ones must be re-issued (see rz_sync)            */
#define RZ_E_NOMEM -7     /* device or host allocation failed                                     */
#define RZ_E_PEER -8      /* a peer GPU did not raise its completion flag within the timeout      */

#define RZ_VS_MVP 0     /* projection * view * world * (x,y,z,1)          main.rs:147-152 */
#define RZ_FS_TEXTURE 0 /* get_texture(0).sample(u, v)                    main.rs:69-71   */
#define RZ_FS_COLOR 1   /* attr.color                                     main.rs:72      */
#define RZ_FS_DEBUG 2   /* Color::grayscale(frag_coords.depths[0])        main.rs:73-75   */
/* Shader registry beyond the crate's three closures.  The reference lets a fragment
 * shader closure read any bound texture (Uniforms::get_texture(index), uniform.rs:35-37) and combine Colors
 * with the operators of color.rs:75-123; as identifiers that becomes:
 *   - a texture index in bits 8..15 of fs_id for the shaders that sample:  RZ_FS_WITH_TEXTURE(fs, index)
 *   - RZ_FS_TEXTURE_BLEND: (get_texture(index).sample(u, v) + attr.color) / 2.0   (Color Add, Div<f32>) */
#define RZ_FS_TEXTURE_BLEND 3
#define RZ_FS_WITH_TEXTURE(fs, index) ((uint32_t)(fs) | ((uint32_t)(index) << 8))
#define RZ_MAX_TEXTURES 32

typedef struct rz_ctx rz_ctx;   /* Renderer + Rasterizer + Uniforms state (render.rs:38-45)  */
typedef struct rz_mesh rz_mesh; /* a device-resident Mesh<WorldSpace>     (mesh.rs:5-12)     */

/* Work counters, identical in meaning to the oracle's. */
typedef struct rz_counters_t {
    uint64_t n_tris_in;         /* triangles entering Rasterizer::rasterize       mod.rs:425      */
    uint64_t n_degenerate;      /* culled by the clip-space area test             clipping.rs:63  */
    uint64_t n_outside;         /* ClipResult::Outside (trivial or after S-H)     clipping.rs:106,175 */
    uint64_t n_inside;          /* ClipResult::Inside                             clipping.rs:110 */
    uint64_t n_clipped_in;      /* ClipResult::Clipped                            clipping.rs:194 */
    uint64_t n_tris_setup;      /* triangles reaching viewport_transform          mod.rs:438-440  */
    uint64_t n_bbox_px;         /* pixels the reference's bbox walk visits        mod.rs:443-444  */
    uint64_t n_covered_px;      /* pixels with any_coverage()                     mod.rs:446      */
    uint64_t n_shaded_px;       /* fragment-shader invocations (post-depth != 0)  mod.rs:465      */
    uint64_t n_samples_written; /* samples written by write_pixel                 mod.rs:390-395  */
    uint64_t n_tex_oob;         /* texture byte reads past the buffer (a panic in the reference;
                                   clamped to the last byte here and in the oracle)               */
    uint64_t n_clip_overflow;   /* clipped polygons that outgrew the fixed vertex budget (never
                                   expected; the reference's Vec is unbounded)                    */
} rz_counters_t;

/* Per-stage device times of the last completed frame, in milliseconds (CUDA events). */
typedef struct rz_timings_t {
    float geometry_ms; /* vertex transform + clip + setup + small-triangle binning, all draws */
    float bin_ms;      /* large-triangle binning                                              */
    float tile_ms;     /* per-tile raster + depth + shade + resolve                           */
    float total_ms;    /* first launch to last launch of the frame                            */
} rz_timings_t;

/* Renderer::new(width, height) (render.rs:48-69) + Rasterizer::new (rasterizer/mod.rs:273-281),
 * minus the minifb window.  `device` is a CUDA ordinal. */
int rz_create(int device, uint32_t width, uint32_t height, rz_ctx **out);
void rz_destroy(rz_ctx *ctx);

/* Run this ctx's work on a caller-owned cudaStream_t (e.g. torch's current stream). NULL = the
 * ctx's own stream.  No reference counterpart (the reference is synchronous). */
int rz_set_stream(rz_ctx *ctx, void *cuda_stream);

/* Uniforms::bind_texture(index, tex) (uniform.rs:29-33): index must equal the number of textures
 * bound so far, otherwise RZ_E_TEXTURE.  The texels are copied. */
int rz_bind_texture(rz_ctx *ctx, uint32_t index, const uint8_t *texels, uint32_t width, uint32_t height,
                    uint32_t texel_width);

/* Uniforms::write_block() (uniform.rs:43-45): overwrite world / view / projection.  A NULL matrix
 * keeps its current value (the reference mutates single fields: main.rs:135-142,171). */
int rz_write_block(rz_ctx *ctx, const float *world, const float *view, const float *projection);
/* Uniforms::read_block() (uniform.rs:39-41) */
int rz_read_block(rz_ctx *ctx, float *world, float *view, float *projection);

/* Upload a Mesh (mesh.rs:5-12) once so static geometry is not re-sent every frame.
 * n_idx = 3 * number of triangles.  The arrays are copied. */
int rz_mesh_create(rz_ctx *ctx, const float *positions, const float *attributes, uint32_t nv,
                   const uint32_t *indices, uint64_t n_idx, rz_mesh **out);
void rz_mesh_destroy(rz_mesh *mesh);

/* Renderer::render(&mesh, vertex_shader, fragment_shader) (render.rs:98-114) ->
 * Rasterizer::rasterize (rasterizer/mod.rs:399-476).  Accumulates into the frame like the
 * reference (several meshes per frame: main.rs:170-173).  Asynchronous: the draw is recorded with
 * a snapshot of the uniform block and executed, in submission order with every other draw of the
 * frame, by the next rz_framebuffer*.  The mesh must stay alive until then. */
int rz_render(rz_ctx *ctx, const rz_mesh *mesh, uint32_t vs_id, uint32_t fs_id);

/* Same call with a HOST mesh, exactly the reference signature (the mesh is borrowed for the call:
 * it is copied to the device before returning). */
int rz_render_host(rz_ctx *ctx, const float *positions, const float *attributes, uint32_t nv,
                   const uint32_t *indices, uint64_t n_idx, uint32_t vs_id, uint32_t fs_id);

/* Rasterizer::framebuffer() -> resolve_and_clear (rasterizer/mod.rs:478-522), called by
 * Renderer::display (render.rs:121): executes the frame, box-filters the 4 samples, clears the
 * sample state for the next frame.  Synchronises.  out_host (may be NULL) receives
 * width*height u32; *out_device (may be NULL) receives a device pointer owned by the ctx, valid
 * until the next rz_framebuffer* / rz_destroy.  Grows internal buffers and re-runs the frame if a
 * capacity was exceeded. */
int rz_framebuffer(rz_ctx *ctx, uint32_t *out_host, const uint32_t **out_device);

/* Same, without host synchronisation (for back-to-back frames).  device_dst (may be NULL = the
 * ctx's own buffer) is a caller-owned device buffer of width*height u32 that receives the image,
 * e.g. a slice of an NCCL gather buffer.  Errors of the frame surface at the next rz_sync. */
int rz_framebuffer_async(rz_ctx *ctx, uint32_t *device_dst, const uint32_t **out_device);

/* Streaming form of Renderer::display (render.rs:116-127) for a caller that presents every frame
 * on the host: executes the frame like rz_framebuffer_async and copies the resolved image (this
 * ctx's rows) to out_host (width*height u32, ideally pinned) on a separate copy stream, without host
 * synchronisation.  Together with rz_render_host -- whose uploads run on their own stream into one
 * of two alternating staging sets -- the H2D copy of frame i+1, the kernels of frame i and the
 * D2H copy of frame i-1 overlap.  out_host is valid after the next rz_sync(); at most two frames
 * are in flight per output buffer, so alternate between two host buffers.  With pinned host memory
 * the arrays passed to rz_render_host must stay unchanged until that rz_sync(). */
int rz_framebuffer_host_async(rz_ctx *ctx, uint32_t *out_host);

/* Wait for all enqueued frames and report their sticky status (RZ_E_INDEX, RZ_E_PEER ...).  An async frame cannot know
 * that it outgrew the device buffers (the cursors live on the device); rz_sync finds out, grows the buffers from the
 * counts the device kept and REPLAYS the last async frame into its original destination, so after RZ_OK the latest
 * image is complete.  Only when more than one async frame was issued since the previous rz_sync does it return
 * RZ_E_CAPACITY: the last frame is complete, the ones before it may not be (re-issue them; the buffers are large
 * enough now).  The work counters then include the overflowed attempt. */
int rz_sync(rz_ctx *ctx);
/* Drop the draws recorded since the last rz_framebuffer* without executing them (a rank of a screen-space split that owns
 * no rows of this frame still receives the application's rz_render calls).  No reference counterpart. */
int rz_discard_frame(rz_ctx *ctx);

/* Screen-space sharding (multi-GPU tile ranges): this ctx rasterises and resolves only the pixel
 * rows [row_begin, row_end) (rounded outwards to tile rows by the caller via rz_tile_height()).
 * Rows outside are left untouched in the output.  Default: the whole framebuffer. */
int rz_set_row_range(rz_ctx *ctx, uint32_t row_begin, uint32_t row_end);
/* Screen-space sharding without a separate gather step (one process per GPU, NVLink peer memory).
 * The rank that presents the frame allocates the image with rz_shared_alloc and sends the 64-byte
 * handle to its peers (any transport; the Python mirror uses torch.distributed); they map it with
 * rz_shared_open and pass `mapped + row_begin*width` to rz_framebuffer_async, so their tile kernels
 * store the resolved rows straight into the presenting GPU's memory while they rasterise.  Completion
 * travels the same way: after its frame a rank calls rz_signal({&flags[rank]}, 1, seq) on a mapped flag
 * array of the presenter, which calls rz_wait_flags(flags, n, stride, seq) -- both are stream-ordered
 * kernels, nothing blocks the host.  seq must increase from frame to frame (wrap-around safe).
 * rz_wait_flags gives up after timeout_ms (0 = 2000 ms); the next rz_sync then returns RZ_E_PEER.
 * No reference counterpart (the reference is a single-threaded CPU program). */
int rz_shared_alloc(rz_ctx *ctx, uint64_t bytes, void **dev_ptr, uint8_t *handle64);
int rz_shared_open(rz_ctx *ctx, const uint8_t *handle64, void **dev_ptr);
int rz_shared_close(rz_ctx *ctx, void *dev_ptr);
int rz_shared_free(rz_ctx *ctx, void *dev_ptr);
int rz_signal(rz_ctx *ctx, uint32_t *const *flags, uint32_t n, uint32_t value); /* n <= 16 flags, one kernel */
int rz_wait_flags(rz_ctx *ctx, const uint32_t *flags, uint32_t n, uint32_t stride_bytes, uint32_t value,
                  uint32_t timeout_ms);
/* Scissor rectangle, the extension the reference sketches in Rasterizer::bounding_box
 * (rasterizer/mod.rs:349-350: "the user would supply a scissoring rect that could be used to bound the
 * triangles"): every triangle's pixel bounding box is intersected with [x0,x1) x [y0,y1) instead of the
 * viewport, so nothing outside it is rasterised; those pixels resolve to the clear colour.  The rect is
 * clamped to the viewport; it applies to the draws of the frames executed after the call (one rect per
 * frame, like the resolution).  Default: the whole viewport. */
int rz_set_scissor(rz_ctx *ctx, uint32_t x0, uint32_t y0, uint32_t x1, uint32_t y1);

/* Runtime sample count, SURVEY.md section 8 f-4: the reference fixes N_MSAA_SAMPLES = 4 with the rotated-grid pattern
 * (rasterizer/mod.rs:23,109-114); here 1, 2, 4 (default, the reference) or 8 samples per pixel can be selected between
 * frames.  1 = the pixel centre, 2 and 8 = the D3D11 standard patterns.  Everything that depends on the count in the
 * reference follows it: CoverageMask::all() (mod.rs:42-44), the shading position rule of Fragment::interpolate
 * (mod.rs:70-83), box_filter_color (buffers.rs:111-125).  rz_debug_read then returns [height][width][samples]. */
int rz_set_msaa(rz_ctx *ctx, uint32_t samples);
/* Guard-band clipping, the extension the reference sketches at rasterizer/mod.rs:417-419: the four side clip planes move
 * out to |x|, |y| <= factor * w (near / far stay), so triangles that leave the viewport but stay inside the band are
 * rasterised unclipped -- their pixel boxes are bounded by the viewport / scissor as before (mod.rs:347-361).  Trivial
 * rejection still uses the view frustum itself.  factor = 1 (default) is the reference's clipping bit for bit. */
int rz_set_guard_band(rz_ctx *ctx, float factor);

/* Interleaved screen-space sharding: this ctx owns the bands k of `band_tile_rows` tile rows with
 * k % world == rank (inside its row range, normally the whole frame), which balances a centred object
 * across the GPUs; only owned tiles are rasterised, resolved and written, so with rz_framebuffer_async
 * every rank passes the FULL image (e.g. the peer-mapped one) as device_dst.  band_tile_rows = 0 or
 * world <= 1 switches interleaving off. */
int rz_set_row_interleave(rz_ctx *ctx, uint32_t band_tile_rows, uint32_t rank, uint32_t world);
uint32_t rz_tile_width(void);
uint32_t rz_tile_height(void);

/* Counters accumulated since the last rz_reset_counters (synchronises). */
int rz_counters(rz_ctx *ctx, rz_counters_t *out);
int rz_reset_counters(rz_ctx *ctx);
/* Stage timings of the last frame issued through rz_framebuffer() (synchronous path only). */
int rz_timings(rz_ctx *ctx, rz_timings_t *out);
/* Number of kernels this ctx has launched since creation (for gpu_launches accounting). */
uint64_t rz_launch_count(rz_ctx *ctx);

/* Parity instrumentation: when enabled, the next frames also keep the per-sample state the
 * reference holds in ColorBuffer.buffer / DepthBuffer.buffer (rasterizer/buffers.rs:83-157) as it
 * is just before resolve_and_clear, plus the order key (8*triangle_number + fan_index) of the
 * triangle that last wrote each sample.  Each array is [height][width][4]; any may be NULL. */
int rz_debug_capture(rz_ctx *ctx, int enable);
int rz_debug_read(rz_ctx *ctx, float *depth, uint32_t *color, uint32_t *owner);
/* Profiling aid (capture must be enabled): for every non-empty tile of the last frame, 8 u64 words:
 * [0] tile id | list length << 32, [1],[2] start and end of its processing in GPU nanoseconds,
 * [3] SM id, [4] four 16-bit phase stamps of the first chunk (A0,A1,A2,B done; units of 16 ns
 * after start), [5] ns after start when phase C of the first chunk was done. */
int rz_debug_tile_times(rz_ctx *ctx, uint64_t *out, uint32_t max_tiles, uint32_t *n_written);
/* Vertex stage only (render.rs:104-108): clip-space positions f32[nv][4] of a mesh under the
 * current uniform block. */
int rz_debug_vertex_stage(rz_ctx *ctx, const rz_mesh *mesh, float *out_clip);

const char *rz_last_error(rz_ctx *ctx);
const char *rz_version(void);

#ifdef __cplusplus
}
#endif
#endif /* RZ_H */
