// rusterizer.hpp -- header-only C++ host layer mirroring the reference crate's render/shader API
// surface over the C ABI (include/rz.h).  The reference host is Rust (render.rs driver, mesh.rs,
// uniform.rs, camera.rs); this image has no Rust toolchain, so the same surface is provided in
// C++: same names, argument meaning and failure points (the reference's panics become exceptions).
//
//   rz::Renderer r(1280, 720);                              // Renderer::new           render.rs:48
//   r.uniforms().write_block().view = camera.get_view_matrix();   // main.rs:135-136
//   r.uniforms().write_block().projection = rz::project(1.f, 200.f, 720.f / 1280.f, kPi / 2);
//   r.uniforms().bind_texture(0, tex);                       // uniform.rs:29
//   r.uniforms().write_block().world = m;  r.render(mesh, rz::VS::MVP, rz::FS::Texture);  // main.rs:170-172
//   const std::vector<uint32_t>& fb = r.framebuffer();       // Rasterizer::framebuffer  mod.rs:520
#pragma once
#include <array>
#include <cmath>
#include <cstdint>
#include <stdexcept>
#include <string>
#include <vector>

#include "../../include/rz.h"
#include "rz_image.hpp"

namespace rz {

using Mat4 = std::array<float, 16>; // row-major [[f32;4];4]  (math/matrix.rs:13)

enum class VS : uint32_t { MVP = RZ_VS_MVP };                                            // main.rs:147-152
enum class FS : uint32_t { Texture = RZ_FS_TEXTURE, Color = RZ_FS_COLOR, Debug = RZ_FS_DEBUG, // main.rs:23-27
                           TextureBlend = RZ_FS_TEXTURE_BLEND };                               // registry extension (rz.h)

struct Error : std::runtime_error {
    int code;
    Error(int c, const std::string &m) : std::runtime_error(m), code(c) {}
};

// ---- math (host-only: builds the matrices that cross the boundary) -------------------------------
// One IEEE binary32 rounding per operation in the reference's order; compile the including TU with
// -ffp-contract=off so the host never fuses a multiply-add.
inline float dot4(const float *a, const float *b, int sb) { // math/vector.rs:17-23
    volatile float s = 0.0f;
    for (int k = 0; k < 4; k++) {
        volatile float p = a[k] * b[k * sb];
        s = s + p;
    }
    return s;
}
inline Mat4 mul(const Mat4 &a, const Mat4 &b) { // math/matrix.rs:56-79
    Mat4 r{};
    for (int i = 0; i < 4; i++)
        for (int j = 0; j < 4; j++) r[i * 4 + j] = dot4(&a[i * 4], &b[j], 4);
    return r;
}
inline Mat4 identity() { return {1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1}; }
inline Mat4 translate(float x, float y, float z) { return {1, 0, 0, x, 0, 1, 0, y, 0, 0, 1, z, 0, 0, 0, 1}; } // transform.rs:10-17
inline Mat4 rotate_x(float r) { float c = std::cos(r), s = std::sin(r); return {1, 0, 0, 0, 0, c, -s, 0, 0, s, c, 0, 0, 0, 0, 1}; }
inline Mat4 rotate_y(float r) { float c = std::cos(r), s = std::sin(r); return {c, 0, s, 0, 0, 1, 0, 0, -s, 0, c, 0, 0, 0, 0, 1}; }
inline Mat4 rotate_z(float r) { float c = std::cos(r), s = std::sin(r); return {c, -s, 0, 0, s, c, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1}; }
inline Mat4 rotate(float x, float y, float z) { return mul(mul(rotate_z(z), rotate_y(y)), rotate_x(x)); } // transform.rs:91-96
inline Mat4 project(float near, float far, float aspect_ratio, float vert_fov) { // math/mod.rs:92-122
    if (!(near > 0.0f)) throw Error(RZ_E_INVALID, "project: near must be > 0 (assert! at math/mod.rs:98)");
    const float half_width = std::tan(vert_fov / 2.0f) * near;
    const float half_height = aspect_ratio * half_width;
    return {near / half_width, 0, 0, 0, 0, near / half_height, 0, 0,
            0, 0, -(far + near) / (far - near), -2.0f * far * near / (far - near), 0, 0, -1.0f, 0};
}

// ---- resources -------------------------------------------------------------------------------------
struct VertexAttribute { // graphics_primitives.rs:10-13 + color.rs:7-12
    float r, g, b, a, u, v;
};
static_assert(sizeof(VertexAttribute) == 24, "VertexAttribute must be 6 packed floats");

struct Mesh { // mesh.rs:5-12 (indices narrowed from usize to u32)
    std::vector<std::array<float, 3>> vertices;
    std::vector<uint32_t> indices;
    std::vector<VertexAttribute> attributes;
};

inline Mesh centered_quad(float width) { // mesh.rs:15-41
    const float h = width / 2.0f;
    return {{{-h, h, 2}, {h, h, 2}, {h, -h, 2}, {-h, -h, 2}},
            {0, 1, 2, 0, 2, 3},
            {{1, 0, 0, 1, 0, 0}, {0, 0, 1, 1, 1, 0}, {0, 1, 0, 1, 1, 1}, {1, 1, 1, 1, 0, 1}}};
}
inline Mesh triangle() { // mesh.rs:44-66
    return {{{-1, -1, 2}, {0, 1, 2}, {1, -1, 2}}, {0, 1, 2}, {{1, 0, 0, 1, 0, 1}, {0, 0, 1, 1, 1, 0}, {0, 1, 0, 1, 1, 1}}};
}
inline Mesh cube(float width) { // mesh.rs:69-149
    static const float B[24][3] = {
        {-.5f, .5f, -.5f}, {.5f, .5f, -.5f}, {.5f, -.5f, -.5f}, {-.5f, -.5f, -.5f}, // front
        {.5f, .5f, .5f},   {-.5f, .5f, .5f}, {-.5f, -.5f, .5f}, {.5f, -.5f, .5f},   // back
        {-.5f, .5f, .5f},  {-.5f, .5f, -.5f}, {-.5f, -.5f, -.5f}, {-.5f, -.5f, .5f}, // left
        {.5f, .5f, -.5f},  {.5f, .5f, .5f},  {.5f, -.5f, .5f},  {.5f, -.5f, -.5f},   // right
        {-.5f, .5f, -.5f}, {-.5f, .5f, .5f}, {.5f, .5f, .5f},   {.5f, .5f, -.5f},    // top
        {-.5f, -.5f, .5f}, {-.5f, -.5f, -.5f}, {.5f, -.5f, -.5f}, {.5f, -.5f, .5f}}; // bottom
    static const float UV[4][2] = {{0, 0}, {1, 0}, {1, 1}, {0, 1}};
    static const float COL[3][3] = {{1, 0, 0}, {0, 0, 1}, {0, 1, 0}}; // red, blue, green by i % 3
    Mesh m;
    for (uint32_t i = 0; i < 24; i++) {
        m.vertices.push_back({B[i][0] * width, B[i][1] * width, B[i][2] * width});
        m.attributes.push_back({COL[i % 3][0], COL[i % 3][1], COL[i % 3][2], 1.0f, UV[i % 4][0], UV[i % 4][1]});
    }
    for (uint32_t f = 0; f < 6; f++)
        for (uint32_t v : {0u, 1u, 2u, 0u, 2u, 3u}) m.indices.push_back(f * 4 + v);
    return m;
}
inline Mesh sphere(float radius, uint32_t n_phi = 17, uint32_t n_theta = 9) { // mesh.rs:152-207
    Mesh m;
    const float pi = 3.14159274101257324f; // std::f32::consts::PI
    for (uint32_t i = 0; i < n_theta; i++)
        for (uint32_t j = 0; j < n_phi; j++) {
            const float theta_ratio = (float)i / (float)(n_theta - 1), phi_ratio = (float)j / (float)(n_phi - 1);
            const float phi = pi * 2.0f * phi_ratio, theta = pi * theta_ratio;
            const float x = radius * std::sin(theta) * std::cos(phi), y = radius * std::cos(theta),
                        z = radius * std::sin(theta) * std::sin(phi);
            m.vertices.push_back({x, y, z});
            if (i < n_theta - 1 && j < n_phi - 1) {
                const uint32_t a = n_phi * i + j, b = n_phi * i + j + 1, c = n_phi * (i + 1) + j + 1, d = n_phi * (i + 1) + j;
                for (uint32_t v : {a, b, c, a, c, d}) m.indices.push_back(v);
            }
            m.attributes.push_back({std::fabs(x), std::fabs(y), std::fabs(z), 1.0f, phi_ratio, theta_ratio});
        }
    return m;
}

struct Texture { // texture.rs:8-13: row-major u8[h][w][texel_width], origin top-left
    std::vector<uint8_t> buf;
    uint32_t width = 0, height = 0, texel_width = 4;

    // Texture::from_png_file (texture.rs:26-45).  The reference hard-codes texel_width = 4 whatever the file
    // holds (texture.rs:41), which is only right for RGBA files like images/checkerboard.png; here the width
    // follows the decoded layout (3 for RGB, 4 for RGBA), the two layouts read_texel handles (texture.rs:47-63).
    static Texture from_png_file(const std::string &path) {
        image::Image img = image::read_png(path);
        Texture t;
        t.buf = std::move(img.pixels);
        t.width = img.width; t.height = img.height; t.texel_width = img.channels;
        return t;
    }
    // The decoded content of images/checkerboard.png: 400x400 RGBA, 4x4 squares of 100 px, top-left black.
    static Texture checkerboard() {
        Texture t;
        t.width = t.height = 400;
        t.texel_width = 4;
        t.buf.resize(400 * 400 * 4);
        for (uint32_t y = 0; y < 400; y++)
            for (uint32_t x = 0; x < 400; x++) {
                const uint8_t v = ((x / 100 + y / 100) & 1) ? 255 : 0;
                uint8_t *p = &t.buf[(y * 400 + x) * 4];
                p[0] = p[1] = p[2] = v;
                p[3] = 255;
            }
        return t;
    }
};

struct Camera { // camera.rs:3-54
    std::array<float, 3> pos{0, 0, -5}, up{0, 1, 0}, dir{0, 0, 1};
    // Orbit constructor for camera sweeps (not in the reference, which only has Default): on a circle of
    // `radius` in the xz-plane at `angle`, looking at the origin, up +y.  angle 0 is Camera::default().
    static Camera orbit(float angle, float radius = 5.0f, float height = 0.0f) {
        Camera c;
        c.pos = {radius * std::sin(angle), height, -radius * std::cos(angle)};
        c.dir = {-c.pos[0], -c.pos[1], -c.pos[2]};
        return c;
    }
    Mat4 get_view_matrix() const { // camera.rs:10-43
        auto norm = [](std::array<float, 3> v) {
            float acc = 0.0f;
            for (float e : v) acc = acc + e * e;
            const float l = std::sqrt(acc);
            return std::array<float, 3>{v[0] / l, v[1] / l, v[2] / l};
        };
        auto cross = [](const std::array<float, 3> &a, const std::array<float, 3> &b) {
            return std::array<float, 3>{a[1] * b[2] - a[2] * b[1], a[2] * b[0] - a[0] * b[2], a[0] * b[1] - a[1] * b[0]};
        };
        const std::array<float, 3> d = norm(dir), u = norm(up);
        const std::array<float, 3> cz{d[0] * -1.0f, d[1] * -1.0f, d[2] * -1.0f};
        const std::array<float, 3> cx = norm(cross(cz, u)), cy = norm(cross(cx, cz));
        const Mat4 rotation_inv{cx[0], cx[1], cx[2], 0, cy[0], cy[1], cy[2], 0, cz[0], cz[1], cz[2], 0, 0, 0, 0, 1};
        const Mat4 translation_inv = translate(pos[0] * -1.0f, pos[1] * -1.0f, pos[2] * -1.0f);
        return mul(rotation_inv, translation_inv);
    }
};

// ---- Uniforms / Renderer ---------------------------------------------------------------------------
class Renderer;

struct UniformBlock { // uniform.rs:4-9
    Mat4 world = identity(), view = identity(), projection = identity();
};

class Uniforms { // uniform.rs:11-46
  public:
    UniformBlock &write_block() { return block_; }
    const UniformBlock &read_block() const { return block_; }
    void bind_texture(size_t index, const Texture &tex);
    const Texture &get_texture(size_t index) const { return textures_.at(index); }

  private:
    friend class Renderer;
    explicit Uniforms(rz_ctx *ctx) : ctx_(ctx) {}
    rz_ctx *ctx_;
    UniformBlock block_;
    std::vector<Texture> textures_;
};

class Renderer { // render.rs:38-127 without the minifb window
  public:
    Renderer(size_t width, size_t height, int device = 0) : width_(width), height_(height), uniforms_(nullptr) {
        const int rc = rz_create(device, (uint32_t)width, (uint32_t)height, &ctx_);
        if (rc != RZ_OK) throw Error(rc, rz_last_error(nullptr));
        uniforms_ = Uniforms(ctx_);
        fb_.resize(width * height);
    }
    ~Renderer() { rz_destroy(ctx_); }
    Renderer(const Renderer &) = delete;
    Renderer &operator=(const Renderer &) = delete;

    Uniforms &uniforms() { return uniforms_; } // render.rs:71

    // Renderer::render(&mesh, vertex_shader, fragment_shader), render.rs:98-114
    // texture_index: which bound texture a sampling shader reads (Uniforms::get_texture(index), uniform.rs:35-37)
    void render(const Mesh &mesh, VS vs, FS fs, uint32_t texture_index = 0) {
        if (mesh.vertices.size() != mesh.attributes.size())
            throw Error(RZ_E_INVALID, "Mesh: vertices and attributes must have the same length");
        const UniformBlock &b = uniforms_.block_;
        check(rz_write_block(ctx_, b.world.data(), b.view.data(), b.projection.data()));
        check(rz_render_host(ctx_, mesh.vertices.empty() ? nullptr : mesh.vertices[0].data(),
                             mesh.attributes.empty() ? nullptr : &mesh.attributes[0].r, (uint32_t)mesh.vertices.size(),
                             mesh.indices.data(), mesh.indices.size(), (uint32_t)vs, RZ_FS_WITH_TEXTURE((uint32_t)fs, texture_index)));
    }

    // Rasterizer::framebuffer(), rasterizer/mod.rs:520-522 (what Renderer::display hands to minifb)
    const std::vector<uint32_t> &framebuffer() {
        check(rz_framebuffer(ctx_, fb_.data(), nullptr));
        return fb_;
    }

    // Scissor rect [x0,x1) x [y0,y1): the extension sketched in Rasterizer::bounding_box (rasterizer/mod.rs:349-350)
    void set_scissor(uint32_t x0, uint32_t y0, uint32_t x1, uint32_t y1) { check(rz_set_scissor(ctx_, x0, y0, x1, y1)); }
    // N_MSAA_SAMPLES as a runtime value (rasterizer/mod.rs:23): 1, 2, 4 (the reference) or 8 samples per pixel
    void set_msaa(uint32_t samples) { check(rz_set_msaa(ctx_, samples)); }
    // guard-band clipping (rasterizer/mod.rs:417-419): side clip planes at |x|, |y| <= factor * w; 1 = the reference
    void set_guard_band(float factor) { check(rz_set_guard_band(ctx_, factor)); }
    // drop the draws recorded for the current frame
    void discard_frame() { check(rz_discard_frame(ctx_)); }

    // Headless replacement of Renderer::display (render.rs:116-127): the last framebuffer() as a file
    void save_png(const std::string &path) const { image::write_png(path, fb_.data(), width_, height_); }
    void save_ppm(const std::string &path) const { image::write_ppm(path, fb_.data(), width_, height_); }

    rz_counters_t counters() {
        rz_counters_t c;
        check(rz_counters(ctx_, &c));
        return c;
    }
    rz_ctx *ctx() { return ctx_; }
    size_t width() const { return width_; }
    size_t height() const { return height_; }

  private:
    void check(int rc) {
        if (rc != RZ_OK) throw Error(rc, rz_last_error(ctx_));
    }
    rz_ctx *ctx_ = nullptr;
    size_t width_, height_;
    Uniforms uniforms_;
    std::vector<uint32_t> fb_;
};

inline void Uniforms::bind_texture(size_t index, const Texture &tex) {
    const int rc = rz_bind_texture(ctx_, (uint32_t)index, tex.buf.data(), tex.width, tex.height, tex.texel_width);
    if (rc != RZ_OK) throw Error(rc, rz_last_error(ctx_)); // assert!(textures.len() == index), uniform.rs:31
    textures_.push_back(tex);
}

} // namespace rz
