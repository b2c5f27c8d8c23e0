// demo_main.cpp -- the reference's `Mode::Demo` frame (main.rs:93-105,129-182) at a fixed `elapsed`,
// driven through the C++ mirror of the crate API; writes the image the reference would hand to
// minifb as a PPM instead of opening a window.
//   g++ -std=c++17 -O2 -ffp-contract=off demo_main.cpp -o rz_demo -L.. -lrz_b200 -Wl,-rpath,'$ORIGIN/..'
#include <cstdio>
#include <cstdlib>

#include "rusterizer.hpp"

int main(int argc, char **argv) {
    const float elapsed = argc > 1 ? (float)std::atof(argv[1]) : 1.0f;
    const char *out = argc > 2 ? argv[2] : nullptr;
    const size_t W = 1280, H = 720; // main.rs:20-21
    try {
        rz::Renderer renderer(W, H);
        rz::Camera camera; // Camera::default(), camera.rs:46-54
        auto &block = renderer.uniforms().write_block();
        block.view = camera.get_view_matrix();
        block.projection = rz::project(1.0f, 200.0f, (float)H / (float)W, 1.57079637050628662f); // main.rs:137-142

        rz::Texture tex; // images/checkerboard.png decoded: 400x400 RGBA, 4x4 squares, top-left black
        tex.width = tex.height = 400;
        tex.texel_width = 4;
        tex.buf.resize(400 * 400 * 4);
        for (uint32_t y = 0; y < 400; y++)
            for (uint32_t x = 0; x < 400; x++) {
                const uint8_t v = ((x / 100 + y / 100) & 1) ? 255 : 0;
                uint8_t *p = &tex.buf[(y * 400 + x) * 4];
                p[0] = p[1] = p[2] = v;
                p[3] = 255;
            }
        renderer.uniforms().bind_texture(0, tex);

        const rz::Mesh meshes[2] = {rz::cube(1.0f), rz::sphere(0.5f)};
        const rz::Mat4 matrices[2] = {rz::rotate(elapsed, elapsed, 0.0f),
                                      rz::mul(rz::rotate(elapsed, 0.0f, 0.785398185253143311f), rz::translate(0.0f, 3.0f, 0.0f))};
        for (int i = 0; i < 2; i++) { // main.rs:170-173
            renderer.uniforms().write_block().world = matrices[i];
            renderer.render(meshes[i], rz::VS::MVP, rz::FS::Texture);
        }
        const std::vector<uint32_t> &fb = renderer.framebuffer(); // Renderer::display, render.rs:121

        size_t touched = 0;
        uint64_t sum = 0;
        for (uint32_t px : fb) {
            touched += px != 0xFF191919u;
            sum = sum * 1099511628211ull + px;
        }
        const rz_counters_t c = renderer.counters();
        std::printf("rz_demo: %zux%zu elapsed=%.3f tris_in=%llu samples_written=%llu touched_px=%zu checksum=%016llx\n", W, H,
                    elapsed, (unsigned long long)c.n_tris_in, (unsigned long long)c.n_samples_written, touched,
                    (unsigned long long)sum);
        if (out) {
            FILE *f = std::fopen(out, "wb");
            if (!f) return 2;
            std::fprintf(f, "P6\n%zu %zu\n255\n", W, H);
            for (uint32_t px : fb) {
                const unsigned char rgb[3] = {(unsigned char)(px >> 16), (unsigned char)(px >> 8), (unsigned char)px};
                std::fwrite(rgb, 1, 3, f);
            }
            std::fclose(f);
        }
        return touched > 1000 ? 0 : 3;
    } catch (const rz::Error &e) {
        std::fprintf(stderr, "rz_demo: error %d: %s\n", e.code, e.what());
        return 1;
    }
}
