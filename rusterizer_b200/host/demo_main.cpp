// demo_main.cpp -- the reference's `Mode::Demo` frame (main.rs:93-105,129-182) at a fixed `elapsed`,
// driven through the C++ mirror of the crate API; writes the image the reference would hand to
// minifb as a PNG/PPM instead of opening a window.   usage: rz_demo [elapsed [out.png|out.ppm [texture.png]]]
//   g++ -std=c++17 -O2 -ffp-contract=off demo_main.cpp -o rz_demo -L.. -lrz_b200 -Wl,-rpath,'$ORIGIN/..'
#include <cstdio>
#include <cstdlib>
#include <string>

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

        // images/checkerboard.png (main.rs:144): decoded from a file when given, else regenerated procedurally
        const rz::Texture tex = argc > 3 ? rz::Texture::from_png_file(argv[3]) : rz::Texture::checkerboard();
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
        if (out) { // .png or .ppm by extension
            const std::string path(out);
            if (path.size() > 4 && path.substr(path.size() - 4) == ".png") renderer.save_png(path);
            else renderer.save_ppm(path);
        }
        return touched > 1000 ? 0 : 3;
    } catch (const rz::Error &e) {
        std::fprintf(stderr, "rz_demo: error %d: %s\n", e.code, e.what());
        return 1;
    } catch (const std::exception &e) {
        std::fprintf(stderr, "rz_demo: %s\n", e.what());
        return 1;
    }
}
