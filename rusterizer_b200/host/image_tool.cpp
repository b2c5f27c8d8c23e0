// image_tool.cpp -- exercises rz_image.hpp on the host (no GPU, no library): used by tests/test_host.py.
//   image_tool decode in.png out.raw     -> prints "W H C", writes the decoded u8[H][W][C]
//   image_tool encode W H in.u32 out.png -> encodes a 0xAARRGGBB framebuffer file as PNG
//   image_tool ppm    W H in.u32 out.ppm
//   g++ -std=c++17 -O2 image_tool.cpp -o image_tool
#include <cstdio>
#include <cstdlib>
#include <cstring>

#include "rz_image.hpp"

int main(int argc, char **argv) {
    try {
        if (argc == 4 && !std::strcmp(argv[1], "decode")) {
            const rz::image::Image img = rz::image::read_png(argv[2]);
            rz::image::write_file(argv[3], img.pixels.data(), img.pixels.size());
            std::printf("%u %u %u\n", img.width, img.height, img.channels);
            return 0;
        }
        if (argc == 6 && (!std::strcmp(argv[1], "encode") || !std::strcmp(argv[1], "ppm"))) {
            const size_t W = std::strtoul(argv[2], nullptr, 10), H = std::strtoul(argv[3], nullptr, 10);
            const std::vector<uint8_t> raw = rz::image::read_file(argv[4]);
            if (raw.size() != W * H * 4) throw std::runtime_error("framebuffer file has the wrong size");
            const uint32_t *fb = reinterpret_cast<const uint32_t *>(raw.data());
            if (argv[1][0] == 'e') rz::image::write_png(argv[5], fb, W, H);
            else rz::image::write_ppm(argv[5], fb, W, H);
            return 0;
        }
        std::fprintf(stderr, "usage: image_tool decode in.png out.raw | encode W H in.u32 out.png | ppm W H in.u32 out.ppm\n");
        return 2;
    } catch (const std::exception &e) {
        std::fprintf(stderr, "image_tool: %s\n", e.what());
        return 1;
    }
}
