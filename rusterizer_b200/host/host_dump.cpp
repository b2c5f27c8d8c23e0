// host_dump.cpp -- prints what the C++ host mirror (rusterizer.hpp) builds on the HOST: matrices, camera views and the
// crate's mesh generators, as hex floats / integers, one named record per line.  tests/test_host.py compares it with
// the Python mirror (mathx.py, camera.py, mesh.py): both restate math/mod.rs, math/transform.rs, camera.rs and mesh.rs.
// No GPU, no library call.
#include <cinttypes>
#include <cstdio>
#include <cstring>

#include "rusterizer.hpp"

static void put(const char *name, const float *v, size_t n) {
    std::printf("%s f", name);
    for (size_t i = 0; i < n; i++) {
        uint32_t b;
        std::memcpy(&b, v + i, 4);
        std::printf(" %08" PRIx32, b);
    }
    std::printf("\n");
}
static void put_mesh(const char *name, const rz::Mesh &m) {
    char buf[96];
    std::snprintf(buf, sizeof buf, "%s.vertices", name);
    put(buf, &m.vertices[0][0], m.vertices.size() * 3);
    std::snprintf(buf, sizeof buf, "%s.attributes", name);
    put(buf, &m.attributes[0].r, m.attributes.size() * 6);
    std::printf("%s.indices i", name);
    for (uint32_t i : m.indices) std::printf(" %" PRIu32, i);
    std::printf("\n");
}

int main() {
    const float pi = 3.14159274101257324f;
    put("project", rz::project(1.0f, 200.0f, 720.0f / 1280.0f, pi / 2.0f).data(), 16);
    put("project_1080", rz::project(1.0f, 200.0f, 1080.0f / 1920.0f, pi / 2.0f).data(), 16);
    put("rotate_110", rz::rotate(1.0f, 1.0f, 0.0f).data(), 16);
    put("rotate_0303", rz::rotate(0.3f, 0.3f, 0.0f).data(), 16);
    put("rotate_xyz", rz::rotate(2.5f, -0.7f, 4.1f).data(), 16);
    put("demo_sphere_world", rz::mul(rz::rotate(1.0f, 0.0f, pi / 4.0f), rz::translate(0.0f, 3.0f, 0.0f)).data(), 16);
    put("view_default", rz::Camera().get_view_matrix().data(), 16);
    for (int k : {1, 77, 256, 511, 1000}) {
        char name[32];
        std::snprintf(name, sizeof name, "view_orbit_%d", k);
        put(name, rz::Camera::orbit(2.0f * pi * (float)k / 1024.0f).get_view_matrix().data(), 16);
    }
    put_mesh("centered_quad", rz::centered_quad(9.0f));
    put_mesh("triangle", rz::triangle());
    put_mesh("cube", rz::cube(1.0f));
    put_mesh("sphere_default", rz::sphere(0.5f));
    put_mesh("sphere_65_33", rz::sphere(2.0f, 65, 33));
    return 0;
}
