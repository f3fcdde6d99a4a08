// app_scene.cpp -- the scene construction of the reference's App.tick `.window_open` handler
// (App/App.zig:54-91) on the B200 library: one terrain from a 16-bit heightmap, then polygon1 and
// polygon2, through the C++ mirror of the Zig interface, linked against the STATIC library the way
// build.zig would link it.  Prints FNV-1a hashes of the produced buffers so a test can compare them
// with the oracle's.
//
//   app_scene <heightmap.u16 raw file> <n> [offset prime]
#include <cstdio>
#include <cstdlib>
#include <fstream>
#include <vector>

#include "../myrenderer_b200/host/mr_host.hpp"

static uint64_t fnv1a(const void* p, size_t n) {
    const unsigned char* b = static_cast<const unsigned char*>(p);
    uint64_t h = 1469598103934665603ull;
    for (size_t i = 0; i < n; ++i) h = (h ^ b[i]) * 1099511628211ull;
    return h;
}

int main(int argc, char** argv) {
    if (argc < 3) {
        std::fprintf(stderr, "usage: %s heightmap.u16 n [offset prime]\n", argv[0]);
        return 2;
    }
    const uint32_t n = static_cast<uint32_t>(std::atoi(argv[2]));
    std::vector<uint16_t> pixels(static_cast<size_t>(n) * n);
    std::ifstream f(argv[1], std::ios::binary);
    if (!f.read(reinterpret_cast<char*>(pixels.data()), static_cast<std::streamsize>(pixels.size() * 2))) {
        std::fprintf(stderr, "cannot read %s\n", argv[1]);
        return 2;
    }
    uint32_t op[2] = {1, 1};
    const bool have_op = argc >= 5;
    if (have_op) {
        op[0] = static_cast<uint32_t>(std::atoi(argv[3]));
        op[1] = static_cast<uint32_t>(std::atoi(argv[4]));
    }
    try {
        mr::Context ctx(0);
        mr::Terrain terrain(ctx);
        mr::Polygon polygon(ctx);
        // app.terrain = try terrain.create_terrain(core, full_heightmap_dir);   App.zig:64
        mr::TerrainMesh mesh = terrain.create_terrain(pixels.data(), n);
        std::printf("terrain n=%u vertices=%u indices=%zu bbox=(%g,%g,%g)-(%g,%g,%g) vtx_hash=%016llx idx_hash=%016llx\n", n,
                    mesh.vertex_buffer.vertex_count, mesh.index_buffer.size(), mesh.bounding_box_p0[0], mesh.bounding_box_p0[1],
                    mesh.bounding_box_p0[2], mesh.bounding_box_p1[0], mesh.bounding_box_p1[1], mesh.bounding_box_p1[2],
                    (unsigned long long)fnv1a(mesh.vertex_buffer.mapped.data(), mesh.vertex_buffer.mapped.size()),
                    (unsigned long long)fnv1a(mesh.index_buffer.data(), mesh.index_buffer.size() * 4));
        // app.polygon1 / app.polygon2   App.zig:68-83
        const std::vector<mr::Point> polygon1 = {{{62.742857f, 106.97143f}}, {{93.085712f, 65.828571f}}, {{147.08571f, 85.628572f}},
                                                 {{122.14285f, 144.77143f}}, {{102.34286f, 93.857142f}}, {{79.199998f, 130.37143f}},
                                                 {{81.00000f, 105.17143f}}};
        const std::vector<mr::Point> polygon2 = {{{10.0f, 10.0f}}, {{40.0f, 10.0f}}, {{40.0f, 40.0f}}, {{10.0f, 40.0f}}};
        int which = 1;
        for (const auto* poly : {&polygon1, &polygon2}) {
            mr::PolygonObj o = polygon.create_polygon(*poly, 0x5EED, which, have_op ? op : nullptr);
            std::printf("polygon%d n=%zu status=%u ntri=%u vertex_count=%u bbox=(%g,%g)-(%g,%g) vtx_hash=%016llx\n", which, poly->size(),
                        o.status, o.ntri, o.vertex_buffer.vertex_count, o.bounding_box_p0[0], o.bounding_box_p0[1], o.bounding_box_p1[0],
                        o.bounding_box_p1[1], (unsigned long long)fnv1a(o.vertex_buffer.mapped.data(), o.vertex_buffer.mapped.size()));
            ++which;
        }
        // the callback form, Triangulation.create_polygon(points, ctx, emit)
        mr::Triangulation tri(ctx);
        std::vector<mr::Point> emitted;
        std::function<void(std::vector<mr::Point>&, mr::Point)> emit = [](std::vector<mr::Point>& c, mr::Point p) { c.push_back(p); };
        const uint32_t lin[2] = {0, 1};
        tri.create_polygon<std::vector<mr::Point>>(polygon2, emitted, emit, 0, 0, lin);
        std::printf("emit:");
        for (auto& p : emitted) std::printf(" (%g,%g)", p.v[0], p.v[1]);
        std::printf("\n");
    } catch (const mr::Error& e) {
        std::fprintf(stderr, "error %d: %s\n", e.code, e.what());
        return 1;
    }
    return 0;
}
