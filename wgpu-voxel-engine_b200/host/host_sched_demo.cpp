// host_sched_demo.cpp -- drives the C++ scheduler mirror (chunk_builder.hpp) from stdin so the tests can
// compare it with the Python executable specification and the tick-by-tick reference restatement.
//   host_sched_demo replay                 (no GPU) stdin: entered px py pz | NL (x y z)* | NV (x y z)*
//   host_sched_demo run reference|converged  (GPU)  stdin: one pose per line: ex ey ez p0 .. p23
#include <cstdio>
#include <cstring>
#include <iostream>
#include <string>

#include "chunk_builder.hpp"

static uint32_t fnv(const void *p, size_t n, uint32_t h)
{
    const unsigned char *b = (const unsigned char *)p;
    for (size_t i = 0; i < n; ++i) h = (h ^ b[i]) * 16777619u;
    return h;
}

static int do_replay()
{
    int entered;
    vox::ChunkCoord3D pc;
    size_t nl, nv;
    std::cin >> entered >> pc.x >> pc.y >> pc.z >> nl;
    std::set<vox::ChunkCoord3D> loaded;
    for (size_t i = 0; i < nl; ++i) {
        vox::ChunkCoord3D c;
        std::cin >> c.x >> c.y >> c.z;
        loaded.insert(c);
    }
    std::cin >> nv;
    std::vector<vox::ChunkCoord3D> ring(nv);
    for (auto &c : ring) std::cin >> c.x >> c.y >> c.z;
    vox::ChunkGenerator gen;
    const auto snap = gen.replay(entered != 0, pc, loaded, ring);
    std::printf("ticks %zu\n", gen.replayed_ticks);
    for (const auto &kv : snap) std::printf("snap %d %d %d %u\n", kv.first.x, kv.first.y, kv.first.z, (unsigned)kv.second);
    return 0;
}

static int do_run(const std::string &mode)
{
    vox::Device device;
    vox::ChunkGenerator gen(mode == "converged" ? vox::ChunkGenerator::Mode::Converged : vox::ChunkGenerator::Mode::Reference);
    vox::World world;
    std::unique_ptr<vox::Player> player;
    float v[27];
    int pose = 0;
    while (std::cin >> v[0]) {
        for (int i = 1; i < 27; ++i) std::cin >> v[i];
        if (!player) player.reset(new vox::Player({v[0], v[1], v[2]}));
        player->update_pos({v[0], v[1], v[2]});
        const vox::Frustum frustum = vox::Frustum::from_planes(v + 3);
        gen.build_chunks(device, *player, world, frustum);
        const size_t batch = gen.last_batch, ticks = gen.replayed_ticks;
        gen.build_chunks(device, *player, world, frustum);            // idempotent once settled
        uint32_t h = 2166136261u;
        size_t quads = 0;
        for (const auto &kv : world.chunks) {                          // ordered by (x, y, z)
            h = fnv(&kv.first, 12, h);
            h = fnv(kv.second->cubes.data(), kv.second->cubes.size(), h);
            const vox::ChunkMesh &m = world.meshes.at(kv.first);
            quads += m.instances_len;
            for (const vx_instance &q : m.instance_buffer) {
                float t[16];
                std::memcpy(t, q.t_matrix, 64);
                for (float &f : t) f += 0.0f;                          // -0 -> +0 (not pinned)
                h = fnv(t, 64, h);
                h = fnv(&q.texture_index, 4, h);
            }
        }
        const size_t drawn = world.visible_meshes(device, frustum).size();
        std::printf("pose %d chunks %zu meshes %zu batch %zu second_batch %zu ticks %zu quads %zu drawn %zu digest %08x\n", pose++,
                    world.chunks.size(), world.meshes.size(), batch, gen.last_batch, ticks, quads, drawn, h);
    }
    return 0;
}

int main(int argc, char **argv)
{
    try {
        if (argc >= 2 && std::string(argv[1]) == "replay") return do_replay();
        if (argc >= 3 && std::string(argv[1]) == "run") return do_run(argv[2]);
        std::fprintf(stderr, "usage: host_sched_demo replay | run reference|converged\n");
        return 2;
    } catch (const vox::Error &e) {
        std::printf("error %d: %s\n", e.code, e.what());
        return 1;
    }
}
