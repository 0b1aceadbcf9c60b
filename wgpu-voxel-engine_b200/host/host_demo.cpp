// host_demo.cpp -- the reference's call site (chunk_builder.rs:80-84) written against the
// C++ host mirror.  Prints a digest the GPU test compares with the oracle.
//   g++ -std=c++17 host_demo.cpp -L.. -lvoxb200 -Wl,-rpath,'$ORIGIN/..' -o host_demo
#include <cstdio>
#include <cstring>
#include <memory>

#include "voxb200.hpp"

static uint32_t fnv(const void *p, size_t n, uint32_t h = 2166136261u)
{
    const unsigned char *b = (const unsigned char *)p;
    for (size_t i = 0; i < n; ++i) h = (h ^ b[i]) * 16777619u;
    return h;
}

int main()
{
    try {
        vox::Device device;
        auto pos = vox::ChunkCoord3D::new_(0, 0, 0);
        // let data = Arc::new(Chunk::new(pos)); let mesh = data.create_mesh(device.clone(), adjacent_chunks);
        auto data = std::make_shared<vox::Chunk>(vox::Chunk::new_(device, pos));
        auto left = std::make_shared<vox::Chunk>(vox::Chunk::new_(device, vox::ChunkCoord3D::new_(-1, 0, 0)));
        auto front = std::make_shared<vox::Chunk>(vox::Chunk::new_(device, vox::ChunkCoord3D::new_(0, 0, 1)));
        vox::ChunkMesh mesh = data->create_mesh(device, {left, nullptr, nullptr, front});
        std::printf("cubes_fnv %08x\n", fnv(data->cubes.data(), data->cubes.size()));
        std::printf("instances_len %zu\n", mesh.instances_len);
        uint32_t h = 2166136261u;
        for (const vx_instance &q : mesh.instance_buffer) {
            float m[16];
            std::memcpy(m, q.t_matrix, 64);
            for (float &v : m) v += 0.0f;                     // -0 -> +0 (not pinned)
            h = fnv(m, 64, h);
            h = fnv(&q.texture_index, 4, h);
        }
        std::printf("instances_fnv %08x\n", h);
        auto built = vox::build_chunks(device, {pos, vox::ChunkCoord3D::new_(-1, 0, 0)});
        std::printf("batched_len %zu %zu\n", built.meshes[0].instances_len, built.meshes[1].instances_len);
        float ident[16] = {1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1};
        auto frustum = vox::Frustum::new_(ident);
        std::printf("frustum %d %d\n", (int)frustum.contains(device, pos),
                    (int)frustum.contains(device, vox::ChunkCoord3D::new_(5, 0, 0)));
    } catch (const vox::Error &e) {
        std::printf("error %d: %s\n", e.code, e.what());
        return 1;
    }
    return 0;
}
