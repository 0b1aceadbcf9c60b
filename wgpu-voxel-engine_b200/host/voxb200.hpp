// voxb200.hpp -- C++ host mirror of the reference's Rust interface for the hot
// path, over the C ABI (include/voxb200.h).  Header only; link libvoxb200.so.
//
// The reference is compiled Rust and there is no rustc in this image, so the
// host side above the C ABI is written in C++ with the reference's names,
// argument meaning and error behaviour:
//   Chunk::new(position)                            chunk.rs:21-25
//   Chunk::create_mesh(device, adjacent_chunks)     chunk.rs:48-57
//   ChunkMesh { instance_buffer, instances_len }    chunk.rs:248-266 (bytes for create_buffer_init)
//   Frustum::new / update / contains                frustum_culling.rs:10-45
//   ChunkGenerator worker closure, batched          chunk_builder.rs:80-84
//   (the batched scheduler itself -- World, Player, ChunkGenerator -- is chunk_builder.hpp)
// Failures throw vox::Error (the reference panics: chunk_builder.rs:83,107,128).
#pragma once
#include <algorithm>
#include <array>
#include <cstdint>
#include <memory>
#include <optional>
#include <stdexcept>
#include <string>
#include <vector>

#include "../../include/voxb200.h"

namespace vox {

struct Error : std::runtime_error {
    int code;
    Error(int c, const std::string &m) : std::runtime_error(m), code(c) {}
};
inline void check(int rc)
{
    if (rc != VX_OK) throw Error(rc, vx_last_error());
}

// coordinate.rs:41-46 (derived Ord is lexicographic x, y, z)
struct ChunkCoord3D {
    int32_t x, y, z;
    static ChunkCoord3D new_(int32_t x, int32_t y, int32_t z) { return {x, y, z}; }
    bool operator==(const ChunkCoord3D &o) const { return x == o.x && y == o.y && z == o.z; }
    bool operator<(const ChunkCoord3D &o) const
    {
        return x != o.x ? x < o.x : (y != o.y ? y < o.y : z < o.z);
    }
    // coordinate.rs:60-65 (wrapping i32 multiply)
    std::array<int32_t, 3> to_world_position_i32() const
    {
        return {(int32_t)((uint32_t)x * 32u), (int32_t)((uint32_t)y * 32u), (int32_t)((uint32_t)z * 32u)};
    }
};

// The analogue of Arc<wgpu::Device> at the call sites: the GPU context the work runs on.
class Device {
  public:
    explicit Device(int n_devices = 1) { check(vx_create(nullptr, n_devices, &ctx_)); }
    explicit Device(const std::vector<int> &ids) { check(vx_create(ids.data(), (int)ids.size(), &ctx_)); }
    ~Device() { vx_destroy(ctx_); }
    Device(const Device &) = delete;
    Device &operator=(const Device &) = delete;
    vx_ctx *raw() const { return ctx_; }

  private:
    vx_ctx *ctx_ = nullptr;
};

// chunk.rs:248-251; instance_buffer holds the InstanceRaw bytes ChunkMesh::new uploads (chunk.rs:255-260)
struct ChunkMesh {
    std::vector<vx_instance> instance_buffer;
    size_t instances_len = 0;
};

class Chunk {
  public:
    ChunkCoord3D position;
    std::vector<uint8_t> cubes;   // Vec<Cube>: CubeType per voxel, idx = x + 32 z + 1024 y

    // chunk.rs:21-25 (noise parameters default to perlin_noise.rs:14-18)
    static Chunk new_(const Device &dev, ChunkCoord3D position, const vx_noise_params *params = nullptr)
    {
        Chunk c;
        c.position = position;
        c.cubes.resize(VX_VOXELS);
        const int32_t xyz[3] = {position.x, position.y, position.z};
        check(vx_fill(dev.raw(), xyz, 1, params, c.cubes.data(), nullptr, nullptr, nullptr));
        return c;
    }

    // chunk.rs:48-57; adjacent_chunks = [-x, +x, -z, +z], nullopt/null = None (chunk_builder.rs:134-169)
    ChunkMesh create_mesh(const Device &dev, const std::vector<std::shared_ptr<Chunk>> &adjacent_chunks) const
    {
        if (adjacent_chunks.size() != 4) throw Error(VX_ERR_BAD_ARG, "adjacent_chunks must have 4 entries");
        std::vector<int32_t> coords = {position.x, position.y, position.z};
        std::vector<uint8_t> blocks(cubes);
        std::vector<int32_t> nbr(4, -1);
        for (int k = 0; k < 4; ++k) {
            if (!adjacent_chunks[k]) continue;
            nbr[k] = (int32_t)(coords.size() / 3);
            const Chunk &a = *adjacent_chunks[k];
            coords.insert(coords.end(), {a.position.x, a.position.y, a.position.z});
            blocks.insert(blocks.end(), a.cubes.begin(), a.cubes.end());
        }
        const size_t n = coords.size() / 3;
        std::vector<int32_t> nbr_idx(n * 4, -1);
        std::copy(nbr.begin(), nbr.end(), nbr_idx.begin());
        std::vector<uint8_t> mesh_mask(n, 0);
        mesh_mask[0] = 1;
        std::vector<uint32_t> counts(n);
        std::vector<uint64_t> offsets(n + 1);
        ChunkMesh mesh;
        mesh.instance_buffer.resize(VX_MAX_QUADS_PER_CHUNK);
        uint64_t total = 0;
        check(vx_mesh(dev.raw(), coords.data(), n, blocks.data(), nbr_idx.data(), mesh_mask.data(), counts.data(),
                      offsets.data(), mesh.instance_buffer.data(), mesh.instance_buffer.size(), &total));
        mesh.instance_buffer.resize(total);
        mesh.instances_len = total;
        return mesh;
    }
};

// The three worker closures of chunk_builder.rs:80-84 / :105-108 / :125-129 for a whole
// coordinate list in one launch: Chunk::new + create_mesh per entry.
struct BuiltChunks {
    std::vector<Chunk> chunks;
    std::vector<ChunkMesh> meshes;
};
// Pinned (page-locked) host memory from vx_host_alloc, grown on demand, one per thread: a download into it runs at
// link speed (~50 GB/s), a download into a fresh std::vector is bound by the first-touch page faults of the destination.
class PinnedArena {
  public:
    PinnedArena() = default;
    PinnedArena(const PinnedArena &) = delete;
    PinnedArena &operator=(const PinnedArena &) = delete;
    ~PinnedArena() { if (p_) vx_host_free(p_); }
    unsigned char *ensure(size_t bytes)
    {
        if (bytes > bytes_) {
            if (p_) vx_host_free(p_);
            p_ = nullptr;
            bytes_ = 0;
            check(vx_host_alloc(&p_, bytes + bytes / 4));
            bytes_ = bytes + bytes / 4;
        }
        return static_cast<unsigned char *>(p_);
    }

  private:
    void *p_ = nullptr;
    size_t bytes_ = 0;
};

// ONE pass: the instance buffer is sized from the terrain's typical density (about 4.6 k quads per chunk with
// neighbours, at most 7.9 k without) and the call is repeated only if the library answers VX_ERR_CAPACITY, in which
// case counts, offsets and the exact total are already valid.
inline BuiltChunks build_chunks(const Device &dev, const std::vector<ChunkCoord3D> &positions,
                                const uint8_t *nbr_mask = nullptr, const vx_noise_params *params = nullptr)
{
    static thread_local PinnedArena arena;
    const size_t n = positions.size();
    std::vector<int32_t> coords(n * 3);
    for (size_t i = 0; i < n; ++i) {
        coords[3 * i] = positions[i].x;
        coords[3 * i + 1] = positions[i].y;
        coords[3 * i + 2] = positions[i].z;
    }
    std::vector<uint32_t> counts(n);
    std::vector<uint64_t> offsets(n + 1);
    uint64_t total = 0;
    size_t cap = n * 8192 + 1;
    const vx_instance *inst = nullptr;
    const uint8_t *blocks = nullptr;
    for (;;) {
        const size_t inst_bytes = cap * sizeof(vx_instance);
        unsigned char *base = arena.ensure(inst_bytes + n * VX_VOXELS);        // [instances][cube ids]
        const int rc = vx_gen_mesh(dev.raw(), coords.data(), n, params, nbr_mask, counts.data(), offsets.data(),
                                   reinterpret_cast<vx_instance *>(base), cap, &total, base + inst_bytes);
        if (rc == VX_ERR_CAPACITY) {
            cap = (size_t)total;       // exact now
            continue;
        }
        check(rc);
        inst = reinterpret_cast<const vx_instance *>(base);
        blocks = base + inst_bytes;
        break;
    }
    BuiltChunks out;
    out.chunks.resize(n);
    out.meshes.resize(n);
    for (size_t i = 0; i < n; ++i) {
        out.chunks[i].position = positions[i];
        out.chunks[i].cubes.assign(blocks + i * VX_VOXELS, blocks + (i + 1) * VX_VOXELS);
        out.meshes[i].instance_buffer.assign(inst + offsets[i], inst + offsets[i + 1]);
        out.meshes[i].instances_len = counts[i];
    }
    return out;
}

// frustum_culling.rs:6-45
class Frustum {
  public:
    // Frustum::new(&camera): camera.global_matrix as column-major Matrix4<f32> data
    static Frustum new_(const float global_matrix[16])
    {
        Frustum f;
        f.update(global_matrix);
        return f;
    }
    // from six planes (near, far, left, right, top, bottom; a, b, c, d each) computed elsewhere
    static Frustum from_planes(const float planes[24])
    {
        Frustum f;
        std::copy(planes, planes + 24, f.planes_.begin());
        return f;
    }
    void update(const float global_matrix[16]) { check(vx_planes_from_matrix(global_matrix, planes_.data())); }
    bool contains(const Device &dev, const ChunkCoord3D &pos) const
    {
        const int32_t xyz[3] = {pos.x, pos.y, pos.z};
        uint8_t v = 0;
        check(vx_frustum_cull(dev.raw(), planes_.data(), 1, xyz, 1, &v));
        return v != 0;
    }
    // batched form: one launch for a whole list (world.rs:34-37, chunk_builder.rs:217-255)
    std::vector<uint8_t> contains_many(const Device &dev, const std::vector<ChunkCoord3D> &pos) const
    {
        std::vector<int32_t> coords(pos.size() * 3);
        for (size_t i = 0; i < pos.size(); ++i) {
            coords[3 * i] = pos[i].x;
            coords[3 * i + 1] = pos[i].y;
            coords[3 * i + 2] = pos[i].z;
        }
        std::vector<uint8_t> v(pos.size());
        check(vx_frustum_cull(dev.raw(), planes_.data(), 1, coords.data(), pos.size(), v.data()));
        return v;
    }
    const std::array<float, 24> &planes() const { return planes_; }

  private:
    std::array<float, 24> planes_{};
};

}  // namespace vox
