// chunk_builder.hpp -- C++ host mirror of the reference's chunk scheduler (chunk_builder.rs,
// world.rs, player.rs), batched: the visible-set delta of one call goes to the GPU in ONE
// vx_frustum_cull (enqueue gate) and ONE vx_gen_mesh (every new chunk and every re-mesh).
// Same decisions as the Python executable specification wgpu-voxel-engine_b200/chunk_builder.py
// (SURVEY 8(f) rank 1); tests/test_scheduler.py checks both against a tick-by-tick restatement of
// the reference.  Header only, over voxb200.hpp.
#pragma once
#include <algorithm>
#include <cmath>
#include <map>
#include <set>

#include "voxb200.hpp"

namespace vox {

constexpr int32_t RENDER_DISTANCE = 5;   // world.rs:19

// world.rs:14-17 (HashMap<ChunkCoord3D, ..>; an ordered map keeps iteration deterministic)
struct World {
    std::map<ChunkCoord3D, std::shared_ptr<Chunk>> chunks;
    std::map<ChunkCoord3D, ChunkMesh> meshes;

    // The draw gate of World::render (world.rs:34-37) for every mesh in ONE launch.
    std::vector<ChunkCoord3D> visible_meshes(const Device &dev, const Frustum &frustum) const
    {
        std::vector<ChunkCoord3D> keys, out;
        for (const auto &kv : meshes) keys.push_back(kv.first);
        if (keys.empty()) return out;
        const std::vector<uint8_t> v = frustum.contains_many(dev, keys);
        for (size_t i = 0; i < keys.size(); ++i)
            if (v[i]) out.push_back(keys[i]);
        return out;
    }
};

// player.rs: position -> chunk coordinate, "entered a new chunk" predicate
class Player {
  public:
    std::array<float, 3> pos;
    ChunkCoord3D chunk;
    explicit Player(std::array<float, 3> eye) : pos(eye), chunk(to_chunk(eye)) {}   // player.rs:10-14
    static ChunkCoord3D to_chunk(const std::array<float, 3> &p)                      // coordinate.rs:33-38
    {
        return {(int32_t)std::floor(p[0]) >> 5, (int32_t)std::floor(p[1]) >> 5, (int32_t)std::floor(p[2]) >> 5};
    }
    void update_pos(std::array<float, 3> eye) { pos = eye; }                         // player.rs:16-18
    bool is_in_new_chunk_pos() const { return !(chunk == to_chunk(pos)); }           // player.rs:24-27
    void update_chunk_pos() { chunk = to_chunk(pos); }                               // player.rs:20-22
};

// chunk_builder.rs:178-214 (r = 1, preceded by the centre) and :217-247 (r = 2..RENDER_DISTANCE-1),
// in the reference's enumeration order
inline void ring_into(std::vector<ChunkCoord3D> &out, const ChunkCoord3D &pc, int32_t r)
{
    for (int32_t z = -r; z <= r; ++z) {
        out.push_back({pc.x + r, 0, pc.z + z});
        out.push_back({pc.x - r, 0, pc.z + z});
    }
    for (int32_t x = -r + 1; x < r; ++x) {
        out.push_back({pc.x + x, 0, pc.z + r});
        out.push_back({pc.x + x, 0, pc.z - r});
    }
}
inline std::vector<ChunkCoord3D> primary_coords(const ChunkCoord3D &pc)
{
    std::vector<ChunkCoord3D> out = {{pc.x, 0, pc.z}};
    ring_into(out, pc, 1);
    return out;
}
inline std::vector<ChunkCoord3D> ring_coords(const ChunkCoord3D &pc)
{
    std::vector<ChunkCoord3D> out;
    for (int32_t r = 2; r < RENDER_DISTANCE; ++r) ring_into(out, pc, r);
    return out;
}
// chunk_builder.rs:134-169 order: -x, +x, -z, +z
inline std::array<ChunkCoord3D, 4> neighbours(const ChunkCoord3D &c)
{
    return {{{c.x - 1, c.y, c.z}, {c.x + 1, c.y, c.z}, {c.x, c.y, c.z - 1}, {c.x, c.y, c.z + 1}}};
}

class ChunkGenerator {
  public:
    enum class Mode { Reference, Converged };
    // Reference: the world after build_chunks is EXACTLY the one the reference settles in (workers finishing
    // within the tick), including its artefact: a chunk dispatched while an adjacent chunk is still
    // undelivered keeps the border faces towards it (the 9-chunk burst of :178-214 is delivered one per
    // tick, :311).  Converged: every touched chunk is meshed against all loaded neighbours.
    explicit ChunkGenerator(Mode m = Mode::Reference) : mode(m) {}
    Mode mode;
    size_t last_batch = 0, gpu_calls = 0, replayed_ticks = 0;

    using Snapshot = std::vector<std::pair<ChunkCoord3D, uint8_t>>;   // insertion-ordered {coord, neighbour mask}

    // Symbolic replay of chunk_builder.rs:45-352 (bookkeeping only, no meshing): the neighbour snapshot
    // of every chunk at its last (re)build.  `loaded` is the post-eviction world.
    Snapshot replay(bool entered, const ChunkCoord3D &player_chunk, std::set<ChunkCoord3D> loaded,
                    const std::vector<ChunkCoord3D> &visible_ring)
    {
        Snapshot snap;
        std::vector<ChunkCoord3D> load_queue, rebuild_queue, in_process, chan;
        size_t mesh_chan = 0;
        auto has = [](const std::vector<ChunkCoord3D> &v, const ChunkCoord3D &c) {
            return std::find(v.begin(), v.end(), c) != v.end();
        };
        auto mask_now = [&](const ChunkCoord3D &c) {
            uint8_t m = 0;
            const auto nb = neighbours(c);
            for (int k = 0; k < 4; ++k)
                if (loaded.count(nb[k])) m |= (uint8_t)(1u << k);
            return m;
        };
        auto set_snap = [&](const ChunkCoord3D &c, uint8_t m) {
            for (auto &kv : snap)
                if (kv.first == c) {
                    kv.second = m;
                    return;
                }
            snap.push_back({c, m});
        };
        auto is_loaded = [&](const ChunkCoord3D &c) { return loaded.count(c) || has(load_queue, c) || has(in_process, c); };
        auto spawn = [&](const ChunkCoord3D &c) {
            set_snap(c, mask_now(c));              // adjacent_chunks snapshot at dispatch (:77, :122)
            in_process.push_back(c);
            chan.push_back(c);
        };
        size_t ticks = 0;
        bool first = true;
        for (;;) {
            ++ticks;
            if (first && entered)                  // load_primary_chunks (:171-215)
                for (const auto &c : primary_coords(player_chunk))
                    if (!is_loaded(c)) spawn(c);
            first = false;
            for (const auto &c : visible_ring)     // enqueue_chunks_in_frustum (:217-255)
                if (!is_loaded(c)) load_queue.push_back(c);
            if (!load_queue.empty()) {             // process_chunk_loading_queue (:69-87)
                const ChunkCoord3D c = load_queue.front();
                load_queue.erase(load_queue.begin());
                spawn(c);
            }
            if (in_process.empty() && !rebuild_queue.empty()) {   // process_rebuild_queue (:89-112)
                const ChunkCoord3D c = rebuild_queue.front();
                rebuild_queue.erase(rebuild_queue.begin());
                if (loaded.count(c)) {
                    set_snap(c, mask_now(c));
                    ++mesh_chan;
                }
            }
            if (!chan.empty()) {                   // update_world (:310-324)
                const ChunkCoord3D c = chan.front();
                chan.erase(chan.begin());
                loaded.insert(c);
                in_process.erase(std::find(in_process.begin(), in_process.end(), c));
                const ChunkCoord3D order[4] = {{c.x + 1, c.y, c.z}, {c.x - 1, c.y, c.z}, {c.x, c.y, c.z + 1},
                                               {c.x, c.y, c.z - 1}};               // :269-300
                for (const auto &n : order)
                    if (loaded.count(n)) rebuild_queue.push_back(n);
            }
            if (mesh_chan) --mesh_chan;            // (:325-331)
            load_queue.clear();                    // clean_up_queues (:302-308): sort + dedup
            std::sort(rebuild_queue.begin(), rebuild_queue.end());
            rebuild_queue.erase(std::unique(rebuild_queue.begin(), rebuild_queue.end()), rebuild_queue.end());
            bool ring_done = true;
            for (const auto &c : visible_ring) ring_done = ring_done && loaded.count(c);
            if (in_process.empty() && rebuild_queue.empty() && chan.empty() && !mesh_chan && ring_done) break;
            if (ticks > 100000) throw Error(VX_ERR_BAD_ARG, "scheduler replay did not settle");
        }
        replayed_ticks = ticks;
        return snap;
    }

    // chunk_builder.rs:45-67, batched
    void build_chunks(const Device &dev, Player &player, World &world, const Frustum &frustum)
    {
        const bool entered = player.is_in_new_chunk_pos();
        if (entered) player.update_chunk_pos();
        // filter_unseen_chunks (:334-352) first: in the reference every load after the first and every re-mesh
        // is dispatched against the post-eviction world, and no wanted chunk can touch an evicted one
        const ChunkCoord3D pc = player.chunk;
        for (auto it = world.chunks.begin(); it != world.chunks.end();) {
            const ChunkCoord3D &c = it->first;
            const bool keep = std::abs(c.x - pc.x) <= RENDER_DISTANCE && std::abs(c.z - pc.z) <= RENDER_DISTANCE &&
                              std::abs(c.y - pc.y) <= RENDER_DISTANCE;
            if (!keep) {
                world.meshes.erase(c);
                it = world.chunks.erase(it);
            } else {
                ++it;
            }
        }
        // enqueue gate (:217-255): ONE batched frustum test instead of one contains() per coordinate
        std::vector<ChunkCoord3D> cand, visible_ring, primaries;
        for (const auto &c : ring_coords(pc))
            if (!world.chunks.count(c)) cand.push_back(c);
        if (!cand.empty()) {
            const std::vector<uint8_t> vis = frustum.contains_many(dev, cand);
            for (size_t i = 0; i < cand.size(); ++i)
                if (vis[i]) visible_ring.push_back(cand[i]);
        }
        if (entered)
            for (const auto &c : primary_coords(pc))
                if (!world.chunks.count(c)) primaries.push_back(c);
        if (visible_ring.empty() && primaries.empty()) {
            last_batch = 0;
            replayed_ticks = 0;
            return;
        }
        Snapshot snap;
        if (mode == Mode::Reference) {
            std::set<ChunkCoord3D> loaded;
            for (const auto &kv : world.chunks) loaded.insert(kv.first);
            snap = replay(entered, pc, loaded, visible_ring);
        } else {
            std::vector<ChunkCoord3D> wanted;
            auto push_unique = [](std::vector<ChunkCoord3D> &v, const ChunkCoord3D &c) {
                if (std::find(v.begin(), v.end(), c) == v.end()) v.push_back(c);
            };
            for (const auto &c : primaries) push_unique(wanted, c);
            for (const auto &c : visible_ring) push_unique(wanted, c);
            std::set<ChunkCoord3D> present(wanted.begin(), wanted.end());
            for (const auto &kv : world.chunks) present.insert(kv.first);
            std::vector<ChunkCoord3D> touched = wanted;
            for (const auto &c : wanted)
                for (const auto &n : neighbours(c))
                    if (world.chunks.count(n)) push_unique(touched, n);
            for (const auto &c : touched) {
                uint8_t m = 0;
                const auto nb = neighbours(c);
                for (int k = 0; k < 4; ++k)
                    if (present.count(nb[k])) m |= (uint8_t)(1u << k);
                snap.push_back({c, m});
            }
        }
        std::vector<ChunkCoord3D> batch;
        std::vector<uint8_t> mask;
        for (const auto &kv : snap) {
            batch.push_back(kv.first);
            mask.push_back(kv.second);
        }
        BuiltChunks built = vox::build_chunks(dev, batch, mask.data());      // ONE GPU call (count pass + data pass)
        ++gpu_calls;
        last_batch = batch.size();
        for (size_t i = 0; i < batch.size(); ++i) {
            if (!world.chunks.count(batch[i]))                                // update_world (:310-324)
                world.chunks[batch[i]] = std::make_shared<Chunk>(std::move(built.chunks[i]));
            world.meshes[batch[i]] = std::move(built.meshes[i]);              // and :325-331 for re-meshes
        }
    }
};

}  // namespace vox
