// The reference's un-batched call shape from NATIVE pool threads (uvth, main.rs:34-36): T threads share one
// context and each builds ONE chunk per call (the fused form of Chunk::new + create_mesh, chunk_builder.rs:80-84).
// The Python twin (unbatched_pool_probe.py) is limited by the interpreter lock; this is what a compiled host sees.
//   g++ -std=c++17 -O2 -pthread tools/probes/pool_probe.cpp -Iinclude -Lwgpu-voxel-engine_b200 -lvoxb200 \
//       -Wl,-rpath,$PWD/wgpu-voxel-engine_b200 -o tools/probes/pool_probe
#include <atomic>
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <thread>
#include <vector>

#include "voxb200.h"

static double run(vx_ctx *ctx, int T, int reps, bool pinned)
{
    std::vector<void *> outs(T);
    for (int t = 0; t < T; ++t) {
        if (pinned) { if (vx_host_alloc(&outs[t], (size_t)VX_MAX_QUADS_PER_CHUNK * sizeof(vx_instance))) exit(2); }
        else outs[t] = malloc((size_t)VX_MAX_QUADS_PER_CHUNK * sizeof(vx_instance));
    }
    std::atomic<int> bad{0};
    auto worker = [&](int t) {
        const uint8_t mask = 15;
        for (int r = 0; r < reps; ++r) {
            const int32_t c[3] = {t * 7 + r % 5, 0, r};
            uint32_t count;
            uint64_t off[2], total;
            if (vx_gen_mesh(ctx, c, 1, nullptr, &mask, &count, off, (vx_instance *)outs[t], VX_MAX_QUADS_PER_CHUNK, &total,
                            nullptr) != 0) { bad++; return; }
        }
    };
    worker(0);
    const auto t0 = std::chrono::steady_clock::now();
    std::vector<std::thread> th;
    for (int t = 0; t < T; ++t) th.emplace_back(worker, t);
    for (auto &x : th) x.join();
    const double dt = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
    for (int t = 0; t < T; ++t) { if (pinned) vx_host_free(outs[t]); else free(outs[t]); }
    if (bad) { fprintf(stderr, "error: %s\n", vx_last_error()); exit(1); }
    return T * (double)reps / dt;
}

// the two-call form of the minimal patch: Chunk::new -> vx_fill of the chunk, then
// create_mesh -> vx_mesh of the chunk and its four neighbours' block grids (already in World.chunks) from pageable memory
static double run2(vx_ctx *ctx, int T, int reps)
{
    std::atomic<int> bad{0};
    auto worker = [&](int t) {
        std::vector<uint8_t> blocks(5 * 32768);
        std::vector<vx_instance> out(VX_MAX_QUADS_PER_CHUNK);
        const int32_t nb[5][3] = {{t * 9, 0, 3}, {t * 9 - 1, 0, 3}, {t * 9 + 1, 0, 3}, {t * 9, 0, 2}, {t * 9, 0, 4}};
        if (vx_fill(ctx, &nb[0][0], 5, nullptr, blocks.data(), nullptr, nullptr, nullptr) != 0) { bad++; return; }
        const int32_t wiring[20] = {1, 2, 3, 4, -1, -1, -1, -1, -1, -1, -1, -1, -1, -1, -1, -1, -1, -1, -1, -1};
        const uint8_t mesh_mask[5] = {1, 0, 0, 0, 0};
        for (int r = 0; r < reps; ++r) {
            uint32_t counts[5];
            uint64_t off[6], total;
            if (vx_fill(ctx, &nb[0][0], 1, nullptr, blocks.data(), nullptr, nullptr, nullptr) != 0 ||
                vx_mesh(ctx, &nb[0][0], 5, blocks.data(), wiring, mesh_mask, counts, off, out.data(), out.size(), &total) != 0) {
                bad++;
                return;
            }
        }
    };
    worker(0);
    const auto t0 = std::chrono::steady_clock::now();
    std::vector<std::thread> th;
    for (int t = 0; t < T; ++t) th.emplace_back(worker, t);
    for (auto &x : th) x.join();
    const double dt = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
    if (bad) { fprintf(stderr, "error: %s\n", vx_last_error()); exit(1); }
    return T * (double)reps / dt;
}

int main()
{
    vx_ctx *ctx = nullptr;
    const int dev = 0;
    if (vx_create(&dev, 1, &ctx)) { fprintf(stderr, "%s\n", vx_last_error()); return 1; }
    for (int T : {1, 2, 4, 8, 16, 32})
        printf("threads %2d on one shared context: %8.0f chunks/s (pinned output)  %8.0f chunks/s (pageable output)\n", T,
               run(ctx, T, 400, true), run(ctx, T, 400, false));
    for (int T : {1, 4, 16})
        printf("two-call form (vx_fill + vx_mesh of 5 block grids, pageable), threads %2d: %8.0f chunks/s\n", T, run2(ctx, T, 300));
    vx_destroy(ctx);
    return 0;
}
