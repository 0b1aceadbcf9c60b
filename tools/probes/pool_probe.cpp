// The reference's un-batched call shape from NATIVE pool threads (uvth, main.rs:34-36): T threads share one
// context and each builds ONE chunk per call (the fused form of Chunk::new + create_mesh, chunk_builder.rs:80-84).
// The Python twin (unbatched_pool_probe.py) is limited by the interpreter lock; this is what a compiled host sees.
//   g++ -std=c++17 -O2 -pthread tools/probes/pool_probe.cpp -Iinclude -Lwgpu-voxel-engine_b200 -lvoxb200 \
//       -Wl,-rpath,$PWD/wgpu-voxel-engine_b200 -o tools/probes/pool_probe
#include <atomic>
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
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

// one batched call (the INTEGRATION section 3 shape: ~80-chunk visible-set delta) into fresh pageable memory (a new Vec
// per call: first-touch page faults included), reused pageable memory, and pinned memory
static void batch(vx_ctx *ctx, int side)
{
    const size_t n = (size_t)side * side, cap = n * 8192;
    std::vector<int32_t> c(3 * n);
    for (size_t i = 0; i < n; ++i) { c[3 * i] = (int)(i % side); c[3 * i + 1] = 0; c[3 * i + 2] = (int)(i / side); }
    std::vector<uint32_t> counts(n);
    std::vector<uint64_t> off(n + 1);
    uint64_t total = 0;
    void *pinned = nullptr;
    if (vx_host_alloc(&pinned, cap * sizeof(vx_instance))) exit(2);
    vx_instance *reused = (vx_instance *)malloc(cap * sizeof(vx_instance));
    uint8_t *reused_blk = (uint8_t *)malloc(n * 32768);
    void *pinned_blk = nullptr;
    if (vx_host_alloc(&pinned_blk, n * 32768)) exit(2);
    auto time_it = [&](const char *name, int kind) {
        double best = 1e9;
        for (int r = 0; r < 6; ++r) {
            vx_instance *dst = kind == 0 ? (vx_instance *)malloc(cap * sizeof(vx_instance)) : (kind == 1 ? reused : (vx_instance *)pinned);
            uint8_t *blk = kind == 0 ? (uint8_t *)malloc(n * 32768) : (kind == 1 ? reused_blk : (uint8_t *)pinned_blk);
            const auto t0 = std::chrono::steady_clock::now();
            if (vx_gen_mesh(ctx, c.data(), n, nullptr, nullptr, counts.data(), off.data(), dst, cap, &total, blk)) {
                fprintf(stderr, "error: %s\n", vx_last_error());
                exit(1);
            }
            const double dt = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
            if (r > 0 && dt < best) best = dt;
            if (kind == 0) { free(dst); free(blk); }
        }
        printf("one call, %4zu chunks, %6.1f MB of instances (+ blocks) into %-22s: %8.0f us = %6.1f GB/s, %8.0f chunks/s\n", n,
               total * 68 / 1e6, name, best * 1e6, total * 68 / best / 1e9, n / best);
    };
    time_it("fresh pageable memory", 0);
    time_it("reused pageable memory", 1);
    time_it("pinned memory", 2);
    free(reused);
    free(reused_blk);
    vx_host_free(pinned_blk);
    vx_host_free(pinned);
}

// vx_mesh of a large list: block grids uploaded from pageable or pinned memory, instances into pinned memory
static void mesh_batch(vx_ctx *ctx, int side)
{
    const size_t n = (size_t)side * side, cap = n * 8192;
    std::vector<int32_t> c(3 * n);
    for (size_t i = 0; i < n; ++i) { c[3 * i] = (int)(i % side); c[3 * i + 1] = 0; c[3 * i + 2] = (int)(i / side); }
    std::vector<uint32_t> counts(n);
    std::vector<uint64_t> off(n + 1);
    uint64_t total = 0;
    void *pinned = nullptr, *pin_blk = nullptr;
    if (vx_host_alloc(&pinned, cap * sizeof(vx_instance)) || vx_host_alloc(&pin_blk, n * 32768)) exit(2);
    std::vector<uint8_t> blk(n * 32768);
    if (vx_fill(ctx, c.data(), n, nullptr, blk.data(), nullptr, nullptr, nullptr)) exit(3);
    memcpy(pin_blk, blk.data(), n * 32768);
    for (int kind = 0; kind < 2; ++kind) {
        double best = 1e9;
        for (int r = 0; r < 5; ++r) {
            const auto t0 = std::chrono::steady_clock::now();
            if (vx_mesh(ctx, c.data(), n, kind ? (const uint8_t *)pin_blk : blk.data(), nullptr, nullptr, counts.data(), off.data(),
                        (vx_instance *)pinned, cap, &total)) {
                fprintf(stderr, "error: %s\n", vx_last_error());
                exit(1);
            }
            const double dt = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
            if (r > 0 && dt < best) best = dt;
        }
        printf("vx_mesh, %4zu chunks, %5.1f MB of blocks up from %-8s memory, %6.1f MB of instances into pinned memory: %8.0f us, %8.0f chunks/s\n",
               n, n * 32768 / 1e6, kind ? "pinned" : "pageable", total * 68 / 1e6, best * 1e6, n / best);
    }
    vx_host_free(pinned);
    vx_host_free(pin_blk);
}

int main(int argc, char **argv)
{
    if (argc > 1) {   // any argument: the batched-call measurement only
        vx_ctx *ctx = nullptr;
        const int dev = 0;
        if (vx_create(&dev, 1, &ctx)) { fprintf(stderr, "%s\n", vx_last_error()); return 1; }
        for (int side : {9, 32}) batch(ctx, side);
        for (int side : {9, 32}) mesh_batch(ctx, side);
        vx_destroy(ctx);
        return 0;
    }

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
