"""Where do the ~150 us of a one-chunk vx_gen_mesh (config 1) go?"""
import os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import voxb200
ctx = voxb200.Context(devices=[0])
for n in (1, 9, 81):
    side = int(np.sqrt(n))
    c = np.array([(x, 0, z) for z in range(side) for x in range(side)], np.int32)
    m = np.full(n, 15 if n == 1 else 0, np.uint8)
    kw = dict(nbr_mask=m) if n == 1 else {}
    for _ in range(20):
        ctx.gen_mesh(c, **kw)
    t0 = time.perf_counter()
    for _ in range(200):
        r = ctx.gen_mesh(c, **kw)
    dt = (time.perf_counter() - t0) / 200
    ctx.set_timing(True)
    for _ in range(50):
        ctx.gen_mesh(c, **kw)
    kt = ctx.get_timing()
    ctx.set_timing(False)
    print("n=%d quads=%d host latency %.1f us; kernels (us): %s" % (
        n, r.total, dt * 1e6, {k: round(1e3 * ms / ln, 1) for k, (ms, ln) in kt.items() if ln}))
    t0 = time.perf_counter()
    for _ in range(200):
        ctx.fill(c, blocks=True)
    print("      vx_fill (blocks to host) %.1f us" % ((time.perf_counter() - t0) / 200 * 1e6))

# the two-call shape of the reference: Chunk::new (vx_fill) then create_mesh (vx_mesh) with 4 adjacent chunks
pos = voxb200.ChunkCoord3D(0, 0, 0)
data = voxb200.Chunk.new(pos, ctx)
adj = [voxb200.Chunk.new(voxb200.ChunkCoord3D(*c), ctx) for c in ((-1, 0, 0), (1, 0, 0), (0, 0, -1), (0, 0, 1))]
for _ in range(20):
    data.create_mesh(ctx, adj)
t0 = time.perf_counter()
for _ in range(200):
    mesh = data.create_mesh(ctx, adj)
print("Chunk.create_mesh with 4 adjacent chunks (vx_mesh, 5 x 32 KiB up): %.1f us, %d quads" % (
    (time.perf_counter() - t0) / 200 * 1e6, mesh.instances_len))
t0 = time.perf_counter()
for _ in range(200):
    voxb200.Chunk.new(pos, ctx)
print("Chunk.new (vx_fill, 32 KiB down): %.1f us" % ((time.perf_counter() - t0) / 200 * 1e6))

# Frustum::contains for one AABB, and the per-frame draw gate over ~80 meshes (world.rs:34-37)
planes = np.random.default_rng(0).standard_normal((1, 6, 4)).astype(np.float32)
for n in (1, 81, 4096):
    c = np.random.default_rng(n).integers(-40, 40, size=(n, 3)).astype(np.int32)
    for _ in range(20):
        ctx.frustum_cull(planes, c)
    t0 = time.perf_counter()
    for _ in range(300):
        ctx.frustum_cull(planes, c)
    print("vx_frustum_cull, 1 frame x %d AABBs: %.1f us" % (n, (time.perf_counter() - t0) / 300 * 1e6))
