"""The reference's un-batched call shape: T pool threads (uvth, main.rs:34-36), each building ONE chunk per call
(Chunk::new + create_mesh, chunk_builder.rs:80-84).  Aggregate chunks/s with one shared context vs one context
per thread."""
import os, sys, time, threading
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import voxb200

def run(T, own, reps=150):
    shared = voxb200.Context(devices=[0])
    ctxs = [voxb200.Context(devices=[0]) if own else shared for _ in range(T)]
    m = np.array([15], np.uint8)
    outs = [voxb200.PinnedBuffer(98304 * 68).view(np.uint8) for _ in range(T)]
    def worker(t):
        for r in range(reps):
            c = np.array([[t * 7 + r % 5, 0, r]], np.int32)
            ctxs[t].gen_mesh(c, nbr_mask=m, out=outs[t])
    for t in range(T):
        worker_t = threading.Thread(target=worker, args=(t,))
    ths = [threading.Thread(target=worker, args=(t,)) for t in range(T)]
    for c in ctxs[:1]:
        c.gen_mesh(np.array([[0, 0, 0]], np.int32), nbr_mask=m)
    t0 = time.perf_counter()
    for th in ths: th.start()
    for th in ths: th.join()
    dt = time.perf_counter() - t0
    return T * reps / dt

for T in (1, 4, 16):
    print("threads %2d: shared context %8.0f chunks/s   context per thread %8.0f chunks/s" % (T, run(T, False), run(T, True)))

# the two-call form (Chunk::new then create_mesh with its 4 neighbours) from T threads on one shared context
def run2(T, reps=60):
    ctx = voxb200.Context(devices=[0])
    def worker(t):
        for r in range(reps):
            pos = (t * 9 + r % 4, 0, r)
            data = voxb200.Chunk.new(voxb200.ChunkCoord3D(*pos), ctx)
            adj = [voxb200.Chunk.new(voxb200.ChunkCoord3D(pos[0] - 1, 0, pos[2]), ctx), None, None, None]
            data.create_mesh(ctx, adj)
    ths = [threading.Thread(target=worker, args=(t,)) for t in range(T)]
    worker(0)
    t0 = time.perf_counter()
    for th in ths: th.start()
    for th in ths: th.join()
    return T * reps / (time.perf_counter() - t0)
for T in (1, 4, 16):
    print("two-call form, threads %2d on one shared context: %8.0f chunks/s (2 fills + 1 mesh each)" % (T, run2(T)))
