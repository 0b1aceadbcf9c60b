#!/usr/bin/env python
"""One host process, one context over k GPUs (vx_create(NULL, k)): a single vx_gen_mesh call shards the
coordinate list across the GPUs and their PCIe links (host buffers, pinned).  End-to-end chunks/s.

  python tools/bench_multidev_host.py [--rings 4] [--steps 10]
"""
import argparse
import json
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--rings", type=int, default=4, help="region = 32 x 32*rings chunks per step")
    ap.add_argument("--steps", type=int, default=10)
    args = ap.parse_args()
    import voxb200
    ndev = voxb200._lib.lib().vx_device_count()
    nz = 32 * args.rings

    def coords(step):
        xs = np.arange(5 * step - 16, 5 * step + 16, dtype=np.int32)
        zs = np.arange(3 * step - 16, 3 * step - 16 + nz, dtype=np.int32)
        z, x = np.meshgrid(zs, xs, indexing="ij")
        return np.stack([x.reshape(-1), np.zeros(x.size, np.int32), z.reshape(-1)], 1).astype(np.int32)
    n = 32 * nz
    pinned = voxb200.PinnedBuffer(n * 8192 * 68)
    out = pinned.view(np.uint8)
    ref_bytes = None
    for k in [d for d in (1, 2, 4, 8) if d <= ndev]:
        ctx = voxb200.Context(n_devices=k)
        r = ctx.gen_mesh(coords(0), out=out)
        digest = hash(r.instances.tobytes()[:1 << 20]) ^ r.total
        if ref_bytes is None:
            ref_bytes = digest
        assert digest == ref_bytes, "multi-device result differs from single device"
        for s in range(1, 3):
            ctx.gen_mesh(coords(s), out=out)
        t0 = time.perf_counter()
        quads = 0
        for s in range(3, 3 + args.steps):
            quads += ctx.gen_mesh(coords(s), out=out).total
        dt = time.perf_counter() - t0
        print(json.dumps({"config": "one process, one context over %d GPU(s), vx_gen_mesh with host buffers" % k,
                          "chunks_per_step": n, "chunks_per_s": n * args.steps / dt, "quads_per_s": quads / dt,
                          "GBps_to_host": quads * 68 / dt / 1e9}))
        ctx.close()


if __name__ == "__main__":
    main()
