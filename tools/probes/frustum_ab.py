#!/usr/bin/env python
"""A/B aid for the frustum kernels: time config 5 (10^6 AABBs, F frames per launch, bit output) on a given build of
the library -- python tools/probes/frustum_ab.py path/to/libvoxb200.so -- as CUDA-graph replays (device time only)."""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch  # noqa: E402
import voxb200  # noqa: E402

if len(sys.argv) > 1:
    voxb200._lib.SO_PATH = os.path.abspath(sys.argv[1])
from tools.bench_configs import camera_planes  # noqa: E402

dev = torch.device("cuda", 0)
ctx = voxb200.Context(devices=[0])
g = np.arange(-500, 500, dtype=np.int32)
z, x = np.meshgrid(g, g, indexing="ij")
c = np.stack([x.reshape(-1), np.zeros(x.size, np.int32), z.reshape(-1)], 1).astype(np.int32)
pl = camera_planes(1000, voxb200)
d_c, d_p = torch.from_numpy(c).to(dev), torch.from_numpy(pl).to(dev)
n = len(c)
out = []
for F in (1, 8, 64, 1000):
    launches = 1 if F == 1000 else min(1000 // F, 64)
    d_bits = torch.zeros(F * ((n + 31) // 32), dtype=torch.int32, device=dev)
    ts = torch.cuda.Stream(device=dev)
    gr = torch.cuda.CUDAGraph()
    with torch.cuda.graph(gr, stream=ts):
        for j in range(launches):
            ctx.frustum_cull_device(d_p.data_ptr() + j * F * 96, F, d_c.data_ptr(), n, bits_ptr=d_bits.data_ptr(),
                                    stream=ts.cuda_stream)
    gr.replay()
    torch.cuda.synchronize()
    best = 1e9
    for rep in range(5):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(5):
            gr.replay()
        e1.record()
        torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1) / 5 / launches)
    out.append("F=%d: %.2f us, %.0f G tests/s" % (F, best * 1e3, n * F / best / 1e6))
print(os.path.basename(sys.argv[1]) if len(sys.argv) > 1 else "default", "|", " | ".join(out))
