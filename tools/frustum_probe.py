#!/usr/bin/env python
"""One large frustum launch (10^6 AABBs x F frames) for ncu / timing."""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402
import voxb200  # noqa: E402
from tools.bench_configs import camera_planes  # noqa: E402

F = int(sys.argv[1]) if len(sys.argv) > 1 else 256
dev = torch.device("cuda", 0)
ctx = voxb200.Context(devices=[0])
g = np.arange(-500, 500, dtype=np.int32)
z, x = np.meshgrid(g, g, indexing="ij")
c = np.stack([x.reshape(-1), np.zeros(x.size, np.int32), z.reshape(-1)], 1).astype(np.int32)
pl = camera_planes(1000, voxb200)[:F]
d_c, d_p = torch.from_numpy(c).to(dev), torch.from_numpy(pl).to(dev)
d_bits = torch.zeros(F * ((len(c) + 31) // 32), dtype=torch.int32, device=dev)
ts = torch.cuda.Stream(device=dev)
for rep in range(3):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    with torch.cuda.stream(ts):
        e0.record()
        ctx.frustum_cull_device(d_p.data_ptr(), F, d_c.data_ptr(), len(c), bits_ptr=d_bits.data_ptr(),
                                stream=ts.cuda_stream)
        e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1)
print("F", F, "ms", ms, "Gtests/s", len(c) * F / ms / 1e6)
