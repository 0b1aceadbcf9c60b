#!/usr/bin/env python
"""Times the noise fill alone (vx_fill_device, occupancy out) on one 8192-chunk batch of the region and
checks the occupancy against the exact kernel.  VX_FILL_VARIANT picks the fast kernel (kernels_fill.cu)."""
import json
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch

import bench
import voxb200

if len(sys.argv) > 1:                       # a variant build of the library (tools/probes/build_fill_variants.sh)
    voxb200._lib.SO_PATH = os.path.abspath(sys.argv[1])

dev = torch.device("cuda", 0)
ctx = voxb200.Context(devices=[0])
c, _ = bench.region_rows(-8, 0, 1024)
n = len(c)
d_c = torch.from_numpy(c).to(dev)
occ = torch.zeros(n * 1024, dtype=torch.int32, device=dev)
st = torch.cuda.Stream(device=dev)
torch.cuda.set_stream(st)
for _ in range(3):
    ctx.fill_device(d_c.data_ptr(), n, occ_ptr=occ.data_ptr(), stream=st.cuda_stream)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
reps = 20
e0.record(st)
for _ in range(reps):
    ctx.fill_device(d_c.data_ptr(), n, occ_ptr=occ.data_ptr(), stream=st.cuda_stream)
e1.record(st)
torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / reps
# exact kernel (noise output forces it) on the first 512 chunks
m = 512
noise = torch.empty(m * 32768, dtype=torch.float32, device=dev)
occ2 = torch.zeros(m * 1024, dtype=torch.int32, device=dev)
ctx.fill_device(d_c.data_ptr(), m, occ_ptr=occ2.data_ptr(), noise_ptr=noise.data_ptr(), stream=st.cuda_stream)
torch.cuda.synchronize()
same = bool(torch.equal(occ[: m * 1024], occ2))
print(json.dumps({"variant": os.path.basename(sys.argv[1]) if len(sys.argv) > 1 else os.environ.get("VX_FILL_VARIANT", "default"), "chunks": n, "ms": ms,
                  "ms_per_1024": ms * 1024 / n, "tflops_equiv": 234 * n * 32768 / (ms * 1e-3) / 1e12,
                  "occupancy_equals_exact_kernel": same}))
