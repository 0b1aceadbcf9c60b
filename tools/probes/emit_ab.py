#!/usr/bin/env python
"""A/B aid for the emit kernel: one 8192-chunk batch of the region is generated once, then the emit alone is re-run
(vx_emit_again_device) 16 times back to back and timed with CUDA events.  argv[1] = a variant build of the library."""
import json
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch

import bench
import voxb200

if len(sys.argv) > 1:
    voxb200._lib.SO_PATH = os.path.abspath(sys.argv[1])
dev = torch.device("cuda", 0)
ctx = voxb200.Context(devices=[0])
c, m = bench.region_rows(-8, 0, 1024)
n = len(c)
d_c, d_m = torch.from_numpy(c).to(dev), torch.from_numpy(m).to(dev)
cap = n * 8192
d_counts = torch.zeros(n, dtype=torch.int32, device=dev)
d_off = torch.zeros(n + 1, dtype=torch.int64, device=dev)
d_inst = torch.empty(cap * 17, dtype=torch.int32, device=dev)
d_st = torch.zeros(4, dtype=torch.int32, device=dev)
st = torch.cuda.Stream(device=dev)
torch.cuda.set_stream(st)
ctx.gen_mesh_device(d_c.data_ptr(), n, d_counts.data_ptr(), d_off.data_ptr(), d_inst.data_ptr(), cap,
                    nbr_mask_ptr=d_m.data_ptr(), status_ptr=d_st.data_ptr(), stream=st.cuda_stream)
torch.cuda.synchronize()
ref = d_inst[: int(d_off[n].item()) * 17].clone()
best = 1e9
for rep in range(3):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(st)
    for _ in range(16):
        ctx.emit_again_device(d_c.data_ptr(), n, d_counts.data_ptr(), d_off.data_ptr(), d_inst.data_ptr(), cap,
                              status_ptr=d_st.data_ptr(), stream=st.cuda_stream)
    e1.record(st)
    torch.cuda.synchronize()
    best = min(best, e0.elapsed_time(e1) / 16)
q = int(d_off[n].item())
print(json.dumps({"variant": os.path.basename(sys.argv[1]) if len(sys.argv) > 1 else "default", "chunks": n, "quads": q,
                  "emit_ms": best, "GBps": (n * 36876 + 68 * q) / (best * 1e-3) / 1e9,
                  "same_bytes": bool(torch.equal(ref, d_inst[: q * 17]))}))
