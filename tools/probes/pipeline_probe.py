"""Probe: does alternating steps between k contexts / streams on ONE GPU overlap the HBM-bound emit of
step i with the issue-bound fill of step i+1?  Prints ms per step for k = 1, 2, 3."""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch
import voxb200
import bench

dev = torch.device("cuda", 0)
n = 1024
cap = n * 8192
K, W = 24, 6
coords = [torch.from_numpy(bench.ring_coords(s)).to(dev) for s in range(K + W)]
for k in (1, 2, 3):
    ctxs = [voxb200.Context(devices=[0]) for _ in range(k)]
    streams = [torch.cuda.Stream(device=dev) for _ in range(k)]
    bufs = [dict(c=torch.zeros(n, dtype=torch.int32, device=dev), o=torch.zeros(n + 1, dtype=torch.int64, device=dev),
                 i=torch.empty(cap * 17, dtype=torch.int32, device=dev), s=torch.zeros(4, dtype=torch.int32, device=dev))
            for _ in range(k)]

    def step(i):
        j = i % k
        b = bufs[j]
        ctxs[j].gen_mesh_device(coords[i].data_ptr(), n, b["c"].data_ptr(), b["o"].data_ptr(), b["i"].data_ptr(), cap,
                                status_ptr=b["s"].data_ptr(), stream=streams[j].cuda_stream)
    for i in range(W):
        step(i)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ends = [torch.cuda.Event() for _ in range(k)]
    with torch.cuda.stream(streams[0]):
        e0.record()
    for s in streams[1:]:
        s.wait_event(e0)
    for i in range(W, W + K):
        step(i)
    for j in range(k):
        ends[j].record(streams[j])
    for j in range(1, k):
        streams[0].wait_event(ends[j])
    with torch.cuda.stream(streams[0]):
        e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / K
    print("contexts %d: %.4f ms/step  %.3f M chunks/s  overflow %s" % (k, ms, n / ms / 1e3, [int(b["s"][0].item()) for b in bufs]))
    del ctxs, bufs
