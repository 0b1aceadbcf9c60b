#!/usr/bin/env python
"""BASELINE configs[2]: one R x R-chunk region generated + meshed, sharded by chunk coordinate
(contiguous z bands) over the ranks of one box; device resident, 8192-chunk batches through a reused
instance arena; no data-path collective (cross-band neighbours are halo slabs recomputed locally).

  python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 \
         --master-port 29533 tools/bench_region_multigpu.py [--region 1024]
"""
import argparse
import json
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--region", type=int, default=1024)
    args = ap.parse_args()
    real_stdout = os.dup(1)
    os.dup2(2, 1)
    import torch
    import torch.distributed as dist
    import voxb200
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    # batches alternate between two contexts / streams (own scratch and arenas): the emit of batch i
    # overlaps the fill of batch i + 1, as in bench.py
    PIPE = 2
    ctxs = [voxb200.Context(devices=[local]) for _ in range(PIPE)]
    tss = [torch.cuda.Stream(device=dev) for _ in range(PIPE)]
    ts = tss[0]
    torch.cuda.set_stream(ts)
    R = args.region
    band = R // world
    z_lo = -R // 2 + rank * band
    z_hi = z_lo + band if rank < world - 1 else R // 2
    xs = np.arange(-R // 2, R // 2, dtype=np.int32)
    rows = max(1, 8192 // R)
    nb = rows * R
    cap = nb * 8192
    d_counts = [torch.zeros(nb, dtype=torch.int32, device=dev) for _ in range(PIPE)]
    d_offsets = [torch.zeros(nb + 1, dtype=torch.int64, device=dev) for _ in range(PIPE)]
    d_inst = [torch.empty(cap * 17, dtype=torch.int32, device=dev) for _ in range(PIPE)]
    d_status = [torch.zeros(4, dtype=torch.int32, device=dev) for _ in range(PIPE)]
    batches = []
    for z0 in range(z_lo, z_hi, rows):
        zs = np.arange(z0, min(z0 + rows, z_hi), dtype=np.int32)
        z, x = np.meshgrid(zs, xs, indexing="ij")
        c = np.stack([x.reshape(-1), np.zeros(x.size, np.int32), z.reshape(-1)], 1).astype(np.int32)
        m = ((c[:, 0] > -R // 2) * 1 + (c[:, 0] < R // 2 - 1) * 2 + (c[:, 2] > -R // 2) * 4 +
             (c[:, 2] < R // 2 - 1) * 8).astype(np.uint8)          # present iff inside the WHOLE region
        batches.append((torch.from_numpy(c).to(dev), torch.from_numpy(m).to(dev), len(c)))
    totals = torch.zeros(len(batches), dtype=torch.int64, device=dev)

    def run():
        for i, (dc, dm, n) in enumerate(batches):
            j = i % PIPE
            ctxs[j].gen_mesh_device(dc.data_ptr(), n, d_counts[j].data_ptr(), d_offsets[j].data_ptr(),
                                    d_inst[j].data_ptr(), cap, nbr_mask_ptr=dm.data_ptr(),
                                    status_ptr=d_status[j].data_ptr(), stream=tss[j].cuda_stream)
            with torch.cuda.stream(tss[j]):
                totals[i:i + 1].copy_(d_offsets[j][n:n + 1], non_blocking=True)

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize()
    run()
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(tss[0])
    for t_ in tss[1:]:
        t_.wait_event(e0)
    run()
    for t_ in tss[1:]:
        done = torch.cuda.Event()
        done.record(t_)
        tss[0].wait_event(done)
    e1.record(tss[0])
    barrier()
    t = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device=dev)
    q = totals.sum().to(torch.float64).reshape(1)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dist.all_reduce(q, op=dist.ReduceOp.SUM)
    if rank == 0:
        ms = float(t.item())
        line = {"config": "3: %dx%d-chunk region, strong scaling, z bands, device resident" % (R, R), "n_gpus": world,
                "chunks": R * R, "quads": int(q.item()), "ms_max_over_ranks": ms, "chunks_per_s": R * R / (ms * 1e-3),
                "quads_per_s": float(q.item()) / (ms * 1e-3), "overflow": int(sum(int(st[0].item()) for st in d_status)), "contexts_per_gpu": PIPE}
        os.write(real_stdout, (json.dumps(line) + "\n").encode())
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
