"""CPU, world_size 2 over gloo: the chunk-coordinate sharding bench.py uses for N > 1.
There is no collective on the data path; the process group only carries the barrier and the
max-over-ranks timing, so this test covers exactly that plus the shard/halo bookkeeping."""
import os
import socket
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
RINGT = 4          # small ring so the oracle finishes in seconds


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _neighbours(c):
    cx, cy, cz = (int(v) for v in c)
    return [(cx - 1, cy, cz), (cx + 1, cy, cz), (cx, cy, cz - 1), (cx, cy, cz + 1)]


def _worker(rank, world, port, q):
    sys.path.insert(0, ROOT)
    import torch
    import torch.distributed as dist

    import bench
    from oracle import oracle
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        step = 3
        coords = bench.ring_coords(step, rank, ring=RINGT)
        masks = bench.region_masks(coords, step, world, ring=RINGT)
        # this rank's shard, meshed the way the GPU path does it: own chunks from the noise, every
        # neighbour the mask declares present is regenerated (halo), everything else is None
        counts = []
        for c, m in zip(coords, masks):
            cubes = oracle.generate_terrain(c)
            adj = [oracle.generate_terrain(nc) if (m >> k) & 1 else None for k, nc in enumerate(_neighbours(c))]
            counts.append(len(oracle.cull_unseen_triangles(c, cubes, adj)))
        # gather shards on every rank (test plumbing only)
        all_coords = [torch.zeros(coords.shape, dtype=torch.int32) for _ in range(world)]
        all_counts = [torch.zeros(len(counts), dtype=torch.int64) for _ in range(world)]
        dist.all_gather(all_coords, torch.from_numpy(coords))
        dist.all_gather(all_counts, torch.tensor(counts))
        dist.barrier()
        tmax = bench.reduce_over_ranks(float(rank + 1), "max", dist)
        tsum = bench.reduce_over_ranks(float(rank + 1), "sum", dist)
        if rank == 0:
            q.put(dict(coords=torch.cat(all_coords).numpy(), counts=torch.cat(all_counts).numpy(),
                       tmax=tmax, tsum=tsum, masks0=masks))
    finally:
        dist.destroy_process_group()


def test_sharded_region_equals_whole_region():
    import torch.multiprocessing as mp
    world = 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = q.get(timeout=120)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    sys.path.insert(0, ROOT)
    from oracle import oracle
    coords = res["coords"]
    # shards tile the ring x ring*world region exactly once
    assert len({tuple(c) for c in coords.tolist()}) == RINGT * RINGT * world == len(coords)
    zs = sorted(set(coords[:, 2].tolist()))
    assert zs == list(range(zs[0], zs[0] + RINGT * world))
    # whole-region result on one "device": neighbours present iff in the list
    whole = oracle.gen_mesh_batch(coords, want_instances=False)
    assert np.array_equal(whole["counts"].astype(np.int64), res["counts"])
    assert res["tmax"] == 2.0 and res["tsum"] == 3.0          # timing = max over ranks
    # rank 0's +z border row sees rank 1 (bit 3), its -z border row sees nothing (bit 2 clear)
    m0 = res["masks0"].reshape(RINGT, RINGT)                 # [z][x]
    assert np.all(m0[-1] & 8) and not np.any(m0[0] & 4)


def test_ring_and_masks_single_rank():
    sys.path.insert(0, ROOT)
    import bench
    c = bench.ring_coords(7)
    assert c.shape == (1024, 3) and c.dtype == np.int32 and np.all(c[:, 1] == 0)
    assert len({tuple(x) for x in c.tolist()}) == 1024
    m = bench.region_masks(c, 7, 1)
    inside = {tuple(x) for x in c.tolist()}
    for ci, mi in list(zip(c.tolist(), m.tolist()))[::37]:
        want = sum(1 << k for k, nc in enumerate(_neighbours(ci)) if nc in inside)
        assert mi == want
    # the player moves every step: no two steps share their coordinate set
    assert not np.array_equal(bench.ring_coords(1), bench.ring_coords(2))


# ---- bench.py's configs[2] form: z bands of a square region, strong scaling ----------------------------------
REGT = 8           # 8 x 8 chunks, 2 bands of 4 rows


def _band_worker(rank, world, port, q):
    sys.path.insert(0, ROOT)
    import torch
    import torch.distributed as dist

    import bench
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        z_lo, z_hi = bench.band_rows(rank, world, REGT)
        coords, masks = bench.region_rows(z_lo, z_hi, REGT)
        # this rank's band the way the GPU path meshes it: present neighbours outside the band are halo
        counts, _, _ = bench.oracle_with_halo(coords, masks)
        parts = [torch.zeros(len(counts), dtype=torch.int32) for _ in range(world)]
        dist.all_gather(parts, torch.from_numpy(counts.astype(np.int32)))
        dist.barrier()
        if rank == 0:
            q.put(dict(counts=torch.cat(parts).numpy().astype(np.uint32), digest=bench.counts_digest(counts)))
    finally:
        dist.destroy_process_group()


def test_z_band_sharding_equals_whole_region():
    import torch.multiprocessing as mp
    world = 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_band_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = q.get(timeout=120)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    sys.path.insert(0, ROOT)
    import bench
    from oracle import oracle
    coords, masks = bench.region_rows(-REGT // 2, REGT // 2, REGT)
    whole = oracle.gen_mesh_batch(coords, want_instances=False)          # in-list neighbours == in-region neighbours
    assert np.array_equal(whole["counts"], res["counts"])                 # the concatenation of the bands, in order
    assert bench.counts_digest(res["counts"]) == bench.counts_digest(whole["counts"])
    assert res["digest"]["chunks"] == REGT * REGT // 2


def test_region_geometry():
    sys.path.insert(0, ROOT)
    import bench
    for world in (1, 2, 4, 8):
        rows = [bench.band_rows(r, world, 1024) for r in range(world)]
        assert rows[0][0] == -512 and rows[-1][1] == 512
        assert all(a[1] == b[0] for a, b in zip(rows, rows[1:]))
        assert len({hi - lo for lo, hi in rows}) == 1 and (rows[0][1] - rows[0][0]) % 8 == 0
    c, m = bench.region_rows(-4, -2, 8)
    assert c.shape == (16, 3) and c[0].tolist() == [-4, 0, -4] and c[-1].tolist() == [3, 0, -3]
    inside = {(x, 0, z) for x in range(-4, 4) for z in range(-4, 4)}
    for ci, mi in zip(c.tolist(), m.tolist()):
        assert mi == sum(1 << k for k, nc in enumerate(_neighbours(ci)) if nc in inside)
    idx = bench.sample_indices(1024, 8)
    pts = {(int(i % 1024) - 512, int(i // 1024) - 512) for i in idx}
    assert {(-512, -512), (511, 511), (-512, 511), (511, -512)} <= pts
    for r in range(1, 8):
        zb = bench.band_rows(r, 8, 1024)[0]
        assert any(z == zb for _, z in pts) and any(z == zb - 1 for _, z in pts)
    assert len(idx) >= 64 and len(set(idx.tolist())) == len(idx)
