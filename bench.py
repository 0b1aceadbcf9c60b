#!/usr/bin/env python
"""bench.py -- chunks/sec gen+mesh on B200 (BASELINE.json metric), one JSON line.

  python bench.py --gpus N --steps K --warmup W            # this repo's CUDA path
  python bench.py --impl reference --gpus N --steps K ...  # the reference's CPU work function (oracle port)

Workload = BASELINE configs[2]: the 1024 x 1024-chunk region (cx, cz in [-512, 511], cy = 0, seed 1,
default noise, a neighbour is present iff it lies inside the region), generated and meshed in 8192-chunk
batches (8 z rows) through reused output arenas.  A step = ONE PASS OVER THE WHOLE REGION (1 048 576
chunks, 128 batches).  With N GPUs the region is sharded by chunk coordinate into N contiguous z bands
(STRONG scaling: the total work is fixed), one process per GPU, no data-path collective: neighbours in
another band are halo slabs recomputed from the noise (chunk_builder.rs:181-241 keeps every chunk on the
y = 0 plane, so a band is a strip of the (x, z) grid).

  value     chunks/s = region chunks x K / device time (CUDA events, max over ranks); coordinates resident
            in HBM, instances written to HBM (vx_gen_mesh_device), two batches in flight per GPU
  e2e       the same region through the host-buffer C ABI call (vx_gen_mesh): H2D of coordinates + masks,
            D2H of counts / offsets / 68-byte instances into pinned memory inside the timed region
  roofline  the dominant kernel (noise fill): algorithmic FP32 ops / CUDA-event duration per launch
            against the FP32 peak measured on this GPU; roofline_mesh: the emit kernel against HBM
  parity    every rank's per-chunk quad counts are digested (CRC-32); rank 0 checks the concatenation
            against its own single-GPU pass over the whole region (N > 1) and against the CPU oracle on
            sampled chunks incl. the region corners and every band border
"""
import argparse
import json
import os
import socket
import statistics
import subprocess
import sys
import threading
import time
import zlib

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

REGION = 1024                   # configs[2]: 1024 x 1024 chunks
BATCH_CHUNKS = 8192             # chunks per vx_gen_mesh_device call
RING = 32                       # configs[1]: 32 x 32 chunks (secondary figure)
VOXELS = 32768
# Algorithmic work per voxel, counted from the oracle source (oracle/vox_oracle.c):
# 73 f32 add/sub/mul per simplex_3d, 3 octaves, 3 freq products, 6 per extra octave.
FLOP_PER_VOXEL = 3 * 73 + 3 + 2 * 6            # = 234
# mesher bytes per chunk = 36 864 in + 68 per quad out + 12 bookkeeping
MESH_BYTES_FIXED = 36864 + 12
INSTANCE_BYTES = 68
METRIC = "chunks/sec gen+mesh (device-timed)"
REF_ROWS_PER_STEP = 4           # reference arm: 4 z rows (4096 chunks) of the region per step


def workload_config(region: int) -> dict:
    """`config` of the JSON line -- identical in both arms."""
    return {
        "workload": "region%dx%d (configs[2]): %d chunks, cx,cz in [%d,%d], cy=0, seed 1, default noise (freq 0.09, "
                    "gain 2.5, lacunarity 0.5, 3 octaves), a neighbour is present iff it lies inside the region"
                    % (region, region, region * region, -region // 2, region // 2 - 1),
        "region": region,
        "step": "one pass over the whole region (the reference arm times a bounded slab of it per step)",
        "l2": "every batch writes ~2.5 GB of instances (> 126 MB L2) and reads 98 KB of coordinates that differ "
              "from batch to batch; the 32 MB occupancy intermediate is produced and consumed inside the batch",
    }


# ---- workload geometry (pure numpy: also used by the CPU tests) ----------------------------------------------
def band_rows(rank: int, world: int, region: int = REGION):
    """z rows [z_lo, z_hi) of rank's band: contiguous, equal, tiling [-region/2, region/2)."""
    per = region // world
    z_lo = -region // 2 + rank * per
    z_hi = z_lo + per if rank < world - 1 else region // 2
    return z_lo, z_hi


def region_rows(z_lo: int, z_hi: int, region: int = REGION):
    """Coordinates (z-major, x fastest) and neighbour masks of the z rows [z_lo, z_hi) of the region.
    Mask bit k (-x,+x,-z,+z, chunk_builder.rs:134-169) is set iff that neighbour is inside the WHOLE region."""
    h = region // 2
    xs = np.arange(-h, h, dtype=np.int32)
    zs = np.arange(z_lo, z_hi, dtype=np.int32)
    z, x = np.meshgrid(zs, xs, indexing="ij")
    c = np.stack([x.reshape(-1), np.zeros(x.size, np.int32), z.reshape(-1)], axis=1).astype(np.int32)
    m = ((c[:, 0] > -h) * 1 + (c[:, 0] < h - 1) * 2 + (c[:, 2] > -h) * 4 + (c[:, 2] < h - 1) * 8).astype(np.uint8)
    return c, m


def sample_indices(region: int, world: int, extra: int = 48, seed: int = 0) -> np.ndarray:
    """Region-linear indices (z-major) of the chunks the oracle re-checks: the four corners, both rows of
    every band border (at a few x), plus `extra` random ones."""
    h = region // 2
    pts = {(-h, -h), (h - 1, -h), (-h, h - 1), (h - 1, h - 1), (0, 0)}
    xs = [-h, -1, 0, h - 1, (7 * h) // 16]
    for r in range(1, world):
        zb = band_rows(r, world, region)[0]
        for x in xs:
            pts.add((x, zb - 1))
            pts.add((x, zb))
    rng = np.random.default_rng(seed)
    for x, z in rng.integers(-h, h, size=(extra, 2)):
        pts.add((int(x), int(z)))
    return np.array(sorted((z + h) * region + (x + h) for x, z in pts), np.int64)


def ring_coords(step: int, rank: int = 0, ring: int = RING) -> np.ndarray:
    """configs[1]: the ring x ring tile around the player's chunk; the player walks (+5, +3) chunks per step.
    Rank r owns the tile r*ring further along z."""
    cx, cz = 5 * step, 3 * step + ring * rank
    xs = np.arange(cx - ring // 2, cx + ring // 2, dtype=np.int32)
    zs = np.arange(cz - ring // 2, cz + ring // 2, dtype=np.int32)
    z, x = np.meshgrid(zs, xs, indexing="ij")
    return np.stack([x.reshape(-1), np.zeros(x.size, np.int32), z.reshape(-1)], axis=1).astype(np.int32)


def region_masks(coords: np.ndarray, step: int, world: int, ring: int = RING) -> np.ndarray:
    """Ring variant of the presence mask: neighbour k present iff inside the ring x ring*world region."""
    cx, cz = 5 * step, 3 * step
    x0, x1 = cx - ring // 2, cx + ring // 2 - 1
    z0, z1 = cz - ring // 2, cz - ring // 2 + ring * world - 1
    x, z = coords[:, 0], coords[:, 2]
    m = (x - 1 >= x0) * 1 + (x + 1 <= x1) * 2 + (z - 1 >= z0) * 4 + (z + 1 <= z1) * 8
    return m.astype(np.uint8)


def counts_digest(counts: np.ndarray) -> dict:
    c = np.ascontiguousarray(counts, dtype="<u4")
    return {"chunks": int(c.size), "quads": int(c.astype(np.uint64).sum()), "crc32_counts": "%08x" % zlib.crc32(c.tobytes())}


def reduce_over_ranks(x: float, op: str, dist=None, device=None) -> float:
    """max / sum of a scalar over all ranks (identity without a process group)."""
    if dist is None:
        return x
    import torch
    t = torch.tensor([x], dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.MAX if op == "max" else dist.ReduceOp.SUM)
    return float(t.item())


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled while the GPU is under load."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu: int):
        self.rows, self.proc, self.gpu = [], None, gpu

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", "-i", str(self.gpu), "--query-gpu=" + self.Q, "--format=csv,noheader,nounits",
                 "-lms", "50"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except OSError:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append((time.time(), [c.strip() for c in line.split(",")]))

    def stop(self, t0: float, t1: float) -> dict:
        """Only the samples taken inside [t0, t1] (the device-timed region) count."""
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        rows = [r for t, r in self.rows if t0 <= t <= t1 and len(r) >= 7]
        sm, mx, reasons = [], None, set()
        for r in rows:
            try:
                sm.append(float(r[0]))
                mx = float(r[1])
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": mx, "reasons": sorted(reasons),
                "samples": len(sm), "window": "the K timed region passes only (50 ms period, %.2f s)" % (t1 - t0)}


# ---- the reference's CPU work function (oracle port) ----------------------------------------------------------
def cpu_slab(step: int, nthreads: int, region: int, rows: int = REF_ROWS_PER_STEP):
    """The reference's worker closure (chunk_builder.rs:80-84: Chunk::new, then create_mesh) over `rows` z rows
    of the region: simdnoise-style AVX2 lanes, reference-order mesher, per-quad sin/cos + 4x4 product.  The
    slab's -z / +z neighbour rows are part of the work list only as neighbours (the oracle meshes what it is
    given; their cost is counted in the time but not in the chunk count -- conservative for the CPU)."""
    from oracle import oracle
    n_slabs = region // rows
    z0 = -region // 2 + (step % n_slabs) * rows
    c, m = region_rows(z0, z0 + rows, region)
    t0 = time.perf_counter()
    res = oracle.gen_mesh_batch(c, nbr_mask=m, nthreads=nthreads, use_simd=True, want_instances=False)
    return time.perf_counter() - t0, len(c), res["total"]


def run_reference(args, rank: int, world: int):
    if rank != 0:
        return
    cores = os.cpu_count() or 1
    region = args.region
    for s in range(args.warmup):
        cpu_slab(s, cores, region)
    dt = chunks = quads = 0
    for s in range(args.warmup, args.warmup + args.steps):
        d, c, q = cpu_slab(s, cores, region)
        dt += d
        chunks += c
        quads += q
    v = chunks / dt
    sample = ("%d slabs of %d z rows x %d chunks of the region (%d chunks/step), C oracle port of the reference work "
              "function with AVX2 noise lanes, one chunk per task on %d threads (the reference is Rust + un-vendored "
              "crates: not buildable here); in-slab and in-region neighbours present, neighbours outside the slab "
              "absent" % (args.steps, REF_ROWS_PER_STEP, region, chunks // max(args.steps, 1), cores))
    emit_json(({
        "impl": "reference", "metric": METRIC, "value": v, "unit": "chunks/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * dt / max(args.steps, 1),
        "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "quads_per_s": quads / dt,
        "config": workload_config(region),
        "detail": {"chunks_per_step": chunks // max(args.steps, 1),
                   "note": "a step of this arm is a bounded slab of the region, not the whole region; at N > 1 rank 0 "
                           "alone runs it on the host cores, so the rate does not depend on N"},
        "cpu_baseline": {"value": v, "unit": "chunks/s", "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": v, "unit": "chunks/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }))


def bind_to_gpu_numa_node(gpu: int):
    """Pin this process to the CPU cores NVML reports as local to its GPU, so that the pinned host
    buffers (first touch) and the copy threads sit on the GPU's NUMA node.  Plumbing only."""
    try:
        import pynvml
        pynvml.nvmlInit()
        h = pynvml.nvmlDeviceGetHandleByIndex(gpu)
        words = pynvml.nvmlDeviceGetCpuAffinity(h, (os.cpu_count() + 63) // 64)
        cpus = [64 * i + b for i, w in enumerate(words) for b in range(64) if (w >> b) & 1]
        if cpus:
            os.sched_setaffinity(0, cpus)
            return len(cpus)
    except Exception:
        pass
    return None


def oracle_with_halo(coords: np.ndarray, masks: np.ndarray):
    """Oracle meshes of `coords` where every neighbour the mask declares present is generated from the noise
    (the oracle only sees neighbours that are in its list, so they are appended as mesh-less extras)."""
    from oracle import oracle
    have = {tuple(c) for c in coords.tolist()}
    extra = []
    for c, m in zip(coords.tolist(), masks.tolist()):
        for k, d in enumerate(((-1, 0), (1, 0), (0, -1), (0, 1))):
            nc = (c[0] + d[0], c[1], c[2] + d[1])
            if (m >> k) & 1 and nc not in have:
                have.add(nc)
                extra.append(nc)
    n = len(coords)
    lst = np.concatenate([coords, np.array(extra, np.int32).reshape(-1, 3)]) if extra else coords
    msk = np.concatenate([masks, np.zeros(len(extra), np.uint8)])
    res = oracle.gen_mesh_batch(lst, nbr_mask=msk, want_instances=True)
    end = int(res["offsets"][n])
    return res["counts"][:n], res["offsets"][:n + 1], res["instances"][:end]


def run_ours(args, rank: int, world: int, local_rank: int):
    import torch
    import voxb200

    R = args.region
    if R % world or (R // world) % (max(1, BATCH_CHUNKS // R)):
        raise SystemExit("region %d does not split into %d bands of whole batches" % (R, world))
    numa_cpus = bind_to_gpu_numa_node(local_rank) if world > 1 else None
    torch.cuda.set_device(local_rank)
    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    dev = torch.device("cuda", local_rank)
    K, W = args.steps, args.warmup
    rows_b = max(1, BATCH_CHUNKS // R)          # z rows per batch
    nb = rows_b * R                             # chunks per batch
    cap = nb * 8192                             # quads per output arena (terrain averages ~4.6 k per chunk)

    # Two contexts and two streams on the GPU, batches alternate between them: each context owns its scratch
    # and output arena, so the HBM-bound emit of batch i overlaps the issue-bound fill of batch i + 1
    # (what a host does with two batches in flight).
    PIPE = int(os.environ.get("VX_BENCH_PIPE", "2"))       # (experiments only; the reported configuration is 2)
    ctxs = [voxb200.Context(devices=[local_rank]) for _ in range(PIPE)]
    ctx = ctxs[0]
    d_counts_l = [torch.zeros(nb, dtype=torch.int32, device=dev) for _ in range(PIPE)]
    d_offsets_l = [torch.zeros(nb + 1, dtype=torch.int64, device=dev) for _ in range(PIPE)]
    d_inst_l = [torch.empty(cap * 17, dtype=torch.int32, device=dev) for _ in range(PIPE)]
    d_status_l = [torch.zeros(4, dtype=torch.int32, device=dev) for _ in range(PIPE)]
    # real (non-default) torch streams: handle 0 would mean "the context's own stream" to the C ABI and
    # torch.cuda.Event only sees torch's current stream
    tstreams = [torch.cuda.Stream(device=dev) for _ in range(PIPE)]
    torch.cuda.set_stream(tstreams[0])
    assert all(t.cuda_stream != 0 for t in tstreams)

    class Band:
        """All batches of the z rows [z_lo, z_hi): coordinates and masks resident in HBM."""

        def __init__(self, z_lo, z_hi):
            self.z_lo, self.z_hi = z_lo, z_hi
            self.h_coords, self.h_masks = region_rows(z_lo, z_hi, R)
            self.d_coords = torch.from_numpy(self.h_coords).to(dev)
            self.d_masks = torch.from_numpy(self.h_masks).to(dev)
            self.n = len(self.h_coords)
            self.n_batches = self.n // nb
            self.d_all_counts = torch.zeros(self.n, dtype=torch.int32, device=dev)
            self.d_totals = torch.zeros(self.n_batches, dtype=torch.int64, device=dev)

        def batch(self, b, record=False, lanes=PIPE):
            j = b % lanes
            lo = b * nb
            ctxs[j].gen_mesh_device(self.d_coords[lo:lo + nb].data_ptr(), nb, d_counts_l[j].data_ptr(),
                                    d_offsets_l[j].data_ptr(), d_inst_l[j].data_ptr(), cap,
                                    nbr_mask_ptr=self.d_masks[lo:lo + nb].data_ptr(),
                                    status_ptr=d_status_l[j].data_ptr(), stream=tstreams[j].cuda_stream)
            if record:
                with torch.cuda.stream(tstreams[j]):
                    self.d_all_counts[lo:lo + nb].copy_(d_counts_l[j], non_blocking=True)
                    self.d_totals[b:b + 1].copy_(d_offsets_l[j][nb:nb + 1], non_blocking=True)

        def region_pass(self, record=False, lanes=PIPE):
            for b in range(self.n_batches):
                self.batch(b, record, lanes)

    z_lo, z_hi = band_rows(rank, world, R)
    band = Band(z_lo, z_hi)

    def barrier():
        torch.cuda.synchronize()
        if dist is not None:
            dist.barrier()
            torch.cuda.synchronize()

    def max_over_ranks(x: float) -> float:
        return reduce_over_ranks(x, "max", dist, dev)

    def sum_over_ranks(x: float) -> float:
        return reduce_over_ranks(x, "sum", dist, dev)

    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()

    def timed(fn, lanes, reps, warm):
        """`warm` untimed calls, then exactly `reps` between barriers; device time from the first stream's start
        event to an end event that waits for every stream."""
        for _ in range(warm):
            fn()
        barrier()
        l0 = sum(c.launch_count for c in ctxs)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t0 = time.time()
        e0.record(tstreams[0])
        for t in tstreams[1:lanes]:
            t.wait_event(e0)
        for _ in range(reps):
            fn()
        for t in tstreams[1:lanes]:
            done = torch.cuda.Event()
            done.record(t)
            tstreams[0].wait_event(done)
        e1.record(tstreams[0])
        barrier()
        return max_over_ranks(e0.elapsed_time(e1)), sum(c.launch_count for c in ctxs) - l0, t0, time.time()

    # ---- device-timed: K passes over the region (this rank's band), two batches in flight -------------------
    ms_total, launches, t_load0, t_load1 = timed(lambda: band.region_pass(), PIPE, K, W)
    clocks = sampler.stop(t_load0, t_load1) if rank == 0 else None
    value = R * R * K / (ms_total * 1e-3)
    # the same passes back to back on ONE context / stream (no overlap between batches), for reference
    ms_one_stream, _, _, _ = timed(lambda: band.region_pass(lanes=1), 1, max(1, min(K, 3)), 1)
    value_one_stream = R * R * max(1, min(K, 3)) / (ms_one_stream * 1e-3)

    if args.device_only:
        if rank == 0:
            emit_json({"value": value, "ms_per_step": ms_total / K, "one_stream": value_one_stream, "n_gpus": world,
                       "env": {k: v for k, v in os.environ.items() if k.startswith("VX_")}})
        if dist is not None:
            dist.destroy_process_group()
        return

    # ---- untimed digest pass: per-chunk counts of the whole band, status words --------------------------------
    band.region_pass(record=True)
    torch.cuda.synchronize()
    overflow = int(sum(int(s[0].item()) for s in d_status_l))
    assert overflow == 0, "instance capacity overflow in bench"
    my_counts = band.d_all_counts.cpu().numpy().astype(np.uint32)
    my_digest = counts_digest(my_counts)
    quads_region = sum_over_ranks(float(my_digest["quads"]))
    all_counts = my_counts
    rank_digests = [my_digest]
    if dist is not None:
        parts = [torch.zeros(band.n, dtype=torch.int32, device=dev) for _ in range(world)]
        dist.all_gather(parts, band.d_all_counts)
        all_counts = torch.cat(parts).cpu().numpy().astype(np.uint32)
        rank_digests = [counts_digest(p.cpu().numpy().astype(np.uint32)) for p in parts]

    # ---- per-kernel CUDA-event durations (one stream, timing on): 4 batches ------------------------------------
    ctx.set_timing(True)
    for b in range(min(4, band.n_batches)):
        band.batch(b, lanes=1)
    torch.cuda.synchronize()
    kt = ctx.get_timing()
    ctx.set_timing(False)
    k_ms = {k: (ms / ln if ln else None) for k, (ms, ln) in kt.items() if ln}
    q_batch = float(band.d_totals[:min(4, band.n_batches)].double().mean().item())

    # ---- emit over >= 8 back-to-back launches on one stream, so L2 write-back residue is paid inside the window ----
    band.batch(0, lanes=1)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    emit_reps = 8
    e0.record(tstreams[0])
    for _ in range(emit_reps):
        ctx.emit_again_device(band.d_coords.data_ptr(), nb, d_counts_l[0].data_ptr(), d_offsets_l[0].data_ptr(),
                              d_inst_l[0].data_ptr(), cap, status_ptr=d_status_l[0].data_ptr(),
                              stream=tstreams[0].cuda_stream)
    e1.record(tstreams[0])
    torch.cuda.synchronize()
    emit_steady_ms = e0.elapsed_time(e1) / emit_reps

    # ---- end to end through the host-buffer ABI call ------------------------------------------------------------
    # Two host threads, one context and one pinned output buffer each, take alternate sub-batches (the reference
    # runs this call on pool threads, chunk_builder.rs:80-84): the fill of one call runs while the other call's
    # instances cross PCIe.  Every call is one complete vx_gen_mesh: coordinates + masks from host memory in,
    # counts / offsets / instances in host memory out.
    e2e_nb = nb // 2                                           # 4096 chunks per call: 1.3 GB of instances
    e2e_cap = e2e_nb * 6144
    pinned_l = [voxb200.PinnedBuffer(e2e_cap * INSTANCE_BYTES) for _ in range(PIPE)]
    outs = [p_.view(np.uint8) for p_ in pinned_l]
    n_calls = band.n // e2e_nb

    def e2e_pass(first, last, fn_name, views):
        totals, errs = [0] * PIPE, []

        def worker(j):
            try:
                fn = getattr(ctxs[j], fn_name)
                for i in range(first + j, last, PIPE):
                    lo = i * e2e_nb
                    totals[j] += fn(band.h_coords[lo:lo + e2e_nb], nbr_mask=band.h_masks[lo:lo + e2e_nb],
                                    out=views[j]).total
            except Exception as e:                       # pragma: no cover
                errs.append(repr(e))
        th = [threading.Thread(target=worker, args=(j,)) for j in range(PIPE)]
        for t in th:
            t.start()
        for t in th:
            t.join()
        if errs:
            raise RuntimeError("; ".join(errs))
        return sum(totals)

    e2e_pass(0, min(2, n_calls), "gen_mesh", outs)             # warm-up: arenas, pinned pages
    barrier()
    t0 = time.perf_counter()
    e2e_quads = e2e_pass(0, n_calls, "gen_mesh", outs)
    torch.cuda.synchronize()
    e2e_s = max_over_ranks(time.perf_counter() - t0)
    barrier()
    e2e_quads_all = sum_over_ranks(float(e2e_quads))
    e2e_value = R * R / e2e_s
    d2h_bytes = int(4 * R * R + 8 * (R * R + n_calls * world) + INSTANCE_BYTES * e2e_quads_all)
    h2d_bytes = 13 * R * R
    # raw pinned D2H ceiling of this box, all ranks copying at once: 8 x 256 MB cudaMemcpyAsync per rank
    link_bytes = 256 << 20
    src = d_inst_l[0][: link_bytes // 4]
    dst = torch.from_numpy(outs[0][:link_bytes]).view(torch.int32)
    dst.copy_(src, non_blocking=True)
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(tstreams[0])
    for _ in range(8):
        dst.copy_(src, non_blocking=True)
    e1.record(tstreams[0])
    barrier()
    link_ms = max_over_ranks(e0.elapsed_time(e1))
    link_peak_gbs = world * 8 * link_bytes / (link_ms * 1e-3) / 1e9

    # ---- extra: the same region with 4-byte packed quads on the wire (extension, not the headline) --------------
    e2e_packed = None
    if hasattr(ctx, "gen_mesh_packed"):
        pqs = [p_.view(np.uint32) for p_ in pinned_l]
        e2e_pass(0, min(2, n_calls), "gen_mesh_packed", pqs)
        barrier()
        t0 = time.perf_counter()
        pk_quads = e2e_pass(0, n_calls, "gen_mesh_packed", pqs)
        torch.cuda.synchronize()
        pk_s = max_over_ranks(time.perf_counter() - t0)
        barrier()
        e2e_packed = {"value": R * R / pk_s, "unit": "chunks/s",
                      "d2h_bytes_per_step": int(4 * R * R + 8 * (R * R + n_calls * world) + 4 * sum_over_ranks(float(pk_quads))),
                      "api": "vx_gen_mesh_packed: 4-byte quad records, expanded by the renderer's vertex shader; "
                             "NOT the reference's 68-byte wire format (SURVEY 8f rank 2 extension)"}

    # ---- extra: a bounded sample of the same calls into ordinary PAGEABLE memory (what a Rust Vec is): the library's
    # pinned bounce buffers + copy threads instead of cudaMemcpyAsync's own pageable path -------------------------
    pg_calls = min(16, n_calls)
    pg_out = np.empty(e2e_cap * INSTANCE_BYTES, np.uint8)
    pg_out[:] = 0                                                # touched once: a reused destination
    ctxs[0].gen_mesh(band.h_coords[:e2e_nb], nbr_mask=band.h_masks[:e2e_nb], out=pg_out)
    barrier()
    t0 = time.perf_counter()
    pg_quads = 0
    for i in range(pg_calls):
        lo = i * e2e_nb
        pg_quads += ctxs[0].gen_mesh(band.h_coords[lo:lo + e2e_nb], nbr_mask=band.h_masks[lo:lo + e2e_nb], out=pg_out).total
    pg_s = max_over_ranks(time.perf_counter() - t0)
    barrier()
    pg_quads_all = sum_over_ranks(float(pg_quads))
    e2e_pageable = {"value": world * pg_calls * e2e_nb / pg_s, "unit": "chunks/s",
                    "d2h_gbs": INSTANCE_BYTES * pg_quads_all / pg_s / 1e9,
                    "sample": "%d calls of %d chunks per rank into one reused pageable buffer, one host thread per rank "
                              "(not the headline: `e2e` uses pinned buffers)" % (pg_calls, e2e_nb)}
    del pg_out

    if rank != 0:
        if dist is not None:
            dist.barrier()                                   # rank 0's single-GPU verification pass
            dist.destroy_process_group()
        return

    # ---- parity (rank 0) ----------------------------------------------------------------------------------------
    parity = {"region_digest": counts_digest(all_counts), "per_rank": rank_digests,
              "quads_match_e2e_pass": int(quads_region) == int(e2e_quads_all)}
    if world > 1:
        # the N-rank concatenation against this GPU's own pass over the WHOLE region (N = 1 sharding)
        whole = Band(-R // 2, R // 2)
        whole.region_pass(record=True)
        torch.cuda.synchronize()
        single = whole.d_all_counts.cpu().numpy().astype(np.uint32)
        parity["single_gpu_digest"] = counts_digest(single)
        parity["sharded_equals_single_gpu"] = bool(np.array_equal(single, all_counts))
        assert parity["sharded_equals_single_gpu"], "N-rank region differs from the single-GPU region"
        del whole
    if dist is not None:
        dist.barrier()
    if not args.no_cpu_baseline:
        idx = sample_indices(R, world, extra=48)
        h = R // 2
        sc = np.stack([(idx % R) - h, np.zeros(len(idx), np.int64), (idx // R) - h], axis=1).astype(np.int32)
        sm = ((sc[:, 0] > -h) * 1 + (sc[:, 0] < h - 1) * 2 + (sc[:, 2] > -h) * 4 + (sc[:, 2] < h - 1) * 8).astype(np.uint8)
        o_counts, o_off, o_inst = oracle_with_halo(sc, sm)
        g = ctx.gen_mesh(sc, nbr_mask=sm)
        a, b_ = g.instances.copy(), o_inst.copy()
        a["t_matrix"] += np.float32(0)
        b_["t_matrix"] += np.float32(0)
        parity["oracle_sample"] = {
            "chunks": int(len(idx)), "includes": "4 region corners, both rows of every band border, 48 random",
            "pass_counts_equal_oracle": bool(np.array_equal(all_counts[idx], o_counts)),
            "instance_bytes_equal_oracle": bool(a.tobytes() == b_.tobytes()), "quads": int(o_off[-1])}
        assert parity["oracle_sample"]["pass_counts_equal_oracle"], "region pass differs from the oracle"
        assert parity["oracle_sample"]["instance_bytes_equal_oracle"], "instances differ from the oracle"

    # ---- rooflines ------------------------------------------------------------------------------------------------
    peaks_path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    hbm_peak, hbm_src = 6650.0, "fallback (B200_PROFILING.md)"
    if os.path.exists(peaks_path):
        hbm_peak, hbm_src = float(json.load(open(peaks_path))["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    ffma_peak = ctx.measure_peak(0)          # TFLOP/s, FFMA = 2 flop
    nofuse_peak = ctx.measure_peak(1)        # TFLOP/s, FMUL + FADD issued separately
    int_peak = ctx.measure_peak(2)           # Tera int-ops/s
    traffic = {}
    tpath = os.path.join(ROOT, "profiles", "ncu_traffic.json")
    if os.path.exists(tpath):
        traffic = json.load(open(tpath))
    fill_ms = k_ms.get("fill")
    fill_tflops = FLOP_PER_VOXEL * nb * VOXELS / (fill_ms * 1e-3) / 1e12 if fill_ms else None
    emit_ms = k_ms.get("emit")
    mesh_bytes = nb * MESH_BYTES_FIXED + INSTANCE_BYTES * q_batch
    mesh_gbs = mesh_bytes / (emit_ms * 1e-3) / 1e9 if emit_ms else None
    mesh_gbs_steady = mesh_bytes / (emit_steady_ms * 1e-3) / 1e9 if emit_steady_ms else None

    # ---- configs[1] beside it: the 32 x 32 ring around a moving player (1024 chunks per step) ---------------------
    ring = None
    try:
        n_r = RING * RING
        rsteps = 20
        rc = [torch.from_numpy(ring_coords(s)).to(dev) for s in range(rsteps + 3)]

        def ring_step(i, lanes):
            j = i % lanes
            ctxs[j].gen_mesh_device(rc[i].data_ptr(), n_r, d_counts_l[j].data_ptr(), d_offsets_l[j].data_ptr(),
                                    d_inst_l[j].data_ptr(), cap, status_ptr=d_status_l[j].data_ptr(),
                                    stream=tstreams[j].cuda_stream)
        ring = {}
        for lanes, key in ((PIPE, "two_in_flight"), (1, "one_stream")):
            for i in range(3):
                ring_step(i, lanes)
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(tstreams[0])
            for t in tstreams[1:lanes]:
                t.wait_event(e0)
            for i in range(3, 3 + rsteps):
                ring_step(i, lanes)
            for t in tstreams[1:lanes]:
                done = torch.cuda.Event()
                done.record(t)
                tstreams[0].wait_event(done)
            e1.record(tstreams[0])
            torch.cuda.synchronize()
            ms = e0.elapsed_time(e1) / rsteps
            ring[key] = {"ms_per_ring": ms, "chunks_per_s": n_r / (ms * 1e-3)}
        ring["note"] = "configs[1] on this GPU: 1024 chunks per vx_gen_mesh_device call, 20 calls"
    except Exception as e:                               # pragma: no cover
        ring = {"error": repr(e)}

    # ---- CPU baseline beside it (bounded sample, N = 1 only) ---------------------------------------------------
    cpu = None
    if world == 1 and not args.no_cpu_baseline:
        cores = os.cpu_count() or 1
        dt1, c1, _ = cpu_slab(0, cores, R)
        reps = int(min(24, max(1, round(5.0 / max(dt1, 1e-3)))))
        dt = chunks = 0
        for s in range(1, 1 + reps):
            d, c, _ = cpu_slab(s * 7, cores, R)
            dt += d
            chunks += c
        cpu = {"value": chunks / dt, "unit": "chunks/s", "cores": cores, "kind": "port",
               "sample": "%d slabs of %d z rows x %d chunks of the region (%.1f s wall), C oracle port of the reference "
                         "work function with AVX2 noise lanes, one chunk per task on %d threads"
                         % (reps, REF_ROWS_PER_STEP, R, dt, cores)}

    census = None
    try:
        census = ctx.fill(band.h_coords[:1024], blocks=False, census=True)["near_zero"]
    except Exception:
        pass

    cfg = workload_config(R)
    line = {
        "metric": METRIC, "value": value, "unit": "chunks/s", "n_gpus": world, "steps": K, "warmup": W,
        "ms_per_step": ms_total / K, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
        "dtype": "f32", "data": "synthetic",
        "quads_per_s": quads_region * K / (ms_total * 1e-3),
        "config": cfg,
        "detail": {
            "chunks_per_step": R * R, "quads_per_step": quads_region, "batch_chunks": nb,
            "batches_per_step_per_gpu": band.n_batches,
            "parallelism": "chunk-coordinate sharding into %d contiguous z bands, one process per GPU, no collective "
                           "on the data path (cross-band neighbours are halo slabs recomputed from the noise)" % world,
            "pipelining": "batches alternate between 2 contexts / CUDA streams per GPU (own scratch and output arena): "
                          "the HBM-bound emit of batch i overlaps the issue-bound fill of batch i+1; every batch is a "
                          "complete vx_gen_mesh_device call",
            "numa_bound_cpus_rank0": numa_cpus,
        },
        "one_stream": {"value": value_one_stream, "ms_per_step": ms_one_stream / max(1, min(K, 3)),
                       "note": "the same region passes with every batch back to back on ONE context / stream; the "
                               "per-kernel times and rooflines below are measured in this mode"},
        "ring32x32": ring,
        "e2e": {"value": e2e_value, "unit": "chunks/s", "h2d_bytes_per_step": h2d_bytes, "d2h_bytes_per_step": d2h_bytes,
                "steps": 1, "seconds": e2e_s, "d2h_gbs": d2h_bytes / e2e_s / 1e9, "link_peak_gbs": link_peak_gbs,
                "link_frac": (d2h_bytes / e2e_s / 1e9) / link_peak_gbs if link_peak_gbs else None,
                "link_peak_how": "8 x 256 MB cudaMemcpyAsync device -> pinned host per rank, all %d ranks at once, "
                                 "CUDA events, max over ranks" % world,
                "api": "vx_gen_mesh (host buffers, pinned instance buffers), %d chunks per call; two host threads with "
                       "one context each take alternate calls; one full pass over the region" % e2e_nb},
        "e2e_packed_extension": e2e_packed,
        "e2e_pageable_sample": e2e_pageable,
        "gpu_launches": launches,
        "clocks": clocks,
        "roofline": {
            "kernel": "vx_fill_lattice_kernel (noise fill: FP32 + INT32 instruction issue; not HBM- or tensor-bound)",
            "bound": "fp32", "achieved": fill_tflops, "peak": ffma_peak, "unit": "TFLOP/s",
            "frac": (fill_tflops / ffma_peak) if fill_tflops and ffma_peak else None,
            "peak_source": "FP32 FFMA peak measured live by vx_measure_peak (MEASURED_PEAKS.json has no FP32 figure)",
            "peak_nonfused": nofuse_peak,
            "frac_nonfused": (fill_tflops / nofuse_peak) if fill_tflops and nofuse_peak else None,
            "note": "achieved = the reference algorithm's flops (3 octaves x 78 per voxel) / fill time; the kernel skips "
                    "octaves that cannot flip the sign, so it executes fewer (throughput-equivalent, not pipe "
                    "utilisation); the cell selection may not contract a*b+c (bit-exact), the smooth part uses FMA "
                    "under a certified error margin",
            "flop_per_voxel": FLOP_PER_VOXEL, "voxels_per_launch": nb * VOXELS, "ms_per_launch": fill_ms,
            "traffic": traffic.get("fill"),
        },
        "roofline_mesh": {
            "kernel": "vx_mesh_emit_kernel", "bound": "hbm", "achieved": mesh_gbs, "peak": hbm_peak, "unit": "GB/s",
            "frac": (mesh_gbs / hbm_peak) if mesh_gbs else None, "peak_source": hbm_src,
            "bytes_per_launch": mesh_bytes, "ms_per_launch": emit_ms, "traffic": traffic.get("emit"),
            "steady_state": {"ms_per_launch": emit_steady_ms, "achieved": mesh_gbs_steady,
                             "frac": (mesh_gbs_steady / hbm_peak) if mesh_gbs_steady else None,
                             "note": "8 emit launches back to back on one stream: the write-back of L2-resident lines "
                                     "of launch i is paid inside the window of launch i+1"},
            "note": "a single launch's event window ends while part of its output is still dirty in L2; "
                    "`traffic` (ncu dram bytes) below the algorithmic bytes shows how much",
        },
        "kernel_ms": k_ms,
        "peaks": {"fp32_ffma_tflops": ffma_peak, "fp32_fmul_fadd_tflops": nofuse_peak, "int32_tops": int_peak,
                  "hbm_gbs": hbm_peak},
        "cpu_baseline": cpu,
        "parity": parity,
        "near_threshold_voxels": {"abs_noise_le_1e-6": census, "of": 1024 * VOXELS,
                                  "note": "first 1024 chunks of the band; counted with the exact noise kernel"},
    }
    emit_json(line)
    if dist is not None:
        dist.destroy_process_group()


_REAL_STDOUT = None


def emit_json(obj) -> None:
    """Exactly one JSON line on the real stdout (library chatter such as NCCL's version banner
    goes to stderr: fd 1 is pointed at fd 2 for the duration of the run)."""
    data = (json.dumps(obj) + "\n").encode()
    if _REAL_STDOUT is None:
        sys.stdout.write(data.decode())
        sys.stdout.flush()
    else:
        os.write(_REAL_STDOUT, data)


def redirect_stdout():
    """fd 1 -> fd 2 for library chatter; the JSON line goes to the saved real stdout (after a possible re-exec
    under torchrun, so that the ranks inherit the caller's stdout)."""
    global _REAL_STDOUT
    sys.stdout.flush()
    _REAL_STDOUT = os.dup(1)
    os.dup2(2, 1)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--region", type=int, default=REGION, help="edge of the square region in chunks (tests use less)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--device-only", action="store_true",
                    help="A/B aid: only the device-timed region passes, a short line (not the driver's contract)")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        redirect_stdout()
        run_reference(args, rank, world)
        return
    if args.gpus > 1 and world == 1:
        # launched without torchrun: re-exec under it (one rank per GPU, 127.0.0.1 rendezvous)
        s = socket.socket()
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
        s.close()
        os.execvp(sys.executable, [sys.executable, "-m", "torch.distributed.run", "--nnodes=1",
                                   "--nproc-per-node", str(args.gpus), "--master-addr", "127.0.0.1",
                                   "--master-port", str(port), os.path.abspath(__file__)] + sys.argv[1:])
    redirect_stdout()
    run_ours(args, rank, world, local_rank)


if __name__ == "__main__":
    main()
