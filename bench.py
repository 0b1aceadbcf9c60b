#!/usr/bin/env python
"""bench.py -- chunks/sec gen+mesh on B200 (BASELINE.json metric), one JSON line.

  python bench.py --gpus N --steps K --warmup W            # this repo's CUDA path
  python bench.py --impl reference --gpus N --steps K ...  # the reference's CPU work function (oracle port)

A step = one pass of the hot path (noise fill -> face culling -> InstanceRaw
emission) over the configs[1] workload: the 32x32 render-distance ring (1024
chunks) around a player that moves every step, seed 1, default noise,
neighbours present iff inside the region.  With N GPUs every rank owns one
such ring tile of a 32 x 32N region (weak scaling, sharded by chunk
coordinate, no data-path collective; cross-rank neighbours are halo slabs
recomputed from the noise).

  value     chunks/s, device-timed with CUDA events, coordinates already in HBM,
            instances written to HBM (vx_gen_mesh_device)
  e2e       chunks/s through the host-buffer C ABI call (vx_gen_mesh): H2D of the
            coordinates and D2H of counts/offsets/instances (pinned) inside the
            timed region
  roofline  the dominant kernel (noise fill): algorithmic FP32 ops / CUDA-event
            duration against the FP32 peak measured on this GPU
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

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

RING = 32                       # configs[1]: 32 x 32 chunks
CHUNKS_PER_STEP = RING * RING
VOXELS = 32768
# Algorithmic work per voxel, counted from the oracle source (oracle/vox_oracle.c):
# 73 f32 add/sub/mul per simplex_3d, 3 octaves, 3 freq products, 6 per extra octave.
FLOP_PER_VOXEL = 3 * 73 + 3 + 2 * 6            # = 234
# mesher bytes per chunk = 36 864 in + 68 per quad out + 12 bookkeeping
MESH_BYTES_FIXED = 36864 + 12
INSTANCE_BYTES = 68
METRIC = "chunks/sec gen+mesh (device-timed)"


def ring_coords(step: int, rank: int = 0, ring: int = RING) -> np.ndarray:
    """The ring x ring tile around the player's chunk; the player walks (+5, +3) chunks per step.
    Rank r owns the tile r*ring further along z (chunk-coordinate sharding)."""
    cx, cz = 5 * step, 3 * step + ring * rank
    xs = np.arange(cx - ring // 2, cx + ring // 2, dtype=np.int32)
    zs = np.arange(cz - ring // 2, cz + ring // 2, dtype=np.int32)
    z, x = np.meshgrid(zs, xs, indexing="ij")
    return np.stack([x.reshape(-1), np.zeros(x.size, np.int32), z.reshape(-1)], axis=1).astype(np.int32)


def region_masks(coords: np.ndarray, step: int, world: int, ring: int = RING) -> np.ndarray:
    """Neighbour k (-x,+x,-z,+z) present iff inside the whole ring x ring*world region of this step;
    a present neighbour owned by another rank becomes a halo slab recomputed locally."""
    cx, cz = 5 * step, 3 * step
    x0, x1 = cx - ring // 2, cx + ring // 2 - 1
    z0, z1 = cz - ring // 2, cz - ring // 2 + ring * world - 1
    x, z = coords[:, 0], coords[:, 2]
    m = (x - 1 >= x0) * 1 + (x + 1 <= x1) * 2 + (z - 1 >= z0) * 4 + (z + 1 <= z1) * 8
    return m.astype(np.uint8)


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
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.06)
        self.proc.terminate()
        rows = [r for t, r in self.rows if t0 <= t <= t1 and len(r) >= 7] or [r for _, r in self.rows if len(r) >= 7]
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
                "samples": len(sm), "window": "warm-up + timed + e2e passes (50 ms period)"}


def cpu_work_function(steps, nthreads: int):
    """The reference's worker closure (chunk_builder.rs:80-84) over whole rings on the host cores:
    simdnoise-style AVX2 lanes, reference-order mesher, per-quad sin/cos + 4x4 product (oracle port)."""
    from oracle import oracle
    quads = 0
    t0 = time.perf_counter()
    for s in steps:
        res = oracle.gen_mesh_batch(ring_coords(s), nthreads=nthreads, use_simd=True, want_instances=False)
        quads += res["total"]
    return time.perf_counter() - t0, quads


def run_reference(args, rank: int):
    if rank != 0:
        return
    cores = os.cpu_count() or 1
    cpu_work_function(range(args.warmup), cores) if args.warmup else None
    dt, quads = cpu_work_function(range(args.warmup, args.warmup + args.steps), cores)
    v = CHUNKS_PER_STEP * args.steps / dt
    emit_json(({
        "impl": "reference", "metric": METRIC, "value": v, "unit": "chunks/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * dt / args.steps,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "quads_per_s": quads / dt,
        "config": {"workload": "ring32x32 (configs[1]): 1024 chunks around a moving player, seed 1, default "
                               "noise, neighbours present iff in the region",
                   "chunks_per_step": CHUNKS_PER_STEP},
        "cpu_baseline": {"value": v, "unit": "chunks/s", "cores": cores, "kind": "port",
                         "sample": "%d rings of 32x32 chunks, C oracle port with AVX2 lanes, %d threads "
                                   "(the reference is Rust + un-vendored crates: not buildable here)"
                                   % (args.steps, cores)},
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


def run_ours(args, rank: int, world: int, local_rank: int):
    import torch
    import voxb200

    numa_cpus = bind_to_gpu_numa_node(local_rank) if world > 1 else None
    torch.cuda.set_device(local_rank)
    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    dev = torch.device("cuda", local_rank)
    ctx = voxb200.Context(devices=[local_rank])
    K, W = args.steps, args.warmup
    n = CHUNKS_PER_STEP
    steps = list(range(W + K))
    h_coords = [ring_coords(s, rank) for s in steps]
    h_masks = [region_masks(c, s, world) for c, s in zip(h_coords, steps)] if world > 1 else None
    d_coords = [torch.from_numpy(c).to(dev) for c in h_coords]
    d_masks = [torch.from_numpy(m).to(dev) for m in h_masks] if h_masks else None
    cap = n * 8192
    # Two contexts and two streams on the GPU, steps alternate between them: each context owns its scratch
    # and output buffers, so the HBM-bound emit of step i overlaps the issue-bound fill of step i + 1
    # (what a host would do with two batches in flight).  The one-stream figure is
    # measured as well and reported beside it.
    PIPE = 2
    ctxs = [ctx] + [voxb200.Context(devices=[local_rank]) for _ in range(PIPE - 1)]
    d_counts_l = [torch.zeros(n, dtype=torch.int32, device=dev) for _ in range(PIPE)]
    d_offsets_l = [torch.zeros(n + 1, dtype=torch.int64, device=dev) for _ in range(PIPE)]
    d_inst_l = [torch.empty(cap * 17, dtype=torch.int32, device=dev) for _ in range(PIPE)]
    d_status_l = [torch.zeros(4, dtype=torch.int32, device=dev) for _ in range(PIPE)]
    d_counts, d_offsets, d_inst, d_status = d_counts_l[0], d_offsets_l[0], d_inst_l[0], d_status_l[0]
    # real (non-default) torch streams: handle 0 would mean "the context's own stream" to the
    # C ABI and torch.cuda.Event only sees torch's current stream
    tstreams = [torch.cuda.Stream(device=dev) for _ in range(PIPE)]
    tstream = tstreams[0]
    torch.cuda.set_stream(tstream)
    stream = tstream.cuda_stream
    assert stream != 0

    def step_device(i, lanes=1):
        j = i % lanes
        ctxs[j].gen_mesh_device(d_coords[i].data_ptr(), n, d_counts_l[j].data_ptr(), d_offsets_l[j].data_ptr(),
                                d_inst_l[j].data_ptr(), cap, nbr_mask_ptr=d_masks[i].data_ptr() if d_masks else None,
                                status_ptr=d_status_l[j].data_ptr(), stream=tstreams[j].cuda_stream)

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
    t_load0 = time.time()

    def timed_steps(lanes):
        """W warm-up steps, then exactly K steps between barriers; device time from the first stream's
        start event to an end event that waits for every stream."""
        for i in range(W):
            step_device(i, lanes)
        barrier()
        l0 = sum(c.launch_count for c in ctxs)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(tstreams[0])
        for t in tstreams[1:lanes]:
            t.wait_event(e0)
        for i in range(W, W + K):
            step_device(i, lanes)
        for t in tstreams[1:lanes]:
            done = torch.cuda.Event()
            done.record(t)
            tstreams[0].wait_event(done)
        e1.record(tstreams[0])
        barrier()
        return max_over_ranks(e0.elapsed_time(e1)), sum(c.launch_count for c in ctxs) - l0

    # ---- device-timed -------------------------------------------------------------------------
    ms_one_stream, _ = timed_steps(1)
    ms_total, launches = timed_steps(PIPE)
    value = world * n * K / (ms_total * 1e-3)
    value_one_stream = world * n * K / (ms_one_stream * 1e-3)

    # ---- untimed replay: quads per step, status, per-kernel CUDA-event durations -------------
    ctx.set_timing(True)
    quads = 0
    for i in range(W, W + K):
        step_device(i)
        torch.cuda.synchronize()
        assert int(d_status[0].item()) == 0, "instance capacity overflow in bench"
        quads += int(d_offsets[n].item())
    kt = ctx.get_timing()
    ctx.set_timing(False)
    quads_all = sum_over_ranks(float(quads))
    k_ms = {k: (ms / ln if ln else None) for k, (ms, ln) in kt.items() if ln}

    # ---- end to end through the host-buffer ABI call ------------------------------------------
    # Two host threads, one context and one pinned output buffer each, take alternate steps (the
    # reference runs this call on pool threads, chunk_builder.rs:80-84): the fill of one step runs
    # while the other step's instances cross PCIe, so the link never idles.  Every step is one
    # complete vx_gen_mesh call: coordinates from host memory in, instances in host memory out.
    import threading
    pinned_l = [voxb200.PinnedBuffer(cap * INSTANCE_BYTES) for _ in range(PIPE)]
    pinned = pinned_l[0]
    outs = [p_.view(np.uint8) for p_ in pinned_l]
    out = outs[0]

    def e2e_pass(first, last, fn_name, views):
        totals, errs = [0] * PIPE, []

        def worker(j):
            try:
                fn = getattr(ctxs[j], fn_name)
                for i in range(first + j, last, PIPE):
                    totals[j] += fn(h_coords[i], nbr_mask=h_masks[i] if h_masks else None, out=views[j]).total
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

    e2e_pass(0, min(W, 4), "gen_mesh", outs)
    barrier()
    t0 = time.perf_counter()
    e2e_quads = e2e_pass(W, W + K, "gen_mesh", outs)
    torch.cuda.synchronize()
    e2e_s = max_over_ranks(time.perf_counter() - t0)
    barrier()
    e2e_value = world * n * K / e2e_s
    # one thread, one context, strictly one call after the other (reported beside the headline)
    t0 = time.perf_counter()
    for i in range(W, W + K):
        ctx.gen_mesh(h_coords[i], nbr_mask=h_masks[i] if h_masks else None, out=out)
    torch.cuda.synchronize()
    e2e_serial_s = max_over_ranks(time.perf_counter() - t0)
    barrier()
    # ---- extra: the same call with 4-byte packed quads on the wire (extension, not the headline) ---
    e2e_packed = None
    if hasattr(ctx, "gen_mesh_packed"):
        pqs = [p_.view(np.uint32) for p_ in pinned_l]
        e2e_pass(0, min(W, 4), "gen_mesh_packed", pqs)
        barrier()
        t0 = time.perf_counter()
        pk_quads = e2e_pass(W, W + K, "gen_mesh_packed", pqs)
        torch.cuda.synchronize()
        pk_s = max_over_ranks(time.perf_counter() - t0)
        barrier()
        e2e_packed = {"value": world * n * K / pk_s, "unit": "chunks/s",
                      "d2h_bytes_per_step": int(4 * n + 8 * (n + 1) + 4 * pk_quads / K),
                      "api": "vx_gen_mesh_packed: 4-byte quad records, expanded by the renderer's vertex shader; "
                             "NOT the reference's 68-byte wire format (SURVEY 8f rank 2 extension)"}
    t_load1 = time.time()
    clocks = sampler.stop(t_load0, t_load1) if rank == 0 else None

    if rank != 0:
        if dist is not None:
            dist.destroy_process_group()
        return

    # ---- rooflines -----------------------------------------------------------------------------
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
    fill_tflops = FLOP_PER_VOXEL * n * VOXELS / (fill_ms * 1e-3) / 1e12 if fill_ms else None
    emit_ms = k_ms.get("emit")
    q_step = quads / K
    mesh_bytes = n * MESH_BYTES_FIXED + INSTANCE_BYTES * q_step
    mesh_gbs = mesh_bytes / (emit_ms * 1e-3) / 1e9 if emit_ms else None

    # ---- CPU baseline beside it (bounded sample, N = 1 only) ---------------------------------
    cpu = None
    if world == 1 and not args.no_cpu_baseline:
        cores = os.cpu_count() or 1
        dt1, _ = cpu_work_function([0], cores)
        reps = int(min(16, max(1, round(4.0 / max(dt1, 1e-3)))))
        dt, _ = cpu_work_function(range(1, 1 + reps), cores)
        cpu = {"value": n * reps / dt, "unit": "chunks/s", "cores": cores, "kind": "port",
               "sample": "%d rings of 32x32 chunks (%.1f s wall), C oracle port of the reference work function "
                         "with AVX2 noise lanes, one chunk per task on %d threads" % (reps, dt, cores)}

    # ---- live parity spot check: per-chunk quad counts of the first timed ring, GPU vs oracle ----
    parity = None
    if world == 1 and not args.no_cpu_baseline:
        from oracle import oracle
        g = ctx.gen_mesh(h_coords[W], instances=False)
        o = oracle.gen_mesh_batch(h_coords[W], want_instances=False)
        parity = {"ring_step": W, "gpu_quads": g.total, "oracle_quads": o["total"],
                  "per_chunk_counts_equal": bool(np.array_equal(g.counts, o["counts"]))}

    census = None
    try:
        census = ctx.fill(h_coords[W], blocks=False, census=True)["near_zero"]
    except Exception:
        pass

    line = {
        "metric": METRIC, "value": value, "unit": "chunks/s", "n_gpus": world, "steps": K, "warmup": W,
        "ms_per_step": ms_total / K, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f32", "data": "synthetic",
        "quads_per_s": quads_all / (ms_total * 1e-3),
        "one_stream": {"value": value_one_stream, "ms_per_step": ms_one_stream / K,
                       "note": "same K steps back to back on ONE context / stream (no overlap between steps); the "
                               "per-kernel times and rooflines below are measured in this mode"},
        "config": {
            "workload": "ring32x32 (configs[1]): 1024 chunks around a moving player per GPU, seed 1, default noise "
                        "(freq 0.09, gain 2.5, lacunarity 0.5, 3 octaves), neighbours present iff in the region",
            "chunks_per_step": n * world, "quads_per_step": quads_all / K,
            "parallelism": "chunk-coordinate sharding, one process per GPU, no collective on the data path",
            "pipelining": "steps alternate between 2 contexts / CUDA streams per GPU (own scratch and output buffers): "
                          "the HBM-bound emit of step i overlaps the issue-bound fill of step i+1; every step is a "
                          "complete vx_gen_mesh_device call",
            "numa_bound_cpus_rank0": numa_cpus,
            "l2": "outputs %.0f MB/step exceed the 126 MB L2; inputs are 12 KB of coordinates that change every "
                  "step; the 4 MB occupancy intermediate is produced and consumed inside the step" % (
                      q_step * INSTANCE_BYTES / 1e6),
        },
        "e2e": {"value": e2e_value, "unit": "chunks/s", "h2d_bytes_per_step": 12 * n + (n if h_masks else 0),
                "d2h_bytes_per_step": int(4 * n + 8 * (n + 1) + INSTANCE_BYTES * e2e_quads / K),
                "api": "vx_gen_mesh (host buffers, pinned instance buffers); two host threads with one context each "
                       "take alternate steps",
                "one_thread": {"value": world * n * K / e2e_serial_s,
                               "note": "one context, one call after the other"}},
        "e2e_packed_extension": e2e_packed,
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
                    "octaves that cannot flip the sign, so it executes fewer; the cell selection may not contract "
                    "a*b+c (bit-exact), the smooth part uses FMA under a certified error margin",
            "flop_per_voxel": FLOP_PER_VOXEL, "voxels_per_launch": n * VOXELS, "ms_per_launch": fill_ms,
            "traffic": traffic.get("fill"),
        },
        "roofline_mesh": {
            "kernel": "vx_mesh_emit_kernel", "bound": "hbm", "achieved": mesh_gbs, "peak": hbm_peak, "unit": "GB/s",
            "frac": (mesh_gbs / hbm_peak) if mesh_gbs else None, "peak_source": hbm_src,
            "bytes_per_launch": mesh_bytes, "ms_per_launch": emit_ms, "traffic": traffic.get("emit"),
        },
        "kernel_ms": k_ms,
        "peaks": {"fp32_ffma_tflops": ffma_peak, "fp32_fmul_fadd_tflops": nofuse_peak, "int32_tops": int_peak,
                  "hbm_gbs": hbm_peak},
        "cpu_baseline": cpu,
        "parity_spot_check": parity,
        "near_threshold_voxels": {"abs_noise_le_1e-6": census, "of": n * VOXELS},
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


def main():
    global _REAL_STDOUT
    sys.stdout.flush()
    _REAL_STDOUT = os.dup(1)
    os.dup2(2, 1)
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=50)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        run_reference(args, rank)
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
    run_ours(args, rank, world, local_rank)


if __name__ == "__main__":
    main()
