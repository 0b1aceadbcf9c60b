#!/usr/bin/env python
"""Measurements for the BASELINE.json configs that bench.py does not headline
(bench.py = configs[1]).  One JSON line per config; results are copied to profiles/.

  python tools/bench_configs.py [--region 1024] [--quick]

  config1  single chunk at the origin through the host API (latency)
  config3  region x region chunks generated + meshed in 8192-chunk batches, device resident
  config4  worst-case face density through the mesher only (checkerboard, caves)
  config5  frustum test of 1M chunk AABBs over a camera sweep, F frames per launch
"""
import argparse
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

HBM = 6553.9
if os.path.exists(os.path.join(ROOT, "MEASURED_PEAKS.json")):
    HBM = float(json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"])


def camera_planes(n_frames, vox):
    """Yaw sweep 0..360 deg, eye (0,16,0), pitch 0, perspective(aspect 1.6, fovy 60 deg, 0.1, 2000) in the
    camera.rs:41-51 composition (OPENGL_TO_WGPU * proj * view); harness side only."""
    f32 = np.float32
    gl = np.array([[1, 0, 0, 0], [0, 1, 0, 0], [0, 0, .5, 0], [0, 0, .5, 1]], dtype=f32)
    t = f32(1.0) / np.tan(f32(np.deg2rad(60.0)) / f32(2))
    near, far, aspect = f32(0.1), f32(2000.0), f32(1.6)
    proj = np.zeros((4, 4), f32)
    proj[0, 0], proj[1, 1] = t / aspect, t
    proj[2, 2], proj[2, 3], proj[3, 2] = (far + near) / (near - far), 2 * far * near / (near - far), -1
    out = np.empty((n_frames, 6, 4), f32)
    eye = np.array([0, 16, 0], f32)
    for i in range(n_frames):
        a = f32(2 * np.pi * i / n_frames)
        fwd = np.array([np.cos(a), 0, np.sin(a)], f32)
        s = np.cross(fwd, [0, 1, 0]).astype(f32)
        s /= np.linalg.norm(s)
        u = np.cross(s, fwd).astype(f32)
        view = np.eye(4, dtype=f32)
        view[0, :3], view[1, :3], view[2, :3] = s, u, -fwd
        view[:3, 3] = [-(s @ eye), -(u @ eye), fwd @ eye]
        m = (gl @ proj @ view).astype(f32)
        out[i] = vox.planes_from_matrix(m.T.reshape(-1))
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--region", type=int, default=1024)
    ap.add_argument("--quick", action="store_true")
    args = ap.parse_args()
    import torch
    import voxb200
    dev = torch.device("cuda", 0)
    ctx = voxb200.Context(devices=[0])
    ts = torch.cuda.Stream(device=dev)
    torch.cuda.set_stream(ts)
    stream = ts.cuda_stream

    def timed(fn, reps=1):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize()
        e0.record()
        for _ in range(reps):
            fn()
        e1.record()
        torch.cuda.synchronize()
        return e0.elapsed_time(e1) / reps

    # ---- config 1: single chunk latency through the host API --------------------------------
    c0 = np.array([[0, 0, 0]], np.int32)
    for mask in (0, 15):
        nm = np.array([mask], np.uint8)
        ctx.gen_mesh(c0, nbr_mask=nm)
        t0 = time.perf_counter()
        for _ in range(50):
            r = ctx.gen_mesh(c0, nbr_mask=nm)
        dt = (time.perf_counter() - t0) / 50
        print(json.dumps({"config": "1: single chunk at origin, host API vx_gen_mesh", "nbr_mask": mask,
                          "quads": r.total, "latency_us": dt * 1e6}))

    # ---- config 3: region batch ----------------------------------------------------------------
    R = 64 if args.quick else args.region
    rows_per_batch = max(1, 8192 // R)
    nb = rows_per_batch * R
    cap = nb * 8192
    d_counts = torch.zeros(nb, dtype=torch.int32, device=dev)
    d_offsets = torch.zeros(nb + 1, dtype=torch.int64, device=dev)
    d_inst = torch.empty(cap * 17, dtype=torch.int32, device=dev)
    d_status = torch.zeros(4, dtype=torch.int32, device=dev)
    xs = np.arange(-R // 2, R // 2, dtype=np.int32)
    batches = []
    for z0 in range(-R // 2, R // 2, rows_per_batch):
        zs = np.arange(z0, min(z0 + rows_per_batch, R // 2), dtype=np.int32)
        z, x = np.meshgrid(zs, xs, indexing="ij")
        c = np.stack([x.reshape(-1), np.zeros(x.size, np.int32), z.reshape(-1)], 1).astype(np.int32)
        m = ((c[:, 0] > -R // 2) * 1 + (c[:, 0] < R // 2 - 1) * 2 + (c[:, 2] > -R // 2) * 4 +
             (c[:, 2] < R // 2 - 1) * 8).astype(np.uint8)
        batches.append((torch.from_numpy(c).to(dev), torch.from_numpy(m).to(dev), len(c)))
    totals = torch.zeros(len(batches), dtype=torch.int64, device=dev)

    def run_region():
        for i, (dc, dm, n) in enumerate(batches):
            ctx.gen_mesh_device(dc.data_ptr(), n, d_counts.data_ptr(), d_offsets.data_ptr(), d_inst.data_ptr(), cap,
                                nbr_mask_ptr=dm.data_ptr(), status_ptr=d_status.data_ptr(), stream=stream)
            totals[i:i + 1].copy_(d_offsets[n:n + 1], non_blocking=True)
    run_region()
    ms = timed(run_region)
    q = int(totals.sum().item())
    print(json.dumps({"config": "3: %dx%d-chunk region, %d-chunk batches, device resident, 1 GPU" % (R, R, nb),
                      "chunks": R * R, "quads": q, "ms": ms, "chunks_per_s": R * R / (ms * 1e-3),
                      "quads_per_s": q / (ms * 1e-3), "instance_bytes": q * 68, "overflow": int(d_status[0].item())}))

    # ---- config 4: mesher only, worst-case face density -----------------------------------------
    n4 = 256 if args.quick else 1024
    g = np.indices((32, 32, 32)).sum(axis=0) & 1
    cb = np.where(g == 0, 0, 2).astype(np.uint8).reshape(-1)
    side = int(np.sqrt(n4))
    coords4 = np.array([(x, 0, z) for z in range(side) for x in range(side)], np.int32)
    caves = ctx.fill(coords4, params=voxb200.default_noise_params(freq=0.9, octaves=1))["blocks"]
    for name, blocks in (("checkerboard", np.broadcast_to(cb, (n4, 32768)).copy()), ("caves freq 0.9 oct 1", caves)):
        d_b = torch.from_numpy(blocks).to(dev)
        d_c = torch.from_numpy(coords4).to(dev)
        cap4 = n4 * 98304
        d_i4 = torch.empty(cap4 * 17, dtype=torch.int32, device=dev)
        d_cn = torch.zeros(n4, dtype=torch.int32, device=dev)
        d_of = torch.zeros(n4 + 1, dtype=torch.int64, device=dev)

        def run4():
            ctx.mesh_device(d_c.data_ptr(), n4, d_b.data_ptr(), d_cn.data_ptr(), d_of.data_ptr(), d_i4.data_ptr(),
                            cap4, status_ptr=d_status.data_ptr(), stream=stream)
        run4()
        ctx.set_timing(True)
        ms = timed(run4, 5)
        kt = ctx.get_timing()
        ctx.set_timing(False)
        q = int(d_of[n4].item())
        emit_ms = kt["emit"][0] / kt["emit"][1]
        bytes_ = n4 * (36864 + 12) + 68 * q
        print(json.dumps({"config": "4: mesher only, " + name, "chunks": n4, "quads": q, "quads_per_chunk": q / n4,
                          "ms_total": ms, "ms_emit": emit_ms, "ms_pack": kt["pack"][0] / kt["pack"][1],
                          "ms_count": kt["count"][0] / kt["count"][1],
                          "chunks_per_s": n4 / (ms * 1e-3), "quads_per_s": q / (ms * 1e-3),
                          "emit_GBps": bytes_ / (emit_ms * 1e-3) / 1e9, "emit_frac_of_hbm": bytes_ / (emit_ms * 1e-3) / 1e9 / HBM}))
        del d_i4

    # ---- config 5: frustum ----------------------------------------------------------------------
    side = 300 if args.quick else 1000
    gx = np.arange(-side // 2, side // 2, dtype=np.int32)
    z, x = np.meshgrid(gx, gx, indexing="ij")
    c5 = np.stack([x.reshape(-1), np.zeros(x.size, np.int32), z.reshape(-1)], 1).astype(np.int32)
    n5 = len(c5)
    planes = camera_planes(1000, voxb200)
    d_c5 = torch.from_numpy(c5).to(dev)
    d_pl = torch.from_numpy(planes).to(dev)
    for F in (1, 8, 64, 1000):
        d_bits = torch.zeros(F * ((n5 + 31) // 32), dtype=torch.int32, device=dev)
        d_vis = torch.zeros(F * n5, dtype=torch.uint8, device=dev)
        for mode, vp, bp in (("bits", None, d_bits.data_ptr()), ("bytes", d_vis.data_ptr(), None)):
            launches = 1 if F == 1000 else min(1000 // F, 64)

            def run5():
                for j in range(launches):
                    ctx.frustum_cull_device(d_pl.data_ptr() + j * F * 96, F, d_c5.data_ptr(), n5, visible_ptr=vp,
                                            bits_ptr=bp, stream=stream)
            out_b = (n5 * F if mode == "bytes" else ((n5 + 31) // 32) * 4 * F)
            run5()
            ms = timed(run5)
            # the loop above is bound by the host's launch path for small F (ctypes call + lock + launch ~ 10 us);
            # device time per launch: CUDA events around every launch (vx_set_timing) and a CUDA-graph replay
            ctx.set_timing(True)
            run5()
            torch.cuda.synchronize()
            kt5 = ctx.get_timing()
            ctx.set_timing(False)
            dev_ms = kt5["frustum"][0] / max(kt5["frustum"][1], 1)
            graph_ms = None
            try:
                g = torch.cuda.CUDAGraph()
                ts5 = torch.cuda.Stream(device=dev)
                with torch.cuda.graph(g, stream=ts5):
                    for j in range(launches):
                        ctx.frustum_cull_device(d_pl.data_ptr() + j * F * 96, F, d_c5.data_ptr(), n5, visible_ptr=vp,
                                                bits_ptr=bp, stream=ts5.cuda_stream)
                g.replay()
                torch.cuda.synchronize()
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                for _ in range(5):
                    g.replay()
                e1.record()
                torch.cuda.synchronize()
                graph_ms = e0.elapsed_time(e1) / 5 / launches
            except Exception as e:                       # pragma: no cover
                graph_ms = repr(e)[:80]
            tests = n5 * F * launches
            vis = int(torch.count_nonzero(d_vis[:n5]).item()) if mode == "bytes" else None
            print(json.dumps({"config": "5: frustum, %d AABBs x %d frames per launch, %s out" % (n5, F, mode),
                              "launches": launches, "ms_per_launch": ms / launches,
                              "device_ms_per_launch_events": dev_ms, "ms_per_launch_in_cuda_graph": graph_ms,
                              "GBps_algorithmic_device": (12 * n5 + out_b) / (dev_ms * 1e-3) / 1e9,
                              "tests_per_s": tests / (ms * 1e-3), "visible_frame0": vis,
                              "GBps_algorithmic": (12 * n5 + out_b) / (ms / launches * 1e-3) / 1e9}))


if __name__ == "__main__":
    main()
