#!/usr/bin/env python
"""Full-size check of the fast fill kernel (sign-only octave skipping + fused arithmetic + table
gradient) against the exact kernel: occupancy of every voxel of an R x R chunk region, compared on
the device.  The exact kernel is the one the noise-parity tests pin bit for bit against the oracle.

  python tools/verify_sign_only.py [--region 1024] [--seeds 1,2,12345]
"""
import argparse
import json
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--region", type=int, default=1024)
    ap.add_argument("--seeds", default="1")
    ap.add_argument("--batch", type=int, default=8192)
    ap.add_argument("--freq", type=float, default=None)
    ap.add_argument("--lacunarity", type=float, default=None)
    ap.add_argument("--gain", type=float, default=None)
    ap.add_argument("--scale", type=float, default=None, help="vx_noise_params.simplex_scale (default: the library's)")
    ap.add_argument("--origin", default="0,0,0", help="chunk coordinate of the region centre (x,y,z)")
    args = ap.parse_args()
    import torch
    import voxb200
    dev = torch.device("cuda", 0)
    ctx = voxb200.Context(devices=[0])
    ts = torch.cuda.Stream(device=dev)
    torch.cuda.set_stream(ts)
    st = ts.cuda_stream
    R = args.region
    xs = np.arange(-R // 2, R // 2, dtype=np.int32)
    rows = max(1, args.batch // R)
    nb = rows * R
    occ_fast = torch.zeros(nb * 1024, dtype=torch.int32, device=dev)
    occ_exact = torch.zeros(nb * 1024, dtype=torch.int32, device=dev)
    nz = torch.zeros(1, dtype=torch.int64, device=dev)
    for seed in [int(v) for v in args.seeds.split(",")]:
        kw = {k: v for k, v in (("freq", args.freq), ("lacunarity", args.lacunarity), ("gain", args.gain),
                                    ("simplex_scale", args.scale)) if v is not None}
        p = voxb200.default_noise_params(seed=seed, **kw)
        ox, oy, oz = (int(v) for v in args.origin.split(","))
        voxels = mismatched_words = near = solid = 0
        t0 = time.time()
        for z0 in range(-R // 2, R // 2, rows):
            zs = np.arange(z0, min(z0 + rows, R // 2), dtype=np.int32)
            z, x = np.meshgrid(zs, xs, indexing="ij")
            c = np.stack([x.reshape(-1) + ox, np.full(x.size, oy, np.int32), z.reshape(-1) + oz], 1).astype(np.int32)
            n = len(c)
            d_c = torch.from_numpy(c).to(dev)
            ctx.fill_device(d_c.data_ptr(), n, params=p, occ_ptr=occ_fast.data_ptr(), stream=st)
            ctx.fill_device(d_c.data_ptr(), n, params=p, occ_ptr=occ_exact.data_ptr(), near_zero_ptr=nz.data_ptr(),
                            stream=st)
            torch.cuda.synchronize()
            a, b = occ_fast[:n * 1024], occ_exact[:n * 1024]
            mismatched_words += int((a != b).sum().item())
            near += int(nz.item())
            voxels += n * 32768
        print(json.dumps({"check": "fast fill kernel vs exact kernel, occupancy of every voxel",
                          "region": "%dx%d chunks" % (R, R), "origin": [ox, oy, oz], "params": kw, "seed": seed, "voxels": voxels,
                          "mismatched_occupancy_words": mismatched_words,
                          "near_threshold_voxels_abs_noise_le_1e-6": near, "seconds": round(time.time() - t0, 2)}))
        assert mismatched_words == 0


if __name__ == "__main__":
    main()
