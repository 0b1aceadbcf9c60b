#!/usr/bin/env python
"""Small end-to-end exercise of every kernel for compute-sanitizer (memcheck / racecheck / synccheck).
  compute-sanitizer --tool racecheck python tools/sanitize_smoke.py
"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import voxb200  # noqa: E402

ctx = voxb200.Context()
coords = np.array([(x, 0, z) for z in range(-2, 2) for x in range(-2, 2)], np.int32)
r = ctx.gen_mesh(coords, blocks=True)                                   # resolve, lattice fill, count+scan, emit
f5 = ctx.fill(coords[:2], params=voxb200.default_noise_params(freq=0.9), blocks=True)         # hash (brick2) fill
f6 = ctx.fill(np.array([(40000000, 0, -40000000)], np.int32), blocks=True)   # lattice kernel, boxes that do not fit
m = np.full(len(coords), 15, np.uint8)
r2 = ctx.gen_mesh(coords[:5], nbr_mask=m[:5])                           # halo fill; short-list path, pageable output
pin = voxb200.PinnedBuffer(5 * 98304 * 68)
r3 = ctx.gen_mesh(coords[:5], nbr_mask=m[:5], out=pin.view(np.uint8))   # short-list path, emit straight into pinned host memory
assert r3.total == r2.total and r3.instances.tobytes() == r2.instances.tobytes()
f = ctx.fill(coords[:3], noise=True, census=True, occupancy=True)       # exact fill
f4 = ctx.fill(coords[:2], params=voxb200.default_noise_params(octaves=4))   # generic octave loop
g = np.indices((32, 32, 32)).sum(axis=0) & 1
cb = np.where(g == 0, 0, 2).astype(np.uint8).reshape(-1)
mm = ctx.mesh(coords[:2], np.stack([cb, r.blocks[0]]))                   # pack, count, emit (98304 quads)
planes = np.random.default_rng(0).standard_normal((70, 6, 4)).astype(np.float32)
planes[:40, :, 3] += 300.0            # frames decided by the extreme corners alone / frames full of straddling chunks
planes[3, 2, 0] = np.nan              # a frame that takes the corner-by-corner test for every chunk
v = ctx.frustum_cull(planes, coords)
b = ctx.frustum_cull(planes, coords, bits=True)
# single-frame kernel: ragged tail, and more than one wave of CTAs
big = np.zeros((700003, 3), np.int32)
big[:, 0] = np.arange(700003) % 997 - 400
big[:, 2] = np.arange(700003) // 997 - 300
v1 = ctx.frustum_cull(planes[:1], coords[:13])
b1 = ctx.frustum_cull(planes[:1], big, bits=True)
v2 = ctx.frustum_cull(planes[:1], big)
assert int(np.unpackbits(b1.view(np.uint8), bitorder="little")[:700003].sum()) == int(v2.sum())
# extension: vertical neighbours (in-list, halo y planes) + DIRT rule, blocks from the occupancy planes
st = np.array([(x, y, 0) for y in (-1, 0, 1) for x in (0, 1)], np.int32)
e1 = ctx.gen_mesh(st, blocks=True, vertical=True, dirt_rule=True)
e2 = ctx.gen_mesh(st[:3], nbr_mask=np.full(3, 63, np.uint8), blocks=True, vertical=True, dirt_rule=True)
f7 = ctx.fill(coords[:2], params=voxb200.default_noise_params(simplex_scale=1.0), blocks=True)   # raw simplex sum
print("ok", r.total, r2.total, f["near_zero"], mm.total, int(v.sum()), int(v2.sum()), e1.total, e2.total, ctx.launch_count)
