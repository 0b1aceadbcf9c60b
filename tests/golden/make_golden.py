#!/usr/bin/env python
"""Independent numpy restatement of the hot path -> tests/golden/golden.json.

The reference (Rust + un-vendored simdnoise 3.1.6) cannot be built or run in
this image and holds no tests or golden vectors, so these fixtures come from a
SECOND, independently written implementation (vectorised numpy float32, array
shifts for the mesher) of the same published algorithm.  They pin the C oracle
(oracle/vox_oracle.c) against transcription drift; they do NOT pin it against
the real crate ("parity unpinned" for the noise values).

SURVEY.md App. A lists figures from a third throwaway restatement; those are
embedded below as `survey_app_a` and must match too.

Run:  python tests/golden/make_golden.py     (rewrites golden.json)
"""
import json
import os
import zlib

import numpy as np

f32 = np.float32
F3 = f32(1.0) / f32(3.0)
G3 = f32(1.0) / f32(6.0)
G33 = f32(3.0) / f32(6.0) - f32(1.0)
XP, YP, ZP = np.uint32(1619), np.uint32(31337), np.uint32(6971)
# simplex_3d's normaliser in simdnoise's later 3.1.x sources; the "raw" records use 1.0 (SURVEY App. A form)
SCALED = f32(32.69587493801679)


def _hash(seed, i, j, k):
    h = i ^ np.uint32(seed & 0xFFFFFFFF) ^ j ^ k
    h = h * h * np.uint32(60493) * h
    return (h.view(np.int32) >> 13).view(np.uint32) ^ h


def _grad(seed, i, j, k, x, y, z):
    h = _hash(seed, i, j, k)
    h13 = h & np.uint32(13)
    u = np.where(h13 < 8, x, y)
    v = np.where(h13 < 2, y, np.where(h13 == 12, x, z))
    su = (u.view(np.uint32) ^ (h << np.uint32(31))).view(f32)
    sv = (v.view(np.uint32) ^ ((h & np.uint32(2)) << np.uint32(30))).view(f32)
    return su + sv


def simplex3(x, y, z, seed):
    f = F3 * ((x + y) + z)
    x0 = np.floor(x + f)
    y0 = np.floor(y + f)
    z0 = np.floor(z + f)
    i = x0.astype(np.int32).view(np.uint32) * XP
    j = y0.astype(np.int32).view(np.uint32) * YP
    k = z0.astype(np.int32).view(np.uint32) * ZP
    g = G3 * ((x0 + y0) + z0)
    x0 = x - (x0 - g)
    y0 = y - (y0 - g)
    z0 = z - (z0 - g)
    xy, yz, xz = x0 >= y0, y0 >= z0, x0 >= z0
    i1, j1, k1 = xy & xz, ~xy & yz, ~xz & ~yz
    i2, j2, k2 = xy | xz, ~xy | yz, ~(xz & yz)
    m1 = f32(-1.0)
    zero = f32(0.0)

    def off(a, m, c):
        return (a + np.where(m, m1, zero)) + c

    x1, y1, z1 = off(x0, i1, G3), off(y0, j1, G3), off(z0, k1, G3)
    x2, y2, z2 = off(x0, i2, F3), off(y0, j2, F3), off(z0, k2, F3)
    x3, y3, z3 = x0 + G33, y0 + G33, z0 + G33

    def tw(a, b, c):
        t = ((f32(0.6) - a * a) - b * b) - c * c
        t = np.where(t >= 0, t, zero)
        t = t * t
        return t * t

    def sel(m, p):
        return np.where(m, p, np.uint32(0))

    v0 = tw(x0, y0, z0) * _grad(seed, i, j, k, x0, y0, z0)
    v1 = tw(x1, y1, z1) * _grad(seed, i + sel(i1, XP), j + sel(j1, YP), k + sel(k1, ZP), x1, y1, z1)
    v2 = tw(x2, y2, z2) * _grad(seed, i + sel(i2, XP), j + sel(j2, YP), k + sel(k2, ZP), x2, y2, z2)
    v3 = tw(x3, y3, z3) * _grad(seed, i + XP, j + YP, k + ZP, x3, y3, z3)
    return ((v3 + v2) + v1) + v0


def noise_chunk(coord, seed=1, freq=0.09, gain=2.5, lac=0.5, octaves=3, scale=SCALED):
    cx, cy, cz = (int(c) for c in coord)
    ar = np.arange(32, dtype=f32)
    sx = f32(np.int32(cx * 32)) + ar          # simdnoise x = engine x
    sy = f32(np.int32(cz * 32)) + ar          # simdnoise y = engine z
    sz = f32(np.int32(cy * 32)) + ar          # simdnoise z = engine y
    Z, Y, X = np.meshgrid(sz, sy, sx, indexing="ij")   # out[z][y][x] == engine idx
    freq, gain, lac, scale = f32(freq), f32(gain), f32(lac), f32(scale)
    x, y, z = X * freq, Y * freq, Z * freq
    amp = f32(1.0)
    with np.errstate(over="ignore"):
        r = simplex3(x, y, z, seed) * scale
        for _ in range(1, octaves):
            x, y, z = x * lac, y * lac, z * lac
            amp = amp * gain
            r = (simplex3(x, y, z, seed) * scale) * amp + r
    assert r.dtype == f32
    return r.reshape(-1)


def face_masks(solid, adj):
    """solid: bool [y][z][x]; adj: 4 bool grids or None, order -x,+x,-z,+z."""
    s = solid
    air = ~s

    def nb(axis, shift, k, border_index):
        n = np.roll(air, shift, axis=axis)
        sl = [slice(None)] * 3
        sl[axis] = border_index
        if k is None:                      # y borders always emit
            n[tuple(sl)] = True
        elif adj[k] is None:
            n[tuple(sl)] = True
        else:
            other = [slice(None)] * 3
            other[axis] = 31 - border_index
            n[tuple(sl)] = ~adj[k][tuple(other)]
        return s & n

    left = nb(2, 1, 0, 0)
    right = nb(2, -1, 1, 31)
    back = nb(1, 1, 2, 0)
    front = nb(1, -1, 3, 31)
    down = nb(0, 1, None, 0)
    up = nb(0, -1, None, 31)
    return [left, right, back, front, down, up]


def chunk_record(coord, scale=SCALED):
    noise = noise_chunk(coord, scale=scale)
    occ = np.where(noise < 0, 0, 2).astype(np.uint8)
    solid = (occ != 2).reshape(32, 32, 32)
    cx, cy, cz = coord
    nbc = [(cx - 1, cy, cz), (cx + 1, cy, cz), (cx, cy, cz - 1), (cx, cy, cz + 1)]
    adj = [(noise_chunk(c, scale=scale) < 0).reshape(32, 32, 32) for c in nbc]
    none = [int(m.sum()) for m in face_masks(solid, [None] * 4)]
    allf = [int(m.sum()) for m in face_masks(solid, adj)]
    # emission-order digest: for every emitted quad (y,z,x major; L,R,B,F,D,U)
    # the tuple (idx, face); crc over the packed uint32 stream idx*8+face
    masks = np.stack(face_masks(solid, adj), axis=-1)            # [y][z][x][f]
    flat = masks.reshape(-1, 6)
    idx, face = np.nonzero(flat)
    stream = (idx.astype(np.uint32) * 8 + face.astype(np.uint32)).astype("<u4")
    sample = noise[::997]
    # full InstanceRaw stream (quad.rs:47-54, renderer/instance.rs:1-6) built independently:
    # t_matrix[c][r] = column c of R(face) for c < 3, (x, y, z, 1) for c = 3; texture per slot
    rot = _rotation_matrices()
    tex_of_face = np.array([0, 0, 0, 0, 1, 2], np.uint32)     # GRASS [0,0,0,0,2,1] via slots 0,1,2,3,5,4
    inst = np.zeros(len(idx), dtype=[("t_matrix", "<f4", (4, 4)), ("texture_index", "<u4")])
    inst["t_matrix"][:, :3, :3] = np.transpose(rot[face], (0, 2, 1))
    inst["t_matrix"][:, 3, 0] = cx * 32 + (idx & 31)
    inst["t_matrix"][:, 3, 1] = cy * 32 + (idx >> 10)
    inst["t_matrix"][:, 3, 2] = cz * 32 + ((idx >> 5) & 31)
    inst["t_matrix"][:, 3, 3] = 1.0
    inst["texture_index"] = tex_of_face[face]
    return {
        "instances_crc32_all": "%08x" % zlib.crc32(inst.tobytes()),
        "coord": list(coord),
        "solid": int(solid.sum()),
        "occ_crc32": "%08x" % zlib.crc32(occ.tobytes()),
        "faces_none": none,
        "faces_all": allf,
        "quad_stream_crc32_all": "%08x" % zlib.crc32(stream.tobytes()),
        "noise_1000": "%08x" % int(noise[1000:1001].view(np.uint32)[0]),
        "noise_sample_997": ["%08x" % int(v) for v in sample.view(np.uint32)],
        "near_zero_1e-6": int((np.abs(noise) <= f32(1e-6)).sum()),
    }


SURVEY_APP_A = [   # SURVEY.md App. A table (third restatement)
    {"coord": [0, 0, 0], "solid": 19303, "occ_crc32": "f8ac6530",
     "faces_none": [1041, 1041, 1088, 1088, 1039, 1039], "faces_all": [306, 864, 507, 830, 1039, 1039],
     "noise_1000": "3db7963d"},
    {"coord": [1, 0, 0], "solid": 7233, "occ_crc32": "ab5cb55e",
     "faces_none": [900, 900, 582, 582, 886, 886], "faces_all": [723, 346, 270, 500, 886, 886],
     "noise_1000": "3d58dbd5"},
    {"coord": [-1, 0, -1], "solid": 24415, "occ_crc32": "004a1285",
     "faces_none": [1077, 1077, 1407, 1407, 1075, 1075], "faces_all": [627, 547, 816, 432, 1075, 1075],
     "noise_1000": "3d0a47f7"},
    {"coord": [511, 0, -512], "solid": 11880, "occ_crc32": "250757d2",
     "faces_none": [1507, 1507, 957, 957, 1160, 1160], "faces_all": [726, 962, 845, 592, 1160, 1160],
     "noise_1000": "3d7a96d7"},
]


def _rotation_matrices():
    """[face][row][col] float32, faces L R B F D U, signed zeros canonicalised."""
    r = rotations()
    return np.stack([np.array([int(h, 16) for h in r[str(f)]], np.uint32).view(f32).reshape(3, 3) for f in range(6)])


def rotations():
    """quad.rs:16-40 with nalgebra's Rz*Ry*Rx, float32 sin/cos."""
    def rad(d):
        return f32(d) * (f32(np.pi) / f32(180.0))

    def euler(roll, pitch, yaw):
        sr, cr = np.sin(f32(roll)), np.cos(f32(roll))
        sp, cp = np.sin(f32(pitch)), np.cos(f32(pitch))
        sy, cy = np.sin(f32(yaw)), np.cos(f32(yaw))
        R = np.array([[cy * cp, cy * sp * sr - sy * cr, cy * sp * cr + sy * sr],
                      [sy * cp, sy * sp * sr + cy * cr, sy * sp * cr - cy * sr],
                      [-sp, cp * sr, cp * cr]], dtype=f32)
        return R + f32(0.0)   # canonicalise -0
    z = f32(0.0)
    faces = {  # oracle face ids: L R B F D U
        0: euler(z, rad(-90.0), z), 1: euler(z, rad(90.0), z), 2: euler(z, rad(180.0), z),
        3: euler(z, z, z), 4: euler(rad(90.0), z, z), 5: euler(rad(-90.0), z, z)}
    return {str(k): ["%08x" % int(v) for v in R.reshape(-1).view(np.uint32)] for k, R in faces.items()}


def frustum_fixture():
    """Planes + visibility for a deterministic look-at matrix (harness side)."""
    def perspective(aspect, fovy, near, far):
        t = f32(1.0) / np.tan(f32(fovy) / f32(2.0))
        m = np.zeros((4, 4), dtype=f32)
        m[0, 0] = t / f32(aspect)
        m[1, 1] = t
        m[2, 2] = (f32(far) + f32(near)) / (f32(near) - f32(far))
        m[2, 3] = (f32(2.0) * f32(far) * f32(near)) / (f32(near) - f32(far))
        m[3, 2] = f32(-1.0)
        return m

    def look_at_rh(eye, target, up):
        eye, target, up = (np.asarray(v, dtype=f32) for v in (eye, target, up))
        fwd = target - eye
        fwd = fwd / np.sqrt((fwd * fwd).sum(dtype=f32))
        s = np.cross(fwd, up).astype(f32)
        s = s / np.sqrt((s * s).sum(dtype=f32))
        u = np.cross(s, fwd).astype(f32)
        m = np.eye(4, dtype=f32)
        m[0, :3], m[1, :3], m[2, :3] = s, u, -fwd
        m[0, 3] = -(s * eye).sum(dtype=f32)
        m[1, 3] = -(u * eye).sum(dtype=f32)
        m[2, 3] = (fwd * eye).sum(dtype=f32)
        return m

    gl = np.array([[1, 0, 0, 0], [0, 1, 0, 0], [0, 0, .5, 0], [0, 0, .5, 1]], dtype=f32)  # camera.rs:18-23
    out = []
    for yaw in (0.0, 37.0, 123.0, 270.0):
        r = np.deg2rad(f32(yaw)).astype(f32)
        eye = np.array([3.0, 16.0, -7.0], dtype=f32)
        tgt = eye + np.array([np.cos(r), f32(-0.1), np.sin(r)], dtype=f32)
        M = (gl @ perspective(1.6, 1.0472, 0.1, 600.0) @ look_at_rh(eye, tgt, [0, 1, 0])).astype(f32)
        colmajor = M.T.reshape(-1).copy()      # data[c][r]
        d = colmajor.reshape(4, 4)
        pl = {
            "left": d[:, 3] + d[:, 0], "right": d[:, 3] - d[:, 0], "top": d[:, 3] - d[:, 1],
            "bottom": d[:, 3] + d[:, 1], "near": d[:, 3] + d[:, 2], "far": d[:, 3] - d[:, 2]}
        planes = np.stack([pl[k] for k in ("near", "far", "left", "right", "top", "bottom")]).astype(f32)
        g = np.arange(-12, 12, dtype=np.int32)
        cz, cx = np.meshgrid(g, g, indexing="ij")
        coords = np.stack([cx.reshape(-1), np.zeros(cx.size, np.int32), cz.reshape(-1)], axis=1)
        p = coords.astype(f32) * f32(32.0)
        vis = np.zeros(coords.shape[0], dtype=bool)
        for ox, oy, oz in [(0, 0, 0), (1, 0, 0), (0, 0, 1), (1, 0, 1), (0, 1, 0), (1, 1, 0), (0, 1, 1), (1, 1, 1)]:
            ex = p[:, 0] + f32(32.0) * f32(ox)
            ey = p[:, 1] + f32(32.0) * f32(oy)
            ez = p[:, 2] + f32(32.0) * f32(oz)
            inside = np.ones(coords.shape[0], dtype=bool)
            for a, b, c, dd in planes:
                dist = ((a * ex + b * ey) + c * ez) + dd
                inside &= ~(dist < 0)
            vis |= inside
        out.append({
            "matrix_colmajor": ["%08x" % int(v) for v in colmajor.view(np.uint32)],
            "planes": ["%08x" % int(v) for v in planes.reshape(-1).view(np.uint32)],
            "grid": [-12, 12],
            "visible_count": int(vis.sum()),
            "vis_crc32": "%08x" % zlib.crc32(vis.astype(np.uint8).tobytes()),
        })
    return out


def main():
    coords = [(0, 0, 0), (1, 0, 0), (-1, 0, -1), (511, 0, -512), (-16, 0, 15), (7, 1, -3), (0, -2, 0)]
    gold = {
        "about": "numpy restatement (tests/golden/make_golden.py); noise parity vs simdnoise 3.1.6 is UNPINNED",
        "simplex_scale": float(SCALED),
        "chunks": [chunk_record(c) for c in coords],                      # default: simplex_3d normalised
        "chunks_raw": [chunk_record(c, scale=1.0) for c in coords],       # simplex_scale = 1 (raw sum)
        "survey_app_a": SURVEY_APP_A,
        "rotations": rotations(),
        "frustum": frustum_fixture(),
    }
    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden.json")
    with open(path, "w") as fh:
        json.dump(gold, fh, indent=1)
    print("wrote", path)
    for rec, ref in zip(gold["chunks_raw"], SURVEY_APP_A):
        for key in ("solid", "occ_crc32", "faces_none", "faces_all", "noise_1000"):
            status = "ok" if rec[key] == ref[key] else "MISMATCH"
            print(rec["coord"], key, rec[key], ref[key], status)


if __name__ == "__main__":
    main()
