"""CPU: the C oracle against the committed golden vectors (tests/golden/golden.json,
numpy restatement + SURVEY App. A) and against its own AVX2 lane path."""
import zlib

import numpy as np
import pytest


def _neighbours(c):
    cx, cy, cz = c
    return [(cx - 1, cy, cz), (cx + 1, cy, cz), (cx, cy, cz - 1), (cx, cy, cz + 1)]


def test_survey_app_a_vectors(oracle, golden):
    # SURVEY App. A restates the raw simplex sum: simplex_scale = 1
    raw = oracle.default_params(simplex_scale=1.0)
    for ref in golden["survey_app_a"]:
        c = ref["coord"]
        noise = oracle.noise_chunk(c, raw)
        cubes = oracle.terrain_from_noise(noise)
        assert int((cubes != oracle.AIR).sum()) == ref["solid"]
        assert "%08x" % zlib.crc32(cubes.tobytes()) == ref["occ_crc32"]
        assert "%08x" % int(noise[1000:1001].view(np.uint32)[0]) == ref["noise_1000"]
        q = oracle.cull_unseen_triangles(c, cubes, None)
        assert np.bincount(q["face"], minlength=6).tolist() == ref["faces_none"]
        adj = [oracle.generate_terrain(n, raw) for n in _neighbours(c)]
        q = oracle.cull_unseen_triangles(c, cubes, adj)
        assert np.bincount(q["face"], minlength=6).tolist() == ref["faces_all"]


@pytest.mark.parametrize("key,scale", [("chunks", None), ("chunks_raw", 1.0)])
def test_numpy_golden_chunks(oracle, golden, key, scale):
    prm = oracle.default_params() if scale is None else oracle.default_params(simplex_scale=scale)
    assert np.float32(oracle.default_params().simplex_scale) == np.float32(golden["simplex_scale"])
    for rec in golden[key]:
        c = rec["coord"]
        noise = oracle.noise_chunk(c, prm)
        assert ["%08x" % int(v) for v in noise[::997].view(np.uint32)] == rec["noise_sample_997"]
        assert int((np.abs(noise) <= np.float32(1e-6)).sum()) == rec["near_zero_1e-6"]
        cubes = oracle.terrain_from_noise(noise)
        assert "%08x" % zlib.crc32(cubes.tobytes()) == rec["occ_crc32"]
        assert int((cubes != oracle.AIR).sum()) == rec["solid"]
        adj = [oracle.generate_terrain(n, prm) for n in _neighbours(c)]
        q = oracle.cull_unseen_triangles(c, cubes, adj)
        assert np.bincount(q["face"], minlength=6).tolist() == rec["faces_all"]
        # emission order: (y,z,x)-major then L,R,B,F,D,U  (chunk.rs:65-69,77-240)
        w = np.array(c, dtype=np.int64) * 32
        idx = (q["x"] - w[0]) + 32 * (q["z"] - w[2]) + 1024 * (q["y"] - w[1])
        stream = (idx.astype(np.uint32) * 8 + q["face"]).astype("<u4")
        assert "%08x" % zlib.crc32(stream.tobytes()) == rec["quad_stream_crc32_all"]
        # the full 68-byte instance stream against the numpy restatement's packing
        inst = oracle.create_mesh(c, cubes, adj)
        inst["t_matrix"] += np.float32(0)
        assert "%08x" % zlib.crc32(inst.tobytes()) == rec["instances_crc32_all"]
        q0 = oracle.cull_unseen_triangles(c, cubes, None)
        assert np.bincount(q0["face"], minlength=6).tolist() == rec["faces_none"]


def test_avx2_lanes_equal_scalar(oracle):
    if not oracle.lib().vxo_have_avx2():
        pytest.skip("no AVX2 on this host")
    for c in [(0, 0, 0), (-3, 2, 9), (511, 0, -512), (100000, -7, -99999)]:
        a = oracle.noise_chunk(c, simd=False)
        b = oracle.noise_chunk(c, simd=True)
        assert np.array_equal(a.view(np.uint32), b.view(np.uint32))


def test_noise_properties(oracle):
    n0 = oracle.noise_chunk((0, 0, 0))
    assert n0[0] == 0.0 and n0[1] == 0.0            # noise[0..2] exactly 0 -> AIR
    assert oracle.terrain_from_noise(n0)[0] == oracle.AIR
    assert -0.35 * 32.7 < float(n0.min()) < float(n0.max()) < 0.35 * 32.7
    raw = oracle.noise_chunk((0, 0, 0), oracle.default_params(simplex_scale=1.0))
    assert -0.35 < float(raw.min()) < float(raw.max()) < 0.35
    # the normalised form is the raw form times 32.69587 up to the per-octave rounding
    assert np.allclose(n0, raw * np.float32(32.69587493801679), rtol=1e-5, atol=1e-5)
    assert (np.sign(n0) != np.sign(raw)).sum() <= 4      # only near-zero sums can flip
    # seed sensitivity and determinism
    p2 = oracle.default_params(seed=2)
    assert not np.array_equal(oracle.noise_chunk((0, 0, 0), p2), n0)
    assert np.array_equal(oracle.noise_chunk((0, 0, 0)), n0)
    # continuity across a chunk border: world-adjacent voxels differ only slightly
    n1 = oracle.noise_chunk((1, 0, 0)).reshape(32, 32, 32)
    d = np.abs(n0.reshape(32, 32, 32)[:, :, 31] - n1[:, :, 0])
    assert float(d.max()) < 0.05 * 32.7
    # octaves=1 equals plain simplex at freq-scaled coords
    p1 = oracle.default_params(octaves=1)
    n = oracle.noise_chunk((2, 0, 1), p1)
    L = oracle.lib()
    f = np.float32(0.09)
    x, y, z = np.float32(64 + 5), np.float32(3), np.float32(32 + 7)     # engine (x,y,z) = (5,3,7)
    want = L.vxo_simplex_3d(x * f, z * f, y * f, 1)                    # simdnoise (x,y,z)=(x,z,y)
    assert np.float32(want) * np.float32(p1.simplex_scale) == n[5 + 32 * 7 + 1024 * 3]


def test_threshold_rule(oracle):
    # chunk.rs:38: strictly `< 0.` ; +0, -0 and NaN are AIR
    noise = np.zeros(32768, dtype=np.float32)
    noise[1] = -0.0
    noise[2] = -1e-30
    noise[3] = np.nan
    noise[4] = 1e-30
    cubes = oracle.terrain_from_noise(noise)
    assert cubes[:5].tolist() == [2, 2, 0, 2, 2]


def test_rotations_match_golden_and_survey_app_b(oracle, golden):
    e = np.float32(-4.371139e-8)
    d = np.float32(8.742278e-8)
    want = {
        oracle.UP: [[1, 0, 0], [0, e, 1], [0, -1, e]],
        oracle.DOWN: [[1, 0, 0], [0, e, -1], [0, 1, e]],
        oracle.LEFT: [[e, 0, -1], [0, 1, 0], [1, 0, e]],
        oracle.RIGHT: [[e, 0, 1], [0, 1, 0], [-1, 0, e]],
        oracle.FRONT: [[1, 0, 0], [0, 1, 0], [0, 0, 1]],
        oracle.BACK: [[-1, 0, -d], [0, 1, 0], [d, 0, -1]],
    }
    for face, R in want.items():
        got = oracle.face_rotation(face)
        assert np.array_equal(got, np.array(R, dtype=np.float32)), face
        gold = np.array([int(h, 16) for h in golden["rotations"][str(face)]], dtype=np.uint32).view(np.float32)
        assert np.array_equal(got.reshape(-1) + np.float32(0), gold)
        # the unit quad lies in z=+0.5 (quad.rs:66-87): normal = R*(0,0,1)
    normals = {oracle.LEFT: (-1, 0, 0), oracle.RIGHT: (1, 0, 0), oracle.BACK: (0, 0, -1),
               oracle.FRONT: (0, 0, 1), oracle.DOWN: (0, -1, 0), oracle.UP: (0, 1, 0)}
    for face, nrm in normals.items():
        assert np.allclose(oracle.face_rotation(face)[:, 2], nrm, atol=1e-6)


def test_instance_packing(oracle):
    cubes = np.full(32768, oracle.AIR, dtype=np.uint8)
    cubes[5 + 32 * 6 + 1024 * 7] = oracle.GRASS
    cubes[0] = oracle.DIRT
    inst = oracle.create_mesh((2, -1, 3), cubes, None)
    assert inst.dtype.itemsize == 68 and len(inst) == 12
    # first cube in (y,z,x) order is the DIRT one at (0,0,0): textures all 1 (texture.rs:242)
    assert inst["texture_index"][:6].tolist() == [1] * 6
    # GRASS: sides 0, DOWN uses slot 5 -> 1, UP uses slot 4 -> 2 (texture.rs:241, chunk.rs:205-240)
    assert inst["texture_index"][6:].tolist() == [0, 0, 0, 0, 1, 2]
    m = inst["t_matrix"][6]                       # LEFT face of the grass cube
    assert m[3].tolist() == [64 + 5, -32 + 7, 96 + 6, 1.0]
    assert m[0][3] == 0 and m[1][3] == 0 and m[2][3] == 0
    R = oracle.face_rotation(oracle.LEFT)
    assert np.array_equal(m[:3, :3].T, R + np.float32(0))


def test_mesher_border_rules(oracle):
    solid = np.zeros(32768, dtype=np.uint8)            # all GRASS
    air = np.full(32768, oracle.AIR, dtype=np.uint8)
    # no neighbours: every border face emitted (chunk.rs:102-108 etc.), y borders always
    q = oracle.cull_unseen_triangles((0, 0, 0), solid, None)
    assert np.bincount(q["face"], minlength=6).tolist() == [1024] * 6
    # solid neighbours hide x/z borders, never y borders (chunk.rs:216-222,234-240)
    q = oracle.cull_unseen_triangles((0, 0, 0), solid, [solid] * 4)
    assert np.bincount(q["face"], minlength=6).tolist() == [0, 0, 0, 0, 1024, 1024]
    q = oracle.cull_unseen_triangles((0, 0, 0), solid, [air, None, solid, air])
    assert np.bincount(q["face"], minlength=6).tolist() == [1024, 1024, 0, 1024, 1024, 1024]
    # neighbour cell actually consulted is the opposite border (x=31 of -x neighbour, ...)
    nb = solid.copy().reshape(32, 32, 32)
    nb[4, 9, 31] = oracle.AIR                                  # y=4, z=9, x=31
    q = oracle.cull_unseen_triangles((0, 0, 0), solid, [nb.reshape(-1), solid, solid, solid])
    left = q[q["face"] == oracle.LEFT]
    assert len(left) == 1 and (left["x"][0], left["y"][0], left["z"][0]) == (0, 4, 9)
    # checkerboard: 16384 cubes * 6 faces
    g = np.indices((32, 32, 32)).sum(axis=0) & 1            # [y][z][x]
    cb = np.where(g == 0, oracle.GRASS, oracle.AIR).astype(np.uint8).reshape(-1)
    assert len(oracle.cull_unseen_triangles((0, 0, 0), cb, None)) == 98304


def test_frustum_golden(oracle, golden):
    for fx in golden["frustum"]:
        m = np.array([int(h, 16) for h in fx["matrix_colmajor"]], dtype=np.uint32).view(np.float32)
        planes = oracle.matrix_to_planes(m)
        gold_planes = np.array([int(h, 16) for h in fx["planes"]], dtype=np.uint32).view(np.float32)
        assert np.array_equal(planes.reshape(-1).view(np.uint32), gold_planes.view(np.uint32))
        lo, hi = fx["grid"]
        g = np.arange(lo, hi, dtype=np.int32)
        cz, cx = np.meshgrid(g, g, indexing="ij")
        coords = np.stack([cx.reshape(-1), np.zeros(cx.size, np.int32), cz.reshape(-1)], axis=1)
        vis = oracle.frustum_cull(planes.reshape(1, 24), coords)[0]
        assert int(vis.sum()) == fx["visible_count"]
        assert "%08x" % zlib.crc32(vis.tobytes()) == fx["vis_crc32"]
        assert 0 < fx["visible_count"] < coords.shape[0]


def test_frustum_any_corner_semantics(oracle):
    # frustum_culling.rs:37: true iff ANY corner is inside ALL planes; NaN distance counts as inside
    planes = np.zeros((6, 4), dtype=np.float32)
    planes[:, 3] = 1.0                                         # all distances = 1 -> inside
    assert oracle.frustum_contains(planes, (5, 0, -3))
    planes[2] = [1, 0, 0, -40.0]                               # x >= 40
    assert oracle.frustum_contains(planes, (1, 0, 0))          # corner x=64 passes
    assert not oracle.frustum_contains(planes, (0, 0, 0))      # max corner x=32 fails
    planes[3] = [-1, 0, 0, 50.0]                               # x <= 50: slab 40..50 has no corner
    assert not oracle.frustum_contains(planes, (1, 0, 0))      # straddling chunk is culled
    planes[4] = [np.nan, 0, 0, 0]
    planes[3] = [0, 0, 0, 1]
    assert oracle.frustum_contains(planes, (1, 0, 0))


def test_batch_driver_matches_single_calls(oracle):
    coords = np.array([(x, 0, z) for z in range(-1, 2) for x in range(-1, 2)], dtype=np.int32)
    res = oracle.gen_mesh_batch(coords, nthreads=3, want_cubes=True)
    cubes = {tuple(c): res["cubes"][i] for i, c in enumerate(coords)}
    for i, c in enumerate(coords):
        assert np.array_equal(res["cubes"][i], oracle.generate_terrain(c))
        adj = [cubes.get(n) for n in _neighbours(tuple(c))]
        want = oracle.create_mesh(c, res["cubes"][i], adj)
        lo, hi = int(res["offsets"][i]), int(res["offsets"][i + 1])
        assert hi - lo == len(want) == res["counts"][i]
        assert res["instances"][lo:hi].tobytes() == want.tobytes()
    # neighbour mask 0 = isolated chunks
    res0 = oracle.gen_mesh_batch(coords, nbr_mask=np.zeros(len(coords), np.uint8), nthreads=2)
    for i, c in enumerate(coords):
        assert res0["counts"][i] == len(oracle.cull_unseen_triangles(c, cubes[tuple(c)], None))


def test_extension_vertical_neighbours_and_dirt_rule(oracle):
    """SURVEY 8(f4) extension semantics (no reference behaviour): vertical neighbours follow the x / z border
    rule of chunk.rs:89-108; a covered solid voxel is DIRT."""
    solid = np.zeros(32768, dtype=np.uint8)
    air = np.full(32768, oracle.AIR, dtype=np.uint8)
    L = oracle.lib()

    def cull6(cubes, adj6):
        a6 = (oracle.C.c_void_p * 6)()
        keep = []
        for k in range(6):
            if adj6[k] is None:
                a6[k] = None
            else:
                keep.append(np.ascontiguousarray(adj6[k], dtype=np.uint8))
                a6[k] = keep[-1].ctypes.data
        out = np.empty(98304, dtype=oracle.INSTANCE_DTYPE)
        c = np.zeros(3, np.int32)
        n = L.vxo_create_mesh_ex(c.ctypes.data, cubes.ctypes.data, a6, out.ctypes.data, 98304)
        return out[:n]
    # all six neighbours solid: nothing is visible; none: the reference's 6 x 1024
    assert len(cull6(solid, [solid] * 6)) == 0
    assert len(cull6(solid, [None] * 6)) == 6 * 1024
    # only the vertical ones solid: x / z borders stay, y borders go
    assert len(cull6(solid, [None] * 4 + [solid, solid])) == 4 * 1024
    # the cell consulted is the facing plane: y = 31 of the chunk below, y = 0 of the chunk above
    below = solid.copy().reshape(32, 32, 32)
    below[31, 9, 4] = oracle.AIR
    inst = cull6(solid, [solid] * 4 + [below.reshape(-1), solid])
    assert len(inst) == 1 and inst["t_matrix"][0][3].tolist() == [4.0, 0.0, 9.0, 1.0] and inst["texture_index"][0] == 1
    # NULL vertical slots == the reference function
    got = cull6(solid, [air, None, solid, air, None, None])
    ref = oracle.create_mesh((0, 0, 0), solid, [air, None, solid, air])
    assert got.tobytes() == ref.tobytes()
    # DIRT rule: covered -> DIRT (1), uncovered -> GRASS (0); the top layer looks into the chunk above
    col = air.copy().reshape(32, 32, 32)
    col[0:3, 5, 5] = oracle.GRASS
    col[31, 7, 7] = oracle.GRASS
    d = oracle.dirt_rule(col.reshape(-1)).reshape(32, 32, 32)
    assert d[0, 5, 5] == oracle.DIRT and d[1, 5, 5] == oracle.DIRT and d[2, 5, 5] == oracle.GRASS
    assert d[31, 7, 7] == oracle.GRASS
    assert oracle.dirt_rule(col.reshape(-1), above=solid).reshape(32, 32, 32)[31, 7, 7] == oracle.DIRT
    assert (oracle.dirt_rule(col.reshape(-1)) == oracle.AIR).sum() == (col == oracle.AIR).sum()
    # batch driver: 2 stacked chunks, vertical on/off
    coords = np.array([(0, 0, 0), (0, 1, 0)], np.int32)
    off = oracle.gen_mesh_batch(coords)
    on = oracle.gen_mesh_batch(coords, vertical=True, dirt_rule=True, want_cubes=True)
    assert on["total"] < off["total"] and (on["cubes"] == oracle.DIRT).any()
    plain = oracle.gen_mesh_batch(coords, vertical=False, dirt_rule=False)
    assert plain["instances"].tobytes() == off["instances"].tobytes()
