"""GPU: parity of the CUDA path (through the C ABI) with the CPU oracle and the
golden vectors.  Bit-exact for blocks, occupancy, quad sequences and frustum
bits; noise within 1e-6 absolute (north_star) -- in fact bit-exact -- with the
near-threshold census reported."""
import zlib

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ctx(vox):
    c = vox.Context()
    yield c
    c.close()


def _neighbours(c):
    cx, cy, cz = (int(v) for v in c)
    return [(cx - 1, cy, cz), (cx + 1, cy, cz), (cx, cy, cz - 1), (cx, cy, cz + 1)]


def _canon(inst):
    """Signed zeros of the rotation products are not pinned: compare with -0 == +0."""
    out = inst.copy()
    out["t_matrix"] = out["t_matrix"] + np.float32(0)
    return out


COORDS = [(0, 0, 0), (1, 0, 0), (-1, 0, -1), (511, 0, -512), (-16, 0, 15), (7, 1, -3), (0, -2, 0),
          (100000, 3, -99999)]


def test_fill_blocks_noise_bit_exact(ctx, oracle):
    res = ctx.fill(COORDS, blocks=True, noise=True, occupancy=True, census=True)
    near = 0
    for i, c in enumerate(COORDS):
        want = oracle.noise_chunk(c)
        got = res["noise"][i]
        assert float(np.max(np.abs(got - want))) <= 1e-6          # north_star tolerance
        assert np.array_equal(got.view(np.uint32), want.view(np.uint32))   # and in fact bit-exact
        cubes = oracle.terrain_from_noise(want)
        assert np.array_equal(res["blocks"][i], cubes)
        bits = np.unpackbits(res["occupancy"][i].view(np.uint8), bitorder="little").reshape(32768)
        assert np.array_equal(bits.astype(bool), cubes != oracle.AIR)
        near += int((np.abs(want) <= np.float32(1e-6)).sum())
    assert res["near_zero"] == near           # tolerance census: counted and reported, not hidden
    print("near-threshold voxels (|noise| <= 1e-6): %d of %d" % (near, 32768 * len(COORDS)))


def test_fill_golden_crc(ctx, golden):
    recs = golden["chunks"]
    res = ctx.fill([r["coord"] for r in recs], blocks=True, noise=True, census=False)
    for i, r in enumerate(recs):
        assert "%08x" % zlib.crc32(res["blocks"][i].tobytes()) == r["occ_crc32"]
        assert ["%08x" % int(v) for v in res["noise"][i][::997].view(np.uint32)] == r["noise_sample_997"]
    # fast kernel variant (no noise output) must give the same blocks
    res2 = ctx.fill([r["coord"] for r in recs], blocks=True)
    assert np.array_equal(res2["blocks"], res["blocks"])
    # full InstanceRaw streams against the numpy restatement (no oracle involved)
    for r in recs:
        m = ctx.gen_mesh([r["coord"]], nbr_mask=[15])
        assert np.bincount(np.zeros(1, int)).size and m.total == sum(r["faces_all"])
        assert "%08x" % zlib.crc32(_canon(m.instances).tobytes()) == r["instances_crc32_all"]


@pytest.mark.parametrize("seed,octaves,freq", [(2, 3, 0.09), (12345, 1, 0.9), (1, 4, 0.05), (7, 2, 0.09)])
def test_fill_other_params(ctx, oracle, vox, seed, octaves, freq):
    p = vox.default_noise_params(seed=seed, octaves=octaves, freq=freq)
    po = oracle.default_params(seed=seed, octaves=octaves, freq=freq)
    cs = [(0, 0, 0), (-5, 0, 9)]
    res = ctx.fill(cs, params=p, noise=True)
    for i, c in enumerate(cs):
        want = oracle.noise_chunk(c, po)
        assert np.array_equal(res["noise"][i].view(np.uint32), want.view(np.uint32))
        assert np.array_equal(res["blocks"][i], oracle.terrain_from_noise(want))


@pytest.mark.parametrize("p_solid,p_dirt", [(0.1, 0.0), (0.5, 0.3), (0.9, 0.5)])
def test_mesh_random_grids_exact_sequence(ctx, oracle, p_solid, p_dirt):
    rng = np.random.default_rng(int(p_solid * 100))
    n = 7
    coords = np.array([(i - 3, rng.integers(-2, 3), 2 * i) for i in range(n)], np.int32)
    solid = rng.random((n, 32768)) < p_solid
    dirt = rng.random((n, 32768)) < p_dirt
    blocks = np.where(solid, np.where(dirt, 1, 0), 2).astype(np.uint8)
    nbr = rng.integers(-1, n, size=(n, 4)).astype(np.int32)        # arbitrary neighbour wiring, -1 = None
    got = ctx.mesh(coords, blocks, nbr_idx=nbr)
    want = oracle.mesh_batch(coords, blocks, nbr)
    assert np.array_equal(got.counts, want["counts"])
    assert np.array_equal(got.offsets, want["offsets"])
    assert got.total == want["total"]
    assert _canon(got.instances).tobytes() == _canon(want["instances"]).tobytes()


def test_mesh_checkerboard_worst_case(ctx, oracle):
    g = np.indices((32, 32, 32)).sum(axis=0) & 1
    cb = np.where(g == 0, 0, 2).astype(np.uint8).reshape(-1)
    coords = np.array([(0, 0, 0), (1, 0, 0)], np.int32)
    # the checkerboard continues across the chunk border (32 is even), second chunk is DIRT
    blocks = np.stack([cb, np.where(g == 0, 1, 2).astype(np.uint8).reshape(-1)])
    got = ctx.mesh(coords, blocks)                       # neighbours resolved by coordinate
    assert got.counts.tolist() == [98304, 98304]
    # opposite parity next door hides the 512 border faces on each side
    anti = np.stack([cb, np.where(g == 1, 1, 2).astype(np.uint8).reshape(-1)])
    assert ctx.mesh(coords, anti, instances=False).counts.tolist() == [97792, 97792]
    want = oracle.mesh_batch(coords, blocks, np.array([[-1, 1, -1, -1], [0, -1, -1, -1]], np.int32))
    assert _canon(got.instances).tobytes() == _canon(want["instances"]).tobytes()


def test_mesh_edge_cases(ctx, vox):
    empty = ctx.mesh(np.zeros((0, 3), np.int32), np.zeros((0, 32768), np.uint8))
    assert empty.total == 0 and empty.offsets.tolist() == [0]
    air = np.full((1, 32768), 2, np.uint8)
    r = ctx.mesh([(0, 0, 0)], air)
    assert r.total == 0 and r.counts.tolist() == [0]
    solid = np.zeros((1, 32768), np.uint8)
    r = ctx.mesh([(0, 0, 0)], solid)
    assert r.total == 6 * 1024
    # mesh_mask = 0: the chunk only serves as a neighbour
    both = np.zeros((2, 32768), np.uint8)
    r = ctx.mesh([(0, 0, 0), (1, 0, 0)], both, mesh_mask=[1, 0])
    assert r.counts.tolist() == [5 * 1024, 0]
    # capacity overflow is an error, counts stay valid
    with pytest.raises(vox.VoxError) as e:
        ctx.mesh([(0, 0, 0)], solid, capacity=100)
    assert e.value.code == vox._lib.VX_ERR_CAPACITY
    # a block id outside {0,1,2} is rejected (cube.rs:16-21)
    bad = solid.copy()
    bad[0, 777] = 9
    with pytest.raises(vox.VoxError) as e:
        ctx.mesh([(0, 0, 0)], bad)
    assert e.value.code == vox._lib.VX_ERR_BAD_BLOCK
    with pytest.raises(vox.VoxError):
        ctx.fill([(0, 0, 0)], params=vox.default_noise_params(octaves=0))


def test_large_downloads_into_pageable_memory_equal_pinned(ctx, vox):
    """Downloads of 1 MB and more into pageable caller memory go through the library's pinned bounce buffers and copy
    threads (4 MB pieces): instance streams, blocks, noise and occupancy equal what lands in pinned buffers, for sizes
    below, at and across the piece boundaries, from fresh and from reused destinations."""
    for n_side in ((6, 6), (9, 9), (20, 13)):
        coords = _grid(-4, 2, *n_side)
        n = len(coords)
        pin = vox.PinnedBuffer(n * 8192 * 68)
        a = ctx.gen_mesh(coords, blocks=True)                               # fresh numpy destinations
        b = ctx.gen_mesh(coords, out=pin.view(np.uint8))
        assert a.total == b.total and a.total * 68 > (1 << 20)
        assert np.array_equal(a.counts, b.counts) and a.instances.tobytes() == b.instances.tobytes()
        reuse = np.empty(a.total * 68 + 4096, np.uint8)
        reuse[:] = 0x5A
        c = ctx.gen_mesh(coords, out=reuse[:a.total * 68])                  # touched destination, exact capacity
        assert c.instances.tobytes() == a.instances.tobytes() and (reuse[a.total * 68:] == 0x5A).all()
        f = ctx.fill(coords, blocks=True, noise=True, occupancy=True)
        assert np.array_equal(f["blocks"], a.blocks)
        small = ctx.fill(coords[:3], blocks=True, noise=True, occupancy=True)      # below the bounce threshold
        assert np.array_equal(f["noise"][:3].view(np.uint32), small["noise"].view(np.uint32))
        assert np.array_equal(f["occupancy"][:3], small["occupancy"]) and np.array_equal(f["blocks"][:3], small["blocks"])


def test_mesh_short_lists_one_sync_path(ctx, vox, oracle):
    """vx_mesh lists of up to 32 chunks (Chunk::create_mesh: the chunk and its four neighbours, chunk.rs:48-57) go up
    in one pinned copy and come back after one synchronisation: into pageable memory through the lane's pinned staging,
    into a pinned caller buffer directly, count-only, with explicit wiring + mesh mask, and past the staging size."""
    rng = np.random.default_rng(77)
    pinned = vox.PinnedBuffer(5 * 98304 * 68)
    for n, p_solid in ((1, 0.5), (5, 0.3), (5, 0.7), (32, 0.05)):
        coords = np.array([(i % 3 - 1, 0, i // 3) for i in range(n)], np.int32)
        blocks = np.where(rng.random((n, 32768)) < p_solid, rng.integers(0, 2, (n, 32768)), 2).astype(np.uint8)
        nbr = rng.integers(-1, n, size=(n, 4)).astype(np.int32)
        mask = (rng.random(n) < 0.6).astype(np.uint8)
        mask[0] = 1
        want = oracle.mesh_batch(coords, blocks, nbr)
        keep = np.concatenate([np.arange(want["offsets"][i], want["offsets"][i + 1]) for i in range(n) if mask[i]]).astype(np.int64)
        a = ctx.mesh(coords, blocks, nbr_idx=nbr, mesh_mask=mask)                          # pageable
        assert a.counts.tolist() == [int(c) if mask[i] else 0 for i, c in enumerate(want["counts"])]
        assert _canon(a.instances).tobytes() == _canon(want["instances"][keep]).tobytes()
        if a.total * 68 <= 5 * 98304 * 68:
            b = ctx.mesh(coords, blocks, nbr_idx=nbr, mesh_mask=mask, out=pinned.view(np.uint8))   # pinned: zero-copy emit
            assert b.counts.tolist() == a.counts.tolist() and b.instances.tobytes() == a.instances.tobytes()
        c = ctx.mesh(coords, blocks, nbr_idx=nbr, mesh_mask=mask, instances=False)          # count only
        assert c.counts.tolist() == a.counts.tolist() and c.total == a.total
        by_coord = ctx.mesh(coords, blocks)                                                 # neighbours by coordinate
        assert by_coord.total > 0
    # the reference's create_mesh shape: the chunk plus four neighbour grids, only the chunk meshed
    c5 = np.array([(0, 0, 0), (-1, 0, 0), (1, 0, 0), (0, 0, -1), (0, 0, 1)], np.int32)
    cubes = np.stack([oracle.generate_terrain(c) for c in c5])
    wiring = np.full((5, 4), -1, np.int32)
    wiring[0] = (1, 2, 3, 4)
    got = ctx.mesh(c5, cubes, nbr_idx=wiring, mesh_mask=[1, 0, 0, 0, 0])
    want = oracle.create_mesh(c5[0], cubes[0], [cubes[1], cubes[2], cubes[3], cubes[4]])
    assert got.counts.tolist() == [len(want), 0, 0, 0, 0]
    assert _canon(got.instances).tobytes() == _canon(want).tobytes()
    # out-of-range wiring is rejected before anything is enqueued; bad block ids are reported
    with pytest.raises(vox.VoxError) as e:
        ctx.mesh(c5, cubes, nbr_idx=np.full((5, 4), 5, np.int32))
    assert e.value.code == vox._lib.VX_ERR_BAD_ARG
    bad = cubes.copy()
    bad[3, 5] = 7
    with pytest.raises(vox.VoxError) as e:
        ctx.mesh(c5, bad, nbr_idx=wiring)
    assert e.value.code == vox._lib.VX_ERR_BAD_BLOCK


def _grid(x0, z0, nx, nz, y=0):
    return np.array([(x, y, z) for z in range(z0, z0 + nz) for x in range(x0, x0 + nx)], np.int32)


def test_gen_mesh_region_in_list_policy(ctx, oracle):
    coords = _grid(-3, -2, 6, 5)
    got = ctx.gen_mesh(coords, blocks=True)
    want = oracle.gen_mesh_batch(coords, want_cubes=True)
    assert np.array_equal(got.blocks, want["cubes"])
    assert np.array_equal(got.counts, want["counts"])
    assert np.array_equal(got.offsets, want["offsets"])
    assert _canon(got.instances).tobytes() == _canon(want["instances"]).tobytes()


def test_gen_mesh_neighbour_masks_and_halo(ctx, oracle):
    # neighbours declared present although they are NOT in the list -> halo slabs from noise;
    # neighbours in the list but masked out -> None (chunk_builder.rs:134-169 snapshots)
    rng = np.random.default_rng(5)
    coords = np.array([(0, 0, 0), (1, 0, 0), (5, 0, 5), (-7, 1, 2), (0, 0, 1)], np.int32)
    masks = rng.integers(0, 16, size=len(coords)).astype(np.uint8)
    masks[2] = 15
    got = ctx.gen_mesh(coords, nbr_mask=masks)
    for i, c in enumerate(coords):
        cubes = oracle.generate_terrain(c)
        adj = [oracle.generate_terrain(nc) if (masks[i] >> k) & 1 else None
               for k, nc in enumerate(_neighbours(c))]
        want = oracle.create_mesh(c, cubes, adj)
        assert got.counts[i] == len(want)
        assert _canon(got.chunk(i)).tobytes() == _canon(want).tobytes()


def test_gen_mesh_sharded_equals_whole(ctx):
    # coordinate sharding: meshing two halves with explicit presence masks
    # reproduces the whole-region result byte for byte
    coords = _grid(-4, -4, 8, 8)
    whole = ctx.gen_mesh(coords)
    inside = {tuple(c) for c in coords.tolist()}
    mask = np.array([sum(1 << k for k, nc in enumerate(_neighbours(c)) if nc in inside) for c in coords], np.uint8)
    a = ctx.gen_mesh(coords[:32], nbr_mask=mask[:32])
    b = ctx.gen_mesh(coords[32:], nbr_mask=mask[32:])
    assert np.array_equal(np.concatenate([a.counts, b.counts]), whole.counts)
    assert a.instances.tobytes() + b.instances.tobytes() == whole.instances.tobytes()
    again = ctx.gen_mesh(coords)
    assert again.instances.tobytes() == whole.instances.tobytes()       # deterministic


def test_gen_mesh_ring_32x32_properties(ctx, oracle):
    # BASELINE config 2 at full size: size-independent properties + sampled oracle checks
    coords = _grid(-16, -16, 32, 32)
    res = ctx.gen_mesh(coords, blocks=True)
    assert res.total == int(res.counts.sum()) == int(res.offsets[-1])
    assert np.array_equal(np.diff(res.offsets.astype(np.int64)), res.counts.astype(np.int64))
    solid = (res.blocks != 2)
    inst = res.instances
    # every instance sits on a solid voxel of its chunk, w row is (0,0,0,1), texture in {0,1,2}
    assert np.all(inst["t_matrix"][:, 3, 3] == 1.0) and np.all(inst["t_matrix"][:, :3, 3] == 0.0)
    assert set(np.unique(inst["texture_index"]).tolist()) <= {0, 1, 2}
    # inside the region, faces between two chunks cancel: #LEFT over all == #RIGHT over all
    col0 = inst["t_matrix"][:, 0, :3]
    left = int(np.sum(col0[:, 2] == 1.0) - 0)     # LEFT column 0 = (e, 0, 1)
    right = int(np.sum(col0[:, 2] == -1.0))       # RIGHT column 0 = (e, 0, -1)
    assert left == right
    for i in (0, 31, 500, 1023):
        c = coords[i]
        cubes = oracle.generate_terrain(c)
        assert np.array_equal(res.blocks[i], cubes)
        inside = {tuple(x) for x in coords.tolist()}
        adj = [oracle.generate_terrain(nc) if nc in inside else None for nc in _neighbours(c)]
        want = oracle.create_mesh(c, cubes, adj)
        assert _canon(res.chunk(i)).tobytes() == _canon(want).tobytes()
    assert solid.any() and not solid.all()


def test_fill_sign_only_shortcuts_are_exact(ctx, vox):
    # the fast kernel skips octaves that cannot flip the sign; the exact kernel (noise output)
    # evaluates everything: blocks must agree on every voxel, for the default terrain at
    # BASELINE size and for parameters that move the octave amplitudes around
    coords = _grid(-16, -16, 32, 32)
    fast = ctx.fill(coords, blocks=True, occupancy=True)
    slow = ctx.fill(coords, blocks=True, noise=True)
    assert np.array_equal(fast["blocks"], slow["blocks"])
    assert np.array_equal(fast["blocks"] == 0, slow["noise"] < 0)
    bits = np.unpackbits(fast["occupancy"].view(np.uint8), axis=1, bitorder="little")
    assert np.array_equal(bits.astype(bool), fast["blocks"] != 2)
    rng = np.random.default_rng(3)
    far = rng.integers(-200000, 200000, size=(48, 3)).astype(np.int32)
    for gain, lac, freq, seed in [(2.5, 0.5, 0.09, 7), (0.5, 2.0, 0.03, 1), (1.0, 0.5, 0.2, 99), (-3.0, 0.7, 0.09, 5),
                                  (40.0, 0.5, 0.09, 1)]:
        p = vox.default_noise_params(gain=gain, lacunarity=lac, freq=freq, seed=seed)
        a = ctx.fill(far, params=p, blocks=True)["blocks"]
        b = ctx.fill(far, params=p, blocks=True, noise=True)
        assert np.array_equal(a, b["blocks"]), (gain, lac, freq, seed)
        assert np.array_equal(a == 0, b["noise"] < 0)
    # non-finite parameters go to the exact kernel (NaN noise is never < 0: all AIR where it reaches the sum)
    for kw in (dict(lacunarity=float("nan")), dict(gain=float("inf")), dict(freq=float("nan")), dict(lacunarity=float("inf"))):
        p = vox.default_noise_params(**kw)
        a = ctx.fill(far[:4], params=p, blocks=True)["blocks"]
        b = ctx.fill(far[:4], params=p, blocks=True, noise=True)
        assert np.array_equal(a, b["blocks"]), kw
        assert np.array_equal(a == 0, b["noise"] < 0)
    # coordinates where the f32 spacing of the noise coordinate reaches the simplex cell size: some CTAs of
    # the lattice-table kernel no longer fit their index box and take the in-kernel exact path
    huge = np.array([(40000000, 0, -40000000), (3000000, 1, 3000000), (20000000, -5, 123), (-1500000, 0, 900000),
                     (6000000, 2, -6000001), (66000000, 0, 66000000)], np.int32)
    a = ctx.fill(huge, blocks=True)["blocks"]
    b = ctx.fill(huge, blocks=True, noise=True)
    assert np.array_equal(a, b["blocks"])
    assert np.array_equal(a == 0, b["noise"] < 0)


def test_reference_interface_mirror(ctx, vox, oracle):
    # reads like the reference's call site chunk_builder.rs:80-84
    pos = vox.ChunkCoord3D(0, 0, 0)
    data = vox.Chunk.new(pos, ctx)
    left = vox.Chunk.new(vox.ChunkCoord3D(-1, 0, 0), ctx)
    front = vox.Chunk.new(vox.ChunkCoord3D(0, 0, 1), ctx)
    mesh = data.create_mesh(ctx, [left, None, None, front])
    cubes = oracle.generate_terrain(pos)
    assert np.array_equal(data.cubes, cubes)
    want = oracle.create_mesh(pos, cubes, [oracle.generate_terrain((-1, 0, 0)), None, None,
                                           oracle.generate_terrain((0, 0, 1))])
    assert mesh.instances_len == len(want)
    assert _canon(mesh.instance_buffer).tobytes() == _canon(want).tobytes()
    chunks, meshes = vox.build_chunks([(0, 0, 0), (-1, 0, 0)], ctx)
    assert np.array_equal(chunks[0].cubes, cubes) and meshes[0].instances_len > 0


def test_frustum_golden_and_oracle(ctx, vox, oracle, golden):
    for fx in golden["frustum"]:
        m = np.array([int(h, 16) for h in fx["matrix_colmajor"]], np.uint32).view(np.float32)
        fr = vox.Frustum.new(m, ctx)
        lo, hi = fx["grid"]
        g = np.arange(lo, hi, dtype=np.int32)
        cz, cx = np.meshgrid(g, g, indexing="ij")
        coords = np.stack([cx.reshape(-1), np.zeros(cx.size, np.int32), cz.reshape(-1)], axis=1)
        vis = fr.contains_many(coords)
        assert int(vis.sum()) == fx["visible_count"]
        assert "%08x" % zlib.crc32(vis.astype(np.uint8).tobytes()) == fx["vis_crc32"]
        assert fr.contains(tuple(coords[0])) == bool(vis[0])


def test_frustum_random_planes_batched(ctx, oracle):
    rng = np.random.default_rng(11)
    F, n = 70, 1000                       # > 64 frames (two frame blocks), ragged n
    planes = rng.standard_normal((F, 6, 4)).astype(np.float32)
    planes[:, :, 3] += 200.0
    planes[3, 2, 0] = np.nan              # NaN distance counts as inside (frustum_culling.rs:33)
    planes[5, :, :] = 0.0
    coords = rng.integers(-40, 40, size=(n, 3)).astype(np.int32)
    coords[0] = (1 << 26, 0, -(1 << 26))  # wrapping world coordinates
    want = oracle.frustum_cull(planes, coords)
    got = ctx.frustum_cull(planes, coords)
    assert np.array_equal(got, want)
    bits = ctx.frustum_cull(planes, coords, bits=True)
    unpacked = np.unpackbits(bits.view(np.uint8), axis=1, bitorder="little")[:, :n]
    assert np.array_equal(unpacked, want)
    assert 0 < want.sum() < want.size


def test_device_resident_api(ctx, vox):
    torch = pytest.importorskip("torch")
    coords = _grid(-2, -2, 4, 4)
    host = ctx.gen_mesh(coords)
    dev = torch.device("cuda:0")
    n = len(coords)
    d_coords = torch.from_numpy(coords).to(dev)
    d_counts = torch.zeros(n, dtype=torch.int32, device=dev)
    d_offsets = torch.zeros(n + 1, dtype=torch.int64, device=dev)
    cap = host.total + 10
    d_inst = torch.zeros(cap * 17, dtype=torch.int32, device=dev)
    d_status = torch.zeros(4, dtype=torch.int32, device=dev)
    tstream = torch.cuda.Stream(device=dev)        # handle 0 would mean "the context's stream"
    stream = tstream.cuda_stream
    tstream.wait_stream(torch.cuda.current_stream())
    ctx.gen_mesh_device(d_coords.data_ptr(), n, d_counts.data_ptr(), d_offsets.data_ptr(), d_inst.data_ptr(), cap,
                        status_ptr=d_status.data_ptr(), stream=stream)
    torch.cuda.synchronize()
    assert d_status[0].item() == 0
    assert np.array_equal(d_counts.cpu().numpy().view(np.uint32), host.counts)
    got = d_inst.cpu().numpy().view(np.uint8)[:host.total * 68]
    assert got.tobytes() == host.instances.tobytes()
    # overflow is flagged, not written
    ctx.gen_mesh_device(d_coords.data_ptr(), n, d_counts.data_ptr(), d_offsets.data_ptr(), d_inst.data_ptr(), 5,
                        status_ptr=d_status.data_ptr(), stream=stream)
    torch.cuda.synchronize()
    assert d_status[0].item() & 1


@pytest.mark.parametrize("slots", [2, 3])
def test_multi_gpu_context_equals_single(ctx, vox, slots):
    """One context over several device slots: the list is split by contiguous index range, each slot is driven by its
    own host thread, neighbours across a cut come from halo slabs (fused call) or ghost chunks (vx_mesh).  On a box
    with fewer GPUs than slots the same GPU is named more than once -- every slot still has its own streams and
    scratch, so the whole multi-device path runs (and is what the driver's 1-GPU box exercises)."""
    have = vox._lib.lib().vx_device_count()
    devices = [i % have for i in range(slots)]
    coords = _grid(-5, -5, 10, 10)
    one = ctx.gen_mesh(coords, blocks=True)
    with vox.Context(devices=devices) as c2:
        assert c2.num_devices == slots
        two = c2.gen_mesh(coords, blocks=True)
        assert np.array_equal(one.counts, two.counts) and np.array_equal(one.offsets, two.offsets)
        assert one.instances.tobytes() == two.instances.tobytes()
        assert np.array_equal(one.blocks, two.blocks)
        f1 = ctx.fill(coords, noise=True)
        f2 = c2.fill(coords, noise=True)
        assert np.array_equal(f1["noise"], f2["noise"])
        # vx_mesh over two GPUs (ghost chunks for the neighbours across the cut): by-coordinate lookup, explicit
        # wiring with a mesh mask, and the vertical-neighbour extension through the fused call
        m1 = ctx.mesh(coords, one.blocks)
        m2 = c2.mesh(coords, one.blocks)
        assert np.array_equal(m1.counts, m2.counts) and m1.instances.tobytes() == m2.instances.tobytes()
        assert m1.instances.tobytes() == one.instances.tobytes()
        rng = np.random.default_rng(2)
        n = len(coords)
        wiring = rng.integers(-1, n, size=(n, 4)).astype(np.int32)
        mask = (rng.random(n) < 0.7).astype(np.uint8)
        blocks = rng.choice(np.array([0, 1, 2], np.uint8), size=(n, 32768), p=[0.4, 0.1, 0.5])
        w1 = ctx.mesh(coords, blocks, nbr_idx=wiring, mesh_mask=mask)
        w2 = c2.mesh(coords, blocks, nbr_idx=wiring, mesh_mask=mask)
        assert np.array_equal(w1.counts, w2.counts) and w1.instances.tobytes() == w2.instances.tobytes()
        st = np.array([(x, y, z) for y in (0, 1) for z in range(6) for x in range(6)], np.int32)
        e1 = ctx.gen_mesh(st, blocks=True, vertical=True, dirt_rule=True)
        e2 = c2.gen_mesh(st, blocks=True, vertical=True, dirt_rule=True)
        assert np.array_equal(e1.blocks, e2.blocks) and e1.instances.tobytes() == e2.instances.tobytes()


def test_cpp_host_mirror_matches_oracle(vox, oracle, tmp_path):
    """The reference's call site written against host/voxb200.hpp (C++), checked against the oracle."""
    import os
    import subprocess
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    host = os.path.join(root, "wgpu-voxel-engine_b200", "host")
    exe = str(tmp_path / "host_demo")
    subprocess.check_call(["g++", "-std=c++17", "-O1", os.path.join(host, "host_demo.cpp"),
                           "-L" + os.path.dirname(vox.SO_PATH), "-lvoxb200",
                           "-Wl,-rpath," + os.path.dirname(vox.SO_PATH), "-o", exe])
    out = dict(line.split(None, 1) for line in subprocess.check_output([exe], text=True).splitlines())

    def fnv(b, h=2166136261):
        for v in b:
            h = ((h ^ v) * 16777619) & 0xFFFFFFFF
        return h
    cubes = oracle.generate_terrain((0, 0, 0))
    assert out["cubes_fnv"] == "%08x" % fnv(cubes.tobytes())
    want = _canon(oracle.create_mesh((0, 0, 0), cubes, [oracle.generate_terrain((-1, 0, 0)), None, None,
                                                       oracle.generate_terrain((0, 0, 1))]))
    assert int(out["instances_len"]) == len(want)
    assert out["instances_fnv"] == "%08x" % fnv(want.tobytes())
    assert out["frustum"].split() == ["1", "0"]


def test_packed_quads_expand_to_the_instance_stream(ctx, oracle):
    """Extension: 4-byte quad records carry the same quads in the same order."""
    coords = _grid(-2, -2, 4, 4, y=1)
    masks = np.random.default_rng(9).integers(0, 16, size=len(coords)).astype(np.uint8)
    full = ctx.gen_mesh(coords, nbr_mask=masks)
    pk = ctx.gen_mesh_packed(coords, nbr_mask=masks)
    assert np.array_equal(pk.counts, full.counts) and np.array_equal(pk.offsets, full.offsets)
    d = pk.instances.astype(np.int64)
    chunk = np.repeat(np.arange(len(coords)), full.counts.astype(np.int64))
    x, z, y, face, typ = d & 31, (d >> 5) & 31, (d >> 10) & 31, (d >> 15) & 7, (d >> 18) & 1
    assert face.max() <= 5 and typ.max() == 0
    want = np.zeros(len(d), dtype=full.instances.dtype)
    rot = np.stack([oracle.face_rotation(f) for f in range(6)]) + np.float32(0)       # [face][row][col]
    want["t_matrix"][:, :3, :3] = np.transpose(rot[face], (0, 2, 1))                  # t_matrix[c][r]
    world = coords[chunk].astype(np.int64) * 32
    want["t_matrix"][:, 3, 0] = world[:, 0] + x
    want["t_matrix"][:, 3, 1] = world[:, 1] + y
    want["t_matrix"][:, 3, 2] = world[:, 2] + z
    want["t_matrix"][:, 3, 3] = 1.0
    tex = np.array([[0, 0, 0, 0, 1, 2], [1, 1, 1, 1, 1, 1]], np.uint32)             # [type][face]: slots 0,1,2,3,5,4
    want["texture_index"] = tex[typ, face]
    assert _canon(full.instances).tobytes() == want.tobytes()


def test_large_list_uses_the_global_hash_path(ctx, oracle):
    """> 4096 chunks: neighbour resolution runs as vx_ht_insert + vx_resolve over a global-memory table
    instead of the single-CTA shared-memory kernel.  Same answers as the small path and the oracle."""
    coords = _grid(-36, -36, 72, 72)                      # 5184 chunks
    big = ctx.gen_mesh(coords, instances=False)
    assert big.total == int(big.counts.sum())
    # the same region in two calls (small path each? no: 2592 each -> small path) with explicit masks
    inside = np.zeros((74, 74), bool)
    inside[1:-1, 1:-1] = True
    ix, iz = coords[:, 0] + 37, coords[:, 2] + 37
    mask = (inside[iz, ix - 1] * 1 + inside[iz, ix + 1] * 2 + inside[iz - 1, ix] * 4 + inside[iz + 1, ix] * 8).astype(np.uint8)
    h = len(coords) // 2
    a = ctx.gen_mesh(coords[:h], nbr_mask=mask[:h], instances=False)
    b = ctx.gen_mesh(coords[h:], nbr_mask=mask[h:], instances=False)
    assert np.array_equal(np.concatenate([a.counts, b.counts]), big.counts)
    # and the large path with masks + halo slabs (present neighbours outside the list)
    full_mask = np.full(len(coords), 15, np.uint8)
    c = ctx.gen_mesh(coords, nbr_mask=full_mask, instances=False)
    for i in (0, 71, 2600, 5183):
        cubes = oracle.generate_terrain(coords[i])
        adj_in = [oracle.generate_terrain(n) if (mask[i] >> k) & 1 else None for k, n in enumerate(_neighbours(coords[i]))]
        assert big.counts[i] == len(oracle.cull_unseen_triangles(coords[i], cubes, adj_in))
        adj_all = [oracle.generate_terrain(n) for n in _neighbours(coords[i])]
        assert c.counts[i] == len(oracle.cull_unseen_triangles(coords[i], cubes, adj_all))
    # duplicates in the list are legal: every copy gets the same mesh
    dup = np.concatenate([coords[:3], coords[:3]])
    d = ctx.gen_mesh(dup)
    assert np.array_equal(d.counts[:3], d.counts[3:]) and d.chunk(0).tobytes() == d.chunk(3).tobytes()


def test_device_entry_points_match_host_entry_points(ctx, vox):
    torch = pytest.importorskip("torch")
    dev = torch.device("cuda:0")
    ts = torch.cuda.Stream(device=dev)
    st = ts.cuda_stream
    coords = _grid(-2, 3, 3, 3, y=-1)
    n = len(coords)
    d_coords = torch.from_numpy(coords).to(dev)
    # vx_fill_device
    host = ctx.fill(coords, blocks=True, noise=True, occupancy=True, census=True)
    d_blocks = torch.zeros(n * 32768, dtype=torch.uint8, device=dev)
    d_noise = torch.zeros(n * 32768, dtype=torch.float32, device=dev)
    d_occ = torch.zeros(n * 1024, dtype=torch.int32, device=dev)
    d_nz = torch.zeros(1, dtype=torch.int64, device=dev)
    ctx.fill_device(d_coords.data_ptr(), n, blocks_ptr=d_blocks.data_ptr(), noise_ptr=d_noise.data_ptr(),
                    occ_ptr=d_occ.data_ptr(), near_zero_ptr=d_nz.data_ptr(), stream=st)
    torch.cuda.synchronize()
    assert np.array_equal(d_blocks.cpu().numpy().reshape(n, -1), host["blocks"])
    assert np.array_equal(d_noise.cpu().numpy().view(np.uint32).reshape(n, -1), host["noise"].view(np.uint32))
    assert np.array_equal(d_occ.cpu().numpy().view(np.uint32).reshape(n, -1), host["occupancy"])
    assert int(d_nz.item()) == host["near_zero"]
    # fast (sign-only) device fill gives the same blocks
    d_blocks2 = torch.zeros_like(d_blocks)
    ctx.fill_device(d_coords.data_ptr(), n, blocks_ptr=d_blocks2.data_ptr(), stream=st)
    torch.cuda.synchronize()
    assert torch.equal(d_blocks, d_blocks2)
    # vx_mesh_device
    hm = ctx.mesh(coords, host["blocks"])
    d_counts = torch.zeros(n, dtype=torch.int32, device=dev)
    d_offsets = torch.zeros(n + 1, dtype=torch.int64, device=dev)
    d_inst = torch.zeros((hm.total + 3) * 17, dtype=torch.int32, device=dev)
    d_status = torch.zeros(4, dtype=torch.int32, device=dev)
    ctx.mesh_device(d_coords.data_ptr(), n, d_blocks.data_ptr(), d_counts.data_ptr(), d_offsets.data_ptr(),
                    d_inst.data_ptr(), hm.total + 3, status_ptr=d_status.data_ptr(), stream=st)
    torch.cuda.synchronize()
    assert int(d_status[0].item()) == 0
    assert d_inst.cpu().numpy().view(np.uint8)[:hm.total * 68].tobytes() == hm.instances.tobytes()
    # vx_frustum_cull_device, bytes and bits in one launch
    planes = np.random.default_rng(2).standard_normal((5, 6, 4)).astype(np.float32)
    planes[:, :, 3] += 150
    hv = ctx.frustum_cull(planes, coords)
    d_pl = torch.from_numpy(planes).to(dev)
    d_vis = torch.zeros(5 * n, dtype=torch.uint8, device=dev)
    d_bits = torch.zeros(5 * ((n + 31) // 32), dtype=torch.int32, device=dev)
    ctx.frustum_cull_device(d_pl.data_ptr(), 5, d_coords.data_ptr(), n, visible_ptr=d_vis.data_ptr(),
                            bits_ptr=d_bits.data_ptr(), stream=st)
    torch.cuda.synchronize()
    assert np.array_equal(d_vis.cpu().numpy().reshape(5, n), hv)
    bits = np.unpackbits(d_bits.cpu().numpy().view(np.uint8).reshape(5, -1), axis=1, bitorder="little")[:, :n]
    assert np.array_equal(bits, hv)


def test_host_streaming_reuses_arenas_correctly(ctx, vox):
    """~14 M quads = 4+ slices through the two 256 MiB instance arenas: every arena is reused while
    the copy stream is still draining the other one.  Must equal the single-launch device path."""
    torch = pytest.importorskip("torch")
    coords = _grid(-32, -24, 64, 48)                       # 3072 chunks
    host = ctx.gen_mesh(coords)
    assert host.total * 68 > 3 * (256 << 20)
    dev = torch.device("cuda:0")
    ts = torch.cuda.Stream(device=dev)
    n = len(coords)
    d_coords = torch.from_numpy(coords).to(dev)
    d_counts = torch.zeros(n, dtype=torch.int32, device=dev)
    d_offsets = torch.zeros(n + 1, dtype=torch.int64, device=dev)
    d_inst = torch.zeros(host.total * 17, dtype=torch.int32, device=dev)
    d_status = torch.zeros(4, dtype=torch.int32, device=dev)
    ctx.gen_mesh_device(d_coords.data_ptr(), n, d_counts.data_ptr(), d_offsets.data_ptr(), d_inst.data_ptr(),
                        host.total, status_ptr=d_status.data_ptr(), stream=ts.cuda_stream)
    torch.cuda.synchronize()
    assert int(d_status[0].item()) == 0
    assert np.array_equal(d_offsets.cpu().numpy().view(np.uint64), host.offsets)
    got = d_inst.cpu().numpy().view(np.uint8)
    want = host.instances.view(np.uint8).reshape(-1)
    assert got.shape == want.shape and np.array_equal(got, want)
    # boundary sizes around the shared-memory / global-memory resolve switch (4096)
    for m in (4095, 4096, 4097):
        c = _grid(-40, -30, 80, 60)[:m]
        a = ctx.gen_mesh(c, instances=False)
        b = ctx.gen_mesh(c[::-1].copy(), instances=False)          # order must not matter for the counts
        assert np.array_equal(a.counts, b.counts[::-1])


def test_host_api_output_slicing(ctx, oracle):
    """Outputs larger than the 256 MiB staging buffers go through the slice loops of vx_fill and
    vx_frustum_cull; sampled against the oracle."""
    # vx_fill: 2200 chunks x (128 KiB noise + 32 KiB blocks + 4 KiB occupancy) = 360 MB -> 2 slices
    coords = _grid(-25, -22, 50, 44)
    res = ctx.fill(coords, blocks=True, noise=True, occupancy=True)
    for i in (0, 1637, 1638, 1639, 2199):
        want = oracle.noise_chunk(coords[i])
        assert np.array_equal(res["noise"][i].view(np.uint32), want.view(np.uint32))
        assert np.array_equal(res["blocks"][i], oracle.terrain_from_noise(want))
        bits = np.unpackbits(res["occupancy"][i].view(np.uint8), bitorder="little")
        assert np.array_equal(bits.astype(bool), res["blocks"][i] != 2)
    # vx_frustum_cull: 300 000 chunks x 1000 frames of bytes = 300 MB -> 2 frame slices
    rng = np.random.default_rng(8)
    g = np.arange(-274, 274, dtype=np.int32)
    z, x = np.meshgrid(g, g, indexing="ij")
    c = np.stack([x.reshape(-1), np.zeros(x.size, np.int32), z.reshape(-1)], 1).astype(np.int32)[:300000]
    planes = rng.standard_normal((1000, 6, 4)).astype(np.float32)
    planes[:, :, 3] += 3000.0
    vis = ctx.frustum_cull(planes, c)
    assert vis.shape == (1000, 300000)
    for f in (0, 893, 894, 895, 999):
        assert np.array_equal(vis[f], oracle.frustum_cull(planes[f:f + 1], c)[0])


def test_concurrent_callers_like_the_uvth_pool(ctx, vox):
    """chunk_builder.rs:80-84 runs Chunk::new / create_mesh on pool threads.  Calls on one context are
    serialised by the library, separate contexts run side by side; both must give the serial answer."""
    import threading
    jobs = [_grid(x0, z0, 3, 3) for x0 in (-9, 0, 9) for z0 in (-9, 0, 9)]
    serial = [ctx.gen_mesh(j).instances.tobytes() for j in jobs]
    planes = np.random.default_rng(1).standard_normal((3, 6, 4)).astype(np.float32)
    serial_vis = [ctx.frustum_cull(planes, j).tobytes() for j in jobs]
    results, errors = {}, []

    def worker(tid, context, own):
        try:
            c = vox.Context() if own else context
            for rep in range(3):
                for k, j in enumerate(jobs):
                    if (k + tid) % 2 == 0:
                        results[(tid, rep, k, "m")] = c.gen_mesh(j).instances.tobytes()
                    else:
                        results[(tid, rep, k, "f")] = c.frustum_cull(planes, j).tobytes()
                    ch = vox.Chunk.new((k, 0, tid), c)                  # single-chunk calls in between
                    assert ch.cubes.shape == (32768,)
            if own:
                c.close()
        except Exception as e:                                          # pragma: no cover
            errors.append(repr(e))
    threads = [threading.Thread(target=worker, args=(t, ctx, t >= 4)) for t in range(8)]
    for t in threads:
        t.start()
    for t in threads:
        t.join()
    assert not errors, errors
    assert len(results) == 8 * 3 * len(jobs)
    for (tid, rep, k, kind), b in results.items():
        assert b == (serial[k] if kind == "m" else serial_vis[k]), (tid, rep, k, kind)


def test_two_batches_in_flight_on_two_contexts(ctx, vox):
    """bench.py and INTEGRATION.md 3 keep two batches in flight: device-resident calls alternate between two
    contexts, each on its own stream with its own outputs, and nothing synchronises in between.  Every batch
    must equal the result of the same batch run alone."""
    import torch
    dev = torch.device("cuda", 0)
    n, cap = 256, 256 * 8192
    batches = [_grid(16 * (i % 3) - 16, 16 * (i // 3) - 8, 16, 16) for i in range(6)]
    alone = [ctx.gen_mesh(b) for b in batches]
    ctxs = [vox.Context(devices=[0]) for _ in range(2)]
    streams = [torch.cuda.Stream(device=dev) for _ in range(2)]
    d_coords = [torch.from_numpy(b).to(dev) for b in batches]
    outs = [dict(c=torch.zeros(n, dtype=torch.int32, device=dev), o=torch.zeros(n + 1, dtype=torch.int64, device=dev),
                 i=torch.empty(cap * 17, dtype=torch.int32, device=dev), s=torch.zeros(4, dtype=torch.int32, device=dev))
            for _ in batches]                                            # one output set per batch: read after the flush
    torch.cuda.synchronize()
    for rep in range(2):
        for k in range(len(batches)):
            j, o = k % 2, outs[k]
            ctxs[j].gen_mesh_device(d_coords[k].data_ptr(), n, o["c"].data_ptr(), o["o"].data_ptr(), o["i"].data_ptr(), cap,
                                    status_ptr=o["s"].data_ptr(), stream=streams[j].cuda_stream)
    torch.cuda.synchronize()
    for k, o in enumerate(outs):
        assert int(o["s"][0].item()) == 0
        assert np.array_equal(o["c"].cpu().numpy().astype(np.uint32), alone[k].counts)
        total = int(o["o"][n].item())
        assert total == alone[k].total
        got = o["i"][:total * 17].cpu().numpy().view(np.uint8)
        assert got.tobytes() == alone[k].instances.tobytes()
    for c in ctxs:
        c.close()


def test_fast_fill_paths_agree_with_the_exact_kernel_on_random_parameters(ctx, vox):
    """Both fast kernels (lattice tables for |freq * lac^k| <= 0.1, per-voxel hashing above) and the lattice
    kernel's per-CTA exact path, against the exact kernel, over random seeds / frequencies / lacunarities /
    gains (either sign) and coordinates from the origin out to where f32 loses the cell."""
    rng = np.random.default_rng(2024)
    scales = [8, 300, 20000, 1500000, 50000000]
    for case in range(24):
        freq = float(np.exp(rng.uniform(np.log(0.004), np.log(0.4)))) * (1 if rng.random() < 0.8 else -1)
        lac = float(rng.uniform(0.25, 2.2)) * (1 if rng.random() < 0.8 else -1)
        gain = float(rng.uniform(0.4, 4.0)) * (1 if rng.random() < 0.8 else -1)
        p = vox.default_noise_params(seed=int(rng.integers(0, 2 ** 31 - 1)), freq=freq, lacunarity=lac, gain=gain)
        s = scales[case % len(scales)]
        coords = np.stack([rng.integers(-s, s + 1, 6), rng.integers(-3, 4, 6), rng.integers(-s, s + 1, 6)], 1).astype(np.int32)
        fast = ctx.fill(coords, params=p, blocks=True, occupancy=True)
        slow = ctx.fill(coords, params=p, blocks=True, noise=True)
        assert np.array_equal(fast["blocks"], slow["blocks"]), (case, freq, lac, gain, s)
        assert np.array_equal(fast["blocks"] == 0, slow["noise"] < 0)
        bits = np.unpackbits(fast["occupancy"].view(np.uint8), axis=1, bitorder="little")
        assert np.array_equal(bits.astype(bool), fast["blocks"] != 2)


def test_short_list_path_pinned_and_pageable_outputs(ctx, vox, oracle):
    """Lists of up to 32 chunks take the one-synchronisation path of vx_gen_mesh: into a pinned caller buffer
    the emit kernel writes directly over PCIe, into pageable memory through the arena -- same bytes either
    way, equal to the oracle, for 1..32 chunks, with and without block output, also in packed form."""
    pinned = vox.PinnedBuffer(32 * 98304 * 68)
    for n in (1, 2, 7, 32):
        coords = _grid(-3, 5, n, 1) if n < 32 else _grid(-4, -4, 8, 4)
        masks = np.random.default_rng(n).integers(0, 16, size=n).astype(np.uint8)
        cubes = [oracle.generate_terrain(c) for c in coords]
        want = [oracle.create_mesh(c, cubes[i], [oracle.generate_terrain(nc) if (masks[i] >> k) & 1 else None
                                                 for k, nc in enumerate(_neighbours(c))]) for i, c in enumerate(coords)]
        want_all = np.concatenate(want)
        a = ctx.gen_mesh(coords, nbr_mask=masks, blocks=True)                          # pageable numpy output
        b = ctx.gen_mesh(coords, nbr_mask=masks, out=pinned.view(np.uint8))            # pinned: zero-copy emit
        for r in (a, b):
            assert r.counts.tolist() == [len(w) for w in want]
            assert r.offsets.tolist() == np.concatenate([[0], np.cumsum([len(w) for w in want])]).tolist()
            assert _canon(r.instances).tobytes() == _canon(want_all).tobytes()
        assert np.array_equal(a.blocks, np.stack(cubes))
        pk = ctx.gen_mesh_packed(coords, nbr_mask=masks)
        pk2 = ctx.gen_mesh_packed(coords, nbr_mask=masks, out=pinned.view(np.uint32))
        assert pk.total == a.total and np.array_equal(pk.instances, pk2.instances)
    # a pageable list whose instances do not fit the lane's pinned staging buffer (caves: ~50 k quads per chunk, 3 chunks
    # = 10 MB > 8 MB) falls back to the arena: same bytes as the zero-copy form
    caves = vox.default_noise_params(freq=0.9, octaves=1)
    c3 = _grid(2, -1, 3, 1)
    pa = ctx.gen_mesh(c3, params=caves, nbr_mask=np.full(3, 15, np.uint8))
    pb = ctx.gen_mesh(c3, params=caves, nbr_mask=np.full(3, 15, np.uint8), out=pinned.view(np.uint8))
    assert pa.total == pb.total and pa.total * 68 > (8 << 20)
    assert pa.counts.tolist() == pb.counts.tolist() and pa.instances.tobytes() == pb.instances.tobytes()
    # pageable destination one quad short: error, nothing written
    one = ctx.gen_mesh(_grid(0, 0, 1, 1))
    short = np.full((one.total - 1) * 68 + 64, 0xCD, np.uint8)
    with pytest.raises(vox.VoxError) as e:
        ctx.gen_mesh(_grid(0, 0, 1, 1), out=short[:(one.total - 1) * 68])
    assert e.value.code == vox._lib.VX_ERR_CAPACITY and (short == 0xCD).all()
    # capacity too small: error, counts still valid, nothing written past the buffer
    small = vox.PinnedBuffer(100 * 68)
    guard = small.view(np.uint8)
    guard[:] = 0xAB
    with pytest.raises(vox.VoxError) as e:
        ctx.gen_mesh(_grid(0, 0, 2, 1), out=guard)
    assert e.value.code == vox._lib.VX_ERR_CAPACITY
    assert (guard == 0xAB).all()


@pytest.mark.parametrize("n_side,masked", [((32, 32), False), ((37, 21), True), ((16, 32), True)])
def test_large_device_batches_equal_the_host_api(ctx, vox, n_side, masked):
    """vx_gen_mesh_device on batches of 512+ chunks (the bench's call): counts, offsets, blocks and instance bytes
    equal the host-API result, repeated calls reuse the scratch, and a capacity one quad short only raises the
    overflow flag (counts and offsets stay complete)."""
    import torch
    dev = torch.device("cuda", 0)
    coords = _grid(-7, 3, n_side[0], n_side[1])
    n = len(coords)
    assert n >= 512
    rng = np.random.default_rng(n)
    masks = rng.integers(0, 16, size=n).astype(np.uint8) if masked else None
    want = ctx.gen_mesh(coords, nbr_mask=masks, blocks=True)                   # host API: one piece
    st = torch.cuda.Stream(device=dev)
    d_c = torch.from_numpy(coords).to(dev)
    d_m = torch.from_numpy(masks).to(dev) if masked else None
    cap = want.total + 5
    d_counts = torch.zeros(n, dtype=torch.int32, device=dev)
    d_off = torch.zeros(n + 1, dtype=torch.int64, device=dev)
    d_inst = torch.zeros(cap * 17, dtype=torch.int32, device=dev)
    d_blocks = torch.zeros(n * 32768, dtype=torch.uint8, device=dev)
    d_status = torch.zeros(4, dtype=torch.int32, device=dev)
    torch.cuda.synchronize()
    for rep in range(2):
        ctx.gen_mesh_device(d_c.data_ptr(), n, d_counts.data_ptr(), d_off.data_ptr(), d_inst.data_ptr(), cap,
                            nbr_mask_ptr=d_m.data_ptr() if masked else None, status_ptr=d_status.data_ptr(),
                            blocks_ptr=d_blocks.data_ptr(), stream=st.cuda_stream)
        st.synchronize()
        assert int(d_status[0].item()) == 0
        assert np.array_equal(d_counts.cpu().numpy().astype(np.uint32), want.counts)
        assert np.array_equal(d_off.cpu().numpy().astype(np.uint64), want.offsets)
        assert np.array_equal(d_blocks.cpu().numpy().reshape(n, 32768), want.blocks)
        assert d_inst[:want.total * 17].cpu().numpy().view(np.uint8).tobytes() == want.instances.tobytes()
        d_inst.zero_()
    # capacity one quad short: overflow flagged, counts and offsets still complete
    ctx.gen_mesh_device(d_c.data_ptr(), n, d_counts.data_ptr(), d_off.data_ptr(), d_inst.data_ptr(), want.total - 1,
                        nbr_mask_ptr=d_m.data_ptr() if masked else None, status_ptr=d_status.data_ptr(), stream=st.cuda_stream)
    st.synchronize()
    assert int(d_status[0].item()) & 1
    assert np.array_equal(d_off.cpu().numpy().astype(np.uint64), want.offsets)


# ---- round 2: full-size streams, the 1024 x 1024 region, caves through the fused call, both noise scales -----
def test_gen_mesh_ring_32x32_full_stream_equals_oracle(ctx, oracle):
    """configs[1] at full size, everything compared: 1024 chunks, ~4.7 M instances = 321 MB, byte for byte."""
    coords = _grid(-16, -16, 32, 32)
    got = ctx.gen_mesh(coords, blocks=True)
    want = oracle.gen_mesh_batch(coords, want_cubes=True)
    assert np.array_equal(got.blocks, want["cubes"])
    assert np.array_equal(got.counts, want["counts"]) and np.array_equal(got.offsets, want["offsets"])
    assert got.total == want["total"] > 4_000_000
    assert _canon(got.instances).tobytes() == _canon(want["instances"]).tobytes()


def test_region_1024_samples_equal_oracle(ctx, oracle):
    """configs[2]: chunks sampled from the 1024 x 1024 region -- the four corners (+-512), both rows of every
    band border of the 8-GPU sharding, 200 random ones -- meshed with the region's presence masks (neighbours
    outside the list are halo slabs) equal the oracle in counts and instance bytes."""
    import bench
    R, h = 1024, 512
    idx = bench.sample_indices(R, 8, extra=200, seed=3)
    assert len(idx) >= 256
    sc = np.stack([(idx % R) - h, np.zeros(len(idx), np.int64), (idx // R) - h], axis=1).astype(np.int32)
    sm = ((sc[:, 0] > -h) * 1 + (sc[:, 0] < h - 1) * 2 + (sc[:, 2] > -h) * 4 + (sc[:, 2] < h - 1) * 8).astype(np.uint8)
    assert {0b1010, 0b1001, 0b0110, 0b0101} <= set(sm.tolist())          # the four corner masks occur
    o_counts, o_off, o_inst = bench.oracle_with_halo(sc, sm)
    got = ctx.gen_mesh(sc, nbr_mask=sm)
    assert np.array_equal(got.counts, o_counts) and np.array_equal(got.offsets, o_off)
    assert _canon(got.instances).tobytes() == _canon(o_inst).tobytes()
    # and the same rows inside a full 8192-chunk batch of the region (the bench's call shape): band border row
    z0 = bench.band_rows(3, 8, R)[0]
    c, m = bench.region_rows(z0 - 4, z0 + 4, R)
    big = ctx.gen_mesh(c, nbr_mask=m, instances=False)
    pick = np.arange(0, len(c), 61)
    pc, _, _ = bench.oracle_with_halo(c[pick], m[pick])
    assert np.array_equal(big.counts[pick], pc)


def test_caves_through_gen_mesh(ctx, oracle, vox):
    """configs[4] 'caves' (freq 0.9, one octave: ~50 % solid, ~10x the terrain's quads) through the FUSED call,
    not only through vx_fill: blocks, counts and the instance stream equal the oracle."""
    p = vox.default_noise_params(freq=0.9, octaves=1)
    po = oracle.default_params(freq=0.9, octaves=1)
    coords = _grid(-2, 3, 4, 4)
    got = ctx.gen_mesh(coords, params=p, blocks=True)
    want = oracle.gen_mesh_batch(coords, params=po, want_cubes=True)
    assert np.array_equal(got.blocks, want["cubes"])
    assert np.array_equal(got.counts, want["counts"])
    assert got.total > 16 * 30000                                        # dense: stresses compaction
    assert _canon(got.instances).tobytes() == _canon(want["instances"]).tobytes()
    # with halo neighbours (present but outside the list)
    m = np.full(len(coords), 15, np.uint8)
    got2 = ctx.gen_mesh(coords[:3], params=p, nbr_mask=m[:3])
    import bench
    oc, oo, oi = _oracle_halo_params(oracle, coords[:3], m[:3], po)
    assert np.array_equal(got2.counts, oc)
    assert _canon(got2.instances).tobytes() == _canon(oi).tobytes()


def _oracle_halo_params(oracle, coords, masks, params):
    out_c, out_i = [], []
    for c, m in zip(coords, masks):
        cubes = oracle.generate_terrain(c, params)
        adj = [oracle.generate_terrain(nc, params) if (m >> k) & 1 else None for k, nc in enumerate(_neighbours(c))]
        inst = oracle.create_mesh(c, cubes, adj)
        out_c.append(len(inst))
        out_i.append(inst)
    return np.array(out_c, np.uint32), None, np.concatenate(out_i)


@pytest.mark.parametrize("scale", [1.0, 32.69587493801679, -7.5])
def test_simplex_scale_variants(ctx, oracle, vox, scale):
    """vx_noise_params.simplex_scale: the raw simplex sum (1.0), simdnoise's normalised form (default) and an
    arbitrary factor -- noise bit-exact, fast (sign-only) blocks equal exact blocks, meshes equal the oracle."""
    p = vox.default_noise_params(simplex_scale=scale)
    po = oracle.default_params(simplex_scale=scale)
    cs = [(0, 0, 0), (-5, 0, 9), (511, 0, -512), (100000, 3, -99999)]
    res = ctx.fill(cs, params=p, noise=True)
    for i, c in enumerate(cs):
        want = oracle.noise_chunk(c, po)
        assert np.array_equal(res["noise"][i].view(np.uint32), want.view(np.uint32))
        assert np.array_equal(res["blocks"][i], oracle.terrain_from_noise(want))
    coords = _grid(-8, -8, 16, 16)
    fast = ctx.fill(coords, params=p, blocks=True)["blocks"]
    slow = ctx.fill(coords, params=p, blocks=True, noise=True)
    assert np.array_equal(fast, slow["blocks"]) and np.array_equal(fast == 0, slow["noise"] < 0)
    got = ctx.gen_mesh(coords[:40], params=p)
    want = oracle.gen_mesh_batch(coords[:40], params=po)
    assert np.array_equal(got.counts, want["counts"])
    assert _canon(got.instances).tobytes() == _canon(want["instances"]).tobytes()
    if scale > 0:
        # a positive scale leaves almost every block alone; the few near-zero sums it re-rounds may flip
        base = ctx.fill(coords, blocks=True)["blocks"]
        assert (base != fast).mean() < 1e-4
    zero = ctx.fill(cs[:1], params=vox.default_noise_params(simplex_scale=0.0), noise=True)      # 0 = default
    assert np.array_equal(zero["noise"][0], ctx.fill(cs[:1], noise=True)["noise"][0])


def test_emit_again_device(ctx, vox):
    torch = pytest.importorskip("torch")
    dev = torch.device("cuda", 0)
    coords = _grid(-3, 5, 12, 9)
    n = len(coords)
    host = ctx.gen_mesh(coords)
    d_coords = torch.from_numpy(coords).to(dev)
    d_counts = torch.zeros(n, dtype=torch.int32, device=dev)
    d_off = torch.zeros(n + 1, dtype=torch.int64, device=dev)
    cap = host.total + 5
    a = torch.zeros(cap * 17, dtype=torch.int32, device=dev)
    b = torch.zeros(cap * 17, dtype=torch.int32, device=dev)
    st = torch.zeros(1, dtype=torch.int32, device=dev)
    s = torch.cuda.Stream(device=dev)
    with vox.Context() as c2:
        c2.gen_mesh_device(d_coords.data_ptr(), n, d_counts.data_ptr(), d_off.data_ptr(), a.data_ptr(), cap,
                           status_ptr=st.data_ptr(), stream=s.cuda_stream)
        c2.emit_again_device(d_coords.data_ptr(), n, d_counts.data_ptr(), d_off.data_ptr(), b.data_ptr(), cap,
                             status_ptr=st.data_ptr(), stream=s.cuda_stream)
        s.synchronize()
        assert int(st.item()) == 0 and torch.equal(a, b)
        assert a[:host.total * 17].cpu().numpy().tobytes() == host.instances.tobytes()
        with pytest.raises(vox.VoxError):
            c2.emit_again_device(d_coords.data_ptr(), n - 1, d_counts.data_ptr(), d_off.data_ptr(), b.data_ptr(), cap)


# ---- SURVEY 8(f4) extension: vertical neighbours and the DIRT rule ------------------------------------------
def _stack(x0, z0, nx, nz, ys):
    return np.array([(x, y, z) for y in ys for z in range(z0, z0 + nz) for x in range(x0, x0 + nx)], np.int32)


@pytest.mark.parametrize("vertical,dirt", [(True, False), (False, True), (True, True)])
def test_vertical_neighbours_and_dirt_rule(ctx, oracle, vertical, dirt):
    """vx_gen_mesh_ex: in-list vertical neighbours, masked ones, halo ones (present but not in the list), DIRT under
    cover incl. across the chunk border -- blocks, counts and instance bytes equal the oracle's extension."""
    coords = _stack(-1, 2, 3, 2, (-1, 0, 1))                        # 3 layers of 3 x 2 chunks
    got = ctx.gen_mesh(coords, blocks=True, vertical=vertical, dirt_rule=dirt)
    want = oracle.gen_mesh_batch(coords, want_cubes=True, vertical=vertical, dirt_rule=dirt)
    assert np.array_equal(got.blocks, want["cubes"])
    assert np.array_equal(got.counts, want["counts"])
    assert _canon(got.instances).tobytes() == _canon(want["instances"]).tobytes()
    if dirt:
        assert (got.blocks == 1).any() and set(np.unique(got.instances["texture_index"]).tolist()) == {0, 1, 2}
    if vertical:
        ref = ctx.gen_mesh(coords)                                   # reference rule: y borders always emitted
        assert got.total < ref.total
    # explicit masks: random 6-bit masks; present neighbours that are not in the list become halo slabs
    rng = np.random.default_rng(11)
    sub = coords[[0, 4, 7, 9, 13, 17]]
    masks = rng.integers(0, 64, size=len(sub)).astype(np.uint8)
    masks[1] = 63
    got2 = ctx.gen_mesh(sub, nbr_mask=masks, blocks=True, vertical=vertical, dirt_rule=dirt)
    for i, c in enumerate(sub):
        cx, cy, cz = (int(v) for v in c)
        nb6 = [(cx - 1, cy, cz), (cx + 1, cy, cz), (cx, cy, cz - 1), (cx, cy, cz + 1), (cx, cy - 1, cz), (cx, cy + 1, cz)]
        use = 6 if vertical else 4
        adj = [oracle.generate_terrain(nc) if (k < use and (masks[i] >> k) & 1) else None for k, nc in enumerate(nb6)]
        cubes = oracle.generate_terrain(c)
        if dirt:
            cubes = oracle.dirt_rule(cubes, adj[5])
        assert np.array_equal(got2.blocks[i], cubes)
        arr, keep = oracle._adj_array(adj[:4])
        a6 = (oracle.C.c_void_p * 6)()
        keep6 = []
        for k in range(6):
            if adj[k] is None:
                a6[k] = None
            else:
                # neighbour TYPES do not matter to the mesher (only != AIR), solidity does
                keep6.append(np.ascontiguousarray(adj[k]))
                a6[k] = keep6[-1].ctypes.data
        out = np.empty(98304, dtype=oracle.INSTANCE_DTYPE)
        cc = np.ascontiguousarray(c, np.int32)
        nq = oracle.lib().vxo_create_mesh_ex(cc.ctypes.data, np.ascontiguousarray(cubes).ctypes.data, a6, out.ctypes.data, 98304)
        assert got2.counts[i] == nq
        assert _canon(got2.chunk(i)).tobytes() == _canon(out[:nq]).tobytes()


def test_terrain_opts_off_is_the_reference_path(ctx):
    coords = _stack(0, 0, 2, 2, (0, 1))
    a = ctx.gen_mesh(coords, blocks=True)
    import ctypes as C
    from voxb200._lib import TerrainOpts
    n = len(coords)
    counts = np.zeros(n, np.uint32)
    offsets = np.zeros(n + 1, np.uint64)
    inst = np.empty(a.total, a.instances.dtype)
    blocks = np.empty((n, 32768), np.uint8)
    total = C.c_uint64(0)
    opts = TerrainOpts(0, 0)
    rc = ctx._L.vx_gen_mesh_ex(ctx._h, coords.ctypes.data, n, None, C.byref(opts), None, counts.ctypes.data,
                               offsets.ctypes.data, inst.ctypes.data, len(inst), C.addressof(total), blocks.ctypes.data)
    assert rc == 0 and total.value == a.total
    assert inst.tobytes() == a.instances.tobytes() and np.array_equal(blocks, a.blocks)


def test_out_of_range_neighbour_index(ctx, vox, oracle):
    """an nbr_idx entry >= n must not make the mesher read another allocation.  Host entry point:
    VX_ERR_BAD_ARG before anything is enqueued; device entry point: treated as None, status bit 2 raised."""
    torch = pytest.importorskip("torch")
    rng = np.random.default_rng(9)
    coords = _grid(0, 0, 2, 2)
    n = len(coords)
    blocks = rng.choice(np.array([0, 2], np.uint8), size=(n, 32768), p=[0.5, 0.5])
    wiring = np.full((n, 4), -1, np.int32)
    wiring[1, 0] = 0
    bad = wiring.copy()
    bad[2, 3] = n + 5
    with pytest.raises(vox.VoxError) as e:
        ctx.mesh(coords, blocks, nbr_idx=bad)
    assert e.value.code == vox._lib.VX_ERR_BAD_ARG
    good = ctx.mesh(coords, blocks, nbr_idx=wiring)
    dev = torch.device("cuda", 0)
    d = {k: torch.from_numpy(v).to(dev) for k, v in dict(c=coords, b=blocks, w=bad).items()}
    d_counts = torch.zeros(n, dtype=torch.int32, device=dev)
    d_off = torch.zeros(n + 1, dtype=torch.int64, device=dev)
    d_inst = torch.zeros((good.total + 8) * 17, dtype=torch.int32, device=dev)
    st = torch.zeros(1, dtype=torch.int32, device=dev)
    s = torch.cuda.Stream(device=dev)
    with vox.Context() as c2:
        c2.mesh_device(d["c"].data_ptr(), n, d["b"].data_ptr(), d_counts.data_ptr(), d_off.data_ptr(), d_inst.data_ptr(),
                       good.total + 8, nbr_idx_ptr=d["w"].data_ptr(), status_ptr=st.data_ptr(), stream=s.cuda_stream)
        s.synchronize()
    assert int(st.item()) & 4                                   # VX_ST_BAD_NBR
    assert np.array_equal(d_counts.cpu().numpy().astype(np.uint32), good.counts)     # the bad entry meshed as None
    assert d_inst[:good.total * 17].cpu().numpy().tobytes() == good.instances.tobytes()


def test_frustum_extreme_corner_classes(ctx, vox, oracle):
    """The frustum kernels decide most chunks from two corners per plane (kernels_frustum.cu: the corners the
    coefficient signs pick) and run the reference's corner-by-corner test only on chunks that straddle a plane, or
    on every chunk when a coefficient is NaN, infinite or above 2^90.  Every sign class, signed zeros, denormal and
    huge coefficients, the 2^90 boundary from both sides, plane sets that cut through the grid so that whole rows
    straddle, and the wrapping limit of the coordinates -- one frame per launch and many -- equal the oracle."""
    from tools.bench_configs import camera_planes
    rng = np.random.default_rng(41)
    lim = np.float32(2.0 ** 90)
    sets = []
    for k in range(8):                                      # the eight sign classes on every plane slot
        pl = np.abs(rng.standard_normal((6, 4))).astype(np.float32) + np.float32(0.05)
        pl[:, 0] *= -1 if k & 1 else 1
        pl[:, 1] *= -1 if k & 2 else 1
        pl[:, 2] *= -1 if k & 4 else 1
        pl[:, 3] = rng.uniform(-400, 1500, 6).astype(np.float32)
        sets.append(pl)
    for scale in (1e-42, 1e-30, 1e20, 2.0 ** 89):
        pl = np.clip((rng.standard_normal((6, 4)) * scale).astype(np.float32), -lim, lim)
        pl[:, 3] = (rng.standard_normal(6) * scale * 900).astype(np.float32).clip(-lim, lim)
        sets.append(pl)
    for special in (-0.0, 0.0, lim, -lim, np.nextafter(lim, np.float32(np.inf)), 3e38, np.inf, -np.inf, np.nan):
        pl = rng.standard_normal((6, 4)).astype(np.float32)
        pl[:, 3] += 600.0
        pl[rng.integers(6), rng.integers(3)] = special
        sets.append(pl)
    pl = rng.standard_normal((6, 4)).astype(np.float32)
    pl[:, 3] += 600.0
    pl[2, 3] = np.nan                                        # NaN in d: that plane never rejects (:33)
    sets.append(pl)
    sets.extend(camera_planes(9, vox).reshape(9, 6, 4))      # thin bands of straddlers along the side planes
    planes = np.stack(sets).astype(np.float32)
    g = np.arange(-45, 45, dtype=np.int32)
    z, x = np.meshgrid(g, g, indexing="ij")
    grid = np.stack([x.reshape(-1), rng.integers(-1, 2, x.size).astype(np.int32), z.reshape(-1)], 1)
    far = rng.integers(-(1 << 26), 1 << 26, size=(515, 3)).astype(np.int32)
    far[:3] = [((1 << 26) - 1, 0, -(1 << 26)), (1 << 26, -1, (1 << 26) + 5), (-(1 << 26) - 1, 1, 0)]
    coords = np.ascontiguousarray(np.concatenate([grid, far]).astype(np.int32))
    want = oracle.frustum_cull(planes, coords)
    assert 0 < want.sum() < want.size
    got = ctx.frustum_cull(planes, coords)                                      # general kernel, all frames
    assert np.array_equal(got, want), np.argwhere(got != want)[:8]
    bits = ctx.frustum_cull(planes, coords, bits=True)
    assert np.array_equal(np.unpackbits(bits.view(np.uint8), axis=1, bitorder="little")[:, :len(coords)], want)
    for f in range(len(planes)):                                                # single-frame kernel
        one = ctx.frustum_cull(planes[f:f + 1], coords)
        assert np.array_equal(one[0], want[f]), (f, planes[f])
        b1 = ctx.frustum_cull(planes[f:f + 1], coords, bits=True)
        assert np.array_equal(np.unpackbits(b1.view(np.uint8), axis=1, bitorder="little")[0, :len(coords)], want[f]), f


def test_frustum_single_frame_kernel(ctx, vox, oracle):
    """One plane set per launch takes its own kernel (four chunks per thread group, 128-bit coordinate loads, one or
    two groups per thread): ragged sizes around every granularity, NaN / zero / wrapping inputs, both outputs, and
    misaligned device pointers (which must fall back to the general kernel) -- all equal to the oracle."""
    torch = pytest.importorskip("torch")
    rng = np.random.default_rng(23)
    for n in (1, 3, 4, 5, 31, 32, 33, 127, 128, 129, 1000, 4099, 606208, 606211, 700003):
        planes = rng.standard_normal((1, 6, 4)).astype(np.float32)
        planes[:, :, 3] += 300.0
        if n % 3 == 0:
            planes[0, 4, 1] = np.nan              # NaN distance counts as inside (frustum_culling.rs:33)
        coords = rng.integers(-60, 60, size=(n, 3)).astype(np.int32)
        coords[0] = (1 << 26, 0, -(1 << 26))      # wrapping world coordinates
        want = oracle.frustum_cull(planes, coords)
        got = ctx.frustum_cull(planes, coords)
        assert np.array_equal(got, want), n
        bits = ctx.frustum_cull(planes, coords, bits=True)
        assert np.array_equal(np.unpackbits(bits.view(np.uint8), axis=1, bitorder="little")[:, :n], want), n
        assert bits.shape[1] == (n + 31) // 32
    # device pointers that are not 16-byte (coords) / 4-byte (visible) aligned
    dev = torch.device("cuda", 0)
    n = 5001
    coords = rng.integers(-60, 60, size=(n, 3)).astype(np.int32)
    planes = rng.standard_normal((1, 6, 4)).astype(np.float32)
    planes[:, :, 3] += 300.0
    want = oracle.frustum_cull(planes, coords)[0]
    d_p = torch.from_numpy(planes).to(dev)
    raw_c = torch.zeros(3 * n + 8, dtype=torch.int32, device=dev)
    raw_v = torch.zeros(n + 64, dtype=torch.uint8, device=dev)
    d_bits = torch.zeros((n + 31) // 32, dtype=torch.int32, device=dev)
    s = torch.cuda.Stream(device=dev)
    with vox.Context() as c2:
        for c_off, v_off in ((0, 0), (1, 0), (0, 3), (3, 1)):
            raw_c[c_off:c_off + 3 * n] = torch.from_numpy(coords.reshape(-1)).to(dev)
            raw_v.zero_()
            c2.frustum_cull_device(d_p.data_ptr(), 1, raw_c.data_ptr() + 4 * c_off, n, visible_ptr=raw_v.data_ptr() + v_off,
                                   bits_ptr=d_bits.data_ptr(), stream=s.cuda_stream)
            s.synchronize()
            assert np.array_equal(raw_v[v_off:v_off + n].cpu().numpy(), want), (c_off, v_off)
            assert int(raw_v[v_off + n:].sum().item()) == 0 and int(raw_v[:v_off].sum().item()) == 0
            b = np.unpackbits(d_bits.cpu().numpy().view(np.uint8), bitorder="little")[:n]
            assert np.array_equal(b, want)
