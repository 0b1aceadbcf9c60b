"""Property tests: random block grids and neighbour wiring."""
import numpy as np
import pytest
from hypothesis import HealthCheck, given, settings, strategies as st


def _grid(rng, p_solid, p_dirt):
    solid = rng.random(32768) < p_solid
    dirt = rng.random(32768) < p_dirt
    return np.where(solid, np.where(dirt, 1, 0), 2).astype(np.uint8)


@settings(max_examples=12, deadline=None, suppress_health_check=[HealthCheck.function_scoped_fixture])
@given(seed=st.integers(0, 2 ** 31 - 1), p=st.sampled_from([0.02, 0.1, 0.5, 0.9, 0.98]))
def test_oracle_face_identities(oracle, seed, p):
    """Without neighbours every solid run along an axis has exactly one face on each end, so the two
    faces of an axis balance (chunk.rs:77-240 with all border faces emitted), and a face exists
    exactly where a solid voxel meets air or the chunk border."""
    rng = np.random.default_rng(seed)
    cubes = _grid(rng, p, 0.3)
    q = oracle.cull_unseen_triangles((3, -1, 2), cubes, None)
    n = np.bincount(q["face"], minlength=6)
    assert n[oracle.LEFT] == n[oracle.RIGHT] and n[oracle.BACK] == n[oracle.FRONT] and n[oracle.DOWN] == n[oracle.UP]
    s = (cubes != 2).reshape(32, 32, 32)                 # [y][z][x]
    pad = np.pad(s, 1, constant_values=False)
    want = sum(int((s & ~np.roll(pad, sh, ax)[1:-1, 1:-1, 1:-1]).sum()) for ax in range(3) for sh in (1, -1))
    assert len(q) == want
    # emission order is (y, z, x)-major with faces ascending inside a cube
    idx = (q["x"] - 96) + 32 * (q["z"] - 64) + 1024 * (q["y"] + 32)
    key = idx.astype(np.int64) * 8 + q["face"]
    assert np.all(np.diff(key) > 0)
    # textures: DIRT always 1; GRASS sides 0, DOWN 1, UP 2 (texture.rs:238-243 through slots 0,1,2,3,5,4)
    typ = cubes[idx]
    grass_tex = np.array([0, 0, 0, 0, 1, 2])[q["face"]]
    assert np.array_equal(q["tex"], np.where(typ == 1, 1, grass_tex))


@pytest.mark.gpu
@settings(max_examples=10, deadline=None, suppress_health_check=[HealthCheck.function_scoped_fixture])
@given(seed=st.integers(0, 2 ** 31 - 1), p=st.sampled_from([0.05, 0.5, 0.95]), n=st.integers(1, 6))
def test_gpu_mesher_equals_oracle_on_random_wiring(vox, oracle, seed, p, n):
    rng = np.random.default_rng(seed)
    coords = rng.integers(-1000, 1000, size=(n, 3)).astype(np.int32)
    blocks = np.stack([_grid(rng, p, 0.5) for _ in range(n)])
    nbr = rng.integers(-1, n, size=(n, 4)).astype(np.int32)
    mask = (rng.random(n) < 0.8).astype(np.uint8)
    ctx = vox.default_context()
    got = ctx.mesh(coords, blocks, nbr_idx=nbr, mesh_mask=mask)
    want = oracle.mesh_batch(coords, blocks, nbr)
    for i in range(n):
        if mask[i]:
            lo, hi = int(want["offsets"][i]), int(want["offsets"][i + 1])
            a = got.chunk(i).copy()
            b = want["instances"][lo:hi].copy()
            a["t_matrix"] += np.float32(0)
            b["t_matrix"] += np.float32(0)
            assert a.tobytes() == b.tobytes()
        else:
            assert got.counts[i] == 0


@pytest.mark.gpu
@settings(max_examples=8, deadline=None, suppress_health_check=[HealthCheck.function_scoped_fixture])
@given(seed=st.integers(0, 2 ** 31 - 1), F=st.integers(1, 9))
def test_gpu_frustum_equals_oracle_on_random_planes(vox, oracle, seed, F):
    rng = np.random.default_rng(seed)
    planes = (rng.standard_normal((F, 6, 4)) * rng.choice([1e-3, 1.0, 1e3])).astype(np.float32)
    planes[:, :, 3] += rng.choice([0.0, 100.0, -50.0])
    coords = rng.integers(-60, 60, size=(int(rng.integers(1, 700)), 3)).astype(np.int32)
    got = vox.default_context().frustum_cull(planes, coords)
    assert np.array_equal(got, oracle.frustum_cull(planes, coords))
