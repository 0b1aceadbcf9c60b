"""CPU: the DEVICE mesher core (csrc/vx_mesh_rows.cuh -- the face masks of a 32-voxel row with every border rule of
chunk.rs:59-245, and the rotation / texture tables of quad.rs:16-40 and texture.rs:238-243) compiled for the host must
give the oracle's instance stream byte for byte.  The kernels around it (staging, scan, warp-level quad decode, bulk
stores) are covered by the `-m gpu` tests."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
INST = np.dtype([("t_matrix", np.float32, (4, 4)), ("texture_index", np.uint32)])


@pytest.fixture(scope="module")
def shim(tmp_path_factory):
    so = str(tmp_path_factory.mktemp("shim") / "libmesh_host.so")
    src = os.path.join(ROOT, "tests", "host_shim", "mesh_host.cpp")
    subprocess.check_call(["g++", "-O2", "-std=c++17", "-fPIC", "-shared", "-o", so, src])
    L = C.CDLL(so)
    L.shim_mesh_chunk.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_void_p]
    L.shim_mesh_chunk.restype = C.c_uint32
    return L


def _device(shim, coord, cubes, adj, use4):
    c = np.array(coord, np.int32)
    cubes = np.ascontiguousarray(cubes, np.uint8)
    keep = [None if a is None else np.ascontiguousarray(a, np.uint8) for a in adj]
    ptrs = (C.c_void_p * 4)(*[None if a is None else a.ctypes.data for a in keep])
    out = np.zeros(98304, INST)
    n = shim.shim_mesh_chunk(c.ctypes.data, cubes.ctypes.data, ptrs, int(use4), out.ctypes.data)
    assert n != 0xffffffff, "row_count disagrees with the face masks"
    return out[:n]


def _canon(a):
    b = a.copy()
    b["t_matrix"] += np.float32(0)          # signed zeros of the rotation products are nalgebra-internal
    return b.tobytes()


@pytest.mark.parametrize("use4", [False, True])
def test_terrain_chunks_with_every_neighbour_pattern(shim, oracle, use4):
    coord = (3, 0, -2)
    cubes = oracle.generate_terrain(coord)
    nbrs = [oracle.generate_terrain(c) for c in ((2, 0, -2), (4, 0, -2), (3, 0, -3), (3, 0, -1))]
    for mask in range(16):
        adj = [nbrs[k] if (mask >> k) & 1 else None for k in range(4)]
        want = oracle.create_mesh(coord, cubes, adj)
        got = _device(shim, coord, cubes, adj, use4)
        assert len(got) == len(want) and _canon(got) == _canon(want), mask


@pytest.mark.parametrize("use4", [False, True])
def test_random_grids_with_dirt_and_far_coordinates(shim, oracle, use4):
    rng = np.random.default_rng(9)
    for p_solid, coord in ((0.08, (0, 0, 0)), (0.5, (-7, 2, 511)), (0.93, ((1 << 26) - 1, -1, -(1 << 26)))):
        grids = [np.where(rng.random(32768) < p_solid, rng.integers(0, 2, 32768), 2).astype(np.uint8) for _ in range(5)]
        adj = [grids[1], None, grids[3], grids[4]]
        want = oracle.create_mesh(coord, grids[0], adj)
        got = _device(shim, coord, grids[0], adj, use4)
        assert len(got) == len(want) and _canon(got) == _canon(want)


def test_checkerboard_worst_case(shim, oracle):
    g = np.indices((32, 32, 32)).sum(axis=0) & 1          # [y][z][x] order does not matter for a checkerboard
    cb = np.where(g == 0, 0, 2).astype(np.uint8).reshape(-1)
    got = _device(shim, (0, 0, 0), cb, [None] * 4, True)
    assert len(got) == 98304
    assert _canon(got) == _canon(oracle.create_mesh((0, 0, 0), cb, [None] * 4))
