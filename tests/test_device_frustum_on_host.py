"""CPU: the DEVICE frustum logic (csrc/vx_frustum.cuh -- plane classes, extreme-corner classification, the
corner-by-corner test and the decision that combines them) compiled for the host through an intrinsic shim (a warp of
one lane) must give the oracle's `Frustum::contains` (frustum_culling.rs:16-40) on every input.  Catches slips in the
CUDA source without a GPU; the kernels around it are covered by the `-m gpu` tests."""
import ctypes as C
import os
import subprocess
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
f32 = np.float32
LIM = f32(2.0 ** 90)


@pytest.fixture(scope="module", params=[0, 1], ids=["plane-by-plane", "plane-pairs"])
def shim(tmp_path_factory, request):
    # both compile-time forms of the classification loop (VX_FR_PAIR_MIN: two planes join the running minimum at once)
    so = str(tmp_path_factory.mktemp("shim") / "libfrustum_host.so")
    src = os.path.join(ROOT, "tests", "host_shim", "frustum_host.cpp")
    subprocess.check_call(["g++", "-O2", "-std=c++17", "-fPIC", "-shared", "-ffp-contract=off",
                           "-DVX_FR_PAIR_MIN=%d" % request.param, "-o", so, src])
    L = C.CDLL(so)
    L.shim_frustum.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_int]
    L.shim_corner_test.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_void_p]
    L.shim_plane_classes.argtypes = [C.c_void_p]
    L.shim_plane_classes.restype = C.c_uint32
    return L


def _device(shim, planes, coords, unroll=6):
    planes = np.ascontiguousarray(planes, f32).reshape(24)
    coords = np.ascontiguousarray(coords, np.int32)
    out = np.zeros(len(coords), np.uint8)
    shim.shim_frustum(planes.ctypes.data, coords.ctypes.data, len(coords), out.ctypes.data, unroll)
    return out


def _coords(rng):
    return np.concatenate([
        rng.integers(-40, 40, size=(1201, 3)),                               # ragged: 1201 = 300 groups + 1
        rng.integers(-(1 << 26), 1 << 26, size=(300, 3)),                    # out to the wrapping limit
        rng.integers(-(1 << 31), 1 << 31, size=(100, 3)),                    # any i32
        [[(1 << 26) - 1, 0, -(1 << 26)], [0, 0, 0], [-1, -1, -1]],
    ]).astype(np.int32)


def test_device_frustum_source_matches_oracle(shim, oracle):
    from test_frustum_extremes import plane_sets
    rng = np.random.default_rng(17)
    coords = _coords(rng)
    n_vis = 0
    for k, pl in enumerate(plane_sets(rng, 120)):
        want = oracle.frustum_cull(pl.reshape(1, 6, 4), coords)[0]
        got = _device(shim, pl, coords, unroll=(6, 2, 1)[k % 3])
        assert np.array_equal(got, want), (k, pl)
        n_vis += int(want.sum())
        assert shim.shim_plane_classes(np.ascontiguousarray(pl).ctypes.data) & 0x888888 == 0
    assert n_vis > 0


@pytest.mark.parametrize("special", [np.nan, np.inf, -np.inf, float(np.nextafter(LIM, f32(np.inf))), 3e38, -3e38])
def test_unusable_coefficients_take_the_corner_test(shim, oracle, special):
    rng = np.random.default_rng(3)
    coords = _coords(rng)
    for slot in range(4):
        pl = rng.standard_normal((6, 4)).astype(f32)
        pl[:, 3] += f32(600.0)
        pl[rng.integers(6), slot] = special
        assert shim.shim_plane_classes(pl.ctypes.data) & 0x888888
        want = oracle.frustum_cull(pl.reshape(1, 6, 4), coords)[0]
        assert np.array_equal(_device(shim, pl, coords), want), (special, slot)


def test_camera_frusta_and_the_corner_test_alone(shim, oracle):
    import voxb200
    from tools.bench_configs import camera_planes
    g = np.arange(-60, 60, dtype=np.int32)
    z, x = np.meshgrid(g, g, indexing="ij")
    coords = np.stack([x.reshape(-1), np.zeros(x.size, np.int32), z.reshape(-1)], 1).astype(np.int32)
    for pl in camera_planes(16, voxb200).reshape(16, 6, 4):
        want = oracle.frustum_cull(pl.reshape(1, 6, 4), coords)[0]
        assert 0 < want.sum() < want.size
        assert np.array_equal(_device(shim, pl, coords), want)
        alone = np.zeros(len(coords), np.uint8)
        p = np.ascontiguousarray(pl, f32)
        shim.shim_corner_test(p.ctypes.data, coords.ctypes.data, len(coords), alone.ctypes.data)
        assert np.array_equal(alone, want)                    # the straddlers' path is the reference's test for every chunk
