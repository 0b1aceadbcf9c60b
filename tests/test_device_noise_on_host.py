"""CPU: the DEVICE noise code (csrc/vx_noise.cuh) compiled for the host through a
tiny intrinsic shim must reproduce the oracle bit for bit.  Catches arithmetic
slips in the CUDA source without a GPU."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def shim(tmp_path_factory):
    so = str(tmp_path_factory.mktemp("shim") / "libnoise_host.so")
    src = os.path.join(ROOT, "tests", "host_shim", "noise_host.cpp")
    subprocess.check_call(["g++", "-O2", "-std=c++17", "-fPIC", "-shared", "-ffp-contract=off", "-o", so, src])
    L = C.CDLL(so)
    L.shim_noise_chunk.argtypes = [C.c_void_p, C.c_int32, C.c_float, C.c_float, C.c_float, C.c_int, C.c_void_p]
    L.shim_noise_chunk_sign.argtypes = [C.c_void_p, C.c_int32, C.c_float, C.c_float, C.c_float, C.c_void_p]
    L.shim_noise_chunk_fused.argtypes = [C.c_void_p, C.c_int32, C.c_float, C.c_float, C.c_float, C.c_void_p]
    L.shim_fused_err_ulps.restype = C.c_float
    L.shim_set_scale.argtypes = [C.c_float]
    L.shim_noise_chunk_lattice.argtypes = [C.c_void_p, C.c_int32, C.c_float, C.c_float, C.c_float, C.c_void_p]
    for f in (L.shim_noise_chunk_x2, L.shim_noise_chunk_x1):
        f.argtypes = [C.c_void_p, C.c_int32, C.c_float, C.c_float, C.c_float, C.c_int, C.c_void_p]
    yield L


SCALED = 32.69587493801679          # vx_noise_params.simplex_scale default; 1.0 = the raw sum


@pytest.fixture(params=[1.0, SCALED], ids=["raw", "scaled"])
def scale(request, shim):
    shim.shim_set_scale(request.param)
    yield request.param
    shim.shim_set_scale(1.0)


@pytest.mark.parametrize("coord", [(0, 0, 0), (-1, 0, -1), (511, 0, -512), (7, 1, -3), (100000, 3, -99999)])
def test_device_noise_source_matches_oracle(shim, oracle, coord, scale):
    c = np.array(coord, np.int32)
    out = np.empty(32768, np.float32)
    shim.shim_noise_chunk(c.ctypes.data, 1, 0.09, 2.5, 0.5, 3, out.ctypes.data)
    want = oracle.noise_chunk(coord, oracle.default_params(simplex_scale=scale))
    assert np.array_equal(out.view(np.uint32), want.view(np.uint32))


def test_device_noise_other_params(shim, oracle, scale):
    c = np.array((3, -1, 4), np.int32)
    out = np.empty(32768, np.float32)
    shim.shim_noise_chunk(c.ctypes.data, 12345, 0.9, 1.7, 0.4, 5, out.ctypes.data)
    want = oracle.noise_chunk((3, -1, 4), oracle.default_params(seed=12345, freq=0.9, gain=1.7, lacunarity=0.4, octaves=5,
                                                                simplex_scale=scale))
    assert np.array_equal(out.view(np.uint32), want.view(np.uint32))


@pytest.mark.parametrize("coord,seed,gain,lac,freq", [((0, 0, 0), 1, 2.5, 0.5, 0.09), ((-1, 0, -1), 1, 2.5, 0.5, 0.09),
                                                       ((511, 0, -512), 1, 2.5, 0.5, 0.09), ((9, -3, 4), 77, -3.0, 0.7, 0.2),
                                                       ((100000, 3, -99999), 5, 1.0, 2.0, 0.03)])
def test_sign_only_gradient_variant(shim, oracle, coord, seed, gain, lac, freq, scale):
    """vxn_simplex3_sign (coefficient-table gradient on the FMA pipe) equals the exact noise in value;
    only the sign of a zero may differ, which `noise < 0` cannot see."""
    c = np.array(coord, np.int32)
    out = np.empty(32768, np.float32)
    shim.shim_noise_chunk_sign(c.ctypes.data, seed, freq, gain, lac, out.ctypes.data)
    want = oracle.noise_chunk(coord, oracle.default_params(seed=seed, freq=freq, gain=gain, lacunarity=lac,
                                                           simplex_scale=scale))
    assert np.array_equal(out, want)                       # value equality (+0 == -0)
    nz = want != 0
    assert np.array_equal(out.view(np.uint32)[nz], want.view(np.uint32)[nz])
    assert np.array_equal(out < 0, want < 0)


def test_fused_evaluation_stays_inside_its_certified_bound(shim, oracle, scale):
    """The FMA evaluation used for sign decisions differs from the exact noise by far less than the
    margin the kernel demands before trusting it."""
    eps = 2.0 ** -24
    ulps = float(shim.shim_fused_err_ulps())
    rng = np.random.default_rng(4)
    worst = 0.0
    for coord in [(0, 0, 0), (511, 0, -512), (-77, 2, 33)] + [tuple(int(v) for v in rng.integers(-9000, 9000, 3)) for _ in range(9)]:
        c = np.array(coord, np.int32)
        out = np.empty(32768, np.float32)
        shim.shim_noise_chunk_fused(c.ctypes.data, 1, 0.09, 2.5, 0.5, out.ctypes.data)
        want = oracle.noise_chunk(coord, oracle.default_params(simplex_scale=scale))
        worst = max(worst, float(np.max(np.abs(out.astype(np.float64) - want.astype(np.float64)))))
    amps = 1 + 2.5 + 6.25
    # vx_launch_fill's err_margin / 3: raw-sum bound x |scale| x amplitudes + the combination's roundings
    bound = (ulps * scale + 5 * max(1.0, 0.0306 * scale)) * amps * eps
    assert worst <= bound, (worst, bound)
    assert worst > 0                                  # it really is a different rounding sequence
    print("max |fused - exact| = %.3g, certified bound %.3g, kernel margin %.3g" % (worst, bound, 3 * bound))


@pytest.mark.parametrize("coord,seed,gain,lac,freq", [((0, 0, 0), 1, 2.5, 0.5, 0.09), ((511, 0, -512), 1, 2.5, 0.5, 0.09),
                                                       ((9, -3, 4), 77, -3.0, 0.7, 0.2), ((100000, 3, -99999), 5, 1.0, 2.0, 0.03)])
def test_packed_two_voxel_variant_is_bit_equal_to_the_scalar_fused_one(shim, coord, seed, gain, lac, freq):
    """vxn_simplex3_fused_x2 (two voxels per thread on FADD2/FMUL2/FFMA2) runs each element through the scalar
    fused variant's operation sequence: bit-equal, so the certified error bound carries over unchanged."""
    c = np.array(coord, np.int32)
    a = np.empty(32768, np.float32)
    b = np.empty(32768, np.float32)
    shim.shim_noise_chunk_x2(c.ctypes.data, seed, freq, gain, lac, 1, a.ctypes.data)
    shim.shim_noise_chunk_x1(c.ctypes.data, seed, freq, gain, lac, 1, b.ctypes.data)
    assert np.array_equal(a.view(np.uint32), b.view(np.uint32))


@pytest.mark.parametrize("coord,seed,gain,lac,freq,expect_fit", [
    ((0, 0, 0), 1, 2.5, 0.5, 0.09, True), ((511, 0, -512), 1, 2.5, 0.5, 0.09, True), ((-1, -1, -1), 1, 2.5, 0.5, 0.09, True),
    ((9, -3, 4), 77, -3.0, 0.7, 0.1, True), ((100000, 3, -99999), 5, 1.0, 2.0, 0.02, True),
    ((-200000, 0, 200000), 1, 2.5, 0.5, 0.09, True), ((-30000, 2, 7), 9, 2.5, -0.5, -0.09, True),
    ((3, 0, 3), 1, 2.5, 0.5, 0.9, False), ((40000000, 0, -40000000), 1, 2.5, 0.5, 0.09, None)])
def test_lattice_table_variant_is_bit_equal_to_the_scalar_fused_one(shim, coord, seed, gain, lac, freq, expect_fit):
    """vxn_simplex3_lattice_x2: gradients from a per-CTA table hashed once per lattice point (index box from
    the 8 box corners) give exactly the per-voxel-hash values; boxes that do not fit are reported, never
    silently truncated (an unbuilt entry is NaN in this host build)."""
    c = np.array(coord, np.int32)
    a = np.full(32768, np.nan, np.float32)
    b = np.empty(32768, np.float32)
    unfit = shim.shim_noise_chunk_lattice(c.ctypes.data, seed, freq, gain, lac, a.ctypes.data)
    shim.shim_noise_chunk_x1(c.ctypes.data, seed, freq, gain, lac, 1, b.ctypes.data)
    if expect_fit is True:
        assert unfit == 0
    elif expect_fit is False:
        assert unfit == 8
    got = ~np.isnan(a)
    zb = (np.arange(32768) // 32 % 32) // 4
    assert got.reshape(-1).sum() == (8 - unfit) * 4096
    assert np.array_equal(a.view(np.uint32)[got], b.view(np.uint32)[got])
    assert len(set(zb[~got])) == unfit


def test_lattice_index_box_holds_for_random_chunks(shim):
    """The 8-corner index box must contain the floor indices of every voxel of the z block (monotone rounding
    argument, vx_noise.cuh): random chunks from the origin out to |chunk| = 6e7 and random parameters of either
    sign.  An index outside its box would read a NaN-poisoned table entry (or miss the x1 values)."""
    rng = np.random.default_rng(77)
    fitted = unfitted = 0
    for case in range(60):
        scale = [40, 5000, 400000, 3000000, 60000000][case % 5]
        coord = tuple(int(v) for v in (rng.integers(-scale, scale + 1), rng.integers(-4, 5), rng.integers(-scale, scale + 1)))
        freq = float(np.exp(rng.uniform(np.log(0.003), np.log(0.12)))) * (1 if rng.random() < 0.75 else -1)
        lac = float(rng.uniform(0.2, 1.6)) * (1 if rng.random() < 0.75 else -1)
        gain = float(rng.uniform(0.3, 4.0))
        seed = int(rng.integers(0, 2 ** 31 - 1))
        c = np.array(coord, np.int32)
        a = np.full(32768, np.nan, np.float32)
        b = np.empty(32768, np.float32)
        unfit = shim.shim_noise_chunk_lattice(c.ctypes.data, seed, freq, gain, lac, a.ctypes.data)
        shim.shim_noise_chunk_x1(c.ctypes.data, seed, freq, gain, lac, 1, b.ctypes.data)
        got = ~np.isnan(a)
        assert got.sum() == (8 - unfit) * 4096, (case, coord, freq, lac)
        assert np.array_equal(a.view(np.uint32)[got], b.view(np.uint32)[got]), (case, coord, freq, lac)
        fitted += 8 - unfit
        unfitted += unfit
    assert fitted > 100 and unfitted > 20          # both outcomes are exercised
