"""The ordering argument behind the frustum kernels' extreme-corner classification (csrc/kernels_frustum.cu):
with finite coefficients the reference's separately rounded distance ((a*x + b*y) + c*z) + d
(frustum_culling.rs:29-32) is monotone in every coordinate, so over an AABB's eight corners it is smallest at the
corner picked by the coefficient signs and largest at the opposite one.  Checked here on the CPU against the
oracle's corner-by-corner `Frustum::contains` (:16-40), on inputs chosen to break it if it could break."""
import numpy as np
import pytest

f32 = np.float32
LIM = f32(2.0 ** 90)          # kernels_frustum.cu plane_class: larger coefficients take the corner-by-corner test


def classify(planes, coords):
    """(rejected, inside) per chunk as the kernels decide them: planes f32[6][4], coords i32[n][3].  `inside` also
    covers a chunk that straddles exactly ONE plane: all corners pass the other five and the p corner passes that one."""
    lo = (coords.astype(np.int64) * 32).astype(np.int32).astype(f32)        # coordinate.rs:53-58, wrapping
    hi = lo + f32(32.0)                                                     # frustum_culling.rs:19-26
    rejected = np.zeros(len(coords), bool)
    straddled = np.zeros(len(coords), np.int32)
    with np.errstate(all="ignore"):
        for a, b, c, d in planes:
            n = [np.where(k < 0, hi[:, i], lo[:, i]) for i, k in enumerate((a, b, c))]
            p = [np.where(k < 0, lo[:, i], hi[:, i]) for i, k in enumerate((a, b, c))]
            dn = ((a * n[0] + b * n[1]) + c * n[2]) + d
            dp = ((a * p[0] + b * p[1]) + c * p[2]) + d
            assert dn.dtype == f32 and np.all(dn <= dp)
            rejected |= dp < 0
            straddled += np.signbit(dn)                      # the kernels count sign bits (-0 included)
    return rejected, ~rejected & (straddled <= 1)


def usable(planes):
    return bool(np.all(np.abs(planes) <= LIM))            # False for NaN as well


def plane_sets(rng, count):
    for k in range(count):
        scale = f32(rng.choice([1e-42, 1e-30, 1e-3, 1.0, 1e3, 1e20, 2.0 ** 89]))
        pl = (rng.standard_normal((6, 4)) * scale).astype(f32)
        pl[:, 3] = (rng.standard_normal(6) * scale * rng.choice([0.0, 1.0, 50.0, 1000.0])).astype(f32)
        pl = np.clip(pl, -LIM, LIM)
        if k % 5 == 0:
            pl[rng.integers(6), rng.integers(3)] = f32(-0.0)
        if k % 7 == 0:
            pl[rng.integers(6), rng.integers(3)] = f32(0.0)
        if k % 11 == 0:
            pl[rng.integers(6), rng.integers(4)] = LIM * f32(rng.choice([-1, 1]))
        yield pl


def test_extreme_corners_decide_exactly_like_the_reference(oracle):
    rng = np.random.default_rng(5)
    coords = np.concatenate([
        rng.integers(-40, 40, size=(1500, 3)),
        rng.integers(-(1 << 26), 1 << 26, size=(400, 3)),                    # out to the wrapping limit
        [[(1 << 26) - 1, 0, -(1 << 26)], [1 << 26, -1, (1 << 26) + 5], [0, 0, 0], [-1, -1, -1]],
    ]).astype(np.int32)
    decided = straddling = 0
    for pl in plane_sets(rng, 150):
        assert usable(pl)
        want = oracle.frustum_cull(pl.reshape(1, 6, 4), coords)[0].astype(bool)
        rejected, inside = classify(pl, coords)
        assert not np.any(rejected & inside)
        assert not np.any(want[rejected]), "a rejected chunk has a corner inside all planes"
        assert np.all(want[inside]), "an all-inside chunk fails the reference test"
        decided += int((rejected | inside).sum())
        straddling += int((~(rejected | inside)).sum())
    assert decided > 20 * straddling > 0


def test_camera_frustum_mostly_decided_without_the_corner_test(oracle):
    import voxb200
    from tools.bench_configs import camera_planes
    g = np.arange(-100, 100, dtype=np.int32)
    z, x = np.meshgrid(g, g, indexing="ij")
    coords = np.stack([x.reshape(-1), np.zeros(x.size, np.int32), z.reshape(-1)], 1).astype(np.int32)
    planes = camera_planes(12, voxb200).reshape(12, 6, 4)
    for pl in planes:
        assert usable(pl)
        want = oracle.frustum_cull(pl.reshape(1, 6, 4), coords)[0].astype(bool)
        rejected, inside = classify(pl, coords)
        assert not np.any(want[rejected]) and np.all(want[inside])
        assert (~(rejected | inside)).mean() < 0.002           # what is left for the corner test: frustum edges and apex


def test_unusable_plane_sets_are_recognised():
    pl = np.ones((6, 4), f32)
    assert usable(pl)
    for bad in (np.nan, np.inf, -np.inf, np.nextafter(LIM, f32(np.inf)), f32(-3e38)):
        q = pl.copy()
        q[3, 1] = bad
        assert not usable(q)
    # where the bound comes from: |coordinate| <= 2^31, so coefficients up to 2^90 keep products and sums finite
    big = f32(2.0 ** 31) * LIM
    assert np.isfinite(((big + big) + big) + LIM)


from hypothesis import HealthCheck, given, settings  # noqa: E402
from hypothesis import strategies as st  # noqa: E402

_coef = st.floats(min_value=-float(LIM), max_value=float(LIM), width=32, allow_nan=False, allow_infinity=False)
_chunk = st.integers(min_value=-(1 << 31), max_value=(1 << 31) - 1)       # any i32: the world coordinate wraps


@settings(max_examples=120, deadline=None, suppress_health_check=[HealthCheck.function_scoped_fixture])
@given(planes=st.lists(_coef, min_size=24, max_size=24), coords=st.lists(st.tuples(_chunk, _chunk, _chunk), min_size=1, max_size=24))
def test_any_finite_plane_set_and_any_i32_chunk(oracle, planes, coords):
    """hypothesis: whatever finite coefficients (subnormals, signed zeros, up to 2^90) and whatever i32 chunk coordinates,
    a chunk the two extreme corners reject is invisible for the reference and a chunk they accept is visible."""
    pl = np.array(planes, f32).reshape(6, 4)
    c = np.array(coords, np.int64).astype(np.int32)
    want = oracle.frustum_cull(pl.reshape(1, 6, 4), c)[0].astype(bool)
    rejected, inside = classify(pl, c)
    assert not np.any(rejected & inside)
    assert not np.any(want[rejected]) and np.all(want[inside])
