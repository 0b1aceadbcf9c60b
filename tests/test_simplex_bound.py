"""CPU: the bound behind the fill kernel's sign-only shortcuts (kernels_fill.cu):
|simplex_3d| <= 1/32.69587 = 0.030585 for ANY gradient assignment, so skipping an
octave whose maximal contribution cannot flip the sign keeps blocks bit-exact."""
import numpy as np

B = 0.0306          # the constant the kernel uses (0.05 % above the supremum)


def _upper_envelope(p):
    """sum_n max(0, 0.6 - |d_n|^2)^4 * max_g |g . d_n| with g over the 12 (+-u +-v) directions:
    an upper bound of |simplex_3d| whatever the lattice hashes are."""
    x, y, z = p[..., 0], p[..., 1], p[..., 2]
    f = (x + y + z) / 3.0
    i, j, k = np.floor(x + f), np.floor(y + f), np.floor(z + f)
    g = (i + j + k) / 6.0
    x0, y0, z0 = x - (i - g), y - (j - g), z - (k - g)
    xy, yz, xz = x0 >= y0, y0 >= z0, x0 >= z0
    i1, j1, k1 = xy & xz, ~xy & yz, ~xz & ~yz
    i2, j2, k2 = xy | xz, ~xy | yz, ~(xz & yz)
    G3, F3 = 1 / 6.0, 1 / 3.0
    corners = [(x0, y0, z0), (x0 - i1 + G3, y0 - j1 + G3, z0 - k1 + G3),
               (x0 - i2 + F3, y0 - j2 + F3, z0 - k2 + F3), (x0 - 0.5, y0 - 0.5, z0 - 0.5)]
    tot = 0.0
    for dx, dy, dz in corners:
        t = np.maximum(0.6 - dx * dx - dy * dy - dz * dz, 0.0)
        a = np.sort(np.abs(np.stack([dx, dy, dz], -1)), axis=-1)
        tot = tot + t ** 4 * (a[..., 1] + a[..., 2])
    return tot


def test_envelope_supremum_is_the_normalisation_constant():
    n = 96
    g = (np.arange(n) + 0.5) / n
    p = np.stack(np.meshgrid(g, g, g, indexing="ij"), -1).reshape(-1, 3)
    v = _upper_envelope(p)
    best, at = float(v.max()), p[v.argmax()]
    rng = np.random.default_rng(0)
    for it in range(30):                       # local refinement around the grid maximum
        cand = at + rng.normal(0, 0.02 / (1 + 0.3 * it), size=(20000, 3))
        v = _upper_envelope(cand)
        if v.max() > best:
            best, at = float(v.max()), cand[v.argmax()]
    assert abs(best - 1 / 32.69587) < 2e-6      # FastNoise's simplex normaliser is exactly this supremum
    assert best < B


def test_oracle_samples_respect_the_bound(oracle):
    L = oracle.lib()
    rng = np.random.default_rng(1)
    pts = rng.uniform(-2000, 2000, size=(200000, 3)).astype(np.float32)
    m = 0.0
    for seed in (1, 2, 12345):
        for x, y, z in pts[:60000]:
            m = max(m, abs(L.vxo_simplex_3d(x, y, z, seed)))
    assert 0.025 < m <= 1 / 32.69587 + 1e-6
