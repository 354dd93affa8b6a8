"""Exactness of the two arithmetic shortcuts the kernels take on the decision-carrying path (mt_math.cuh,
mt_tex.cuh), checked exhaustively / at scale on the host with the same IEEE operations the GPU executes."""
import numpy as np


import pytest


@pytest.mark.parametrize("divisor", [12500.0, 10.0, 9.0, 5.0])
def test_division_by_constant_is_exact_for_every_significand(divisor):
    """div_thickness / MT_DIV_CONST: fma(fma(-q, d, x), r, q), q = x*r, r = RN(1/d) equals x / d for all 2^23
    significands (the result is scale-invariant across binades in the normal range), both signs."""
    d = np.float32(divisor)
    r = np.float32(1.0) / d
    bits = (np.uint32(140) << np.uint32(23)) | np.arange(1 << 23, dtype=np.uint32)  # one full binade (~8192..16384)
    for sign in (1.0, -1.0):
        x = bits.view(np.float32) * np.float32(sign)
        q = x * r
        # fma in float64 is exact here: products of two binary32 values fit in 48 bits
        e = (x.astype(np.float64) - q.astype(np.float64) * np.float64(d)).astype(np.float32)
        assert np.array_equal(e.astype(np.float64), x.astype(np.float64) - q.astype(np.float64) * np.float64(d))  # residual is exact
        got = (q.astype(np.float64) + e.astype(np.float64) * np.float64(r)).astype(np.float32)  # one rounding, like fmaf
        assert np.array_equal(got, x / d)


def test_denormal_scaled_texel_unpack_is_linear():
    """The bit pattern c << 16 is the float c * 2^-133 for every byte c (binary32 is linear across the
    denormal/normal boundary), so one PRMT replaces the int->float conversion."""
    c = np.arange(256, dtype=np.uint32)
    f = (c << np.uint32(16)).view(np.float32)
    assert np.array_equal(f.astype(np.float64), c.astype(np.float64) * 2.0**-133)
    # scaling the z weight by 2^120 and the final 1/255 by 2^13 leaves a two-term filter bit-identical
    rng = np.random.default_rng(11)
    n = 2_000_000
    c0 = rng.integers(0, 256, n).astype(np.uint32)
    c1 = rng.integers(0, 256, n).astype(np.uint32)
    w0 = rng.random(n, dtype=np.float32) * np.float32(2.0) ** rng.integers(-20, 1, n).astype(np.float32)
    w1 = rng.random(n, dtype=np.float32)
    wz = rng.random(n, dtype=np.float32)

    def fma32(a, b, acc):  # exact product + one rounding (operands are binary32, so float64 holds the product)
        return (a.astype(np.float64) * b.astype(np.float64) + acc.astype(np.float64)).astype(np.float32)

    ref = fma32(w1 * wz, c1.astype(np.float32), (w0 * wz) * c0.astype(np.float32)) * np.float32(1.0 / 255.0)
    wzs = wz * np.float32(2.0**120)
    t0 = (c0 << np.uint32(16)).view(np.float32)
    t1 = (c1 << np.uint32(16)).view(np.float32)
    with np.errstate(under="ignore"):
        prod0 = ((w0 * wzs).astype(np.float64) * t0.astype(np.float64)).astype(np.float32)
        got = fma32(w1 * wzs, t1, prod0) * (np.float32(1.0 / 255.0) * np.float32(8192.0))
    # float64 addition of a 48-bit product and a 24-bit addend can itself round; compare where it did not matter
    assert (got == ref).mean() > 0.9999


def test_magic_constant_floor_equals_floorf():
    """lin_axes_xy<MAGIC> / the god-ray tap: t = RD(u + 1.5 * 2^23) holds floor(u) in its low mantissa bits (two's complement
    below the constant's own bits, so `& (n - 1)` is the REPEAT wrap) and t - 1.5 * 2^23 == floorf(u), for |u| < 2^22."""
    rng = np.random.default_rng(5)
    u = np.concatenate([
        (rng.random(2_000_000) * 2.0 - 1.0) * 2.0 ** rng.integers(-10, 22, 2_000_000),   # all magnitudes, both signs
        np.arange(-4096, 4096, dtype=np.float64), np.arange(-4096, 4096, dtype=np.float64) + 0.5,
        np.nextafter(np.arange(-512, 512, dtype=np.float32), np.float32(-np.inf)).astype(np.float64),
        np.array([-0.0, 0.0, 2.0**22 - 0.5, -(2.0**22) + 0.5]),
    ]).astype(np.float32)
    u = u[(u == 0) | (np.abs(u) >= 2.0**-20)]                    # (binary64 holds u + magic exactly only down to 2^-29; the hardware's sum is exact)
    magic = np.float32(12582912.0)
    exact = u.astype(np.float64) + float(magic)                 # exact in binary64
    t = exact.astype(np.float32)                                # round to nearest ...
    t = np.where(t.astype(np.float64) > exact, np.nextafter(t, np.float32(-np.inf)), t).astype(np.float32)  # ... made round-down
    fl = np.floor(u)
    assert np.array_equal(t - magic, fl)                        # the subtraction is exact
    low = t.view(np.uint32).astype(np.int64) - 0x4B400000       # the mantissa field relative to the constant's bits
    assert np.array_equal(low, fl.astype(np.int64))
    for n in (32, 128):
        assert np.array_equal(t.view(np.uint32) & np.uint32(n - 1), fl.astype(np.int64) & (n - 1))


def test_rf_lerp_filter_is_within_guard_of_the_exact_filter():
    """rf_filter (mt_tex.cuh, MT_CONE_LERP): seven nested lerps on the denormal-scaled (r, F) words.  Its distance from the
    real-number trilinear filter must stay far inside MT_RF_GUARD / 4 = 2e-6 (the guard band of the light-cone decisions)."""
    rng = np.random.default_rng(1)
    n = 200_000
    r = rng.integers(0, 256, (n, 8)).astype(np.uint32)
    F = rng.integers(0, 2041, (n, 8)).astype(np.uint32)
    f = rng.random((n, 3)).astype(np.float32)
    rw = (r << np.uint32(16)).view(np.float32)      # MT_B3: r * 2^-133
    Fw = (F << np.uint32(13)).view(np.float32)      # MT_F13: F * 2^-136

    def lerp(a, b, t):  # fma(t, b - a, a): the difference is exact, one rounding in the fma
        with np.errstate(under="ignore"):
            d = (b - a).astype(np.float32)
            return (t.astype(np.float64) * d.astype(np.float64) + a.astype(np.float64)).astype(np.float32)

    def tri(w):
        l00, l01 = lerp(w[:, 0], w[:, 1], f[:, 0]), lerp(w[:, 2], w[:, 3], f[:, 0])
        l10, l11 = lerp(w[:, 4], w[:, 5], f[:, 0]), lerp(w[:, 6], w[:, 7], f[:, 0])
        return lerp(lerp(l00, l01, f[:, 1]), lerp(l10, l11, f[:, 1]), f[:, 2])

    def exact(v):
        v = v.astype(np.float64)
        fx, fy, fz = (f[:, i].astype(np.float64) for i in range(3))
        l00, l01 = v[:, 0] + fx * (v[:, 1] - v[:, 0]), v[:, 2] + fx * (v[:, 3] - v[:, 2])
        l10, l11 = v[:, 4] + fx * (v[:, 5] - v[:, 4]), v[:, 6] + fx * (v[:, 7] - v[:, 6])
        m0, m1 = l00 + fy * (l01 - l00), l10 + fy * (l11 - l10)
        return m0 + fz * (m1 - m0)

    k = np.float32(2.0**133 / 255.0)
    assert np.abs(tri(rw) * k - exact(r) / 255.0).max() < 4e-7
    assert np.abs(tri(Fw) * k - exact(F) / 2040.0).max() < 4e-7
