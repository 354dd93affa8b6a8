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
