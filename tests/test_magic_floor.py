"""The round-toward-minus-infinity "magic number" floor the FAST kernels use (fb_kernels_fast.cu: __fadd_rd(t, MAGIC);
fb_render.cu rtex_axis under -DFB_RENDER_MAGIC_FLOOR=1), emulated on the host: for -2^22 <= t < 2^22 the fp32 sum
t + 1.5 * 2^23 rounded DOWN equals floor(t) + 1.5 * 2^23 exactly, so both the float floor and the integer in the low
mantissa bits agree with floorf / (int)floorf — which is what makes those paths bit-identical to tex_axis()
(shaders: the sampler's texel index and fraction, src/precompute.rs:85-98)."""
import numpy as np

MAGIC = np.float32(12582912.0)          # 1.5 * 2^23
MAGIC_BITS = 0x4B400000


def fadd_rd(t: np.ndarray, m: np.float32) -> np.ndarray:
    """fp32 addition rounded toward minus infinity, for |t| < 2^24: the exact sum fits a double unless t is tiny, and a
    tiny t only decides on which side of m the sum lies."""
    exact = t.astype(np.float64) + np.float64(m)
    near = exact.astype(np.float32)
    too_high = near.astype(np.float64) > exact
    # doubles swallow |t| < 2^-29: the exact sum is then m + (tiny), below m iff t < 0
    too_high |= (near.astype(np.float64) == exact) & (np.abs(t) < 2.0 ** -28) & (t < 0) & (exact == np.float64(m))
    return np.where(too_high, np.nextafter(near, np.float32(-np.inf)), near).astype(np.float32)


def cases() -> np.ndarray:
    rng = np.random.default_rng(7)
    parts = [rng.uniform(-2.0 ** 22, 2.0 ** 22, 200000), rng.uniform(-300, 300, 200000), rng.uniform(-2, 2, 100000),
             np.arange(-1024, 1025, dtype=np.float64), np.arange(-1024, 1025, dtype=np.float64) + 0.5,
             np.array([0.0, -0.5, 0.5, 1e-30, -1e-30, 1e-45, -1e-45, 2.0 ** 22 - 0.5, -(2.0 ** 22), 4194303.5, -4194303.5])]
    t = np.concatenate(parts).astype(np.float32)
    ints = np.arange(-4096, 4097, dtype=np.float32)
    edges = np.concatenate([np.nextafter(ints, np.float32(np.inf)), np.nextafter(ints, np.float32(-np.inf))])
    return np.concatenate([t, edges])


def test_magic_floor_matches_floorf_bit_for_bit():
    t = cases()
    s = fadd_rd(t, MAGIC)
    fl = (s - MAGIC).astype(np.float32)                      # exact: both are integers below 2^24
    want = np.floor(t).astype(np.float32)
    assert np.array_equal(fl, want)
    frac_magic = (t - fl).astype(np.float32)
    frac_ref = (t - want).astype(np.float32)
    assert np.array_equal(frac_magic.view(np.uint32), frac_ref.view(np.uint32))
    i = s.view(np.int32) - np.int32(MAGIC_BITS)
    assert np.array_equal(i, want.astype(np.int32))


def test_magic_floor_infinities_clamp_like_tex_axis():
    """t = +inf / -inf: tex_axis clamps floor(t) to [-1, n] before the conversion, then to [0, n - 1]; the magic form
    clamps the (garbage but correctly signed) integer — same indices."""
    n = 32
    for t, want in ((np.float32(np.inf), (n - 1, n - 1)), (np.float32(-np.inf), (0, 0))):
        s = np.float32(t + MAGIC)
        i = int(np.array([s]).view(np.int32)[0]) - MAGIC_BITS
        i0 = min(max(i, 0), n - 1)
        i1 = min(max(i + 1, 0), n - 1)
        assert (i0, i1) == want
