"""CPU restatements (numpy) of the exact arithmetic shortcuts the round-2 kernels use — the claims DESIGN.md section 4 makes about
`k_granule`, checked independently of the device (the exhaustive device-side proof is `mp3b_selftest`,
tests/test_gpu_parity.py::test_exact_arithmetic_shortcuts):

  * widen_normal: float -> double of a positive normal number by integer operations (csrc/kernels.cu);
  * quant30m / pair_index: floor(min(t, 30.5)) as `t + 2^23` rounded down, the pair's table index from the bit patterns;
  * div192: x / 192 = (x / 3) / 64 whenever the quotient is normal;
  * gain_from_peak: an estimate + a walk on the exact threshold table lands on the index the bisection finds;
  * lane_tree_pair: two butterfly sums sharing their shuffles give the butterfly's bits.
"""
import numpy as np


def _normal_positive_floats(n, seed):
    rng = np.random.default_rng(seed)
    bits = rng.integers(0x00800000, 0x7F800000, size=n, dtype=np.uint32)      # exponent field 1 ... 254: every positive normal float
    edge = np.array([0x00800000, 0x7F7FFFFF, 0x3F800000, 0x2EDBE6FF, 0x2EDBE700, 0x00FFFFFF], np.uint32)   # FLT_MIN, FLT_MAX, 1, ~1e-10
    return np.concatenate([bits, edge])


def test_widen_normal_is_the_exact_conversion():
    b = _normal_positive_floats(2_000_000, 1)
    hi = (b >> np.uint32(3)) + np.uint32(0x38000000)
    lo = b << np.uint32(29)                                                     # (mod 2^32)
    wide = ((hi.astype(np.uint64) << np.uint64(32)) | lo.astype(np.uint64)).view(np.float64)
    assert np.array_equal(wide, b.view(np.float32).astype(np.float64))


def test_magic_floor_and_pair_index():
    # bit pattern of 2^23 + k as a float is 0x4B000000 + k for 0 <= k < 2^23: the low bits ARE floor(t)
    k = np.arange(0, 31, dtype=np.uint32)
    as_float = (np.float32(8388608.0) + k.astype(np.float32))
    assert np.array_equal(as_float.view(np.uint32), np.uint32(0x4B000000) + k)
    # t + 2^23 rounded DOWN = 2^23 + floor(t) for 0 <= t < 2^23 (the ulp there is 1): in exact arithmetic
    rng = np.random.default_rng(2)
    t = np.concatenate([rng.uniform(0, 40, 1_000_000), np.array([0.0, 0.999999, 1.0, 29.9999, 30.0, 30.49, 30.5, 1e9, np.inf])]).astype(np.float32)
    clamped = np.minimum(t, np.float32(30.5))
    exact = clamped.astype(np.float64) + 8388608.0                             # exact in double
    rd = np.floor(exact)                                                       # round down to the float grid (spacing 1)
    u = (rd - 8388608.0).astype(np.int64)
    assert np.array_equal(u, np.minimum(np.floor(t.astype(np.float64)), 30).astype(np.int64))
    # index of a pair from the two bit patterns: ((ux << 5) + uy) & 1023 = 32 qx + qy
    qx, qy = np.meshgrid(k, k, indexing="ij")
    ux, uy = np.uint32(0x4B000000) + qx, np.uint32(0x4B000000) + qy
    idx = ((ux << np.uint32(5)) + uy) & np.uint32(1023)
    assert np.array_equal(idx, 32 * qx + qy)


def test_div192_is_div3_then_an_exact_scaling():
    rng = np.random.default_rng(3)
    x = np.concatenate([np.exp(rng.uniform(np.log(1e-30), np.log(1e30), 2_000_000)), np.array([1e-30, 192.0, 3.0, 1.0, 3.4e38])]).astype(np.float32)
    x = x[x >= np.float32(1e-30)]
    a = (x / np.float32(3.0)) * np.float32(0.015625)
    assert np.array_equal(a.view(np.uint32), (x / np.float32(192.0)).view(np.uint32))


def test_gain_walk_equals_bisection():
    thr = np.exp2((np.arange(256) - 210) / 4.0)                                 # c_gain_thr: 2^((g - 210) / 4) in double
    rng = np.random.default_rng(4)
    ratio = np.concatenate([np.exp(rng.uniform(np.log(thr[0]), np.log(thr[255] * 4), 200_000)), thr, np.nextafter(thr, 0), np.nextafter(thr, np.inf)]).astype(np.float32)
    ratio = ratio[ratio.astype(np.float64) >= thr[0]]
    r = ratio.astype(np.float64)
    want = np.searchsorted(thr, r, side="right") - 1                            # largest index with thr[i] <= r
    for skew in (0.0, -1.7, +2.3):                                              # "any estimate would do": also from a bad one
        lo = np.clip(np.floor(4.0 * np.log2(r) + 210.0 + skew), 0, 255).astype(np.int64)
        for _ in range(8):                                                      # the walk (the kernel's loops run 0 or 1 step)
            up = (lo < 255) & (thr[np.minimum(lo + 1, 255)] <= r)
            lo = np.where(up, lo + 1, lo)
        for _ in range(8):
            down = (lo > 0) & (thr[lo] > r)
            lo = np.where(down, lo - 1, lo)
        assert np.array_equal(lo, want)
    est = np.clip(np.floor(4.0 * np.log2(r) + 210.0), 0, 255).astype(np.int64)
    assert np.abs(est - want).max() <= 1                                        # the honest estimate is never more than one step off


def _butterfly(p):
    p = p.copy()
    for m in (16, 8, 4, 2, 1):
        p = p + p[np.arange(32) ^ m]
    return p


def test_lane_tree_pair_gives_the_butterflys_bits():
    rng = np.random.default_rng(5)
    lanes = np.arange(32)
    for _ in range(2000):
        a = (rng.standard_normal(32) * 10.0 ** rng.uniform(-6, 6)).astype(np.float32) ** 2
        b = (rng.standard_normal(32) * 10.0 ** rng.uniform(-6, 6)).astype(np.float32) ** 2
        up = (lanes & 16) != 0
        keep = np.where(up, b, a)
        give = np.where(up, a, b)
        keep = keep + give[lanes ^ 16]
        for m in (8, 4, 2, 1):
            keep = keep + keep[lanes ^ m]
        ra, rb = _butterfly(a), _butterfly(b)
        assert len(set(ra.view(np.uint32))) == 1 and len(set(rb.view(np.uint32))) == 1     # every lane ends with the same bits
        assert keep[0].view(np.uint32) == ra[0].view(np.uint32) and keep[16].view(np.uint32) == rb[0].view(np.uint32)


def test_pow34_newton_step_truncation_error():
    """|x|^0.75 = d * q with q = q0 (1 + e / 4 + 5 e^2 / 32), e = 1 - d q0^4: in exact arithmetic, a seed q0 within 2^-19 of
    d^-1/4 (the device's two rsqrt approximations give about 2^-21) leaves a relative error below 2^-52 — what remains on the
    device is the rounding of its eight FP64 operations, and whether that ever reaches the float is what mp3b_selftest decides
    exhaustively."""
    from decimal import Decimal, getcontext
    getcontext().prec = 60
    rng = np.random.default_rng(6)
    worst = Decimal(0)
    for _ in range(1500):
        d = Decimal(float(np.float32(np.exp(rng.uniform(np.log(1e-10), np.log(1e10))))))
        true_q = 1 / d.sqrt().sqrt()
        for rel in (Decimal(2) ** -19, -(Decimal(2) ** -19), Decimal(float(rng.uniform(-1, 1))) * Decimal(2) ** -21):
            q0 = true_q * (1 + rel)
            e = 1 - d * q0 ** 4
            q = q0 * (1 + e / 4 + 5 * e * e / 32)
            worst = max(worst, abs(q / true_q - 1))
    assert worst < Decimal(2) ** -52, worst
