"""The conservative slab test on 16-bit box codes (csrc/vkrt_device.cuh qray_setup / qslab_test, DESIGN.md
"Conservative coded boxes"), restated with exact rational arithmetic and explicitly rounded binary32 operations.

Claim under test: for every ray and every exact float box inside a coded box, the coded test's entry bound is <=
the entry parameter the exact float slab test (slab_test) computes for the inner box, and its exit bound is >= the
exact test's exit parameter.  Then an ancestor is never rejected when a sphere's own exact test passes, which is
all rule S needs from the hierarchy.  No GPU involved: this checks the arithmetic of the design itself.
"""
from fractions import Fraction

import numpy as np
import pytest

F32 = np.float32


def _round(x, mode):
    """exact Fraction -> binary32 with rounding mode 'n' (nearest even), 'd' (down), 'u' (up)"""
    if x == 0:
        return F32(0.0)
    c = F32(float(x))                     # float64 nearest then float32 nearest: a candidate within 1 ulp
    lo = c if Fraction(float(c)) <= x else np.nextafter(c, F32(-np.inf))
    while Fraction(float(lo)) > x:
        lo = np.nextafter(lo, F32(-np.inf))
    while Fraction(float(np.nextafter(lo, F32(np.inf)))) <= x:
        lo = np.nextafter(lo, F32(np.inf))
    if Fraction(float(lo)) == x:
        return lo
    hi = np.nextafter(lo, F32(np.inf))
    if mode == 'd':
        return lo
    if mode == 'u':
        return hi
    dl, dh = x - Fraction(float(lo)), Fraction(float(hi)) - x
    if dl != dh:
        return lo if dl < dh else hi
    return lo if (lo.view(np.uint32) & 1) == 0 else hi


def fr(v):
    return Fraction(float(v))


def fma(a, b, c, mode='n'):
    return _round(fr(a) * fr(b) + fr(c), mode)


def mul(a, b, mode='n'):
    return _round(fr(a) * fr(b), mode)


def add(a, b, mode='n'):
    return _round(fr(a) + fr(b), mode)


def safe_inv(d):
    dd = d if abs(float(d)) > 1e-20 else F32(np.copysign(F32(1e-20), d))
    return _round(Fraction(1) / fr(dd), 'n')


def exact_slab(o, d, lo, hi):
    """vkrt_device.cuh::slab_setup + slab_test in binary32"""
    tn, tf = None, None
    for k in range(3):
        inv = safe_inv(d[k])
        oinv = mul(o[k], inv)
        t0, t1 = fma(lo[k], inv, -oinv), fma(hi[k], inv, -oinv)
        a, b = min(t0, t1), max(t0, t1)
        tn = a if tn is None else max(tn, a)
        tf = b if tf is None else min(tf, b)
    return tn, tf


def coded_slab(o, d, qlo, qhi, s, b2):
    """qray_setup + qslab_test"""
    tn, tf = None, None
    for k in range(3):
        inv = safe_inv(d[k])
        oinv = mul(o[k], inv)
        a = mul(s[k], inv)
        m = mul(F32(0.51), F32(abs(a)), 'u')
        cdn = add(fma(b2[k], inv, -oinv, 'd'), -m, 'd')
        cup = add(fma(b2[k], inv, -oinv, 'u'), m, 'u')
        near, far = (qhi[k], qlo[k]) if inv < 0 else (qlo[k], qhi[k])
        t_near = fma(F32(8388608 + near), a, cdn, 'd')
        t_far = fma(F32(8388608 + far), a, cup, 'u')
        tn = t_near if tn is None else max(tn, t_near)
        tf = t_far if tf is None else min(tf, t_far)
    return tn, tf


def quantize(x, s, b2, lower, margin=1):
    """k_quantize: the tightest code whose coordinate X(q) = (2^23 + q) s + b2 encloses x (exact arithmetic), moved
    `margin` steps outward (the builder uses 1 because it evaluates X in double; the bound must hold with 0 too)"""
    X = lambda q: Fraction(8388608 + q) * fr(s) + fr(b2)
    q = int((fr(x) - X(0)) / fr(s))
    q = min(max(q, 0), 65535)
    if lower:
        while q > 0 and X(q) > fr(x):
            q -= 1
        while q < 65535 and X(q + 1) <= fr(x):
            q += 1
        return max(q - margin, 0)
    while q < 65535 and X(q) < fr(x):
        q += 1
    while q > 0 and X(q - 1) >= fr(x):
        q -= 1
    return min(q + margin, 65535)


@pytest.mark.parametrize("margin", [1, 0])
def test_coded_slab_bounds_enclose_exact_slab_test(margin):
    rng = np.random.default_rng(11 + margin)
    checked = 0
    for trial in range(400):
        # a scene grid like k_qgrid's: 65024 steps across the scene box, 128 steps of margin
        lo_s = rng.uniform(-200, 50, 3).astype(F32)
        ext = rng.uniform(1.0, 300.0, 3).astype(F32)
        s = (ext / F32(65024.0)).astype(F32)
        base = (lo_s - F32(128.0) * s).astype(F32)
        b2 = np.array([fma(F32(-8388608.0), s[k], base[k]) for k in range(3)], dtype=F32)
        # an exact box inside the scene box and its codes
        c = (lo_s + rng.uniform(0.05, 0.95, 3).astype(F32) * ext).astype(F32)
        half = (rng.uniform(0.001, 0.05, 3).astype(F32) * ext).astype(F32)
        lo, hi = np.maximum(c - half, lo_s).astype(F32), np.minimum(c + half, lo_s + ext).astype(F32)
        qlo = [quantize(lo[k], s[k], b2[k], True, margin) for k in range(3)]
        qhi = [quantize(hi[k], s[k], b2[k], False, margin) for k in range(3)]
        for k in range(3):
            assert Fraction(8388608 + qlo[k]) * fr(s[k]) + fr(b2[k]) <= fr(lo[k])
            assert Fraction(8388608 + qhi[k]) * fr(s[k]) + fr(b2[k]) >= fr(hi[k])
        for r in range(6):
            kind = (trial + r) % 6
            o = (lo_s + rng.uniform(-0.2, 1.2, 3).astype(F32) * ext).astype(F32)
            d = rng.normal(size=3).astype(F32)
            if kind == 1:
                d[rng.integers(3)] = F32(0.0)                       # axis-parallel: |d| clamped to 1e-20
            elif kind == 2:
                d[rng.integers(3)] = F32(-0.0)
            elif kind == 3:
                d[rng.integers(3)] = F32(1e-30) * F32(rng.choice([-1.0, 1.0]))
            elif kind == 4:
                o = (c + (rng.uniform(-1, 1, 3) * half * 1.001).astype(F32)).astype(F32)   # origin at / inside the box
            elif kind == 5:
                k = rng.integers(3)
                o[k] = lo[k] if rng.random() < 0.5 else hi[k]      # origin exactly on a face plane
            d = (d / max(float(np.linalg.norm(d)), 1e-30)).astype(F32)
            tn_e, tf_e = exact_slab(o, d, lo, hi)
            tn_q, tf_q = coded_slab(o, d, qlo, qhi, s, b2)
            assert np.isfinite(tn_q) and np.isfinite(tf_q)
            assert tn_q <= tn_e, (trial, r, o, d, tn_q, tn_e)
            assert tf_q >= tf_e, (trial, r, o, d, tf_q, tf_e)
            checked += 1
    assert checked == 2400


def test_rounding_helper():
    third = Fraction(1, 3)
    d, n, u = _round(third, 'd'), _round(third, 'n'), _round(third, 'u')
    assert fr(d) < third < fr(u) and n in (d, u) and np.nextafter(d, F32(1)) == u
    assert _round(Fraction(1, 2), 'd') == F32(0.5) == _round(Fraction(1, 2), 'u')
    assert _round(-third, 'd') == -u and _round(-third, 'u') == -d
    # ties to even
    x = fr(F32(1.0)) + Fraction(1, 2 ** 24)
    assert _round(x, 'n') == F32(1.0)
