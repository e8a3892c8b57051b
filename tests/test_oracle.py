"""CPU tests pinning the oracle to everything the reference holds for the MSM path."""
import random

import pytest

from oracle.glv import GlvScalar, signed_digits, window_size
from oracle.msm import msm, msm_naive
from oracle.params import BLS12_377, BLS12_381, ED_ON_BLS12_377, KAT_BLS12_377_POINT, KAT_ED377_POINT, PALLAS
from oracle.twisted_edwards import TwistedEdwardsCurve
from oracle.weierstrass import AffineCurve, ProjectiveCurve


@pytest.mark.parametrize("prm", [BLS12_377, PALLAS, BLS12_381], ids=lambda p: p.label)
def test_generator_on_curve_and_subgroup(prm):
    # src/bigint/curves.test.ts:20-51
    A = AffineCurve(prm)
    assert A.is_on_curve(prm.G) and A.is_in_subgroup(prm.G)
    P = ProjectiveCurve(prm)
    assert P.to_affine(P.scale(prm.q - 1, P.one)) == A.negate(prm.G)   # (q-1) P = -P


def test_ed377_generator():
    T = TwistedEdwardsCurve(ED_ON_BLS12_377)
    assert T.is_on_curve(T.one) and T.is_in_subgroup(T.one)
    assert T.is_equal(T.scale(T.q - 1, T.one), T.negate(T.one))


def test_kat_bls12_377():
    # scripts/zprize23/submission-test-bls377.ts:6-26 : msm([P, P], [2, q-1]) == P
    A, P = AffineCurve(BLS12_377), ProjectiveCurve(BLS12_377)
    pt = KAT_BLS12_377_POINT
    assert A.is_on_curve(pt) and A.is_in_subgroup(pt)
    r = P.to_affine(msm(P, [2, BLS12_377.q - 1], [P.from_affine(pt)] * 2))
    assert r == pt
    # :28-45 : same point, random scalars == (sum s) * P
    rnd = random.Random(1)
    sc = [rnd.randrange(BLS12_377.q) for _ in range(100)]
    r2 = P.to_affine(msm(P, sc, [P.from_affine(pt)] * 100))
    assert r2 == P.to_affine(P.scale(sum(sc) % BLS12_377.q, P.from_affine(pt)))


def test_kat_ed377():
    # scripts/zprize23/submission-test.ts:5-21
    T = TwistedEdwardsCurve(ED_ON_BLS12_377)
    x, y, t = KAT_ED377_POINT
    pt = (x, y, 1, t)
    assert T.is_on_curve(pt)
    assert T.to_affine(msm(T, [2, T.q - 1], [pt, pt])) == (x, y)


@pytest.mark.parametrize("label", ["bls12-377", "pallas", "ed-on-bls12-377", "bls12-381"])
def test_msm_identities(label):
    # src/bigint/msm.test.ts:18-59
    rnd = random.Random(7)
    if label == "ed-on-bls12-377":
        C = TwistedEdwardsCurve(ED_ON_BLS12_377)
        eq, zero = C.is_equal, C.zero
    else:
        C = ProjectiveCurve({"bls12-377": BLS12_377, "pallas": PALLAS, "bls12-381": BLS12_381}[label])
        eq, zero = C.is_equal, C.zero
    q = C.q
    n = 12
    sc = [rnd.randrange(q) for _ in range(n)]
    pts = [C.scale(rnd.randrange(1, q), C.one) for _ in range(n)]
    # msm(s_i, P) = (sum s_i) P
    assert eq(msm(C, sc, [pts[0]] * n), C.scale(sum(sc) % q, pts[0]))
    # msm([...s, -sum s], P) = 0
    assert eq(msm(C, sc + [(-sum(sc)) % q], [pts[0]] * (n + 1)), zero)
    # msm(s, P_i) = s * sum P_i
    tot = zero
    for Pt in pts:
        tot = C.add(tot, Pt)
    assert eq(msm(C, [sc[0]] * n, pts), C.scale(sc[0], tot))
    # Pippenger == defining sum
    assert eq(msm(C, sc, pts), msm_naive(C, sc, pts))


def test_golden_vectors_match_oracle(golden):
    for label, C in (("bls12-377", ProjectiveCurve(BLS12_377)), ("pallas", ProjectiveCurve(PALLAS)), ("bls12-381", ProjectiveCurve(BLS12_381))):
        g = golden[label]
        pts = [(int(x, 16), int(y, 16), 1) for x, y in g["points"]]
        sc = [int(s, 16) for s in g["scalars"]]
        for n, exp in g["results"].items():
            n = int(n)
            r = C.to_affine(msm(C, sc[:n], pts[:n]))
            assert [hex(r[0]), hex(r[1])] == exp
    T = TwistedEdwardsCurve(ED_ON_BLS12_377)
    g = golden["ed-on-bls12-377"]
    pts = [T.from_affine((int(x, 16), int(y, 16))) for x, y in g["points"]]
    sc = [int(s, 16) for s in g["scalars"]]
    r = T.to_affine(msm(T, sc[:16], pts[:16]))
    assert [hex(r[0]), hex(r[1])] == g["results"]["16"]


@pytest.mark.parametrize("prm", [BLS12_377, PALLAS, BLS12_381], ids=lambda p: p.label)
def test_glv_decomposition(prm):
    # src/scalar-glv.ts:92-103 : s0 + s1*lambda = s (mod q), halves below maxBits
    g = GlvScalar(prm.q, prm.lam)
    rnd = random.Random(3)
    for s in [0, 1, prm.q - 1] + [rnd.randrange(prm.q) for _ in range(2000)]:
        s0, s1 = g.decompose(s)
        assert (s0 + s1 * prm.lam - s) % prm.q == 0
        assert abs(s0).bit_length() <= g.max_bits and abs(s1).bit_length() <= g.max_bits
    # the endomorphism really is multiplication by lambda
    A = AffineCurve(prm)
    assert A.scale(prm.lam, prm.G) == (prm.beta * prm.G[0] % prm.p, prm.G[1])


def test_window_table_and_signed_digits():
    assert window_size(377, 16) == 14 and window_size(377, 20) == 18 and window_size(377, 24) == 23
    assert window_size(255, 16) == 12 and window_size(255, 18) == 17
    rnd = random.Random(5)
    for c in (5, 13, 16, 18):
        K = -(-(127 + 1) // c)
        L = 1 << (c - 1)
        for _ in range(200):
            s = rnd.randrange(1 << 127)
            d = signed_digits(s, c, K)
            assert all(0 <= l <= L for l, _ in d)
            assert sum((-l if neg else l) << (c * k) for k, (l, neg) in enumerate(d)) == s


def test_cpu_restatement_matches_python_oracle(golden):
    """oracle/msm_cpu.cpp (the CPU baseline / reference-algorithm port) against the bigint oracle."""
    import numpy as np
    from oracle import cpu_ref
    from tests.helpers import OracleCurve, points_to_bytes, scalars_to_bytes
    for label, cb in (("bls12-377", 48), ("pallas", 32), ("ed-on-bls12-377", 32), ("bls12-381", 48)):
        g = golden[label]
        pts = [(int(x, 16), int(y, 16)) for x, y in g["points"]]
        sc = [int(s, 16) for s in g["scalars"]]
        xy, _ = points_to_bytes(pts, cb)
        for n, exp in g["results"].items():
            n = int(n)
            for threads, c in ((1, 0), (3, 0), (8, 5)):
                r, _ = cpu_ref.msm(label, scalars_to_bytes(sc[:n]), xy, n, threads=threads, c=c)
                assert [hex(r["x"]), hex(r["y"])] == exp, (label, n, threads, c)
        # seeded known-dlog generator == a_i * G, and MSM over it == closed form
        from montgomery_b200 import inputs
        O = OracleCurve(label)
        n = 512
        pb = cpu_ref.known_dlog_points(label, 5, n, threads=2)
        a = inputs.known_dlogs(5, n)
        P0 = O.scale(int(a[0]), O.G)
        assert (int.from_bytes(pb[0, :cb].tobytes(), "little"), int.from_bytes(pb[0, cb:].tobytes(), "little")) == P0
        scb = inputs.random_scalars(O.q, n, 6)
        r, _ = cpu_ref.msm(label, scb, pb, n, threads=4)
        k = sum(s * int(ai) for s, ai in zip(inputs.scalars_to_ints(scb), a)) % O.q
        assert r == O.result_of(O.scale(k, O.G))
