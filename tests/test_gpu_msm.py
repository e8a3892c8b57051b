"""GPU parity of the MSM path, through the C ABI, against the oracle.

Mirrors the reference's own checks: src/msm.test.ts:22-82 (msm == bigint msm for n in 0..12 on each
curve), scripts/zprize23/submission-test*.ts (known-answer identities), src/bigint/msm.test.ts
(algebraic identities), plus the size-independent closed form for known-dlog points at large N.
Everything is bit-exact: canonical affine coordinates and the is-zero flag."""
import numpy as np
import pytest

import montgomery_b200 as m
from montgomery_b200 import inputs
from oracle.params import KAT_BLS12_377_POINT, KAT_ED377_POINT
from tests.helpers import OracleCurve, points_to_bytes, scalars_to_bytes

pytestmark = pytest.mark.gpu
CURVES = {"bls12-377": m.curves.BLS12_377, "pallas": m.curves.PALLAS, "ed-on-bls12-377": m.curves.ED_ON_BLS12_377,
          "bls12-381": m.curves.BLS12_381}


@pytest.fixture(scope="module")
def engines():
    cache = {}

    def get(label, n):
        key = label
        if key not in cache or cache[key].max_points < n:
            if key in cache:
                cache[key].close()
            cache[key] = m.MsmEngine(CURVES[label], 0, max(n, 1 << 14))
        return cache[key]

    yield get
    for e in cache.values():
        e.close()


def _points_from_engine(eng, n):
    xy, z = eng.get_points(0, n)
    cb = eng.curve.coord_bytes
    return [None if f else (int.from_bytes(r[:cb].tobytes(), "little"), int.from_bytes(r[cb:].tobytes(), "little")) for r, f in zip(xy, z)]


@pytest.mark.parametrize("label", list(CURVES))
def test_golden_vectors(engines, golden, label):
    g = golden[label]
    eng = engines(label, 64)
    pts = [(int(x, 16), int(y, 16)) for x, y in g["points"]]
    sc = [int(s, 16) for s in g["scalars"]]
    xy, _ = points_to_bytes(pts, eng.curve.coord_bytes)
    eng.set_points(xy)
    for n, exp in g["results"].items():
        n = int(n)
        res, _ = eng.msm(scalars_to_bytes(sc[:n]), n=n)
        assert not res["isZero"]
        assert [hex(res["x"]), hex(res["y"])] == exp, (label, n)
        for c in (5, 7, 11):   # explicit window sizes (the reference's {c} option)
            res2, _ = eng.msm(scalars_to_bytes(sc[:n]), n=n, c=c)
            assert res2 == res, (label, n, c)


@pytest.mark.parametrize("label", list(CURVES))
def test_random_points_generator_matches_oracle(engines, label):
    eng = engines(label, 64)
    O = OracleCurve(label)
    eng.random_points(16, seed=99)
    a = inputs.known_dlogs(99, 16)
    pts = _points_from_engine(eng, 16)
    for i in (0, 1, 7, 15):
        assert pts[i] == O.scale(int(a[i]), O.G)


@pytest.mark.parametrize("label", list(CURVES))
@pytest.mark.parametrize("logn", [0, 1, 2, 3, 5, 8, 10])
def test_msm_vs_oracle_small(engines, label, logn):
    # src/msm.test.ts:39-41,65-82
    n = 1 << logn
    eng = engines(label, n)
    O = OracleCurve(label)
    eng.random_points(n, seed=1000 + logn)
    pts = _points_from_engine(eng, n)
    sc = inputs.random_scalars(O.q, n, seed=2000 + logn)
    res, tm = eng.msm(sc, n=n)
    assert res == O.msm(inputs.scalars_to_ints(sc), pts), (label, n, tm)


@pytest.mark.parametrize("label", list(CURVES))
def test_ragged_and_empty(engines, label):
    eng = engines(label, 1 << 10)
    O = OracleCurve(label)
    eng.random_points(777, seed=5)
    pts = _points_from_engine(eng, 777)
    sc = inputs.random_scalars(O.q, 777, seed=6)
    for n in (0, 3, 100, 777):     # n = 0: neutral element; n not a power of two; n < number of points set
        res, _ = eng.msm(sc[:n], n=n)
        assert res == O.msm(inputs.scalars_to_ints(sc[:n]), pts[:n]), (label, n)


def test_kat_bls12_377(engines):
    # scripts/zprize23/submission-test-bls377.ts: 2P + (q-1)P = P ; 1000 x same point
    O = OracleCurve("bls12-377")
    compute_msm = m.make_compute_msm(m.curves.BLS12_377)
    P = {"x": KAT_BLS12_377_POINT[0], "y": KAT_BLS12_377_POINT[1], "isZero": False}
    r = compute_msm([P, P], [2, O.q - 1])
    assert (r["x"], r["y"], r["isZero"]) == (P["x"], P["y"], False)
    sc = inputs.scalars_to_ints(inputs.random_scalars(O.q, 1000, 77))
    r2 = compute_msm([P] * 1000, sc)
    r3 = compute_msm([P], [sum(sc) % O.q])
    assert r2 == r3 == O.result_of(O.scale(sum(sc), KAT_BLS12_377_POINT))


def test_kat_ed377(engines):
    # scripts/zprize23/submission-test.ts
    O = OracleCurve("ed-on-bls12-377")
    compute_msm = m.make_compute_msm(m.curves.ED_ON_BLS12_377)
    x, y, _ = KAT_ED377_POINT
    pts = np.frombuffer((x.to_bytes(32, "little") + y.to_bytes(32, "little")) * 2, dtype=np.uint8)
    sc = np.frombuffer((2).to_bytes(32, "little") + (O.q - 1).to_bytes(32, "little"), dtype=np.uint8)
    r = compute_msm(pts.tobytes(), sc.tobytes())
    assert (r["x"], r["y"]) == (x, y)


@pytest.mark.parametrize("label", list(CURVES))
def test_degenerate_inputs(engines, label):
    """Duplicated points, P and -P, zero scalars, scalar q-1, all-equal scalars, sum to zero."""
    eng = engines(label, 1 << 10)
    O = OracleCurve(label)
    q = O.q
    G = O.G
    P1, P2 = O.scale(12345, G), O.scale(99999, G)
    neg = (lambda P: (P[0], (-P[1]) % O.prm.p)) if O.kind == "weierstrass" else (lambda P: ((-P[0]) % O.prm.p, P[1]))
    cases = [
        ([P1, P1, P1, P1], [1, 1, 1, 1]),
        ([P1, neg(P1), P2], [5, 5, 0]),                       # cancels to the neutral element
        ([P1, P2, P1, P2], [0, 0, 0, 0]),                     # all-zero scalars
        ([P1, P2, neg(P1), neg(P2)], [q - 1, q - 1, q - 1, q - 1]),
        ([P1] * 64, [7] * 64),                                # one bucket holds everything
        ([P1, P1], [3, q - 3]),                               # sum of scalars = 0 mod q
        ([P1, P2] * 32, list(range(1, 65))),
    ]
    for pts, sc in cases:
        xy, z = points_to_bytes(pts, eng.curve.coord_bytes)
        eng.set_points(xy)
        res, _ = eng.msm(scalars_to_bytes(sc), n=len(sc))
        assert res == O.msm(sc, pts), (label, sc[:4])
    if O.kind == "weierstrass":   # points at infinity among the inputs (isNonZero = 0 in the reference layout)
        pts = [P1, None, P2, None]
        xy, z = points_to_bytes(pts, eng.curve.coord_bytes)
        eng.set_points(xy, z)
        res, _ = eng.msm(scalars_to_bytes([3, 4, 5, 6]), n=4)
        assert res == O.msm([3, 5], [P1, P2])


@pytest.mark.parametrize("label,logn", [("bls12-377", 14), ("bls12-377", 16), ("pallas", 16), ("ed-on-bls12-377", 16),
                                        ("bls12-377", 20), ("pallas", 18), ("ed-on-bls12-377", 18), ("bls12-381", 16)])
def test_closed_form_large(engines, label, logn):
    """Known-dlog points P_i = a_i G: result must equal [(sum s_i a_i) mod q] G."""
    n = 1 << logn
    eng = engines(label, n)
    O = OracleCurve(label)
    eng.random_points(n, seed=31337 + logn)
    a = inputs.known_dlogs(31337 + logn, n)
    sc = inputs.random_scalars(O.q, n, seed=4242 + logn)
    res, tm = eng.msm(sc, n=n)
    s_ints = inputs.scalars_to_ints(sc)
    k = sum(s * int(ai) for s, ai in zip(s_ints, a)) % O.q
    assert res == O.result_of(O.scale(k, O.G)), (label, logn, tm)
    # linearity in the scalars: msm(s + s') = msm(s) + msm(s') is implied by the closed form; check
    # additionally that a permuted pairing changes the answer (guards against ignoring the scalars)
    res_dev, _ = eng.msm(sc[::-1].copy(), n=n)
    k2 = sum(s * int(ai) for s, ai in zip(reversed(s_ints), a)) % O.q
    assert res_dev == O.result_of(O.scale(k2, O.G))


@pytest.mark.parametrize("label", list(CURVES))
def test_all_window_sizes(engines, label):
    """Every window size c (sparse top windows, clipped reduction digits, tiny bucket counts) gives the
    same point -- the reference exposes c as an option (msm-batched-affine.ts:74-77)."""
    n = 300
    eng = engines(label, 1 << 10)
    O = OracleCurve(label)
    eng.random_points(n, seed=777)
    pts = _points_from_engine(eng, n)
    sc = inputs.random_scalars(O.q, n, seed=778)
    exp = O.msm(inputs.scalars_to_ints(sc), pts)
    for c in range(2, 23):
        res, tm = eng.msm(sc, n=n, c=c)
        assert res == exp, (label, c, tm)


@pytest.mark.parametrize("label", ["bls12-377", "pallas", "bls12-381"])
def test_msm_projective_equals_batched_affine(label):
    """src/msm.test.ts:73-82: `msmProjective` (msm-basic, no GLV) == `msmUnsafe` (batched affine + GLV),
    through the reference-shaped host API."""
    mod = m.Weierstrass.create(CURVES[label])
    O = OracleCurve(label)
    for logn in (0, 3, 7, 12):
        n = 1 << logn
        pts = mod.Parallel.randomPointsFast(n, seed=60 + logn)
        sc = mod.Parallel.randomScalars(n, seed=61 + logn)
        a = mod.Parallel.msmUnsafe(sc, pts, n)["result"]
        b = mod.Parallel.msmProjective(sc, pts, n)["result"]
        assert a == b, (label, logn)
        if logn <= 7:
            P = [None if q["isZero"] else (q["x"], q["y"]) for q in pts.toBigints()]
            assert a == O.msm(inputs.scalars_to_ints(sc), P)


@pytest.mark.parametrize("label", ["bls12-377", "ed-on-bls12-377"])
def test_extreme_bucket_skew(engines, label):
    """All scalars equal: every point of a window lands in ONE bucket of 2^14 elements, so the bucket
    trees run deep (the engine leaves at most 16 elements of a bucket to the reduction: >= 10 rounds)
    instead of the usual ~log2(average).  Expected: s * sum(P_i) = s * (sum a_i) G."""
    n = 1 << 14
    eng = engines(label, n)
    O = OracleCurve(label)
    eng.random_points(n, seed=4321)
    a = inputs.known_dlogs(4321, n)
    for s in (7, O.q - 2):
        sc = scalars_to_bytes([s] * n)
        res, tm = eng.msm(sc, n=n)
        assert tm["rounds"] >= 10
        assert res == O.result_of(O.scale(s * sum(int(v) for v in a), O.G)), (label, s, tm)
