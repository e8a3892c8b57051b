"""GPU tests of the rest of the C ABI and of the reference-shaped host API (include/montgomery_b200.h):
point ingestion round trip, scalars already on the device, the partial / combine pair that the
multi-GPU path is made of (run here as two shards on one GPU), linearity at full size, state reuse
across calls of different sizes, error codes, and `compute_msm` (scripts/zprize23/submission-bls377.ts)."""
import numpy as np
import pytest
import torch

import montgomery_b200 as m
from montgomery_b200 import inputs
from montgomery_b200.api import MsmError as MgbError
from tests.helpers import OracleCurve, points_to_bytes, scalars_to_bytes

pytestmark = pytest.mark.gpu
CURVES = {"bls12-377": m.curves.BLS12_377, "pallas": m.curves.PALLAS, "ed-on-bls12-377": m.curves.ED_ON_BLS12_377,
          "bls12-381": m.curves.BLS12_381}


def _oracle_points(O, n, seed):
    a = inputs.known_dlogs(seed, n)
    return [O.scale(int(v), O.G) for v in a]


@pytest.mark.parametrize("label", list(CURVES))
def test_set_get_points_round_trip(label):
    """pointsFromBytes -> toBigint (src/parallel.ts:97-116, src/curve-affine.ts:220-233): canonical bytes come back
    unchanged, infinity flags are kept, a sub-range read is the same slice."""
    O = OracleCurve(label)
    cv = CURVES[label]
    pts = _oracle_points(O, 40, 7)
    if O.kind == "weierstrass":
        pts[3] = None
        pts[17] = None
    xy, z = points_to_bytes(pts, cv.coord_bytes)
    with_flags = O.kind == "weierstrass"
    eng = m.MsmEngine(cv, 0, 64)
    try:
        assert eng.set_points(xy, z if with_flags else None) == 40
        got, gz = eng.get_points(0, 40)
        assert np.array_equal(got.reshape(-1), xy)
        if with_flags:
            assert gz.tolist() == z.tolist()
        sub, sz = eng.get_points(10, 12)
        assert np.array_equal(sub, got[10:22]) and sz.tolist() == gz[10:22].tolist()
    finally:
        eng.close()


@pytest.mark.parametrize("label", ["bls12-377", "ed-on-bls12-377"])
def test_device_scalars_equal_host_scalars(label):
    """mgb_msm_device (scalars already in HBM) returns what mgb_msm returns for the same bytes."""
    cv = CURVES[label]
    n = 1 << 13
    eng = m.MsmEngine(cv, 0, n)
    try:
        eng.random_points(n, seed=99)
        sc = inputs.random_scalars(cv.q, n, 5)
        host, _ = eng.msm(sc, n=n)
        dev = torch.from_numpy(np.ascontiguousarray(sc)).cuda()
        torch.cuda.synchronize()
        on_dev, _ = eng.msm(None, n=n, device_ptr=dev.data_ptr())
        assert host == on_dev
    finally:
        eng.close()


@pytest.mark.parametrize("label", list(CURVES))
def test_partial_and_combine_equal_whole(label):
    """The multi-GPU decomposition on one GPU: two shards in two contexts, mgb_msm_partial each,
    mgb_combine_partials of the two accumulators == the MSM over all points."""
    cv = CURVES[label]
    n, cut = 3000, 1234
    whole = m.MsmEngine(cv, 0, n)
    a = m.MsmEngine(cv, 0, n)
    b = m.MsmEngine(cv, 0, n)
    try:
        whole.random_points(n, seed=31)
        xy, z = whole.get_points(0, n)
        flags = z if cv.kind == "weierstrass" else None
        a.set_points(xy[:cut].reshape(-1), None if flags is None else flags[:cut])
        b.set_points(xy[cut:].reshape(-1), None if flags is None else flags[cut:])
        sc = inputs.random_scalars(cv.q, n, 8)
        expect, _ = whole.msm(sc, n=n)
        pb = a.partial_bytes
        assert pb == b.partial_bytes and pb > 0
        parts = torch.zeros(2 * pb, dtype=torch.uint8, device="cuda")
        sa = torch.from_numpy(np.ascontiguousarray(sc[:cut])).cuda()
        sb = torch.from_numpy(np.ascontiguousarray(sc[cut:])).cuda()
        torch.cuda.synchronize()
        a.msm_partial(sa.data_ptr(), True, cut, parts.data_ptr())
        b.msm_partial(sb.data_ptr(), True, n - cut, parts.data_ptr() + pb)
        assert a.combine_partials(parts.data_ptr(), 2) == expect
        assert b.combine_partials(parts.data_ptr(), 2) == expect
    finally:
        for e in (whole, a, b):
            e.close()


@pytest.mark.parametrize("label", ["bls12-377", "pallas", "ed-on-bls12-377"])
def test_linearity_at_full_size(label):
    """Size-independent property at 2^18: msm(s) + msm(t) == msm(s + t mod q) (group addition by the oracle)."""
    cv = CURVES[label]
    O = OracleCurve(label)
    n = 1 << 18
    eng = m.MsmEngine(cv, 0, n)
    try:
        eng.random_points(n, seed=2024)
        s = inputs.scalars_to_ints(inputs.random_scalars(cv.q, n, 1))
        t = inputs.scalars_to_ints(inputs.random_scalars(cv.q, n, 2))
        u = [(x + y) % cv.q for x, y in zip(s, t)]
        rs, _ = eng.msm(inputs.ints_to_le_bytes(s, 32), n=n)
        rt, _ = eng.msm(inputs.ints_to_le_bytes(t, 32), n=n)
        ru, _ = eng.msm(inputs.ints_to_le_bytes(u, 32), n=n)
        if O.kind == "weierstrass":
            P = O.P
            total = P.to_affine(P.add(P.from_affine((rs["x"], rs["y"])), P.from_affine((rt["x"], rt["y"]))))
        else:
            T = O.T
            total = T.to_affine(T.add(T.from_affine((rs["x"], rs["y"])), T.from_affine((rt["x"], rt["y"]))))
        assert ru == O.result_of(total)
    finally:
        eng.close()


def test_state_reuse_across_sizes_and_options():
    """One context, calls of different sizes / window sizes / paths interleaved: every call is independent of the
    buffers the previous one left behind (idempotence)."""
    cv = CURVES["bls12-377"]
    n = 1 << 15
    eng = m.MsmEngine(cv, 0, n)
    try:
        eng.random_points(n, seed=5)
        sc = inputs.random_scalars(cv.q, n, 77)
        ref = {}
        for k in (n, 100, 1 << 12, 7):
            ref[k], _ = eng.msm(sc[:k], n=k)
        for k, c, proj in [(7, None, False), (n, 9, False), (100, None, True), (1 << 12, 15, False), (n, None, False), (100, 6, False),
                           (n, None, True), (1 << 12, None, False)]:
            got, _ = eng.msm(sc[:k], n=k, c=c, projective=proj)
            assert got == ref[k], (k, c, proj)
    finally:
        eng.close()


def test_error_codes_on_device():
    """Error behaviour of the boundary: more scalars than uploaded points, window size out of range, MSM before any
    point upload, more points than the context was created for -- negative code + message, and the context stays usable."""
    cv = CURVES["bls12-377"]
    eng = m.MsmEngine(cv, 0, 1 << 10)
    try:
        sc = inputs.random_scalars(cv.q, 64, 3)
        with pytest.raises(MgbError):
            eng.msm(sc, n=64)                      # no points yet
        eng.random_points(32, seed=1)
        with pytest.raises(MgbError):
            eng.msm(sc, n=64)                      # n > points held
        with pytest.raises(MgbError):
            eng.msm(sc[:32], n=32, c=40)           # window size out of range
        with pytest.raises(MgbError):
            eng.random_points((1 << 10) + 1, seed=1)   # beyond max_points
        ok, _ = eng.msm(sc[:32], n=32)
        O = OracleCurve("bls12-377")
        a = inputs.known_dlogs(1, 32)
        s = inputs.scalars_to_ints(sc[:32])
        assert ok == O.result_of(O.scale(sum(int(x) * int(y) for x, y in zip(s, a)), O.G))
    finally:
        eng.close()


def test_compute_msm_bigint_interface():
    """compute_msm(points, scalars) with bigint points {x, y, isZero} and bigint scalars, as the zprize harness calls it
    (scripts/zprize23/submission-bls377.ts:20-65), against the oracle's bigint Pippenger."""
    cv = CURVES["bls12-377"]
    O = OracleCurve("bls12-377")
    pts = _oracle_points(O, 33, 11)
    pts[5] = None
    sc = [int(x) for x in inputs.scalars_to_ints(inputs.random_scalars(cv.q, 33, 4))]
    compute = m.make_compute_msm(cv)
    big = [{"x": 0, "y": 0, "isZero": True} if P is None else {"x": P[0], "y": P[1], "isZero": False} for P in pts]
    got = compute(big, sc)
    assert got == O.msm(sc, pts)
    xy, _ = points_to_bytes([P for P in pts if P is not None], cv.coord_bytes)
    sc2 = [s for s, P in zip(sc, pts) if P is not None]
    assert compute(xy.tobytes(), scalars_to_bytes(sc2).tobytes()) == got    # byte interface, the infinity point dropped


@pytest.mark.parametrize("label", ["bls12-377", "ed-on-bls12-377"])
def test_tree_depth_and_finish_paths(label, monkeypatch):
    """The depth of the bucket trees and the choice between k_bucket_finish and direct leftover sums are tuning
    decisions: every combination (0 rounds = the scatter materialises the points, .., full depth; finish on / off)
    must give the same canonical point as the default."""
    cv = CURVES[label]
    n = 1 << 14
    eng = m.MsmEngine(cv, 0, n)
    try:
        eng.random_points(n, seed=77)
        sc = inputs.random_scalars(cv.q, n, 9)
        ref, tm = eng.msm(sc, n=n)
        for rounds in (0, 1, 2, 3, 6):
            for finish in ("0", "1"):
                monkeypatch.setenv("MGB_DEBUG_NROUNDS", str(rounds))
                monkeypatch.setenv("MGB_DEBUG_FINISH", finish)
                got, tm2 = eng.msm(sc, n=n)
                assert got == ref, (label, rounds, finish)
                assert tm2["rounds"] <= rounds
        monkeypatch.delenv("MGB_DEBUG_NROUNDS")
        monkeypatch.delenv("MGB_DEBUG_FINISH")
        assert eng.msm(sc, n=n)[0] == ref
    finally:
        eng.close()
