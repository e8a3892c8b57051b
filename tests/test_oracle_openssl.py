"""The oracle pinned by a SECOND arbitrary-precision implementation (round-1 verdict: the golden vectors were only
self-consistent).  oracle/bn_check.c computes an MSM from its definition with OpenSSL BIGNUMs and the textbook affine
group laws -- no code, number representation or formula shared with oracle/*.py or with the engine.  Checked here:
every golden vector of tests/golden/msm_vectors.json on all four curves, the two known-answer identities the reference
holds (scripts/zprize23/submission-test-bls377.ts:17-26, submission-test.ts:12-21), the known-dlog point construction
the large-size closed-form tests rest on, and the Python oracle's Pippenger on fresh random inputs."""
import json
import os
import random
import shutil
import subprocess

import pytest

from oracle.params import CURVES, KAT_BLS12_377_POINT, KAT_ED377_POINT
from tests.helpers import OracleCurve

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
pytestmark = pytest.mark.skipif(shutil.which("gcc") is None or not os.path.exists("/usr/include/openssl/bn.h"), reason="needs gcc + OpenSSL headers")


@pytest.fixture(scope="module")
def bn(tmp_path_factory):
    exe = str(tmp_path_factory.mktemp("bn") / "bn_check")
    subprocess.check_call(["gcc", "-O2", "-Wall", "-Werror", "-o", exe, os.path.join(ROOT, "oracle", "bn_check.c"), "-lcrypto"])

    def msm(label, points, scalars):
        """points: affine tuples or None; returns an affine tuple, or None for the Weierstrass point at infinity"""
        prm = CURVES[label]
        lines = ["w %x" % prm.p if prm.kind == "weierstrass" else "te %x %x" % (prm.p, prm.d), str(len(points))]
        for P, s in zip(points, scalars):
            lines.append("inf 0 %x" % s if P is None else "%x %x %x" % (P[0], P[1], s))
        out = subprocess.run([exe], input="\n".join(lines) + "\n", capture_output=True, text=True, check=True).stdout.split()
        return None if out[0] == "inf" else (int(out[0], 16), int(out[1], 16))

    return msm


@pytest.mark.parametrize("label", list(CURVES))
def test_golden_vectors_against_openssl(bn, label):
    g = json.load(open(os.path.join(ROOT, "tests", "golden", "msm_vectors.json")))[label]
    pts = [(int(x, 16), int(y, 16)) for x, y in g["points"]]
    sc = [int(s, 16) for s in g["scalars"]]
    for n, exp in g["results"].items():
        n = int(n)
        assert bn(label, pts[:n], sc[:n]) == (int(exp[0], 16), int(exp[1], 16)), (label, n)


def test_reference_known_answers_against_openssl(bn):
    q = CURVES["bls12-377"].q
    P = KAT_BLS12_377_POINT
    assert bn("bls12-377", [P, P], [2, q - 1]) == P                        # submission-test-bls377.ts:17-26
    assert bn("bls12-377", [P, P], [1, q - 1]) is None
    s = [random.Random(5).randrange(q) for _ in range(40)]
    assert bn("bls12-377", [P] * 40, s) == bn("bls12-377", [P], [sum(s) % q])   # :28-45 (N = 1000 there)
    qe = CURVES["ed-on-bls12-377"].q
    E = KAT_ED377_POINT[:2]
    assert bn("ed-on-bls12-377", [E, E], [2, qe - 1]) == E                 # submission-test.ts:12-21
    assert bn("ed-on-bls12-377", [E, E], [1, qe - 1]) == (0, 1)


@pytest.mark.parametrize("label", list(CURVES))
def test_oracle_and_known_dlog_points_against_openssl(bn, label):
    from montgomery_b200 import inputs
    O = OracleCurve(label)
    prm = CURVES[label]
    a = [int(v) for v in inputs.known_dlogs(77, 6)]
    pts = [bn(label, [tuple(prm.G)], [v]) for v in a]                      # a_i G by OpenSSL double-and-add
    assert pts == [O.scale(v, O.G) for v in a]                             # == the oracle's scalar multiplication
    rnd = random.Random(9)
    sc = [rnd.randrange(prm.q) for _ in range(6)] + [0, prm.q - 1]
    pts2 = pts + [pts[0], pts[1]]
    if prm.kind == "weierstrass":
        pts2[3] = None
    exp = O.msm(sc, pts2)                                                  # the oracle's Pippenger (src/bigint/msm.ts)
    got = bn(label, pts2, sc)
    assert exp == O.result_of(got)
    # closed form used at sizes nothing else can check: sum s_i (a_i G) == [(sum s_i a_i) mod q] G
    k = sum(s * v for s, v, P in zip(sc, a + [a[0], a[1]], pts2) if P is not None) % prm.q
    assert got == bn(label, [tuple(prm.G)], [k])
