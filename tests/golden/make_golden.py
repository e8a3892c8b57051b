#!/usr/bin/env python3
"""Generates tests/golden/msm_vectors.json with the oracle (run from the repo root):

    python tests/golden/make_golden.py

For each curve: 48 points sampled the reference's way (random x, "try x+1, x+2, ..." until on
curve, cofactor cleared -- src/bigint/affine-weierstrass.ts:141-155, twisted-edwards.ts:174-191),
48 scalars < q, and the MSM result for several prefix lengths, computed twice: with the
reference-shaped Pippenger (`oracle.msm.msm`, src/bigint/msm.ts:8-53) and with the defining
double-and-add sum; both must agree before the vector is written.  The reference itself cannot
run in this image (no Node), so these are oracle-generated vectors, not reference outputs.
"""
import json
import os
import random
import sys

sys.path.insert(0, os.path.join(os.path.dirname(__file__), "..", ".."))
from oracle.params import BLS12_377, BLS12_381, PALLAS, ED_ON_BLS12_377  # noqa: E402
from oracle.weierstrass import AffineCurve, ProjectiveCurve  # noqa: E402
from oracle.twisted_edwards import TwistedEdwardsCurve  # noqa: E402
from oracle.msm import msm, msm_naive  # noqa: E402

NPTS = 48
PREFIXES = [1, 2, 3, 7, 16, 48]


def main():
    rnd = random.Random(0x6D6F6E74)
    out = {}
    for prm in (BLS12_377, PALLAS, BLS12_381):
        A, P = AffineCurve(prm), ProjectiveCurve(prm)
        pts = [A.point_from_x(rnd.randrange(prm.p)) for _ in range(NPTS)]
        assert all(A.is_on_curve(Q) and A.is_in_subgroup(Q) for Q in pts[:4])
        sc = [rnd.randrange(prm.q) for _ in range(NPTS)]
        sc[5] = 0
        sc[6] = prm.q - 1
        sc[7] = 1
        res = {}
        for n in PREFIXES:
            r1 = P.to_affine(msm(P, sc[:n], [P.from_affine(Q) for Q in pts[:n]]))
            r2 = P.to_affine(msm_naive(P, sc[:n], [P.from_affine(Q) for Q in pts[:n]]))
            assert r1 == r2
            res[str(n)] = None if r1 is None else [hex(r1[0]), hex(r1[1])]
        out[prm.label] = {"points": [[hex(x), hex(y)] for x, y in pts], "scalars": [hex(s) for s in sc], "results": res}
    prm = ED_ON_BLS12_377
    T = TwistedEdwardsCurve(prm)
    pts = [T.point_from_x(rnd.randrange(prm.p)) for _ in range(NPTS)]
    assert all(T.is_on_curve(Q) and T.is_in_subgroup(Q) for Q in pts[:4])
    sc = [rnd.randrange(prm.q) for _ in range(NPTS)]
    sc[5] = 0
    sc[6] = prm.q - 1
    sc[7] = 1
    res = {}
    for n in PREFIXES:
        r1 = T.to_affine(msm(T, sc[:n], pts[:n]))
        r2 = T.to_affine(msm_naive(T, sc[:n], pts[:n]))
        assert r1 == r2
        res[str(n)] = [hex(r1[0]), hex(r1[1])]
    out[prm.label] = {"points": [[hex(v) for v in T.to_affine(Q)] for Q in pts], "scalars": [hex(s) for s in sc], "results": res}
    path = os.path.join(os.path.dirname(__file__), "msm_vectors.json")
    with open(path, "w") as fh:
        json.dump(out, fh, indent=0)
    print("wrote", path)


if __name__ == "__main__":
    main()
