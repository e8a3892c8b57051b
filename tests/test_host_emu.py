"""CPU test: the engine's math headers (field.cuh, ec.cuh), compiled for the host with an emulated
carry flag (tests/host_emu/emu_lib.cpp), against the oracle.  Mirrors src/field.test.ts and the
curve unit tests of the reference; the GPU versions of the same checks are in test_gpu_field.py."""
import ctypes
import os
import random
import subprocess

import pytest

from oracle.params import BLS12_377, BLS12_381, ED_ON_BLS12_377, PALLAS
from oracle.twisted_edwards import TwistedEdwardsCurve
from oracle.weierstrass import AffineCurve

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
FIELDS = [(BLS12_377.p, 12), (BLS12_377.q, 8), (PALLAS.p, 8), (BLS12_381.p, 12)]


@pytest.fixture(scope="module")
def emu(tmp_path_factory):
    so = str(tmp_path_factory.mktemp("emu") / "emu.so")
    subprocess.check_call(["g++", "-O2", "-shared", "-fPIC", "-Wno-unknown-pragmas", "-x", "c++", "-o", so,
                           os.path.join(ROOT, "tests", "host_emu", "emu_lib.cpp")])
    return ctypes.CDLL(so)


def L(xs, n):
    out = []
    for x in xs:
        out += [(x >> (32 * i)) & 0xFFFFFFFF for i in range(n)]
    return (ctypes.c_uint32 * len(out))(*out)


def I(a, n, k):
    return [sum(int(a[j * n + i]) << (32 * i) for i in range(n)) for j in range(k)]


@pytest.mark.parametrize("fid", [0, 1, 2, 3])
def test_field_ops(emu, fid):
    p, n = FIELDS[fid]
    R = 1 << (32 * n)
    Ri = pow(R, -1, p)
    rnd = random.Random(fid)
    edge = [(0, 0), (1, p - 1), (p - 1, p - 1), (p - 1, 1), (0, 5), (2, p - 2)]
    out = (ctypes.c_uint32 * n)()
    for it in range(200):
        a, b = edge[it] if it < len(edge) else (rnd.randrange(p), rnd.randrange(p))
        for op, exp in [(0, a * b * Ri % p), (1, (a + b) % p), (2, (a - b) % p), (4, a * R % p), (5, a * Ri % p),
                        (6, a * a * Ri % p), (7, (-a) % p)]:
            emu.emu_fe_op(fid, op, out, L([a], n), L([b], n))
            assert I(out, n, 1)[0] == exp, (fid, op, hex(a), hex(b))
    for it in range(20):
        a = [1, 2, p - 1, R % p][it] if it < 4 else rnd.randrange(1, p)
        exp = pow(a * Ri, -1, p) * R % p
        for op in (3, 8, 9):  # Fermat, binary-gcd and division-step inverses (src/field.test.ts: inverse)
            if op == 3 and it >= 6:
                continue
            emu.emu_fe_op(fid, op, out, L([a], n), L([0], n))
            assert I(out, n, 1)[0] == exp
    for it in range(400):   # the division-step inverse is the one on the hot path: more cases
        a = rnd.randrange(1, p)
        emu.emu_fe_op(fid, 9, out, L([a], n), L([0], n))
        assert I(out, n, 1)[0] == pow(a * Ri, -1, p) * R % p
    for op in (8, 9):
        emu.emu_fe_op(fid, op, out, L([0], n), L([0], n))
        assert I(out, n, 1)[0] == 0


@pytest.mark.parametrize("fid", [0, 1, 2, 3])
def test_dedicated_squaring(emu, fid):
    """Field::sqr is its own routine (doubled off-diagonal products, skipped low products): operands with
    all-ones limbs, top bits set in every limb, single limbs, and random ones, against a*a/R mod p."""
    p, n = FIELDS[fid]
    R = 1 << (32 * n)
    Ri = pow(R, -1, p)
    rnd = random.Random(100 + fid)
    cases = [p - 1, p - 2, (p - 1) // 2, (p + 1) // 2, R % p, (R - 1) % p]
    cases += [((1 << (32 * k + 32)) - 1) % p for k in range(n)]                    # low k+1 limbs all ones
    cases += [sum(0x80000000 << (32 * i) for i in range(n)) % p, sum(0xffffffff << (32 * i) for i in range(0, n, 2)) % p]
    cases += [(0xffffffff << (32 * k)) % p for k in range(n)] + [(1 << (32 * k + 31)) % p for k in range(n)]
    cases += [rnd.randrange(p) for _ in range(600)]
    out = (ctypes.c_uint32 * n)()
    for a in cases:
        emu.emu_fe_op(fid, 6, out, L([a], n), L([0], n))
        assert I(out, n, 1)[0] == a * a * Ri % p, (fid, hex(a))


@pytest.mark.parametrize("cid,prm,n", [(0, BLS12_377, 12), (1, PALLAS, 8), (2, BLS12_381, 12)], ids=["bls12-377", "pallas", "bls12-381"])
def test_weierstrass_ops(emu, cid, prm, n):
    p = prm.p
    R = 1 << (32 * n)
    Ri = pow(R, -1, p)
    A = AffineCurve(prm)
    rnd = random.Random(11)
    INF = 1 << (32 * n - 1)
    M = lambda x: x * R % p

    def enc_aff(P):
        return [INF, 0] if P is None else [M(P[0]), M(P[1])]

    def dec_aff(v):
        return None if v[0] & INF else (v[0] * Ri % p, v[1] * Ri % p)

    def enc_x(P):
        if P is None:
            return [0, M(1), 0, 0]
        z = rnd.randrange(1, p)
        return [M(P[0] * z * z % p), M(P[1] * z * z * z % p), M(z * z % p), M(z * z * z % p)]

    def dec_x(v):
        X, Y, ZZ, ZZZ = [t * Ri % p for t in v]
        return None if ZZ == 0 else (X * pow(ZZ, -1, p) % p, Y * pow(ZZZ, -1, p) % p)

    pts = [A.scale(rnd.randrange(1, prm.q), prm.G) for _ in range(6)]
    cases = [(pts[0], pts[1]), (pts[2], pts[2]), (pts[3], A.negate(pts[3])), (None, pts[4]), (pts[5], None), (None, None)]
    out = (ctypes.c_uint32 * (4 * n))()
    for P, Q in cases:
        exp = A.add(P, Q)
        emu.emu_w_op(cid, 0, out, L(enc_aff(P), n), L(enc_aff(Q), n)); assert dec_aff(I(out, n, 2)) == exp
        emu.emu_w_op(cid, 1, out, L(enc_x(P), n), L(enc_aff(Q), n)); assert dec_x(I(out, n, 4)) == exp
        emu.emu_w_op(cid, 2, out, L(enc_x(P), n), L(enc_x(Q), n)); assert dec_x(I(out, n, 4)) == exp
        emu.emu_w_op(cid, 3, out, L(enc_x(P), n), L(enc_x(Q), n)); assert dec_x(I(out, n, 4)) == A.double(P)
        emu.emu_w_op(cid, 4, out, L(enc_x(P), n), L(enc_x(Q), n)); assert dec_aff(I(out, n, 2)) == P
        emu.emu_w_op(cid, 5, out, L(enc_aff(P), n), L(enc_aff(Q), n)); assert dec_x(I(out, n, 4)) == P


def test_twisted_edwards_ops(emu):
    prm, n = ED_ON_BLS12_377, 8
    p = prm.p
    R = 1 << (32 * n)
    Ri = pow(R, -1, p)
    T = TwistedEdwardsCurve(prm)
    rnd = random.Random(13)
    M = lambda x: x * R % p

    def enc_e(P):
        z = rnd.randrange(1, p)
        x, y = T.to_affine(P)
        return [M(x * z % p), M(y * z % p), M(z), M(x * y * z % p)]

    def enc_a(P):
        x, y = T.to_affine(P)
        return [M(x), M(y), M(2 * prm.d * x * y % p)]

    def dec_e(v):
        X, Y, Z, Tt = [t * Ri % p for t in v]
        assert (X * Y - Tt * Z) % p == 0
        zi = pow(Z, -1, p)
        return (X * zi % p, Y * zi % p)

    pts = [T.scale(rnd.randrange(1, prm.q), T.one) for _ in range(4)]
    out = (ctypes.c_uint32 * (4 * n))()
    for P, Q in [(pts[0], pts[1]), (pts[2], pts[2]), (pts[3], T.negate(pts[3])), (T.zero, pts[0]), (pts[1], T.zero)]:
        exp = T.to_affine(T.add(P, Q))
        emu.emu_te_op(0, out, L(enc_e(P), n), L(enc_e(Q), n)); assert dec_e(I(out, n, 4)) == exp
        emu.emu_te_op(1, out, L(enc_e(P), n), L(enc_a(Q), n)); assert dec_e(I(out, n, 4)) == exp
        emu.emu_te_op(2, out, L(enc_a(P), n), L(enc_a(Q), n)); assert dec_e(I(out, n, 4)) == exp
        emu.emu_te_op(3, out, L(enc_e(P), n), L(enc_e(Q), n)); assert dec_e(I(out, n, 4)) == T.to_affine(T.double(P))
        emu.emu_te_op(4, out, L(enc_e(P), n), L(enc_e(Q), n)); assert tuple(t * Ri % p for t in I(out, n, 2)) == T.to_affine(P)
