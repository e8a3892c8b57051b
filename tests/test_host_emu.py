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
# the warp emulation deadlocks (like the GPU would misbehave) if lanes diverge around a shuffle / ballot: fail, don't hang
pytestmark = pytest.mark.timeout(300)
FIELDS = [(BLS12_377.p, 12), (BLS12_377.q, 8), (PALLAS.p, 8), (BLS12_381.p, 12)]


@pytest.fixture(scope="module")
def emu(tmp_path_factory):
    so = str(tmp_path_factory.mktemp("emu") / "emu.so")
    subprocess.check_call(["g++", "-O2", "-shared", "-fPIC", "-pthread", "-Wno-unknown-pragmas", "-x", "c++", "-o", so,
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


# ---- warp-cooperative multiplication (csrc/warp.cuh): one limb per lane, run on a 32-thread lockstep
# emulation of a warp (tests/host_emu/simt_emu.h) -- the same source the Horner kernels compile for the GPU

@pytest.mark.parametrize("fid", [0, 1, 2, 3])
def test_warp_cooperative_mul(emu, fid):
    p, n = FIELDS[fid]
    R = 1 << (32 * n)
    Ri = pow(R, -1, p)
    rnd = random.Random(200 + fid)
    edge = [0, 1, p - 1, p - 2, (p - 1) // 2, R % p, (R - 1) % p, (1 << 32) - 1,
            sum(0xffffffff << (32 * i) for i in range(n)) % p, sum(0xffffffff << (32 * i) for i in range(n - 1))]
    A = [x for x in edge for _ in edge] + [rnd.randrange(p) for _ in range(501)]      # odd count: the second half-warp idles once
    B = [y for _ in edge for y in edge] + [rnd.randrange(p) for _ in range(501)]
    out = (ctypes.c_uint32 * (n * len(A)))()
    emu.emu_warp_mul(fid, out, L(A, n), L(B, n), len(A))
    got = I(out, n, len(A))
    for a, b, g in zip(A, B, got):
        assert g == a * b * Ri % p, (fid, hex(a), hex(b))


@pytest.mark.parametrize("fid", [0, 1, 2, 3])
def test_warp_carry_and_borrow_lookahead(emu, fid):
    """WarpField::finish alone, on inputs a random product practically never produces: carries that ripple through
    runs of 0xffffffff limbs, borrows that ripple through limbs equal to the modulus's, pending carries at their maximum."""
    p, n = FIELDS[fid]
    rnd = random.Random(300 + fid)
    plimbs = [(p >> (32 * i)) & 0xFFFFFFFF for i in range(n)]
    cases = []   # (t limbs, c limbs) with sum (t_i + c_i) 2^(32 i) < 2p

    def val(t, c):
        return sum((ti + ci) << (32 * i) for i, (ti, ci) in enumerate(zip(t, c)))

    for lo_run in range(0, n - 1):                       # carry generated below a run of all-ones limbs
        for hi_run in range(lo_run, n - 1):
            t = [rnd.getrandbits(32) for _ in range(n)]
            c = [0] * n
            for i in range(lo_run, hi_run + 1):
                t[i] = 0xFFFFFFFF
            t[n - 1] = rnd.randrange(plimbs[n - 1])      # keeps the value below 2p
            c[lo_run] = rnd.choice([1, 2, (1 << 32) + 1, (1 << 33) + 8])
            if lo_run > 0:
                t[lo_run - 1], c[lo_run - 1] = 0xFFFFFFFF, rnd.choice([1, 1 << 32])    # ...and a carry arriving through the shuffle
            cases.append((t, c))
    for k in range(1, n):                                # value = p + 2^(32k) - 1 - (anything below limb 0..): borrows through equal limbs
        v = p + (1 << (32 * k)) - 1 - rnd.randrange(2)
        if v < 2 * p:
            cases.append(([(v >> (32 * i)) & 0xFFFFFFFF for i in range(n)], [0] * n))
    for v in (p, p - 1, p + 1, 2 * p - 1, 0, 1):
        cases.append(([(v >> (32 * i)) & 0xFFFFFFFF for i in range(n)], [0] * n))
    for _ in range(60):                                  # random split of a random value into limbs and pending carries
        c = [rnd.randrange((1 << 33) + 9) for _ in range(n - 2)] + [rnd.randrange(4), 0]
        v = rnd.randrange(sum(ci << (32 * i) for i, ci in enumerate(c)), 2 * p)
        rest = v - sum(ci << (32 * i) for i, ci in enumerate(c))
        cases.append(([(rest >> (32 * i)) & 0xFFFFFFFF for i in range(n)], c))
    if len(cases) % 2:
        cases.append(cases[0])
    for j in range(0, len(cases), 2):                    # two elements per warp (lanes 0-15, 16-31)
        t = (ctypes.c_uint32 * 32)()
        c = (ctypes.c_uint64 * 32)()
        for g in range(2):
            for i in range(n):
                t[16 * g + i] = cases[j + g][0][i]
                c[16 * g + i] = cases[j + g][1][i]
        out = (ctypes.c_uint32 * 32)()
        emu.emu_warp_finish(fid, out, t, c)
        for g in range(2):
            v = val(*cases[j + g])
            assert v < 2 * p
            got = sum(int(out[16 * g + i]) << (32 * i) for i in range(n))
            assert got == (v - p if v >= p else v), (fid, j + g, hex(v))
            assert all(out[16 * g + i] == 0 for i in range(n, 16))


@pytest.mark.parametrize("fid", [0, 1, 2, 3])
def test_warp_parallel_inverse(emu, fid):
    """WarpField::inv (lane-parallel division steps with lazy signed limbs; an experiment the MSM does not use yet)
    == a^-1 in Montgomery form, on the emulated warp.  600 random elements per field were run once by hand."""
    p, n = FIELDS[fid]
    R = 1 << (32 * n)
    Ri = pow(R, -1, p)
    rnd = random.Random(400 + fid)
    A = [1, 2, p - 1, R % p, (R - 1) % p, p - 2, (p + 1) // 2, 1 << 30, (1 << 60) % p, 0] + [rnd.randrange(1, p) for _ in range(30)]
    out = (ctypes.c_uint32 * (n * len(A)))()
    emu.emu_warp_inv(fid, out, L(A, n), len(A))
    for a, g in zip(A, I(out, n, len(A))):
        assert g == (pow(a * Ri, -1, p) * R % p if a else 0), (fid, hex(a))


# ---- one-warp point arithmetic of the Horner kernels (csrc/onewarp.cuh) on an emulated warp

@pytest.mark.parametrize("cid,prm,n", [(0, BLS12_377, 12), (1, PALLAS, 8), (2, BLS12_381, 12)])
def test_onewarp_weierstrass_horner_steps(emu, cid, prm, n):
    """onewarp.cuh, what k_final / k_window_assemble run: the whole XYZZ point in one warp (no shared memory, no barriers)."""
    run = emu.emu_onewarp_w
    p = prm.p
    R = 1 << (32 * n)
    Ri = pow(R, -1, p)
    A = AffineCurve(prm)
    rnd = random.Random(40 + cid)
    M = lambda x: x * R % p

    def enc_x(P):
        if P is None:
            return [0, M(1), 0, 0]
        z = rnd.randrange(1, p)
        return [M(P[0] * z * z % p), M(P[1] * z * z * z % p), M(z * z % p), M(z * z * z % p)]

    def dec_x(v):
        X, Y, ZZ, ZZZ = [t * Ri % p for t in v]
        return None if ZZ == 0 else (X * pow(ZZ, -1, p) % p, Y * pow(ZZZ, -1, p) % p)

    def dbl_k(P, k):
        for _ in range(k):
            P = A.double(P)
        return P

    P, Q = (A.scale(rnd.randrange(1, prm.q), prm.G) for _ in range(2))
    out = (ctypes.c_uint32 * (4 * n))()
    for count in (0, 1, 2, 5):                       # runs of doublings: Y stays pending between them
        run(cid, 0, count, out, L(enc_x(P), n), L(enc_x(Q), n))
        assert dec_x(I(out, n, 4)) == dbl_k(P, count), count
    run(cid, 0, 3, out, L(enc_x(None), n), L(enc_x(Q), n))      # doubling the neutral element
    assert dec_x(I(out, n, 4)) is None
    for X, Y in [(P, Q), (P, P), (P, A.negate(P)), (None, Q), (P, None), (None, None)]:
        run(cid, 1, 0, out, L(enc_x(X), n), L(enc_x(Y), n))
        assert dec_x(I(out, n, 4)) == A.add(X, Y)
    run(cid, 2, 4, out, L(enc_x(P), n), L(enc_x(Q), n))         # one Horner step: 2^4 P + Q
    assert dec_x(I(out, n, 4)) == A.add(dbl_k(P, 4), Q)
    run(cid, 2, 3, out, L(enc_x(None), n), L(enc_x(Q), n))      # empty top window
    assert dec_x(I(out, n, 4)) == Q


def test_onewarp_twisted_edwards_horner_steps(emu):
    """OneWarpTwistedEdwards: the whole extended point in one warp, dedicated doubling (what k_final runs for the
    twisted-Edwards curve)."""
    run_te = emu.emu_onewarp_te
    prm, n = ED_ON_BLS12_377, 8
    p = prm.p
    R = 1 << (32 * n)
    Ri = pow(R, -1, p)
    T = TwistedEdwardsCurve(prm)
    rnd = random.Random(51)
    M = lambda x: x * R % p

    def enc_e(P):
        z = rnd.randrange(1, p)
        x, y = T.to_affine(P)
        return [M(x * z % p), M(y * z % p), M(z), M(x * y * z % p)]

    def dec_e(v):
        X, Y, Z, Tt = [t * Ri % p for t in v]
        assert (X * Y - Tt * Z) % p == 0
        zi = pow(Z, -1, p)
        return (X * zi % p, Y * zi % p)

    P, Q = (T.scale(rnd.randrange(1, prm.q), T.one) for _ in range(2))
    out = (ctypes.c_uint32 * (4 * n))()
    D = P
    for _ in range(3):
        D = T.double(D)
    run_te(0, 3, out, L(enc_e(P), n), L(enc_e(Q), n))
    assert dec_e(I(out, n, 4)) == T.to_affine(D)
    for X, Y in [(P, Q), (P, P), (P, T.negate(P)), (T.zero, Q), (P, T.zero)]:
        run_te(1, 0, out, L(enc_e(X), n), L(enc_e(Y), n))
        assert dec_e(I(out, n, 4)) == T.to_affine(T.add(X, Y))
    run_te(2, 3, out, L(enc_e(P), n), L(enc_e(Q), n))
    assert dec_e(I(out, n, 4)) == T.to_affine(T.add(D, Q))
    for count in (0, 1, 2, 7):
        E2 = P
        for _ in range(count):
            E2 = T.double(E2)
        run_te(0, count, out, L(enc_e(P), n), L(enc_e(Q), n))
        assert dec_e(I(out, n, 4)) == T.to_affine(E2), count
    run_te(0, 4, out, L(enc_e(T.zero), n), L(enc_e(Q), n))      # doubling the neutral element
    assert dec_e(I(out, n, 4)) == (0, 1)
    run_te(2, 5, out, L(enc_e(T.zero), n), L(enc_e(Q), n))      # empty top window
    assert dec_e(I(out, n, 4)) == T.to_affine(Q)


@pytest.mark.parametrize("cid,prm,n", [(0, BLS12_377, 12), (1, PALLAS, 8)])
def test_quad_cooperative_addition(emu, cid, prm, n):
    """QuadWeierstrass::add: four lanes hold X, Y, ZZ, ZZZ of one point; eight additions per warp, among them the
    rare cases (doubling, cancellation, neutral operands) next to generic ones in the same warp."""
    p = prm.p
    R = 1 << (32 * n)
    Ri = pow(R, -1, p)
    A = AffineCurve(prm)
    rnd = random.Random(60 + cid)
    M = lambda x: x * R % p

    def enc_x(P):
        if P is None:
            return [0, M(1), 0, 0]
        z = rnd.randrange(1, p)
        return [M(P[0] * z * z % p), M(P[1] * z * z * z % p), M(z * z % p), M(z * z * z % p)]

    def dec_x(v):
        X, Y, ZZ, ZZZ = [t * Ri % p for t in v]
        return None if ZZ == 0 else (X * pow(ZZ, -1, p) % p, Y * pow(ZZZ, -1, p) % p)

    pts = [A.scale(rnd.randrange(1, prm.q), prm.G) for _ in range(8)]
    pairs = [(pts[0], pts[1]), (pts[2], pts[2]), (pts[3], A.negate(pts[3])), (None, pts[4]), (pts[5], None), (None, None),
             (pts[6], pts[7]), (pts[7], pts[0])]
    a = [c for X, _ in pairs for c in enc_x(X)]
    b = [c for _, Y in pairs for c in enc_x(Y)]
    out = (ctypes.c_uint32 * (32 * n))()
    emu.emu_quad_add(cid, out, L(a, n), L(b, n))
    got = I(out, n, 32)
    for j, (X, Y) in enumerate(pairs):
        assert dec_x(got[4 * j:4 * j + 4]) == A.add(X, Y), j


# ---- two limbs (one 64-bit digit) per lane: four products per warp (csrc/warp.cuh WarpField2; an experiment)

@pytest.mark.parametrize("fid", [0, 1, 2, 3])
def test_warp_cooperative_mul_two_limbs_per_lane(emu, fid):
    p, n = FIELDS[fid]
    R = 1 << (32 * n)
    Ri = pow(R, -1, p)
    rnd = random.Random(600 + fid)
    edge = [0, 1, p - 1, p - 2, (p - 1) // 2, R % p, (R - 1) % p, (1 << 64) - 1,
            sum(0xffffffff << (32 * i) for i in range(n)) % p, sum(0xffffffff << (32 * i) for i in range(n - 2))]
    A = [x for x in edge for _ in edge] + [rnd.randrange(p) for _ in range(1001)]      # 1101: the last warp runs one product
    B = [y for _ in edge for y in edge] + [rnd.randrange(p) for _ in range(1001)]
    out = (ctypes.c_uint32 * (n * len(A)))()
    emu.emu_warp_mul2(fid, out, L(A, n), L(B, n), len(A))
    for a, b, g in zip(A, B, I(out, n, len(A))):
        assert g == a * b * Ri % p, (fid, hex(a), hex(b))


@pytest.mark.parametrize("fid", [0, 1, 2, 3])
def test_warp_carry_and_borrow_lookahead_two_limbs_per_lane(emu, fid):
    """WarpField2::finish on carries rippling through runs of all-ones DIGITS and borrows rippling through digits
    equal to the modulus's (a random product never gets there), four elements per warp."""
    p, n = FIELDS[fid]
    D = n // 2
    rnd = random.Random(700 + fid)
    W64 = (1 << 64) - 1
    pd = [(p >> (64 * i)) & W64 for i in range(D)]
    cases = []        # (t digits, clo digits, chi digits): sum (t + clo + chi 2^64) 2^(64 i) < 2p

    def val(t, clo, chi):
        return sum((a + b + (c << 64)) << (64 * i) for i, (a, b, c) in enumerate(zip(t, clo, chi)))

    for lo_run in range(0, D - 1):
        for hi_run in range(lo_run, D - 1):
            t = [rnd.getrandbits(64) for _ in range(D)]
            clo, chi = [0] * D, [0] * D
            for i in range(lo_run, hi_run + 1):
                t[i] = W64
            t[D - 1] = rnd.randrange(pd[D - 1])
            clo[lo_run] = rnd.choice([1, 2, W64])
            if lo_run > 0:
                t[lo_run - 1], chi[lo_run - 1] = W64, rnd.choice([1, 2])         # a carry arriving through the shuffle
            if val(t, clo, chi) < 2 * p:
                cases.append((t, clo, chi))
    for k in range(1, D):
        v = p + (1 << (64 * k)) - 1 - rnd.randrange(2)
        if v < 2 * p:
            cases.append(([(v >> (64 * i)) & W64 for i in range(D)], [0] * D, [0] * D))
    for v in (p, p - 1, p + 1, 2 * p - 1, 0, 1):
        cases.append(([(v >> (64 * i)) & W64 for i in range(D)], [0] * D, [0] * D))
    for _ in range(40):
        clo = [rnd.getrandbits(64) for _ in range(D - 2)] + [rnd.randrange(4), 0]
        chi = [rnd.randrange(3) for _ in range(D - 2)] + [0, 0]
        low = val([0] * D, clo, chi)
        v = rnd.randrange(low, 2 * p)
        rest = v - low
        cases.append(([(rest >> (64 * i)) & W64 for i in range(D)], clo, chi))
    while len(cases) % 4:
        cases.append(cases[0])
    for j in range(0, len(cases), 4):
        t = (ctypes.c_uint64 * 32)()
        clo = (ctypes.c_uint64 * 32)()
        chi = (ctypes.c_uint32 * 32)()
        for g in range(4):
            for i in range(D):
                t[8 * g + i], clo[8 * g + i], chi[8 * g + i] = cases[j + g][0][i], cases[j + g][1][i], cases[j + g][2][i]
        out = (ctypes.c_uint64 * 32)()
        emu.emu_warp_finish2(fid, out, t, clo, chi)
        for g in range(4):
            v = val(*cases[j + g])
            assert v < 2 * p
            got = sum(int(out[8 * g + i]) << (64 * i) for i in range(D))
            assert got == (v - p if v >= p else v), (fid, j + g, hex(v))
            assert all(out[8 * g + i] == 0 for i in range(D, 8))


@pytest.mark.parametrize("fid", [0, 1, 2, 3])
def test_warp_distributed_add_sub(emu, fid):
    """WarpField2::add / sub / dbl / is_zero on elements spread over 8-lane groups: carries and borrows cross the
    lanes by lookahead, so the operands are chosen to make them ripple (all-ones digits, digits equal to p's)."""
    p, n = FIELDS[fid]
    rnd = random.Random(800 + fid)
    W64 = (1 << 64) - 1
    D = n // 2
    special = [0, 1, p - 1, p - 2, (p - 1) // 2, (p + 1) // 2, W64, (1 << 64), (1 << (64 * (D - 1))) - 1, (1 << (64 * (D - 1))),
               p - (1 << 64), p - W64, sum(W64 << (64 * i) for i in range(D - 1))]
    special = [x % p for x in special]
    A = [x for x in special for _ in special] + [rnd.randrange(p) for _ in range(400)]
    B = [y for _ in special for y in special] + [rnd.randrange(p) for _ in range(400)]
    A += [x for x in special]                           # a - a, a + (p - a)
    B += [x for x in special]
    A += [x for x in special]
    B += [(p - x) % p for x in special]
    out = (ctypes.c_uint32 * (n * len(A)))()
    for op, f in ((0, lambda a, b: (a + b) % p), (1, lambda a, b: (a - b) % p), (2, lambda a, b: 2 * a % p), (3, lambda a, b: int(a == 0))):
        emu.emu_warp_addsub2(fid, op, out, L(A, n), L(B, n), len(A))
        got = I(out, n, len(A))
        for a, b, g in zip(A, B, got):
            if op == 3:
                g &= W64                               # the flag comes back in every digit of the group
            assert g == f(a, b), (fid, op, hex(a), hex(b))
