"""CPU test of a whole KERNEL: k_batch_add (batched-affine bucket accumulation, csrc/engine.cuh) compiled for the host
against the CUDA stand-ins of tests/host_emu/cuda_emu.h and run as 4 blocks x 128 lockstep host threads -- shuffles,
ballots, atomics, shared-memory staging, dynamic tile hand-out and the per-tile pair-list reservation included.  Two tree
rounds over synthetic buckets of 1..4 points, with endomorphism / negation references, a doubling and a cancellation,
against the oracle's affine arithmetic.  (The GPU versions of this check are the MSM parity tests of test_gpu_msm.py.)"""
import ctypes
import os
import random
import subprocess

import pytest

from oracle.params import BLS12_377, BLS12_381, ED_ON_BLS12_377, PALLAS
from oracle.twisted_edwards import TwistedEdwardsCurve
from oracle.weierstrass import AffineCurve

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
pytestmark = pytest.mark.timeout(600)
REF_NEG, REF_ENDO, REF_EMPTY = 0x80000000, 0x40000000, 0xFFFFFFFF
U32 = ctypes.c_uint32


# the kernels as shipped (lane-parallel inversion of a tile's total, csrc/warp.cuh; Horner kernels in one warp, csrc/onewarp.cuh)
@pytest.fixture(scope="module")
def emu_k(tmp_path_factory):
    so = str(tmp_path_factory.mktemp("emu_k") / "emu_k.so")
    subprocess.check_call(["g++", "-O2", "-std=c++17", "-shared", "-fPIC", "-pthread", "-Wno-unknown-pragmas",
                           "-o", so, os.path.join(ROOT, "tests", "host_emu", "emu_kernels.cpp")])
    lib = ctypes.CDLL(so)
    lib.variant = "product"
    return lib


@pytest.mark.parametrize("cid,prm,n,e_big", [(0, BLS12_377, 12, 2), (0, BLS12_377, 12, 8), (1, PALLAS, 8, 4)],
                         ids=["bls12-377-E2", "bls12-377-E8", "pallas-E4"])
def test_batch_add_two_rounds(emu_k, cid, prm, n, e_big):
    p = prm.p
    R = 1 << (32 * n)
    Ri = pow(R, -1, p)
    A = AffineCurve(prm)
    rnd = random.Random(90 + cid + e_big)
    limbs = lambda x: [(x >> (32 * i)) & 0xFFFFFFFF for i in range(n)]
    M = lambda x: x * R % p

    # ---- point table: x | y | beta*x, Montgomery form
    npts = 64
    pts = [A.scale(rnd.randrange(1, prm.q), prm.G) for _ in range(npts)]
    table = []
    for x, y in pts:
        table += limbs(M(x)) + limbs(M(y)) + limbs(M(prm.beta * x % p))

    def ref_point(ref):
        x, y = pts[ref & 0x3FFFFFFF]
        if ref & REF_ENDO:
            x = prm.beta * x % p
        if ref & REF_NEG:
            y = (-y) % p
        return (x, y)

    def rand_ref():
        return rnd.randrange(npts) | rnd.choice([0, REF_NEG]) | rnd.choice([0, REF_ENDO])

    # ---- buckets of 1..4 elements at 4-slot strides; slot pair q = slot / 2
    nb = 150
    buckets = []
    for j in range(nb):
        k = rnd.choice([1, 2, 3, 4, 4, 4])
        refs = [rand_ref() for _ in range(k)]
        if j == 3:                                         # doubling in round 0
            refs = [refs[0], refs[0]] + [rand_ref(), rand_ref()]
        if j == 5:                                         # cancellation in round 0, then infinity + point in round 1
            refs = [refs[0], refs[0] ^ REF_NEG] + [rand_ref(), rand_ref()]
        if j == 7:                                         # doubling in round 1: the two halves of the bucket are equal
            a, b = rand_ref(), rand_ref()
            refs = [a, b, a, b]
        buckets.append(refs)
    npairs = 2 * nb
    recs, lifes = [], []
    for refs in buckets:
        k = len(refs)
        recs += [refs[0], refs[1] if k > 1 else REF_EMPTY]
        lifes.append(2 if k > 2 else 1)
        if k > 2:
            recs += [refs[2], refs[3] if k > 3 else REF_EMPTY]
        else:
            recs += [REF_EMPTY, REF_EMPTY]                 # no pair at all
        lifes.append(1)
    while len(lifes) % 4:
        lifes.append(0)

    blocks = 4
    V = (U32 * (2 * npairs * 2 * n))()
    pairs_a = (U32 * (2 * npairs))()
    pairs_b = (U32 * (2 * npairs))()
    cnt_a, cnt_b, tiles = U32(0), U32(0), U32(0)
    scratch = (U32 * (blocks * 4 * 8 * (n // 4) * 32 * 4))()
    offs = (U32 * 2)(0, 2 * npairs)
    dummy = U32(0)
    emu_k.emu_batch_add(cid, 1, V, pairs_a, ctypes.byref(dummy), 0, e_big, 1000, pairs_a, ctypes.byref(cnt_a), ctypes.byref(tiles),
                        (U32 * len(recs))(*recs), (ctypes.c_uint8 * len(lifes))(*lifes), (U32 * len(table))(*table), offs, 0, 1, scratch, blocks)

    def slot(s):
        w = [int(V[s * 2 * n + i]) for i in range(2 * n)]
        if w[n - 1] & 0x80000000:
            return None
        x = sum(w[i] << (32 * i) for i in range(n)) * Ri % p
        y = sum(w[n + i] << (32 * i) for i in range(n)) * Ri % p
        return (x, y)

    def add_all(refs):
        acc = None
        for r_ in refs:
            acc = A.add(acc, ref_point(r_))
        return acc

    exp_pairs = set()
    for j, refs in enumerate(buckets):
        assert slot(4 * j) == add_all(refs[:2]), ("round 0, first half of bucket", j)
        if len(refs) > 2:
            assert slot(4 * j + 2) == add_all(refs[2:]), ("round 0, second half of bucket", j)
            exp_pairs.add((4 * j, 2))
    got_pairs = {(int(pairs_a[2 * i]), int(pairs_a[2 * i + 1])) for i in range(cnt_a.value)}
    assert cnt_a.value == len(exp_pairs) and got_pairs == exp_pairs        # the next round's pair list, no duplicates, no holes

    # ---- round 1: V[slot] += V[slot + 2] for the listed slots
    tiles.value = 0
    emu_k.emu_batch_add(cid, 0, V, pairs_a, ctypes.byref(cnt_a), 1, e_big, 1000, pairs_b, ctypes.byref(cnt_b), ctypes.byref(tiles),
                        (U32 * len(recs))(*recs), (ctypes.c_uint8 * len(lifes))(*lifes), (U32 * len(table))(*table), offs, 0, 1, scratch, blocks)
    for j, refs in enumerate(buckets):
        assert slot(4 * j) == add_all(refs), ("round 1, bucket", j)
    assert cnt_b.value == 0                                                 # nobody lives past round 1


@pytest.mark.parametrize("cid,prm,n", [(0, BLS12_377, 12), (1, PALLAS, 8)], ids=["bls12-377", "pallas"])
def test_final_horner_kernel(emu_k, cid, prm, n):
    """k_final: sum_w 2^(c w) S_w over K window sums given as XYZZ accumulators (one of them the neutral element)."""
    p = prm.p
    R = 1 << (32 * n)
    Ri = pow(R, -1, p)
    A = AffineCurve(prm)
    rnd = random.Random(70 + cid)
    M = lambda x: x * R % p
    limbs = lambda x: [(x >> (32 * i)) & 0xFFFFFFFF for i in range(n)]
    K, c = 4, 5
    S = [A.scale(rnd.randrange(1, prm.q), prm.G) for _ in range(K)]
    S[2] = None
    words = []
    for P in S:
        if P is None:
            coords = [0, M(1), 0, 0]
        else:
            z = rnd.randrange(1, p)
            coords = [M(P[0] * z * z % p), M(P[1] * z * z * z % p), M(z * z % p), M(z * z * z % p)]
        for v in coords:
            words += limbs(v)
    out = (U32 * (4 * n))()
    xy, flag = (U32 * (2 * n))(), (U32 * 1)(7)
    emu_k.emu_final(cid, K, c, (U32 * len(words))(*words), out, xy, flag)
    X, Y, ZZ, ZZZ = [sum(int(out[k * n + i]) << (32 * i) for i in range(n)) * Ri % p for k in range(4)]
    got = None if ZZ == 0 else (X * pow(ZZ, -1, p) % p, Y * pow(ZZZ, -1, p) % p)
    exp = None
    for w in reversed(range(K)):
        for _ in range(c):
            exp = A.double(exp)
        exp = A.add(exp, S[w])
    assert got == exp
    # the fused normalisation (single-GPU path): canonical plain coordinates + flag from the same launch
    assert flag[0] == 0 and tuple(sum(int(xy[k * n + i]) << (32 * i) for i in range(n)) for k in range(2)) == exp
    none = (U32 * (4 * n))()
    emu_k.emu_final(cid, K, c, (U32 * len(words))(*words), none, None, None)       # multi-GPU path: accumulator only
    assert list(none) == list(out)
    # a sum that is the neutral element: flag 1, zero coordinates
    zwords = []
    for _ in range(2):
        for v in (0, M(1), 0, 0):
            zwords += limbs(v)
    emu_k.emu_final(cid, 2, c, (U32 * len(zwords))(*zwords), out, xy, flag)
    assert flag[0] == 1 and not any(xy)


def test_final_horner_kernel_twisted_edwards(emu_k):
    prm, n = ED_ON_BLS12_377, 8
    p = prm.p
    R = 1 << (32 * n)
    Ri = pow(R, -1, p)
    T = TwistedEdwardsCurve(prm)
    rnd = random.Random(77)
    M = lambda x: x * R % p
    limbs = lambda x: [(x >> (32 * i)) & 0xFFFFFFFF for i in range(n)]
    K, c = 3, 4
    S = [T.scale(rnd.randrange(1, prm.q), T.one) for _ in range(K)]
    S[1] = T.zero
    words = []
    for P in S:
        x, y = T.to_affine(P)
        z = rnd.randrange(1, p)
        for v in (M(x * z % p), M(y * z % p), M(z), M(x * y * z % p)):
            words += limbs(v)
    out = (U32 * (4 * n))()
    xy, flag = (U32 * (2 * n))(), (U32 * 1)(7)
    emu_k.emu_final(3, K, c, (U32 * len(words))(*words), out, xy, flag)
    X, Y, Z, Tt = [sum(int(out[k * n + i]) << (32 * i) for i in range(n)) * Ri % p for k in range(4)]
    zi = pow(Z, -1, p)
    exp = T.zero
    for w in reversed(range(K)):
        for _ in range(c):
            exp = T.double(exp)
        exp = T.add(exp, S[w])
    assert (X * zi % p, Y * zi % p) == T.to_affine(exp) and (X * Y - Tt * Z) % p == 0
    assert flag[0] == 0 and tuple(sum(int(xy[k * n + i]) << (32 * i) for i in range(n)) for k in range(2)) == T.to_affine(exp)


def test_pair_add_two_rounds_twisted_edwards(emu_k):
    """k_pair_add (the accumulation kernel of the twisted-Edwards curve): buckets of 2..4 points laid out back to back
    as the scatter kernel does, negated references, P + P and P + (-P), two tree rounds, next round's pair list."""
    prm, n = ED_ON_BLS12_377, 8
    p = prm.p
    R = 1 << (32 * n)
    Ri = pow(R, -1, p)
    T = TwistedEdwardsCurve(prm)
    rnd = random.Random(95)
    M = lambda x: x * R % p
    limbs = lambda x: [(x >> (32 * i)) & 0xFFFFFFFF for i in range(n)]
    npts = 48
    pts = [T.scale(rnd.randrange(1, prm.q), T.one) for _ in range(npts)]
    table = []
    for P in pts:
        x, y = T.to_affine(P)
        table += limbs(M(x)) + limbs(M(y)) + limbs(M(x * y % p)) + limbs(M(2 * prm.d * x * y % p))      # x | y | t | k t
    ref_point = lambda ref: T.negate(pts[ref & 0x3FFFFFFF]) if ref & REF_NEG else pts[ref & 0x3FFFFFFF]
    rand_ref = lambda: rnd.randrange(npts) | rnd.choice([0, REF_NEG])
    buckets = []
    for j in range(100):
        refs = [rand_ref() for _ in range(rnd.choice([2, 3, 4, 4]))]
        if j == 2:
            refs = [refs[0], refs[0], rand_ref(), rand_ref()]                 # doubling (the addition law is unified)
        if j == 4:
            refs = [refs[0], refs[0] ^ REF_NEG, rand_ref()]                   # neutral element as an intermediate sum
        buckets.append(refs)
    recs, lifes, first_slot = [], [], []
    for refs in buckets:
        first_slot.append(len(recs))                                          # = 2 * (index of the bucket's first slot pair)
        recs += [refs[0], refs[1]]
        lifes.append(2 if len(refs) > 2 else 1)
        if len(refs) > 2:
            recs += [refs[2], refs[3] if len(refs) > 3 else REF_EMPTY]
            lifes.append(1)
    npairs = len(lifes)
    while len(lifes) % 4:
        lifes.append(0)
    V = (U32 * (2 * npairs * 4 * n))()
    pairs_a, pairs_b = (U32 * (2 * npairs))(), (U32 * (2 * npairs))()
    cnt_a, cnt_b, dummy = U32(0), U32(0), U32(0)
    offs = (U32 * 2)(0, 2 * npairs)
    args = ((U32 * len(recs))(*recs), (ctypes.c_uint8 * len(lifes))(*lifes), (U32 * len(table))(*table), offs, 0, 1, 3)
    emu_k.emu_pair_add_te(1, V, pairs_a, ctypes.byref(dummy), 0, pairs_a, ctypes.byref(cnt_a), *args)

    def slot(s):
        X, Y, Z, Tt = [sum(int(V[(s * 4 + k) * n + i]) << (32 * i) for i in range(n)) * Ri % p for k in range(4)]
        assert (X * Y - Tt * Z) % p == 0
        zi = pow(Z, -1, p)
        return (X * zi % p, Y * zi % p)

    def add_all(refs):
        acc = T.zero
        for r_ in refs:
            acc = T.add(acc, ref_point(r_))
        return T.to_affine(acc)

    exp_pairs = set()
    for refs, s0 in zip(buckets, first_slot):
        assert slot(s0) == add_all(refs[:2])
        if len(refs) > 2:
            assert slot(s0 + 2) == add_all(refs[2:])
            exp_pairs.add((s0, 2))
    assert {(int(pairs_a[2 * i]), int(pairs_a[2 * i + 1])) for i in range(cnt_a.value)} == exp_pairs and cnt_a.value == len(exp_pairs)
    emu_k.emu_pair_add_te(0, V, pairs_a, ctypes.byref(cnt_a), 1, pairs_b, ctypes.byref(cnt_b), *args)
    for refs, s0 in zip(buckets, first_slot):
        assert slot(s0) == add_all(refs)
    assert cnt_b.value == 0


@pytest.mark.parametrize("cid,prm,n", [(0, BLS12_377, 12), (1, PALLAS, 8), (2, BLS12_381, 12)], ids=["bls12-377", "pallas", "bls12-381"])
def test_set_get_points_and_combine_weierstrass(emu_k, cid, prm, n):
    """k_set_points (canonical bytes -> Montgomery table entry x | y | beta x, infinity flag), k_get_points (back), and
    k_normalize summing three partial accumulators (the multi-GPU combine) into the canonical affine point."""
    p = prm.p
    R = 1 << (32 * n)
    A = AffineCurve(prm)
    rnd = random.Random(33 + cid)
    limbs = lambda x: [(x >> (32 * i)) & 0xFFFFFFFF for i in range(n)]
    val = lambda w: sum(int(v) << (32 * i) for i, v in enumerate(w))
    pts = [A.scale(rnd.randrange(1, prm.q), prm.G) for _ in range(130)]          # two blocks, the second one partly idle
    flags = [1 if i in (5, 129) else 0 for i in range(len(pts))]
    xy = []
    for x, y in pts:
        xy += limbs(x) + limbs(y)
    table = (U32 * (len(pts) * 3 * n))()
    back = (U32 * len(xy))()
    zback = (ctypes.c_uint8 * len(pts))()
    emu_k.emu_set_get_points(cid, len(pts), (U32 * len(xy))(*xy), (ctypes.c_uint8 * len(pts))(*flags), table, back, zback)
    for i, (x, y) in enumerate(pts):
        e = [int(v) for v in table[i * 3 * n:(i + 1) * 3 * n]]
        if flags[i]:
            assert e[n - 1] & 0x80000000 and zback[i] == 1 and val(back[i * 2 * n:i * 2 * n + 2 * n]) == 0
            continue
        assert (val(e[:n]), val(e[n:2 * n]), val(e[2 * n:])) == (x * R % p, y * R % p, prm.beta * x * R % p)
        assert (val(back[i * 2 * n:i * 2 * n + n]), val(back[i * 2 * n + n:(i + 1) * 2 * n]), zback[i]) == (x, y, 0)
    # combine: three XYZZ partial sums (one the neutral element) -> affine
    parts = [pts[0], None, pts[1]]
    words = []
    for P in parts:
        z = rnd.randrange(1, p)
        coords = [0, R % p, 0, 0] if P is None else [P[0] * z * z * R % p, P[1] * z * z * z * R % p, z * z * R % p, z * z * z * R % p]
        for v in coords:
            words += limbs(v)
    out, flag = (U32 * (2 * n))(), (U32 * 2)(7, 7)          # flag[0]: the sum is the neutral element, flag[1]: a rank reported a failure
    emu_k.emu_normalize(cid, (U32 * len(words))(*words), 3, 4 * n, out, flag)
    assert (val(out[:n]), val(out[n:]), flag[0], flag[1]) == (*A.add(pts[0], pts[1]), 0, 0)
    # the records mgb_msm_sharded gathers: accumulator + status word (padded to four limbs); any non-zero status is reported
    for bad in (None, 1):
        rec = []
        for k in range(3):
            rec += words[k * 4 * n:(k + 1) * 4 * n] + [0xFFFFFFFF if k == bad else 0, 9, 9, 9]
        emu_k.emu_normalize(cid, (U32 * len(rec))(*rec), 3, 4 * n + 4, out, flag)
        assert (val(out[:n]), val(out[n:]), flag[0]) == (*A.add(pts[0], pts[1]), 0) and (flag[1] != 0) == (bad is not None)
    words2 = words[:4 * n] + [w for w in words[:4 * n]]
    neg = A.negate(pts[0])
    words2[4 * n:] = sum((limbs(v) for v in (neg[0] * R % p, neg[1] * R % p, R % p, R % p)), [])
    emu_k.emu_normalize(cid, (U32 * len(words2))(*words2), 2, 4 * n, out, flag)                # P + (-P): the neutral element
    assert flag[0] == 1 and val(out[:]) == 0
    # the eight-GPU shape and beyond: 8, 11 and 17 partials (quad q sums q, q + 8, ..., then a tree over the eight quads),
    # with a duplicate (doubling inside the tree) and neutral elements among them
    for count in (8, 11, 17):
        sel = [pts[i % len(pts)] for i in range(count)]
        sel[3] = None
        sel[5] = sel[4]
        words3 = []
        for P in sel:
            z = rnd.randrange(1, p)
            coords = [0, R % p, 0, 0] if P is None else [P[0] * z * z * R % p, P[1] * z * z * z * R % p, z * z * R % p, z * z * z * R % p]
            for v in coords:
                words3 += limbs(v)
        exp = None
        for P in sel:
            exp = A.add(exp, P)
        emu_k.emu_normalize(cid, (U32 * len(words3))(*words3), count, 4 * n, out, flag)
        assert (val(out[:n]), val(out[n:]), flag[0]) == ((*exp, 0) if exp is not None else (0, 0, 1)), count
