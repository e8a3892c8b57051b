"""GPU parity of the field layer against Python integers (mirrors src/field.test.ts: multiply,
square, add, subtract, inverse vs the bigint field, randomized plus edge values)."""
import ctypes
import random

import numpy as np
import pytest

from oracle.params import BLS12_377, BLS12_381, PALLAS

pytestmark = pytest.mark.gpu
FIELDS = [(BLS12_377.p, 12), (BLS12_377.q, 8), (PALLAS.p, 8), (BLS12_381.p, 12)]


def _run(lib, fid, op, a_vals, b_vals, nlimbs):
    nb = nlimbs * 4
    a = np.frombuffer(b"".join(v.to_bytes(nb, "little") for v in a_vals), dtype=np.uint8).copy()
    b = np.frombuffer(b"".join(v.to_bytes(nb, "little") for v in b_vals), dtype=np.uint8).copy()
    out = np.zeros_like(a)
    vp = ctypes.c_void_p
    rc = lib.mgb_field_op(0, fid, op, a.ctypes.data_as(vp), b.ctypes.data_as(vp), out.ctypes.data_as(vp), len(a_vals))
    assert rc == 0
    return [int.from_bytes(out[i * nb:(i + 1) * nb].tobytes(), "little") for i in range(len(a_vals))]


@pytest.mark.parametrize("fid", [0, 1, 2, 3])
def test_field_ops_gpu(fid):
    from montgomery_b200 import _native
    lib = _native.lib()
    p, n = FIELDS[fid]
    rnd = random.Random(100 + fid)
    N = 4096
    a = [0, 1, p - 1, p - 1, 0, 2] + [rnd.randrange(p) for _ in range(N - 6)]
    b = [0, p - 1, p - 1, 1, 5, p - 2] + [rnd.randrange(p) for _ in range(N - 6)]
    assert _run(lib, fid, 0, a, b, n) == [x * y % p for x, y in zip(a, b)]
    assert _run(lib, fid, 1, a, b, n) == [(x + y) % p for x, y in zip(a, b)]
    assert _run(lib, fid, 2, a, b, n) == [(x - y) % p for x, y in zip(a, b)]
    assert _run(lib, fid, 4, a, b, n) == [x * x % p for x in a]
    assert _run(lib, fid, 6, a, b, n) == [(-x) % p for x in a]
    inv_exp = [pow(x, -1, p) if x else 0 for x in a[:512]]
    assert _run(lib, fid, 3, a[:512], b[:512], n) == inv_exp     # Fermat
    assert _run(lib, fid, 5, a[:512], b[:512], n) == inv_exp     # binary gcd
    inv_all = [pow(x, -1, p) if x else 0 for x in a]
    assert _run(lib, fid, 7, a, b, n) == inv_all                  # division steps (used by the MSM)


@pytest.mark.parametrize("fid", [0, 1, 2, 3])
def test_warp_cooperative_mul_gpu(fid):
    """csrc/warp.cuh (one limb per lane, used by the Horner kernels) == a*b mod p, including an element count
    that leaves half a warp and most of the last block idle (the shuffles still run on every lane)."""
    from montgomery_b200 import _native
    lib = _native.lib()
    p, n = FIELDS[fid]
    rnd = random.Random(500 + fid)
    R = 1 << (32 * n)
    edge = [0, 1, 2, p - 1, p - 2, (p - 1) // 2, R % p, (R - 1) % p, (1 << 32) - 1, sum(0xffffffff << (32 * i) for i in range(n - 1))]
    a = [x for x in edge for _ in edge] + [rnd.randrange(p) for _ in range(4001)]
    b = [y for _ in edge for y in edge] + [rnd.randrange(p) for _ in range(4001)]
    assert _run(lib, fid, 8, a, b, n) == [x * y % p for x, y in zip(a, b)]
    assert _run(lib, fid, 8, a[:1], b[:1], n) == [a[0] * b[0] % p]
    assert _run(lib, fid, 8, a[100:103], b[100:103], n) == [x * y % p for x, y in zip(a[100:103], b[100:103])]
