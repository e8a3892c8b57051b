"""Seeded synthetic inputs (the reference's generators are unseeded `crypto.getRandomValues`,
src/util.ts:201-208, so the seeds are defined here).

  * random_scalars: uniform in [0, q) by rejection sampling of 32 LE bytes with the top byte masked
    to the bit length -- the method of src/curve-random.ts:151-189 / src/bigint/field-random.ts:29-36.
  * known_dlogs: the 64-bit multipliers a_i of `mgb_random_points` (P_i = a_i * G), so a test can
    form the closed-form expected result [(sum s_i a_i) mod q] G.
"""
import numpy as np

_GOLD = 0x9E3779B97F4A7C15
_M64 = (1 << 64) - 1


def splitmix64(seed: int, idx: np.ndarray) -> np.ndarray:
    """Vectorised mirror of the device function in csrc/engine.cuh."""
    with np.errstate(over="ignore"):
        z = (np.uint64(seed & _M64) + (idx.astype(np.uint64) + np.uint64(1)) * np.uint64(_GOLD))
        z = (z ^ (z >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)
        z = (z ^ (z >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)
        z = z ^ (z >> np.uint64(31))
    return z


def known_dlogs(seed: int, n: int) -> np.ndarray:
    a = splitmix64(seed, np.arange(n, dtype=np.uint64))
    a[a == 0] = 1
    return a


def random_scalars(q: int, n: int, seed: int) -> np.ndarray:
    """n scalars < q as an (n, 32) uint8 array, little endian."""
    bits = q.bit_length()
    top_mask = (1 << (bits - 8 * 31)) - 1 if bits < 256 else 0xFF
    qwords = np.array([(q >> (64 * i)) & _M64 for i in range(4)], dtype=np.uint64)
    rng = np.random.Generator(np.random.Philox(seed))
    out = np.empty((n, 32), dtype=np.uint8)
    filled = 0
    while filled < n:
        m = max(1024, int((n - filled) * 1.3))
        cand = rng.integers(0, 256, size=(m, 32), dtype=np.uint8)
        cand[:, 31] &= top_mask
        w = cand.view("<u8").reshape(m, 4)
        # lexicographic compare from the most significant word: keep cand < q
        lt = np.zeros(m, dtype=bool)
        eq = np.ones(m, dtype=bool)
        for k in (3, 2, 1, 0):
            lt |= eq & (w[:, k] < qwords[k])
            eq &= w[:, k] == qwords[k]
        good = cand[lt]
        take = min(len(good), n - filled)
        out[filled:filled + take] = good[:take]
        filled += take
    return out


def scalars_to_ints(sc: np.ndarray):
    return [int.from_bytes(row.tobytes(), "little") for row in sc]


def ints_to_le_bytes(vals, nbytes: int) -> np.ndarray:
    buf = b"".join(int(v).to_bytes(nbytes, "little") for v in vals)
    return np.frombuffer(buf, dtype=np.uint8).reshape(len(vals), nbytes).copy()


def dot_known_dlogs(sc: np.ndarray, a: np.ndarray) -> int:
    """sum_i s_i * a_i as an exact Python int, for (n, 32) uint8 little-endian scalars and uint64 multipliers --
    the scalar of the closed form  msm(s, a_i G) = [(sum s_i a_i) mod q] G  used by the parity checks at sizes where
    a per-element Python loop would take minutes.  Vectorised: every 32x32-bit partial product is exact in uint64,
    and its two 32-bit halves are summed separately (n < 2^31 keeps those sums below 2^63)."""
    n = sc.shape[0]
    assert sc.shape == (n, 32) and a.shape == (n,) and n < (1 << 31)
    w = np.ascontiguousarray(sc).view("<u4").reshape(n, 8).astype(np.uint64)      # 8 words of 32 bits per scalar
    a = a.astype(np.uint64)
    halves = (a & np.uint64(0xFFFFFFFF), a >> np.uint64(32))
    total = 0
    m32 = np.uint64(0xFFFFFFFF)
    for k in range(8):
        for h in (0, 1):
            prod = w[:, k] * halves[h]                                            # < 2^64, exact
            lo = int((prod & m32).sum(dtype=np.uint64))
            hi = int((prod >> np.uint64(32)).sum(dtype=np.uint64))
            total += (lo + (hi << 32)) << (32 * (k + h))
    return total
