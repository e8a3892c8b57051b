"""Plain modular arithmetic (oracle = test infrastructure).

Follows /root/reference/src/bigint/field.ts:12-70 (ops), :92-122 (egcd inverse),
:127-156 (Tonelli-Shanks) and src/bigint/field-util.ts:8-11 (mod).
"""


def mod(x: int, p: int) -> int:
    # field-util.ts:8-11 (Python's % is already non-negative for p > 0)
    return x % p


def egcd(a: int, p: int):
    """field.ts:92-110 -- returns (d, x, y) with a*x + p*y = d."""
    if a > p:
        d, y, x = egcd(p, a)
        return d, x, y
    r0, r1 = p, a
    s0, s1 = 1, 0
    t0, t1 = 0, 1
    while r1 != 0:
        quo = r0 // r1
        r0, r1 = r1, r0 - quo * r1
        s0, s1 = s1, s0 - quo * s1
        t0, t1 = t1, t0 - quo * t1
    return r0, t0, s0


def inverse(x: int, p: int) -> int:
    """field.ts:117-122 -- throws on 0, like the reference."""
    if x % p == 0:
        raise ZeroDivisionError("cannot invert 0")
    d, xinv, _ = egcd(x % p, p)
    if d != 1:
        raise ZeroDivisionError("inverting failed (no inverse)")
    return xinv % p


def exp(x: int, n: int, p: int) -> int:
    """field.ts:77-85 square-and-multiply (kept literal, not pow(), as a cross-check)."""
    x %= p
    u = 1
    while n > 0:
        if n & 1:
            u = u * x % p
        x = x * x % p
        n >>= 1
    return u


class Field:
    """createField(p), field.ts:12-70."""

    def __init__(self, p: int):
        self.p = p
        self.size_in_bits = log2_ceil(p)
        self.size_in_bytes = (self.size_in_bits + 7) // 8
        # rootsOfUnity, field.ts:161-187
        t, M = p - 1, 0
        while t & 1 == 0:
            t >>= 1
            M += 1
        z = 2
        while exp(z, (p - 1) >> 1, p) == 1:
            z += 1
        roots = [exp(z, t, p)]
        for i in range(M):
            roots.append(roots[i] * roots[i] % p)
        self.t, self.M, self.roots = t, M, roots

    def add(self, x, y):
        return (x + y) % self.p

    def sub(self, x, y):
        return (x - y) % self.p

    def neg(self, x):
        return (-x) % self.p

    def mul(self, x, y):
        return x * y % self.p

    def sqr(self, x):
        return x * x % self.p

    def inv(self, x):
        return inverse(x, self.p)

    def is_equal(self, x, y):
        return (x - y) % self.p == 0

    def sqrt(self, x):
        """Tonelli-Shanks, field.ts:127-156; returns None if x is a non-residue."""
        p, t, M, roots = self.p, self.t, self.M, self.roots
        x %= p
        if x == 0:
            return 0
        i = M
        u = exp(x, (t - 1) // 2, p)
        sqrtx = x * u % p
        u = u * sqrtx % p
        while True:
            if u == 1:
                return sqrtx
            i_ = 1
            s = u * u % p
            while s != 1:
                s = s * s % p
                i_ += 1
            if i == i_:
                return None
            assert i_ < i
            i = i_
            sqrtx = sqrtx * roots[M - i - 1] % p
            u = u * roots[M - i] % p


def log2_ceil(n: int) -> int:
    """util.ts:134-142: ceil(log2(n)) = smallest k with n <= 2^k."""
    if n == 1:
        return 0
    return (n - 1).bit_length()


def montgomery_params(p: int, w: int, min_extra_bits: int = 2):
    """field-util.ts:18-42: limbs n, K = n*w, R = 2^K."""
    length_p = log2_ceil(p)
    n = -(-(length_p + min_extra_bits) // w)
    K = n * w
    return {"n": n, "K": K, "R": 1 << K, "lengthP": length_p}
