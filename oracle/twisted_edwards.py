"""Twisted Edwards curve -x^2 + y^2 = 1 + d x^2 y^2 in extended coordinates (oracle).

Follows /root/reference/src/bigint/twisted-edwards.ts: zero :34, from/toAffine :36-45,
add-2008-hwcd-3 :52-85, negate :99-101, isZero :117-124, scale :129-137, isOnCurve :160-168.
Points are (X, Y, Z, T).
"""
from .field import Field, inverse


class TwistedEdwardsCurve:
    def __init__(self, params):
        self.params = params
        self.p, self.q, self.d, self.h = params.p, params.q, params.d, params.h
        self.k = 2 * params.d
        self.Fp = Field(params.p)
        self.zero = (0, 1, 1, 0)
        self.one = self.from_affine(params.G)
        self.scalar_bits = (params.q - 1).bit_length()

    def from_affine(self, xy):
        x, y = xy
        return (x, y, 1, x * y % self.p)

    def to_affine(self, P):
        X, Y, Z, _ = P
        assert Z % self.p != 0, "Not an affine point"
        zi = inverse(Z, self.p)
        return (X * zi % self.p, Y * zi % self.p)

    def add(self, P1, P2):
        p, k = self.p, self.k
        X1, Y1, Z1, T1 = P1
        X2, Y2, Z2, T2 = P2
        A = (Y1 - X1) * (Y2 - X2) % p
        B = (Y1 + X1) * (Y2 + X2) % p
        C = T1 * T2 % p * k % p
        D = 2 * Z1 * Z2 % p
        E = (B - A) % p
        F = (D - C) % p
        G = (D + C) % p
        H = (B + A) % p
        return (E * F % p, G * H % p, F * G % p, E * H % p)

    def double(self, P):
        return self.add(P, P)

    def negate(self, P):
        return ((-P[0]) % self.p, P[1], P[2], (-P[3]) % self.p)

    def is_zero(self, P):
        X, Y, Z, T = P
        p = self.p
        return Z % p != 0 and X % p == 0 and T % p == 0 and (Y - Z) % p == 0

    def is_equal(self, P1, P2):
        p = self.p
        return (
            P1[2] % p != 0 and P2[2] % p != 0
            and (P1[0] * P2[2] - P2[0] * P1[2]) % p == 0
            and (P1[1] * P2[2] - P2[1] * P1[2]) % p == 0
            and (P1[3] * P2[2] - P2[3] * P1[2]) % p == 0
        )

    def scale(self, s, P):
        Q = self.zero
        for i in range(s.bit_length() - 1, -1, -1):
            Q = self.double(Q)
            if (s >> i) & 1:
                Q = self.add(Q, P)
        return Q

    def is_on_curve(self, P):
        X, Y, Z, T = P
        p = self.p
        if Z % p == 0:
            return False
        if (T * Z - X * Y) % p != 0:
            return False
        return (-X * X + Y * Y - Z * Z - self.d * T * T) % p == 0

    def is_in_subgroup(self, P):
        return self.is_zero(self.scale(self.q, P))

    def to_subgroup(self, P):
        return P if self.h == 1 else self.scale(self.h, P)

    def point_from_x(self, x):
        """twisted-edwards.ts:174-191: y^2 = (1 + x^2) / (1 - d x^2); try x+1, x+2, ..."""
        p = self.p
        while True:
            x = (x + 1) % p
            den = (1 - self.d * x * x) % p
            if den == 0:
                continue
            y = self.Fp.sqrt((1 + x * x) * inverse(den, p) % p)
            if y is not None:
                return self.to_subgroup((x, y, 1, x * y % p))
