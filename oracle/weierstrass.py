"""Short Weierstrass curve y^2 = x^3 + b (a = 0), affine and projective (oracle).

Affine:     /root/reference/src/bigint/affine-weierstrass.ts:44-89 (add/double), :116-155
Projective: /root/reference/src/bigint/projective-weierstrass.ts:33-80 (add-1998-cmo-2),
            :85-115 (dbl-1998-cmo-2), :205-209 (toAffine)
Affine points are (x, y) tuples or None for the point at infinity; projective are (X, Y, Z)
with Z == 0 for infinity (projective-weierstrass.ts:25).
"""
from .field import Field, inverse


class AffineCurve:
    def __init__(self, params):
        assert params.a == 0
        self.params = params
        self.p, self.q, self.b, self.h = params.p, params.q, params.b, params.h
        self.Fp = Field(params.p)
        self.zero = None
        self.one = params.G

    def add(self, P1, P2):
        # affine-weierstrass.ts:44-69
        if P1 is None:
            return P2
        if P2 is None:
            return P1
        p = self.p
        x1, y1 = P1
        x2, y2 = P2
        if (x1 - x2) % p == 0:
            if (y1 - y2) % p == 0:
                return self.double(P1)
            assert (y1 + y2) % p == 0
            return None
        d = inverse(x2 - x1, p)
        m = (y2 - y1) * d % p
        x3 = (m * m - x1 - x2) % p
        y3 = (m * (x1 - x3) - y1) % p
        return (x3, y3)

    def double(self, P):
        # affine-weierstrass.ts:74-87.  (y = 0 has order 2; the reference would throw in inverse.)
        if P is None:
            return None
        p = self.p
        x, y = P
        if y % p == 0:
            return None
        d = inverse(2 * y, p)
        m = 3 * x * x * d % p
        x2 = (m * m - 2 * x) % p
        y2 = (m * (x - x2) - y) % p
        return (x2, y2)

    def negate(self, P):
        if P is None:
            return None
        return (P[0], (-P[1]) % self.p)

    def scale(self, s, P):
        # affine-weierstrass.ts:116-124 (MSB-first double-and-add)
        Q = None
        for i in range(s.bit_length() - 1, -1, -1):
            Q = self.double(Q)
            if (s >> i) & 1:
                Q = self.add(Q, P)
        return Q

    def is_on_curve(self, P):
        if P is None:
            return True
        x, y = P
        return (y * y - x * x * x - self.b) % self.p == 0

    def is_in_subgroup(self, P):
        return self.scale(self.q, P) is None

    def to_subgroup(self, P):
        return P if self.h == 1 else self.scale(self.h, P)

    def point_from_x(self, x):
        """affine-weierstrass.ts:141-155: try x+1, x+2, ... until x^3+b is a square; clear cofactor."""
        while True:
            x = (x + 1) % self.p
            y = self.Fp.sqrt((x * x * x + self.b) % self.p)
            if y is not None:
                return self.to_subgroup((x, y))


class ProjectiveCurve:
    def __init__(self, params):
        assert params.a == 0
        self.params = params
        self.p, self.q, self.b, self.h = params.p, params.q, params.b, params.h
        self.zero = (0, 1, 0)
        self.one = (params.G[0], params.G[1], 1)
        self.scalar_bits = (params.q - 1).bit_length()  # Scalar.sizeInBits = log2(q)

    def add(self, P1, P2):
        # projective-weierstrass.ts:33-80
        p = self.p
        X1, Y1, Z1 = P1
        X2, Y2, Z2 = P2
        if Z1 % p == 0:
            return P2
        if Z2 % p == 0:
            return P1
        Y1Z2 = Y1 * Z2 % p
        X1Z2 = X1 * Z2 % p
        Z1Z2 = Z1 * Z2 % p
        u = (Y2 * Z1 - Y1Z2) % p
        uu = u * u % p
        v = (X2 * Z1 - X1Z2) % p
        if v == 0:
            if u == 0:
                return self.double(P1)
            return self.zero
        vv = v * v % p
        vvv = v * vv % p
        R = vv * X1Z2 % p
        A = (uu * Z1Z2 - vvv - 2 * R) % p
        X3 = v * A % p
        Y3 = (u * (R - A) - vvv * Y1Z2) % p
        Z3 = vvv * Z1Z2 % p
        return (X3, Y3, Z3)

    def double(self, P):
        # projective-weierstrass.ts:85-115
        p = self.p
        X1, Y1, Z1 = P
        if Z1 % p == 0:
            return self.zero
        w = 3 * X1 * X1 % p
        s = Y1 * Z1 % p
        ss = s * s % p
        sss = s * ss
        R = Y1 * s % p
        B = X1 * R % p
        h = (w * w - 8 * B) % p
        X3 = 2 * h * s % p
        Y3 = (w * (4 * B - h) - 8 * R * R) % p
        Z3 = 8 * sss % p
        return (X3, Y3, Z3)

    def negate(self, P):
        return (P[0], (-P[1]) % self.p, P[2])

    def scale(self, s, P):
        Q = self.zero
        for i in range(s.bit_length() - 1, -1, -1):
            Q = self.double(Q)
            if (s >> i) & 1:
                Q = self.add(Q, P)
        return Q

    def is_equal(self, P1, P2):
        # projective-weierstrass.ts:124-137
        p = self.p
        X1, Y1, Z1 = P1
        X2, Y2, Z2 = P2
        if Z1 % p == 0:
            return Z2 % p == 0
        if Z2 % p == 0:
            return False
        return (X1 * Z2 - X2 * Z1) % p == 0 and (Y1 * Z2 - Y2 * Z1) % p == 0

    def from_affine(self, P):
        return self.zero if P is None else (P[0], P[1], 1)

    def to_affine(self, P):
        # projective-weierstrass.ts:205-209; infinity -> None
        X, Y, Z = P
        if Z % self.p == 0:
            return None
        zi = inverse(Z, self.p)
        return (X * zi % self.p, Y * zi % self.p)
