"""GLV decomposition and signed-digit slicing, restated (oracle = test infrastructure).

  egcd_stop_early   /root/reference/src/glv/glv.ts:21-50
  GlvScalar         /root/reference/src/wasm/glv.ts:35-63 (constants), :77-80 (algorithm),
                    :187-214 (multiplyMsb = rounded high product), :216-226 (maxBits)
  signed_digits     /root/reference/src/msm-batched-affine.ts:183-193
  window_size       /root/reference/src/msm-common.ts:8-41
"""
from .field import log2_ceil, montgomery_params


def egcd_stop_early(l: int, p: int):
    assert l <= p
    r0, r1 = p, l
    s0, s1 = 1, 0
    t0, t1 = 0, 1
    while r1 * r1 > p:
        quo = r0 // r1
        r0, r1 = r1, r0 - quo * r1
        s0, s1 = s1, s0 - quo * s1
        t0, t1 = t1, t0 - quo * t1
    quo = r0 // r1
    r2 = r0 - quo * r1
    t2 = t0 - quo * t1
    v00, v10 = r1, -t1
    if max(r0, abs(t0)) <= max(r2, abs(t2)):
        v01, v11 = r0, -t0
    else:
        v01, v11 = r2, -t2
    return (v00, v01), (v10, v11)


def _trunc_div(a: int, b: int) -> int:
    """JS BigInt division truncates toward zero."""
    qt = abs(a) // abs(b)
    return qt if (a >= 0) == (b >= 0) else -qt


class GlvScalar:
    """createGlvScalar({q, lambda, w}) -> decompose / maxBits (scalar-glv.ts:19-51)."""

    def __init__(self, q: int, lam: int, w: int = 29):
        self.q, self.lam, self.w = q, lam, w
        n = montgomery_params(q, w, 1)["n"]          # scalar-glv.ts:36 (minExtraBits = 1)
        n0 = -(-n // 2)
        self.n, self.n0 = n, n0
        self.m = n0 * w
        self.k = (n - n0) * w
        (v00, v01), (v10, v11) = egcd_stop_early(lam, q)
        det = v00 * v11 - v10 * v01
        self.V = (v00, v01, v10, v11)
        self.det = det
        self.m0 = _trunc_div((1 << (self.m + self.k)) * -v11, det)
        self.m1 = _trunc_div((1 << (self.m + self.k)) * v10, det)
        # upper bound on |s0|, |s1| in bits (wasm/glv.ts:216-226), computed exactly here
        x0err = 0.5 + abs(self.m0) / 2 ** self.m + 1.0 * q / 2 ** (self.m + self.k)
        x1err = 0.5 + abs(self.m1) / 2 ** self.m + 1.0 * q / 2 ** (self.m + self.k)
        max_s0 = x0err * abs(v00) + x1err * abs(v01)
        max_s1 = x0err * abs(v10) + x1err * abs(v11)
        self.max_bits = max(log2_ceil(int(max_s0) + 1), log2_ceil(int(max_s1) + 1))

    def _mul_msb(self, x: int, y: int) -> int:
        """round(x*y / 2^m): floor plus the bit below the cut (wasm/glv.ts:187-214)."""
        prod = abs(x) * abs(y)
        r = (prod >> self.m) + ((prod >> (self.m - 1)) & 1)
        return r if (x >= 0) == (y >= 0) else -r

    def decompose(self, s: int):
        """s -> (s0, s1) signed, with s0 + s1*lambda = s (mod q) (wasm/glv.ts:77-80)."""
        v00, v01, v10, v11 = self.V
        shi = s >> self.k
        x0 = self._mul_msb(shi, self.m0)
        x1 = self._mul_msb(shi, self.m1)
        s0 = v00 * x0 + v01 * x1 + s
        s1 = v10 * x0 + v11 * x1
        return s0, s1


def window_size(field_bits: int, n: int) -> int:
    """msm-common.ts:8-41."""
    table = {
        "large": {14: 13, 15: 14, 16: 14, 17: 14, 18: 14, 19: 18, 20: 18},
        "small": {16: 12},
    }["large" if field_bits > 260 else "small"]
    return table.get(n, max(n - 1, 1))


def signed_digits(s: int, c: int, K: int):
    """msm-batched-affine.ts:183-193: digits l in [0, L] with a carry flag meaning 'negated'."""
    L = 1 << (c - 1)
    out = []
    carry = 0
    for k in range(K):
        l = ((s >> (k * c)) & ((1 << c) - 1)) + carry
        if l > L:
            l = 2 * L - l
            carry = 1
        else:
            carry = 0
        out.append((l, carry))
    assert carry == 0, "K windows must absorb the final carry"
    return out
