"""Bigint MSM (oracle = test infrastructure).

`msm` follows /root/reference/src/bigint/msm.ts:8-53 literally: unsigned c-bit windows with
c = max(ceil(log2 N) - 1, 1), K = ceil(b / c), 2^c - 1 buckets, running-sum reduction and a
Horner final sum.  `msm_naive` is the defining sum, used to check `msm` itself.
"""
from .field import log2_ceil


def msm(curve, scalars, points):
    N = len(scalars)
    assert N == len(points), "matching length"
    if N == 0:
        return curve.zero
    b = curve.scalar_bits
    c = max(log2_ceil(N) - 1, 1)
    cmask = (1 << c) - 1
    K = -(-b // c)
    L = 1 << c
    partition_sums = []
    for k in range(K):
        buckets = [curve.zero] * (L - 1)
        for i in range(N):
            l = (scalars[i] >> (k * c)) & cmask
            if l == 0:
                continue
            buckets[l - 1] = curve.add(buckets[l - 1], points[i])
        running = curve.zero
        triangle = curve.zero
        for l in range(L - 2, -1, -1):
            running = curve.add(running, buckets[l])
            triangle = curve.add(triangle, running)
        partition_sums.append(triangle)
    result = partition_sums[K - 1]
    for k in range(K - 2, -1, -1):
        for _ in range(c):
            result = curve.double(result)
        result = curve.add(result, partition_sums[k])
    return result


def msm_naive(curve, scalars, points):
    acc = curve.zero
    for s, P in zip(scalars, points):
        acc = curve.add(acc, curve.scale(s, P))
    return acc
