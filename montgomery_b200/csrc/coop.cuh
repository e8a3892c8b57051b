// Quad-cooperative point addition for the latency-bound levels of the bucket reduction (the last levels of the group
// trees, the suffix scans over the 32 digit values, the multi-GPU combine).
// (Round 1 also had block-cooperative doubling / addition routines here for the Horner chain -- four warps sharing the
// products of a formula level through shared memory; round 2 replaced them by the one-warp arithmetic of onewarp.cuh,
// measured faster, and removed them.)
#pragma once
#include "ec.cuh"
#include "warp.cuh"

namespace mgb {

// Quad-cooperative XYZZ addition: FOUR LANES of a warp share one point addition.  A dependent chain of
// point additions by one thread costs 14 sequential field multiplications per step (~11 us at the
// ~0.8 us latency of a lone 377-bit product); the latency-bound levels of the bucket reduction (the
// last levels of the group trees, the suffix scans over the 32 digit values) have far fewer
// additions than lanes.  Here the point is DISTRIBUTED over a quad -- lane k = 0..3 holds X, Y, ZZ,
// ZZZ -- every lane computes one of the (up to four) independent products of a formula level in the
// same SIMT multiplication, and the operands travel by shuffles: 4 multiplication latencies per
// addition instead of 14 (add-2008-s has 14 products in 4 dependent levels).
template <class P>
struct QuadWeierstrass {
  typedef Field<P> F;
  typedef Fe<P> fe;
  typedef Weierstrass<P> G;
  static constexpr uint32_t FULL = 0xffffffffu;

  MGB_DEV static fe shfl(const fe& a, int src) {
    fe r;
    _Pragma("unroll") for (int i = 0; i < P::N; i++) r.v[i] = __shfl_sync(FULL, a.v[i], src);
    return r;
  }
  MGB_DEV static fe pick(int k, const fe& a0, const fe& a1, const fe& a2, const fe& a3) {
    fe r;
    _Pragma("unroll") for (int i = 0; i < P::N; i++) r.v[i] = k == 0 ? a0.v[i] : (k == 1 ? a1.v[i] : (k == 2 ? a2.v[i] : a3.v[i]));
    return r;
  }
  // coordinate k of the neutral element (0, 1, 0, 0)
  MGB_DEV static fe zero_coord(int k) { return k == 1 ? F::one() : F::zero(); }

  // a + b for distributed points; must be called by all 32 lanes (8 independent additions per warp)
  MGB_DEV static fe add(const fe& a, const fe& b) {
    const int lane = threadIdx.x & 31, k = lane & 3, q = lane & ~3;
    const bool infA = __shfl_sync(FULL, (int)F::is_zero(a), q + 2) != 0;
    const bool infB = __shfl_sync(FULL, (int)F::is_zero(b), q + 2) != 0;
    // level 1:  k0: U1 = X1*ZZ2   k1: S1 = Y1*ZZZ2   k2: U2 = ZZ1*X2   k3: S2 = ZZZ1*Y2
    const fe t1 = F::mul(a, shfl(b, q + (k ^ 2)));
    // level 2:  k0: PP = (U2-U1)^2   k1: RR = (S2-S1)^2   k2: ZZ1*ZZ2   k3: ZZZ1*ZZZ2
    const fe diff = F::sub(shfl(t1, q + (k ^ 2)), t1);           // k0: P, k1: R
    const fe t2 = F::mul(k < 2 ? diff : a, k < 2 ? diff : b);
    const bool pz = __shfl_sync(FULL, (int)F::is_zero(diff), q) != 0;
    // level 3:  k0: PPP = P*PP   k1: Q = U1*PP   k2: ZZ3 = ZZ1ZZ2*PP
    const fe pp = shfl(t2, q), u1 = shfl(t1, q);
    const fe t3 = F::mul(k == 0 ? diff : (k == 1 ? u1 : t2), k == 0 ? t2 : pp);
    // level 4:  k0: S1*PPP   k1: R*(Q - X3), X3 = RR - PPP - 2Q   k3: ZZZ3 = ZZZ1ZZZ2*PPP
    const fe ppp = shfl(t3, q), s1 = shfl(t1, q + 1);
    const fe x3 = F::sub(F::sub(t2, ppp), F::dbl(t3));           // on k1
    const fe t4 = F::mul(k == 0 ? s1 : (k == 1 ? diff : t2), k == 0 ? t3 : (k == 1 ? F::sub(t3, x3) : ppp));
    const fe ysub = shfl(t4, q), xs = shfl(x3, q + 1);
    fe res = pick(k, xs, F::sub(t4, ysub), t3, t4);
    // same x (doubling or cancellation): rare, one lane of the quad runs the complete serial formula
    const bool rare = pz && !infA && !infB;
    if (__any_sync(FULL, rare)) {
      typename G::acc A, B;
      A.X = shfl(a, q); A.Y = shfl(a, q + 1); A.ZZ = shfl(a, q + 2); A.ZZZ = shfl(a, q + 3);
      B.X = shfl(b, q); B.Y = shfl(b, q + 1); B.ZZ = shfl(b, q + 2); B.ZZZ = shfl(b, q + 3);
      if (rare && k == 0) A = G::add(A, B);
      const fe r0 = shfl(A.X, q), r1 = shfl(A.Y, q), r2 = shfl(A.ZZ, q), r3 = shfl(A.ZZZ, q);
      if (rare) res = pick(k, r0, r1, r2, r3);
    }
    if (infB) res = a;
    if (infA) res = b;
    return res;
  }
};

}  // namespace mgb
