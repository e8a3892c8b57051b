// One-warp point arithmetic for the Horner kernels (XYZZ Weierstrass, extended twisted Edwards).
//
// A dependent chain of point operations (Horner over the windows, assembly of a window sum) has ONE point operation in
// flight.  Round 1 ran it on a 128-thread block: the four warps took the (up to four) products of a formula level, each
// spread over 12 lanes of its warp, and met at a block barrier after every level.  Here the whole point lives in ONE warp: 8-lane group g holds
// coordinate g (X, Y, ZZ, ZZZ), one 64-bit digit per lane (WarpField2), the four products of a
// level are one WarpField2::mul, the additions / subtractions between them run on the distributed
// digits (carry lookahead over ballots), and operands move between groups by shuffles -- no
// shared memory and no barrier anywhere in the chain.
#pragma once
#include "ec.cuh"
#include "warp.cuh"

namespace mgb {

template <class P>
struct OneWarpWeierstrass {
  typedef WarpField2<P> WF;
  typedef typename WF::u64 u64;
  typedef Field<P> F;
  typedef Fe<P> fe;
  typedef Weierstrass<P> G;
  static constexpr int N = P::N, D = WF::D, W = WF::W;

  MGB_DEV static int group() { return warp::lane() >> 3; }
  // the element held by group sg, on every group (digit l stays on digit lane l)
  MGB_DEV static u64 from_group(u64 v, int sg) {
    const int src = sg * W + (warp::lane() & (W - 1));
    return ((u64)warp::shfl((uint32_t)(v >> 32), src, 32) << 32) | warp::shfl((uint32_t)v, src, 32);
  }
  MGB_DEV static u64 pick(int g, u64 a0, u64 a1, u64 a2, u64 a3) { return g == 0 ? a0 : (g == 1 ? a1 : (g == 2 ? a2 : a3)); }
  // bit g = "the element of group g is zero", the same on all 32 lanes
  MGB_DEV static uint32_t zero_mask(u64 v) {
    const uint32_t nz = warp::ballot(v != 0);
    return ((nz & 0xffu) == 0 ? 1u : 0u) | ((nz & 0xff00u) == 0 ? 2u : 0u) | ((nz & 0xff0000u) == 0 ? 4u : 0u) | ((nz & 0xff000000u) == 0 ? 8u : 0u);
  }
  // digit of a full element for this lane (0 on the spare lanes)
  MGB_DEV static u64 digit_of(const fe& a) {
    const int l = warp::lane() & (W - 1);
    u64 r = 0;
    _Pragma("unroll") for (int k = 0; k < D; k++) r = (l == k) ? (((u64)a.v[2 * k + 1] << 32) | a.v[2 * k]) : r;
    return r;
  }
  MGB_DEV static u64 spread(const typename G::acc& A) {
    return pick(group(), digit_of(A.X), digit_of(A.Y), digit_of(A.ZZ), digit_of(A.ZZZ));
  }
  MGB_DEV static typename G::acc gather(u64 v) {      // the whole point on every lane
    typename G::acc A;
    fe* c[4] = {&A.X, &A.Y, &A.ZZ, &A.ZZZ};
    _Pragma("unroll") for (int g = 0; g < 4; g++) {
      _Pragma("unroll") for (int k = 0; k < D; k++) {
        const int src = g * W + k;
        const uint32_t lo = warp::shfl((uint32_t)v, src, 32), hi = warp::shfl((uint32_t)(v >> 32), src, 32);
        c[g]->v[2 * k] = lo; c[g]->v[2 * k + 1] = hi;
      }
    }
    return A;
  }
  MGB_DEV static u64 neutral() { return spread(G::acc_zero()); }

  // 2P  (dbl-2008-s-1, a = 0): three products levels
  MGB_DEV static u64 dbl(u64 v) {
    const int g = group();
    if (zero_mask(v) & (4u | 2u)) return neutral();                  // infinity (ZZ = 0) or 2-torsion (Y = 0)
    const u64 dv = WF::dbl(v);                                       // group 1: U = 2Y
    const u64 a1 = g == 1 ? dv : v;
    const u64 t1 = WF::mul(a1, a1);                                  // [XX, V = U^2, *, *]
    const u64 m = WF::add(WF::dbl(t1), t1);                          // group 0: M = 3 XX
    const u64 vb = from_group(t1, 1), xb = from_group(v, 0);
    const u64 t2 = WF::mul(pick(g, m, dv, v, xb), g == 0 ? m : vb);  // [M^2, W = U V, ZZ' = ZZ V, S = X V]
    const u64 sb = from_group(t2, 3), wb = from_group(t2, 1);
    const u64 x3 = WF::sub(t2, WF::dbl(sb));                         // group 0: X' = M^2 - 2S
    const u64 sx = WF::sub(sb, x3);                                  // group 0: S - X'
    const u64 t3 = WF::mul(pick(g, m, t2, v, wb), g == 0 ? sx : v);  // [M (S - X'), W Y, *, ZZZ' = W ZZZ]
    const u64 y3 = WF::sub(from_group(t3, 0), t3);                   // group 1: Y' = M (S - X') - W Y
    return pick(g, x3, y3, t2, t3);
  }

  // A + B  (add-2008-s, the schedule of QuadWeierstrass::add); complete: the rare cases take the serial formula
  MGB_DEV static u64 add(u64 a, u64 b) {
    const int g = group();
    const bool infA = (zero_mask(a) & 4u) != 0, infB = (zero_mask(b) & 4u) != 0;
    // level 1:  g0: U1 = X1 ZZ2   g1: S1 = Y1 ZZZ2   g2: U2 = ZZ1 X2   g3: S2 = ZZZ1 Y2
    const u64 t1 = WF::mul(a, from_group(b, g ^ 2));
    // level 2:  g0: PP = (U2 - U1)^2   g1: RR = (S2 - S1)^2   g2: ZZ1 ZZ2   g3: ZZZ1 ZZZ2
    const u64 diff = WF::sub(from_group(t1, g ^ 2), t1);             // g0: P, g1: R
    const bool pz = (zero_mask(diff) & 1u) != 0;
    const u64 t2 = WF::mul(g < 2 ? diff : a, g < 2 ? diff : b);
    // level 3:  g0: PPP = P PP   g1: Q = U1 PP   g2: ZZ3 = ZZ1 ZZ2 PP
    const u64 pp = from_group(t2, 0), u1 = from_group(t1, 0);
    const u64 t3 = WF::mul(g == 0 ? diff : (g == 1 ? u1 : t2), g == 0 ? t2 : pp);
    // level 4:  g0: S1 PPP   g1: R (Q - X3), X3 = RR - PPP - 2Q   g3: ZZZ3 = ZZZ1 ZZZ2 PPP
    const u64 ppp = from_group(t3, 0), s1 = from_group(t1, 1);
    const u64 x3 = WF::sub(WF::sub(t2, ppp), WF::dbl(t3));           // on g1
    const u64 qx = WF::sub(t3, x3);                                  // on g1: Q - X3
    const u64 t4 = WF::mul(g == 0 ? s1 : (g == 1 ? diff : t2), g == 0 ? t3 : (g == 1 ? qx : ppp));
    const u64 ysub = from_group(t4, 0), xs = from_group(x3, 1);
    u64 res = pick(g, xs, WF::sub(t4, ysub), t3, t4);
    if (pz && !infA && !infB) res = spread(G::add(gather(a), gather(b)));   // same x: doubling or cancellation (uniform branch)
    if (infB) res = a;
    if (infA) res = b;
    return res;
  }
};

// The same for extended twisted-Edwards points (a = -1): 8-lane group g holds coordinate g of (X, Y, Z, T).  A doubling
// is dbl-2008-hwcd -- four squares, then four products: TWO product levels instead of the three of the unified
// addition round 1's block-cooperative routine ran through shared memory (5.6 us per doubling there: the 2^18-point
// ed-on-BLS12-377 MSM spent 1.33 of its 2.73 ms in the 238 doublings of its Horner chain; now 0.31 ms).  The
// addition is the strongly unified add-2008-hwcd-3 (complete: doubling, neutral element, inverses), three levels.
template <class P, class C>
struct OneWarpTwistedEdwards {
  typedef WarpField2<P> WF;
  typedef typename WF::u64 u64;
  typedef Field<P> F;
  typedef Fe<P> fe;
  typedef TwistedEdwards<P, C> G;
  static constexpr int N = P::N, D = WF::D, W = WF::W;

  MGB_DEV static int group() { return warp::lane() >> 3; }
  MGB_DEV static u64 from_group(u64 v, int sg) {
    const int src = sg * W + (warp::lane() & (W - 1));
    return ((u64)warp::shfl((uint32_t)(v >> 32), src, 32) << 32) | warp::shfl((uint32_t)v, src, 32);
  }
  MGB_DEV static u64 pick(int g, u64 a0, u64 a1, u64 a2, u64 a3) { return g == 0 ? a0 : (g == 1 ? a1 : (g == 2 ? a2 : a3)); }
  MGB_DEV static u64 digit_of(const fe& a) {
    const int l = warp::lane() & (W - 1);
    u64 r = 0;
    _Pragma("unroll") for (int k = 0; k < D; k++) r = (l == k) ? (((u64)a.v[2 * k + 1] << 32) | a.v[2 * k]) : r;
    return r;
  }
  MGB_DEV static u64 spread(const typename G::acc& A) { return pick(group(), digit_of(A.X), digit_of(A.Y), digit_of(A.Z), digit_of(A.T)); }
  MGB_DEV static typename G::acc gather(u64 v) {      // the whole point on every lane
    typename G::acc A;
    fe* c[4] = {&A.X, &A.Y, &A.Z, &A.T};
    _Pragma("unroll") for (int g = 0; g < 4; g++) {
      _Pragma("unroll") for (int k = 0; k < D; k++) {
        const int src = g * W + k;
        const uint32_t lo = warp::shfl((uint32_t)v, src, 32), hi = warp::shfl((uint32_t)(v >> 32), src, 32);
        c[g]->v[2 * k] = lo; c[g]->v[2 * k + 1] = hi;
      }
    }
    return A;
  }

  // [E F, G H, F G, E H] = (X3, Y3, Z3, T3) from E, F, G, H known on every group
  MGB_DEV static u64 finish(u64 E, u64 Ff, u64 Gg, u64 H) {
    const int g = group();
    return WF::mul(pick(g, E, Gg, Ff, E), pick(g, Ff, H, Gg, H));
  }
  // 2P  (dbl-2008-hwcd with a = -1): A = X^2, B = Y^2, C = 2 Z^2, E = (X + Y)^2 - A - B, G = B - A, F = G - C, H = -A - B
  MGB_DEV static u64 dbl(u64 v) {
    const int g = group();
    const u64 xb = from_group(v, 0), yb = from_group(v, 1);
    const u64 xy = WF::add(xb, yb);                                  // (every lane runs the ballots inside)
    const u64 a1 = g == 3 ? xy : v;                                  // [X, Y, Z, X + Y]
    const u64 t1 = WF::mul(a1, a1);                                  // [A, B, ZZ, (X + Y)^2]
    const u64 A = from_group(t1, 0), B = from_group(t1, 1), ZZ = from_group(t1, 2), S = from_group(t1, 3);
    const u64 AB = WF::add(A, B);
    const u64 E = WF::sub(S, AB), Gg = WF::sub(B, A);
    const u64 Ff = WF::sub(Gg, WF::dbl(ZZ)), H = WF::sub(0ull, AB);
    return finish(E, Ff, Gg, H);
  }
  // P + Q  (add-2008-hwcd-3): A = (Y1 - X1)(Y2 - X2), B = (Y1 + X1)(Y2 + X2), C = k T1 T2, D = 2 Z1 Z2
  MGB_DEV static u64 add(u64 a, u64 b) {
    const int g = group();
    const u64 x1 = from_group(a, 0), y1 = from_group(a, 1), x2 = from_group(b, 0), y2 = from_group(b, 1);
    const u64 l1 = pick(g, WF::sub(y1, x1), WF::add(y1, x1), a, a);  // [Y1 - X1, Y1 + X1, Z1, T1]
    const u64 r1 = pick(g, WF::sub(y2, x2), WF::add(y2, x2), b, b);
    const u64 t1 = WF::mul(l1, r1);                                  // [A, B, Z1 Z2, T1 T2]
    const u64 t2 = WF::mul(t1, digit_of(G::k2d()));                  // group 3: C = k T1 T2
    const u64 A = from_group(t1, 0), B = from_group(t1, 1), Cc = from_group(t2, 3);
    const u64 Dd = WF::dbl(from_group(t1, 2));
    return finish(WF::sub(B, A), WF::sub(Dd, Cc), WF::add(Dd, Cc), WF::add(B, A));
  }
};

}  // namespace mgb
