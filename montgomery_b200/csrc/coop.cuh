// Block-cooperative point doubling / addition for the serial tail of the MSM (Horner over the
// windows: ~110 dependent doublings, reference msm-batched-affine.ts:322-334).
//
// A dependent chain of field multiplications on ONE warp is bound by that warp's SMSP multiplier
// pipe (277 IMAD.WIDE x 4 cycles per 377-bit product), so instruction-level parallelism inside a
// thread buys nothing.  The four warps of a 128-thread block sit on the four SMSPs of an SM: each
// takes one of the independent multiplications of a formula level and the operands travel through
// shared memory.  XYZZ doubling = 3 levels instead of 9 sequential products, addition = 4 instead
// of 14; extended twisted-Edwards addition = 3 instead of 9.
//
// Every one of those products is itself spread over the lanes of its warp (warp.cuh: one limb per lane, ~12
// dependent steps instead of 277 dependent MADs; 554 ns instead of 974 ns for a lone 377-bit product); the warp's
// lane 0 does the additions / subtractions between the products.
#pragma once
#include "ec.cuh"
#include "warp.cuh"

namespace mgb {

template <class P>
struct CoopMem {  // slots of N limbs in shared memory
  uint32_t* base;
  MGB_DEV Fe<P> ld(int slot) const {
    Fe<P> r;
    _Pragma("unroll") for (int i = 0; i < P::N; i++) r.v[i] = base[slot * P::N + i];
    return r;
  }
  MGB_DEV void st(int slot, const Fe<P>& a) const {
    _Pragma("unroll") for (int i = 0; i < P::N; i++) base[slot * P::N + i] = a.v[i];
  }
  // slot[out] = slot[a] * slot[b] by ALL lanes of the calling warp (lane l moves limb l); what lane 0 wrote
  // before the call is visible to the others, and the result is visible to lane 0 after it.  Out of line:
  // one copy of the product per kernel, so the Horner loop stays inside the instruction cache.
  MGB_NOINLINE_DEV static void wmul_impl(uint32_t* base, int out, int a, int b) {
    const int l = threadIdx.x & 31;
    __syncwarp();
    const uint32_t x = l < P::N ? base[a * P::N + l] : 0u, y = l < P::N ? base[b * P::N + l] : 0u;
    const uint32_t r = WarpField<P>::mul(x, y);
    if (l < P::N) base[out * P::N + l] = r;
    __syncwarp();
  }
  MGB_DEV void wmul(int out, int a, int b) const { wmul_impl(base, out, a, b); }
};

// slot map: 0..3 = P (accumulator), 4..7 = Q (second operand), 8.. = temporaries, flags after the slots
static constexpr int COOP_SLOTS = 24;

template <class P>
struct CoopWeierstrass {
  typedef Field<P> F;
  typedef Fe<P> fe;
  typedef Weierstrass<P> G;
  enum { X = 0, Y = 1, ZZ = 2, ZZZ = 3, X2 = 4, Y2 = 5, ZZ2 = 6, ZZZ2 = 7, T = 8 };

  // P <- 2^count P  (dbl-2008-s-1, a = 0).  Per doubling three formula levels and three barriers, nothing else:
  //  * the infinity / 2-torsion test runs on warp 2 during level 1 (the level's results are only temporaries);
  //  * X', ZZ', ZZZ' are written straight into the accumulator slots by the level that produces them (no other
  //    warp reads those slots in that level);
  //  * Y' = M (S - X') - W Y needs two products of level 3 from different warps: it stays pending as the pair
  //    (T+10, T+7) and the subtraction is done by the first reader -- lane 0 of warp 0 at the start of the next
  //    doubling's level 1 (warp 2 repeats it for its test) -- or by the epilogue after the last doubling.
  MGB_DEV static void dbl_n(CoopMem<P> m, volatile int* flag, int count) {
    const int warp = threadIdx.x >> 5;
    const bool act = (threadIdx.x & 31) == 0;
    for (int i = 0; i < count; i++) {
      if (warp == 0) {
        if (act) {
          fe y = m.ld(Y);
          if (i) { y = F::sub(m.ld(T + 10), m.ld(T + 7)); m.st(Y, y); }
          m.st(T + 0, F::dbl(y));
        }
        m.wmul(T + 1, T + 0, T + 0);                                         // U = 2Y, V = U^2
      }
      if (warp == 1) { m.wmul(T + 2, X, X); if (act) { fe xx = m.ld(T + 2); m.st(T + 2, F::add(F::dbl(xx), xx)); } }   // M = 3 X^2
      if (warp == 2 && act) {
        const fe y = i ? F::sub(m.ld(T + 10), m.ld(T + 7)) : m.ld(Y);
        *flag = (F::is_zero(m.ld(ZZ)) || F::is_zero(y)) ? 1 : 0;              // infinity or 2-torsion
      }
      __syncthreads();
      if (*flag) {                             // the result of this and of every further doubling is the neutral element
        if (threadIdx.x == 0) { m.st(X, F::zero()); m.st(Y, F::one()); m.st(ZZ, F::zero()); m.st(ZZZ, F::zero()); }
        __syncthreads();
        return;
      }
      if (warp == 0) m.wmul(T + 3, T + 0, T + 1);                            // W = U*V
      if (warp == 1) m.wmul(T + 4, X, T + 1);                                // S = X*V
      if (warp == 2) m.wmul(T + 5, T + 2, T + 2);                            // M^2
      if (warp == 3) m.wmul(ZZ, T + 1, ZZ);                                  // ZZ' = V*ZZ, in place
      __syncthreads();
      if (warp == 0) m.wmul(T + 7, T + 3, Y);                                // W*Y
      if (warp == 1) m.wmul(ZZZ, T + 3, ZZZ);                                // ZZZ' = W*ZZZ, in place
      if (warp == 2) {
        if (act) {
          fe S = m.ld(T + 4);
          fe x3 = F::sub(m.ld(T + 5), F::dbl(S));
          m.st(X, x3);                                                       // X', in place
          m.st(T + 11, F::sub(S, x3));
        }
        m.wmul(T + 10, T + 2, T + 11);                                       // M*(S - X')
      }
      __syncthreads();
    }
    if (count > 0) {
      if (threadIdx.x == 0) m.st(Y, F::sub(m.ld(T + 10), m.ld(T + 7)));
      __syncthreads();
    }
  }

  // P <- P + Q  (add-2008-s; the degenerate cases fall back to the complete serial formula)
  MGB_DEV static void add(CoopMem<P> m, volatile int* flag) {
    const int warp = threadIdx.x >> 5;
    const bool act = (threadIdx.x & 31) == 0;
    if (threadIdx.x == 0) *flag = (F::is_zero(m.ld(ZZ)) ? 1 : 0) | (F::is_zero(m.ld(ZZ2)) ? 2 : 0);
    __syncthreads();
    int f = *flag;
    if (f) {
      if (threadIdx.x == 0 && (f & 1) && !(f & 2)) { m.st(X, m.ld(X2)); m.st(Y, m.ld(Y2)); m.st(ZZ, m.ld(ZZ2)); m.st(ZZZ, m.ld(ZZZ2)); }
      __syncthreads();
      return;
    }
    if (warp == 0) m.wmul(T + 0, X, ZZ2);      // U1
    if (warp == 1) m.wmul(T + 1, X2, ZZ);      // U2
    if (warp == 2) m.wmul(T + 2, Y, ZZZ2);     // S1
    if (warp == 3) m.wmul(T + 3, Y2, ZZZ);     // S2
    __syncthreads();
    if (threadIdx.x == 0) *flag = F::is_zero(F::sub(m.ld(T + 1), m.ld(T + 0))) ? 1 : 0;
    __syncthreads();
    if (*flag) {   // same x: doubling or cancellation -- rare, serial
      if (threadIdx.x == 0) {
        typename G::acc a, b;
        a.X = m.ld(X); a.Y = m.ld(Y); a.ZZ = m.ld(ZZ); a.ZZZ = m.ld(ZZZ);
        b.X = m.ld(X2); b.Y = m.ld(Y2); b.ZZ = m.ld(ZZ2); b.ZZZ = m.ld(ZZZ2);
        a = G::add(a, b);
        m.st(X, a.X); m.st(Y, a.Y); m.st(ZZ, a.ZZ); m.st(ZZZ, a.ZZZ);
      }
      __syncthreads();
      return;
    }
    if (warp == 0) { if (act) m.st(T + 4, F::sub(m.ld(T + 1), m.ld(T + 0))); m.wmul(T + 5, T + 4, T + 4); }   // P, PP
    if (warp == 1) { if (act) m.st(T + 6, F::sub(m.ld(T + 3), m.ld(T + 2))); m.wmul(T + 7, T + 6, T + 6); }   // R, RR
    if (warp == 2) m.wmul(T + 8, ZZ, ZZ2);
    if (warp == 3) m.wmul(T + 9, ZZZ, ZZZ2);
    __syncthreads();
    if (warp == 0) m.wmul(T + 10, T + 4, T + 5);    // PPP
    if (warp == 1) m.wmul(T + 11, T + 0, T + 5);    // Q = U1*PP
    if (warp == 2) m.wmul(T + 12, T + 8, T + 5);    // ZZ3
    __syncthreads();
    if (warp == 0) {
      if (act) {
        fe Q = m.ld(T + 11);
        fe x3 = F::sub(F::sub(m.ld(T + 7), m.ld(T + 10)), F::dbl(Q));
        m.st(T + 13, x3);
        m.st(T + 5, F::sub(Q, x3));                 // (T+5 = PP is dead)
      }
      m.wmul(T + 14, T + 6, T + 5);                 // R*(Q - X3)
    }
    if (warp == 1) m.wmul(T + 15, T + 2, T + 10);   // S1*PPP
    if (warp == 2) m.wmul(T + 4, T + 9, T + 10);    // ZZZ3 (T+4 = P is dead)
    __syncthreads();
    if (threadIdx.x == 0) {
      m.st(X, m.ld(T + 13));
      m.st(Y, F::sub(m.ld(T + 14), m.ld(T + 15)));
      m.st(ZZ, m.ld(T + 12));
      m.st(ZZZ, m.ld(T + 4));
    }
    __syncthreads();
  }
};


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

template <class P, class C>
struct CoopTwistedEdwards {
  typedef Field<P> F;
  typedef Fe<P> fe;
  typedef TwistedEdwards<P, C> G;
  enum { X = 0, Y = 1, Z = 2, Tt = 3, T = 8 };

  // P <- P + Q with Q at slot qbase (4 = second operand, 0 = P itself, i.e. doubling); unified, complete
  MGB_DEV static void add_from(CoopMem<P> m, int qb) {
    const int warp = threadIdx.x >> 5;
    const bool act = (threadIdx.x & 31) == 0;
    const int t0 = T + 4 + 2 * warp, t1 = t0 + 1;       // two private operand slots per warp
    if (act) {
      if (warp == 0) { m.st(t0, F::sub(m.ld(Y), m.ld(X))); m.st(t1, F::sub(m.ld(qb + Y), m.ld(qb + X))); }
      if (warp == 1) { m.st(t0, F::add(m.ld(Y), m.ld(X))); m.st(t1, F::add(m.ld(qb + Y), m.ld(qb + X))); }
      if (warp == 2) m.st(t1, G::k2d());
    }
    if (warp == 0) m.wmul(T + 0, t0, t1);                                          // A
    if (warp == 1) m.wmul(T + 1, t0, t1);                                          // B
    if (warp == 2) { m.wmul(t0, Tt, qb + Tt); m.wmul(T + 2, t0, t1); }             // C
    if (warp == 3) { m.wmul(T + 3, Z, qb + Z); if (act) m.st(T + 3, F::dbl(m.ld(T + 3))); }   // D
    __syncthreads();
    if (act) {
      fe A = m.ld(T + 0), B = m.ld(T + 1), Cc = m.ld(T + 2), D = m.ld(T + 3);
      fe E = F::sub(B, A), Ff = F::sub(D, Cc), Gg = F::add(D, Cc), H = F::add(B, A);
      if (warp == 0) { m.st(t0, E); m.st(t1, Ff); }
      if (warp == 1) { m.st(t0, Gg); m.st(t1, H); }
      if (warp == 2) { m.st(t0, E); m.st(t1, H); }
      if (warp == 3) { m.st(t0, Ff); m.st(t1, Gg); }
    }
    m.wmul(warp == 0 ? X : (warp == 1 ? Y : (warp == 2 ? Tt : Z)), t0, t1);
    __syncthreads();
  }
  MGB_DEV static void dbl_n(CoopMem<P> m, volatile int*, int count) { for (int i = 0; i < count; i++) add_from(m, 0); }
  MGB_DEV static void add(CoopMem<P> m, volatile int*) { add_from(m, 4); }
};

}  // namespace mgb
