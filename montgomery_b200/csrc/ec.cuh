// Group operations for the two curve families of the MSM path.
//
//  * Short Weierstrass y^2 = x^3 + b (a = 0): affine chord/tangent pieces for the batched-affine
//    accumulation (reference: src/wasm/curve.ts:32-84, src/curve-affine.ts:90-109,376-458) and an
//    extended-Jacobian "XYZZ" accumulator (x = X/ZZ, y = Y/ZZZ) for the bucket / window reduction.
//    The reference uses homogeneous projective coordinates there (src/curve-projective.ts:51-253);
//    XYZZ has a cheaper mixed add and the result is the same group element.
//  * Twisted Edwards -x^2 + y^2 = 1 + d x^2 y^2 in extended coordinates, add-2008-hwcd-3 with
//    k = 2d (reference: src/curve-twisted-edwards.ts:84-165).
// All adds here are complete for the inputs the pipeline can produce (infinity, P = Q, P = -Q).
#pragma once
#include "field.cuh"

namespace mgb {

// ---------------------------------------------------------------- short Weierstrass, a = 0
template <class P>
struct AffinePt {  // infinity is flagged by bit 31 of x's top limb (free in all supported fields)
  Fe<P> x, y;
};

template <class P>
struct XyzzPt {  // infinity iff ZZ == 0
  Fe<P> X, Y, ZZ, ZZZ;
};

template <class P>
struct Weierstrass {
  typedef Field<P> F;
  typedef Fe<P> fe;
  typedef AffinePt<P> affine;
  typedef XyzzPt<P> acc;
  static constexpr int N = P::N;
  static constexpr uint32_t INF_BIT = 0x80000000u;

  MGB_DEV static bool is_inf(const affine& p) { return (p.x.v[N - 1] & INF_BIT) != 0; }
  MGB_DEV static affine affine_inf() { affine r; r.x = F::zero(); r.y = F::zero(); r.x.v[N - 1] = INF_BIT; return r; }
  MGB_DEV static affine neg(const affine& p) { affine r; r.x = p.x; r.y = F::neg(p.y); return r; }

  MGB_DEV static acc acc_zero() { acc r; r.X = F::zero(); r.Y = F::one(); r.ZZ = F::zero(); r.ZZZ = F::zero(); return r; }
  MGB_DEV static bool acc_is_zero(const acc& p) { return F::is_zero(p.ZZ); }
  MGB_DEV static acc from_affine(const affine& p) {
    if (is_inf(p)) return acc_zero();
    acc r; r.X = p.x; r.Y = p.y; r.ZZ = F::one(); r.ZZZ = F::one(); return r;
  }
  MGB_DEV static acc acc_neg(const acc& p) { acc r = p; r.Y = F::neg(p.Y); return r; }

  // dbl-2008-s-1 with a = 0
  MGB_DEV static acc dbl(const acc& p) {
    if (acc_is_zero(p)) return p;
    fe U = F::dbl(p.Y);
    if (F::is_zero(U)) return acc_zero();  // order-2 point
    fe V = F::sqr(U);
    fe W = F::mul(U, V);
    fe S = F::mul(p.X, V);
    fe XX = F::sqr(p.X);
    fe M = F::add(F::dbl(XX), XX);
    acc r;
    r.X = F::sub(F::sqr(M), F::dbl(S));
    r.Y = F::sub(F::mul(M, F::sub(S, r.X)), F::mul(W, p.Y));
    r.ZZ = F::mul(V, p.ZZ);
    r.ZZZ = F::mul(W, p.ZZZ);
    return r;
  }

  // add-2008-s, complete
  MGB_DEV static acc add(const acc& p, const acc& q) {
    if (acc_is_zero(p)) return q;
    if (acc_is_zero(q)) return p;
    fe U1 = F::mul(p.X, q.ZZ);
    fe U2 = F::mul(q.X, p.ZZ);
    fe S1 = F::mul(p.Y, q.ZZZ);
    fe S2 = F::mul(q.Y, p.ZZZ);
    fe Pd = F::sub(U2, U1);
    fe R = F::sub(S2, S1);
    if (F::is_zero(Pd)) {
      if (F::is_zero(R)) return dbl(p);
      return acc_zero();
    }
    fe PP = F::sqr(Pd);
    fe PPP = F::mul(Pd, PP);
    fe Q = F::mul(U1, PP);
    acc r;
    r.X = F::sub(F::sub(F::sqr(R), PPP), F::dbl(Q));
    r.Y = F::sub(F::mul(R, F::sub(Q, r.X)), F::mul(S1, PPP));
    r.ZZ = F::mul(F::mul(p.ZZ, q.ZZ), PP);
    r.ZZZ = F::mul(F::mul(p.ZZZ, q.ZZZ), PPP);
    return r;
  }

  // madd-2008-s, complete
  MGB_DEV static acc madd(const acc& p, const affine& q) {
    if (is_inf(q)) return p;
    if (acc_is_zero(p)) return from_affine(q);
    fe U2 = F::mul(q.x, p.ZZ);
    fe S2 = F::mul(q.y, p.ZZZ);
    fe Pd = F::sub(U2, p.X);
    fe R = F::sub(S2, p.Y);
    if (F::is_zero(Pd)) {
      if (F::is_zero(R)) return dbl(p);
      return acc_zero();
    }
    fe PP = F::sqr(Pd);
    fe PPP = F::mul(Pd, PP);
    fe Q = F::mul(p.X, PP);
    acc r;
    r.X = F::sub(F::sub(F::sqr(R), PPP), F::dbl(Q));
    r.Y = F::sub(F::mul(R, F::sub(Q, r.X)), F::mul(p.Y, PPP));
    r.ZZ = F::mul(p.ZZ, PP);
    r.ZZZ = F::mul(p.ZZZ, PPP);
    return r;
  }

  // mmadd-2008-s: both operands affine (ZZ = ZZZ = 1), result XYZZ; 4M + 2S instead of the 8M + 2S of madd.  Complete.
  MGB_DEV static acc mmadd(const affine& p, const affine& q) {
    if (is_inf(q)) return from_affine(p);
    if (is_inf(p)) return from_affine(q);
    fe Pd = F::sub(q.x, p.x);
    fe R = F::sub(q.y, p.y);
    if (F::is_zero(Pd)) {
      if (F::is_zero(R)) return dbl(from_affine(p));
      return acc_zero();
    }
    fe PP = F::sqr(Pd);
    fe PPP = F::mul(Pd, PP);
    fe Q = F::mul(p.x, PP);
    acc r;
    r.X = F::sub(F::sub(F::sqr(R), PPP), F::dbl(Q));
    r.Y = F::sub(F::mul(R, F::sub(Q, r.X)), F::mul(p.y, PPP));
    r.ZZ = PP;
    r.ZZZ = PPP;
    return r;
  }

  // x = X/ZZ, y = Y/ZZZ with one inversion: t = 1/ZZZ, 1/ZZ = t^2 * ZZ^2
  MGB_DEV static affine to_affine(const acc& p) {
    if (acc_is_zero(p)) return affine_inf();
    fe t = F::inv_divsteps(p.ZZZ);
    fe izz = F::mul(F::sqr(t), F::sqr(p.ZZ));
    affine r;
    r.x = F::mul(p.X, izz);
    r.y = F::mul(p.Y, t);
    return r;
  }

  // ---- pieces of the batched-affine addition S = A + B with a shared inversion
  // kind: 0 = generic chord (den = xB - xA), 1 = tangent (A == B, den = 2y), 2 = result is A,
  //       3 = result is B, 4 = result is infinity.  For kinds 2..4 den is 1 (keeps the batch product invertible).
  MGB_DEV static int add_prepare(const affine& A, const affine& B, fe& den) {
    if (is_inf(B)) { den = F::one(); return 2; }
    if (is_inf(A)) { den = F::one(); return 3; }
    den = F::sub(B.x, A.x);
    if (F::is_zero(den)) {
      if (F::eq(A.y, B.y) && !F::is_zero(A.y)) { den = F::dbl(A.y); return 1; }
      den = F::one();
      return 4;
    }
    return 0;
  }
  // Fast path of add_prepare from the x coordinates alone: returns false (den untouched) when an
  // operand is infinity or xA == xB, in which case the caller falls back to add_prepare.
  MGB_DEV static bool prepare_x(const fe& xa, const fe& xb, fe& den) {
    if ((xa.v[N - 1] | xb.v[N - 1]) & INF_BIT) return false;
    fe d = F::sub(xb, xa);
    if (F::is_zero(d)) return false;
    den = d;
    return true;
  }
  // INL = true inlines the three multiplications (hot batched-addition loop: no call marshalling)
  template <bool INL = false>
  MGB_DEV static affine add_finish(int kind, const affine& A, const affine& B, const fe& inv_den) {
    if (kind == 2) return A;
    if (kind == 3) return B;
    if (kind == 4) return affine_inf();
    fe num;
    if (kind == 1) { fe xx = F::sqr(A.x); num = F::add(F::dbl(xx), xx); }
    else num = F::sub(B.y, A.y);
    fe m = INL ? F::mul_inl(num, inv_den) : F::mul(num, inv_den);
    affine r;
    fe mm = INL ? F::mul_inl(m, m) : F::sqr(m);
    r.x = F::sub(F::sub(mm, A.x), B.x);   // kind 1: B == A, so this is m^2 - 2x
    fe t = F::sub(A.x, r.x);
    r.y = F::sub(INL ? F::mul_inl(m, t) : F::mul(m, t), A.y);
    return r;
  }
};

// ---------------------------------------------------------------- twisted Edwards, a = -1
template <class P>
struct ExtPt {  // (X : Y : Z : T), x = X/Z, y = Y/Z, T = XY/Z; neutral = (0, 1, 1, 0)
  Fe<P> X, Y, Z, T;
};
template <class P>
struct TeAffine {  // input table entry: x, y and kt = 2d*x*y (Montgomery form)
  Fe<P> x, y, kt;
};

template <class P, class C>
struct TwistedEdwards {
  typedef Field<P> F;
  typedef Fe<P> fe;
  typedef ExtPt<P> acc;
  typedef TeAffine<P> affine;
  static constexpr int N = P::N;

  MGB_DEV static fe k2d() { fe r; _Pragma("unroll") for (int i = 0; i < N; i++) r.v[i] = C::k(i); return r; }
  MGB_DEV static acc acc_zero() { acc r; r.X = F::zero(); r.Y = F::one(); r.Z = F::one(); r.T = F::zero(); return r; }
  MGB_DEV static acc acc_neg(const acc& p) { acc r = p; r.X = F::neg(p.X); r.T = F::neg(p.T); return r; }
  MGB_DEV static affine neg(const affine& p) { affine r; r.x = F::neg(p.x); r.y = p.y; r.kt = F::neg(p.kt); return r; }

  // add-2008-hwcd-3, strongly unified (9M)
  MGB_DEV static acc add(const acc& p, const acc& q) {
    fe A = F::mul(F::sub(p.Y, p.X), F::sub(q.Y, q.X));
    fe B = F::mul(F::add(p.Y, p.X), F::add(q.Y, q.X));
    fe Cc = F::mul(F::mul(p.T, q.T), k2d());
    fe D = F::dbl(F::mul(p.Z, q.Z));
    fe E = F::sub(B, A), Ff = F::sub(D, Cc), G = F::add(D, Cc), H = F::add(B, A);
    acc r;
    r.X = F::mul(E, Ff); r.Y = F::mul(G, H); r.T = F::mul(E, H); r.Z = F::mul(Ff, G);
    return r;
  }
  MGB_DEV static acc dbl(const acc& p) { return add(p, p); }
  // mixed add, Z2 = 1 and k*T2 precomputed (7M)
  MGB_DEV static acc madd(const acc& p, const affine& q) {
    fe A = F::mul(F::sub(p.Y, p.X), F::sub(q.y, q.x));
    fe B = F::mul(F::add(p.Y, p.X), F::add(q.y, q.x));
    fe Cc = F::mul(p.T, q.kt);
    fe D = F::dbl(p.Z);
    fe E = F::sub(B, A), Ff = F::sub(D, Cc), G = F::add(D, Cc), H = F::add(B, A);
    acc r;
    r.X = F::mul(E, Ff); r.Y = F::mul(G, H); r.T = F::mul(E, H); r.Z = F::mul(Ff, G);
    return r;
  }
  // both operands affine table entries (Z1 = Z2 = 1), 6M
  MGB_DEV static acc add_affine(const affine& p, const affine& q) {
    fe A = F::mul(F::sub(p.y, p.x), F::sub(q.y, q.x));
    fe B = F::mul(F::add(p.y, p.x), F::add(q.y, q.x));
    fe Cc = F::mul(F::mul(p.x, p.y), q.kt);
    fe D = F::dbl(F::one());
    fe E = F::sub(B, A), Ff = F::sub(D, Cc), G = F::add(D, Cc), H = F::add(B, A);
    acc r;
    r.X = F::mul(E, Ff); r.Y = F::mul(G, H); r.T = F::mul(E, H); r.Z = F::mul(Ff, G);
    return r;
  }
  MGB_DEV static acc from_affine(const affine& p) {
    acc r; r.X = p.x; r.Y = p.y; r.Z = F::one(); r.T = F::mul(p.x, p.y); return r;
  }
  // -> canonical affine (x, y); Z != 0 for every point of the odd-order subgroup times cofactor 4 group
  MGB_DEV static void to_affine(const acc& p, fe& x, fe& y) {
    fe zi = F::inv_divsteps(p.Z);
    x = F::mul(p.X, zi);
    y = F::mul(p.Y, zi);
  }
};

}  // namespace mgb
