// TEST-ONLY: compiles the engine's math headers (field.cuh, ec.cuh) for the host with an emulated
// carry flag, so the formulas can be checked against the oracle without a GPU.  Not shipped, not
// linked into the product library.
#define MGB_HOST_EMU 1
#include "simt_emu.h"
#include "../../montgomery_b200/csrc/ec.cuh"
#include "../../montgomery_b200/csrc/warp.cuh"
#include "../../montgomery_b200/csrc/coop.cuh"
#include "../../montgomery_b200/csrc/onewarp.cuh"
#include <cstring>
using namespace mgb;

template <class P> static Fe<P> ld(const uint32_t* p) { Fe<P> r; memcpy(r.v, p, sizeof(r.v)); return r; }
template <class P> static void st(uint32_t* p, const Fe<P>& a) { memcpy(p, a.v, sizeof(a.v)); }

template <class P> static void fe_op(int op, uint32_t* out, const uint32_t* a, const uint32_t* b) {
  typedef Field<P> F;
  Fe<P> x = ld<P>(a), y = ld<P>(b), r;
  switch (op) {
    case 0: r = F::mul(x, y); break;
    case 1: r = F::add(x, y); break;
    case 2: r = F::sub(x, y); break;
    case 3: r = F::inv(x); break;
    case 4: r = F::to_mont(x); break;
    case 5: r = F::from_mont(x); break;
    case 6: r = F::sqr(x); break;
    case 7: r = F::neg(x); break;
    case 8: r = F::inv_bgcd(x); break;
    case 9: r = F::inv_divsteps(x); break;
    default: r = F::zero();
  }
  st<P>(out, r);
}

// Weierstrass ops on Montgomery-form coordinates.  Affine = 2N limbs (+inf flag in x top bit), XYZZ = 4N limbs.
template <class P> static void w_op(int op, uint32_t* out, const uint32_t* a, const uint32_t* b) {
  typedef Weierstrass<P> W;
  typedef Field<P> F;
  constexpr int N = P::N;
  auto lda = [&](const uint32_t* p) { typename W::affine r; r.x = ld<P>(p); r.y = ld<P>(p + N); return r; };
  auto ldx = [&](const uint32_t* p) { typename W::acc r; r.X = ld<P>(p); r.Y = ld<P>(p + N); r.ZZ = ld<P>(p + 2 * N); r.ZZZ = ld<P>(p + 3 * N); return r; };
  auto sta = [&](const typename W::affine& r) { st<P>(out, r.x); st<P>(out + N, r.y); };
  auto stx = [&](const typename W::acc& r) { st<P>(out, r.X); st<P>(out + N, r.Y); st<P>(out + 2 * N, r.ZZ); st<P>(out + 3 * N, r.ZZZ); };
  switch (op) {
    case 0: {  // affine + affine through prepare / invert / finish
      auto A = lda(a), B = lda(b);
      Fe<P> den; int kind = W::add_prepare(A, B, den);
      sta(W::add_finish(kind, A, B, F::inv(den)));
      break;
    }
    case 1: stx(W::madd(ldx(a), lda(b))); break;
    case 2: stx(W::add(ldx(a), ldx(b))); break;
    case 3: stx(W::dbl(ldx(a))); break;
    case 4: sta(W::to_affine(ldx(a))); break;
    case 5: stx(W::from_affine(lda(a))); break;
  }
}

template <class P, class C> static void te_op(int op, uint32_t* out, const uint32_t* a, const uint32_t* b) {
  typedef TwistedEdwards<P, C> T;
  constexpr int N = P::N;
  auto lde = [&](const uint32_t* p) { typename T::acc r; r.X = ld<P>(p); r.Y = ld<P>(p + N); r.Z = ld<P>(p + 2 * N); r.T = ld<P>(p + 3 * N); return r; };
  auto lda = [&](const uint32_t* p) { typename T::affine r; r.x = ld<P>(p); r.y = ld<P>(p + N); r.kt = ld<P>(p + 2 * N); return r; };
  auto ste = [&](const typename T::acc& r) { st<P>(out, r.X); st<P>(out + N, r.Y); st<P>(out + 2 * N, r.Z); st<P>(out + 3 * N, r.T); };
  switch (op) {
    case 0: ste(T::add(lde(a), lde(b))); break;
    case 1: ste(T::madd(lde(a), lda(b))); break;
    case 2: ste(T::add_affine(lda(a), lda(b))); break;
    case 3: ste(T::dbl(lde(a))); break;
    case 4: { Fe<P> x, y; T::to_affine(lde(a), x, y); st<P>(out, x); st<P>(out + N, y); break; }
  }
}

// n products through the warp-cooperative multiplication (warp.cuh), two at a time: lanes 0-15 take pair 2j,
// lanes 16-31 pair 2j+1 (a zero pair when n is odd); operands and results are Montgomery-form limbs.
template <class P> static void warp_mul(uint32_t* out, const uint32_t* a, const uint32_t* b, int n) {
  constexpr int N = P::N;
  simt::run_warp([&](int lane) {
    const int g = lane >> 4, l = lane & 15;
    for (int j = 0; j < n; j += 2) {
      const int e = j + g;
      const bool live = e < n && l < N;
      // lanes above the top limb pass garbage on purpose: the routine must ignore it
      const uint32_t x = live ? a[e * N + l] : (e < n ? 0xdeadbeefu : 0u), y = live ? b[e * N + l] : 0x12345678u;
      const uint32_t r = WarpField<P>::mul(x, y);
      if (live) out[e * N + l] = r;
      else if (e < n && r != 0) out[e * N] ^= 0xffffffffu;   // poison: spare lanes must return 0
    }
  });
}

// n products through the two-limbs-per-lane variant, four at a time (8-lane groups)
template <class P> static void warp_mul2(uint32_t* out, const uint32_t* a, const uint32_t* b, int n) {
  constexpr int N = P::N, D = N / 2;
  typedef unsigned long long u64;
  simt::run_warp([&](int lane) {
    const int g = lane >> 3, l = lane & 7;
    for (int j = 0; j < n; j += 4) {
      const int e = j + g;
      const bool live = e < n && l < D;
      const u64 x = live ? (((u64)a[e * N + 2 * l + 1] << 32) | a[e * N + 2 * l]) : (e < n ? 0xdeadbeefcafef00dull : 0ull);
      const u64 y = live ? (((u64)b[e * N + 2 * l + 1] << 32) | b[e * N + 2 * l]) : 0x123456789abcdef0ull;
      const u64 r = WarpField2<P>::mul(x, y);
      if (live) { out[e * N + 2 * l] = (uint32_t)r; out[e * N + 2 * l + 1] = (uint32_t)(r >> 32); }
      else if (e < n && r != 0) out[e * N] ^= 0xffffffffu;   // poison: spare lanes must return 0
    }
  });
}

// op 0: a + b, 1: a - b, 2: 2a, 3: (a == 0) on distributed elements, four at a time
template <class P> static void warp_addsub2(int op, uint32_t* out, const uint32_t* a, const uint32_t* b, int n) {
  constexpr int N = P::N, D = N / 2;
  typedef unsigned long long u64;
  typedef WarpField2<P> WF;
  simt::run_warp([&](int lane) {
    const int g = lane >> 3, l = lane & 7;
    for (int j = 0; j < n; j += 4) {
      const int e = j + g;
      const bool live = e < n && l < D;
      const u64 x = live ? (((u64)a[e * N + 2 * l + 1] << 32) | a[e * N + 2 * l]) : 0ull;
      const u64 y = live ? (((u64)b[e * N + 2 * l + 1] << 32) | b[e * N + 2 * l]) : 0ull;
      const u64 r = op == 0 ? WF::add(x, y) : op == 1 ? WF::sub(x, y) : op == 2 ? WF::dbl(x) : (u64)WF::is_zero(x);
      if (live) { out[e * N + 2 * l] = (uint32_t)r; out[e * N + 2 * l + 1] = (uint32_t)(r >> 32); }
    }
  });
}

template <class P> static void warp_finish2(uint64_t* out, const uint64_t* t, const uint64_t* clo, const uint32_t* chi) {
  simt::run_warp([&](int lane) { out[lane] = WarpField2<P>::finish(t[lane], clo[lane], chi[lane]); });
}

// the carry / borrow resolution alone: lane l of group g gets t[g*16 + l] and c[g*16 + l] (c as 64-bit)
template <class P> static void warp_finish(uint32_t* out, const uint32_t* t, const uint64_t* c) {
  simt::run_warp([&](int lane) { out[lane] = WarpField<P>::finish(t[lane], c[lane]); });
}

// n inversions through the lane-parallel division-step inverse: every lane passes the same element
template <class P> static void warp_inv(uint32_t* out, const uint32_t* a, int n) {
  constexpr int N = P::N;
  simt::run_warp([&](int lane) {
    for (int j = 0; j < n; j++) {
      const Fe<P> r = WarpField<P>::inv(ld<P>(a + j * N));
      if (lane == (j & 31)) st<P>(out + j * N, r);                 // every lane holds the result: take a different one each time
    }
  });
}

// ---- the one-warp point arithmetic of the Horner kernels (onewarp.cuh), driven the way k_final drives it.
// op 0: P <- 2^count P, op 1: P <- P + Q, op 2: P <- 2^count P + Q (what one Horner step does)
// the same three operations on ONE emulated warp (onewarp.cuh): the point distributed over the lanes
template <class P> static void onewarp_op(int op, int count, uint32_t* out, const uint32_t* a, const uint32_t* b) {
  typedef OneWarpWeierstrass<P> OW;
  typedef Weierstrass<P> W;
  constexpr int N = P::N;
  auto ldx = [&](const uint32_t* p) { typename W::acc r; r.X = ld<P>(p); r.Y = ld<P>(p + N); r.ZZ = ld<P>(p + 2 * N); r.ZZZ = ld<P>(p + 3 * N); return r; };
  simt::run_warp([&](int lane) {
    auto v = OW::spread(ldx(a));
    const auto q = OW::spread(ldx(b));
    if (op == 0 || op == 2) for (int i = 0; i < count; i++) v = OW::dbl(v);
    if (op == 1 || op == 2) v = OW::add(v, q);
    const typename W::acc r = OW::gather(v);
    if (lane == 7) { st<P>(out, r.X); st<P>(out + N, r.Y); st<P>(out + 2 * N, r.ZZ); st<P>(out + 3 * N, r.ZZZ); }
  });
}

// extended twisted-Edwards points in one warp (OneWarpTwistedEdwards): same three operations
template <class P, class C> static void onewarp_te_op(int op, int count, uint32_t* out, const uint32_t* a, const uint32_t* b) {
  typedef OneWarpTwistedEdwards<P, C> OW;
  typedef TwistedEdwards<P, C> T;
  constexpr int N = P::N;
  auto ldx = [&](const uint32_t* p) { typename T::acc r; r.X = ld<P>(p); r.Y = ld<P>(p + N); r.Z = ld<P>(p + 2 * N); r.T = ld<P>(p + 3 * N); return r; };
  simt::run_warp([&](int lane) {
    auto v = OW::spread(ldx(a));
    const auto q = OW::spread(ldx(b));
    if (op == 0 || op == 2) for (int i = 0; i < count; i++) v = OW::dbl(v);
    if (op == 1 || op == 2) v = OW::add(v, q);
    const typename T::acc r = OW::gather(v);
    if (lane == 13) { st<P>(out, r.X); st<P>(out + N, r.Y); st<P>(out + 2 * N, r.Z); st<P>(out + 3 * N, r.T); }
  });
}

// quad-cooperative XYZZ addition: lane 4j + k holds coordinate k of the j-th of 8 independent additions
template <class P> static void quad_add(uint32_t* out, const uint32_t* a, const uint32_t* b) {
  constexpr int N = P::N;
  simt::run_warp([&](int lane) {
    const Fe<P> x = ld<P>(a + lane * N), y = ld<P>(b + lane * N);
    st<P>(out + lane * N, QuadWeierstrass<P>::add(x, y));
  });
}

extern "C" {
void emu_onewarp_w(int curve, int op, int count, uint32_t* out, const uint32_t* a, const uint32_t* b) {
  if (curve == 0) onewarp_op<Fp377>(op, count, out, a, b);
  else if (curve == 1) onewarp_op<FpPallas>(op, count, out, a, b);
  else onewarp_op<Fp381>(op, count, out, a, b);
}
void emu_onewarp_te(int op, int count, uint32_t* out, const uint32_t* a, const uint32_t* b) {
  onewarp_te_op<Fr377, Ed377Consts>(op, count, out, a, b);
}
void emu_quad_add(int curve, uint32_t* out, const uint32_t* a, const uint32_t* b) {
  if (curve == 0) quad_add<Fp377>(out, a, b);
  else if (curve == 1) quad_add<FpPallas>(out, a, b);
  else quad_add<Fp381>(out, a, b);
}
void emu_warp_inv(int field, uint32_t* out, const uint32_t* a, int n) {
  if (field == 0) warp_inv<Fp377>(out, a, n);
  else if (field == 1) warp_inv<Fr377>(out, a, n);
  else if (field == 2) warp_inv<FpPallas>(out, a, n);
  else warp_inv<Fp381>(out, a, n);
}
void emu_warp_mul2(int field, uint32_t* out, const uint32_t* a, const uint32_t* b, int n) {
  if (field == 0) warp_mul2<Fp377>(out, a, b, n);
  else if (field == 1) warp_mul2<Fr377>(out, a, b, n);
  else if (field == 2) warp_mul2<FpPallas>(out, a, b, n);
  else warp_mul2<Fp381>(out, a, b, n);
}
void emu_warp_addsub2(int field, int op, uint32_t* out, const uint32_t* a, const uint32_t* b, int n) {
  if (field == 0) warp_addsub2<Fp377>(op, out, a, b, n);
  else if (field == 1) warp_addsub2<Fr377>(op, out, a, b, n);
  else if (field == 2) warp_addsub2<FpPallas>(op, out, a, b, n);
  else warp_addsub2<Fp381>(op, out, a, b, n);
}
void emu_warp_finish2(int field, uint64_t* out, const uint64_t* t, const uint64_t* clo, const uint32_t* chi) {
  if (field == 0) warp_finish2<Fp377>(out, t, clo, chi);
  else if (field == 1) warp_finish2<Fr377>(out, t, clo, chi);
  else if (field == 2) warp_finish2<FpPallas>(out, t, clo, chi);
  else warp_finish2<Fp381>(out, t, clo, chi);
}
void emu_warp_finish(int field, uint32_t* out, const uint32_t* t, const uint64_t* c) {
  if (field == 0) warp_finish<Fp377>(out, t, c);
  else if (field == 1) warp_finish<Fr377>(out, t, c);
  else if (field == 2) warp_finish<FpPallas>(out, t, c);
  else warp_finish<Fp381>(out, t, c);
}
void emu_warp_mul(int field, uint32_t* out, const uint32_t* a, const uint32_t* b, int n) {
  if (field == 0) warp_mul<Fp377>(out, a, b, n);
  else if (field == 1) warp_mul<Fr377>(out, a, b, n);
  else if (field == 2) warp_mul<FpPallas>(out, a, b, n);
  else warp_mul<Fp381>(out, a, b, n);
}
void emu_fe_op(int field, int op, uint32_t* out, const uint32_t* a, const uint32_t* b) {
  if (field == 0) fe_op<Fp377>(op, out, a, b);
  else if (field == 1) fe_op<Fr377>(op, out, a, b);
  else if (field == 2) fe_op<FpPallas>(op, out, a, b);
  else fe_op<Fp381>(op, out, a, b);
}
void emu_w_op(int curve, int op, uint32_t* out, const uint32_t* a, const uint32_t* b) {
  if (curve == 0) w_op<Fp377>(op, out, a, b);
  else if (curve == 1) w_op<FpPallas>(op, out, a, b);
  else w_op<Fp381>(op, out, a, b);
}
void emu_te_op(int op, uint32_t* out, const uint32_t* a, const uint32_t* b) { te_op<Fr377, Ed377Consts>(op, out, a, b); }
}
