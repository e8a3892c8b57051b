// Warp-cooperative multi-limb field arithmetic: ONE field element spread over the lanes of a
// 16-lane group, one limb per lane -- the Montgomery product (used by the Horner kernels) and a
// lane-parallel division-step inverse (experiment, see WarpField::inv).
//
// Field::mul (field.cuh) keeps a whole element in one thread: 277 carry-dependent IMAD.WIDE for a
// 377-bit product, ~1.2 us when the warp runs alone.  That is the right shape for the throughput
// kernels (32 independent products per warp instruction) and the wrong one for the serial tail of
// the MSM -- the Horner chain over the windows (~110 dependent doublings, reference
// msm-batched-affine.ts:322-334) and the assembly of a window sum -- where one product at a time
// is in flight and 31 lanes of the warp idle.  Here the N limbs of a product are computed by N
// lanes at once (the reference's only step in this direction is the two-lane Wasm-SIMD experiment
// src/wasm/experiments/multiply-schoolbook-simd.ts; here the lanes can exchange data):
//
//   lane i holds a_i, b_i, p_i and one limb t_i of the running sum plus a pending carry c_i that
//   belongs one limb higher.  Step k (k = 0..N-1, interleaved reduction, CIOS):
//       s_i  = t_i + lo(a_i * b_k) + c_i
//       m    = s_0 * (-1/p mod 2^32)                       (lane 0, broadcast)
//       s_i += lo(m * p_i)                                  (now s_0 = 0 mod 2^32)
//       c_i  = hi(a_i * b_k) + hi(m * p_i) + (s_i >> 32)    (weight of limb i + 1)
//       t_i  = low word of s_(i+1)                          (shift down one lane = divide by 2^32:
//                                                            c_i now has the weight of t_i)
//   The carries are never rippled inside the loop (c_i < 2^33 + 8, kept in 64 bits).  After N
//   steps the value is sum (t_i + c_i) 2^(32 i) < 2p; one shuffle moves the high parts up, the
//   remaining single-bit carries are resolved by a carry-lookahead over two ballots (generate /
//   propagate masks added as integers), and the conditional subtraction of p uses the same trick
//   for the borrows.  Dependent path per step: 2 IMAD.WIDE + 1 IMAD + 2 shuffles; 12 steps instead
//   of 277 dependent MADs.
//
// Two independent products fit in one warp (lanes 0-15 and 16-31).  Result is the canonical
// representative in [0, p) -- bit-identical to Field::mul.
#pragma once
#include "field.cuh"

namespace mgb {
namespace warp {
#ifdef MGB_HOST_EMU
// host emulation (tests only): tests/host_emu/simt_emu.h, included first, runs 32 host threads in
// lockstep and defines lane / shfl / shfl_up / shfl_down / ballot in this namespace
#ifndef MGB_SIMT_EMU
#error "MGB_HOST_EMU: include tests/host_emu/simt_emu.h before warp.cuh"
#endif
#else
MGB_DEV int lane() { return (int)(threadIdx.x & 31u); }
MGB_DEV uint32_t shfl(uint32_t v, int src, int width) { return __shfl_sync(0xffffffffu, v, src, width); }
MGB_DEV uint32_t shfl_up(uint32_t v, int delta, int width) { return __shfl_up_sync(0xffffffffu, v, (unsigned)delta, width); }
MGB_DEV uint32_t shfl_down(uint32_t v, int delta, int width) { return __shfl_down_sync(0xffffffffu, v, (unsigned)delta, width); }
MGB_DEV uint32_t ballot(bool pred) { return __ballot_sync(0xffffffffu, pred); }
#endif
}  // namespace warp

template <class P>
struct WarpField {
  static constexpr int N = P::N;
  static constexpr int W = 16;          // lanes per element (one limb per lane, lanes N..W-1 carry zeros)
  static_assert(N < W, "one spare lane above the top limb");

  // limb l of the modulus for a run-time lane index (0 for l >= N)
  MGB_DEV static uint32_t mod_limb(int l) {
    uint32_t r = 0;
    _Pragma("unroll") for (int k = 0; k < N; k++) r = (l == k) ? P::mod(k) : r;
    return r;
  }

  // carry-lookahead over a 16-lane group: gen / prop are this group's ballot bits (bit i = lane i);
  // returns the mask of carries INTO each lane: c_0 = 0, c_(i+1) = gen_i | (prop_i & c_i).
  // (Adding A = gen | prop and B = gen as integers performs exactly that recurrence; the carry into
  // bit i of a sum is (A + B) ^ A ^ B.)  gen and prop never hold the same bit.
  MGB_DEV static uint32_t lookahead(uint32_t gen, uint32_t prop) {
    const uint32_t A = gen | prop, B = gen;
    return (A + B) ^ A ^ B;
  }

  // All 32 lanes call.  Lane l (within its 16-lane group) passes limb l of a and b (ignored for
  // l >= N) and receives limb l of a*b/R mod p (0 for l >= N).  a, b < p.
  MGB_DEV static uint32_t mul(uint32_t a, uint32_t b) {
    const int ln = warp::lane();
    const int l = ln & (W - 1);
    if (l >= N) a = 0;
    const uint32_t p = mod_limb(l);
    const uint32_t minv = c_mgb_minv[P::ID];
    uint32_t bk[N];
    _Pragma("unroll") for (int k = 0; k < N; k++) bk[k] = warp::shfl(b, k, W);
    uint32_t t = 0;
    uint64_t c = 0;
    _Pragma("unroll") for (int k = 0; k < N; k++) {
      const uint64_t p1 = (uint64_t)a * bk[k];
      uint64_t s = (uint64_t)t + (uint32_t)p1 + c;
      const uint32_t m = warp::shfl((uint32_t)s * minv, 0, W);
      const uint64_t p2 = (uint64_t)m * p;
      s += (uint32_t)p2;
      c = (p1 >> 32) + (p2 >> 32) + (s >> 32);
      t = warp::shfl_down((uint32_t)s, 1, W);   // lane W-1 keeps its own value, which is 0
    }
    return finish(t, c);
  }

  // sum (t_i + c_i) 2^(32 i) < 2p, c_i < 2^34, spread over the lanes as in mul -> canonical limbs.
  // One shuffle moves the high parts one lane up; what is left are single-bit carries.
  MGB_DEV static uint32_t finish(uint32_t t, uint64_t c) {
    const int ln = warp::lane();
    const int l = ln & (W - 1);
    const int gsh = ln & ~(W - 1);
    const uint32_t p = mod_limb(l);
    const uint64_t s = (uint64_t)t + c;
    uint32_t up = warp::shfl_up((uint32_t)(s >> 32), 1, W);
    if (l == 0) up = 0;
    const uint64_t s2 = (uint64_t)(uint32_t)s + up;
    uint32_t lo = (uint32_t)s2;
    const uint32_t gen = (warp::ballot((s2 >> 32) != 0) >> gsh) & 0xffffu;
    const uint32_t prop = (warp::ballot(lo == 0xffffffffu) >> gsh) & 0xffffu;
    lo += (lookahead(gen, prop) >> l) & 1u;
    // canonical: subtract p unless that borrows out of the top limb
    const uint32_t lt = (warp::ballot(lo < p) >> gsh) & 0xffffu;
    const uint32_t eq = (warp::ballot(lo == p) >> gsh) & 0xffffu;
    const uint32_t bw = lookahead(lt, eq);
    const uint32_t d = lo - p - ((bw >> l) & 1u);
    return ((bw >> N) & 1u) ? lo : d;
  }

  // ---------------------------------------------------------------------------------------------
  // Lane-parallel division-step inverse (experiment for the next round; Field::inv_divsteps is the
  // one in use).  Field::inv_divsteps runs on ONE lane: per batch of 30 division steps it finds the
  // 2x2 transition matrix from the low words (inherently serial) and then applies it to the four
  // L-limb numbers f, g, d, e -- two thirds of its instructions, and 10 % of all instructions the
  // batched-addition kernel issues, with 1 of 32 lanes active.  Here the matrix is found by every
  // lane redundantly (uniform control flow), and the application is spread over the warp: lane k of
  // the first half-warp holds limb k of (f, g), lane 16 + k limb k of (d, e).  Limbs are signed and
  // LAZY: after u*x + v*y (+ p_k*m) each lane keeps the low 30 bits, hands them one lane down
  // (= division by 2^30) and the high part stays; a second one-lane-up hop of the small overflow
  // leaves every limb in [-6, 2^30 + 5] (the top limb carries the sign).  Values are exact, only
  // their representation is not canonical -- so "g == 0" is tested on the low words first (exact
  // mod 2^32) and, when those vanish, by one sequential carry walk over the lanes.
  // All 32 lanes call with the SAME a (Montgomery form); every lane returns a^-1.  a = 0 -> 0.
  MGB_DEV static Fe<P> inv(const Fe<P>& a) {
    typedef Field<P> F;
    constexpr int L = P::N30;
    constexpr int32_t M30 = 0x3fffffff;
    static_assert(L < W, "one spare lane above the top limb");
    if (F::is_zero(a)) return a;
    const int ln = warp::lane();
    const int k = ln & (W - 1);
    const bool de = (ln & W) != 0;                 // second half-warp: (d, e); first: (f, g)
    int32_t x = 0, y = 0, pk = 0;                  // this lane's limb of (f | d) and (g | e); modulus limb for the d, e lanes
    _Pragma("unroll") for (int i = 0; i < L; i++) {
      const int bit = 30 * i, w = bit >> 5, sh = bit & 31;
      const uint32_t lo = a.v[w], hi = (w + 1 < N) ? a.v[w + 1] : 0u;
      const uint32_t gi = (sh ? ((lo >> sh) | (sh > 2 ? (hi << (32 - sh)) : 0u)) : lo) & (uint32_t)M30;
      if (k == i) {
        x = de ? 0 : P::mod30(i);
        y = de ? (i == 0 ? 1 : 0) : (int32_t)gi;
        pk = de ? P::mod30(i) : 0;
      }
    }
    int32_t eta = -1;
    for (int iter = 0; iter < 64; iter++) {
      // low words of f, g (exact mod 2^32 whatever the representation), low and top limbs of d, e
      const uint32_t f0 = warp::shfl((uint32_t)x, 0, 32) + (warp::shfl((uint32_t)x, 1, 32) << 30);
      const uint32_t g0 = warp::shfl((uint32_t)y, 0, 32) + (warp::shfl((uint32_t)y, 1, 32) << 30);
      if (g0 == 0) {                               // g == 0?  one exact carry walk over the limbs (uniform branch)
        int32_t carry = 0, nz = 0;
        for (int i = 0; i < L; i++) {
          const int32_t v = (int32_t)warp::shfl((uint32_t)y, i, 32) + carry;
          nz |= (i < L - 1) ? (v & M30) : v;
          carry = v >> 30;
        }
        if (nz == 0) break;
      }
      uint32_t ff = f0, gg = g0;
      int32_t u = 1, v = 0, q = 0, r = 1;
      int i = 30;
      auto neg_inv = [](uint32_t fo) -> uint32_t { uint32_t t = fo; t *= 2u - fo * t; t *= 2u - fo * t; return 0u - t; };
      uint32_t ninv = neg_inv(ff);
      while (true) {                               // 30 division steps on the low words (as Field::inv_divsteps)
        const uint32_t lim = gg | (0xffffffffu << i);
        const int zeros = MGB_CTZ(lim);
        gg >>= zeros; u <<= zeros; v <<= zeros; eta -= zeros; i -= zeros;
        if (i == 0) break;
        if (eta < 0) {
          eta = -eta;
          const uint32_t tf = ff; ff = gg; gg = 0u - tf;
          const int32_t tu = u; u = q; q = -tu;
          const int32_t tv = v; v = r; r = -tv;
          ninv = neg_inv(ff);
        }
        const int limit = (eta + 1 < i) ? eta + 1 : i;
        const uint32_t m = (0xffffffffu >> (32 - limit)) & 255u;
        const uint32_t w = (gg * ninv) & m;
        gg += ff * w; q += u * (int32_t)w; r += v * (int32_t)w;
      }
      // multiples of p that make the low 30 bits of the new d, e vanish (zero for the f, g lanes through pk = 0)
      const int32_t d0 = (int32_t)warp::shfl((uint32_t)x, W, 32), e0 = (int32_t)warp::shfl((uint32_t)y, W, 32);
      const int32_t dt = (int32_t)warp::shfl((uint32_t)x, W + L - 1, 32), et = (int32_t)warp::shfl((uint32_t)y, W + L - 1, 32);
      const int32_t sd = dt >> 31, se = et >> 31;
      int32_t md = (u & sd) + (v & se), me = (q & sd) + (r & se);
      const int64_t cd = (int64_t)u * d0 + (int64_t)v * e0, ce = (int64_t)q * d0 + (int64_t)r * e0;
      md -= (int32_t)((P::MINV30 * (uint32_t)cd + (uint32_t)md) & (uint32_t)M30);
      me -= (int32_t)((P::MINV30 * (uint32_t)ce + (uint32_t)me) & (uint32_t)M30);
      // this lane's limb of both numbers, then / 2^30: low parts one lane down, small overflow one lane up
      const int64_t tx = (int64_t)u * x + (int64_t)v * y + (int64_t)pk * md;
      const int64_t ty = (int64_t)q * x + (int64_t)r * y + (int64_t)pk * me;
      const uint32_t lx = warp::shfl_down((uint32_t)tx & (uint32_t)M30, 1, W);      // lane W-1 gets its own 0 back
      const uint32_t ly = warp::shfl_down((uint32_t)ty & (uint32_t)M30, 1, W);
      const int64_t wx = (tx >> 30) + (int64_t)lx, wy = (ty >> 30) + (int64_t)ly;
      const bool top = k >= L - 1;                 // the top limb is not split: it carries the sign
      int32_t cx = (int32_t)warp::shfl_up((uint32_t)(top ? 0 : (int32_t)(wx >> 30)), 1, W);
      int32_t cy = (int32_t)warp::shfl_up((uint32_t)(top ? 0 : (int32_t)(wy >> 30)), 1, W);
      if (k == 0) { cx = 0; cy = 0; }
      x = (top ? (int32_t)wx : (int32_t)((uint32_t)wx & (uint32_t)M30)) + cx;
      y = (top ? (int32_t)wy : (int32_t)((uint32_t)wy & (uint32_t)M30)) + cy;
    }
    // f = +-1 (its low word tells which); inverse = sign(f) * d, gathered to every lane and brought to [0, p)
    const bool fneg = (warp::shfl((uint32_t)x, 0, 32) + (warp::shfl((uint32_t)x, 1, 32) << 30)) != 1u;
    int32_t d[L];
    _Pragma("unroll") for (int i = 0; i < L; i++) d[i] = (int32_t)warp::shfl((uint32_t)x, W + i, 32);
    int64_t c = 0;
    _Pragma("unroll") for (int i = 0; i < L; i++) {
      c += fneg ? -(int64_t)d[i] : (int64_t)d[i];
      d[i] = (i < L - 1) ? ((int32_t)c & M30) : (int32_t)c;
      c >>= 30;
    }
    _Pragma("unroll 1") for (int rep = 0; rep < 6; rep++) {   // d in (-3p, 3p) -> [0, p)
      const bool neg = d[L - 1] < 0;
      int32_t t[L];
      int64_t cc = 0;
      _Pragma("unroll") for (int i = 0; i < L; i++) {
        cc += (int64_t)d[i] + (neg ? (int64_t)P::mod30(i) : -(int64_t)P::mod30(i));
        t[i] = (i < L - 1) ? ((int32_t)cc & M30) : (int32_t)cc;
        cc >>= 30;
      }
      if (!(neg || t[L - 1] >= 0)) break;          // negative: add p; non-negative and >= p: subtract p
      _Pragma("unroll") for (int i = 0; i < L; i++) d[i] = t[i];
    }
    Fe<P> out;
    _Pragma("unroll") for (int w = 0; w < N; w++) {
      const int bit = 32 * w, j = bit / 30, sh = bit - 30 * j;
      uint32_t val = (uint32_t)d[j] >> sh;
      if (j + 1 < L) val |= (uint32_t)d[j + 1] << (30 - sh);
      if (sh > 28 && j + 2 < L) val |= (uint32_t)d[j + 2] << (60 - sh);
      out.v[w] = val;
    }
    Fe<P> r3;
    _Pragma("unroll") for (int i = 0; i < N; i++) r3.v[i] = P::r3(i);
    return F::mul(out, r3);                        // (aR)^-1 * R^3 / R = a^-1 R
  }
  // out-of-line copy for kernels that are short of registers and instruction cache
  MGB_NOINLINE_DEV static Fe<P> inv_call(Fe<P> a) { return inv(a); }
};

// ---------------------------------------------------------------------------------------------
// The same product with TWO limbs (one 64-bit digit) per lane: an element occupies an 8-lane group,
// so FOUR independent products fit in one warp -- all the products of a formula level of a point
// doubling / addition, which would let the Horner chain run inside one warp without block barriers
// (experiment for the next round; the Horner kernels use WarpField).  Half as many steps (N/2,
// each with its two shuffles on the dependent path), more arithmetic per step: 64x64->128 products
// and a pending carry that needs 66 bits (64-bit word + a small count).
template <class P>
struct WarpField2 {
  typedef unsigned long long u64;
  static constexpr int N = P::N, D = P::N / 2;     // limbs, 64-bit digits
  static constexpr int W = 8;                      // lanes per element (lanes D..W-1 carry zeros)
  static_assert(D < W, "one spare lane above the top digit");

  MGB_DEV static u64 mod_digit(int l) {
    u64 r = 0;
    _Pragma("unroll") for (int k = 0; k < D; k++) r = (l == k) ? (((u64)P::mod(2 * k + 1) << 32) | P::mod(2 * k)) : r;
    return r;
  }
  MGB_DEV static u64 shfl64(u64 v, int src) {
    return ((u64)warp::shfl((uint32_t)(v >> 32), src, W) << 32) | warp::shfl((uint32_t)v, src, W);
  }
  MGB_DEV static void mul128(u64 a, u64 b, u64& lo, u64& hi) {
#ifdef MGB_HOST_EMU
    const unsigned __int128 t = (unsigned __int128)a * b;
    lo = (u64)t; hi = (u64)(t >> 64);
#else
    lo = a * b; hi = __umul64hi(a, b);
#endif
  }
  MGB_DEV static uint32_t lookahead(uint32_t gen, uint32_t prop) {   // as WarpField::lookahead
    const uint32_t A = gen | prop, B = gen;
    return (A + B) ^ A ^ B;
  }

  // All 32 lanes call.  Lane l of an 8-lane group passes digit l of a and b (ignored for l >= D) and receives
  // digit l of a*b/R mod p, canonical (0 for l >= D).  a, b < p.
  MGB_DEV static u64 mul(u64 a, u64 b) {
    const int ln = warp::lane();
    const int l = ln & (W - 1);
    const int gsh = ln & ~(W - 1);
    if (l >= D) a = 0;
    const u64 p = mod_digit(l);
    const u64 p0 = ((u64)P::mod(1) << 32) | P::mod(0);
    u64 y = (u64)(0u - c_mgb_minv[P::ID]);           // p^-1 mod 2^32 -> one Newton step -> mod 2^64
    y *= 2ull - p0 * y;
    const u64 minv = 0ull - y;
    u64 bk[D];
    _Pragma("unroll") for (int k = 0; k < D; k++) bk[k] = shfl64(b, k);
    u64 t = 0, clo = 0;
    uint32_t chi = 0;                                // pending carry clo + chi 2^64, one digit above t
    _Pragma("unroll") for (int k = 0; k < D; k++) {
      u64 p1l, p1h, p2l, p2h;
      mul128(a, bk[k], p1l, p1h);
      uint32_t sc = 0, ch = 0;                       // carries out of the low / the high word sums (explicit carry chains:
      const u64 s = ptx::add64_count(t, p1l, sc);    //  a third fewer instructions in this latency-bound routine)
      const u64 s1 = ptx::add64_count(s, clo, sc);
      const u64 m = shfl64(s1 * minv, 0);
      mul128(m, p, p2l, p2h);
      const u64 s2 = ptx::add64_count(s1, p2l, sc);  // digit 0: 0 by the choice of m
      const u64 c = ptx::add64_count(p1h, p2h, ch);
      const u64 c2 = ptx::add64_count(c, (u64)(sc + chi), ch);
      clo = c2; chi = ch;
      const uint32_t dl = warp::shfl_down((uint32_t)s2, 1, W), dh = warp::shfl_down((uint32_t)(s2 >> 32), 1, W);
      t = ((u64)dh << 32) | dl;                      // lane W-1 keeps its own value, which is 0
    }
    return finish(t, clo, chi);
  }

  // sum (t_i + clo_i + chi_i 2^64) 2^(64 i) < 2p spread over the lanes as in mul -> canonical digits:
  // high parts one lane up, then single-bit carries by lookahead, then the conditional subtraction of p
  MGB_DEV static u64 finish(u64 t, u64 clo, uint32_t chi) {
    const int ln = warp::lane();
    const int l = ln & (W - 1);
    const int gsh = ln & ~(W - 1);
    const u64 p = mod_digit(l);
    const u64 s = t + clo;
    uint32_t up = warp::shfl_up(chi + (uint32_t)(s < t), 1, W);
    if (l == 0) up = 0;
    u64 lo = s + up;
    const uint32_t gen = (warp::ballot(lo < s) >> gsh) & 0xffu;
    const uint32_t prop = (warp::ballot(lo == ~0ull) >> gsh) & 0xffu;
    lo += (lookahead(gen, prop) >> l) & 1u;
    const uint32_t lt = (warp::ballot(lo < p) >> gsh) & 0xffu;
    const uint32_t eq = (warp::ballot(lo == p) >> gsh) & 0xffu;
    const uint32_t bw = lookahead(lt, eq);
    const u64 d = lo - p - ((bw >> l) & 1u);
    return ((bw >> D) & 1u) ? lo : d;
  }

  // ---- additions on distributed elements (all four groups of a warp at once); canonical in, canonical out
  MGB_DEV static uint32_t group_ballot(bool pred) { return (warp::ballot(pred) >> (warp::lane() & ~(W - 1))) & 0xffu; }
  // a - b mod p
  MGB_DEV static u64 sub(u64 a, u64 b) {
    const int l = warp::lane() & (W - 1);
    const u64 p = mod_digit(l);
    const uint32_t bw = lookahead(group_ballot(a < b), group_ballot(a == b));
    const u64 d = a - b - ((bw >> l) & 1u);
    // a < b (borrow out of the top digit): add p back; its carry out of the top digit cancels the borrow.  No branch:
    // the four groups of a warp decide differently and the ballots below need all 32 lanes.
    const u64 s = d + p;
    const u64 r = s + ((lookahead(group_ballot(s < d), group_ballot(s == ~0ull)) >> l) & 1u);
    return ((bw >> D) & 1u) ? r : d;
  }
  // a + b mod p
  MGB_DEV static u64 add(u64 a, u64 b) {
    const int l = warp::lane() & (W - 1);
    const u64 p = mod_digit(l);
    u64 s = a + b;                                   // 2p < 2^(64 D): no carry out of the top digit
    s += (lookahead(group_ballot(s < a), group_ballot(s == ~0ull)) >> l) & 1u;
    const uint32_t bw = lookahead(group_ballot(s < p), group_ballot(s == p));
    const u64 d = s - p - ((bw >> l) & 1u);
    return ((bw >> D) & 1u) ? s : d;
  }
  MGB_DEV static u64 dbl(u64 a) { return add(a, a); }
  // is the element of this lane's group zero?  (same answer on the 8 lanes of a group)
  MGB_DEV static bool is_zero(u64 a) { return group_ballot(a != 0) == 0; }
};

}  // namespace mgb
