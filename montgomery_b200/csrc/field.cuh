// Prime-field arithmetic on 32-bit limbs in Montgomery form, fully reduced to [0, p).
//
// Replaces the reference's runtime-generated Wasm field module (29-bit limbs in i64 locals):
//   multiply / square      src/wasm/multiply-montgomery.ts:58-215
//   add / subtract / ...   src/wasm/field-arithmetic.ts:32-176
//   inverse                src/wasm/inverse.ts:191-218 (here: Fermat, a^(p-2))
//   to/from Montgomery     src/field-msm.ts:179-185
// Design differs on purpose: 32-bit limbs, CIOS with two accumulators ("even"/"odd" limb
// alignment) so that every 32x32->64 product lands on an aligned register pair and ptxas emits
// IMAD.WIDE.U32(.X) carry chains; values are kept canonical (< p) so equality is a limb compare.
// All three base primes satisfy p = 1 (mod 2^32): the Montgomery quotient digit is m = -t0 and
// the p[0]*m product is replaced by a carry (same shortcut as multiply-montgomery.ts:324-334).
#pragma once
#include <type_traits>
#include "ptx.cuh"
#include "constants_gen.cuh"

namespace mgb {

// -p^-1 mod 2^32 as a RUNTIME value, one entry per field (P::ID).  For the primes with p = 1 mod 2^32
// the factor is 2^32 - 1, and if ptxas can see that the Montgomery quotient digit is a plain
// negation it folds the sign into the modulus immediates and splits every modulus product into
// IMAD.X + IMAD.HI.U32.X (6 issue cycles on the multiplier pipe) instead of one IMAD.WIDE.U32.X
// (4 cycles) -- measured +24% on the whole multiplication.
// c_mgb_zero: a zero in constant memory, i.e. one the compiler must treat as unknown.  "x + 0 + carry" with a literal
// zero becomes IMAD.X (multiplier pipe), with this zero as the addend IADD3.X (ALU pipe); the multiplier pipe is what
// bounds the accumulation kernel (ncu: 88 % busy), the ALU pipe is a quarter busy.  See mul_impl.
#ifdef MGB_HOST_EMU
static const uint32_t c_mgb_minv[4] = MGB_MINV_TABLE;
static const uint32_t c_mgb_zero = 0;
#else
static __device__ __constant__ uint32_t c_mgb_minv[4] = MGB_MINV_TABLE;
static __device__ __constant__ uint32_t c_mgb_zero = 0;
#endif

template <class P>
struct Fe {
  uint32_t v[P::N];
};

template <class P>
struct Field {
  static constexpr int N = P::N;
  typedef Fe<P> fe;
  static_assert(N % 2 == 0, "even limb count");

  MGB_DEV static fe zero() { fe r; _Pragma("unroll") for (int i = 0; i < N; i++) r.v[i] = 0; return r; }
  MGB_DEV static fe one() { fe r; _Pragma("unroll") for (int i = 0; i < N; i++) r.v[i] = P::one(i); return r; }
  MGB_DEV static fe modulus() { fe r; _Pragma("unroll") for (int i = 0; i < N; i++) r.v[i] = P::mod(i); return r; }

  MGB_DEV static bool is_zero(const fe& a) {
    uint32_t o = 0;
    _Pragma("unroll") for (int i = 0; i < N; i++) o |= a.v[i];
    return o == 0;
  }
  MGB_DEV static bool eq(const fe& a, const fe& b) {
    uint32_t o = 0;
    _Pragma("unroll") for (int i = 0; i < N; i++) o |= a.v[i] ^ b.v[i];
    return o == 0;
  }
  MGB_DEV static fe select(bool c, const fe& a, const fe& b) {  // c ? a : b
    fe r;
    _Pragma("unroll") for (int i = 0; i < N; i++) r.v[i] = c ? a.v[i] : b.v[i];
    return r;
  }

  // t (N limbs, < 2p) -> canonical
  MGB_DEV static fe reduce_once(const uint32_t* t) {
    uint32_t u[N];
    u[0] = ptx::sub_cc(t[0], P::mod(0));
    _Pragma("unroll") for (int i = 1; i < N; i++) u[i] = ptx::subc_cc(t[i], P::mod(i));
    uint32_t borrow = ptx::subc(0, 0);
    fe r;
    _Pragma("unroll") for (int i = 0; i < N; i++) r.v[i] = borrow ? t[i] : u[i];
    return r;
  }

  MGB_DEV static fe add(const fe& a, const fe& b) {
    uint32_t t[N];
    t[0] = ptx::add_cc(a.v[0], b.v[0]);
    _Pragma("unroll") for (int i = 1; i < N - 1; i++) t[i] = ptx::addc_cc(a.v[i], b.v[i]);
    t[N - 1] = ptx::addc(a.v[N - 1], b.v[N - 1]);  // 2p < 2^(32N): no carry out
    return reduce_once(t);
  }
  MGB_DEV static fe dbl(const fe& a) { return add(a, a); }

  MGB_DEV static fe sub(const fe& a, const fe& b) {
    uint32_t t[N];
    t[0] = ptx::sub_cc(a.v[0], b.v[0]);
    _Pragma("unroll") for (int i = 1; i < N; i++) t[i] = ptx::subc_cc(a.v[i], b.v[i]);
    uint32_t borrow = ptx::subc(0, 0);  // 0xffffffff if a < b
    fe r;
    r.v[0] = ptx::add_cc(t[0], P::mod(0) & borrow);
    _Pragma("unroll") for (int i = 1; i < N - 1; i++) r.v[i] = ptx::addc_cc(t[i], P::mod(i) & borrow);
    r.v[N - 1] = ptx::addc(t[N - 1], P::mod(N - 1) & borrow);
    return r;
  }
  MGB_DEV static fe neg(const fe& a) { return sub(zero(), a); }

  // Montgomery product a*b/R mod p, a, b < p.
  // X/Y alternate as E (pairs on limbs 0,1|2,3|...) and O (pairs on limbs 1,2|3,4|...): t = E + O*2^32.
  // SQR = true: fb is ignored, the result is fa^2.  Row i then only forms the products with j >= i -- the
  // diagonal a_i*a_i and a_i*(2 a_{>i})_j for j > i (2a < 2^(32 N) for every modulus here) -- so 66 of the 144
  // operand products of a 12-limb squaring disappear; the accumulator limbs below the first product
  // of a row are only shifted, with the carry rippling through plain add-with-carry (ALU pipe).
  // compile-time loop: f(std::integral_constant<int, I>) for I = Begin, Begin + Step, .. < End
  template <int I, int End, int Step, class Fn>
  MGB_DEV static void static_for(Fn&& f) {
    if constexpr (I < End) {
      f(std::integral_constant<int, I>{});
      static_for<I + Step, End, Step>(f);
    }
  }
  // compile-time classification of a modulus limb (see mul_impl)
  // zero or a single bit.  (Two-bit limbs such as BLS12-377's p_2 = 3 << 28 were tried: ptxas turns the second
  // shifted add back into IMAD.WIDE x, 2^k, so nothing is saved.)
  MGB_DEV static constexpr bool cheap_limb(uint32_t c) { return (c & (c - 1)) == 0; }
  MGB_DEV static constexpr int low_bit(uint32_t c) { int b = 0; while (b < 31 && !((c >> b) & 1)) b++; return b; }
  MGB_DEV static constexpr int high_bit(uint32_t c) { int b = 31; while (b > 0 && !((c >> b) & 1)) b--; return b; }
  template <bool SQR>
  MGB_DEV static fe mul_impl(const fe& fa, const fe& fb) {
    const uint32_t* a = fa.v;
    const uint32_t* b = SQR ? fa.v : fb.v;
    uint32_t a2[N];
    if constexpr (SQR) {
      a2[0] = a[0] << 1;
      _Pragma("unroll") for (int j = 1; j < N; j++) a2[j] = (a[j] << 1) | (a[j - 1] >> 31);
    }
    // A zero the compiler cannot see through (c_mgb_zero).  The two carry absorbers of every row
    // are "x + 0 + carry": with a literal 0 ptxas emits IMAD.X (multiplier pipe, the bottleneck), with an opaque
    // register addend it emits IADD3.X (ALU pipe) -- 22 multiplier-pipe instructions fewer per product.
    const uint32_t zr = c_mgb_zero;
    uint32_t X[N], Y[N];
    _Pragma("unroll") for (int i = 0; i < N; i++) {
      uint32_t* E = (i & 1) ? Y : X;
      uint32_t* O = (i & 1) ? X : Y;
      const uint32_t bi = b[i];
      // multiplicand limb j of row i, and whether the product exists at all
      // (the limb right above the diagonal must not take the top bit of a_i: that bit belongs to 2*a_i, not to 2*(a >> 32(i+1)))
      auto mc = [&](int j) -> uint32_t { return SQR ? (j > i + 1 ? a2[j] : (j == i + 1 ? a[j] << 1 : a[j])) : a[j]; };
      auto has = [&](int j) -> bool { return !SQR || j >= i; };
      if (i == 0) {
        _Pragma("unroll") for (int j = 0; j < N; j += 2) { E[j] = ptx::mul_lo(mc(j), bi); E[j + 1] = ptx::mul_hi(mc(j), bi); }
        _Pragma("unroll") for (int j = 0; j < N; j += 2) { O[j] = ptx::mul_lo(mc(j + 1), bi); O[j + 1] = ptx::mul_hi(mc(j + 1), bi); }
      } else {
        // O is last round's E: its limb 0 was cancelled, its limb 1 now has the weight of E[0];
        // the carry of that add has the weight of O's (shifted) pair 0 and enters the chain.
        E[0] = ptx::add_cc(E[0], O[1]);
        _Pragma("unroll") for (int j = 0; j < N - 2; j += 2) {
          if (has(j + 1)) {
            O[j] = ptx::madc_lo_cc(mc(j + 1), bi, O[j + 2]);
            O[j + 1] = ptx::madc_hi_cc(mc(j + 1), bi, O[j + 3]);
          } else {
            O[j] = ptx::addc_cc(O[j + 2], 0);
            O[j + 1] = ptx::addc_cc(O[j + 3], 0);
          }
        }
        O[N - 2] = ptx::madc_lo_cc(mc(N - 1), bi, 0);
        O[N - 1] = ptx::madc_hi(mc(N - 1), bi, 0);
        bool first = true;
        _Pragma("unroll") for (int j = 0; j < N; j += 2) {
          if (!has(j)) continue;
          E[j] = first ? ptx::mad_lo_cc(mc(j), bi, E[j]) : ptx::madc_lo_cc(mc(j), bi, E[j]);
          E[j + 1] = ptx::madc_hi_cc(mc(j), bi, E[j + 1]);
          first = false;
        }
        if (first) O[N - 1] = ptx::add_cc(O[N - 1], 0);   // (cannot happen: row N-1 still has a product with j >= i)
        O[N - 1] = ptx::addc(O[N - 1], zr);
      }
      // Quotient digit m = -t0 / p mod 2^32.  For the moduli with p = 1 mod 2^32 that is a negation: formed as
      // zr - t0 with the opaque zero (ptxas must not learn that m = -t0: see the note at c_mgb_minv), it is one IADD3 on the
      // ALU pipe instead of an IMAD on the multiplier pipe (12 of a product's 320 multiplier-pipe instructions).
      uint32_t m;
      if constexpr (P::mod(0) == 1u) m = zr - E[0];
      else m = E[0] * c_mgb_minv[P::ID];
      // Modulus limbs that are zero or a power of two (Pallas: p_4..p_6 = 0, p_7 = 1 << 30) do not go through the
      // multiplier: the 64-bit product m * p_j is two shifts on the ALU pipe and enters the carry chain as a plain addend.
      uint32_t clo[N], chi[N];
      static_for<1, N, 1>([&](auto J) {
        constexpr int j = decltype(J)::value;
        constexpr uint32_t c = P::mod(j);
        clo[j] = 0; chi[j] = 0;
        if constexpr (cheap_limb(c) && c != 0) {
          constexpr int b0 = low_bit(c), b1 = high_bit(c);
          clo[j] = m << b0;
          if constexpr (b0 != 0) chi[j] = m >> (32 - b0);
          if constexpr (b1 != b0) {
            clo[j] = ptx::add_cc(clo[j], m << b1);
            chi[j] = ptx::addc(chi[j], m >> (32 - b1));
          }
        }
      });
      static_for<0, N, 2>([&](auto J) {
        constexpr int j = decltype(J)::value;
        if constexpr (cheap_limb(P::mod(j + 1))) {
          O[j] = (j == 0) ? ptx::add_cc(O[j], clo[j + 1]) : ptx::addc_cc(O[j], clo[j + 1]);
          O[j + 1] = ptx::addc_cc(O[j + 1], chi[j + 1]);
        } else {
          O[j] = (j == 0) ? ptx::mad_lo_cc(P::mod(j + 1), m, O[j]) : ptx::madc_lo_cc(P::mod(j + 1), m, O[j]);
          O[j + 1] = ptx::madc_hi_cc(P::mod(j + 1), m, O[j + 1]);
        }
      });
      if constexpr (P::mod(0) == 1u) {
        // E pair 0 += p[0]*m with p[0] = 1: limb 0 becomes 0, carry (E[0] != 0) goes into limb 1
        (void)ptx::add_cc(E[0], 0xffffffffu);
        E[1] = ptx::addc_cc(E[1], 0);
      } else {
        E[0] = ptx::mad_lo_cc(P::mod(0), m, E[0]);      // becomes 0 by the choice of m
        E[1] = ptx::madc_hi_cc(P::mod(0), m, E[1]);
      }
      static_for<2, N, 2>([&](auto J) {
        constexpr int j = decltype(J)::value;
        if constexpr (cheap_limb(P::mod(j))) {
          E[j] = ptx::addc_cc(E[j], clo[j]);
          E[j + 1] = ptx::addc_cc(E[j + 1], chi[j]);
        } else {
          E[j] = ptx::madc_lo_cc(P::mod(j), m, E[j]);
          E[j + 1] = ptx::madc_hi_cc(P::mod(j), m, E[j + 1]);
        }
      });
      O[N - 1] = ptx::addc(O[N - 1], zr);
    }
    uint32_t* E = ((N - 1) & 1) ? Y : X;
    uint32_t* O = ((N - 1) & 1) ? X : Y;
    uint32_t t[N];
    t[0] = ptx::add_cc(O[0], E[1]);
    _Pragma("unroll") for (int j = 1; j < N - 1; j++) t[j] = ptx::addc_cc(O[j], E[j + 1]);
    t[N - 1] = ptx::addc(O[N - 1], 0);
    return reduce_once(t);
  }
  MGB_DEV static fe mul_inl(const fe& fa, const fe& fb) { return mul_impl<false>(fa, fb); }
  // Out-of-line copy: one body per kernel keeps the instruction footprint inside the SM's
  // instruction cache (an inlined multiplication is ~5 KB of SASS).
  MGB_NOINLINE_DEV static fe mul(fe a, fe b) { return mul_inl(a, b); }  // by value: register ABI, no stack traffic
  MGB_NOINLINE_DEV static fe sqr(fe a) { return mul_impl<true>(a, a); }

  MGB_DEV static fe to_mont(const fe& a) { fe r2; _Pragma("unroll") for (int i = 0; i < N; i++) r2.v[i] = P::r2(i); return mul(a, r2); }
  MGB_DEV static fe from_mont(const fe& a) { fe o = zero(); o.v[0] = 1; return mul(a, o); }

  // Montgomery inverse a -> a^-1 (both in Montgomery form) by the plain binary extended Euclid on
  // the stored integer aR (shifts and adds only, ALU pipe), then one multiplication by R^3.
  // Variable time, meant to run on one lane per batch (reference: Kaliski almost-inverse plus
  // corrections, src/wasm/inverse.ts:136-218).  a = 0 -> 0.
  MGB_NOINLINE_DEV static fe inv_bgcd(fe a) {
    if (is_zero(a)) return a;
    uint32_t u[N], v[N], x1[N], x2[N];
    _Pragma("unroll") for (int i = 0; i < N; i++) { u[i] = a.v[i]; v[i] = P::mod(i); x1[i] = 0; x2[i] = 0; }
    x1[0] = 1;
    while (true) {
      uint32_t ou = u[0] ^ 1, ov = v[0] ^ 1;
      _Pragma("unroll") for (int i = 1; i < N; i++) { ou |= u[i]; ov |= v[i]; }
      if (ou == 0 || ov == 0) {
        fe x;
        _Pragma("unroll") for (int i = 0; i < N; i++) x.v[i] = (ou == 0) ? x1[i] : x2[i];
        fe r3; _Pragma("unroll") for (int i = 0; i < N; i++) r3.v[i] = P::r3(i);
        return mul(x, r3);
      }
      if ((u[0] & 1) == 0) { shr1(u); halve(x1); }
      else if ((v[0] & 1) == 0) { shr1(v); halve(x2); }
      else {
        // both odd: subtract the smaller from the larger
        uint32_t t[N];
        t[0] = ptx::sub_cc(u[0], v[0]);
        _Pragma("unroll") for (int i = 1; i < N; i++) t[i] = ptx::subc_cc(u[i], v[i]);
        uint32_t borrow = ptx::subc(0, 0);
        if (borrow == 0) {  // u >= v
          _Pragma("unroll") for (int i = 0; i < N; i++) u[i] = t[i];
          submod(x1, x2);
        } else {
          v[0] = ptx::sub_cc(v[0], u[0]);
          _Pragma("unroll") for (int i = 1; i < N - 1; i++) v[i] = ptx::subc_cc(v[i], u[i]);
          v[N - 1] = ptx::subc(v[N - 1], u[N - 1]);
          submod(x2, x1);
        }
      }
    }
  }
  MGB_DEV static void shr1(uint32_t* x) {
    _Pragma("unroll") for (int i = 0; i < N - 1; i++) x[i] = (x[i] >> 1) | (x[i + 1] << 31);
    x[N - 1] >>= 1;
  }
  MGB_DEV static void halve(uint32_t* x) {  // x/2 mod p for x < p
    uint32_t msk = 0u - (x[0] & 1);
    x[0] = ptx::add_cc(x[0], P::mod(0) & msk);
    _Pragma("unroll") for (int i = 1; i < N - 1; i++) x[i] = ptx::addc_cc(x[i], P::mod(i) & msk);
    x[N - 1] = ptx::addc(x[N - 1], P::mod(N - 1) & msk);
    shr1(x);
  }
  MGB_DEV static void submod(uint32_t* x, const uint32_t* y) {  // x = x - y mod p
    x[0] = ptx::sub_cc(x[0], y[0]);
    _Pragma("unroll") for (int i = 1; i < N; i++) x[i] = ptx::subc_cc(x[i], y[i]);
    uint32_t borrow = ptx::subc(0, 0);
    x[0] = ptx::add_cc(x[0], P::mod(0) & borrow);
    _Pragma("unroll") for (int i = 1; i < N - 1; i++) x[i] = ptx::addc_cc(x[i], P::mod(i) & borrow);
    x[N - 1] = ptx::addc(x[N - 1], P::mod(N - 1) & borrow);
  }

  // Montgomery inverse by Bernstein-Yang division steps in batches of 30 (the "safegcd" scheme:
  // the 2x2 transition matrix of 30 steps is found from the low words alone, then applied to the
  // full-size f, g and, modulo p, to d, e; signed 30-bit limbs make every /2^30 a limb shift).
  // ~4x fewer instructions than inv_bgcd; the matrix products run on the multiplier pipe.
  // Variable time, one lane per batch of additions.  Replaces the reference's Kaliski
  // almost-inverse (src/wasm/inverse.ts:136-218) and the hi/lo-approximation experiment of
  // src/inverse/faster-inverse.ts.  a = 0 -> 0.
  MGB_NOINLINE_DEV static fe inv_divsteps(fe a) {
    constexpr int L = P::N30;
    constexpr int32_t M30 = 0x3fffffff;
    if (is_zero(a)) return a;
    int32_t f[L], g[L], d[L], e[L];
    // packed 32-bit words -> 30-bit limbs
    _Pragma("unroll") for (int i = 0; i < L; i++) {
      const int bit = 30 * i, w = bit >> 5, sh = bit & 31;
      uint32_t lo = a.v[w], hi = (w + 1 < N) ? a.v[w + 1] : 0u;
      uint32_t x = sh ? ((lo >> sh) | (sh > 2 ? (hi << (32 - sh)) : 0u)) : lo;
      g[i] = (int32_t)(x & (uint32_t)M30);
      f[i] = P::mod30(i);
      d[i] = 0;
      e[i] = 0;
    }
    e[0] = 1;
    int32_t eta = -1;
    for (int iter = 0; iter < 64; iter++) {
      // ---- 30 division steps on the low words -> transition matrix (u v; q r), |entries| <= 2^30
      uint32_t f0 = (uint32_t)f[0] | ((uint32_t)f[1] << 30), g0 = (uint32_t)g[0] | ((uint32_t)g[1] << 30);
      int32_t u = 1, v = 0, q = 0, r = 1;
      int i = 30;
      // -(f0^-1) mod 2^12 by two Newton steps from x = f0 (f0 odd: f0*f0 = 1 mod 8)
      auto neg_inv = [](uint32_t fo) -> uint32_t { uint32_t x = fo; x *= 2u - fo * x; x *= 2u - fo * x; return 0u - x; };
      uint32_t ninv = neg_inv(f0);
      while (true) {
        uint32_t lim = g0 | (0xffffffffu << i);
        const int zeros = MGB_CTZ(lim);   // trailing zeros of g0, at most i
        g0 >>= zeros; u <<= zeros; v <<= zeros; eta -= zeros; i -= zeros;
        if (i == 0) break;
        if (eta < 0) {
          eta = -eta;
          uint32_t tf = f0; f0 = g0; g0 = 0u - tf;
          int32_t tu = u; u = q; q = -tu;
          int32_t tv = v; v = r; r = -tv;
          ninv = neg_inv(f0);
        }
        // eta >= 0: up to min(eta + 1, i, 8) low bits of g can be cancelled at once by adding the
        // multiple w of f with w = -g/f mod 2^bits (that many division steps with g odd, eta not
        // changing sign) -- several times fewer iterations than one bit at a time
        const int limit = (eta + 1 < i) ? eta + 1 : i;
        const uint32_t m = (0xffffffffu >> (32 - limit)) & 255u;
        const uint32_t w = (g0 * ninv) & m;
        g0 += f0 * w; q += u * (int32_t)w; r += v * (int32_t)w;
      }
      // ---- (f, g) <- (u f + v g, q f + r g) / 2^30   (exact)
      {
        int64_t cf = (int64_t)u * f[0] + (int64_t)v * g[0];
        int64_t cg = (int64_t)q * f[0] + (int64_t)r * g[0];
        cf >>= 30; cg >>= 30;
        _Pragma("unroll") for (int k = 1; k < L; k++) {
          int32_t fk = f[k], gk = g[k];
          cf += (int64_t)u * fk + (int64_t)v * gk;
          cg += (int64_t)q * fk + (int64_t)r * gk;
          f[k - 1] = (int32_t)cf & M30; cf >>= 30;
          g[k - 1] = (int32_t)cg & M30; cg >>= 30;
        }
        f[L - 1] = (int32_t)cf;
        g[L - 1] = (int32_t)cg;
      }
      // ---- (d, e) <- (u d + v e, q d + r e) / 2^30 mod p, kept in (-2p, p)
      {
        int32_t sd = d[L - 1] >> 31, se = e[L - 1] >> 31;
        int32_t md = (u & sd) + (v & se), me = (q & sd) + (r & se);
        int64_t cd = (int64_t)u * d[0] + (int64_t)v * e[0];
        int64_t ce = (int64_t)q * d[0] + (int64_t)r * e[0];
        // pick md, me so that the low 30 bits of cd + p*md (ce + p*me) cancel
        md -= (int32_t)((P::MINV30 * (uint32_t)cd + (uint32_t)md) & (uint32_t)M30);
        me -= (int32_t)((P::MINV30 * (uint32_t)ce + (uint32_t)me) & (uint32_t)M30);
        cd += (int64_t)P::mod30(0) * md;
        ce += (int64_t)P::mod30(0) * me;
        cd >>= 30; ce >>= 30;
        _Pragma("unroll") for (int k = 1; k < L; k++) {
          int32_t dk = d[k], ek = e[k];
          cd += (int64_t)u * dk + (int64_t)v * ek + (int64_t)P::mod30(k) * md;
          ce += (int64_t)q * dk + (int64_t)r * ek + (int64_t)P::mod30(k) * me;
          d[k - 1] = (int32_t)cd & M30; cd >>= 30;
          e[k - 1] = (int32_t)ce & M30; ce >>= 30;
        }
        d[L - 1] = (int32_t)cd;
        e[L - 1] = (int32_t)ce;
      }
      int32_t og = 0;
      _Pragma("unroll") for (int k = 0; k < L; k++) og |= g[k];
      if (og == 0) break;
    }
    // f = +-1; inverse = sign(f) * d, brought to [0, p)
    const bool fneg = f[L - 1] < 0;
    int64_t c = 0;
    _Pragma("unroll") for (int k = 0; k < L; k++) {   // d <- +-d, limbs renormalised
      c += fneg ? -(int64_t)d[k] : (int64_t)d[k];
      d[k] = (k < L - 1) ? ((int32_t)c & M30) : (int32_t)c;
      c >>= 30;
    }
    _Pragma("unroll 1") for (int rep = 0; rep < 3; rep++) {   // d in (-2p, 2p) -> [0, p)
      const bool neg = d[L - 1] < 0;
      // t = d - p (if d >= 0) to test d >= p
      int32_t t[L];
      int64_t cc = 0;
      _Pragma("unroll") for (int k = 0; k < L; k++) {
        cc += (int64_t)d[k] + (neg ? (int64_t)P::mod30(k) : -(int64_t)P::mod30(k));
        t[k] = (k < L - 1) ? ((int32_t)cc & M30) : (int32_t)cc;
        cc >>= 30;
      }
      const bool take = neg || t[L - 1] >= 0;     // negative: add p; non-negative and >= p: subtract p
      if (!take) break;
      _Pragma("unroll") for (int k = 0; k < L; k++) d[k] = t[k];
    }
    // 30-bit limbs -> packed words, then out of the plain domain: (aR)^-1 * R^3 / R = a^-1 R
    fe x;
    _Pragma("unroll") for (int w = 0; w < N; w++) {
      const int bit = 32 * w, k = bit / 30, sh = bit - 30 * k;
      uint32_t val = (uint32_t)d[k] >> sh;
      if (k + 1 < L) val |= (uint32_t)d[k + 1] << (30 - sh);
      if (sh > 28 && k + 2 < L) val |= (uint32_t)d[k + 2] << (60 - sh);
      x.v[w] = val;
    }
    fe r3; _Pragma("unroll") for (int i = 0; i < N; i++) r3.v[i] = P::r3(i);
    return mul(x, r3);
  }

  // a^(p-2); a = 0 -> 0.  (Not inlined: one copy per kernel.)
  MGB_NOINLINE_DEV static fe inv(fe a) {
    fe r = one();
    _Pragma("unroll 1") for (int k = N - 1; k >= 0; k--) {
      uint32_t w = 0;
      _Pragma("unroll") for (int kk = 0; kk < N; kk++) if (kk == k) w = P::pm2(kk);
      _Pragma("unroll 1") for (int bit = 31; bit >= 0; bit--) {
        r = sqr(r);
        if ((w >> bit) & 1) r = mul(r, a);
      }
    }
    return r;
  }
};

}  // namespace mgb
