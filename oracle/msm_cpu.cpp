// oracle/msm_cpu.cpp -- CPU restatement of the reference's MSM algorithms.  TEST / BASELINE
// INFRASTRUCTURE ONLY: used by tests/ as an independent cross-check of the Python oracle and by
// bench.py's `cpu_baseline` / `--impl reference` legs.  Never linked into the product library.
//
// The reference itself (TypeScript + runtime-generated Wasm) cannot run in this image (no Node, no
// Wasm runtime), so its CPU path is restated here natively, phase by phase:
//   msm (batched-affine, GLV)       src/msm-batched-affine.ts:69-340
//     preparePointsAndScalars       :350-421   (G, -G, endo(G), -endo(G) with sign folding)
//     signed-digit slicing + counts :175-205
//     integrateBucketCounts         :423-447
//     sortPoints (copies points)    :456-502
//     accumulation rounds           :243-283   with batchAdd (src/curve-affine.ts:376-522)
//     reduceBucketsColumnProjective :556-583   with add-1998-cmo-2 / dbl-1998-cmo-2
//                                              (src/curve-projective.ts:51-253)
//     partition + final sum         :311-334
//   msmBasic (twisted Edwards)      src/msm-basic.ts:45-211, add-2008-hwcd-3 (src/curve-twisted-edwards.ts:84-165)
//   windowSize table                src/msm-common.ts:8-41
//   GLV decomposition               src/wasm/glv.ts:68-169 (constants passed in by the caller, computed by oracle/glv.py)
// Differences, on purpose: 64-bit limbs with unsigned __int128 instead of 29-bit limbs in i64 (a
// native build has 64x64->128 multiplies); values fully reduced; std::thread instead of workers.
// kind = "port" in bench.py's cpu_baseline: a native restatement, NOT the reference's Wasm.
#include <atomic>
#include <chrono>
#include <cmath>
#include <algorithm>
#include <cstdint>
#include <cstring>
#include <cstdio>
#include <cstdlib>
#include <thread>
#include <vector>

typedef unsigned __int128 u128;
typedef uint64_t u64;

namespace {

// ------------------------------------------------------------------ field, NL 64-bit limbs
template <int NL>
struct Fp {
  u64 p[NL], r1[NL], r2[NL];
  u64 inv;  // -p^-1 mod 2^64

  void init(const u64* mod) {
    memcpy(p, mod, sizeof(p));
    u64 x = 1;
    for (int i = 0; i < 6; i++) x *= 2 - p[0] * x;  // Newton: p^-1 mod 2^64
    inv = (u64)0 - x;
    // r1 = 2^(64 NL) mod p by doubling 1, r2 = 2^(128 NL) mod p
    u64 t[NL] = {0};
    t[0] = 1;
    for (int i = 0; i < 64 * NL; i++) dbl_mod(t);
    memcpy(r1, t, sizeof(t));
    for (int i = 0; i < 64 * NL; i++) dbl_mod(t);
    memcpy(r2, t, sizeof(t));
  }
  static bool geq(const u64* a, const u64* b) {
    for (int i = NL - 1; i >= 0; i--) {
      if (a[i] != b[i]) return a[i] > b[i];
    }
    return true;
  }
  static u64 add_n(u64* r, const u64* a, const u64* b) {
    u128 c = 0;
    for (int i = 0; i < NL; i++) { c += (u128)a[i] + b[i]; r[i] = (u64)c; c >>= 64; }
    return (u64)c;
  }
  static u64 sub_n(u64* r, const u64* a, const u64* b) {
    u64 borrow = 0;
    for (int i = 0; i < NL; i++) {
      u128 t = (u128)a[i] - b[i] - borrow;
      r[i] = (u64)t;
      borrow = (u64)(t >> 64) & 1;
    }
    return borrow;
  }
  void dbl_mod(u64* t) const {
    u64 c = add_n(t, t, t);
    if (c || geq(t, p)) sub_n(t, t, p);
  }
  void add(u64* r, const u64* a, const u64* b) const {
    u64 c = add_n(r, a, b);
    if (c || geq(r, p)) sub_n(r, r, p);
  }
  void sub(u64* r, const u64* a, const u64* b) const {
    if (sub_n(r, a, b)) add_n(r, r, p);
  }
  void neg(u64* r, const u64* a) const {
    bool z = true;
    for (int i = 0; i < NL; i++) z &= a[i] == 0;
    if (z) { memset(r, 0, 8 * NL); return; }
    sub_n(r, p, a);
  }
  static bool is_zero(const u64* a) {
    u64 o = 0;
    for (int i = 0; i < NL; i++) o |= a[i];
    return o == 0;
  }
  static bool eq(const u64* a, const u64* b) { return memcmp(a, b, 8 * NL) == 0; }
  // CIOS Montgomery product (the native analogue of src/wasm/multiply-montgomery.ts:58-136)
  void mul(u64* r, const u64* a, const u64* b) const {
    u64 t[NL + 2] = {0};
    for (int i = 0; i < NL; i++) {
      u128 c = 0;
      for (int j = 0; j < NL; j++) { c += (u128)a[j] * b[i] + t[j]; t[j] = (u64)c; c >>= 64; }
      c += t[NL];
      t[NL] = (u64)c;
      t[NL + 1] = (u64)(c >> 64);
      u64 m = t[0] * inv;
      c = (u128)m * p[0] + t[0];
      c >>= 64;
      for (int j = 1; j < NL; j++) { c += (u128)m * p[j] + t[j]; t[j - 1] = (u64)c; c >>= 64; }
      c += t[NL];
      t[NL - 1] = (u64)c;
      t[NL] = t[NL + 1] + (u64)(c >> 64);
    }
    if (t[NL] || geq(t, p)) sub_n(r, t, p);
    else memcpy(r, t, 8 * NL);
  }
  void sqr(u64* r, const u64* a) const { mul(r, a, a); }
  void to_mont(u64* r, const u64* a) const { mul(r, a, r2); }
  void from_mont(u64* r, const u64* a) const {
    u64 one[NL] = {0};
    one[0] = 1;
    mul(r, a, one);
  }
  // Montgomery inverse via binary extended Euclid (reference: Kaliski almost-inverse, src/wasm/inverse.ts:136-218)
  void inverse(u64* r, const u64* a) const {
    u64 u[NL], v[NL], x1[NL] = {0}, x2[NL] = {0}, one[NL] = {0};
    one[0] = 1;
    memcpy(u, a, sizeof(u));
    memcpy(v, p, sizeof(v));
    x1[0] = 1;
    auto halve = [&](u64* x) {
      u64 c = 0;
      if (x[0] & 1) c = add_n(x, x, p);
      for (int i = 0; i < NL - 1; i++) x[i] = (x[i] >> 1) | (x[i + 1] << 63);
      x[NL - 1] = (x[NL - 1] >> 1) | (c << 63);
    };
    auto shr = [&](u64* x) {
      for (int i = 0; i < NL - 1; i++) x[i] = (x[i] >> 1) | (x[i + 1] << 63);
      x[NL - 1] >>= 1;
    };
    while (!eq(u, one) && !eq(v, one)) {
      while (!(u[0] & 1)) { shr(u); halve(x1); }
      while (!(v[0] & 1)) { shr(v); halve(x2); }
      if (geq(u, v)) { sub_n(u, u, v); sub(x1, x1, x2); }
      else { sub_n(v, v, u); sub(x2, x2, x1); }
    }
    u64 x[NL];
    memcpy(x, eq(u, one) ? x1 : x2, sizeof(x));   // (aR)^-1
    u64 t[NL];
    mul(t, x, r2);                                 // a^-1 R^-1 * R^2 / R = a^-1
    mul(r, t, r2);                                 // a^-1 R
  }
};

struct GlvConsts {  // from oracle/glv.py (GlvScalar): all magnitudes, little-endian 64-bit limbs
  u64 m0[3], m1[3];          // |m_i| < 2^145
  u64 v[4][2];               // |v00|, |v01|, |v10|, |v11| < 2^128
  int sm0, sm1, sv[4];       // signs (+1 / -1)
  int m_bits, k_bits;        // m = 145, k = 116 for the 9x29-bit layout
  int max_bits;              // bound on |s0|, |s1| in bits
};

// signed 320-bit helpers for the GLV arithmetic
struct S320 {
  u64 w[5];
};
static void s_addmul(S320& acc, const u64* a, int na, const u64* b, int nb, int sign) {
  u64 prod[6] = {0};
  for (int i = 0; i < na; i++) {
    u128 c = 0;
    for (int j = 0; j < nb && i + j < 6; j++) { c += (u128)a[i] * b[j] + prod[i + j]; prod[i + j] = (u64)c; c >>= 64; }
    if (i + nb < 6) prod[i + nb] = (u64)c;
  }
  if (sign > 0) {
    u128 c = 0;
    for (int i = 0; i < 5; i++) { c += (u128)acc.w[i] + prod[i]; acc.w[i] = (u64)c; c >>= 64; }
  } else {
    u64 borrow = 0;
    for (int i = 0; i < 5; i++) { u128 t = (u128)acc.w[i] - prod[i] - borrow; acc.w[i] = (u64)t; borrow = (u64)(t >> 64) & 1; }
  }
}
// round(|x|*|y| / 2^m) : floor plus the bit below the cut (src/wasm/glv.ts:187-214)
static void mul_msb(u64* out3, const u64* x, int nx, const u64* y, int ny, int m) {
  u64 prod[8] = {0};
  for (int i = 0; i < nx; i++) {
    u128 c = 0;
    for (int j = 0; j < ny; j++) { c += (u128)x[i] * y[j] + prod[i + j]; prod[i + j] = (u64)c; c >>= 64; }
    prod[i + ny] = (u64)c;
  }
  auto bit_shift = [&](int sh, u64* o, int no) {
    int wq = sh / 64, b = sh % 64;
    for (int i = 0; i < no; i++) {
      u64 lo = (wq + i < 8) ? prod[wq + i] : 0, hi = (wq + i + 1 < 8) ? prod[wq + i + 1] : 0;
      o[i] = b ? (lo >> b) | (hi << (64 - b)) : lo;
    }
  };
  u64 fl[3], rb[1];
  bit_shift(m, fl, 3);
  bit_shift(m - 1, rb, 1);
  u128 c = (u128)fl[0] + (rb[0] & 1);
  out3[0] = (u64)c; c >>= 64;
  c += fl[1]; out3[1] = (u64)c; c >>= 64;
  c += fl[2]; out3[2] = (u64)c;
}
// s (4 limbs) -> |s0|, |s1| (2 limbs each) + negative flags
static void glv_decompose(const GlvConsts& g, const u64* s, u64* s0, u64* s1, bool& n0, bool& n1) {
  // shi = s >> k
  u64 shi[3];
  {
    int wq = g.k_bits / 64, b = g.k_bits % 64;
    for (int i = 0; i < 3; i++) {
      u64 lo = (wq + i < 4) ? s[wq + i] : 0, hi = (wq + i + 1 < 4) ? s[wq + i + 1] : 0;
      shi[i] = b ? (lo >> b) | (hi << (64 - b)) : lo;
    }
  }
  u64 x0[3], x1[3];
  mul_msb(x0, shi, 3, g.m0, 3, g.m_bits);
  mul_msb(x1, shi, 3, g.m1, 3, g.m_bits);
  int sx0 = g.sm0, sx1 = g.sm1;
  S320 a0 = {{s[0], s[1], s[2], s[3], 0}}, a1 = {{0, 0, 0, 0, 0}};
  s_addmul(a0, x0, 3, g.v[0], 2, sx0 * g.sv[0]);
  s_addmul(a0, x1, 3, g.v[1], 2, sx1 * g.sv[1]);
  s_addmul(a1, x0, 3, g.v[2], 2, sx0 * g.sv[2]);
  s_addmul(a1, x1, 3, g.v[3], 2, sx1 * g.sv[3]);
  auto fin = [](S320& a, u64* out, bool& neg) {
    neg = (a.w[4] >> 63) != 0;
    if (neg) {
      u128 c = 1;
      for (int i = 0; i < 5; i++) { c += (u64)~a.w[i]; a.w[i] = (u64)c; c >>= 64; }
    }
    out[0] = a.w[0];
    out[1] = a.w[1];
  };
  fin(a0, s0, n0);
  fin(a1, s1, n1);
}

static int window_size(int field_bits, int n) {  // src/msm-common.ts:8-41
  if (field_bits > 260) {
    switch (n) {
      case 14: return 13;
      case 15: case 16: case 17: case 18: return 14;
      case 19: case 20: return 18;
    }
  } else if (n == 16) return 12;
  return n - 1 > 1 ? n - 1 : 1;
}
static int log2_ceil(size_t n) {
  int l = 0;
  while (((size_t)1 << l) < n) l++;
  return l;
}
static inline uint32_t bit_slice(const u64* x, int nl, int start, int len) {
  int wq = start / 64, b = start % 64;
  if (wq >= nl) return 0;
  u64 lo = x[wq], hi = (wq + 1 < nl) ? x[wq + 1] : 0;
  u64 v = b ? (lo >> b) | (hi << (64 - b)) : lo;
  return (uint32_t)(v & (((u64)1 << len) - 1));
}

struct Tracer {
  bool on;
  std::chrono::steady_clock::time_point t;
  Tracer() : on(getenv("MSM_CPU_TRACE") != nullptr), t(std::chrono::steady_clock::now()) {}
  void mark(const char* label) {
    if (!on) return;
    auto n = std::chrono::steady_clock::now();
    fprintf(stderr, "  [msm_cpu] %-28s %8.2f ms\n", label, std::chrono::duration<double, std::milli>(n - t).count());
    t = n;
  }
};

template <class F>
static void run_threads(int T, F f) {
  std::vector<std::thread> th;
  for (int t = 1; t < T; t++) th.emplace_back(f, t);
  f(0);
  for (auto& x : th) x.join();
}
static void range(size_t n, int t, int T, size_t& lo, size_t& hi) {  // src/threads/threads.ts:354-359
  size_t per = (n + T - 1) / T;
  lo = std::min(n, per * t);
  hi = std::min(n, lo + per);
}

// ------------------------------------------------------------------ Weierstrass, batched affine
template <int NL>
struct WCurve {
  Fp<NL> F;
  u64 beta[NL];  // Montgomery form

  struct Aff { u64 x[NL], y[NL]; uint32_t nz; };
  struct Proj { u64 X[NL], Y[NL], Z[NL]; uint32_t nz; };

  void proj_zero(Proj& P) const { memset(&P, 0, sizeof(P)); }
  void from_affine(Proj& P, const Aff& A) const {
    if (!A.nz) { proj_zero(P); return; }
    memcpy(P.X, A.x, 8 * NL); memcpy(P.Y, A.y, 8 * NL); memcpy(P.Z, F.r1, 8 * NL); P.nz = 1;
  }
  // dbl-1998-cmo-2, a = 0 (src/curve-projective.ts:202-253)
  void dbl(Proj& R, const Proj& P) const {
    if (!P.nz) { proj_zero(R); return; }
    u64 w[NL], s[NL], ss[NL], sss[NL], Rr[NL], B[NL], h[NL], t[NL], t2[NL];
    F.sqr(t, P.X); F.add(w, t, t); F.add(w, w, t);
    F.mul(s, P.Y, P.Z);
    F.sqr(ss, s); F.mul(sss, s, ss);
    F.mul(Rr, P.Y, s);
    F.mul(B, P.X, Rr);
    F.sqr(h, w);
    F.add(t, B, B); F.add(t, t, t); F.add(t2, t, t);   // t = 4B, t2 = 8B
    F.sub(h, h, t2);
    Proj O;
    F.mul(O.X, h, s); F.add(O.X, O.X, O.X);
    F.sub(t, t, h); F.mul(t, w, t);
    F.sqr(t2, Rr); F.add(t2, t2, t2); F.add(t2, t2, t2); F.add(t2, t2, t2);
    F.sub(O.Y, t, t2);
    F.add(O.Z, sss, sss); F.add(O.Z, O.Z, O.Z); F.add(O.Z, O.Z, O.Z);
    O.nz = F.is_zero(O.Z) ? 0 : 1;
    R = O;
  }
  // add-1998-cmo-2 (src/curve-projective.ts:51-200), complete
  void add(Proj& R, const Proj& P, const Proj& Q) const {
    if (!P.nz) { R = Q; return; }
    if (!Q.nz) { R = P; return; }
    u64 Y1Z2[NL], X1Z2[NL], Z1Z2[NL], u[NL], uu[NL], v[NL], vv[NL], vvv[NL], Rr[NL], A[NL], t[NL];
    F.mul(Y1Z2, P.Y, Q.Z); F.mul(X1Z2, P.X, Q.Z); F.mul(Z1Z2, P.Z, Q.Z);
    F.mul(u, Q.Y, P.Z); F.sub(u, u, Y1Z2);
    F.mul(v, Q.X, P.Z); F.sub(v, v, X1Z2);
    if (F.is_zero(v)) {
      if (F.is_zero(u)) { dbl(R, P); return; }
      proj_zero(R); return;
    }
    F.sqr(uu, u); F.sqr(vv, v); F.mul(vvv, v, vv); F.mul(Rr, vv, X1Z2);
    F.mul(A, uu, Z1Z2); F.sub(A, A, vvv); F.sub(A, A, Rr); F.sub(A, A, Rr);
    Proj O;
    F.mul(O.X, v, A);
    F.sub(t, Rr, A); F.mul(t, u, t); F.mul(A, vvv, Y1Z2); F.sub(O.Y, t, A);
    F.mul(O.Z, vvv, Z1Z2);
    O.nz = 1;
    R = O;
  }
  // one batch of independent affine additions S_i = G_i + H_i sharing one inversion, result into G_i
  // (batchAddNew, src/curve-affine.ts:376-458: handles zero, doubling and cancellation)
  void batch_add(Aff** G, Aff** H, size_t n, std::vector<u64>& scratch) const {
    if (n == 0) return;
    scratch.resize(2 * n * NL);
    u64* den = scratch.data();
    u64* pre = den + n * NL;
    std::vector<uint8_t> kind(n);
    u64 run[NL];
    memcpy(run, F.r1, sizeof(run));
    for (size_t i = 0; i < n; i++) {
      u64* d = den + i * NL;
      const Aff &A = *G[i], &B = *H[i];
      if (!B.nz) { kind[i] = 2; memcpy(d, F.r1, 8 * NL); }
      else if (!A.nz) { kind[i] = 3; memcpy(d, F.r1, 8 * NL); }
      else {
        F.sub(d, B.x, A.x);
        if (F.is_zero(d)) {
          if (F.eq(A.y, B.y) && !F.is_zero(A.y)) { kind[i] = 1; F.add(d, A.y, A.y); }
          else { kind[i] = 4; memcpy(d, F.r1, 8 * NL); }
        } else kind[i] = 0;
      }
      memcpy(pre + i * NL, run, 8 * NL);
      F.mul(run, run, d);
    }
    u64 u[NL];
    F.inverse(u, run);
    for (size_t i = n; i-- > 0;) {
      u64 inv[NL], m[NL], num[NL], t[NL];
      F.mul(inv, u, pre + i * NL);
      F.mul(u, u, den + i * NL);
      Aff& A = *G[i];
      const Aff& B = *H[i];
      switch (kind[i]) {
        case 2: break;
        case 3: A = B; break;
        case 4: A.nz = 0; break;
        default: {
          if (kind[i] == 1) { F.sqr(t, A.x); F.add(num, t, t); F.add(num, num, t); }
          else F.sub(num, B.y, A.y);
          F.mul(m, num, inv);
          u64 x3[NL], y3[NL];
          F.sqr(x3, m); F.sub(x3, x3, A.x); F.sub(x3, x3, B.x);
          F.sub(t, A.x, x3); F.mul(y3, m, t); F.sub(y3, y3, A.y);
          memcpy(A.x, x3, 8 * NL); memcpy(A.y, y3, 8 * NL);
        }
      }
    }
  }
};

template <int NL>
static int msm_weierstrass(const u64* mod, const u64* beta_plain, int field_bits, const GlvConsts& glv,
                           const uint8_t* scalars, const uint8_t* points, size_t N, int T, int c_opt,
                           uint8_t* out_xy, int* out_zero, double* ms_out) {
  typedef WCurve<NL> C;
  typedef typename C::Aff Aff;
  typedef typename C::Proj Proj;
  C cv;
  cv.F.init(mod);
  cv.F.to_mont(cv.beta, beta_plain);
  const Fp<NL>& F = cv.F;
  const int CB = (field_bits + 7) / 8;
  // --- untimed: bytes -> Montgomery affine points (Parallel.pointsFromBytes, src/parallel.ts:97-116)
  std::vector<Aff> pts(N);
  run_threads(T, [&](int t) {
    size_t lo, hi; range(N, t, T, lo, hi);
    for (size_t i = lo; i < hi; i++) {
      u64 x[NL] = {0}, y[NL] = {0};
      memcpy(x, points + i * 2 * CB, CB); memcpy(y, points + i * 2 * CB + CB, CB);
      F.to_mont(pts[i].x, x); F.to_mont(pts[i].y, y); pts[i].nz = 1;
    }
  });
  auto t_start = std::chrono::steady_clock::now();
  Tracer tr;
  Proj result;
  cv.proj_zero(result);
  if (N > 0) {
    const int n = log2_ceil(N);
    const int c = c_opt > 0 ? c_opt : window_size(field_bits, n);
    const int b = glv.max_bits;
    const int K = (b + 1 + c - 1) / c;
    const size_t L = (size_t)1 << (c - 1);
    // --- preparePointsAndScalars: 4 variants per point, two half scalars
    std::vector<Aff> prep(4 * N);
    std::vector<u64> half(2 * 2 * N);
    run_threads(T, [&](int t) {
      size_t lo, hi; range(N, t, T, lo, hi);
      for (size_t i = lo; i < hi; i++) {
        u64 s[4];
        memcpy(s, scalars + 32 * i, 32);
        bool n0, n1;
        glv_decompose(glv, s, &half[4 * i], &half[4 * i + 2], n0, n1);
        Aff* q = &prep[4 * i];
        q[0] = pts[i];
        q[1] = pts[i];
        F.neg(q[1].y, pts[i].y);
        if (n0) std::swap(q[0], q[1]);          // sign folding: slot 0 is what a positive digit adds
        q[2] = q[0]; F.mul(q[2].x, q[0].x, cv.beta);
        q[3] = q[1]; q[3].x[0] = q[2].x[0]; memcpy(q[3].x, q[2].x, 8 * NL);
        if (n1 != n0) std::swap(q[2], q[3]);
      }
    });
    tr.mark("prepare points & scalars");
    // --- signed digits + bucket counts
    std::vector<std::vector<uint32_t>> slices(K, std::vector<uint32_t>(2 * N));
    std::vector<std::atomic<uint32_t>> counts((size_t)K * (L + 1));
    for (auto& x : counts) x.store(0, std::memory_order_relaxed);
    std::vector<uint32_t> maxsz(T, 0);
    run_threads(T, [&](int t) {
      size_t lo, hi; range(N, t, T, lo, hi);
      uint32_t mx = 0;
      for (size_t i = 2 * lo; i < 2 * hi; i++) {
        uint32_t carry = 0;
        for (int k = 0; k < K; k++) {
          uint32_t l = bit_slice(&half[2 * i], 2, k * c, c) + carry;
          if (l > L) { l = (uint32_t)(2 * L) - l; carry = 1; } else carry = 0;
          slices[k][i] = l | (carry << 31);
          if (l) { uint32_t v = counts[(size_t)k * (L + 1) + l].fetch_add(1, std::memory_order_relaxed) + 1; if (v > mx) mx = v; }
        }
      }
      maxsz[t] = mx;
    });
    uint32_t max_bucket = 0;
    for (int t = 0; t < T; t++) max_bucket = std::max(max_bucket, maxsz[t]);
    tr.mark("slice scalars & count");
    // --- integrate counts -> bucket bounds (main thread)
    std::vector<std::vector<uint32_t>> start(K, std::vector<uint32_t>(L + 2));
    std::vector<std::vector<uint32_t>> cursor(K, std::vector<uint32_t>(L + 2));
    for (int k = 0; k < K; k++) {
      uint32_t run = 0;
      for (size_t l = 1; l <= L; l++) {
        start[k][l] = run; cursor[k][l] = run;
        run += counts[(size_t)k * (L + 1) + l].load(std::memory_order_relaxed);
      }
      start[k][L + 1] = run;
    }
    tr.mark("integrate bucket counts");
    // --- sortPoints: copy points into bucket order, one window per thread slice
    std::vector<std::vector<Aff>> sorted(K);
    for (int k = 0; k < K; k++) sorted[k].resize(start[k][L + 1]);
    run_threads(T, [&](int t) {
      size_t lo, hi; range(K, t, T, lo, hi);
      for (size_t k = lo; k < hi; k++)
        for (size_t i = 0; i < 2 * N; i++) {
          uint32_t l = slices[k][i], carry = l >> 31;
          l &= 0x7fffffffu;
          if (!l) continue;
          sorted[k][cursor[k][l]++] = prep[2 * i + carry];
        }
    });
    tr.mark("sort points");
    // --- accumulation rounds (implicit binary tree per bucket)
    run_threads(T, [&](int t) {
      size_t lo, hi; range((size_t)K * L, t, T, lo, hi);
      std::vector<Aff*> G, H;
      std::vector<u64> scratch;
      auto tt0 = std::chrono::steady_clock::now();
      size_t tot = 0;
      for (uint32_t m = 1; m < max_bucket; m *= 2) {
        G.clear(); H.clear();
        for (size_t i = lo; i < hi; i++) {
          size_t k = i / L, l = i % L + 1;
          uint32_t b0 = start[k][l], b1 = start[k][l + 1];
          for (uint32_t a = b0; a + m < b1; a += 2 * m) { G.push_back(&sorted[k][a]); H.push_back(&sorted[k][a + m]); }
        }
        cv.batch_add(G.data(), H.data(), G.size(), scratch);
        tot += G.size();
      }
      if (tr.on) fprintf(stderr, "    thread %d: buckets [%zu,%zu) %zu adds %.1f ms\n", t, lo, hi, tot,
                         std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - tt0).count());
    });
    tr.mark("bucket accumulation");
    // --- bucket reduction per chunk (computeBucketsSplit: K*L buckets split evenly over threads)
    const size_t total = (size_t)K * L, per = (total + T - 1) / T;
    struct Chunk { int k; size_t lstart, len; Proj sum; };
    std::vector<std::vector<Chunk>> chunks(T);
    {
      int t = 0; size_t rem = per;
      for (int k = 0; k < K; k++) {
        size_t remL = L, lstart = 1;
        while (remL > 0) {
          size_t len = std::min(remL, rem);
          chunks[t].push_back({k, lstart, len, Proj()});
          remL -= len; lstart += len; rem -= len;
          if (rem == 0) { t++; rem = per; }
        }
      }
    }
    run_threads(T, [&](int t) {
      for (auto& ch : chunks[t]) {
        Proj row, tri, bp;
        cv.proj_zero(row); cv.proj_zero(tri);
        for (size_t j = ch.len; j-- > 0;) {
          size_t l = ch.lstart + j;
          uint32_t b0 = start[ch.k][l], b1 = start[ch.k][l + 1];
          if (b0 != b1) { cv.from_affine(bp, sorted[ch.k][b0]); cv.add(row, row, bp); }
          cv.add(tri, tri, row);
        }
        size_t ls = ch.lstart - 1;
        while (true) {
          if (ls & 1) cv.add(tri, tri, row);
          if ((ls >>= 1) == 0) break;
          cv.dbl(row, row);
        }
        ch.sum = tri;
      }
    });
    tr.mark("bucket reduction");
    // --- partition sums + Horner
    std::vector<Proj> part(K);
    for (int k = 0; k < K; k++) cv.proj_zero(part[k]);
    for (int t = 0; t < T; t++) for (auto& ch : chunks[t]) cv.add(part[ch.k], part[ch.k], ch.sum);
    result = part[K - 1];
    for (int k = K - 2; k >= 0; k--) {
      for (int j = 0; j < c; j++) cv.dbl(result, result);
      cv.add(result, result, part[k]);
    }
  }
  // toAffine + out of Montgomery form (src/curve-projective.ts:335-349, src/curve-affine.ts:220-233)
  memset(out_xy, 0, 2 * CB);
  if (!result.nz || F.is_zero(result.Z)) *out_zero = 1;
  else {
    u64 zi[NL], x[NL], y[NL];
    F.inverse(zi, result.Z);
    F.mul(x, result.X, zi); F.mul(y, result.Y, zi);
    F.from_mont(x, x); F.from_mont(y, y);
    memcpy(out_xy, x, CB); memcpy(out_xy + CB, y, CB);
    *out_zero = 0;
  }
  auto t_end = std::chrono::steady_clock::now();
  if (ms_out) *ms_out = std::chrono::duration<double, std::milli>(t_end - t_start).count();
  return 0;
}

// ------------------------------------------------------------------ twisted Edwards, msmBasic
template <int NL>
struct TECurve {
  Fp<NL> F;
  u64 k[NL];  // 2d, Montgomery form
  struct Ext { u64 X[NL], Y[NL], Z[NL], T[NL]; };
  void zero(Ext& P) const { memset(&P, 0, sizeof(P)); memcpy(P.Y, F.r1, 8 * NL); memcpy(P.Z, F.r1, 8 * NL); }
  // add-2008-hwcd-3 (src/curve-twisted-edwards.ts:110-164); sub: swap (Y2 -/+ X2), negate T2
  void add(Ext& R, const Ext& P, const Ext& Q, bool subtract = false) const {
    u64 A[NL], B[NL], Cc[NL], D[NL], E[NL], Ff[NL], G[NL], H[NL], t1[NL], t2[NL];
    F.sub(t1, P.Y, P.X);
    if (subtract) F.add(t2, Q.Y, Q.X); else F.sub(t2, Q.Y, Q.X);
    F.mul(A, t1, t2);
    F.add(t1, P.Y, P.X);
    if (subtract) F.sub(t2, Q.Y, Q.X); else F.add(t2, Q.Y, Q.X);
    F.mul(B, t1, t2);
    F.mul(Cc, P.T, Q.T); F.mul(Cc, Cc, k);
    if (subtract) F.neg(Cc, Cc);
    F.mul(D, P.Z, Q.Z); F.add(D, D, D);
    F.sub(E, B, A); F.sub(Ff, D, Cc); F.add(G, D, Cc); F.add(H, B, A);
    Ext O;
    F.mul(O.X, E, Ff); F.mul(O.Y, G, H); F.mul(O.T, E, H); F.mul(O.Z, Ff, G);
    R = O;
  }
};

template <int NL>
static int msm_te(const u64* mod, u64 d_small, int scalar_bits, const uint8_t* scalars, const uint8_t* points, size_t N, int T,
                  int c_opt, uint8_t* out_xy, int* out_zero, double* ms_out) {
  typedef TECurve<NL> C;
  typedef typename C::Ext Ext;
  C cv;
  cv.F.init(mod);
  const Fp<NL>& F = cv.F;
  {
    u64 kk[NL] = {0};
    kk[0] = 2 * d_small;
    F.to_mont(cv.k, kk);
  }
  const int CB = 8 * NL;
  std::vector<Ext> pts(N);  // extended with Z = 1 (Parallel.pointsFromBytes, src/parallel.ts:209-232)
  run_threads(T, [&](int t) {
    size_t lo, hi; range(N, t, T, lo, hi);
    for (size_t i = lo; i < hi; i++) {
      u64 x[NL], y[NL];
      memcpy(x, points + i * 2 * CB, CB); memcpy(y, points + i * 2 * CB + CB, CB);
      F.to_mont(pts[i].X, x); F.to_mont(pts[i].Y, y);
      memcpy(pts[i].Z, F.r1, 8 * NL);
      F.mul(pts[i].T, pts[i].X, pts[i].Y);
    }
  });
  auto t_start = std::chrono::steady_clock::now();
  Ext result;
  cv.zero(result);
  if (N > 0) {
    const int n = log2_ceil(N);
    const int b = scalar_bits;
    const int c = c_opt > 0 ? c_opt : window_size(b, n);
    const int K = (b + 1 + c - 1) / c;
    const size_t L = (size_t)1 << (c - 1);
    // digits: the reference slices on the main thread only (src/msm-basic.ts:72-91)
    std::vector<std::vector<uint32_t>> dig(K, std::vector<uint32_t>(N));
    for (size_t i = 0; i < N; i++) {
      u64 s[4];
      memcpy(s, scalars + 32 * i, 32);
      uint32_t carry = 0;
      for (int k = 0; k < K; k++) {
        uint32_t l = bit_slice(s, 4, k * c, c) + carry;
        if (l > L) { l = (uint32_t)(2 * L) - l; carry = 1; } else carry = 0;
        dig[k][i] = l | (carry << 31);
      }
    }
    // splitBuckets (src/msm-common.ts:72-172) incl. the reduced weight of the top window
    struct Chunk { int k; size_t lstart, len; Ext sum; };
    std::vector<std::vector<Chunk>> chunks(T);
    {
      int overlap = b % c;
      double Ll = (double)((size_t)1 << overlap);
      double wl = overlap == 0 ? (double)L / Ll / 32.0 : (double)L / Ll;
      double totalWork = (double)(K - 1) * L + Ll * wl;
      double per = std::ceil(totalWork / T);
      int t = 0;
      double rem = per;
      for (int k = 0; k < K - 1; k++) {
        double remL = (double)L;
        size_t lstart = 1;
        while (remL > 0) {
          double len = std::min(remL, rem);
          chunks[std::min(t, T - 1)].push_back({k, lstart, (size_t)len, Ext()});
          remL -= len; rem -= len; lstart += (size_t)len;
          if (rem <= 0) { t++; rem = per; }
        }
      }
      {
        int k = K - 1;
        double remWork = Ll * wl, remB = (double)L;
        size_t lstart = 1;
        while (remWork > 0 && remB > 0) {
          double len = std::min(std::ceil(rem / wl), remB);
          chunks[std::min(t, T - 1)].push_back({k, lstart, (size_t)len, Ext()});
          remWork -= wl * len; remB -= len; rem -= wl * len; lstart += (size_t)len;
          if (rem <= 0) { t++; rem = per; }
        }
        if (remB > 0) chunks[T - 1].push_back({k, lstart, (size_t)remB, Ext()});  // buckets the reference proves empty; kept for safety
      }
    }
    run_threads(T, [&](int t) {
      for (auto& ch : chunks[t]) {
        std::vector<Ext> buckets(ch.len);
        for (auto& B : buckets) cv.zero(B);
        for (size_t i = 0; i < N; i++) {   // every chunk scans all N points (src/msm-basic.ts:110-122)
          uint32_t l = dig[ch.k][i], carry = l >> 31;
          l &= 0x7fffffffu;
          if (l < ch.lstart || l >= ch.lstart + ch.len) continue;
          cv.add(buckets[l - ch.lstart], buckets[l - ch.lstart], pts[i], carry == 1);
        }
        Ext row, tri;
        cv.zero(row); cv.zero(tri);
        for (size_t j = ch.len; j-- > 0;) { cv.add(row, row, buckets[j]); cv.add(tri, tri, row); }
        size_t ls = ch.lstart - 1;
        while (true) {
          if (ls & 1) cv.add(tri, tri, row);
          if ((ls >>= 1) == 0) break;
          cv.add(row, row, row);
        }
        ch.sum = tri;
      }
    });
    std::vector<Ext> part(K);
    for (int k = 0; k < K; k++) cv.zero(part[k]);
    for (int t = 0; t < T; t++) for (auto& ch : chunks[t]) cv.add(part[ch.k], part[ch.k], ch.sum);
    result = part[K - 1];
    for (int k = K - 2; k >= 0; k--) {
      for (int j = 0; j < c; j++) cv.add(result, result, result);
      cv.add(result, result, part[k]);
    }
  }
  u64 zi[NL], x[NL], y[NL], one[NL] = {0};
  one[0] = 1;
  F.inverse(zi, result.Z);
  F.mul(x, result.X, zi); F.mul(y, result.Y, zi);
  F.from_mont(x, x); F.from_mont(y, y);
  memcpy(out_xy, x, CB); memcpy(out_xy + CB, y, CB);
  *out_zero = (F.is_zero(x) && F.eq(y, one)) ? 1 : 0;
  auto t_end = std::chrono::steady_clock::now();
  if (ms_out) *ms_out = std::chrono::duration<double, std::milli>(t_end - t_start).count();
  return 0;
}

// ------------------------------------------------------------------ seeded known-dlog points
// P_i = a_i * G with a_i = splitmix64(seed, i): the same point set the GPU arm generates with
// mgb_random_points, built here the way the reference builds its test points (sums of window-table
// multiples of a basis point, batch-normalised: src/curve-random.ts:36-92).
static inline u64 splitmix64(u64 seed, u64 i) {
  u64 z = seed + (i + 1) * 0x9E3779B97F4A7C15ull;
  z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
  z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
  return z ^ (z >> 31);
}

template <int NL>
static void gen_points_w(const u64* mod, const u64* gx, const u64* gy, int CB, u64 seed, size_t n, int T, uint8_t* out_xy) {
  typedef WCurve<NL> C;
  typedef typename C::Proj Proj;
  C cv;
  cv.F.init(mod);
  const Fp<NL>& F = cv.F;
  const int W = 16, KW = 4;
  std::vector<std::vector<Proj>> tab(KW, std::vector<Proj>((size_t)1 << W));
  Proj base;
  F.to_mont(base.X, gx); F.to_mont(base.Y, gy); memcpy(base.Z, F.r1, 8 * NL); base.nz = 1;
  std::vector<Proj> bases(KW);
  for (int k = 0; k < KW; k++) { bases[k] = base; for (int j = 0; j < W; j++) cv.dbl(base, base); }
  run_threads(std::min(T, KW), [&](int t) {
    for (int k = t; k < KW; k += std::min(T, KW)) {
      cv.proj_zero(tab[k][0]);
      for (size_t j = 1; j < ((size_t)1 << W); j++) cv.add(tab[k][j], tab[k][j - 1], bases[k]);
    }
  });
  run_threads(T, [&](int t) {
    size_t lo, hi; range(n, t, T, lo, hi);
    if (lo >= hi) return;
    std::vector<Proj> P(hi - lo);
    std::vector<u64> pre((hi - lo) * NL);
    u64 run[NL];
    memcpy(run, F.r1, sizeof(run));
    for (size_t i = lo; i < hi; i++) {
      u64 a = splitmix64(seed, i);
      if (a == 0) a = 1;
      Proj acc = tab[0][a & 0xffff];
      for (int k = 1; k < KW; k++) cv.add(acc, acc, tab[k][(a >> (16 * k)) & 0xffff]);
      P[i - lo] = acc;
      memcpy(&pre[(i - lo) * NL], run, 8 * NL);
      F.mul(run, run, acc.Z);
    }
    u64 u[NL];
    F.inverse(u, run);
    for (size_t i = hi; i-- > lo;) {
      u64 zi[NL], x[NL], y[NL];
      F.mul(zi, u, &pre[(i - lo) * NL]);
      F.mul(u, u, P[i - lo].Z);
      F.mul(x, P[i - lo].X, zi); F.mul(y, P[i - lo].Y, zi);
      F.from_mont(x, x); F.from_mont(y, y);
      memcpy(out_xy + i * 2 * CB, x, CB); memcpy(out_xy + i * 2 * CB + CB, y, CB);
    }
  });
}

template <int NL>
static void gen_points_te(const u64* mod, u64 d_small, const u64* gx, const u64* gy, u64 seed, size_t n, int T, uint8_t* out_xy) {
  typedef TECurve<NL> C;
  typedef typename C::Ext Ext;
  C cv;
  cv.F.init(mod);
  const Fp<NL>& F = cv.F;
  { u64 kk[NL] = {0}; kk[0] = 2 * d_small; F.to_mont(cv.k, kk); }
  const int CB = 8 * NL, W = 16, KW = 4;
  std::vector<std::vector<Ext>> tab(KW, std::vector<Ext>((size_t)1 << W));
  Ext base;
  F.to_mont(base.X, gx); F.to_mont(base.Y, gy); memcpy(base.Z, F.r1, 8 * NL); F.mul(base.T, base.X, base.Y);
  std::vector<Ext> bases(KW);
  for (int k = 0; k < KW; k++) { bases[k] = base; for (int j = 0; j < W; j++) cv.add(base, base, base); }
  run_threads(std::min(T, KW), [&](int t) {
    for (int k = t; k < KW; k += std::min(T, KW)) {
      cv.zero(tab[k][0]);
      for (size_t j = 1; j < ((size_t)1 << W); j++) cv.add(tab[k][j], tab[k][j - 1], bases[k]);
    }
  });
  run_threads(T, [&](int t) {
    size_t lo, hi; range(n, t, T, lo, hi);
    for (size_t i = lo; i < hi; i++) {
      u64 a = splitmix64(seed, i);
      if (a == 0) a = 1;
      Ext acc = tab[0][a & 0xffff];
      for (int k = 1; k < KW; k++) cv.add(acc, acc, tab[k][(a >> (16 * k)) & 0xffff]);
      u64 zi[NL], x[NL], y[NL];
      F.inverse(zi, acc.Z);
      F.mul(x, acc.X, zi); F.mul(y, acc.Y, zi);
      F.from_mont(x, x); F.from_mont(y, y);
      memcpy(out_xy + i * 2 * CB, x, CB); memcpy(out_xy + i * 2 * CB + CB, y, CB);
    }
  });
}

}  // namespace

extern "C" {

struct ref_glv_consts {
  uint64_t m0[3], m1[3];
  uint64_t v[4][2];
  int32_t sm0, sm1, sv[4];
  int32_t m_bits, k_bits, max_bits;
};

// curve: 0 = BLS12-377 G1, 1 = Pallas, 2 = ed-on-BLS12-377, 3 = BLS12-381 G1.  mod/beta: little-endian 64-bit limbs.
// Returns 0; *ms_out = wall time of the MSM proper (inputs already converted, like the reference's timing).
int ref_msm(int curve, const uint64_t* mod, const uint64_t* beta_or_d, const ref_glv_consts* glv,
            const uint8_t* scalars, const uint8_t* points, size_t n, int threads, int c,
            uint8_t* out_xy, int* out_zero, double* ms_out) {
  if (threads < 1) threads = 1;
  GlvConsts g;
  if (glv) {
    memcpy(g.m0, glv->m0, sizeof(g.m0)); memcpy(g.m1, glv->m1, sizeof(g.m1)); memcpy(g.v, glv->v, sizeof(g.v));
    g.sm0 = glv->sm0; g.sm1 = glv->sm1;
    for (int i = 0; i < 4; i++) g.sv[i] = glv->sv[i];
    g.m_bits = glv->m_bits; g.k_bits = glv->k_bits; g.max_bits = glv->max_bits;
  }
  switch (curve) {
    case 0: return msm_weierstrass<6>(mod, beta_or_d, 377, g, scalars, points, n, threads, c, out_xy, out_zero, ms_out);
    case 1: return msm_weierstrass<4>(mod, beta_or_d, 255, g, scalars, points, n, threads, c, out_xy, out_zero, ms_out);
    case 2: return msm_te<4>(mod, beta_or_d[0], 251, scalars, points, n, threads, c, out_xy, out_zero, ms_out);
    case 3: return msm_weierstrass<6>(mod, beta_or_d, 381, g, scalars, points, n, threads, c, out_xy, out_zero, ms_out);
  }
  return -1;
}

// Fills out_xy with n canonical x||y points a_i*G (see gen_points_w); gx, gy = generator, plain form.
int ref_known_dlog_points(int curve, const uint64_t* mod, const uint64_t* d_or_null, const uint64_t* gx, const uint64_t* gy,
                          uint64_t seed, size_t n, int threads, uint8_t* out_xy) {
  if (threads < 1) threads = 1;
  switch (curve) {
    case 0: gen_points_w<6>(mod, gx, gy, 48, seed, n, threads, out_xy); return 0;
    case 1: gen_points_w<4>(mod, gx, gy, 32, seed, n, threads, out_xy); return 0;
    case 2: gen_points_te<4>(mod, d_or_null[0], gx, gy, seed, n, threads, out_xy); return 0;
    case 3: gen_points_w<6>(mod, gx, gy, 48, seed, n, threads, out_xy); return 0;
  }
  return -1;
}
}
