// Warp-cooperative multi-limb Montgomery multiplication: ONE field element spread over the lanes
// of a 16-lane group, one 32-bit limb per lane.
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
};

}  // namespace mgb
