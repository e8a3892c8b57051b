// Inline-PTX carry-chain primitives (sm_100a).  ptxas fuses a mad.lo.cc / madc.hi.cc pair on
// the same operands into IMAD.WIDE.U32(.X) with a predicate carry; see profiles/ for the SASS mix.
#pragma once
#include <cstdint>


#ifdef MGB_HOST_EMU
// Test-only host emulation of the carry-flag primitives (tests/host_emu/*.cpp compile the math
// headers with g++ to check formulas without a GPU).  Never part of the shipped library.
#define MGB_DEV inline
#define MGB_NOINLINE_DEV inline
#define MGB_CTZ(x) __builtin_ctz(x)
#ifndef __CUDACC__
#define __host__
#define __device__
#endif
namespace mgb {
namespace ptx {
static thread_local uint32_t CF = 0;
inline uint32_t mul_lo(uint32_t a, uint32_t b) { return (uint32_t)((uint64_t)a * b); }
inline uint32_t mul_hi(uint32_t a, uint32_t b) { return (uint32_t)(((uint64_t)a * b) >> 32); }
inline uint32_t add3(uint32_t a, uint32_t b, uint32_t cin, bool setcf) { uint64_t t = (uint64_t)a + b + cin; if (setcf) CF = (uint32_t)(t >> 32); return (uint32_t)t; }
inline uint32_t mad_lo_cc(uint32_t a, uint32_t b, uint32_t c) { return add3(mul_lo(a, b), c, 0, true); }
inline uint32_t mad_hi_cc(uint32_t a, uint32_t b, uint32_t c) { return add3(mul_hi(a, b), c, 0, true); }
inline uint32_t madc_lo_cc(uint32_t a, uint32_t b, uint32_t c) { return add3(mul_lo(a, b), c, CF, true); }
inline uint32_t madc_hi_cc(uint32_t a, uint32_t b, uint32_t c) { return add3(mul_hi(a, b), c, CF, true); }
inline uint32_t madc_hi(uint32_t a, uint32_t b, uint32_t c) { return add3(mul_hi(a, b), c, CF, false); }
inline uint32_t add_cc(uint32_t a, uint32_t b) { return add3(a, b, 0, true); }
inline uint32_t addc_cc(uint32_t a, uint32_t b) { return add3(a, b, CF, true); }
inline uint32_t addc(uint32_t a, uint32_t b) { return add3(a, b, CF, false); }
// PTX borrow semantics: CF holds the *borrow* after sub.cc (1 = borrow occurred)
inline uint32_t sub3(uint32_t a, uint32_t b, uint32_t bin, bool setcf) { uint64_t t = (uint64_t)a - b - bin; if (setcf) CF = (uint32_t)((t >> 32) & 1); return (uint32_t)t; }
inline uint32_t sub_cc(uint32_t a, uint32_t b) { return sub3(a, b, 0, true); }
inline uint32_t subc_cc(uint32_t a, uint32_t b) { return sub3(a, b, CF, true); }
inline uint32_t subc(uint32_t a, uint32_t b) { return sub3(a, b, CF, false); }
// x + y (64 bit); `carries` is incremented by the carry out
inline unsigned long long add64_count(unsigned long long x, unsigned long long y, uint32_t& carries) {
  const unsigned long long r = x + y;
  carries += r < x ? 1u : 0u;
  return r;
}
}  // namespace ptx
}  // namespace mgb
#else
#define MGB_DEV __device__ __forceinline__
#define MGB_NOINLINE_DEV __device__ __noinline__
#define MGB_CTZ(x) (__ffs((int)(x)) - 1)

namespace mgb {
namespace ptx {
MGB_DEV uint32_t mul_lo(uint32_t a, uint32_t b) { uint32_t r; asm volatile("mul.lo.u32 %0,%1,%2;" : "=r"(r) : "r"(a), "r"(b)); return r; }
MGB_DEV uint32_t mul_hi(uint32_t a, uint32_t b) { uint32_t r; asm volatile("mul.hi.u32 %0,%1,%2;" : "=r"(r) : "r"(a), "r"(b)); return r; }
MGB_DEV uint32_t mad_lo_cc(uint32_t a, uint32_t b, uint32_t c) { uint32_t r; asm volatile("mad.lo.cc.u32 %0,%1,%2,%3;" : "=r"(r) : "r"(a), "r"(b), "r"(c)); return r; }
MGB_DEV uint32_t mad_hi_cc(uint32_t a, uint32_t b, uint32_t c) { uint32_t r; asm volatile("mad.hi.cc.u32 %0,%1,%2,%3;" : "=r"(r) : "r"(a), "r"(b), "r"(c)); return r; }
MGB_DEV uint32_t madc_lo_cc(uint32_t a, uint32_t b, uint32_t c) { uint32_t r; asm volatile("madc.lo.cc.u32 %0,%1,%2,%3;" : "=r"(r) : "r"(a), "r"(b), "r"(c)); return r; }
MGB_DEV uint32_t madc_hi_cc(uint32_t a, uint32_t b, uint32_t c) { uint32_t r; asm volatile("madc.hi.cc.u32 %0,%1,%2,%3;" : "=r"(r) : "r"(a), "r"(b), "r"(c)); return r; }
MGB_DEV uint32_t madc_hi(uint32_t a, uint32_t b, uint32_t c) { uint32_t r; asm volatile("madc.hi.u32 %0,%1,%2,%3;" : "=r"(r) : "r"(a), "r"(b), "r"(c)); return r; }
MGB_DEV uint32_t add_cc(uint32_t a, uint32_t b) { uint32_t r; asm volatile("add.cc.u32 %0,%1,%2;" : "=r"(r) : "r"(a), "r"(b)); return r; }
MGB_DEV uint32_t addc_cc(uint32_t a, uint32_t b) { uint32_t r; asm volatile("addc.cc.u32 %0,%1,%2;" : "=r"(r) : "r"(a), "r"(b)); return r; }
MGB_DEV uint32_t addc(uint32_t a, uint32_t b) { uint32_t r; asm volatile("addc.u32 %0,%1,%2;" : "=r"(r) : "r"(a), "r"(b)); return r; }
MGB_DEV uint32_t sub_cc(uint32_t a, uint32_t b) { uint32_t r; asm volatile("sub.cc.u32 %0,%1,%2;" : "=r"(r) : "r"(a), "r"(b)); return r; }
MGB_DEV uint32_t subc_cc(uint32_t a, uint32_t b) { uint32_t r; asm volatile("subc.cc.u32 %0,%1,%2;" : "=r"(r) : "r"(a), "r"(b)); return r; }
MGB_DEV uint32_t subc(uint32_t a, uint32_t b) { uint32_t r; asm volatile("subc.u32 %0,%1,%2;" : "=r"(r) : "r"(a), "r"(b)); return r; }
// x + y (64 bit); `carries` is incremented by the carry out.  One carry chain of three instructions (IADD3, IADD3.X,
// IADD3.X) instead of the add + two-word compare + select the C idiom `r = x + y; carries += r < x` compiles to.
MGB_DEV unsigned long long add64_count(unsigned long long x, unsigned long long y, uint32_t& carries) {
  uint32_t lo, hi, c;
  asm("add.cc.u32 %0,%3,%5;\n\taddc.cc.u32 %1,%4,%6;\n\taddc.u32 %2,%7,0;"
      : "=r"(lo), "=r"(hi), "=r"(c)
      : "r"((uint32_t)x), "r"((uint32_t)(x >> 32)), "r"((uint32_t)y), "r"((uint32_t)(y >> 32)), "r"(carries));
  carries = c;
  return ((unsigned long long)hi << 32) | lo;
}
}  // namespace ptx
}  // namespace mgb
#endif  // MGB_HOST_EMU
