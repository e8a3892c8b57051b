// TEST-ONLY: stand-ins for the CUDA built-ins engine.cuh uses, on top of the lockstep thread emulation of simt_emu.h,
// so that whole KERNELS run on the CPU: run_grid(blocks, threads, kernel) executes the blocks one after the other,
// each as `threads` host threads.  __shared__ variables become function-local statics (one block at a time).
#pragma once
#define MGB_CUDA_EMU 1
#include "simt_emu.h"
#include <algorithm>
#include <cstring>

#define __global__
#define __launch_bounds__(...)
#define __shared__ static

struct uint2 { uint32_t x, y; };
struct alignas(16) uint4 { uint32_t x, y, z, w; };
inline uint2 make_uint2(uint32_t x, uint32_t y) { return uint2{x, y}; }
inline uint4 make_uint4(uint32_t x, uint32_t y, uint32_t z, uint32_t w) { return uint4{x, y, z, w}; }
struct EmuDim { unsigned x, y, z; };
static EmuDim blockIdx = {0, 0, 0}, gridDim = {1, 1, 1}, blockDim = {32, 1, 1};   // uniform within the running block

template <class T> inline T __ldg(const T* p) { return *p; }
inline int __popc(uint32_t v) { return __builtin_popcount(v); }
inline int __ffs(int v) { return __builtin_ffs(v); }
inline int __clz(int v) { return v ? __builtin_clz((unsigned)v) : 32; }
inline void __threadfence_block() {}
inline void __threadfence_system() {}
inline uint32_t __ballot_sync(uint32_t, bool pred) { return simt::vote(pred); }
inline uint32_t __shfl_up_sync(uint32_t, uint32_t v, unsigned d) { return simt::exchange(v, simt::t_lane < (int)d ? simt::t_lane : simt::t_lane - (int)d); }
inline uint32_t __reduce_add_sync(uint32_t, uint32_t v) { uint32_t s = 0; for (int l = 0; l < 32; l++) s += simt::exchange(v, l); return s; }
inline uint32_t __reduce_max_sync(uint32_t, uint32_t v) { uint32_t s = 0; for (int l = 0; l < 32; l++) s = std::max(s, simt::exchange(v, l)); return s; }
inline uint32_t atomicAdd(uint32_t* p, uint32_t v) { return __atomic_fetch_add(p, v, __ATOMIC_SEQ_CST); }
inline uint32_t atomicOr(uint32_t* p, uint32_t v) { return __atomic_fetch_or(p, v, __ATOMIC_SEQ_CST); }
inline uint32_t atomicMax(uint32_t* p, uint32_t v) {
  uint32_t o = __atomic_load_n(p, __ATOMIC_SEQ_CST);
  while (o < v && !__atomic_compare_exchange_n(p, &o, v, false, __ATOMIC_SEQ_CST, __ATOMIC_SEQ_CST)) {}
  return o;
}
using std::max;
using std::min;
inline uint32_t min(uint32_t a, int b) { return std::min<uint32_t>(a, (uint32_t)b); }

namespace simt {
template <class K>
inline void run_grid(unsigned blocks, unsigned threads, K kernel) {
  gridDim.x = blocks; blockDim.x = threads;
  for (unsigned b = 0; b < blocks; b++) { blockIdx.x = b; run_block((int)threads, [&](int) { kernel(); }); }
}
}  // namespace simt
