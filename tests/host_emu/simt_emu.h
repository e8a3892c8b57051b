// TEST-ONLY: a lockstep emulator for warp- and block-synchronous device code (shuffles, ballots, __syncwarp,
// __syncthreads, shared memory).  run_block(n, f) starts n host threads, one per CUDA thread of a block of n / 32
// warps; every warp primitive is a rendezvous of the 32 threads of a warp: each lane publishes its operand, all
// wait, each reads what it needs, all wait again -- exactly the semantics of the *_sync intrinsics with a full mask
// when no lane diverges around the call.  __syncthreads is a rendezvous of the whole block.  Shared memory is any
// buffer the threads share.  run_warp(f) = one warp.
#pragma once
#define MGB_SIMT_EMU 1
#include <atomic>
#include <cstdint>
#include <functional>
#include <thread>
#include <vector>

namespace simt {
struct Barrier {   // generation barrier on atomics: the threads outnumber the cores, so waiters yield instead of sleeping
  std::atomic<int> count{0}, gen{0};
  int need = 32;
  void wait() {
    const int g = gen.load(std::memory_order_acquire);
    if (count.fetch_add(1, std::memory_order_acq_rel) + 1 == need) {
      count.store(0, std::memory_order_relaxed);
      gen.store(g + 1, std::memory_order_release);
    } else {
      while (gen.load(std::memory_order_acquire) == g) std::this_thread::yield();
    }
  }
};
constexpr int MAX_WARPS = 32;   // blocks of up to 1024 threads (the scan kernels)
static Barrier g_warp_bar[MAX_WARPS];
static Barrier g_block_bar;
static uint32_t g_slot[MAX_WARPS][32];
static thread_local int t_lane = 0, t_warp = 0;

inline uint32_t exchange(uint32_t v, int src) {   // value published by lane src of the calling thread's warp
  g_slot[t_warp][t_lane] = v;
  g_warp_bar[t_warp].wait();
  const uint32_t r = g_slot[t_warp][src];
  g_warp_bar[t_warp].wait();
  return r;
}
inline uint32_t vote(bool pred) {
  uint32_t r = 0;
  g_slot[t_warp][t_lane] = pred ? 1u : 0u;
  g_warp_bar[t_warp].wait();
  for (int l = 0; l < 32; l++) r |= g_slot[t_warp][l] << l;
  g_warp_bar[t_warp].wait();
  return r;
}
struct Idx { unsigned x; };
static thread_local Idx t_idx = {0};
inline void run_block(int nthreads, const std::function<void(int)>& f) {
  g_block_bar.need = nthreads;
  std::vector<std::thread> th;
  for (int t = 0; t < nthreads; t++)
    th.emplace_back([&f, t] { t_lane = t & 31; t_warp = t >> 5; t_idx.x = (unsigned)t; f(t); });
  for (auto& t : th) t.join();
}
inline void run_warp(const std::function<void(int)>& f) { run_block(32, f); }
}  // namespace simt

// the CUDA names the cooperative headers use directly
#define threadIdx simt::t_idx
inline void __syncthreads() { simt::g_block_bar.wait(); }
inline void __syncwarp() { simt::g_warp_bar[simt::t_warp].wait(); }
inline uint32_t __shfl_sync(uint32_t, uint32_t v, int src) { return simt::exchange(v, src & 31); }
inline int __shfl_sync(uint32_t, int v, int src) { return (int)simt::exchange((uint32_t)v, src & 31); }
inline bool __any_sync(uint32_t, bool pred) { return simt::vote(pred) != 0; }

namespace mgb {
namespace warp {
inline int lane() { return simt::t_lane; }
inline uint32_t shfl(uint32_t v, int src, int width) {
  const int base = simt::t_lane & ~(width - 1);
  return simt::exchange(v, base | (src & (width - 1)));
}
inline uint32_t shfl_up(uint32_t v, int delta, int width) {
  const int sub = simt::t_lane & (width - 1);
  return simt::exchange(v, sub - delta < 0 ? simt::t_lane : simt::t_lane - delta);
}
inline uint32_t shfl_down(uint32_t v, int delta, int width) {
  const int sub = simt::t_lane & (width - 1);
  return simt::exchange(v, sub + delta >= width ? simt::t_lane : simt::t_lane + delta);
}
inline uint32_t ballot(bool pred) { return simt::vote(pred); }
}  // namespace warp
}  // namespace mgb
