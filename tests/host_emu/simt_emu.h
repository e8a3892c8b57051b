// TEST-ONLY: a 32-lane lockstep emulator for warp-synchronous device code (shuffles, ballots).
// run_warp(f) starts 32 host threads, one per lane; every warp primitive is a rendezvous: each lane
// publishes its operand, all wait, each reads what it needs, all wait again.  That is exactly the
// semantics of the *_sync intrinsics with a full mask when no lane diverges around the call.
#pragma once
#define MGB_SIMT_EMU 1
#include <condition_variable>
#include <cstdint>
#include <functional>
#include <mutex>
#include <thread>
#include <vector>

namespace simt {
struct Barrier {
  std::mutex mu;
  std::condition_variable cv;
  int count = 0, gen = 0;
  void wait() {
    std::unique_lock<std::mutex> lk(mu);
    const int g = gen;
    if (++count == 32) { count = 0; gen++; cv.notify_all(); }
    else cv.wait(lk, [&] { return gen != g; });
  }
};
static Barrier g_bar;
static uint32_t g_slot[32];
static thread_local int t_lane = 0;

inline uint32_t exchange(uint32_t v, int src) {   // value published by lane src
  g_slot[t_lane] = v;
  g_bar.wait();
  const uint32_t r = g_slot[src];
  g_bar.wait();
  return r;
}
inline void run_warp(const std::function<void(int)>& f) {
  std::vector<std::thread> th;
  for (int l = 0; l < 32; l++) th.emplace_back([&f, l] { t_lane = l; f(l); });
  for (auto& t : th) t.join();
}
}  // namespace simt

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
inline uint32_t ballot(bool pred) {
  uint32_t r = 0;
  simt::g_slot[simt::t_lane] = pred ? 1u : 0u;
  simt::g_bar.wait();
  for (int l = 0; l < 32; l++) r |= simt::g_slot[l] << l;
  simt::g_bar.wait();
  return r;
}
}  // namespace warp
}  // namespace mgb
