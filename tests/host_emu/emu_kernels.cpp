// TEST-ONLY: whole kernels of the engine on the CPU.  engine.cuh is compiled for the host against the CUDA stand-ins of
// cuda_emu.h; a launch becomes simt::run_grid (blocks one after the other, each as 128 lockstep host threads with real
// shuffles / ballots / barriers / atomics).  Not shipped, not linked into the product library.
#define MGB_HOST_EMU 1
#include "cuda_emu.h"
#include "../../montgomery_b200/csrc/engine.cuh"
using namespace mgb;

// k_batch_add<curve, EMAX = 8, MINB = 4, FIRST>: one round of batched-affine additions (see engine.cuh).  All buffers are
// the caller's; `blocks` must be a multiple of 4 (the static first tile of a warp assumes MINB blocks per SM).
template <class CV, bool FIRST>
static void batch_add(uint32_t* V, const PairEnt* pairs, const uint32_t* npairs_ptr, int r, int E_big, uint32_t n_big, PairEnt* pairs_out,
                      uint32_t* npairs_out, uint32_t* tile_counter, const uint2* recs, const uint8_t* lifes, const uint32_t* table,
                      const uint32_t* offs, uint32_t b_begin, uint32_t b_end, uint4* scratch, int blocks) {
  simt::run_grid((unsigned)blocks, 128, [&] {
    k_batch_add<CV, 8, 4, FIRST>(V, pairs, npairs_ptr, r, E_big, E_big >= 16 ? E_big / 4 : E_big, n_big, pairs_out, npairs_out, tile_counter, recs, lifes, table, offs, b_begin, b_end, scratch);
  });
}

extern "C" void emu_batch_add(int curve, int first, uint32_t* V, const uint32_t* pairs, const uint32_t* npairs_ptr, int r, int E_big, uint32_t n_big,
                              uint32_t* pairs_out, uint32_t* npairs_out, uint32_t* tile_counter, const uint32_t* recs, const uint8_t* lifes,
                              const uint32_t* table, const uint32_t* offs, uint32_t b_begin, uint32_t b_end, uint32_t* scratch, int blocks) {
  auto P = reinterpret_cast<const PairEnt*>(pairs);
  auto PO = reinterpret_cast<PairEnt*>(pairs_out);
  auto RC = reinterpret_cast<const uint2*>(recs);
  auto SC = reinterpret_cast<uint4*>(scratch);
#define GO(CV) (first ? batch_add<CV, true>(V, P, npairs_ptr, r, E_big, n_big, PO, npairs_out, tile_counter, RC, lifes, table, offs, b_begin, b_end, SC, blocks) \
                      : batch_add<CV, false>(V, P, npairs_ptr, r, E_big, n_big, PO, npairs_out, tile_counter, RC, lifes, table, offs, b_begin, b_end, SC, blocks))
  if (curve == 0) GO(CurveBls377);
  else if (curve == 1) GO(CurvePallas);
  else GO(CurveBls381);
#undef GO
}

// k_final: result = sum_w 2^(c w) S_w by Horner in one warp
// out_xy / out_flag (2 N + 1 words, may be NULL): the fused normalisation of the single-GPU path
extern "C" void emu_final(int curve, int K, int c, const uint32_t* Sw, uint32_t* out_acc, uint32_t* out_xy, uint32_t* out_flag) {
  simt::run_grid(1, 32, [&] {
    if (curve == 0) k_final<CurveBls377>(K, c, Sw, out_acc, out_xy, out_flag);
    else if (curve == 1) k_final<CurvePallas>(K, c, Sw, out_acc, out_xy, out_flag);
    else if (curve == 2) k_final<CurveBls381>(K, c, Sw, out_acc, out_xy, out_flag);
    else k_final<CurveEd377>(K, c, Sw, out_acc, out_xy, out_flag);
  });
}

// k_pair_add<CurveEd377, FIRST>: one tree round of the twisted-Edwards accumulation (unified additions, no inversion)
extern "C" void emu_pair_add_te(int first, uint32_t* V, const uint32_t* pairs, const uint32_t* npairs_ptr, int r, uint32_t* pairs_out,
                                uint32_t* npairs_out, const uint32_t* recs, const uint8_t* lifes, const uint32_t* table, const uint32_t* offs,
                                uint32_t b_begin, uint32_t b_end, int blocks) {
  auto P = reinterpret_cast<const PairEnt*>(pairs);
  auto PO = reinterpret_cast<PairEnt*>(pairs_out);
  auto RC = reinterpret_cast<const uint2*>(recs);
  simt::run_grid((unsigned)blocks, 128, [&] {
    if (first) k_pair_add<CurveEd377, true>(V, P, npairs_ptr, r, PO, npairs_out, RC, lifes, table, offs, b_begin, b_end);
    else k_pair_add<CurveEd377, false>(V, P, npairs_ptr, r, PO, npairs_out, RC, lifes, table, offs, b_begin, b_end);
  });
}

// byte ingestion and read-back (k_set_points / k_get_points) and the combine + normalise kernel (k_normalize: sums `count`
// partial accumulators -- the multi-GPU combine -- and returns the canonical affine point)
template <class CV> static void set_get(uint32_t n, const uint32_t* xy, const uint8_t* is_zero, uint32_t* table, uint32_t* xy_back, uint8_t* zero_back) {
  const unsigned blocks = (n + 127) / 128;
  simt::run_grid(blocks, 128, [&] { k_set_points<CV>(n, xy, is_zero, table); });
  simt::run_grid(blocks, 128, [&] { k_get_points<CV>(0, n, table, xy_back, zero_back); });
}
extern "C" void emu_set_get_points(int curve, uint32_t n, const uint32_t* xy, const uint8_t* is_zero, uint32_t* table, uint32_t* xy_back, uint8_t* zero_back) {
  if (curve == 0) set_get<CurveBls377>(n, xy, is_zero, table, xy_back, zero_back);
  else if (curve == 1) set_get<CurvePallas>(n, xy, is_zero, table, xy_back, zero_back);
  else if (curve == 2) set_get<CurveBls381>(n, xy, is_zero, table, xy_back, zero_back);
  else set_get<CurveEd377>(n, xy, is_zero, table, xy_back, zero_back);
}
extern "C" void emu_normalize(int curve, const uint32_t* accs, int count, int stride, uint32_t* out_xy, uint32_t* out_flag /* 2 words */) {
  simt::run_grid(1, 32, [&] {
    if (curve == 0) k_normalize<CurveBls377>(accs, count, stride, out_xy, out_flag);
    else if (curve == 1) k_normalize<CurvePallas>(accs, count, stride, out_xy, out_flag);
    else if (curve == 2) k_normalize<CurveBls381>(accs, count, stride, out_xy, out_flag);
    else k_normalize<CurveEd377>(accs, count, stride, out_xy, out_flag);
  });
}
