// Field-layer test hooks and integer-pipe microbenchmarks (see include/montgomery_b200.h).
#include <cstdio>
#include <string>
#include "engine.cuh"
#include "../../include/montgomery_b200.h"

using namespace mgb;

namespace {

template <class P>
__global__ void __launch_bounds__(128) k_field_op(int op, uint32_t n, const uint32_t* __restrict__ a, const uint32_t* __restrict__ b, uint32_t* __restrict__ out) {
  typedef Field<P> F;
  uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  Fe<P> x = F::to_mont(ld_fe<P>(a + (size_t)i * P::N));
  Fe<P> y = F::to_mont(ld_fe<P>(b + (size_t)i * P::N));
  Fe<P> r;
  switch (op) {
    case 0: r = F::mul(x, y); break;
    case 1: r = F::add(x, y); break;
    case 2: r = F::sub(x, y); break;
    case 3: r = F::inv(x); break;
    case 4: r = F::sqr(x); break;
    case 5: r = F::inv_bgcd(x); break;
    case 6: r = F::neg(x); break;
    case 7: r = F::inv_divsteps(x); break;
    default: r = F::zero();
  }
  st_fe<P>(out + (size_t)i * P::N, F::from_mont(r));
}

// op 8: the warp-cooperative multiplication (warp.cuh), one element per 16-lane group.  The conversions in and
// out of Montgomery form are products with R^2 and with 1 and go through the same routine.
template <class P>
__global__ void __launch_bounds__(128) k_field_op_warp(uint32_t n, const uint32_t* __restrict__ a, const uint32_t* __restrict__ b, uint32_t* __restrict__ out) {
  typedef WarpField<P> WF;
  const uint32_t e = (blockIdx.x * blockDim.x + threadIdx.x) >> 4;     // element of this 16-lane group
  const int l = threadIdx.x & 15;
  const bool live = e < n && l < P::N;                                   // every lane runs the shuffles, dead ones on zeros
  uint32_t r2 = 0;
  _Pragma("unroll") for (int k = 0; k < P::N; k++) r2 = (l == k) ? P::r2(k) : r2;
  const uint32_t x = WF::mul(live ? a[(size_t)e * P::N + l] : 0u, r2);
  const uint32_t y = WF::mul(live ? b[(size_t)e * P::N + l] : 0u, r2);
  const uint32_t r = WF::mul(WF::mul(x, y), l == 0 ? 1u : 0u);
  if (live) out[(size_t)e * P::N + l] = r;
}

// op 9: the lane-parallel division-step inverse (warp.cuh), one element per WARP (every lane passes the same value)
template <class P>
__global__ void __launch_bounds__(128) k_field_op_warp_inv(uint32_t n, const uint32_t* __restrict__ a, uint32_t* __restrict__ out) {
  typedef Field<P> F;
  const uint32_t e = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  Fe<P> x = F::zero();
  if (e < n) x = F::to_mont(ld_fe<P>(a + (size_t)e * P::N));       // warp-uniform
  const Fe<P> r = WarpField<P>::inv_call(x);
  if (e < n && (threadIdx.x & 31) == (e & 31)) st_fe<P>(out + (size_t)e * P::N, F::from_mont(r));   // any lane holds it
}

// latency of a dependent chain of products on one warp: lane 0 alone (Field::mul) or the lanes together (WarpField::mul)
template <class P, bool COOP>
__global__ void __launch_bounds__(32) k_mullat(uint32_t* out, uint32_t seed, int iters) {
  typedef Field<P> F;
  const int l = threadIdx.x & 31;
  Fe<P> a = F::one(), b = F::one();
  a.v[0] ^= seed & 0xffff;
  b.v[1] ^= blockIdx.x & 0xffff;
  uint32_t s = 0;
  if (COOP) {
    uint32_t x = 0, y = 0;
    _Pragma("unroll") for (int k = 0; k < P::N; k++) { x = (l == k) ? a.v[k] : x; y = (l == k) ? b.v[k] : y; }
#pragma unroll 1
    for (int it = 0; it < iters; it++) { x = WarpField<P>::mul(x, y); y = WarpField<P>::mul(y, x); }
    s = x ^ y;
  } else if (l == 0) {
#pragma unroll 1
    for (int it = 0; it < iters; it++) { a = F::mul(a, b); b = F::mul(b, a); }
    _Pragma("unroll") for (int k = 0; k < P::N; k++) s ^= a.v[k] ^ b.v[k];
  }
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

// ---- integer pipe microbenchmarks: 8 independent dependency chains per thread
template <int MODE>
__global__ void __launch_bounds__(1024) k_imad(uint32_t* out, uint32_t seed, int iters) {
  uint32_t a = seed ^ threadIdx.x, b = seed * 2654435761u + blockIdx.x;
  if (MODE <= 1) {
    uint32_t x[8];
#pragma unroll
    for (int k = 0; k < 8; k++) x[k] = a + k;
#pragma unroll 1
    for (int it = 0; it < iters; it++) {
#pragma unroll
      for (int u = 0; u < 16; u++) {
#pragma unroll
        for (int k = 0; k < 8; k++) {
          if (MODE == 0) asm volatile("mad.lo.u32 %0,%0,%1,%2;" : "+r"(x[k]) : "r"(a), "r"(b));
          else asm volatile("mad.hi.u32 %0,%0,%1,%2;" : "+r"(x[k]) : "r"(a), "r"(b));
        }
      }
    }
    uint32_t s = 0;
#pragma unroll
    for (int k = 0; k < 8; k++) s ^= x[k];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
  } else if (MODE == 2) {
    // 32x32+64->64 with the multiplicand taken from the accumulator itself, so ptxas cannot hoist
    // the product out of the loop (a loop-invariant product turns into plain 64-bit adds).
    uint64_t x[8];
#pragma unroll
    for (int k = 0; k < 8; k++) x[k] = a + k;
#pragma unroll 1
    for (int it = 0; it < iters; it++) {
#pragma unroll
      for (int u = 0; u < 16; u++) {
#pragma unroll
        for (int k = 0; k < 8; k++) { uint32_t m = (uint32_t)x[k]; asm volatile("mad.wide.u32 %0,%1,%2,%0;" : "+l"(x[k]) : "r"(m), "r"(b)); }
      }
    }
    uint64_t s = 0;
#pragma unroll
    for (int k = 0; k < 8; k++) s ^= x[k];
    out[blockIdx.x * blockDim.x + threadIdx.x] = (uint32_t)(s ^ (s >> 32));
  } else {
    // carry chain: 8 aligned pairs, one IMAD.WIDE.U32.X each, carry rippling through; the
    // multiplicand of pair k is the (changing) low word of pair k+1
    uint32_t x[16];
#pragma unroll
    for (int k = 0; k < 16; k++) x[k] = a + k;
#pragma unroll 1
    for (int it = 0; it < iters; it++) {
#pragma unroll
      for (int u = 0; u < 16; u++) {
        x[0] = ptx::mad_lo_cc(x[2], b, x[0]);
        x[1] = ptx::madc_hi_cc(x[2], b, x[1]);
#pragma unroll
        for (int k = 2; k < 16; k += 2) {
          x[k] = ptx::madc_lo_cc(x[(k + 2) & 15], b, x[k]);
          x[k + 1] = ptx::madc_hi_cc(x[(k + 2) & 15], b, x[k + 1]);
        }
      }
    }
    uint32_t s = 0;
#pragma unroll
    for (int k = 0; k < 16; k++) s ^= x[k];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
  }
}

template <class P, bool INL>
__global__ void __launch_bounds__(1024) k_mulbench(uint32_t* out, uint32_t seed, int iters) {
  typedef Field<P> F;
  Fe<P> a = F::one(), b = F::one();
  a.v[0] ^= (seed ^ threadIdx.x) & 0xffff;
  b.v[1] ^= blockIdx.x & 0xffff;
#pragma unroll 1
  for (int it = 0; it < iters; it++) {
    if (INL) { a = F::mul_inl(a, b); b = F::mul_inl(b, a); }
    else { a = F::mul(a, b); b = F::mul(b, a); }
  }
  uint32_t s = 0;
#pragma unroll
  for (int k = 0; k < P::N; k++) s ^= a.v[k] ^ b.v[k];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

// two independent products per call, instruction streams free to interleave (one basic block): does a warp that runs
// alone on its scheduler keep the multiplier pipe busier with two carry-chain sets in flight than with one?
template <class P>
struct Fe2 { Fe<P> a, b; };
template <class P>
MGB_NOINLINE_DEV Fe2<P> mul2_call(Fe<P> a, Fe<P> b, Fe<P> c, Fe<P> d) {
  Fe2<P> r;
  r.a = Field<P>::mul_inl(a, b);
  r.b = Field<P>::mul_inl(c, d);
  return r;
}
template <class P>
__global__ void __launch_bounds__(256) k_mul2bench(uint32_t* out, uint32_t seed, int iters) {
  typedef Field<P> F;
  Fe<P> a = F::one(), b = F::one(), c = F::one(), d = F::one();
  a.v[0] ^= (seed ^ threadIdx.x) & 0xffff;
  b.v[1] ^= blockIdx.x & 0xffff;
  c.v[2] ^= (seed ^ threadIdx.x) & 0xfff;
  d.v[3] ^= blockIdx.x & 0xfff;
#pragma unroll 1
  for (int it = 0; it < iters; it++) {
    Fe2<P> r = mul2_call<P>(a, b, c, d);
    a = r.a; c = r.b;
    r = mul2_call<P>(b, a, d, c);
    b = r.a; d = r.b;
  }
  uint32_t s = 0;
#pragma unroll
  for (int k = 0; k < P::N; k++) s ^= a.v[k] ^ b.v[k] ^ c.v[k] ^ d.v[k];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

// inversion latency: dependent chain of division-step inverses, all lanes or lane 0 only
template <class P, bool LANE0, bool COOP = false>
__global__ void __launch_bounds__(128) k_invbench(uint32_t* out, uint32_t seed, int iters) {
  typedef Field<P> F;
  Fe<P> a = F::one();
  a.v[0] ^= (seed ^ (COOP ? threadIdx.x >> 5 : threadIdx.x) ^ (blockIdx.x << 8)) & 0xffffff;
  if (COOP) {                                            // every lane of a warp holds the same element
#pragma unroll 1
    for (int it = 0; it < iters; it++) { a = WarpField<P>::inv_call(a); a.v[0] ^= 5; }
  } else if (!LANE0 || (threadIdx.x & 31) == 0) {
#pragma unroll 1
    for (int it = 0; it < iters; it++) { a = F::inv_divsteps(a); a.v[0] ^= 5; }
  }
  uint32_t s = 0;
#pragma unroll
  for (int k = 0; k < P::N; k++) s ^= a.v[k];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

thread_local std::string g_err;
#define CUT(call) do { cudaError_t e_ = (call); if (e_ != cudaSuccess) { g_err = std::string(#call) + ": " + cudaGetErrorString(e_); return MGB_E_CUDA; } } while (0)

}  // namespace

extern "C" {

int mgb_field_op(int device, int field, int op, const uint8_t* a, const uint8_t* b, uint8_t* out, size_t n) {
  if (!a || !b || !out || field < 0 || field > 3) return MGB_E_INVALID;
  CUT(cudaSetDevice(device));
  const int N = (field == 0 || field == 3) ? 12 : 8;
  size_t bytes = n * N * 4;
  uint32_t *da, *db, *dout;
  CUT(cudaMalloc(&da, bytes)); CUT(cudaMalloc(&db, bytes)); CUT(cudaMalloc(&dout, bytes));
  CUT(cudaMemcpy(da, a, bytes, cudaMemcpyHostToDevice));
  CUT(cudaMemcpy(db, b, bytes, cudaMemcpyHostToDevice));
  unsigned grid = (unsigned)((n + 127) / 128);
  if (op == 9) {
    grid = (unsigned)((n * 32 + 127) / 128);
    if (field == 0) k_field_op_warp_inv<Fp377><<<grid, 128>>>((uint32_t)n, da, dout);
    else if (field == 1) k_field_op_warp_inv<Fr377><<<grid, 128>>>((uint32_t)n, da, dout);
    else if (field == 2) k_field_op_warp_inv<FpPallas><<<grid, 128>>>((uint32_t)n, da, dout);
    else k_field_op_warp_inv<Fp381><<<grid, 128>>>((uint32_t)n, da, dout);
  }
  else if (op == 8) {
    grid = (unsigned)((n * 16 + 127) / 128);
    if (field == 0) k_field_op_warp<Fp377><<<grid, 128>>>((uint32_t)n, da, db, dout);
    else if (field == 1) k_field_op_warp<Fr377><<<grid, 128>>>((uint32_t)n, da, db, dout);
    else if (field == 2) k_field_op_warp<FpPallas><<<grid, 128>>>((uint32_t)n, da, db, dout);
    else k_field_op_warp<Fp381><<<grid, 128>>>((uint32_t)n, da, db, dout);
  }
  else if (field == 0) k_field_op<Fp377><<<grid, 128>>>(op, (uint32_t)n, da, db, dout);
  else if (field == 1) k_field_op<Fr377><<<grid, 128>>>(op, (uint32_t)n, da, db, dout);
  else if (field == 2) k_field_op<FpPallas><<<grid, 128>>>(op, (uint32_t)n, da, db, dout);
  else k_field_op<Fp381><<<grid, 128>>>(op, (uint32_t)n, da, db, dout);
  CUT(cudaGetLastError());
  CUT(cudaMemcpy(out, dout, bytes, cudaMemcpyDeviceToHost));
  cudaFree(da); cudaFree(db); cudaFree(dout);
  return 0;
}

int mgb_microbench(int device, int mode, int blocks_per_sm, int threads, int iters, double* ops_per_s, float* ms_out) {
  if (!ops_per_s || threads < 32 || threads > 1024 || blocks_per_sm < 1 || iters < 1) return MGB_E_INVALID;
  CUT(cudaSetDevice(device));
  cudaDeviceProp prop;
  CUT(cudaGetDeviceProperties(&prop, device));
  int grid = prop.multiProcessorCount * blocks_per_sm;
  uint32_t* d;
  CUT(cudaMalloc(&d, (size_t)grid * threads * 4));
  cudaEvent_t e0, e1;
  CUT(cudaEventCreate(&e0)); CUT(cudaEventCreate(&e1));
  auto launch = [&](int it) {
    switch (mode) {
      case 0: k_imad<0><<<grid, threads>>>(d, 12345u, it); break;
      case 1: k_imad<1><<<grid, threads>>>(d, 12345u, it); break;
      case 2: k_imad<2><<<grid, threads>>>(d, 12345u, it); break;
      case 3: k_imad<3><<<grid, threads>>>(d, 12345u, it); break;
      case 4: k_mulbench<Fp377, false><<<grid, threads>>>(d, 12345u, it); break;
      case 5: k_mulbench<Fr377, false><<<grid, threads>>>(d, 12345u, it); break;
      case 6: k_mulbench<Fp377, true><<<grid, threads>>>(d, 12345u, it); break;
      case 7: k_mulbench<Fr377, true><<<grid, threads>>>(d, 12345u, it); break;
      case 8: k_invbench<Fp377, false><<<grid, threads>>>(d, 12345u, it); break;
      case 9: k_invbench<Fp377, true><<<grid, threads>>>(d, 12345u, it); break;
      case 10: k_mullat<Fp377, false><<<grid, 32>>>(d, 12345u, it); break;
      case 11: k_mullat<Fp377, true><<<grid, 32>>>(d, 12345u, it); break;
      case 12: k_invbench<Fp377, true, true><<<grid, threads>>>(d, 12345u, it); break;
      case 13: k_mul2bench<Fp377><<<grid, threads>>>(d, 12345u, it); break;
      default: break;
    }
  };
  if (mode < 0 || mode > 13) return MGB_E_INVALID;
  if (mode >= 8 && mode <= 12 && threads > 128) return MGB_E_INVALID;
  if (mode == 10 || mode == 11) threads = 32;                         // one warp per block: the chain runs alone on its scheduler
  launch(iters / 8 + 1);  // warm-up
  CUT(cudaDeviceSynchronize());
  CUT(cudaEventRecord(e0));
  launch(iters);
  CUT(cudaEventRecord(e1));
  CUT(cudaEventSynchronize(e1));
  CUT(cudaGetLastError());
  float ms = 0;
  CUT(cudaEventElapsedTime(&ms, e0, e1));
  double per_thread;
  if (mode <= 2) per_thread = 16.0 * 8 * iters;        // instructions per thread
  else if (mode == 3) per_thread = 16.0 * 8 * iters;   // wide MADs per thread
  else if (mode == 13) per_thread = 4.0 * iters;        // field multiplications per thread (two per call)
  else if (mode == 12) per_thread = 1.0 / 32 * iters;   // inversions per WARP
  else if (mode >= 10) per_thread = 2.0 * iters / 32;   // products per WARP (one chain per warp)
  else if (mode >= 8) per_thread = (mode == 9 ? 1.0 / 32 : 1.0) * iters;   // inversions per thread
  else per_thread = 2.0 * iters;                        // field multiplications per thread
  *ops_per_s = per_thread * (double)grid * threads / (ms * 1e-3);
  if (ms_out) *ms_out = ms;
  cudaFree(d);
  cudaEventDestroy(e0); cudaEventDestroy(e1);
  return 0;
}

}  // extern "C"
