// TEST-ONLY: a stand-in for the CUDA runtime, so that the HOST side of the engine (montgomery_b200/csrc/msm.cu: the C ABI,
// buffer management, round planning, every kernel launch) compiles with g++ and runs on the CPU on top of the SIMT
// emulation of cuda_emu.h.  "Device" memory is host memory, streams are synchronous, a kernel launch runs the grid block
// after block (emu_launch).  tests/host_emu/make_emu_host.py rewrites the <<< >>> launches of msm.cu into emu_launch calls;
// tests/test_host_emu_pipeline.py builds the result and checks whole MSMs through the C ABI against the oracle.
// Never part of the shipped library; the product has no CPU path.
#pragma once
#include "cuda_emu.h"
#include <chrono>
#include <cstdlib>
#include <mutex>

typedef int cudaError_t;
enum { cudaSuccess = 0, cudaErrorMemoryAllocation = 2 };
typedef void* cudaStream_t;
struct EmuEvent { double ms = 0; };
typedef EmuEvent* cudaEvent_t;
enum cudaMemcpyKind { cudaMemcpyHostToHost, cudaMemcpyHostToDevice, cudaMemcpyDeviceToHost, cudaMemcpyDeviceToDevice };
enum { cudaStreamNonBlocking = 1, cudaEventDisableTiming = 2, cudaFuncAttributePreferredSharedMemoryCarveout = 9 };
struct cudaDeviceProp { int multiProcessorCount; };
struct dim3 { unsigned x, y, z; dim3(unsigned x_ = 1, unsigned y_ = 1, unsigned z_ = 1) : x(x_), y(y_), z(z_) {} };

#ifndef MGB_EMU_SM_COUNT
#define MGB_EMU_SM_COUNT 2        // "SMs" of the emulated device: persistent grids are sm_count x blocks-per-SM
#endif
inline cudaError_t cudaSetDevice(int) { return cudaSuccess; }
inline cudaError_t cudaGetDeviceCount(int* n) { *n = 4; return cudaSuccess; }   // four "devices" (all the same host memory) for the mgb_multi_* tests
inline cudaError_t cudaGetDeviceProperties(cudaDeviceProp* p, int) { p->multiProcessorCount = MGB_EMU_SM_COUNT; return cudaSuccess; }
inline cudaError_t cudaGetLastError() { return cudaSuccess; }
inline const char* cudaGetErrorString(cudaError_t) { return "emulated runtime"; }
inline cudaError_t cudaMalloc(void** p, size_t n) { *p = std::aligned_alloc(256, (n + 255) / 256 * 256); return *p ? cudaSuccess : cudaErrorMemoryAllocation; }
template <class T> inline cudaError_t cudaMalloc(T** p, size_t n) { return cudaMalloc((void**)p, n); }
inline cudaError_t cudaFree(void* p) { std::free(p); return cudaSuccess; }
inline cudaError_t cudaMallocHost(void** p, size_t n) { return cudaMalloc(p, n); }
inline cudaError_t cudaFreeHost(void* p) { std::free(p); return cudaSuccess; }
inline cudaError_t cudaMemcpy(void* d, const void* s, size_t n, cudaMemcpyKind) { std::memcpy(d, s, n); return cudaSuccess; }
inline cudaError_t cudaMemcpyAsync(void* d, const void* s, size_t n, cudaMemcpyKind, cudaStream_t) { std::memcpy(d, s, n); return cudaSuccess; }
inline cudaError_t cudaMemsetAsync(void* d, int v, size_t n, cudaStream_t) { std::memset(d, v, n); return cudaSuccess; }
inline cudaError_t cudaStreamCreateWithFlags(cudaStream_t* s, unsigned) { *s = (void*)1; return cudaSuccess; }
inline cudaError_t cudaStreamDestroy(cudaStream_t) { return cudaSuccess; }
inline cudaError_t cudaStreamSynchronize(cudaStream_t) { return cudaSuccess; }
inline cudaError_t cudaStreamWaitEvent(cudaStream_t, cudaEvent_t, unsigned) { return cudaSuccess; }
inline cudaError_t cudaEventCreate(cudaEvent_t* e) { *e = new EmuEvent; return cudaSuccess; }
inline cudaError_t cudaEventCreateWithFlags(cudaEvent_t* e, unsigned) { return cudaEventCreate(e); }
inline cudaError_t cudaEventDestroy(cudaEvent_t e) { delete e; return cudaSuccess; }
inline cudaError_t cudaEventRecord(cudaEvent_t e, cudaStream_t) {
  e->ms = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now().time_since_epoch()).count();
  return cudaSuccess;
}
inline cudaError_t cudaEventSynchronize(cudaEvent_t) { return cudaSuccess; }
inline cudaError_t cudaEventElapsedTime(float* ms, cudaEvent_t a, cudaEvent_t b) { *ms = (float)(b->ms - a->ms); return cudaSuccess; }
template <class K> inline cudaError_t cudaFuncSetAttribute(K, int, int) { return cudaSuccess; }

// the NCCL types msm.cu names (the library itself is bound with dlopen at run time; the tests hand it tests/host_emu/fake_nccl.cpp)
typedef int ncclResult_t;
enum { ncclSuccess = 0 };
typedef struct ncclComm* ncclComm_t;
struct ncclUniqueId { char internal[128]; };
typedef int ncclDataType_t;
enum { ncclUint8 = 1 };

// kernel launch: the blocks of a (one- or two-dimensional) grid one after the other, each as `threads` lockstep host threads.
// One launch at a time in the process: the block state of the emulation (blockIdx, the barriers, __shared__ statics) is
// global, and mgb_multi_* / the sharded tests drive several contexts from several host threads.
inline std::mutex& emu_launch_mutex() { static std::mutex mu; return mu; }   // (not a static of the template below: one per process)
template <class F>
inline void emu_launch(dim3 grid, unsigned threads, F body) {
  std::lock_guard<std::mutex> lk(emu_launch_mutex());
  gridDim.x = grid.x; gridDim.y = grid.y; blockDim.x = threads;
  for (unsigned by = 0; by < grid.y; by++)
    for (unsigned bx = 0; bx < grid.x; bx++) {
      blockIdx.x = bx; blockIdx.y = by;
      simt::run_block((int)threads, [&](int) { body(); });
    }
  gridDim.y = 1; blockIdx.y = 0;
}
