// TEST CODE (CPU only): a minimal functional emulation of the CUDA execution model, enough to run the library's
// kernels (p3dfft_b200/csrc/fft_fast.cuh, fft_kernels.cu, rcopy.h) on the host: one fiber per CUDA thread of a
// CTA (emu_runtime.inc), switched at __syncthreads() and at warp rendezvous, warp shuffles / ballots through a per-warp
// scratch line, CTAs one after the other, "shared memory" = static buffers.  It checks the kernels' index arithmetic, digit reversal,
// twiddle tables and row tables without a GPU; it says nothing about performance, and memory-model questions
// (ordering across CTAs, volatile) are outside its reach.
#pragma once
#include <cuda_runtime.h>      // host-side declarations only (double2, dim3, cudaError_t ...)
#include <pthread.h>

#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <functional>
#include <thread>
#include <vector>

#ifndef __launch_bounds__
#define __launch_bounds__(...)
#endif
// (__noinline__ is defined by the translation unit that needs it, after every standard header: libstdc++ itself
//  spells __attribute__((__noinline__)))

// ---- built-in variables -----------------------------------------------------------------------------------
extern thread_local uint3 threadIdx, blockIdx;
extern thread_local dim3 blockDim, gridDim;

namespace emu {
extern unsigned long long (*warp_scratch)[32];   // [warp][lane]
// runs kernel body `fn` for a grid of CTAs of `nt` threads: CTAs one after the other, the threads of a CTA as fibers
void launch(const std::function<void()>& fn, dim3 grid, unsigned nt);
inline void launch(const std::function<void()>& fn, unsigned grid, unsigned nt) { launch(fn, dim3(grid, 1, 1), nt); }
unsigned grid_for(long long tiles);      // a few CTAs, so that every CTA walks several tiles
// canary behind the dynamic shared memory a launch asked for: a kernel that WRITES past its `smem` bytes (an illegal
// address on the GPU) is caught after the launch
inline void canary_set(unsigned char* base, size_t used, size_t total) {
  const size_t n = std::min<size_t>(total - std::min(used, total), 16384);
  memset(base + std::min(used, total), 0xA5, n);
}
inline void canary_check(const unsigned char* base, size_t used, size_t total, const char* what) {
  const size_t n = std::min<size_t>(total - std::min(used, total), 16384);
  for (size_t i = 0; i < n; i++)
    if (base[std::min(used, total) + i] != 0xA5) {
      fprintf(stderr, "emulated launch of %s wrote shared memory at byte %zu, beyond the %zu bytes it asked for\n", what, used + i, used);
      abort();
    }
}
// Stream work of the mock runtime: a launch is handed to `enqueue_hook` (emu_api.cpp: per-stream queues run at
// synchronisation points, in an order only constrained by stream order and events); without a runtime (libemu_fast.so)
// it runs on the spot.
extern const char* current_kernel;       // name of the kernel being emulated (for the fault report of emu_api.cpp)
extern void (*enqueue_hook)(cudaStream_t, std::function<void()>);
inline void enqueue(cudaStream_t s, std::function<void()> f) {
  if (enqueue_hook) enqueue_hook(s, std::move(f));
  else f();
}
void sync_cta();                         // __syncthreads()
void sync_warp();                        // rendezvous of the calling warp's live threads

// all lanes of the calling warp exchange one 64-bit word
inline unsigned long long warp_exchange(unsigned long long mine, int src_lane) {
  const unsigned w = threadIdx.x >> 5, l = threadIdx.x & 31;
  warp_scratch[w][l] = mine;
  sync_warp();
  const unsigned long long v = warp_scratch[w][src_lane & 31];
  sync_warp();
  return v;
}
}  // namespace emu

// ---- runtime templates that cuda_runtime.h only provides to nvcc ------------------------------------------------
template <class F> inline cudaError_t cudaFuncSetAttribute(F*, cudaFuncAttribute, int) { return cudaSuccess; }

// ---- intrinsics the kernels use ------------------------------------------------------------------------------
inline void __syncthreads() { emu::sync_cta(); }
template <class T> inline T __ldg(const T* p) { return *p; }
inline int __popc(unsigned x) { return __builtin_popcount(x); }
inline int __clz(int x) { return x == 0 ? 32 : __builtin_clz((unsigned)x); }
using std::max;
using std::min;

template <class T> inline T __shfl_up_sync(unsigned, T val, unsigned delta) {
  static_assert(sizeof(T) <= 8, "emulated shuffles move up to 8 bytes");
  unsigned long long bits = 0;
  memcpy(&bits, &val, sizeof(T));
  const int lane = (int)(threadIdx.x & 31), src = lane - (int)delta;
  const unsigned long long got = emu::warp_exchange(bits, src < 0 ? lane : src);
  T out;
  memcpy(&out, &got, sizeof(T));
  return src < 0 ? val : out;
}
inline unsigned __ballot_sync(unsigned, int pred) {
  unsigned m = 0;
  for (int l = 0; l < 32; l++) m |= (emu::warp_exchange(pred ? 1ull : 0ull, l) ? 1u : 0u) << l;
  return m;
}
inline double atomicAdd(double* p, double v) {
  unsigned long long* q = reinterpret_cast<unsigned long long*>(p);
  unsigned long long old = __atomic_load_n(q, __ATOMIC_RELAXED), nw;
  double o;
  do {
    memcpy(&o, &old, 8);
    const double n = o + v;
    memcpy(&nw, &n, 8);
  } while (!__atomic_compare_exchange_n(q, &old, nw, false, __ATOMIC_RELAXED, __ATOMIC_RELAXED));
  return o;
}
