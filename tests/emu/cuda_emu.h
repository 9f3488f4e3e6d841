// TEST CODE (CPU only): a minimal functional emulation of the CUDA execution model, enough to run the library's
// stage kernels (p3dfft_b200/csrc/fft_fast.cuh) on the host: one OS thread per CUDA thread of a CTA, a pthread
// barrier for __syncthreads(), CTAs one after the other, "shared memory" = one static buffer.  It checks the
// kernels' index arithmetic, digit reversal, twiddle tables and row tables without a GPU; it says nothing about
// performance, and memory-model questions (ordering across CTAs, volatile) are outside its reach.
#pragma once
#include <cuda_runtime.h>      // host-side declarations only (double2, dim3, cudaError_t ...)
#include <pthread.h>

#include <cstdint>
#include <functional>
#include <thread>
#include <vector>

#ifndef __launch_bounds__
#define __launch_bounds__(...)
#endif

// ---- built-in variables -----------------------------------------------------------------------------------
extern thread_local uint3 threadIdx, blockIdx;
extern thread_local dim3 blockDim, gridDim;

namespace emu {
extern pthread_barrier_t* cta_barrier;
// runs kernel body `fn` for a grid of `grid` CTAs of `nt` threads, CTAs sequentially, threads concurrently
void launch(const std::function<void()>& fn, unsigned grid, unsigned nt);
unsigned grid_for(long long tiles);      // a few CTAs, so that every CTA walks several tiles
}  // namespace emu

// ---- intrinsics the kernels use ------------------------------------------------------------------------------
inline void __syncthreads() { pthread_barrier_wait(emu::cta_barrier); }
template <class T> inline T __ldg(const T* p) { return *p; }
inline int __popc(unsigned x) { return __builtin_popcount(x); }
inline int __clz(int x) { return x == 0 ? 32 : __builtin_clz((unsigned)x); }
