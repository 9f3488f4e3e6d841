// Descriptor and launcher of the specialised power-of-two stage kernels (fft_fast.cuh).
//
// The generic stage_kernel (fft_kernels.cu) accepts any length; the kernels behind
// launch_fast() cover the lengths the headline configurations use (64 ... 2048) with
// compile-time radix schedules, register butterflies and one shared-memory exchange
// between passes.  Both read the same planner output (P3dStage); to_fast() turns its
// segment lists into "runs" of consecutive LOGICAL rows so that the kernels can map a row
// of the transform axis to an address with one table lookup.
#pragma once
#include <cuda_runtime.h>
#include <stddef.h>
#include <stdint.h>

#include "stage.h"

#define P3D_MAXRUN (P3D_MAXSEG + 1)

namespace p3d {

// element address of logical row k of line (a,b,c), a = ta*TX + t (TX = lines per tile):
//     base + (R(k - korg) + ta*sat + t*sa + B(b) + c*sc) * sizeof(element)     for kstart <= k < kstart+len
//     R(i) = i*ps  (kw <= 1)   or   (i / kw)*psh + (i % kw)*ps  (rows blocked by kw, stage.h)
//     B(b) = b*sb  (bw <= 1)   or   (b / bw)*sbh + (b % bw)*sb
// (sat = TX*sa for a plain strided layout; the tile-blocked internal buffers give it directly)
struct FastRun {
  const void* base;
  int32_t kstart, len;
  int32_t korg, pad_;    // logical index that maps to the block's first stored row (== kstart unless a
                         // pruned segment was split: the upper part keeps the block's origin)
  int32_t kw, bw;
  int64_t ps, psh, sa, sat, sb, sbh, sc;
};

struct FastSide {
  int32_t nrun, pad_;
  FastRun run[P3D_MAXRUN];
};

struct FastStage {
  int32_t na, nb, nc;    // batch extents; CTAs tile a
  int32_t n;             // logical length of the transform axis (nx, ny, nz)
  int32_t mirror;        // 1: DCT-I -- FFT row r >= n reads logical row nfft - r;  2: DST-I (odd extension, DST instantiation)
  int32_t prefetch;      // L2 prefetch of a CTA's next tile for inputs whose row pitch is <= this many bytes (0: off)
  int32_t bord;          // tile order: this many consecutive b are innermost (the tiles that share memory lines of a
                         // gathered input run on neighbouring CTAs at the same time), else 1
  int32_t rowb;          // bytes per tile row (64 or 128) = block width of the internal layouts
  const void* tw;        // device twiddle block of this (kind, nfft), see fast_twiddle_*
  FastSide in, out;
  double scale;          // multiplies every output (SCALED instantiations only; last member: the unscaled kernels'
                         // parameter layout does not depend on it)
  int32_t variant;       // bit 0: two-pass radix-32 c2c schedule; bit 2: bulk-copy stores (fast_variant)
  int32_t sm_cap;        // > 0: the persistent grid uses at most this many SMs (pipelined tail, api.cpp); host side only
};

// true when a specialised kernel exists for this stage (kind, length, strides, alignment)
template <typename T> bool fast_supported(const P3dStage& st);
// number of T2 elements of the twiddle block and its host-side fill
// kernel variant this stage runs with (0 = default; bit 0 decides which twiddle block it needs)
template <typename T> int fast_variant(const P3dStage& st);
// re-reads the P3DFFT_B200_R32 / P3DFFT_B200_BULK switches (p3dfft_setup calls it; launches never touch the environment)
void fast_reload_switches();
template <typename T> size_t fast_twiddle_elems(int kind, int nfft, int variant = 0);
template <typename T> void fast_twiddle_fill(int kind, int nfft, void* host, int variant = 0);
// converts resolved segments (seg.base set) into runs; real_bytes = sizeof(real)
void to_fast(const P3dStage& st, FastStage& f, size_t real_bytes, int variant = 0);
template <typename T> cudaError_t launch_fast(const P3dStage& st, const FastStage& f, cudaStream_t stream);

}  // namespace p3d
