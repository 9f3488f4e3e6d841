// Batched 1D transform stage kernels for sm_100a.
//
// stage_kernel<T,KIND>: one CTA transforms `tile` lines.  Each input element is read once
// from HBM (gathered through the segment list, so the unpack of the preceding all-to-all and
// the zero-padding of pruned modes are free), the lines are transformed in shared memory by
// an in-place decimation-in-frequency mixed-radix FFT, and each output element is written
// once (scattered through the output segment list, so pruning and the pack for the next
// all-to-all are free).  Replaces FFTW's execute calls in build/fft_exec.F90 together with
// the pack/unpack loops of build/fcomm1.F90, fcomm2.F90, bcomm1.F90, bcomm2.F90 and the
// seg_copy_*/seg_zero_* helpers of build/module.F90:427-706.
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#include "stage.h"
#include "kernels.h"
#include "rcopy.h"

// every kernel launch of this file goes through this macro (tests/emu supplies a host emulation of it)
#ifndef P3D_KLAUNCH
#define P3D_KLAUNCH(kernel, grid, block, smem, stream, ...) kernel<<<grid, block, smem, stream>>>(__VA_ARGS__)
#endif

namespace p3d {

template <typename T> struct Cx;
template <> struct Cx<double> { using type = double2; };
template <> struct Cx<float>  { using type = float2; };

template <typename T2> __device__ __forceinline__ T2 cadd(T2 a, T2 b) { return {a.x + b.x, a.y + b.y}; }
template <typename T2> __device__ __forceinline__ T2 csub(T2 a, T2 b) { return {a.x - b.x, a.y - b.y}; }
template <typename T2> __device__ __forceinline__ T2 cmul(T2 a, T2 b) {
  return {a.x * b.x - a.y * b.y, a.x * b.y + a.y * b.x};
}
// multiply by -i
template <typename T2> __device__ __forceinline__ T2 mul_mi(T2 a) { return {a.y, -a.x}; }

// ---- forward DFT butterflies (sign -), natural order in and out --------------------
template <typename T2> __device__ __forceinline__ void bfly2(T2& a, T2& b) {
  T2 t = csub(a, b); a = cadd(a, b); b = t;
}
template <typename T2> __device__ __forceinline__ void bfly4(T2& v0, T2& v1, T2& v2, T2& v3) {
  T2 a = cadd(v0, v2), b = csub(v0, v2), c = cadd(v1, v3), d = mul_mi(csub(v1, v3));
  v0 = cadd(a, c); v2 = csub(a, c); v1 = cadd(b, d); v3 = csub(b, d);
}
template <typename T, typename T2> __device__ __forceinline__ void bfly8(T2* v) {
  const T h = (T)0.70710678118654752440;
  bfly4(v[0], v[2], v[4], v[6]);   // E0..E3 in v0,v2,v4,v6
  bfly4(v[1], v[3], v[5], v[7]);   // O0..O3 in v1,v3,v5,v7
  T2 o1 = {(v[3].x + v[3].y) * h, (v[3].y - v[3].x) * h};        // O1 * (1-i)/sqrt2
  T2 o2 = mul_mi(v[5]);                                         // O2 * (-i)
  T2 o3 = {(v[7].y - v[7].x) * h, -(v[7].x + v[7].y) * h};       // O3 * (-1-i)/sqrt2
  T2 e0 = v[0], e1 = v[2], e2 = v[4], e3 = v[6], o0 = v[1];
  v[0] = cadd(e0, o0); v[4] = csub(e0, o0);
  v[1] = cadd(e1, o1); v[5] = csub(e1, o1);
  v[2] = cadd(e2, o2); v[6] = csub(e2, o2);
  v[3] = cadd(e3, o3); v[7] = csub(e3, o3);
}
template <typename T, typename T2> __device__ __forceinline__ void bfly3(T2& v0, T2& v1, T2& v2) {
  const T s = (T)0.86602540378443864676;
  T2 t1 = cadd(v1, v2);
  T2 t2 = {v0.x - (T)0.5 * t1.x, v0.y - (T)0.5 * t1.y};
  T2 d = csub(v1, v2);
  T2 t3 = {d.y * s, -d.x * s};     // (-i) * s * (v1 - v2)
  v0 = cadd(v0, t1); v1 = cadd(t2, t3); v2 = csub(t2, t3);
}

// two-level strides of the tile-blocked buffers (stage.h)
__device__ __forceinline__ int64_t seg_row_off(const P3dSeg& sg, int i) {
  return sg.kw > 1 ? (int64_t)(i / sg.kw) * sg.psh + (int64_t)(i % sg.kw) * sg.ps : (int64_t)i * sg.ps;
}
__device__ __forceinline__ int64_t seg_line_off(const P3dSeg& sg, int a) {
  return sg.aw > 1 ? (int64_t)(a / sg.aw) * sg.sah + (int64_t)(a % sg.aw) * sg.sa : (int64_t)a * sg.sa;
}

__device__ __forceinline__ int64_t seg_b_off(const P3dSeg& sg, int b) {
  return sg.bw > 1 ? (int64_t)(b / sg.bw) * sg.sbh + (int64_t)(b % sg.bw) * sg.sb : (int64_t)b * sg.sb;
}

struct SmemMap {
  int layx, tile, ldl;
  __device__ __forceinline__ int operator()(int k, int t) const {
    return layx ? t * ldl + k + (k >> 3) + (k >> 6) : k * tile + t;
  }
};

// position of spectrum index k after the in-place DIF passes (mixed-radix digit reversal)
__device__ __forceinline__ int digit_rev(int k, int nfft, int nfac, const int* fac) {
  int pos = 0, ncur = nfft;
  for (int i = 0; i < nfac; i++) {
    int r = fac[i];
    int q = k % r; k /= r; ncur /= r;
    pos += q * ncur;
  }
  return pos;
}

template <typename T, int R>
__device__ __forceinline__ void dif_pass(typename Cx<T>::type* s, const SmemMap& sm, int lines, int nfft,
                                         int ncur, const typename Cx<T>::type* __restrict__ tw) {
  using T2 = typename Cx<T>::type;
  const int m = ncur / R, per_line = nfft / R, total = lines * per_line, twstep = nfft / ncur;
  for (int w = threadIdx.x; w < total; w += blockDim.x) {
    int t, u;
    if (sm.layx) { t = w / per_line; u = w - t * per_line; } else { u = w / lines; t = w - u * lines; }
    int blk = u / m, j = u - blk * m, base = blk * ncur + j;
    T2 v[R];
#pragma unroll
    for (int p = 0; p < R; p++) v[p] = s[sm(base + p * m, t)];
    if (R == 2) bfly2(v[0], v[1]);
    else if (R == 3) bfly3<T>(v[0], v[1], v[2]);
    else if (R == 4) bfly4(v[0], v[1], v[2], v[3]);
    else if (R == 8) bfly8<T>(v);
    if (m > 1) {
#pragma unroll
      for (int q = 1; q < R; q++) v[q] = cmul(v[q], tw[j * q * twstep]);
    }
#pragma unroll
    for (int q = 0; q < R; q++) s[sm(base + q * m, t)] = v[q];
  }
}

// any radix r <= 32 (odd primes): O(r^2) DFT with roots read from the twiddle table
template <typename T>
__device__ __noinline__ void dif_pass_generic(typename Cx<T>::type* s, const SmemMap& sm, int lines, int nfft,
                                              int ncur, int r, const typename Cx<T>::type* __restrict__ tw) {
  using T2 = typename Cx<T>::type;
  const int m = ncur / r, per_line = nfft / r, total = lines * per_line, twstep = nfft / ncur, rstep = nfft / r;
  for (int w = threadIdx.x; w < total; w += blockDim.x) {
    int t, u;
    if (sm.layx) { t = w / per_line; u = w - t * per_line; } else { u = w / lines; t = w - u * lines; }
    int blk = u / m, j = u - blk * m, base = blk * ncur + j;
    T2 v[32];
    for (int p = 0; p < r; p++) v[p] = s[sm(base + p * m, t)];
    for (int q = 0; q < r; q++) {
      T2 acc = v[0];
      int e = 0;
      for (int p = 1; p < r; p++) {
        e += q; if (e >= r) e -= r;
        acc = cadd(acc, cmul(v[p], tw[e * rstep]));
      }
      if (m > 1 && q > 0) acc = cmul(acc, tw[j * q * twstep]);
      s[sm(base + q * m, t)] = acc;
    }
  }
}

// radix r > 32 (large prime factors, e.g. the even extension 2*(nz-1) of an odd-length DCT-I):
// O(r^2) DFT computed out of place into the second half of the tile, then copied back.
template <typename T>
__device__ __noinline__ void dif_pass_big(typename Cx<T>::type* s, typename Cx<T>::type* s2, const SmemMap& sm, int lines,
                                          int nfft, int ncur, int r, const typename Cx<T>::type* __restrict__ tw) {
  using T2 = typename Cx<T>::type;
  const int m = ncur / r, twstep = nfft / ncur, rstep = nfft / r, total = lines * nfft;
  for (int w = threadIdx.x; w < total; w += blockDim.x) {
    int t, k;
    if (sm.layx) { t = w / nfft; k = w - t * nfft; } else { k = w / lines; t = w - k * lines; }
    const int blk = k / ncur, rem = k - blk * ncur, q = rem / m, j = rem - q * m, base = blk * ncur + j;
    T2 acc = s[sm(base, t)];
    int e = 0;
    for (int p = 1; p < r; p++) {
      e += q; if (e >= r) e -= r;
      acc = cadd(acc, cmul(s[sm(base + p * m, t)], tw[e * rstep]));
    }
    if (m > 1 && q > 0) acc = cmul(acc, tw[(int)(((long long)j * q * twstep) % nfft)]);
    s2[sm(k, t)] = acc;
  }
  __syncthreads();
  for (int w = threadIdx.x; w < total; w += blockDim.x) {
    int t, k;
    if (sm.layx) { t = w / nfft; k = w - t * nfft; } else { k = w / lines; t = w - k * lines; }
    s[sm(k, t)] = s2[sm(k, t)];
  }
}

template <typename T>
__device__ __forceinline__ void fft_inplace(typename Cx<T>::type* s, const SmemMap& sm, int lines,
                                            const P3dStage& st, const typename Cx<T>::type* tw) {
  int ncur = st.nfft;
  for (int i = 0; i < st.nfac; i++) {
    int r = st.fac[i];
    switch (r) {
      case 8: dif_pass<T, 8>(s, sm, lines, st.nfft, ncur, tw); break;
      case 4: dif_pass<T, 4>(s, sm, lines, st.nfft, ncur, tw); break;
      case 2: dif_pass<T, 2>(s, sm, lines, st.nfft, ncur, tw); break;
      case 3: dif_pass<T, 3>(s, sm, lines, st.nfft, ncur, tw); break;
      default:
        if (r <= 32) dif_pass_generic<T>(s, sm, lines, st.nfft, ncur, r, tw);
        else dif_pass_big<T>(s, s + st.tile * (st.layx ? sm.ldl : st.nfft), sm, lines, st.nfft, ncur, r, tw);
        break;
    }
    ncur /= r;
    __syncthreads();
  }
}

template <typename T, int KIND>
__global__ void __launch_bounds__(256) stage_kernel(const __grid_constant__ P3dStage st) {
  using T2 = typename Cx<T>::type;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  T2* s = reinterpret_cast<T2*>(smem_raw);
  const T2* __restrict__ tw = reinterpret_cast<const T2*>(st.tw);

  const int tiles_a = (st.na + st.tile - 1) / st.tile;
  int bid = blockIdx.x;
  const int ta = bid % tiles_a; bid /= tiles_a;
  const int b = bid % st.nb;
  const int c = bid / st.nb;
  const int a0 = ta * st.tile;
  const int lines = min(st.tile, st.na - a0);
  const int nfft = st.nfft, n = st.n;
  SmemMap sm{st.layx, lines, nfft + (nfft >> 3) + (nfft >> 6) + 1};
  if (st.layx) sm.tile = st.tile;

  // ---- clear (only when some logical inputs are not stored) -------------------------
  if (st.need_zero) {
    const int tot = st.layx ? st.tile * sm.ldl : nfft * lines;
    for (int i = threadIdx.x; i < tot; i += blockDim.x) s[i] = T2{0, 0};
    __syncthreads();
  }

  // ---- load phase ------------------------------------------------------------------
  {
    const int shift = st.in.logical - st.in.cnt;
    for (int g = 0; g < st.in.nseg; g++) {
      const P3dSeg& sg = st.in.seg[g];
      const int64_t lbase = seg_b_off(sg, b) + (int64_t)c * sg.sc;
      const int tot = sg.len * lines;
      for (int w = threadIdx.x; w < tot; w += blockDim.x) {
        int t, i;
        if (sg.ps == 1) { t = w / sg.len; i = w - t * sg.len; } else { i = w / lines; t = w - i * lines; }
        const int sidx = sg.start + i;
        const int k = sidx < st.in.h1 ? sidx : sidx + shift;
        const int64_t addr = lbase + seg_row_off(sg, i) + seg_line_off(sg, a0 + t);
        if (KIND == P3D_R2C) {
          T v = reinterpret_cast<const T*>(sg.base)[addr];
          s[sm(k, t)] = T2{v, 0};
        } else {
          T2 v = reinterpret_cast<const T2*>(sg.base)[addr];
          if (KIND == P3D_C2C_FWD || KIND == P3D_NOOP) s[sm(k, t)] = v;
          else if (KIND == P3D_C2C_BWD) s[sm(k, t)] = T2{v.x, -v.y};
          else if (KIND == P3D_C2R) {
            // Hermitian completion; x_j = Re FFT(conj X)_j.  imag of DC / Nyquist ignored like c2r.
            if (k == 0 || 2 * k == n) s[sm(k, t)] = T2{v.x, 0};
            else { s[sm(k, t)] = T2{v.x, -v.y}; s[sm(n - k, t)] = v; }
          } else if (KIND == P3D_DCT1) {
            s[sm(k, t)] = v;
            if (k >= 1 && k <= n - 2) s[sm(nfft - k, t)] = v;
          } else if (KIND == P3D_DST1) {
            s[sm(k + 1, t)] = v;
            s[sm(nfft - (k + 1), t)] = T2{-v.x, -v.y};
          }
        }
      }
    }
  }
  __syncthreads();

  // ---- transform -------------------------------------------------------------------
  if (KIND != P3D_NOOP) fft_inplace<T>(s, sm, lines, st, tw);

  // ---- store phase -----------------------------------------------------------------
  {
    const int shift = st.out.logical - st.out.cnt;
    const T scale = (T)st.scale;
    for (int g = 0; g < st.out.nseg; g++) {
      const P3dSeg& sg = st.out.seg[g];
      const int64_t lbase = seg_b_off(sg, b) + (int64_t)c * sg.sc;
      const int tot = sg.len * lines;
      for (int w = threadIdx.x; w < tot; w += blockDim.x) {
        int t, i;
        if (sg.ps == 1) { t = w / sg.len; i = w - t * sg.len; } else { i = w / lines; t = w - i * lines; }
        const int sidx = sg.start + i;
        int k = sidx < st.out.h1 ? sidx : sidx + shift;
        if (KIND == P3D_DST1) k += 1;
        const int pos = (KIND == P3D_NOOP) ? k : digit_rev(k, nfft, st.nfac, st.fac);
        T2 v = s[sm(pos, t)];
        const int64_t addr = lbase + seg_row_off(sg, i) + seg_line_off(sg, a0 + t);
        if (KIND == P3D_C2R) {
          reinterpret_cast<T*>(sg.base)[addr] = v.x * scale;
        } else {
          if (KIND == P3D_C2C_BWD) v.y = -v.y;
          if (KIND == P3D_DST1) v = T2{-v.y, v.x};     // Y_k = i * W_{k+1}
          v.x *= scale; v.y *= scale;
          reinterpret_cast<T2*>(sg.base)[addr] = v;
        }
      }
    }
  }
}

// ------------------------------------------------------------------------------------
// Chebyshev epilogue (ftran.F90:408-451): scale by 1/(nx ny (nzc-1)) then the derivative
// recurrence along z, one thread per (x,y) column, coalesced across columns.
// ------------------------------------------------------------------------------------
template <typename T>
__global__ void cheby_kernel(typename Cx<T>::type* out, int64_t ncol, int nzc, int64_t zstride, int64_t colstride,
                             T norm, T lfac) {
  using T2 = typename Cx<T>::type;
  int64_t col = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (col >= ncol) return;
  T2* p = out + col * colstride;
  auto at = [&](int k) -> T2& { return p[(int64_t)(k - 1) * zstride]; };   // 1-based like the reference
  if (nzc < 2) return;
  T2 a_hi = at(nzc); a_hi.x *= norm; a_hi.y *= norm;          // a(nzc)
  T2 old = at(nzc - 1); old.x *= norm; old.y *= norm;         // a(nzc-1)
  T2 o_k2 = T2{0, 0};                                         // out(k+2), starts as out(nzc) = 0
  T2 o_k1 = T2{lfac * (T)(nzc - 1) * a_hi.x * (T)0.5, lfac * (T)(nzc - 1) * a_hi.y * (T)0.5};  // out(nzc-1)
  at(nzc) = o_k2;
  at(nzc - 1) = o_k1;
  // The recurrence is sequential in k, its loads are not: U coefficients are fetched at once (U independent loads in flight
  // per thread), then U steps run in registers.  With one dependent load per step a column costs nzc DRAM latencies, which is
  // what a rank of a multi-GPU grid pays in full (few columns per GPU: 1.3 ms of the 1.7 ms Z stage of config 5a on 2x4).
  constexpr int U = 16;
  for (int k = nzc - 2; k >= 1; k -= U) {
    const int cnt = k < U ? k : U;                            // steps k, k-1, ..., k-cnt+1
    T2 a[U];
#pragma unroll
    for (int u = 0; u < U; u++) if (u < cnt) a[u] = at(k - u);
#pragma unroll
    for (int u = 0; u < U; u++) {
      if (u < cnt) {
        const int kk = k - u;
        T2 nw = a[u]; nw.x *= norm; nw.y *= norm;             // a(kk)
        T2 o = T2{lfac * (T)kk * old.x + o_k2.x, lfac * (T)kk * old.y + o_k2.y};
        if (kk == 1) { o.x *= (T)0.5; o.y *= (T)0.5; }
        at(kk) = o;
        o_k2 = o_k1; o_k1 = o; old = nw;
      }
    }
  }
  if (nzc == 2) { T2 o = at(1); o.x *= (T)0.5; o.y *= (T)0.5; at(1) = o; }
}

// ------------------------------------------------------------------------------------
// Power spectrum (driver_spec.c:298-384): one warp per row of the contiguous direction, per-CTA privatised
// histograms in shared memory (one copy per group of warps), one global atomic per bin and CTA at the end.
// ------------------------------------------------------------------------------------
__device__ __forceinline__ int spec_wavenumber(const SpecJob& j, int ax, int local) {
  const int s = j.start[ax] + local;                              // stored global index
  int k = s < j.nch[ax] ? s : s + (j.n[ax] - j.nc[ax]);           // the mode it holds (pruned transforms skip the middle)
  if (ax != 0 && k > j.n[ax] / 2) k = j.n[ax] - k;                // driver_spec.c:352-358; kx <= nx/2 is never folded
  return k;
}

template <typename T>
__global__ void __launch_bounds__(256) spectrum_kernel(const typename Cx<T>::type* __restrict__ B,
                                                       const __grid_constant__ SpecJob j, double* __restrict__ E) {
  using T2 = typename Cx<T>::type;
  extern __shared__ double spec_hist[];
  const int nbin = j.kmax + 1;
  for (int i = threadIdx.x; i < nbin * j.ncopy; i += blockDim.x) spec_hist[i] = 0.0;
  __syncthreads();
  const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
  double* h = spec_hist + (wib % j.ncopy) * nbin;
  const long long rows = (long long)j.ext[1] * j.ext[2];
  const long long nwarp = (long long)gridDim.x * (blockDim.x >> 5);
  for (long long row = (long long)blockIdx.x * (blockDim.x >> 5) + wib; row < rows; row += nwarp) {
    const int v = (int)(row % j.ext[1]), w = (int)(row / j.ext[1]);
    const int kv = spec_wavenumber(j, j.axis[1], v), kw = spec_wavenumber(j, j.axis[2], w);
    const int kvw2 = kv * kv + kw * kw;
    const T2* p = B + (long long)v * j.stride[1] + (long long)w * j.stride[2];
    // Four independent 16-byte loads per lane in flight, then per 32 consecutive u: lanes that fall into the same
    // shell are adjacent (k grows with u), so their contributions are summed with a segmented warp scan and only
    // the last lane of each run issues the shared-memory atomic (a 64-bit CAS loop) -- a handful per warp
    // instead of 32 colliding ones.
    for (int u0 = 0; u0 < j.ext[0]; u0 += 128) {      // warp-uniform trip count: the shuffles below need all lanes
      T2 z[4];
#pragma unroll
      for (int i = 0; i < 4; i++) {
        const int u = u0 + 32 * i + lane;
        z[i] = u < j.ext[0] ? p[(long long)u * j.stride[0]] : T2{0, 0};
      }
#pragma unroll
      for (int i = 0; i < 4; i++) {
        const int u = u0 + 32 * i + lane;
        int ik = 0x7fffffff;                            // lanes past the end of the row: a run of their own, value 0
        double val = 0.0;
        if (u < j.ext[0]) {
          const int ku = spec_wavenumber(j, j.axis[0], u);
          const int k2 = ku * ku + kvw2;
          // ik = int(sqrt(k2) + 0.5) exactly: float estimate, then ik is the integer with ik(ik-1) < k2 <= ik(ik+1)
          ik = (int)(sqrtf((float)k2) + 0.5f);
          if (k2 > ik * (ik + 1)) ik++;
          else if (ik > 0 && k2 <= ik * (ik - 1)) ik--;
          val = (double)k2 * ((double)z[i].x * (double)z[i].x + (double)z[i].y * (double)z[i].y) * j.f2;
        }
        const int prev = __shfl_up_sync(0xffffffffu, ik, 1);
        const unsigned heads = __ballot_sync(0xffffffffu, lane == 0 || prev != ik);
        const int first = 31 - __clz((int)(heads & ((2u << lane) - 1u)));      // first lane of this lane's run
#pragma unroll
        for (int dd = 1; dd < 32; dd <<= 1) {
          const double o = __shfl_up_sync(0xffffffffu, val, dd);
          if (lane - dd >= first) val += o;
        }
        const bool tail = lane == 31 || ((heads >> (lane + 1)) & 1u);
        if (tail && ik <= j.kmax && val != 0.0) atomicAdd(h + ik, val);
      }
    }
  }
  __syncthreads();
  for (int i = threadIdx.x; i < nbin; i += blockDim.x) {
    double acc = 0.0;
    for (int c = 0; c < j.ncopy; c++) acc += spec_hist[c * nbin + i];
    if (acc != 0.0) atomicAdd(E + i, acc);
  }
}

template <typename T>
cudaError_t launch_spectrum(const void* B, const SpecJob& job_in, double* E, cudaStream_t stream) {
  SpecJob job = job_in;
  const long long rows = (long long)job.ext[1] * job.ext[2];
  if (rows <= 0 || job.ext[0] <= 0) return cudaSuccess;
  const size_t per = (size_t)(job.kmax + 1) * sizeof(double);
  // histogram copies per CTA: as many as fit in ~28 KB (at most 4), so that 8 CTAs = 64 warps stay resident per SM
  // and the loads of one warp hide under the atomics of the others (r1 ncu: 2 CTAs/SM were latency bound)
  if (per > 200 * 1024) return cudaErrorInvalidValue;
  int ncopy = (int)((28 * 1024) / per);
  if (ncopy < 1) ncopy = 1;
  if (ncopy > 4) ncopy = 4;
  job.ncopy = ncopy;
  const size_t smem = per * ncopy;
  int per_sm = (int)((200 * 1024) / (smem + 1024));
  if (per_sm > 8) per_sm = 8;
  if (per_sm < 1) per_sm = 1;
  static size_t configured = 0;
  if (smem > 48 * 1024 && smem > configured) {
    cudaError_t e = cudaFuncSetAttribute(spectrum_kernel<T>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    configured = smem;
  }
  int dev = 0, sms = 148;
  cudaGetDevice(&dev);
  if (cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || sms <= 0) sms = 148;
  long long grid = (rows + 7) / 8;
  if (grid > (long long)sms * per_sm) grid = (long long)sms * per_sm;
  P3D_KLAUNCH(spectrum_kernel<T>, (unsigned)grid, 256, smem, stream, reinterpret_cast<const typename Cx<T>::type*>(B), job, E);
  return cudaGetLastError();
}

// ------------------------------------------------------------------------------------
// Flag barrier across the GPUs of one NVSwitch box (opt-in replacement of the one-float NCCL all-reduce that
// orders the peer-to-peer transposes, api.cpp world_barrier).  The preceding stage kernel has completed when
// this kernel starts (same stream), so its peer stores are performed; thread r publishes this rank's epoch in
// rank r's array with a system-scope release store and then spins with acquire loads on slot r of the local
// array.  Epochs only grow, so a rank that is one barrier ahead never unblocks a waiter early.
// ------------------------------------------------------------------------------------
__global__ void flag_barrier_kernel(unsigned* const* __restrict__ peers, int me, int nrank, unsigned epoch) {
  const int r = threadIdx.x;
  if (r < nrank) {
    unsigned* dst = peers[r] + (size_t)me * 32;
    const unsigned* src = peers[me] + (size_t)r * 32;
    unsigned v;
#ifndef P3D_EMULATE
    __threadfence_system();
    asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(dst), "r"(epoch) : "memory");
    do {
      asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(src) : "memory");
    } while ((int)(v - epoch) < 0);
#else
    __atomic_store_n(dst, epoch, __ATOMIC_RELEASE);
    do { v = __atomic_load_n(src, __ATOMIC_ACQUIRE); } while ((int)(v - epoch) < 0);
#endif
  }
  __syncthreads();
}

// The two halves of the barrier as kernels of their own, for the pipelined groups (api.cpp): `signal` follows a
// producer chunk on the main stream (its stores are complete when it runs), `wait` precedes the consumer chunk on the
// side stream -- so the next producer chunk need not wait for the slowest peer.
__global__ void flag_signal_kernel(unsigned* const* __restrict__ peers, int me, int nrank, unsigned epoch) {
  const int r = threadIdx.x;
  if (r < nrank) {
    unsigned* dst = peers[r] + (size_t)me * 32;
#ifndef P3D_EMULATE
    __threadfence_system();
    asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(dst), "r"(epoch) : "memory");
#else
    __atomic_store_n(dst, epoch, __ATOMIC_RELEASE);
#endif
  }
}
__global__ void flag_wait_kernel(unsigned* const* __restrict__ peers, int me, int nrank, unsigned epoch) {
  const int r = threadIdx.x;
  if (r < nrank) {
    const unsigned* src = peers[me] + (size_t)r * 32;
    unsigned v;
#ifndef P3D_EMULATE
    do {
      asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(src) : "memory");
    } while ((int)(v - epoch) < 0);
#else
    do { v = __atomic_load_n(src, __ATOMIC_ACQUIRE); } while ((int)(v - epoch) < 0);
#endif
  }
  __syncthreads();
}

// Scoped forms (api.cpp, scoped synchronisation): `signal` always goes to every rank -- a flag is the sender's progress counter --,
// the wait covers only the ranks of `mask` (bit r = world rank r; at most 64 ranks).  One launch: optional signal, then wait.
__global__ void flag_sync_mask_kernel(unsigned* const* __restrict__ peers, int me, int nrank, unsigned signal_epoch, int do_signal,
                                      unsigned long long mask, unsigned wait_epoch) {
  const int r = threadIdx.x;
  if (r < nrank) {
    if (do_signal) {
      unsigned* dst = peers[r] + (size_t)me * 32;
#ifndef P3D_EMULATE
      __threadfence_system();
      asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(dst), "r"(signal_epoch) : "memory");
#else
      __atomic_store_n(dst, signal_epoch, __ATOMIC_RELEASE);
#endif
    }
    if ((mask >> r) & 1ull) {
      const unsigned* src = peers[me] + (size_t)r * 32;
      unsigned v;
#ifndef P3D_EMULATE
      do {
        asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(src) : "memory");
      } while ((int)(v - wait_epoch) < 0);
#else
      do { v = __atomic_load_n(src, __ATOMIC_ACQUIRE); } while ((int)(v - wait_epoch) < 0);
#endif
    }
  }
  __syncthreads();
}

static int flag_threads(int nrank) { return nrank <= 32 ? 32 : ((nrank + 31) / 32) * 32; }
cudaError_t launch_flag_sync_mask(unsigned* const* peers, int me, int nrank, unsigned signal_epoch, int do_signal,
                                  unsigned long long mask, unsigned wait_epoch, cudaStream_t stream) {
  if (nrank > 64) return cudaErrorInvalidValue;
  const int nt = flag_threads(nrank);
  P3D_KLAUNCH(flag_sync_mask_kernel, 1, nt, 0, stream, peers, me, nrank, signal_epoch, do_signal, mask, wait_epoch);
  return cudaGetLastError();
}
cudaError_t launch_flag_barrier(unsigned* const* peers, int me, int nrank, unsigned epoch, cudaStream_t stream) {
  if (nrank > 1024) return cudaErrorInvalidValue;
  const int nt = flag_threads(nrank);
  P3D_KLAUNCH(flag_barrier_kernel, 1, nt, 0, stream, peers, me, nrank, epoch);
  return cudaGetLastError();
}
cudaError_t launch_flag_signal(unsigned* const* peers, int me, int nrank, unsigned epoch, cudaStream_t stream) {
  if (nrank > 1024) return cudaErrorInvalidValue;
  const int nt = flag_threads(nrank);
  P3D_KLAUNCH(flag_signal_kernel, 1, nt, 0, stream, peers, me, nrank, epoch);
  return cudaGetLastError();
}
cudaError_t launch_flag_wait(unsigned* const* peers, int me, int nrank, unsigned epoch, cudaStream_t stream) {
  if (nrank > 1024) return cudaErrorInvalidValue;
  const int nt = flag_threads(nrank);
  P3D_KLAUNCH(flag_wait_kernel, 1, nt, 0, stream, peers, me, nrank, epoch);
  return cudaGetLastError();
}

// ------------------------------------------------------------------------------------
// host launchers
// ------------------------------------------------------------------------------------
static int big_factor(const P3dStage& st) {
  for (int i = 0; i < st.nfac; i++) if (st.fac[i] > 32) return 2;    // out-of-place pass needs a second tile
  return 1;
}

template <typename T>
size_t stage_smem_bytes(const P3dStage& st) {
  using T2 = typename Cx<T>::type;
  size_t per_line = st.layx ? (size_t)(st.nfft + (st.nfft >> 3) + (st.nfft >> 6) + 1) : (size_t)st.nfft;
  return per_line * st.tile * sizeof(T2) * big_factor(st);
}

template <typename T>
int choose_tile(const P3dStage& st) {
  using T2 = typename Cx<T>::type;
  const size_t budget = 64 * 1024, hard = 200 * 1024;
  size_t per_line = (st.layx ? (size_t)(st.nfft + (st.nfft >> 3) + (st.nfft >> 6) + 1) : (size_t)st.nfft) * sizeof(T2) * big_factor(st);
  int want = st.layx ? 8 : (int)(128 / sizeof(T2));   // interleaved layout: one 128 B row of lines
  int tile = (int)(budget / per_line);
  if (tile > want) tile = want;
  if (tile < 4 && !st.layx) { tile = (int)(hard / per_line); if (tile > 4) tile = 4; }
  if (tile < 1) tile = (per_line <= hard) ? 1 : 0;
  if (tile > st.na) tile = st.na;
  return tile;
}

template <typename T, int KIND>
static cudaError_t launch_kind(const P3dStage& st, cudaStream_t stream) {
  size_t smem = stage_smem_bytes<T>(st);
  static size_t configured = 0;
  if (smem > configured) {
    cudaError_t e = cudaFuncSetAttribute(stage_kernel<T, KIND>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    configured = smem;
  }
  long long tiles_a = (st.na + st.tile - 1) / st.tile;
  long long grid = tiles_a * st.nb * st.nc;
  if (grid <= 0) return cudaSuccess;
  if (grid > 0x7fffffffLL) return cudaErrorInvalidConfiguration;
  P3D_KLAUNCH((stage_kernel<T, KIND>), (unsigned)grid, 256, smem, stream, st);
  return cudaGetLastError();
}

template <typename T>
cudaError_t launch_stage(const P3dStage& st, cudaStream_t stream) {
  if (st.kind == P3D_RCOPY) return launch_rcopy<T>(st, stream);
  if (st.tile <= 0) return cudaErrorInvalidValue;
  switch (st.kind) {
    case P3D_C2C_FWD: return launch_kind<T, P3D_C2C_FWD>(st, stream);
    case P3D_C2C_BWD: return launch_kind<T, P3D_C2C_BWD>(st, stream);
    case P3D_R2C: return launch_kind<T, P3D_R2C>(st, stream);
    case P3D_C2R: return launch_kind<T, P3D_C2R>(st, stream);
    case P3D_DCT1: return launch_kind<T, P3D_DCT1>(st, stream);
    case P3D_DST1: return launch_kind<T, P3D_DST1>(st, stream);
    case P3D_NOOP: return launch_kind<T, P3D_NOOP>(st, stream);
    default: break;
  }
  return cudaErrorInvalidValue;
}

template <typename T>
cudaError_t launch_cheby(void* out, long long ncol, int nzc, long long zstride, long long colstride,
                         double norm, double lfac, cudaStream_t stream) {
  if (ncol <= 0) return cudaSuccess;
  unsigned grid = (unsigned)((ncol + 127) / 128);
  P3D_KLAUNCH(cheby_kernel<T>, grid, 128, 0, stream, reinterpret_cast<typename Cx<T>::type*>(out), ncol, nzc, zstride, colstride,
              (T)norm, (T)lfac);
  return cudaGetLastError();
}

// P3D_RCOPY stages (real-data transposes): one launch copies every (input block, output block) intersection
template <typename T>
cudaError_t launch_rcopy(const P3dStage& st, cudaStream_t stream) {
  RcopyJob job;
  if (!rcopy_boxes(st, job, sizeof(T))) return cudaErrorInvalidValue;
  if (job.nbox <= 0 || job.rows_max <= 0) return cudaSuccess;
  static int sms = 0;
  if (!sms) {
    int dev = 0;
    cudaGetDevice(&dev);
    if (cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || sms <= 0) sms = 148;
  }
  // 8 warps per CTA, one row per warp at a time; the grid is a multiple of the SM count unless there is less work
  long long gx = (job.rows_max + 7) / 8;
  const long long cap = (long long)sms * 8;
  if (gx > cap) gx = cap;
  P3D_KLAUNCH(rcopy_kernel<T>, dim3((unsigned)gx, (unsigned)job.nbox), 256, 0, stream, job);
  return cudaGetLastError();
}

template cudaError_t launch_stage<double>(const P3dStage&, cudaStream_t);
template cudaError_t launch_spectrum<double>(const void*, const SpecJob&, double*, cudaStream_t);
template cudaError_t launch_spectrum<float>(const void*, const SpecJob&, double*, cudaStream_t);
template cudaError_t launch_rcopy<double>(const P3dStage&, cudaStream_t);
template cudaError_t launch_rcopy<float>(const P3dStage&, cudaStream_t);
template cudaError_t launch_stage<float>(const P3dStage&, cudaStream_t);
template int choose_tile<double>(const P3dStage&);
template int choose_tile<float>(const P3dStage&);
template size_t stage_smem_bytes<double>(const P3dStage&);
template size_t stage_smem_bytes<float>(const P3dStage&);
template cudaError_t launch_cheby<double>(void*, long long, int, long long, long long, double, double, cudaStream_t);
template cudaError_t launch_cheby<float>(void*, long long, int, long long, long long, double, double, cudaStream_t);

}  // namespace p3d
