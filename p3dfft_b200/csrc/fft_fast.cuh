// Specialised sm_100a stage kernels for power-of-two transform lengths.
//
// One CTA transforms a tile of TX lines that are adjacent along the CONTIGUOUS direction of
// the pencil (x for the Y and Z stages, so every row of the tile is one 64- or 128-byte
// chunk of HBM; for the X stage the lines themselves are contiguous and a warp covers
// 128 bytes of each of four lines).  The transform is an in-place decimation-in-frequency
// FFT with a compile-time radix schedule:
//
//   pass 1      global -> registers (R1 independent 16-byte loads in flight per thread),
//               radix-R1 butterfly, twiddle, registers -> shared memory
//   pass 2..L-1 shared -> registers -> shared
//   pass L      shared -> registers, butterfly, registers -> global
//
// so a line makes ONE trip through HBM in each direction and L-1 trips through shared
// memory.  The unpack of the preceding all-to-all, the pack for the next one, pruning and
// zero padding are row -> address lookups (FastSide runs) on the two global sides.
//
// Shared-memory layout: element (row k, line t) of the tile lives at [k'][t] with TX
// elements (64 or 128 bytes) per row.  With 64-byte rows two rows share one 128-byte bank
// window, so k' = k ^ parity(k >> 1): any two rows whose indices differ in exactly one bit
// (all the pairs adjacent threads touch in a power-of-two DIF pass) land in different
// halves of the window and every pass is bank-conflict free.  Because the rows of one
// butterfly differ only in a bit-field disjoint from the rest of the index, the swizzle of
// row base|p*m is swizzle(base) ^ const(p): one POPC per butterfly, one XOR per access.
//
// X stage (r2c / c2r): the real line of length N is transformed as a complex FFT of length
// H = N/2 on the packed pairs (x[2j], x[2j+1]); the Hermitian post-/pre-processing needs the
// pair (k, H-k), which is produced by two butterflies of the last (r2c) or consumed by two
// butterflies of the first (c2r) pass.  One thread owns both, so the combination happens in
// registers and costs no extra pass (replaces exec_f_r2c / exec_b_c2r, fft_exec.F90:495,298).
//
// Backward transforms use FFT^-1(z) = swap(FFT(swap(z))) (swap = exchange re and im): the
// same forward butterflies and tables, the swap is free at the load and the store.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include <string.h>

#include <type_traits>
#include <utility>

#include "fast.h"
#include "stage.h"

namespace p3d {
namespace fast {

template <typename T> struct Cx;
template <> struct Cx<double> { using type = double2; };
template <> struct Cx<float>  { using type = float2; };

// ---------------------------------------------------------------------------------------
// radix schedules
// ---------------------------------------------------------------------------------------
template <int A, int B = 1, int C = 1, int D = 1>
struct RS {
  static constexpr int L = 1 + (B > 1) + (C > 1) + (D > 1);
  static constexpr int N = A * B * C * D;
  __host__ __device__ static constexpr int r(int i) { return i == 0 ? A : i == 1 ? B : i == 2 ? C : D; }
  // length of the sub-transforms entering pass i
  __host__ __device__ static constexpr int ncur(int i) { int n = N; for (int k = 0; k < i; k++) n /= r(k); return n; }
  __host__ __device__ static constexpr int m(int i) { return ncur(i) / r(i); }
  // offset (elements) of pass i's twiddle table inside the block: sum_{k<i} (r_k-1)*m_k
  __host__ __device__ static constexpr int twoff(int i) { int o = 0; for (int k = 0; k < i; k++) o += (r(k) - 1) * m(k); return o; }
  __host__ __device__ static constexpr int twtotal() { return twoff(L - 1); }   // last pass has m = 1: no table
};

// c2c configuration per (type, length, row bytes): schedule, threads per CTA, CTAs/SM hint.  A tile row is
// RB = 64 or 128 bytes of lines adjacent in x (TX = RB / sizeof(complex) lines = the block width W of the
// planner's internal layouts).  128-byte rows are one full L1 wavefront / DRAM burst pair per row and are
// the default; 64-byte rows halve the tile and exist for lengths whose 128-byte tile exceeds shared memory.
template <typename T, int N, int RB> struct CCfg;
#define P3D_CCFG(T, N, RB, SCHED, NTV, MINBV)                                                    \
  template <> struct CCfg<T, N, RB> {                                                            \
    using S = SCHED;                                                                             \
    static constexpr int TX = RB / (2 * (int)sizeof(T)), NT = NTV, MINB = MINBV;                 \
  };
#define P3D_RS(...) RS<__VA_ARGS__>
P3D_CCFG(double, 64, 64, P3D_RS(8, 8), 32, 16)
P3D_CCFG(double, 128, 64, P3D_RS(16, 8), 64, 8)
P3D_CCFG(double, 256, 64, P3D_RS(16, 16), 64, 6)
P3D_CCFG(double, 512, 64, P3D_RS(8, 8, 8), 256, 3)
P3D_CCFG(double, 1024, 64, P3D_RS(16, 8, 8), 256, 2)
P3D_CCFG(double, 2048, 64, P3D_RS(16, 16, 8), 512, 1)
P3D_CCFG(double, 64, 128, P3D_RS(8, 8), 64, 8)
P3D_CCFG(double, 128, 128, P3D_RS(16, 8), 64, 8)
P3D_CCFG(double, 256, 128, P3D_RS(16, 16), 128, 4)
P3D_CCFG(double, 512, 128, P3D_RS(8, 8, 8), 256, 2)
P3D_CCFG(double, 1024, 128, P3D_RS(16, 8, 8), 512, 1)
P3D_CCFG(float, 64, 64, P3D_RS(8, 8), 64, 16)
P3D_CCFG(float, 128, 64, P3D_RS(16, 8), 128, 8)
P3D_CCFG(float, 256, 64, P3D_RS(16, 16), 128, 6)
P3D_CCFG(float, 512, 64, P3D_RS(8, 8, 8), 256, 3)
P3D_CCFG(float, 1024, 64, P3D_RS(16, 8, 8), 512, 2)
P3D_CCFG(float, 2048, 64, P3D_RS(16, 16, 8), 512, 1)
P3D_CCFG(float, 64, 128, P3D_RS(8, 8), 128, 8)
P3D_CCFG(float, 128, 128, P3D_RS(16, 8), 128, 8)
P3D_CCFG(float, 256, 128, P3D_RS(16, 16), 256, 4)
P3D_CCFG(float, 512, 128, P3D_RS(8, 8, 8), 256, 2)
P3D_CCFG(float, 1024, 128, P3D_RS(16, 8, 8), 512, 1)
// 3 * 2^k and 5 * 2^k: the odd radix comes FIRST, so that every later sub-transform length (the M of the XOR addressing in
// twiddle_store / mid_pass) stays a power of two
P3D_CCFG(double, 384, 64, P3D_RS(3, 16, 8), 128, 4)
P3D_CCFG(double, 768, 64, P3D_RS(3, 16, 16), 256, 3)
P3D_CCFG(double, 1536, 64, P3D_RS(6, 16, 16), 256, 2)
P3D_CCFG(double, 640, 64, P3D_RS(5, 16, 8), 128, 4)
P3D_CCFG(double, 1280, 64, P3D_RS(5, 16, 16), 256, 2)
P3D_CCFG(double, 384, 128, P3D_RS(3, 16, 8), 256, 3)
P3D_CCFG(double, 768, 128, P3D_RS(3, 16, 16), 256, 2)
P3D_CCFG(double, 640, 128, P3D_RS(5, 16, 8), 256, 2)
P3D_CCFG(double, 1280, 128, P3D_RS(5, 16, 16), 512, 1)
P3D_CCFG(double, 1536, 128, P3D_RS(6, 16, 16), 512, 1)
P3D_CCFG(float, 384, 64, P3D_RS(3, 16, 8), 256, 4)
P3D_CCFG(float, 768, 64, P3D_RS(3, 16, 16), 256, 3)
P3D_CCFG(float, 1536, 64, P3D_RS(6, 16, 16), 512, 2)
P3D_CCFG(float, 640, 64, P3D_RS(5, 16, 8), 256, 4)
P3D_CCFG(float, 1280, 64, P3D_RS(5, 16, 16), 512, 2)
P3D_CCFG(float, 384, 128, P3D_RS(3, 16, 8), 256, 3)
P3D_CCFG(float, 768, 128, P3D_RS(3, 16, 16), 512, 2)
P3D_CCFG(float, 640, 128, P3D_RS(5, 16, 8), 512, 2)
P3D_CCFG(float, 1280, 128, P3D_RS(5, 16, 16), 512, 1)
P3D_CCFG(float, 1536, 128, P3D_RS(6, 16, 16), 512, 1)
// 2048 points with 128-byte rows: the tile is 256 KB and exists only for the SPLIT kernel (half of it in shared memory, the
// other half waiting in registers); schedule and line count come from here, threads from SplitCfg
P3D_CCFG(double, 2048, 128, P3D_RS(16, 16, 8), 512, 1)
P3D_CCFG(float, 2048, 128, P3D_RS(16, 16, 8), 512, 1)
#undef P3D_CCFG
// the 128-byte tile of a 2048-point transform (256 KB) does not fit in shared memory; 1536 points (192 KB + 24 KB of row
// tables) still do, with one CTA per SM
constexpr bool csplit_only(int n) { return n == 2048; }
constexpr bool ccfg_exists(int n, int rb) { return rb == 64 || (rb == 128 && (n <= 1536 || csplit_only(n))); }

// Two-pass variants (opt-in, P3DFFT_B200_R32=1): 1024 = 32 x 32 and 512 = 16 x 32 instead of three passes, i.e. ONE round
// trip through shared memory per element instead of two (r1 ncu: the LSU pipe is busy ~55 % of a 1024-point stage, most
// of it shared-memory wavefronts).  The price is 32 complex values per thread in registers, hence fewer threads per CTA.
// 128-byte rows only.  Not yet timed on hardware.
template <typename T, int N> struct CCfgR32;
template <> struct CCfgR32<double, 1024> { using S = RS<32, 32>; static constexpr int TX = 8, NT = 256, MINB = 1; };
template <> struct CCfgR32<double, 512>  { using S = RS<16, 32>; static constexpr int TX = 8, NT = 128, MINB = 3; };
template <> struct CCfgR32<float, 1024>  { using S = RS<32, 32>; static constexpr int TX = 16, NT = 512, MINB = 1; };
template <> struct CCfgR32<float, 512>   { using S = RS<16, 32>; static constexpr int TX = 16, NT = 256, MINB = 3; };
constexpr bool ccfg_r32_exists(int n) { return n == 1024 || n == 512; }

// X-stage configuration per (type, H = nx/2).  The first (c2r) / last (r2c) pass works on
// butterfly PAIRS, i.e. 2R complex values per thread, so those radices stay <= 8.
// The X tile is kept LINE-major in shared memory ([line][k], pitch H + 4 elements) because the
// lines are contiguous in HBM: a warp then loads / stores 512 contiguous bytes of one line.
// Element k of a line sits at k ^ (((k >> SA) ^ (k >> SB)) & 7) (SB = 0: one term): with these
// shifts every pass of the schedule -- including the digit-reversed last pass and the
// (k, H-k) pair pass -- is free of bank conflicts (H = 128: 12 % replays).
template <typename T, int HH> struct XCfg;
#define P3D_XCFG(T, HV, SCHED, TXV, NTV, MINBV, SAV, SBV)                                          \
  template <> struct XCfg<T, HV> {                                                                 \
    using S = SCHED;                                                                               \
    static constexpr int H = HV, TX = TXV, NT = NTV, MINB = MINBV, SA = SAV, SB = SBV, LP = HV + 4; \
  };
P3D_XCFG(double, 32, P3D_RS(4, 8), 4, 32, 8, 3, 0)
P3D_XCFG(double, 64, P3D_RS(8, 8), 4, 32, 8, 3, 0)
P3D_XCFG(double, 128, P3D_RS(4, 4, 8), 4, 64, 8, 3, 5)
P3D_XCFG(double, 256, P3D_RS(8, 4, 8), 4, 128, 4, 5, 0)
P3D_XCFG(double, 512, P3D_RS(8, 8, 8), 4, 128, 4, 6, 0)
P3D_XCFG(double, 1024, P3D_RS(8, 16, 8), 4, 256, 2, 7, 0)
P3D_XCFG(float, 32, P3D_RS(4, 8), 8, 64, 8, 3, 0)
P3D_XCFG(float, 64, P3D_RS(8, 8), 8, 64, 8, 3, 0)
P3D_XCFG(float, 128, P3D_RS(4, 4, 8), 8, 128, 6, 3, 5)
P3D_XCFG(float, 256, P3D_RS(8, 4, 8), 8, 256, 3, 5, 0)
P3D_XCFG(float, 512, P3D_RS(8, 8, 8), 8, 256, 3, 6, 0)
P3D_XCFG(float, 1024, P3D_RS(8, 16, 8), 8, 512, 2, 7, 0)
// nx = 3 * 2^k, 5 * 2^k: the pair passes (first pass of c2r, last pass of r2c) need EVEN radices <= 8, the odd factor sits in
// the middle (the X kernels address shared memory additively, so no pass needs a power-of-two sub-transform length).  The
// swizzle shifts are not tuned for these lengths (some bank conflicts; correctness does not depend on them).
P3D_XCFG(double, 192, P3D_RS(4, 6, 8), 4, 64, 8, 3, 5)
P3D_XCFG(double, 384, P3D_RS(8, 6, 8), 4, 128, 4, 5, 0)
P3D_XCFG(double, 768, P3D_RS(8, 3, 4, 8), 4, 256, 2, 6, 0)
P3D_XCFG(double, 320, P3D_RS(8, 5, 8), 4, 128, 4, 5, 0)
P3D_XCFG(double, 640, P3D_RS(8, 5, 2, 8), 4, 256, 2, 6, 0)
P3D_XCFG(float, 192, P3D_RS(4, 6, 8), 8, 128, 6, 3, 5)
P3D_XCFG(float, 384, P3D_RS(8, 6, 8), 8, 256, 3, 5, 0)
P3D_XCFG(float, 768, P3D_RS(8, 3, 4, 8), 8, 512, 2, 6, 0)
P3D_XCFG(float, 320, P3D_RS(8, 5, 8), 8, 256, 3, 5, 0)
P3D_XCFG(float, 640, P3D_RS(8, 5, 2, 8), 8, 512, 2, 6, 0)
#undef P3D_XCFG
#undef P3D_RS
#if defined(__CUDACC__) || defined(P3D_EMULATE)      // P3D_EMULATE: host emulation of the kernels (tests/emu, CPU tests)
// ---------------------------------------------------------------------------------------
// complex helpers and natural-order forward butterflies
// ---------------------------------------------------------------------------------------
template <typename T2> __device__ __forceinline__ T2 cadd(T2 a, T2 b) { return T2{a.x + b.x, a.y + b.y}; }
template <typename T2> __device__ __forceinline__ T2 csub(T2 a, T2 b) { return T2{a.x - b.x, a.y - b.y}; }
template <typename T2> __device__ __forceinline__ T2 cmul(T2 a, T2 b) {
  return T2{a.x * b.x - a.y * b.y, a.x * b.y + a.y * b.x};
}
template <typename T2> __device__ __forceinline__ T2 cconj(T2 a) { return T2{a.x, -a.y}; }
template <typename T2> __device__ __forceinline__ T2 mul_mi(T2 a) { return T2{a.y, -a.x}; }   // * (-i)
template <typename T2> __device__ __forceinline__ T2 mul_pi(T2 a) { return T2{-a.y, a.x}; }   // * (+i)
template <typename T2> __device__ __forceinline__ T2 cswap(T2 a) { return T2{a.y, a.x}; }
template <typename T2> __device__ __forceinline__ T2 sel(bool c, T2 a, T2 b) { return T2{c ? a.x : b.x, c ? a.y : b.y}; }

template <typename T2> __device__ __forceinline__ void bf2(T2& a, T2& b) { T2 t = csub(a, b); a = cadd(a, b); b = t; }
template <typename T2> __device__ __forceinline__ void bf4(T2& v0, T2& v1, T2& v2, T2& v3) {
  T2 a = cadd(v0, v2), b = csub(v0, v2), c = cadd(v1, v3), d = mul_mi(csub(v1, v3));
  v0 = cadd(a, c); v2 = csub(a, c); v1 = cadd(b, d); v3 = csub(b, d);
}

template <typename T, int R> struct Bfly;
template <typename T> struct Bfly<T, 2> {
  using T2 = typename Cx<T>::type;
  __device__ __forceinline__ static void run(T2* v) { bf2(v[0], v[1]); }
};
template <typename T> struct Bfly<T, 4> {
  using T2 = typename Cx<T>::type;
  __device__ __forceinline__ static void run(T2* v) { bf4(v[0], v[1], v[2], v[3]); }
};
// odd radices (lengths 3 * 2^k, 5 * 2^k: the DNS-typical sizes 384, 768, 1536, 640, 1280): forward DFT, natural order
template <typename T> struct Bfly<T, 3> {
  using T2 = typename Cx<T>::type;
  __device__ __forceinline__ static void run(T2* v) {
    const T s = (T)0.86602540378443864676, hf = (T)0.5;
    const T2 t1 = cadd(v[1], v[2]), d = csub(v[1], v[2]);
    const T2 t2 = T2{v[0].x - hf * t1.x, v[0].y - hf * t1.y};
    const T2 r = T2{s * d.y, -s * d.x};                         // -i s (v1 - v2)
    v[0] = cadd(v[0], t1); v[1] = cadd(t2, r); v[2] = csub(t2, r);
  }
};
template <typename T> struct Bfly<T, 5> {
  using T2 = typename Cx<T>::type;
  __device__ __forceinline__ static void run(T2* v) {
    const T c1 = (T)0.30901699437494742410, c2 = (T)-0.80901699437494742410, s1 = (T)0.95105651629515357212, s2 = (T)0.58778525229247312917;
    const T2 t1 = cadd(v[1], v[4]), t2 = cadd(v[2], v[3]), t3 = csub(v[1], v[4]), t4 = csub(v[2], v[3]);
    const T2 m1 = T2{v[0].x + c1 * t1.x + c2 * t2.x, v[0].y + c1 * t1.y + c2 * t2.y};
    const T2 m2 = T2{v[0].x + c2 * t1.x + c1 * t2.x, v[0].y + c2 * t1.y + c1 * t2.y};
    const T2 n1 = T2{s1 * t3.x + s2 * t4.x, s1 * t3.y + s2 * t4.y};
    const T2 n2 = T2{s2 * t3.x - s1 * t4.x, s2 * t3.y - s1 * t4.y};
    v[0] = T2{v[0].x + t1.x + t2.x, v[0].y + t1.y + t2.y};
    v[1] = T2{m1.x + n1.y, m1.y - n1.x};                        // m1 - i n1
    v[4] = T2{m1.x - n1.y, m1.y + n1.x};                        // m1 + i n1
    v[2] = T2{m2.x + n2.y, m2.y - n2.x};
    v[3] = T2{m2.x - n2.y, m2.y + n2.x};
  }
};
template <typename T> struct Bfly<T, 6> {
  using T2 = typename Cx<T>::type;
  // X[k] = E[k mod 3] + W6^k O[k mod 3],  E = DFT3(v0, v2, v4), O = DFT3(v1, v3, v5)
  __device__ __forceinline__ static void run(T2* v) {
    const T s = (T)0.86602540378443864676, hf = (T)0.5;
    T2 e[3] = {v[0], v[2], v[4]}, o[3] = {v[1], v[3], v[5]};
    Bfly<T, 3>::run(e);
    Bfly<T, 3>::run(o);
    const T2 o1 = T2{hf * o[1].x + s * o[1].y, hf * o[1].y - s * o[1].x};        // O1 * W6^1 = O1 (1/2 - i s)
    const T2 o2 = T2{-hf * o[2].x + s * o[2].y, -hf * o[2].y - s * o[2].x};      // O2 * W6^2 = O2 (-1/2 - i s)
    v[0] = cadd(e[0], o[0]); v[3] = csub(e[0], o[0]);
    v[1] = cadd(e[1], o1);   v[4] = csub(e[1], o1);
    v[2] = cadd(e[2], o2);   v[5] = csub(e[2], o2);
  }
};
template <typename T> struct Bfly<T, 8> {
  using T2 = typename Cx<T>::type;
  __device__ __forceinline__ static void run(T2* v) {
    const T h = (T)0.70710678118654752440;
    bf4(v[0], v[2], v[4], v[6]);
    bf4(v[1], v[3], v[5], v[7]);
    T2 o1 = T2{(v[3].x + v[3].y) * h, (v[3].y - v[3].x) * h};        // O1 * W8^1
    T2 o2 = mul_mi(v[5]);                                            // O2 * W8^2
    T2 o3 = T2{(v[7].y - v[7].x) * h, -(v[7].x + v[7].y) * h};       // O3 * W8^3
    T2 e0 = v[0], e1 = v[2], e2 = v[4], e3 = v[6], o0 = v[1];
    v[0] = cadd(e0, o0); v[4] = csub(e0, o0);
    v[1] = cadd(e1, o1); v[5] = csub(e1, o1);
    v[2] = cadd(e2, o2); v[6] = csub(e2, o2);
    v[3] = cadd(e3, o3); v[7] = csub(e3, o3);
  }
};
template <typename T> struct Bfly<T, 16> {
  using T2 = typename Cx<T>::type;
  // p = 4a + b, q = c + 4d:  W16^(pq) = W4^(ac) * W16^(bc) * W4^(bd)
  __device__ __forceinline__ static void run(T2* v) {
    const T c1 = (T)0.92387953251128675613, s1 = (T)0.38268343236508977173, h = (T)0.70710678118654752440;
#pragma unroll
    for (int b = 0; b < 4; b++) bf4(v[b], v[4 + b], v[8 + b], v[12 + b]);    // v[4c+b] = A_b[c]
    v[5]  = cmul(v[5],  T2{c1, -s1});     // (b,c) = (1,1): W^1
    v[9]  = cmul(v[9],  T2{h, -h});       // (1,2): W^2
    v[13] = cmul(v[13], T2{s1, -c1});     // (1,3): W^3
    v[6]  = cmul(v[6],  T2{h, -h});       // (2,1): W^2
    v[10] = mul_mi(v[10]);                // (2,2): W^4
    v[14] = cmul(v[14], T2{-h, -h});      // (2,3): W^6
    v[7]  = cmul(v[7],  T2{s1, -c1});     // (3,1): W^3
    v[11] = cmul(v[11], T2{-h, -h});      // (3,2): W^6
    v[15] = cmul(v[15], T2{-c1, s1});     // (3,3): W^9
#pragma unroll
    for (int c = 0; c < 4; c++) bf4(v[4 * c], v[4 * c + 1], v[4 * c + 2], v[4 * c + 3]);   // v[4c+d] = y[c+4d]
#pragma unroll
    for (int c = 0; c < 4; c++)
#pragma unroll
      for (int d = c + 1; d < 4; d++) { T2 t = v[4 * c + d]; v[4 * c + d] = v[4 * d + c]; v[4 * d + c] = t; }
  }
};

template <int B, int... I, class F>
__device__ __forceinline__ void static_for_impl(F&& f, std::integer_sequence<int, I...>) { (f(std::integral_constant<int, B + I>{}), ...); }
template <int B, int E, class F>
__device__ __forceinline__ void static_for(F&& f) { static_for_impl<B>(f, std::make_integer_sequence<int, E - B>{}); }

// 32 = 4 x 8:  n = 8 n1 + n2,  k = k1 + 4 k2:   X[k1 + 4 k2] = sum_n2 W8^(n2 k2) W32^(n2 k1) sum_n1 W4^(n1 k1) x[8 n1 + n2]
constexpr double c32tab[9] = {1.0, 0.98078528040323044913, 0.92387953251128675613, 0.83146961230254523708, 0.70710678118654752440,
                              0.55557023301960222474, 0.38268343236508977173, 0.19509032201612826785, 0.0};
constexpr double cos32(int m) { m = ((m % 32) + 32) % 32; if (m > 16) m = 32 - m; return m <= 8 ? c32tab[m] : -c32tab[16 - m]; }
constexpr double sin32(int m) { return cos32(m - 8); }
template <typename T> struct Bfly<T, 32> {
  using T2 = typename Cx<T>::type;
  __device__ __forceinline__ static void run(T2* v) {
    T2 a[8][4];                                   // a[n2][k1]
#pragma unroll
    for (int n2 = 0; n2 < 8; n2++) {
      T2 x0 = v[n2], x1 = v[8 + n2], x2 = v[16 + n2], x3 = v[24 + n2];
      bf4(x0, x1, x2, x3);
      a[n2][0] = x0; a[n2][1] = x1; a[n2][2] = x2; a[n2][3] = x3;
    }
    // W32^(n2 k1): the roots are compile-time constants (constexpr evaluation only -- the table above does not exist on the device)
    static_for<1, 8>([&](auto n2c) {
      constexpr int n2 = decltype(n2c)::value;
      static_for<1, 4>([&](auto k1c) {
        constexpr int k1 = decltype(k1c)::value;
        constexpr double c = cos32(n2 * k1), sn = sin32(n2 * k1);
        a[n2][k1] = cmul(a[n2][k1], T2{(T)c, (T)(-sn)});
      });
    });
#pragma unroll
    for (int k1 = 0; k1 < 4; k1++) {
      T2 w[8];
#pragma unroll
      for (int n2 = 0; n2 < 8; n2++) w[n2] = a[n2][k1];
      Bfly<T, 8>::run(w);
#pragma unroll
      for (int k2 = 0; k2 < 8; k2++) v[k1 + 4 * k2] = w[k2];
    }
  }
};

// ---------------------------------------------------------------------------------------
// memory access
// ---------------------------------------------------------------------------------------
#ifndef P3D_EMULATE
__device__ __forceinline__ double2 ldg_stream(const double2* p) {
  double2 v;
  asm volatile("ld.global.nc.L1::no_allocate.v2.f64 {%0,%1}, [%2];" : "=d"(v.x), "=d"(v.y) : "l"(p));
  return v;
}
__device__ __forceinline__ float2 ldg_stream(const float2* p) {
  float2 v;
  asm volatile("ld.global.nc.L1::no_allocate.v2.f32 {%0,%1}, [%2];" : "=f"(v.x), "=f"(v.y) : "l"(p));
  return v;
}
__device__ __forceinline__ void stg_stream(double2* p, double2 v) {
  asm volatile("st.global.v2.f64 [%0], {%1,%2};" ::"l"(p), "d"(v.x), "d"(v.y) : "memory");
}
__device__ __forceinline__ void stg_stream(float2* p, float2 v) {
  asm volatile("st.global.v2.f32 [%0], {%1,%2};" ::"l"(p), "f"(v.x), "f"(v.y) : "memory");
}
// evict-first ("cache streaming") stores for outputs nobody reads again soon.  Measured on B200, 1024^3 double (profiles/
// r2_ab_1gpu_cache_hints.log): X c2r 4.34 -> 4.17 ms, Z backward 4.38 -> 4.30, Y backward 3.58 -> 3.53; the X r2c stage, whose
// mirrored half rows are completed in L2 by a neighbouring warp, loses (3.46 -> 3.52) and keeps the default policy.
// (The same capture: .L2::128B / .L2::256B prefetch sizes on the loads change nothing / lose 1.5 %.)
__device__ __forceinline__ void stg_cs(double2* p, double2 v) {
  asm volatile("st.global.cs.v2.f64 [%0], {%1,%2};" ::"l"(p), "d"(v.x), "d"(v.y) : "memory");
}
__device__ __forceinline__ void stg_cs(float2* p, float2 v) {
  asm volatile("st.global.cs.v2.f32 [%0], {%1,%2};" ::"l"(p), "f"(v.x), "f"(v.y) : "memory");
}
#else
__device__ __forceinline__ void stg_cs(double2* p, double2 v) { *p = v; }
__device__ __forceinline__ void stg_cs(float2* p, float2 v) { *p = v; }
__device__ __forceinline__ double2 ldg_stream(const double2* p) { return *p; }
__device__ __forceinline__ float2 ldg_stream(const float2* p) { return *p; }
__device__ __forceinline__ void stg_stream(double2* p, double2 v) { *p = v; }
__device__ __forceinline__ void stg_stream(float2* p, float2 v) { *p = v; }
#endif

constexpr int cpar(int x) { int p = 0; while (x) { p ^= x & 1; x >>= 1; } return p; }
// swizzle of a row offset that occupies a bit-field disjoint from the rest of the index
template <bool SWZ> __host__ __device__ constexpr int swz_const(int off) { return SWZ ? (off ^ cpar(off >> 1)) : off; }
template <bool SWZ> __device__ __forceinline__ int swz_base(int row) { return SWZ ? (row ^ (__popc(row >> 1) & 1)) : row; }

// ---------------------------------------------------------------------------------------
// passes over the shared-memory tile
// ---------------------------------------------------------------------------------------
// digit reversal: kappa = q1 + R1*q2 + ... (digits of passes 1..L-1)  ->  block index of the last pass
template <class S> __device__ __forceinline__ int blk_of_kappa(int kappa) {
  int blk = 0;
#pragma unroll
  for (int i = 0; i < S::L - 1; i++) {
    const int R = S::r(i);
    const int q = kappa % R; kappa /= R;
    blk = (i == 0) ? q : blk * R + q;
  }
  return blk;
}

template <typename T, class S, int PI, int TX, bool SWZ>
__device__ __forceinline__ void twiddle_store(typename Cx<T>::type* v, typename Cx<T>::type* s, const typename Cx<T>::type* __restrict__ tw,
                                              int base_row, int j, int t) {
  using T2 = typename Cx<T>::type;
  constexpr int R = S::r(PI), M = S::m(PI);
  const T2* twp = tw + S::twoff(PI) + j;
#pragma unroll
  for (int q = 1; q < R; q++) v[q] = cmul(v[q], __ldg(twp + (q - 1) * M));
  const int bi = swz_base<SWZ>(base_row) * TX + t;
#pragma unroll
  for (int q = 0; q < R; q++) s[bi ^ (swz_const<SWZ>(q * M) * TX)] = v[q];
}

// middle pass PI (0 < PI < L-1): shared -> shared, in place
template <typename T, class S, int PI, int TX, int NT, bool SWZ, int NROWS = S::N>
__device__ __forceinline__ void mid_pass(typename Cx<T>::type* s, const typename Cx<T>::type* __restrict__ tw) {
  using T2 = typename Cx<T>::type;
  constexpr int R = S::r(PI), NCUR = S::ncur(PI), M = S::m(PI), ITEMS = (NROWS / R) * TX;
#pragma unroll 1
  for (int w = threadIdx.x; w < ITEMS; w += NT) {
    const int t = w % TX, u = w / TX;
    const int blk = u / M, j = u % M, base = blk * NCUR + j;
    const int bi = swz_base<SWZ>(base) * TX + t;
    T2 v[R];
#pragma unroll
    for (int p = 0; p < R; p++) v[p] = s[bi ^ (swz_const<SWZ>(p * M) * TX)];
    Bfly<T, R>::run(v);
    twiddle_store<T, S, PI, TX, SWZ>(v, s, tw, base, j, t);
  }
}

template <typename T, class S, int PI, int TX, int NT, bool SWZ>
__device__ __forceinline__ void mid_passes(typename Cx<T>::type* s, const typename Cx<T>::type* __restrict__ tw) {
  if constexpr (PI < S::L - 1) {
    mid_pass<T, S, PI, TX, NT, SWZ>(s, tw);
    __syncthreads();
    mid_passes<T, S, PI + 1, TX, NT, SWZ>(s, tw);
  }
}

// reads the R_L rows of last-pass butterfly kappa (natural output order) and transforms them:
// v[q] = Z[kappa + q*ML]
template <typename T, class S, int TX, bool SWZ>
__device__ __forceinline__ void last_bfly(const typename Cx<T>::type* s, int kappa, int t, typename Cx<T>::type* v) {
  constexpr int RL = S::r(S::L - 1);
  const int base = blk_of_kappa<S>(kappa) * RL;
  const int bi = swz_base<SWZ>(base) * TX + t;
#pragma unroll
  for (int p = 0; p < RL; p++) v[p] = s[bi ^ (swz_const<SWZ>(p) * TX)];
  Bfly<T, RL>::run(v);
}

// ---------------------------------------------------------------------------------------
// c2c stage kernel (Y and Z stages, forward/backward, DCT-I by even extension)
//
// Persistent CTAs: each walks tiles blockIdx.x, blockIdx.x + gridDim.x, ...  While the row
// table of the current tile is built, the rows of the NEXT tile of this CTA are prefetched
// into L2, so that tile's pass-1 loads find their data on chip and HBM stays busy while the
// SM computes.
// ---------------------------------------------------------------------------------------
#ifndef P3D_EMULATE
// bulk asynchronous copy shared -> global (TMA, no tensor map): one instruction moves a whole contiguous block of tile rows
__device__ __forceinline__ void bulk_store(void* gdst, const void* ssrc, unsigned bytes) {
  asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(gdst), "r"((unsigned)__cvta_generic_to_shared(ssrc)), "r"(bytes)
               : "memory");
}
__device__ __forceinline__ void bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
__device__ __forceinline__ void bulk_wait_read() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }   // sources may be overwritten
__device__ __forceinline__ void bulk_wait_all() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }         // writes performed
__device__ __forceinline__ void fence_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }  // generic-proxy writes -> async proxy
#else
__device__ __forceinline__ void bulk_store(void* gdst, const void* ssrc, unsigned bytes) { memcpy(gdst, ssrc, bytes); }
__device__ __forceinline__ void bulk_commit() {}
__device__ __forceinline__ void bulk_wait_read() {}
__device__ __forceinline__ void bulk_wait_all() {}
__device__ __forceinline__ void fence_async_smem() {}
#endif
#ifndef P3D_EMULATE
__device__ __forceinline__ void prefetch_l2(const void* p) { asm volatile("prefetch.global.L2 [%0];" ::"l"(p)); }
#else
__device__ __forceinline__ void prefetch_l2(const void* p) { (void)*reinterpret_cast<const volatile char*>(p); }   // must be a mapped address
#endif

// ---------------------------------------------------------------------------------------
// row -> address
//
// The address of FFT row `row` of line t of a tile is   tb[run] + off(row) + t * line pitch,   where
// off(row) does not depend on the tile and tb[run] (the run's first row, line 0 of the tile) does not depend
// on the row.  The kernels therefore keep one tile-invariant entry per row in shared memory, built once per
// CTA:  (off(row) << 5) | run,  ROW_NONE for rows that are not stored (read as zero / dropped), and a handful
// of per-run tile bases that a few threads refresh per tile (double buffered by tile parity).
// ---------------------------------------------------------------------------------------
constexpr long long ROW_NONE = -1;

// DSTM (DST-I by odd extension, FFT length nrows = 2 (n + 1), n = nlogical): FFT row r holds logical row r - 1 for
// 1 <= r <= n; rows 0 and n + 1 are zero; on the INPUT side (DSTM = 1) row r > n + 1 mirrors logical row nrows - r - 1
// (the kernel negates it), on the OUTPUT side (DSTM = 2) only rows 1 .. n are stored.
template <int NT, int ESZ, int DSTM = 0>
__device__ __forceinline__ void build_rowent(const FastSide& sd, long long* ent, int nrows, int nlogical, int mirror_nfft) {
  for (int row = threadIdx.x; row < nrows; row += NT) {
    int k = row;
    if constexpr (DSTM != 0) {
      if (row >= 1 && row <= nlogical) k = row - 1;
      else if (DSTM == 1 && row > nlogical + 1) k = nrows - row - 1;
      else k = -1;
    } else if (mirror_nfft && row >= nlogical) k = mirror_nfft - row;
    long long e = ROW_NONE;
    for (int i = 0; i < sd.nrun; i++) {
      const FastRun& r = sd.run[i];
      if (k >= r.kstart && k < r.kstart + r.len) {
        const int j = k - r.korg;
        const long long ro = r.kw > 1 ? (long long)(j / r.kw) * r.psh + (long long)(j % r.kw) * r.ps : (long long)j * r.ps;
        e = ((ro * ESZ) << 5) | i;
      }
    }
    ent[row] = e;
  }
}
__device__ __forceinline__ char* row_addr(long long e, char* const* tb) { return tb[(int)e & 31] + (e >> 5); }

// Tile order: a-tiles fastest, then b, then c -- except that bord consecutive b's are innermost when the
// planner asks for it (stage.h, bord).  b >= nb marks a padding slot of the last b-block (no work).
// (32-bit arithmetic: the host falls back to the generic kernel for >= 2^31 tiles)
struct TileIdx { int ta, b, c; };
__device__ __forceinline__ TileIdx tile_decode(unsigned tile, unsigned tiles_a, unsigned nb, unsigned bord) {
  TileIdx x;
  if (bord <= 1) {
    x.ta = (int)(tile % tiles_a);
    const unsigned r = tile / tiles_a;
    x.b = (int)(r % nb);
    x.c = (int)(r / nb);
  } else {
    const unsigned nbb = (nb + bord - 1) / bord;
    const unsigned bl = tile % bord;
    unsigned r = tile / bord;
    x.ta = (int)(r % tiles_a);
    r /= tiles_a;
    x.b = (int)((r % nbb) * bord + bl);
    x.c = (int)(r / nbb);
  }
  return x;
}
__host__ __device__ inline long long tile_count(int tiles_a, int nb, int nc, int bord) {
  const long long nbp = bord <= 1 ? nb : (long long)((nb + bord - 1) / bord) * bord;
  return (long long)tiles_a * nbp * nc;
}

// tile bases of every run of one side: tb[g] = run base + tile offset
template <int ESZ>
__device__ __forceinline__ char* tile_base(const FastRun& r, TileIdx ti) {
  const long long bo = r.bw > 1 ? (long long)(ti.b / r.bw) * r.sbh + (long long)(ti.b % r.bw) * r.sb : (long long)ti.b * r.sb;
  return (char*)r.base + ((long long)ti.ta * r.sat + bo + (long long)ti.c * r.sc) * ESZ;
}

struct RunTab {
  char* tb[2][2][P3D_MAXRUN];            // [tile parity][side][run]
  unsigned char pfmode[P3D_MAXRUN];      // input runs: 0 no L2 prefetch, 1 every row, 2 even rows only (64-byte contiguous rows)
};

template <int NT, int ESZ>
__device__ __forceinline__ void fill_tilebase(const FastStage& st, RunTab& rt, int slot, TileIdx ti) {
  const int nin = st.in.nrun, ntot = nin + st.out.nrun;
  for (int i = threadIdx.x; i < ntot; i += NT) {
    const int side = i >= nin, g = side ? i - nin : i;
    rt.tb[slot][side][g] = tile_base<ESZ>(side ? st.out.run[g] : st.in.run[g], ti);
  }
}

// SCALED: every output is multiplied by st.scale on its way out (the drivers' normalisation pass, mult_array in
// sample/C/driver_*.c, fused into the store); a separate instantiation, so the unscaled kernels are unchanged.
// C: the configuration (schedule, threads); the default is the table above, CCfgR32 selects the two-pass variant
// BULK (opt-in, P3DFFT_B200_BULK=1; 128-byte rows, outputs whose tile rows are contiguous per run -- every stage that
// feeds an exchange): the last pass puts the tile back into shared memory in natural row order and ONE bulk asynchronous
// copy per output run (per peer) moves it to HBM or over NVLink, instead of 16-byte stores from registers.
// DST (P3D_DST1, exec_strans_r2_complex_same, fft_exec.F90:866-921): the sine transform of n = N/2 - 1 points as an N-point
// FFT of the odd extension (row r = 1 .. n holds input r - 1, rows N - r hold its negative, rows 0 and n + 1 are zero) --
// the extension exists only in the row table and one sign at the load; output k is i * W[k + 1].  A separate
// instantiation: the other kernels' code does not change.
template <typename T, int N, int RB, bool SWAP, bool SCALED = false, class C = CCfg<T, N, RB>, bool BULK = false, bool DST = false>
__global__ void __launch_bounds__(C::NT, C::MINB) cstage_kernel(const __grid_constant__ FastStage st) {
  using T2 = typename Cx<T>::type;
  using S = typename C::S;
  constexpr int TX = C::TX, NT = C::NT, L = S::L;
  constexpr bool SWZ = (TX * sizeof(T2) == 64);
  static_assert(TX * sizeof(T2) == 64 || TX * sizeof(T2) == 128, "row must be 64 or 128 bytes");
  static_assert(NT % TX == 0, "t must be constant per thread");
  extern __shared__ __align__(128) unsigned char smem_raw[];
  T2* s = reinterpret_cast<T2*>(smem_raw);
  long long* ent_in = reinterpret_cast<long long*>(smem_raw + sizeof(T2) * N * TX);     // [N]
  long long* ent_out = ent_in + N;                                                       // [N]
  RunTab* rt = reinterpret_cast<RunTab*>(ent_out + N);
  const T2* __restrict__ tw = reinterpret_cast<const T2*>(st.tw);

  const unsigned tiles_a = (st.na + TX - 1) / TX;
  const unsigned ntiles = (unsigned)tile_count(tiles_a, st.nb, st.nc, st.bord);
  const int t = threadIdx.x % TX;
  const long long lin = (long long)t * st.in.run[0].sa * (long long)sizeof(T2);     // line offsets inside a tile
  const long long lout = (long long)t * st.out.run[0].sa * (long long)sizeof(T2);

  static_assert(!DST || (!SWAP && !BULK), "the sine transform is its own inverse and stores from registers");
  if constexpr (DST) {
    build_rowent<NT, sizeof(T2), 1>(st.in, ent_in, N, st.n, 0);
    build_rowent<NT, sizeof(T2), 2>(st.out, ent_out, N, st.n, 0);
  } else {
    build_rowent<NT, sizeof(T2)>(st.in, ent_in, N, st.n, st.mirror ? N : 0);
    build_rowent<NT, sizeof(T2)>(st.out, ent_out, N, N, 0);
  }
  for (int g = threadIdx.x; g < st.in.nrun; g += NT) {
    const long long psb = st.in.run[g].ps * (long long)sizeof(T2);
    rt->pfmode[g] = (st.prefetch && psb <= (long long)st.prefetch) ? (psb == 64 ? 2 : 1) : 0;
  }
  if (blockIdx.x < ntiles) {
    fill_tilebase<NT, sizeof(T2)>(st, *rt, 0, tile_decode(blockIdx.x, tiles_a, st.nb, st.bord));
  }
  __syncthreads();

  int slot = 0;
  for (unsigned tile = blockIdx.x; tile < ntiles; tile += gridDim.x, slot ^= 1) {
    const TileIdx ti = tile_decode(tile, tiles_a, st.nb, st.bord);
    const bool live = ti.ta * TX + t < st.na && ti.b < st.nb;       // b >= nb: padding slot, nothing loaded or stored
    const bool has_next = tile + gridDim.x < ntiles && tile + gridDim.x > tile;
    if (has_next && threadIdx.x < st.in.nrun + st.out.nrun) {
      fill_tilebase<NT, sizeof(T2)>(st, *rt, slot ^ 1, tile_decode(tile + gridDim.x, tiles_a, st.nb, st.bord));
    }
    // ---- pass 1: global -> registers -> shared ---------------------------------------------
    {
      constexpr int R = S::r(0), M = S::m(0), ITEMS = M * TX;
      char* const* tbi = rt->tb[slot][0];
#pragma unroll
      for (int w0 = 0; w0 < ITEMS; w0 += NT) {
        const int w = w0 + threadIdx.x;
        if (ITEMS % NT == 0 || w < ITEMS) {
          const int u = w / TX;
          T2 v[R];
#pragma unroll
          for (int p = 0; p < R; p++) {
            const long long e = ent_in[u + p * M];
            v[p] = (live && e >= 0) ? ldg_stream(reinterpret_cast<const T2*>(row_addr(e, tbi) + lin)) : T2{0, 0};
            if (SWAP) v[p] = cswap(v[p]);
            if constexpr (DST) { if (u + p * M > st.n + 1) v[p] = T2{-v[p].x, -v[p].y}; }      // the mirrored half of the odd extension
          }
          Bfly<T, R>::run(v);
          if constexpr (BULK) {
            static_assert(ITEMS % NT == 0, "bulk variant: every thread reaches the barrier");
            if (w0 == 0 && tile != blockIdx.x) {      // the bulk copies of the previous tile still read the tile buffer
              bulk_wait_read();                       // (only the threads that issued them have groups outstanding)
              __syncthreads();
            }
          }
          twiddle_store<T, S, 0, TX, SWZ>(v, s, tw, u, u, t);
        }
      }
    }
    __syncthreads();
    // ---- L2 prefetch of the rows of this CTA's next tile ---------------------------------------
    if (has_next && st.prefetch) {
      char* const* tbn = rt->tb[slot ^ 1][0];
      for (int row = threadIdx.x; row < (DST ? st.n + 1 : st.mirror ? st.n : N); row += NT) {
        const long long e = ent_in[row];
        if (e < 0) continue;
        const int pm = rt->pfmode[(int)e & 31];
        if (pm == 1 || (pm == 2 && !(row & 1))) prefetch_l2(row_addr(e, tbn));      // one request per 128-byte line
      }
    }
    mid_passes<T, S, 1, TX, NT, SWZ>(s, tw);
    // ---- pass L: shared -> registers -> global ----------------------------------------------
    if constexpr (!BULK) {
      constexpr int RL = S::r(L - 1), ML = N / RL, ITEMS = ML * TX;
      char* const* tbo = rt->tb[slot][1];
#pragma unroll 1
      for (int w = threadIdx.x; w < ITEMS; w += NT) {
        const int kappa = w / TX;
        T2 v[RL];
        last_bfly<T, S, TX, SWZ>(s, kappa, t, v);
        if constexpr (SCALED) {
          const T sc = (T)st.scale;
#pragma unroll
          for (int q = 0; q < RL; q++) { v[q].x *= sc; v[q].y *= sc; }
        }
        if constexpr (DST) {
#pragma unroll
          for (int q = 0; q < RL; q++) v[q] = mul_pi(v[q]);      // Y[k] = i * W[k + 1]
        }
#pragma unroll
        for (int q = 0; q < RL; q++) {
          const long long e = ent_out[kappa + q * ML];
          if (live && e >= 0) stg_cs(reinterpret_cast<T2*>(row_addr(e, tbo) + lout), SWAP ? cswap(v[q]) : v[q]);
        }
      }
    } else {
      // ---- pass L, bulk variant: shared -> registers -> shared (natural row order) -> one bulk copy per output run ----
      constexpr int RL = S::r(L - 1), ML = N / RL, ITEMS = ML * TX, IPT = ITEMS / NT;
      static_assert(ITEMS % NT == 0 && !SWZ, "bulk variant: whole items per thread, unswizzled 128-byte rows");
      char* const* tbo = rt->tb[slot][1];
      T2 v[IPT][RL];
#pragma unroll
      for (int it = 0; it < IPT; it++) {
        last_bfly<T, S, TX, SWZ>(s, (it * NT + (int)threadIdx.x) / TX, t, v[it]);
        if constexpr (SCALED) {
          const T sc = (T)st.scale;
#pragma unroll
          for (int q = 0; q < RL; q++) { v[it][q].x *= sc; v[it][q].y *= sc; }
        }
      }
      __syncthreads();      // every butterfly has read its rows: the buffer may be rewritten
#pragma unroll
      for (int it = 0; it < IPT; it++) {
        const int kappa = (it * NT + (int)threadIdx.x) / TX;
#pragma unroll
        for (int q = 0; q < RL; q++) s[(kappa + q * ML) * TX + t] = SWAP ? cswap(v[it][q]) : v[it][q];
      }
      fence_async_smem();
      __syncthreads();
      if ((int)threadIdx.x < st.out.nrun && ti.b < st.nb) {      // lines past na land in the padding lanes of the blocked layouts
        const FastRun& r = st.out.run[threadIdx.x];
        const long long e = ent_out[r.kstart];
        bulk_store(row_addr(e, tbo), s + (size_t)r.kstart * TX, (unsigned)(r.len * TX * (int)sizeof(T2)));
        bulk_commit();
      }
    }
    __syncthreads();      // the tile buffer and the tile bases of this parity are reused
  }
  if constexpr (BULK) bulk_wait_all();
}

// two-pass variant: the same kernel with the CCfgR32 configuration

template <typename T, int N> constexpr size_t cstage_r32_smem() {
  using T2 = typename Cx<T>::type;
  return sizeof(T2) * N * CCfgR32<T, N>::TX + 2 * sizeof(long long) * N + sizeof(RunTab);
}

// ---------------------------------------------------------------------------------------
// c2c stage kernel, split variant: the tile (N rows x 128 bytes) is processed in two halves that share one
// N/2-row buffer, so that TWO CTAs fit on an SM where the whole tile allows only one (N = 1024: 64 KB
// instead of 128 KB) and one CTA's loads overlap the other's arithmetic.  After the first DIF pass (radix
// R0) the R0 sub-transforms of length N/R0 are independent: pass 1 runs on the whole tile straight from
// global memory, the outputs of sub-transforms 0 .. R0/2-1 go to shared memory and are finished first,
// those of R0/2 .. R0-1 wait in registers (IPT * R0/2 complex values per thread) and follow.
// Three-pass schedules, 128-byte rows only.
// ---------------------------------------------------------------------------------------
template <typename T, int N> struct SplitCfg {
  static constexpr int NT = 256, MINB = 2;
};
// 2048 points: 128 KB of shared memory per CTA, hence one CTA per SM and 512 threads (the half tile that waits in registers
// is 4 x 8 single or 2 x 8 double complex values per thread)
template <typename T> struct SplitCfg<T, 2048> {
  static constexpr int NT = 512, MINB = 1;
};

// DST: the sine transform by odd extension, as in cstage_kernel (only the 2048-point instantiation is built: nz = 1023)
template <typename T, int N, bool SWAP, bool SCALED = false, bool DST = false>
__global__ void __launch_bounds__(SplitCfg<T, N>::NT, SplitCfg<T, N>::MINB) cstage_split_kernel(const __grid_constant__ FastStage st) {
  using T2 = typename Cx<T>::type;
  using S = typename CCfg<T, N, 128>::S;
  constexpr int TX = CCfg<T, N, 128>::TX, NT = SplitCfg<T, N>::NT, L = S::L;
  static_assert(L == 3, "split kernel needs a three-pass schedule");
  constexpr int R0 = S::r(0), M0 = S::m(0), R1 = S::r(1), RL = S::r(2), HB = R0 / 2, NH = N / 2, ML = N / RL;
  constexpr int ITEMS1 = M0 * TX, IPT = ITEMS1 / NT;
  static_assert(ITEMS1 % NT == 0 && NT % TX == 0 && R0 % 2 == 0, "split kernel geometry");
  extern __shared__ __align__(128) unsigned char smem_raw[];
  T2* s = reinterpret_cast<T2*>(smem_raw);
  long long* ent_in = reinterpret_cast<long long*>(smem_raw + sizeof(T2) * NH * TX);     // [N]
  long long* ent_out = ent_in + N;                                                        // [N]
  RunTab* rt = reinterpret_cast<RunTab*>(ent_out + N);
  const T2* __restrict__ tw = reinterpret_cast<const T2*>(st.tw);

  const unsigned tiles_a = (st.na + TX - 1) / TX;
  const unsigned ntiles = (unsigned)tile_count(tiles_a, st.nb, st.nc, st.bord);
  const int t = threadIdx.x % TX;
  const long long lin = (long long)t * st.in.run[0].sa * (long long)sizeof(T2);
  const long long lout = (long long)t * st.out.run[0].sa * (long long)sizeof(T2);

  static_assert(!DST || !SWAP, "the sine transform is its own inverse");
  if constexpr (DST) {
    build_rowent<NT, sizeof(T2), 1>(st.in, ent_in, N, st.n, 0);
    build_rowent<NT, sizeof(T2), 2>(st.out, ent_out, N, st.n, 0);
  } else {
    build_rowent<NT, sizeof(T2)>(st.in, ent_in, N, st.n, st.mirror ? N : 0);
    build_rowent<NT, sizeof(T2)>(st.out, ent_out, N, N, 0);
  }
  for (int g = threadIdx.x; g < st.in.nrun; g += NT) {
    const long long psb = st.in.run[g].ps * (long long)sizeof(T2);
    rt->pfmode[g] = (st.prefetch && psb <= (long long)st.prefetch) ? (psb == 64 ? 2 : 1) : 0;
  }
  if (blockIdx.x < ntiles) fill_tilebase<NT, sizeof(T2)>(st, *rt, 0, tile_decode(blockIdx.x, tiles_a, st.nb, st.bord));
  __syncthreads();

  int slot = 0;
  for (unsigned tile = blockIdx.x; tile < ntiles; tile += gridDim.x, slot ^= 1) {
    const TileIdx ti = tile_decode(tile, tiles_a, st.nb, st.bord);
    const bool live = ti.ta * TX + t < st.na && ti.b < st.nb;
    const bool has_next = tile + gridDim.x < ntiles && tile + gridDim.x > tile;
    if (has_next && threadIdx.x < st.in.nrun + st.out.nrun)
      fill_tilebase<NT, sizeof(T2)>(st, *rt, slot ^ 1, tile_decode(tile + gridDim.x, tiles_a, st.nb, st.bord));
    // ---- pass 1 on the whole tile: lower sub-transforms -> shared, upper ones stay in registers ----
    T2 hold[IPT][HB];
    {
      char* const* tbi = rt->tb[slot][0];
#pragma unroll
      for (int it = 0; it < IPT; it++) {
        const int u = (it * NT + (int)threadIdx.x) / TX;
        T2 v[R0];
#pragma unroll
        for (int p = 0; p < R0; p++) {
          const long long e = ent_in[u + p * M0];
          v[p] = (live && e >= 0) ? ldg_stream(reinterpret_cast<const T2*>(row_addr(e, tbi) + lin)) : T2{0, 0};
          if (SWAP) v[p] = cswap(v[p]);
          if constexpr (DST) { if (u + p * M0 > st.n + 1) v[p] = T2{-v[p].x, -v[p].y}; }
        }
        Bfly<T, R0>::run(v);
        const T2* twp = tw + S::twoff(0) + u;
#pragma unroll
        for (int q = 1; q < R0; q++) v[q] = cmul(v[q], __ldg(twp + (q - 1) * M0));
#pragma unroll
        for (int q = 0; q < HB; q++) {
          s[(u + q * M0) * TX + t] = v[q];
          hold[it][q] = v[HB + q];
        }
      }
    }
    __syncthreads();
    // ---- L2 prefetch of the rows of this CTA's next tile ---------------------------------------
    if (has_next && st.prefetch) {
      char* const* tbn = rt->tb[slot ^ 1][0];
      for (int row = threadIdx.x; row < (DST ? st.n + 1 : st.mirror ? st.n : N); row += NT) {
        const long long e = ent_in[row];
        if (e < 0) continue;
        const int pm = rt->pfmode[(int)e & 31];
        if (pm == 1 || (pm == 2 && !(row & 1))) prefetch_l2(row_addr(e, tbn));
      }
    }
#pragma unroll
    for (int half = 0; half < 2; half++) {
      if (half == 1) {
        __syncthreads();                      // the first half has been read out
#pragma unroll
        for (int it = 0; it < IPT; it++) {
          const int u = (it * NT + (int)threadIdx.x) / TX;
#pragma unroll
          for (int q = 0; q < HB; q++) s[(u + q * M0) * TX + t] = hold[it][q];
        }
        __syncthreads();
      }
      mid_pass<T, S, 1, TX, NT, false, NH>(s, tw);
      __syncthreads();
      // ---- pass 3 of this half: physical butterfly kp = q1' + HB*q2  <->  kappa = q1' + HB*half + R0*q2 ----
      {
        constexpr int ITEMS = (NH / RL) * TX;
        char* const* tbo = rt->tb[slot][1];
#pragma unroll 1
        for (int w = threadIdx.x; w < ITEMS; w += NT) {
          const int kp = w / TX, q1 = kp % HB, q2 = kp / HB;
          const int base = (q1 * R1 + q2) * RL, kappa = q1 + HB * half + R0 * q2;
          T2 v[RL];
#pragma unroll
          for (int p = 0; p < RL; p++) v[p] = s[(base + p) * TX + t];
          Bfly<T, RL>::run(v);
          if constexpr (SCALED) {
            const T sc = (T)st.scale;
#pragma unroll
            for (int q = 0; q < RL; q++) { v[q].x *= sc; v[q].y *= sc; }
          }
          if constexpr (DST) {
#pragma unroll
            for (int q = 0; q < RL; q++) v[q] = mul_pi(v[q]);      // Y[k] = i * W[k + 1]
          }
#pragma unroll
          for (int q = 0; q < RL; q++) {
            const long long e = ent_out[kappa + q * ML];
            if (live && e >= 0) stg_cs(reinterpret_cast<T2*>(row_addr(e, tbo) + lout), SWAP ? cswap(v[q]) : v[q]);
          }
        }
      }
    }
    __syncthreads();      // the tile buffer and the tile bases of this parity are reused
  }
}

// ---------------------------------------------------------------------------------------
// c2c stage kernel, asynchronously staged variant for the 1024-point stages (128-byte rows: one tile = 128 KB, so
// only ONE tile fits on an SM and the plain kernel's load, arithmetic and store phases of a tile run one after the
// other).  Here the input rows reach shared memory by cp.async (LDGSTS: no registers held, no warp waiting) while the
// previous tile is transformed, and the transform is arranged so that the tile can arrive in quarters:
//
//   N = 4 M:   X[k + M m] = sum_q  w_N^(q k) (-i)^(q m)  S_q[k],     S_q = FFT_M of the rows r = q (mod 4)
//
// (decimation in time by 4 on the outside).  Shared memory holds SIX units of M rows (6 x 32 KB): the four quarters of
// the tile being transformed and two quarters of the next one.  Per tile:
//   wait Q0,Q1 -> passes A, B (the M = 16 x 16 sub-transforms, in place) on them
//   wait Q2,Q3 -> passes A, B on them                      (their loads were issued after the previous tile's pass C)
//   pass C: radix-4 combine with the outer twiddles, registers -> global (one HBM write per element)
//   issue the loads of Q2,Q3 of the next tile and Q0,Q1 of the one after it into the four units just freed
// so about one tile (128 KB) of loads is in flight per SM all the time, and every quarter has at least the time of
// half a tile's arithmetic to land.  Shared-memory traffic equals the three-pass kernel's (three writes, three reads
// per element).  Twiddle block: the pass table of RS<16,16>, then wo[(q-1) M + k] = exp(-2 pi i q k / N), q = 1..3.
// ---------------------------------------------------------------------------------------
#ifndef P3D_EMULATE
__device__ __forceinline__ void cp_async16(void* sdst, const void* gsrc, bool valid) {      // !valid: 16 zero bytes
  const unsigned sz = valid ? 16u : 0u;
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"((unsigned)__cvta_generic_to_shared(sdst)), "l"(gsrc), "r"(sz) : "memory");
}
__device__ __forceinline__ void cp_async8(void* sdst, const void* gsrc, bool valid) {
  const unsigned sz = valid ? 8u : 0u;
  asm volatile("cp.async.ca.shared.global [%0], [%1], 8, %2;" ::"r"((unsigned)__cvta_generic_to_shared(sdst)), "l"(gsrc), "r"(sz) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int NPEND> __device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(NPEND) : "memory"); }
#else
__device__ __forceinline__ void cp_async16(void* sdst, const void* gsrc, bool valid) { if (valid) memcpy(sdst, gsrc, 16); else memset(sdst, 0, 16); }
__device__ __forceinline__ void cp_async8(void* sdst, const void* gsrc, bool valid) { if (valid) memcpy(sdst, gsrc, 8); else memset(sdst, 0, 8); }
__device__ __forceinline__ void cp_async_commit() {}
template <int NPEND> __device__ __forceinline__ void cp_async_wait() {}
#endif

template <typename T> struct ACfg {
  using S = RS<16, 16>;                          // schedule of the four M-point sub-transforms
  static constexpr int N = 1024, Q = 4, M = N / Q, TX = 128 / (2 * (int)sizeof(T)), NT = 256, UNITS = 6;
};
constexpr bool acfg_exists(int n) { return n == 1024; }

struct RunTab3 {
  char* tb[3][2][P3D_MAXRUN];            // [tile counter mod 3][side][run]
};

template <typename T, bool SWAP, bool SCALED = false>
__global__ void __launch_bounds__(ACfg<T>::NT, 1) cstage_async_kernel(const __grid_constant__ FastStage st) {
  using T2 = typename Cx<T>::type;
  using C = ACfg<T>;
  using S = typename C::S;
  constexpr int N = C::N, Q = C::Q, M = C::M, TX = C::TX, NT = C::NT, UNITS = C::UNITS;
  constexpr int RA = S::r(0), MA = S::m(0), RB = S::r(1);      // inner passes: RA x RB = M
  static_assert(S::L == 2 && RA * RB == M && Q == 4, "async kernel: two inner passes, radix-4 outside");
  extern __shared__ __align__(128) unsigned char smem_raw[];
  T2* s = reinterpret_cast<T2*>(smem_raw);                                                   // [UNITS][M][TX]
  long long* ent_in = reinterpret_cast<long long*>(smem_raw + sizeof(T2) * UNITS * M * TX);  // [N]
  long long* ent_out = ent_in + N;                                                           // [N]
  RunTab3* rt = reinterpret_cast<RunTab3*>(ent_out + N);
  const T2* __restrict__ tw = reinterpret_cast<const T2*>(st.tw);
  const T2* __restrict__ wo = tw + S::twtotal();

  const unsigned tiles_a = (st.na + TX - 1) / TX;
  const unsigned ntiles = (unsigned)tile_count(tiles_a, st.nb, st.nc, st.bord);
  if (blockIdx.x >= ntiles) return;
  const int t = threadIdx.x % TX;
  const long long lin = (long long)t * st.in.run[0].sa * (long long)sizeof(T2);     // line offsets inside a tile
  const long long lout = (long long)t * st.out.run[0].sa * (long long)sizeof(T2);
  const int nrun_all = st.in.nrun + st.out.nrun;

  build_rowent<NT, sizeof(T2)>(st.in, ent_in, N, st.n, st.mirror ? N : 0);
  build_rowent<NT, sizeof(T2)>(st.out, ent_out, N, N, 0);
  auto fill_tb = [&](unsigned tile, int slot) {
    if ((int)threadIdx.x < nrun_all) {
      const int side = (int)threadIdx.x >= st.in.nrun, g = side ? (int)threadIdx.x - st.in.nrun : (int)threadIdx.x;
      rt->tb[slot][side][g] = tile_base<sizeof(T2)>(side ? st.out.run[g] : st.in.run[g], tile_decode(tile, tiles_a, st.nb, st.bord));
    }
  };
  auto unit_of = [&](int slot, int q) -> T2* { return s + (size_t)((4 * slot + q) % UNITS) * M * TX; };      // slot = tile counter mod 3
  // loads of quarters 2 qp and 2 qp + 1 of a tile (one commit group; an empty group when the tile does not exist)
  auto issue = [&](bool exists, unsigned tile, int slot, int qp) {
    if (exists) {
      const TileIdx ti = tile_decode(tile, tiles_a, st.nb, st.bord);
      const bool live = ti.ta * TX + t < st.na && ti.b < st.nb;
      char* const* tbi = rt->tb[slot][0];
#pragma unroll 4
      for (int w = threadIdx.x; w < 2 * M * TX; w += NT) {
        const int rr = (w / TX) % M, q = 2 * qp + w / (TX * M);
        const long long e = ent_in[Q * rr + q];
        const bool ok = live && e >= 0;
        const char* src = ok ? row_addr(e, tbi) + lin : reinterpret_cast<const char*>(tw);
        T2* dst = unit_of(slot, q) + rr * TX + t;
        if constexpr (sizeof(T2) == 16) cp_async16(dst, src, ok);
        else cp_async8(dst, src, ok);
      }
    }
    cp_async_commit();
  };
  // passes A and B of the sub-transforms of quarters 2 qp, 2 qp + 1 (in place)
  auto inner = [&](int slot, int qp) {
#pragma unroll 1
    for (int w = threadIdx.x; w < 2 * MA * TX; w += NT) {
      const int u = (w / TX) % MA;
      T2* U = unit_of(slot, 2 * qp + w / (TX * MA));
      T2 v[RA];
#pragma unroll
      for (int p = 0; p < RA; p++) { v[p] = U[(u + p * MA) * TX + t]; if (SWAP) v[p] = cswap(v[p]); }
      Bfly<T, RA>::run(v);
      twiddle_store<T, S, 0, TX, false>(v, U, tw, u, u, t);
    }
    __syncthreads();
#pragma unroll 1
    for (int w = threadIdx.x; w < 2 * (M / RB) * TX; w += NT) {
      const int kap = (w / TX) % (M / RB);
      T2* U = unit_of(slot, 2 * qp + w / (TX * (M / RB))) + (size_t)kap * RB * TX + t;
      T2 v[RB];
#pragma unroll
      for (int p = 0; p < RB; p++) v[p] = U[p * TX];
      Bfly<T, RB>::run(v);                                    // v[j] = S_q[kap + (M / RB) j], kept at row kap * RB + j
#pragma unroll
      for (int p = 0; p < RB; p++) U[p * TX] = v[p];
    }
  };

  const unsigned first = blockIdx.x, step = gridDim.x;
  fill_tb(first, 0);
  if (first + step < ntiles && first + step > first) fill_tb(first + step, 1);
  __syncthreads();
  issue(true, first, 0, 0);
  issue(true, first, 0, 1);
  issue(first + step < ntiles && first + step > first, first + step, 1, 0);

  int slot = 0;
  for (unsigned tile = first;; slot = (slot + 1) % 3) {
    const unsigned t1 = tile + step, t2 = tile + 2 * step;
    const bool has1 = t1 < ntiles && t1 > tile, has2 = has1 && t2 < ntiles && t2 > t1;
    const TileIdx ti = tile_decode(tile, tiles_a, st.nb, st.bord);
    const bool live = ti.ta * TX + t < st.na && ti.b < st.nb;
    if (has2) fill_tb(t2, (slot + 2) % 3);        // read by issue() below, behind several barriers
    cp_async_wait<2>();
    __syncthreads();                              // quarters 0, 1 have landed (every thread's copies)
    inner(slot, 0);
    cp_async_wait<1>();
    __syncthreads();                              // quarters 2, 3 have landed; passes A, B of 0, 1 are complete
    inner(slot, 1);
    __syncthreads();
    // ---- pass C: X[k + M m] = sum_q wo_q[k] (-i)^(q m) S_q[k];  S_q[k] sits at row (k % (M/RB)) * RB + k / (M/RB) of unit q
    {
      char* const* tbo = rt->tb[slot][1];
      const T2* U0 = unit_of(slot, 0); const T2* U1 = unit_of(slot, 1); const T2* U2 = unit_of(slot, 2); const T2* U3 = unit_of(slot, 3);
#pragma unroll 2
      for (int w = threadIdx.x; w < M * TX; w += NT) {
        const int k = w / TX;
        const int pos = ((k % (M / RB)) * RB + k / (M / RB)) * TX + t;
        T2 a0 = U0[pos];
        T2 a1 = cmul(U1[pos], __ldg(wo + k));
        T2 a2 = cmul(U2[pos], __ldg(wo + M + k));
        T2 a3 = cmul(U3[pos], __ldg(wo + 2 * M + k));
        bf4(a0, a1, a2, a3);
        T2 o[4] = {a0, a1, a2, a3};
#pragma unroll
        for (int m = 0; m < 4; m++) {
          if constexpr (SCALED) { const T sc = (T)st.scale; o[m].x *= sc; o[m].y *= sc; }
          const long long e = ent_out[k + m * M];
          if (live && e >= 0) stg_cs(reinterpret_cast<T2*>(row_addr(e, tbo) + lout), SWAP ? cswap(o[m]) : o[m]);
        }
      }
    }
    __syncthreads();                              // the four units of this tile are free
    issue(has1, t1, (slot + 1) % 3, 1);
    issue(has2, t2, (slot + 2) % 3, 0);
    if (!has1) break;
    tile = t1;
  }
  cp_async_wait<0>();
}

template <typename T> constexpr size_t cstage_async_smem() {
  using T2 = typename Cx<T>::type;
  return sizeof(T2) * ACfg<T>::UNITS * ACfg<T>::M * ACfg<T>::TX + 2 * sizeof(long long) * ACfg<T>::N + sizeof(RunTab3);
}

template <typename T, int N> constexpr size_t cstage_split_smem() {
  using T2 = typename Cx<T>::type;
  return sizeof(T2) * (N / 2) * CCfg<T, N, 128>::TX + 2 * sizeof(long long) * N + sizeof(RunTab);
}

template <typename T, int N, int RB> constexpr size_t cstage_smem() {
  using T2 = typename Cx<T>::type;
  return sizeof(T2) * N * CCfg<T, N, RB>::TX + 2 * sizeof(long long) * N + sizeof(RunTab);
}

// ---------------------------------------------------------------------------------------
// X stage: line-major shared-memory tile
// ---------------------------------------------------------------------------------------
template <class C> __device__ __forceinline__ int xs_idx(int t, int e) {
  int x = e >> C::SA;
  if constexpr (C::SB > 0) x ^= e >> C::SB;
  return t * C::LP + (e ^ (x & 7));
}

template <typename T, class C, int PI>
__device__ __forceinline__ void xtwiddle_store(typename Cx<T>::type* v, typename Cx<T>::type* s,
                                               const typename Cx<T>::type* __restrict__ tw, int t, int base, int j) {
  using T2 = typename Cx<T>::type;
  using S = typename C::S;
  constexpr int R = S::r(PI), M = S::m(PI);
  const T2* twp = tw + S::twoff(PI) + j;
#pragma unroll
  for (int q = 1; q < R; q++) v[q] = cmul(v[q], __ldg(twp + (q - 1) * M));
#pragma unroll
  for (int q = 0; q < R; q++) s[xs_idx<C>(t, base + q * M)] = v[q];
}

template <typename T, class C, int PI>
__device__ __forceinline__ void xmid_pass(typename Cx<T>::type* s, const typename Cx<T>::type* __restrict__ tw) {
  using T2 = typename Cx<T>::type;
  using S = typename C::S;
  constexpr int R = S::r(PI), NCUR = S::ncur(PI), M = S::m(PI), PER = C::H / R, ITEMS = PER * C::TX;
#pragma unroll 1
  for (int w = threadIdx.x; w < ITEMS; w += C::NT) {
    const int t = w / PER, rest = w % PER;
    const int blk = rest / M, j = rest % M, base = blk * NCUR + j;
    T2 v[R];
#pragma unroll
    for (int p = 0; p < R; p++) v[p] = s[xs_idx<C>(t, base + p * M)];
    Bfly<T, R>::run(v);
    xtwiddle_store<T, C, PI>(v, s, tw, t, base, j);
  }
}

template <typename T, class C, int PI>
__device__ __forceinline__ void xmid_passes(typename Cx<T>::type* s, const typename Cx<T>::type* __restrict__ tw) {
  if constexpr (PI < C::S::L - 1) {
    xmid_pass<T, C, PI>(s, tw);
    __syncthreads();
    xmid_passes<T, C, PI + 1>(s, tw);
  }
}

// v[q] = Z[kappa + q*ML] of line t
template <typename T, class C>
__device__ __forceinline__ void xlast_bfly(const typename Cx<T>::type* s, int t, int kappa, typename Cx<T>::type* v) {
  using S = typename C::S;
  constexpr int RL = S::r(S::L - 1);
  const int base = blk_of_kappa<S>(kappa) * RL;
#pragma unroll
  for (int p = 0; p < RL; p++) v[p] = s[xs_idx<C>(t, base + p)];
  Bfly<T, RL>::run(v);
}

// Lanes of the (k, H-k) pair pass: a warp takes 32 consecutive k of ONE line (fewer when the pass has fewer
// pairs), the lines of the tile go to neighbouring warps.  The k side then touches whole 128-byte rows of the
// [xb][y][xi] buffer and the H-k side, which is shifted by one element against the row grid, whole rows for
// three quarters of its accesses -- full rows are what NVLink peer stores and the L1 want.
// returns log2(lanes along k)
__device__ __forceinline__ int pair_lanes_log(int half) {
  int il = 32;
  while (il > half || half % il) il >>= 1;      // a power of two that divides the pair count (24 pairs: 8 lanes along k)
  return 31 - __clz(il);
}

// ---------------------------------------------------------------------------------------
// X stage, forward: real line (N = 2H) -> H+1 complex, exec_f_r2c (fft_exec.F90:495)
// twiddle block: pass tables of the H-point schedule, then wx[k] = exp(-2 pi i k / N), k < H
// ---------------------------------------------------------------------------------------
// dynamic shared memory of an X-stage CTA without the staging buffer (16-byte aligned)
template <typename T, int H, class C = XCfg<T, H>> constexpr size_t xstage_smem_base() {
  using T2 = typename Cx<T>::type;
  return (sizeof(T2) * C::LP * C::TX + (sizeof(long long) + sizeof(char*)) * (H + 1) + sizeof(char*) * 2 * P3D_MAXRUN +
          sizeof(int) * (H + 2) + 15) / 16 * 16;
}
// E = (Zk + conj Zm)/2, O = -(i/2)(Zk - conj Zm);  X[k] = E + w O,  X[H-k] = conj(E - w O)
template <typename T, typename T2>
__device__ __forceinline__ void r2c_combine(T2 zk, T2 zm, T2 w, T2& xk, T2& xm) {
  const T hf = (T)0.5;
  T2 e = T2{(zk.x + zm.x) * hf, (zk.y - zm.y) * hf};
  T2 o = T2{(zk.y + zm.y) * hf, (zm.x - zk.x) * hf};
  T2 wo = cmul(w, o);
  xk = cadd(e, wo);
  xm = cconj(csub(e, wo));
}

// STAGED (stages whose output goes to a peer over NVLink): the last pass puts the H+1 outputs of every line into a second
// shared-memory buffer in natural order and the CTA then stores whole 128-byte rows of the blocked X<->Y buffer.  Straight
// from registers the mirrored half (H - k) of every pair is shifted by one element against the row grid, i.e. two partial
// rows (112 + 16 bytes) per quarter warp -- two NVLink packets where one would do (measured 2x4, 1024^3: the X stage moved
// 542 GB/s to its row peer where the Y and Z stages reach 630-640 GB/s).
template <typename T, int H, bool STAGED = false, class C = XCfg<T, H>>
__global__ void __launch_bounds__(C::NT, C::MINB) xr2c_kernel(const __grid_constant__ FastStage st) {
  using T2 = typename Cx<T>::type;
  using S = typename C::S;
  constexpr int TX = C::TX, NT = C::NT, L = S::L;
  extern __shared__ __align__(128) unsigned char smem_raw[];
  T2* s = reinterpret_cast<T2*>(smem_raw);
  long long* ent_out = reinterpret_cast<long long*>(smem_raw + sizeof(T2) * C::LP * TX);      // [H+1] output rows (tile invariant)
  char** rowptr = reinterpret_cast<char**>(ent_out + H + 1);                                   // [H+1] row addresses of this tile
  char** tbs = rowptr + H + 1;                                                                 // [tile parity][run] tile bases
  constexpr int LPO = H + 8;                                                                   // line pitch of the staging buffer
  T2* so = reinterpret_cast<T2*>(smem_raw + xstage_smem_base<T, H, C>());                      // [TX][LPO], STAGED only
  const T2* __restrict__ tw = reinterpret_cast<const T2*>(st.tw);
  const T2* __restrict__ wx = tw + S::twtotal();

  const unsigned tiles_a = (st.na + TX - 1) / TX;
  const unsigned ntiles = (unsigned)tile_count(tiles_a, st.nb, st.nc, st.bord);
  const FastRun& rin = st.in.run[0];
  const long long sao = st.out.run[0].sa * (long long)sizeof(T2);
  constexpr int RL = S::r(L - 1), ML = H / RL;
  const int ilog = pair_lanes_log(ML / 2);

  build_rowent<NT, sizeof(T2)>(st.out, ent_out, H + 1, H + 1, 0);
  if (blockIdx.x < ntiles && threadIdx.x < st.out.nrun)
    tbs[threadIdx.x] = tile_base<sizeof(T2)>(st.out.run[threadIdx.x], tile_decode(blockIdx.x, tiles_a, st.nb, st.bord));
  __syncthreads();

  int slot = 0;
  for (unsigned tile = blockIdx.x; tile < ntiles; tile += gridDim.x, slot ^= 1) {
    const TileIdx ti = tile_decode(tile, tiles_a, st.nb, st.bord);
    const T* tbase = reinterpret_cast<const T*>(rin.base) + (long long)ti.b * rin.sb + (long long)ti.c * rin.sc;
    const unsigned nxt = tile + gridDim.x;
    const bool has_next = nxt < ntiles && nxt > tile;
    // ---- output row addresses of this tile (read after the next barrier) from the per-run tile bases;
    //      tile bases and L2 prefetch of the real lines of this CTA's next tile
    {
      char* const* tbo = tbs + slot * P3D_MAXRUN;
      for (int row = threadIdx.x; row <= H; row += NT) {
        const long long e = ent_out[row];
        rowptr[row] = e >= 0 ? row_addr(e, tbo) : nullptr;
      }
    }
    if (has_next) {
      const TileIdx tn = tile_decode(nxt, tiles_a, st.nb, st.bord);
      if (threadIdx.x < st.out.nrun) tbs[(slot ^ 1) * P3D_MAXRUN + threadIdx.x] = tile_base<sizeof(T2)>(st.out.run[threadIdx.x], tn);
      if (st.prefetch) {
        constexpr int PER_LINE = (int)(H * sizeof(T2) / 128);          // 128-byte lines per real line
        const char* nb = reinterpret_cast<const char*>(reinterpret_cast<const T*>(rin.base) + (long long)(tn.ta * TX) * rin.sa +
                                                       (long long)tn.b * rin.sb + (long long)tn.c * rin.sc);
        for (int i = threadIdx.x; i < PER_LINE * TX; i += NT) {
          const int l = i / PER_LINE, j = i % PER_LINE;
          if (tn.ta * TX + l < st.na) prefetch_l2(nb + (long long)l * rin.sa * (long long)sizeof(T) + j * 128);
        }
      }
    }
    // ---- pass 1: packed real pairs -> registers -> shared (a warp reads 512 contiguous bytes) ---
    {
      constexpr int R = S::r(0), M = S::m(0), ITEMS = M * TX;
#pragma unroll
      for (int w0 = 0; w0 < ITEMS; w0 += NT) {
        const int w = w0 + threadIdx.x;
        if (ITEMS % NT == 0 || w < ITEMS) {
          const int t = w / M, u = w % M;
          const bool live = ti.ta * TX + t < st.na;
          const T2* line = reinterpret_cast<const T2*>(tbase + (long long)(ti.ta * TX + t) * rin.sa);
          T2 v[R];
#pragma unroll
          for (int p = 0; p < R; p++) v[p] = live ? ldg_stream(line + u + p * M) : T2{0, 0};
          Bfly<T, R>::run(v);
          xtwiddle_store<T, C, 0>(v, s, tw, t, u, u);
        }
      }
    }
    __syncthreads();
    xmid_passes<T, C, 1>(s, tw);
    // ---- pass L on butterfly pairs (kappa, ML-kappa) + Hermitian post-processing -------------
    {
      constexpr int ITEMS = (ML / 2) * TX;
      static_assert(ML >= 2, "last pass needs at least two butterflies per line");
#pragma unroll 1
      for (int w = threadIdx.x; w < ITEMS; w += NT) {
        const int t = (w >> ilog) % TX, i = (w & ((1 << ilog) - 1)) + (((w >> ilog) / TX) << ilog);
        const bool live = ti.ta * TX + t < st.na;
        const int lout = t * (int)sao;
        auto put = [&](int k, T2 v) {
          if constexpr (STAGED) so[t * LPO + k] = v;
          else {
            char* rp = rowptr[k];
            if (live && rp) stg_stream(reinterpret_cast<T2*>(rp + lout), v);
          }
        };
        // Item i pairs butterflies (i, ML - i).  Butterflies 0 and ML/2 pair with THEMSELVES; the lane with
        // i == 0 takes both and runs the same instruction stream with other operands (selects, no branch):
        //   slots q <  RL/2: (zb[q], zb[RL-1-q])        k = ML/2 + q*ML          (butterfly ML/2)
        //   slots q >= RL/2: (za[j], za[RL-j]), j = q - (RL/2 - 1) = 1..RL/2,  k = j*ML   (butterfly 0)
        // plus the purely real k = 0 and k = H terms.
        const bool sp = i == 0;
        T2 za[RL], zb[RL];
        xlast_bfly<T, C>(s, t, i, za);
        xlast_bfly<T, C>(s, t, sp ? ML / 2 : ML - i, zb);
#pragma unroll
        for (int q = 0; q < RL; q++) {
          constexpr int OFF = RL / 2 - 1;
          const int j = q - OFF;                            // only used for q >= RL/2
          T2 A, B;
          int k;
          if (q < RL / 2) {
            A = sel(sp, zb[q], za[q]);
            B = zb[RL - 1 - q];
            k = (sp ? ML / 2 : i) + q * ML;
          } else {
            A = sel(sp, za[j], za[q]);
            B = sel(sp, za[RL - j], zb[RL - 1 - q]);
            k = sp ? j * ML : i + q * ML;
          }
          T2 xk, xm;
          r2c_combine<T>(A, B, __ldg(wx + k), xk, xm);
          put(k, xk);
          put(H - k, xm);
        }
        if (sp) {
          put(0, T2{za[0].x + za[0].y, 0});
          put(H, T2{za[0].x - za[0].y, 0});
        }
      }
      if constexpr (STAGED) {
        // whole rows: WB consecutive k of one line are one 128-byte row, the TX lines of a row group are adjacent in memory
        constexpr int WB = 128 / (int)sizeof(T2), NG = (H + 1 + WB - 1) / WB;
        __syncthreads();
#pragma unroll 2
        for (int idx = threadIdx.x; idx < NG * TX * WB; idx += NT) {
          const int xi = idx % WB, t = (idx / WB) % TX, k = (idx / (WB * TX)) * WB + xi;
          if (k <= H && ti.ta * TX + t < st.na) {
            char* rp = rowptr[k];
            if (rp) stg_stream(reinterpret_cast<T2*>(rp + (long long)t * sao), so[t * LPO + k]);
          }
        }
      }
    }
    __syncthreads();      // tile buffer and row table are reused by the next tile
  }
}

template <typename T, int H, class C = XCfg<T, H>> constexpr size_t xstage_smem(bool staged = false) {
  using T2 = typename Cx<T>::type;
  return xstage_smem_base<T, H, C>() + (staged ? sizeof(T2) * (H + 8) * C::TX : 0);
}

// ---------------------------------------------------------------------------------------
// X stage, backward: H+1 complex -> real line (N = 2H), exec_b_c2r (fft_exec.F90:298)
// E = Xk + conj Xm, O = conj(w)(Xk - conj Xm);  Z[k] = E + i O,  Z[H-k] = conj(E - i O)
// (unnormalised: the line comes out multiplied by N like FFTW's c2r)
// ---------------------------------------------------------------------------------------
template <typename T, typename T2>
__device__ __forceinline__ void c2r_combine(T2 xk, T2 xm, T2 w, T2& zk, T2& zm) {
  T2 e = T2{xk.x + xm.x, xk.y - xm.y};
  T2 d = T2{xk.x - xm.x, xk.y + xm.y};
  T2 o = cmul(cconj(w), d);
  // stored swapped (re <-> im) for the swap-trick inverse FFT
  T2 a = T2{e.x - o.y, e.y + o.x};          // E + iO
  T2 bb = T2{e.x + o.y, -(e.y - o.x)};      // conj(E - iO)
  zk = cswap(a);
  zm = cswap(bb);
}

template <typename T, int H, bool SCALED = false, class C = XCfg<T, H>>
__global__ void __launch_bounds__(C::NT, C::MINB) xc2r_kernel(const __grid_constant__ FastStage st) {
  using T2 = typename Cx<T>::type;
  using S = typename C::S;
  constexpr int TX = C::TX, NT = C::NT, L = S::L;
  extern __shared__ __align__(128) unsigned char smem_raw[];
  T2* s = reinterpret_cast<T2*>(smem_raw);
  long long* ent_in = reinterpret_cast<long long*>(smem_raw + sizeof(T2) * C::LP * TX);      // [H+1] input rows (tile invariant)
  char** rowptr = reinterpret_cast<char**>(ent_in + H + 1);                                   // [H+1] row addresses of this tile
  char** tbs = rowptr + H + 1;                                                                 // [tile parity][run] tile bases
  int* pfrow = reinterpret_cast<int*>(tbs + 2 * P3D_MAXRUN);                                   // [0] = count, then the rows to prefetch
  const T2* __restrict__ tw = reinterpret_cast<const T2*>(st.tw);
  const T2* __restrict__ wx = tw + S::twtotal();

  const unsigned tiles_a = (st.na + TX - 1) / TX;
  const unsigned ntiles = (unsigned)tile_count(tiles_a, st.nb, st.nc, st.bord);
  const long long sab = st.in.run[0].sa * (long long)sizeof(T2);
  const FastRun& ro = st.out.run[0];
  constexpr int R1 = S::r(0), M1 = S::m(0);
  const int ilog = pair_lanes_log(M1 / 2);
  // L2 prefetch of the next tile, one request per 128-byte line: in the blocked [xb][y][xi] layout the TX lines
  // of kw consecutive rows are one contiguous piece of TX * sab bytes, else (plain) every line is contiguous
  // along the rows.  The rows that start such a piece are listed once per CTA.
  const int kw = st.in.run[0].kw;
  const bool pf_blocked = kw > 1 && sab == (long long)sizeof(T2) * kw, pf_plain = !pf_blocked && st.in.run[0].ps == 1;
  const int pf_per = pf_blocked ? (int)((TX * sab + 127) / 128) : TX;

  build_rowent<NT, sizeof(T2)>(st.in, ent_in, H + 1, H + 1, 0);
  if (blockIdx.x < ntiles && threadIdx.x < st.in.nrun)
    tbs[threadIdx.x] = tile_base<sizeof(T2)>(st.in.run[threadIdx.x], tile_decode(blockIdx.x, tiles_a, st.nb, st.bord));
  __syncthreads();
  if (threadIdx.x == 0) {
    int n = 0;
    if (st.prefetch && (pf_blocked || pf_plain)) {
      const int mask = pf_blocked ? kw - 1 : (int)(128 / sizeof(T2)) - 1;
      for (int row = 0; row <= H; row++) {
        const long long e = ent_in[row];
        if (e >= 0 && ((int)((e >> 5) / (long long)sizeof(T2)) & mask) == 0) pfrow[1 + n++] = row;
      }
    }
    pfrow[0] = n;
  }
  __syncthreads();

  int slot = 0;
  for (unsigned tile = blockIdx.x; tile < ntiles; tile += gridDim.x, slot ^= 1) {
    const TileIdx ti = tile_decode(tile, tiles_a, st.nb, st.bord);
    const unsigned nxt = tile + gridDim.x;
    const bool has_next = nxt < ntiles && nxt > tile;
    // ---- input row addresses of this tile from the per-run tile bases --------------------------------
    {
      char* const* tbi = tbs + slot * P3D_MAXRUN;
      for (int row = threadIdx.x; row <= H; row += NT) {
        const long long e = ent_in[row];
        rowptr[row] = e >= 0 ? row_addr(e, tbi) : nullptr;
      }
    }
    __syncthreads();
    // ---- tile bases of the next tile (read after the next barriers) and its L2 prefetch -----------------
    if (has_next) {
      const TileIdx tn = tile_decode(nxt, tiles_a, st.nb, st.bord);
      if (threadIdx.x < st.in.nrun) tbs[(slot ^ 1) * P3D_MAXRUN + threadIdx.x] = tile_base<sizeof(T2)>(st.in.run[threadIdx.x], tn);
      const int npf = pfrow[0];
      for (int i = threadIdx.x; i < npf * pf_per; i += NT) {
        const int j = i % pf_per;
        const long long e = ent_in[pfrow[1 + i / pf_per]];
        char* q = tile_base<sizeof(T2)>(st.in.run[(int)e & 31], tn) + (e >> 5);
        if (pf_blocked) {
          // only the lines of the tile that exist (a partial last tile would otherwise touch memory behind the buffer)
          const long long have = (long long)st.na - (long long)tn.ta * TX;
          if ((long long)j * 128 < (have < TX ? have : TX) * sab) prefetch_l2(q + j * 128);
        } else if (tn.ta * TX + j < st.na) prefetch_l2(q + j * sab);
      }
    }
    // ---- pass 1 on butterfly pairs (u, M-u) with the Hermitian pre-processing ----------------
    {
      constexpr int R = R1, M = M1, ITEMS = (M / 2) * TX;
      static_assert(M >= 2, "first pass needs at least two butterflies per line");
#pragma unroll 1
      for (int w = threadIdx.x; w < ITEMS; w += NT) {
        const int t = (w >> ilog) % TX, i = (w & ((1 << ilog) - 1)) + (((w >> ilog) / TX) << ilog);
        const bool live = ti.ta * TX + t < st.na;
        const int lin = t * (int)sab;
        auto get = [&](int k) -> T2 {
          const char* rp = rowptr[k];
          return (live && rp) ? ldg_stream(reinterpret_cast<const T2*>(rp + lin)) : T2{0, 0};
        };
        // same operand selection as the r2c pair pass: the lane with i == 0 builds butterflies 0 and M/2
        const bool sp = i == 0;
        T2 x0 = T2{0, 0}, xh = T2{0, 0};
        if (sp) { x0 = get(0); xh = get(H); }              // issued first: their latency overlaps the loads below
        T2 xk[R], xm[R];
        int kk[R];
#pragma unroll
        for (int p = 0; p < R; p++) {
          constexpr int OFF = R / 2 - 1;
          kk[p] = p < R / 2 ? (sp ? M / 2 : i) + p * M : (sp ? (p - OFF) * M : i + p * M);
          xk[p] = get(kk[p]);
          xm[p] = get(H - kk[p]);
        }
        T2 zk[R], zm[R];
#pragma unroll
        for (int p = 0; p < R; p++) c2r_combine<T>(xk[p], xm[p], __ldg(wx + kk[p]), zk[p], zm[p]);
        T2 za[R], zb[R];
#pragma unroll
        for (int j = 0; j < R; j++) {
          constexpr int OFF = R / 2 - 1;
          // general: za[j] = zk[j], zb[j] = zm[R-1-j]
          const T2 zs = j == 0 ? cswap(T2{x0.x + xh.x, x0.x - xh.x}) : (j <= R / 2 ? zk[j + OFF] : zm[R + OFF - j]);
          za[j] = sel(sp, zs, zk[j]);
          zb[j] = j < R / 2 ? sel(sp, zk[j], zm[R - 1 - j]) : zm[R - 1 - j];
        }
        const int ua = i, ub = (i == 0) ? M / 2 : M - i;
        Bfly<T, R>::run(za);
        xtwiddle_store<T, C, 0>(za, s, tw, t, ua, ua);
        Bfly<T, R>::run(zb);
        xtwiddle_store<T, C, 0>(zb, s, tw, t, ub, ub);
      }
    }
    __syncthreads();
    xmid_passes<T, C, 1>(s, tw);
    // ---- pass L: shared -> registers -> packed real pairs (a warp writes 512 contiguous bytes) ---
    {
      T* tbase = const_cast<T*>(reinterpret_cast<const T*>(ro.base)) + (long long)ti.b * ro.sb + (long long)ti.c * ro.sc;
      constexpr int RL = S::r(L - 1), ML = H / RL, ITEMS = ML * TX;
#pragma unroll 1
      for (int w = threadIdx.x; w < ITEMS; w += NT) {
        const int t = w / ML, kappa = w % ML;
        T2* line = reinterpret_cast<T2*>(tbase + (long long)(ti.ta * TX + t) * ro.sa);
        T2 v[RL];
        xlast_bfly<T, C>(s, t, kappa, v);
        if constexpr (SCALED) {
          const T sc = (T)st.scale;
#pragma unroll
          for (int q = 0; q < RL; q++) { v[q].x *= sc; v[q].y *= sc; }
        }
        if (ti.ta * TX + t < st.na) {
#pragma unroll
          for (int q = 0; q < RL; q++) stg_cs(line + kappa + q * ML, cswap(v[q]));
        }
      }
    }
    __syncthreads();      // tile buffer and row table are reused by the next tile
  }
}
#endif  // __CUDACC__ || P3D_EMULATE

}  // namespace fast
}  // namespace p3d
