// Stage descriptors shared by the host planner and the CUDA kernels.
//
// One "stage" = one batched 1D transform along one axis of a pencil, reading every input
// element exactly once from HBM and writing every output element exactly once.  The
// pack/unpack copies the reference performs around its MPI_Alltoallv calls
// (build/fcomm1.F90:239-253,284-320; fcomm2.F90:321-386; bcomm1.F90:309-380;
// bcomm2.F90:233-313) and its pruning / zero-padding copies (module.F90:427-706
// seg_copy_*/seg_zero_*) are expressed as *addressing*: each side of a stage is a list of
// segments along the transform axis, one per peer block of the exchange buffer.
#pragma once
#include <stdint.h>

#define P3D_MAXSEG 16
#define P3D_MAXFAC 24

enum P3dKind {
  P3D_C2C_FWD = 0,   // exec_f_c1 / exec_f_c2_same   (fft_exec.F90:346,401)
  P3D_C2C_BWD = 1,   // exec_b_c1 / exec_b_c2_same   (fft_exec.F90:120,204)
  P3D_R2C = 2,       // exec_f_r2c                   (fft_exec.F90:495)
  P3D_C2R = 3,       // exec_b_c2r                   (fft_exec.F90:298)
  P3D_DCT1 = 4,      // exec_ctrans_r2_complex_same  (fft_exec.F90:646)
  P3D_DST1 = 5,      // exec_strans_r2_complex_same  (fft_exec.F90:866)
  P3D_NOOP = 6,      // op letter 'n' / '0': layout change and pruning only
  P3D_RCOPY = 7      // REAL elements on both sides, no arithmetic: the pack / unpack passes of the real-data
                     // transposes rtran_x2y / y2x / x2z / z2x (module.F90:1061-1361), run by rcopy_kernel (rcopy.h)
};

// buffer ids used by the planner; resolved to pointers at execution time
enum P3dBuf { P3D_BUF_USER_IN = 0, P3D_BUF_USER_OUT = 1, P3D_BUF_A = 2, P3D_BUF_B = 3, P3D_BUF_C = 4 };

// One run of consecutive stored points along the transform axis living in one block.
// Address (in elements of the side's type) of stored point s of line (a,b,c), i = s - start:
//     off + R(i) + A(a) + B(b) + c*sc
//     R(i) = i*ps                       (kw <= 1)     (i / kw)*psh + (i % kw)*ps   (kw > 1)
//     A(a) = a*sa                       (aw <= 1)     (a / aw)*sah + (a % aw)*sa   (aw > 1)
//     B(b) = b*sb                       (bw <= 1)     (b / bw)*sbh + (b % bw)*sb   (bw > 1)
// The two-level forms describe the tile-blocked layouts of the library's own pencil buffers
// (plan.h): aw lines that are adjacent along the contiguous direction form one 64-byte row
// of a kernel tile, and a tile's rows are consecutive in memory.
struct P3dSeg {
  void* base;        // resolved block base (device pointer; may be a peer-mapped pointer)
  int32_t buf;       // P3dBuf the planner refers to
  int32_t peer;      // -1: this rank's memory; >= 0 (peer-to-peer plans): world rank whose buffer `buf` holds it
  int64_t off;       // element offset inside buf
  int32_t start, len;
  int64_t ps, sa, sb, sc;
  int32_t kw, aw;    // block widths along the transform axis / along a (0 or 1: plain strides)
  int64_t psh, sah;  // strides of whole blocks
  int32_t bw, pad_;  // block width along b
  int64_t sbh;
};

// Stored points s in [0,cnt) map to logical indices k in [0,L):
//     k = s            for s <  h1
//     k = s + (L-cnt)  for s >= h1        (logical points not stored are zero / dropped)
struct P3dSide {
  int32_t nseg, cnt, h1, logical;
  P3dSeg seg[P3D_MAXSEG];
};

struct P3dStage {
  int32_t kind;       // P3dKind
  int32_t n;          // logical transform length (nx, ny or nz)
  int32_t nfft;       // length of the complex FFT run on-chip
  int32_t na, nb, nc; // batch extents; CTAs tile dimension a
  int32_t tile;       // lines per CTA
  int32_t layx;       // 1: lines are contiguous in memory (X stage) -> [line][point] smem
  int32_t need_zero;  // smem must be cleared before the load phase
  int32_t nfac;
  int32_t fac[P3D_MAXFAC];
  int32_t timer;      // 1-based slot of the reference's timers(12) this stage books into
  int32_t bord;       // > 1: the input is gathered in rows that tiles adjacent in b share memory lines with ->
                      // run this many consecutive b back to back (tile order hint for the kernels)
  const void* tw;     // device table exp(-2 pi i k / nfft), k < nfft
  double scale;       // multiplies every output
  P3dSide in, out;
};

// alltoallv over the row (comm=0) or column (comm=1) communicator; offsets/counts in elements of
// `ebytes` bytes (0: complex elements of the library's precision).  Mirrors the If/Kf/Jr/Kr tables of
// setup.F90:481-518 and, with real elements, the Ii/Ji/Ij/Kj tables of setup.F90:522-549.
struct P3dExchange {
  int32_t comm, npeer, self;
  int32_t sendbuf, recvbuf;
  int32_t timer;
  int32_t p2p;       // 1: the producing stage already stored every block at its destination; barrier only
  int32_t ebytes;    // bytes per element of the offsets / counts below; 0 = one complex element
  int64_t sndoff[P3D_MAXSEG], sndcnt[P3D_MAXSEG], rcvoff[P3D_MAXSEG], rcvcnt[P3D_MAXSEG];
};
