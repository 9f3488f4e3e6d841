// Real-data sub-box copies: the pack and unpack passes of rtran_x2y / rtran_y2x / rtran_x2z / rtran_z2x
// (build/module.F90:1061-1361) expressed as P3D_RCOPY stages (stage.h).
//
// A P3D_RCOPY stage moves REAL elements from the input segment list to the output segment list without
// arithmetic.  rcopy_boxes() intersects the two lists along the stage axis on the host: every non-empty
// (input segment, output segment) intersection is one BOX, a 4-D strided block whose rows are contiguous
// in memory on both sides (all four transposes keep x fastest).  The kernel then needs no per-element
// search or division: one warp copies one row, lanes along the contiguous direction.
//
// rcopy_boxes() and rcopy_row() are plain host/device functions so that the address arithmetic is
// exercised on a machine without a GPU (tests/c/rcopy_host.cpp drives them over numpy buffers).
#pragma once
#include <stdint.h>

#include "stage.h"

#if defined(__CUDACC__)
#define P3D_HD __host__ __device__ __forceinline__
#else
#define P3D_HD inline
#endif

#define P3D_MAXBOX (2 * P3D_MAXSEG)

namespace p3d {

// rows r = v + nv*(b + nb*c); element u of row r:  src[u*su_in + v*sv_in + b*sb_in + c*sc_in] (elements)
struct RcopyBox {
  const void* src;
  void* dst;
  int32_t nu, nv;
  int64_t su_in, su_out, sv_in, sv_out, sb_in, sb_out, sc_in, sc_out;
};

struct RcopyJob {
  int32_t nbox, nb, nc, pad_;
  long long rows_max;          // largest row count of a box (grid sizing)
  RcopyBox box[P3D_MAXBOX];
};

// Builds the box list of a resolved stage (seg.base set).  esz = bytes per real.  Returns false when the
// stage is not a plain real copy (blocked layouts, pruned sides, more intersections than P3D_MAXBOX).
inline bool rcopy_boxes(const P3dStage& st, RcopyJob& job, size_t esz) {
  job.nbox = 0; job.nb = st.nb; job.nc = st.nc; job.pad_ = 0; job.rows_max = 0;
  if (st.kind != P3D_RCOPY) return false;
  if (st.in.cnt != st.in.logical || st.out.cnt != st.out.logical || st.in.logical != st.out.logical) return false;
  for (int g = 0; g < st.out.nseg; g++) {
    const P3dSeg& og = st.out.seg[g];
    if (og.kw > 1 || og.aw > 1 || og.bw > 1) return false;
    for (int h = 0; h < st.in.nseg; h++) {
      const P3dSeg& ig = st.in.seg[h];
      if (ig.kw > 1 || ig.aw > 1 || ig.bw > 1) return false;
      const int k0 = og.start > ig.start ? og.start : ig.start;
      const int e0 = og.start + og.len, e1 = ig.start + ig.len;
      const int k1 = e0 < e1 ? e0 : e1;
      if (k1 <= k0) continue;
      if (job.nbox >= P3D_MAXBOX) return false;
      RcopyBox& b = job.box[job.nbox++];
      b.src = (const char*)ig.base + (int64_t)(k0 - ig.start) * ig.ps * (int64_t)esz;
      b.dst = (char*)og.base + (int64_t)(k0 - og.start) * og.ps * (int64_t)esz;
      const bool lines_contig = ig.sa == 1 && og.sa == 1 && !(ig.ps == 1 && og.ps == 1);
      if (lines_contig) {       // the lines (a) are adjacent in memory: u = a, v = axis
        b.nu = st.na; b.su_in = ig.sa; b.su_out = og.sa;
        b.nv = k1 - k0; b.sv_in = ig.ps; b.sv_out = og.ps;
      } else {                  // u = axis (contiguous when ps == 1 on both sides), v = a
        b.nu = k1 - k0; b.su_in = ig.ps; b.su_out = og.ps;
        b.nv = st.na; b.sv_in = ig.sa; b.sv_out = og.sa;
      }
      b.sb_in = ig.sb; b.sb_out = og.sb; b.sc_in = ig.sc; b.sc_out = og.sc;
      const long long rows = (long long)b.nv * st.nb * st.nc;
      if (rows > job.rows_max) job.rows_max = rows;
    }
  }
  return true;
}

// element offsets of the first element of row `row` of box `bx`
P3D_HD void rcopy_row(const RcopyBox& bx, int nb, long long row, long long* src_off, long long* dst_off) {
  const long long v = row % bx.nv, r = row / bx.nv;
  const long long b = r % nb, c = r / nb;
  *src_off = v * bx.sv_in + b * bx.sb_in + c * bx.sc_in;
  *dst_off = v * bx.sv_out + b * bx.sb_out + c * bx.sc_out;
}

#if defined(__CUDACC__) || defined(P3D_EMULATE)
// grid = (row groups, boxes); one warp per row at a time, lanes along the contiguous direction u.
template <typename T>
__global__ void __launch_bounds__(256) rcopy_kernel(const __grid_constant__ RcopyJob job) {
  const RcopyBox& bx = job.box[blockIdx.y];
  const long long rows = (long long)bx.nv * job.nb * job.nc;
  const int lane = threadIdx.x & 31;
  const long long warp = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const long long nwarp = (long long)gridDim.x * (blockDim.x >> 5);
  const T* __restrict__ src = reinterpret_cast<const T*>(bx.src);
  T* __restrict__ dst = reinterpret_cast<T*>(bx.dst);
  for (long long row = warp; row < rows; row += nwarp) {
    long long so, dof;
    rcopy_row(bx, job.nb, row, &so, &dof);
    const T* s = src + so;
    T* d = dst + dof;
#pragma unroll 4
    for (int u = lane; u < bx.nu; u += 32) d[(long long)u * bx.su_out] = __ldg(s + (long long)u * bx.su_in);
  }
}
#endif

}  // namespace p3d
