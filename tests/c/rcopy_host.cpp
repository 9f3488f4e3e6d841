// TEST CODE (CPU only): drives the host/device address arithmetic of the real-data copy stages
// (p3dfft_b200/csrc/rcopy.h: rcopy_boxes + rcopy_row, the same functions rcopy_kernel runs on the GPU)
// over host buffers, so that tests/test_rtran_cpu.py can check the P3D_RCOPY plans end to end without a GPU.
// Built by p3dfft_b200/build.py into p3dfft_b200/lib/librcopy_check.so; never linked into the product library.
#include <cstring>

#include "../../p3dfft_b200/csrc/rcopy.h"

extern "C" int rcopy_host_run(const P3dStage* st, int elem_bytes) {
  p3d::RcopyJob job;
  if (!p3d::rcopy_boxes(*st, job, (size_t)elem_bytes)) return -1;
  for (int i = 0; i < job.nbox; i++) {
    const p3d::RcopyBox& bx = job.box[i];
    const long long rows = (long long)bx.nv * job.nb * job.nc;
    for (long long row = 0; row < rows; row++) {
      long long so, dof;
      p3d::rcopy_row(bx, job.nb, row, &so, &dof);
      for (int u = 0; u < bx.nu; u++)
        memcpy((char*)bx.dst + (dof + (long long)u * bx.su_out) * elem_bytes,
               (const char*)bx.src + (so + (long long)u * bx.su_in) * elem_bytes, (size_t)elem_bytes);
    }
  }
  return job.nbox;
}

// number of boxes whose rows are contiguous on both sides (su == 1): the coalescing the kernel relies on
extern "C" int rcopy_host_contiguous_boxes(const P3dStage* st, int elem_bytes) {
  p3d::RcopyJob job;
  if (!p3d::rcopy_boxes(*st, job, (size_t)elem_bytes)) return -1;
  int n = 0;
  for (int i = 0; i < job.nbox; i++) n += job.box[i].su_in == 1 && job.box[i].su_out == 1;
  return n;
}
