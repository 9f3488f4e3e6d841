// Host-callable launchers of the stage kernels (fft_kernels.cu).
#pragma once
#include <cuda_runtime.h>
#include <stddef.h>

#include "stage.h"

namespace p3d {
template <typename T> cudaError_t launch_stage(const P3dStage& st, cudaStream_t stream);
template <typename T> cudaError_t launch_rcopy(const P3dStage& st, cudaStream_t stream);     // P3D_RCOPY stages (rcopy.h)
template <typename T> int choose_tile(const P3dStage& st);
template <typename T> size_t stage_smem_bytes(const P3dStage& st);
template <typename T> cudaError_t launch_cheby(void* out, long long ncol, int nzc, long long zstride,
                                               long long colstride, double norm, double lfac, cudaStream_t stream);

// Power-spectrum epilogue (sample/C/driver_spec.c:298-384 compute_spectrum, on the device): shell sums
// E[ik] += k2 * |B|^2 * f2 over this rank's wavenumber block, ik = int(sqrt(k2) + 0.5).  Loop dimension 0 is the
// contiguous one; axis[i] names the physical axis (0 x, 1 y, 2 z) of loop dimension i.
struct SpecJob {
  int32_t ext[3];          // loop extents
  int64_t stride[3];       // element strides of the complex array
  int32_t axis[3];
  int32_t start[3];        // per PHYSICAL axis: first stored global index held by this rank (0-based)
  int32_t n[3], nc[3], nch[3];   // per physical axis: logical length, stored count, stored count of the lower half
  int32_t kmax, ncopy;     // bins 0..kmax; privatised shared-memory copies of the histogram per CTA
  double f2;               // factor^2
};
// Flag barrier over peer-mapped memory (opt-in, P3DFFT_B200_FLAGBAR=1): every rank stores `epoch` into slot `me` of
// every rank's flag array (128-byte slots) and waits until all slots of its own array have reached it.
// peers = DEVICE array of nrank pointers (entry `me` = the rank's own array).
cudaError_t launch_flag_barrier(unsigned* const* peers, int me, int nrank, unsigned epoch, cudaStream_t stream);
// the barrier's two halves as separate launches (pipelined groups): store `epoch` into every rank's slot `me` / wait for all
cudaError_t launch_flag_signal(unsigned* const* peers, int me, int nrank, unsigned epoch, cudaStream_t stream);
cudaError_t launch_flag_wait(unsigned* const* peers, int me, int nrank, unsigned epoch, cudaStream_t stream);
// scoped form: optionally signal `signal_epoch` to every rank, then wait until the ranks of `mask` (bit r = world rank r, at most
// 64 ranks) have signalled `wait_epoch`
cudaError_t launch_flag_sync_mask(unsigned* const* peers, int me, int nrank, unsigned signal_epoch, int do_signal,
                                  unsigned long long mask, unsigned wait_epoch, cudaStream_t stream);
template <typename T> cudaError_t launch_spectrum(const void* B, const SpecJob& job, double* E, cudaStream_t stream);
}  // namespace p3d
