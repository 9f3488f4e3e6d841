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
}  // namespace p3d
