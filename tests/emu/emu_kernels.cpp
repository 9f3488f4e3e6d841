// TEST CODE (CPU only): the any-length stage kernel, the Chebyshev epilogue, the real-copy kernel, the power-spectrum kernel
// and the flag barrier of p3dfft_b200/csrc/fft_kernels.cu compiled for the host emulation (cuda_emu.h).  Part of
// libp3dfft_emu[_single].so only.
#define P3D_EMULATE 1
#include "cuda_emu.h"

#define __noinline__ __attribute__((noinline))

#define P3D_KLAUNCH(kernel, grid, block, smem, stream, ...)                                   \
  do {                                                                                        \
    const size_t smem__ = (size_t)(smem);                                                     \
    const dim3 grid__ = dim3(grid);                                                           \
    const unsigned block__ = (unsigned)(block);                                               \
    emu::enqueue(stream, [=]() {                                                              \
      emu::canary_set(p3d::smem_raw, smem__, sizeof(p3d::smem_raw));                          \
      emu::canary_set((unsigned char*)p3d::spec_hist, smem__, sizeof(p3d::spec_hist));        \
      emu::current_kernel = #kernel;                                                          \
      emu::launch([&]() { kernel(__VA_ARGS__); }, grid__, block__);                           \
      emu::current_kernel = nullptr;                                                          \
      emu::canary_check(p3d::smem_raw, smem__, sizeof(p3d::smem_raw), #kernel);               \
      emu::canary_check((const unsigned char*)p3d::spec_hist, smem__, sizeof(p3d::spec_hist), #kernel); \
    });                                                                                       \
  } while (0)

namespace p3d {
alignas(128) unsigned char smem_raw[232448];
double spec_hist[232448 / 8];
}  // namespace p3d

#include "../../p3dfft_b200/csrc/fft_kernels.cu"
