// TEST CODE (CPU only): runs the specialised stage kernels of p3dfft_b200/csrc/fft_fast.cuh on the host through the
// emulation in cuda_emu.h.  The product's own host code (eligibility, plan-segment -> run conversion, twiddle
// blocks, length dispatch: p3dfft_b200/csrc/fft_fast.cu) is compiled in unchanged; only the launch macro is replaced.
// Built by p3dfft_b200/build.py into p3dfft_b200/lib/libemu_fast[_single].so; never part of the product library.
#define P3D_EMULATE 1
#include "cuda_emu.h"

#ifndef EMU_NO_RUNTIME
#include "emu_runtime.inc"
#endif

// every launch of fft_fast.cu goes through this macro (the CUDA build defines it with <<< >>>)
#define P3D_LAUNCH(...)                                                             \
  do {                                                                              \
    (void)stream; (void)smem;                                                       \
    static_assert(smem <= sizeof(p3d::fast::smem_raw), "emulated shared memory");   \
    const FastStage fcopy = f;                                                      \
    emu::enqueue(stream, [=]() {                                                    \
      emu::canary_set(p3d::fast::smem_raw, smem, sizeof(p3d::fast::smem_raw));      \
      emu::current_kernel = #__VA_ARGS__;                                           \
      emu::launch([&]() { __VA_ARGS__(fcopy); }, emu::grid_for(tiles), NT);         \
      emu::current_kernel = nullptr;                                                \
      emu::canary_check(p3d::fast::smem_raw, smem, sizeof(p3d::fast::smem_raw), #__VA_ARGS__); \
    });                                                                             \
    e = cudaSuccess;                                                                \
  } while (0)

namespace p3d { namespace fast { alignas(128) unsigned char smem_raw[232448]; } }

// the one CUDA runtime entry point the host code of fft_fast.cu still references: renamed, so that neither a linked
// nor an already loaded libcudart (the product library is in the same process in the tests) is ever called
static inline cudaError_t emu_cudaGetLastError(void) { return cudaSuccess; }
#define cudaGetLastError emu_cudaGetLastError

#include "../../p3dfft_b200/csrc/fft_fast.cu"

#ifdef SINGLE_PREC
typedef float emu_real;
#else
typedef double emu_real;
#endif

// runs one resolved stage (seg.base set by the caller) with the specialised kernels; returns 0, 1 if no specialised
// kernel takes this stage (the generic kernel would), or a negative error
static int g_last_variant = 0;
extern "C" int emu_last_variant(void) { return g_last_variant; }      // kernel variant of the last stage (fast.h, FastStage.variant)

extern "C" int emu_run_fast(const P3dStage* st) {
  g_last_variant = 0;
  if (!p3d::fast_supported<emu_real>(*st)) return 1;
  p3d::FastStage fs;
  p3d::fast_reload_switches();      // the kernel-only test library has no p3dfft_setup: tests flip the switches between runs
  p3d::to_fast(*st, fs, sizeof(emu_real), p3d::fast_variant<emu_real>(*st));
  g_last_variant = fs.variant;
  std::vector<emu_real> tw(2 * p3d::fast_twiddle_elems<emu_real>(st->kind, st->nfft, fs.variant) + 2);
  p3d::fast_twiddle_fill<emu_real>(st->kind, st->nfft, tw.data(), fs.variant);
  fs.tw = tw.data();
  cudaError_t e = p3d::launch_fast<emu_real>(*st, fs, nullptr);
  if (e == cudaErrorMisalignedAddress) return 1;
  return e == cudaSuccess ? 0 : -(int)e;
}
