// TEST CODE (CPU only): the C ABI of the library (p3dfft_b200/csrc/api.cpp, unchanged) on top of a mock CUDA runtime, so
// that the WHOLE library -- planner, executor, staging of host arrays, scaling, epilogues, the auxiliary routines -- runs in
// the CPU test-suite through the emulated kernels (emu_fast.cpp, emu_kernels.cpp).  "Device memory" is host memory, streams
// execute at enqueue time in program order, events are wall-clock stamps.  Several ranks = several processes: emu_mp.inc
// (shared-memory "device" buffers that the mock cudaIpc calls map into the peers, and a mock NCCL bound through dlsym).
// Linked with -Bsymbolic: the mock entry points below carry the real CUDA runtime names and must win over an already
// loaded libcudart inside this library only.
#define P3D_EMULATE 1
#include "cuda_emu.h"

#include <dlfcn.h>
#include <nccl.h>
#include <sys/mman.h>

#include <chrono>
#include <cstdlib>
#include <map>
#include <mutex>

#include "emu_runtime.inc"
#include "emu_streams.inc"
#include "emu_mp.inc"

namespace {
std::map<char*, size_t> g_alloc;          // "device" allocations
std::map<char*, size_t> g_hostalloc;      // page-locked host allocations (cudaHostAlloc)
std::mutex g_mu;
}  // namespace

extern "C" {
cudaError_t cudaMalloc(void** p, size_t n) {
  static bool hooked = false;
  if (emu_mp::use_shm() && !hooked) { atexit(emu_mp::shm_cleanup); hooked = true; }
  *p = emu_mp::shm_create(n, emu_mp::use_shm());      // guarded; shared memory in multi-rank runs
  if (!*p) return cudaErrorMemoryAllocation;
  std::lock_guard<std::mutex> l(g_mu);
  g_alloc[(char*)*p] = n;
  return cudaSuccess;
}
cudaError_t cudaFree(void* p) {
  if (!p) return cudaSuccess;
  emu_rt::sync_all();                 // cudaFree synchronises the device
  { std::lock_guard<std::mutex> l(g_mu); g_alloc.erase((char*)p); }
  emu_mp::shm_release(p);
  return cudaSuccess;
}
// page-locked host memory: ordinary heap memory here (cudaPointerGetAttributes keeps calling it unregistered, so host arrays
// of any size take the pageable path of api.cpp -- the ring of chunks and the copy threads)
// (own mappings, unmapped at cudaFreeHost like the real runtime's: a late access -- a copy thread that outlives its ring --
// faults here as it would on the GPU box instead of scribbling over recycled heap memory)
cudaError_t cudaHostAlloc(void** p, size_t n, unsigned) {
  const size_t len = (n + 4095) / 4096 * 4096 + 4096;
  void* m = mmap(nullptr, len, PROT_READ | PROT_WRITE, MAP_PRIVATE | MAP_ANONYMOUS, -1, 0);
  if (m == MAP_FAILED) { *p = nullptr; return cudaErrorMemoryAllocation; }
  std::lock_guard<std::mutex> l(g_mu);
  g_hostalloc[(char*)m] = len;
  *p = m;
  return cudaSuccess;
}
cudaError_t cudaFreeHost(void* p) {
  emu_rt::sync_all();
  std::lock_guard<std::mutex> l(g_mu);
  auto it = g_hostalloc.find((char*)p);
  if (it == g_hostalloc.end()) return cudaErrorInvalidValue;
  munmap(p, it->second);
  g_hostalloc.erase(it);
  return cudaSuccess;
}
cudaError_t cudaMemcpy(void* d, const void* s, size_t n, cudaMemcpyKind) { emu_rt::sync_legacy(); memmove(d, s, n); return cudaSuccess; }
cudaError_t cudaMemcpyAsync(void* d, const void* s, size_t n, cudaMemcpyKind, cudaStream_t st) {
  emu_rt::enqueue(st, [=]() { memmove(d, s, n); });
  return cudaSuccess;
}
cudaError_t cudaMemset(void* d, int v, size_t n) { emu_rt::sync_legacy(); memset(d, v, n); return cudaSuccess; }
cudaError_t cudaMemsetAsync(void* d, int v, size_t n, cudaStream_t st) {
  emu_rt::enqueue(st, [=]() { memset(d, v, n); });
  return cudaSuccess;
}
cudaError_t cudaStreamCreate(cudaStream_t* s) { *s = emu_rt::new_stream(true); return cudaSuccess; }
cudaError_t cudaStreamCreateWithFlags(cudaStream_t* s, unsigned flags) { *s = emu_rt::new_stream(!(flags & cudaStreamNonBlocking)); return cudaSuccess; }
cudaError_t cudaStreamDestroy(cudaStream_t s) { emu_rt::sync_stream(s); emu_rt::delete_stream(s); return cudaSuccess; }
cudaError_t cudaStreamSynchronize(cudaStream_t s) { emu_rt::sync_stream(s); return cudaSuccess; }
cudaError_t cudaDeviceSynchronize(void) { emu_rt::sync_all(); return cudaSuccess; }
cudaError_t cudaStreamWaitEvent(cudaStream_t s, cudaEvent_t e, unsigned) { emu_rt::wait_event(s, (emu_rt::Event*)e); return cudaSuccess; }
cudaError_t cudaStreamSetAttribute(cudaStream_t, cudaStreamAttrID, const cudaStreamAttrValue*) { return cudaSuccess; }
cudaError_t cudaEventCreate(cudaEvent_t* e) { *e = (cudaEvent_t) new emu_rt::Event; return cudaSuccess; }
cudaError_t cudaEventCreateWithFlags(cudaEvent_t* e, unsigned) { *e = (cudaEvent_t) new emu_rt::Event; return cudaSuccess; }
cudaError_t cudaEventDestroy(cudaEvent_t e) { emu_rt::sync_all(); delete (emu_rt::Event*)e; return cudaSuccess; }
cudaError_t cudaEventRecord(cudaEvent_t e, cudaStream_t s) { emu_rt::record_event(s, (emu_rt::Event*)e); return cudaSuccess; }
cudaError_t cudaEventSynchronize(cudaEvent_t e) { emu_rt::sync_event((emu_rt::Event*)e); return cudaSuccess; }
cudaError_t cudaEventElapsedTime(float* ms, cudaEvent_t a, cudaEvent_t b) {
  emu_rt::Event *ea = (emu_rt::Event*)a, *eb = (emu_rt::Event*)b;
  if (ea->done < ea->issued || eb->done < eb->issued) return cudaErrorNotReady;      // like the real runtime: no implicit wait
  *ms = std::chrono::duration<float, std::milli>(eb->t - ea->t).count();
  return cudaSuccess;
}
cudaError_t cudaPointerGetAttributes(cudaPointerAttributes* a, const void* p) {
  memset(a, 0, sizeof *a);
  a->type = cudaMemoryTypeUnregistered;
  std::lock_guard<std::mutex> l(g_mu);
  auto it = g_alloc.upper_bound((char*)p);
  if (it != g_alloc.begin()) {
    --it;
    if ((char*)p < it->first + it->second) { a->type = cudaMemoryTypeDevice; a->devicePointer = (void*)p; }
  }
  return cudaSuccess;
}
cudaError_t cudaGetLastError(void) { return cudaSuccess; }
const char* cudaGetErrorName(cudaError_t) { return "cudaErrorEmulated"; }
const char* cudaGetErrorString(cudaError_t) { return "error reported by the emulated runtime"; }
cudaError_t cudaSetDevice(int) { return cudaSuccess; }
cudaError_t cudaGetDevice(int* d) { *d = 0; return cudaSuccess; }
cudaError_t cudaGetDeviceCount(int* n) { *n = 1; return cudaSuccess; }
cudaError_t cudaDeviceGetAttribute(int* v, cudaDeviceAttr, int) { *v = 4; return cudaSuccess; }        // a 4-SM "GPU"
cudaError_t cudaDeviceSetLimit(cudaLimit, size_t) { return cudaSuccess; }
cudaError_t cudaCtxResetPersistingL2Cache(void) { return cudaSuccess; }
// peer mappings: a handle names the shared-memory segment of an allocation (multi-process runs only)
cudaError_t cudaIpcGetMemHandle(cudaIpcMemHandle_t* h, void* p) {
  static_assert(sizeof(cudaIpcMemHandle_t) >= 64, "handle size");
  return emu_mp::shm_handle(p, h->reserved) ? cudaSuccess : cudaErrorNotSupported;
}
cudaError_t cudaIpcOpenMemHandle(void** p, cudaIpcMemHandle_t h, unsigned) {
  *p = emu_mp::shm_open_peer(h.reserved);
  if (!*p) return cudaErrorInvalidValue;
  std::lock_guard<std::mutex> l(g_mu);
  g_alloc[(char*)*p] = emu_mp::g_shm[(char*)*p].user_bytes;       // a device pointer of this process from now on
  return cudaSuccess;
}
cudaError_t cudaIpcCloseMemHandle(void* p) {
  { std::lock_guard<std::mutex> l(g_mu); g_alloc.erase((char*)p); }
  return emu_mp::shm_release(p) ? cudaSuccess : cudaErrorInvalidValue;
}
}

// fault report: which kernel touched which guard page (async-signal-unsafe printing is fine for a test that is about to die)
#include <signal.h>
namespace {
struct sigaction g_prev_segv;
void on_segv(int sig, siginfo_t* si, void* uc) {
  char* a = (char*)si->si_addr;
  if (emu::current_kernel)
    fprintf(stderr, "emulated runtime: SIGSEGV at %p in kernel %s (block %u,%u thread %u)\n", (void*)a, emu::current_kernel, blockIdx.x,
            blockIdx.y, threadIdx.x);
  for (auto& kv : emu_mp::g_shm) {
    const emu_mp::ShmBlock& b = kv.second;
    if (a >= b.map && a < b.map + b.map_bytes)
      fprintf(stderr, "  inside the guarded mapping of a %zu-byte %s allocation: %lld bytes past its end (negative: before its start)\n",
              b.user_bytes, b.owner ? "own" : "peer", (long long)(a - (kv.first + b.user_bytes)) < 0 && a < kv.first ? (long long)(a - kv.first) : (long long)(a - (kv.first + b.user_bytes)));
  }
  // hand over to whoever was there before (the other emulated library, Python's faulthandler, or the default action)
  if ((g_prev_segv.sa_flags & SA_SIGINFO) && g_prev_segv.sa_sigaction) { g_prev_segv.sa_sigaction(sig, si, uc); return; }
  if (!(g_prev_segv.sa_flags & SA_SIGINFO) && g_prev_segv.sa_handler != SIG_DFL && g_prev_segv.sa_handler != SIG_IGN) {
    g_prev_segv.sa_handler(sig);
    return;
  }
  signal(SIGSEGV, SIG_DFL);
  raise(SIGSEGV);
}
struct InstallSegv {
  InstallSegv() {
    static char altstack[1 << 16];
    stack_t ss; ss.ss_sp = altstack; ss.ss_size = sizeof altstack; ss.ss_flags = 0;
    sigaltstack(&ss, nullptr);
    struct sigaction sa;
    memset(&sa, 0, sizeof sa);
    sa.sa_sigaction = on_segv;
    sa.sa_flags = SA_SIGINFO | SA_ONSTACK;
    sigaction(SIGSEGV, &sa, &g_prev_segv);
  }
} g_install_segv;
}  // namespace

// test hook: number (and bytes) of live mock cudaMalloc blocks + peer mappings -- p3dfft_clean must leave none behind
extern "C" long long emu_live_allocations(long long* bytes) {
  long long n = 0, b = 0;
  for (auto& kv : emu_mp::g_shm) { n++; b += (long long)kv.second.user_bytes; }
  if (bytes) *bytes = b;
  return n;
}

// extension for the tests: lets a numpy array play the part of a device array (used in place, not staged)
extern "C" void emu_register_device_range(void* p, size_t n) { std::lock_guard<std::mutex> l(g_mu); g_alloc[(char*)p] = n; }
extern "C" void emu_unregister_device_range(void* p) { std::lock_guard<std::mutex> l(g_mu); g_alloc.erase((char*)p); }

// api.cpp binds NCCL with dlopen/dlsym: hand it the mock
#define dlopen emu_mp::emu_dlopen
#define dlsym emu_mp::emu_dlsym
#include "../../p3dfft_b200/csrc/api.cpp"
