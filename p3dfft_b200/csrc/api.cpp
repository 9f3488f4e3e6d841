// C ABI of the B200 build: the eleven entry points of the reference's BIND(C) layer
// (include/p3dfft.h cites the file:line each replaces) plus the p3dfft_b200_* extensions.
//
// State model follows the reference: ONE plan per process held in module-global variables
// (build/module.F90:103-176), p3dfft_setup may be called again only after p3dfft_clean
// (setup.F90:130-135), every rank calls every routine collectively.
#include <cuda_runtime.h>
#include <sched.h>
#include <dlfcn.h>
#include <nccl.h>

#include <cmath>
#include <condition_variable>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <map>
#include <mutex>
#include <string>
#include <thread>
#include <tuple>
#include <vector>

#include "../../include/p3dfft_b200.h"
#include "fast.h"
#include "kernels.h"
#include "plan.h"
#include "procmap.h"
#include "stage.h"

#ifdef SINGLE_PREC
typedef float real_t;
#else
typedef double real_t;
#endif
static const size_t CSIZE = 2 * sizeof(real_t);

// defaults of the multi-GPU switches (decided by the A/B runs under profiles/)
// Measured on 8 B200s, 1024^3 double, 2x4 (profiles/r2_ab_multi_8gpu.log): NCCL barrier, no pipelining 5.67 ms; flag barrier
// 5.62; flag barrier + the last stage-exchange-stage triple in 4 chunks with 74 SMs for the consumer 5.29 (2 chunks 5.34,
// 8 chunks 5.50, 56 SMs 5.41, 92 SMs 5.56); 1x8: 5.02 -> 4.87; 1x2: 14.58 -> 14.07; 2x1: 15.48 -> 13.63.
#ifndef P3D_DEFAULT_FLAGBAR
#define P3D_DEFAULT_FLAGBAR 1
#endif
#ifndef P3D_DEFAULT_SCOPED
#define P3D_DEFAULT_SCOPED 1
#endif
#ifndef P3D_DEFAULT_OVERLAP
#define P3D_DEFAULT_OVERLAP 4
#endif
#ifndef P3D_DEFAULT_OVERLAP_SMS
#define P3D_DEFAULT_OVERLAP_SMS 74
#endif
#ifndef P3D_DEFAULT_OVERLAP_SHAPE
#define P3D_DEFAULT_OVERLAP_SHAPE ""
#endif
// the pipelined group pays a few launches and barriers per chunk: only for stages of at least this many bytes per rank
#ifndef P3D_OVERLAP_MIN_BYTES
#define P3D_OVERLAP_MIN_BYTES (48ll << 20)
#endif

extern "C" int p3dfft_b200_get_unique_id(void* id128);
extern "C" int p3dfft_b200_comm_create(int rank, int size, const void* id128, int device);

namespace {

// ------------------------------------------------------------------------------------
// errors
// ------------------------------------------------------------------------------------
int g_error_mode = 0;
std::string g_last_error;

void report(bool fatal, const char* fmt, ...) {
  char buf[1024];
  va_list ap; va_start(ap, fmt); vsnprintf(buf, sizeof buf, fmt, ap); va_end(ap);
  g_last_error = buf;
  fprintf(stderr, "%s\n", buf);
  if (fatal && g_error_mode == 0) { fflush(stderr); abort(); }     // where the reference calls MPI_Abort
}

#define CUDA_OK(expr)                                                                      \
  do {                                                                                     \
    cudaError_t e__ = (expr);                                                              \
    if (e__ != cudaSuccess) {                                                              \
      report(true, "P3DFFT(B200) CUDA error %s at %s:%d: %s", cudaGetErrorName(e__), __FILE__, __LINE__, \
             cudaGetErrorString(e__));                                                     \
      return false;                                                                        \
    }                                                                                      \
  } while (0)

// ------------------------------------------------------------------------------------
// NCCL, bound at run time so the library loads (and the planner works) without it
// ------------------------------------------------------------------------------------
struct Nccl {
  void* h = nullptr;
  ncclResult_t (*GetUniqueId)(ncclUniqueId*) = nullptr;
  ncclResult_t (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int) = nullptr;
  ncclResult_t (*CommSplit)(ncclComm_t, int, int, ncclComm_t*, ncclConfig_t*) = nullptr;
  ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
  ncclResult_t (*Send)(const void*, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*Recv)(void*, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*AllGather)(const void*, void*, size_t, ncclDataType_t, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*AllReduce)(const void*, void*, size_t, ncclDataType_t, ncclRedOp_t, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*GroupStart)() = nullptr;
  ncclResult_t (*GroupEnd)() = nullptr;
  const char* (*GetErrorString)(ncclResult_t) = nullptr;
  bool load() {
    if (h) return true;
    const char* names[] = {"libnccl.so.2", "libnccl.so", nullptr};
    for (int i = 0; names[i] && !h; i++) h = dlopen(names[i], RTLD_NOW | RTLD_GLOBAL);
    if (!h) { report(true, "P3DFFT(B200): cannot load libnccl.so.2: %s", dlerror()); return false; }
#define SYM(f) *(void**)(&f) = dlsym(h, "nccl" #f); if (!f) { report(true, "P3DFFT(B200): nccl" #f " missing"); return false; }
    SYM(GetUniqueId) SYM(CommInitRank) SYM(CommSplit) SYM(CommDestroy) SYM(Send) SYM(Recv)
    SYM(GroupStart) SYM(GroupEnd) SYM(GetErrorString) SYM(AllGather) SYM(AllReduce)
#undef SYM
    return true;
  }
} g_nccl;

#define NCCL_OK(expr)                                                                      \
  do {                                                                                     \
    ncclResult_t r__ = (expr);                                                             \
    if (r__ != ncclSuccess) {                                                              \
      report(true, "P3DFFT(B200) NCCL error at %s:%d: %s", __FILE__, __LINE__, g_nccl.GetErrorString(r__)); \
      return false;                                                                        \
    }                                                                                      \
  } while (0)

struct CommCtx {
  int rank = 0, size = 1, device = -1;
  ncclComm_t world = nullptr;
};
std::map<int, CommCtx*> g_comms;
int g_next_handle = 0x50330001;      // 'P3' + counter: out of the way of small integers, which are Fortran MPI handles of some MPIs

// ------------------------------------------------------------------------------------
// the plan (module-global state of the reference)
// ------------------------------------------------------------------------------------
struct PlanKey {
  int backward, nv; char op; long long dim_real, dim_cplx; int w, p2p;      // p2p: bit 0 peer-to-peer, bits 8.. overlap chunks
  bool operator<(const PlanKey& o) const {
    return std::tie(backward, nv, op, dim_real, dim_cplx, w, p2p) <
           std::tie(o.backward, o.nv, o.op, o.dim_real, o.dim_cplx, o.w, o.p2p);
  }
};

struct Lib {
  bool set = false;
  p3d::Decomp d;
  bool overwrite = true;
  CommCtx* comm = nullptr;
  ncclComm_t row = nullptr, col = nullptr;
  void* buf[5] = {nullptr, nullptr, nullptr, nullptr, nullptr};   // indexed by P3dBuf (A,B,C used)
  int nv_preset = 0;
  void* stage_in = nullptr; size_t stage_in_bytes = 0;
  void* stage_out = nullptr; size_t stage_out_bytes = 0;
  cudaStream_t own_stream = nullptr, user_stream = nullptr;
  bool has_user_stream = false;
  bool async = false;
  bool force_generic = false;
  // peer-to-peer transposes: stage kernels store each block straight into the destination rank's
  // receive buffer over NVLink (CUDA IPC mappings of the peers' work buffers); exchange = barrier
  bool api_p2p = true;           // p3dfft_b200_set_p2p; the environment variable, when present at setup, wins
  bool want_p2p = true, p2p = false;
  std::vector<void*> peer_buf;   // [world rank * 3 + (buffer id - P3D_BUF_A)], own entries = own buffers
  float* bar_scratch = nullptr;
  // flag barrier over peer-mapped memory (the default since round 2, stress-tested on 2/4/8 B200s; P3DFFT_B200_FLAGBAR=0: NCCL all-reduce)
  bool want_flagbar = false, flagbar = false;
  unsigned* bar_flags = nullptr;                 // this rank's slots: one 128-byte slot per world rank, own 2 MiB allocation
  std::vector<unsigned*> peer_flags;             // [world rank] mapped flag arrays (own entry = bar_flags)
  unsigned** peer_flags_dev = nullptr;           // the same table on the device
  unsigned bar_epoch = 0;
  // Scoped synchronisation (flag barrier only, at most 64 ranks; P3DFFT_B200_SCOPED=0 keeps world barriers).  A flag is its
  // owner's progress counter: every sync point of a plan (each exchange, each chunk of a pipelined group, the end of a
  // transform) signals the next epoch to ALL ranks, but an exchange WAITS only for the ranks of its own row or column.
  // The write-after-read rule becomes exact: release[b] is the epoch whose signal is stream-ordered behind this rank's
  // last read of buffer b -- every rank runs the same step sequence, so it is also the epoch a PEER signals behind ITS last
  // read of its copy of b.  A stage that stores into the peers' copies of b first waits for its target peers to have reached
  // release[b], unless a wait already queued on this stream covers it (passed[c], per communicator).
  bool want_scoped = true, scoped = false;
  unsigned long long scope_mask[2] = {0, 0};      // world ranks of this rank's row / column (bit r)
  unsigned release[5] = {0, 0, 0, 0, 0};
  unsigned passed[2] = {0, 0};
  // dirty[b]: buffer b was the receive buffer of an exchange since the last world barrier, i.e. some rank may still
  // be reading it.  A peer-to-peer stage must not store into the peers' copies of such a buffer before another
  // barrier (run_plan).  Derived from the exchange steps only, so every rank takes the same decisions.
  bool dirty[5] = {false, false, false, false, false};
  long long work_elems_alloc = 0;   // complex elements per work buffer
  bool rtran_sized = false;         // the buffers already cover rtran_work_elems()
  std::map<int, p3d::TransformPlan> aux_plans;     // r2c_1d (key 100) and rtran (key which*2 + p2p) plans
  // pipelined tail of the peer-to-peer plans (P3DFFT_B200_OVERLAP=C chunks, default 4; plan.h split_for_overlap): the consumer
  // chunks run on a side stream, on at most overlap_sms SMs while the producer keeps the rest (A/B on 2/4/8 B200s under profiles/)
  int overlap = P3D_DEFAULT_OVERLAP, overlap_sms = P3D_DEFAULT_OVERLAP_SMS;
  bool overlap_forced = false;   // P3DFFT_B200_OVERLAP given in the environment: also for small transforms (tests)
  int overlap_sms_dir[2] = {0, 0};      // per direction (forward, backward); 0: overlap_sms
  std::vector<int> overlap_shape;       // relative chunk sizes (P3DFFT_B200_OVERLAP_SHAPE=1,3,3,3,1); empty: equal chunks
  cudaStream_t side_stream = nullptr;
  std::vector<cudaEvent_t> chunk_events;
  cudaEvent_t side_done = nullptr;
  double scale_fwd = 1.0, scale_bwd = 1.0;          // fused output normalisation (p3dfft_b200_set_scale)
  double* spec_dev = nullptr; int spec_bins = 0;   // device accumulator of p3dfft_b200_spectrum
  p3d::ProcMap procmap;                            // proc_id2coords / proc_dims tables (setup.F90:224-230, 551-577)
  bool plain_layout = false;     // true: the reference's pack-buffer layouts instead of the tile-blocked ones
  int force_row_bytes = 0;       // 64 / 128: override the planner's choice of the tile row width (p3dfft_b200_row_bytes)
  int row_bytes_plan = 0;        // ... as latched by p3dfft_setup: the live plan and its buffers never see a later change
  int W() const { return plain_layout ? 0 : p3d::pick_W(d.ny, d.nz, (int)CSIZE, row_bytes_plan); }
  long long fast_launches = 0;
  std::map<std::pair<int, int>, void*> fast_twiddles;   // (x-stage?, nfft) -> device block
  double timers[12] = {0};
  std::map<int, void*> twiddles;
  std::map<PlanKey, p3d::TransformPlan> plans;
  std::vector<cudaEvent_t> events;
  long long launches = 0;
  int stride1 =
#ifdef STRIDE1
      1;
#else
      0;
#endif
  int dims_c =
#ifdef DIMS_C
      1;
#else
      0;
#endif
  cudaStream_t stream() { return has_user_stream ? user_stream : own_stream; }
} L;

bool is_device_ptr(const void* p) {
  cudaPointerAttributes a;
  cudaError_t e = cudaPointerGetAttributes(&a, p);
  if (e != cudaSuccess) { cudaGetLastError(); return false; }
  return a.type == cudaMemoryTypeDevice || a.type == cudaMemoryTypeManaged;
}


// ------------------------------------------------------------------------------------
// Host arrays.  The reference's callers pass ordinary heap memory (malloc in every sample driver, driver_sine.c:144-146).
// Page-locked arrays (cudaHostAlloc / cudaHostRegister) are copied by one DMA; pageable ones go through a ring of
// page-locked chunks filled (or drained) by a few copy threads while the previous chunk is on the PCIe bus, instead of
// the driver's single-threaded bounce copy.  Nothing of the caller's memory is pinned or remembered between calls.
// ------------------------------------------------------------------------------------
class CopyPool {
  std::vector<std::thread> th_;
  std::mutex m_;
  std::condition_variable cv_, done_;
  char* dst_ = nullptr; const char* src_ = nullptr; size_t n_ = 0;
  unsigned long gen_ = 0; int pending_ = 0; bool stop_ = false;
  void work(int id, int nth, unsigned long seen) {      // seen: the generation at start -- a pool restarted after p3dfft_clean
    for (;;) {                                          // must not take the last job of its predecessor for a new one
      std::unique_lock<std::mutex> l(m_);
      cv_.wait(l, [&] { return stop_ || gen_ != seen; });
      if (stop_) return;
      seen = gen_;
      char* d = dst_; const char* s = src_; const size_t n = n_;
      l.unlock();
      const size_t per = ((n + nth - 1) / nth + 4095) / 4096 * 4096, a = std::min(n, per * id), b = std::min(n, a + per);
      if (b > a) memcpy(d + a, s + a, b - a);
      l.lock();
      if (--pending_ == 0) done_.notify_all();
    }
  }
 public:
  int threads() const { return (int)th_.size(); }
  void start(int n) {
    if (!th_.empty()) return;
    unsigned long g0;
    { std::lock_guard<std::mutex> l(m_); g0 = gen_; }
    for (int i = 0; i < n; i++) th_.emplace_back([this, i, n, g0] { work(i, n, g0); });
  }
  void copy(void* d, const void* s, size_t n) {      // returns when every slice has been copied
    if (th_.empty() || n < (1u << 16)) { memcpy(d, s, n); return; }
    std::unique_lock<std::mutex> l(m_);
    dst_ = (char*)d; src_ = (const char*)s; n_ = n; pending_ = (int)th_.size(); gen_++;
    cv_.notify_all();
    done_.wait(l, [&] { return pending_ == 0; });
  }
  void shutdown() {
    { std::lock_guard<std::mutex> l(m_); stop_ = true; }
    cv_.notify_all();
    for (auto& t : th_) t.join();
    th_.clear(); stop_ = false;
  }
};

struct HostPipe {
  static constexpr int NSLOT = 4;
  size_t chunk = 8u << 20;       // measured on the B200 box (tools/ab_hostpipe.py, profiles/r2_ab_hostpipe.log): 8 MB chunks, 12 threads:
                                 // 1024^3 pair 805 ms against 665 page-locked; 32 MB / 8 threads (the first setting): 940-1050
  char* slot[NSLOT] = {nullptr, nullptr, nullptr, nullptr};
  cudaEvent_t ev[NSLOT] = {nullptr, nullptr, nullptr, nullptr};
  CopyPool pool;
  bool ready = false, failed = false;
  bool init() {
    if (ready) return true;
    if (failed) return false;
    const char* ct = getenv("P3DFFT_B200_COPY_THREADS");
    const char* ck = getenv("P3DFFT_B200_COPY_CHUNK_KB");      // (tests: small chunks so that small arrays take the ring)
    if (ck && atoi(ck) > 0) chunk = (size_t)atoi(ck) << 10;
    for (int i = 0; i < NSLOT; i++) {
      if (cudaHostAlloc((void**)&slot[i], chunk, cudaHostAllocDefault) != cudaSuccess ||
          cudaEventCreateWithFlags(&ev[i], cudaEventDisableTiming) != cudaSuccess) { cudaGetLastError(); release(); failed = true; return false; }
    }
    unsigned hw = std::thread::hardware_concurrency();
#ifdef __linux__
    cpu_set_t cs;      // the cores this process may use, not the machine's
    if (sched_getaffinity(0, sizeof cs, &cs) == 0 && CPU_COUNT(&cs) > 0) hw = (unsigned)CPU_COUNT(&cs);
#endif
    const unsigned share = (unsigned)std::max(1, std::min(L.d.numtasks, 8));      // ranks of one box share its cores
    const int nth = ct ? atoi(ct) : (int)std::min<unsigned>(12, std::max<unsigned>(1, hw * 3 / 4 / share));
    if (nth > 1) pool.start(nth);
    ready = true;
    return true;
  }
  void release() {
    pool.shutdown();
    for (int i = 0; i < NSLOT; i++) {
      if (slot[i]) { cudaFreeHost(slot[i]); slot[i] = nullptr; }
      if (ev[i]) { cudaEventDestroy(ev[i]); ev[i] = nullptr; }
    }
    ready = false;
  }
} g_pipe;

bool is_pinned_host(const void* p) {
  cudaPointerAttributes a;
  if (cudaPointerGetAttributes(&a, p) != cudaSuccess) { cudaGetLastError(); return false; }
  return a.type == cudaMemoryTypeHost;
}

// host -> device on stream st; on return the host array may be reused by the caller
bool copy_h2d(void* dev, const void* host, size_t bytes, cudaStream_t st) {
  static const bool direct = getenv("P3DFFT_B200_HOSTPIPE") && atoi(getenv("P3DFFT_B200_HOSTPIPE")) == 0;
  if (direct || is_pinned_host(host) || !g_pipe.init() || bytes < 2 * g_pipe.chunk) {
    CUDA_OK(cudaMemcpyAsync(dev, host, bytes, cudaMemcpyHostToDevice, st));
    return true;
  }
  bool used[HostPipe::NSLOT] = {false, false, false, false};
  size_t off = 0;
  for (int k = 0; off < bytes; k++) {
    const int s = k % HostPipe::NSLOT;
    const size_t n = std::min(g_pipe.chunk, bytes - off);
    if (used[s]) CUDA_OK(cudaEventSynchronize(g_pipe.ev[s]));          // the DMA that read this slot has finished
    g_pipe.pool.copy(g_pipe.slot[s], (const char*)host + off, n);
    CUDA_OK(cudaMemcpyAsync((char*)dev + off, g_pipe.slot[s], n, cudaMemcpyHostToDevice, st));
    CUDA_OK(cudaEventRecord(g_pipe.ev[s], st));
    used[s] = true;
    off += n;
  }
  for (int s = 0; s < HostPipe::NSLOT; s++) if (used[s]) CUDA_OK(cudaEventSynchronize(g_pipe.ev[s]));   // slots are free for the next call
  return true;
}

// device -> host on stream st.  Page-locked destination: asynchronous (the caller synchronises the stream); pageable: complete on return
bool copy_d2h(void* host, const void* dev, size_t bytes, cudaStream_t st) {
  static const bool direct = getenv("P3DFFT_B200_HOSTPIPE") && atoi(getenv("P3DFFT_B200_HOSTPIPE")) == 0;
  if (direct || is_pinned_host(host) || !g_pipe.init() || bytes < 2 * g_pipe.chunk) {
    CUDA_OK(cudaMemcpyAsync(host, dev, bytes, cudaMemcpyDeviceToHost, st));
    return true;
  }
  const size_t nchunk = (bytes + g_pipe.chunk - 1) / g_pipe.chunk;
  auto issue = [&](size_t k) -> bool {
    const int s = (int)(k % HostPipe::NSLOT);
    const size_t off = k * g_pipe.chunk, n = std::min(g_pipe.chunk, bytes - off);
    CUDA_OK(cudaMemcpyAsync(g_pipe.slot[s], (const char*)dev + off, n, cudaMemcpyDeviceToHost, st));
    CUDA_OK(cudaEventRecord(g_pipe.ev[s], st));
    return true;
  };
  for (size_t k = 0; k < nchunk && k < (size_t)HostPipe::NSLOT; k++) if (!issue(k)) return false;
  for (size_t k = 0; k < nchunk; k++) {
    const int s = (int)(k % HostPipe::NSLOT);
    const size_t off = k * g_pipe.chunk, n = std::min(g_pipe.chunk, bytes - off);
    CUDA_OK(cudaEventSynchronize(g_pipe.ev[s]));
    g_pipe.pool.copy((char*)host + off, g_pipe.slot[s], n);
    if (k + HostPipe::NSLOT < nchunk && !issue(k + HostPipe::NSLOT)) return false;
  }
  return true;
}

const void* twiddle_table(int nfft) {
  auto it = L.twiddles.find(nfft);
  if (it != L.twiddles.end()) return it->second;
  std::vector<real_t> h(2 * (size_t)nfft);
  const long double twopi = 6.283185307179586476925286766559L;
  for (int k = 0; k < nfft; k++) {
    // exact symmetries keep the table accurate to the last bit for the octant points
    long double ang = -twopi * (long double)k / (long double)nfft;
    h[2 * k] = (real_t)cosl(ang);
    h[2 * k + 1] = (real_t)sinl(ang);
  }
  void* dptr = nullptr;
  if (cudaMalloc(&dptr, h.size() * sizeof(real_t)) != cudaSuccess) return nullptr;
  if (cudaMemcpy(dptr, h.data(), h.size() * sizeof(real_t), cudaMemcpyHostToDevice) != cudaSuccess) return nullptr;
  L.twiddles[nfft] = dptr;
  return dptr;
}

// twiddle block of the specialised kernels (fft_fast.cu), one per (stage class, length)
const void* fast_twiddle_block(int kind, int nfft, int variant = 0) {
  const bool xs = kind == P3D_R2C || kind == P3D_C2R;
  auto key = std::make_pair((xs ? 1 : 0) + 2 * variant, nfft);
  auto it = L.fast_twiddles.find(key);
  if (it != L.fast_twiddles.end()) return it->second;
  const size_t n = p3d::fast_twiddle_elems<real_t>(kind, nfft, variant);
  std::vector<real_t> h(2 * n + 2);
  p3d::fast_twiddle_fill<real_t>(kind, nfft, h.data(), variant);
  void* dptr = nullptr;
  if (cudaMalloc(&dptr, h.size() * sizeof(real_t)) != cudaSuccess) return nullptr;
  if (cudaMemcpy(dptr, h.data(), h.size() * sizeof(real_t), cudaMemcpyHostToDevice) != cudaSuccess) return nullptr;
  L.fast_twiddles[key] = dptr;
  return dptr;
}

void close_peer_maps() {
  const int me = L.comm ? L.comm->rank : 0;
  for (size_t i = 0; i < L.peer_buf.size(); i++)
    if (L.peer_buf[i] && (int)(i / 3) != me) cudaIpcCloseMemHandle(L.peer_buf[i]);
  L.peer_buf.clear();
  L.p2p = false;
  for (size_t i = 0; i < L.peer_flags.size(); i++)
    if (L.peer_flags[i] && (int)i != me) cudaIpcCloseMemHandle(L.peer_flags[i]);
  L.peer_flags.clear();
  L.flagbar = false;
  L.scoped = false;
}

bool world_barrier(cudaStream_t st) {
  if (!L.comm || !L.comm->world) return true;
  if (L.flagbar) {
    CUDA_OK(p3d::launch_flag_barrier(L.peer_flags_dev, L.comm->rank, L.comm->size, ++L.bar_epoch, st));
    for (bool& d : L.dirty) d = false;
    L.passed[0] = L.passed[1] = L.bar_epoch;      // every rank has reached this epoch
    return true;
  }
  if (!L.bar_scratch) { CUDA_OK(cudaMalloc(&L.bar_scratch, 256)); CUDA_OK(cudaMemset(L.bar_scratch, 0, 256)); }
  NCCL_OK(g_nccl.AllReduce(L.bar_scratch, L.bar_scratch + 32, 1, ncclFloat, ncclSum, L.comm->world, st));
  for (bool& d : L.dirty) d = false;
  return true;
}

// Flag barrier: maps every rank's flag array (collective over the world).  The arrays are zeroed before their
// handles travel through the all-gather, so nobody can signal into an array that is still being initialised.
bool open_flag_maps() {
  const int P = L.comm->size, me = L.comm->rank;
  cudaStream_t st = L.stream();
  const size_t bytes = 2u << 20;                 // an allocation of its own (IPC handles name whole allocations)
  if (!L.bar_flags) {
    CUDA_OK(cudaMalloc(&L.bar_flags, bytes));
    CUDA_OK(cudaMemset(L.bar_flags, 0, bytes));
    L.bar_epoch = 0;
    for (unsigned& r : L.release) r = 0;
    L.passed[0] = L.passed[1] = 0;
  }
  if ((size_t)P * 128 > bytes) return false;
  cudaIpcMemHandle_t mine;
  memset(&mine, 0, sizeof mine);
  bool ok = cudaIpcGetMemHandle(&mine, L.bar_flags) == cudaSuccess;
  if (!ok) cudaGetLastError();
  cudaIpcMemHandle_t* dev = nullptr;
  CUDA_OK(cudaMalloc(&dev, sizeof(mine) * (P + 1)));
  CUDA_OK(cudaMemcpy(dev + P, &mine, sizeof mine, cudaMemcpyHostToDevice));
  NCCL_OK(g_nccl.AllGather(dev + P, dev, sizeof(mine), ncclInt8, L.comm->world, st));
  CUDA_OK(cudaStreamSynchronize(st));
  std::vector<cudaIpcMemHandle_t> all(P);
  CUDA_OK(cudaMemcpy(all.data(), dev, sizeof(mine) * P, cudaMemcpyDeviceToHost));
  cudaFree(dev);
  L.peer_flags.assign(P, nullptr);
  for (int r = 0; r < P && ok; r++) {
    if (r == me) { L.peer_flags[r] = L.bar_flags; continue; }
    void* ptr = nullptr;
    if (cudaIpcOpenMemHandle(&ptr, all[r], cudaIpcMemLazyEnablePeerAccess) != cudaSuccess) { cudaGetLastError(); ok = false; }
    L.peer_flags[r] = (unsigned*)ptr;
  }
  float flag = ok ? 0.f : 1.f, sum = 0.f;
  CUDA_OK(cudaMemcpy(L.bar_scratch + 3, &flag, sizeof flag, cudaMemcpyHostToDevice));
  NCCL_OK(g_nccl.AllReduce(L.bar_scratch + 3, L.bar_scratch + 4, 1, ncclFloat, ncclSum, L.comm->world, st));
  CUDA_OK(cudaStreamSynchronize(st));
  CUDA_OK(cudaMemcpy(&sum, L.bar_scratch + 4, sizeof sum, cudaMemcpyDeviceToHost));
  if (sum != 0.f) {
    for (int r = 0; r < P; r++) if (L.peer_flags[r] && r != me) cudaIpcCloseMemHandle(L.peer_flags[r]);
    L.peer_flags.clear();
    return false;
  }
  if (!L.peer_flags_dev) CUDA_OK(cudaMalloc(&L.peer_flags_dev, sizeof(unsigned*) * 1024));
  CUDA_OK(cudaMemcpy(L.peer_flags_dev, L.peer_flags.data(), sizeof(unsigned*) * P, cudaMemcpyHostToDevice));
  L.flagbar = true;
  L.scoped = L.want_scoped && P <= 64;
  L.scope_mask[0] = L.scope_mask[1] = 0;
  for (int r = 0; r < P && L.scoped; r++) {
    const int ip = L.d.dims_c ? r / L.d.jproc : r % L.d.iproc, jp = L.d.dims_c ? r % L.d.jproc : r / L.d.iproc;
    if (jp == L.d.jpid) L.scope_mask[0] |= 1ull << r;      // my row: same jpid
    if (ip == L.d.ipid) L.scope_mask[1] |= 1ull << r;      // my column: same ipid
  }
  return true;
}

// maps the work buffers of every rank in this rank's row and column (collective over the world)
bool open_peer_maps() {
  close_peer_maps();
  const int P = L.comm->size, me = L.comm->rank;
  struct Rec { cudaIpcMemHandle_t h[3]; };
  Rec mine;
  memset(&mine, 0, sizeof mine);
  for (int b = 0; b < 3; b++)
    if (cudaIpcGetMemHandle(&mine.h[b], L.buf[P3D_BUF_A + b]) != cudaSuccess) { cudaGetLastError(); return false; }
  Rec* dev = nullptr;
  CUDA_OK(cudaMalloc(&dev, sizeof(Rec) * (P + 1)));
  CUDA_OK(cudaMemcpy(dev + P, &mine, sizeof mine, cudaMemcpyHostToDevice));
  cudaStream_t st = L.stream();
  NCCL_OK(g_nccl.AllGather(dev + P, dev, sizeof(Rec), ncclInt8, L.comm->world, st));
  CUDA_OK(cudaStreamSynchronize(st));
  std::vector<Rec> all(P);
  CUDA_OK(cudaMemcpy(all.data(), dev, sizeof(Rec) * P, cudaMemcpyDeviceToHost));
  cudaFree(dev);
  L.peer_buf.assign((size_t)P * 3, nullptr);
  bool ok = true;
  for (int r = 0; r < P; r++) {
    const int ip = L.d.dims_c ? r / L.d.jproc : r % L.d.iproc, jp = L.d.dims_c ? r % L.d.jproc : r / L.d.iproc;
    if (r == me) { for (int b = 0; b < 3; b++) L.peer_buf[(size_t)r * 3 + b] = L.buf[P3D_BUF_A + b]; continue; }
    if (ip != L.d.ipid && jp != L.d.jpid) continue;      // neither in my row nor in my column
    for (int b = 0; b < 3 && ok; b++) {
      void* ptr = nullptr;
      if (cudaIpcOpenMemHandle(&ptr, all[r].h[b], cudaIpcMemLazyEnablePeerAccess) != cudaSuccess) { cudaGetLastError(); ok = false; }
      L.peer_buf[(size_t)r * 3 + b] = ptr;
    }
  }
  // all ranks must agree: one failure anywhere disables the path everywhere
  float flag = ok ? 0.f : 1.f, sum = 0.f;
  if (!L.bar_scratch) { CUDA_OK(cudaMalloc(&L.bar_scratch, 256)); CUDA_OK(cudaMemset(L.bar_scratch, 0, 256)); }
  CUDA_OK(cudaMemcpy(L.bar_scratch + 1, &flag, sizeof flag, cudaMemcpyHostToDevice));
  NCCL_OK(g_nccl.AllReduce(L.bar_scratch + 1, L.bar_scratch + 2, 1, ncclFloat, ncclSum, L.comm->world, st));
  CUDA_OK(cudaStreamSynchronize(st));
  CUDA_OK(cudaMemcpy(&sum, L.bar_scratch + 2, sizeof sum, cudaMemcpyDeviceToHost));
  if (sum != 0.f) { close_peer_maps(); return false; }
  L.p2p = true;
  if (L.want_flagbar && !open_flag_maps()) L.flagbar = false;      // the NCCL barrier stays in use
  return true;
}

// for_rtran: the buffers must also stage the real-data transposes (rtran_work_elems, the same bound on every rank).
// Both conditions are identical on all ranks, so re-creating the buffers and the peer mappings stays collective.
bool alloc_work(int nv, bool for_rtran = false) {
  if (nv <= L.nv_preset && (!for_rtran || L.rtran_sized)) return true;
  if (nv < L.nv_preset) nv = L.nv_preset;
  if (for_rtran) L.rtran_sized = true;
  // lazy growth like ftran.F90:133-157 (nv_preset)
  cudaStreamSynchronize(L.stream());
  const bool multi = L.comm && L.comm->size > 1;
  if (multi && !L.peer_buf.empty()) {       // peers may still be storing into the old buffers
    if (!world_barrier(L.stream())) return false;
    cudaStreamSynchronize(L.stream());
    close_peer_maps();
    // an exported allocation must not be freed before every importer has closed its mapping of it (cudaIpcCloseMemHandle)
    if (!world_barrier(L.stream())) return false;
    cudaStreamSynchronize(L.stream());
  }
  long long elems = L.d.work_elems(nv, L.W());
  if (L.rtran_sized) elems = std::max(elems, p3d::rtran_work_elems(L.d));
  size_t bytes = (size_t)elems * CSIZE;
  int nbuf = (L.d.iproc * L.d.jproc > 1) ? 3 : 2;
  for (int b = 0; b < nbuf; b++) {
    int id = P3D_BUF_A + b;
    if (L.buf[id]) { cudaFree(L.buf[id]); L.buf[id] = nullptr; }
    CUDA_OK(cudaMalloc(&L.buf[id], bytes));
  }
  L.nv_preset = nv;
  L.work_elems_alloc = elems;
  L.plans.clear();
  L.aux_plans.clear();
  if (multi && L.want_p2p && L.W() > 0 && nbuf == 3) {
    if (!open_peer_maps() && L.comm->rank == 0 && getenv("P3DFFT_B200_VERBOSE"))
      fprintf(stderr, "P3DFFT(B200): peer mapping unavailable, transposes use ncclSend/ncclRecv\n");
  }
  return true;
}

// tile sizes and twiddle tables of the FFT stages of a freshly built plan
bool finalize_plan(p3d::TransformPlan& tp) {
  for (auto& s : tp.steps) {
    if (s.is_exchange || s.st.kind == P3D_RCOPY) continue;
    if (s.st.na <= 0 || s.st.nb <= 0 || s.st.nc <= 0) { s.st.tile = 1; continue; }      // empty chunk of a pipelined tail: never launched
    s.st.tile = p3d::choose_tile<real_t>(s.st);
    if (s.st.tile <= 0) { report(true, "P3DFFT(B200): transform length %d does not fit on chip", s.st.nfft); return false; }
    if (s.st.kind != P3D_NOOP) {
      s.st.tw = twiddle_table(s.st.nfft);
      if (!s.st.tw) { report(true, "P3DFFT(B200): cannot allocate twiddle table"); return false; }
    }
  }
  return true;
}

p3d::TransformPlan* get_plan(bool backward, const char* op, int nv, long long dim_real, long long dim_cplx) {
  // (decided from the global sizes, so that every rank cuts its plan the same way)
  const long long stage_bytes = (long long)(L.d.nxhpc / L.d.iproc + 1) * L.d.ny * (L.d.nz / L.d.jproc + 1) * (long long)CSIZE * nv;
  const int chunks = (L.p2p && L.overlap > 1 && (stage_bytes >= P3D_OVERLAP_MIN_BYTES || L.overlap_forced)) ? L.overlap : 0;
  PlanKey key{backward ? 1 : 0, nv, backward ? op[0] : op[2], dim_real, dim_cplx, L.W(), (L.p2p ? 1 : 0) | (chunks << 8)};
  auto it = L.plans.find(key);
  if (it != L.plans.end()) return &it->second;
  p3d::TransformPlan tp = p3d::build_plan(L.d, backward, op, nv, dim_real, dim_cplx, L.W(), L.p2p);
  if (!tp.error.empty()) {
    // ftran.F90:640-643: print + MPI_Abort
    report(true, "%s", tp.error.c_str());
    return nullptr;
  }
  if (chunks > 1) p3d::split_for_overlap(tp, chunks, L.W(), L.overlap_shape);
  if (!finalize_plan(tp)) return nullptr;
  auto res = L.plans.emplace(key, std::move(tp));
  return &res.first->second;
}

cudaEvent_t get_event(size_t i) {
  while (L.events.size() <= i) { cudaEvent_t e; cudaEventCreate(&e); L.events.push_back(e); }
  return L.events[i];
}

bool run_exchange(const P3dExchange& e, cudaStream_t st) {
  if (e.p2p) return world_barrier(st);     // data already landed: order it against the consumers
  ncclComm_t c = e.comm == 0 ? L.row : L.col;
  char* sb = (char*)L.buf[e.sendbuf];
  char* rb = (char*)L.buf[e.recvbuf];
  const size_t eb = e.ebytes > 0 ? (size_t)e.ebytes : CSIZE;
  NCCL_OK(g_nccl.GroupStart());
  for (int p = 0; p < e.npeer; p++) {
    if (p == e.self) continue;      // own block was written in place by the producing stage
    if (e.sndcnt[p] > 0) NCCL_OK(g_nccl.Send(sb + e.sndoff[p] * eb, (size_t)e.sndcnt[p] * eb, ncclInt8, p, c, st));
    if (e.rcvcnt[p] > 0) NCCL_OK(g_nccl.Recv(rb + e.rcvoff[p] * eb, (size_t)e.rcvcnt[p] * eb, ncclInt8, p, c, st));
  }
  NCCL_OK(g_nccl.GroupEnd());
  return true;
}

// bytes per element of one side (0 = input, 1 = output) of a stage of this kind
size_t side_elem_bytes(int kind, int side) {
  if (kind == P3D_RCOPY) return sizeof(real_t);
  if ((kind == P3D_R2C && side == 0) || (kind == P3D_C2R && side == 1)) return sizeof(real_t);
  return CSIZE;
}

// Runs a step list.  `in`/`out` may be host or device pointers (host arrays are staged over PCIe inside the
// call).  exchange_ms, when given, receives the device time spent in the exchange steps (rtran's `t`).
bool run_plan(p3d::TransformPlan* tp, const void* in, void* out, size_t in_bytes, size_t out_bytes, int nv,
              long long dim_cplx, bool cheby, double Lz, double* exchange_ms, double out_scale = 1.0) {
  cudaStream_t st = L.stream();
  const p3d::Decomp& d = L.d;
  const bool in_dev = is_device_ptr(in), out_dev = is_device_ptr(out);
  const void* din = in; void* dout = out;
  if (!in_dev) {
    if (L.stage_in_bytes < in_bytes) {
      if (L.stage_in) { cudaFree(L.stage_in); L.stage_in = nullptr; L.stage_in_bytes = 0; }      // (a failed allocation must not leave the old pointer behind)
      CUDA_OK(cudaMalloc(&L.stage_in, in_bytes)); L.stage_in_bytes = in_bytes;
    }
    if (!copy_h2d(L.stage_in, in, in_bytes, st)) return false;
    din = L.stage_in;
  }
  if (!out_dev) {
    if (L.stage_out_bytes < out_bytes) {
      if (L.stage_out) { cudaFree(L.stage_out); L.stage_out = nullptr; L.stage_out_bytes = 0; }
      CUDA_OK(cudaMalloc(&L.stage_out, out_bytes)); L.stage_out_bytes = out_bytes;
    }
    dout = L.stage_out;
  }
  const bool timed = !L.async || exchange_ms;
  size_t nev = 0;
  std::vector<int> slots;
  std::vector<char> is_ex;
  const size_t nsteps = tp->steps.size();
  int nchunks = 0;                 // pipelined tail: chunk ids 0 .. nchunks-1
  for (auto& s : tp->steps) if (s.chunk + 1 > nchunks) nchunks = s.chunk + 1;
  std::vector<unsigned> chunk_epoch((size_t)nchunks, 0u);   // flag-barrier epoch of every chunk of a pipelined group
  bool used_side = false;
  // resolves the segment bases of a stage and launches it on stream sx; sm_cap > 0 limits a persistent grid to that many SMs
  auto launch = [&](const P3dStage& planned, cudaStream_t sx, int sm_cap) -> bool {
    P3dStage stg = planned;
    if (out_scale != 1.0 && stg.out.nseg > 0 && stg.out.seg[0].buf == P3D_BUF_USER_OUT) stg.scale *= out_scale;   // the transform's last stage
    for (int side = 0; side < 2; side++) {
      P3dSide& sd = side ? stg.out : stg.in;
      const size_t esz = side_elem_bytes(stg.kind, side);
      for (int g = 0; g < sd.nseg; g++) {
        P3dSeg& sg = sd.seg[g];
        char* base = sg.buf == P3D_BUF_USER_IN ? (char*)din : sg.buf == P3D_BUF_USER_OUT ? (char*)dout : (char*)L.buf[sg.buf];
        if (sg.peer >= 0) base = (char*)L.peer_buf[(size_t)sg.peer * 3 + (sg.buf - P3D_BUF_A)];     // NVLink peer mapping
        sg.base = base + sg.off * esz;
      }
    }
    cudaError_t e = cudaErrorMisalignedAddress;
    if (stg.kind == P3D_RCOPY) e = p3d::launch_rcopy<real_t>(stg, sx);
    else if (!L.force_generic && p3d::fast_supported<real_t>(stg)) {
      p3d::FastStage fs;
      p3d::to_fast(stg, fs, sizeof(real_t), p3d::fast_variant<real_t>(stg));
      fs.tw = fast_twiddle_block(stg.kind, stg.nfft, fs.variant & 9);      // the block depends on the schedule only
      if (!fs.tw) { report(true, "P3DFFT(B200): cannot allocate twiddle table"); return false; }
      fs.sm_cap = sm_cap;
      e = p3d::launch_fast<real_t>(stg, fs, sx);
      if (e == cudaSuccess) L.fast_launches++;
    }
    if (e == cudaErrorMisalignedAddress) e = p3d::launch_stage<real_t>(stg, sx);   // any length / alignment
    if (e != cudaSuccess) { report(true, "P3DFFT(B200): stage launch failed: %s", cudaGetErrorString(e)); return false; }
    L.launches++;
    return true;
  };
  int sms = 148;
  if (nchunks > 0) {
    int dev = 0;
    cudaGetDevice(&dev);
    if (cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || sms <= 0) sms = 148;
    if (!L.side_stream) CUDA_OK(cudaStreamCreateWithFlags(&L.side_stream, cudaStreamNonBlocking));
    if (!L.side_done) CUDA_OK(cudaEventCreateWithFlags(&L.side_done, cudaEventDisableTiming));
    while ((int)L.chunk_events.size() < nchunks) {
      cudaEvent_t ev; CUDA_OK(cudaEventCreateWithFlags(&ev, cudaEventDisableTiming)); L.chunk_events.push_back(ev);
    }
  }
  auto join_side = [&]() -> bool {       // the main stream waits for everything queued on the side stream
    if (!used_side) return true;
    CUDA_OK(cudaEventRecord(L.side_done, L.side_stream));
    CUDA_OK(cudaStreamWaitEvent(st, L.side_done, 0));
    used_side = false;
    return true;
  };
  // pipelined group (plan.h split_for_overlap): producer chunks P_c on the main stream on `sms - side_sms` SMs (they are bound by
  // NVLink, not by SMs), consumer chunks Q_c on the side stream on the other `side_sms`; the last consumer has the GPU to itself
  const int dir = (!tp->steps.empty() && !tp->steps[0].is_exchange && tp->steps[0].st.timer >= 9) ? 1 : 0;      // 9..12: backward stages
  const int want_sms = L.overlap_sms_dir[dir] > 0 ? L.overlap_sms_dir[dir] : L.overlap_sms;
  const int side_sms = want_sms > 0 && want_sms < sms ? want_sms : sms / 2;
  bool pre_done = false;      // the barrier that protects the receive buffer of the NEXT exchange has been issued
  // scoped synchronisation (see Lib::release): bookkeeping of this call
  const bool scoped = L.scoped && L.flagbar;
  std::vector<int> reads_main, reads_side;      // work buffers read by stages queued since the last signal (main stream) / on the side stream
  bool any_p2p = false;
  int group_comm = -1;
  unsigned group_last_epoch = 0;
  auto note_reads = [&](const P3dStage& stg, bool on_side) {
    for (int g = 0; g < stg.in.nseg; g++) {
      const int b = stg.in.seg[g].buf;
      if (b >= P3D_BUF_A && b <= P3D_BUF_C) (on_side ? reads_side : reads_main).push_back(b);
    }
  };
  auto assign_release = [&](unsigned e) {       // `e` has just been signalled on the main stream, behind everything in reads_main
    for (int b : reads_main) L.release[b] = e;
    reads_main.clear();
  };
  auto hazard_wait = [&](const P3dExchange& ex) -> bool {      // before a stage that stores into the peers' copies of ex.recvbuf
    const unsigned need = L.release[ex.recvbuf];
    if (need != 0 && (int)(need - L.passed[ex.comm]) > 0) {
      CUDA_OK(p3d::launch_flag_sync_mask(L.peer_flags_dev, L.comm->rank, L.comm->size, 0u, 0, L.scope_mask[ex.comm], need, st));
      L.passed[ex.comm] = need;
    }
    return true;
  };
  auto join_group = [&]() -> bool {
    const bool was = used_side;
    if (!join_side()) return false;
    if (was && scoped) {
      reads_main.insert(reads_main.end(), reads_side.begin(), reads_side.end());
      reads_side.clear();
      if (group_comm >= 0 && (int)(group_last_epoch - L.passed[group_comm]) > 0) L.passed[group_comm] = group_last_epoch;
    }
    return true;
  };
  for (size_t i = 0; i < nsteps; i++) {
    auto& s = tp->steps[i];
    if (s.chunk < 0 && used_side && !join_group()) return false;      // a step behind the pipelined group: the group is complete
    if (!s.is_exchange && s.side) {
      // consumer chunk: after every rank's producer chunk (flag wait on this stream, or the event behind the NCCL barrier)
      CUDA_OK(cudaStreamWaitEvent(L.side_stream, L.chunk_events[s.chunk], 0));      // this rank's own chunk (and signal) first
      if (scoped) {
        const P3dExchange& ge = tp->steps[i - 1].ex;      // the chunk's exchange step precedes its consumer
        CUDA_OK(p3d::launch_flag_sync_mask(L.peer_flags_dev, L.comm->rank, L.comm->size, 0u, 0, L.scope_mask[ge.comm], chunk_epoch[s.chunk],
                                           L.side_stream));
        group_comm = ge.comm; group_last_epoch = chunk_epoch[s.chunk];
        note_reads(s.st, true);
      } else if (L.flagbar) {
        CUDA_OK(p3d::launch_flag_wait(L.peer_flags_dev, L.comm->rank, L.comm->size, chunk_epoch[s.chunk], L.side_stream));
      }
      if (!launch(s.st, L.side_stream, s.chunk + 1 < nchunks ? side_sms : 0)) return false;
      used_side = true;
      continue;
    }
    // Peer-to-peer plans: the stage in front of an exchange stores into the peers' receive buffer.  If that buffer
    // was handed to a consumer stage since the last barrier, a peer may still be reading it (write-after-read
    // across ranks): barrier first.  Decided from the exchange steps alone, identically on every rank -- a rank
    // whose producing stage is empty issues the same barrier when it reaches the exchange.  (Chunks after the first
    // of a pipelined group store into other regions of the buffer their group was already cleared for.)
    // With scoped synchronisation the rule is exact instead: wait for the target peers' release epoch of that buffer.
    const P3dExchange* nex = s.is_exchange ? &s.ex : (i + 1 < nsteps && tp->steps[i + 1].is_exchange ? &tp->steps[i + 1].ex : nullptr);
    if (scoped) {
      if (nex && nex->p2p && !s.is_exchange && s.chunk <= 0 && !hazard_wait(*nex)) return false;
    } else if (nex && nex->p2p && !pre_done && L.dirty[nex->recvbuf] && s.chunk <= 0) {
      if (timed) { cudaEventRecord(get_event(nev++), st); slots.push_back(nex->timer); is_ex.push_back(1); }
      if (!world_barrier(st)) return false;
    }
    if (nex) pre_done = !s.is_exchange;
    if (timed) { cudaEventRecord(get_event(nev++), st); }
    if (s.is_exchange) {
      if (s.ex.p2p && scoped) {
        any_p2p = true;
        const unsigned e = ++L.bar_epoch;
        if (s.chunk >= 0) {      // chunk of a pipelined group: signal only, the consumer waits on the side stream
          chunk_epoch[s.chunk] = e;
          CUDA_OK(p3d::launch_flag_signal(L.peer_flags_dev, L.comm->rank, L.comm->size, e, st));
        } else {                 // signal to every rank, wait for the ranks of this exchange's communicator
          CUDA_OK(p3d::launch_flag_sync_mask(L.peer_flags_dev, L.comm->rank, L.comm->size, e, 1, L.scope_mask[s.ex.comm], e, st));
          L.passed[s.ex.comm] = e;
        }
        assign_release(e);
      } else if (s.chunk >= 0 && s.ex.p2p && L.flagbar) {
        // chunk barrier, first half: tell every rank that this rank's producer chunk has been stored; the main stream goes
        // straight on to the next producer chunk, the consumer waits for all ranks' flags on the side stream
        chunk_epoch[s.chunk] = ++L.bar_epoch;
        CUDA_OK(p3d::launch_flag_signal(L.peer_flags_dev, L.comm->rank, L.comm->size, chunk_epoch[s.chunk], st));
      } else if (!run_exchange(s.ex, st)) return false;
      if (s.chunk >= 0) CUDA_OK(cudaEventRecord(L.chunk_events[s.chunk], st));
      if (s.chunk < 0 || s.chunk + 1 == nchunks) L.dirty[s.ex.recvbuf] = true;
      slots.push_back(s.ex.timer); is_ex.push_back(1);
    } else {
      // (the first producer chunk has no consumer beside it yet: it takes the whole GPU)
      if (!launch(s.st, st, s.chunk > 0 ? sms - side_sms : 0)) return false;
      if (scoped) note_reads(s.st, false);
      slots.push_back(s.st.timer); is_ex.push_back(0);
    }
  }
  if (!join_group()) return false;      // everything behind this point (epilogues, copies, the next call) is ordered after the side stream
  if (scoped && any_p2p) {
    // end of the transform: one more signal, so that every read of a work buffer in this call has a release epoch behind it
    const unsigned e = ++L.bar_epoch;
    CUDA_OK(p3d::launch_flag_signal(L.peer_flags_dev, L.comm->rank, L.comm->size, e, st));
    assign_release(e);
  }
  if (cheby) {
    // p3dfft_cheby epilogue, ftran.F90:408-451
    const double norm = 1.0 / ((double)d.nx * (double)d.ny * (double)(d.nzc - 1));
    const double lfac = 4.0 / Lz;
    if (timed) cudaEventRecord(get_event(nev++), st);
    for (int v = 0; v < nv; v++) {
      char* o = (char*)dout + (size_t)v * dim_cplx * CSIZE;
      long long ncol = (long long)d.iisize * d.jjsize;
      cudaError_t e = d.stride1 ? p3d::launch_cheby<real_t>(o, ncol, d.nzc, 1, d.nzc, norm, lfac, st)
                                : p3d::launch_cheby<real_t>(o, ncol, d.nzc, ncol, 1, norm, lfac, st);
      if (e != cudaSuccess) { report(true, "P3DFFT(B200): cheby launch failed: %s", cudaGetErrorString(e)); return false; }
      L.launches++;
    }
    slots.push_back(8); is_ex.push_back(0);
  }
  if (timed) cudaEventRecord(get_event(nev++), st);
  if (!out_dev && !copy_d2h(out, dout, out_bytes, st)) return false;
  if (!L.async || !in_dev || !out_dev || exchange_ms) {
    CUDA_OK(cudaStreamSynchronize(st));
    if (timed) {
      for (size_t i = 0; i + 1 < nev; i++) {
        float ms = 0; cudaEventElapsedTime(&ms, L.events[i], L.events[i + 1]);
        int slot = slots[i];
        if (slot >= 1 && slot <= 12) L.timers[slot - 1] += ms * 1e-3;
        if (exchange_ms && is_ex[i]) *exchange_ms += ms;
      }
    }
  }
  return true;
}

// Runs one transform.
bool run_transform(bool backward, const void* in, void* out, const char* op, int nv, long long dim_real,
                   long long dim_cplx, bool cheby, double Lz) {
  if (!alloc_work(nv)) return false;
  p3d::TransformPlan* tp = get_plan(backward, op, nv, dim_real, dim_cplx);
  if (!tp) return false;
  const p3d::Decomp& d = L.d;
  const size_t real_elems = (size_t)d.nx * d.jisize * d.kjsize, cplx_elems = (size_t)d.iisize * d.jjsize * d.nzc;
  const size_t real_bytes = ((size_t)(nv - 1) * dim_real + real_elems) * sizeof(real_t);
  const size_t cplx_bytes = ((size_t)(nv - 1) * dim_cplx + cplx_elems) * CSIZE;
  // p3dfft_cheby normalises by itself (ftran.F90:408-413): the user scale does not apply to it
  const double sc = cheby ? 1.0 : (backward ? L.scale_bwd : L.scale_fwd);
  return run_plan(tp, in, out, backward ? cplx_bytes : real_bytes, backward ? real_bytes : cplx_bytes, nv, dim_cplx,
                  cheby, Lz, nullptr, sc);
}

// p3dfft_ftran_r2c_1d and the rtran_* transposes: plans cached by key
p3d::TransformPlan* get_aux_plan(int key) {
  auto it = L.aux_plans.find(key);
  if (it != L.aux_plans.end()) return &it->second;
  p3d::TransformPlan tp = key == 100 ? p3d::build_r2c_1d_plan(L.d)
                                     : p3d::build_rtran_plan(L.d, key / 2, (key & 1) != 0, (int)sizeof(real_t));
  if (!tp.error.empty()) { report(true, "%s", tp.error.c_str()); return nullptr; }
  if (!finalize_plan(tp)) return nullptr;
  return &L.aux_plans.emplace(key, std::move(tp)).first->second;
}

bool run_rtran(int which, const void* src, void* dst, int* dstart, int* dend, int* dsize, double* t) {
  const bool multi = L.comm && L.comm->size > 1;
  if (multi && !alloc_work(L.nv_preset, true)) return false;
  p3d::TransformPlan* tp = get_aux_plan(which * 2 + (L.p2p ? 1 : 0));
  if (!tp) return false;
  double ex_ms = 0.0;
  const size_t ib = (size_t)p3d::rtran_elems(L.d, which, false) * sizeof(real_t);
  const size_t ob = (size_t)p3d::rtran_elems(L.d, which, true) * sizeof(real_t);
  if (!run_plan(tp, src, dst, ib, ob, 1, 0, false, 0.0, t ? &ex_ms : nullptr)) return false;
  if (t) *t += ex_ms * 1e-3;          // the reference adds the MPI_Wtime spent in the alltoallv (module.F90:1085-1095)
  if (dstart && dend && dsize) p3d::rtran_dims(L.d, which, dstart, dend, dsize);
  return true;
}


// ------------------------------------------------------------------------------------
// A caller's own MPI.  The reference takes a Fortran MPI handle (MPI_Comm_c2f(MPI_COMM_WORLD), driver_sine.c:126;
// setup.F90:178-261 builds its Cartesian communicators from it).  When `comm` is not a handle of
// p3dfft_b200_comm_create and the process has an MPI loaded, the library borrows four entry points from it through
// dlsym(RTLD_DEFAULT) -- MPI_Comm_f2c, MPI_Comm_rank, MPI_Comm_size, MPI_Bcast -- to learn rank and size and to
// distribute the ncclUniqueId, then creates its own communicator.  Nothing else of MPI is used; the library does not link
// it.  Handle types differ between MPI families: MPICH and its derivatives (Intel MPI, MVAPICH, Cray) use 32-bit
// integers (MPI_BYTE = 0x4c00010d), Open MPI uses pointers to global objects (MPI_BYTE = &ompi_mpi_byte); both pass
// through pointer-sized arguments here.
// ------------------------------------------------------------------------------------
std::map<int, int> g_mpi_comms;      // Fortran MPI handle -> library handle

int local_device_for(int rank) {
  const char* names[] = {"OMPI_COMM_WORLD_LOCAL_RANK", "MV2_COMM_WORLD_LOCAL_RANK", "MPI_LOCALRANKID", "SLURM_LOCALID", "LOCAL_RANK", nullptr};
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev <= 0) { cudaGetLastError(); return -1; }
  for (int i = 0; names[i]; i++)
    if (const char* v = getenv(names[i])) return atoi(v) % ndev;
  return rank % ndev;
}

// returns a library handle (> 0), 0 when no MPI is loaded, < 0 on failure
int comm_from_mpi(int fhandle) {
  auto it = g_mpi_comms.find(fhandle);
  if (it != g_mpi_comms.end() && g_comms.count(it->second)) return it->second;
  typedef uintptr_t (*f2c_t)(int);
  typedef int (*rank_t)(uintptr_t, int*);
  typedef int (*bcast_t)(void*, int, uintptr_t, int, uintptr_t);
  typedef int (*inited_t)(int*);
  f2c_t f2c = (f2c_t)dlsym(RTLD_DEFAULT, "MPI_Comm_f2c");
  rank_t crank = (rank_t)dlsym(RTLD_DEFAULT, "MPI_Comm_rank"), csize = (rank_t)dlsym(RTLD_DEFAULT, "MPI_Comm_size");
  bcast_t bcast = (bcast_t)dlsym(RTLD_DEFAULT, "MPI_Bcast");
  inited_t inited = (inited_t)dlsym(RTLD_DEFAULT, "MPI_Initialized");
  if (!f2c || !crank || !csize || !bcast) return 0;
  int flag = 1;
  if (inited && (inited(&flag) != 0 || !flag)) return 0;
  void* ompi_byte = dlsym(RTLD_DEFAULT, "ompi_mpi_byte");
  uintptr_t c = f2c(fhandle);
  if (!ompi_byte) c &= 0xffffffffu;                                   // integer handles: only the low word is defined
  const uintptr_t byte_t = ompi_byte ? (uintptr_t)ompi_byte : (uintptr_t)0x4c00010d;
  int rank = -1, size = 0;
  if (crank(c, &rank) != 0 || csize(c, &size) != 0 || size < 1 || rank < 0 || rank >= size) return -1;
  unsigned char id[P3DFFT_B200_UNIQUE_ID_BYTES];
  memset(id, 0, sizeof id);
  if (size > 1) {
    if (rank == 0 && p3dfft_b200_get_unique_id(id) != 0) memset(id, 0, sizeof id);      // all-zero id: the others fail with rank 0
    if (bcast(id, (int)sizeof id, byte_t, 0, c) != 0) return -1;
    bool zero = true;
    for (unsigned char b : id) zero = zero && b == 0;
    if (zero) return -1;
  }
  const int h = p3dfft_b200_comm_create(rank, size, id, local_device_for(rank));
  if (h > 0) g_mpi_comms[fhandle] = h;
  return h;
}

bool check_set() {
  if (!L.set) {
    // module.F90:231, ftran.F90:506-509: message and return
    report(false, "P3DFFT error: call setup before other routines");
    return false;
  }
  return true;
}

}  // namespace

// ======================================================================================
// reference entry points
// ======================================================================================
extern "C" {

void p3dfft_setup(int* dims, int* nx, int* ny, int* nz, int* comm, int* nxc, int* nyc, int* nzc, int* ow,
                  int* memsize) {
  if (L.set) {     // setup.F90:130-135
    report(true, "P3DFFT Setup error: the problem is already initialized. \n"
                 "Currently multiple setups not supported.\n"
                 "Quit the library using p3dfft_clean before initializing another setup");
    return;
  }
  CommCtx* cc = nullptr;
  if (comm) { auto it = g_comms.find(*comm); if (it != g_comms.end()) cc = it->second; }
  if (comm && !cc) {
    // Not a handle of p3dfft_b200_comm_create: with an MPI loaded in the process this is the Fortran MPI handle of an unmodified
    // caller (MPI_Comm_c2f(MPI_COMM_WORLD): 0x44000000 with MPICH and its derivatives, 0 with Open MPI) and the library
    // bootstraps itself from that MPI (comm_from_mpi).  Without one, 0 is the implicit one-rank communicator; any other value is
    // an error: running on as a single rank would silently give every process of a job its own full-size transform.
    const bool try_mpi = *comm != 0 || dlsym(RTLD_DEFAULT, "ompi_mpi_comm_world") != nullptr;
    const int h = try_mpi ? comm_from_mpi(*comm) : 0;
    if (h > 0) cc = g_comms[h];
    else if (h < 0 || *comm != 0) {
      report(true, "P3DFFT(B200) setup error: communicator %d is neither a handle of p3dfft_b200_comm_create nor a "
                   "communicator of an MPI loaded in this process%s (see INTEGRATION.md)", *comm,
             h < 0 ? " -- the bootstrap through that MPI failed" : "");
      return;
    }
  }
  const int rank = cc ? cc->rank : 0, ntasks = cc ? cc->size : 1;
  std::string err = L.d.init(*nx, *ny, *nz, dims[0], dims[1], rank, ntasks, nxc ? *nxc : *nx, nyc ? *nyc : *ny,
                             nzc ? *nzc : *nz, L.dims_c != 0, L.stride1 != 0);
  if (!err.empty()) { report(true, "%s", err.c_str()); return; }
  L.overwrite = ow ? (*ow != 0) : true;
  L.comm = cc;
  if (cc && cc->device >= 0) cudaSetDevice(cc->device);
  // a blocking stream: ordered after work the caller queued on the legacy default stream
  if (!L.own_stream && cudaStreamCreate(&L.own_stream) != cudaSuccess) {
    report(true, "P3DFFT(B200): no usable CUDA device (%s)", cudaGetErrorString(cudaGetLastError()));
    return;
  }
  L.row = L.col = nullptr;
  if (ntasks > 1) {
    if (!g_nccl.load()) return;
    // mpi_comm_row: same jpid ordered by ipid; mpi_comm_col: same ipid ordered by jpid (setup.F90:245-261)
    if (L.d.iproc > 1 && g_nccl.CommSplit(cc->world, L.d.jpid, L.d.ipid, &L.row, nullptr) != ncclSuccess) {
      report(true, "P3DFFT(B200): ncclCommSplit(row) failed"); return;
    }
    if (L.d.jproc > 1 && g_nccl.CommSplit(cc->world, L.d.ipid, L.d.jpid, &L.col, nullptr) != ncclSuccess) {
      report(true, "P3DFFT(B200): ncclCommSplit(col) failed"); return;
    }
  }
  // switches: an environment variable wins; without it the value set through the API stays (generic, plain, p2p, row bytes)
  // or the default applies (barrier kind, pipelined tail) -- so a later setup in the same process starts from the defaults
  auto env_int = [](const char* name, int dflt) { const char* v = getenv(name); return v ? atoi(v) : dflt; };
  if (getenv("P3DFFT_B200_GENERIC")) L.force_generic = true;
  if (getenv("P3DFFT_B200_PLAIN")) L.plain_layout = true;
  L.want_p2p = env_int("P3DFFT_B200_P2P", L.api_p2p ? 1 : 0) != 0;
  if (getenv("P3DFFT_B200_ROWB")) L.force_row_bytes = atoi(getenv("P3DFFT_B200_ROWB"));
  L.want_flagbar = env_int("P3DFFT_B200_FLAGBAR", P3D_DEFAULT_FLAGBAR) != 0;
  L.want_scoped = env_int("P3DFFT_B200_SCOPED", P3D_DEFAULT_SCOPED) != 0;
  L.overlap = env_int("P3DFFT_B200_OVERLAP", P3D_DEFAULT_OVERLAP);
  L.overlap_forced = getenv("P3DFFT_B200_OVERLAP") != nullptr;
  L.overlap_sms_dir[0] = env_int("P3DFFT_B200_OVERLAP_SMS_FWD", 0);
  L.overlap_sms_dir[1] = env_int("P3DFFT_B200_OVERLAP_SMS_BWD", 0);
  L.overlap_shape.clear();
  {
    const char* sh = getenv("P3DFFT_B200_OVERLAP_SHAPE");
    if (!sh) sh = P3D_DEFAULT_OVERLAP_SHAPE;
    for (const char* q = sh; q && *q;) {
      L.overlap_shape.push_back(atoi(q));
      q = strchr(q, ',');
      if (q) q++;
    }
    if (L.overlap_shape.size() == 1) L.overlap_shape.clear();
    if (!L.overlap_shape.empty() && L.overlap > 1) L.overlap = (int)L.overlap_shape.size();
  }
  L.overlap_sms = env_int("P3DFFT_B200_OVERLAP_SMS", P3D_DEFAULT_OVERLAP_SMS);
  L.row_bytes_plan = L.force_row_bytes;
  p3d::fast_reload_switches();
  for (int i = 0; i < 12; i++) L.timers[i] = 0.0;      // setup.F90:144
  L.procmap.init(L.d);
  L.nv_preset = 0;
  L.set = true;
  if (!alloc_work(1)) { L.set = false; return; }
  if (memsize) { memsize[0] = L.d.memsize[0]; memsize[1] = L.d.memsize[1]; memsize[2] = L.d.memsize[2]; }
  if (rank == 0 && getenv("P3DFFT_B200_VERBOSE"))
    fprintf(stderr, "P3DFFT(B200): %d x %d x %d on a %d x %d grid%s\n", *nx, *ny, *nz, dims[0], dims[1],
            L.stride1 ? ", stride-1 layout" : "");
}

void p3dfft_get_dims(int* istart, int* iend, int* isize, int* conf) {
  if (!check_set()) return;
  L.d.get_dims(istart, iend, isize, *conf);
}

void p3dfft_ftran_r2c(real_t* A, real_t* B, unsigned char* op) {
  if (!check_set()) return;
  const p3d::Decomp& d = L.d;
  run_transform(false, A, B, (const char*)op, 1, (long long)d.nx * d.jisize * d.kjsize,
                (long long)d.iisize * d.jjsize * d.nzc, false, 0.0);
}

void p3dfft_btran_c2r(real_t* A, real_t* B, unsigned char* op) {
  if (!check_set()) return;
  const p3d::Decomp& d = L.d;
  run_transform(true, A, B, (const char*)op, 1, (long long)d.nx * d.jisize * d.kjsize,
                (long long)d.iisize * d.jjsize * d.nzc, false, 0.0);
}

void p3dfft_ftran_r2c_many(real_t* A, int* dim_in, real_t* B, int* dim_out, int* nv, unsigned char* op) {
  if (!check_set()) return;
  const p3d::Decomp& d = L.d;
  // ftran.F90:118-129: message only in the reference; here the call is also abandoned
  if ((long long)*dim_in < (long long)d.nx * d.jisize * d.kjsize) {
    report(false, "%d: ftran error: input array dimensions are too low: %d while expecting %lld", d.rank, *dim_in,
           (long long)d.nx * d.jisize * d.kjsize);
    return;
  }
  if ((long long)*dim_out < (long long)d.nzc * d.jjsize * d.iisize) {
    report(false, "%d: ftran error: output array dimensions are too low: %d while expecting %lld", d.rank, *dim_out,
           (long long)d.nzc * d.jjsize * d.iisize);
    return;
  }
  if (*nv <= 0) return;
  run_transform(false, A, B, (const char*)op, *nv, *dim_in, *dim_out, false, 0.0);
}

void p3dfft_btran_c2r_many(real_t* A, int* dim_in, real_t* B, int* dim_out, int* nv, unsigned char* op) {
  if (!check_set()) return;
  const p3d::Decomp& d = L.d;
  if ((long long)*dim_in < (long long)d.nzc * d.jjsize * d.iisize) {      // btran.F90:118-129
    report(false, "%d: btran error: input array dimensions are too low: %d while expecting %lld", d.rank, *dim_in,
           (long long)d.nzc * d.jjsize * d.iisize);
    return;
  }
  if ((long long)*dim_out < (long long)d.nx * d.jisize * d.kjsize) {
    report(false, "%d: btran error: output array dimensions are too low: %d while expecting %lld", d.rank, *dim_out,
           (long long)d.nx * d.jisize * d.kjsize);
    return;
  }
  if (*nv <= 0) return;
  run_transform(true, A, B, (const char*)op, *nv, *dim_out, *dim_in, false, 0.0);
}

void p3dfft_cheby(real_t* A, real_t* B, real_t* Lz) {
  if (!check_set()) return;
  const p3d::Decomp& d = L.d;
  run_transform(false, A, B, "ffc", 1, (long long)d.nx * d.jisize * d.kjsize, (long long)d.iisize * d.jjsize * d.nzc,
                true, (double)*Lz);
}

void p3dfft_cheby_many(real_t* A, int* dim_in, real_t* B, int* dim_out, int* nv, real_t* Lz) {
  if (!check_set()) return;
  if (*nv <= 0) return;
  run_transform(false, A, B, "ffc", *nv, *dim_in, *dim_out, true, (double)*Lz);
}

void p3dfft_clean(void) {
  // module.F90:309-420: destroy plans, free buffers, mpi_set = .false.
  if (!L.set) return;
  cudaStreamSynchronize(L.stream());
  if (!L.peer_buf.empty()) {       // nobody frees while a peer may still store into its buffers
    world_barrier(L.stream());
    cudaStreamSynchronize(L.stream());
    close_peer_maps();
    world_barrier(L.stream());     // ... nor before every importer has closed its mapping of them
    cudaStreamSynchronize(L.stream());
  }
  if (L.bar_scratch) { cudaFree(L.bar_scratch); L.bar_scratch = nullptr; }
  if (L.bar_flags) { cudaFree(L.bar_flags); L.bar_flags = nullptr; }
  if (L.peer_flags_dev) { cudaFree(L.peer_flags_dev); L.peer_flags_dev = nullptr; }
  if (L.spec_dev) { cudaFree(L.spec_dev); L.spec_dev = nullptr; L.spec_bins = 0; }
  if (L.side_stream) { cudaStreamSynchronize(L.side_stream); cudaStreamDestroy(L.side_stream); L.side_stream = nullptr; }
  for (auto ev : L.chunk_events) cudaEventDestroy(ev);
  L.chunk_events.clear();
  if (L.side_done) { cudaEventDestroy(L.side_done); L.side_done = nullptr; }
  for (int b = P3D_BUF_A; b <= P3D_BUF_C; b++) if (L.buf[b]) { cudaFree(L.buf[b]); L.buf[b] = nullptr; }
  if (L.stage_in) { cudaFree(L.stage_in); L.stage_in = nullptr; L.stage_in_bytes = 0; }
  if (L.stage_out) { cudaFree(L.stage_out); L.stage_out = nullptr; L.stage_out_bytes = 0; }
  g_pipe.release();
  for (auto& kv : L.twiddles) cudaFree(kv.second);
  L.twiddles.clear();
  for (auto& kv : L.fast_twiddles) cudaFree(kv.second);
  L.fast_twiddles.clear();
  L.plans.clear();
  if (L.row) { g_nccl.CommDestroy(L.row); L.row = nullptr; }
  if (L.col) { g_nccl.CommDestroy(L.col); L.col = nullptr; }
  L.nv_preset = 0;
  L.work_elems_alloc = 0;
  L.rtran_sized = false;
  L.aux_plans.clear();
  for (bool& d : L.dirty) d = false;
  L.set = false;
}

void get_timers(double* timers) { for (int i = 0; i < 12; i++) timers[i] = L.timers[i]; }   // module.F90:726
void set_timers(void) { for (int i = 0; i < 12; i++) L.timers[i] = 0.0; }                   // module.F90:742

// ======================================================================================
// remaining public routines of the reference's Fortran module (module.F90:178-186)
// ======================================================================================
void p3dfft_ftran_r2c_1d(void* rXgYZ, void* cXgYZ) {      // ftran.F90:787-814
  if (!check_set()) return;
  const p3d::Decomp& d = L.d;
  p3d::TransformPlan* tp = get_aux_plan(100);
  if (!tp || tp->steps.empty()) return;
  const size_t lines = (size_t)d.jisize * d.kjsize;
  run_plan(tp, rXgYZ, cXgYZ, lines * d.nx * sizeof(real_t), lines * d.nxhp * CSIZE, 1, 0, false, 0.0, nullptr);
}

void p3dfft_b200_rtran_x2y(const void* src, void* dst, int* dstart, int* dend, int* dsize, double* t) {
  if (check_set()) run_rtran(p3d::RTRAN_X2Y, src, dst, dstart, dend, dsize, t);
}
void p3dfft_b200_rtran_y2x(const void* src, void* dst, int* dstart, int* dend, int* dsize, double* t) {
  if (check_set()) run_rtran(p3d::RTRAN_Y2X, src, dst, dstart, dend, dsize, t);
}
void p3dfft_b200_rtran_x2z(const void* src, void* dst, int* dstart, int* dend, int* dsize, double* t) {
  if (check_set()) run_rtran(p3d::RTRAN_X2Z, src, dst, dstart, dend, dsize, t);
}
void p3dfft_b200_rtran_z2x(const void* src, void* dst, int* dstart, int* dend, int* dsize, double* t) {
  if (check_set()) run_rtran(p3d::RTRAN_Z2X, src, dst, dstart, dend, dsize, t);
}

void p3dfft_get_mpi_info(int* taskid, int* ntasks, int* comm) {      // module.F90:280-297
  if (!check_set()) return;
  *taskid = L.d.rank; *ntasks = L.d.numtasks;
  *comm = 0;
  for (auto& kv : g_comms) if (kv.second == L.comm) *comm = kv.first;
  for (auto& kv : g_mpi_comms) if (kv.second == *comm) *comm = kv.first;      // the caller's own MPI handle, as the reference returns it
}

int p3dfft_b200_proc_id2coords(int id, int* ipid, int* jpid) {
  if (!check_set() || id < 0 || id >= L.procmap.nproc()) return -1;
  *ipid = L.procmap.id2coords[2 * id]; *jpid = L.procmap.id2coords[2 * id + 1];
  return 0;
}
int p3dfft_b200_proc_coords2id(int ipid, int jpid) { return check_set() ? L.procmap.coords2id(ipid, jpid) : -1; }
int p3dfft_b200_proc_dims(int conf, int id, int* out9) {
  if (!check_set() || conf < 1 || conf > 2 || id < 0 || id >= L.procmap.nproc()) return -1;
  for (int k = 1; k <= 9; k++) out9[k - 1] = L.procmap.pd(conf, k, id);
  return 0;
}
int p3dfft_b200_proc_neighb(int base, int orient, int direction) {
  return check_set() ? L.procmap.neighb(base, orient, direction) : -1;
}
int p3dfft_b200_get_proc_parts(int base_x, int base_y, int base_z, int size_x, int size_y, int size_z, int conf,
                               int* parts, int* ierr) {
  if (!check_set()) { if (ierr) *ierr = -2; return 0; }
  int e = 0;
  const int n = L.procmap.parts(base_x, base_y, base_z, size_x, size_y, size_z, conf, parts, &e);
  if (ierr) *ierr = e;
  return n;
}

// ======================================================================================
// extensions
// ======================================================================================
int p3dfft_b200_build_flags(void) {
  int f = 0;
#ifdef SINGLE_PREC
  f |= 1;
#endif
#ifdef STRIDE1
  f |= 2;
#endif
#ifdef DIMS_C
  f |= 4;
#endif
  return f;
}

void p3dfft_b200_set_layout(int stride1, int dims_c) { L.stride1 = stride1 != 0; L.dims_c = dims_c != 0; }

int p3dfft_b200_get_unique_id(void* id128) {
  if (!g_nccl.load()) return -1;
  ncclUniqueId id;
  if (g_nccl.GetUniqueId(&id) != ncclSuccess) return -2;
  static_assert(sizeof(ncclUniqueId) == P3DFFT_B200_UNIQUE_ID_BYTES, "unique id size");
  memcpy(id128, &id, sizeof id);
  return 0;
}

int p3dfft_b200_comm_create(int rank, int size, const void* id128, int device) {
  if (size < 1 || rank < 0 || rank >= size) return -3;
  CommCtx* c = new CommCtx;
  c->rank = rank; c->size = size; c->device = device;
  if (device >= 0 && cudaSetDevice(device) != cudaSuccess) { delete c; return -4; }
  if (size > 1) {
    if (!g_nccl.load()) { delete c; return -1; }
    ncclUniqueId id; memcpy(&id, id128, sizeof id);
    ncclResult_t r = g_nccl.CommInitRank(&c->world, size, id, rank);
    if (r != ncclSuccess) { report(false, "P3DFFT(B200): ncclCommInitRank: %s", g_nccl.GetErrorString(r)); delete c; return -5; }
  }
  int h = g_next_handle++;
  g_comms[h] = c;
  return h;
}

void p3dfft_b200_comm_destroy(int handle) {
  auto it = g_comms.find(handle);
  if (it == g_comms.end()) return;
  if (it->second->world) g_nccl.CommDestroy(it->second->world);
  if (L.comm == it->second) L.comm = nullptr;
  delete it->second;
  g_comms.erase(it);
}

void p3dfft_b200_set_scale(double forward, double backward) { L.scale_fwd = forward; L.scale_bwd = backward; }

void p3dfft_b200_spectrum(const void* B, double factor, double* E, int kmax) {
  if (!check_set()) return;
  if (kmax < 0 || !E) { report(false, "P3DFFT(B200): spectrum needs kmax >= 0 and an output array"); return; }
  const p3d::Decomp& d = L.d;
  cudaStream_t st = L.stream();
  const size_t cplx_bytes = (size_t)d.iisize * d.jjsize * d.nzc * CSIZE;
  const void* dB = B;
  auto run = [&]() -> bool {
    if (!is_device_ptr(B)) {
      if (L.stage_in_bytes < cplx_bytes) {
        if (L.stage_in) cudaFree(L.stage_in);
        L.stage_in = nullptr; L.stage_in_bytes = 0;
        CUDA_OK(cudaMalloc(&L.stage_in, cplx_bytes)); L.stage_in_bytes = cplx_bytes;
      }
      if (!copy_h2d(L.stage_in, B, cplx_bytes, st)) return false;
      dB = L.stage_in;
    }
    if (L.spec_bins < kmax + 1) {
      if (L.spec_dev) cudaFree(L.spec_dev);
      L.spec_dev = nullptr; L.spec_bins = 0;
      CUDA_OK(cudaMalloc(&L.spec_dev, sizeof(double) * (size_t)(kmax + 1))); L.spec_bins = kmax + 1;
    }
    CUDA_OK(cudaMemsetAsync(L.spec_dev, 0, sizeof(double) * (size_t)(kmax + 1), st));
    p3d::SpecJob j;
    memset(&j, 0, sizeof j);
    // physical extents / strides of the wavenumber array (get_dims conf 2)
    const int ext[3] = {d.iisize, d.jjsize, d.nzc};
    long long str[3];
    if (d.stride1) { str[2] = 1; str[1] = d.nzc; str[0] = (long long)d.nzc * d.jjsize; }
    else { str[0] = 1; str[1] = d.iisize; str[2] = (long long)d.iisize * d.jjsize; }
    const int order[3] = {d.stride1 ? 2 : 0, 1, d.stride1 ? 0 : 2};      // contiguous axis first
    for (int i = 0; i < 3; i++) { j.axis[i] = order[i]; j.ext[i] = ext[order[i]]; j.stride[i] = str[order[i]]; }
    j.start[0] = d.iistart - 1; j.start[1] = d.jjstart - 1; j.start[2] = 0;
    j.n[0] = d.nx; j.nc[0] = d.nxhpc; j.nch[0] = d.nxhpc;                 // x: the first nxhpc modes, never folded
    j.n[1] = d.ny; j.nc[1] = d.nyc; j.nch[1] = d.nycph;
    j.n[2] = d.nz; j.nc[2] = d.nzc; j.nch[2] = d.nzcph;
    j.kmax = kmax; j.f2 = factor * factor;
    cudaError_t e = p3d::launch_spectrum<real_t>(dB, j, L.spec_dev, st);
    if (e != cudaSuccess) { report(true, "P3DFFT(B200): spectrum launch failed: %s", cudaGetErrorString(e)); return false; }
    L.launches++;
    if (L.comm && L.comm->size > 1)      // MPI_Reduce(..., MPI_SUM, root) of driver_spec.c:381 -- here every rank gets the sum
      NCCL_OK(g_nccl.AllReduce(L.spec_dev, L.spec_dev, (size_t)(kmax + 1), ncclDouble, ncclSum, L.comm->world, st));
    CUDA_OK(cudaMemcpyAsync(E, L.spec_dev, sizeof(double) * (size_t)(kmax + 1),
                            is_device_ptr(E) ? cudaMemcpyDeviceToDevice : cudaMemcpyDeviceToHost, st));
    CUDA_OK(cudaStreamSynchronize(st));
    return true;
  };
  run();
}

void p3dfft_b200_set_error_mode(int mode) { g_error_mode = mode; }

int p3dfft_b200_last_error(char* buf, int buflen) {
  int n = (int)g_last_error.size();
  if (buf && buflen > 0) { snprintf(buf, buflen, "%s", g_last_error.c_str()); }
  g_last_error.clear();
  return n;
}

void p3dfft_b200_set_stream(void* s) { L.user_stream = (cudaStream_t)s; L.has_user_stream = true; }
void p3dfft_b200_reset_stream(void) { L.user_stream = nullptr; L.has_user_stream = false; }
void p3dfft_b200_force_generic(int on) { L.force_generic = on != 0; }
void p3dfft_b200_set_p2p(int on) { L.api_p2p = L.want_p2p = on != 0; }
int p3dfft_b200_p2p_active(void) { return L.p2p ? 1 : 0; }
void p3dfft_b200_row_bytes(int rb) {      // 0: planner's rule; 64 / 128: forced.  Takes effect at the next p3dfft_setup
  L.force_row_bytes = (rb == 64 || rb == 128) ? rb : 0;
}
void p3dfft_b200_plain_layout(int on) {
  if (L.set && (on != 0) != L.plain_layout) { cudaStreamSynchronize(L.stream()); L.plans.clear(); L.nv_preset = 0; }
  L.plain_layout = on != 0;
}
long long p3dfft_b200_fast_launch_count(int reset) { long long n = L.fast_launches; if (reset) L.fast_launches = 0; return n; }
void p3dfft_b200_set_async(int a) { L.async = a != 0; }
void p3dfft_b200_sync(void) { cudaStreamSynchronize(L.stream()); }
long long p3dfft_b200_launch_count(int reset) { long long n = L.launches; if (reset) L.launches = 0; return n; }

int p3dfft_b200_plan_decomp(const int* dims, int nx, int ny, int nz, int rank, int nxc, int nyc, int nzc, int flags,
                            p3dfft_b200_decomp* o) {
  p3d::Decomp d;
  std::string err = d.init(nx, ny, nz, dims[0], dims[1], rank, dims[0] * dims[1], nxc, nyc, nzc, (flags & 4) != 0,
                           (flags & 2) != 0);
  if (!err.empty()) { g_last_error = err; return -1; }
  o->nx = d.nx; o->ny = d.ny; o->nz = d.nz; o->nxc = d.nxc; o->nyc = d.nyc; o->nzc = d.nzc;
  o->nxhp = d.nxhp; o->nxhpc = d.nxhpc; o->nycph = d.nycph; o->nzcph = d.nzcph;
  o->iproc = d.iproc; o->jproc = d.jproc; o->ipid = d.ipid; o->jpid = d.jpid;
  o->iistart = d.iistart; o->iiend = d.iiend; o->iisize = d.iisize;
  o->jistart = d.jistart; o->jiend = d.jiend; o->jisize = d.jisize;
  o->jjstart = d.jjstart; o->jjend = d.jjend; o->jjsize = d.jjsize;
  o->kjstart = d.kjstart; o->kjend = d.kjend; o->kjsize = d.kjsize;
  o->padi_work = d.padi_work; o->padi = d.padi;
  for (int i = 0; i < 3; i++) o->memsize[i] = d.memsize[i];
  o->nm = d.nm;
  o->work_elems = d.work_elems(1, (flags & 8) ? 0 : p3d::pick_W(ny, nz, (flags & 1) ? 8 : 16, (flags & 32) ? 64 : (flags & 64) ? 128 : 0));
  return 0;
}

struct P3dStepC { int32_t is_exchange; int32_t pad_; P3dStage st; P3dExchange ex; };

int p3dfft_b200_sizeof_step(void) { return (int)sizeof(P3dStepC); }

int p3dfft_b200_plan_steps(const int* dims, int nx, int ny, int nz, int rank, int nxc, int nyc, int nzc, int flags,
                           int backward, const char* op, int nv, int64_t dim_real, int64_t dim_cplx, int elem_bytes,
                           void* steps, int max_steps) {
  p3d::Decomp d;
  std::string err = d.init(nx, ny, nz, dims[0], dims[1], rank, dims[0] * dims[1], nxc, nyc, nzc, (flags & 4) != 0,
                           (flags & 2) != 0);
  if (!err.empty()) { g_last_error = err; return -1; }
  p3d::TransformPlan tp = p3d::build_plan(d, backward != 0, op, nv, dim_real, dim_cplx,
                                          (flags & 8) ? 0 : p3d::pick_W(ny, nz, 2 * elem_bytes, (flags & 32) ? 64 : (flags & 64) ? 128 : 0),
                                          (flags & 16) != 0 && !(flags & 8));
  if (!tp.error.empty()) { g_last_error = tp.error; return -1; }
  const int nchunk = (flags >> 8) & 0xff;       // pipelined tail (split_for_overlap), peer-to-peer plans only
  if (nchunk > 1) p3d::split_for_overlap(tp, nchunk, (flags & 8) ? 0 : p3d::pick_W(ny, nz, 2 * elem_bytes, (flags & 32) ? 64 : (flags & 64) ? 128 : 0));
  if ((int)tp.steps.size() > max_steps) { g_last_error = "step array too small"; return -1; }
  P3dStepC* out = (P3dStepC*)steps;
  for (size_t i = 0; i < tp.steps.size(); i++) {
    memset(&out[i], 0, sizeof out[i]);
    out[i].is_exchange = tp.steps[i].is_exchange ? 1 : 0;
    // bit 0: side stream; bits 8..: chunk + 1
    out[i].pad_ = (tp.steps[i].side ? 1 : 0) | ((tp.steps[i].chunk + 1) << 8);
    out[i].st = tp.steps[i].st;
    out[i].ex = tp.steps[i].ex;
    if (!tp.steps[i].is_exchange)
      out[i].st.tile = elem_bytes == 4 ? p3d::choose_tile<float>(out[i].st) : p3d::choose_tile<double>(out[i].st);
  }
  return (int)tp.steps.size();
}

int p3dfft_b200_plan_aux_steps(const int* dims, int nx, int ny, int nz, int rank, int nxc, int nyc, int nzc, int flags,
                               int which, int elem_bytes, void* steps, int max_steps) {
  p3d::Decomp d;
  std::string err = d.init(nx, ny, nz, dims[0], dims[1], rank, dims[0] * dims[1], nxc, nyc, nzc, (flags & 4) != 0,
                           (flags & 2) != 0);
  if (!err.empty()) { g_last_error = err; return -1; }
  if (which != 100 && (which < 0 || which > 3)) { g_last_error = "unknown auxiliary plan"; return -1; }
  p3d::TransformPlan tp = which == 100 ? p3d::build_r2c_1d_plan(d) : p3d::build_rtran_plan(d, which, (flags & 16) != 0, elem_bytes);
  if (!tp.error.empty()) { g_last_error = tp.error; return -1; }
  if ((int)tp.steps.size() > max_steps) { g_last_error = "step array too small"; return -1; }
  P3dStepC* out = (P3dStepC*)steps;
  for (size_t i = 0; i < tp.steps.size(); i++) {
    memset(&out[i], 0, sizeof out[i]);
    out[i].is_exchange = tp.steps[i].is_exchange ? 1 : 0;
    out[i].st = tp.steps[i].st;
    out[i].ex = tp.steps[i].ex;
    if (!tp.steps[i].is_exchange && tp.steps[i].st.kind != P3D_RCOPY)
      out[i].st.tile = elem_bytes == 4 ? p3d::choose_tile<float>(out[i].st) : p3d::choose_tile<double>(out[i].st);
  }
  return (int)tp.steps.size();
}

long long p3dfft_b200_plan_rtran_info(const int* dims, int nx, int ny, int nz, int rank, int which, int flags, int* dims9) {
  p3d::Decomp d;
  std::string err = d.init(nx, ny, nz, dims[0], dims[1], rank, dims[0] * dims[1], nx, ny, nz, (flags & 4) != 0, (flags & 2) != 0);
  if (!err.empty()) { g_last_error = err; return -1; }
  if (which < 0 || which > 3) { g_last_error = "unknown transpose"; return -1; }
  if (dims9) p3d::rtran_dims(d, which, dims9, dims9 + 3, dims9 + 6);
  return p3d::rtran_work_elems(d);
}

int p3dfft_b200_plan_proc_parts(const int* dims, int nx, int ny, int nz, int nxc, int nyc, int nzc, int flags, int base_x,
                                int base_y, int base_z, int size_x, int size_y, int size_z, int conf, int* parts, int* ierr) {
  p3d::Decomp d;
  std::string err = d.init(nx, ny, nz, dims[0], dims[1], 0, dims[0] * dims[1], nxc, nyc, nzc, (flags & 4) != 0, (flags & 2) != 0);
  if (!err.empty()) { g_last_error = err; return -1; }
  p3d::ProcMap pm;
  pm.init(d);
  int e = 0;
  const int n = pm.parts(base_x, base_y, base_z, size_x, size_y, size_z, conf, parts, &e);
  if (ierr) *ierr = e;
  return n;
}

int p3dfft_b200_plan_proc_neighb(const int* dims, int flags, int base, int orient, int direction) {
  p3d::Decomp d;
  std::string err = d.init(4, 4, 4, dims[0], dims[1], 0, dims[0] * dims[1], 4, 4, 4, (flags & 4) != 0, false);
  if (!err.empty()) { g_last_error = err; return -2; }
  p3d::ProcMap pm;
  pm.init(d);
  return pm.neighb(base, orient, direction);
}

}  // extern "C"
