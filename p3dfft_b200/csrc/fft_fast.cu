// Host side of the specialised stage kernels: eligibility, twiddle blocks, plan-segment ->
// run conversion and the length dispatch.  Instantiated for the library's precision only
// (SINGLE_PREC selects float, as configure --enable-single does for the reference).
#include <cuda_runtime.h>

#include <cmath>
#include <cstdlib>
#include <cstring>
#include <type_traits>
#include <vector>

#include "fast.h"
#include "fft_fast.cuh"

namespace p3d {

using namespace fast;

namespace {

bool c2c_len_ok(int n) {
  return n == 64 || n == 128 || n == 256 || n == 512 || n == 1024 || n == 2048 || n == 384 || n == 768 || n == 1536 || n == 640 || n == 1280;
}
bool x_half_ok(int h) {
  return h == 32 || h == 64 || h == 128 || h == 256 || h == 512 || h == 1024 || h == 192 || h == 384 || h == 768 || h == 320 || h == 640;
}

template <class S>
void fill_pass_tables(long double n_total, std::vector<long double>& re, std::vector<long double>& im) {
  const long double twopi = 6.283185307179586476925286766559L;
  (void)n_total;
  for (int i = 0; i < S::L - 1; i++) {
    const int R = S::r(i), M = S::m(i), NC = S::ncur(i);
    for (int q = 1; q < R; q++)
      for (int j = 0; j < M; j++) {
        long double ang = -twopi * (long double)((long long)q * j % NC) / (long double)NC;
        re.push_back(cosl(ang)); im.push_back(sinl(ang));
      }
  }
}

template <typename T, class S>
void fill_block(bool xstage, void* host) {
  std::vector<long double> re, im;
  fill_pass_tables<S>(S::N, re, im);
  if (xstage) {
    const long double twopi = 6.283185307179586476925286766559L;
    const int H = S::N, N = 2 * H;
    for (int k = 0; k < H; k++) {
      long double ang = -twopi * (long double)k / (long double)N;
      re.push_back(cosl(ang)); im.push_back(sinl(ang));
    }
  }
  T* o = reinterpret_cast<T*>(host);
  for (size_t i = 0; i < re.size(); i++) { o[2 * i] = (T)re[i]; o[2 * i + 1] = (T)im[i]; }
}

template <class S> size_t block_elems(bool xstage) { return (size_t)S::twtotal() + (xstage ? S::N : 0); }

bool is_x(int kind) { return kind == P3D_R2C || kind == P3D_C2R; }

}  // namespace

template <class F>
bool dispatch_c(int n, F&& f) {
  switch (n) {
    case 64:   f(std::integral_constant<int, 64>{});   return true;
    case 128:  f(std::integral_constant<int, 128>{});  return true;
    case 256:  f(std::integral_constant<int, 256>{});  return true;
    case 512:  f(std::integral_constant<int, 512>{});  return true;
    case 1024: f(std::integral_constant<int, 1024>{}); return true;
    case 2048: f(std::integral_constant<int, 2048>{}); return true;
    case 384:  f(std::integral_constant<int, 384>{});  return true;
    case 768:  f(std::integral_constant<int, 768>{});  return true;
    case 1536: f(std::integral_constant<int, 1536>{}); return true;
    case 640:  f(std::integral_constant<int, 640>{});  return true;
    case 1280: f(std::integral_constant<int, 1280>{}); return true;
    default: return false;
  }
}
template <class F>
bool dispatch_x(int h, F&& f) {
  switch (h) {
    case 32:   f(std::integral_constant<int, 32>{});   return true;
    case 64:   f(std::integral_constant<int, 64>{});   return true;
    case 128:  f(std::integral_constant<int, 128>{});  return true;
    case 256:  f(std::integral_constant<int, 256>{});  return true;
    case 512:  f(std::integral_constant<int, 512>{});  return true;
    case 1024: f(std::integral_constant<int, 1024>{}); return true;
    case 192:  f(std::integral_constant<int, 192>{});  return true;
    case 384:  f(std::integral_constant<int, 384>{});  return true;
    case 768:  f(std::integral_constant<int, 768>{});  return true;
    case 320:  f(std::integral_constant<int, 320>{});  return true;
    case 640:  f(std::integral_constant<int, 640>{});  return true;
    default: return false;
  }
}

// bytes per tile row of a c2c stage: fixed by the block width of its tile-blocked segments (plan.h, W);
// plain strided layouts take 128-byte rows whenever that tile fits on chip.  0: inconsistent widths
template <typename T>
static int row_bytes(const P3dStage& st) {
  int aw = 0;
  for (int side = 0; side < 2; side++) {
    const P3dSide& sd = side ? st.out : st.in;
    for (int g = 0; g < sd.nseg; g++) {
      if (sd.seg[g].aw <= 1) continue;
      if (aw && sd.seg[g].aw != aw) return 0;
      aw = sd.seg[g].aw;
    }
  }
  if (!aw) return (ccfg_exists(st.nfft, 128) && !csplit_only(st.nfft)) ? 128 : 64;
  return aw * 2 * (int)sizeof(T);
}

template <typename T>
static int tile_lines(const P3dStage& st, int variant = 0) {
  int tx = 1;
  if (is_x(st.kind)) dispatch_x(st.n / 2, [&](auto h) { tx = XCfg<T, decltype(h)::value>::TX; });
  else tx = row_bytes<T>(st) / (2 * (int)sizeof(T));
  return tx;
}

template <typename T>
bool fast_supported(const P3dStage& st) {
  if (st.scale != 1.0 && st.kind == P3D_R2C) return false;     // fused scaling: c2c / DCT stages and the X c2r stage
  if (st.in.nseg + 1 > P3D_MAXRUN || st.out.nseg + 1 > P3D_MAXRUN) return false;
  switch (st.kind) {
    case P3D_C2C_FWD: case P3D_C2C_BWD: case P3D_DCT1: case P3D_DST1: {
      if (!c2c_len_ok(st.nfft)) return false;
      if (st.kind == P3D_DST1 && (st.scale != 1.0 || st.nfft != 2 * (st.n + 1))) return false;      // no SCALED instantiation
      const int rb = row_bytes<T>(st);
      if ((rb != 64 && rb != 128) || !ccfg_exists(st.nfft, rb)) return false;
      const int tx = tile_lines<T>(st);
      for (int side = 0; side < 2; side++) {          // one line pitch per side (the kernels keep it in a register)
        const P3dSide& sd = side ? st.out : st.in;
        for (int g = 0; g < sd.nseg; g++) {
          const P3dSeg& sg = sd.seg[g];
          if (sg.sa != sd.seg[0].sa || (sg.aw > 1 && sg.aw != tx)) return false;
        }
      }
      return true;
    }
    case P3D_R2C: case P3D_C2R: {
      if (st.n % 2 || !x_half_ok(st.n / 2)) return false;
      const P3dSide& rs = st.kind == P3D_R2C ? st.in : st.out;
      if (rs.nseg != 1 || rs.cnt != rs.logical) return false;
      const P3dSeg& g = rs.seg[0];
      if (g.ps != 1 || g.start != 0 || g.len != st.n) return false;
      if ((g.sa & 1) || (st.nb > 1 && (g.sb & 1)) || (st.nc > 1 && (g.sc & 1)) || (g.off & 1)) return false;
      const P3dSide& cs = st.kind == P3D_R2C ? st.out : st.in;      // complex side: plain lines, one pitch
      for (int i = 0; i < cs.nseg; i++)
        if (cs.seg[i].sa != cs.seg[0].sa || cs.seg[i].aw > 1) return false;
      return true;
    }
    default: return false;
  }
}

template <class F>
bool dispatch_r32(int n, F&& f) {
  switch (n) {
    case 512:  f(std::integral_constant<int, 512>{});  return true;
    case 1024: f(std::integral_constant<int, 1024>{}); return true;
    default: return false;
  }
}

// Kernel variant of a c2c / DCT stage (bit 0: two-pass radix-32 schedule, bit 2: bulk-copy stores), decided from the stage's
// geometry and two switches read at p3dfft_setup (fast_reload_switches), never per launch:
//   P3DFFT_B200_R32  = 0 never / 1 wherever a two-pass schedule exists / unset: the measured rule below
//   P3DFFT_B200_BULK = 0 never / 1 wherever the output rows allow it   / unset: stages that store into a peer's memory
// Measured on B200 (profiles/r2_ab_1gpu_optin_variants.log, 1024^3): the two-pass schedule 1024 = 32 x 32 wins where the
// stage WRITES whole tiles contiguously (Y forward 3.54 -> 3.32 ms, Y backward 3.67 -> 3.56 ms; single 1.87 -> 1.70 ms) and
// loses where the rows are scattered into the user layout or gathered from it (Z stages), and at 512 points.
#ifndef P3D_DEFAULT_ASYNC
#define P3D_DEFAULT_ASYNC 0
#endif
struct FastSwitches { int r32, bulk, async, xstage; bool loaded; };
static FastSwitches g_switches = {-1, -1, -1, -1, false};
void fast_reload_switches() {
  g_switches.r32 = getenv("P3DFFT_B200_R32") ? atoi(getenv("P3DFFT_B200_R32")) : -1;
  g_switches.bulk = getenv("P3DFFT_B200_BULK") ? atoi(getenv("P3DFFT_B200_BULK")) : -1;
  g_switches.async = getenv("P3DFFT_B200_ASYNC") ? atoi(getenv("P3DFFT_B200_ASYNC")) : -1;
  g_switches.xstage = getenv("P3DFFT_B200_XSTAGE") ? atoi(getenv("P3DFFT_B200_XSTAGE")) : -1;
  g_switches.loaded = true;
}
static const FastSwitches& fast_switches() {
  if (!g_switches.loaded) fast_reload_switches();
  return g_switches;
}

template <typename T>
int fast_variant(const P3dStage& st) {
  const FastSwitches& sw = fast_switches();
  if (st.kind == P3D_R2C) {      // bit 4: whole-row stores through a staging buffer (P3DFFT_B200_XSTAGE = 0 / 1 / unset: peers)
    bool peer = false;
    for (int g = 0; g < st.out.nseg; g++) if (st.out.seg[g].peer >= 0) peer = true;
    return (sw.xstage > 0 || (sw.xstage < 0 && peer)) ? 16 : 0;
  }
  if (is_x(st.kind) || st.kind == P3D_DST1 || st.out.nseg == 0 || st.in.nseg == 0) return 0;
  if (!(st.nfft == 1024 || st.nfft == 512) || row_bytes<T>(st) != 128) return 0;
  const int tx = 128 / (2 * (int)sizeof(T));
  // every output run: whole 128-byte tile rows, consecutive in memory (the writer-contiguous internal layouts)
  bool out_contig = true, out_peer = false;
  for (int g = 0; g < st.out.nseg; g++) {
    const P3dSeg& sg = st.out.seg[g];
    if (sg.sa != 1 || sg.kw > 1 || sg.ps != tx || sg.aw != tx) out_contig = false;
    if (sg.peer >= 0) out_peer = true;
  }
  const bool far_in = st.in.seg[0].ps * 2 * (long long)sizeof(T) > 131072;      // the split kernel's case (launch_fast)
  // asynchronously staged 1024-point kernel (bit 3; P3DFFT_B200_ASYNC = 0 never / 1 always / unset: P3D_DEFAULT_ASYNC)
  if (acfg_exists(st.nfft) && (sw.async > 0 || (sw.async < 0 && P3D_DEFAULT_ASYNC))) return 8;
  const bool r32 = ccfg_r32_exists(st.nfft) && (sw.r32 > 0 || (sw.r32 < 0 && st.nfft == 1024 && out_contig && !far_in));
  const bool bulk = out_contig && (sw.bulk > 0 || (sw.bulk < 0 && out_peer));
  return (r32 ? 1 : 0) | (bulk ? 4 : 0);
}

// twiddle block of the asynchronously staged kernel: pass table of the M-point schedule, then the outer twiddles
template <typename T>
static void fill_async_block(void* host) {
  using C = ACfg<T>;
  std::vector<long double> re, im;
  fill_pass_tables<typename C::S>(C::M, re, im);
  const long double twopi = 6.283185307179586476925286766559L;
  for (int q = 1; q < C::Q; q++)
    for (int k = 0; k < C::M; k++) {
      long double ang = -twopi * (long double)((long long)q * k % C::N) / (long double)C::N;
      re.push_back(cosl(ang)); im.push_back(sinl(ang));
    }
  T* o = reinterpret_cast<T*>(host);
  for (size_t i = 0; i < re.size(); i++) { o[2 * i] = (T)re[i]; o[2 * i + 1] = (T)im[i]; }
}

template <typename T>
size_t fast_twiddle_elems(int kind, int nfft, int variant) {
  size_t n = 0;
  if ((variant & 8) && !is_x(kind) && acfg_exists(nfft)) return (size_t)ACfg<T>::S::twtotal() + (ACfg<T>::Q - 1) * ACfg<T>::M;
  if ((variant & 1) && !is_x(kind)) { dispatch_r32(nfft, [&](auto nn) { n = block_elems<typename CCfgR32<T, decltype(nn)::value>::S>(false); }); return n; }
  if (is_x(kind)) dispatch_x(nfft / 2, [&](auto h) { n = block_elems<typename XCfg<T, decltype(h)::value>::S>(true); });
  else dispatch_c(nfft, [&](auto nn) { n = block_elems<typename CCfg<T, decltype(nn)::value, 64>::S>(false); });
  return n;
}

template <typename T>
void fast_twiddle_fill(int kind, int nfft, void* host, int variant) {
  if ((variant & 8) && !is_x(kind) && acfg_exists(nfft)) { fill_async_block<T>(host); return; }
  if ((variant & 1) && !is_x(kind)) { dispatch_r32(nfft, [&](auto nn) { fill_block<T, typename CCfgR32<T, decltype(nn)::value>::S>(false, host); }); return; }
  if (is_x(kind)) dispatch_x(nfft / 2, [&](auto h) { fill_block<T, typename XCfg<T, decltype(h)::value>::S>(true, host); });
  else dispatch_c(nfft, [&](auto nn) { fill_block<T, typename CCfg<T, decltype(nn)::value, 64>::S>(false, host); });
}

// stored index s -> logical k:  k = s (s < h1),  k = s + (logical - cnt) (s >= h1); a segment that
// straddles h1 becomes two runs.
static void side_to_runs(const P3dSide& sd, FastSide& f, size_t esz, int tx) {
  f.nrun = 0;
  const int shift = sd.logical - sd.cnt;
  for (int g = 0; g < sd.nseg; g++) {
    const P3dSeg& sg = sd.seg[g];
    const int s0 = sg.start, s1 = sg.start + sg.len;
    const int cut = shift > 0 ? sd.h1 : s1;
    const int parts[2][2] = {{s0, s1 < cut ? s1 : cut}, {s0 > cut ? s0 : cut, s1}};
    for (int p = 0; p < 2; p++) {
      const int a = parts[p][0], b = parts[p][1];
      if (b <= a) continue;
      FastRun& r = f.run[f.nrun++];
      const int sh = (shift > 0 && a >= sd.h1) ? shift : 0;
      r.base = sg.base;
      r.kstart = a + sh;
      r.korg = s0 + sh;
      r.len = b - a;
      r.ps = sg.ps; r.sa = sg.sa; r.sb = sg.sb; r.sc = sg.sc;
      r.kw = sg.kw; r.psh = sg.psh; r.bw = sg.bw; r.sbh = sg.sbh;
      r.sat = sg.aw > 1 ? sg.sah : sg.sa * tx;         // fast_supported() guarantees aw == tx when blocked
    }
  }
}

void to_fast(const P3dStage& st, FastStage& f, size_t real_bytes, int variant) {
  memset(&f, 0, sizeof f);
  const int tx = real_bytes == 4 ? tile_lines<float>(st, variant) : tile_lines<double>(st, variant);
  f.variant = variant;
  f.na = st.na; f.nb = st.nb; f.nc = st.nc; f.n = st.n;
  f.mirror = st.kind == P3D_DCT1 ? 1 : st.kind == P3D_DST1 ? 2 : 0;
  static const int pf = getenv("P3DFFT_B200_PREFETCH") ? atoi(getenv("P3DFFT_B200_PREFETCH")) : 131072;
  f.prefetch = pf;
  static const int bo = getenv("P3DFFT_B200_BORD") ? atoi(getenv("P3DFFT_B200_BORD")) : -1;
  f.bord = 1;
  if (!is_x(st.kind) && st.bord > 1) f.bord = bo >= 0 ? (bo > 1 ? bo : 1) : st.bord;
  f.rowb = is_x(st.kind) ? 0 : (real_bytes == 4 ? row_bytes<float>(st) : row_bytes<double>(st));
  f.tw = nullptr;
  f.scale = st.scale;
  side_to_runs(st.in, f.in, st.kind == P3D_R2C ? real_bytes : 2 * real_bytes, tx);
  side_to_runs(st.out, f.out, st.kind == P3D_C2R ? real_bytes : 2 * real_bytes, tx);
}

template <typename K>
static cudaError_t launch_cfg(K kernel, size_t smem, bool& configured) {
  if (configured) return cudaSuccess;
  cudaError_t e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e == cudaSuccess) configured = true;
  return e;
}

static int sm_count() {
  static int n = 0;
  if (!n) {
    int dev = 0;
    cudaGetDevice(&dev);
    if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0) n = 148;
  }
  return n;
}

// persistent grid: as many CTAs as stay resident (occupancy query, cached), never more than tiles
template <typename K>
static unsigned persistent_grid(K kernel, int nt, size_t smem, long long tiles, int& per_sm, int sm_cap = 0) {
  if (per_sm <= 0) {
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kernel, nt, smem) != cudaSuccess || per_sm <= 0) per_sm = 1;
  }
  const int sms = sm_cap > 0 && sm_cap < sm_count() ? sm_cap : sm_count();
  const long long g = (long long)per_sm * sms;
  return (unsigned)(tiles < g ? tiles : g);
}

// one (configured, CTAs per SM) pair per kernel instantiation; the enclosing function defines smem, tiles, NT, f, stream, e
#ifndef P3D_LAUNCH      // (tests/emu/emu_fast.cpp supplies a host emulation of the launch)
#define P3D_LAUNCH(...)                                                                                       \
  do {                                                                                                        \
    static bool cfg = false;                                                                                  \
    static int per_sm = 0;                                                                                    \
    if ((e = launch_cfg(__VA_ARGS__, smem, cfg)) != cudaSuccess) return e;                                    \
    __VA_ARGS__<<<persistent_grid(__VA_ARGS__, NT, smem, tiles, per_sm, f.sm_cap), NT, smem, stream>>>(f);              \
  } while (0)
#endif

template <typename T, int HH>
static cudaError_t launch_x_staged(const P3dStage& st, const FastStage& f, cudaStream_t stream) {
  constexpr int TX = XCfg<T, HH>::TX, NT = XCfg<T, HH>::NT;
  constexpr size_t smem = xstage_smem<T, HH>(true);
  const long long tiles = (long long)((st.na + TX - 1) / TX) * st.nb * st.nc;
  if (tiles <= 0) return cudaSuccess;
  if (tiles >= (1LL << 31)) return cudaErrorMisalignedAddress;
  cudaError_t e;
  P3D_LAUNCH(xr2c_kernel<T, HH, true>);
  return cudaGetLastError();
}

template <typename T, int HH>
static cudaError_t launch_x(const P3dStage& st, const FastStage& f, cudaStream_t stream) {
  constexpr int TX = XCfg<T, HH>::TX, NT = XCfg<T, HH>::NT;
  constexpr size_t smem = xstage_smem<T, HH>();
  const long long tiles = (long long)((st.na + TX - 1) / TX) * st.nb * st.nc;
  if (tiles <= 0) return cudaSuccess;
  if (tiles >= (1LL << 31)) return cudaErrorMisalignedAddress;      // 32-bit tile counters: the generic kernel takes over
  if (st.kind == P3D_R2C && (f.variant & 16)) return launch_x_staged<T, HH>(st, f, stream);
  cudaError_t e;
  if (st.kind == P3D_R2C) P3D_LAUNCH(xr2c_kernel<T, HH>);
  else if (f.scale != 1.0) P3D_LAUNCH(xc2r_kernel<T, HH, true>);
  else P3D_LAUNCH(xc2r_kernel<T, HH>);
  return cudaGetLastError();
}

template <typename T, int NN, int RB>
static cudaError_t launch_c(const P3dStage& st, const FastStage& f, cudaStream_t stream) {
  constexpr int TX = CCfg<T, NN, RB>::TX, NT = CCfg<T, NN, RB>::NT;
  constexpr size_t smem = cstage_smem<T, NN, RB>();
  const long long nbp = f.bord > 1 ? (long long)((st.nb + f.bord - 1) / f.bord) * f.bord : st.nb;
  const long long tiles = (long long)((st.na + TX - 1) / TX) * nbp * st.nc;
  if (tiles <= 0) return cudaSuccess;
  if (tiles >= (1LL << 31)) return cudaErrorMisalignedAddress;      // 32-bit tile counters: the generic kernel takes over
  cudaError_t e;
  const bool scaled = f.scale != 1.0;
  if (st.kind == P3D_DST1) {
    P3D_LAUNCH(cstage_kernel<T, NN, RB, false, false, CCfg<T, NN, RB>, false, true>);
  } else if (st.kind == P3D_C2C_BWD) {
    if (scaled) P3D_LAUNCH(cstage_kernel<T, NN, RB, true, true>);
    else P3D_LAUNCH(cstage_kernel<T, NN, RB, true>);
  } else {
    if (scaled) P3D_LAUNCH(cstage_kernel<T, NN, RB, false, true>);
    else P3D_LAUNCH(cstage_kernel<T, NN, RB, false>);
  }
  return cudaGetLastError();
}

// variants of the 128-byte-row c2c kernel: C = CCfg (three passes) or CCfgR32 (two passes); BULK = tile stored by cp.async.bulk
template <typename T, int NN, class C, bool BULK>
static cudaError_t launch_cv(const P3dStage& st, const FastStage& f, cudaStream_t stream) {
  using T2 = typename Cx<T>::type;
  constexpr int TX = C::TX, NT = C::NT;
  constexpr size_t smem = sizeof(T2) * NN * TX + 2 * sizeof(long long) * NN + sizeof(RunTab);
  const long long nbp = f.bord > 1 ? (long long)((st.nb + f.bord - 1) / f.bord) * f.bord : st.nb;
  const long long tiles = (long long)((st.na + TX - 1) / TX) * nbp * st.nc;
  if (tiles <= 0) return cudaSuccess;
  if (tiles >= (1LL << 31)) return cudaErrorMisalignedAddress;
  cudaError_t e;
  const bool scaled = f.scale != 1.0;
  if (st.kind == P3D_C2C_BWD) {
    if (scaled) P3D_LAUNCH(cstage_kernel<T, NN, 128, true, true, C, BULK>);
    else P3D_LAUNCH(cstage_kernel<T, NN, 128, true, false, C, BULK>);
  } else {
    if (scaled) P3D_LAUNCH(cstage_kernel<T, NN, 128, false, true, C, BULK>);
    else P3D_LAUNCH(cstage_kernel<T, NN, 128, false, false, C, BULK>);
  }
  return cudaGetLastError();
}

// asynchronously staged variant of the 1024-point stages (cp.async input staging, one CTA per SM)
template <typename T>
static cudaError_t launch_async(const P3dStage& st, const FastStage& f, cudaStream_t stream) {
  constexpr int TX = ACfg<T>::TX, NT = ACfg<T>::NT;
  constexpr size_t smem = cstage_async_smem<T>();
  const long long nbp = f.bord > 1 ? (long long)((st.nb + f.bord - 1) / f.bord) * f.bord : st.nb;
  const long long tiles = (long long)((st.na + TX - 1) / TX) * nbp * st.nc;
  if (tiles <= 0) return cudaSuccess;
  if (tiles >= (1LL << 31)) return cudaErrorMisalignedAddress;
  cudaError_t e;
  const bool scaled = f.scale != 1.0;
  if (st.kind == P3D_C2C_BWD) {
    if (scaled) P3D_LAUNCH(cstage_async_kernel<T, true, true>);
    else P3D_LAUNCH(cstage_async_kernel<T, true>);
  } else {
    if (scaled) P3D_LAUNCH(cstage_async_kernel<T, false, true>);
    else P3D_LAUNCH(cstage_async_kernel<T, false>);
  }
  return cudaGetLastError();
}

// split variant (two half tiles per CTA, two CTAs per SM) for the lengths whose 128-byte tile fills an SM
template <typename T, int NN>
static cudaError_t launch_split(const P3dStage& st, const FastStage& f, cudaStream_t stream) {
  constexpr int TX = CCfg<T, NN, 128>::TX, NT = SplitCfg<T, NN>::NT;
  constexpr size_t smem = cstage_split_smem<T, NN>();
  const long long nbp = f.bord > 1 ? (long long)((st.nb + f.bord - 1) / f.bord) * f.bord : st.nb;
  const long long tiles = (long long)((st.na + TX - 1) / TX) * nbp * st.nc;
  if (tiles <= 0) return cudaSuccess;
  if (tiles >= (1LL << 31)) return cudaErrorMisalignedAddress;
  cudaError_t e;
  const bool scaled = f.scale != 1.0;
  if (st.kind == P3D_DST1) {
    if constexpr (csplit_only(NN)) P3D_LAUNCH(cstage_split_kernel<T, NN, false, false, true>);
    else return cudaErrorInvalidValue;
  } else if (st.kind == P3D_C2C_BWD) {
    if (scaled) P3D_LAUNCH(cstage_split_kernel<T, NN, true, true>);
    else P3D_LAUNCH(cstage_split_kernel<T, NN, true>);
  } else {
    if (scaled) P3D_LAUNCH(cstage_split_kernel<T, NN, false, true>);
    else P3D_LAUNCH(cstage_split_kernel<T, NN, false>);
  }
  return cudaGetLastError();
}

template <typename T>
cudaError_t launch_fast(const P3dStage& st, const FastStage& f, cudaStream_t stream) {
  using T2 = typename Cx<T>::type;
  // alignment of every block base: the kernels move whole complex elements
  for (int side = 0; side < 2; side++) {
    const FastSide& sd = side ? f.out : f.in;
    for (int g = 0; g < sd.nrun; g++)
      if (reinterpret_cast<uintptr_t>(sd.run[g].base) % sizeof(T2)) return cudaErrorMisalignedAddress;
  }
  cudaError_t err = cudaErrorInvalidValue;
  if ((f.variant & 8) && !is_x(st.kind) && f.rowb == 128 && acfg_exists(st.nfft)) return launch_async<T>(st, f, stream);
  if (f.variant != 0 && !is_x(st.kind) && f.rowb == 128) {
    const bool r32 = (f.variant & 1) != 0, bulk = (f.variant & 4) != 0;
    if (st.nfft == 1024) {
      if (r32 && bulk) return launch_cv<T, 1024, CCfgR32<T, 1024>, true>(st, f, stream);
      if (r32) return launch_cv<T, 1024, CCfgR32<T, 1024>, false>(st, f, stream);
      if (bulk) return launch_cv<T, 1024, CCfg<T, 1024, 128>, true>(st, f, stream);
    } else if (st.nfft == 512) {
      if (r32 && bulk) return launch_cv<T, 512, CCfgR32<T, 512>, true>(st, f, stream);
      if (r32) return launch_cv<T, 512, CCfgR32<T, 512>, false>(st, f, stream);
      if (bulk) return launch_cv<T, 512, CCfg<T, 512, 128>, true>(st, f, stream);
    }
  }
  if (is_x(st.kind)) dispatch_x(st.n / 2, [&](auto h) { err = launch_x<T, decltype(h)::value>(st, f, stream); });
  else dispatch_c(st.nfft, [&](auto nn) {
    constexpr int NN = decltype(nn)::value;
    // Split variant (two CTAs per SM, more loads in flight): measured on B200 it wins when the INPUT rows are far
    // apart (the user layout: pitch of megabytes, no L2 prefetch) and loses for the other patterns, scattered
    // writes above all.  P3DFFT_B200_SPLIT = 0 never / 1 always / unset: by that rule.
    static const int split_env = getenv("P3DFFT_B200_SPLIT") ? atoi(getenv("P3DFFT_B200_SPLIT")) : -1;
    if (f.rowb == 64) err = launch_c<T, NN, 64>(st, f, stream);
    else if constexpr (csplit_only(NN)) { if (f.rowb == 128) err = launch_split<T, NN>(st, f, stream); }
    else if constexpr (NN == 1024) {
      const bool far_rows = f.in.nrun > 0 && f.in.run[0].ps * (long long)sizeof(T2) > (long long)(f.prefetch > 0 ? f.prefetch : 131072);
      const bool split = st.kind != P3D_DST1 && (split_env < 0 ? far_rows : split_env != 0);
      if (f.rowb == 128) err = split ? launch_split<T, NN>(st, f, stream) : launch_c<T, NN, 128>(st, f, stream);
    } else if constexpr (ccfg_exists(NN, 128)) { if (f.rowb == 128) err = launch_c<T, NN, 128>(st, f, stream); }
  });
  return err;
}

#ifdef SINGLE_PREC
typedef float fast_real_t;
#else
typedef double fast_real_t;
#endif
template bool fast_supported<fast_real_t>(const P3dStage&);
template int fast_variant<fast_real_t>(const P3dStage&);
template size_t fast_twiddle_elems<fast_real_t>(int, int, int);
template void fast_twiddle_fill<fast_real_t>(int, int, void*, int);
template cudaError_t launch_fast<fast_real_t>(const P3dStage&, const FastStage&, cudaStream_t);

}  // namespace p3d
