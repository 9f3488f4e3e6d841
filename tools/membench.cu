// Access-pattern microbenchmark for the stage kernels' two global sides (no FFT arithmetic).
//
// A persistent CTA moves tiles of R rows x CB bytes: all of a thread's 16-byte loads are issued
// first (like pass 1 of the stage kernels), then all its stores.  Each side of a tile is
//     base + (t % ta) * sa + (t / ta) * sb + row * pitch
// which covers the contiguous internal tiles, the [xb][y][z][xi] buffer and the user layout
// (x fastest, 513 complex per row -> 16-byte misalignment of the 64-byte tile rows).
//
//   nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o membench tools/membench.cu && ./membench
#include <cuda_runtime.h>

#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

struct Side { long long sa, sb, sc, pitch; int ta, tb; };
struct Job { Side in, out; int rows, cb; long long ntiles; int fin, fout; };
// flat tile t of the blocked [xb][row][y][xi] buffer: the 8 points f = 8 t + j of the contiguous (x, y) plane (x fastest,
// 513 per y), each at its own place (x / 8, y, x % 8) -- a 128-byte tile row becomes two partial rows of neighbouring blocks
__device__ __forceinline__ long long flat_off(long long t, int col, long long xbs) {
  const long long f = t * 8 + col, y = f / 513, x = f - y * 513;
  return (x >> 3) * xbs + y * 128 + (x & 7) * 16;
}

template <int LD>
__device__ __forceinline__ double2 ld16(const char* p) {
  double2 v;
  if (LD == 0) asm volatile("ld.global.nc.L1::no_allocate.v2.f64 {%0,%1}, [%2];" : "=d"(v.x), "=d"(v.y) : "l"(p));
  else if (LD == 1) asm volatile("ld.global.v2.f64 {%0,%1}, [%2];" : "=d"(v.x), "=d"(v.y) : "l"(p));
  else if (LD == 2) asm volatile("ld.global.nc.L1::no_allocate.L2::128B.v2.f64 {%0,%1}, [%2];" : "=d"(v.x), "=d"(v.y) : "l"(p));
  else if (LD == 3) asm volatile("ld.global.cg.v2.f64 {%0,%1}, [%2];" : "=d"(v.x), "=d"(v.y) : "l"(p));
  else asm volatile("ld.global.nc.L1::no_allocate.L2::64B.v2.f64 {%0,%1}, [%2];" : "=d"(v.x), "=d"(v.y) : "l"(p));
  return v;
}

// store flavour (P3D_MEMBENCH_ST = 0 plain, 1 .cs evict-first, 2 .cg L2 only, 3 .wt write-through); a run-time uniform branch
__device__ int g_st = 0;
__device__ __forceinline__ void st16(char* p, double2 v, int st) {
  if (st == 1) asm volatile("st.global.cs.v2.f64 [%0], {%1,%2};" ::"l"(p), "d"(v.x), "d"(v.y) : "memory");
  else if (st == 2) asm volatile("st.global.cg.v2.f64 [%0], {%1,%2};" ::"l"(p), "d"(v.x), "d"(v.y) : "memory");
  else if (st == 3) asm volatile("st.global.wt.v2.f64 [%0], {%1,%2};" ::"l"(p), "d"(v.x), "d"(v.y) : "memory");
  else asm volatile("st.global.v2.f64 [%0], {%1,%2};" ::"l"(p), "d"(v.x), "d"(v.y) : "memory");
}

// PER = 16-byte elements per thread per tile
template <int LD, int PER>
__global__ void __launch_bounds__(256) copy_tiles(const char* __restrict__ src, char* __restrict__ dst, Job j) {
  extern __shared__ char pad[];
  const int epr = j.cb / 16;                       // elements per row
  const int stf = g_st;
  for (long long t = blockIdx.x; t < j.ntiles; t += gridDim.x) {
    const char* s = src + (t % j.in.ta) * j.in.sa + ((t / j.in.ta) % j.in.tb) * j.in.sb + (t / j.in.ta / j.in.tb) * j.in.sc;
    char* d = dst + (t % j.out.ta) * j.out.sa + ((t / j.out.ta) % j.out.tb) * j.out.sb + (t / j.out.ta / j.out.tb) * j.out.sc;
    double2 v[PER];
#pragma unroll
    for (int i = 0; i < PER; i++) {
      const int e = i * 256 + threadIdx.x;
      if (j.fin) v[i] = ld16<LD>(src + flat_off(t, e % epr, j.in.sa) + (long long)(e / epr) * j.in.pitch);
      else v[i] = ld16<LD>(s + (long long)(e / epr) * j.in.pitch + (e % epr) * 16);
    }
#pragma unroll
    for (int i = 0; i < PER; i++) {
      const int e = i * 256 + threadIdx.x;
      if (j.fout) st16(dst + flat_off(t, e % epr, j.out.sa) + (long long)(e / epr) * j.out.pitch, v[i], stf);
      else st16(d + (long long)(e / epr) * j.out.pitch + (e % epr) * 16, v[i], stf);
    }
  }
  if (j.ntiles < 0) pad[0] = 0;
}

static float run(int ld, const char* src, char* dst, const Job& j, int ctas_per_sm, int reps) {
  int per = j.rows * (j.cb / 16) / 256;
  size_t smem = ctas_per_sm >= 4 ? 48 * 1024 : ctas_per_sm == 3 ? 72 * 1024 : ctas_per_sm == 2 ? 100 * 1024 : 200 * 1024;
  int grid = 148 * ctas_per_sm;
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0); cudaEventCreate(&e1);
  float best = 1e9f;
  for (int r = 0; r < reps + 1; r++) {
    cudaEventRecord(e0);
#define LAUNCH(LDV, PERV)                                                                              \
    {                                                                                                  \
      cudaFuncSetAttribute(copy_tiles<LDV, PERV>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem); \
      copy_tiles<LDV, PERV><<<grid, 256, smem>>>(src, dst, j);                                         \
    }
#define BYLD(PERV)                                                              \
    switch (ld) { case 0: LAUNCH(0, PERV) break; case 1: LAUNCH(1, PERV) break; \
                  case 2: LAUNCH(2, PERV) break; case 3: LAUNCH(3, PERV) break; default: LAUNCH(4, PERV) break; }
    if (per == 16) BYLD(16) else if (per == 8) BYLD(8) else if (per == 32) BYLD(32) else { printf("bad per %d\n", per); return 0; }
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    if (r > 0 && ms < best) best = ms;
  }
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) printf("CUDA error %s\n", cudaGetErrorString(e));
  return best;
}

int main(int argc, char** argv) {
  const long long NX = 513, NY = 1024, NZ = 1024;
  const size_t bytes = (size_t)(NX + 3) * NY * NZ * 16 + (1 << 20);
  char *a, *b;
  cudaMalloc(&a, bytes); cudaMalloc(&b, bytes);
  cudaMemset(a, 1, bytes); cudaMemset(b, 0, bytes);
  const char* only = argc > 1 ? argv[1] : "";
  if (getenv("P3D_MEMBENCH_ST")) {
    const int stv = atoi(getenv("P3D_MEMBENCH_ST"));
    cudaMemcpyToSymbol(g_st, &stv, sizeof stv);
    printf("# store flavour %d\n", stv);
  }
  int gran = argc > 2 ? atoi(argv[2]) : 0;
  if (gran) {
    cudaError_t e = cudaDeviceSetLimit(cudaLimitMaxL2FetchGranularity, gran);
    size_t g = 0; cudaDeviceGetLimit(&g, cudaLimitMaxL2FetchGranularity);
    printf("# L2 fetch granularity set %d -> %zu (%s)\n", gran, g, cudaGetErrorString(e));
  } else {
    size_t g = 0; cudaDeviceGetLimit(&g, cudaLimitMaxL2FetchGranularity);
    printf("# L2 fetch granularity default %zu\n", g);
  }
  const long long plane = NX * NY * 16;            // bytes of one z-plane of the user array
  struct Case { std::string name; Job j; };
  std::vector<Case> cases;
  auto contig = [&](int rows, int cb) { Side s; s.ta = 1 << 30; s.tb = 1; s.sa = (long long)rows * cb; s.sb = 0; s.sc = 0; s.pitch = cb; return s; };
  // user layout, tile (xb, y): cb bytes of x at row y, rows z
  auto user = [&](int cb, bool mis, int rows = 1024) {
    Side s; const int tpr = (int)(512 * 16 / cb);
    s.ta = tpr; s.sa = cb; s.tb = (int)NY; s.sb = mis ? NX * 16 : 512 * 16; s.sc = rows * plane; s.pitch = plane; return s;
  };
  // [xb][y][z][xi] buffer, tile (z, xb) (a = xb fastest as in the Y stage), rows y
  auto ybuf = [&](int cb, int rows = 1024) { Side s; s.ta = (int)(512 * 16 / cb); s.sa = NY * NZ * cb; s.tb = (int)NZ; s.sb = cb; s.sc = (long long)rows * NZ * cb; s.pitch = NZ * cb; return s; };
  // same buffer but tile order z fastest: tiles (z, z+1, ...) of one xb are neighbours in time
  auto ybuf_zfast = [&](int cb) { Side s; s.ta = (int)NZ; s.sa = cb; s.tb = 1 << 30; s.sb = NY * NZ * cb; s.sc = 0; s.pitch = NZ * cb; return s; };
  auto add = [&](const std::string& n, Side in, Side out, int rows, int cb) {
    Job j; j.in = in; j.out = out; j.rows = rows; j.cb = cb; j.fin = j.fout = 0; j.ntiles = (long long)(512 * 16 / cb) * NY * (NZ / rows) * 1;
    j.ntiles = (512LL * 16 * NY * NZ) / ((long long)rows * cb);
    cases.push_back({n, j});
  };
  add("contig->contig 64", contig(1024, 64), contig(1024, 64), 1024, 64);
  add("user64mis->contig (z_bwd)", user(64, true), contig(1024, 64), 1024, 64);
  add("user64aligned->contig", user(64, false), contig(1024, 64), 1024, 64);
  add("contig->user64mis (z_fwd)", contig(1024, 64), user(64, true), 1024, 64);
  add("contig->user64aligned", contig(1024, 64), user(64, false), 1024, 64);
  add("ybuf64->contig (y_bwd)", ybuf(64), contig(1024, 64), 1024, 64);
  add("ybuf64 zfast->contig", ybuf_zfast(64), contig(1024, 64), 1024, 64);
  add("contig->ybuf64 (y_fwd)", contig(1024, 64), ybuf(64), 1024, 64);
  add("contig->ybuf64 zfast", contig(1024, 64), ybuf_zfast(64), 1024, 64);
  add("user128mis->contig", user(128, true, 512), contig(512, 128), 512, 128);
  add("contig->user128mis", contig(512, 128), user(128, true, 512), 512, 128);
  add("user256mis->contig", user(256, true, 256), contig(256, 256), 256, 256);
  add("contig->user256mis", contig(256, 256), user(256, true, 256), 256, 256);
  // 128-byte rows of the user layout: rows aligned (as if nxhp were 512), and FLAT tiles -- 8 consecutive points of the
  // contiguous (x, y) plane, whatever row they fall into: aligned rows although nxhp = 513 is odd
  auto userflat = [&](int cb, int rows) {
    Side s; s.ta = (int)(NX * NY * 16 / cb); s.sa = cb; s.tb = 1; s.sb = 0; s.sc = rows * plane; s.pitch = plane; return s;
  };
  add("contig->contig 128", contig(512, 128), contig(512, 128), 512, 128);
  add("user128aligned->contig", user(128, false, 512), contig(512, 128), 512, 128);
  add("contig->user128aligned", contig(512, 128), user(128, false, 512), 512, 128);
  add("userflat128->contig", userflat(128, 512), contig(512, 128), 512, 128);
  add("contig->userflat128", contig(512, 128), userflat(128, 512), 512, 128);
  add("ybuf128->user128mis (z_fwd now)", ybuf(128, 512), user(128, true, 512), 512, 128);
  add("ybuf128->userflat128 (z_fwd flat)", ybuf(128, 512), userflat(128, 512), 512, 128);
  // the flat tiles on the blocked side: rows of 512 (the upper half of the rows is not visited -- same traffic per tile)
  {
    Side fb; fb.ta = 1 << 30; fb.tb = 1; fb.sa = NY * NZ * 128; fb.sb = 0; fb.sc = 0; fb.pitch = NY * 128;
    Side uf = userflat(128, 512); uf.sc = 0;      // tile t <-> flat tile t of the first 512 planes
    add("ybufflat128->userflat128 (z_fwd flat both)", fb, uf, 512, 128); cases.back().j.fin = 1; cases.back().j.ntiles = NX * NY / 8;
    add("userflat128->ybufflat128 (z_bwd flat both)", uf, fb, 512, 128); cases.back().j.fout = 1; cases.back().j.ntiles = NX * NY / 8;
    add("ybufflat128->contig", fb, contig(512, 128), 512, 128); cases.back().j.fin = 1; cases.back().j.ntiles = NX * NY / 8;
    add("contig->ybufflat128", contig(512, 128), fb, 512, 128); cases.back().j.fout = 1; cases.back().j.ntiles = NX * NY / 8;
    Side um = user(128, true, 512); um.sc = 0;
    add("ybuf128->user128mis half (z_fwd now)", ybuf(128, 512), um, 512, 128); cases.back().j.ntiles = 64 * NY;
    add("user128mis->contig half (z_bwd now)", um, contig(512, 128), 512, 128); cases.back().j.ntiles = 64 * NY;
  }
  add("ybuf128->contig", ybuf(128, 512), contig(512, 128), 512, 128);
  add("contig->ybuf128", contig(512, 128), ybuf(128, 512), 512, 128);
  add("ybuf256->contig", ybuf(256, 256), contig(256, 256), 256, 256);
  add("contig->ybuf256", contig(256, 256), ybuf(256, 256), 256, 256);
  const double gb = 2.0 * 512 * 16 * NY * NZ / 1e9;
  printf("%-30s %4s %4s %8s %8s\n", "pattern", "ld", "cta", "ms", "GB/s");
  for (auto& c : cases) {
    if (*only) {      // comma-separated substrings
      bool hit = false;
      std::string o(only);
      for (size_t p0 = 0; p0 <= o.size();) {
        size_t p1 = o.find(',', p0); if (p1 == std::string::npos) p1 = o.size();
        if (p1 > p0 && c.name.find(o.substr(p0, p1 - p0)) != std::string::npos) hit = true;
        p0 = p1 + 1;
      }
      if (!hit) continue;
    }
    for (int ld : {0, 1, 2, 4}) {
      if (ld != 0 && c.name.find("->contig") == std::string::npos) continue;     // load flavour matters on scattered reads
      for (int cps : {2, 4}) {
        float ms = run(ld, a, b, c.j, cps, 2);
        const double gbc = 2.0 * (double)c.j.ntiles * c.j.rows * c.j.cb / 1e9;
        (void)gb;
        printf("%-44s %4d %4d %8.3f %8.1f\n", c.name.c_str(), ld, cps, ms, gbc / ms * 1e3);
        fflush(stdout);
      }
    }
  }
  return 0;
}
