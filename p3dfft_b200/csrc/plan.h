// Host-side planner: decomposition arithmetic and the stage/exchange sequence of one
// forward or backward transform.  Pure integer code, no CUDA -- it is exercised on CPU by
// tests/ through the p3dfft_b200_plan_* entry points.
//
// Restates build/setup.F90:148-176 (derived extents), :192-219 (rank -> (ipid,jpid)),
// :279-312 (block maps), :382-398 (padi,nm), :481-518 (alltoallv tables), :580-603
// (memsize), :608-635 (MapDataToProc); build/module.F90:225-273 (get_dims); stage order of
// build/ftran.F90:489-780 and build/btran.F90:396-679.
#pragma once
#include <algorithm>
#include <cstdio>
#include <cstring>
#include <string>
#include <vector>

#include "stage.h"

namespace p3d {

struct BlockMap {
  std::vector<int> st, en, sz;   // 1-based starts/ends like the reference
};

// MapDataToProc (setup.F90:608-635): the LAST (data mod proc) ranks get one extra element.
inline BlockMap map_data_to_proc(int data, int proc) {
  BlockMap m;
  m.st.assign(proc, 0); m.en.assign(proc, 0); m.sz.assign(proc, 0);
  int size = data / proc, nu = data - size * proc, nl = proc - nu;
  m.st[0] = 1; m.sz[0] = size; m.en[0] = size;
  for (int i = 1; i < nl; i++) { m.st[i] = m.st[i-1] + size; m.sz[i] = size; m.en[i] = m.en[i-1] + size; }
  for (int i = std::max(nl, 1); i < proc; i++) {
    m.st[i] = m.en[i-1] + 1; m.sz[i] = size + 1; m.en[i] = m.en[i-1] + size + 1;
  }
  m.en[proc-1] = data;
  m.sz[proc-1] = data - m.st[proc-1] + 1;
  return m;
}

struct Decomp {
  int nx = 0, ny = 0, nz = 0, nxc = 0, nyc = 0, nzc = 0;
  int nxh, nxhp, nxhc, nxhpc, nyh, nzh, nyhc, nzhc, nycph, nzcph;
  int iproc = 1, jproc = 1, ipid = 0, jpid = 0, rank = 0, numtasks = 1;
  bool dims_c = false, stride1 = false;
  BlockMap ii, ji, jj, kj;   // x(nxhpc)/iproc, y(ny)/iproc, y(nyc)/jproc, z(nz)/jproc
  BlockMap iii, ij;          // real-space x(nx)/iproc and x(nx)/jproc of the rtran_* transposes (setup.F90:299-312)
  int iistart, iiend, iisize, jistart, jiend, jisize, jjstart, jjend, jjsize, kjstart, kjend, kjsize;
  int iiistart, iiiend, iiisize, ijstart, ijend, ijsize;
  int padi_work = 0, padi = 0;
  long long nm = 0;
  int memsize[3] = {0, 0, 0};

  // returns empty string on success, else the reference's error text
  std::string init(int nx_, int ny_, int nz_, int d0, int d1, int rank_, int ntasks,
                   int nxc_, int nyc_, int nzc_, bool dims_c_, bool stride1_) {
    char msg[256];
    if (nx_ <= 0 || ny_ <= 0 || nz_ <= 0) {                       // setup.F90:125-128
      snprintf(msg, sizeof msg, "Invalid dimensions : %d %d %d", nx_, ny_, nz_);
      return msg;
    }
    if (d0 <= 0 || d1 <= 0 || d0 * d1 != ntasks) {                // setup.F90:181-184
      snprintf(msg, sizeof msg, "Invalid processor geometry: %d %d for %d tasks", d0, d1, ntasks);
      return msg;
    }
    nx = nx_; ny = ny_; nz = nz_; nxc = nxc_; nyc = nyc_; nzc = nzc_;
    if (nxc <= 0 || nxc > nx || nyc <= 0 || nyc > ny || nzc <= 0 || nzc > nz) {
      snprintf(msg, sizeof msg, "Invalid pruned dimensions : %d %d %d", nxc, nyc, nzc);
      return msg;
    }
    dims_c = dims_c_; stride1 = stride1_;
    iproc = d0; jproc = d1; rank = rank_; numtasks = ntasks;
    nxh = nx / 2; nxhp = nxh + 1; nxhc = nxc / 2; nxhpc = nxhc + 1;
    nyh = ny / 2; nzh = nz / 2; nyhc = nyc / 2; nzhc = nzc / 2;
    nycph = (nyc + 1) / 2; nzcph = (nzc + 1) / 2;
    if (dims_c) { ipid = rank / jproc; jpid = rank % jproc; }     // setup.F90:195-219
    else        { ipid = rank % iproc; jpid = rank / iproc; }
    ii = map_data_to_proc(nxhpc, iproc);
    ji = map_data_to_proc(ny, iproc);
    jj = map_data_to_proc(nyc, jproc);
    kj = map_data_to_proc(nz, jproc);
    iistart = ii.st[ipid]; iiend = ii.en[ipid]; iisize = ii.sz[ipid];
    jistart = ji.st[ipid]; jiend = ji.en[ipid]; jisize = ji.sz[ipid];
    jjstart = jj.st[jpid]; jjend = jj.en[jpid]; jjsize = jj.sz[jpid];
    kjstart = kj.st[jpid]; kjend = kj.en[jpid]; kjsize = kj.sz[jpid];
    iii = map_data_to_proc(nx, iproc);                             // setup.F90:305-312
    ij = map_data_to_proc(nx, jproc);
    iiistart = iii.st[ipid]; iiiend = iii.en[ipid]; iiisize = iii.sz[ipid];
    ijstart = ij.st[jpid]; ijend = ij.en[jpid]; ijsize = ij.sz[jpid];
    // setup.F90:382-398
    long long padd = std::max((long long)iisize * jjsize * nz, (long long)iisize * ny * kjsize)
                     - (long long)nxhp * jisize * kjsize;
    long long d = (long long)nxhp * jisize;
    if (padd <= 0 || d == 0) padi_work = 0;
    else padi_work = (int)(padd / d + (padd % d ? 1 : 0));
    nm = (long long)nxhp * jisize * (kjsize + padi_work);
    // setup.F90:580-603
    long long pad1 = 2 * std::max((long long)nz * jjsize * iisize, (long long)ny * kjsize * iisize)
                     - (long long)nx * jisize * kjsize;
    if (pad1 < 0) pad1 = 0;
    long long dm = (long long)nx * jisize;
    padi = dm > 0 ? (int)(pad1 / dm + (pad1 % dm ? 1 : 0)) : 0;
    memsize[0] = nx; memsize[1] = jisize; memsize[2] = kjsize + padi;
    return "";
  }

  int rank_of(int ip, int jp) const { return dims_c ? ip * jproc + jp : jp * iproc + ip; }

  // p3dfft_get_dims (module.F90:225-273)
  bool get_dims(int* st, int* en, int* sz, int conf) const {
    if (conf == 1) {
      st[0] = 1; en[0] = nx; sz[0] = nx;
      st[1] = jistart; en[1] = jiend; sz[1] = jisize;
      st[2] = kjstart; en[2] = kjend; sz[2] = kjsize;
    } else if (conf == 2) {
      if (stride1) {
        st[0] = 1; en[0] = nzc; sz[0] = nzc;
        st[1] = jjstart; en[1] = jjend; sz[1] = jjsize;
        st[2] = iistart; en[2] = iiend; sz[2] = iisize;
      } else {
        st[0] = iistart; en[0] = iiend; sz[0] = iisize;
        st[1] = jjstart; en[1] = jjend; sz[1] = jjsize;
        st[2] = 1; en[2] = nzc; sz[2] = nzc;
      }
    } else if (conf == 3) {
      for (int i = 0; i < 3; i++) { st[i] = 0; en[i] = memsize[i]; sz[i] = memsize[i]; }
    } else return false;
    return true;
  }

  // complex elements each work buffer must hold for nv variables.  W = lines per 64-byte tile
  // row of the blocked internal layouts (0: the reference's plain layouts)
  long long work_elems(int nv, int W = 0) const {
    long long m;
    if (W <= 0) {
      m = (long long)nxhpc * jisize * kjsize;
      m = std::max(m, (long long)iisize * ny * kjsize);
      m = std::max(m, (long long)iisize * nyc * kjsize);
      m = std::max(m, (long long)iisize * jjsize * nz);
    } else {
      long long nxb_all = 0;
      for (int p = 0; p < iproc; p++) nxb_all += (ii.sz[p] + W - 1) / W;
      const long long nxb = (iisize + W - 1) / W;
      m = nxb_all * W * jisize * kjsize;                         // X side of the X<->Y buffer
      m = std::max(m, nxb * W * (long long)ny * kjsize);         // Y side of it
      m = std::max(m, nxb * W * (long long)nyc * kjsize);        // Y side of the Y<->Z buffer
      m = std::max(m, nxb * W * (long long)jjsize * nz);         // Z side of it
    }
    return std::max<long long>(m, 1) * nv;
  }
};

// Block width W of the tile-blocked internal layouts (lines per tile row) for complex elements of `csize`
// bytes: 128-byte rows unless the 128-byte tile of the longest Y/Z transform would not fit in shared
// memory (fft_fast.cuh, CCfg); `force_row_bytes` (64 or 128) overrides the rule.
inline int pick_W(int ny, int nz, int csize, int force_row_bytes = 0) {
  // 128-byte rows whenever a kernel exists for the 128-byte tile of the longest Y/Z transform: up to 1024 points for any length
  // (the any-length kernel then takes up to 8 lines), the specialised 1280- and 1536-point kernels (160 / 192 KB tiles;
  // measured at 1280^3: 64-byte rows 76.6 ms per pair, 2.6 TB/s -- the Z stages crawl on half-line accesses) and 2048 points
  // through the split kernel (half of the 256 KB tile waits in registers; measured on one B200, profiles/
  // r2_ab_1gpu_2048_rows.log: 256 x 2048 x 2048 double 33.2 -> 28.5 ms per pair, 512 x 2048 x 2048 single 31.1 -> 29.2: the Z
  // stages gain 17-29 %, the single-precision Y stages, whose 64-byte-row tiles are contiguous, lose 4-11 %).
  // nz = 1023 / 1025: the sine / cosine (Chebyshev) transform of the third dimension is a 2048-point FFT of the odd / even
  // extension (split kernel); the plain c2c of such an nz is the any-length kernel's
  auto fits128 = [](int n) { return n <= 1024 || n == 1280 || n == 1536 || n == 2048; };
  const int rb = (force_row_bytes == 64 || force_row_bytes == 128) ? force_row_bytes
               : ((fits128(ny) && (fits128(nz) || nz == 1025)) ? 128 : 64);
  return rb / csize;
}

// ------------------------------------------------------------------------------------
// FFT length factorisation for the on-chip engine
// ------------------------------------------------------------------------------------
inline bool factorize(int n, int* fac, int* nfac, int maxprime = 4096) {
  int k = 0;
  while (n % 8 == 0) { fac[k++] = 8; n /= 8; }
  while (n % 4 == 0) { fac[k++] = 4; n /= 4; }
  while (n % 2 == 0) { fac[k++] = 2; n /= 2; }
  for (int p = 3; p <= maxprime && n > 1; p += 2)
    while (n % p == 0) { if (k >= P3D_MAXFAC) return false; fac[k++] = p; n /= p; }
  *nfac = k;
  return n == 1;
}

struct Step {
  bool is_exchange = false;
  int chunk = -1;        // >= 0: this step belongs to chunk `chunk` of a pipelined group (split_for_overlap)
  bool side = false;     // consumer stage of a pipelined group: runs on the side stream after its chunk's barrier
  P3dStage st{};
  P3dExchange ex{};
};

struct TransformPlan {
  std::vector<Step> steps;
  std::string error;
};

inline int kind_of_letter(char ch, bool backward) {
  switch (ch) {
    case 't': case 'f': return backward ? P3D_C2C_BWD : P3D_C2C_FWD;
    case 'c': return P3D_DCT1;
    case 's': return P3D_DST1;
    case 'n': case '0': return P3D_NOOP;
    default: return -1;      // "Unknown transform type" (ftran.F90:640-643)
  }
}

inline int fft_len(int kind, int n) {
  switch (kind) {
    case P3D_DCT1: return n > 1 ? 2 * (n - 1) : 1;
    case P3D_DST1: return 2 * (n + 1);
    default: return n;
  }
}

inline void stage_init(P3dStage& s, int kind, int n, int na, int nb, int nc, int timer) {
  memset(&s, 0, sizeof s);
  s.kind = kind; s.n = n; s.nfft = fft_len(kind, n);
  s.na = na; s.nb = nb; s.nc = nc; s.timer = timer; s.scale = 1.0;
}

inline void side_init(P3dSide& sd, int logical, int cnt, int h1) {
  sd.nseg = 0; sd.logical = logical; sd.cnt = cnt; sd.h1 = h1;
}

inline void add_seg(P3dSide& sd, int buf, int peer, long long off, int start, int len,
                    long long ps, long long sa, long long sb, long long sc) {
  if (len <= 0) return;
  P3dSeg& g = sd.seg[sd.nseg++];
  g.base = nullptr; g.buf = buf; g.peer = peer; g.off = off; g.start = start; g.len = len;
  g.ps = ps; g.sa = sa; g.sb = sb; g.sc = sc;
  g.kw = 0; g.aw = 0; g.psh = 0; g.sah = 0; g.bw = 0; g.pad_ = 0; g.sbh = 0;
}

// segment of a tile-blocked buffer: rows blocked by kw (stride psh) and/or lines blocked by aw (stride sah)
inline void add_seg_blk(P3dSide& sd, int buf, int peer, long long off, int start, int len, long long ps, int kw,
                        long long psh, long long sa, int aw, long long sah, long long sb, long long sc) {
  if (len <= 0) return;
  add_seg(sd, buf, peer, off, start, len, ps, sa, sb, sc);
  P3dSeg& g = sd.seg[sd.nseg - 1];
  g.kw = kw; g.psh = psh; g.aw = aw; g.sah = sah;
}

// Forward: X r2c -> T1(row) -> Y c2c -> T2(col) -> Z.   ftran.F90:489-780.
// Backward: Z -> T3(col) -> Y c2c -> T4(row) -> X c2r.   btran.F90:396-679.
// dim_real / dim_cplx: per-variable element strides of the user arrays (reals for the
// real-space array, complex elements for the wavenumber array), as in the *_many API.
//
// Work buffers rotate among A, B, C (the reference also owns three: buf, buf1, buf2,
// setup.F90:401-423): a stage never writes a buffer it reads, every non-self block of an
// exchange is written into the send buffer `snd`, and the rank's own block is written
// straight to its landing place in the receive buffer `rcv`.
//
// W > 0 selects the B200 layouts of the internal buffers: W lines that are adjacent in x (W * sizeof(complex)
// = 128 bytes, or 64 when a 128-byte tile of the longest transform would not fit in shared memory) form one
// row of a kernel tile --
//   X<->Y buffer   [z][x/W][y][x%W]   Y-stage tile = (z, x/W), rows y: contiguous; the X stage stores / loads
//                                     pieces of (lines per X tile) * W elements
//   Y->Z buffer    [x/W][z][y][x%W]   forward:  the Y stage WRITES its tile contiguously, the Z stage gathers rows
//   Z->Y buffer    [x/W][y][z][x%W]   backward: the Z stage WRITES its tile contiguously, the Y stage gathers rows
// per peer block (blocks padded in x to a multiple of W).  Measured on B200 (tools/membench.cu): scattered
// WRITES of one row are the slow side (DRAM write locality), scattered READS recover most of the bandwidth
// when the tiles that share memory lines run at the same time -- hence writer-contiguous layouts and the
// `bord` tile-order hint on the gathering stage.  The exchange still moves one
// contiguous block per peer; only the order of the elements inside a block differs from the
// reference's pack buffers, which no caller can observe.  W = 0 keeps the reference's plain
// layouts and its exact alltoallv tables (setup.F90:481-518).
//
// p2p (needs W > 0): a stage writes the block destined to peer p straight into p's RECEIVE buffer
// through a peer-mapped pointer (NVLink stores issued by the FFT kernel itself), at the offset where
// p expects the block from this rank; the exchange step degenerates to a barrier.  seg.peer then
// holds the WORLD rank owning the memory.  Every rank runs the same step sequence, so buffer ids
// rotate identically everywhere and a world barrier after each producing stage orders both the
// read-after-write and the write-after-read hazards (api.cpp, run_exchange).
inline TransformPlan build_plan(const Decomp& d, bool backward, const char* op, int nv,
                                long long dim_real, long long dim_cplx, int W = 0, bool p2p = false) {
  TransformPlan tp;
  const int M1 = d.iproc, M2 = d.jproc;
  if (M1 > P3D_MAXSEG || M2 > P3D_MAXSEG) { tp.error = "processor grid dimension exceeds P3D_MAXSEG"; return tp; }
  const long long ii = d.iisize, ji = d.jisize, jj = d.jjsize, kj = d.kjsize;
  const long long nx = d.nx, ny = d.ny, nz = d.nz, nxhpc = d.nxhpc, nyc = d.nyc, nzc = d.nzc;
  const int zkind = kind_of_letter(backward ? op[0] : op[2], backward);
  if (zkind < 0) { tp.error = std::string("Unknown transform type: ") + (backward ? op[0] : op[2]); return tp; }
  int cur = -1, snd = P3D_BUF_A, rcv = P3D_BUF_B;
  auto rotate = [&]() {   // choose snd/rcv different from the buffer holding the data (cur)
    int f[2], k = 0;
    for (int b = P3D_BUF_A; b <= P3D_BUF_C; b++) if (b != cur && k < 2) f[k++] = b;
    snd = f[0]; rcv = f[1];
  };

  auto push_stage = [&](P3dStage& s) {
    if ((long long)s.na * s.nb * s.nc <= 0) return;
    if (s.kind == P3D_NOOP) s.nfac = 0;
    else if (!factorize(s.nfft, s.fac, &s.nfac)) {
      char m[128]; snprintf(m, sizeof m, "transform length %d needs a prime factor > 4096 (unsupported)", s.nfft);
      tp.error = m; return;
    }
    s.need_zero = (s.in.cnt < s.in.logical) || s.kind == P3D_DST1;
    Step st; st.is_exchange = false; st.st = s; tp.steps.push_back(st);
  };
  auto push_exchange = [&](int comm, int npeer, int self, int timer, const BlockMap& sm, long long sunit,
                           const BlockMap& rm, long long runit) {
    Step st; st.is_exchange = true; P3dExchange& e = st.ex; memset(&e, 0, sizeof e);
    e.comm = comm; e.npeer = npeer; e.self = self; e.sendbuf = snd; e.recvbuf = rcv; e.timer = timer;
    for (int p = 0; p < npeer; p++) {
      e.sndoff[p] = nv * (long long)(sm.st[p] - 1) * sunit; e.sndcnt[p] = nv * (long long)sm.sz[p] * sunit;
      e.rcvoff[p] = nv * (long long)(rm.st[p] - 1) * runit; e.rcvcnt[p] = nv * (long long)rm.sz[p] * runit;
    }
    tp.steps.push_back(st);
  };
  // Piecewise side over the blocks of `bm` (one per peer of an exchange).  Block p holds
  // nv variables of extent unit*bm.sz[p]; `strides(p)` gives (ps,sa,sb) and sc = unit*sz.
  // send=true : stage OUTPUT feeding an exchange (self block redirected into rcv at selfoff)
  // send=false: stage INPUT reading what an exchange delivered into `cur`.
  auto piecewise = [&](P3dSide& sd, const BlockMap& bm, int npeer, int self, long long unit, bool send,
                       long long selfoff, long long ps_mul, int mode) {
    for (int p = 0; p < npeer; p++) {
      long long sz = bm.sz[p];
      long long ps, sa, sb;
      if (mode == 0)      { ps = 1;      sa = sz;  sb = sz * ji; }        // x-blocks   (sz, ji, kj)
      else if (mode == 1) { ps = ii;     sa = 1;   sb = ii * sz; }        // y-blocks   (ii, sz, kj)
      else                { ps = ii*jj;  sa = 1;   sb = d.stride1 ? ii : 0; }   // z-slabs (ii, jj, sz)
      (void)ps_mul;
      long long sc = unit * sz;
      long long off = nv * (long long)(bm.st[p] - 1) * unit;
      if (send && p == self) add_seg(sd, rcv, -1, selfoff, bm.st[p] - 1, (int)sz, ps, sa, sb, sc);
      else add_seg(sd, send ? snd : cur, -1, off, bm.st[p] - 1, (int)sz, ps, sa, sb, sc);
    }
  };

  P3dStage s;
  if (W > 0) {
    // ================= tile-blocked internal buffers =====================================
    auto cdiv = [&](long long a) { return (a + W - 1) / W; };
    const long long nxb = cdiv(ii);
    // exchange whose per-peer element counts are given explicitly
    auto push_exchange_cnt = [&](int comm, int npeer, int self, int timer, const std::vector<long long>& sc_,
                                 const std::vector<long long>& rc_) {
      Step st; st.is_exchange = true; P3dExchange& e = st.ex; memset(&e, 0, sizeof e);
      e.comm = comm; e.npeer = npeer; e.self = self; e.sendbuf = snd; e.recvbuf = rcv; e.timer = timer;
      e.p2p = p2p ? 1 : 0;
      long long so = 0, ro = 0;
      for (int p = 0; p < npeer; p++) {
        e.sndoff[p] = so; e.sndcnt[p] = nv * sc_[p]; so += e.sndcnt[p];
        e.rcvoff[p] = ro; e.rcvcnt[p] = nv * rc_[p]; ro += e.rcvcnt[p];
      }
      tp.steps.push_back(st);
    };
    // ---- per-peer block sizes (elements per variable) --------------------------------------
    std::vector<long long> xy_x(M1), xy_y(M1), yz_y(M2), yz_z(M2);
    for (int p = 0; p < M1; p++) { xy_x[p] = kj * cdiv(d.ii.sz[p]) * ji * W; xy_y[p] = kj * nxb * d.ji.sz[p] * W; }
    for (int p = 0; p < M2; p++) { yz_y[p] = nxb * d.jj.sz[p] * kj * W; yz_z[p] = nxb * jj * d.kj.sz[p] * W; }
    auto offs = [&](const std::vector<long long>& v, int p) { long long o = 0; for (int q = 0; q < p; q++) o += nv * v[q]; return o; };
    // X side of the X<->Y buffer: rows = x (blocked by W), lines = y, b = z.   [z][xb][y][xi] per peer block
    auto xy_xside = [&](P3dSide& sd, bool send) {
      for (int p = 0; p < M1; p++) {
        const long long nxbp = cdiv(d.ii.sz[p]);
        const bool self_redirect = send && M1 > 1 && p == d.ipid;
        const bool remote = send && !self_redirect && M1 > 1 && p2p;
        const int buf = send ? ((self_redirect || remote) ? rcv : snd) : cur;
        long long off = self_redirect ? offs(xy_y, d.ipid) : offs(xy_x, p);
        if (remote) {     // where peer p expects the block from this rank (its Y side of the buffer)
          off = 0;
          for (int q = 0; q < d.ipid; q++) off += nv * kj * nxbp * d.ji.sz[q] * W;
        }
        add_seg_blk(sd, buf, remote ? d.rank_of(p, d.jpid) : -1, off, d.ii.st[p] - 1, d.ii.sz[p],
                    1, W, ji * W, W, 0, 0, nxbp * ji * W, xy_x[p]);
      }
    };
    // Y side of the X<->Y buffer: rows = y, lines = x (blocked by W), b = z
    auto xy_yside = [&](P3dSide& sd, bool send) {
      for (int q = 0; q < M1; q++) {
        const long long nyq = d.ji.sz[q];
        const bool self_redirect = send && M1 > 1 && q == d.ipid;
        const bool remote = send && !self_redirect && M1 > 1 && p2p;
        const int buf = send ? ((self_redirect || remote) ? rcv : snd) : cur;
        long long off = self_redirect ? offs(xy_x, d.ipid) : offs(xy_y, q);
        if (remote) {     // peer q's X side: blocks ordered by sender, sender p' holds x in ii(p'), y in ji(q)
          off = 0;
          for (int pp = 0; pp < d.ipid; pp++) off += nv * kj * cdiv(d.ii.sz[pp]) * nyq * W;
        }
        add_seg_blk(sd, buf, remote ? d.rank_of(q, d.jpid) : -1, off, d.ji.st[q] - 1, (int)nyq,
                    W, 0, 0, 1, W, nyq * W, nxb * nyq * W, xy_y[q]);
      }
    };
    // Y<->Z buffer, writer-contiguous (see the header comment).  Block sizes do not depend on the direction.
    auto yz_yside = [&](P3dSide& sd, bool send) {
      for (int p = 0; p < M2; p++) {
        const long long nyp = d.jj.sz[p];
        const bool self_redirect = send && M2 > 1 && p == d.jpid;
        const bool remote = send && !self_redirect && M2 > 1 && p2p;
        const int buf = send ? ((self_redirect || remote) ? rcv : snd) : cur;
        if (send) {     // forward: this rank's Y stage writes [xb][z][y in jj(p)][xi]
          long long off = self_redirect ? offs(yz_z, d.jpid) : offs(yz_y, p);
          if (remote) {     // peer p's Z side: block from sender q' holds y in jj(p), z in kj(q')
            off = 0;
            for (int q = 0; q < d.jpid; q++) off += nv * nxb * nyp * d.kj.sz[q] * W;
          }
          add_seg_blk(sd, buf, remote ? d.rank_of(d.ipid, p) : -1, off, d.jj.st[p] - 1, (int)nyp,
                      W, 0, 0, 1, W, kj * nyp * W, nyp * W, yz_y[p]);
        } else {        // backward: this rank's Y stage gathers rows y from [xb][y in jj(p)][z][xi]
          add_seg_blk(sd, buf, -1, offs(yz_y, p), d.jj.st[p] - 1, (int)nyp,
                      kj * W, 0, 0, 1, W, nyp * kj * W, W, yz_y[p]);
        }
      }
    };
    auto yz_zside = [&](P3dSide& sd, bool send) {
      for (int q = 0; q < M2; q++) {
        const long long nzq = d.kj.sz[q];
        const bool self_redirect = send && M2 > 1 && q == d.jpid;
        const bool remote = send && !self_redirect && M2 > 1 && p2p;
        const int buf = send ? ((self_redirect || remote) ? rcv : snd) : cur;
        if (send) {     // backward: this rank's Z stage writes [xb][y][z in kj(q)][xi]
          long long off = self_redirect ? offs(yz_y, d.jpid) : offs(yz_z, q);
          if (remote) {     // peer q's Y side: block from sender p' holds y in jj(p'), z in kj(q)
            off = 0;
            for (int pp = 0; pp < d.jpid; pp++) off += nv * nxb * d.jj.sz[pp] * nzq * W;
          }
          add_seg_blk(sd, buf, remote ? d.rank_of(d.ipid, q) : -1, off, d.kj.st[q] - 1, (int)nzq,
                      W, 0, 0, 1, W, jj * nzq * W, nzq * W, yz_z[q]);
        } else {        // forward: this rank's Z stage gathers rows z from [xb][z in kj(q)][y][xi]
          add_seg_blk(sd, buf, -1, offs(yz_z, q), d.kj.st[q] - 1, (int)nzq,
                      jj * W, 0, 0, 1, W, nzq * jj * W, W, yz_z[q]);
        }
      }
    };
    // tiles that are adjacent in b share the memory lines of a gathered input: run GATHER_B of them back to back
    const int GATHER_B = 1;      // measured on B200 (1024^3): the x-fastest tile order wins (the far side of the gathering stage is x-contiguous)
    if (!backward) {
      stage_init(s, P3D_R2C, d.nx, (int)ji, (int)kj, nv, 5); s.layx = 1;
      side_init(s.in, d.nx, d.nx, d.nx);
      add_seg(s.in, P3D_BUF_USER_IN, -1, 0, 0, d.nx, 1, nx, nx * ji, dim_real);
      side_init(s.out, d.nxhp, d.nxhpc, d.nxhpc);
      rotate();
      xy_xside(s.out, true);
      push_stage(s);
      if (M1 > 1) { push_exchange_cnt(0, M1, d.ipid, 1, xy_x, xy_y); cur = rcv; } else cur = snd;
      stage_init(s, P3D_C2C_FWD, d.ny, (int)ii, (int)kj, nv, 7);
      side_init(s.in, d.ny, d.ny, d.ny);
      xy_yside(s.in, false);
      side_init(s.out, d.ny, d.nyc, d.nycph);
      rotate();
      yz_yside(s.out, true);
      push_stage(s);
      if (M2 > 1) { push_exchange_cnt(1, M2, d.jpid, 2, yz_y, yz_z); cur = rcv; } else cur = snd;
      stage_init(s, zkind, d.nz, (int)ii, (int)jj, nv, 8); s.bord = GATHER_B;
      side_init(s.in, d.nz, d.nz, d.nz);
      yz_zside(s.in, false);
      side_init(s.out, d.nz, d.nzc, d.nzcph);
      if (d.stride1) add_seg(s.out, P3D_BUF_USER_OUT, -1, 0, 0, d.nzc, 1, nzc * jj, nzc, dim_cplx);
      else           add_seg(s.out, P3D_BUF_USER_OUT, -1, 0, 0, d.nzc, ii * jj, 1, ii, dim_cplx);
      push_stage(s);
    } else {
      stage_init(s, zkind, d.nz, (int)ii, (int)jj, nv, 9);
      side_init(s.in, d.nz, d.nzc, d.nzcph);
      if (d.stride1) add_seg(s.in, P3D_BUF_USER_IN, -1, 0, 0, d.nzc, 1, nzc * jj, nzc, dim_cplx);
      else           add_seg(s.in, P3D_BUF_USER_IN, -1, 0, 0, d.nzc, ii * jj, 1, ii, dim_cplx);
      side_init(s.out, d.nz, d.nz, d.nz);
      rotate();
      yz_zside(s.out, true);
      push_stage(s);
      if (M2 > 1) { push_exchange_cnt(1, M2, d.jpid, 3, yz_z, yz_y); cur = rcv; } else cur = snd;
      stage_init(s, P3D_C2C_BWD, d.ny, (int)ii, (int)kj, nv, 10); s.bord = GATHER_B;
      side_init(s.in, d.ny, d.nyc, d.nycph);
      yz_yside(s.in, false);
      side_init(s.out, d.ny, d.ny, d.ny);
      rotate();
      xy_yside(s.out, true);
      push_stage(s);
      if (M1 > 1) { push_exchange_cnt(0, M1, d.ipid, 4, xy_y, xy_x); cur = rcv; } else cur = snd;
      stage_init(s, P3D_C2R, d.nx, (int)ji, (int)kj, nv, 12); s.layx = 1;
      side_init(s.in, d.nxhp, d.nxhpc, d.nxhpc);
      xy_xside(s.in, false);
      side_init(s.out, d.nx, d.nx, d.nx);
      add_seg(s.out, P3D_BUF_USER_OUT, -1, 0, 0, d.nx, 1, nx, nx * ji, dim_real);
      push_stage(s);
    }
    return tp;
  }
  if (!backward) {
    // ---- K1: X r2c (+ X-prune, + pack for T1).
    // exec_f_r2c ftran.F90:530; fcomm1.F90:239-253; seg_copy_x ftran.F90:554
    stage_init(s, P3D_R2C, d.nx, (int)ji, (int)kj, nv, 5); s.layx = 1;
    side_init(s.in, d.nx, d.nx, d.nx);
    add_seg(s.in, P3D_BUF_USER_IN, -1, 0, 0, d.nx, 1, nx, nx * ji, dim_real);
    side_init(s.out, d.nxhp, d.nxhpc, d.nxhpc);
    rotate();
    if (M1 == 1) { add_seg(s.out, snd, -1, 0, 0, d.nxhpc, 1, nxhpc, nxhpc * ji, nxhpc * ji * kj); cur = snd; }
    else piecewise(s.out, d.ii, M1, d.ipid, ji * kj, true, nv * (long long)(d.ji.st[d.ipid] - 1) * ii * kj, 0, 0);
    push_stage(s);
    if (M1 > 1) { push_exchange(0, M1, d.ipid, 1, d.ii, ji * kj, d.ji, ii * kj); cur = rcv; }   // T1 fcomm1.F90:271
    // ---- K2: Y c2c (+ unpack of T1, + Y-prune and pack for T2).
    // fcomm1.F90:284-320; ftran.F90:581-583; pack_fcomm2 fcomm2.F90:321-386; seg_copy_y ftran.F90:756-757
    stage_init(s, P3D_C2C_FWD, d.ny, (int)ii, (int)kj, nv, 7);
    side_init(s.in, d.ny, d.ny, d.ny);
    if (M1 == 1) add_seg(s.in, cur, -1, 0, 0, d.ny, ii, 1, ii * ny, ii * ny * kj);
    else piecewise(s.in, d.ji, M1, d.ipid, ii * kj, false, 0, 0, 1);
    side_init(s.out, d.ny, d.nyc, d.nycph);
    rotate();
    if (M2 == 1) add_seg(s.out, snd, -1, 0, 0, d.nyc, ii, 1, ii * nyc, ii * nyc * kj);
    else piecewise(s.out, d.jj, M2, d.jpid, ii * kj, true, nv * (long long)(d.kj.st[d.jpid] - 1) * ii * jj, 0, 1);
    push_stage(s);
    if (M2 == 1) cur = snd;
    else { push_exchange(1, M2, d.jpid, 2, d.jj, ii * kj, d.kj, ii * jj); cur = rcv; }          // T2 fcomm2.F90:313
    // ---- K3: Z transform (+ Z-prune, + STRIDE1 output reorder).  ftran.F90:605-683; seg_copy_z :645-646
    if (d.stride1) stage_init(s, zkind, d.nz, (int)ii, (int)jj, nv, 8);
    else           stage_init(s, zkind, d.nz, (int)(ii * jj), 1, nv, 8);
    side_init(s.in, d.nz, d.nz, d.nz);
    if (M2 == 1) add_seg(s.in, cur, -1, 0, 0, d.nz, ii * jj, 1, d.stride1 ? ii : 0, ii * jj * nz);
    else piecewise(s.in, d.kj, M2, d.jpid, ii * jj, false, 0, 0, 2);
    side_init(s.out, d.nz, d.nzc, d.nzcph);
    if (d.stride1) add_seg(s.out, P3D_BUF_USER_OUT, -1, 0, 0, d.nzc, 1, nzc * jj, nzc, dim_cplx);
    else           add_seg(s.out, P3D_BUF_USER_OUT, -1, 0, 0, d.nzc, ii * jj, 1, 0, dim_cplx);
    push_stage(s);
  } else {
    // ---- K4: Z inverse (+ zero-pad of pruned Z, + pack for T3).  btran.F90:437-509
    if (d.stride1) stage_init(s, zkind, d.nz, (int)ii, (int)jj, nv, 9);
    else           stage_init(s, zkind, d.nz, (int)(ii * jj), 1, nv, 9);
    side_init(s.in, d.nz, d.nzc, d.nzcph);
    if (d.stride1) add_seg(s.in, P3D_BUF_USER_IN, -1, 0, 0, d.nzc, 1, nzc * jj, nzc, dim_cplx);
    else           add_seg(s.in, P3D_BUF_USER_IN, -1, 0, 0, d.nzc, ii * jj, 1, 0, dim_cplx);
    side_init(s.out, d.nz, d.nz, d.nz);
    rotate();
    if (M2 == 1) { add_seg(s.out, snd, -1, 0, 0, d.nz, ii * jj, 1, d.stride1 ? ii : 0, ii * jj * nz); cur = snd; }
    else piecewise(s.out, d.kj, M2, d.jpid, ii * jj, true, nv * (long long)(d.jj.st[d.jpid] - 1) * ii * kj, 0, 2);
    push_stage(s);
    if (M2 > 1) { push_exchange(1, M2, d.jpid, 3, d.kj, ii * jj, d.jj, ii * kj); cur = rcv; }   // T3 bcomm1.F90:295
    // ---- K5: Y inverse (+ unpack of T3 with zero-fill of the pruned Y band, + pack for T4)
    // unpack_bcomm1 bcomm1.F90:309-380; btran.F90:618-623; bcomm2.F90:233-273
    stage_init(s, P3D_C2C_BWD, d.ny, (int)ii, (int)kj, nv, 10);
    side_init(s.in, d.ny, d.nyc, d.nycph);
    if (M2 == 1) add_seg(s.in, cur, -1, 0, 0, d.nyc, ii, 1, ii * nyc, ii * nyc * kj);
    else piecewise(s.in, d.jj, M2, d.jpid, ii * kj, false, 0, 0, 1);
    side_init(s.out, d.ny, d.ny, d.ny);
    rotate();
    if (M1 == 1) add_seg(s.out, snd, -1, 0, 0, d.ny, ii, 1, ii * ny, ii * ny * kj);
    else piecewise(s.out, d.ji, M1, d.ipid, ii * kj, true, nv * (long long)(d.ii.st[d.ipid] - 1) * ji * kj, 0, 1);
    push_stage(s);
    if (M1 == 1) cur = snd;
    else { push_exchange(0, M1, d.ipid, 4, d.ji, ii * kj, d.ii, ji * kj); cur = rcv; }          // T4 bcomm2.F90:281
    // ---- K6: X c2r (+ unpack of T4 with zero-fill of x > nxhpc).  bcomm2.F90:290-313; btran.F90:655
    stage_init(s, P3D_C2R, d.nx, (int)ji, (int)kj, nv, 12); s.layx = 1;
    side_init(s.in, d.nxhp, d.nxhpc, d.nxhpc);
    if (M1 == 1) add_seg(s.in, cur, -1, 0, 0, d.nxhpc, 1, nxhpc, nxhpc * ji, nxhpc * ji * kj);
    else piecewise(s.in, d.ii, M1, d.ipid, ji * kj, false, 0, 0, 0);
    side_init(s.out, d.nx, d.nx, d.nx);
    add_seg(s.out, P3D_BUF_USER_OUT, -1, 0, 0, d.nx, 1, nx, nx * ji, dim_real);
    push_stage(s);
  }
  return tp;
}

// ------------------------------------------------------------------------------------
// Pipelined group of a peer-to-peer plan (P3DFFT_B200_OVERLAP=C).
//
// On a multi-GPU grid the stage in front of the LAST exchange of a transform is bound by NVLink (its stores go to
// the peers), the stage behind it is local and HBM-bound:  forward  Y -> T2 -> Z,  backward  Y -> T4 -> X.  Both
// stages are split into C chunks along the batch axis they share (forward: the x blocks, dimension a; backward: the
// z planes, dimension b); chunk c of the consumer only needs chunk c of the producer from every peer, so the list
//     P_0  E_0  Q_0   P_1  E_1  Q_1  ...                       (E_c = barrier; Q_c marked `side`)
// lets the executor run Q_c on a second stream while P_(c+1) is storing.  Chunks are equal on every rank in COUNT
// (empty ones stay in the list, their barrier is still collective); a chunk is the same stage with smaller batch
// extents and shifted segment offsets -- the kernels do not know about it.
// Returns false (plan untouched) when the plan holds no  stage, p2p exchange, stage  triple.
// ------------------------------------------------------------------------------------
inline void shift_side(P3dSide& sd, bool along_a, long long x0) {
  for (int g = 0; g < sd.nseg; g++) {
    P3dSeg& s = sd.seg[g];
    if (along_a) s.off += s.aw > 1 ? (x0 / s.aw) * s.sah + (x0 % s.aw) * s.sa : x0 * s.sa;
    else         s.off += s.bw > 1 ? (x0 / s.bw) * s.sbh + (x0 % s.bw) * s.sb : x0 * s.sb;
  }
}

// `shape` (optional): relative sizes of the chunks, e.g. {1, 3, 3, 3, 1} -- a small first chunk lets the consumer start early,
// a small last one shortens the tail that nothing overlaps; empty = equal chunks.
inline bool split_for_overlap(TransformPlan& tp, int nchunk, int W, const std::vector<int>& shape = std::vector<int>()) {
  const size_t n = tp.steps.size();
  if (!shape.empty()) nchunk = (int)shape.size();
  if (nchunk < 2 || n < 3) return false;
  // the LAST  stage, peer-to-peer exchange, stage  triple of the plan (forward: Y T2 Z; backward: Y T4 X, or Z T3 Y
  // when the row communicator has one rank and no exchange precedes the X stage)
  size_t at = n;
  for (size_t i = n - 3;; i--) {
    if (!tp.steps[i].is_exchange && tp.steps[i + 1].is_exchange && tp.steps[i + 1].ex.p2p && !tp.steps[i + 2].is_exchange &&
        tp.steps[i].chunk < 0 && tp.steps[i + 2].chunk < 0) { at = i; break; }
    if (i == 0) break;
  }
  if (at == n) return false;
  const P3dStage P = tp.steps[at].st, Q = tp.steps[at + 2].st;
  const P3dExchange E = tp.steps[at + 1].ex;
  if (P.nc != Q.nc) return false;
  // the batch axis the two stages share, by their places in the transform (timer slots of the reference, module.F90:106):
  //   X -> Y (5, 7) and Y -> X (10, 12): the z planes, dimension b;   Y -> Z (7, 8) and Z -> Y (9, 10): the x lines, dimension a
  bool along_a;
  int total, gran;
  const bool zplanes = (P.timer == 5 && Q.timer == 7) || (P.timer == 10 && Q.timer == 12);
  const bool xlines = (P.timer == 7 && Q.timer == 8) || (P.timer == 9 && Q.timer == 10);
  if (zplanes && P.nb == Q.nb) { along_a = false; total = P.nb; gran = 1; }
  else if (xlines && P.na == Q.na) { along_a = true; total = P.na; gran = W > 0 ? W : 1; }
  else return false;
  for (const P3dStage* st : {&P, &Q})       // blocked line groups must not be cut
    for (const P3dSide* sd : {&st->in, &st->out})
      for (int g = 0; g < sd->nseg; g++)
        if (along_a && sd->seg[g].aw > 1 && gran % sd->seg[g].aw) return false;
  const long long nblk = (total + gran - 1) / gran;
  std::vector<long long> cum(nchunk + 1, 0);      // cumulative weights: chunk c covers blocks [nblk cum[c] / cum[C], nblk cum[c+1] / cum[C])
  for (int c = 0; c < nchunk; c++) cum[c + 1] = cum[c] + (shape.empty() ? 1 : std::max(shape[c], 1));
  std::vector<Step> group;
  for (int c = 0; c < nchunk; c++) {
    const long long b0 = nblk * cum[c] / cum[nchunk] * gran, b1 = std::min<long long>(nblk * cum[c + 1] / cum[nchunk] * gran, total);
    const int cnt = (int)std::max<long long>(b1 - b0, 0);
    Step a; a.is_exchange = false; a.chunk = c; a.st = P;
    Step e; e.is_exchange = true;  e.chunk = c; e.ex = E;
    Step b; b.is_exchange = false; b.chunk = c; b.side = true; b.st = Q;
    for (P3dStage* st : {&a.st, &b.st}) {
      if (along_a) st->na = cnt; else st->nb = cnt;
      shift_side(st->in, along_a, b0);
      shift_side(st->out, along_a, b0);
    }
    group.push_back(a); group.push_back(e); group.push_back(b);
  }
  std::vector<Step> out(tp.steps.begin(), tp.steps.begin() + at);
  out.insert(out.end(), group.begin(), group.end());
  out.insert(out.end(), tp.steps.begin() + at + 3, tp.steps.end());
  tp.steps.swap(out);
  return true;
}

// ------------------------------------------------------------------------------------
// p3dfft_ftran_r2c_1d (ftran.F90:787-814): the X stage alone, real (nx, jisize, kjsize) -> complex
// (nxhp, jisize, kjsize) [= nx+2 reals per line], no pruning, no transpose.
// ------------------------------------------------------------------------------------
inline TransformPlan build_r2c_1d_plan(const Decomp& d) {
  TransformPlan tp;
  const long long nx = d.nx, ji = d.jisize, kj = d.kjsize, nxhp = d.nxhp;
  if (ji * kj <= 0) return tp;                                   // ftran.F90:809
  P3dStage s;
  stage_init(s, P3D_R2C, d.nx, (int)ji, (int)kj, 1, 5); s.layx = 1;
  side_init(s.in, d.nx, d.nx, d.nx);
  add_seg(s.in, P3D_BUF_USER_IN, -1, 0, 0, d.nx, 1, nx, nx * ji, nx * ji * kj);
  side_init(s.out, d.nxhp, d.nxhp, d.nxhp);
  add_seg(s.out, P3D_BUF_USER_OUT, -1, 0, 0, d.nxhp, 1, nxhp, nxhp * ji, nxhp * ji * kj);
  if (!factorize(s.nfft, s.fac, &s.nfac)) {
    char m[128]; snprintf(m, sizeof m, "transform length %d needs a prime factor > 4096 (unsupported)", s.nfft);
    tp.error = m; return tp;
  }
  Step st; st.is_exchange = false; st.st = s; tp.steps.push_back(st);
  return tp;
}

// ------------------------------------------------------------------------------------
// Real-data pencil transposes rtran_x2y / rtran_y2x / rtran_x2z / rtran_z2x (module.F90:1061-1361):
//   x2y: (nx, jisize, kjsize)      -> (iiisize, ny, kjsize)     over the row communicator
//   y2x: (iiisize, ny, kjsize)     -> (nx, jisize, kjsize)
//   x2z: (nx, jisize, kjsize)      -> (ijsize, jisize, nz)      over the column communicator
//   z2x: (ijsize, jisize, nz)      -> (nx, jisize, kjsize)
// Each is pack -> alltoallv -> unpack in the reference (three passes over the data plus the network).  Here
// the pack is ONE P3D_RCOPY stage that writes every block where its consumer reads it (the peer's receive
// buffer with peer-to-peer plans, else the send buffer; the rank's own block straight into its own receive
// buffer) and the unpack is a second one; a 1-rank communicator degenerates to a single direct copy.  Blocks
// keep the reference's layouts and its Ii/Ji/Ij/Kj counts and displacements (setup.F90:522-549), in REAL
// elements.
// ------------------------------------------------------------------------------------
enum RtranKind { RTRAN_X2Y = 0, RTRAN_Y2X = 1, RTRAN_X2Z = 2, RTRAN_Z2X = 3 };

// real elements of the source / destination array of a transpose on this rank
inline long long rtran_elems(const Decomp& d, int which, bool dest) {
  const long long x = (long long)d.nx * d.jisize * d.kjsize;
  const long long y = (long long)d.iiisize * d.ny * d.kjsize;
  const long long z = (long long)d.ijsize * d.jisize * d.nz;
  switch (which) {
    case RTRAN_X2Y: return dest ? y : x;
    case RTRAN_Y2X: return dest ? x : y;
    case RTRAN_X2Z: return dest ? z : x;
    default:        return dest ? x : z;
  }
}

// complex elements a work buffer must hold so that every rank can stage any of the four transposes
// (the same bound on all ranks, so that growing the buffers is a collective decision)
inline long long rtran_work_elems(const Decomp& d) {
  auto mx = [](const BlockMap& m) { int v = 0; for (int x : m.sz) v = std::max(v, x); return (long long)v; };
  const long long ji = mx(d.ji), kj = mx(d.kj), iii = mx(d.iii), ij = mx(d.ij);
  long long r = (long long)d.nx * ji * kj;
  r = std::max(r, iii * d.ny * kj);
  r = std::max(r, ij * ji * d.nz);
  return (r + 1) / 2;
}

// dstart / dend / dsize of the destination array (module.F90:1118-1127, 1195-1204, 1273-1282, 1349-1358)
inline void rtran_dims(const Decomp& d, int which, int* st, int* en, int* sz) {
  if (which == RTRAN_X2Y) {
    st[0] = d.iiistart; en[0] = d.iiiend; sz[0] = d.iiisize;
    st[1] = 1; en[1] = d.ny; sz[1] = d.ny;
    st[2] = d.kjstart; en[2] = d.kjend; sz[2] = d.kjsize;
  } else if (which == RTRAN_X2Z) {
    st[0] = d.ijstart; en[0] = d.ijend; sz[0] = d.ijsize;
    st[1] = d.jistart; en[1] = d.jiend; sz[1] = d.jisize;
    st[2] = 1; en[2] = d.nz; sz[2] = d.nz;
  } else {
    st[0] = 1; en[0] = d.nx; sz[0] = d.nx;
    st[1] = d.jistart; en[1] = d.jiend; sz[1] = d.jisize;
    st[2] = d.kjstart; en[2] = d.kjend; sz[2] = d.kjsize;
  }
}

inline TransformPlan build_rtran_plan(const Decomp& d, int which, bool p2p, int real_bytes) {
  TransformPlan tp;
  const bool row = which == RTRAN_X2Y || which == RTRAN_Y2X;      // row communicator (x <-> y), else column (x <-> z)
  const bool from_x = which == RTRAN_X2Y || which == RTRAN_X2Z;   // the source is the X pencil
  const int M = row ? d.iproc : d.jproc, me = row ? d.ipid : d.jpid;
  if (M > P3D_MAXSEG) { tp.error = "processor grid dimension exceeds P3D_MAXSEG"; return tp; }
  const long long nx = d.nx, ji = d.jisize, kj = d.kjsize;
  const BlockMap& xb = row ? d.iii : d.ij;          // x blocks of the far pencil
  const BlockMap& fb = row ? d.ji : d.kj;           // blocks of the far pencil's long axis (y or z) = this rank's peers' shares
  const long long xs = xb.sz[me];                   // iiisize / ijsize
  const int nfar = row ? d.ny : d.nz;
  const int snd = P3D_BUF_A, rcv = P3D_BUF_B;
  auto world_rank = [&](int p) { return row ? d.rank_of(p, d.jpid) : d.rank_of(d.ipid, p); };
  // X-pencil side: axis x, lines y (a), planes z (b); block p = (xb.sz[p], ji, kj) at (xb.st[p]-1)*ji*kj
  auto xside_user = [&](P3dSide& sd, int buf) {
    side_init(sd, d.nx, d.nx, d.nx);
    add_seg(sd, buf, -1, 0, 0, d.nx, 1, nx, nx * ji, nx * ji * kj);
  };
  auto xside_blocks = [&](P3dSide& sd, bool send) {
    side_init(sd, d.nx, d.nx, d.nx);
    for (int p = 0; p < M; p++) {
      const long long w = xb.sz[p];
      int buf = send ? snd : rcv, peer = -1;
      long long off = (long long)(xb.st[p] - 1) * ji * kj;
      if (send && p == me) { buf = rcv; off = (long long)(fb.st[me] - 1) * xs * (row ? kj : ji); }
      else if (send && p2p) { buf = rcv; peer = world_rank(p); off = (long long)(fb.st[me] - 1) * w * (row ? kj : ji); }
      add_seg(sd, buf, peer, off, xb.st[p] - 1, (int)w, 1, w, w * ji, w * ji * kj);
    }
  };
  // far-pencil side: axis = y (row) or z (column).  row: lines x (a, xs of them), planes z (b);
  // column: lines (x,y) merged (a, xs*ji of them, contiguous), one plane.
  // block q = (xs, fb.sz[q], kj) [row] / (xs, ji, fb.sz[q]) [column] at (fb.st[q]-1) * xs * (kj | ji)
  const long long fps = row ? xs : xs * ji;         // stride of the far axis in the user array and inside a block
  auto fside_user = [&](P3dSide& sd, int buf) {
    side_init(sd, nfar, nfar, nfar);
    add_seg(sd, buf, -1, 0, 0, nfar, fps, 1, row ? xs * d.ny : 0, 0);
  };
  auto fside_blocks = [&](P3dSide& sd, bool send) {
    side_init(sd, nfar, nfar, nfar);
    for (int q = 0; q < M; q++) {
      const long long w = fb.sz[q];
      int buf = send ? snd : rcv, peer = -1;
      long long off = (long long)(fb.st[q] - 1) * xs * (row ? kj : ji);
      if (send && q == me) { buf = rcv; off = (long long)(xb.st[me] - 1) * ji * kj; }
      else if (send && p2p) {     // peer q's X side: the block from this rank holds x in xb(me), the peer's y / z share
        buf = rcv; peer = world_rank(q);
        off = (long long)(xb.st[me] - 1) * (row ? w * kj : ji * w);
      }
      add_seg(sd, buf, peer, off, fb.st[q] - 1, (int)w, fps, 1, row ? xs * w : 0, 0);
    }
  };
  auto push = [&](P3dStage& s) {
    if ((long long)s.na * s.nb * s.nc <= 0 || s.n <= 0) return;
    s.nfac = 0;
    Step st; st.is_exchange = false; st.st = s; tp.steps.push_back(st);
  };
  const int fa = row ? (int)xs : (int)(xs * ji), fbn = row ? (int)kj : 1;   // batch extents of the far-axis stages
  P3dStage s;
  if (M == 1) {         // one rank along this communicator: the two pencils coincide up to the axis naming
    stage_init(s, P3D_RCOPY, d.nx, (int)ji, (int)kj, 1, 0); s.layx = 1;
    xside_user(s.in, P3D_BUF_USER_IN);
    xside_user(s.out, P3D_BUF_USER_OUT);
    push(s);
    return tp;
  }
  Step ex; ex.is_exchange = true; P3dExchange& e = ex.ex; memset(&e, 0, sizeof e);
  e.comm = row ? 0 : 1; e.npeer = M; e.self = me; e.sendbuf = snd; e.recvbuf = rcv; e.timer = 0;
  e.p2p = p2p ? 1 : 0; e.ebytes = real_bytes;
  for (int p = 0; p < M; p++) {
    const long long xo = (long long)(xb.st[p] - 1) * ji * kj, xc = (long long)xb.sz[p] * ji * kj;                  // Ii / Ij
    const long long fo = (long long)(fb.st[p] - 1) * xs * (row ? kj : ji), fc = (long long)fb.sz[p] * xs * (row ? kj : ji);   // Ji / Kj
    e.sndoff[p] = from_x ? xo : fo; e.sndcnt[p] = from_x ? xc : fc;
    e.rcvoff[p] = from_x ? fo : xo; e.rcvcnt[p] = from_x ? fc : xc;
  }
  if (from_x) {
    stage_init(s, P3D_RCOPY, d.nx, (int)ji, (int)kj, 1, 0); s.layx = 1;
    xside_user(s.in, P3D_BUF_USER_IN);
    xside_blocks(s.out, true);
    push(s);
    tp.steps.push_back(ex);
    stage_init(s, P3D_RCOPY, nfar, fa, fbn, 1, 0);
    fside_blocks(s.in, false);
    fside_user(s.out, P3D_BUF_USER_OUT);
    push(s);
  } else {
    stage_init(s, P3D_RCOPY, nfar, fa, fbn, 1, 0);
    fside_user(s.in, P3D_BUF_USER_IN);
    fside_blocks(s.out, true);
    push(s);
    tp.steps.push_back(ex);
    stage_init(s, P3D_RCOPY, d.nx, (int)ji, (int)kj, 1, 0); s.layx = 1;
    xside_blocks(s.in, false);
    xside_user(s.out, P3D_BUF_USER_OUT);
    push(s);
  }
  return tp;
}


}  // namespace p3d
