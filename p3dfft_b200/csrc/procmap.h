// Host-side process-map queries of the reference's "trans2proc support" (build/module.F90:168-176, 788-1054;
// tables filled in build/setup.F90:224-230, 551-577): proc_id2coords, proc_coords2id, proc_dims, proc_neighb,
// search_proc, get_proc_parts.  Pure integer code derived from the Decomp of every rank -- the reference
// gathers the same numbers with MPI_Allgather; here every rank can compute every other rank's extents.
#pragma once
#include <vector>

#include "plan.h"

namespace p3d {

struct ProcMap {
  int iproc = 1, jproc = 1;
  std::vector<int> id2coords;          // [2*id] = ipid, [2*id+1] = jpid       (setup.F90:224-225)
  std::vector<int> dims;               // [(id*2 + (conf-1))*9 + k], k = 0..8: start(3), end(3), size(3)   (setup.F90:552-574)

  void init(const Decomp& me) {
    iproc = me.iproc; jproc = me.jproc;
    const int P = iproc * jproc;
    id2coords.assign(2 * (size_t)P, 0);
    dims.assign((size_t)P * 18, 0);
    for (int r = 0; r < P; r++) {
      Decomp d;
      d.init(me.nx, me.ny, me.nz, iproc, jproc, r, P, me.nxc, me.nyc, me.nzc, me.dims_c, me.stride1);
      id2coords[2 * r] = d.ipid; id2coords[2 * r + 1] = d.jpid;
      for (int conf = 1; conf <= 2; conf++) {
        int* o = &dims[((size_t)r * 2 + (conf - 1)) * 9];
        d.get_dims(o, o + 3, o + 6, conf);
      }
    }
  }
  int nproc() const { return iproc * jproc; }
  // proc_coords2id (setup.F90:226-230); -1 outside the grid
  int coords2id(int ip, int jp) const {
    if (ip < 0 || ip >= iproc || jp < 0 || jp >= jproc) return -1;
    for (int r = 0; r < nproc(); r++) if (id2coords[2 * r] == ip && id2coords[2 * r + 1] == jp) return r;
    return -1;
  }
  // proc_dims(conf, k, id), k 1-based as in the reference
  int pd(int conf, int k, int id) const { return dims[((size_t)id * 2 + (conf - 1)) * 9 + (k - 1)]; }

  // proc_neighb (module.F90:788-825).  The reference accepts coord+orient == iproc (resp. jproc) and then reads
  // proc_coords2id out of bounds; here a neighbour outside the grid is -1.
  int neighb(int base, int orient, int direction) const {
    if (base < 0 || base >= nproc()) return -1;
    if (orient != 1 && orient != -1) return -1;
    if (direction != 1 && direction != 2) return -1;
    const int ci = id2coords[2 * base], cj = id2coords[2 * base + 1];
    return direction == 1 ? coords2id(ci + orient, cj) : coords2id(ci, cj + orient);
  }

  // search_proc (module.F90:832-881): the rank whose block contains (point_i, point_j)
  int search(int point_i, int point_j, int di, int dj, int conf) const {
    if (di < 1 || di > 3 || dj < 1 || dj > 3 || conf < 1 || conf > 2) return -1;
    int id = coords2id(0, 0);
    while (!(point_i < pd(conf, di, id) + pd(conf, di + 6, id))) { id = neighb(id, 1, 1); if (id < 0) return -1; }
    while (!(point_j < pd(conf, dj, id) + pd(conf, dj + 6, id))) { id = neighb(id, 1, 2); if (id < 0) return -1; }
    return id;
  }

  // get_proc_parts (module.F90:888-1054): splits the box (base, size) of the conf-1 (X-pencil, physical space) or
  // conf-2 (Z-pencil, wavenumber space) decomposition into the parts owned by each rank.  parts = nproc rows of
  // 7 ints {proc id, base x, base y, base z, size x, size y, size z}, unused rows -1 -- the contents the
  // reference leaves in proc_parts, including its omission: a part that continues a box in the j direction
  // does not get its i-direction base (column 2 in the i,j,k frame) and keeps -1 there (module.F90:978-985).
  // Returns the number of parts; ierr as the reference (0 ok, 1 bad conf, -1 base point outside every block).
  int parts(int base_x, int base_y, int base_z, int size_x, int size_y, int size_z, int conf, int* out, int* ierr) const {
    const int P = nproc();
    for (int i = 0; i < P * 7; i++) out[i] = -1;
    *ierr = 0;
    int base_i, base_j, base_k, size_i, size_j, size_k, di, dj;
    if (conf == 1) { base_i = base_y; base_j = base_z; base_k = base_x; size_i = size_y; size_j = size_z; size_k = size_x; di = 2; dj = 3; }
    else if (conf == 2) { base_i = base_x; base_j = base_y; base_k = base_z; size_i = size_x; size_j = size_y; size_k = size_z; di = 1; dj = 2; }
    else { *ierr = 1; return 0; }
    int found = search(base_i, base_j, di, dj, conf);
    if (found < 0) { *ierr = -1; return 0; }
    auto at = [&](int part, int col) -> int& { return out[(part - 1) * 7 + (col - 1)]; };   // 1-based like proc_parts
    auto hi = [&](int off, int id) { return pd(conf, off, id) + pd(conf, off + 6, id); };  // first index past the block
    int n = 0;
    bool end_i = false;
    while (!end_i && n < P) {
      n++;
      at(n, 1) = found;
      const int start_id_j = found, start_base_j = base_j, start_size_j = size_j;
      if (base_i + size_i <= hi(di, found)) { at(n, 2) = base_i; at(n, 5) = size_i; end_i = true; }
      else {
        at(n, 2) = base_i; at(n, 5) = hi(di, found) - base_i;
        size_i = base_i + size_i - hi(di, found);
        base_i = hi(di, found);
      }
      at(n, 4) = base_k; at(n, 7) = size_k;
      bool end_j = false, first = true;
      base_j = start_base_j; size_j = start_size_j;
      while (!end_j) {
        if (first) first = false;
        else {
          if (n >= P) break;
          n++;
          at(n, 1) = found; at(n, 4) = at(n - 1, 4); at(n, 5) = at(n - 1, 5); at(n, 7) = at(n - 1, 7);
        }
        if (base_j + size_j <= hi(dj, found)) { at(n, 3) = base_j; at(n, 6) = size_j; end_j = true; }
        else {
          at(n, 3) = base_j; at(n, 6) = hi(dj, found) - base_j;
          size_j = base_j + size_j - hi(dj, found);
          base_j = hi(dj, found);
        }
        if (!end_j) { found = neighb(found, 1, 2); if (found < 0) break; }
      }
      if (!end_i) { found = neighb(start_id_j, 1, 1); if (found < 0) break; }
    }
    if (conf == 1) {          // i,j,k -> x,y,z (module.F90:1024-1052); conf 2 needs no translation
      for (int p = 1; p <= P; p++) {
        const int bi = at(p, 2), bj = at(p, 3), bk = at(p, 4), si = at(p, 5), sj = at(p, 6), sk = at(p, 7);
        at(p, 2) = bk; at(p, 3) = bi; at(p, 4) = bj; at(p, 5) = sk; at(p, 6) = si; at(p, 7) = sj;
      }
    }
    return n;
  }
};

}  // namespace p3d
