/* spec_epilogue.c -- C acceptance driver for the routines beside the transform path of the B200 build.
 *
 * Written for this repository.  WHAT is checked follows the reference's sample/C/driver_spec.c (forward transform of a
 * product of sines plus noise, normalisation by 1/N, shell-summed power spectrum) and the module routines
 * rtran_x2y / rtran_y2x / rtran_x2z / rtran_z2x and p3dfft_ftran_r2c_1d (build/module.F90:1061-1361, build/ftran.F90:787):
 *
 *   1. p3dfft_b200_set_scale(1/N, 1) + Cp3dfft_ftran_r2c  ==  Cp3dfft_ftran_r2c followed by a host normalisation loop
 *   2. p3dfft_b200_spectrum(B)  ==  the host shell sum over every rank's block (summed with MPI_Allreduce)
 *   3. y2x(x2y(A)) == A and z2x(x2z(A)) == A bit for bit, the reported extents tile the global array, and the
 *      transposed arrays hold the field values their global coordinates say they hold
 *   4. p3dfft_ftran_r2c_1d: the kx = 0 coefficient of every line is the sum of the line
 *
 *   usage: spec_epilogue [nx ny nz [m1 m2]]          (default 64 48 40, grid from MPI_Dims_create)
 */
#include <math.h>
#include <mpi.h>
#include <stdio.h>
#include <stdlib.h>

#include "p3dfft.h"
#include "p3dfft_b200.h"

static double field(int x, int y, int z, int nx, int ny, int nz) {      /* global 0-based coordinates */
  const double twopi = 8.0 * atan(1.0);
  const unsigned h = (unsigned)x * 2654435761u ^ (unsigned)y * 40503u ^ (unsigned)z * 2246822519u;
  return sin(twopi * x / nx) * sin(twopi * y / ny) * sin(twopi * z / nz) + 1e-3 * (double)(h % 1000u);
}

int main(int argc, char** argv) {
  int nproc, rank, nx = 64, ny = 48, nz = 40, dims[2] = {0, 0}, fails = 0;
  MPI_Init(&argc, &argv);
  MPI_Comm_size(MPI_COMM_WORLD, &nproc);
  MPI_Comm_rank(MPI_COMM_WORLD, &rank);
  if (argc > 3) { nx = atoi(argv[1]); ny = atoi(argv[2]); nz = atoi(argv[3]); }
  if (argc > 5) { dims[0] = atoi(argv[4]); dims[1] = atoi(argv[5]); }
  if (dims[0] * dims[1] != nproc) {
    dims[0] = dims[1] = 0;
    MPI_Dims_create(nproc, 2, dims);
    if (dims[0] > dims[1]) { int t = dims[0]; dims[0] = dims[1]; dims[1] = t; }
  }
  int memsize[3], is[3], ie[3], isz[3], fs[3], fe[3], fsz[3];
  Cp3dfft_setup(dims, nx, ny, nz, MPI_Comm_c2f(MPI_COMM_WORLD), nx, ny, nz, 1, memsize);
  Cp3dfft_get_dims(is, ie, isz, 1);
  Cp3dfft_get_dims(fs, fe, fsz, 2);
  const long nreal = (long)isz[0] * isz[1] * isz[2], ncplx = (long)fsz[0] * fsz[1] * fsz[2];
  const double ntot = (double)nx * ny * nz;
#ifdef SINGLE_PREC
  const double tol = 1e-5;
#else
  const double tol = 1e-12;
#endif
  p3dfft_real* a = (p3dfft_real*)malloc(sizeof(p3dfft_real) * (size_t)nreal);
  p3dfft_real* b = (p3dfft_real*)malloc(sizeof(p3dfft_real) * (size_t)ncplx * 2);
  p3dfft_real* b2 = (p3dfft_real*)malloc(sizeof(p3dfft_real) * (size_t)ncplx * 2);
  for (int k = 0; k < isz[2]; k++)
    for (int j = 0; j < isz[1]; j++)
      for (int i = 0; i < isz[0]; i++)
        a[((long)k * isz[1] + j) * isz[0] + i] = (p3dfft_real)field(i + is[0] - 1, j + is[1] - 1, k + is[2] - 1, nx, ny, nz);
  unsigned char fwd[] = "fft";

  /* ---- 1. fused normalisation ------------------------------------------------------------------- */
  Cp3dfft_ftran_r2c(a, b, fwd);
  for (long i = 0; i < 2 * ncplx; i++) b[i] = (p3dfft_real)(b[i] / ntot);           /* the drivers' mult_array */
  p3dfft_b200_set_scale(1.0 / ntot, 1.0);
  Cp3dfft_ftran_r2c(a, b2, fwd);
  p3dfft_b200_set_scale(1.0, 1.0);
  double num = 0.0, den = 0.0;
  for (long i = 0; i < 2 * ncplx; i++) { num += ((double)b[i] - b2[i]) * ((double)b[i] - b2[i]); den += (double)b[i] * b[i]; }
  double nd[2] = {num, den}, ndg[2];
  MPI_Allreduce(nd, ndg, 2, MPI_DOUBLE, MPI_SUM, MPI_COMM_WORLD);
  const double e1 = sqrt(ndg[0] / ndg[1]);
  if (!(e1 <= tol)) fails++;
  if (rank == 0) printf("fused normalisation      : rel-L2 %.2e %s\n", e1, e1 <= tol ? "ok" : "FAIL");

  /* ---- 2. power spectrum ------------------------------------------------------------------------ */
  const int kmax = (int)(sqrt((double)nx * nx + (double)ny * ny + (double)nz * nz) * 0.5 + 0.5);
  double* el = (double*)calloc((size_t)kmax + 1, sizeof(double));
  double* eh = (double*)calloc((size_t)kmax + 1, sizeof(double));
  double* ed = (double*)calloc((size_t)kmax + 1, sizeof(double));
  for (int z = 0; z < fsz[2]; z++) {
    int kz = z + fs[2] - 1; if (kz > nz / 2) kz = nz - kz;
    for (int y = 0; y < fsz[1]; y++) {
      int ky = y + fs[1] - 1; if (ky > ny / 2) ky = ny - ky;
      for (int x = 0; x < fsz[0]; x++) {
        const int kx = x + fs[0] - 1, k2 = kx * kx + ky * ky + kz * kz, ik = (int)(sqrt((double)k2) + 0.5);
        const p3dfft_real* p = b + 2 * (((long)z * fsz[1] + y) * fsz[0] + x);
        if (ik <= kmax) el[ik] += (double)k2 * ((double)p[0] * p[0] + (double)p[1] * p[1]);
      }
    }
  }
  MPI_Allreduce(el, eh, kmax + 1, MPI_DOUBLE, MPI_SUM, MPI_COMM_WORLD);
  p3dfft_b200_spectrum(b, 1.0, ed, kmax);
  double emax = 0.0, ediff = 0.0;
  for (int i = 0; i <= kmax; i++) { if (eh[i] > emax) emax = eh[i]; if (fabs(eh[i] - ed[i]) > ediff) ediff = fabs(eh[i] - ed[i]); }
  if (!(ediff <= tol * 10 * emax)) fails++;
  if (rank == 0) printf("power spectrum (%4d bins): max |dE| / max E %.2e %s\n", kmax + 1, ediff / emax, ediff <= tol * 10 * emax ? "ok" : "FAIL");

  /* ---- 3. real-data transposes ------------------------------------------------------------------ */
  int ds[3], de[3], dz[3], bad = 0;
  double t = 0.0;
  const long ymax = (long)((nx + dims[0] - 1) / dims[0] + 1) * ny * isz[2], zmax = (long)((nx + dims[1] - 1) / dims[1] + 1) * isz[1] * nz;
  p3dfft_real* ty = (p3dfft_real*)malloc(sizeof(p3dfft_real) * (size_t)(ymax > zmax ? ymax : zmax));
  p3dfft_real* back = (p3dfft_real*)malloc(sizeof(p3dfft_real) * (size_t)nreal);
  for (int pass = 0; pass < 2; pass++) {
    if (pass == 0) p3dfft_b200_rtran_x2y(a, ty, ds, de, dz, &t); else p3dfft_b200_rtran_x2z(a, ty, ds, de, dz, &t);
    for (int d = 0; d < 3; d++) if (de[d] - ds[d] + 1 != dz[d]) bad++;
    for (int k = 0; k < dz[2]; k++)
      for (int j = 0; j < dz[1]; j++)
        for (int i = 0; i < dz[0]; i++)
          if (ty[((long)k * dz[1] + j) * dz[0] + i] != (p3dfft_real)field(i + ds[0] - 1, j + ds[1] - 1, k + ds[2] - 1, nx, ny, nz)) bad++;
    for (long i = 0; i < nreal; i++) back[i] = (p3dfft_real)-1.0;
    if (pass == 0) p3dfft_b200_rtran_y2x(ty, back, ds, de, dz, &t); else p3dfft_b200_rtran_z2x(ty, back, ds, de, dz, &t);
    for (long i = 0; i < nreal; i++) if (back[i] != a[i]) bad++;
    for (int d = 0; d < 3; d++) if (ds[d] != is[d] || de[d] != ie[d] || dz[d] != isz[d]) bad++;
  }
  int badg = 0;
  MPI_Allreduce(&bad, &badg, 1, MPI_INT, MPI_SUM, MPI_COMM_WORLD);
  if (badg) fails++;
  if (rank == 0) printf("rtran x2y/y2x/x2z/z2x     : %d mismatches, %.3f ms in exchanges %s\n", badg, t * 1e3, badg ? "FAIL" : "ok");

  /* ---- 4. X transform alone ---------------------------------------------------------------------- */
  const int nxhp = nx / 2 + 1;
  p3dfft_real* c1 = (p3dfft_real*)malloc(sizeof(p3dfft_real) * (size_t)2 * nxhp * isz[1] * isz[2]);
  p3dfft_ftran_r2c_1d(a, c1);
  double dcerr = 0.0;
  for (long l = 0; l < (long)isz[1] * isz[2]; l++) {
    double s = 0.0, s1 = 0.0;
    for (int i = 0; i < nx; i++) { s += a[l * nx + i]; s1 += (i & 1) ? -(double)a[l * nx + i] : (double)a[l * nx + i]; }
    double e = fabs(c1[2 * l * nxhp] - s) + fabs(c1[2 * l * nxhp + 1]);
    if (nx % 2 == 0) e += fabs(c1[2 * (l * nxhp + nx / 2)] - s1);            /* Nyquist coefficient = alternating sum */
    if (e > dcerr) dcerr = e;
  }
  double dcg = 0.0;
  MPI_Allreduce(&dcerr, &dcg, 1, MPI_DOUBLE, MPI_MAX, MPI_COMM_WORLD);
  if (!(dcg <= tol * 100 * nx)) fails++;
  if (rank == 0) printf("p3dfft_ftran_r2c_1d      : max |DC/Nyquist error| %.2e %s\n", dcg, dcg <= tol * 100 * nx ? "ok" : "FAIL");

  Cp3dfft_clean();
  if (rank == 0) printf(fails ? "Results are INCORRECT (%d checks failed)\n" : "Results are correct\n", fails);
  free(a); free(b); free(b2); free(el); free(eh); free(ed); free(ty); free(back); free(c1);
  MPI_Finalize();
  return fails ? 1 : 0;
}
