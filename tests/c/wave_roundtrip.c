/* wave_roundtrip.c -- C acceptance driver for the P3DFFT C interface of the B200 build.
 *
 * Written for this repository (not a copy of the reference's samples); it exercises the same calls in the
 * same order as a P3DFFT user code does (reference sample/C/driver_sine.c is the model for WHAT is checked):
 * MPI_Init -> Cp3dfft_setup on MPI_Comm_c2f(MPI_COMM_WORLD) -> Cp3dfft_get_dims(1|2) -> forward "fft" of a
 * product of sines on host arrays -> known answer: exactly four non-zero modes of modulus N/8 (two of them
 * on this half-spectrum's kx = 1 plane ... see below) -> normalise -> backward "tff" -> compare with the input.
 * Pass criterion is the reference's: max |error| <= 1e-14 * N / 4 (1e-5 * N / 4 in single precision).
 *
 *   usage: wave_roundtrip [nx ny nz [m1 m2 [repeats]]]          (default 64 64 64, grid from MPI_Dims_create)
 */
#include <math.h>
#include <mpi.h>
#include <stdio.h>
#include <stdlib.h>

#include "p3dfft.h"

int main(int argc, char** argv) {
  int nproc, rank, nx = 64, ny = 64, nz = 64, dims[2] = {0, 0}, reps = 1;
  MPI_Init(&argc, &argv);
  MPI_Comm_size(MPI_COMM_WORLD, &nproc);
  MPI_Comm_rank(MPI_COMM_WORLD, &rank);
  if (argc > 3) { nx = atoi(argv[1]); ny = atoi(argv[2]); nz = atoi(argv[3]); }
  if (argc > 5) { dims[0] = atoi(argv[4]); dims[1] = atoi(argv[5]); }
  if (argc > 6) reps = atoi(argv[6]);
  if (dims[0] * dims[1] != nproc) {
    dims[0] = dims[1] = 0;
    MPI_Dims_create(nproc, 2, dims);
    if (dims[0] > dims[1]) { int t = dims[0]; dims[0] = dims[1]; dims[1] = t; }     /* small M1, as the user guide advises */
  }
  if (rank == 0) printf("wave_roundtrip: %d x %d x %d on a %d x %d grid, %d repetition(s)\n", nx, ny, nz, dims[0], dims[1], reps);

  int memsize[3], is[3], ie[3], isz[3], fs[3], fe[3], fsz[3];
  Cp3dfft_setup(dims, nx, ny, nz, MPI_Comm_c2f(MPI_COMM_WORLD), nx, ny, nz, 1, memsize);
  Cp3dfft_get_dims(is, ie, isz, 1);
  Cp3dfft_get_dims(fs, fe, fsz, 2);
  const long nreal = (long)isz[0] * isz[1] * isz[2], ncplx = (long)fsz[0] * fsz[1] * fsz[2];
  p3dfft_real* a = (p3dfft_real*)malloc(sizeof(p3dfft_real) * (size_t)nreal);
  p3dfft_real* b = (p3dfft_real*)malloc(sizeof(p3dfft_real) * (size_t)ncplx * 2);
  p3dfft_real* c = (p3dfft_real*)malloc(sizeof(p3dfft_real) * (size_t)nreal);
  if (!a || !b || !c) { fprintf(stderr, "out of memory\n"); MPI_Abort(MPI_COMM_WORLD, 2); }

  /* u(x,y,z) = sin(2 pi x/nx) sin(2 pi y/ny) sin(2 pi z/nz) on this rank's X-pencil (x fastest, 1-based starts) */
  const double twopi = 8.0 * atan(1.0);
  for (int k = 0; k < isz[2]; k++)
    for (int j = 0; j < isz[1]; j++)
      for (int i = 0; i < isz[0]; i++)
        a[((long)k * isz[1] + j) * isz[0] + i] = (p3dfft_real)(sin(twopi * (i + is[0] - 1) / nx) * sin(twopi * (j + is[1] - 1) / ny) *
                                                               sin(twopi * (k + is[2] - 1) / nz));
  unsigned char fwd[] = "fft", bwd[] = "tff";
  const double ntot = (double)nx * ny * nz;
  int nspike = 0, nspike_all = 0, bad_spike = 0, bad_all = 0;
  double err = 0.0, err_all = 0.0, t0 = MPI_Wtime();
  for (int r = 0; r < reps; r++) {
    Cp3dfft_ftran_r2c(a, b, fwd);
    if (r == 0) {
      /* half spectrum (kx = 0..nx/2): the wave lives on kx = 1, ky = +-1, kz = +-1, each mode i*(-+)N/8 -> modulus N/8 */
      for (int k = 0; k < fsz[2]; k++)
        for (int j = 0; j < fsz[1]; j++)
          for (int i = 0; i < fsz[0]; i++) {
            const long o = 2 * (((long)k * fsz[1] + j) * fsz[0] + i);
            const double m = hypot((double)b[o], (double)b[o + 1]);
            if (m > 1e-6 * ntot) {
              const int gx = i + fs[0], gy = j + fs[1], gz = k + fs[2];          /* 1-based global mode indices */
              nspike++;
              if (gx != 2 || (gy != 2 && gy != ny) || (gz != 2 && gz != nz) || fabs(m - ntot / 8.0) > 1e-4 * ntot) bad_spike++;
            }
          }
    }
    for (long o = 0; o < 2 * ncplx; o++) b[o] = (p3dfft_real)(b[o] / ntot);
    Cp3dfft_btran_c2r(b, c, bwd);
  }
  const double dt = MPI_Wtime() - t0;
  for (long o = 0; o < nreal; o++) { const double d = fabs((double)c[o] - (double)a[o]); if (d > err) err = d; }
  MPI_Reduce(&nspike, &nspike_all, 1, MPI_INT, MPI_SUM, 0, MPI_COMM_WORLD);
  MPI_Reduce(&bad_spike, &bad_all, 1, MPI_INT, MPI_SUM, 0, MPI_COMM_WORLD);
  MPI_Reduce(&err, &err_all, 1, MPI_DOUBLE, MPI_MAX, 0, MPI_COMM_WORLD);
  double timers[12], tsum[12];
  Cget_timers(timers);
  MPI_Reduce(timers, tsum, 12, MPI_DOUBLE, MPI_SUM, 0, MPI_COMM_WORLD);
  int ok = 1;
  if (rank == 0) {
#ifdef SINGLE_PREC
    const double prec = 1e-5;
#else
    const double prec = 1e-14;
#endif
    ok = nspike_all == 4 && bad_all == 0 && err_all <= prec * ntot * 0.25;
    printf("forward: %d non-zero modes (expected 4, %d misplaced); round trip max error %.3e (limit %.3e)\n", nspike_all, bad_all,
           err_all, prec * ntot * 0.25);
    printf("time per forward+backward pair %.4f s (host arrays, staged over PCIe); stage timers:", dt / reps);
    for (int i = 0; i < 12; i++) printf(" %.4f", tsum[i] / nproc / reps);
    printf("\nResults are %s\n", ok ? "correct" : "incorrect");
  }
  MPI_Bcast(&ok, 1, MPI_INT, 0, MPI_COMM_WORLD);
  Cp3dfft_clean();
  free(a); free(b); free(c);
  MPI_Finalize();
  return ok ? 0 : 1;
}
