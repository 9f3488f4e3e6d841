/* shim_selftest.c -- checks the collectives of include/mpi_shim/mpi.h (no GPU, no transform). */
#include <mpi.h>
#include <stdio.h>

int main(int argc, char** argv) {
  int n, r, fail = 0;
  MPI_Init(&argc, &argv);
  MPI_Comm_size(MPI_COMM_WORLD, &n);
  MPI_Comm_rank(MPI_COMM_WORLD, &r);
  for (int root = 0; root < n; root++) {                      /* broadcast from every root */
    int v[3] = {r == root ? 100 + root : -1, r == root ? 7 : -1, r == root ? root * root : -1};
    MPI_Bcast(v, 3, MPI_INT, root, MPI_COMM_WORLD);
    if (v[0] != 100 + root || v[1] != 7 || v[2] != root * root) fail++;
  }
  for (int root = 0; root < n; root++) {                      /* reductions to every root */
    double x[2] = {r + 1.0, -(double)r}, s[2] = {0, 0}, mx[2] = {0, 0}, mn[2] = {0, 0};
    MPI_Reduce(x, s, 2, MPI_DOUBLE, MPI_SUM, root, MPI_COMM_WORLD);
    MPI_Reduce(x, mx, 2, MPI_DOUBLE, MPI_MAX, root, MPI_COMM_WORLD);
    MPI_Reduce(x, mn, 2, MPI_DOUBLE, MPI_MIN, root, MPI_COMM_WORLD);
    if (r == root && (s[0] != n * (n + 1) / 2.0 || s[1] != -n * (n - 1) / 2.0 || mx[0] != n || mx[1] != 0 || mn[0] != 1 || mn[1] != -(n - 1.0))) fail++;
  }
  float f = (float)r, fs = 0;
  MPI_Allreduce(&f, &fs, 1, MPI_REAL, MPI_SUM, MPI_COMM_WORLD);
  if (fs != n * (n - 1) / 2.0f) fail++;
  MPI_Barrier(MPI_COMM_WORLD);
  static const int cases[][3] = {{1, 1, 1}, {2, 2, 1}, {4, 2, 2}, {6, 3, 2}, {8, 4, 2}, {12, 4, 3}, {16, 4, 4}, {7, 7, 1}, {64, 8, 8}};
  for (unsigned i = 0; i < sizeof cases / sizeof cases[0]; i++) {
    int d[2] = {0, 0};
    MPI_Dims_create(cases[i][0], 2, d);
    if (d[0] != cases[i][1] || d[1] != cases[i][2]) { fail++; if (r == 0) printf("dims_create(%d) = %d x %d\n", cases[i][0], d[0], d[1]); }
  }
  int d2[2] = {2, 0};
  MPI_Dims_create(8, 2, d2);
  if (d2[0] != 2 || d2[1] != 4) fail++;
  int total = 0;
  MPI_Allreduce(&fail, &total, 1, MPI_INT, MPI_SUM, MPI_COMM_WORLD);
  double t = MPI_Wtime();
  if (r == 0) printf("shim selftest on %d rank(s): %s (t=%.0f)\n", n, total ? "FAILED" : "passed", t);
  MPI_Finalize();
  return total ? 1 : 0;
}
