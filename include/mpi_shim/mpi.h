/* mpi.h -- minimal single-header MPI stand-in for running P3DFFT's C drivers on a box without MPI.
 *
 * The reference's sample drivers (sample/C/driver_*.c) use MPI only for bookkeeping: they read the
 * problem size on rank 0 and broadcast it, reduce error norms and timers, and hand
 * MPI_Comm_c2f(MPI_COMM_WORLD) to p3dfft_setup (driver_sine.c:60-126, :233-266).  All data movement
 * of the transform itself happens inside the library (NCCL / NVLink peer stores), so this header
 * only has to provide those bookkeeping calls:
 *
 *   MPI_Init, MPI_Finalize, MPI_Abort, MPI_Comm_size, MPI_Comm_rank, MPI_Barrier, MPI_Bcast, MPI_Reduce,
 *   MPI_Allreduce, MPI_Wtime, MPI_Dims_create, MPI_Comm_c2f, MPI_Comm_f2c
 *
 * Processes are started by any launcher that sets RANK, WORLD_SIZE, LOCAL_RANK, MASTER_ADDR and
 * MASTER_PORT (tools/p3drun.py, or `python -m torch.distributed.run --no-python`); without those
 * variables the program is a single rank.  Collectives go through rank 0 over TCP (a star: they carry
 * a few bytes).  MPI_Comm_c2f(MPI_COMM_WORLD) creates the library's communicator
 * (p3dfft_b200_comm_create, one GPU per rank = LOCAL_RANK) and returns its handle, which is what
 * p3dfft_setup expects in the place of the Fortran MPI handle -- so a driver compiles UNCHANGED:
 *
 *   cc -Iinclude/mpi_shim -Iinclude driver_sine.c -Lp3dfft_b200/lib -lp3dfft -lm -o driver_sine
 *   python tools/p3drun.py -n 4 ./driver_sine
 *
 * Everything is `static`: include it from one translation unit per program (the drivers are single files).
 */
#ifndef P3DFFT_B200_MPI_SHIM_H
#define P3DFFT_B200_MPI_SHIM_H

#include <arpa/inet.h>
#include <errno.h>
#include <netdb.h>
#include <netinet/in.h>
#include <netinet/tcp.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <sys/socket.h>
#include <time.h>
#include <unistd.h>

#ifdef __cplusplus
extern "C" {
#endif

/* library side (include/p3dfft_b200.h) */
int p3dfft_b200_get_unique_id(void* id128);
int p3dfft_b200_comm_create(int rank, int size, const void* id128, int device);

typedef int MPI_Comm;
typedef int MPI_Datatype;
typedef int MPI_Op;
typedef int MPI_Fint;
typedef struct { int MPI_SOURCE, MPI_TAG, MPI_ERROR; } MPI_Status;

#define MPI_SUCCESS 0
#ifdef P3D_MPI_STUB_LIBRARY
/* Built as a stand-alone shared library (tests/c/mpi_stub.c -> libmpi_stub.so): an "application's own MPI" with
 * MPICH-family handle values, whose MPI_Comm_c2f knows nothing about the P3DFFT library -- the library then
 * bootstraps itself from this MPI through dlsym (api.cpp, comm_from_mpi). */
#define MPI_COMM_WORLD 0x44000000
#define MPI_COMM_NULL 0x04000000
enum { MPI_CHAR = 0x4c000101, MPI_BYTE = 0x4c00010d, MPI_INT = 0x4c000405, MPI_LONG = 0x4c000807, MPI_FLOAT = 0x4c00040a,
       MPI_DOUBLE = 0x4c00080b, MPI_UNSIGNED = 0x4c000406, MPI_LONG_LONG = 0x4c000809 };
#else
#define MPI_COMM_WORLD 1
#define MPI_COMM_NULL 0
enum { MPI_CHAR = 1, MPI_BYTE, MPI_INT, MPI_LONG, MPI_FLOAT, MPI_DOUBLE, MPI_UNSIGNED, MPI_LONG_LONG };
#endif
#define MPI_REAL MPI_FLOAT
#define MPI_DOUBLE_PRECISION MPI_DOUBLE
#define MPI_INTEGER MPI_INT
enum { MPI_SUM = 1, MPI_MAX, MPI_MIN, MPI_PROD };

static struct {
  int init, rank, size, local;
  int* fd;          /* rank 0: socket of every other rank; others: fd[0] = socket to rank 0 */
  int lib_comm;     /* handle from p3dfft_b200_comm_create, 0 until MPI_Comm_c2f is called */
} p3d_mpi_ = {0, 0, 1, 0, NULL, 0};

static void p3d_mpi_die_(const char* what) {
  fprintf(stderr, "mpi shim (rank %d): %s: %s\n", p3d_mpi_.rank, what, strerror(errno));
  exit(1);
}
static void p3d_mpi_send_(int fd, const void* buf, size_t n) {
  const char* p = (const char*)buf;
  while (n) {
    ssize_t k = send(fd, p, n, MSG_NOSIGNAL);
    if (k <= 0) { if (errno == EINTR) continue; p3d_mpi_die_("send"); }
    p += k; n -= (size_t)k;
  }
}
static void p3d_mpi_recv_(int fd, void* buf, size_t n) {
  char* p = (char*)buf;
  while (n) {
    ssize_t k = recv(fd, p, n, 0);
    if (k <= 0) { if (k < 0 && errno == EINTR) continue; p3d_mpi_die_("recv (peer gone?)"); }
    p += k; n -= (size_t)k;
  }
}
static size_t p3d_mpi_size_(MPI_Datatype t) {
  switch (t) {
    case MPI_CHAR: case MPI_BYTE: return 1;
    case MPI_INT: case MPI_UNSIGNED: case MPI_FLOAT: return 4;
    default: return 8;
  }
}
static int p3d_mpi_env_(const char* a, const char* b, int dflt) {
  const char* v = getenv(a);
  if (!v && b) v = getenv(b);
  return v ? atoi(v) : dflt;
}

static int MPI_Init(int* argc, char*** argv) {
  (void)argc; (void)argv;
  if (p3d_mpi_.init) return MPI_SUCCESS;
  p3d_mpi_.rank = p3d_mpi_env_("RANK", "P3D_RANK", 0);
  p3d_mpi_.size = p3d_mpi_env_("WORLD_SIZE", "P3D_WORLD_SIZE", 1);
  p3d_mpi_.local = p3d_mpi_env_("LOCAL_RANK", "P3D_LOCAL_RANK", p3d_mpi_.rank);
  p3d_mpi_.init = 1;
  if (p3d_mpi_.size <= 1) { p3d_mpi_.size = 1; p3d_mpi_.rank = 0; return MPI_SUCCESS; }
  const char* addr = getenv("MASTER_ADDR") ? getenv("MASTER_ADDR") : "127.0.0.1";
  /* next to the launcher's own rendezvous port, not on it */
  int port = p3d_mpi_env_("P3D_SHIM_PORT", NULL, p3d_mpi_env_("MASTER_PORT", NULL, 29500) + 1);
  int one = 1;
  if (p3d_mpi_.rank == 0) {
    int ls = socket(AF_INET, SOCK_STREAM, 0);
    if (ls < 0) p3d_mpi_die_("socket");
    setsockopt(ls, SOL_SOCKET, SO_REUSEADDR, &one, sizeof one);
    struct sockaddr_in sa; memset(&sa, 0, sizeof sa);
    sa.sin_family = AF_INET; sa.sin_addr.s_addr = htonl(INADDR_ANY); sa.sin_port = htons((unsigned short)port);
    if (bind(ls, (struct sockaddr*)&sa, sizeof sa) < 0) p3d_mpi_die_("bind");
    if (listen(ls, p3d_mpi_.size) < 0) p3d_mpi_die_("listen");
    p3d_mpi_.fd = (int*)calloc((size_t)p3d_mpi_.size, sizeof(int));
    for (int i = 1; i < p3d_mpi_.size; i++) {
      int c = accept(ls, NULL, NULL);
      if (c < 0) p3d_mpi_die_("accept");
      setsockopt(c, IPPROTO_TCP, TCP_NODELAY, &one, sizeof one);
      int r = -1;
      p3d_mpi_recv_(c, &r, sizeof r);
      if (r <= 0 || r >= p3d_mpi_.size || p3d_mpi_.fd[r]) { errno = EINVAL; p3d_mpi_die_("bad rank in handshake"); }
      p3d_mpi_.fd[r] = c;
    }
    close(ls);
  } else {
    struct addrinfo hints, *res = NULL;
    char ps[16];
    memset(&hints, 0, sizeof hints);
    hints.ai_family = AF_INET; hints.ai_socktype = SOCK_STREAM;
    snprintf(ps, sizeof ps, "%d", port);
    if (getaddrinfo(addr, ps, &hints, &res) != 0 || !res) p3d_mpi_die_("getaddrinfo(MASTER_ADDR)");
    int c = -1;
    for (int attempt = 0; attempt < 600; attempt++) {          /* rank 0 may not be listening yet */
      c = socket(AF_INET, SOCK_STREAM, 0);
      if (c < 0) p3d_mpi_die_("socket");
      if (connect(c, res->ai_addr, res->ai_addrlen) == 0) break;
      close(c); c = -1;
      usleep(100000);
    }
    freeaddrinfo(res);
    if (c < 0) p3d_mpi_die_("connect to rank 0");
    setsockopt(c, IPPROTO_TCP, TCP_NODELAY, &one, sizeof one);
    p3d_mpi_send_(c, &p3d_mpi_.rank, sizeof(int));
    p3d_mpi_.fd = (int*)calloc(1, sizeof(int));
    p3d_mpi_.fd[0] = c;
  }
  return MPI_SUCCESS;
}

static int MPI_Initialized(int* flag) { *flag = p3d_mpi_.init; return MPI_SUCCESS; }
static int MPI_Comm_size(MPI_Comm c, int* n) { (void)c; *n = p3d_mpi_.size; return MPI_SUCCESS; }
static int MPI_Comm_rank(MPI_Comm c, int* r) { (void)c; *r = p3d_mpi_.rank; return MPI_SUCCESS; }

static int MPI_Barrier(MPI_Comm c) {
  (void)c;
  char b = 0;
  if (p3d_mpi_.size == 1) return MPI_SUCCESS;
  if (p3d_mpi_.rank == 0) {
    for (int i = 1; i < p3d_mpi_.size; i++) p3d_mpi_recv_(p3d_mpi_.fd[i], &b, 1);
    for (int i = 1; i < p3d_mpi_.size; i++) p3d_mpi_send_(p3d_mpi_.fd[i], &b, 1);
  } else {
    p3d_mpi_send_(p3d_mpi_.fd[0], &b, 1);
    p3d_mpi_recv_(p3d_mpi_.fd[0], &b, 1);
  }
  return MPI_SUCCESS;
}

static int MPI_Bcast(void* buf, int count, MPI_Datatype t, int root, MPI_Comm c) {
  (void)c;
  const size_t n = (size_t)count * p3d_mpi_size_(t);
  if (p3d_mpi_.size == 1 || n == 0) return MPI_SUCCESS;
  if (p3d_mpi_.rank == 0) {
    if (root != 0) p3d_mpi_recv_(p3d_mpi_.fd[root], buf, n);
    for (int i = 1; i < p3d_mpi_.size; i++) if (i != root) p3d_mpi_send_(p3d_mpi_.fd[i], buf, n);
  } else if (p3d_mpi_.rank == root) {
    p3d_mpi_send_(p3d_mpi_.fd[0], buf, n);
  } else {
    p3d_mpi_recv_(p3d_mpi_.fd[0], buf, n);
  }
  return MPI_SUCCESS;
}

static void p3d_mpi_combine_(void* acc, const void* in, int count, MPI_Datatype t, MPI_Op op) {
#define P3D_MPI_LOOP(TY)                                                                 \
  { TY* a = (TY*)acc; const TY* b = (const TY*)in;                                       \
    for (int i = 0; i < count; i++)                                                      \
      a[i] = op == MPI_SUM ? (TY)(a[i] + b[i]) : op == MPI_PROD ? (TY)(a[i] * b[i])      \
           : op == MPI_MAX ? (a[i] > b[i] ? a[i] : b[i]) : (a[i] < b[i] ? a[i] : b[i]); }
  switch (t) {
    case MPI_INT: P3D_MPI_LOOP(int) break;
    case MPI_UNSIGNED: P3D_MPI_LOOP(unsigned) break;
    case MPI_LONG: P3D_MPI_LOOP(long) break;
    case MPI_LONG_LONG: P3D_MPI_LOOP(long long) break;
    case MPI_FLOAT: P3D_MPI_LOOP(float) break;
    case MPI_DOUBLE: P3D_MPI_LOOP(double) break;
    default: P3D_MPI_LOOP(char) break;
  }
#undef P3D_MPI_LOOP
}

static int MPI_Reduce(const void* sbuf, void* rbuf, int count, MPI_Datatype t, MPI_Op op, int root, MPI_Comm c) {
  (void)c;
  const size_t n = (size_t)count * p3d_mpi_size_(t);
  if (p3d_mpi_.size == 1) { if (rbuf != sbuf) memmove(rbuf, sbuf, n); return MPI_SUCCESS; }
  if (p3d_mpi_.rank == 0) {
    void* acc = malloc(n ? n : 1);
    void* tmp = malloc(n ? n : 1);
    memcpy(acc, sbuf, n);
    for (int i = 1; i < p3d_mpi_.size; i++) {                 /* fixed order: reproducible sums */
      p3d_mpi_recv_(p3d_mpi_.fd[i], tmp, n);
      p3d_mpi_combine_(acc, tmp, count, t, op);
    }
    if (root == 0) memcpy(rbuf, acc, n); else p3d_mpi_send_(p3d_mpi_.fd[root], acc, n);
    free(acc); free(tmp);
  } else {
    p3d_mpi_send_(p3d_mpi_.fd[0], sbuf, n);
    if (p3d_mpi_.rank == root) p3d_mpi_recv_(p3d_mpi_.fd[0], rbuf, n);
  }
  return MPI_SUCCESS;
}

static int MPI_Allreduce(const void* sbuf, void* rbuf, int count, MPI_Datatype t, MPI_Op op, MPI_Comm c) {
  MPI_Reduce(sbuf, rbuf, count, t, op, 0, c);
  return MPI_Bcast(rbuf, count, t, 0, c);
}

static double MPI_Wtime(void) {
  struct timespec ts;
  clock_gettime(CLOCK_MONOTONIC, &ts);
  return (double)ts.tv_sec + 1e-9 * (double)ts.tv_nsec;
}

/* balanced factorisation in non-increasing order; entries that are already non-zero are kept */
static int MPI_Dims_create(int nnodes, int ndims, int* dims) {
  int fixed = 1, nfree = 0;
  for (int i = 0; i < ndims; i++) { if (dims[i] > 0) fixed *= dims[i]; else nfree++; }
  if (nfree == 0 || nnodes % fixed) return nfree == 0 ? MPI_SUCCESS : 1;
  int rest = nnodes / fixed;
  int* f = (int*)calloc((size_t)nfree, sizeof(int));
  for (int i = 0; i < nfree; i++) f[i] = 1;
  for (int p = 2; rest > 1;) {                                 /* hand the prime factors, largest first, to the smallest entry */
    while (rest % p) p++;
    int big = p, r = rest;
    for (int q = p; q <= r; q++) if (r % q == 0) { int isprime = 1; for (int d = 2; d * d <= q; d++) if (q % d == 0) isprime = 0; if (isprime) big = q; }
    int m = 0;
    for (int i = 1; i < nfree; i++) if (f[i] < f[m]) m = i;
    f[m] *= big; rest /= big;
  }
  for (int i = 0; i < nfree; i++) for (int j = i + 1; j < nfree; j++) if (f[j] > f[i]) { int x = f[i]; f[i] = f[j]; f[j] = x; }
  for (int i = 0, k = 0; i < ndims; i++) if (dims[i] <= 0) dims[i] = f[k++];
  free(f);
  return MPI_SUCCESS;
}

/* The "Fortran handle" p3dfft_setup receives is the library's communicator handle. */
static MPI_Fint MPI_Comm_c2f(MPI_Comm c) {
#ifdef P3D_MPI_STUB_LIBRARY
  return (MPI_Fint)c;      /* an MPI of its own: the handle value, nothing else */
#endif
  (void)c;
  if (!p3d_mpi_.init) MPI_Init(NULL, NULL);
  if (!p3d_mpi_.lib_comm) {
    unsigned char id[128];
    memset(id, 0, sizeof id);
    if (p3d_mpi_.size > 1) {
      if (p3d_mpi_.rank == 0 && p3dfft_b200_get_unique_id(id) != 0) { errno = EIO; p3d_mpi_die_("p3dfft_b200_get_unique_id"); }
      MPI_Bcast(id, (int)sizeof id, MPI_BYTE, 0, MPI_COMM_WORLD);
    }
    p3d_mpi_.lib_comm = p3dfft_b200_comm_create(p3d_mpi_.rank, p3d_mpi_.size, id, p3d_mpi_.local);
    if (p3d_mpi_.lib_comm <= 0) { errno = EIO; p3d_mpi_die_("p3dfft_b200_comm_create"); }
  }
  return p3d_mpi_.lib_comm;
}
static MPI_Comm MPI_Comm_f2c(MPI_Fint f) { (void)f; return MPI_COMM_WORLD; }

static int MPI_Finalize(void) {
  if (p3d_mpi_.fd) {
    MPI_Barrier(MPI_COMM_WORLD);
    const int n = p3d_mpi_.rank == 0 ? p3d_mpi_.size : 1;
    for (int i = 0; i < n; i++) if (p3d_mpi_.fd[i] > 0) close(p3d_mpi_.fd[i]);
    free(p3d_mpi_.fd);
    p3d_mpi_.fd = NULL;
  }
  return MPI_SUCCESS;
}
static int MPI_Abort(MPI_Comm c, int code) { (void)c; fflush(NULL); _exit(code ? code : 1); return 0; }

/* a program need not use every call */
static inline void p3d_mpi_unused_(void) {
  (void)MPI_Initialized; (void)MPI_Comm_size; (void)MPI_Comm_rank; (void)MPI_Barrier; (void)MPI_Bcast; (void)MPI_Reduce;
  (void)MPI_Allreduce; (void)MPI_Wtime; (void)MPI_Dims_create; (void)MPI_Comm_c2f; (void)MPI_Comm_f2c; (void)MPI_Finalize;
  (void)MPI_Abort; (void)p3d_mpi_unused_;
}

#ifdef __cplusplus
}
#endif
#endif
