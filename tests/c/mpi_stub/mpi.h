/* TEST CODE: the header of the stand-alone MPI stub (tests/c/mpi_stub.c): prototypes and MPICH-family handle values only.
 * A program compiled against it contains no MPI code of its own, exactly like one compiled against a real <mpi.h>. */
#ifndef P3D_MPI_STUB_CLIENT_H
#define P3D_MPI_STUB_CLIENT_H
#ifdef __cplusplus
extern "C" {
#endif
typedef int MPI_Comm;
typedef int MPI_Datatype;
typedef int MPI_Op;
typedef int MPI_Fint;
typedef struct { int MPI_SOURCE, MPI_TAG, MPI_ERROR; } MPI_Status;
#define MPI_SUCCESS 0
#define MPI_COMM_WORLD 0x44000000
#define MPI_COMM_NULL 0x04000000
enum { MPI_CHAR = 0x4c000101, MPI_BYTE = 0x4c00010d, MPI_INT = 0x4c000405, MPI_LONG = 0x4c000807, MPI_FLOAT = 0x4c00040a,
       MPI_DOUBLE = 0x4c00080b, MPI_UNSIGNED = 0x4c000406, MPI_LONG_LONG = 0x4c000809 };
#define MPI_REAL MPI_FLOAT
#define MPI_DOUBLE_PRECISION MPI_DOUBLE
#define MPI_INTEGER MPI_INT
enum { MPI_SUM = 1, MPI_MAX, MPI_MIN, MPI_PROD };
int MPI_Init(int* argc, char*** argv);
int MPI_Initialized(int* flag);
int MPI_Finalize(void);
int MPI_Abort(MPI_Comm c, int code);
int MPI_Comm_size(MPI_Comm c, int* n);
int MPI_Comm_rank(MPI_Comm c, int* r);
int MPI_Barrier(MPI_Comm c);
int MPI_Bcast(void* buf, int count, MPI_Datatype t, int root, MPI_Comm c);
int MPI_Reduce(const void* sbuf, void* rbuf, int count, MPI_Datatype t, MPI_Op op, int root, MPI_Comm c);
int MPI_Allreduce(const void* sbuf, void* rbuf, int count, MPI_Datatype t, MPI_Op op, MPI_Comm c);
double MPI_Wtime(void);
int MPI_Dims_create(int nnodes, int ndims, int* dims);
MPI_Fint MPI_Comm_c2f(MPI_Comm c);
MPI_Comm MPI_Comm_f2c(MPI_Fint f);
#ifdef __cplusplus
}
#endif
#endif
