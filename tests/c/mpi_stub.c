/* TEST CODE: a stand-alone "MPI library" (libmpi_stub.so) made from the single-header stand-in include/mpi_shim/mpi.h
 * with external linkage and MPICH-family handle values.  A driver compiled against tests/c/mpi_stub/mpi.h and linked
 * with it is in the position of an application built with its own MPI: it hands MPI_Comm_c2f(MPI_COMM_WORLD) -- a value
 * the P3DFFT library has never seen -- to p3dfft_setup, and the library finds MPI_Comm_f2c / MPI_Comm_rank /
 * MPI_Comm_size / MPI_Bcast in the process through dlsym to bootstrap its own (NCCL) communicator. */
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
#define P3D_MPI_STUB_LIBRARY 1
#define static          /* every function and the state of the header get external linkage in this one translation unit */
#define inline
#include "../../include/mpi_shim/mpi.h"
#undef static
#undef inline
