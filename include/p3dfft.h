/* p3dfft.h -- C interface of the B200 build of the P3DFFT r2c/c2r hot path.
 *
 * Drop-in for the reference's include/p3dfft.h: the same eleven unmangled entry points
 * (all arguments by reference, Fortran style) and the same Cp3dfft_* convenience wrappers,
 * so the reference's C drivers (sample/C/driver_*.c) compile against it unchanged.
 *
 *   entry point                reference definition it replaces
 *   ------------------------   -------------------------------------------------------
 *   p3dfft_setup               build/setup.F90:83   (BIND(C) shim of p3dfft_setup :107)
 *   p3dfft_get_dims            build/module.F90:214 (shim of p3dfft_get_dims :225)
 *   p3dfft_ftran_r2c           build/ftran.F90:469  (shim of p3dfft_ftran_r2c :489)
 *   p3dfft_btran_c2r           build/btran.F90:377  (shim of p3dfft_btran_c2r :396)
 *   p3dfft_ftran_r2c_many      build/ftran.F90:84   (shim of p3dfft_ftran_r2c_many :104)
 *   p3dfft_btran_c2r_many      build/btran.F90:84   (shim of p3dfft_btran_c2r_many :104)
 *   p3dfft_cheby               build/ftran.F90:359  (shim of p3dfft_cheby :383)
 *   p3dfft_cheby_many          build/ftran.F90:321  (shim of p3dfft_cheby_many :339)
 *   p3dfft_clean               build/module.F90:301 (shim of p3dfft_clean :309)
 *   get_timers / set_timers    build/module.F90:709 / :719
 *
 * Build variants mirror the reference's configure switches: compile user code with
 * -DSINGLE_PREC to bind the float library, -DSTRIDE1 when linking the stride-1 build.
 * Arrays may live in host memory (staged over PCIe) or in device memory (used in place).
 */
#ifndef P3DFFT_H_B200
#define P3DFFT_H_B200

#include <stdlib.h>

#ifdef SINGLE_PREC
typedef float p3dfft_real;
#else
typedef double p3dfft_real;
#endif

/* the reference decorates names per Fortran compiler (IBM/INTEL/PGI/CRAY/GNU macros);
 * BIND(C) names need none, the macros are kept so -DGNU etc. on old command lines are harmless */
#define FORT_MOD_NAME(NAME) NAME
#define FORTNAME(NAME) NAME

#ifdef __cplusplus
extern "C" {
#endif

/* ---- library entry points (Fortran calling convention: everything by reference) -------- */
void p3dfft_setup(int* dims, int* nx, int* ny, int* nz, int* comm, int* nxc, int* nyc, int* nzc, int* ow,
                  int* memsize);
void p3dfft_get_dims(int* istart, int* iend, int* isize, int* conf);
void p3dfft_ftran_r2c(p3dfft_real* A, p3dfft_real* B, unsigned char* op);
void p3dfft_btran_c2r(p3dfft_real* A, p3dfft_real* B, unsigned char* op);
void p3dfft_ftran_r2c_many(p3dfft_real* A, int* dim_in, p3dfft_real* B, int* dim_out, int* nv, unsigned char* op);
void p3dfft_btran_c2r_many(p3dfft_real* A, int* dim_in, p3dfft_real* B, int* dim_out, int* nv, unsigned char* op);
void p3dfft_cheby(p3dfft_real* A, p3dfft_real* B, p3dfft_real* Lz);
void p3dfft_cheby_many(p3dfft_real* A, int* dim_in, p3dfft_real* B, int* dim_out, int* nv, p3dfft_real* Lz);
void p3dfft_clean(void);
void get_timers(double* timers);
void set_timers(void);

/* ---- by-value wrappers used by C callers ------------------------------------------------ */
static inline void Cp3dfft_setup(int* dims, int nx, int ny, int nz, int comm, int nxc, int nyc, int nzc,
                                 int overwrite, int* memsize) {
  p3dfft_setup(dims, &nx, &ny, &nz, &comm, &nxc, &nyc, &nzc, &overwrite, memsize);
}
static inline void Cp3dfft_clean(void) { p3dfft_clean(); }
static inline void Cp3dfft_get_dims(int* start, int* end, int* size, int conf) {
  p3dfft_get_dims(start, end, size, &conf);
}
static inline void Cget_timers(double* timers) { get_timers(timers); }
static inline void Cset_timers(void) { set_timers(); }
static inline void Cp3dfft_ftran_r2c(p3dfft_real* A, p3dfft_real* B, unsigned char* op) {
  p3dfft_ftran_r2c(A, B, op);
}
static inline void Cp3dfft_btran_c2r(p3dfft_real* A, p3dfft_real* B, unsigned char* op) {
  p3dfft_btran_c2r(A, B, op);
}
static inline void Cp3dfft_cheby(p3dfft_real* A, p3dfft_real* B, p3dfft_real Lz) { p3dfft_cheby(A, B, &Lz); }
static inline void Cp3dfft_cheby_many(p3dfft_real* A, int dim_in, p3dfft_real* B, int dim_out, int nv,
                                      p3dfft_real Lz) {
  p3dfft_cheby_many(A, &dim_in, B, &dim_out, &nv, &Lz);
}
static inline void Cp3dfft_ftran_r2c_many(p3dfft_real* A, int dim_in, p3dfft_real* B, int dim_out, int nv,
                                          unsigned char* op) {
  p3dfft_ftran_r2c_many(A, &dim_in, B, &dim_out, &nv, op);
}
static inline void Cp3dfft_btran_c2r_many(p3dfft_real* A, int dim_in, p3dfft_real* B, int dim_out, int nv,
                                          unsigned char* op) {
  p3dfft_btran_c2r_many(A, &dim_in, B, &dim_out, &nv, op);
}

#ifdef __cplusplus
}
#endif
#endif
