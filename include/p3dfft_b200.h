/* p3dfft_b200.h -- non-breaking extensions of the B200 build of the P3DFFT C ABI.
 *
 * The reference library takes a Fortran MPI communicator handle in p3dfft_setup
 * (build/setup.F90:83-107) and relies on MPI for process bootstrap.  The B200 build runs
 * one process per GPU and moves data with NCCL / NVLink peer memory, so the `comm` integer
 * passed to p3dfft_setup is a handle returned by p3dfft_b200_comm_create(), or -- with an MPI
 * loaded in the process -- the caller's own Fortran MPI handle, from which the library then
 * bootstraps itself; 0 without an MPI is the single-rank
 * communicator, any other value is an error.
 *
 * Everything here is plain C: pointers, ints and sizes only.
 *
 * Limits (sizes the reference accepts and this build refuses with a message through the error
 * path below): a processor-grid dimension above 16; a transform length with a prime factor above
 * 4096; a transform whose single line exceeds 200 KB of shared memory (double: ny, nz <= 12800,
 * nx <= ~11000; single: twice that).  Specialised kernels exist for the lengths 64 ... 2048 in
 * powers of two, 384, 768, 1536, 640, 1280 (nx: twice those; sine / cosine third dimension:
 * nz = those / 2 -+ 1); every other length is correct on the any-length kernel at about three
 * times the cost per byte (DESIGN.md sections 4 and 6).
 */
#ifndef P3DFFT_B200_H
#define P3DFFT_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define P3DFFT_B200_UNIQUE_ID_BYTES 128

/* build facts: bit0 = SINGLE_PREC, bit1 = STRIDE1, bit2 = DIMS_C (configure.ac:171-329) */
int p3dfft_b200_build_flags(void);
/* run-time override of the STRIDE1 / DIMS_C build defaults; must precede p3dfft_setup */
void p3dfft_b200_set_layout(int stride1, int dims_c);

/* ---- process bootstrap (replaces MPI_Init/MPI_Comm_*) ------------- */
/* rank 0 obtains an id and ships its 128 bytes to the other ranks by any side channel */
int p3dfft_b200_get_unique_id(void* id128);
/* collective over all ranks; binds this process to CUDA device `device` (<0: current).
 * Returns a handle >= 1 to pass as `comm` to p3dfft_setup, or a negative error code. */
int p3dfft_b200_comm_create(int rank, int size, const void* id128, int device);
void p3dfft_b200_comm_destroy(int handle);

/* ---- error reporting ---------------------------------------------------------------- */
/* mode 0 (default) = the reference's behaviour: print and abort the process where the
 * reference calls MPI_Abort (setup.F90:125-135,181-184; ftran.F90:640-643).
 * mode 1 = record the message, return to the caller.                                      */
void p3dfft_b200_set_error_mode(int mode);
/* copies the last recorded message (empty string if none), clears it, returns its length */
int p3dfft_b200_last_error(char* buf, int buflen);

/* ---- execution control ---------------------------------------------------------------- */
/* transforms are issued on this CUDA stream (cudaStream_t passed as void*; NULL = own)    */
void p3dfft_b200_set_stream(void* cuda_stream);
/* back to the library's own stream                                                         */
void p3dfft_b200_reset_stream(void);
/* async != 0: ftran/btran return after enqueueing (device pointers only, timers not
 * updated); the caller synchronises with p3dfft_b200_sync() or on its stream.            */
void p3dfft_b200_set_async(int async);
void p3dfft_b200_sync(void);
/* number of kernels launched by the library since the last call with reset != 0          */
long long p3dfft_b200_launch_count(int reset);
/* ... of which launches of the specialised power-of-two kernels (fft_fast.cuh)              */
long long p3dfft_b200_fast_launch_count(int reset);
/* on != 0: use only the any-length kernel (A/B checks; also env P3DFFT_B200_GENERIC)        */
void p3dfft_b200_force_generic(int on);
/* Peer-to-peer transposes (default on when every rank of a row/column can map its peers' work
 * buffers with CUDA IPC): the stage kernels store each block straight into the destination rank's
 * receive buffer over NVLink and the exchange step is only a barrier.  on = 0 keeps grouped
 * ncclSend/ncclRecv.  Must precede p3dfft_setup (also env P3DFFT_B200_P2P=0/1).               */
void p3dfft_b200_set_p2p(int on);
/* env P3DFFT_B200_OVERLAP=C (default 4; 0 or 1: off): the last stage-exchange-stage triple of a peer-to-peer transform runs as C
 * chunks, the local consumer chunks on a second stream beside the NVLink-bound producer (P3DFFT_B200_OVERLAP_SMS = SMs left to
 * them, default 74); the per-stage timers then only cover the producer side.
 * env P3DFFT_B200_R32 = 0 / 1 / unset: two-pass (radix-32) 512/1024-point c2c schedules never / wherever they exist / by the
 * measured rule (1024-point stages that write whole tiles contiguously).  P3DFFT_B200_BULK = 0 / 1 / unset: bulk asynchronous
 * (TMA, cp.async.bulk) tile stores never / wherever the output rows allow / for stages storing into a peer's memory.
 * Both are read at p3dfft_setup.                                                                                          */
/* env P3DFFT_B200_FLAGBAR=0: the barrier that orders the peer-to-peer transposes is a one-float NCCL all-reduce instead of the
 * default one-CTA kernel exchanging epoch flags through peer-mapped memory.  P3DFFT_B200_SCOPED=0: every such barrier spans
 * the world instead of signalling to all ranks and waiting only for the ranks of the exchange's row or column (default).  */
int p3dfft_b200_p2p_active(void);
/* on != 0: keep the reference's pack-buffer layouts and exact alltoallv counts in the
 * library's own work buffers instead of the tile-blocked B200 layouts (plan.h); results are
 * identical, only the order of elements inside the internal buffers changes
 * (also env P3DFFT_B200_PLAIN)                                                              */
void p3dfft_b200_plain_layout(int on);
/* Width of one tile row of the internal layouts: 0 = planner's rule (128 bytes unless a Y/Z length exceeds
 * 1024), 64 or 128 = forced.  Takes effect at the next p3dfft_setup (also env P3DFFT_B200_ROWB)      */
void p3dfft_b200_row_bytes(int rb);

/* ---- remaining public routines of the reference's Fortran module (build/module.F90:178-186) ----------
 * The reference exports these from `module p3dfft` only (no BIND(C) shim); fortran/p3dfft.F90 binds the
 * module names to the symbols below.  `real` = double, or float in libp3dfft_single.so; arrays may be
 * host or device pointers like everywhere else.                                                        */
/* p3dfft_ftran_r2c_1d (build/ftran.F90:787-814): X transform only, real (nx, jisize, kjsize) ->
 * complex (nxhp, jisize, kjsize), i.e. nx+2 reals per line; no pruning, no transpose                   */
void p3dfft_ftran_r2c_1d(void* rXgYZ, void* cXgYZ);
/* rtran_x2y / rtran_y2x / rtran_x2z / rtran_z2x (build/module.F90:1061, 1137, 1214, 1292): real-data pencil
 * transposes  (nx, jisize, kjsize) <-> (iiisize, ny, kjsize)  over the row communicator and
 * (nx, jisize, kjsize) <-> (ijsize, jisize, nz)  over the column communicator, iii / ij = MapDataToProc(nx,
 * iproc / jproc) (build/setup.F90:305-312).  dstart/dend/dsize receive the 1-based extents of `dest` (may be
 * NULL); *t (may be NULL) is increased by the seconds spent in the exchange, as the reference adds MPI_Wtime
 * around its alltoallv.  The reference's scratch arguments rbuf1 / rbuf2 do not exist: the library stages the
 * blocks in its own work buffers (or straight in the peers' buffers over NVLink).                           */
void p3dfft_b200_rtran_x2y(const void* source, void* dest, int* dstart, int* dend, int* dsize, double* t);
void p3dfft_b200_rtran_y2x(const void* source, void* dest, int* dstart, int* dend, int* dsize, double* t);
void p3dfft_b200_rtran_x2z(const void* source, void* dest, int* dstart, int* dend, int* dsize, double* t);
void p3dfft_b200_rtran_z2x(const void* source, void* dest, int* dstart, int* dend, int* dsize, double* t);
/* p3dfft_get_mpi_info (build/module.F90:280-297): rank, number of ranks and the communicator handle in use  */
void p3dfft_get_mpi_info(int* taskid, int* ntasks, int* comm);
/* proc_id2coords / proc_coords2id / proc_dims tables (build/setup.F90:224-230, 551-577) and proc_neighb
 * (build/module.F90:788-825).  proc_dims: out9 = start(3), end(3), size(3) of rank `id` in conf 1 or 2.
 * Return -1 for arguments outside the grid (where the reference reads out of bounds).                       */
int p3dfft_b200_proc_id2coords(int id, int* ipid, int* jpid);
int p3dfft_b200_proc_coords2id(int ipid, int jpid);
int p3dfft_b200_proc_dims(int conf, int id, int* out9);
int p3dfft_b200_proc_neighb(int base_proc_id, int orient, int direction);
/* get_proc_parts (build/module.F90:888-1054): which ranks own which part of the box (base, size) of the conf-1
 * (physical space, X pencils) or conf-2 (wavenumber space, Z pencils) decomposition.  parts = iproc*jproc rows
 * of 7 ints {proc id, base x, y, z, size x, y, z} as the reference leaves them in proc_parts (unused rows -1).
 * Returns the number of parts; *ierr as the reference (0, 1 = bad conf, -1 = base point outside the grid).    */
int p3dfft_b200_get_proc_parts(int base_x, int base_y, int base_z, int size_x, int size_y, int size_z, int conf,
                               int* parts, int* ierr);

/* ---- wave-space epilogues (what the sample drivers do on the host after a transform) -------------------- */
/* Fused normalisation: every output of p3dfft_ftran_r2c[_many] is multiplied by `forward` and every output of
 * p3dfft_btran_c2r[_many] by `backward` inside the store of the transform's last stage -- the drivers' mult_array
 * pass (sample/C/driver_rand.c:224, :298, driver_spec.c:223) without a second trip through memory.  Default 1, 1
 * (the reference's unnormalised transforms).  p3dfft_cheby keeps its own normalisation (ftran.F90:408-413).      */
void p3dfft_b200_set_scale(double forward, double backward);
/* Power spectrum of this rank's wavenumber array B (get_dims conf 2 layout; host or device pointer), summed over
 * all ranks: E[ik] = sum k2 * |factor * B|^2, ik = int(sqrt(k2) + 0.5) <= kmax, k2 = kx^2 + ky^2 + kz^2 with ky,
 * kz folded about n/2 -- compute_spectrum + MPI_Reduce of sample/C/driver_spec.c:298-384 on the device.  E receives
 * kmax+1 doubles (host or device) on EVERY rank.  Pruned transforms: stored indices are mapped to their modes.   */
void p3dfft_b200_spectrum(const void* B, double factor, double* E, int kmax);

/* ---- host-only planner queries (no GPU needed; used by the CPU test-suite) ------------- */
typedef struct {
  int32_t nx, ny, nz, nxc, nyc, nzc;
  int32_t nxhp, nxhpc, nycph, nzcph;
  int32_t iproc, jproc, ipid, jpid;
  int32_t iistart, iiend, iisize, jistart, jiend, jisize;
  int32_t jjstart, jjend, jjsize, kjstart, kjend, kjsize;
  int32_t padi_work, padi;
  int32_t memsize[3];
  int64_t nm;
  int64_t work_elems;
} p3dfft_b200_decomp;

/* flags: bit0 = single precision (sizes the blocked layouts), bit1 = STRIDE1, bit2 = DIMS_C,
 * bit3 = plain (reference) internal layouts, bit4 = peer-to-peer plan, bit5 / bit6 = force 64- / 128-byte
 * tile rows (default: the planner's rule); bits 8-15 (plan_steps only): number of chunks of the pipelined tail
 * (P3DFFT_B200_OVERLAP, peer-to-peer plans).  Returns 0, or -1 and records the reference's
 * error text (retrievable with p3dfft_b200_last_error).                                   */
int p3dfft_b200_plan_decomp(const int* dims, int nx, int ny, int nz, int rank, int nxc, int nyc, int nzc,
                            int flags, p3dfft_b200_decomp* out);

/* Writes the step sequence of one transform into `steps` (array of P3dStep, see
 * p3dfft_b200/csrc/stage.h: {int32 is_exchange; P3dStage st; P3dExchange ex;}) and returns
 * the number of steps, or -1 on error.  elem_bytes = 8 (double) or 4 (single).            */
int p3dfft_b200_plan_steps(const int* dims, int nx, int ny, int nz, int rank, int nxc, int nyc, int nzc,
                           int flags, int backward, const char* op, int nv, int64_t dim_real, int64_t dim_cplx,
                           int elem_bytes, void* steps, int max_steps);
/* sizeof(P3dStep) as compiled, so bindings can check their struct mirror                  */
int p3dfft_b200_sizeof_step(void);
/* step list of p3dfft_ftran_r2c_1d (which = 100) or of a real-data transpose (which = 0 x2y, 1 y2x, 2 x2z,
 * 3 z2x); flags as above (bit4 = peer-to-peer plan)                                                        */
int p3dfft_b200_plan_aux_steps(const int* dims, int nx, int ny, int nz, int rank, int nxc, int nyc, int nzc, int flags,
                               int which, int elem_bytes, void* steps, int max_steps);
/* dims9 = dstart(3), dend(3), dsize(3) of the destination of transpose `which` on `rank`; returns the complex
 * elements per work buffer the transposes need (the same bound on every rank), -1 on error                */
long long p3dfft_b200_plan_rtran_info(const int* dims, int nx, int ny, int nz, int rank, int which, int flags, int* dims9);
/* get_proc_parts / proc_neighb without a plan (any rank count)                                             */
int p3dfft_b200_plan_proc_parts(const int* dims, int nx, int ny, int nz, int nxc, int nyc, int nzc, int flags, int base_x,
                                int base_y, int base_z, int size_x, int size_y, int size_z, int conf, int* parts, int* ierr);
int p3dfft_b200_plan_proc_neighb(const int* dims, int flags, int base_proc_id, int orient, int direction);

#ifdef __cplusplus
}
#endif
#endif
