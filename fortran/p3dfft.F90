! p3dfft.F90 -- Fortran surface of the B200 build: `use p3dfft` keeps working.
!
! The reference's module p3dfft (build/module.F90:84-207) contains the whole library; here it is
! a thin set of ISO_C_BINDING interfaces over the C ABI of libp3dfft.so (include/p3dfft.h).
! Public names, argument order, optional arguments and defaults follow the reference:
!   p3dfft_setup   build/setup.F90:107   (nxcut,nycut,nzcut,overwrite,memsize optional; overwrite
!                                         defaults to .true. when absent, setup.F90:137-141)
!   p3dfft_get_dims module.F90:225, p3dfft_ftran_r2c ftran.F90:489, p3dfft_btran_c2r btran.F90:396,
!   *_many ftran.F90:104 / btran.F90:104, p3dfft_cheby ftran.F90:383, p3dfft_cheby_many ftran.F90:339,
!   p3dfft_clean module.F90:309, get_timers / set_timers module.F90:726 / :742.
!   p3dfft_get_mpi_info module.F90:280, p3dfft_ftran_r2c_1d ftran.F90:787, rtran_x2y / y2x / x2z / z2x
!   module.F90:1061-1361 (rbuf1/rbuf2 are accepted and ignored: the library stages the blocks itself),
!   get_proc_parts module.F90:888 with the public tables proc_id2coords, proc_coords2id, proc_dims, proc_parts
!   (setup.F90:224-230, 551-577) filled at setup, print_buf / print_buf_real module.F90:747 / :765.
! The external (non-module) wrappers of build/wrap.F90:82-150 follow the module.
!
! NOT COMPILED IN THE BUILD IMAGE (no Fortran compiler there: `gfortran -cpp -c p3dfft.F90` is the check to run where
! one exists); shipped as source because a .mod file is compiler specific.  Build with  -DSINGLE_PREC  to bind libp3dfft_single.so.
      module p3dfft
      use iso_c_binding
      implicit none
      private

#ifdef SINGLE_PREC
      integer, parameter, public :: p3dfft_type = c_float
#else
      integer, parameter, public :: p3dfft_type = c_double
#endif
      integer, parameter, public :: r8 = c_double, i8 = c_long_long
      integer, public :: num_thr = 1, padi = 0
      real(r8), public :: timers(12) = 0.0d0
      integer, public :: real_size = 0, complex_size = 0
#ifdef P3DFFT_WITH_MPI
      ! MPI datatype handles as in module.F90:92-100; need the application's MPI (compile with -DP3DFFT_WITH_MPI)
      include 'mpif.h'
#ifdef SINGLE_PREC
      integer, parameter, public :: p3dfft_mpireal = MPI_REAL, p3dfft_mpicomplex = MPI_COMPLEX
#else
      integer, parameter, public :: p3dfft_mpireal = MPI_DOUBLE_PRECISION, p3dfft_mpicomplex = MPI_DOUBLE_COMPLEX
#endif
#endif
      ! trans2proc tables (module.F90:168-176), indexed exactly like the reference's
      integer, public, allocatable :: proc_id2coords(:)        ! (0:2*P-1)
      integer, public, allocatable :: proc_coords2id(:,:)      ! (0:iproc-1, 0:jproc-1)
      integer, public, allocatable :: proc_dims(:,:,:)         ! (2, 9, 0:P-1)
      integer, public, allocatable :: proc_parts(:,:)          ! (P, 7)

      public :: p3dfft_setup, p3dfft_get_dims, p3dfft_ftran_r2c, p3dfft_btran_c2r, &
                p3dfft_ftran_r2c_many, p3dfft_btran_c2r_many, p3dfft_cheby, p3dfft_cheby_many, &
                p3dfft_clean, get_timers, set_timers, p3dfft_get_mpi_info, p3dfft_ftran_r2c_1d, &
                rtran_x2y, rtran_y2x, rtran_x2z, rtran_z2x, get_proc_parts, print_buf, print_buf_real

      interface
        subroutine c_setup(dims,nx,ny,nz,comm,nxc,nyc,nzc,ow,memsize) bind(C,name='p3dfft_setup')
          import :: c_int
          integer(c_int) :: dims(2), nx, ny, nz, comm, nxc, nyc, nzc, ow, memsize(3)
        end subroutine
        subroutine c_get_dims(istart,iend,isize,conf) bind(C,name='p3dfft_get_dims')
          import :: c_int
          integer(c_int) :: istart(3), iend(3), isize(3), conf
        end subroutine
        subroutine c_ftran(a,b,op) bind(C,name='p3dfft_ftran_r2c')
          import :: c_ptr, c_char
          type(c_ptr), value :: a, b
          character(kind=c_char) :: op(*)
        end subroutine
        subroutine c_btran(a,b,op) bind(C,name='p3dfft_btran_c2r')
          import :: c_ptr, c_char
          type(c_ptr), value :: a, b
          character(kind=c_char) :: op(*)
        end subroutine
        subroutine c_ftran_many(a,dim_in,b,dim_out,nv,op) bind(C,name='p3dfft_ftran_r2c_many')
          import :: c_ptr, c_char, c_int
          type(c_ptr), value :: a, b
          integer(c_int) :: dim_in, dim_out, nv
          character(kind=c_char) :: op(*)
        end subroutine
        subroutine c_btran_many(a,dim_in,b,dim_out,nv,op) bind(C,name='p3dfft_btran_c2r_many')
          import :: c_ptr, c_char, c_int
          type(c_ptr), value :: a, b
          integer(c_int) :: dim_in, dim_out, nv
          character(kind=c_char) :: op(*)
        end subroutine
        subroutine c_cheby(a,b,Lz) bind(C,name='p3dfft_cheby')
          import :: c_ptr, p3dfft_type
          type(c_ptr), value :: a, b
          real(p3dfft_type) :: Lz
        end subroutine
        subroutine c_cheby_many(a,dim_in,b,dim_out,nv,Lz) bind(C,name='p3dfft_cheby_many')
          import :: c_ptr, c_int, p3dfft_type
          type(c_ptr), value :: a, b
          integer(c_int) :: dim_in, dim_out, nv
          real(p3dfft_type) :: Lz
        end subroutine
        subroutine c_clean() bind(C,name='p3dfft_clean')
        end subroutine
        subroutine c_get_timers(t) bind(C,name='get_timers')
          import :: c_double
          real(c_double) :: t(12)
        end subroutine
        subroutine c_set_timers() bind(C,name='set_timers')
        end subroutine
        subroutine c_get_mpi_info(taskid,ntasks,comm) bind(C,name='p3dfft_get_mpi_info')
          import :: c_int
          integer(c_int) :: taskid, ntasks, comm
        end subroutine
        subroutine c_r2c_1d(a,b) bind(C,name='p3dfft_ftran_r2c_1d')
          import :: c_ptr
          type(c_ptr), value :: a, b
        end subroutine
        subroutine c_rtran_x2y(s,d,dstart,dend,dsize,t) bind(C,name='p3dfft_b200_rtran_x2y')
          import :: c_ptr, c_int, c_double
          type(c_ptr), value :: s, d
          integer(c_int) :: dstart(3), dend(3), dsize(3)
          real(c_double) :: t
        end subroutine
        subroutine c_rtran_y2x(s,d,dstart,dend,dsize,t) bind(C,name='p3dfft_b200_rtran_y2x')
          import :: c_ptr, c_int, c_double
          type(c_ptr), value :: s, d
          integer(c_int) :: dstart(3), dend(3), dsize(3)
          real(c_double) :: t
        end subroutine
        subroutine c_rtran_x2z(s,d,dstart,dend,dsize,t) bind(C,name='p3dfft_b200_rtran_x2z')
          import :: c_ptr, c_int, c_double
          type(c_ptr), value :: s, d
          integer(c_int) :: dstart(3), dend(3), dsize(3)
          real(c_double) :: t
        end subroutine
        subroutine c_rtran_z2x(s,d,dstart,dend,dsize,t) bind(C,name='p3dfft_b200_rtran_z2x')
          import :: c_ptr, c_int, c_double
          type(c_ptr), value :: s, d
          integer(c_int) :: dstart(3), dend(3), dsize(3)
          real(c_double) :: t
        end subroutine
        integer(c_int) function c_proc_id2coords(id,ipid,jpid) bind(C,name='p3dfft_b200_proc_id2coords')
          import :: c_int
          integer(c_int), value :: id
          integer(c_int) :: ipid, jpid
        end function
        integer(c_int) function c_proc_dims(conf,id,out9) bind(C,name='p3dfft_b200_proc_dims')
          import :: c_int
          integer(c_int), value :: conf, id
          integer(c_int) :: out9(9)
        end function
        integer(c_int) function c_get_proc_parts(bx,by,bz,sx,sy,sz,conf,parts,ierr) bind(C,name='p3dfft_b200_get_proc_parts')
          import :: c_int
          integer(c_int), value :: bx, by, bz, sx, sy, sz, conf
          integer(c_int) :: parts(*), ierr
        end function
      end interface

      contains

      subroutine p3dfft_setup(dims,nx,ny,nz,mpi_comm_in,nxcut,nycut,nzcut,overwrite,memsize)
        integer :: dims(2), nx, ny, nz, mpi_comm_in
        integer, optional, intent(in) :: nxcut, nycut, nzcut
        logical, optional, intent(in) :: overwrite
        integer, optional, intent(out) :: memsize(3)
        integer(c_int) :: nxc, nyc, nzc, ow, mem(3)
        integer(c_int) :: me, np, cm, i, ip, jp, k, rc, d9(9)
        nxc = nx; nyc = ny; nzc = nz; ow = 1
        if (present(nxcut)) nxc = nxcut
        if (present(nycut)) nyc = nycut
        if (present(nzcut)) nzc = nzcut
        if (present(overwrite)) then
          if (.not. overwrite) ow = 0
        end if
        call c_setup(dims, nx, ny, nz, mpi_comm_in, nxc, nyc, nzc, ow, mem)
        if (present(memsize)) memsize = mem
        ! public variables the reference sets in setup (setup.F90:143-147, 224-230, 551-577, 597)
        real_size = int(storage_size(1.0_p3dfft_type) / 8)
        complex_size = 2 * real_size
        call c_get_mpi_info(me, np, cm)
        if (allocated(proc_id2coords)) deallocate(proc_id2coords, proc_coords2id, proc_dims, proc_parts)
        allocate(proc_id2coords(0:2*np-1), proc_coords2id(0:dims(1)-1, 0:dims(2)-1), proc_dims(2, 9, 0:np-1), proc_parts(np, 7))
        proc_parts = -1
        do i = 0, np - 1
          rc = c_proc_id2coords(i, ip, jp)
          proc_id2coords(2*i) = ip; proc_id2coords(2*i+1) = jp
          proc_coords2id(ip, jp) = i
          do k = 1, 2
            rc = c_proc_dims(k, i, d9)
            proc_dims(k, :, i) = d9
          end do
        end do
        call c_get_dims(d9(1:3), d9(4:6), d9(7:9), 1)
        padi = mem(3) - d9(9)                  ! memsize(3) = kjsize + padi (setup.F90:597-602)
      end subroutine

      subroutine p3dfft_get_mpi_info(mpi_taskid, mpi_tasks, mpi_comm)
        integer, intent(out) :: mpi_taskid, mpi_tasks, mpi_comm
        call c_get_mpi_info(mpi_taskid, mpi_tasks, mpi_comm)
      end subroutine

      subroutine p3dfft_ftran_r2c_1d(rXgYZ, cXgYZ)
        real(p3dfft_type), target :: rXgYZ(*), cXgYZ(*)
        call c_r2c_1d(c_loc(rXgYZ), c_loc(cXgYZ))
      end subroutine

      ! rbuf1 / rbuf2: the reference's caller-provided scratch (module.F90:1066); unused here
      subroutine rtran_x2y(source, dest, rbuf1, rbuf2, dstart, dend, dsize, t)
        real(p3dfft_type), target :: source(*), dest(*)
        real(p3dfft_type) :: rbuf1(*), rbuf2(*)
        integer, intent(out) :: dstart(3), dend(3), dsize(3)
        real(r8) :: t
        call c_rtran_x2y(c_loc(source), c_loc(dest), dstart, dend, dsize, t)
      end subroutine
      subroutine rtran_y2x(source, dest, rbuf1, rbuf2, dstart, dend, dsize, t)
        real(p3dfft_type), target :: source(*), dest(*)
        real(p3dfft_type) :: rbuf1(*), rbuf2(*)
        integer, intent(out) :: dstart(3), dend(3), dsize(3)
        real(r8) :: t
        call c_rtran_y2x(c_loc(source), c_loc(dest), dstart, dend, dsize, t)
      end subroutine
      subroutine rtran_x2z(source, dest, rbuf1, rbuf2, dstart, dend, dsize, t)
        real(p3dfft_type), target :: source(*), dest(*)
        real(p3dfft_type) :: rbuf1(*), rbuf2(*)
        integer, intent(out) :: dstart(3), dend(3), dsize(3)
        real(r8) :: t
        call c_rtran_x2z(c_loc(source), c_loc(dest), dstart, dend, dsize, t)
      end subroutine
      subroutine rtran_z2x(source, dest, rbuf1, rbuf2, dstart, dend, dsize, t)
        real(p3dfft_type), target :: source(*), dest(*)
        real(p3dfft_type) :: rbuf1(*), rbuf2(*)
        integer, intent(out) :: dstart(3), dend(3), dsize(3)
        real(r8) :: t
        call c_rtran_z2x(c_loc(source), c_loc(dest), dstart, dend, dsize, t)
      end subroutine

      ! result in the public proc_parts(P,7) like the reference (module.F90:888-1054)
      subroutine get_proc_parts(base_x, base_y, base_z, size_x, size_y, size_z, conf, ierr)
        integer, intent(in) :: base_x, base_y, base_z, size_x, size_y, size_z, conf
        integer, intent(out) :: ierr
        integer(c_int), allocatable :: flat(:)
        integer(c_int) :: n, np, p
        np = size(proc_parts, 1)
        allocate(flat(7*np))
        n = c_get_proc_parts(base_x, base_y, base_z, size_x, size_y, size_z, conf, flat, ierr)
        do p = 1, np
          proc_parts(p, :) = flat(7*(p-1)+1 : 7*p)        ! the C side stores one part per row
        end do
      end subroutine

      subroutine print_buf(A, lx, ly, lz)                 ! module.F90:747-763
        integer :: lx, ly, lz, x, y, z, me, np, cm
        complex(p3dfft_type) :: A(lx, ly, lz)
        call c_get_mpi_info(me, np, cm)
        do z = 1, lz
          do y = 1, ly
            do x = 1, lx
              if (abs(A(x,y,z)) .gt. 0.0000005) print *, me, ': (', x, y, z, ') =', A(x,y,z)
            end do
          end do
        end do
      end subroutine

      subroutine print_buf_real(A, lx, ly, lz)            ! module.F90:765-781
        integer :: lx, ly, lz, x, y, z, me, np, cm
        real(p3dfft_type) :: A(lx, ly, lz)
        call c_get_mpi_info(me, np, cm)
        do z = 1, lz
          do y = 1, ly
            do x = 1, lx
              if (abs(A(x,y,z)) .gt. 0.0000005) print *, me, ': (', x, y, z, ') =', A(x,y,z)
            end do
          end do
        end do
      end subroutine

      subroutine p3dfft_get_dims(istart,iend,isize,conf)
        integer :: istart(3), iend(3), isize(3), conf
        call c_get_dims(istart, iend, isize, conf)
      end subroutine

      ! Dummy types as in the reference (ftran.F90:494-499, btran.F90:401-406): the real-space array is real, the
      ! wavenumber array complex -- the drivers pass complex arrays (driver_rand.F90:61,219).  Callers that pass a real
      ! array posing as complex (in-place calls) go through the external wrappers behind the module (wrap.F90:82-150),
      ! which have an implicit interface.
      subroutine p3dfft_ftran_r2c(XgYZ,XYZg,op)
        real(p3dfft_type), target :: XgYZ(*)
        complex(p3dfft_type), target :: XYZg(*)
        character(len=3) :: op
        call c_ftran(c_loc(XgYZ), c_loc(XYZg), op//c_null_char)
        call c_get_timers(timers)
      end subroutine

      subroutine p3dfft_btran_c2r(XYZg,XgYZ,op)
        complex(p3dfft_type), target :: XYZg(*)
        real(p3dfft_type), target :: XgYZ(*)
        character(len=3) :: op
        call c_btran(c_loc(XYZg), c_loc(XgYZ), op//c_null_char)
        call c_get_timers(timers)
      end subroutine

      subroutine p3dfft_ftran_r2c_many(XgYZ,dim_in,XYZg,dim_out,nv,op)
        integer :: dim_in, dim_out, nv
        real(p3dfft_type), target :: XgYZ(*)
        complex(p3dfft_type), target :: XYZg(*)
        character(len=3) :: op
        call c_ftran_many(c_loc(XgYZ), dim_in, c_loc(XYZg), dim_out, nv, op//c_null_char)
        call c_get_timers(timers)
      end subroutine

      subroutine p3dfft_btran_c2r_many(XYZg,dim_in,XgYZ,dim_out,nv,op)
        integer :: dim_in, dim_out, nv
        complex(p3dfft_type), target :: XYZg(*)
        real(p3dfft_type), target :: XgYZ(*)
        character(len=3) :: op
        call c_btran_many(c_loc(XYZg), dim_in, c_loc(XgYZ), dim_out, nv, op//c_null_char)
        call c_get_timers(timers)
      end subroutine

      subroutine p3dfft_cheby(in,out,Lz)
        real(p3dfft_type), target :: in(*)
        complex(p3dfft_type), target :: out(*)
        real(p3dfft_type) :: Lz
        call c_cheby(c_loc(in), c_loc(out), Lz)
      end subroutine

      subroutine p3dfft_cheby_many(in,dim_in,out,dim_out,nv,Lz)
        integer :: dim_in, dim_out, nv
        real(p3dfft_type), target :: in(*)
        complex(p3dfft_type), target :: out(*)
        real(p3dfft_type) :: Lz
        call c_cheby_many(c_loc(in), dim_in, c_loc(out), dim_out, nv, Lz)
      end subroutine

      subroutine p3dfft_clean()
        call c_clean()
      end subroutine

      subroutine get_timers(t)
        real(r8) :: t(12)
        call c_get_timers(t)
      end subroutine

      subroutine set_timers()
        call c_set_timers()
        timers = 0.0d0
      end subroutine

      end module p3dfft

! ---- external wrappers of build/wrap.F90:82-150 (real arrays posing as complex, in-place calls) ----
! Not module procedures: callers reach them through an implicit interface, so any array type is accepted (that is what
! the reference's wrap.F90 is for).  They bind the C ABI themselves and refresh the module's public timers.
      subroutine ftran_r2c(IN,OUT,op)
        use iso_c_binding
        use p3dfft, only : p3dfft_type, timers, get_timers
        real(p3dfft_type), target :: IN(*), OUT(*)
        character(len=3) :: op
        interface
          subroutine c_ftran(a,b,op) bind(C,name='p3dfft_ftran_r2c')
            import :: c_ptr, c_char
            type(c_ptr), value :: a, b
            character(kind=c_char) :: op(*)
          end subroutine
        end interface
        call c_ftran(c_loc(IN), c_loc(OUT), op//c_null_char)
        call get_timers(timers)
      end subroutine

      subroutine btran_c2r(IN,OUT,op)
        use iso_c_binding
        use p3dfft, only : p3dfft_type, timers, get_timers
        real(p3dfft_type), target :: IN(*), OUT(*)
        character(len=3) :: op
        interface
          subroutine c_btran(a,b,op) bind(C,name='p3dfft_btran_c2r')
            import :: c_ptr, c_char
            type(c_ptr), value :: a, b
            character(kind=c_char) :: op(*)
          end subroutine
        end interface
        call c_btran(c_loc(IN), c_loc(OUT), op//c_null_char)
        call get_timers(timers)
      end subroutine

      subroutine ftran_r2c_many(IN,dim_in,OUT,dim_out,nv,op)
        use iso_c_binding
        use p3dfft, only : p3dfft_type, timers, get_timers
        integer :: dim_in, dim_out, nv
        real(p3dfft_type), target :: IN(*), OUT(*)
        character(len=3) :: op
        interface
          subroutine c_ftran_many(a,dim_in,b,dim_out,nv,op) bind(C,name='p3dfft_ftran_r2c_many')
            import :: c_ptr, c_char, c_int
            type(c_ptr), value :: a, b
            integer(c_int) :: dim_in, dim_out, nv
            character(kind=c_char) :: op(*)
          end subroutine
        end interface
        call c_ftran_many(c_loc(IN), dim_in, c_loc(OUT), dim_out, nv, op//c_null_char)
        call get_timers(timers)
      end subroutine

      subroutine btran_c2r_many(IN,dim_in,OUT,dim_out,nv,op)
        use iso_c_binding
        use p3dfft, only : p3dfft_type, timers, get_timers
        integer :: dim_in, dim_out, nv
        real(p3dfft_type), target :: IN(*), OUT(*)
        character(len=3) :: op
        interface
          subroutine c_btran_many(a,dim_in,b,dim_out,nv,op) bind(C,name='p3dfft_btran_c2r_many')
            import :: c_ptr, c_char, c_int
            type(c_ptr), value :: a, b
            integer(c_int) :: dim_in, dim_out, nv
            character(kind=c_char) :: op(*)
          end subroutine
        end interface
        call c_btran_many(c_loc(IN), dim_in, c_loc(OUT), dim_out, nv, op//c_null_char)
        call get_timers(timers)
      end subroutine
