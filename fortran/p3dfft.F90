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
! The external (non-module) wrappers of build/wrap.F90:82-150 follow the module.
!
! NOT COMPILED IN THE BUILD IMAGE (no Fortran compiler there); shipped as source because a .mod
! file is compiler specific.  Build with  -DSINGLE_PREC  to bind libp3dfft_single.so.
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

      public :: p3dfft_setup, p3dfft_get_dims, p3dfft_ftran_r2c, p3dfft_btran_c2r, &
                p3dfft_ftran_r2c_many, p3dfft_btran_c2r_many, p3dfft_cheby, p3dfft_cheby_many, &
                p3dfft_clean, get_timers, set_timers

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
      end interface

      contains

      subroutine p3dfft_setup(dims,nx,ny,nz,mpi_comm_in,nxcut,nycut,nzcut,overwrite,memsize)
        integer :: dims(2), nx, ny, nz, mpi_comm_in
        integer, optional, intent(in) :: nxcut, nycut, nzcut
        logical, optional, intent(in) :: overwrite
        integer, optional, intent(out) :: memsize(3)
        integer(c_int) :: nxc, nyc, nzc, ow, mem(3)
        nxc = nx; nyc = ny; nzc = nz; ow = 1
        if (present(nxcut)) nxc = nxcut
        if (present(nycut)) nyc = nycut
        if (present(nzcut)) nzc = nzcut
        if (present(overwrite)) then
          if (.not. overwrite) ow = 0
        end if
        call c_setup(dims, nx, ny, nz, mpi_comm_in, nxc, nyc, nzc, ow, mem)
        if (present(memsize)) memsize = mem
      end subroutine

      subroutine p3dfft_get_dims(istart,iend,isize,conf)
        integer :: istart(3), iend(3), isize(3), conf
        call c_get_dims(istart, iend, isize, conf)
      end subroutine

      ! assumed-size arguments: the reference's explicit-shape dummies (ftran.F90:494-499) accept any
      ! contiguous actual argument, including real arrays posing as complex (in-place calls)
      subroutine p3dfft_ftran_r2c(XgYZ,XYZg,op)
        real(p3dfft_type), target :: XgYZ(*)
        real(p3dfft_type), target :: XYZg(*)
        character(len=3) :: op
        call c_ftran(c_loc(XgYZ), c_loc(XYZg), op//c_null_char)
        call c_get_timers(timers)
      end subroutine

      subroutine p3dfft_btran_c2r(XYZg,XgYZ,op)
        real(p3dfft_type), target :: XYZg(*)
        real(p3dfft_type), target :: XgYZ(*)
        character(len=3) :: op
        call c_btran(c_loc(XYZg), c_loc(XgYZ), op//c_null_char)
        call c_get_timers(timers)
      end subroutine

      subroutine p3dfft_ftran_r2c_many(XgYZ,dim_in,XYZg,dim_out,nv,op)
        integer :: dim_in, dim_out, nv
        real(p3dfft_type), target :: XgYZ(*), XYZg(*)
        character(len=3) :: op
        call c_ftran_many(c_loc(XgYZ), dim_in, c_loc(XYZg), dim_out, nv, op//c_null_char)
        call c_get_timers(timers)
      end subroutine

      subroutine p3dfft_btran_c2r_many(XYZg,dim_in,XgYZ,dim_out,nv,op)
        integer :: dim_in, dim_out, nv
        real(p3dfft_type), target :: XYZg(*), XgYZ(*)
        character(len=3) :: op
        call c_btran_many(c_loc(XYZg), dim_in, c_loc(XgYZ), dim_out, nv, op//c_null_char)
        call c_get_timers(timers)
      end subroutine

      subroutine p3dfft_cheby(in,out,Lz)
        real(p3dfft_type), target :: in(*), out(*)
        real(p3dfft_type) :: Lz
        call c_cheby(c_loc(in), c_loc(out), Lz)
      end subroutine

      subroutine p3dfft_cheby_many(in,dim_in,out,dim_out,nv,Lz)
        integer :: dim_in, dim_out, nv
        real(p3dfft_type), target :: in(*), out(*)
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
      subroutine ftran_r2c(IN,OUT,op)
        use p3dfft
        real(p3dfft_type) :: IN(*), OUT(*)
        character(len=3) :: op
        call p3dfft_ftran_r2c(IN, OUT, op)
      end subroutine

      subroutine btran_c2r(IN,OUT,op)
        use p3dfft
        real(p3dfft_type) :: IN(*), OUT(*)
        character(len=3) :: op
        call p3dfft_btran_c2r(IN, OUT, op)
      end subroutine

      subroutine ftran_r2c_many(IN,dim_in,OUT,dim_out,nv,op)
        use p3dfft
        integer :: dim_in, dim_out, nv
        real(p3dfft_type) :: IN(*), OUT(*)
        character(len=3) :: op
        call p3dfft_ftran_r2c_many(IN, dim_in, OUT, dim_out, nv, op)
      end subroutine

      subroutine btran_c2r_many(IN,dim_in,OUT,dim_out,nv,op)
        use p3dfft
        integer :: dim_in, dim_out, nv
        real(p3dfft_type) :: IN(*), OUT(*)
        character(len=3) :: op
        call p3dfft_btran_c2r_many(IN, dim_in, OUT, dim_out, nv, op)
      end subroutine
