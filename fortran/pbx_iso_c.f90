!!! fortran/pbx_iso_c.f90
!
! ISO_C_BINDING interface to libpbx.so (include/pbx.h): the thin layer through which the
! reference's Fortran host code reaches the CUDA implementation.  One `bind(C)` interface per
! C entry point the Fortran side uses; the replacement module bodies (compact_schemes_pbx.f90,
! tridsol_pbx.f90) and the re-pointed MATSHELL callback (INTEGRATION.md) are written on top.
!
! NOTE: this image has no Fortran compiler, so this file has not been compiled here; it is kept
! deliberately mechanical (argument-for-argument copies of the C prototypes).
module pbx_iso_c

  use, intrinsic :: iso_c_binding

  implicit none

  integer(c_int), parameter :: PBX_OK = 0
  integer(c_int), parameter :: PBX_ERR_SIZE = 7        ! the reference's `stop 7`
  integer(c_int), parameter :: PBX_MODE_FAST = 0
  integer(c_int), parameter :: PBX_MODE_REFERENCE = 1

  interface

     ! --- lifecycle (include/pbx.h: pbx_create, pbx_destroy, pbx_set_mode, pbx_set_stream)
     integer(c_int) function pbx_create(nx, ny, nz, dx, device, nccl_comm, h) bind(C, name="pbx_create")
       import :: c_int, c_double, c_ptr
       integer(c_int), value :: nx, ny, nz
       real(c_double), intent(in) :: dx(3)
       integer(c_int), value :: device
       type(c_ptr), value :: nccl_comm     ! ncclComm_t or c_null_ptr
       type(c_ptr), intent(out) :: h
     end function pbx_create

     integer(c_int) function pbx_destroy(h) bind(C, name="pbx_destroy")
       import :: c_int, c_ptr
       type(c_ptr), value :: h
     end function pbx_destroy

     integer(c_int) function pbx_set_mode(h, mode) bind(C, name="pbx_set_mode")
       import :: c_int, c_ptr
       type(c_ptr), value :: h
       integer(c_int), value :: mode
     end function pbx_set_mode

     integer(c_int) function pbx_set_stream(h, stream) bind(C, name="pbx_set_stream")
       import :: c_int, c_ptr
       type(c_ptr), value :: h, stream
     end function pbx_set_stream

     ! --- device-pointer operators (used by the MATSHELL MatMult)
     integer(c_int) function pbx_lapl_device(h, f, d2f) bind(C, name="pbx_lapl_device")
       import :: c_int, c_ptr
       type(c_ptr), value :: h, f, d2f      ! device pointers (VecCUDAGetArrayRead / Write)
     end function pbx_lapl_device

     integer(c_int) function pbx_cg_solve_device(h, b, x, rtol, abstol, maxit, its, rnorm, reason, hist, nhist) &
          bind(C, name="pbx_cg_solve_device")
       import :: c_int, c_double, c_ptr
       type(c_ptr), value :: h, b, x
       real(c_double), value :: rtol, abstol
       integer(c_int), value :: maxit
       integer(c_int), intent(out) :: its, reason
       real(c_double), intent(out) :: rnorm
       type(c_ptr), value :: hist
       integer(c_int), value :: nhist
     end function pbx_cg_solve_device

     ! --- host-pointer variants: what the module bodies call
     integer(c_int) function pbx_lapl_host(nx, ny, nz, f, dx, d2f, mode) bind(C, name="pbx_lapl_host")
       import :: c_int, c_double
       integer(c_int), value :: nx, ny, nz, mode
       real(c_double), intent(in) :: f(*), dx(3)
       real(c_double), intent(out) :: d2f(*)
     end function pbx_lapl_host

     integer(c_int) function pbx_grad_host(nx, ny, nz, f, dx, df) bind(C, name="pbx_grad_host")
       import :: c_int, c_double
       integer(c_int), value :: nx, ny, nz
       real(c_double), intent(in) :: f(*), dx(3)
       real(c_double), intent(out) :: df(*)
     end function pbx_grad_host

     integer(c_int) function pbx_div_host(nx, ny, nz, f, dx, df) bind(C, name="pbx_div_host")
       import :: c_int, c_double
       integer(c_int), value :: nx, ny, nz
       real(c_double), intent(in) :: f(*), dx(3)
       real(c_double), intent(out) :: df(*)
     end function pbx_div_host

     integer(c_int) function pbx_interp_host(nx, ny, nz, f, fi, stagger) bind(C, name="pbx_interp_host")
       import :: c_int, c_double
       integer(c_int), value :: nx, ny, nz, stagger
       real(c_double), intent(in) :: f(*)
       real(c_double), intent(out) :: fi(*)
     end function pbx_interp_host

     integer(c_int) function pbx_grad_1d_host(nf, f, dx, ndf, df, stagger) bind(C, name="pbx_grad_1d_host")
       import :: c_int, c_double
       integer(c_int), value :: nf, ndf, stagger
       real(c_double), value :: dx
       real(c_double), intent(in) :: f(*)
       real(c_double), intent(out) :: df(*)
     end function pbx_grad_1d_host

     integer(c_int) function pbx_interp_1d_host(nf, f, nfi, fi, stagger) bind(C, name="pbx_interp_1d_host")
       import :: c_int, c_double
       integer(c_int), value :: nf, nfi, stagger
       real(c_double), intent(in) :: f(*)
       real(c_double), intent(out) :: fi(*)
     end function pbx_interp_1d_host

     integer(c_int) function pbx_tdma_host(n, a, b, c, d) bind(C, name="pbx_tdma_host")
       import :: c_int, c_double
       integer(c_int), value :: n
       real(c_double), intent(in) :: a(*), c(*)
       real(c_double), intent(inout) :: b(*), d(*)
     end function pbx_tdma_host

     integer(c_int) function pbx_tdma_periodic_host(n, a, b, c, d) bind(C, name="pbx_tdma_periodic_host")
       import :: c_int, c_double
       integer(c_int), value :: n
       real(c_double), intent(in) :: a(*), b(*), c(*)
       real(c_double), intent(inout) :: d(*)
     end function pbx_tdma_periodic_host

     integer(c_int) function pbx_fwd_sweep_host(n, a, b, c, d) bind(C, name="pbx_fwd_sweep_host")
       import :: c_int, c_double
       integer(c_int), value :: n
       real(c_double), intent(in) :: a(*), c(*)
       real(c_double), intent(inout) :: b(*), d(*)
     end function pbx_fwd_sweep_host

     integer(c_int) function pbx_bwd_sweep_host(n, b, c, d) bind(C, name="pbx_bwd_sweep_host")
       import :: c_int, c_double
       integer(c_int), value :: n
       real(c_double), intent(in) :: b(*), c(*)
       real(c_double), intent(inout) :: d(*)
     end function pbx_bwd_sweep_host

  end interface

end module pbx_iso_c
