!!! fortran/tridsol_pbx.f90
!
! Drop-in replacement for src/tridsol.f90 (public list :16-18): single-line calls forwarded to the
! batched GPU kernels with a batch of one.  For production use call the batch entry points
! (pbx_tdma_batch_device, ...) directly; this module exists so that the reference's tests/tridiag
! programs link and run unchanged.  Not compiled in this image (no Fortran compiler).
module tridsol

  use constants
  use pbx_iso_c

  implicit none

  private
  public :: tdma, tdma_periodic
  public :: fwd_sweep
  public :: bwd_sweep

contains

  subroutine chk(ierr)
    integer(c_int), intent(in) :: ierr
    if (ierr /= PBX_OK) then
       print *, "ERROR: libpbx returned ", ierr
       stop 2
    end if
  end subroutine chk

  subroutine tdma(a, b, c, d)
    real(pb_dp), dimension(:), intent(in), contiguous :: a      ! sub-diagonal
    real(pb_dp), dimension(:), intent(inout), contiguous :: b   ! DIAGONAL (overwritten with the pivots)
    real(pb_dp), dimension(:), intent(in), contiguous :: c      ! super-diagonal
    real(pb_dp), dimension(:), intent(inout), contiguous :: d   ! RHS/solution
    call chk(pbx_tdma_host(size(d), a, b, c, d))
  end subroutine tdma

  subroutine tdma_periodic(a, b, c, d)
    real(pb_dp), dimension(:), intent(in), contiguous :: a
    real(pb_dp), dimension(:), intent(inout), contiguous :: b   ! left untouched, as in the reference
    real(pb_dp), dimension(:), intent(in), contiguous :: c
    real(pb_dp), dimension(:), intent(inout), contiguous :: d
    call chk(pbx_tdma_periodic_host(size(a), a, b, c, d))
  end subroutine tdma_periodic

  subroutine fwd_sweep(a, b, c, d)
    real(pb_dp), dimension(:), intent(in), contiguous :: a
    real(pb_dp), dimension(:), intent(inout), contiguous :: b
    real(pb_dp), dimension(:), intent(in), contiguous :: c
    real(pb_dp), dimension(:), intent(inout), contiguous :: d
    call chk(pbx_fwd_sweep_host(size(d), a, b, c, d))
  end subroutine fwd_sweep

  subroutine bwd_sweep(b, c, d)
    real(pb_dp), dimension(:), intent(in), contiguous :: b
    real(pb_dp), dimension(:), intent(in), contiguous :: c
    real(pb_dp), dimension(:), intent(inout), contiguous :: d
    call chk(pbx_bwd_sweep_host(size(d), b, c, d))
  end subroutine bwd_sweep

end module tridsol
