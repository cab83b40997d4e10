!!! fortran/compact_schemes_pbx.f90
!
! Drop-in replacement for src/compact_schemes.f90: same module name, same public list
! (src/compact_schemes.f90:9-13), same dummy arguments, bodies forwarded to libpbx.so through
! pbx_iso_c.  Assumed-shape dummies are not C-interoperable, so each body takes the extents with
! size(), makes sure the actual arguments are contiguous (the `contiguous` copies gfortran would
! make anyway for the strided sections the reference passes around) and calls the host-pointer
! C entry point.  A length mismatch still ends in `stop 7` (src/compact_schemes.f90:177-180).
!
! Not compiled in this image (no Fortran compiler); see INTEGRATION.md.
module compact_schemes

  use constants
  use pbx_iso_c

  implicit none

  private
  public :: grad, grad_1d
  public :: interp, interp_1d
  public :: div, div_1d
  public :: interp_div, interp_1d_div
  public :: lapl

  integer(c_int), save :: lapl_mode = PBX_MODE_FAST   ! set to PBX_MODE_REFERENCE for bit-faithful runs

contains

  subroutine check(ierr)
    integer(c_int), intent(in) :: ierr
    if (ierr == PBX_ERR_SIZE) then
       print *, "ERROR: periodic gradient is same length as field!"
       stop 7
    else if (ierr /= PBX_OK) then
       print *, "ERROR: libpbx returned ", ierr
       stop 2
    end if
  end subroutine check

  subroutine lapl(f, dx, d2fdx2)
    real(pb_dp), dimension(:, :, :), intent(in), contiguous :: f
    real(pb_dp), dimension(3), intent(in) :: dx
    real(pb_dp), dimension(:, :, :), intent(out), contiguous :: d2fdx2
    call check(pbx_lapl_host(size(f, 1), size(f, 2), size(f, 3), f, dx, d2fdx2, lapl_mode))
  end subroutine lapl

  subroutine grad(f, dx, df)
    real(pb_dp), dimension(:, :, :), intent(in), contiguous :: f
    real(pb_dp), dimension(3), intent(in) :: dx
    real(pb_dp), dimension(:, :, :, :), intent(out), contiguous :: df
    call check(pbx_grad_host(size(f, 1), size(f, 2), size(f, 3), f, dx, df))
  end subroutine grad

  subroutine div(f, dx, df)
    real(pb_dp), dimension(:, :, :, :), intent(in), contiguous :: f
    real(pb_dp), dimension(3), intent(in) :: dx
    real(pb_dp), dimension(:, :, :), intent(out), contiguous :: df
    call check(pbx_div_host(size(f, 1), size(f, 2), size(f, 3), f, dx, df))
  end subroutine div

  subroutine interp(f, fi, opt_stagger)
    real(pb_dp), dimension(:, :, :), intent(in), contiguous :: f
    real(pb_dp), dimension(:, :, :), intent(out), contiguous :: fi
    integer, intent(in), optional :: opt_stagger
    integer(c_int) :: stagger
    stagger = -1
    if (present(opt_stagger)) stagger = opt_stagger
    call check(pbx_interp_host(size(f, 1), size(f, 2), size(f, 3), f, fi, stagger))
  end subroutine interp

  subroutine interp_div(f, fi)
    real(pb_dp), dimension(:, :, :), intent(in), contiguous :: f
    real(pb_dp), dimension(:, :, :), intent(out), contiguous :: fi
    call interp(f, fi, +1)
  end subroutine interp_div

  subroutine grad_1d(f, dx, df, opt_stagger)
    real(pb_dp), dimension(:), intent(in), contiguous :: f
    real(pb_dp), intent(in) :: dx
    real(pb_dp), dimension(:), intent(out), contiguous :: df
    integer, intent(in), optional :: opt_stagger
    integer(c_int) :: stagger
    stagger = -1
    if (present(opt_stagger)) stagger = opt_stagger
    call check(pbx_grad_1d_host(size(f), f, dx, size(df), df, stagger))
  end subroutine grad_1d

  subroutine div_1d(f, dx, df)
    real(pb_dp), dimension(:), intent(in), contiguous :: f
    real(pb_dp), intent(in) :: dx
    real(pb_dp), dimension(:), intent(out), contiguous :: df
    call grad_1d(f, dx, df, +1)
  end subroutine div_1d

  subroutine interp_1d(f, fi, opt_stagger)
    real(pb_dp), dimension(:), intent(in), contiguous :: f
    real(pb_dp), dimension(:), intent(out), contiguous :: fi
    integer, intent(in), optional :: opt_stagger
    integer(c_int) :: stagger
    stagger = -1
    if (present(opt_stagger)) stagger = opt_stagger
    call check(pbx_interp_1d_host(size(f), f, size(fi), fi, stagger))
  end subroutine interp_1d

  subroutine interp_1d_div(f, fi)
    real(pb_dp), dimension(:), intent(in), contiguous :: f
    real(pb_dp), dimension(:), intent(out), contiguous :: fi
    call interp_1d(f, fi, +1)
  end subroutine interp_1d_div

end module compact_schemes
