!!! fortran/poissbox_matfree_pbx.f90
!
! Replacement bodies for the two routines of module `poissbox` (src/poissbox.f90) that sit on the
! hot path's boundary.  Everything else in src/poissbox.f90 -- the modules matfree_types and
! matfree (:9-69: `mat_ctx` and the explicit MatCreateShell / MatShellSetContext /
! MatShellGetContext interfaces), compute_lapl, initialise_grid, initialise_linear_system, solve --
! stays exactly as it is.  Signatures, dummy names and the printed lines are the reference's.
!
! How to apply: in src/poissbox.f90 add `use pbx_petsc_iso_c` to the two routines and replace
! their bodies by the ones below (a two-hunk patch: one added call in initialise_matrix_free
! (:261-263), one replaced call in mfmult (:316)).
!
! What does NOT change: the shell context is still the Fortran `mat_ctx` {da, grid_deltas}, so
! `MatShellGetContext(M, ctx, ierr)` in mfmult and in src/example.f90:201-233 (check_lapl, which
! calls compute_lapl_pointwise with ctx%da and ctx%grid_deltas) keep working; the device operator
! rides on the Mat as a composed PetscContainer and is freed by MatDestroy.
!
! NOTE: not compiled in this image (no Fortran compiler, no PETSc); tests/petsc_mock acts out this
! exact call sequence against the C layer on the GPU (tests/test_petsc_glue.py).

  subroutine initialise_matrix_free(ctx, P, A)
    !! Create a matrix free object

    use matfree_types
    use matfree
    use pbx_petsc_iso_c                                    ! new
    use, intrinsic :: iso_c_binding, only : c_null_ptr     ! new

    type(mat_ctx) :: ctx
    type(tMat), intent(in) :: P
    type(tMat), intent(out) :: A

    integer :: m, n

    integer :: ierr

    print *, "- Initialising matrix-free system"

    call MatGetLocalSize(P, m, n, ierr)

    call MatCreateShell(PETSC_COMM_WORLD, m, n, PETSC_DETERMINE, PETSC_DETERMINE, ctx, A, ierr)
    call MatShellSetContext(A, ctx, ierr) ! Is this necessary?
    ! new: the CUDA operator for this rank's brick of ctx%da (a z-slab DMDA:
    ! -da_processors_x 1 -da_processors_y 1), composed on A; A now hands out VECCUDA vectors.
    ! One rank: no communicator.  Several ranks: pass the ncclComm_t made with pbx_comm_unique_id
    ! (rank 0) + MPI_Bcast of its 128 bytes + pbx_comm_init_rank (every rank), INTEGRATION.md 3.
    ierr = PbxShellAttach(A%v, ctx%da%v, ctx%grid_deltas, c_null_ptr)
    if (ierr /= 0) then
       print *, "ERROR: PbxShellAttach returned ", ierr
       stop 2
    end if
    call MatShellSetOperation(A, MATOP_MULT, mfmult, ierr)

    print *, "- Done"

  end subroutine initialise_matrix_free

  subroutine mfmult(M, x, f, ierr)
    !! Computes the matrix vector product f = Mx, matrix-free

    use matfree_types
    use matfree
    use pbx_petsc_iso_c                                    ! new (instead of `use compute_lapl`)

    type(tMat) :: M ! The operator
    type(tVec) :: x ! The input vector
    type(tVec) :: f ! The output vector
    integer :: ierr ! Error status (0 indicates success)

    type(mat_ctx), pointer :: ctx

    call MatShellGetContext(M, ctx, ierr)                  ! unchanged: the Fortran context
    ! was: call compute_lapl_pointwise(ctx%da, ctx%grid_deltas, x, f)   (the 2nd-order star, on the host)
    ! now: the operator the handle is set to -- the compact Laplacian `lapl` of compact_schemes by
    ! default, the same 2nd-order star after pbx_set_operator(h, PBX_OPERATOR_STAR) -- on the device
    ierr = PbxShellMult(M%v, x%v, f%v)

  end subroutine mfmult
