!!! fortran/pbx_petsc_iso_c.f90
!
! ISO_C_BINDING interfaces to the Fortran-callable layer of petsc/pbx_matshell.c: the MATSHELL
! registration and MatMult of the reference (src/poissbox.f90:242-267, 300-322) re-pointed at the
! CUDA operator, with the Fortran derived type `mat_ctx` (src/poissbox.f90:17-20) left in place
! as the shell context.
!
! A PETSc Fortran object (type(tMat), type(tVec), type(tDM)) is a derived type whose only
! component `v` holds the C object's address (PetscFortranAddr = integer(c_intptr_t) on 64-bit
! systems); PETSc's own Fortran stubs receive such objects by reference, i.e. as `Mat *`.  The
! interfaces below do the same with the `v` component, which is C-interoperable where the derived
! type itself is not: call them as  PbxShellMult(M%v, x%v, f%v).
!
! NOTE: this image has neither a Fortran compiler nor PETSc; the file has not been compiled here.
! The C side of every interface is compiled and run on the GPU through tests/petsc_mock, whose
! driver acts out this calling convention (objects by reference, Fortran mat_ctx as context).
module pbx_petsc_iso_c

  use, intrinsic :: iso_c_binding

  implicit none

  interface

     ! after MatCreateShell + MatShellSetContext: create the device operator for this rank's z-slab
     ! brick of `da`, compose it on A (freed by MatDestroy), make A hand out VECCUDA vectors.
     ! nccl_comm: c_null_ptr on one rank, else an ncclComm_t over PETSC_COMM_WORLD in rank order.
     integer(c_int) function PbxShellAttach(A, da, deltas, nccl_comm) bind(C, name="PbxShellAttach")
       import :: c_int, c_intptr_t, c_double, c_ptr
       integer(c_intptr_t), intent(in) :: A, da      ! A%v, ctx%da%v
       real(c_double), intent(in) :: deltas(3)       ! ctx%grid_deltas
       type(c_ptr), value :: nccl_comm
     end function PbxShellAttach

     ! f = A x on the device (VECCUDA arrays, PETSc's current stream): the body of mfmult
     integer(c_int) function PbxShellMult(M, x, f) bind(C, name="PbxShellMult")
       import :: c_int, c_intptr_t
       integer(c_intptr_t), intent(in) :: M, x, f    ! M%v, x%v, f%v
     end function PbxShellMult

     ! the pbx handle composed on A (for pbx_set_operator / pbx_set_mode / pbx_set_pc)
     integer(c_int) function PbxShellGetHandle(A, h) bind(C, name="PbxShellGetHandle")
       import :: c_int, c_intptr_t, c_ptr
       integer(c_intptr_t), intent(in) :: A
       type(c_ptr), intent(out) :: h
     end function PbxShellGetHandle

     ! optional replacement of KSPSolve (src/poissbox.f90:296) by the library's fused device CG
     integer(c_int) function PbxShellSolveCG(A, b, x, rtol, maxit, its, reason) bind(C, name="PbxShellSolveCG")
       import :: c_int, c_intptr_t, c_double
       integer(c_intptr_t), intent(in) :: A, b, x
       real(c_double), value :: rtol
       integer(c_int), value :: maxit
       integer(c_int), intent(out) :: its, reason
     end function PbxShellSolveCG

  end interface

end module pbx_petsc_iso_c
