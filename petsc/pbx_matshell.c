/*
 * petsc/pbx_matshell.c -- PETSc glue: the compact Laplacian as a MATSHELL MatMult over
 * device-resident VECCUDA vectors, so that `-ksp_type cg` runs entirely on the GPU.
 *
 * Replaces the body of `mfmult` (src/poissbox.f90:300-322), which today calls the 2nd-order
 * pointwise star, and extends `mat_ctx` (src/poissbox.f90:17-20) with the pbx handle.
 * Registration mirrors initialise_matrix_free (src/poissbox.f90:242-267).
 *
 * NOT COMPILED HERE: this image has neither PETSc nor MPI.  The file is intentionally small
 * and only moves pointers; every call below the PETSc API is a tested entry point of
 * include/pbx.h.  Requirements on the caller: DMDA z-slab layout
 * (-da_processors_x 1 -da_processors_y 1) so that each rank's block is one contiguous
 * nx x ny x nz_local brick in f(i,j,k) order, and -dm_vec_type cuda.
 */
#include <petscdmda.h>
#include <petscksp.h>
#include <petscdevice_cuda.h>

#include "pbx.h"

typedef struct {
    DM da;                   /* as in mat_ctx, src/poissbox.f90:18 */
    PetscReal grid_deltas[3]; /* :19 */
    pbx_handle h;            /* new: the device operator */
} PbxMatCtx;

/* MatMult callback: f = A x on the device (the counterpart of mfmult, src/poissbox.f90:300) */
static PetscErrorCode PbxMatMult(Mat M, Vec x, Vec f)
{
    PbxMatCtx *ctx;
    const PetscScalar *px;
    PetscScalar *pf;

    PetscFunctionBeginUser;
    PetscCall(MatShellGetContext(M, &ctx));
    PetscCall(VecCUDAGetArrayRead(x, &px));
    PetscCall(VecCUDAGetArrayWrite(f, &pf));
    /* the handle's operator: the compact Laplacian (default) or, after
     * pbx_set_operator(h, PBX_OPERATOR_STAR), the 2nd-order star mfmult applies today */
    PetscCheck(pbx_matmult_device(ctx->h, (const double *)px, (double *)pf) == PBX_OK, PETSC_COMM_SELF,
               PETSC_ERR_LIB, "pbx_matmult_device: %s", pbx_last_error());
    PetscCall(VecCUDARestoreArrayWrite(f, &pf));
    PetscCall(VecCUDARestoreArrayRead(x, &px));
    PetscFunctionReturn(PETSC_SUCCESS);
}

/* counterpart of initialise_matrix_free (src/poissbox.f90:242-267).  nccl_comm: NULL on one
 * rank, else an ncclComm_t spanning the ranks of PETSC_COMM_WORLD in rank order (z-slabs). */
PetscErrorCode PbxCreateShell(DM da, const PetscReal deltas[3], void *nccl_comm, Mat P, Mat *A)
{
    PbxMatCtx *ctx;
    PetscInt m, n, xs, ys, zs, xm, ym, zm, M, N, Q;
    int device = 0;
    double dx[3];

    PetscFunctionBeginUser;
    PetscCall(PetscNew(&ctx));
    ctx->da = da;
    for (int d = 0; d < 3; ++d) dx[d] = ctx->grid_deltas[d] = deltas[d];
    PetscCall(DMDAGetInfo(da, NULL, &M, &N, &Q, NULL, NULL, NULL, NULL, NULL, NULL, NULL, NULL, NULL));
    PetscCall(DMDAGetCorners(da, &xs, &ys, &zs, &xm, &ym, &zm));
    PetscCheck(xm == M && ym == N, PETSC_COMM_WORLD, PETSC_ERR_SUP,
               "pbx needs a z-slab DMDA: run with -da_processors_x 1 -da_processors_y 1");
    PetscCallCUDA(cudaGetDevice(&device));
    PetscCheck(pbx_create((int)xm, (int)ym, (int)zm, dx, device, nccl_comm, &ctx->h) == PBX_OK,
               PETSC_COMM_SELF, PETSC_ERR_LIB, "pbx_create: %s", pbx_last_error());
    PetscCall(MatGetLocalSize(P, &m, &n));                                   /* :259 */
    PetscCall(MatCreateShell(PETSC_COMM_WORLD, m, n, PETSC_DETERMINE, PETSC_DETERMINE, ctx, A)); /* :261 */
    PetscCall(MatShellSetVecType(*A, VECCUDA));
    PetscCall(MatShellSetOperation(*A, MATOP_MULT, (void (*)(void))PbxMatMult)); /* :263 */
    PetscFunctionReturn(PETSC_SUCCESS);
}

/* Optional: bypass KSP and run the library's own fused device CG with the same semantics
 * (-ksp_type cg -pc_type none, constant null space, preconditioned norm). */
PetscErrorCode PbxSolveCG(Mat A, Vec b, Vec x, PetscReal rtol, PetscInt maxit, PetscInt *its,
                          KSPConvergedReason *reason)
{
    PbxMatCtx *ctx;
    const PetscScalar *pb;
    PetscScalar *px;
    int it = 0, why = 0;
    double rnorm = 0;

    PetscFunctionBeginUser;
    PetscCall(MatShellGetContext(A, &ctx));
    PetscCall(VecCUDAGetArrayRead(b, &pb));
    PetscCall(VecCUDAGetArrayWrite(x, &px));
    PetscCheck(pbx_cg_solve_device(ctx->h, (const double *)pb, (double *)px, rtol, 1e-50, (int)maxit,
                                   &it, &rnorm, &why, NULL, 0) == PBX_OK,
               PETSC_COMM_SELF, PETSC_ERR_LIB, "pbx_cg_solve_device: %s", pbx_last_error());
    PetscCall(VecCUDARestoreArrayWrite(x, &px));
    PetscCall(VecCUDARestoreArrayRead(b, &pb));
    if (its) *its = it;
    if (reason) *reason = (KSPConvergedReason)why;
    PetscFunctionReturn(PETSC_SUCCESS);
}
