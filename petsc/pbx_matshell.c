/*
 * petsc/pbx_matshell.c -- PETSc glue: the compact Laplacian as a MATSHELL MatMult over
 * device-resident VECCUDA vectors, so that `-ksp_type cg` runs entirely on the GPU.
 *
 * The reference registers the shell matrix in Fortran (initialise_matrix_free,
 * src/poissbox.f90:242-267), with the Fortran derived type `mat_ctx` {da, grid_deltas}
 * (src/poissbox.f90:17-20) as the shell CONTEXT, and other reference code fetches that context
 * with MatShellGetContext (mfmult :313-315, src/example.f90:201-233).  The glue therefore leaves the shell
 * context alone: the device operator (a pbx handle) rides on the Mat as a composed PetscContainer
 * ("pbx_handle"), destroyed with the matrix.  Two layers:
 *
 *   Fortran-callable (bind(C) interfaces in fortran/pbx_petsc_iso_c.f90; PETSc's Fortran objects
 *   are passed by reference, i.e. arrive here as Mat*, Vec*, DM*, exactly as in PETSc's own Fortran stubs):
 *     PbxShellAttach(A, da, deltas, nccl_comm)  after MatCreateShell/MatShellSetContext: create the handle
 *                                               for this rank's brick, compose it on A, make A hand out
 *                                               VECCUDA vectors
 *     PbxShellMult(M, x, f)                     the body of mfmult (src/poissbox.f90:300-322): f = A x on the device
 *     PbxShellSolveCG(A, b, x, ...)             optional: the library's own fused device CG instead of KSPSolve
 *   C hosts:
 *     PbxCreateShell(da, deltas, comm, P, &A)   initialise_matrix_free for a C caller (context = PbxMatCtx,
 *                                               layout-compatible with mat_ctx: {da, grid_deltas})
 *
 * NOT COMPILED AGAINST PETSc HERE: this image has neither PETSc nor MPI.  The file only moves
 * pointers; every call below the PETSc API is a tested entry point of include/pbx.h, and the file is
 * compiled (-Wall -Werror) and run on the GPU against a mock of the PETSc calls it makes
 * (tests/petsc_mock).  Requirements on the caller: DMDA z-slab layout
 * (-da_processors_x 1 -da_processors_y 1) so that each rank's block is one contiguous
 * nx x ny x nz_local brick in f(i,j,k) order, and -dm_vec_type cuda.
 */
#include <petscdmda.h>
#include <petscksp.h>
#include <petscdevice_cuda.h>

#include "pbx.h"

#define PBX_CONTAINER_KEY "pbx_handle"

typedef struct {
    DM da;                    /* as in mat_ctx, src/poissbox.f90:18 */
    PetscReal grid_deltas[3]; /* :19 */
} PbxMatCtx;

/* PetscContainer destructor: the handle dies with the matrix (MatDestroy drops the composed objects) */
static PetscErrorCode PbxHandleDestroy(void *p)
{
    PetscFunctionBeginUser;
    if (p) (void)pbx_destroy((pbx_handle)p);
    PetscFunctionReturn(PETSC_SUCCESS);
}

static PetscErrorCode PbxGetHandle(Mat A, pbx_handle *h)
{
    PetscContainer c = NULL;
    void *p = NULL;

    PetscFunctionBeginUser;
    PetscCall(PetscObjectQuery((PetscObject)A, PBX_CONTAINER_KEY, (PetscObject *)&c));
    PetscCheck(c != NULL, PETSC_COMM_SELF, PETSC_ERR_ARG_WRONGSTATE, "this Mat carries no pbx handle (PbxShellAttach was not called)");
    PetscCall(PetscContainerGetPointer(c, &p));
    *h = (pbx_handle)p;
    PetscFunctionReturn(PETSC_SUCCESS);
}

/* The handle launches on the stream PETSc's current device context uses, so that the MatMult is ordered
 * with PETSc's own vector kernels also when PETSc runs on a non-blocking stream
 * (-device_context_stream_type nonblocking / default_blocking). */
static PetscErrorCode PbxFollowPetscStream(pbx_handle h)
{
    PetscDeviceContext dctx;
    void *sh = NULL;

    PetscFunctionBeginUser;
    PetscCall(PetscDeviceContextGetCurrentContext(&dctx));
    PetscCall(PetscDeviceContextGetStreamHandle(dctx, &sh));
    PetscCheck(pbx_set_stream(h, sh ? *(void **)sh : NULL) == PBX_OK, PETSC_COMM_SELF, PETSC_ERR_LIB, "pbx_set_stream: %s",
               pbx_last_error());
    PetscFunctionReturn(PETSC_SUCCESS);
}

static PetscErrorCode PbxAttach(Mat A, DM da, const PetscReal deltas[3], void *nccl_comm)
{
    PetscInt xs, ys, zs, xm, ym, zm, M, N, Q;
    PetscContainer c;
    pbx_handle h = NULL;
    int device = 0;
    double dx[3];

    PetscFunctionBeginUser;
    for (int d = 0; d < 3; ++d) dx[d] = deltas[d];
    PetscCall(DMDAGetInfo(da, NULL, &M, &N, &Q, NULL, NULL, NULL, NULL, NULL, NULL, NULL, NULL, NULL));
    PetscCall(DMDAGetCorners(da, &xs, &ys, &zs, &xm, &ym, &zm));
    PetscCheck(xm == M && ym == N, PETSC_COMM_WORLD, PETSC_ERR_SUP,
               "pbx needs a z-slab DMDA: run with -da_processors_x 1 -da_processors_y 1");
    PetscCallCUDA(cudaGetDevice(&device));
    PetscCheck(pbx_create((int)xm, (int)ym, (int)zm, dx, device, nccl_comm, &h) == PBX_OK, PETSC_COMM_SELF,
               PETSC_ERR_LIB, "pbx_create: %s", pbx_last_error());
    PetscCall(PetscContainerCreate(PETSC_COMM_SELF, &c));
    PetscCall(PetscContainerSetPointer(c, h));
    PetscCall(PetscContainerSetUserDestroy(c, PbxHandleDestroy));
    PetscCall(PetscObjectCompose((PetscObject)A, PBX_CONTAINER_KEY, (PetscObject)c));
    PetscCall(PetscContainerDestroy(&c)); /* A holds the remaining reference */
    PetscCall(MatShellSetVecType(A, VECCUDA));
    PetscFunctionReturn(PETSC_SUCCESS);
}

/* f = A x on the device for the operator the handle is set to: the compact Laplacian (default) or,
 * after pbx_set_operator(h, PBX_OPERATOR_STAR), the 2nd-order star mfmult applies today */
static PetscErrorCode PbxMult(Mat M, Vec x, Vec f)
{
    pbx_handle h;
    const PetscScalar *px;
    PetscScalar *pf;

    PetscFunctionBeginUser;
    PetscCall(PbxGetHandle(M, &h));
    PetscCall(PbxFollowPetscStream(h));
    PetscCall(VecCUDAGetArrayRead(x, &px));
    PetscCall(VecCUDAGetArrayWrite(f, &pf));
    PetscCheck(pbx_matmult_device(h, (const double *)px, (double *)pf) == PBX_OK, PETSC_COMM_SELF, PETSC_ERR_LIB,
               "pbx_matmult_device: %s", pbx_last_error());
    PetscCall(VecCUDARestoreArrayWrite(f, &pf));
    PetscCall(VecCUDARestoreArrayRead(x, &px));
    PetscFunctionReturn(PETSC_SUCCESS);
}

/* ---- Fortran-callable layer (objects by reference; int return = ierr) -------------------------- */

/* called by the replacement body of initialise_matrix_free (fortran/poissbox_matfree_pbx.f90) right
 * after MatCreateShell + MatShellSetContext (src/poissbox.f90:261-262): the Fortran mat_ctx stays
 * the shell context.  nccl_comm: NULL on one rank, else an ncclComm_t over the ranks of
 * PETSC_COMM_WORLD in rank order (z-slabs). */
int PbxShellAttach(Mat *A, DM *da, const double deltas[3], void *nccl_comm)
{
    PetscReal d[3] = {deltas[0], deltas[1], deltas[2]};
    return (int)PbxAttach(*A, *da, d, nccl_comm);
}

/* the body of mfmult (src/poissbox.f90:300-322) */
int PbxShellMult(Mat *M, Vec *x, Vec *f) { return (int)PbxMult(*M, *x, *f); }

/* the handle itself, e.g. for pbx_set_operator / pbx_set_mode / pbx_set_pc from the host code */
int PbxShellGetHandle(Mat *A, pbx_handle *h) { return (int)PbxGetHandle(*A, h); }

/* Optional: bypass KSPSolve (src/poissbox.f90:296) and run the library's own fused device CG with the
 * same semantics (-ksp_type cg, constant null space, preconditioned norm). */
static PetscErrorCode PbxSolve(Mat A, Vec b, Vec x, PetscReal rtol, PetscInt maxit, PetscInt *its,
                               KSPConvergedReason *reason)
{
    pbx_handle h;
    const PetscScalar *pb;
    PetscScalar *px;
    int it = 0, why = 0;
    double rnorm = 0;

    PetscFunctionBeginUser;
    PetscCall(PbxGetHandle(A, &h));
    PetscCall(PbxFollowPetscStream(h));
    PetscCall(VecCUDAGetArrayRead(b, &pb));
    PetscCall(VecCUDAGetArrayWrite(x, &px));
    PetscCheck(pbx_cg_solve_device(h, (const double *)pb, (double *)px, rtol, 1e-50, (int)maxit, &it, &rnorm, &why,
                                   NULL, 0) == PBX_OK,
               PETSC_COMM_SELF, PETSC_ERR_LIB, "pbx_cg_solve_device: %s", pbx_last_error());
    PetscCall(VecCUDARestoreArrayWrite(x, &px));
    PetscCall(VecCUDARestoreArrayRead(b, &pb));
    if (its) *its = it;
    if (reason) *reason = (KSPConvergedReason)why;
    PetscFunctionReturn(PETSC_SUCCESS);
}

int PbxShellSolveCG(Mat *A, Vec *b, Vec *x, double rtol, int maxit, int *its, int *reason)
{
    PetscInt k = 0;
    KSPConvergedReason why = (KSPConvergedReason)0;
    const int rc = (int)PbxSolve(*A, *b, *x, rtol, maxit, &k, &why);
    if (its) *its = (int)k;
    if (reason) *reason = (int)why;
    return rc;
}

/* ---- C hosts ------------------------------------------------------------------------------------ */

static PetscErrorCode PbxMatMult(Mat M, Vec x, Vec f) { return PbxMult(M, x, f); }

/* shell context of the C path: freed with the matrix */
static PetscErrorCode PbxShellDestroy(Mat A)
{
    PbxMatCtx *ctx = NULL;

    PetscFunctionBeginUser;
    PetscCall(MatShellGetContext(A, &ctx));
    PetscCall(PetscFree(ctx));
    PetscFunctionReturn(PETSC_SUCCESS);
}

/* counterpart of initialise_matrix_free (src/poissbox.f90:242-267) for a C caller */
PetscErrorCode PbxCreateShell(DM da, const PetscReal deltas[3], void *nccl_comm, Mat P, Mat *A)
{
    PbxMatCtx *ctx;
    PetscInt m, n;

    PetscFunctionBeginUser;
    PetscCall(PetscNew(&ctx));
    ctx->da = da;
    for (int d = 0; d < 3; ++d) ctx->grid_deltas[d] = deltas[d];
    PetscCall(MatGetLocalSize(P, &m, &n));                                                        /* :259 */
    PetscCall(MatCreateShell(PETSC_COMM_WORLD, m, n, PETSC_DETERMINE, PETSC_DETERMINE, ctx, A));  /* :261 */
    PetscCall(MatShellSetOperation(*A, MATOP_DESTROY, (void (*)(void))PbxShellDestroy));
    PetscCall(MatShellSetOperation(*A, MATOP_MULT, (void (*)(void))PbxMatMult));                  /* :263 */
    {
        const PetscErrorCode e = PbxAttach(*A, da, deltas, nccl_comm);
        if (e) {
            (void)MatDestroy(A);
            return e;
        }
    }
    PetscFunctionReturn(PETSC_SUCCESS);
}

PetscErrorCode PbxSolveCG(Mat A, Vec b, Vec x, PetscReal rtol, PetscInt maxit, PetscInt *its,
                          KSPConvergedReason *reason)
{
    return PbxSolve(A, b, x, rtol, maxit, its, reason);
}
