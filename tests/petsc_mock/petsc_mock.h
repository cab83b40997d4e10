/*
 * petsc_mock.h -- a few dozen lines standing in for the part of the PETSc API that
 * petsc/pbx_matshell.c uses, so that the glue can be COMPILED and its MatMult / solve paths
 * EXECUTED on a GPU box that has no PETSc (this image has neither PETSc nor MPI).
 * TEST INFRASTRUCTURE ONLY.  Signatures are restated from PETSc's public headers (petscmat.h,
 * petscvec.h, petscdmda.h, petscksp.h, petscdevice_cuda.h; PETSc >= 3.19 naming: PetscCall,
 * PETSC_SUCCESS); a Vec here is a bare device array, a Mat a shell context with one operation, a DM
 * the extents of one rank's brick of a z-slab DMDA.
 */
#ifndef PETSC_MOCK_H
#define PETSC_MOCK_H

#include <cuda_runtime.h>
#include <stdio.h>
#include <stdlib.h>

typedef int PetscErrorCode;
typedef int PetscInt;
typedef double PetscReal;
typedef double PetscScalar;
typedef int MPI_Comm;
typedef const char *VecType;
typedef int KSPConvergedReason;
#define PETSC_SUCCESS 0
#define PETSC_ERR_LIB 76
#define PETSC_ERR_SUP 56
#define PETSC_COMM_SELF 1
#define PETSC_COMM_WORLD 2
#define PETSC_DETERMINE (-1)
#define VECCUDA "cuda"
typedef enum { MATOP_MULT = 3 } MatOperation;

typedef struct _p_Vec {
    double *dev;
    PetscInt n;
    int reads, writes;   /* outstanding Get without Restore */
} *Vec;
typedef struct _p_Mat {
    void *ctx;
    PetscErrorCode (*mult)(struct _p_Mat *, Vec, Vec);
    PetscInt m, n;
    VecType vtype;
} *Mat;
typedef struct _p_DM {
    PetscInt M, N, P;             /* global extents */
    PetscInt xs, ys, zs, xm, ym, zm;   /* this rank's corner and widths */
} *DM;

#define PetscFunctionBeginUser
#define PetscFunctionReturn(x) return (x)
#define PetscCall(call)                         \
    do {                                        \
        PetscErrorCode ierr__ = (call);         \
        if (ierr__) return ierr__;              \
    } while (0)
#define PetscCallCUDA(call)                     \
    do {                                        \
        if ((call) != cudaSuccess) return PETSC_ERR_LIB; \
    } while (0)
#define PetscCheck(cond, comm, err, ...)        \
    do {                                        \
        if (!(cond)) {                          \
            fprintf(stderr, __VA_ARGS__);       \
            fprintf(stderr, "\n");              \
            return (err);                       \
        }                                       \
    } while (0)
#define PetscNew(p) ((*(void **)(p) = calloc(1, sizeof(**(p)))) ? PETSC_SUCCESS : PETSC_ERR_LIB)

static inline PetscErrorCode MatShellGetContext(Mat A, void *ctx)
{
    *(void **)ctx = A->ctx;
    return PETSC_SUCCESS;
}
static inline PetscErrorCode MatGetLocalSize(Mat A, PetscInt *m, PetscInt *n)
{
    *m = A->m;
    *n = A->n;
    return PETSC_SUCCESS;
}
static inline PetscErrorCode MatCreateShell(MPI_Comm c, PetscInt m, PetscInt n, PetscInt M, PetscInt N, void *ctx, Mat *A)
{
    (void)c; (void)M; (void)N;
    *A = (Mat)calloc(1, sizeof(**A));
    (*A)->ctx = ctx;
    (*A)->m = m;
    (*A)->n = n;
    return PETSC_SUCCESS;
}
static inline PetscErrorCode MatShellSetVecType(Mat A, VecType t)
{
    A->vtype = t;
    return PETSC_SUCCESS;
}
static inline PetscErrorCode MatShellSetOperation(Mat A, MatOperation op, void (*f)(void))
{
    if (op != MATOP_MULT) return PETSC_ERR_SUP;
    A->mult = (PetscErrorCode (*)(struct _p_Mat *, Vec, Vec))f;
    return PETSC_SUCCESS;
}
static inline PetscErrorCode MatMult(Mat A, Vec x, Vec y) { return A->mult(A, x, y); }
static inline PetscErrorCode VecCUDAGetArrayRead(Vec v, const PetscScalar **a)
{
    ++v->reads;
    *a = v->dev;
    return PETSC_SUCCESS;
}
static inline PetscErrorCode VecCUDARestoreArrayRead(Vec v, const PetscScalar **a)
{
    --v->reads;
    *a = NULL;
    return PETSC_SUCCESS;
}
static inline PetscErrorCode VecCUDAGetArrayWrite(Vec v, PetscScalar **a)
{
    ++v->writes;
    *a = v->dev;
    return PETSC_SUCCESS;
}
static inline PetscErrorCode VecCUDARestoreArrayWrite(Vec v, PetscScalar **a)
{
    --v->writes;
    *a = NULL;
    return PETSC_SUCCESS;
}
static inline PetscErrorCode DMDAGetInfo(DM da, PetscInt *dim, PetscInt *M, PetscInt *N, PetscInt *P, PetscInt *m,
                                         PetscInt *n, PetscInt *p, PetscInt *dof, PetscInt *s, void *bx, void *by,
                                         void *bz, void *st)
{
    (void)dim; (void)m; (void)n; (void)p; (void)dof; (void)s; (void)bx; (void)by; (void)bz; (void)st;
    *M = da->M;
    *N = da->N;
    *P = da->P;
    return PETSC_SUCCESS;
}
static inline PetscErrorCode DMDAGetCorners(DM da, PetscInt *xs, PetscInt *ys, PetscInt *zs, PetscInt *xm,
                                            PetscInt *ym, PetscInt *zm)
{
    *xs = da->xs; *ys = da->ys; *zs = da->zs;
    *xm = da->xm; *ym = da->ym; *zm = da->zm;
    return PETSC_SUCCESS;
}
#endif
