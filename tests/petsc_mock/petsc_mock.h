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
#include <string.h>

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
#define PETSC_ERR_ARG_WRONGSTATE 73
#define PETSC_COMM_SELF 1
#define PETSC_COMM_WORLD 2
#define PETSC_DETERMINE (-1)
#define VECCUDA "cuda"
typedef enum { MATOP_MULT = 3, MATOP_DESTROY = 60 } MatOperation;

/* PetscObject / PetscContainer: one composed object per Mat is all the glue needs */
typedef struct _p_PetscContainer {
    void *ptr;
    PetscErrorCode (*destroy)(void *);
    int refs;
} *PetscContainer;
typedef void *PetscObject;
typedef struct _p_PetscDeviceContext {
    cudaStream_t stream;
} *PetscDeviceContext;

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
    PetscErrorCode (*destroy)(struct _p_Mat *);
    char key[32];            /* the one composed object */
    PetscContainer composed;
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
    if (op == MATOP_MULT)
        A->mult = (PetscErrorCode (*)(struct _p_Mat *, Vec, Vec))f;
    else if (op == MATOP_DESTROY)
        A->destroy = (PetscErrorCode (*)(struct _p_Mat *))f;
    else
        return PETSC_ERR_SUP;
    return PETSC_SUCCESS;
}
static inline PetscErrorCode PetscFree_(void *p)
{
    free(p);
    return PETSC_SUCCESS;
}
#define PetscFree(p) PetscFree_((void *)(p))
static inline PetscErrorCode PetscContainerCreate(MPI_Comm c, PetscContainer *out)
{
    (void)c;
    *out = (PetscContainer)calloc(1, sizeof(**out));
    (*out)->refs = 1;
    return PETSC_SUCCESS;
}
static inline PetscErrorCode PetscContainerSetPointer(PetscContainer c, void *p)
{
    c->ptr = p;
    return PETSC_SUCCESS;
}
static inline PetscErrorCode PetscContainerGetPointer(PetscContainer c, void **p)
{
    *p = c->ptr;
    return PETSC_SUCCESS;
}
static inline PetscErrorCode PetscContainerSetUserDestroy(PetscContainer c, PetscErrorCode (*d)(void *))
{
    c->destroy = d;
    return PETSC_SUCCESS;
}
static inline PetscErrorCode PetscContainerDestroy(PetscContainer *c)
{
    if (*c && --(*c)->refs == 0) {
        if ((*c)->destroy) (*c)->destroy((*c)->ptr);
        free(*c);
    }
    *c = NULL;
    return PETSC_SUCCESS;
}
/* only Mats are composed on in the glue */
static inline PetscErrorCode PetscObjectCompose(PetscObject obj, const char *key, PetscObject what)
{
    Mat A = (Mat)obj;
    PetscContainer c = (PetscContainer)what;
    if (A->composed) PetscContainerDestroy(&A->composed);
    snprintf(A->key, sizeof A->key, "%s", key);
    A->composed = c;
    if (c) ++c->refs;
    return PETSC_SUCCESS;
}
static inline PetscErrorCode PetscObjectQuery(PetscObject obj, const char *key, PetscObject *what)
{
    Mat A = (Mat)obj;
    *what = (A->composed && strcmp(A->key, key) == 0) ? (PetscObject)A->composed : NULL;
    return PETSC_SUCCESS;
}
static inline PetscErrorCode MatDestroy(Mat *A)
{
    if (!*A) return PETSC_SUCCESS;
    if ((*A)->destroy) (*A)->destroy(*A);
    if ((*A)->composed) PetscContainerDestroy(&(*A)->composed);
    free(*A);
    *A = NULL;
    return PETSC_SUCCESS;
}
/* the stream PETSc's vector kernels run on; the test switches it to a non-blocking stream */
static struct _p_PetscDeviceContext petsc_mock_dctx = {0};
static inline PetscErrorCode PetscDeviceContextGetCurrentContext(PetscDeviceContext *d)
{
    *d = &petsc_mock_dctx;
    return PETSC_SUCCESS;
}
static inline PetscErrorCode PetscDeviceContextGetStreamHandle(PetscDeviceContext d, void **handle)
{
    *handle = &d->stream;
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
