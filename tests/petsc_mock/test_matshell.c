/* Drives petsc/pbx_matshell.c through the mock of tests/petsc_mock/petsc_mock.h on a real GPU:
 * PbxCreateShell on a one-rank z-slab "DMDA", MatMult against pbx_lapl_host of the same field,
 * PbxSolveCG on b = A x.  Exit status 0 = pass, 2 = no CUDA device. */
#include <math.h>
#include <string.h>

#include "../../petsc/pbx_matshell.c"

/* ---- the Fortran side, acted out in C -----------------------------------------------------------
 * What gfortran + PETSc's Fortran stubs do with fortran/poissbox_matfree_pbx.f90: the shell context is
 * the Fortran derived type mat_ctx {da, grid_deltas} (src/poissbox.f90:17-20), objects travel by
 * reference, the MatMult callback has the Fortran signature (M, x, f, ierr) (src/poissbox.f90:300-311). */
typedef struct {
    DM da;
    double grid_deltas[3];
} fortran_mat_ctx;

static int mfmult_calls = 0;
static void mfmult_fortran(Mat *M, Vec *x, Vec *f, int *ierr)
{
    fortran_mat_ctx *ctx = NULL;
    *ierr = MatShellGetContext(*M, &ctx);            /* :315 -- the Fortran context is still what comes back */
    if (*ierr || !ctx || ctx->da == NULL || ctx->grid_deltas[0] <= 0.0) {
        *ierr = 99;
        return;
    }
    ++mfmult_calls;
    *ierr = PbxShellMult(M, x, f);                   /* replaces compute_lapl_pointwise, :316 */
}
/* PETSc's MatMult_Shell calling a Fortran callback */
static PetscErrorCode mfmult_trampoline(Mat M, Vec x, Vec f)
{
    int ierr = 0;
    mfmult_fortran(&M, &x, &f, &ierr);
    return ierr;
}

/* initialise_matrix_free (src/poissbox.f90:242-267) with the one added call */
static int initialise_matrix_free_fortran(fortran_mat_ctx *ctx, Mat P, Mat *A)
{
    PetscInt m, n;
    int ierr;
    if ((ierr = MatGetLocalSize(P, &m, &n))) return ierr;
    if ((ierr = MatCreateShell(PETSC_COMM_WORLD, m, n, PETSC_DETERMINE, PETSC_DETERMINE, ctx, A))) return ierr;
    if ((ierr = PbxShellAttach(A, &ctx->da, ctx->grid_deltas, NULL))) return ierr;   /* new */
    return MatShellSetOperation(*A, MATOP_MULT, (void (*)(void))mfmult_trampoline);
}

int main(void)
{
    enum { NX = 32, NY = 16, NZ = 48, N = NX * NY * NZ };
    if (pbx_device_count() <= 0) {
        fprintf(stderr, "no CUDA device (there is no CPU fallback)\n");
        return 2;
    }
    static double f[N], want[N], got[N], sol[N];
    unsigned s = 12345u;
    for (int i = 0; i < N; ++i) {
        s = s * 1664525u + 1013904223u;
        f[i] = (double)(s >> 8) / (double)(1u << 24) * 2.0 - 1.0;
    }
    const PetscReal dx[3] = {1.0 / NX, 0.5 / NY, 2.0 / NZ};
    if (pbx_lapl_host(NX, NY, NZ, f, dx, want, PBX_MODE_FAST) != PBX_OK) return 1;

    struct _p_DM da = {NX, NY, NZ, 0, 0, 0, NX, NY, NZ};
    struct _p_Mat pmat = {NULL, NULL, N, N, NULL};
    struct _p_Vec x = {NULL, N, 0, 0}, y = {NULL, N, 0, 0}, z = {NULL, N, 0, 0};
    cudaMalloc((void **)&x.dev, sizeof f);
    cudaMalloc((void **)&y.dev, sizeof f);
    cudaMalloc((void **)&z.dev, sizeof f);
    cudaMemcpy(x.dev, f, sizeof f, cudaMemcpyHostToDevice);
    Mat A = NULL;
    if (PbxCreateShell(&da, dx, NULL, &pmat, &A) != PETSC_SUCCESS) return 1;
    if (strcmp(A->vtype, VECCUDA) != 0 || A->m != N) return 1;
    if (MatMult(A, &x, &y) != PETSC_SUCCESS) return 1;
    cudaMemcpy(got, y.dev, sizeof f, cudaMemcpyDeviceToHost);
    if (x.reads || x.writes || y.reads || y.writes) {
        fprintf(stderr, "FAIL: array access not restored\n");
        return 1;
    }
    if (memcmp(got, want, sizeof f) != 0) {
        fprintf(stderr, "FAIL: MatMult differs from pbx_lapl_host\n");
        return 1;
    }
    /* solve A sol = A f and check the true residual with another MatMult (A annihilates the
     * constants and, on even grids, the modes that sit at Nyquist in two directions, so sol is
     * compared through A, not with f) */
    PetscInt its = 0;
    KSPConvergedReason why = 0;
    if (PbxSolveCG(A, &y, &z, 1e-10, 10000, &its, &why) != PETSC_SUCCESS) return 1;
    if (MatMult(A, &z, &x) != PETSC_SUCCESS) return 1;   /* x <- A sol */
    cudaMemcpy(sol, x.dev, sizeof f, cudaMemcpyDeviceToHost);
    double err = 0, nrm = 0;
    for (int i = 0; i < N; ++i) {
        const double d = sol[i] - want[i];
        err += d * d;
        nrm += want[i] * want[i];
    }
    printf("MatMult identical to pbx_lapl_host; CG %d its, reason %d, true residual %.2e\n", its, why, sqrt(err / nrm));
    if (why != 2 || sqrt(err / nrm) > 1e-7) return 1;
    /* a DMDA that is not z-slabs is refused */
    struct _p_DM bad = {NX, NY, NZ, 0, 0, 0, NX / 2, NY, NZ};
    Mat B = NULL;
    if (PbxCreateShell(&bad, dx, NULL, &pmat, &B) != PETSC_ERR_SUP || B != NULL) return 1;

    /* the same through the Fortran-callable layer, on a NON-BLOCKING stream standing in for PETSc's device
     * context: the Fortran mat_ctx stays the shell context, the handle rides on the Mat */
    cudaStream_t ps;
    cudaStreamCreateWithFlags(&ps, cudaStreamNonBlocking);
    petsc_mock_dctx.stream = ps;
    fortran_mat_ctx fctx = {&da, {dx[0], dx[1], dx[2]}};
    Mat F = NULL;
    if (initialise_matrix_free_fortran(&fctx, &pmat, &F) != 0) return 1;
    fortran_mat_ctx *back = NULL;
    MatShellGetContext(F, &back);
    if (back != &fctx || strcmp(F->vtype, VECCUDA) != 0) {
        fprintf(stderr, "FAIL: the Fortran mat_ctx is not the shell context any more\n");
        return 1;
    }
    cudaMemcpyAsync(x.dev, f, sizeof f, cudaMemcpyHostToDevice, ps);   /* "PETSc's" work on its own stream */
    cudaMemsetAsync(y.dev, 0, sizeof f, ps);
    if (MatMult(F, &x, &y) != PETSC_SUCCESS || mfmult_calls != 1) return 1;
    cudaMemcpyAsync(got, y.dev, sizeof f, cudaMemcpyDeviceToHost, ps);  /* ordered after the MatMult by the stream */
    cudaStreamSynchronize(ps);
    if (memcmp(got, want, sizeof f) != 0) {
        fprintf(stderr, "FAIL: Fortran-path MatMult differs from pbx_lapl_host\n");
        return 1;
    }
    pbx_handle hh = NULL;
    if (PbxShellGetHandle(&F, &hh) != 0 || hh == NULL) return 1;
    int fits = 0, fwhy = 0;
    Vec vy = &y, vz = &z;
    if (PbxShellSolveCG(&F, &vy, &vz, 1e-10, 10000, &fits, &fwhy) != 0 || fwhy != 2 || fits != its) return 1;
    /* a Mat nobody attached a handle to is refused, and destroying the matrices frees the handles */
    Mat bare = NULL;
    MatCreateShell(PETSC_COMM_WORLD, N, N, PETSC_DETERMINE, PETSC_DETERMINE, &fctx, &bare);
    if (PbxShellMult(&bare, &vy, &vz) != PETSC_ERR_ARG_WRONGSTATE) return 1;
    MatDestroy(&bare);
    MatDestroy(&F);
    MatDestroy(&A);
    petsc_mock_dctx.stream = 0;
    cudaStreamDestroy(ps);
    printf("Fortran-callable layer: mat_ctx kept as shell context, MatMult identical on a non-blocking stream, CG %d its\n", fits);
    printf("PASS\n");
    return 0;
}
