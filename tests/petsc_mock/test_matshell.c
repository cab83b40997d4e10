/* Drives petsc/pbx_matshell.c through the mock of tests/petsc_mock/petsc_mock.h on a real GPU:
 * PbxCreateShell on a one-rank z-slab "DMDA", MatMult against pbx_lapl_host of the same field,
 * PbxSolveCG on b = A x.  Exit status 0 = pass, 2 = no CUDA device. */
#include <math.h>
#include <string.h>

#include "../../petsc/pbx_matshell.c"

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
    if (PbxCreateShell(&bad, dx, NULL, &pmat, &B) != PETSC_ERR_SUP) return 1;
    printf("PASS\n");
    return 0;
}
