/*
 * pbx_oracle.h -- CPU ORACLE for the poissbox compact-Laplacian hot path.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing under oracle/ is part of the product: only tests/,
 * __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may load it, and
 * there only as the checker / the CPU baseline, never as the thing shipped.
 *
 * What it is: an op-for-op C restatement of the reference's Fortran
 *   src/tridsol.f90:22-115          (tdma, tdma_periodic, fwd_sweep, bwd_sweep)
 *   src/compact_schemes.f90:17-372  (lapl, grad, interp, interp_div, grad_1d, div, div_1d,
 *                                    interp_1d, interp_1d_div, eval_1d_rhs)
 * plus a conjugate-gradient loop that follows the public semantics of PETSc's KSPCG as the
 * reference drives it (src/poissbox.f90:285-296): third-party, not vendored, version unpinned.
 *
 * Pinning status.  The reference cannot be compiled in this image (no Fortran compiler, no MPI, no
 * PETSc) and its tests hold no stored golden vectors; they hold analytic known-answer tests.  The
 * oracle is pinned against every in-scope one of those (tests/test_oracle_kat.py restates
 * tests/tridiag, tests/coefficients/test_compact, tests/grad, tests/div, tests/lapl with their
 * tolerances) and, independently, against a dense/spectral numpy evaluation of the same operators.
 * The CG has no reference-side test or recorded output at all: **CG parity is unpinned**.
 *
 * Array layout: Fortran column-major.  f(i,j,k) <-> f[i + nx*(j + ny*k)] (0-based here);
 * df(i,j,k,c) adds c*nx*ny*nz (src/compact_schemes.f90:19-23,46).
 * Arithmetic: IEEE fp64, built with -O2 -ffp-contract=off (the reference build has no FMA
 * contraction on baseline x86-64 and no fast-math, CMakeLists.txt:16).
 */
#ifndef PBX_ORACLE_H
#define PBX_ORACLE_H

#ifdef __cplusplus
extern "C" {
#endif

/* ---- src/tridsol.f90 ---- */
void orc_fwd_sweep(int n, const double *a, double *b, const double *c, double *d);
void orc_bwd_sweep(int n, const double *b, const double *c, double *d);
void orc_tdma(int n, const double *a, double *b, const double *c, double *d);
void orc_tdma_periodic(int n, const double *a, const double *b, const double *c, double *d);

/* ---- src/compact_schemes.f90, 1-D ---- */
void orc_eval_1d_rhs(double a, double b, int opsign, int stagger, int n, const double *f,
                     double *rhs);
/* return 0, or 7 on size mismatch (the reference does `stop 7`, compact_schemes.f90:177-180) */
int orc_grad_1d(int n, const double *f, double dx, int ndf, double *df, int stagger);
int orc_div_1d(int n, const double *f, double dx, int ndf, double *df);
int orc_interp_1d(int n, const double *f, int nfi, double *fi, int stagger);
int orc_interp_1d_div(int n, const double *f, int nfi, double *fi);

/* ---- src/compact_schemes.f90, 3-D ---- */
void orc_grad(int nx, int ny, int nz, const double *f, const double dx[3], double *df);
void orc_div(int nx, int ny, int nz, const double *f, const double dx[3], double *df);
void orc_interp(int nx, int ny, int nz, const double *f, double *fi, int stagger);
void orc_interp_div(int nx, int ny, int nz, const double *f, double *fi);
void orc_lapl(int nx, int ny, int nz, const double *f, const double dx[3], double *d2f);

/* ---- the 2nd-order 7-point star: what the reference's MatMult callback computes TODAY ----
 * src/coefficients.f90:22-48 (lapl_1d_coeffs, lapl_star_coeffs) and src/poissbox.f90:84-148
 * (compute_lapl_pointwise, evaluate_laplacian_pointwise), called by mfmult :300-322; periodic
 * ghosts as the DMDA provides them (DM_BOUNDARY_PERIODIC). */
void orc_lapl_1d_coeffs(double dx, double coeffs[3]);
void orc_lapl_star_coeffs(double dx, double dy, double dz, double coeffs[27]);
double orc_evaluate_laplacian_pointwise(const double f[27], const double grid_deltas[3]);
void orc_star(int nx, int ny, int nz, const double *x, const double dx[3], double *b);

/* thread count used by the 3-D routines' loops over lines (1 = the reference's serial behaviour;
 * >1 is the "all host cores" courtesy baseline, results are bit-identical either way). */
void orc_set_threads(int nthreads);
int orc_get_threads(void);

/* ---- CG with PETSc KSPCG semantics on the oracle Laplacian (PC none, constant null space) ----
 * x0 = 0.  Returns the iteration count; *reason > 0 converged (2 = rtol, 3 = atol),
 * < 0 diverged (-3 = max_it, -4 = dtol, -8 = indefinite PC, -9 = NaN, -10 = indefinite operator).  hist (may be NULL) receives the
 * residual norm at iterations 0..its, at most nhist entries. */
int orc_cg_solve(int nx, int ny, int nz, const double dx[3], const double *b, double *x,
                 double rtol, double abstol, int maxit, double *rnorm, int *reason, double *hist,
                 int nhist);

int orc_cg_solve_op(int op, int nx, int ny, int nz, const double dx[3], const double *b, double *x,
                    double rtol, double abstol, int maxit, double *rnorm, int *reason, double *hist,
                    int nhist);

#ifdef __cplusplus
}
#endif
#endif
