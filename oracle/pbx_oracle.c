/*
 * pbx_oracle.c -- CPU oracle (TEST INFRASTRUCTURE, see pbx_oracle.h for scope and pinning status).
 *
 * Each function restates, operation for operation and in the same evaluation order, the Fortran
 * routine cited above it.  Build: gcc -O2 -ffp-contract=off -pthread (oracle/Makefile).
 */
#include "pbx_oracle.h"

#include <math.h>
#include <stdlib.h>
#include <string.h>
#include <pthread.h>

static int g_threads = 1;

/* Minimal static-schedule parallel-for on pthreads (this image's gcc ships no libgomp).  With
 * g_threads == 1 the body runs inline on the caller: the reference's serial behaviour. */
typedef void (*range_fn)(long lo, long hi, int tid, void *ctx);
typedef struct { range_fn fn; long lo, hi; int tid; void *ctx; } par_task;
static void *par_tramp(void *p)
{
    par_task *t = (par_task *)p;
    t->fn(t->lo, t->hi, t->tid, t->ctx);
    return NULL;
}
static void par_for(long n, range_fn fn, void *ctx)
{
    int nt = g_threads;
    if (nt > n) nt = (int)(n > 0 ? n : 1);
    if (nt <= 1) { fn(0, n, 0, ctx); return; }
    pthread_t *th = (pthread_t *)malloc(sizeof(pthread_t) * (size_t)nt);
    par_task *tk = (par_task *)malloc(sizeof(par_task) * (size_t)nt);
    for (int t = 0; t < nt; ++t) {
        tk[t].fn = fn; tk[t].ctx = ctx; tk[t].tid = t;
        tk[t].lo = n * t / nt; tk[t].hi = n * (t + 1) / nt;
        pthread_create(&th[t], NULL, par_tramp, &tk[t]);
    }
    for (int t = 0; t < nt; ++t) pthread_join(th[t], NULL);
    free(tk);
    free(th);
}
#define PBX_MAX_THREADS 256

void orc_set_threads(int nthreads) { g_threads = nthreads < 1 ? 1 : nthreads; }
int orc_get_threads(void) { return g_threads; }

/* ------------------------------------------------------------------------------------------ */
/* src/tridsol.f90                                                                            */
/* ------------------------------------------------------------------------------------------ */

/* src/tridsol.f90:76-96.  NB the reference's dummy names: a = sub-diagonal, b = DIAGONAL
 * (overwritten with the pivots), c = super-diagonal, d = rhs. */
void orc_fwd_sweep(int n, const double *a, double *b, const double *c, double *d)
{
    for (int i = 1; i < n; ++i) {
        double w = a[i] / b[i - 1];
        b[i] = b[i] - w * c[i - 1];
        d[i] = d[i] - w * d[i - 1];
    }
}

/* src/tridsol.f90:98-115 */
void orc_bwd_sweep(int n, const double *b, const double *c, double *d)
{
    d[n - 1] = d[n - 1] / b[n - 1];
    for (int i = n - 2; i >= 0; --i)
        d[i] = (d[i] - c[i] * d[i + 1]) / b[i];
}

/* src/tridsol.f90:22-32 */
void orc_tdma(int n, const double *a, double *b, const double *c, double *d)
{
    orc_fwd_sweep(n, a, b, c, d);
    orc_bwd_sweep(n, b, c, d);
}

/* src/tridsol.f90:34-74.  Sherman-Morrison closure; b is not modified (the sweeps run on the
 * copy bmod, :54-57 and :59-66); the combine at :69-70 is an array assignment, so the right-hand
 * side sees the pre-assignment d(1), d(n). */
void orc_tdma_periodic(int n, const double *a, const double *b, const double *c, double *d)
{
    double *bmod = (double *)malloc(sizeof(double) * (size_t)n);
    double *u = (double *)malloc(sizeof(double) * (size_t)n);
    double gamma = -b[0]; /* :51 */

    memcpy(bmod, b, sizeof(double) * (size_t)n); /* :54 */
    bmod[0] = bmod[0] - gamma;
    bmod[n - 1] = bmod[n - 1] - c[n - 1] * a[0] / gamma;
    orc_tdma(n, a, bmod, c, d); /* :57 */

    memcpy(bmod, b, sizeof(double) * (size_t)n); /* :59 */
    bmod[0] = bmod[0] - gamma;
    bmod[n - 1] = bmod[n - 1] - c[n - 1] * a[0] / gamma;
    for (int i = 0; i < n; ++i) u[i] = 0.0; /* :63 */
    u[0] = gamma;
    u[n - 1] = c[n - 1];
    orc_tdma(n, a, bmod, c, u); /* :66 */

    {
        double d1 = d[0], dn = d[n - 1];
        double fac = d1 + (a[0] / gamma) * dn;                 /* :69 */
        double den = 1.0 + (u[0] + (a[0] / gamma) * u[n - 1]); /* :70 */
        for (int i = 0; i < n; ++i) d[i] = d[i] - (u[i] * fac) / den;
    }
    free(u);
    free(bmod);
}

/* ------------------------------------------------------------------------------------------ */
/* src/compact_schemes.f90, 1-D                                                               */
/* ------------------------------------------------------------------------------------------ */

/* src/compact_schemes.f90:332-372.  1-based index helper keeps the wrap cases readable. */
void orc_eval_1d_rhs(double a, double b, int opsign, int stagger, int n, const double *f,
                     double *rhs)
{
#define F(i) f[(i)-1]
#define R(i) rhs[(i)-1]
    const double s = (double)opsign; /* integer * real promotes the integer, :357 */
    int shift = (stagger == -1) ? 0 : 1;

    if (stagger == -1) {
        R(1) = a * (F(1) + s * F(n)) + b * (F(2) + s * F(n - 1));
        R(2) = a * (F(2) + s * F(1)) + b * (F(3) + s * F(n));
    } else {
        R(1) = a * (F(2) + s * F(1)) + b * (F(3) + s * F(n));
    }
    for (int i = 3 - shift; i <= n - 1 - shift; ++i)
        R(i) = a * (F(i + shift) + s * F(i - 1 + shift)) + b * (F(i + 1 + shift) + s * F(i - 2 + shift));
    if (stagger == -1) {
        R(n) = a * (F(n) + s * F(n - 1)) + b * (F(1) + s * F(n - 2));
    } else {
        R(n - 1) = a * (F(n) + s * F(n - 1)) + b * (F(1) + s * F(n - 2));
        R(n) = a * (F(1) + s * F(n)) + b * (F(2) + s * F(n - 1));
    }
#undef F
#undef R
}

/* shared tail of grad_1d (:183-202) and interp_1d (:298-317): constant [alpha,1,alpha] system */
static void solve_const_periodic(int n, double alpha, double *x)
{
    double *ld = (double *)malloc(sizeof(double) * (size_t)n);
    double *dg = (double *)malloc(sizeof(double) * (size_t)n);
    double *ud = (double *)malloc(sizeof(double) * (size_t)n);
    for (int i = 0; i < n; ++i) {
        ld[i] = alpha;
        dg[i] = 1.0;
        ud[i] = alpha;
    }
    orc_tdma_periodic(n, ld, dg, ud, x);
    free(ud);
    free(dg);
    free(ld);
}

/* src/compact_schemes.f90:155-204 */
int orc_grad_1d(int n, const double *f, double dx, int ndf, double *df, int stagger)
{
    if (ndf != n) return 7; /* :177-180 */
    double a = 63.0 / 62.0 / dx;         /* :188 */
    double b = 17.0 / 62.0 / (3.0 * dx); /* :189 */
    double alpha = 9.0 / 62.0;           /* :190 */
    orc_eval_1d_rhs(a, b, -1, stagger, n, f, df);
    solve_const_periodic(n, alpha, df);
    return 0;
}

/* src/compact_schemes.f90:260-268 */
int orc_div_1d(int n, const double *f, double dx, int ndf, double *df)
{
    return orc_grad_1d(n, f, dx, ndf, df, +1);
}

/* src/compact_schemes.f90:271-319 */
int orc_interp_1d(int n, const double *f, int nfi, double *fi, int stagger)
{
    if (nfi != n) return 7; /* :292-295 */
    double a = 0.75;         /* :303 */
    double b = 1.0 / 20.0;   /* :304 */
    double alpha = 3.0 / 10.0; /* :305 */
    orc_eval_1d_rhs(a, b, +1, stagger, n, f, fi);
    solve_const_periodic(n, alpha, fi);
    return 0;
}

/* src/compact_schemes.f90:322-329 */
int orc_interp_1d_div(int n, const double *f, int nfi, double *fi)
{
    return orc_interp_1d(n, f, nfi, fi, +1);
}

/* ------------------------------------------------------------------------------------------ */
/* src/compact_schemes.f90, 3-D.  The Fortran passes strided array sections f(i,j,:) to the     */
/* 1-D routines; here a line is gathered into a contiguous buffer, operated on, and scattered  */
/* (which is also what gfortran does for non-contiguous actual arguments).                    */
/* ------------------------------------------------------------------------------------------ */

typedef enum { OP_INTERP, OP_GRAD } op_t;

static void gather(const double *src, size_t stride, int n, double *dst)
{
    for (int i = 0; i < n; ++i) dst[i] = src[(size_t)i * stride];
}
static void scatter(const double *src, int n, double *dst, size_t stride)
{
    for (int i = 0; i < n; ++i) dst[(size_t)i * stride] = src[i];
}

/* apply one 1-D operator to every line of a 3-D field along dir (0=x,1=y,2=z) */
typedef struct {
    int nn[3]; size_t st[3]; int dir, d1, d2; op_t op; int stagger; double h;
    const double *in; double *out;
} lines_ctx;

static void lines_range(long lo, long hi, int tid, void *p)
{
    const lines_ctx *c = (const lines_ctx *)p;
    const int n = c->nn[c->dir];
    double *lin = (double *)malloc(sizeof(double) * (size_t)n);
    double *lout = (double *)malloc(sizeof(double) * (size_t)n);
    (void)tid;
    for (long l = lo; l < hi; ++l) {
        size_t off = (size_t)(l % c->nn[c->d1]) * c->st[c->d1] + (size_t)(l / c->nn[c->d1]) * c->st[c->d2];
        gather(c->in + off, c->st[c->dir], n, lin);
        if (c->op == OP_GRAD)
            orc_grad_1d(n, lin, c->h, n, lout, c->stagger);
        else
            orc_interp_1d(n, lin, n, lout, c->stagger);
        scatter(lout, n, c->out + off, c->st[c->dir]);
    }
    free(lout);
    free(lin);
}

static void lines_apply(int nx, int ny, int nz, int dir, op_t op, int stagger, double h,
                        const double *in, double *out)
{
    lines_ctx c = {{nx, ny, nz}, {1, (size_t)nx, (size_t)nx * (size_t)ny}, dir, (dir + 1) % 3,
                   (dir + 2) % 3, op, stagger, h, in, out};
    par_for((long)c.nn[c.d1] * c.nn[c.d2], lines_range, &c);
}

/* src/compact_schemes.f90:42-88.  Z -> Y -> X, backward stagger (cell -> vertex). */
void orc_grad(int nx, int ny, int nz, const double *f, const double dx[3], double *df)
{
    const size_t N = (size_t)nx * ny * nz;
    double *dff = (double *)malloc(sizeof(double) * 3 * N); /* :59 */
    double *dfe = (double *)malloc(sizeof(double) * 3 * N); /* :69 */
    /* :60-66 */
    lines_apply(nx, ny, nz, 2, OP_INTERP, -1, 0.0, f, dff);
    memcpy(dff + N, dff, sizeof(double) * N); /* :63 */
    lines_apply(nx, ny, nz, 2, OP_GRAD, -1, dx[2], f, dff + 2 * N);
    /* :70-76 */
    lines_apply(nx, ny, nz, 1, OP_INTERP, -1, 0.0, dff, dfe);
    lines_apply(nx, ny, nz, 1, OP_GRAD, -1, dx[1], dff + N, dfe + N);
    lines_apply(nx, ny, nz, 1, OP_INTERP, -1, 0.0, dff + 2 * N, dfe + 2 * N);
    /* :80-86 */
    lines_apply(nx, ny, nz, 0, OP_GRAD, -1, dx[0], dfe, df);
    lines_apply(nx, ny, nz, 0, OP_INTERP, -1, 0.0, dfe + N, df + N);
    lines_apply(nx, ny, nz, 0, OP_INTERP, -1, 0.0, dfe + 2 * N, df + 2 * N);
    free(dfe);
    free(dff);
}

/* src/compact_schemes.f90:207-257.  X -> Y -> Z, forward stagger (vertex -> cell). */
void orc_div(int nx, int ny, int nz, const double *f, const double dx[3], double *df)
{
    const size_t N = (size_t)nx * ny * nz;
    double *dfe = (double *)malloc(sizeof(double) * 3 * N); /* :225 */
    double *dff = (double *)malloc(sizeof(double) * 3 * N); /* :235 */
    double *tmp = (double *)malloc(sizeof(double) * N);
    /* :226-232 */
    lines_apply(nx, ny, nz, 0, OP_GRAD, +1, dx[0], f, dfe);
    lines_apply(nx, ny, nz, 0, OP_INTERP, +1, 0.0, f + N, dfe + N);
    lines_apply(nx, ny, nz, 0, OP_INTERP, +1, 0.0, f + 2 * N, dfe + 2 * N);
    /* :236-242 */
    lines_apply(nx, ny, nz, 1, OP_INTERP, +1, 0.0, dfe, dff);
    lines_apply(nx, ny, nz, 1, OP_GRAD, +1, dx[1], dfe + N, dff + N);
    lines_apply(nx, ny, nz, 1, OP_INTERP, +1, 0.0, dfe + 2 * N, dff + 2 * N);
    /* :247-253: the sum dff1 + dff2 is formed first (:249), interpolated into dfc, the z
     * derivative of dff3 lands in df, then df = df + dfc (:251). */
    for (size_t i = 0; i < N; ++i) tmp[i] = dff[i] + dff[N + i];
    lines_apply(nx, ny, nz, 2, OP_INTERP, +1, 0.0, tmp, dfe); /* dfe[0..N) reused as dfc */
    lines_apply(nx, ny, nz, 2, OP_GRAD, +1, dx[2], dff + 2 * N, df);
    for (size_t i = 0; i < N; ++i) df[i] = df[i] + dfe[i];
    free(tmp);
    free(dff);
    free(dfe);
}

/* src/compact_schemes.f90:93-142.  Z -> Y -> X with the caller's stagger. */
void orc_interp(int nx, int ny, int nz, const double *f, double *fi, int stagger)
{
    const size_t N = (size_t)nx * ny * nz;
    double *ff = (double *)malloc(sizeof(double) * N);
    double *fe = (double *)malloc(sizeof(double) * N);
    lines_apply(nx, ny, nz, 2, OP_INTERP, stagger, 0.0, f, ff);
    lines_apply(nx, ny, nz, 1, OP_INTERP, stagger, 0.0, ff, fe);
    lines_apply(nx, ny, nz, 0, OP_INTERP, stagger, 0.0, fe, fi);
    free(fe);
    free(ff);
}

/* src/compact_schemes.f90:144-152 */
void orc_interp_div(int nx, int ny, int nz, const double *f, double *fi)
{
    orc_interp(nx, ny, nz, f, fi, +1);
}

/* src/compact_schemes.f90:17-37 */
void orc_lapl(int nx, int ny, int nz, const double *f, const double dx[3], double *d2f)
{
    const size_t N = (size_t)nx * ny * nz;
    double *df = (double *)malloc(sizeof(double) * 3 * N); /* :30 */
    orc_grad(nx, ny, nz, f, dx, df);
    orc_div(nx, ny, nz, df, dx, d2f);
    free(df);
}

/* ------------------------------------------------------------------------------------------ */
/* The 2nd-order star (the operator mfmult applies today, src/poissbox.f90:300-322).            */
/* ------------------------------------------------------------------------------------------ */

/* src/coefficients.f90:22-37 */
void orc_lapl_1d_coeffs(double dx, double coeffs[3])
{
    const double invdx2 = 1.0 / (dx * dx); /* :29  1 / dx**2 */
    coeffs[0] = invdx2;                    /* :31 */
    coeffs[1] = -2.0 * invdx2;             /* :32 */
    coeffs[2] = invdx2;                    /* :33 */
}

/* src/coefficients.f90:40-50; coeffs(ii,jj,kk) <-> coeffs[ii + 3*(jj + 3*kk)], 0-based */
void orc_lapl_star_coeffs(double dx, double dy, double dz, double coeffs[27])
{
    double c[3];
    for (int i = 0; i < 27; ++i) coeffs[i] = 0.0; /* :45 */
    orc_lapl_1d_coeffs(dx, c);
    for (int i = 0; i < 3; ++i) coeffs[i + 3 * (1 + 3 * 1)] = coeffs[i + 3 * (1 + 3 * 1)] + c[i]; /* :46 */
    orc_lapl_1d_coeffs(dy, c);
    for (int j = 0; j < 3; ++j) coeffs[1 + 3 * (j + 3 * 1)] = coeffs[1 + 3 * (j + 3 * 1)] + c[j]; /* :47 */
    orc_lapl_1d_coeffs(dz, c);
    for (int k = 0; k < 3; ++k) coeffs[1 + 3 * (1 + 3 * k)] = coeffs[1 + 3 * (1 + 3 * k)] + c[k]; /* :48 */
}

/* src/poissbox.f90:128-148: dot_product over the flattened 3x3x3 boxes, in array element order */
double orc_evaluate_laplacian_pointwise(const double f[27], const double grid_deltas[3])
{
    double coeffs[27], s = 0.0;
    orc_lapl_star_coeffs(grid_deltas[0], grid_deltas[1], grid_deltas[2], coeffs);
    for (int i = 0; i < 27; ++i) s = s + f[i] * coeffs[i];
    return s;
}

typedef struct { int nx, ny, nz; const double *x; const double *dx; double *b; } star_ctx;

static void star_range(long lo, long hi, int tid, void *p)
{
    star_ctx *c = (star_ctx *)p;
    (void)tid;
    const int nx = c->nx, ny = c->ny, nz = c->nz;
    for (long k = lo; k < hi; ++k)
        for (int j = 0; j < ny; ++j)
            for (int i = 0; i < nx; ++i) {
                double f[27]; /* xdof(i-1:i+1, j-1:j+1, k-1:k+1), :107, periodic ghosts */
                for (int kk = 0; kk < 3; ++kk)
                    for (int jj = 0; jj < 3; ++jj)
                        for (int ii = 0; ii < 3; ++ii) {
                            const int gi = (i + ii - 1 + nx) % nx, gj = (j + jj - 1 + ny) % ny;
                            const int gk = (int)((k + kk - 1 + nz) % nz);
                            f[ii + 3 * (jj + 3 * kk)] = c->x[gi + (size_t)nx * (gj + (size_t)ny * gk)];
                        }
                c->b[i + (size_t)nx * (j + (size_t)ny * k)] = orc_evaluate_laplacian_pointwise(f, c->dx);
            }
}

/* src/poissbox.f90:84-126 */
void orc_star(int nx, int ny, int nz, const double *x, const double dx[3], double *b)
{
    star_ctx c = {nx, ny, nz, x, dx, b};
    par_for(nz, star_range, &c);
}

/* ------------------------------------------------------------------------------------------ */
/* CG.  Follows PETSc KSPCG (third-party; src/poissbox.f90:293-296 creates the KSP, :285-291     */
/* attaches a constant MatNullSpace to A and P).  With -pc_type none the "preconditioner        */
/* apply" is a copy followed by null-space removal (z = r - mean(r)), the default norm is the   */
/* preconditioned one (||z||_2), beta = z.r, and the default convergence test compares against  */
/* max(rtol * ||z_0||, abstol) with a divergence guard at 1e4 * ||z_0||.  PARITY UNPINNED: the   */
/* reference records no CG output for any operator.                                            */
/* ------------------------------------------------------------------------------------------ */

typedef struct { const double *a, *b; double *y; double s1, s2; double part[PBX_MAX_THREADS]; } vec_ctx;

static void dot_range(long lo, long hi, int tid, void *p)
{
    vec_ctx *c = (vec_ctx *)p;
    double s = 0.0;
    for (long i = lo; i < hi; ++i) s += c->a[i] * c->b[i];
    c->part[tid] = s;
}
static void sum_range(long lo, long hi, int tid, void *p)
{
    vec_ctx *c = (vec_ctx *)p;
    double s = 0.0;
    for (long i = lo; i < hi; ++i) s += c->a[i];
    c->part[tid] = s;
}
static double collect(vec_ctx *c)
{
    double s = 0.0;
    int nt = g_threads < PBX_MAX_THREADS ? g_threads : PBX_MAX_THREADS;
    for (int t = 0; t < nt; ++t) s += c->part[t];
    return s;
}
static double vdot(size_t N, const double *a, const double *b)
{
    vec_ctx c;
    memset(&c, 0, sizeof c);
    c.a = a; c.b = b;
    par_for((long)N, dot_range, &c);
    return collect(&c);
}
static double vsum(size_t N, const double *a)
{
    vec_ctx c;
    memset(&c, 0, sizeof c);
    c.a = a;
    par_for((long)N, sum_range, &c);
    return collect(&c);
}
static void shift_range(long lo, long hi, int tid, void *p)
{
    vec_ctx *c = (vec_ctx *)p;
    (void)tid;
    for (long i = lo; i < hi; ++i) c->y[i] = c->a[i] + c->s1;
}
/* z = r with the constant removed: VecSum, scale by -1/N, VecShift (MatNullSpaceRemove) */
static void pc_apply(size_t N, const double *r, double *z)
{
    vec_ctx c;
    memset(&c, 0, sizeof c);
    c.a = r; c.y = z;
    c.s1 = vsum(N, r) / (-1.0 * (double)N);
    par_for((long)N, shift_range, &c);
}
/* p = z + bb * p */
static void aypx_range(long lo, long hi, int tid, void *p)
{
    vec_ctx *c = (vec_ctx *)p;
    (void)tid;
    for (long i = lo; i < hi; ++i) c->y[i] = c->a[i] + c->s1 * c->y[i];
}
/* y = y + s1 * a */
static void axpy_range(long lo, long hi, int tid, void *p)
{
    vec_ctx *c = (vec_ctx *)p;
    (void)tid;
    for (long i = lo; i < hi; ++i) c->y[i] = c->y[i] + c->s1 * c->a[i];
}
static void vaxpy(size_t N, double s, const double *a, double *y)
{
    vec_ctx c;
    memset(&c, 0, sizeof c);
    c.a = a; c.y = y; c.s1 = s;
    par_for((long)N, axpy_range, &c);
}
static void vaypx(size_t N, double s, const double *a, double *y)
{
    vec_ctx c;
    memset(&c, 0, sizeof c);
    c.a = a; c.y = y; c.s1 = s;
    par_for((long)N, aypx_range, &c);
}

int orc_cg_solve(int nx, int ny, int nz, const double dx[3], const double *b, double *x,
                 double rtol, double abstol, int maxit, double *rnorm, int *reason, double *hist,
                 int nhist)
{
    return orc_cg_solve_op(0, nx, ny, nz, dx, b, x, rtol, abstol, maxit, rnorm, reason, hist, nhist);
}

/* op = 0: the compact Laplacian (orc_lapl); 1: the 2nd-order star (orc_star), the operator the
 * reference's shell matrix applies today */
int orc_cg_solve_op(int op, int nx, int ny, int nz, const double dx[3], const double *b, double *x,
                    double rtol, double abstol, int maxit, double *rnorm, int *reason, double *hist,
                    int nhist)
{
    const size_t N = (size_t)nx * ny * nz;
    double *r = (double *)malloc(sizeof(double) * N);
    double *z = (double *)malloc(sizeof(double) * N);
    double *p = (double *)malloc(sizeof(double) * N);
    double *w = (double *)malloc(sizeof(double) * N);
    double dp, dp0, beta, betaold = 0.0, dpi = 0.0, dpiold, ttol;
    const double dtol = 1.0e4;
    int i = 0, its = 0, why = 0;

    for (size_t k = 0; k < N; ++k) x[k] = 0.0; /* zero initial guess */
    memcpy(r, b, sizeof(double) * N);
    pc_apply(N, r, z);
    dp = sqrt(vdot(N, z, z));
    dp0 = dp;
    if (hist && nhist > 0) hist[0] = dp;
    ttol = fmax(rtol * dp, abstol);
    if (dp != dp) why = -9;
    else if (dp <= ttol) why = (dp < abstol) ? 3 : 2;
    if (!why) {
        beta = vdot(N, z, r);
        do {
            its = i + 1;
            if (beta == 0.0) { why = 3; break; }
            if (i > 0 && beta * betaold < 0.0) { why = -8; break; }
            if (i == 0) {
                memcpy(p, z, sizeof(double) * N);
            } else {
                double bb = beta / betaold;
                vaypx(N, bb, z, p); /* VecAYPX: p <- z + b p */
            }
            dpiold = dpi;
            if (op == 1)
                orc_star(nx, ny, nz, p, dx, w);
            else
                orc_lapl(nx, ny, nz, p, dx, w);
            dpi = vdot(N, p, w);
            betaold = beta;
            if (dpi == 0.0 || (i > 0 && ((dpi > 0) - (dpi < 0)) * ((dpiold > 0) - (dpiold < 0)) < 0)) {
                why = -10; /* KSP_DIVERGED_INDEFINITE_MAT */
                break;
            }
            {
                double a = beta / dpi;
                vaxpy(N, a, p, x);  /* VecAXPY: x <- x + a p */
                vaxpy(N, -a, w, r); /* VecAXPY: r <- r - a w */
            }
            pc_apply(N, r, z);
            dp = sqrt(vdot(N, z, z));
            if (hist && i + 1 < nhist) hist[i + 1] = dp;
            if (dp != dp) { why = -9; break; }
            if (dp <= ttol) { why = (dp < abstol) ? 3 : 2; break; }
            if (dp >= dtol * dp0) { why = -4; break; }
            beta = vdot(N, z, r);
            ++i;
        } while (i < maxit);
        if (!why && i >= maxit) why = -3;
    }
    if (rnorm) *rnorm = dp;
    if (reason) *reason = why;
    free(w);
    free(p);
    free(z);
    free(r);
    return its;
}
