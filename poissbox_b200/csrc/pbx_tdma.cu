// pbx_tdma.cu -- batched general-coefficient tridiagonal solves, module tridsol of the reference
// (src/tridsol.f90:22-115), one thread per line.  Arithmetic is the reference's, operation for
// operation (round-to-nearest intrinsics: no FMA contraction, IEEE division), so a line solved here
// carries the same bits as the CPU oracle's.  Element i of line l is base[l*ls + i*es]; choose
// es = number of lines, ls = 1 for fully coalesced access.
#include "pbx_internal.h"

namespace pbx {

namespace {

// fwd_sweep, src/tridsol.f90:76-96 (a sub-diagonal, b DIAGONAL <- pivots, c super-diagonal)
__global__ void __launch_bounds__(128)
fwd_kernel(int n, long long nl, long long es, long long ls, const double *__restrict__ a,
           double *__restrict__ b, const double *__restrict__ c, double *__restrict__ d)
{
    long long l = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (l >= nl) return;
    const long long o = l * ls;
    double bp = b[o], dp = d[o], cp = c[o];
    for (int i = 1; i < n; ++i) {
        const long long q = o + i * es;
        double w = __ddiv_rn(a[q], bp);
        bp = __dsub_rn(b[q], __dmul_rn(w, cp));
        dp = __dsub_rn(d[q], __dmul_rn(w, dp));
        cp = c[q];
        b[q] = bp;
        d[q] = dp;
    }
}

// bwd_sweep, src/tridsol.f90:98-115
__global__ void __launch_bounds__(128)
bwd_kernel(int n, long long nl, long long es, long long ls, const double *__restrict__ b,
           const double *__restrict__ c, double *__restrict__ d)
{
    long long l = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (l >= nl) return;
    const long long o = l * ls;
    long long q = o + (long long)(n - 1) * es;
    double x = __ddiv_rn(d[q], b[q]);
    d[q] = x;
    for (int i = n - 2; i >= 0; --i) {
        q = o + i * es;
        x = __ddiv_rn(__dsub_rn(d[q], __dmul_rn(c[q], x)), b[q]);
        d[q] = x;
    }
}

// tdma_periodic, src/tridsol.f90:34-74.  bmod and u live in a workspace laid out [i][line] so
// that the sweeps stay coalesced whatever the caller's strides are.
__global__ void __launch_bounds__(128)
periodic_kernel(int n, long long nl, long long es, long long ls, const double *__restrict__ a,
                const double *__restrict__ b, const double *__restrict__ c, double *__restrict__ d,
                double *__restrict__ bmod, double *__restrict__ u)
{
    long long l = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (l >= nl) return;
    const long long o = l * ls;
    const long long qn = o + (long long)(n - 1) * es;
    const double gamma = -b[o];                                      // :51
    const double a1 = a[o], cn = c[qn];
    const double b1m = __dsub_rn(b[o], gamma);                       // :55
    const double bnm = __dsub_rn(b[qn], __ddiv_rn(__dmul_rn(cn, a1), gamma));   // :56

    // two forward sweeps (:57 on d, :66 on u) share the pivots; pivots are stored for the
    // backward sweeps
    double bp = b1m, dp = d[o], up = gamma, cp = c[o];
    bmod[l] = bp;
    u[l] = up;
    for (int i = 1; i < n; ++i) {
        const long long q = o + i * es;
        double bi = (i == n - 1) ? bnm : b[q];
        double ui = (i == n - 1) ? cn : 0.0;                         // :63-65
        double w = __ddiv_rn(a[q], bp);
        bp = __dsub_rn(bi, __dmul_rn(w, cp));
        dp = __dsub_rn(d[q], __dmul_rn(w, dp));
        up = __dsub_rn(ui, __dmul_rn(w, up));
        cp = c[q];
        bmod[i * nl + l] = bp;
        u[i * nl + l] = up;
        d[q] = dp;
    }
    // backward sweeps
    double xd = __ddiv_rn(dp, bp), xu = __ddiv_rn(up, bp);
    d[qn] = xd;
    u[(long long)(n - 1) * nl + l] = xu;
    const double dn = xd, un = xu;
    for (int i = n - 2; i >= 0; --i) {
        const long long q = o + i * es;
        const double ci = c[q], bi = bmod[i * nl + l];
        xd = __ddiv_rn(__dsub_rn(d[q], __dmul_rn(ci, xd)), bi);
        xu = __ddiv_rn(__dsub_rn(u[i * nl + l], __dmul_rn(ci, xu)), bi);
        d[q] = xd;
        u[i * nl + l] = xu;
    }
    // :69-70
    const double a1g = __ddiv_rn(a1, gamma);
    const double fac = __dadd_rn(xd, __dmul_rn(a1g, dn));
    const double den = __dadd_rn(1.0, __dadd_rn(xu, __dmul_rn(a1g, un)));
    for (int i = 0; i < n; ++i) {
        const long long q = o + i * es;
        d[q] = __dsub_rn(d[q], __ddiv_rn(__dmul_rn(u[i * nl + l], fac), den));
    }
}

inline unsigned nblocks(long long nl, int bs) { return (unsigned)((nl + bs - 1) / bs); }

}  // namespace

int tdma_fwd_batch(cudaStream_t s, int n, long long nl, long long es, long long ls, const double *a,
                   double *b, const double *c, double *d)
{
    if (n < 1 || nl < 0) return PBX_ERR_ARG;
    if (nl == 0) return PBX_OK;
    fwd_kernel<<<nblocks(nl, 128), 128, 0, s>>>(n, nl, es, ls, a, b, c, d);
    PBX_CUDA(cudaGetLastError());
    return PBX_OK;
}

int tdma_bwd_batch(cudaStream_t s, int n, long long nl, long long es, long long ls, const double *b,
                   const double *c, double *d)
{
    if (n < 1 || nl < 0) return PBX_ERR_ARG;
    if (nl == 0) return PBX_OK;
    bwd_kernel<<<nblocks(nl, 128), 128, 0, s>>>(n, nl, es, ls, b, c, d);
    PBX_CUDA(cudaGetLastError());
    return PBX_OK;
}

int tdma_periodic_batch(cudaStream_t s, int n, long long nl, long long es, long long ls,
                        const double *a, const double *b, const double *c, double *d)
{
    if (n < 2 || nl < 0) return PBX_ERR_ARG;
    if (nl == 0) return PBX_OK;
    double *ws = nullptr;
    PBX_CUDA(cudaMallocAsync(&ws, sizeof(double) * 2 * (size_t)n * (size_t)nl, s));
    periodic_kernel<<<nblocks(nl, 128), 128, 0, s>>>(n, nl, es, ls, a, b, c, d, ws,
                                                     ws + (size_t)n * (size_t)nl);
    cudaError_t e = cudaGetLastError();
    cudaFreeAsync(ws, s);
    PBX_CUDA(e);
    return PBX_OK;
}

}  // namespace pbx
