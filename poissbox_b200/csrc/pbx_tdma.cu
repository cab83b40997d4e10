// pbx_tdma.cu -- batched general-coefficient tridiagonal solves, module tridsol of the reference
// (src/tridsol.f90:22-115), one thread per line.  Arithmetic is the reference's, operation for
// operation (round-to-nearest intrinsics: no FMA contraction, IEEE division), so a line solved here
// carries the same bits as the CPU oracle's.  Element i of line l is base[l*ls + i*es]; choose
// es = number of lines, ls = 1 for fully coalesced access.
#include "pbx_internal.h"

namespace pbx {

namespace {

// The sweeps are serial recurrences with a division per point, so a thread can only hide the
// latency of its global loads by issuing them ahead: every sweep walks its line in blocks of PF
// points and loads block k+1 into registers before it computes block k (same operations in the
// same order as the plain loop -- the results carry the same bits).
constexpr int PF = 8;

// fwd_sweep, src/tridsol.f90:76-96 (a sub-diagonal, b DIAGONAL <- pivots, c super-diagonal)
__global__ void __launch_bounds__(128)
fwd_kernel(int n, long long nl, long long es, long long ls, const double *__restrict__ a,
           double *__restrict__ b, const double *__restrict__ c, double *__restrict__ d)
{
    long long l = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (l >= nl) return;
    const long long o = l * ls;
    double bp = b[o], dp = d[o], cp = c[o];
    double an[PF], bn[PF], cn[PF], dn[PF];
#pragma unroll
    for (int u = 0; u < PF; ++u) {
        const long long q = o + (long long)(1 + u) * es;
        const bool in = 1 + u < n;
        an[u] = in ? a[q] : 0.0;
        bn[u] = in ? b[q] : 1.0;
        cn[u] = in ? c[q] : 0.0;
        dn[u] = in ? d[q] : 0.0;
    }
    for (int i0 = 1; i0 < n; i0 += PF) {
        double ac[PF], bc[PF], cc[PF], dc[PF];
#pragma unroll
        for (int u = 0; u < PF; ++u) {
            ac[u] = an[u];
            bc[u] = bn[u];
            cc[u] = cn[u];
            dc[u] = dn[u];
            const int i = i0 + PF + u;
            const long long q = o + (long long)i * es;
            const bool in = i < n;
            an[u] = in ? a[q] : 0.0;
            bn[u] = in ? b[q] : 1.0;
            cn[u] = in ? c[q] : 0.0;
            dn[u] = in ? d[q] : 0.0;
        }
#pragma unroll
        for (int u = 0; u < PF; ++u) {
            const int i = i0 + u;
            if (i < n) {
                const long long q = o + (long long)i * es;
                double w = __ddiv_rn(ac[u], bp);
                bp = __dsub_rn(bc[u], __dmul_rn(w, cp));
                dp = __dsub_rn(dc[u], __dmul_rn(w, dp));
                cp = cc[u];
                b[q] = bp;
                d[q] = dp;
            }
        }
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
    double bn[PF], cn[PF], dn[PF];
#pragma unroll
    for (int u = 0; u < PF; ++u) {
        const int i = n - 2 - u;
        const long long qq = o + (long long)i * es;
        const bool in = i >= 0;
        bn[u] = in ? b[qq] : 1.0;
        cn[u] = in ? c[qq] : 0.0;
        dn[u] = in ? d[qq] : 0.0;
    }
    for (int i0 = n - 2; i0 >= 0; i0 -= PF) {
        double bc[PF], cc[PF], dc[PF];
#pragma unroll
        for (int u = 0; u < PF; ++u) {
            bc[u] = bn[u];
            cc[u] = cn[u];
            dc[u] = dn[u];
            const int i = i0 - PF - u;
            const long long qq = o + (long long)i * es;
            const bool in = i >= 0;
            bn[u] = in ? b[qq] : 1.0;
            cn[u] = in ? c[qq] : 0.0;
            dn[u] = in ? d[qq] : 0.0;
        }
#pragma unroll
        for (int u = 0; u < PF; ++u) {
            const int i = i0 - u;
            if (i >= 0) {
                x = __ddiv_rn(__dsub_rn(dc[u], __dmul_rn(cc[u], x)), bc[u]);
                d[o + (long long)i * es] = x;
            }
        }
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
    const double a1 = a[o], cn_ = c[qn];
    const double b1m = __dsub_rn(b[o], gamma);                       // :55
    const double bnm = __dsub_rn(b[qn], __ddiv_rn(__dmul_rn(cn_, a1), gamma));   // :56

    // two forward sweeps (:57 on d, :66 on u) share the pivots; pivots are stored for the
    // backward sweeps
    double bp = b1m, dp = d[o], up = gamma, cp = c[o];
    bmod[l] = bp;
    u[l] = up;
    {
        double an[PF], bn[PF], cn[PF], dn[PF];
#pragma unroll
        for (int k = 0; k < PF; ++k) {
            const long long q = o + (long long)(1 + k) * es;
            const bool in = 1 + k < n;
            an[k] = in ? a[q] : 0.0;
            bn[k] = in ? b[q] : 1.0;
            cn[k] = in ? c[q] : 0.0;
            dn[k] = in ? d[q] : 0.0;
        }
        for (int i0 = 1; i0 < n; i0 += PF) {
            double ac[PF], bc[PF], cc[PF], dc[PF];
#pragma unroll
            for (int k = 0; k < PF; ++k) {
                ac[k] = an[k];
                bc[k] = bn[k];
                cc[k] = cn[k];
                dc[k] = dn[k];
                const int i = i0 + PF + k;
                const long long q = o + (long long)i * es;
                const bool in = i < n;
                an[k] = in ? a[q] : 0.0;
                bn[k] = in ? b[q] : 1.0;
                cn[k] = in ? c[q] : 0.0;
                dn[k] = in ? d[q] : 0.0;
            }
#pragma unroll
            for (int k = 0; k < PF; ++k) {
                const int i = i0 + k;
                if (i < n) {
                    const long long q = o + (long long)i * es;
                    double bi = (i == n - 1) ? bnm : bc[k];
                    double ui = (i == n - 1) ? cn_ : 0.0;                // :63-65
                    double w = __ddiv_rn(ac[k], bp);
                    bp = __dsub_rn(bi, __dmul_rn(w, cp));
                    dp = __dsub_rn(dc[k], __dmul_rn(w, dp));
                    up = __dsub_rn(ui, __dmul_rn(w, up));
                    cp = cc[k];
                    bmod[i * nl + l] = bp;
                    u[i * nl + l] = up;
                    d[q] = dp;
                }
            }
        }
    }
    // backward sweeps
    double xd = __ddiv_rn(dp, bp), xu = __ddiv_rn(up, bp);
    d[qn] = xd;
    u[(long long)(n - 1) * nl + l] = xu;
    const double dn_ = xd, un = xu;
    {
        double cn[PF], bn[PF], dn[PF], vn[PF];
#pragma unroll
        for (int k = 0; k < PF; ++k) {
            const int i = n - 2 - k;
            const long long q = o + (long long)i * es;
            const bool in = i >= 0;
            cn[k] = in ? c[q] : 0.0;
            bn[k] = in ? bmod[i * nl + l] : 1.0;
            dn[k] = in ? d[q] : 0.0;
            vn[k] = in ? u[i * nl + l] : 0.0;
        }
        for (int i0 = n - 2; i0 >= 0; i0 -= PF) {
            double cc[PF], bc[PF], dc[PF], vc[PF];
#pragma unroll
            for (int k = 0; k < PF; ++k) {
                cc[k] = cn[k];
                bc[k] = bn[k];
                dc[k] = dn[k];
                vc[k] = vn[k];
                const int i = i0 - PF - k;
                const long long q = o + (long long)i * es;
                const bool in = i >= 0;
                cn[k] = in ? c[q] : 0.0;
                bn[k] = in ? bmod[i * nl + l] : 1.0;
                dn[k] = in ? d[q] : 0.0;
                vn[k] = in ? u[i * nl + l] : 0.0;
            }
#pragma unroll
            for (int k = 0; k < PF; ++k) {
                const int i = i0 - k;
                if (i >= 0) {
                    xd = __ddiv_rn(__dsub_rn(dc[k], __dmul_rn(cc[k], xd)), bc[k]);
                    xu = __ddiv_rn(__dsub_rn(vc[k], __dmul_rn(cc[k], xu)), bc[k]);
                    d[o + (long long)i * es] = xd;
                    u[i * nl + l] = xu;
                }
            }
        }
    }
    // :69-70
    const double a1g = __ddiv_rn(a1, gamma);
    const double fac = __dadd_rn(xd, __dmul_rn(a1g, dn_));
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
    const int rc = tdma_fwd_batch_lm(s, n, nl, es, ls, a, b, c, d);
    if (rc != PBX_ERR_UNSUPPORTED) return rc;
    fwd_kernel<<<nblocks(nl, 128), 128, 0, s>>>(n, nl, es, ls, a, b, c, d);
    PBX_CUDA(cudaGetLastError());
    return PBX_OK;
}

int tdma_bwd_batch(cudaStream_t s, int n, long long nl, long long es, long long ls, const double *b,
                   const double *c, double *d)
{
    if (n < 1 || nl < 0) return PBX_ERR_ARG;
    if (nl == 0) return PBX_OK;
    const int rc = tdma_bwd_batch_lm(s, n, nl, es, ls, b, c, d);
    if (rc != PBX_ERR_UNSUPPORTED) return rc;
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
    const int rc = tdma_periodic_batch_lm(s, n, nl, es, ls, a, b, c, d, ws);
    if (rc != PBX_ERR_UNSUPPORTED) {
        cudaFreeAsync(ws, s);
        return rc;
    }
    periodic_kernel<<<nblocks(nl, 128), 128, 0, s>>>(n, nl, es, ls, a, b, c, d, ws,
                                                     ws + (size_t)n * (size_t)nl);
    cudaError_t e = cudaGetLastError();
    cudaFreeAsync(ws, s);
    PBX_CUDA(e);
    return PBX_OK;
}

}  // namespace pbx
