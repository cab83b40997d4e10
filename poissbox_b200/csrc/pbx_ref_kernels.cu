// pbx_ref_kernels.cu -- REFERENCE-order schedule: every 1-D compact operator is the reference's
// own sequence (eval_1d_rhs, then tdma_periodic = Thomas forward/backward sweep + Sherman-Morrison
// combine), one thread per line, lines coalesced across threads.  All floating-point operations use
// the round-to-nearest intrinsics (__dadd_rn, __dmul_rn, __ddiv_rn), which nvcc never contracts
// into FMAs, so the results are bit-identical to the CPU oracle's.
//
// Follows: src/compact_schemes.f90:332-372 (eval_1d_rhs), :155-204 (grad_1d), :271-319
// (interp_1d); src/tridsol.f90:34-115 (tdma_periodic, fwd_sweep, bwd_sweep).
#include "pbx_internal.h"

namespace pbx {

namespace {

__device__ __forceinline__ double sgn_mul(int s, double v) { return s > 0 ? v : -v; }

// rhs(i) of eval_1d_rhs (0-based i), periodic wrap (:356-370)
__device__ __forceinline__ double rhs_at(const double *__restrict__ f, long long es, int n, int i,
                                         int shift, int s, double a, double b)
{
    // stagger -1 (shift 0): a (f(i) + s f(i-1)) + b (f(i+1) + s f(i-2))
    // stagger +1 (shift 1): a (f(i+1) + s f(i)) + b (f(i+2) + s f(i-1))
    int i0 = i + shift, i1 = i - 1 + shift, i2 = i + 1 + shift, i3 = i - 2 + shift;
    if (i0 >= n) i0 -= n;
    if (i1 < 0) i1 += n;
    if (i2 >= n) i2 -= n;
    if (i3 < 0) i3 += n;
    double t1 = __dmul_rn(a, __dadd_rn(f[i0 * es], sgn_mul(s, f[i1 * es])));
    double t2 = __dmul_rn(b, __dadd_rn(f[i2 * es], sgn_mul(s, f[i3 * es])));
    return __dadd_rn(t1, t2);
}

// One thread per line.  Line (l1, l2) starts at l1*ls1 + l2*ls2; element stride es.
__global__ void __launch_bounds__(128)
ref_line_kernel(int n, long long nl1, long long nl2, long long es, long long ls1, long long ls2,
                int shift, int s, double a, double b, const double *__restrict__ w,
                const double *__restrict__ piv, const double *__restrict__ u, double a1g, double den,
                const double *__restrict__ in, double *__restrict__ out)
{
    long long l1 = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    long long l2 = blockIdx.y;
    if (l1 >= nl1 || l2 >= nl2) return;
    const double *f = in + l1 * ls1 + l2 * ls2;
    double *d = out + l1 * ls1 + l2 * ls2;

    // eval_1d_rhs fused with fwd_sweep's data recurrence d(i) = d(i) - w d(i-1)   (tridsol.f90:93)
    double prev = rhs_at(f, es, n, 0, shift, s, a, b);
    d[0] = prev;
    for (int i = 1; i < n; ++i) {
        double r = rhs_at(f, es, n, i, shift, s, a, b);
        prev = __dsub_rn(r, __dmul_rn(__ldg(w + i), prev));
        d[i * es] = prev;
    }
    // bwd_sweep (tridsol.f90:108-113); c(i) = alpha folded by the caller into `a1g`'s sibling
    const double alpha = -a1g;   // a(1)/gamma = alpha/(-1)
    double x = __ddiv_rn(prev, __ldg(piv + n - 1));
    d[(long long)(n - 1) * es] = x;
    const double dn = x;
    for (int i = n - 2; i >= 0; --i) {
        double di = d[i * es];
        x = __ddiv_rn(__dsub_rn(di, __dmul_rn(alpha, x)), __ldg(piv + i));
        d[i * es] = x;
    }
    // Sherman-Morrison combine (tridsol.f90:69-70), old d(1), d(n) on the right-hand side
    const double fac = __dadd_rn(x, __dmul_rn(a1g, dn));
    for (int i = 0; i < n; ++i) {
        double di = d[i * es];
        d[i * es] = __dsub_rn(di, __ddiv_rn(__dmul_rn(__ldg(u + i), fac), den));
    }
}

__global__ void add_kernel(size_t N, const double *__restrict__ a, const double *__restrict__ b,
                           double *__restrict__ o)
{
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    size_t st = (size_t)gridDim.x * blockDim.x;
    for (; i < N; i += st) o[i] = __dadd_rn(a[i], b[i]);
}

}  // namespace

int ref_line_op(cudaStream_t st, int n, long long nl1, long long nl2, long long es, long long ls1,
                long long ls2, OpKind kind, int stagger, double dx, const RefLineTables &tab,
                const double *in, double *out, long long *launches)
{
    if (n < 3 || tab.n != n) {
        set_last_error("ref_line_op: line length must be >= 3 and match the table");
        return PBX_ERR_ARG;
    }
    double a, b;
    scheme_ab(kind, dx, &a, &b);
    const int s = (kind == OP_DERIV) ? -1 : +1;
    const int shift = (stagger == PBX_STAGGER_BACKWARD) ? 0 : 1;
    // gridDim.y is limited to 65535: fold the outer line index when it is larger
    const int bs = 128;
    for (long long l2 = 0; l2 < nl2; l2 += 65535) {
        long long cnt = nl2 - l2 < 65535 ? nl2 - l2 : 65535;
        dim3 grid((unsigned)((nl1 + bs - 1) / bs), (unsigned)cnt);
        ref_line_kernel<<<grid, bs, 0, st>>>(n, nl1, cnt, es, ls1, ls2, shift, s, a, b, tab.w,
                                             tab.piv, tab.u, tab.a1g, tab.den, in + l2 * ls2,
                                             out + l2 * ls2);
        if (launches) ++*launches;
    }
    PBX_CUDA(cudaGetLastError());
    return PBX_OK;
}

int ref_add(cudaStream_t st, size_t N, const double *a, const double *b, double *out,
            long long *launches)
{
    int bs = 256;
    size_t nb = (N + bs - 1) / bs;
    if (nb > 148 * 16) nb = 148 * 16;
    add_kernel<<<(unsigned)nb, bs, 0, st>>>(N, a, b, out);
    if (launches) ++*launches;
    PBX_CUDA(cudaGetLastError());
    return PBX_OK;
}

}  // namespace pbx
