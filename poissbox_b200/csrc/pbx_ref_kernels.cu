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

// One thread per line.  Line (l1, l2) starts at l1*ls1 + l2*ls2; element stride es.
// The three sweeps are serial recurrences; each walks its line in blocks of PF points and loads
// block k+1 into registers before it computes block k, so that the divide chain does not wait
// on memory (same operations in the same order: same bits).
constexpr int PF = 8;

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
    // rhs(i) = a (f(i+sh) + s f(i-1+sh)) + b (f(i+1+sh) + s f(i-2+sh)): a sliding window of four values
    auto at = [&](int i) {
        i += shift;
        if (i < 0) i += n;
        if (i >= n) i -= n;
        return f[(long long)i * es];
    };
    double fm2 = at(-2), fm1 = at(-1), f0 = at(0);      // f(i-2+sh), f(i-1+sh), f(i+sh) for i = 0
    double fn[PF], wn[PF];
#pragma unroll
    for (int k = 0; k < PF; ++k) {
        fn[k] = k < n ? at(k + 1) : 0.0;                // f(i+1+sh) for i = k
        wn[k] = (k >= 1 && k < n) ? __ldg(w + k) : 0.0;
    }
    double prev = 0.0;
    for (int i0 = 0; i0 < n; i0 += PF) {
        double fc[PF], wc[PF];
#pragma unroll
        for (int k = 0; k < PF; ++k) {
            fc[k] = fn[k];
            wc[k] = wn[k];
            const int i = i0 + PF + k;
            fn[k] = i < n ? at(i + 1) : 0.0;
            wn[k] = i < n ? __ldg(w + i) : 0.0;
        }
#pragma unroll
        for (int k = 0; k < PF; ++k) {
            const int i = i0 + k;
            if (i < n) {
                const double t1 = __dmul_rn(a, __dadd_rn(f0, sgn_mul(s, fm1)));
                const double t2 = __dmul_rn(b, __dadd_rn(fc[k], sgn_mul(s, fm2)));
                const double r = __dadd_rn(t1, t2);
                prev = i == 0 ? r : __dsub_rn(r, __dmul_rn(wc[k], prev));
                d[(long long)i * es] = prev;
                fm2 = fm1;
                fm1 = f0;
                f0 = fc[k];
            }
        }
    }
    // bwd_sweep (tridsol.f90:108-113); c(i) = alpha folded by the caller into `a1g`'s sibling
    const double alpha = -a1g;   // a(1)/gamma = alpha/(-1)
    double x = __ddiv_rn(prev, __ldg(piv + n - 1));
    d[(long long)(n - 1) * es] = x;
    const double dn = x;
    {
        double dnx[PF], pn[PF];
#pragma unroll
        for (int k = 0; k < PF; ++k) {
            const int i = n - 2 - k;
            dnx[k] = i >= 0 ? d[(long long)i * es] : 0.0;
            pn[k] = i >= 0 ? __ldg(piv + i) : 1.0;
        }
        for (int i0 = n - 2; i0 >= 0; i0 -= PF) {
            double dc[PF], pc[PF];
#pragma unroll
            for (int k = 0; k < PF; ++k) {
                dc[k] = dnx[k];
                pc[k] = pn[k];
                const int i = i0 - PF - k;
                dnx[k] = i >= 0 ? d[(long long)i * es] : 0.0;
                pn[k] = i >= 0 ? __ldg(piv + i) : 1.0;
            }
#pragma unroll
            for (int k = 0; k < PF; ++k) {
                const int i = i0 - k;
                if (i >= 0) {
                    x = __ddiv_rn(__dsub_rn(dc[k], __dmul_rn(alpha, x)), pc[k]);
                    d[(long long)i * es] = x;
                }
            }
        }
    }
    // Sherman-Morrison combine (tridsol.f90:69-70), old d(1), d(n) on the right-hand side
    const double fac = __dadd_rn(x, __dmul_rn(a1g, dn));
    for (int i0 = 0; i0 < n; i0 += PF) {
        double dc[PF], uc[PF];
#pragma unroll
        for (int k = 0; k < PF; ++k) {
            const int i = i0 + k;
            dc[k] = i < n ? d[(long long)i * es] : 0.0;
            uc[k] = i < n ? __ldg(u + i) : 0.0;
        }
#pragma unroll
        for (int k = 0; k < PF; ++k) {
            const int i = i0 + k;
            if (i < n) d[(long long)i * es] = __dsub_rn(dc[k], __ddiv_rn(__dmul_rn(uc[k], fac), den));
        }
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
