// pbx_tdma.cu -- batched general-coefficient tridiagonal solves, module tridsol of the reference
// (src/tridsol.f90:22-115), one thread per line.  Arithmetic is the reference's, operation for
// operation (round-to-nearest intrinsics: no FMA contraction, IEEE division), so a line solved here
// carries the same bits as the CPU oracle's.  Element i of line l is base[l*ls + i*es]; choose
// es = number of lines, ls = 1 for fully coalesced access.
#include <mutex>

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

// tdma_periodic, src/tridsol.f90:34-74, as THREE kernels over a workspace laid out [i][line] (so that the
// sweeps stay coalesced whatever the caller's strides are): both Thomas solves of the Sherman-Morrison
// closure -- on d and on u = (gamma, 0 .. 0, c(n)) -- share the pivots and run together.
//   per_fwd_kernel   reads a, b, c, d          writes bmod', u', d'            7 words / point
//   per_bwd_kernel   reads c, bmod', d', u'    writes y (in d), q (in u)       6 words / point; fac, den per line
//   per_cmb_kernel   reads y, q                writes d = y - q fac / den      3 words / point, a thread per POINT
// 128 B / point against the 48 B / point of SURVEY 8(d): the reference's algorithm is a serial recurrence with
// a division per point, so every line must be in flight at once and its intermediate state (pivot, reduced
// d, reduced u: 24 B / point) cannot stay on chip.  One kernel doing all three legs per thread (round 1) needed
// 184 registers and ran its legs back to back at 8 warps per SM: 3.6 TB/s on the moved bytes for 512-point
// lines; the legs as separate kernels move the same bytes at the rate of the plain tdma sweeps.
__global__ void __launch_bounds__(128)
per_fwd_kernel(int n, long long nl, long long es, long long ls, const double *__restrict__ a,
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
    double bp = b1m, dp = d[o], up = gamma, cp = c[o];
    bmod[l] = bp;
    u[l] = up;
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

// backward sweeps of both systems; facden[l], facden[nl + l] = the two scalars of :69-70
__global__ void __launch_bounds__(128)
per_bwd_kernel(int n, long long nl, long long es, long long ls, const double *__restrict__ a,
               const double *__restrict__ b, const double *__restrict__ c, double *__restrict__ d,
               const double *__restrict__ bmod, double *__restrict__ u, double *__restrict__ facden)
{
    long long l = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (l >= nl) return;
    const long long o = l * ls;
    const long long qn = o + (long long)(n - 1) * es;
    const long long wn = (long long)(n - 1) * nl + l;
    double xd = __ddiv_rn(d[qn], bmod[wn]), xu = __ddiv_rn(u[wn], bmod[wn]);
    d[qn] = xd;
    u[wn] = xu;
    const double dn_ = xd, un = xu;
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
    // :69-70 -- xd, xu now hold the solutions at point 1 of the line
    const double gamma = -b[o];
    const double a1g = __ddiv_rn(a[o], gamma);
    facden[l] = __dadd_rn(xd, __dmul_rn(a1g, dn_));
    facden[nl + l] = __dadd_rn(1.0, __dadd_rn(xu, __dmul_rn(a1g, un)));
}

// d = y - (q * fac) / den, a thread per point (lines fastest: coalesced in the workspace, and in the
// caller's array when its lines are the fast index)
__global__ void __launch_bounds__(256)
per_cmb_kernel(int n, long long nl, long long es, long long ls, double *__restrict__ d,
               const double *__restrict__ u, const double *__restrict__ facden)
{
    const long long total = (long long)n * nl;
    for (long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x; t < total;
         t += (long long)gridDim.x * blockDim.x) {
        const long long i = t / nl, l = t - i * nl;
        const long long q = l * ls + i * es;
        d[q] = __dsub_rn(d[q], __ddiv_rn(__dmul_rn(u[t], facden[l]), facden[nl + l]));
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

// Stream-ordered workspace from a pool of the library's own, whose release threshold keeps freed blocks
// cached: with the device's default pool every synchronisation handed the 2 n nl doubles back to the driver
// and the next call paid ~0.3 ms to map them again (measured: 0.65 ms per call around a 0.34 ms kernel).
// pbx_host_cache_clear() trims it.
static cudaMemPool_t g_pool[64];
static std::mutex g_pool_mutex;

static int ws_pool(cudaMemPool_t *pool)
{
    int dev = 0;
    PBX_CUDA(cudaGetDevice(&dev));
    std::lock_guard<std::mutex> lk(g_pool_mutex);
    if (!g_pool[dev & 63]) {
        cudaMemPoolProps props = {};
        props.allocType = cudaMemAllocationTypePinned;
        props.handleTypes = cudaMemHandleTypeNone;
        props.location.type = cudaMemLocationTypeDevice;
        props.location.id = dev;
        PBX_CUDA(cudaMemPoolCreate(&g_pool[dev & 63], &props));
        unsigned long long keep = ~0ull;
        PBX_CUDA(cudaMemPoolSetAttribute(g_pool[dev & 63], cudaMemPoolAttrReleaseThreshold, &keep));
    }
    *pool = g_pool[dev & 63];
    return PBX_OK;
}

void tdma_trim_workspace()
{
    std::lock_guard<std::mutex> lk(g_pool_mutex);
    for (auto &p : g_pool)
        if (p) cudaMemPoolTrimTo(p, 0);
}

int tdma_periodic_batch(cudaStream_t s, int n, long long nl, long long es, long long ls,
                        const double *a, const double *b, const double *c, double *d)
{
    if (n < 2 || nl < 0) return PBX_ERR_ARG;
    if (nl == 0) return PBX_OK;
    cudaMemPool_t pool;
    PBX_TRY(ws_pool(&pool));
    double *ws = nullptr;
    const size_t pts = (size_t)n * (size_t)nl;
    PBX_CUDA(cudaMallocFromPoolAsync(&ws, sizeof(double) * (2 * pts + 2 * (size_t)nl), pool, s));
    int rc = tdma_periodic_batch_lm(s, n, nl, es, ls, a, b, c, d, ws);
    if (rc == PBX_ERR_UNSUPPORTED) {
        double *bmod = ws, *u = ws + pts, *facden = ws + 2 * pts;
        per_fwd_kernel<<<nblocks(nl, 128), 128, 0, s>>>(n, nl, es, ls, a, b, c, d, bmod, u);
        per_bwd_kernel<<<nblocks(nl, 128), 128, 0, s>>>(n, nl, es, ls, a, b, c, d, bmod, u, facden);
        const long long nb = (long long)((pts + 255) / 256);
        per_cmb_kernel<<<(unsigned)(nb < 148 * 16 ? nb : 148 * 16), 256, 0, s>>>(n, nl, es, ls, d, u, facden);
        rc = cudaGetLastError() == cudaSuccess ? PBX_OK : PBX_ERR_CUDA;
        if (rc != PBX_OK) set_last_error("tdma_periodic kernels failed to launch");
    }
    cudaFreeAsync(ws, s);
    return rc;
}

}  // namespace pbx
