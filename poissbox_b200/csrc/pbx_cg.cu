// pbx_cg.cu -- conjugate gradients on the compact Laplacian, resident on the device.
//
// Semantics: PETSc KSPCG as the reference's solve() configures it (src/poissbox.f90:269-298:
// constant MatNullSpace attached to A, KSPSetFromOptions, KSPSolve) with `-ksp_type cg
// -pc_type none`: zero initial guess, z = r - mean(r) (the "preconditioner apply" followed by
// null-space removal), preconditioned norm ||z||_2, beta = z.r, default convergence test.
// PETSc is third-party and not vendored by the reference; the loop below follows its published
// algorithm; the test-side CPU restatement of the same loop lives under oracle/.
//
// Per iteration (PC none) the vector work is two fused kernels,
//   update   r -= a w ; partial sums of (r - m0), (r - m0)^2                    24 B/DoF
//   pupdate  x += a p ; p = (r - m) + b p                                      40 B/DoF
// and p.w comes out of the Laplacian's z pass (which reads p for it: 8 B/DoF), so an iteration moves
// 88 + 64 = 152 B/DoF.  (x += a p rides with the p update, which reads p anyway, instead of with the
// r update: 8 B/DoF less than the textbook grouping.  The iteration that converges still owes x its update when the status word is
// already set; k_pupdate_x recognises it by the iteration stamp SC_XIT that the scalar step leaves
// whenever it computes a new step length.)
// All scalars (a, b, mean, norms, status) stay in a device block; the host only reads back the
// status word, one iteration late, so there is no host synchronisation on the critical path.
// Reductions are two-stage with a fixed shape (per-CTA partials, then one CTA), so the iteration
// count is reproducible run to run.
//
// Because mean(r) is invariant in exact arithmetic, the sums are taken about the mean of b (m0):
//   mean = m0 + S1/N ,  ||z||^2 = S2 - S1^2/N   with S1 = sum(r - m0), S2 = sum (r - m0)^2,
// which avoids the cancellation a raw sum(r^2) - N mean^2 would suffer when mean(b) != 0.
#include <cmath>
#include <cstdlib>

#include "pbx_cg_dev.cuh"

namespace pbx {

using namespace cgdev;

namespace {

// (iteration number, status) in one 8-byte word: the host must never see one without the other
__host__ __device__ inline unsigned long long status_word(int itag, double status)
{
    return ((unsigned long long)(unsigned)itag << 16) | (unsigned long long)(unsigned)((int)status + 1024);
}

// each CTA owns a contiguous slice so that the summation order is fixed
__device__ __forceinline__ void slice(size_t N, size_t *lo, size_t *hi)
{
    size_t per = (N + gridDim.x - 1) / gridDim.x;
    per = (per + 1) & ~(size_t)1;
    *lo = per * blockIdx.x;
    *hi = *lo + per < N ? *lo + per : N;
    if (*lo > N) *lo = N;
}

__global__ void __launch_bounds__(VT) k_sum(size_t N, const double *__restrict__ b,
                                            double *__restrict__ part)
{
    __shared__ double sh[VT / 32];
    size_t lo, hi;
    slice(N, &lo, &hi);
    double s = 0.0;
    for (size_t i = lo + threadIdx.x; i < hi; i += VT) s += b[i];
    s = block_sum(s, sh);
    if (threadIdx.x == 0) part[blockIdx.x] = s;
}

// r = b ; x = 0 ; p = b - m0 ; partial sums about m0
__global__ void __launch_bounds__(VT)
k_init(size_t N, const double *__restrict__ b, double *__restrict__ x, double *__restrict__ r,
       double *__restrict__ p, const double *__restrict__ sc, double *__restrict__ part, int np)
{
    __shared__ double sh[VT / 32];
    size_t lo, hi;
    slice(N, &lo, &hi);
    const double m0 = sc[SC_M0];
    double s1 = 0.0, s2 = 0.0;
    for (size_t i = lo + threadIdx.x; i < hi; i += VT) {
        double v = b[i];
        double t = v - m0;
        x[i] = 0.0;
        r[i] = v;
        p[i] = t;
        s1 += t;
        s2 = fma(t, t, s2);
    }
    s1 = block_sum(s1, sh);
    s2 = block_sum(s2, sh);
    if (threadIdx.x == 0) {
        part[blockIdx.x] = s1;
        part[np + blockIdx.x] = s2;
    }
}

// x += a p ; r -= a w ; partial sums of r - m0
__global__ void __launch_bounds__(VT)
k_update(size_t N, double *__restrict__ x, double *__restrict__ r, const double *__restrict__ p,
         const double *__restrict__ w, const double *__restrict__ sc, double *__restrict__ part,
         int np)
{
    __shared__ double sh[VT / 32];
    if (sc[SC_STATUS] != 0.0) return;
    size_t lo, hi;
    slice(N, &lo, &hi);
    const double a = sc[SC_A], m0 = sc[SC_M0];
    double s1 = 0.0, s2 = 0.0;
    // two elements per thread per trip: slices start at even indices and the fields are 16-byte
    // aligned, so the double2 accesses are aligned
    size_t i = lo + 2 * (size_t)threadIdx.x;
    for (; i + 1 < hi; i += 2 * VT) {
        double2 xv = *reinterpret_cast<const double2 *>(x + i);
        double2 rv = *reinterpret_cast<const double2 *>(r + i);
        double2 pv = *reinterpret_cast<const double2 *>(p + i);
        double2 wv = *reinterpret_cast<const double2 *>(w + i);
        xv.x = fma(a, pv.x, xv.x);
        xv.y = fma(a, pv.y, xv.y);
        rv.x = fma(-a, wv.x, rv.x);
        rv.y = fma(-a, wv.y, rv.y);
        *reinterpret_cast<double2 *>(x + i) = xv;
        *reinterpret_cast<double2 *>(r + i) = rv;
        double t0 = rv.x - m0, t1 = rv.y - m0;
        s1 += t0 + t1;
        s2 = fma(t0, t0, fma(t1, t1, s2));
    }
    if (i < hi) {
        double xv = fma(a, p[i], x[i]);
        double rv = fma(-a, w[i], r[i]);
        x[i] = xv;
        r[i] = rv;
        double t = rv - m0;
        s1 += t;
        s2 = fma(t, t, s2);
    }
    s1 = block_sum(s1, sh);
    s2 = block_sum(s2, sh);
    if (threadIdx.x == 0) {
        part[blockIdx.x] = s1;
        part[np + blockIdx.x] = s2;
    }
}

// r -= a w ; partial sums of r - m0   (PC none: x is updated together with p, k_pupdate_x)
__global__ void __launch_bounds__(VT)
k_update_r(size_t N, double *__restrict__ r, const double *__restrict__ w, const double *__restrict__ sc,
           double *__restrict__ part, int np, const __grid_constant__ RedTail tail)
{
    __shared__ double sh[VT / 32];
    if (sc[SC_STATUS] != 0.0) return;
    size_t lo, hi;
    slice(N, &lo, &hi);
    const double a = sc[SC_A], m0 = sc[SC_M0];
    double s1 = 0.0, s2 = 0.0;
    size_t i = lo + 2 * (size_t)threadIdx.x;
    for (; i + 1 < hi; i += 2 * VT) {
        double2 rv = *reinterpret_cast<const double2 *>(r + i);
        const double2 wv = *reinterpret_cast<const double2 *>(w + i);
        rv.x = fma(-a, wv.x, rv.x);
        rv.y = fma(-a, wv.y, rv.y);
        *reinterpret_cast<double2 *>(r + i) = rv;
        const double t0 = rv.x - m0, t1 = rv.y - m0;
        s1 += t0 + t1;
        s2 = fma(t0, t0, fma(t1, t1, s2));
    }
    if (i < hi) {
        const double rv = fma(-a, w[i], r[i]);
        r[i] = rv;
        const double t = rv - m0;
        s1 += t;
        s2 = fma(t, t, s2);
    }
    s1 = block_sum(s1, sh);
    s2 = block_sum(s2, sh);
    if (threadIdx.x == 0) {
        part[blockIdx.x] = s1;
        part[np + blockIdx.x] = s2;
    }
    // PBX_FUSE_TAIL: the CTA that finishes last sums the partials, all-reduces them and runs the scalar step
    if (tail.on) red_tail<VT>(tail);
}

// x += a p (iteration `itag`, if its step length was computed: also when that iteration converged or
// hit a limit) ; p = (r - mean) + b p (while the solve is still running)
// Thread 0 also posts (iteration number, status) as ONE word straight into the host's pinned status word: the host
// follows the solve one iteration behind by reading that word, with no copy and no event in the stream (a
// stream-ordered 200-byte cudaMemcpyAsync between two kernels cost ~5 us of idle GPU per iteration).
__global__ void __launch_bounds__(VT)
k_pupdate_x(size_t N, const double *__restrict__ r, double *__restrict__ p, double *__restrict__ x,
            const double *__restrict__ sc, int itag, unsigned long long *__restrict__ host_word)
{
    if (host_word && blockIdx.x == 0 && threadIdx.x == 0)
        *(volatile unsigned long long *)host_word = status_word(itag, sc[SC_STATUS]);
    const bool dox = sc[SC_XIT] == (double)itag, dop = sc[SC_STATUS] == 0.0;
    if (!dox && !dop) return;
    const double a = sc[SC_A], m = sc[SC_MEAN], b = sc[SC_B];
    const size_t st = (size_t)gridDim.x * VT * 2;
    size_t i = ((size_t)blockIdx.x * VT + threadIdx.x) * 2;
    for (; i + 1 < N; i += st) {
        double2 pv = *reinterpret_cast<const double2 *>(p + i);
        if (dox) {
            double2 xv = *reinterpret_cast<const double2 *>(x + i);
            xv.x = fma(a, pv.x, xv.x);
            xv.y = fma(a, pv.y, xv.y);
            *reinterpret_cast<double2 *>(x + i) = xv;
        }
        if (dop) {
            const double2 rv = *reinterpret_cast<const double2 *>(r + i);
            pv.x = fma(b, pv.x, rv.x - m);
            pv.y = fma(b, pv.y, rv.y - m);
            *reinterpret_cast<double2 *>(p + i) = pv;
        }
    }
    if (i < N) {
        const double pv = p[i];
        if (dox) x[i] = fma(a, pv, x[i]);
        if (dop) p[i] = fma(b, pv, r[i] - m);
    }
}

// sum `cnt` partials of each of `narr` arrays (`stride` apart) into dst[0..narr) (one CTA, fixed order)
__global__ void __launch_bounds__(VT)
k_reduce(const double *__restrict__ part, int cnt, int stride, int narr, double *__restrict__ dst,
         const double *__restrict__ sc, int guarded)
{
    __shared__ double sh[VT / 32];
    if (guarded && sc[SC_STATUS] != 0.0) return;
    for (int a = 0; a < narr; ++a) {
        double s = 0.0;
        for (int i = threadIdx.x; i < cnt; i += VT) s += part[a * stride + i];
        s = block_sum(s, sh);
        if (threadIdx.x == 0) dst[a] = s;
        __syncthreads();
    }
}

// preconditioned CG: partial sums of z, z^2, z (r - m0), (r - m0) into four arrays `np` apart
__global__ void __launch_bounds__(VT)
k_pcdots(size_t N, const double *__restrict__ z, const double *__restrict__ r,
         const double *__restrict__ sc, double *__restrict__ part, int np)
{
    __shared__ double sh[VT / 32];
    size_t lo, hi;
    slice(N, &lo, &hi);
    const double m0 = sc[SC_M0];
    double a0 = 0.0, a1 = 0.0, a2 = 0.0, a3 = 0.0;
    for (size_t i = lo + threadIdx.x; i < hi; i += VT) {
        const double zv = z[i], t = r[i] - m0;
        a0 += zv;
        a1 = fma(zv, zv, a1);
        a2 = fma(zv, t, a2);
        a3 += t;
    }
    a0 = block_sum(a0, sh);
    a1 = block_sum(a1, sh);
    a2 = block_sum(a2, sh);
    a3 = block_sum(a3, sh);
    if (threadIdx.x == 0) {
        part[blockIdx.x] = a0;
        part[np + blockIdx.x] = a1;
        part[2 * np + blockIdx.x] = a2;
        part[3 * np + blockIdx.x] = a3;
    }
}

// p = (z - mean z) + b p   (b = 0 on the first iteration)
__global__ void __launch_bounds__(VT)
k_pupdate_pc(size_t N, const double *__restrict__ z, double *__restrict__ p,
             const double *__restrict__ sc)
{
    if (sc[SC_STATUS] != 0.0) return;
    const double m = sc[SC_MZ], b = sc[SC_B];
    size_t i = (size_t)blockIdx.x * VT + threadIdx.x;
    const size_t st = (size_t)gridDim.x * VT;
    for (; i < N; i += st) p[i] = b == 0.0 ? z[i] - m : fma(b, p[i], z[i] - m);
}

// v -= *m
__global__ void __launch_bounds__(VT)
k_center(size_t N, double *__restrict__ v, const double *__restrict__ m)
{
    const double mm = *m;
    size_t i = (size_t)blockIdx.x * VT + threadIdx.x;
    const size_t st = (size_t)gridDim.x * VT;
    for (; i < N; i += st) v[i] -= mm;
}

// r = b ; x = 0 (preconditioned CG: p comes from the first preconditioner application)
__global__ void __launch_bounds__(VT)
k_init_pc(size_t N, const double *__restrict__ b, double *__restrict__ x, double *__restrict__ r)
{
    size_t i = (size_t)blockIdx.x * VT + threadIdx.x;
    const size_t st = (size_t)gridDim.x * VT;
    for (; i < N; i += st) {
        x[i] = 0.0;
        r[i] = b[i];
    }
}

__global__ void k_scalar(double *__restrict__ sc, int phase, double *__restrict__ hist, int nhist)
{
    if (threadIdx.x != 0 || blockIdx.x != 0) return;
    scalar_phase(sc, phase, hist, nhist);
}

// Multi-rank CG over the peer boards: ONE kernel sums the per-CTA partials of `narr` arrays, all-
// reduces the sums with the other ranks over NVLink (peer_exchange_sum: same bits on every rank)
// and runs the scalar step `phase` of the loop on them -- what k_reduce + ncclAllReduce + k_scalar
// do in three launches.  narr <= PEER_VALS; dst points into the scalar block.
__global__ void __launch_bounds__(VT)
k_reduce_peer(const double *__restrict__ part, int cnt, int stride, int narr, double *__restrict__ dst,
              double *__restrict__ sc, int guarded, const __grid_constant__ PeerLinks L,
              unsigned long long seq, int phase, double *__restrict__ hist, int nhist)
{
    __shared__ double sh[VT / 32];
    __shared__ double mine[PEER_VALS];
    __shared__ double all[PEER_MAXR][PEER_VALS];
    // the status word is the same on every rank (it derives from all-reduced sums), so either all
    // ranks skip this exchange or none does
    if (guarded && sc[SC_STATUS] != 0.0) return;
    for (int a = 0; a < narr; ++a) {
        double s = 0.0;
        for (int i = threadIdx.x; i < cnt; i += VT) s += part[a * stride + i];
        s = block_sum(s, sh);
        if (threadIdx.x == 0) mine[a] = s;
        __syncthreads();
    }
    double res[PEER_VALS];
    peer_exchange_sum(L, seq, mine, narr, all, res);
    if (threadIdx.x == 0) {
        for (int a = 0; a < narr; ++a) dst[a] = res[a];
        if (phase >= 0) scalar_phase(sc, phase, hist, nhist);
    }
}

int vec_grid(size_t N)
{
    size_t nb = (N + (size_t)VT * 8 - 1) / ((size_t)VT * 8);
    if (nb > 148 * 8) nb = 148 * 8;
    if (nb < 1) nb = 1;
    return (int)nb;
}

// the device residual history holds what the caller will read back: min(maxit + 1, nhist) entries when a
// history was asked for, one entry otherwise (scalar_phase guards every store with `it < nhist`), so that a
// legal -ksp_max_it of 1e9 does not turn into an 8 GB allocation
static int hist_entries(int maxit, const double *hist, int nhist)
{
    if (!hist || nhist <= 0) return 1;
    const long long want = (long long)maxit + 1;
    return (int)(want < (long long)nhist ? want : (long long)nhist);
}

int cg_alloc(pbx_handle_s *h, int nhist)
{
    const size_t N = (size_t)h->nx * h->ny * h->nz;
    if (!h->cg_r) {
        PBX_CUDA(cudaMalloc(&h->cg_r, N * sizeof(double)));
        PBX_CUDA(cudaMalloc(&h->cg_p, N * sizeof(double)));
        PBX_CUDA(cudaMalloc(&h->cg_w, N * sizeof(double)));
        int np = vec_grid(N);
        Brick g{h->nx, h->ny, h->nz};
        if (h->fast_ok) {
            int nz = fast_zpass_max_partials(g);
            if (nz > np) np = nz;
        }
        h->cg_npartials = np;
        PBX_CUDA(cudaMalloc(&h->cg_partials, 4 * (size_t)np * sizeof(double)));
        PBX_CUDA(cudaMalloc(&h->cg_scal, SC_COUNT * sizeof(double)));
        PBX_CUDA(cudaMalloc(&h->cg_ticket, sizeof(unsigned)));
        PBX_CUDA(cudaMemset(h->cg_ticket, 0, sizeof(unsigned)));
        // pinned and mapped: the last two doubles' worth is the status word k_pupdate_x posts into
        PBX_CUDA(cudaHostAlloc(&h->cg_host, (2 * SC_COUNT + 2) * sizeof(double), cudaHostAllocMapped));
    }
    if (h->pc != PBX_PC_NONE && !h->cg_z) PBX_CUDA(cudaMalloc(&h->cg_z, N * sizeof(double)));
    if (nhist < 1) nhist = 1;
    if (h->cg_hist_cap < nhist) {
        if (h->cg_hist) cudaFree(h->cg_hist);
        h->cg_hist = nullptr;
        h->cg_hist_cap = 0;
        PBX_CUDA(cudaMalloc(&h->cg_hist, (size_t)nhist * sizeof(double)));
        h->cg_hist_cap = nhist;
    }
    return PBX_OK;
}

// Sums of `narr` arrays of per-CTA partials into dst[0 .. narr) (inside the scalar block), summed
// over all ranks, followed by scalar step `phase` of the loop (phase < 0: none).  Single rank or
// NCCL: k_reduce, ncclAllReduce, k_scalar.  Peer boards: the one fused kernel k_reduce_peer.
int reduce_step(pbx_handle_s *h, const double *part, int cnt, int stride, int narr, double *dst,
                int guarded, int phase)
{
    cudaStream_t s = h->stream;
    double *sc = h->cg_scal;
    PeerLinks L;
    unsigned long long seq;
    if (h->nranks > 1 && narr <= PEER_VALS && dist_peer_next(h, &L, &seq)) {
        k_reduce_peer<<<1, VT, 0, s>>>(part, cnt, stride, narr, dst, sc, guarded, L, seq, phase,
                                       h->cg_hist, h->cg_hist_cap);
        ++h->launches;
    } else {
        k_reduce<<<1, VT, 0, s>>>(part, cnt, stride, narr, dst, sc, guarded);
        ++h->launches;
        if (h->nranks > 1) PBX_TRY(dist_allreduce_sum(h, dst, narr));
        if (phase >= 0) {
            k_scalar<<<1, 1, 0, s>>>(sc, phase, h->cg_hist, h->cg_hist_cap);
            ++h->launches;
        }
    }
    PBX_CUDA(cudaGetLastError());
    return PBX_OK;
}

__global__ void k_dot_generic(size_t N, const double *__restrict__ a, const double *__restrict__ b,
                              double *__restrict__ part)
{
    __shared__ double sh[VT / 32];
    size_t lo, hi;
    slice(N, &lo, &hi);
    double s = 0.0;
    for (size_t i = lo + threadIdx.x; i < hi; i += VT) s = fma(a[i], b[i], s);
    s = block_sum(s, sh);
    if (threadIdx.x == 0) part[blockIdx.x] = s;
}

// Reduction tails (default; PBX_FUSE_TAIL=0 turns them off): the reduction of the per-CTA partial sums, its all-reduce and the scalar step run in
// the tail of the kernel that wrote the partials (cgdev::red_tail) instead of in further launches.
// Possible on one rank and with the peer boards (an NCCL all-reduce cannot be called from a kernel).
bool make_tail(pbx_handle_s *h, double *dst, int guarded, int phase, RedTail *t)
{
    // default since the runs of round 2 (one rank: 9 -> 5 launches per iteration; eight ranks with the peer boards:
    // CG 0.775 -> 0.722 s); PBX_FUSE_TAIL=0 keeps the reductions in launches of their own
    if (!env_switch("PBX_FUSE_TAIL", true) || !h->cg_ticket) return false;
    *t = RedTail();
    if (h->nranks > 1 && !dist_peer_next(h, &t->L, &t->seq)) return false;
    t->on = 1;
    t->ticket = h->cg_ticket;
    t->sc = h->cg_scal;
    t->dst = dst;
    t->hist = h->cg_hist;
    t->nhist = h->cg_hist_cap;
    t->phase = phase;
    t->guarded = guarded;
    return true;
}

// w = A p and p.w (over all ranks) into dst (device), then scalar step `phase` (< 0: none)
int matmult_dot(pbx_handle_s *h, const double *p, double *w, double *dst, int guarded, int phase)
{
    const size_t N = (size_t)h->nx * h->ny * h->nz;
    cudaStream_t s = h->stream;
    int np;
    if (h->op == PBX_OPERATOR_STAR) {
        PBX_TRY(matmult(h, p, w));
        np = vec_grid(N);
        k_dot_generic<<<np, VT, 0, s>>>(N, p, w, h->cg_partials);
        ++h->launches;
    } else if (h->nranks > 1 || h->mode == PBX_MODE_FAST) {
        // the z pass may reduce its partial sums itself (it takes h->pending_tail if it can)
        RedTail tail;
        const bool offered = make_tail(h, dst, guarded, phase, &tail);
        h->pending_tail = offered ? &tail : nullptr;
        const int rc = h->nranks > 1 ? dist_lapl(h, p, w, p, h->cg_partials) : lapl_fast(h, p, w, p, h->cg_partials);
        const bool taken = offered && h->pending_tail == nullptr;
        h->pending_tail = nullptr;
        if (offered && !taken && h->nranks > 1) dist_peer_unget(h, tail.seq);   // reduce_step draws it again
        PBX_TRY(rc);
        if (taken) return PBX_OK;
        np = fast_zpass_max_partials(Brick{h->nx, h->ny, h->nz});
    } else {
        PBX_TRY(lapl_reference(h, p, w));
        np = vec_grid(N);
        k_dot_generic<<<np, VT, 0, s>>>(N, p, w, h->cg_partials);
        ++h->launches;
    }
    PBX_CUDA(cudaGetLastError());
    return reduce_step(h, h->cg_partials, np, np, 1, dst, guarded, phase);
}

}  // namespace

void cg_free(pbx_handle_s *h)
{
    if (h->cg_r) cudaFree(h->cg_r);
    if (h->cg_p) cudaFree(h->cg_p);
    if (h->cg_w) cudaFree(h->cg_w);
    if (h->cg_partials) cudaFree(h->cg_partials);
    if (h->cg_scal) cudaFree(h->cg_scal);
    if (h->cg_ticket) cudaFree(h->cg_ticket);
    h->cg_ticket = nullptr;
    if (h->cg_host) cudaFreeHost(h->cg_host);
    if (h->cg_hist) cudaFree(h->cg_hist);
    if (h->cg_z) cudaFree(h->cg_z);
    h->cg_r = h->cg_p = h->cg_w = h->cg_partials = h->cg_scal = h->cg_host = h->cg_hist = h->cg_z = nullptr;
    h->cg_hist_cap = 0;
}

int cg_lapl_dot(pbx_handle_s *h, const double *f, double *out, double *dot_dev)
{
    PBX_TRY(cg_alloc(h, 1));
    return matmult_dot(h, f, out, dot_dev, 0, -1);
}

// z = M^-1 (r - *mean) for the handle's preconditioner (no mean removal of z)
static int pc_raw(pbx_handle_s *h, const double *r, const double *mean_dev, double *z)
{
    if (h->pc == PBX_PC_MG) return mg_vcycle(h, r, mean_dev, z);
    const size_t N = (size_t)h->nx * h->ny * h->nz;
    PBX_CUDA(cudaMemcpyAsync(z, r, N * sizeof(double), cudaMemcpyDeviceToDevice, h->stream));
    k_center<<<vec_grid(N), VT, 0, h->stream>>>(N, z, mean_dev);
    ++h->launches;
    return PBX_OK;
}

int pc_apply(pbx_handle_s *h, const double *r, double *z)
{
    const size_t N = (size_t)h->nx * h->ny * h->nz;
    PBX_TRY(cg_alloc(h, 1));
    cudaStream_t s = h->stream;
    double *sc = h->cg_scal, *part = h->cg_partials;
    const int np = h->cg_npartials, nb = vec_grid(N);
    const double ntot = (double)N * (double)h->nranks;
    PBX_CUDA(cudaMemcpyAsync(sc + SC_NTOT, &ntot, sizeof ntot, cudaMemcpyHostToDevice, s));
    k_sum<<<nb, VT, 0, s>>>(N, r, part);
    ++h->launches;
    PBX_TRY(reduce_step(h, part, nb, np, 1, sc + SC_S1, 0, 7));
    PBX_TRY(pc_raw(h, r, sc + SC_MEAN, z));
    k_sum<<<nb, VT, 0, s>>>(N, z, part);
    ++h->launches;
    PBX_TRY(reduce_step(h, part, nb, np, 1, sc + SC_S1, 0, 8));
    k_center<<<nb, VT, 0, s>>>(N, z, sc + SC_MZ);
    ++h->launches;
    PBX_CUDA(cudaGetLastError());
    return PBX_OK;
}

// KSPCG with a preconditioner: z = M^-1 r (constant removed), preconditioned norm, beta = z.r
static int cg_solve_pc(pbx_handle_s *h, const double *b, double *x, double rtol, double abstol,
                       int maxit, int *its, double *rnorm, int *reason, double *hist, int nhist)
{
    const size_t N = (size_t)h->nx * h->ny * h->nz;
    PBX_TRY(cg_alloc(h, hist_entries(maxit, hist, nhist)));
    cudaStream_t s = h->stream;
    double *sc = h->cg_scal, *part = h->cg_partials;
    const int np = h->cg_npartials, nb = vec_grid(N);
    double *r = h->cg_r, *p = h->cg_p, *w = h->cg_w, *z = h->cg_z;

    double init[SC_COUNT];
    for (int i = 0; i < SC_COUNT; ++i) init[i] = 0.0;
    init[SC_RTOL] = rtol;
    init[SC_ABSTOL] = abstol;
    init[SC_MAXIT] = (double)maxit;
    init[SC_NTOT] = (double)N * (double)h->nranks;
    PBX_CUDA(cudaMemcpyAsync(sc, init, sizeof init, cudaMemcpyHostToDevice, s));
    k_sum<<<nb, VT, 0, s>>>(N, b, part);
    ++h->launches;
    PBX_TRY(reduce_step(h, part, nb, np, 1, sc + SC_S1, 0, 0));   // m0 = mean(b) = mean(r) from here on
    k_init_pc<<<nb, VT, 0, s>>>(N, b, x, r);
    ++h->launches;
    PBX_TRY(pc_raw(h, r, sc + SC_MEAN, z));
    k_pcdots<<<nb, VT, 0, s>>>(N, z, r, sc, part, np);
    ++h->launches;
    PBX_TRY(reduce_step(h, part, nb, np, 4, sc + SC_SZ, 0, 4));
    k_pupdate_pc<<<nb, VT, 0, s>>>(N, z, p, sc);
    ++h->launches;
    PBX_CUDA(cudaGetLastError());

    cudaEvent_t ev[2];
    PBX_CUDA(cudaEventCreateWithFlags(&ev[0], cudaEventDisableTiming));
    PBX_CUDA(cudaEventCreateWithFlags(&ev[1], cudaEventDisableTiming));
    double *hs = h->cg_host;
    PBX_CUDA(cudaMemcpyAsync(hs, sc, SC_COUNT * sizeof(double), cudaMemcpyDeviceToHost, s));
    PBX_CUDA(cudaStreamSynchronize(s));
    int rc = PBX_OK;
    bool done = hs[SC_STATUS] != 0.0;
    int issued = 0;
    while (!done && issued < maxit) {
        const int slot = issued & 1;
        if ((rc = matmult_dot(h, p, w, sc + SC_PW, 1, 2)) != PBX_OK) break;
        k_update<<<nb, VT, 0, s>>>(N, x, r, p, w, sc, part, np);
        ++h->launches;
        if ((rc = reduce_step(h, part, nb, np, 2, sc + SC_S1, 1, 5)) != PBX_OK) break;
        if ((rc = pc_raw(h, r, sc + SC_MEAN, z)) != PBX_OK) break;
        k_pcdots<<<nb, VT, 0, s>>>(N, z, r, sc, part, np);
        ++h->launches;
        if ((rc = reduce_step(h, part, nb, np, 4, sc + SC_SZ, 1, 6)) != PBX_OK) break;
        k_pupdate_pc<<<nb, VT, 0, s>>>(N, z, p, sc);
        ++h->launches;
        cudaMemcpyAsync(hs + slot * SC_COUNT, sc, SC_COUNT * sizeof(double), cudaMemcpyDeviceToHost, s);
        cudaEventRecord(ev[slot], s);
        ++issued;
        if (issued >= 2) {
            cudaEventSynchronize(ev[slot ^ 1]);
            if (hs[(slot ^ 1) * SC_COUNT + SC_STATUS] != 0.0) done = true;
        }
    }
    cudaError_t e = cudaStreamSynchronize(s);
    cudaEventDestroy(ev[0]);
    cudaEventDestroy(ev[1]);
    if (rc != PBX_OK) return rc;
    PBX_CUDA(e);
    PBX_CUDA(cudaMemcpy(hs, sc, SC_COUNT * sizeof(double), cudaMemcpyDeviceToHost));
    int st = (int)hs[SC_STATUS];
    const int nit = (int)hs[SC_IT];
    if (st == 0) st = PBX_DIVERGED_ITS;
    if (its) *its = nit;
    if (rnorm) *rnorm = hs[SC_DP];
    if (reason) *reason = st;
    if (hist && nhist > 0) {
        const int cnt = nit + 1 < nhist ? nit + 1 : nhist;
        PBX_CUDA(cudaMemcpy(hist, h->cg_hist, cnt * sizeof(double), cudaMemcpyDeviceToHost));
    }
    PBX_CUDA(cudaGetLastError());
    return PBX_OK;
}

int cg_solve(pbx_handle_s *h, const double *b, double *x, double rtol, double abstol, int maxit,
             int *its, double *rnorm, int *reason, double *hist, int nhist)
{
    if (h->pc != PBX_PC_NONE) {
        if (((reinterpret_cast<uintptr_t>(b) | reinterpret_cast<uintptr_t>(x)) & 15) != 0) {
            set_last_error("pbx_cg_solve: b and x must be 16-byte aligned");
            return PBX_ERR_ARG;
        }
        return cg_solve_pc(h, b, x, rtol, abstol, maxit, its, rnorm, reason, hist, nhist);
    }
    const size_t N = (size_t)h->nx * h->ny * h->nz;
    if (((reinterpret_cast<uintptr_t>(b) | reinterpret_cast<uintptr_t>(x)) & 15) != 0) {
        set_last_error("pbx_cg_solve: b and x must be 16-byte aligned");
        return PBX_ERR_ARG;
    }
    PBX_TRY(cg_alloc(h, hist_entries(maxit, hist, nhist)));
    cudaStream_t s = h->stream;
    double *sc = h->cg_scal, *part = h->cg_partials;
    const int np = h->cg_npartials;
    const int nb = vec_grid(N);
    double *r = h->cg_r, *p = h->cg_p, *w = h->cg_w;

    double init[SC_COUNT];
    for (int i = 0; i < SC_COUNT; ++i) init[i] = 0.0;
    init[SC_RTOL] = rtol;
    init[SC_ABSTOL] = abstol;
    init[SC_MAXIT] = (double)maxit;
    init[SC_NTOT] = (double)N * (double)h->nranks;
    PBX_CUDA(cudaMemcpyAsync(sc, init, sizeof init, cudaMemcpyHostToDevice, s));

    // mean of b, then r = b, x = 0, p = z = b - mean, ||z||
    k_sum<<<nb, VT, 0, s>>>(N, b, part);
    ++h->launches;
    PBX_TRY(reduce_step(h, part, nb, np, 1, sc + SC_S1, 0, 0));
    k_init<<<nb, VT, 0, s>>>(N, b, x, r, p, sc, part, np);
    ++h->launches;
    PBX_TRY(reduce_step(h, part, nb, np, 2, sc + SC_S1, 0, 1));
    PBX_CUDA(cudaGetLastError());

    // The host runs one iteration ahead of the status it has seen: every kernel that changes
    // solver state checks the device status word first, so iterations issued after convergence
    // are no-ops.  The status travels as one word that the p-update posts into pinned host memory.
    double *hs = h->cg_host;
    volatile unsigned long long *hword = reinterpret_cast<volatile unsigned long long *>(hs + 2 * SC_COUNT);
    unsigned long long *dword = nullptr;
    PBX_CUDA(cudaHostGetDevicePointer((void **)&dword, (void *)(hs + 2 * SC_COUNT), 0));
    *hword = 0;
    PBX_CUDA(cudaMemcpyAsync(hs, sc, SC_COUNT * sizeof(double), cudaMemcpyDeviceToHost, s));
    PBX_CUDA(cudaStreamSynchronize(s));
    int rc = PBX_OK;
    bool done = hs[SC_STATUS] != 0.0;
    int issued = 0;
    while (!done && issued < maxit) {
        rc = matmult_dot(h, p, w, sc + SC_PW, 1, 2);
        if (rc != PBX_OK) break;
        RedTail tail;
        const bool tailed = make_tail(h, sc + SC_S1, 1, 3, &tail);
        if (tailed) {
            tail.part = part;
            tail.cnt = nb;
            tail.stride = np;
            tail.narr = 2;
        } else {
            tail = RedTail();
        }
        k_update_r<<<nb, VT, 0, s>>>(N, r, w, sc, part, np, tail);
        ++h->launches;
        if (!tailed) {
            if ((rc = reduce_step(h, part, nb, np, 2, sc + SC_S1, 1, 3)) != PBX_OK) break;
        }
        k_pupdate_x<<<vec_grid(N), VT, 0, s>>>(N, r, p, x, sc, issued + 1, dword);
        ++h->launches;
        ++issued;
        if (issued >= 2) {
            // wait for the word of iteration issued - 1 (bounded: a failed kernel ends the wait through the stream's
            // error state, a finished stream through the word itself)
            const unsigned long long want = (unsigned long long)(issued - 1);
            unsigned long long wv;
            long long spins = 0;
            while (((wv = *hword) >> 16) < want) {
#if defined(__x86_64__) || defined(__i386__)
                __builtin_ia32_pause();   // the word arrives within an iteration (0.6-4 ms): be a polite spinner
#endif
                if ((++spins & 0xfff) == 0) {
                    const cudaError_t q = cudaStreamQuery(s);
                    if (q != cudaErrorNotReady && ((*hword) >> 16) < want) {
                        if (q == cudaSuccess) set_last_error("pbx_cg_solve: the status word never arrived");
                        rc = q == cudaSuccess ? PBX_ERR_CUDA : cuda_fail(q, "cudaStreamQuery", __FILE__, __LINE__);
                        break;
                    }
                }
            }
            if (rc != PBX_OK) break;
            if ((int)(wv & 0xffff) != 1024) done = true;
        }
    }
    cudaError_t e = cudaStreamSynchronize(s);
    if (rc != PBX_OK) return rc;
    PBX_CUDA(e);
    PBX_CUDA(cudaMemcpy(hs, sc, SC_COUNT * sizeof(double), cudaMemcpyDeviceToHost));
    int st = (int)hs[SC_STATUS];
    int nit = (int)hs[SC_IT];
    if (st == 0) st = PBX_DIVERGED_ITS;   // maxit == 0 or loop exhausted without a verdict
    if (its) *its = nit;
    if (rnorm) *rnorm = hs[SC_DP];
    if (reason) *reason = st;
    if (hist && nhist > 0) {
        int cnt = nit + 1 < nhist ? nit + 1 : nhist;
        PBX_CUDA(cudaMemcpy(hist, h->cg_hist, cnt * sizeof(double), cudaMemcpyDeviceToHost));
    }
    PBX_CUDA(cudaGetLastError());
    return PBX_OK;
}

}  // namespace pbx
