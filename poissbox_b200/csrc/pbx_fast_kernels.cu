// pbx_fast_kernels.cu -- FAST schedule of the compact Laplacian: one sweep per axis.
//
//   lapl = Dxx Myy Mzz + Mxx Dyy Mzz + Mxx Myy Dzz ,   Dxx = D+x G-x ,  Mxx = I+x I-x , ...
//
// which is src/compact_schemes.f90:17-37 (lapl = div(grad), stages :60-86 and :226-253) with the
// 1-D operators of different axes commuted next to each other (they are circulants acting on
// different indices, so they commute exactly in exact arithmetic).  Three kernels:
//
//   x pass   f        -> A = Dxx f            , B = Mxx f           ( 8 B read, 16 B written / DoF)
//   y pass   A, B     -> C = Myy A + Dyy B    , D = Myy B           (16 B read, 16 B written / DoF)
//   z pass   C, D     -> out = Mzz C + Dzz D  (+ optional p.out)    (16 B read,  8 B written / DoF)
//
// = 80 B/DoF, the algorithmic minimum of a one-sweep-per-axis schedule.
//
// Each composite 1-D operator O = A^-2 S (see pbx_internal.h) is applied to a periodic line that
// is cut into chunks of LC = 16 points; a thread owns one chunk of one line IN REGISTERS:
//   1. 7-point symmetric stencil S on the raw input (3-point halos from the neighbouring chunks),
//   2. local causal double recursion   y_k = s_k + r y_{k-1},  z_k = y_k + r z_{k-1}  from zero state,
//   3. the true incoming state is the neighbours' local end states, looked back nlook (2-3) chunks
//      (the influence of chunk t-m decays as r^(16 m)); the chunk is corrected with the
//      homogeneous solution  r^(k+1) (Z + (k+1) Y),
//   4. the same, anti-causal.
// Periodicity needs no Sherman-Morrison step: the look-back simply wraps around the line.
// Truncating the look-back at |16 m r^(16 m)| < 1e-19 makes the result exact to fp64 rounding.
//
// Memory access: y and z passes load/store straight between global memory and registers, a warp
// covering 4 chunks x 8 consecutive x (64-byte segments); the x pass stages R whole lines through
// shared memory (coalesced 16-byte accesses, chunk-padded so that the per-thread 128-bit reads are
// bank-conflict free).  Neighbour exchange (halos, chunk states) goes through shared memory.
#include "pbx_fast_common.cuh"

namespace pbx {

using namespace fast;

namespace {

constexpr int XW = 8;             // x-width of a y/z-pass tile (64-byte row segments)
constexpr int CPAD = LC + 2;      // x pass: doubles per chunk in the padded staging buffer

// ---------------------------------------------------------------------------------------------
// y / z pass.  Tile: XW consecutive x, T chunks (one whole line), G lines in the remaining
// direction.  blockDim = (XW, T, G).  sl = stride along the line, sg = stride between lines.
// ---------------------------------------------------------------------------------------------
struct YZParams {
    CompositeCoef M, D;
    int nx, T, ng;          // x extent, chunks per CTA, number of lines in the "group" direction
    long long sl, sg;       // strides (in doubles) along the line / between lines of a group
    SegGeom seg;            // long lines: blockIdx.z numbers the segment (pbx_internal.h)
};

// y pass:  C = M a + D b ;  Dd = M b
__global__ void __launch_bounds__(NT, 2)
ypass_kernel(const __grid_constant__ YZParams p, const double *__restrict__ A,
             const double *__restrict__ B, double *__restrict__ C, double *__restrict__ Dd)
{
    extern __shared__ double sm[];
    const int tx = threadIdx.x, t = threadIdx.y, tz = threadIdx.z;
    const int x = blockIdx.x * XW + tx;
    const int g = blockIdx.y * blockDim.z + tz;
    const bool live = (x < p.nx) && (g < p.ng);
    const SegChunk sc = seg_chunk(p.seg, blockIdx.z, t);
    Xchg xc{sm, (tz * p.T + t) * XW + tx, t, p.T, XW, p.seg.nseg > 1 ? 1 : 0};
    const long long base = (long long)x + (long long)(sc.chunk * LC) * p.sl + (long long)g * p.sg;

    double a[LC], b[LC];
#pragma unroll
    for (int k = 0; k < LC; ++k) {
        a[k] = live ? __ldg(A + base + k * p.sl) : 0.0;
        b[k] = live ? __ldg(B + base + k * p.sl) : 0.0;
    }
    put_halo(xc, Y_SLOTS, b);
    __syncthreads();
    double eb[LC + 6], c[LC], d[LC];
    get_halo(xc, Y_SLOTS, b, eb);
    ypass_body(p.M, p.D, xc, a, eb, c, d, BarAll());
    if (live && sc.interior) {
#pragma unroll
        for (int k = 0; k < LC; ++k) {
            C[base + k * p.sl] = c[k];
            Dd[base + k * p.sl] = d[k];
        }
    }
}

// z pass:  out = M c + D d ; optionally partial sums of pv . out (one per CTA, deterministic)
__global__ void __launch_bounds__(NT, 2)
zpass_kernel(const __grid_constant__ YZParams p, const __grid_constant__ ZOpen zo,
             const double *__restrict__ Cc, const double *__restrict__ Dd,
             double *__restrict__ out, const double *__restrict__ pv,
             double *__restrict__ partials)
{
    extern __shared__ double sm[];
    const int tx = threadIdx.x, t = threadIdx.y, tz = threadIdx.z;
    const int x = blockIdx.x * XW + tx;
    const int g = blockIdx.y * blockDim.z + tz;
    const bool live = (x < p.nx) && (g < p.ng);
    const SegChunk sc = seg_chunk(p.seg, blockIdx.z, t);
    Xchg xc{sm, (tz * p.T + t) * XW + tx, t, p.T, XW, (zo.open || p.seg.nseg > 1) ? 1 : 0};
    const long long base = (long long)x + (long long)(sc.chunk * LC) * p.sl + (long long)g * p.sg;

    double c[LC], d[LC];
#pragma unroll
    for (int k = 0; k < LC; ++k) {
        c[k] = live ? __ldg(Cc + base + k * p.sl) : 0.0;
        d[k] = live ? __ldg(Dd + base + k * p.sl) : 0.0;
    }
    double lo9[DIST_MSG], up9[DIST_MSG];
    if (zo.open)
        slab_load_messages(zo, t == 0, t == p.T - 1, live ? (long long)x + (long long)p.nx * g : 0, lo9, up9);
    put_halo(xc, ZNAT_SLOTS, d);
    __syncthreads();
    double ed[LC + 6], o[LC];
    get_halo(xc, ZNAT_SLOTS, d, ed);
    if (zo.open) {
        zpass_body_slab(p.M, p.D, zo, xc, lo9, up9, c, ed, o, BarAll());
    } else {
        zpass_body(p.M, p.D, xc, c, ed, o, BarAll());
    }

    double dot = 0.0;
    if (live && sc.interior) {
        if (pv != nullptr) {
#pragma unroll
            for (int k = 0; k < LC; ++k) {
                dot = fma(__ldg(pv + base + k * p.sl), o[k], dot);
                out[base + k * p.sl] = o[k];
            }
        } else {
#pragma unroll
            for (int k = 0; k < LC; ++k) out[base + k * p.sl] = o[k];
        }
    }
    if (pv != nullptr) {
        const int nthr = blockDim.x * blockDim.y * blockDim.z;
        // same association as the TMA kernel whenever the CTA consists of full warps
        double tot = (nthr & 31) == 0 ? block_sum_warps(dot, sm + (Y_SLOTS + 5) * NT, xc.q, nthr, BarAll())
                                      : block_sum_fixed(dot, sm, xc.q, nthr, BarAll());
        if (xc.q == 0) partials[(blockIdx.z * gridDim.y + blockIdx.y) * gridDim.x + blockIdx.x] = tot;
    }
}

// ---------------------------------------------------------------------------------------------
// x pass.  Tile: R whole lines (contiguous in memory).  blockDim = (T, R).
// ---------------------------------------------------------------------------------------------
struct XParams {
    CompositeCoef M, D;
    int n, T, R;            // line length, chunks per line, lines per CTA
    long long nlines;
};

constexpr int XK_SLOTS = X_SLOTS + 6;

__global__ void __launch_bounds__(NT, 2)
xpass_kernel(const __grid_constant__ XParams p, const double *__restrict__ F,
             double *__restrict__ A, double *__restrict__ B)
{
    extern __shared__ double sm[];
    const int t = threadIdx.x, r = threadIdx.y;
    const int tid = r * p.T + t;
    const int nthr = p.T * p.R;
    const long long line0 = (long long)blockIdx.x * p.R;
    const int rlive = (int)((p.nlines - line0) < p.R ? (p.nlines - line0) : p.R);
    const int rowpad = p.T * CPAD;
    double *xs = sm;                         // exchange area
    double *buf0 = sm + XK_SLOTS * NT;       // staging, R * rowpad doubles
    double *buf1 = buf0 + p.R * rowpad;
    Xchg xc{xs, tid, t, p.T, 1};

    // stage in: coalesced 16-byte loads, chunk-padded rows
    {
        const double2 *src = reinterpret_cast<const double2 *>(F + line0 * p.n);
        const int n2 = rlive * p.n / 2;
        for (int q = tid; q < n2; q += nthr) {
            double2 val = __ldg(src + q);
            int i = 2 * q;
            int rr = i / p.n, ii = i - rr * p.n;
            *reinterpret_cast<double2 *>(buf0 + rr * rowpad + (ii >> 4) * CPAD + (ii & 15)) = val;
        }
    }
    __syncthreads();
    double f[LC];
    {
        const double2 *src = reinterpret_cast<const double2 *>(buf0 + r * rowpad + t * CPAD);
#pragma unroll
        for (int k = 0; k < LC / 2; ++k) {
            double2 val = (r < rlive) ? src[k] : make_double2(0.0, 0.0);
            f[2 * k] = val.x;
            f[2 * k + 1] = val.y;
        }
    }
    put_halo(xc, X_SLOTS, f);
    __syncthreads();
    double ef[LC + 6], va[LC], vb[LC];
    get_halo(xc, X_SLOTS, f, ef);
    xpass_body(p.M, p.D, xc, ef, va, vb, BarAll());

    // stage out through the padded buffers, then coalesced 16-byte stores
    {
        double2 *d0 = reinterpret_cast<double2 *>(buf0 + r * rowpad + t * CPAD);
        double2 *d1 = reinterpret_cast<double2 *>(buf1 + r * rowpad + t * CPAD);
#pragma unroll
        for (int k = 0; k < LC / 2; ++k) {
            d0[k] = make_double2(va[2 * k], va[2 * k + 1]);
            d1[k] = make_double2(vb[2 * k], vb[2 * k + 1]);
        }
    }
    __syncthreads();
    {
        double2 *da = reinterpret_cast<double2 *>(A + line0 * p.n);
        double2 *db = reinterpret_cast<double2 *>(B + line0 * p.n);
        const int n2 = rlive * p.n / 2;
        for (int q = tid; q < n2; q += nthr) {
            int i = 2 * q;
            int rr = i / p.n, ii = i - rr * p.n;
            int off = rr * rowpad + (ii >> 4) * CPAD + (ii & 15);
            da[q] = *reinterpret_cast<const double2 *>(buf0 + off);
            db[q] = *reinterpret_cast<const double2 *>(buf1 + off);
        }
    }
}

constexpr size_t YZ_SMEM = sizeof(double) * (Y_SLOTS + 6) * NT;

}  // namespace

bool fast_supported(int nx, int ny, int nz)
{
    // chunks of 16 points; an x line must fit one CTA (T <= 256 chunks); y and z lines of more
    // than 32 chunks are cut into overlapping segments (SegGeom)
    auto ok = [](int n, int tmax) { return n >= LC && n % LC == 0 && n / LC <= tmax; };
    return ok(nx, NT) && ok(ny, 1 << 20) && ok(nz, 1 << 20);
}

int fast_xpass(cudaStream_t s, const Brick &g, const FastCoefs &fc, const double *f, double *A,
               double *B, long long *launches)
{
    XParams p;
    p.M = fc.M;
    p.D = fc.D[0];
    p.n = g.nx;
    p.T = g.nx / LC;
    p.R = NT / p.T;
    p.nlines = (long long)g.ny * g.nz;
    if (p.R > p.nlines) p.R = (int)p.nlines;
    const size_t smem = sizeof(double) * (XK_SLOTS * NT + 2 * (size_t)p.R * p.T * CPAD);
    static std::atomic<bool> attr_set[64];   // per device: the attribute belongs to the context
    int dev_ = 0;
    cudaGetDevice(&dev_);
    if (!attr_set[dev_ & 63]) {
        PBX_CUDA(cudaFuncSetAttribute(xpass_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                      112 * 1024));
        attr_set[dev_ & 63] = true;
    }
    dim3 block(p.T, p.R);
    unsigned grid = (unsigned)((p.nlines + p.R - 1) / p.R);
    xpass_kernel<<<grid, block, smem, s>>>(p, f, A, B);
    if (launches) ++*launches;
    PBX_CUDA(cudaGetLastError());
    return PBX_OK;
}

static void yz_geometry(const Brick &g, int dir, YZParams *p, dim3 *grid, dim3 *block)
{
    const int n = dir == 1 ? g.ny : g.nz;
    p->nx = g.nx;
    p->seg = seg_geometry(n / LC);
    p->T = p->seg.T;
    p->ng = dir == 1 ? g.nz : g.ny;
    p->sl = dir == 1 ? (long long)g.nx : (long long)g.nx * g.ny;
    p->sg = dir == 1 ? (long long)g.nx * g.ny : (long long)g.nx;
    int G = NT / (XW * p->T);
    if (G < 1) G = 1;
    if (G > p->ng) G = p->ng;
    *block = dim3(XW, p->T, G);
    *grid = dim3((g.nx + XW - 1) / XW, (p->ng + G - 1) / G, p->seg.nseg);
}

int fast_ypass(cudaStream_t s, const Brick &g, const FastCoefs &fc, const double *A,
               const double *B, double *C, double *D, long long *launches)
{
    YZParams p;
    p.M = fc.M;
    p.D = fc.D[1];
    dim3 grid, block;
    yz_geometry(g, 1, &p, &grid, &block);
    static std::atomic<bool> attr_set[64];   // per device: the attribute belongs to the context
    int dev_ = 0;
    cudaGetDevice(&dev_);
    if (!attr_set[dev_ & 63]) {
        PBX_CUDA(cudaFuncSetAttribute(ypass_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                      (int)YZ_SMEM));
        attr_set[dev_ & 63] = true;
    }
    ypass_kernel<<<grid, block, YZ_SMEM, s>>>(p, A, B, C, D);
    if (launches) ++*launches;
    PBX_CUDA(cudaGetLastError());
    return PBX_OK;
}

int fast_zpass_max_partials(const Brick &g)
{
    YZParams p;
    dim3 grid, block;
    yz_geometry(g, 2, &p, &grid, &block);
    return (int)(grid.x * grid.y * grid.z);
}

int fast_zpass(cudaStream_t s, const Brick &g, const FastCoefs &fc, const double *C,
               const double *D, double *out, const double *pvec, double *dot_partials,
               int *n_partials, const ZOpen &zo, long long *launches)
{
    YZParams p;
    p.M = fc.M;
    p.D = fc.D[2];
    dim3 grid, block;
    yz_geometry(g, 2, &p, &grid, &block);
    static std::atomic<bool> attr_set[64];   // per device: the attribute belongs to the context
    int dev_ = 0;
    cudaGetDevice(&dev_);
    if (!attr_set[dev_ & 63]) {
        PBX_CUDA(cudaFuncSetAttribute(zpass_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                      (int)YZ_SMEM));
        attr_set[dev_ & 63] = true;
    }
    if (zo.open && p.seg.nseg > 1) {
        set_last_error("a z slab of more than 512 planes per rank is not supported");
        return PBX_ERR_UNSUPPORTED;
    }
    zpass_kernel<<<grid, block, YZ_SMEM, s>>>(p, zo, C, D, out, pvec, dot_partials);
    if (n_partials) *n_partials = (int)(grid.x * grid.y * grid.z);
    if (launches) ++*launches;
    PBX_CUDA(cudaGetLastError());
    return PBX_OK;
}

}  // namespace pbx
